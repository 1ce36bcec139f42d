"""GPU (-m gpu), needs >= 2 GPUs in the box (skipped otherwise): the sharded path on real devices --
tools/gpu_multi_check.py under torchrun, one rank per GPU over NCCL + symmetric memory.

  camera split (B < ranks)   forward slices bit-identical to the single-GPU forward; summed position / colour
                             gradients and gathered sh rows equal the single-GPU gradients over all cameras
                             (NCCL, one-shot peer-load all-reduce, NVLS) to rel-L2 <= 2e-5 (atomic order only)
  batch split, shared params the all-reduce fused into the backward's last kernel == NCCL == the sum one GPU computes
                             alone over every rank's batch element; identical bits on every rank
  public layer               CudaRendererGpu(sharedGrads_attr=...) captured in CUDA graphs returns the summed gradients
The single-GPU half of the camera split (two handles on one device) runs everywhere: test_camera_split_on_one_device."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT
from gvv_differentiable_cuda_renderer_b200 import _native, sharding, synthetic

pytestmark = pytest.mark.gpu
KEYS = ("vertex_pos", "vertex_color", "texture", "sh_coeff", "target_image", "extrinsics", "intrinsics")


def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm())


def test_camera_split_on_one_device():
    """plan_views / shard_views with the real kernels, the collectives replaced by local sums: the cameras of one batch
    element rendered by three handles (as three ranks would) give the single-handle result."""
    dev = torch.device("cuda:0")
    C, W, H = 7, 200, 168
    sc = synthetic.make_scene(kind="sphere", rings=40, segments=48, cameras=C, width=W, height=H, batch=1, tex=16, seed=31)
    N = sc["num_vertices"]
    full = {k: torch.as_tensor(sc[k], device=dev) for k in KEYS}
    G = torch.randn((1, C, H, W, 3), generator=torch.Generator().manual_seed(1)).to(dev)
    whole = _native.NativeRenderer(sc["faces"], sc["texcoords"], N, C, W, H, "vertexColor", "shaded", 1, 1, False, dev)
    fo = whole.forward(*[full[k] for k in KEYS])
    fg = whole.backward(G, None, *[full[k] for k in KEYS[:5]], fo[3], fo[0], fo[1], full["extrinsics"], full["intrinsics"])
    plan = sharding.plan_views(1, C, 3)
    assert sharding.camera_teams(plan) == [[0, 1, 2]]
    gpos = torch.zeros_like(fg[0]); gcol = torch.zeros_like(fg[1]); gsh = torch.zeros_like(fg[3])
    for (b0, b1, c0, c1) in plan:
        loc = sharding.shard_views(full, C, (b0, b1, c0, c1))
        part = _native.NativeRenderer(sc["faces"], sc["texcoords"], N, c1 - c0, W, H, "vertexColor", "shaded", 1, 1, False, dev)
        po = part.forward(*[loc[k] for k in KEYS])
        for a, b in zip(po[:3], fo[:3]):                                  # bary, face, render of the slice: same bits
            assert torch.equal(a, b[:, c0:c1])
        assert torch.equal(po[3], fo[3][:, c0:c1])                         # vertex normals are per (b, c) copies
        g = part.backward(G[:, c0:c1].contiguous(), None, *[loc[k] for k in KEYS[:5]], po[3], po[0], po[1], loc["extrinsics"], loc["intrinsics"])
        gpos += g[0]; gcol += g[1]; gsh[:, c0:c1] = g[3]
        part.close()
    assert rel(gpos, fg[0]) <= 2e-5 and rel(gcol, fg[1]) <= 2e-5 and rel(gsh, fg[3]) <= 2e-5
    whole.close()


def test_shared_batch_grads_option_sums_over_the_batch():
    """shared_batch_grads = 1: colour / SH (/ texture) gradients of all batch elements land in ONE [1, ...] slice."""
    dev = torch.device("cuda:0")
    sc = synthetic.make_scene(kind="sphere", rings=20, segments=24, cameras=2, width=96, height=80, batch=3, tex=16, seed=5)
    N = sc["num_vertices"]
    ins = [torch.as_tensor(sc[k], device=dev) for k in KEYS]
    G = torch.randn((3, 2, 80, 96, 3), generator=torch.Generator().manual_seed(1)).to(dev)
    for albedo in ("vertexColor", "textured"):
        r = _native.NativeRenderer(sc["faces"], sc["texcoords"], N, 2, 96, 80, albedo, "shaded", 1, 1, False, dev)
        o = r.forward(*ins)
        g = r.backward(G, None, ins[0], ins[1], ins[2], ins[3], ins[4], o[3], o[0], o[1], ins[5], ins[6])
        r.set_option("shared_batch_grads", 1)
        s = r.backward(G, None, ins[0], ins[1], ins[2], ins[3], ins[4], o[3], o[0], o[1], ins[5], ins[6])
        assert s[1].shape == (1, N, 3) and s[3].shape == (1, 2, 27) and s[2].shape[0] == 1 and s[0].shape == g[0].shape
        assert rel(s[0], g[0]) <= 1e-5 and rel(s[3], g[3].sum(0, keepdim=True)) <= 1e-5
        if albedo == "vertexColor":
            assert rel(s[1], g[1].sum(0, keepdim=True)) <= 1e-5
        else:
            assert rel(s[2], g[2].sum(0, keepdim=True)) <= 1e-5
        r.close()


@pytest.mark.skipif(torch.cuda.is_available() and torch.cuda.device_count() < 2, reason="needs at least two GPUs")
def test_two_ranks_camera_split_and_fused_allreduce(tmp_path):
    out = tmp_path / "multi.json"
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="0,1")
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29517", os.path.join(ROOT, "tools", "gpu_multi_check.py"), str(out)],
                       capture_output=True, text=True, timeout=900, env=env, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-3000:]
    d = json.load(open(out))
    cs = d["camera_split"]
    assert cs["forward_slices_bit_equal"]
    for mode in ("nccl", "p2p", "nvls"):
        if "unavailable" in cs[mode]:
            assert mode != "nccl"
            continue
        assert all(v <= 2e-5 for v in cs[mode].values()), (mode, cs[mode])
    bs = d["batch_split_shared"]
    for mode in ("nccl", "p2p", "nvls"):
        if "unavailable" in bs[mode]:
            continue
        assert bs[mode]["vertex_color_grad"] <= 1e-5 and bs[mode]["sh_coeff_grad"] <= 1e-5, (mode, bs[mode])
        assert bs[mode].get("bit_identical_across_ranks", True)
    assert "unavailable" not in bs["p2p"], bs["p2p"]            # the peer-load path must work on an NVLink box
    lg = d["layer_in_cuda_graph"]
    assert lg["vertex_color_grad"] <= 1e-5 and lg["sh_coeff_grad"] <= 1e-5
