#!/usr/bin/env python3
"""Multi-GPU correctness + latency check of the sharded path (run under torchrun, one rank per GPU):

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/gpu_multi_check.py [out.json]

  1. camera split (B < ranks): every rank renders its cameras of ONE batch element; position / colour gradients summed
     and sh rows gathered (a) by NCCL (sharding.reduce_camera_split), (b) by the one-shot all-reduce over symmetric
     memory (peer loads, and NVLS when the allocation has a multicast mapping) -- each compared with the gradients a
     single GPU computes over all cameras; forward slices must be bit-identical to the single-GPU forward;
  2. batch split with shared SH / colours (bench.py's configuration): the all-reduce fused into the backward's last
     kernel vs NCCL vs the sum rank 0 computes alone over every rank's batch element;
  3. the same through the public layer (CudaRendererGpu(sharedGrads_attr=...)) captured in CUDA graphs;
  4. device-timed cost per step of: no collective / NCCL after the backward / fused one-shot (p2p, nvls).
Prints one JSON object on rank 0 (and writes it to the path given)."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gvv_differentiable_cuda_renderer_b200 import CudaRendererGpu, _native, sharding, synthetic   # noqa: E402

KEYS = ("vertex_pos", "vertex_color", "texture", "sh_coeff", "target_image", "extrinsics", "intrinsics")


def rel(a, b):
    a, b = a.double(), b.double()
    d = float(b.norm())
    return float((a - b).norm()) / d if d > 0 else float(a.norm())


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    dist.init_process_group("nccl", device_id=dev)
    out = {"world": world}
    T = lambda x: torch.as_tensor(np.ascontiguousarray(x), device=dev)

    # ---------------- 1. camera split ----------------
    C = 2 * world + 1                                    # uneven on purpose: the first rank gets one camera more
    W = H = 256
    sc = synthetic.make_scene(kind="sphere", rings=60, segments=64, cameras=C, width=W, height=H, batch=1, tex=16, seed=21)
    rng = np.random.default_rng(2)
    sc["sh_coeff"] = (sc["sh_coeff"] + 0.2 * rng.random(sc["sh_coeff"].shape, dtype=np.float32)).astype(np.float32)
    N = sc["num_vertices"]
    full_in = {k: T(sc[k]) for k in KEYS}
    G = torch.randn((1, C, H, W, 3), generator=torch.Generator().manual_seed(3)).to(dev)
    whole = _native.NativeRenderer(sc["faces"], sc["texcoords"], N, C, W, H, "vertexColor", "shaded", 1, 1, False, dev)
    fo = whole.forward(*[full_in[k] for k in KEYS])
    fg = whole.backward(G, None, *[full_in[k] for k in KEYS[:5]], fo[3], fo[0], fo[1], full_in["extrinsics"], full_in["intrinsics"])
    plan = sharding.plan_views(1, C, world)
    groups = sharding.make_team_groups(plan)
    b0, b1, c0, c1 = plan[rank]
    loc = sharding.shard_views(full_in, C, plan[rank])
    Cl = c1 - c0
    part = _native.NativeRenderer(sc["faces"], sc["texcoords"], N, Cl, W, H, "vertexColor", "shaded", 1, 1, False, dev)
    po = part.forward(*[loc[k] for k in KEYS])
    Gl = G[:, c0:c1].contiguous()
    bit = all(torch.equal(a.view(torch.int32) if a.dtype == torch.float32 else a, (b[:, c0:c1].contiguous()).view(torch.int32) if b.dtype == torch.float32 else b[:, c0:c1])
              for a, b in zip(po[:4], fo[:4]))
    bargs = lambda: (Gl, None, *[loc[k] for k in KEYS[:5]], po[3], po[0], po[1], loc["extrinsics"], loc["intrinsics"])
    g = part.backward(*bargs())
    gpos, gcol, gtex, gsh = sharding.reduce_camera_split(g, plan, C, rank=rank, groups=groups)
    cs = {"forward_slices_bit_equal": bool(bit), "cameras": C, "plan": plan,
          "nccl": {"vertex_pos_grad": rel(gpos, fg[0]), "vertex_color_grad": rel(gcol, fg[1]), "sh_coeff_grad": rel(gsh, fg[3])}}
    for mode in ("p2p", "nvls"):
        try:
            buf = sharding.SymmetricGradBuffer([(1, N, 3), (1, N, 3), (1, C, 27)], dev, mode=mode, after_backward=True)
        except Exception as e:
            cs[mode] = {"unavailable": repr(e)[:200]}
            continue
        worst = {}
        for step in range(5):
            sh = sharding.SharedGrads(buf, step % 2, ("vertex_pos", "vertex_color", "sh_coeff"), sh_rows=(c0, c1))
            buf.attach(part, sh.slot)
            part.backward(*bargs(), out=sh.outputs())
            part.set_allreduce(None)
            rp, rc, rs = buf.results(sh.slot)
            torch.cuda.synchronize()
            for name, a, b in (("vertex_pos_grad", rp, fg[0]), ("vertex_color_grad", rc, fg[1]), ("sh_coeff_grad", rs, fg[3])):
                worst[name] = max(worst.get(name, 0.0), rel(a, b))
        cs[mode] = worst
        del buf
    out["camera_split"] = cs

    # ---------------- 2. batch split, shared SH + colours (bench.py's shape, smaller) ----------------
    C2 = 4
    base = synthetic.make_scene(kind="sphere", rings=60, segments=64, cameras=C2, width=W, height=H, batch=1, tex=16, seed=22)

    def element(r):
        e = dict(base)
        if r:
            e["vertex_pos"] = (base["vertex_pos"] + np.random.default_rng(100 + r).normal(0, 0.3, base["vertex_pos"].shape)).astype(np.float32)
        return e
    mine = element(rank)
    N2 = mine["num_vertices"]
    ins = [T(mine[k]) for k in KEYS]
    G2 = torch.randn((1, C2, H, W, 3), generator=torch.Generator().manual_seed(4)).to(dev)
    r2 = _native.NativeRenderer(mine["faces"], mine["texcoords"], N2, C2, W, H, "vertexColor", "shaded", 1, 1, False, dev)

    def fwd_bwd(inp, outs=None):
        o = r2.forward(*inp)
        return r2.backward(G2, None, inp[0], inp[1], inp[2], inp[3], inp[4], o[3], o[0], o[1], inp[5], inp[6], out=outs)
    # what rank 0 computes alone over every rank's element
    want_col = torch.zeros((1, N2, 3), device=dev, dtype=torch.float64)
    want_sh = torch.zeros((1, C2, 27), device=dev, dtype=torch.float64)
    for r in range(world):
        e = element(r)
        gg = fwd_bwd([T(e[k]) for k in KEYS])
        want_col += gg[1].double(); want_sh += gg[3].double()
    gg = fwd_bwd(ins)
    nc_sh, nc_col = gg[3].clone(), gg[1].clone()
    sharding.allreduce_shared_grads([nc_sh, nc_col])
    bs = {"nccl": {"vertex_color_grad": rel(nc_col, want_col), "sh_coeff_grad": rel(nc_sh, want_sh)}}
    bufs = {}
    for mode in ("p2p", "nvls"):
        try:
            buf = sharding.SymmetricGradBuffer([(1, C2, 27), (1, N2, 3)], dev, mode=mode)
        except Exception as e:
            bs[mode] = {"unavailable": repr(e)[:200]}
            continue
        bufs[mode] = buf
        worst = {}
        for step in range(6):
            sh = sharding.SharedGrads(buf, step % 2, ("sh_coeff", "vertex_color"))
            buf.attach(r2, sh.slot)
            fwd_bwd(ins, sh.outputs())
            r2.set_allreduce(None)
            rs, rc = buf.results(sh.slot)
            torch.cuda.synchronize()
            worst["vertex_color_grad"] = max(worst.get("vertex_color_grad", 0.0), rel(rc, want_col))
            worst["sh_coeff_grad"] = max(worst.get("sh_coeff_grad", 0.0), rel(rs, want_sh))
            worst["bit_identical_across_ranks"] = True
            chk = rc.clone()
            dist.broadcast(chk, 0)
            worst["bit_identical_across_ranks"] = worst["bit_identical_across_ranks"] and bool(torch.equal(chk, rc))
        bs[mode] = worst
    out["batch_split_shared"] = bs

    # ---------------- 3. through the public layer, captured in CUDA graphs ----------------
    if "p2p" in bufs:
        buf = bufs["p2p"]
        faces_l, tcs_l = mine["faces"].reshape(-1), mine["texcoords"].reshape(-1)
        static = {k: ins[i].clone() for i, k in enumerate(KEYS)}

        def user_step(slot):
            leaves = {k: static[k].detach().requires_grad_(True) for k in ("vertex_pos", "vertex_color", "sh_coeff")}
            layer = CudaRendererGpu(faces_attr=faces_l, texCoords_attr=tcs_l, numberOfVertices_attr=N2, numberOfCameras_attr=C2,
                                    renderResolutionU_attr=W, renderResolutionV_attr=H, albedoMode_attr="vertexColor", shadingMode_attr="shaded",
                                    vertexPos_input=leaves["vertex_pos"], vertexColor_input=leaves["vertex_color"], texture_input=static["texture"],
                                    shCoeff_input=leaves["sh_coeff"], targetImage_input=static["target_image"], extrinsics_input=static["extrinsics"],
                                    intrinsics_input=static["intrinsics"], device=dev,
                                    sharedGrads_attr=sharding.SharedGrads(buf, slot, ("sh_coeff", "vertex_color")))
            loss = (layer.getRenderBufferTF() * G2).sum()
            return torch.autograd.grad(loss, [leaves["vertex_pos"], leaves["vertex_color"], leaves["sh_coeff"]])
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for s in (0, 1, 0, 1):
                user_step(s)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize(); dist.barrier()
        graphs, outs = [], []
        for s in (0, 1):
            gph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gph):
                outs.append(user_step(s))
            graphs.append(gph)
        worst = {"vertex_color_grad": 0.0, "sh_coeff_grad": 0.0}
        for step in range(8):
            graphs[step % 2].replay()
            torch.cuda.synchronize()
            worst["vertex_color_grad"] = max(worst["vertex_color_grad"], rel(outs[step % 2][1], want_col))
            worst["sh_coeff_grad"] = max(worst["sh_coeff_grad"], rel(outs[step % 2][2], want_sh))
        out["layer_in_cuda_graph"] = worst
        del graphs

    # ---------------- 4. cost per step (fwd+bwd of this small scene; the DIFFERENCE between the rows is the collective) ----------------
    def timed(fn, k=200):
        for _ in range(10):
            fn()
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            fn()
        e1.record(); torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / k], device=dev)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms) * 1e3
    flat, (fsh, fcol) = sharding.shared_grad_buffer([(1, C2, 27), (1, N2, 3)], dev)
    t = {"no_collective_us": timed(lambda: fwd_bwd(ins, (None, fcol, None, fsh)))}

    def nccl_step():
        fwd_bwd(ins, (None, fcol, None, fsh))
        sharding.allreduce_shared_grads([fsh, fcol])
    t["nccl_after_backward_us"] = timed(nccl_step)
    for mode, buf in bufs.items():
        cnt = [0]

        def fused_step():
            sh = sharding.SharedGrads(buf, cnt[0] % 2, ("sh_coeff", "vertex_color"))
            cnt[0] += 1
            buf.attach(r2, sh.slot)
            fwd_bwd(ins, sh.outputs())
        t[f"fused_{mode}_us"] = timed(fused_step)
        r2.set_allreduce(None)
    t["message_bytes"] = (C2 * 27 + N2 * 3) * 4
    out["cost_per_step"] = t

    if rank == 0:
        txt = json.dumps(out, indent=1)
        print(txt)
        if len(sys.argv) > 1:
            open(sys.argv[1], "w").write(txt)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
