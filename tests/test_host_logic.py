"""CPU: host-side logic -- synthetic scenes, batch partitioning, gloo world_size-2 collectives."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gvv_differentiable_cuda_renderer_b200 import sharding, synthetic


def test_uv_sphere_is_closed_and_outward():
    v, f, t = synthetic.uv_sphere(12, 16)
    assert t.shape == (len(f), 3, 2) and f.min() == 0 and f.max() == len(v) - 1
    e = np.sort(np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]]), 1)
    _, cnt = np.unique(e, axis=0, return_counts=True)
    assert (cnt == 2).all()
    n = np.cross(v[f[:, 1]] - v[f[:, 0]], v[f[:, 2]] - v[f[:, 0]])
    assert (np.einsum("ij,ij->i", n, v[f].mean(1)) > 0).all()
    assert ((t >= 0) & (t <= 1)).all()


def test_headline_mesh_size():
    v, f, _ = synthetic.uv_sphere(187, 188)
    assert abs(len(v) - 35000) < 500 and abs(len(f) - 70000) < 500


def test_make_scene_shapes():
    sc = synthetic.make_scene(cameras=3, width=40, height=24, batch=2)
    assert sc["extrinsics"].shape == (2, 36) and sc["intrinsics"].shape == (2, 27)
    assert sc["target_image"].shape == (2, 3, 24, 40, 3) and sc["sh_coeff"].shape == (2, 3, 27)
    E = sc["extrinsics"][0].reshape(3, 3, 4)
    for c in range(3):
        R = E[c, :, :3]
        assert np.allclose(R @ R.T, np.eye(3), atol=1e-5) and np.linalg.det(R) > 0


@pytest.mark.parametrize("B,W", [(64, 8), (7, 2), (3, 4), (1, 1), (9, 8)])
def test_partition_batch(B, W):
    parts = sharding.partition_batch(B, W)
    assert len(parts) == W and parts[0][0] == 0 and parts[-1][1] == B
    sizes = [e - s for s, e in parts]
    assert sum(sizes) == B and max(sizes) - min(sizes) <= 1
    assert all(parts[i][1] == parts[i + 1][0] for i in range(W - 1))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, B, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        full = {"vertex_pos": torch.arange(B * 4 * 3, dtype=torch.float32).reshape(B, 4, 3),
                "sh_coeff": torch.arange(B * 2 * 27, dtype=torch.float32).reshape(B, 2, 27)}
        loc = sharding.shard_inputs(full)
        s, e = sharding.local_slice(B)
        assert loc["vertex_pos"].shape[0] == e - s
        # per-rank gradient of batch-shared parameters = sum over the local slice
        g_sh = loc["sh_coeff"].sum(0)
        g_v = loc["vertex_pos"].sum(0)
        h = sharding.allreduce_shared_grads([g_sh, g_v], async_op=True)
        h.wait()
        ok1 = torch.allclose(g_sh, full["sh_coeff"].sum(0)) and torch.allclose(g_v, full["vertex_pos"].sum(0))
        # the same through ONE flat buffer (views handed to the backward as out=): reduced in place, no packing
        flat, (f_sh, f_v) = sharding.shared_grad_buffer([(2, 27), (4, 3)], "cpu")
        f_sh.copy_(loc["sh_coeff"].sum(0)); f_v.copy_(loc["vertex_pos"].sum(0))
        assert sharding._common_flat_buffer([f_sh, f_v]).data_ptr() == flat.data_ptr()
        assert sharding._common_flat_buffer([f_v, f_sh]) is None and sharding._common_flat_buffer([g_sh, g_v]) is None
        sharding.allreduce_shared_grads([f_sh, f_v])
        ok1 = ok1 and torch.allclose(f_sh, full["sh_coeff"].sum(0)) and torch.allclose(f_v, full["vertex_pos"].sum(0))
        ok1 = ok1 and torch.allclose(flat[:54], full["sh_coeff"].sum(0).reshape(-1))
        # per-batch-element rows are disjoint: gather restores the full batch in order
        gathered = sharding.gather_batch(loc["sh_coeff"] * 2.0, B)
        ok2 = torch.equal(gathered, full["sh_coeff"] * 2.0)
        q.put((rank, bool(ok1), bool(ok2)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("B", [4, 5])
def test_gloo_world2_shared_grad_allreduce_and_gather(B):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, B, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert sorted(r[0] for r in res) == [0, 1] and all(r[1] and r[2] for r in res)


@pytest.mark.parametrize("B,C,W", [(1, 8, 8), (2, 16, 8), (1, 3, 4), (3, 5, 8), (8, 16, 8), (64, 16, 8), (1, 1, 2)])
def test_plan_views_covers_every_view_exactly_once(B, C, W):
    plan = sharding.plan_views(B, C, W)
    assert len(plan) == W
    seen = np.zeros((B, C), int)
    for p in plan:
        if p is not None:
            b0, b1, c0, c1 = p
            assert 0 <= b0 < b1 <= B and 0 <= c0 < c1 <= C
            seen[b0:b1, c0:c1] += 1
    assert (seen == 1).all()                                     # every (batch element, camera) rendered by exactly one rank
    if B >= W:
        assert sharding.camera_teams(plan) == []                 # whole batch elements per rank: no collective inside the op
    else:
        for team in sharding.camera_teams(plan):
            assert len({plan[r][:2] for r in team}) == 1         # a team shares one batch element


def _fake_backward(inp):
    """Stand-in for the op's backward with its accumulation structure (CudaRendererGrad.cpp:264-283): position / colour /
    texture gradients are SUMS over the cameras of a batch element, sh_coeff_grad has one row per (b, c)."""
    w = inp["sh_coeff"].sum(-1)                                           # [b, c]: a per-view weight
    e = inp["extrinsics"].reshape(inp["extrinsics"].shape[0], -1, 12).sum(-1) + inp["intrinsics"].reshape(inp["intrinsics"].shape[0], -1, 9).sum(-1)
    per_view = w * e + inp["target_image"].sum((2, 3, 4))                # [b, c]
    gpos = inp["vertex_pos"] * per_view.sum(1)[:, None, None]
    gcol = inp["vertex_color"] * (per_view ** 2).sum(1)[:, None, None]
    gtex = inp["texture"] * per_view.sum(1)[:, None, None, None]
    gsh = inp["sh_coeff"] * per_view[:, :, None]
    return gpos, gcol, gtex, gsh


def _camera_split_worker(rank, world, port, B, C, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(5)
        full = {"vertex_pos": torch.randn(B, 6, 3, generator=g), "vertex_color": torch.rand(B, 6, 3, generator=g),
                "texture": torch.rand(B, 4, 4, 3, generator=g), "sh_coeff": torch.randn(B, C, 27, generator=g),
                "target_image": torch.rand(B, C, 5, 7, 3, generator=g), "extrinsics": torch.randn(B, C * 12, generator=g),
                "intrinsics": torch.randn(B, C * 9, generator=g)}
        plan = sharding.plan_views(B, C, world)
        groups = sharding.make_team_groups(plan)
        mine = plan[rank]
        want = _fake_backward(full)
        ok = True
        if mine is not None:
            loc = sharding.shard_views(full, C, mine)
            assert loc["sh_coeff"].shape == (mine[1] - mine[0], mine[3] - mine[2], 27)
            assert loc["extrinsics"].shape == (mine[1] - mine[0], (mine[3] - mine[2]) * 12)
            grads = sharding.reduce_camera_split(_fake_backward(loc), plan, C, rank=rank, groups=groups)
            for got, ref in zip(grads, want):
                ok = ok and torch.allclose(got, ref[mine[0]:mine[1]], rtol=1e-5, atol=1e-5)
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("B,C", [(1, 4), (1, 3), (2, 5)])
def test_gloo_world2_camera_split_reduces_like_one_process(B, C):
    """Cameras of one batch element split over two ranks: after reduce_camera_split every rank holds the gradients
    a single process computes over all cameras (B >= world degenerates to the batch split, no collective)."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_camera_split_worker, args=(r, 2, port, B, C, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert sorted(r[0] for r in res) == [0, 1] and all(r[1] for r in res)
