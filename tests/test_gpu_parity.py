"""GPU (-m gpu): parity of the CUDA path, called through the C ABI.

  1. vs the golden fixtures = outputs of the reference's own CUDA core (tests/golden/README.md):
     face buffer bit-exact except exact depth ties (the reference's own choice there is a data
     race, CUDABasedRasterization.cu:295-301; ours is the smallest triangle id) -- every mismatch
     is PROVEN to be an exact tie by re-evaluating both candidates with gvv_debug_eval and
     comparing with the reference's depth buffer; barycentrics bit-exact; render / normals to
     1e-6; gradients rel-L2 <= 1e-4 and max-abs <= 1e-3 * max|g| (atomic order differs).
  2. vs Oracle 2 (CPU) on fresh seeded inputs, near-tie protocol of tests/test_oracle_golden.py.
  3. vs Oracle 1 (oracle/_ref/libgvv_ref.so) directly when it was shipped: larger scenes.
  4. size-independent properties at the headline size: determinism, tile-size invariance,
     linearity in colour / SH, finite differences of the linear inputs.
"""
import numpy as np
import pytest
import torch

from conftest import golden_files, golden_ids, load_golden, rel_l2
from gvv_differentiable_cuda_renderer_b200 import _native, synthetic

pytestmark = pytest.mark.gpu
INPUT_KEYS = ("vertex_pos", "vertex_color", "texture", "sh_coeff", "target_image", "extrinsics", "intrinsics")


def dev():
    return torch.device("cuda:0")


def T(x):
    return torch.as_tensor(np.ascontiguousarray(x), device=dev())


def make(sc, albedo, shading, tile=32, image_filter=1):
    r = _native.NativeRenderer(sc["faces"], sc["texcoords"], sc["num_vertices"], sc["num_cameras"], sc["width"], sc["height"],
                               albedo, shading, image_filter, 1, False, dev())
    r.set_option("tile", tile)
    return r


def assert_faces_equal_up_to_exact_ties(r, face, ref_face, ref_depth=None):
    """Every differing pixel must be an exact tie: both candidates hit with the same depth key
    (and that key equals the reference's depth buffer when given)."""
    face = face.cpu().numpy() if torch.is_tensor(face) else face
    ref_face = ref_face.cpu().numpy() if torch.is_tensor(ref_face) else ref_face
    mism = np.argwhere(face != ref_face)
    if len(mism) == 0:
        return 0
    C, H, W = face.shape[1:]
    mine = face[tuple(mism.T)]
    theirs = ref_face[tuple(mism.T)]
    assert (mine >= 0).all() and (theirs >= 0).all(), "coverage differs from the reference"
    view = mism[:, 0] * C + mism[:, 1]
    q1 = np.stack([view, mism[:, 3], mism[:, 2], mine], 1)
    q2 = np.stack([view, mism[:, 3], mism[:, 2], theirs], 1)
    k1, _ = r.eval_pairs(q1)
    k2, _ = r.eval_pairs(q2)
    assert np.array_equal(k1, k2), "face mismatch that is not an exact depth tie"
    assert (mine < theirs).all(), "tie not resolved to the smallest triangle id"
    if ref_depth is not None:
        rd = ref_depth.cpu().numpy() if torch.is_tensor(ref_depth) else ref_depth
        assert np.array_equal(k1, rd[tuple(mism.T)])
    return len(mism)


def grads_close(mine, ref, rel=1e-4, mx=1e-3):
    for name, a, b in zip(("vertex_pos_grad", "vertex_color_grad", "texture_grad", "sh_coeff_grad"), mine, ref):
        a = a.cpu().numpy() if torch.is_tensor(a) else a
        b = b.cpu().numpy() if torch.is_tensor(b) else b
        assert rel_l2(a, b) <= rel, (name, rel_l2(a, b))
        if np.abs(b).max() > 0:
            assert np.abs(a - b).max() <= mx * np.abs(b).max(), name


@pytest.mark.skipif(not golden_files(), reason="no golden fixtures")
@pytest.mark.parametrize("tile", [32, 16])
@pytest.mark.parametrize("path", golden_files(), ids=golden_ids())
def test_cuda_matches_reference_golden(path, tile):
    g = load_golden(path)
    r = make(g, g["albedo"], g["shading"], tile)
    ins = [T(g[k]) for k in INPUT_KEYS]
    bary, face, render, vn, tout, _ = r.forward(*ins)
    torch.cuda.synchronize()
    assert_faces_equal_up_to_exact_ties(r, face, g["ref_face"], g["ref_depth"])
    same = torch.as_tensor(face.cpu().numpy() == g["ref_face"])
    b_, rb = bary.cpu(), torch.as_tensor(g["ref_bary"])
    assert torch.equal(b_.view(torch.int32)[same], rb.view(torch.int32)[same]), "barycentrics not bit-exact"
    assert np.abs(render.cpu().numpy() - g["ref_render"])[same.numpy()].max() <= 1e-6
    rvn = g["ref_vertex_normal"]
    assert np.abs(vn.cpu().numpy() - rvn).max() <= 1e-6 * np.abs(rvn).max()
    assert torch.equal(tout.cpu(), torch.as_tensor(g["target_image"]))
    if "ref_vertex_pos_grad" in g:
        tg = T(g["target_grad"]) if "target_grad" in g else None
        gm = r.backward(T(g["render_grad"]), tg, ins[0], ins[1], ins[2], ins[3], ins[4], T(g["ref_vertex_normal"]),
                        T(g["ref_bary"]), T(g["ref_face"]), ins[5], ins[6])
        grads_close(gm, [g["ref_vertex_pos_grad"], g["ref_vertex_color_grad"], g["ref_texture_grad"], g["ref_sh_coeff_grad"]])
    r.close()


SCENES = [
    ("sphere", dict(rings=20, segments=26, cameras=2, width=96, height=80, batch=2, tex=32), "vertexColor", "shaded"),
    ("sphere", dict(rings=14, segments=18, cameras=1, width=70, height=70, tex=48), "textured", "shaded"),
    ("pyramid", dict(cameras=2, width=100, height=60, tex=16), "vertexColor", "shadeless"),
    ("triangle", dict(cameras=1, width=33, height=47, tex=16), "normal", "shaded"),
    ("sphere", dict(rings=10, segments=12, cameras=1, width=64, height=64, tex=16), "lighting", "shaded"),
    ("sphere", dict(rings=10, segments=12, cameras=1, width=64, height=64, tex=16), "foregroundMask", "shaded"),
]


@pytest.mark.parametrize("kind,kw,albedo,shading", SCENES, ids=[f"{s[0]}-{s[2]}-{s[3]}" for s in SCENES])
def test_cuda_matches_cpu_oracle(kind, kw, albedo, shading):
    from oracle import cpu
    sc = synthetic.make_scene(kind=kind, seed=11, **kw)
    N, C, W, H = sc["num_vertices"], sc["num_cameras"], sc["width"], sc["height"]
    o = cpu.forward(sc["faces"], sc["texcoords"], N, C, W, H, albedo, shading, sc["vertex_pos"], sc["vertex_color"],
                    sc["texture"], sc["sh_coeff"], sc["extrinsics"], sc["intrinsics"])
    r = make(sc, albedo, shading)
    ins = [T(sc[k]) for k in INPUT_KEYS]
    bary, face, render, vn, _, _ = r.forward(*ins)
    f = face.cpu().numpy()
    mism = f != o["face"]
    both = (f >= 0) & (o["face"] >= 0)
    gap = o["second_depth"].astype(np.int64) - o["best_depth"].astype(np.int64)
    assert np.all((gap[mism & both] <= 8) | (o["tie"][mism & both] == 1))
    assert (mism & ~both).sum() <= max(2, 0.002 * both.sum())
    same = ~mism & (f >= 0)
    assert np.abs(bary.cpu().numpy() - o["bary"])[same].max() <= 5e-4
    assert np.abs(render.cpu().numpy() - o["render"])[same].max() <= 5e-4
    assert np.abs(vn.cpu().numpy() - o["vertex_normal"]).max() <= 1e-5 * np.abs(o["vertex_normal"]).max()
    if albedo in ("vertexColor", "textured", "foregroundMask"):
        rng = np.random.default_rng(3)
        B = sc["vertex_pos"].shape[0]
        rg = rng.standard_normal((B, C, H, W, 3)).astype(np.float32)
        tg = rng.standard_normal((B, C, H, W, 3)).astype(np.float32)
        tgt = rng.random((B, C, H, W, 3), dtype=np.float32)
        # same forward buffers for both, so this isolates the backward
        gm = r.backward(T(rg), T(tg), ins[0], ins[1], ins[2], ins[3], T(tgt), vn, bary, face, ins[5], ins[6])
        go = cpu.backward(sc["faces"], sc["texcoords"], N, C, W, H, albedo, shading, 1, rg, tg, sc["vertex_pos"], sc["vertex_color"],
                          sc["texture"], sc["sh_coeff"], tgt, vn.cpu().numpy(), bary.cpu().numpy(), f, sc["extrinsics"], sc["intrinsics"])
        grads_close(gm, go)
    r.close()


def _ref_available():
    from oracle import ref
    return ref.available()


@pytest.mark.skipif(not _ref_available(), reason="oracle/_ref/libgvv_ref.so not shipped")
@pytest.mark.parametrize("albedo,shading,size,rings,segs", [("vertexColor", "shaded", 512, 96, 128), ("textured", "shadeless", 384, 64, 80)])
def test_cuda_matches_reference_cuda_core_live(albedo, shading, size, rings, segs):
    from oracle import ref as oref
    sc = synthetic.make_scene(kind="sphere", rings=rings, segments=segs, cameras=3, width=size, height=size, batch=2, tex=128, seed=5)
    N, C, W, H = sc["num_vertices"], sc["num_cameras"], size, size
    ins = [T(sc[k]) for k in INPUT_KEYS]
    ref = oref.RefRenderer(sc["faces"], sc["texcoords"], N, C, W, H, albedo, shading)
    rr = ref.forward(*ins, intermediates=True)
    r = make(sc, albedo, shading)
    bary, face, render, vn, _, _ = r.forward(*ins)
    n_tie = assert_faces_equal_up_to_exact_ties(r, face, rr["face"], rr["depth"])
    assert n_tie <= 1e-4 * face.numel()
    same = face == rr["face"]
    assert torch.equal(bary.view(torch.int32)[same], rr["bary"].view(torch.int32)[same])
    assert float((render - rr["render"]).abs()[same].max()) <= 1e-6
    assert torch.equal(vn.view(torch.int32), rr["vertex_normal"].view(torch.int32))
    g = torch.Generator().manual_seed(1)
    rg = torch.randn((2, C, H, W, 3), generator=g).to(dev())
    gm = r.backward(rg, None, ins[0], ins[1], ins[2], ins[3], ins[4], rr["vertex_normal"], rr["bary"], rr["face"], ins[5], ins[6])
    gr = ref.backward(rg, ins[0], ins[1], ins[2], ins[3], ins[4], rr["vertex_normal"], rr["bary"], rr["face"], None, ins[5], ins[6])
    grads_close(gm, gr)
    r.close()


@pytest.mark.skipif(not _ref_available(), reason="oracle/_ref/libgvv_ref.so not shipped")
@pytest.mark.parametrize("albedo,shuffled", [("normal", False), ("textured", False), ("lighting", False), ("normal", True)])
def test_stress_resolution_4k_matches_reference(albedo, shuffled):
    """BASELINE.json config 5 shape (3840x2160, forward, visibility bound) at a triangle count the
    reference's O(N*F) constructor can still digest.  8160 tiles: the binning kernels keep their histograms over a
    window of tile rows per block (ring-ordered faces: short bands); with the face order shuffled every block's band
    is the whole image, which does not fit the window and takes the global-atomic fallback."""
    from oracle import ref as oref
    sc = synthetic.make_scene(kind="sphere", rings=160, segments=200, cameras=1, width=3840, height=2160, tex=256, seed=3,
                              coverage_radius_frac=0.26)
    if shuffled:
        perm = np.random.default_rng(5).permutation(len(sc["faces"]))
        sc["faces"] = np.ascontiguousarray(sc["faces"][perm])
        sc["texcoords"] = np.ascontiguousarray(sc["texcoords"][perm])      # [F, 3, 2], per-corner
    N, W, H = sc["num_vertices"], 3840, 2160
    ins = [T(sc[k]) for k in INPUT_KEYS]
    ref = oref.RefRenderer(sc["faces"], sc["texcoords"], N, 1, W, H, albedo, "shaded", with_backward=False)
    rr = ref.forward(*ins, intermediates=True)
    r = make(sc, albedo, "shaded")
    bary, face, render, vn, _, _ = r.forward(*ins)
    n_tie = assert_faces_equal_up_to_exact_ties(r, face, rr["face"], rr["depth"])
    assert n_tie <= 1e-4 * face.numel()
    same = face == rr["face"]
    assert float(same.float().mean()) > 0.9999 and 0.3 < float((face >= 0).float().mean()) < 0.7
    assert torch.equal(bary.view(torch.int32)[same], rr["bary"].view(torch.int32)[same])
    assert float((render - rr["render"]).abs()[same].max()) <= 1e-6
    r.close()


@pytest.mark.skipif(not _ref_available(), reason="oracle/_ref/libgvv_ref.so not shipped")
def test_huge_triangles_big_list_path_matches_reference():
    """BASELINE.json config 1 shape: a handful of triangles with image-sized bounding boxes (they go
    through the per-view big-triangle list, not the tile bins), 1024x1024, B=2."""
    from oracle import ref as oref
    sc = synthetic.make_scene(kind="pyramid", cameras=1, width=1024, height=1024, batch=2, tex=16, seed=1, distance=900.0)
    N = sc["num_vertices"]
    ins = [T(sc[k]) for k in INPUT_KEYS]
    ref = oref.RefRenderer(sc["faces"], sc["texcoords"], N, 1, 1024, 1024, "vertexColor", "shaded")
    rr = ref.forward(*ins, intermediates=True)
    r = make(sc, "vertexColor", "shaded")
    bary, face, render, vn, _, _ = r.forward(*ins)
    assert_faces_equal_up_to_exact_ties(r, face, rr["face"], rr["depth"])
    same = face == rr["face"]
    assert float((face >= 0).float().mean()) > 0.1
    assert torch.equal(bary.view(torch.int32)[same], rr["bary"].view(torch.int32)[same])
    assert float((render - rr["render"]).abs()[same].max()) <= 1e-6
    g = torch.Generator().manual_seed(1)
    rg = torch.randn(render.shape, generator=g).to(dev())
    gm = r.backward(rg, None, ins[0], ins[1], ins[2], ins[3], ins[4], rr["vertex_normal"], rr["bary"], rr["face"], ins[5], ins[6])
    gr = ref.backward(rg, ins[0], ins[1], ins[2], ins[3], ins[4], rr["vertex_normal"], rr["bary"], rr["face"], None, ins[5], ins[6])
    grads_close(gm, gr)
    r.close()


@pytest.fixture(scope="module")
def headline():
    """SURVEY.md 8d config 2 at full size: ~35k verts / 70k tris, 8 cameras, 1024^2."""
    sc = synthetic.make_scene(kind="sphere", rings=187, segments=188, cameras=8, width=1024, height=1024, batch=1, tex=64)
    return sc, [T(sc[k]) for k in INPUT_KEYS]


def test_headline_size_properties(headline):
    sc, ins = headline
    C, H, W = 8, 1024, 1024
    r32 = make(sc, "vertexColor", "shaded", 32)
    out1 = r32.forward(*ins)
    out2 = r32.forward(*ins)
    for a, b in zip(out1[:4], out2[:4]):                       # determinism (no racy ties): identical bits
        assert torch.equal(a.view(torch.int32) if a.dtype == torch.float32 else a, b.view(torch.int32) if b.dtype == torch.float32 else b)
    r16 = make(sc, "vertexColor", "shaded", 16)
    out3 = r16.forward(*ins)
    for a, b in zip(out1[:4], out3[:4]):                       # tiling changes who evaluates, never what
        assert torch.equal(a.view(torch.int32) if a.dtype == torch.float32 else a, b.view(torch.int32) if b.dtype == torch.float32 else b)
    # the conservative screen-space pre-test only skips pairs that fail the exact test anyway:
    # identical bits with it switched off (every bbox pixel tested, like the reference) or tightened
    for margin in (-1, 16):
        rc = make(sc, "vertexColor", "shaded", 32)
        rc.set_option("cull_margin_milli", margin)
        for a, b in zip(out1[:4], rc.forward(*ins)[:4]):
            assert torch.equal(a.view(torch.int32) if a.dtype == torch.float32 else a, b.view(torch.int32) if b.dtype == torch.float32 else b)
        rc.close()
    bary, face, render, vn = out1[:4]
    cov = float((face >= 0).float().mean())
    assert 0.45 < cov < 0.60                                   # ~50 % coverage, as the workload is specified
    bg = face < 0
    assert torch.all(render[bg] == torch.tensor([0.0, 1.0, 0.0], device=dev())) and torch.all(bary[bg] == 0)
    fg = ~bg
    assert float(bary[fg].min()) >= -0.0011 and float(bary[fg].sum(-1).max()) <= 1.0011
    assert int(face.max()) < len(sc["faces"])
    # linearity of the render buffer in vertex colour: R(c1 + c2) = R(c1) + R(c2) on covered pixels
    c2 = torch.rand_like(ins[1])
    ra = r32.forward(ins[0], ins[1] + c2, *ins[2:])[2]
    rb = r32.forward(ins[0], c2, *ins[2:])[2]
    assert float((ra - render - rb)[fg].abs().max()) <= 2e-5
    # backward: sums of colour gradient weights = sums of (g * light) over covered pixels, per channel
    g = torch.randn((1, C, H, W, 3), generator=torch.Generator().manual_seed(3)).to(dev())
    gpos, gcol, gtex, gsh = r32.backward(g, None, ins[0], ins[1], ins[2], ins[3], ins[4], vn, bary, face, ins[5], ins[6])
    # dL/dcolour contracted with the colours = <g, render> (render is linear and homogeneous in colour)
    lhs = float((gcol.double() * ins[1].double()).sum())
    rhs = float((g.double() * render.double())[fg].sum())
    assert abs(lhs - rhs) <= 1e-4 * max(1.0, abs(rhs)), (lhs, rhs)
    # same for SH (render is linear and homogeneous in the SH coefficients in shaded mode)
    lhs = float((gsh.double() * ins[3].double()).sum())
    assert abs(lhs - rhs) <= 1e-4 * max(1.0, abs(rhs)), (lhs, rhs)
    assert torch.isfinite(gpos).all() and float(gpos.abs().max()) > 0 and float(gtex.abs().max()) == 0
    gpos2 = r32.backward(g, None, ins[0], ins[1], ins[2], ins[3], ins[4], vn, bary, face, ins[5], ins[6])[0]
    assert rel_l2(gpos2.cpu().numpy(), gpos.cpu().numpy()) <= 1e-5    # atomic order only
    r32.close(); r16.close()


@pytest.mark.parametrize("albedo", ["vertexColor", "textured"])
def test_finite_differences_of_linear_inputs_gpu(albedo):
    sc = synthetic.make_scene(kind="sphere", rings=12, segments=16, cameras=2, width=64, height=64, tex=24, seed=2)
    r = make(sc, albedo, "shaded")
    ins = [T(sc[k]) for k in INPUT_KEYS]
    bary, face, render, vn, _, _ = r.forward(*ins)
    rg = torch.randn(render.shape, generator=torch.Generator().manual_seed(0)).to(dev())
    gpos, gcol, gtex, gsh = r.backward(rg, None, ins[0], ins[1], ins[2], ins[3], ins[4], vn, bary, face, ins[5], ins[6])

    def loss(i, x):
        a = list(ins); a[i] = x
        return float((r.forward(*a)[2].double() * rg.double()).sum())

    eps = 1e-2
    which = [(1, gcol), (3, gsh)] if albedo == "vertexColor" else [(2, gtex)]
    for i, grad in which:
        k = int(grad.abs().argmax())
        d = torch.zeros_like(ins[i]).view(-1)
        d[k] = eps
        d = d.view_as(ins[i])
        fd = (loss(i, ins[i] + d) - loss(i, ins[i] - d)) / (2 * eps)
        an = float(grad.view(-1)[k])
        assert abs(fd - an) <= 5e-3 * abs(an) + 1e-4, (albedo, i, fd, an)
    r.close()


def test_batched_call_equals_per_element_calls():
    """One call over B batch elements (what replaces the reference's serial host loop,
    CudaRenderer.cpp:309-328) gives the same bits as B separate calls, forward and backward."""
    sc = synthetic.make_scene(kind="sphere", rings=40, segments=50, cameras=3, width=160, height=120, batch=5, tex=32, seed=8)
    N, C, W, H = sc["num_vertices"], 3, 160, 120
    rng = np.random.default_rng(0)
    sc["sh_coeff"] = (sc["sh_coeff"] + rng.random(sc["sh_coeff"].shape, dtype=np.float32) * 0.2).astype(np.float32)
    ins = [T(sc[k]) for k in INPUT_KEYS]
    r = make(sc, "textured", "shaded")
    out = r.forward(*ins)
    rg = torch.randn(out[2].shape, generator=torch.Generator().manual_seed(2)).to(dev())
    g_all = r.backward(rg, None, ins[0], ins[1], ins[2], ins[3], ins[4], out[3], out[0], out[1], ins[5], ins[6])
    for b in range(5):
        one = [t[b:b + 1].contiguous() for t in ins]
        ob = r.forward(*one)
        for full, part in zip(out[:4], ob[:4]):
            a_, b_ = full[b:b + 1], part
            assert torch.equal(a_.view(torch.int32) if a_.dtype == torch.float32 else a_, b_.view(torch.int32) if b_.dtype == torch.float32 else b_)
        gb = r.backward(rg[b:b + 1].contiguous(), None, one[0], one[1], one[2], one[3], one[4], ob[3], ob[0], ob[1], one[5], one[6])
        for full, part in zip(g_all, gb):
            assert rel_l2(part.cpu().numpy(), full[b:b + 1].cpu().numpy()) <= 1e-5     # atomic order only
    r.close()


@pytest.mark.parametrize("albedo,shading,with_target", [("vertexColor", "shaded", False), ("textured", "shaded", True), ("textured", "shadeless", False)])
def test_persistent_backward_variant_matches_one_tile_per_cta(albedo, shading, with_target):
    """Option bwd_persistent (persistent CTAs pulling tiles from a counter, face tiles through a TMA ring on mbarriers;
    off by default, measured slower) computes the same gradients as the one-tile-per-CTA kernel; it only engages on
    images made of whole 32x32 tiles, so the odd-sized call must fall back silently."""
    for (W, H) in ((256, 192), (200, 168)):
        sc = synthetic.make_scene(kind="sphere", rings=48, segments=56, cameras=3, width=W, height=H, batch=2, tex=32, seed=5)
        ins = [T(sc[k]) for k in INPUT_KEYS]
        r = make(sc, albedo, shading)
        out = r.forward(*ins)
        rg = torch.randn(out[2].shape, generator=torch.Generator().manual_seed(4)).to(dev())
        tg = torch.randn(out[2].shape, generator=torch.Generator().manual_seed(6)).to(dev()) if with_target else None
        args = (rg, tg, ins[0], ins[1], ins[2], ins[3], ins[4], out[3], out[0], out[1], ins[5], ins[6])
        g0 = [t.clone() for t in r.backward(*args)]
        r.set_option("bwd_persistent", 1)
        n0 = r.launch_count
        g1 = r.backward(*args)
        assert r.launch_count - n0 == (3 if shading == "shaded" else 2)
        for name, a, b in zip(("vertex_pos_grad", "vertex_color_grad", "texture_grad", "sh_coeff_grad"), g1, g0):
            assert rel_l2(a.cpu().numpy(), b.cpu().numpy()) <= 2e-5, (name, W, H)     # summation order only
        r.close()


def test_uv_space_normal_map_matches_cpu_oracle():
    """compute_normal_map: rasterisation is replaced by the UV-space normal map."""
    from oracle import cpu
    sc = synthetic.make_scene(kind="sphere", rings=14, segments=18, cameras=2, width=32, height=32, batch=2, tex=96, seed=9)
    N, C = sc["num_vertices"], sc["num_cameras"]
    r = _native.NativeRenderer(sc["faces"], sc["texcoords"], N, C, 32, 32, "textured", "shaded", 1, 1, True, dev())
    ins = [T(sc[k]) for k in INPUT_KEYS]
    bary, face, render, vn, _, nmap = r.forward(*ins)
    o = cpu.normal_map(sc["faces"], sc["texcoords"], N, C, sc["vertex_pos"], 96, 96)
    assert nmap.shape == (2, 96, 96, 3) and int((face >= 0).sum()) == 0        # no rasterisation in this mode
    assert np.abs(vn.cpu().numpy() - o["vertex_normal"]).max() <= 1e-5 * np.abs(o["vertex_normal"]).max()
    # texels covered by exactly one UV triangle are unambiguous; shared UV edges may pick either neighbour
    clean = (o["tie"] == 0)
    d = np.abs(nmap.cpu().numpy() - o["normal_map"])[:, clean]
    assert d.max() <= 2e-4, float(d.max())
    assert clean.mean() > 0.8 and o["covered"].mean() > 0.5
    r.close()


def test_edge_cases():
    # fully off-screen mesh, degenerate triangle, non-multiple-of-tile resolution, isolated vertex
    verts = np.array([[0, 0, 0], [50, 0, 0], [0, 50, 0], [10, 10, 10], [10, 10, 10], [999, 999, 999]], np.float32)
    faces = np.array([[0, 1, 2], [3, 4, 3]], np.int32)
    tcs = np.zeros((2, 3, 2), np.float32)
    E, K = synthetic.ring_cameras(1, 800.0, 37, 21, 300.0)
    sc = dict(faces=faces, texcoords=tcs, num_vertices=6, num_cameras=1, width=37, height=21)
    ins = [T(verts[None]), T(np.ones((1, 6, 3), np.float32)), T(np.ones((1, 4, 4, 3), np.float32)), T(synthetic.base_sh()[None, None]),
           T(np.zeros((1, 1, 21, 37, 3), np.float32)), T(E.reshape(1, -1)), T(K.reshape(1, -1))]
    r = make(sc, "vertexColor", "shaded")
    bary, face, render, vn, _, _ = r.forward(*ins)
    assert set(np.unique(face.cpu().numpy())) <= {-1, 0}           # the degenerate face never wins
    assert float(vn[0, 0, 5].abs().max()) == 0.0                    # isolated vertex: defined as 0
    far = [T(verts[None] + 1e6)] + ins[1:]
    face2 = r.forward(*far)[1]
    assert int((face2 >= 0).sum()) == 0                             # off screen -> all background
    g = r.backward(torch.ones_like(render), None, far[0], far[1], far[2], far[3], far[4], vn, bary, face2, far[5], far[6])
    assert all(float(x.abs().max()) == 0 for x in g)                # nothing visible -> zero gradients
    r.close()
    with pytest.raises(_native.GvvError):
        _native.NativeRenderer(faces, None, 6, 1, 8, 8, "textured", "shaded", device=dev())


def test_bilinear_texture_variant():
    """texture_bilinear = 1 (non-default): the bilinear fetch (CUDABasedRasterization.cu:365-372) and the four
    weighted texture-gradient adds (CUDABasedRasterizationGrad.cu:361-378) the reference has commented out.
    Checked against Oracle 2 with the same switch, by finite differences (the texture is a linear input and the
    scatter is now the exact adjoint of the fetch) and by the weights summing to one."""
    from oracle import cpu
    sc = synthetic.make_scene(kind="sphere", rings=14, segments=18, cameras=2, width=72, height=64, tex=20, seed=5, noise=0.0)
    N, C, W, H = sc["num_vertices"], sc["num_cameras"], sc["width"], sc["height"]
    ins = [T(sc[k]) for k in INPUT_KEYS]
    r = make(sc, "textured", "shaded")
    bary0, face0, render0, vn0, _, _ = r.forward(*ins)
    rg = torch.randn(render0.shape, generator=torch.Generator().manual_seed(1)).to(dev())
    g_nearest = r.backward(rg, None, ins[0], ins[1], ins[2], ins[3], ins[4], vn0, bary0, face0, ins[5], ins[6])
    r.set_option("texture_bilinear", 1)
    bary, face, render, vn, _, _ = r.forward(*ins)
    assert torch.equal(face, face0) and torch.equal(bary, bary0)          # visibility does not depend on the texture filter
    assert not torch.equal(render, render0)
    g_bil = r.backward(rg, None, ins[0], ins[1], ins[2], ins[3], ins[4], vn, bary, face, ins[5], ins[6])
    # (a) oracle with the same switch
    cpu.set_texture_bilinear(True)
    try:
        o = cpu.forward(sc["faces"], sc["texcoords"], N, C, W, H, "textured", "shaded", sc["vertex_pos"], sc["vertex_color"],
                        sc["texture"], sc["sh_coeff"], sc["extrinsics"], sc["intrinsics"])
        f = face.cpu().numpy()
        same = (f == o["face"]) & (f >= 0)
        assert same.sum() > 0.95 * (f >= 0).sum()
        assert np.abs(render.cpu().numpy() - o["render"])[same].max() <= 5e-4
        go = cpu.backward(sc["faces"], sc["texcoords"], N, C, W, H, "textured", "shaded", 1, rg.cpu().numpy(), None, sc["vertex_pos"],
                          sc["vertex_color"], sc["texture"], sc["sh_coeff"], sc["target_image"], vn.cpu().numpy(), bary.cpu().numpy(), f,
                          sc["extrinsics"], sc["intrinsics"])
    finally:
        cpu.set_texture_bilinear(False)
    grads_close(g_bil, go)
    # (b) the four weights sum to one: the same total lands in the texture gradient as with the nearest-texel add
    tot_b, tot_n = g_bil[2].double().sum((0, 1, 2)), g_nearest[2].double().sum((0, 1, 2))
    assert torch.allclose(tot_b, tot_n, rtol=1e-4, atol=1e-3), (tot_b, tot_n)
    # (c) finite differences along a random texture direction (fp64 accumulation of the loss).  The reference
    # skips the texture gradient where the shading normal was flipped (grazing pixels at the silhouette,
    # CUDABasedRasterizationGrad.cu:345), so the loss only looks at the central disc of the sphere.
    yy, xx = np.mgrid[0:H, 0:W]
    disc = ((xx - W / 2) ** 2 + (yy - H / 2) ** 2) <= (0.6 * 0.4 * W) ** 2
    rgc = rg * T(disc.astype(np.float32))[None, None, :, :, None]
    g_c = r.backward(rgc, None, ins[0], ins[1], ins[2], ins[3], ins[4], vn, bary, face, ins[5], ins[6])
    D = torch.randn(ins[2].shape, generator=torch.Generator().manual_seed(2)).to(dev())

    def loss(tex):
        a = list(ins); a[2] = tex
        return float((r.forward(*a)[2].double() * rgc.double()).sum())

    eps = 1e-2
    fd = (loss(ins[2] + eps * D) - loss(ins[2] - eps * D)) / (2 * eps)
    an = float((g_c[2].double() * D.double()).sum())
    assert abs(an) > 1.0 and abs(fd - an) <= 2e-3 * abs(an) + 1e-3, (fd, an)
    r.close()


def test_forward_backward_are_cuda_graph_capturable():
    """Steady-state gvv_forward / gvv_backward never allocate or synchronise (include/gvv_b200.h), so one
    fwd+bwd step can be captured in a CUDA graph and replayed: same bits in the integer/forward outputs,
    gradients equal up to atomic order."""
    sc = synthetic.make_scene(kind="sphere", rings=24, segments=30, cameras=2, width=160, height=128, tex=16, seed=6)
    r = make(sc, "vertexColor", "shaded")
    ins = [T(sc[k]) for k in INPUT_KEYS]
    rg = torch.randn((1, sc["num_cameras"], sc["height"], sc["width"], 3), generator=torch.Generator().manual_seed(0)).to(dev())
    bary0, face0, render0, vn0, _, _ = r.forward(*ins)                       # eager: also sizes the scratch
    g0 = r.backward(rg, None, ins[0], ins[1], ins[2], ins[3], ins[4], vn0, bary0, face0, ins[5], ins[6])
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        bary, face, render, vn, _, _ = r.forward(*ins)
        g = r.backward(rg, None, ins[0], ins[1], ins[2], ins[3], ins[4], vn, bary, face, ins[5], ins[6])
    for _ in range(3):
        face.fill_(-7); render.zero_(); g[0].fill_(123.0)
        graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(face, face0) and torch.equal(bary.view(torch.int32), bary0.view(torch.int32))
    assert torch.equal(render.view(torch.int32), render0.view(torch.int32)) and torch.equal(vn, vn0)
    grads_close(g, g0, rel=1e-5)
    # the inputs are read at replay time: new colours -> new image, same visibility
    ins[1].mul_(0.5)
    graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(face, face0) and not torch.equal(render, render0)
    del graph
    r.close()


OPTION_SETS = [
    {"hiz": 0}, {"span_z": 0}, {"span_z": 1}, {"hiz": 0, "span_z": 0, "cull_margin_milli": 250},
    {"split_unit": 48}, {"split_unit": 16, "hiz": 0}, {"heavy_mode": 2, "heavy_thr": 32}, {"heavy_mode": 2, "heavy_thr": 32, "heavy_slots": 3},
    {"heavy_mode": 0}, {"spread_empty": 1}, {"spread_empty": 1, "split_unit": 32}, {"cta_threads": 128}, {"tile": 16}, {"tile": 16, "split_unit": 24},
    {"interleave": 0}, {"batch_div": 3}, {"batch_div": 24}, {"ray_cache": 1}, {"resolve_prefetch": 1}, {"bulk_out": 1},
]


@pytest.mark.parametrize("opts", OPTION_SETS, ids=[",".join(f"{k}={v}" for k, v in o.items()) for o in OPTION_SETS])
def test_every_tuning_knob_keeps_the_bits(opts):
    """All scheduling / culling knobs of gvv_set_option only change WHO evaluates a (pixel, triangle) pair or
    whether a pair that provably cannot win is evaluated at all: face, barycentric and render buffers must be
    bit-identical to the configuration that tests every bbox pixel in one pass (the reference's schedule)."""
    sc = synthetic.make_scene(kind="sphere", rings=60, segments=64, cameras=3, width=200, height=168, tex=16, seed=8)
    ins = [T(sc[k]) for k in INPUT_KEYS]
    base = make(sc, "vertexColor", "shaded")
    for k, v in {"cull_margin_milli": -1, "hiz": 0, "span_z": 0, "heavy_mode": 0}.items():
        base.set_option(k, v)
    b0, f0, r0, v0, _, _ = base.forward(*ins)
    r = make(sc, "vertexColor", "shaded", tile=opts.get("tile", 32))
    for k, v in opts.items():
        if k != "tile":
            r.set_option(k, v)
    for _ in range(2):                                   # twice: the self-cleaning scratch must be back in its initial state
        b1, f1, r1, v1, _, _ = r.forward(*ins)
        assert torch.equal(f1, f0)
        assert torch.equal(b1.view(torch.int32), b0.view(torch.int32))
        assert torch.equal(r1.view(torch.int32), r0.view(torch.int32))
        assert torch.equal(v1, v0)
    base.close(); r.close()


def test_heavy_tile_pair_stress_changing_geometry():
    """The heavy-tile launch and the 256-thread launch that follows it as a programmatic dependent launch overlap on
    purpose; the small CTAs read what the binning kernels wrote BEFORE the pair without a griddepcontrol.wait of their
    own.  Stress: 60 calls with geometry that changes every call (so a stale bin, offset or camera record
    would show), each compared bit for bit with a handle that never uses the pair."""
    sc = synthetic.make_scene(kind="sphere", rings=90, segments=96, cameras=2, width=320, height=256, tex=16, seed=12)
    pair, plain = make(sc, "vertexColor", "shaded"), make(sc, "vertexColor", "shaded")
    for k, v in {"heavy_mode": 2, "heavy_thr": 24, "heavy_slots": 16}.items():
        pair.set_option(k, v)
    plain.set_option("heavy_mode", 0)
    ins = [T(sc[k]) for k in INPUT_KEYS]
    gen = torch.Generator(device="cpu").manual_seed(21)
    base_pos = ins[0].clone()
    n0 = pair.launch_count
    for it in range(60):
        ins[0] = base_pos * (1.0 + 0.15 * float(torch.rand((), generator=gen))) + (torch.rand(base_pos.shape, generator=gen) * 4.0).to(dev())
        a, b = pair.forward(*ins), plain.forward(*ins)
        for x, y in zip(a[:4], b[:4]):
            assert torch.equal(x.view(torch.int32) if x.dtype == torch.float32 else x, y.view(torch.int32) if y.dtype == torch.float32 else y), it
    assert pair.launch_count - n0 == 60 * 7          # the pair really ran: seven launches per forward instead of six
    pair.close(); plain.close()


@pytest.mark.skipif(torch.cuda.is_available() and torch.cuda.device_count() < 2, reason="needs two GPUs in one process")
def test_two_devices_in_one_process():
    """Handles on two devices of the same process (the shared-memory opt-ins are per device): same bits on both."""
    sc = synthetic.make_scene(kind="sphere", rings=40, segments=48, cameras=2, width=160, height=128, tex=16, seed=12)
    outs = []
    for d in (0, 1):
        dv = torch.device("cuda", d)
        ins = [torch.as_tensor(np.ascontiguousarray(sc[k]), device=dv) for k in INPUT_KEYS]
        r = _native.NativeRenderer(sc["faces"], sc["texcoords"], sc["num_vertices"], sc["num_cameras"], sc["width"], sc["height"],
                                   "vertexColor", "shaded", 1, 1, False, dv)
        with torch.cuda.device(dv):
            bary, face, render, vn, _, _ = r.forward(*ins)
            g = r.backward(torch.ones_like(render), None, ins[0], ins[1], ins[2], ins[3], ins[4], vn, bary, face, ins[5], ins[6])
            torch.cuda.synchronize(dv)
        outs.append((face.cpu(), render.cpu(), g[1].cpu()))
        r.close()
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    assert float((outs[0][2] - outs[1][2]).abs().max()) <= 1e-4 * float(outs[0][2].abs().max())
