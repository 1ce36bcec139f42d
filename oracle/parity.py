"""TEST INFRASTRUCTURE ONLY -- the parity protocol of SURVEY.md 8(c), CUDA path vs Oracle 1 on the same GPU.

Oracle 1 (oracle/_ref/libgvv_ref.so) is the reference's own renderer core compiled unmodified for sm_100a, so
the comparison is bit-level wherever the reference is deterministic:

  camera inverses, projected vertices, vertex normals      bit-equal
  face buffer                                              equal except EXACT depth ties -- the reference's own
                                                           choice there is a data race (CUDABasedRasterization.cu:
                                                           295-301), ours is the smallest triangle id; every differing
                                                           pixel is proven to be a tie by re-evaluating both
                                                           candidates (gvv_debug_eval) against the reference's depth buffer
  barycentrics, render buffer (where the face is equal)    bit-equal
  gradients (atomic accumulation order differs)            rel-L2 <= 1e-4 and max-abs <= 1e-3 * max|g| per tensor.
                                                           Where a tensor misses 1e-4 -- the position gradient on meshes of
                                                           millimetre triangles seen from metres away: the reference's
                                                           dJBCDVerpos chain (RendererUtil.h:670-861) cancels terms ~1e5 x
                                                           their difference, so its OWN fp32 result is 1e-4..1e-3 away from
                                                           its formula's exact value -- both are compared with the fp64
                                                           evaluation of the reference's formulas on the same inputs (Oracle 2
                                                           built with -DGVVO_FP64): |ours - ref| must stay below HALF of the
                                                           reference's own distance |ref - exact| to that value, ours must not
                                                           be farther from it than 1.25 x the reference is, and |ours - ref|
                                                           <= 1e-3 in any case.  Measured at the headline configuration
                                                           (profiles/r02_grad_noise.json): reference run-to-run 1.3e-6, ours
                                                           1.8e-7, ours-reference 1.76e-4, reference-exact 9.0e-4, ours-exact 8.96e-4

Used by tests/ (-m gpu) and by the untimed `parity` leg of bench.py; never by the product.
"""
import numpy as np
import torch

from . import ref as oref

INPUT_KEYS = ("vertex_pos", "vertex_color", "texture", "sh_coeff", "target_image", "extrinsics", "intrinsics")
GRAD_NAMES = ("vertex_pos_grad", "vertex_color_grad", "texture_grad", "sh_coeff_grad")


def bits(t):
    return t.contiguous().view(torch.int32)


def rel_l2(a, b):
    a, b = a.double(), b.double()
    d = float(b.norm())
    return float((a - b).norm()) / d if d > 0 else float(a.norm())


def prove_exact_ties(renderer, face, ref_face, ref_depth):
    """Number of pixels whose face id differs; raises unless each is an exact depth tie resolved to the smaller id."""
    mism = (face != ref_face).nonzero()
    if mism.numel() == 0:
        return 0
    mism = mism.cpu().numpy()
    C = face.shape[1]
    f, rf = face.cpu().numpy(), ref_face.cpu().numpy()
    mine, theirs = f[tuple(mism.T)], rf[tuple(mism.T)]
    if not ((mine >= 0).all() and (theirs >= 0).all()):
        raise AssertionError("coverage differs from the reference")
    view = mism[:, 0] * C + mism[:, 1]
    k1, _ = renderer.eval_pairs(np.stack([view, mism[:, 3], mism[:, 2], mine], 1))
    k2, _ = renderer.eval_pairs(np.stack([view, mism[:, 3], mism[:, 2], theirs], 1))
    if not np.array_equal(k1, k2):
        raise AssertionError("face mismatch that is not an exact depth tie")
    if not (mine < theirs).all():
        raise AssertionError("tie not resolved to the smallest triangle id")
    if ref_depth is not None and not np.array_equal(k1, ref_depth.cpu().numpy()[tuple(mism.T)]):
        raise AssertionError("tie key differs from the reference's depth buffer")
    return int(len(mism))


def compare_forward(renderer, ref_out, out, N):
    """renderer: NativeRenderer after forward() -> out; ref_out: RefRenderer.forward(..., intermediates=True).
    Returns a dict of counts; raises AssertionError on any violation of the protocol."""
    bary, face, render, vn = out[:4]
    B, C = face.shape[0], face.shape[1]
    V = B * C
    dev = face.device
    res = {"pixels": int(face.numel()), "covered": int((ref_out["face"] >= 0).sum())}
    cams = torch.from_numpy(renderer.debug_copy(0, V * 256).view(np.float32).reshape(-1, 64)[:V].copy()).to(dev)
    mycam = torch.cat([cams[:, 21:37], cams[:, 37:53]], 1).contiguous()          # Einv | Pinv
    res["cam_bit_mismatch"] = int((bits(mycam) != bits(ref_out["cam"].reshape(V, 32))).sum())
    proj = torch.from_numpy(renderer.debug_copy(1, V * N * 16).view(np.float32).reshape(-1, N, 4)[:V].copy()).to(dev)
    res["proj_bit_mismatch"] = int((bits(proj[..., :3]) != bits(ref_out["proj"].reshape(V, N, 3))).sum())
    res["exact_tie_pixels"] = prove_exact_ties(renderer, face, ref_out["face"], ref_out["depth"])
    same = face == ref_out["face"]
    res["bary_bit_mismatch"] = int((bits(bary) != bits(ref_out["bary"]))[same].sum())
    res["render_bit_mismatch"] = int((bits(render) != bits(ref_out["render"]))[same].sum())
    res["render_maxabs"] = float((render - ref_out["render"]).abs()[same].max())
    res["vertex_normal_bit_mismatch"] = int((bits(vn) != bits(ref_out["vertex_normal"])).sum())
    bad = {k: v for k, v in res.items() if k.endswith("bit_mismatch") and v}
    if bad:
        raise AssertionError(f"not bit-equal to the reference: {bad}")
    if res["exact_tie_pixels"] > 1e-4 * res["pixels"]:
        raise AssertionError(f"{res['exact_tie_pixels']} tie pixels")
    return res


def compare_backward(grads, ref_grads, rel=1e-4, mx=1e-3, exact=None, hard_cap=1e-3):
    """exact: optional callable returning the four gradients of the fp64 evaluation (numpy float64), called only
    when a tensor misses `rel` (see the module docstring)."""
    res = {}
    truth = None
    for i, (name, a, b) in enumerate(zip(GRAD_NAMES, grads, ref_grads)):
        e = rel_l2(a, b)
        res[name + "_rel_l2"] = e
        if not e <= rel:
            if exact is None or not e <= hard_cap:
                raise AssertionError(f"{name}: rel-L2 {e:.3e} > {rel}")
            if truth is None:
                truth = [torch.as_tensor(t) for t in exact()]
            t = truth[i].reshape(b.shape)
            e_ours, e_ref = rel_l2(a.cpu(), t), rel_l2(b.cpu(), t)
            res[name + "_rel_l2_ours_vs_fp64"], res[name + "_rel_l2_ref_vs_fp64"] = e_ours, e_ref
            if not (e <= 0.5 * e_ref and e_ours <= 1.25 * e_ref):
                raise AssertionError(f"{name}: rel-L2 {e:.3e} to the reference exceeds half of the reference's own fp32 error "
                                     f"({e_ref:.3e} to the fp64 evaluation of its formulas; ours {e_ours:.3e})")
        m = float(b.abs().max())
        if m > 0 and not float((a - b).abs().max()) <= mx * m:
            raise AssertionError(f"{name}: max-abs {float((a - b).abs().max()):.3e} > {mx} * {m:.3e}")
    return res


def fp64_backward(sc, albedo, shading, image_filter, render_grad, target_grad, rr):
    """Deferred fp64 evaluation of the reference's backward formulas (Oracle 2, GVVO_FP64 build) on the reference's
    forward buffers rr -- the same inputs both GPU backwards were given."""
    def run():
        from . import cpu
        n = lambda t: None if t is None else t.detach().cpu().numpy()
        return cpu.backward(sc["faces"], sc["texcoords"], sc["num_vertices"], sc["num_cameras"], sc["width"], sc["height"], albedo, shading,
                            image_filter, n(render_grad), n(target_grad), sc["vertex_pos"], sc["vertex_color"], sc["texture"], sc["sh_coeff"],
                            sc["target_image"], n(rr["vertex_normal"]), n(rr["bary"]), n(rr["face"]), sc["extrinsics"], sc["intrinsics"], fp64=True)
    return run


def check_scene(sc, albedo, shading, renderer=None, backward=True, render_grad=None, target_grad=None, image_filter=1,
                options=None, dev=None):
    """Full protocol on one scene dict (synthetic.make_scene layout).  Returns (result dict, our outputs, reference outputs)."""
    from gvv_differentiable_cuda_renderer_b200 import _native
    dev = dev or torch.device("cuda:0")
    N, C, W, H = sc["num_vertices"], sc["num_cameras"], sc["width"], sc["height"]
    ins = [torch.as_tensor(np.ascontiguousarray(sc[k]), device=dev) for k in INPUT_KEYS]
    B = ins[0].shape[0]
    own = renderer is None
    if own:
        renderer = _native.NativeRenderer(sc["faces"], sc["texcoords"], N, C, W, H, albedo, shading, image_filter, 1, False, dev)
        for k, v in (options or {}).items():
            renderer.set_option(k, v)
    do_bwd = backward and albedo in ("vertexColor", "textured", "foregroundMask")
    ref = oref.RefRenderer(sc["faces"], sc["texcoords"], N, C, W, H, albedo, shading, image_filter, with_backward=do_bwd)
    rr = ref.forward(*ins, intermediates=True)
    out = renderer.forward(*ins)
    torch.cuda.synchronize()
    res = compare_forward(renderer, rr, out, N)
    grads = ref_grads = None
    if do_bwd:
        if render_grad is None:
            render_grad = torch.randn((B, C, H, W, 3), generator=torch.Generator().manual_seed(3)).to(dev)
        # both backwards are fed the REFERENCE's forward buffers, so this isolates the backward
        grads = renderer.backward(render_grad, target_grad, ins[0], ins[1], ins[2], ins[3], ins[4], rr["vertex_normal"], rr["bary"],
                                  rr["face"], ins[5], ins[6])
        ref_grads = ref.backward(render_grad, ins[0], ins[1], ins[2], ins[3], ins[4], rr["vertex_normal"], rr["bary"], rr["face"],
                                 target_grad, ins[5], ins[6])
        torch.cuda.synchronize()
        res.update(compare_backward(grads, ref_grads, exact=fp64_backward(sc, albedo, shading, image_filter, render_grad, target_grad, rr)))
    if own:
        renderer.close()
    return res, out, rr, grads, ref_grads
