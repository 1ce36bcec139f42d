#!/usr/bin/env python3
"""Randomised parity sweep on a B200: the CUDA path (through the C ABI) against Oracle 1 = the reference's own CUDA
core compiled in place (oracle/_ref/libgvv_ref.so), on seeded random scenes: mesh kind and resolution, image size
(including sizes that are no multiple of the tile), cameras, batch, distance, coverage, radial noise, albedo /
shading mode, target gradient on / off, image filter size.  TEST INFRASTRUCTURE: prints one JSON line per scene and
a summary; exits non-zero on the first violation of the parity protocol:
  face buffer  : equal everywhere except exact depth ties (both candidates re-evaluated with gvv_debug_eval: equal
                 keys, equal to the reference's depth buffer, ours = the smaller triangle id)
  barycentrics : bit-identical where the faces agree;  vertex normals: bit-identical
  render       : |diff| <= 1e-6 where the faces agree
  gradients    : rel-L2 <= 1e-4 and max-abs <= 1e-3 max|g| per tensor (atomic accumulation order differs); where the
                 reference itself is further than that from the fp64-accumulated CPU oracle (ill-conditioned position
                 gradients), at most a quarter of the reference's own distance to that oracle"""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from gvv_differentiable_cuda_renderer_b200 import _native, synthetic
from oracle import ref as oref
import test_gpu_parity as tp

n_scenes = int(sys.argv[1]) if len(sys.argv) > 1 else 24
rng = np.random.default_rng(2026)
MODES = [("vertexColor", "shaded"), ("vertexColor", "shadeless"), ("textured", "shaded"), ("textured", "shadeless"),
         ("normal", "shaded"), ("lighting", "shaded"), ("foregroundMask", "shaded")]
tot = dict(scenes=0, pixels=0, covered=0, exact_ties=0, max_render_diff=0.0, max_grad_rel_l2=0.0)
for it in range(n_scenes):
    kind = ["sphere", "sphere", "sphere", "pyramid"][int(rng.integers(4))]
    albedo, shading = MODES[int(rng.integers(len(MODES)))]
    kw = dict(cameras=int(rng.integers(1, 4)), width=int(rng.integers(70, 640)), height=int(rng.integers(70, 520)),
              batch=int(rng.integers(1, 3)), tex=int(rng.choice([8, 32, 96])), seed=int(rng.integers(1 << 20)),
              distance=float(rng.uniform(600, 4000)), coverage_radius_frac=float(rng.uniform(0.15, 0.7)))
    if kind == "sphere":
        kw.update(rings=int(rng.integers(6, 90)), segments=int(rng.integers(6, 110)), noise=float(rng.choice([0.0, 0.005, 0.02, 0.05])))
    fs = int(rng.integers(1, 3))
    sc = synthetic.make_scene(kind=kind, **kw)
    N, C, W, H = sc["num_vertices"], sc["num_cameras"], sc["width"], sc["height"]
    B = sc["vertex_pos"].shape[0]
    ins = [tp.T(sc[k]) for k in tp.INPUT_KEYS]
    tgt = tp.T(rng.random((B, C, H, W, 3), dtype=np.float32))
    ins[4] = tgt
    ref = oref.RefRenderer(sc["faces"], sc["texcoords"], N, C, W, H, albedo, shading, image_filter=fs)
    rr = ref.forward(*ins, intermediates=True)
    r = _native.NativeRenderer(sc["faces"], sc["texcoords"], N, C, W, H, albedo, shading, fs, 1, False, tp.dev())
    bary, face, render, vn, _, _ = r.forward(*ins)
    ties = tp.assert_faces_equal_up_to_exact_ties(r, face, rr["face"], rr["depth"])
    same = face == rr["face"]
    assert torch.equal(bary.view(torch.int32)[same], rr["bary"].view(torch.int32)[same]), "barycentrics differ"
    assert torch.equal(vn.view(torch.int32), rr["vertex_normal"].view(torch.int32)), "vertex normals differ"
    rd = float((render - rr["render"]).abs()[same].max()) if bool(same.any()) else 0.0
    assert rd <= 1e-6, ("render", rd)
    gl2 = 0.0
    if albedo in ("vertexColor", "textured", "foregroundMask"):
        rg = tp.T(rng.standard_normal((B, C, H, W, 3)).astype(np.float32))
        tg = tp.T(rng.standard_normal((B, C, H, W, 3)).astype(np.float32)) if rng.random() < 0.5 else None
        gm = r.backward(rg, tg, ins[0], ins[1], ins[2], ins[3], tgt, rr["vertex_normal"], rr["bary"], rr["face"], ins[5], ins[6])
        gr = ref.backward(rg, ins[0], ins[1], ins[2], ins[3], tgt, rr["vertex_normal"], rr["bary"], rr["face"], tg, ins[5], ins[6])
        try:
            tp.grads_close(gm, gr)
        except AssertionError as e:
            # diagnose: which of the two is closer to the fp64-accumulating CPU oracle on the same forward buffers?
            from oracle import cpu
            go = cpu.backward(sc["faces"], sc["texcoords"], N, C, W, H, albedo, shading, fs, rg.cpu().numpy(), None if tg is None else tg.cpu().numpy(),
                              sc["vertex_pos"], sc["vertex_color"], sc["texture"], sc["sh_coeff"], tgt.cpu().numpy(), rr["vertex_normal"].cpu().numpy(),
                              rr["bary"].cpu().numpy(), rr["face"].cpu().numpy(), sc["extrinsics"], sc["intrinsics"])
            names = ("vertex_pos_grad", "vertex_color_grad", "texture_grad", "sh_coeff_grad")
            diag = {n: dict(ours_vs_ref=tp.rel_l2(a.cpu().numpy(), b.cpu().numpy()), ours_vs_fp64=tp.rel_l2(a.cpu().numpy(), o), ref_vs_fp64=tp.rel_l2(b.cpu().numpy(), o),
                            max_abs=float(np.abs(o).max())) for n, a, b, o in zip(names, gm, gr, go) if np.abs(o).max() > 0}
            # Ill-conditioned position gradients (far camera, small triangles: fp32 cancellation in the ray/plane hit)
            # make the REFERENCE itself deviate from the fp64-accumulated oracle by far more than 1e-4; there the
            # protocol accepts a distance to the reference of at most a quarter of the reference's own noise.
            ok = all(v["ours_vs_ref"] <= max(1e-4, 0.25 * v["ref_vs_fp64"]) for v in diag.values())
            print(json.dumps(dict(scene=it, kind=kind, mode=albedo + "+" + shading, kw=kw, image_filter=fs, target_grad=tg is not None,
                                  strict_tolerance_exceeded=str(e), diag=diag, accepted_as_within_reference_noise=ok)), flush=True)
            tot["within_reference_noise_only"] = tot.get("within_reference_noise_only", 0) + 1
            if not ok:
                raise
        gl2 = max([tp.rel_l2(a.cpu().numpy(), b.cpu().numpy()) for a, b in zip(gm, gr) if float(b.abs().max()) > 0] or [0.0])
    cov = int((face >= 0).sum())
    print(json.dumps(dict(scene=it, kind=kind, mode=albedo + "+" + shading, verts=N, tris=len(sc["faces"]), views=B * C, res=[W, H],
                          image_filter=fs, covered_px=cov, exact_tie_px=int(ties), max_render_diff=rd, max_grad_rel_l2=float(gl2))), flush=True)
    tot["scenes"] += 1; tot["pixels"] += face.numel(); tot["covered"] += cov; tot["exact_ties"] += int(ties)
    tot["max_render_diff"] = max(tot["max_render_diff"], rd); tot["max_grad_rel_l2"] = max(tot["max_grad_rel_l2"], float(gl2))
    r.close()
print(json.dumps(dict(summary=tot, verdict="all scenes within the parity protocol")), flush=True)
