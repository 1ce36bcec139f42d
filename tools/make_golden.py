#!/usr/bin/env python3
"""Generates tests/golden/*.npz by running the REFERENCE's own CUDA core (Oracle 1,
oracle/_ref/libgvv_ref.so = unmodified reference sources + oracle/ref_harness.cu) on a B200.

Run on the GPU box:   gpurun -- python tools/make_golden.py     (writes gpurun_out/golden/)
then copy gpurun_out/golden/*.npz into tests/golden/.  Inputs are stored next to the outputs so the
fixtures do not depend on numpy's random streams."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gvv_differentiable_cuda_renderer_b200 import synthetic
from oracle import ref as oref

dev = torch.device("cuda:0")
T = lambda x: torch.as_tensor(np.ascontiguousarray(x), device=dev)

CASES = [
    # name, scene kwargs, albedo, shading, with target gradient
    ("triangle48_vertexColor_shaded", dict(kind="triangle", cameras=1, width=48, height=48, tex=16), "vertexColor", "shaded", False),
    ("pyramid64_vertexColor_shaded", dict(kind="pyramid", cameras=2, width=64, height=64, tex=16), "vertexColor", "shaded", True),
    ("sphere64_vertexColor_shaded", dict(kind="sphere", rings=10, segments=14, cameras=2, width=64, height=64, tex=16), "vertexColor", "shaded", False),
    ("sphere64_vertexColor_shadeless", dict(kind="sphere", rings=10, segments=14, cameras=1, width=64, height=64, tex=16), "vertexColor", "shadeless", False),
    ("sphere64_textured_shaded", dict(kind="sphere", rings=10, segments=14, cameras=2, width=64, height=64, tex=32), "textured", "shaded", False),
    ("sphere64_textured_shadeless", dict(kind="sphere", rings=10, segments=14, cameras=1, width=64, height=64, tex=32), "textured", "shadeless", False),
    ("sphere64_normal", dict(kind="sphere", rings=10, segments=14, cameras=1, width=64, height=64, tex=16), "normal", "shaded", False),
    ("sphere64_lighting", dict(kind="sphere", rings=10, segments=14, cameras=1, width=64, height=64, tex=16), "lighting", "shadeless", False),
    ("sphere64_foregroundMask", dict(kind="sphere", rings=10, segments=14, cameras=1, width=64, height=64, tex=16), "foregroundMask", "shaded", True),
    ("sphere56x40_vertexColor_shaded_B2", dict(kind="sphere", rings=8, segments=12, cameras=2, width=56, height=40, batch=2, tex=16), "vertexColor", "shaded", True),
]

def main():
    out = os.path.join("gpurun_out", "golden")
    os.makedirs(out, exist_ok=True)
    for name, kw, albedo, shading, tgrad in CASES:
        sc = synthetic.make_scene(**kw)
        N, C, W, H = sc["num_vertices"], sc["num_cameras"], sc["width"], sc["height"]
        B = sc["vertex_pos"].shape[0]
        rng = np.random.default_rng(7)
        if tgrad:
            sc["target_image"] = rng.random((B, C, H, W, 3), dtype=np.float32)
        ins = [T(sc[k]) for k in ("vertex_pos", "vertex_color", "texture", "sh_coeff", "target_image", "extrinsics", "intrinsics")]
        ref = oref.RefRenderer(sc["faces"], sc["texcoords"], N, C, W, H, albedo, shading)
        r = ref.forward(*ins, intermediates=True)
        data = dict(albedo=albedo, shading=shading, **{k: np.asarray(v) for k, v in sc.items()})
        for k in ("bary", "face", "render", "vertex_normal", "depth"):
            data["ref_" + k] = r[k].cpu().numpy()
        if albedo in ("vertexColor", "textured", "foregroundMask"):
            rg = rng.standard_normal((B, C, H, W, 3)).astype(np.float32)
            tg = rng.standard_normal((B, C, H, W, 3)).astype(np.float32) if tgrad else None
            g = ref.backward(T(rg), ins[0], ins[1], ins[2], ins[3], ins[4], r["vertex_normal"], r["bary"], r["face"],
                             T(tg) if tgrad else None, ins[5], ins[6])
            data["render_grad"] = rg
            if tgrad:
                data["target_grad"] = tg
            for k, v in zip(("vertex_pos_grad", "vertex_color_grad", "texture_grad", "sh_coeff_grad"), g):
                data["ref_" + k] = v.cpu().numpy()
        np.savez_compressed(os.path.join(out, name + ".npz"), **data)
        print(name, "covered", int((r["face"] >= 0).sum()), "of", r["face"].numel(), flush=True)

if __name__ == "__main__":
    main()
