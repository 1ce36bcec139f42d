#!/usr/bin/env python3
"""compute-sanitizer target: small forward+backward calls through every kernel path (all albedo modes,
16/32 tiles, big-triangle list, strip split, heavy-tile launch, bilinear texture, the TMA options (persistent backward with
its face-tile ring, bulk output tile), normal map, helpers).
  compute-sanitizer --tool memcheck  python tools/gpu_sanitize_target.py
  compute-sanitizer --tool racecheck python tools/gpu_sanitize_target.py"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gvv_differentiable_cuda_renderer_b200 import _native, synthetic
from gvv_differentiable_cuda_renderer_b200.utils import GaussianSmoothingGpu
dev = torch.device("cuda:0")
T = lambda x: torch.as_tensor(np.ascontiguousarray(x), device=dev)
KEYS = ("vertex_pos", "vertex_color", "texture", "sh_coeff", "target_image", "extrinsics", "intrinsics")
runs = 0
for kind, kw in (("sphere", dict(rings=16, segments=20, cameras=2, width=100, height=76, batch=2, tex=16)),
                 ("pyramid", dict(cameras=1, width=96, height=96, tex=8)),
                 ("sphere", dict(rings=40, segments=48, cameras=1, width=64, height=64, tex=8))):
    sc = synthetic.make_scene(kind=kind, seed=3, **kw)
    N, C, W, H = sc["num_vertices"], sc["num_cameras"], sc["width"], sc["height"]
    ins = [T(sc[k]) for k in KEYS]
    B = ins[0].shape[0]
    for albedo, shading in (("vertexColor", "shaded"), ("textured", "shaded"), ("textured", "shadeless"), ("normal", "shaded"), ("foregroundMask", "shaded")):
        for opts in ({}, {"tile": 16}, {"split_unit": 8, "heavy_thr": 16, "heavy_mode": 2}, {"texture_bilinear": 1, "cull_margin_milli": -1}, {"span_z": 1, "hiz": 0}, {"bwd_persistent": 1, "bulk_out": 1}):
            r = _native.NativeRenderer(sc["faces"], sc["texcoords"], N, C, W, H, albedo, shading, 1, 1, False, dev)
            for k, v in opts.items():
                r.set_option(k, v)
            bary, face, render, vn, _, _ = r.forward(*ins)
            if albedo in ("vertexColor", "textured", "foregroundMask"):
                g = torch.randn(render.shape, device=dev)
                r.backward(g, g, ins[0], ins[1], ins[2], ins[3], ins[4], vn, bary, face, ins[5], ins[6])
            torch.cuda.synchronize()
            r.close()
            runs += 1
    r = _native.NativeRenderer(sc["faces"], sc["texcoords"], N, C, W, H, "textured", "shaded", 1, 1, True, dev)
    r.forward(*ins)
    torch.cuda.synchronize()
    r.close()
img = torch.rand((1, 2, 33, 47, 3), device=dev, requires_grad=True)
GaussianSmoothingGpu.smoothImage(img, 2, 0.0, 1.0).sum().backward()
_native.image_gradient(img.detach(), 2)
torch.cuda.synchronize()
print(f"sanitize target done: {runs} renderer configurations")
