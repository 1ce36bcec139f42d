#!/usr/bin/env python3
"""GPU-box probe: raster throughput on views that see NOTHING (every tile takes the background path: 24 B/px of
constant stores) and on the config-2 scene, per option set -- how far the empty-tile path is from the HBM store rate.
  python tools/gpu_empty_probe.py [key=value ...]"""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gvv_differentiable_cuda_renderer_b200 import _native, synthetic
dev = torch.device("cuda:0")
T = lambda x: torch.as_tensor(np.ascontiguousarray(x), device=dev)
for C in (8, 32):
    sc = synthetic.make_scene("sphere", rings=187, segments=188, cameras=C, width=1024, height=1024, batch=1, tex=64)
    N, W, H = sc["num_vertices"], 1024, 1024
    for empty in (True, False):
        s2 = dict(sc)
        if empty:
            s2["vertex_pos"] = sc["vertex_pos"] + np.float32(1.0e6)      # far outside every frustum
        ins = [T(s2[k]) for k in ("vertex_pos", "vertex_color", "texture", "sh_coeff", "target_image", "extrinsics", "intrinsics")]
        r = _native.NativeRenderer(sc["faces"], sc["texcoords"], N, C, W, H, "vertexColor", "shaded", 1, 1, False, dev)
        for a in sys.argv[1:]:
            k, v = a.split("="); r.set_option(k, int(v))
        for _ in range(3):
            out = r.forward(*ins)
        r.set_option("time_kernels", 1)
        for _ in range(20):
            out = r.forward(*ins)
        kt = {k.replace("_kernel", ""): round(v[0] / v[1], 4) for k, v in r.kernel_times().items()}
        cov = float((out[1] >= 0).float().mean())
        ms = kt["raster"]
        print(json.dumps({"views": C, "empty": empty, "coverage": round(cov, 4), "raster_ms": ms, "store_GBps": round(24 * W * H * C / (ms * 1e-3) / 1e9, 1), "ms": kt, "opts": sys.argv[1:]}), flush=True)
        r.close()
