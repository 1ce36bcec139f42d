#!/usr/bin/env python3
"""GPU-box probe: timings of the other BASELINE.json workloads (parity cases, not bench lines).
Prints one JSON line per workload with per-kernel ms (CUDA events) and views/s."""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gvv_differentiable_cuda_renderer_b200 import _native, synthetic
dev = torch.device("cuda:0")
T = lambda x: torch.as_tensor(np.ascontiguousarray(x), device=dev)
KEYS = ("vertex_pos", "vertex_color", "texture", "sh_coeff", "target_image", "extrinsics", "intrinsics")

def run(name, sc, albedo, shading, backward=True, iters=10, opts=None):
    N, C, W, H = sc["num_vertices"], sc["num_cameras"], sc["width"], sc["height"]
    B = sc["vertex_pos"].shape[0]
    ins = [T(sc[k]) for k in KEYS]
    t0 = time.time()
    r = _native.NativeRenderer(sc["faces"], sc["texcoords"], N, C, W, H, albedo, shading, 1, 1, False, dev)
    create_s = time.time() - t0
    for k, v in (opts or {}).items():
        r.set_option(k, v)
    G = torch.randn((B, C, H, W, 3), generator=torch.Generator().manual_seed(3)).to(dev) if backward else None
    def step():
        out = r.forward(*ins)
        if backward:
            r.backward(G, None, ins[0], ins[1], ins[2], ins[3], ins[4], out[3], out[0], out[1], ins[5], ins[6])
        return out
    for _ in range(3):
        out = step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        step()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    r.set_option("time_kernels", 1)
    for _ in range(3):
        step()
    kt = {k: round(v[0] / v[1], 4) for k, v in r.kernel_times().items()}
    cov = float((out[1] >= 0).float().mean())
    print(json.dumps({"workload": name, "views": B * C, "verts": N, "tris": len(sc["faces"]), "res": [W, H], "mode": albedo + "+" + shading,
                      "pass": "fwd+bwd" if backward else "fwd", "ms_per_step": round(ms, 4), "views_per_s": round(B * C / ms * 1e3, 1),
                      "coverage": round(cov, 3), "create_s": round(create_s, 3), "options": opts or {}, "kernel_ms": kt}), flush=True)
    r.close()
    del ins, out, G
    torch.cuda.empty_cache()

# config 1: a few huge triangles, 1 camera, 1024^2 and 512^2, vertexColor, forward
run("config1 pyramid 1024^2 (test_render.py shape)", synthetic.make_scene(kind="pyramid", cameras=1, width=1024, height=1024, batch=2, distance=900.0), "vertexColor", "shaded", backward=False)
run("config1 pyramid 512^2", synthetic.make_scene(kind="pyramid", cameras=1, width=512, height=512, batch=1, distance=900.0), "vertexColor", "shadeless", backward=False)
# config 3: textured template, 1024^2 texture, 32 views at 1024^2, fwd+bwd (magdalena-sized: ~5k verts / 10k tris)
run("config3 textured 32 views (10k tris, 1024^2 texture)", synthetic.make_scene(kind="sphere", rings=72, segments=72, cameras=32, width=1024, height=1024, batch=1, tex=1024, coverage_radius_frac=0.3), "textured", "shaded")
run("config3 textured 32 views, bilinear fetch + weighted 4-texel gradient scatter (non-default variant)", synthetic.make_scene(kind="sphere", rings=72, segments=72, cameras=32, width=1024, height=1024, batch=1, tex=1024, coverage_radius_frac=0.3), "textured", "shaded", opts={"texture_bilinear": 1})
run("config3 textured shadeless 32 views", synthetic.make_scene(kind="sphere", rings=72, segments=72, cameras=32, width=1024, height=1024, batch=1, tex=1024, coverage_radius_frac=0.3), "textured", "shadeless")
# config 4 (one GPU's share): 8 batch elements x 16 cameras = 128 views of the 70k-triangle mesh
run("config4 share of one GPU: B=8 x C=16 (128 views, 70k tris)", synthetic.make_scene(kind="sphere", rings=187, segments=188, cameras=16, width=1024, height=1024, batch=8, tex=64), "vertexColor", "shaded", iters=3)
# config 5: ~1M triangles at 3840x2160, forward, three modes
sc5 = synthetic.make_scene(kind="sphere", rings=708, segments=708, cameras=1, width=3840, height=2160, batch=1, tex=1024, coverage_radius_frac=0.26)
for alb, shd in (("normal", "shaded"), ("textured", "shaded"), ("lighting", "shaded")):
    run("config5 1M tris 3840x2160", sc5, alb, shd, backward=False, iters=5)
run("config5 1M tris 3840x2160", sc5, "normal", "shaded", backward=False, iters=5, opts={"heavy_mode": 0})
run("config5 1M tris 3840x2160", sc5, "normal", "shaded", backward=False, iters=5, opts={"heavy_mode": 2})
