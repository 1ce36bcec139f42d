#!/usr/bin/env python3
"""GPU-box tuning probe: per-kernel times of config 2 for a few knob settings (prints JSON lines).
Own settings: `python tools/gpu_tune.py hiz=0 tile=16,span_z=0` (gvv_set_option keys)."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gvv_differentiable_cuda_renderer_b200 import _native, synthetic
dev = torch.device("cuda:0")
T = lambda x: torch.as_tensor(np.ascontiguousarray(x), device=dev)
sc = synthetic.make_scene("sphere", rings=187, segments=188, cameras=8, width=1024, height=1024, batch=1, tex=64)
N, C, W, H = sc["num_vertices"], 8, 1024, 1024
ins = [T(sc[k]) for k in ("vertex_pos", "vertex_color", "texture", "sh_coeff", "target_image", "extrinsics", "intrinsics")]
G = torch.randn((1, C, H, W, 3), generator=torch.Generator().manual_seed(3)).to(dev)
ref = None
DEFAULTS = {"tile": 32, "cull_margin_milli": 62, "ray_cache": 0, "batch_div": 8, "cta_threads": 256, "interleave": 1, "hiz": 1, "span_z": 2, "heavy_thr": 768, "heavy_mode": 1, "spread_empty": 0}
# the first configuration is the reference behaviour: every bbox pixel tested, no depth culling
configs = [dict(DEFAULTS, cull_margin_milli=-1, hiz=0, span_z=0, heavy_mode=0)]
for a in sys.argv[1:] or ["", "span_z=0", "hiz=0", "hiz=0,span_z=0"]:     # key=value[,key=value...]
    configs.append(dict(DEFAULTS, **{k: int(v) for k, v in (kv.split("=") for kv in a.split(",") if kv)}))
for cfg in configs:
    r = _native.NativeRenderer(sc["faces"], sc["texcoords"], N, C, W, H, "vertexColor", "shaded", 1, 1, False, dev)
    for k, v in cfg.items():
        r.set_option(k, v)
    for _ in range(3):
        out = r.forward(*ins)
        r.backward(G, None, ins[0], ins[1], ins[2], ins[3], ins[4], out[3], out[0], out[1], ins[5], ins[6])
    r.set_option("time_kernels", 1)
    for _ in range(10):
        out = r.forward(*ins)
        r.backward(G, None, ins[0], ins[1], ins[2], ins[3], ins[4], out[3], out[0], out[1], ins[5], ins[6])
    kt = {k: round(v[0] / v[1], 4) for k, v in r.kernel_times().items()}
    if ref is None:
        ref = (out[1].clone(), out[0].clone(), out[2].clone())
    same = bool(torch.equal(out[1], ref[0]) and torch.equal(out[0].view(torch.int32), ref[1].view(torch.int32))
                and torch.equal(out[2].view(torch.int32), ref[2].view(torch.int32)))
    print(json.dumps(dict(cfg, identical_to_unculled=same, ms=kt)), flush=True)
    r.close()
