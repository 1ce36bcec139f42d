#!/usr/bin/env python3
"""GPU-box tuning probe: per-kernel times of config 2 for a few knob settings (prints JSON lines).
Own settings: `python tools/gpu_tune.py 32,62,1 32,62,0` (tile,margin_milli,ray_cache)."""
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
configs = [(32, -1, 1), (32, 250, 1), (32, 62, 1), (32, 16, 1), (16, 62, 1), (16, 250, 1)]
if len(sys.argv) > 1:
    configs = [(32, -1, 1)] + [tuple(map(int, a.split(","))) for a in sys.argv[1:]]   # tile,margin_milli,ray_cache[,batch_div]
for cfg in configs:
    tile, margin, rc = cfg[:3]
    bd = cfg[3] if len(cfg) > 3 else 8
    nth = cfg[4] if len(cfg) > 4 else 256
    guided = cfg[5] if len(cfg) > 5 else 1
    hiz = cfg[6] if len(cfg) > 6 else 1
    r = _native.NativeRenderer(sc["faces"], sc["texcoords"], N, C, W, H, "vertexColor", "shaded", 1, 1, False, dev)
    r.set_option("tile", tile)
    r.set_option("cull_margin_milli", margin)
    r.set_option("ray_cache", rc)
    r.set_option("batch_div", bd)
    r.set_option("cta_threads", nth)
    r.set_option("interleave", guided)
    r.set_option("hiz", hiz)
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
    print(json.dumps({"tile": tile, "margin_milli": margin, "ray_cache": rc, "batch_div": bd, "cta_threads": nth, "interleave": guided, "hiz": hiz, "identical_to_unculled": same, "ms": kt}), flush=True)
    r.close()
