#!/usr/bin/env python3
"""Per-CTA timeline of raster_kernel on config 2 (debug option cta_trace): how long the heaviest
tiles run compared with the whole launch, and how many CTAs are resident over time."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gvv_differentiable_cuda_renderer_b200 import _native, synthetic
dev = torch.device("cuda:0")
T = lambda x: torch.as_tensor(np.ascontiguousarray(x), device=dev)
sc = synthetic.make_scene("sphere", rings=187, segments=188, cameras=8, width=1024, height=1024, batch=1, tex=64)
N, C, W, H = sc["num_vertices"], 8, 1024, 1024
ins = [T(sc[k]) for k in ("vertex_pos", "vertex_color", "texture", "sh_coeff", "target_image", "extrinsics", "intrinsics")]
r = _native.NativeRenderer(sc["faces"], sc["texcoords"], N, C, W, H, "vertexColor", "shaded", 1, 1, False, dev)
for a in sys.argv[1:]:
    k, v = a.split("=")
    r.set_option(k, int(v))
r.set_option("cta_trace", 1)
for _ in range(4):
    r.forward(*ins)
torch.cuda.synchronize()
nT = ((W + 31) // 32) * ((H + 31) // 32) if not any(a.startswith("tile=16") for a in sys.argv[1:]) else ((W + 15) // 16) * ((H + 15) // 16)
nI = nT + nT // 2
tr = r.debug_copy(2, C * nI * 4 * 8).view(np.uint64).reshape(-1, 4)[: C * nI]
tr = tr[tr[:, 0] > 0]
t0, t1, cnt, sm = tr[:, 0].astype(np.int64), tr[:, 1].astype(np.int64), tr[:, 2].astype(np.int64), tr[:, 3].astype(np.int64)
base = t0.min()
span = (t1.max() - base) / 1e3
dur = (t1 - t0) / 1e3
order = np.argsort(-dur)
print(json.dumps({"kernel_span_us": round(span, 1), "ctas": int(len(dur)), "nonempty": int((cnt > 0).sum()),
                  "sum_cta_us": round(float(dur.sum()), 1), "mean_nonempty_us": round(float(dur[cnt > 0].mean()), 2),
                  "longest": [{"us": round(float(dur[i]), 1), "bin": int(cnt[i]), "start_us": round(float((t0[i] - base) / 1e3), 1), "blockIdx": int(i)} for i in order[:12]]}))
# resident CTAs over time (20 slices)
edges = np.linspace(0, span, 21)
res = [int(((t0 - base) / 1e3 < e1) .astype(int).dot(((t1 - base) / 1e3 > e0).astype(int))) for e0, e1 in zip(edges[:-1], edges[1:])]
print(json.dumps({"resident_ctas_per_slice": res}))
# work vs bin size
for lo, hi in ((1, 64), (64, 256), (256, 1024), (1024, 1 << 30)):
    m = (cnt >= lo) & (cnt < hi)
    if m.any():
        print(json.dumps({"bin_range": [lo, hi], "ctas": int(m.sum()), "mean_us": round(float(dur[m].mean()), 2), "us_per_tri": round(float(dur[m].sum() / cnt[m].sum()), 4)}))
r.close()
