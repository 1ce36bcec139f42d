#!/usr/bin/env python3
"""GPU-box A/B probe on config 2: for every option set (key=value[,key=value...]; "" = defaults) the step time (CUDA
events over 60 steps, kernels chained as in production), the per-kernel times (time_kernels) and whether the forward
is bit-identical / the gradients agree with the FIRST set.  JSON lines.
  python tools/gpu_ab.py "" bwd_prefetch=0 ..."""
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
for a in sys.argv[1:] or [""]:
    cfg = {k: int(v) for k, v in (kv.split("=") for kv in a.split(",") if kv)}
    r = _native.NativeRenderer(sc["faces"], sc["texcoords"], N, C, W, H, "vertexColor", "shaded", 1, 1, False, dev)
    for k, v in cfg.items():
        r.set_option(k, v)
    def step():
        out = r.forward(*ins)
        g = r.backward(G, None, ins[0], ins[1], ins[2], ins[3], ins[4], out[3], out[0], out[1], ins[5], ins[6])
        return out, g
    for _ in range(5):
        out, g = step()
    best = 1e9
    for rep in range(3):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(60):
            step()
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 60)
    r.set_option("time_kernels", 1)
    for _ in range(20):
        out, g = step()
    kt = {k.replace("_kernel", ""): round(v[0] / v[1], 4) for k, v in r.kernel_times().items()}
    r.set_option("time_kernels", 0)
    if ref is None:
        ref = ([t.clone() for t in out[:4]], [t.clone() for t in g])
    same = all(torch.equal(x.view(torch.int32) if x.dtype == torch.float32 else x, y.view(torch.int32) if y.dtype == torch.float32 else y) for x, y in zip(out[:4], ref[0]))
    gerr = max(float((x.double() - y.double()).norm() / max(float(y.double().norm()), 1e-30)) for x, y in zip(g, ref[1]))
    print(json.dumps({"opts": a, "step_ms": round(best, 4), "fwd_identical": same, "grad_rel_vs_first": gerr, "ms": kt}), flush=True)
    r.close()
