#!/usr/bin/env python3
"""GPU-box probe: the headline forward (config 2) rendered N times on fresh and reused handles, every output compared bit for
bit with the first render and with the reference's -- looks for run-to-run nondeterminism (stale scratch, ordering races).
  python tools/gpu_repeat_parity.py [repeats] [key=value ...]"""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gvv_differentiable_cuda_renderer_b200 import _native, synthetic
from oracle import parity, ref as oref
dev = torch.device("cuda:0")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
opts = dict(kv.split("=") for kv in sys.argv[2:])
sc = synthetic.make_scene(kind="sphere", rings=187, segments=188, cameras=8, width=1024, height=1024, batch=1, tex=64)
ins = [torch.as_tensor(sc[k], device=dev) for k in parity.INPUT_KEYS]
N = sc["num_vertices"]
ref = oref.RefRenderer(sc["faces"], sc["texcoords"], N, 8, 1024, 1024, "vertexColor", "shaded", 1, with_backward=False)
rr = ref.forward(*ins, intermediates=True)
bits = lambda t: t.contiguous().view(torch.int32) if t.dtype == torch.float32 else t
first = None
bad = 0
r = None
for it in range(n):
    if it % 5 == 0:
        if r is not None: r.close()
        r = _native.NativeRenderer(sc["faces"], sc["texcoords"], N, 8, 1024, 1024, "vertexColor", "shaded", 1, 1, False, dev)
        for k, v in opts.items(): r.set_option(k, int(v))
    out = [t.clone() for t in r.forward(*ins)[:4]]
    torch.cuda.synchronize()
    if first is None:
        first = out
    d_first = [int((bits(a) != bits(b)).sum()) for a, b in zip(out, first)]
    same = out[1] == rr["face"]
    d_ref = {"face": int((~same).sum()), "bary": int((bits(out[0]) != bits(rr["bary"]))[same].sum()), "render": int((bits(out[2]) != bits(rr["render"]))[same].sum())}
    if any(d_first) or d_ref["bary"] or d_ref["render"]:
        bad += 1
        idx = (bits(out[2]) != bits(rr["render"])).nonzero()[:6].tolist()
        print(json.dumps({"it": it, "vs_first": d_first, "vs_ref": d_ref, "where": idx}), flush=True)
print(json.dumps({"repeats": n, "bad": bad, "opts": opts}))
