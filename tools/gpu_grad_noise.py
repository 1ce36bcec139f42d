#!/usr/bin/env python3
"""GPU-box diagnostic: how noisy is the reference's own position gradient?  Runs the reference backward (Oracle 1)
several times on identical inputs, ours twice, and the fp64 evaluation of the reference's formulas (Oracle 2,
GVVO_FP64), and prints the pairwise rel-L2 distances per gradient tensor.  Output -> profiles/ via the caller."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gvv_differentiable_cuda_renderer_b200 import _native, synthetic
from oracle import ref as oref, parity, cpu

dev = torch.device("cuda:0")
cfg = dict(kind="sphere", rings=187, segments=188, cameras=8, width=1024, height=1024, batch=1, tex=64)
if len(sys.argv) > 1:
    cfg.update(json.loads(sys.argv[1]))
sc = synthetic.make_scene(**cfg)
N, C, W, H = sc["num_vertices"], sc["num_cameras"], sc["width"], sc["height"]
ins = [torch.as_tensor(sc[k], device=dev) for k in parity.INPUT_KEYS]
ref = oref.RefRenderer(sc["faces"], sc["texcoords"], N, C, W, H, "vertexColor", "shaded")
rr = ref.forward(*ins, intermediates=True)
G = torch.randn((1, C, H, W, 3), generator=torch.Generator().manual_seed(3)).to(dev)
mine = _native.NativeRenderer(sc["faces"], sc["texcoords"], N, C, W, H, "vertexColor", "shaded", 1, 1, False, dev)
a = (ins[0], ins[1], ins[2], ins[3], ins[4], rr["vertex_normal"], rr["bary"], rr["face"])
refs = [ref.backward(G, *a, None, ins[5], ins[6]) for _ in range(3)]
ours = [mine.backward(G, None, *a, ins[5], ins[6]) for _ in range(2)]
truth = [torch.as_tensor(t) for t in parity.fp64_backward(sc, "vertexColor", "shaded", 1, G, None, rr)()]
o32 = [torch.as_tensor(t) for t in cpu.backward(sc["faces"], sc["texcoords"], N, C, W, H, "vertexColor", "shaded", 1, G.cpu().numpy(), None,
                                                sc["vertex_pos"], sc["vertex_color"], sc["texture"], sc["sh_coeff"], sc["target_image"],
                                                rr["vertex_normal"].cpu().numpy(), rr["bary"].cpu().numpy(), rr["face"].cpu().numpy(), sc["extrinsics"], sc["intrinsics"])]
out = {"config": cfg}
for i, name in enumerate(parity.GRAD_NAMES):
    if i == 2:
        continue
    t = truth[i]
    d = {"ref_vs_ref": [parity.rel_l2(refs[0][i].cpu(), refs[k][i].cpu()) for k in (1, 2)],
         "ours_vs_ours": parity.rel_l2(ours[0][i].cpu(), ours[1][i].cpu()),
         "ours_vs_ref": [parity.rel_l2(ours[0][i].cpu(), refs[k][i].cpu()) for k in range(3)],
         "ref_vs_fp64": [parity.rel_l2(refs[k][i].cpu(), t) for k in range(3)],
         "ours_vs_fp64": [parity.rel_l2(ours[k][i].cpu(), t) for k in range(2)],
         "cpu_fp32_literal_vs_fp64": parity.rel_l2(o32[i], t),
         "norm": float(t.norm()), "maxabs": float(t.abs().max())}
    out[name] = d
# where does the position-gradient difference live?
dpos = (ours[0][0].cpu().double() - refs[0][0].cpu().double()).reshape(-1, 3).norm(dim=1)
tpos = truth[0].reshape(-1, 3).norm(dim=1)
top = torch.topk(dpos, 8)
out["top_vertex_diffs"] = [{"vertex": int(v), "diff": float(x), "truth_norm": float(tpos[v]),
                            "ref_err": float((refs[0][0].cpu().double().reshape(-1, 3)[v] - truth[0].reshape(-1, 3)[v]).norm()),
                            "ours_err": float((ours[0][0].cpu().double().reshape(-1, 3)[v] - truth[0].reshape(-1, 3)[v]).norm())} for x, v in zip(top.values, top.indices)]
out["diff_energy_in_top8"] = float((top.values ** 2).sum() / (dpos ** 2).sum())
print(json.dumps(out, indent=1))
