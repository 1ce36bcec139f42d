#!/usr/bin/env python3
"""ncu target: three fwd+bwd steps of config 2 with the default knobs (or key=value overrides).
  ncu --set full --clock-control none --import-source on -k regex:"raster_kernel|pixel_grad_kernel" \
      --launch-skip 4 --launch-count 2 -o gpurun_out/prof python tools/gpu_profile_target.py"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gvv_differentiable_cuda_renderer_b200 import _native, synthetic
dev = torch.device("cuda:0")
T = lambda x: torch.as_tensor(np.ascontiguousarray(x), device=dev)
sc = synthetic.make_scene("sphere", rings=187, segments=188, cameras=8, width=1024, height=1024, batch=1, tex=64)
N, C, W, H = sc["num_vertices"], 8, 1024, 1024
ins = [T(sc[k]) for k in ("vertex_pos", "vertex_color", "texture", "sh_coeff", "target_image", "extrinsics", "intrinsics")]
G = torch.randn((1, C, H, W, 3), generator=torch.Generator().manual_seed(3)).to(dev)
r = _native.NativeRenderer(sc["faces"], sc["texcoords"], N, C, W, H, "vertexColor", "shaded", 1, 1, False, dev)
for a in sys.argv[1:]:
    k, v = a.split("=")
    r.set_option(k, int(v))
for _ in range(3):
    out = r.forward(*ins)
    r.backward(G, None, ins[0], ins[1], ins[2], ins[3], ins[4], out[3], out[0], out[1], ins[5], ins[6])
torch.cuda.synchronize()
r.close()
