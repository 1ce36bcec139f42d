#!/usr/bin/env python3
"""GPU-box helper: cProfile of the host side of one end-to-end step through the Python layer."""
import cProfile, os, pstats, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gvv_differentiable_cuda_renderer_b200 import CudaRendererGpu, synthetic
dev = torch.device("cuda:0")
sc = synthetic.make_scene("sphere", rings=187, segments=188, cameras=8, width=1024, height=1024, batch=1, tex=64)
N, C, W, H = sc["num_vertices"], 8, 1024, 1024
ins = {k: torch.as_tensor(sc[k], device=dev) for k in ("texture", "target_image")}
host = {k: torch.as_tensor(sc[k]).pin_memory() for k in ("vertex_pos", "vertex_color", "sh_coeff", "extrinsics", "intrinsics")}
G = torch.randn((1, C, H, W, 3), generator=torch.Generator().manual_seed(3)).to(dev).reshape(-1)
faces_l, tcs_l = sc["faces"].reshape(-1), sc["texcoords"].reshape(-1)
out_host = {k: torch.empty_like(host[k]).pin_memory() for k in ("vertex_pos", "vertex_color", "sh_coeff")}

def step():
    d = {k: v.to(dev, non_blocking=True) for k, v in host.items()}
    for k in ("vertex_pos", "vertex_color", "sh_coeff"):
        d[k].requires_grad_(True)
    layer = CudaRendererGpu(faces_attr=faces_l, texCoords_attr=tcs_l, numberOfVertices_attr=N, numberOfCameras_attr=C,
                            renderResolutionU_attr=W, renderResolutionV_attr=H, albedoMode_attr="vertexColor", shadingMode_attr="shaded",
                            vertexPos_input=d["vertex_pos"], vertexColor_input=d["vertex_color"], texture_input=ins["texture"],
                            shCoeff_input=d["sh_coeff"], targetImage_input=ins["target_image"], extrinsics_input=d["extrinsics"],
                            intrinsics_input=d["intrinsics"], device=dev)
    loss = torch.dot(layer.getRenderBufferTF().reshape(-1), G)
    loss.backward()
    for k in out_host:
        out_host[k].copy_(d[k].grad, non_blocking=True)

for _ in range(20):
    step()
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(200):
    step()
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
