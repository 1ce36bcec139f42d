#!/usr/bin/env python3
"""Host cost of one end-to-end step through the Python layer, on a scene so small that the GPU never back-pressures:
what is left is pure Python / ctypes / torch dispatch time per step."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gvv_differentiable_cuda_renderer_b200 import CudaRendererGpu, _native, synthetic
dev = torch.device("cuda:0")
sc = synthetic.make_scene("sphere", rings=8, segments=10, cameras=8, width=64, height=64, batch=1, tex=8)
N, C, W, H = sc["num_vertices"], 8, 64, 64
ins = {k: torch.as_tensor(sc[k], device=dev) for k in ("texture", "target_image")}
dv = {k: torch.as_tensor(sc[k], device=dev) for k in ("vertex_pos", "vertex_color", "sh_coeff", "extrinsics", "intrinsics")}
G = torch.randn((1, C, H, W, 3), device=dev).reshape(-1)
faces_l, tcs_l = sc["faces"].reshape(-1), sc["texcoords"].reshape(-1)

def layer_step():
    d = {k: v.detach() for k, v in dv.items()}
    for k in ("vertex_pos", "vertex_color", "sh_coeff"):
        d[k].requires_grad_(True)
    layer = CudaRendererGpu(faces_attr=faces_l, texCoords_attr=tcs_l, numberOfVertices_attr=N, numberOfCameras_attr=C,
                            renderResolutionU_attr=W, renderResolutionV_attr=H, albedoMode_attr="vertexColor", shadingMode_attr="shaded",
                            vertexPos_input=d["vertex_pos"], vertexColor_input=d["vertex_color"], texture_input=ins["texture"],
                            shCoeff_input=d["sh_coeff"], targetImage_input=ins["target_image"], extrinsics_input=d["extrinsics"],
                            intrinsics_input=d["intrinsics"], device=dev)
    loss = torch.dot(layer.getRenderBufferTF().reshape(-1), G)
    loss.backward()

r = _native.NativeRenderer(sc["faces"], sc["texcoords"], N, C, W, H, "vertexColor", "shaded", 1, 1, False, dev)
a = [dv["vertex_pos"], dv["vertex_color"], ins["texture"], dv["sh_coeff"], ins["target_image"], dv["extrinsics"], dv["intrinsics"]]
Gr = G.view(1, C, H, W, 3)
def native_step():
    bary, face, render, vn, _, _ = r.forward(*a)
    r.backward(Gr, None, a[0], a[1], a[2], a[3], a[4], vn, bary, face, a[5], a[6])

for name, fn in (("python layer + autograd (fwd, dot, bwd)", layer_step), ("NativeRenderer.forward + backward", native_step)):
    for _ in range(50):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 2000
    for _ in range(n):
        fn()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    print(f"{name}: {(t1 - t0) / n * 1e6:.1f} us host per step")
