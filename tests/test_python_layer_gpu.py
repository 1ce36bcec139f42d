"""GPU (-m gpu): the drop-in Python layer (CudaRendererGpu) -- headless restatements of the
reference's four scripts (python/test_render.py, python/test_gradients_*.py): same call pattern,
assertions instead of cv.imshow."""
import numpy as np
import pytest
import torch

from gvv_differentiable_cuda_renderer_b200 import CudaRendererGpu, synthetic
from gvv_differentiable_cuda_renderer_b200.CudaRenderer import _HANDLE_CACHE

pytestmark = pytest.mark.gpu


def scene(**kw):
    sc = synthetic.make_scene(**kw)
    dev = torch.device("cuda:0")
    t = {k: torch.as_tensor(v, device=dev) for k, v in sc.items() if isinstance(v, np.ndarray) and v.dtype == np.float32 and k != "texcoords"}
    return sc, t


def layer(sc, t, albedo, shading, **over):
    a = dict(vertexPos_input=t["vertex_pos"], vertexColor_input=t["vertex_color"], texture_input=t["texture"],
             shCoeff_input=t["sh_coeff"], targetImage_input=t["target_image"], extrinsics_input=t["extrinsics"],
             intrinsics_input=t["intrinsics"])
    a.update(over)
    return CudaRendererGpu(faces_attr=sc["faces"].reshape(-1).tolist(), texCoords_attr=sc["texcoords"].reshape(-1).tolist(),
                           numberOfVertices_attr=sc["num_vertices"], numberOfCameras_attr=sc["num_cameras"],
                           renderResolutionU_attr=sc["width"], renderResolutionV_attr=sc["height"],
                           albedoMode_attr=albedo, shadingMode_attr=shading, **a)


def test_render_script_restated():
    """python/test_render.py: B=2, vertexColor + shaded, forward only."""
    sc, t = scene(kind="pyramid", cameras=1, width=128, height=128, batch=2)
    r = layer(sc, t, "vertexColor", "shaded")
    img = r.getRenderBufferTF()
    assert img.shape == (2, 1, 128, 128, 3) and r.getFaceBufferTF().dtype == torch.int32
    mask = r.getModelMaskTF()
    assert mask.shape == img.shape and 0.02 < float(mask.mean()) < 0.9
    assert r.getRenderBufferOpenCV(0, 0).shape == (128, 128, 3)
    assert torch.equal(r.getTargetBufferTF(), t["target_image"])
    assert r.getNormalMap() is None


def test_sh_fitting_restated():
    """python/test_gradients_SphericalHarmonics.py: Adam(lr 0.01) on l2 loss, shared SH [1,1,27]."""
    sc, t = scene(kind="sphere", rings=16, segments=20, cameras=2, width=96, height=96, batch=3)
    base = torch.as_tensor(synthetic.base_sh(), device="cuda:0").reshape(1, 1, 27)
    target = layer(sc, t, "vertexColor", "shaded", shCoeff_input=base.expand(3, 2, 27).contiguous()).getRenderBufferTF().detach()
    sh = (base + 0.5 * torch.rand(1, 1, 27, device="cuda:0", generator=torch.Generator("cuda").manual_seed(0))).requires_grad_(True)
    opt = torch.optim.Adam([sh], lr=0.01)
    losses = []
    n0 = len(_HANDLE_CACHE)
    for _ in range(30):
        opt.zero_grad()
        out = layer(sc, t, "vertexColor", "shaded", shCoeff_input=sh.expand(3, 2, 27)).getRenderBufferTF()
        loss = 0.5 * ((out - target) ** 2).sum()
        loss.backward()
        opt.step()
        losses.append(float(loss.detach()))
    assert losses[-1] < 0.6 * losses[0], losses
    assert len(_HANDLE_CACHE) <= n0 + 1          # the handle is cached across iterations, like a TF kernel


def test_vertex_colour_fitting_restated():
    """python/test_gradients_VertexColor.py: SGD on sum((out-target)^2)/(C*N) from zero colours."""
    sc, t = scene(kind="sphere", rings=16, segments=20, cameras=2, width=96, height=96)
    target = layer(sc, t, "vertexColor", "shaded").getRenderBufferTF().detach()
    col = torch.zeros_like(t["vertex_color"]).requires_grad_(True)
    opt = torch.optim.SGD([col], lr=1.0)   # the script's lr=10 is tuned to its own mesh/pixel ratio
    losses = []
    for _ in range(10):
        opt.zero_grad()
        out = layer(sc, t, "vertexColor", "shaded", vertexColor_input=col).getRenderBufferTF()
        loss = ((out - target) ** 2).sum() / (sc["num_cameras"] * sc["num_vertices"])
        loss.backward()
        opt.step()
        losses.append(float(loss.detach()))
    assert losses[-1] < 0.7 * losses[0], losses


def test_texture_fitting_restated():
    """python/test_gradients_Texture.py: textured + shadeless, the target render passed as targetImage."""
    sc, t = scene(kind="sphere", rings=16, segments=20, cameras=1, width=96, height=96, tex=32)
    target = layer(sc, t, "textured", "shadeless").getRenderBufferTF().detach()
    tex = torch.ones_like(t["texture"]).requires_grad_(True)
    opt = torch.optim.Adam([tex], lr=0.05)   # texels near the UV poles collect hundreds of pixels: plain SGD needs a tiny lr
    losses = []
    for _ in range(60):
        opt.zero_grad()
        r = layer(sc, t, "textured", "shadeless", texture_input=tex, targetImage_input=target)
        loss = ((r.getRenderBufferTF() - r.getTargetBufferTF()) ** 2).sum()
        loss.backward()
        opt.step()
        losses.append(float(loss.detach()))
    assert losses[-1] < 0.5 * losses[0], losses


def test_gradient_rules_of_the_python_layer():
    sc, t = scene(kind="sphere", rings=10, segments=12, cameras=1, width=48, height=48)
    leaves = {k: t[k].clone().requires_grad_(True) for k in ("vertex_pos", "vertex_color", "texture", "sh_coeff", "target_image", "extrinsics", "intrinsics")}
    kw = dict(vertexPos_input=leaves["vertex_pos"], vertexColor_input=leaves["vertex_color"], texture_input=leaves["texture"],
              shCoeff_input=leaves["sh_coeff"], targetImage_input=leaves["target_image"], extrinsics_input=leaves["extrinsics"],
              intrinsics_input=leaves["intrinsics"])
    # normal / lighting albedo: all-zero gradients (CudaRenderer.py:207-213)
    layer(sc, t, "normal", "shaded", **kw).getRenderBufferTF().sum().backward()
    assert all(float(v.grad.abs().max()) == 0 for v in leaves.values())
    for v in leaves.values():
        v.grad = None
    # vertexColor: position/colour/SH get gradients, target/extrinsics/intrinsics zeros (CudaRenderer.py:215)
    r = layer(sc, t, "vertexColor", "shaded", **kw)
    (r.getRenderBufferTF().sum() + r.getTargetBufferTF().sum()).backward()
    assert float(leaves["vertex_color"].grad.abs().max()) > 0 and float(leaves["sh_coeff"].grad.abs().max()) > 0
    assert float(leaves["vertex_pos"].grad.abs().max()) > 0
    for k in ("target_image", "extrinsics", "intrinsics"):
        assert float(leaves[k].grad.abs().max()) == 0


def test_texture_fitting_with_the_bilinear_variant_and_smoothing():
    """BASELINE.json config 3 wording: textured mode with the bilinear texture-gradient scatter
    (textureBilinear_attr=True, the variant the reference has commented out), and the loss-side smoothImage helper
    (python/utils/GaussianSmoothingGpu.py) on both images: the fit must converge like the nearest-texel one."""
    from gvv_differentiable_cuda_renderer_b200.utils import GaussianSmoothingGpu
    sc, t = scene(kind="sphere", rings=16, segments=20, cameras=2, width=96, height=96, tex=32)
    target = layer(sc, t, "textured", "shadeless", textureBilinear_attr=True).getRenderBufferTF().detach()
    target_s = GaussianSmoothingGpu.smoothImage(target, 1, 0.0, 0.8)
    assert len(_HANDLE_CACHE) >= 1
    tex = torch.ones_like(t["texture"]).requires_grad_(True)
    opt = torch.optim.Adam([tex], lr=0.05)
    losses = []
    for _ in range(60):
        opt.zero_grad()
        r = layer(sc, t, "textured", "shadeless", texture_input=tex, targetImage_input=target, textureBilinear_attr=True)
        loss = ((GaussianSmoothingGpu.smoothImage(r.getRenderBufferTF(), 1, 0.0, 0.8) - target_s) ** 2).sum()
        loss.backward()
        opt.step()
        losses.append(float(loss.detach()))
    assert losses[-1] < 0.3 * losses[0], losses
    # the default layer on the same attributes is a different handle (nearest texel): different image
    near = layer(sc, t, "textured", "shadeless").getRenderBufferTF()
    assert not torch.equal(near, target)


def test_layer_step_is_cuda_graph_capturable_through_autograd():
    """A whole fit step through the public layer -- CudaRendererGpu(...), loss, autograd -- captured with
    torch.cuda.graph and replayed on new input values reproduces the eager loss and gradients (bench.py's e2e leg)."""
    sc, t = scene(kind="sphere", rings=20, segments=24, cameras=2, width=128, height=96, tex=16)
    dev = t["vertex_pos"].device
    g_img = torch.randn((1, 2, 96, 128, 3), generator=torch.Generator().manual_seed(0)).to(dev)
    static = {k: t[k].clone() for k in ("vertex_pos", "vertex_color", "sh_coeff")}

    def step(d):
        leaves = {k: d[k].detach().requires_grad_(True) for k in d}
        r = layer(sc, t, "vertexColor", "shaded", vertexPos_input=leaves["vertex_pos"], vertexColor_input=leaves["vertex_color"],
                  shCoeff_input=leaves["sh_coeff"])
        loss = (r.getRenderBufferTF() * g_img).sum()
        return loss, torch.autograd.grad(loss, list(leaves.values()))

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            step(static)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        loss_s, grads_s = step(static)
    for trial in range(2):
        static["vertex_color"].copy_(torch.rand_like(static["vertex_color"]))       # new values in the static inputs
        static["sh_coeff"].mul_(1.0 + 0.1 * trial)
        graph.replay()
        torch.cuda.synchronize()
        loss_e, grads_e = step({k: v.clone() for k, v in static.items()})
        assert abs(float(loss_s) - float(loss_e)) <= 1e-4 * max(1.0, abs(float(loss_e)))
        for a, b in zip(grads_s, grads_e):
            assert float((a - b).norm()) <= 1e-4 * float(b.norm()) + 1e-6
    del graph


def test_handle_cache_eviction_keeps_live_handles_usable():
    """More distinct attribute sets than the cache holds (16): the evicted handle must stay alive while a layer or a
    pending autograd graph still uses it -- eviction only drops the cache's reference."""
    from gvv_differentiable_cuda_renderer_b200 import CudaRenderer as mod
    sc, t = scene(kind="sphere", rings=8, segments=10, cameras=1, width=40, height=40)
    col = t["vertex_color"].clone().requires_grad_(True)
    first = layer(sc, t, "vertexColor", "shaded", vertexColor_input=col)
    h0 = first._handle
    loss = first.getRenderBufferTF().sum()                  # graph outstanding across the evictions below
    for k in range(mod._HANDLE_CACHE_MAX + 3):              # a resolution pyramid: a new handle per size
        sc2, t2 = scene(kind="sphere", rings=8, segments=10, cameras=1, width=41 + k, height=40)
        layer(sc2, t2, "vertexColor", "shaded")
    assert len(mod._HANDLE_CACHE) <= mod._HANDLE_CACHE_MAX and all(h is not h0 for h in mod._HANDLE_CACHE.values())
    assert h0._h                                            # evicted from the cache, not destroyed
    loss.backward()
    assert float(col.grad.abs().max()) > 0


def test_device_without_index_and_argument_validation():
    """device='cuda' means the current device; wrong dtypes / sizes raise instead of reading out of bounds."""
    from gvv_differentiable_cuda_renderer_b200 import _native
    sc, t = scene(kind="sphere", rings=8, segments=10, cameras=1, width=48, height=32)
    r = layer(sc, t, "vertexColor", "shaded", device="cuda")
    assert r.getRenderBufferTF().device == torch.device("cuda", torch.cuda.current_device())
    nr = _native.NativeRenderer(sc["faces"], sc["texcoords"], sc["num_vertices"], 1, 48, 32, "vertexColor", "shaded", device="cuda")
    ins = [t[k] for k in ("vertex_pos", "vertex_color", "texture", "sh_coeff", "target_image", "extrinsics", "intrinsics")]
    bary, face, render, vn, _, _ = nr.forward(*ins)
    g = torch.ones_like(render)
    with pytest.raises(_native.GvvError, match="int32"):
        nr.backward(g, None, ins[0], ins[1], ins[2], ins[3], ins[4], vn, bary, face.long(), ins[5], ins[6])
    with pytest.raises(_native.GvvError, match="render_buffer_grad"):
        nr.backward(g[..., :2].contiguous(), None, ins[0], ins[1], ins[2], ins[3], ins[4], vn, bary, face, ins[5], ins[6])
    with pytest.raises(_native.GvvError, match="vertex_color"):
        nr.forward(ins[0], ins[1][:, :-1].contiguous(), *ins[2:])
    with pytest.raises(_native.GvvError, match="CUDA device"):
        _native.NativeRenderer(sc["faces"], sc["texcoords"], sc["num_vertices"], 1, 48, 32, "vertexColor", "shaded", device="cpu")
    nr.close()


def test_scratch_growth_is_refused_once_graphs_hold_its_pointers():
    """gvv_reserve sizes the scratch before capture; growing it afterwards would free memory that captured graphs
    still use, so the call is refused with an error instead."""
    from gvv_differentiable_cuda_renderer_b200 import _native
    sc1, t1 = scene(kind="sphere", rings=10, segments=12, cameras=2, width=64, height=48, batch=1)
    sc3, t3 = scene(kind="sphere", rings=10, segments=12, cameras=2, width=64, height=48, batch=3)
    keys = ("vertex_pos", "vertex_color", "texture", "sh_coeff", "target_image", "extrinsics", "intrinsics")
    nr = _native.NativeRenderer(sc1["faces"], sc1["texcoords"], sc1["num_vertices"], 2, 64, 48, "vertexColor", "shaded")
    nr.reserve(2)
    ins1 = [t1[k] for k in keys]
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        nr.forward(*ins1)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        out = nr.forward(*ins1)
    graph.replay()
    torch.cuda.synchronize()
    eager = nr.forward(*ins1)
    assert torch.equal(out[1], eager[1])
    with pytest.raises(_native.GvvError, match="captured"):
        nr.forward(*[t3[k] for k in keys])                    # batch 3 > reserved 2: refused, the graph's scratch stays valid
    graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(out[1], eager[1])
    del graph
    nr.close()
