"""GPU (-m gpu): loss-side helpers next to the op against their numpy restatements."""
import numpy as np
import pytest
import torch

from gvv_differentiable_cuda_renderer_b200 import _native, synthetic
from gvv_differentiable_cuda_renderer_b200.utils import GaussianSmoothingGpu

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("size,mean,std", [(1, 0.0, 0.8), (3, 0.0, 1.5), (2, 0.3, 1.0)])
def test_smooth_image_matches_numpy(size, mean, std):
    from oracle import helpers
    rng = np.random.default_rng(size)
    img = rng.random((2, 3, 37, 29, 3), dtype=np.float32)
    out = GaussianSmoothingGpu.smoothImage(torch.as_tensor(img, device=DEV), size, mean, std)
    ref = helpers.smooth_image(img, size, mean, std)
    assert out.shape == img.shape
    assert np.abs(out.cpu().numpy() - ref).max() <= 2e-6            # fp32 separable vs fp64 2-D: tolerance 2e-6 on [0,1) data
    same = GaussianSmoothingGpu.smoothImage(torch.as_tensor(img, device=DEV), 0, 0.0, 1.0)
    assert torch.equal(same, torch.as_tensor(img, device=DEV))       # size 0 / std 0 return the input (reference :14-15)


def test_smooth_image_matches_the_golden_fixture():
    """tests/golden/helpers/smooth_image.npz (tools/make_smooth_golden.py: the reference's smoothImage restated with
    torch.distributions.Normal + conv2d in fp64) -- pins the CUDA helper, tolerance 2e-6 on [0,1) data."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "helpers", "smooth_image.npz"))
    x = torch.as_tensor(g["image"], device=DEV)
    for i, (size, mean, std) in enumerate(g["cases"]):
        out = GaussianSmoothingGpu.smoothImage(x, int(size), float(mean), float(std))
        assert np.abs(out.cpu().numpy() - g[f"smoothed_{i}"]).max() <= 2e-6


def test_smooth_image_gradient_is_the_adjoint():
    g = torch.Generator().manual_seed(0)
    x = torch.randn((1, 2, 20, 24, 3), generator=g).to(DEV).requires_grad_(True)
    y = torch.randn((1, 2, 20, 24, 3), generator=g).to(DEV)
    out = GaussianSmoothingGpu.smoothImage(x, 2, 0.4, 1.1)           # asymmetric kernel: adjoint != forward
    (out * y).sum().backward()
    # <S x, y> = <x, S^T y> for a second random x2
    x2 = torch.randn((1, 2, 20, 24, 3), generator=g).to(DEV)
    lhs = float((GaussianSmoothingGpu.smoothImage(x2, 2, 0.4, 1.1).double() * y.double()).sum())
    rhs = float((x2.double() * x.grad.double()).sum())
    assert abs(lhs - rhs) <= 1e-5 * max(1.0, abs(lhs)), (lhs, rhs)


@pytest.mark.parametrize("fs", [1, 2])
def test_image_gradient_matches_numpy(fs):
    from oracle import helpers
    rng = np.random.default_rng(7 + fs)
    img = rng.random((2, 2, 31, 45, 3), dtype=np.float32)
    du, dv = _native.image_gradient(torch.as_tensor(img, device=DEV), fs)
    ru, rv = helpers.image_gradient(img, fs)
    assert np.abs(du.cpu().numpy() - ru).max() <= 1e-6 and np.abs(dv.cpu().numpy() - rv).max() <= 1e-6
    assert float(du[..., : fs + 1, :, :].abs().max()) == 0 and float(dv[..., :, : fs + 1, :].abs().max()) == 0


def test_backward_with_precomputed_target_gradient():
    """gvv_set_target_gradient: the model-to-data term reads the precomputed imageGradient instead of evaluating
    the (2s+1)^2 window per covered pixel per call -- same numbers."""
    sc = synthetic.make_scene(kind="sphere", rings=14, segments=18, cameras=2, width=80, height=72, tex=16, seed=9)
    N, C, W, H = sc["num_vertices"], sc["num_cameras"], sc["width"], sc["height"]
    dev = torch.device(DEV)
    T = lambda a: torch.as_tensor(np.ascontiguousarray(a), device=dev)
    ins = [T(sc[k]) for k in ("vertex_pos", "vertex_color", "texture", "sh_coeff", "target_image", "extrinsics", "intrinsics")]
    rng = np.random.default_rng(1)
    target = T(rng.random((1, C, H, W, 3), dtype=np.float32))
    rg = T(rng.standard_normal((1, C, H, W, 3)).astype(np.float32))
    tg = T(rng.standard_normal((1, C, H, W, 3)).astype(np.float32))
    for fs in (1, 2):
        r = _native.NativeRenderer(sc["faces"], sc["texcoords"], N, C, W, H, "vertexColor", "shaded", fs, 1, False, dev)
        bary, face, render, vn, _, _ = r.forward(ins[0], ins[1], ins[2], ins[3], target, ins[5], ins[6])
        g0 = r.backward(rg, tg, ins[0], ins[1], ins[2], ins[3], target, vn, bary, face, ins[5], ins[6])
        du, dv = _native.image_gradient(target, fs)
        r.set_target_gradient(du, dv)
        g1 = r.backward(rg, tg, ins[0], ins[1], ins[2], ins[3], target, vn, bary, face, ins[5], ins[6])
        r.set_target_gradient(None, None)
        g2 = r.backward(rg, tg, ins[0], ins[1], ins[2], ins[3], target, vn, bary, face, ins[5], ins[6])
        assert float(g0[0].abs().max()) > 0
        for a, b, c in zip(g0, g1, g2):
            d = float(a.abs().max())
            assert float((a - b).abs().max()) <= 1e-5 * max(d, 1e-6)      # atomic accumulation order only
            assert float((a - c).abs().max()) <= 1e-5 * max(d, 1e-6)
        r.close()
