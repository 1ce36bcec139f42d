"""CPU: the numpy restatements of the loss-side helpers (oracle/helpers.py) on known answers."""
import numpy as np

from oracle import helpers


def test_gaussian_kernel_matches_the_reference_construction():
    k = helpers.gaussian_kernel_2d(2, 0.0, 1.0)
    assert k.shape == (5, 5) and abs(k.sum() - 1.0) < 1e-12
    assert np.allclose(k, k.T) and np.allclose(k, k[::-1, ::-1])         # symmetric for mean 0
    v = np.exp(-0.5 * np.arange(-2, 3) ** 2.0)
    assert np.allclose(k, np.outer(v, v) / np.outer(v, v).sum())


def test_smooth_image_constant_interior_and_zero_padding():
    img = np.ones((1, 1, 9, 9, 3), np.float32)
    out = helpers.smooth_image(img, 1, 0.0, 1.0)
    assert np.allclose(out[0, 0, 1:-1, 1:-1], 1.0)                       # interior: kernel sums to one
    assert out[0, 0, 0, 0, 0] < 1.0                                      # corner: zero padding leaks in
    assert helpers.smooth_image(img, 0, 0.0, 1.0) is img                # reference :14-15


def test_image_gradient_of_a_ramp():
    # I = 2x + 3y: the filter of RendererUtil.h:566-620 with fs = 1 (weights x/(x^2+y^2), normalised by sum|G_u| = 4)
    H, W = 12, 14
    y, x = np.mgrid[0:H, 0:W].astype(np.float32)
    img = np.repeat((2 * x + 3 * y)[..., None], 3, -1)
    du, dv = helpers.image_gradient(img, 1)
    # sum_x x*G_u = 2 (x=+-1,y=0) + 4*0.5 (corners) = 4 -> /4 = 1 per unit slope
    assert np.allclose(du[2:-2, 2:-2], 2.0, atol=1e-5) and np.allclose(dv[2:-2, 2:-2], 3.0, atol=1e-5)
    assert np.all(du[:2] == 0) and np.all(du[:, :2] == 0) and np.all(du[-2:] == 0) and np.all(du[:, -2:] == 0)


def test_smooth_image_oracle_is_pinned_by_the_golden_fixture():
    """tests/golden/helpers/smooth_image.npz: smoothImage (python/utils/GaussianSmoothingGpu.py:12-37) restated with
    torch.distributions.Normal + torch.nn.functional.conv2d in fp64 by tools/make_smooth_golden.py (an independent
    implementation of tfp Normal.prob / tf.nn.depthwise_conv2d SAME); pins oracle/helpers.smooth_image."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "helpers", "smooth_image.npz"))
    for i, (size, mean, std) in enumerate(g["cases"]):
        out = helpers.smooth_image(g["image"], int(size), float(mean), float(std))
        assert np.abs(out - g[f"smoothed_{i}"]).max() <= 2e-7, (i, np.abs(out - g[f"smoothed_{i}"]).max())
