"""TEST INFRASTRUCTURE ONLY -- numpy restatements of the loss-side helpers.

  smooth_image    python/utils/GaussianSmoothingGpu.py:12-37 (tf.nn.depthwise_conv2d, SAME padding, 2-D kernel
                  outer(vals, vals) / sum) -- evaluated here literally as a 2-D cross-correlation, NOT separably
  image_gradient  imageGradient, cpp/src/Utils/RendererUtil.h:566-620, at every integer pixel
smooth_image is pinned by tests/golden/helpers/smooth_image.npz, a fixture produced by an independent restatement of the
reference function on library primitives (tools/make_smooth_golden.py: torch.distributions.Normal + conv2d in fp64;
TensorFlow itself is not installable here); image_gradient is the same code path the pinned backward oracle
(gvv_oracle.cpp) uses for its model-to-data term.
"""
import math

import numpy as np


def gaussian_kernel_2d(size, mean, std):
    x = np.arange(-size, size + 1, dtype=np.float64)
    vals = np.exp(-0.5 * ((x - mean) / std) ** 2) / (std * math.sqrt(2.0 * math.pi))
    k = np.einsum("i,j->ij", vals, vals)
    return k / k.sum()


def smooth_image(image, size, mean, std):
    """image [..., H, W, 3] -> same shape; zero padding, cross-correlation (kernel index i <-> offset i - size)."""
    if size == 0 or std == 0.0:
        return image
    k = gaussian_kernel_2d(size, mean, std)
    img = np.asarray(image, np.float64)
    H, W = img.shape[-3], img.shape[-2]
    pad = [(0, 0)] * (img.ndim - 3) + [(size, size), (size, size), (0, 0)]
    p = np.pad(img, pad)
    out = np.zeros_like(img)
    for i in range(2 * size + 1):
        for j in range(2 * size + 1):
            out += k[i, j] * p[..., i:i + H, j:j + W, :]
    return out


def image_gradient(image, fs):
    """image [..., H, W, 3] -> (dI/du, dI/dv), zero within fs+1 pixels of the border."""
    img = np.asarray(image, np.float32)
    H, W = img.shape[-3], img.shape[-2]
    du, dv = np.zeros_like(img), np.zeros_like(img)
    norm = np.float32(0)
    ys, xs = slice(fs + 1, H - (fs + 1)), slice(fs + 1, W - (fs + 1))
    if H - 2 * (fs + 1) <= 0 or W - 2 * (fs + 1) <= 0:
        return du, dv
    au, av = np.zeros_like(img[..., ys, xs, :]), np.zeros_like(img[..., ys, xs, :])
    for yy in range(-fs, fs + 1):
        for xx in range(-fs, fs + 1):
            den = np.float32(xx * xx + yy * yy)
            gu = np.float32(xx) / den if den != 0 else np.float32(0)
            gv = np.float32(yy) / den if den != 0 else np.float32(0)
            I = img[..., fs + 1 + yy:H - (fs + 1) + yy, fs + 1 + xx:W - (fs + 1) + xx, :]
            au = au + I * gu
            av = av + I * gv
            norm = norm + abs(gu)
    du[..., ys, xs, :] = au / norm
    dv[..., ys, xs, :] = av / norm
    return du, dv
