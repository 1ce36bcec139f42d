"""smoothImage of the reference (python/utils/GaussianSmoothingGpu.py:12-37) on our CUDA kernels.

Same signature and semantics: `image` [B, C, H, W, 3], a (2*size+1)^2 Gaussian N(mean, std) sampled at the
integers -size..size, normalised to sum one, applied per channel with zero "SAME" padding
(tf.nn.depthwise_conv2d = cross-correlation).  The 2-D kernel is outer(vals, vals)/sum(outer), i.e. two
normalised 1-D passes (gvv_gaussian_smooth).  Differentiable: the adjoint of a zero-padded correlation
is the correlation with the reversed taps.
"""
import math

import numpy as np
import torch

from .. import _native


def gaussian_taps(size, mean, std):
    x = np.arange(-size, size + 1, dtype=np.float64)
    vals = np.exp(-0.5 * ((x - mean) / std) ** 2) / (std * math.sqrt(2.0 * math.pi))   # tfp Normal(mean, std).prob
    return (vals / vals.sum()).astype(np.float32)


class _Smooth(torch.autograd.Function):
    @staticmethod
    def forward(ctx, image, taps):
        ctx.taps = taps
        return _native.gaussian_smooth(image, taps)

    @staticmethod
    def backward(ctx, g):
        return _native.gaussian_smooth(g, ctx.taps[::-1].copy()), None


def smoothImage(image, size: int, mean: float, std: float):
    if size == 0 or std == 0.0:
        return image
    if not (isinstance(image, torch.Tensor) and image.is_cuda):
        raise _native.GvvError("smoothImage needs a CUDA tensor: there is no CPU path")
    return _Smooth.apply(image, gaussian_taps(size, mean, std))
