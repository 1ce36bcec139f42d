#!/usr/bin/env python3
"""TEST INFRASTRUCTURE -- golden fixture for smoothImage (python/utils/GaussianSmoothingGpu.py:12-37 of the reference).

TensorFlow / tensorflow_probability cannot run here, so the reference function is restated with an INDEPENDENT library
implementation of the same two primitives, step by step as the reference writes them:

    d.prob(tf.range(-size, size + 1, dtype=tf.float32))    -> torch.distributions.Normal(mean, std).log_prob(...).exp() in fp32
    tf.einsum('i,j->ij', vals, vals) / tf.reduce_sum(...)  -> torch.einsum + division in fp32
    tf.nn.depthwise_conv2d(x, k[:, :, None, None] tiled to 3 channels, strides 1, padding "SAME")
                                                           -> torch.nn.functional.conv2d(groups = 3, padding = size): both are
                                                              cross-correlations; "SAME" with an odd kernel and stride 1 pads
                                                              `size` zeros on every side.  Evaluated in fp64.

Neither oracle/helpers.py (numpy, explicit shifted sums) nor the CUDA kernel (two separable passes) shares code with
this script.  Writes tests/golden/helpers/smooth_image.npz:  python tools/make_smooth_golden.py
"""
import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = [(1, 0.0, 0.8), (2, 0.0, 1.0), (3, 0.5, 2.0), (4, 0.0, 1.5)]


def smooth_reference(image, size, mean, std):
    d = torch.distributions.Normal(torch.tensor(mean, dtype=torch.float32), torch.tensor(std, dtype=torch.float32))
    vals = d.log_prob(torch.arange(-size, size + 1, dtype=torch.float32)).exp()
    k = torch.einsum("i,j->ij", vals, vals)
    k = k / k.sum()
    B, C, H, W, _ = image.shape
    x = torch.as_tensor(image, dtype=torch.float64).reshape(B * C, H, W, 3).permute(0, 3, 1, 2)
    w = k.to(torch.float64)[None, None].repeat(3, 1, 1, 1)
    y = torch.nn.functional.conv2d(x, w, padding=size, groups=3)
    return y.permute(0, 2, 3, 1).reshape(B, C, H, W, 3).numpy()


def main():
    rng = np.random.default_rng(7)
    image = rng.random((2, 2, 17, 13, 3), dtype=np.float32)
    out = {"image": image, "cases": np.asarray(CASES, np.float64)}
    for i, (size, mean, std) in enumerate(CASES):
        out[f"smoothed_{i}"] = smooth_reference(image, size, mean, std)
    path = os.path.join(ROOT, "tests", "golden", "helpers", "smooth_image.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
