"""TEST INFRASTRUCTURE ONLY (oracle).  CPU restatement of the reference's gradient smoothing pass,
seistorch/signal.py:247-319 (gaussian_filter, 2D input): numpy 'reflect' padding by kernel_size // 2 on both axes, a
normalised float32 Gaussian of 2*radius+1 taps (even radius) convolved along one axis, the padding cropped again.
Never imported by the product path."""
from __future__ import annotations

import numpy as np


def gaussian_weights(sigma, radius):
    k = 2 * radius + 1 if radius % 2 == 0 else 2 * radius
    w = np.exp(-(np.arange(k) - k // 2) ** 2 / (2 * sigma ** 2)).astype(np.float32)
    return w / w.sum(dtype=np.float32)


def gaussian_filter(x, sigma, radius, axis):
    """x: 2D float array; returns float32 (conv2d accumulates in float32; float64 here, rounded at the end)."""
    if radius % 2 != 0:
        raise ValueError("odd radius: the reference returns a different shape (signal.py:263-266)")
    w = gaussian_weights(sigma, radius).astype(np.float64)
    p = len(w) // 2
    xp = np.pad(np.asarray(x, dtype=np.float64), ((p, p), (p, p)), mode="reflect")
    out = np.zeros_like(np.asarray(x, dtype=np.float64))
    n0, n1 = out.shape
    for j, wj in enumerate(w):
        if axis == 0:
            out += wj * xp[j:j + n0, p:p + n1]
        else:
            out += wj * xp[p:p + n0, j:j + n1]
    return out.astype(np.float32)


def smooth_gradient(g, counts, sigma, radius):
    """process.py:66-112 for one 2D gradient: `counts` x (z pass, x pass)."""
    for _ in range(counts):
        g = gaussian_filter(g, sigma["z"], radius["z"], 0)
        g = gaussian_filter(g, sigma["x"], radius["x"], 1)
    return g
