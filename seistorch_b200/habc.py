"""HABC masks and blend weights -- same values as seistorch/habc.py:4-81.

The sm_100a kernels evaluate side ownership in closed form (csrc/st_wave2d_math.cuh,
w2_side_weights); ``bound_mask`` is kept because WaveCell.setup_habc and the per-step
``_time_step(..., habcs=...)`` signature of the reference pass the masks around."""
from __future__ import annotations

import torch


def _masks(nz, nx, w, dev, multiple):
    r = torch.arange(w, device=dev)[:, None]
    c = torch.arange(nx, device=dev)[None, :]
    top = (r <= c) & (c <= nx - 1 - r)
    bottom = torch.flip(top, [0])
    i = torch.arange(nz, device=dev)[:, None]
    j = torch.arange(w, device=dev)[None, :]
    left = (j <= i) & (j <= nz - 1 - i)
    right = torch.flip(left, [1])
    if multiple:
        left = left.clone()
        right = right.clone()
        left[:w] = True
        right[:w] = True
    return top, bottom, left, right


def bound_mask(nz, nx, w, dev, batchsize=1, return_idx=False, multiple=False):
    """habc.py:4-40.  Float masks (return_idx=False) or boolean masks repeated over the
    batch (return_idx=True); the top mask is None for `multiple`."""
    top, bottom, left, right = _masks(nz, nx, w, dev, multiple)
    if not return_idx:
        f = lambda m: m.to(torch.get_default_dtype())
        return (None if multiple else f(top)), f(bottom), f(left), f(right)
    rep = (lambda m: m.repeat(batchsize, 1, 1)) if batchsize > 1 else (lambda m: m)
    return (None if multiple else rep(top)), rep(bottom), rep(left), rep(right)


def generate_habc_coefficients_2d(domain_shape, N=50, multiple=False, device="cpu"):
    """habc.py:42-81: linear weight 1 (outer edge) -> 0 (depth N-1), written side by side
    in the order top, bottom, left, right."""
    nz, nx = domain_shape
    d = torch.zeros(nz, nx, device=device)
    vals = torch.flip(torch.linspace(0.0, N, N, device=device) / N, [0])
    top, bottom, left, right = _masks(nz, nx, N, device, multiple)
    if N > 0:
        if not multiple:
            d[:N][top] = vals[:, None].expand(N, nx)[top]
        d[nz - N:][bottom] = torch.flip(vals, [0])[:, None].expand(N, nx)[bottom]
        d[:, :N][left] = vals[None, :].expand(nz, N)[left]
        d[:, nx - N:][right] = torch.flip(vals, [0])[None, :].expand(nz, N)[right]
    return d
