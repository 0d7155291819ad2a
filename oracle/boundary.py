"""TEST INFRASTRUCTURE ONLY (oracle).  CPU restatement of the reference's
absorbing-boundary coefficient generators.  Never imported by the product path.

Follows:
  * seistorch/pml.py:29-59   generate_pml_coefficients_2d
  * seistorch/pml.py:61-81   generate_pml_coefficients_3d
  * seistorch/habc.py:4-40   bound_mask
  * seistorch/habc.py:42-81  generate_habc_coefficients_2d
"""
from __future__ import annotations

import numpy as np
import torch


def pml_coefficients_2d(domain_shape, N=50, multiple=False, dtype=torch.float32):
    """pml.py:29-59.  domain_shape = (nz, nx) (called (Nx, Ny) there); quadratic
    profile d0*(k/N)^2, d0 = 1.5*1500/N*log10(1e4); the two directions are
    combined as sqrt(dx^2 + dy^2).  Computed in the *default* dtype then cast, so
    the fp64 oracle sees the same fp32-rounded profile as the reference's own
    fp64 run would only if default dtype is fp64 there too; we therefore take a
    dtype argument and build directly in it (reference builds under
    torch.set_default_dtype(cfg dtype), utils.py:251-257)."""
    nz, nx = domain_shape
    d0 = (1.5 * 1500.0 / N) * np.log10(1.0 / 1e-4)
    prof = d0 * torch.linspace(0.0, 1.0, N + 1, dtype=dtype) ** 2
    prof = torch.flip(prof, [0])  # N+1 values, strongest at the outer edge
    dz = torch.zeros(nz, nx, dtype=dtype)
    dx = torch.zeros(nz, nx, dtype=dtype)
    if N > 0:
        if not multiple:
            dz[0:N + 1, :] = prof[:, None]
        dz[nz - N - 1:nz, :] = torch.flip(prof, [0])[:, None]
        dx[:, 0:N + 1] = prof[None, :]
        dx[:, nx - N - 1:nx] = torch.flip(prof, [0])[None, :]
    return torch.sqrt(dz ** 2 + dx ** 2)


def pml_coefficients_3d(domain_shape, N=50, B=100.0, dtype=torch.float32):
    """pml.py:61-81.  Cosine profile B*(1-cos(pi*idx)), three directions combined
    by the Euclidean norm.  Axis naming follows the reference literally:
    domain_shape is unpacked as (nz, ny, nx) = tensor dims (0, 1, 2)."""
    n0, n1, n2 = domain_shape
    idx = (torch.ones(N + 1, dtype=dtype) * (N + 1)
           - torch.linspace(0.0, (N + 1), N + 1, dtype=dtype)) / (2 * (N + 1))
    vals = torch.cos(torch.pi * idx)
    vals = B * (1.0 - vals)
    b1 = torch.zeros((n0, n1, n2), dtype=dtype)
    b2 = torch.zeros((n0, n1, n2), dtype=dtype)
    b0 = torch.zeros((n0, n1, n2), dtype=dtype)
    b1[:, 0:N + 1, :] = vals[None, :, None]
    b1[:, n1 - N - 1:n1, :] = torch.flip(vals, [0])[None, :, None]
    b2[:, :, 0:N + 1] = vals[None, None, :]
    b2[:, :, n2 - N - 1:n2] = torch.flip(vals, [0])[None, None, :]
    b0[0:N + 1, :, :] = vals[:, None, None]
    b0[n0 - N - 1:n0, :, :] = torch.flip(vals, [0])[:, None, None]
    return torch.sqrt(b0 ** 2 + b1 ** 2 + b2 ** 2)


def habc_masks(nz, nx, w=50, multiple=False):
    """habc.py:4-40 in closed form.  Returns four boolean arrays
    top (w, nx), bottom (w, nx), left (nz, w), right (nz, w)."""
    r = np.arange(w)[:, None]
    c = np.arange(nx)[None, :]
    top = (r <= c) & (c <= nx - 1 - r)
    bottom = top[::-1].copy()
    i = np.arange(nz)[:, None]
    j = np.arange(w)[None, :]
    left = (j <= i) & (j <= nz - 1 - i)
    right = left[:, ::-1].copy()
    if multiple:
        left[:w] = True
        right[:w] = True
        top = None
    return top, bottom, left, right


def habc_coefficients_2d(domain_shape, N=50, multiple=False, dtype=torch.float32):
    """habc.py:42-81.  Linear blend weight: 1 at the outer edge, 0 at depth N-1,
    written side by side in the order top, bottom, left, right."""
    nz, nx = domain_shape
    d = torch.zeros(nz, nx, dtype=dtype)
    vals = torch.flip(torch.linspace(0.0, N, N, dtype=dtype) / N, [0])
    tm, bm, lm, rm = habc_masks(nz, nx, N, multiple)
    if N > 0:
        if not multiple:
            t = torch.from_numpy(tm)
            d[:N][t] = vals[:, None].expand(N, nx)[t]
        b = torch.from_numpy(bm)
        d[nz - N:][b] = torch.flip(vals, [0])[:, None].expand(N, nx)[b]
        l = torch.from_numpy(lm)
        d[:, :N][l] = vals[None, :].expand(nz, N)[l]
        r_ = torch.from_numpy(rm)
        d[:, nx - N:][r_] = torch.flip(vals, [0])[None, :].expand(nz, N)[r_]
    return d
