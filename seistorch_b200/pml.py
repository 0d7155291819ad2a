"""PML damping profiles -- same values as seistorch/pml.py:29-81 (setup-time inputs of
the kernels)."""
from __future__ import annotations

import numpy as np
import torch


def generate_pml_coefficients_2d(domain_shape, N=50, B=100.0, multiple=False):
    """pml.py:29-59: d0*(k/N)^2 with d0 = 1.5*1500/N*log10(1e4), both directions combined
    as sqrt(dz^2 + dx^2); `multiple` drops the top layer.  (`B` is unused there too.)"""
    nz, nx = domain_shape
    d0 = (1.5 * 1500.0 / N) * np.log10(1.0 / 1e-4)
    prof = torch.flip(d0 * torch.linspace(0.0, 1.0, N + 1) ** 2, [0])
    dz = torch.zeros(nz, nx)
    dx = torch.zeros(nz, nx)
    if N > 0:
        if not multiple:
            dz[0:N + 1, :] = prof[:, None]
        dz[nz - N - 1:nz, :] = torch.flip(prof, [0])[:, None]
        dx[:, 0:N + 1] = prof[None, :]
        dx[:, nx - N - 1:nx] = torch.flip(prof, [0])[None, :]
    return torch.sqrt(dz ** 2 + dx ** 2)


def generate_pml_coefficients_3d(domain_shape, N=50, B=100.0, multiple=False):
    """pml.py:61-81: cosine profile B*(1-cos(pi*idx)) on the three tensor dims."""
    n0, n1, n2 = domain_shape
    idx = (torch.ones(N + 1) * (N + 1) - torch.linspace(0.0, (N + 1), N + 1)) / (2 * (N + 1))
    vals = B * (1.0 - torch.cos(torch.pi * idx))
    rv = torch.flip(vals, [0])
    b0 = torch.zeros((n0, n1, n2))
    b1 = torch.zeros((n0, n1, n2))
    b2 = torch.zeros((n0, n1, n2))
    b1[:, 0:N + 1, :] = vals[None, :, None]
    b1[:, n1 - N - 1:n1, :] = rv[None, :, None]
    b2[:, :, 0:N + 1] = vals[None, None, :]
    b2[:, :, n2 - N - 1:n2] = rv[None, None, :]
    b0[0:N + 1, :, :] = vals[:, None, None]
    b0[n0 - N - 1:n0, :, :] = rv[:, None, None]
    return torch.sqrt(b0 ** 2 + b1 ** 2 + b2 ** 2)
