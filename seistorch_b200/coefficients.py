"""Model parameters -> dimensionless kernel coefficients (once per forward call).

The kernels never see vp/eps/...: they take coefficient planes (r = vp*dt/h, cxx, czz,
...) and return d loss / d coefficient; the O(N) maps below are plain differentiable
torch ops evaluated ONCE per call on the device (in fp64, then rounded to fp32), so
autograd chains the kernel gradients back to the reference's parameters.  They restate
the coefficient algebra that the reference re-evaluates every time step:

  acoustic / acoustic_habc      equations2d/acoustic.py:73-86, acoustic_habc.py:206-221
  vti_habc2 / tti_habc          equations2d/vti_habc2.py:37-55, tti_habc.py:31-57
  acoustic_lsrtm_habc           equations2d/acoustic_lsrtm_habc.py:10-30
  acoustic_rho_habc             equations2d/acoustic_rho_habc.py:32-57
  acoustic_{vti,tti}_lsrtm_habc equations2d/acoustic_vti_lsrtm_habc.py:33-62, ..tti..:31-68
  acoustic_fwim_habc            equations2d/acoustic_fwim_habc.py:38-60
  elastic                       equations2d/elastic.py:11-13,20-24,33-35
  acoustic (3D)                 equations3d/acoustic.py:72-85
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from .engine import EQ_BORN, EQ_G1, EQ_HABC, EQ_ISO, EQ_PML, EQ_XZ

# equation name -> (family, flags, wavefield names, parameter names, field-channel of each wavefield)
EQUATIONS = {
    "acoustic": ("wave2d", EQ_ISO | EQ_PML),
    "acoustic_habc": ("wave2d", EQ_ISO | EQ_HABC),
    "vti_habc2": ("wave2d", EQ_HABC),
    "tti_habc": ("wave2d", EQ_HABC | EQ_XZ),
    "acoustic_fwim_habc": ("wave2d", EQ_ISO | EQ_HABC | EQ_G1),
    "acoustic_lsrtm_habc": ("wave2d", EQ_HABC | EQ_BORN),          # isotropic Born pair: cxx = czz
    "acoustic_rho_habc": ("wave2d", EQ_HABC | EQ_G1),              # variable density: four different neighbour weights
    "acoustic_vti_lsrtm_habc": ("wave2d", EQ_HABC | EQ_BORN),
    "acoustic_tti_lsrtm_habc": ("wave2d", EQ_HABC | EQ_XZ | EQ_BORN),
    "elastic": ("elastic2d", 0),
}


def _f64(x):
    return x.to(torch.float64)


def _kgrid(shape, h, dev):
    """vti_habc2.py:43-46: fftfreq grids indexed by grid position; 'k_x' runs along
    rows (dim -2), 'k_z' along columns -- replicated literally."""
    kx = torch.fft.fftfreq(shape[0], d=float(h), dtype=torch.float64, device=dev)
    kz = torch.fft.fftfreq(shape[1], d=float(h), dtype=torch.float64, device=dev)
    return torch.meshgrid(kx, kz, indexing="ij")


def _ddx(v, h):
    """centred first difference along the last dim, zero padding (convkernel.py:78)."""
    p = F.pad(v, (1, 1))
    return (p[..., 2:] - p[..., :-2]) / (2 * h)


def _ddz(v, h):
    """centred first difference along dim -2, zero padding (convkernel.py:79)."""
    p = F.pad(v, (0, 0, 1, 1))
    return (p[..., 2:, :] - p[..., :-2, :]) / (2 * h)


def wave2d_coefficients(equation, params, dt, h, d):
    """Returns (coef tensors fp32, coef slots) for the 2D second-order family.
    Slots: 0 r, 1 b, 2 cxx (ciso for ISO equations), 3 czz (alpha for PML), 4 cxz, 5 ax, 6 az, 7 m."""
    dt, h = float(dt), float(h)
    vp = _f64(params[0])
    r = vp * (dt / h)
    out = {0: r, 1: _f64(d)}
    if equation == "acoustic":
        bd = _f64(d) * dt
        out[2] = r * r / (1 + bd)              # ciso
        out[3] = (1 - bd) / (1 + bd)           # alpha
    elif equation == "acoustic_habc":
        out[2] = r * r
    elif equation == "acoustic_rho_habc":
        # acoustic_rho_habc.py:32-57:  K [bxp (E-C) - bxn (C-W) + bzp (S-C) - bzn (C-N)],  K = dt^2/h^2 rho vp^2,
        # b*p / b*n = mean of the buoyancy 1/rho with the next / previous cell (zero outside the grid)
        #   = cxx ((E-C)+(W-C)) + ax (E-W) + czz ((N-C)+(S-C)) + az (S-N),  cxx +- ax = K bxp / K bxn, czz +- az = K bzp / K bzn
        rho = _f64(params[1])
        K = r * r * rho
        bu = 1.0 / rho
        px = F.pad(bu, (1, 1))
        pz = F.pad(bu, (0, 0, 1, 1))
        bxp, bxn = (bu + px[..., 2:]) / 2, (px[..., :-2] + bu) / 2
        bzp, bzn = (bu + pz[..., 2:, :]) / 2, (pz[..., :-2, :] + bu) / 2
        out[2] = K * (bxp + bxn) / 2
        out[3] = K * (bzp + bzn) / 2
        out[5] = K * (bxp - bxn) / 2
        out[6] = K * (bzp - bzn) / 2
    elif equation == "acoustic_lsrtm_habc":    # acoustic_lsrtm_habc.py:10-30: A = vp^2 dt^2 Lap for both fields, + m A[h1]
        out[2] = r * r
        out[3] = r * r
        out[7] = _f64(params[1])
    elif equation in ("vti_habc2", "acoustic_vti_lsrtm_habc"):
        eps, delta = _f64(params[1]), _f64(params[2])
        kx, kz = _kgrid(vp.shape, h, vp.device)
        num = -2 * (eps - delta) * kx ** 2 * kz ** 2
        den = (1 + 2 * eps) * kx ** 4 + kz ** 4 + 2 * (1 + delta) * kx ** 2 * kz ** 2
        S = num / (den + 1e-26)
        r2 = r * r
        out[2] = r2 * ((1 + 2 * eps) + S)
        out[3] = r2 * (1 + S)
        if equation == "acoustic_vti_lsrtm_habc":
            out[7] = _f64(params[3])
    elif equation in ("tti_habc", "acoustic_tti_lsrtm_habc"):
        eps, delta, theta = _f64(params[1]), _f64(params[2]), _f64(params[3])
        th = torch.deg2rad(theta)
        s0, c0, s20 = torch.sin(th), torch.cos(th), torch.sin(2 * th)
        kx, kz = _kgrid(vp.shape, h, vp.device)
        a = kx * c0 - kz * s0
        b = kx * s0 + kz * c0
        num = -2 * (eps - delta) * a ** 2 * b ** 2
        den = (1 + 2 * eps) * a ** 4 + b ** 4 + 2 * (1 + delta) * a ** 2 * b ** 2
        S = num / (den + 1e-26)
        r2 = r * r
        out[2] = r2 * ((1 + 2 * eps) * c0 ** 2 + s0 ** 2 + S)
        out[3] = r2 * ((1 + 2 * eps) * s0 ** 2 + c0 ** 2 + S)
        # -2 eps vp^2 dt^2 sin(2 theta) * d2p/dxdz, d2/dxdz = 4-corner stencil / (4 h^2)
        out[4] = -2 * eps * r2 * s20 / 4
        if equation == "acoustic_tti_lsrtm_habc":
            out[7] = _f64(params[4])
    elif equation == "acoustic_fwim_habc":
        rx, rz = _f64(params[1]), _f64(params[2])
        out[2] = r * r
        v_x, v_z = _ddx(vp, h), _ddz(vp, h)
        # term2 - term3 with p_x = (E-W)/(2h):   coefficient of (E-W) and (S-N)
        out[5] = (vp * dt ** 2 * v_x - 2 * vp ** 2 * dt ** 2 * rx) / (2 * h)
        out[6] = (vp * dt ** 2 * v_z - 2 * vp ** 2 * dt ** 2 * rz) / (2 * h)
    else:
        raise ValueError(f"seistorch_b200: equation '{equation}' is not a 2D second-order equation")
    slots = tuple(sorted(out))
    return [out[s].to(torch.float32) for s in slots], slots


def elastic_coefficients(params, dt, h, d):
    """(ca, cl2m, cl, cm, cb) of include/seistorch_b200.h from (vp, vs, rho)."""
    dt, h = float(dt), float(h)
    vp, vs, rho = (_f64(p) for p in params)
    lam = rho * (vp ** 2 - 2 * vs ** 2)
    mu = rho * vs ** 2
    c = 0.5 * dt * _f64(d)
    ic = 1.0 / (1.0 + c)
    s = dt / h
    coefs = [(1 - c) * ic, (lam + 2 * mu) * s * ic, lam * s * ic, mu * s * ic, (s / rho) * ic]
    return [x.to(torch.float32) for x in coefs]


def acoustic3d_coefficients(params, dt, h, d):
    """(ciso, alpha) of the 3D damped acoustic update  y = h1 + alpha (h1-h2) + ciso lap7(h1)
    (equations3d/acoustic.py:72-85 in increment form)."""
    dt, h = float(dt), float(h)
    vp = _f64(params[0])
    bd = _f64(d) * dt
    r = vp * (dt / h)
    return [(r * r / (1 + bd)).to(torch.float32), ((1 - bd) / (1 + bd)).to(torch.float32)]
