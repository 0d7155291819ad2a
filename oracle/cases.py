"""TEST INFRASTRUCTURE ONLY.  Seeded synthetic cases shared by the golden
generator, the parity tests and bench.py (seed 20230503, dt=1e-3,
h=10, Ricker fm=10 Hz, boundary width 50).  A case is a plain dict; see
oracle/ref_runner.run_reference for the schema.
"""
from __future__ import annotations

import numpy as np

from .loop import ricker_wave

SEED = 20230503


def _smooth(a, n=3):
    for _ in range(n):
        p = np.pad(a, 1, mode="edge")
        a = (p[1:-1, 1:-1] * 4 + p[:-2, 1:-1] + p[2:, 1:-1] + p[1:-1, :-2] + p[1:-1, 2:]) / 8.0
    return a


def vp_model(nz, nx, rng, vmin=1500.0, vmax=3000.0, noise=150.0):
    z = np.linspace(0.0, 1.0, nz)[:, None]
    vp = vmin + (vmax - vmin) * z + _smooth(rng.standard_normal((nz, nx)), 4) * noise * 3
    return np.clip(vp, vmin, vmax + 300).astype(np.float32)


def make_case(equation, nz=40, nx=60, nshots=2, nt=200, rec_step=3, dt=1e-3, h=10.0,
              fm=10.0, delay=60, multiple=False, seed=SEED, ny=None):
    """Small/medium parity case for one equation family."""
    rng = np.random.default_rng(seed)
    if ny is not None:  # 3D: model file (nx, nz, ny) -> tensor layout (x, z, y)
        vp = np.full((nx, nz, ny), 1500.0, np.float32)
        vp[:, nz // 2:, :] = 2000.0
        vp += (rng.standard_normal(vp.shape) * 20).astype(np.float32)
        models = {"vp": vp}
        sx = np.linspace(4, nx - 5, nshots)
        sources = [[float(x) + 0.3, float(ny // 2), 1.0] for x in sx]
        rx, ry = np.meshgrid(np.arange(1, nx, rec_step), np.arange(1, ny, rec_step), indexing="ij")
        recs = [[rx.ravel().tolist(), ry.ravel().tolist(), [2] * rx.size]] * nshots
        boundary, st, rt = "pml", ["h1"], ["h1"]
        inv = {"vp": True}
    else:
        vp = vp_model(nz, nx, rng)
        vpn = (vp - vp.min()) / (vp.max() - vp.min())
        sx = np.linspace(4, nx - 5, nshots)
        sources = [[float(x) + 0.7, 1.2] for x in sx]
        rxs = list(range(1, nx, rec_step))
        recs = [[rxs, [2] * len(rxs)] for _ in range(nshots)]
        if equation == "acoustic":
            models, boundary, st, rt, inv = {"vp": vp}, "pml", ["h1"], ["h1"], {"vp": True}
        elif equation == "acoustic_habc":
            models, boundary, st, rt, inv = {"vp": vp}, "habc", ["h1"], ["h1"], {"vp": True}
        elif equation == "elastic":
            models = {"vp": vp, "vs": (vp / 1.73).astype(np.float32),
                      "rho": (2000.0 + 200 * vpn).astype(np.float32)}
            boundary, st, rt = "pml", ["vz"], ["vx", "vz"]
            inv = {"vp": True, "vs": True, "rho": True}
        elif equation in ("vti_habc2", "tti_habc"):
            models = {"vp": vp, "epsilon": (0.1 * vpn + 0.02).astype(np.float32),
                      "delta": (0.05 * vpn + 0.01).astype(np.float32)}
            inv = {"vp": True, "epsilon": True, "delta": True}
            if equation == "tti_habc":
                models["theta"] = (15.0 + 10 * vpn).astype(np.float32)
                inv["theta"] = True
            boundary, st, rt = "habc", ["p1"], ["p1"]
        elif equation == "acoustic_rho_habc":
            models = {"vp": vp, "rho": (2000.0 + 400 * vpn).astype(np.float32)}
            boundary, st, rt, inv = "habc", ["h1"], ["h1"], {"vp": True, "rho": True}
        elif equation == "acoustic_lsrtm_habc":
            m = np.zeros_like(vp)
            m[1:] = (vp[1:] - vp[:-1]) / vp[1:] * 5
            models, inv = {"vp": vp, "m": m.astype(np.float32)}, {"vp": True, "m": True}
            boundary, st, rt = "habc", ["h1"], ["sh1"]
        elif equation in ("acoustic_vti_lsrtm_habc", "acoustic_tti_lsrtm_habc"):
            m = np.zeros_like(vp)
            m[1:] = (vp[1:] - vp[:-1]) / vp[1:] * 5
            models = {"vp": vp, "epsilon": (0.1 * vpn + 0.02).astype(np.float32),
                      "delta": (0.05 * vpn + 0.01).astype(np.float32), "m": m.astype(np.float32)}
            inv = {"vp": True, "epsilon": True, "delta": True, "m": True}
            if "tti" in equation:
                models["theta"] = (15.0 + 10 * vpn).astype(np.float32)
                inv["theta"] = True
            boundary, st, rt = "habc", ["p1"], ["sp1"]
        elif equation == "acoustic_fwim_habc":
            rz = np.zeros_like(vp)
            rz[1:] = (vp[1:] - vp[:-1]) / vp[1:] * 0.02
            rx_ = np.zeros_like(vp)
            rx_[:, 1:] = (vp[:, 1:] - vp[:, :-1]) / vp[:, 1:] * 0.02
            models = {"vp": vp, "rx": rx_.astype(np.float32), "rz": rz.astype(np.float32)}
            boundary, st, rt = "habc", ["h1"], ["h1"]
            inv = {"vp": True, "rx": True, "rz": True}
        else:
            raise ValueError(equation)
    return dict(equation=equation, models=models, invlist=inv, sources=sources, receivers=recs,
                nt=nt, dt=dt, h=h, wavelet=ricker_wave(fm, dt, nt, delay), source_type=st,
                receiver_type=rt, boundary=boundary, multiple=multiple)
