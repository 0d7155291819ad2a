"""TEST INFRASTRUCTURE ONLY.  CPU driver around tests/hostcheck/libhostcheck.so: it runs
the very same per-cell functions the sm_100a kernels call (csrc/*_math.cuh) in a plain
time loop, so forward/adjoint algebra can be compared with the oracle without a GPU."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(HERE, "hostcheck", "hostcheck.cpp")
LIB = os.path.join(HERE, "hostcheck", "libhostcheck.so")


class HcW2(C.Structure):
    _fields_ = [("flags", C.c_int), ("B", C.c_int), ("nz", C.c_int), ("nx", C.c_int), ("ld", C.c_int),
                ("bw", C.c_int), ("multiple", C.c_int), ("dt", C.c_float), ("coef", C.c_void_p * 8)]


class HcE2(C.Structure):
    _fields_ = [("B", C.c_int), ("nz", C.c_int), ("nx", C.c_int), ("ld", C.c_int), ("coef", C.c_void_p * 5)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        deps = [SRC] + [os.path.join(ROOT, "seistorch_b200", "csrc", f) for f in
                        ("st_wave2d_math.cuh", "st_elastic2d.cuh", "st_common.cuh")]
        if not os.path.exists(LIB) or any(os.path.getmtime(d) > os.path.getmtime(LIB) for d in deps):
            subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-Wno-unknown-pragmas",
                                   "-I", os.path.join(ROOT, "seistorch_b200", "csrc"), SRC, "-o", LIB])
        _lib = C.CDLL(LIB)
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def run_case(case, want_grad=True, grad_rec=None):
    """Forward (+ adjoint with d loss/d rec = grad_rec or 2*rec) of a 2D case through the
    host emulation.  Returns records list and {param: grad} (numpy)."""
    from oracle import loop
    from seistorch_b200 import coefficients as cf
    from seistorch_b200.eqconfigure import field_channels

    L = lib()
    eq = case["equation"]
    multiple = bool(case.get("multiple", False))
    names, params, d, src, bidx, rec, counts = loop.build_geometry(case, torch.float32)
    params = [p.clone().requires_grad_(True) for p in params]
    nz, nx = params[0].shape
    B, nt, dt, h = len(case["sources"]), int(case["nt"]), float(case["dt"]), float(case["h"])
    ld = nx
    chan = field_channels(eq)
    elastic = eq == "elastic"
    if elastic:
        coefs = cf.elastic_coefficients(params, dt, h, d)
        slots = tuple(range(5))
        nf = 5
    else:
        _, flags = cf.EQUATIONS[eq]
        coefs, slots = cf.wave2d_coefficients(eq, params, dt, h, d)
        nf = 2 if flags & 32 else 1
    cnp = [np.ascontiguousarray(c.detach().numpy(), dtype=np.float32) for c in coefs]
    if elastic:
        P = HcE2(B, nz, nx, ld)
        for k in range(5):
            P.coef[k] = _p(cnp[k])
    else:
        P = HcW2(flags, B, nz, nx, ld, 50, int(multiple), dt)
        for k, s in enumerate(slots):
            P.coef[s] = _p(cnp[k])
    w = np.asarray(case["wavelet"], np.float32)
    fs = nz * nx
    smask = [chan[s] for s in case["source_type"]]
    rchan = [chan[r] for r in case["receiver_type"]]
    R = len(bidx)
    S = np.zeros((nt + 2, nf, B, nz, nx), np.float32)   # S[j+2] = state after step j
    Spre = np.zeros((nf, B, nz, nx), np.float32)
    pre_hist = []
    recs = np.zeros((nt, R, len(rchan)), np.float32)
    for i in range(nt):
        if elastic:
            L.hc_elastic2d_forward(C.byref(P), _p(S[i + 1]), _p(S[i + 2]))
            if want_grad:
                pre_hist.append(S[i + 2].copy())
        else:
            L.hc_wave2d_forward(C.byref(P), _p(S[i]), _p(S[i + 1]), _p(S[i + 2]))
        for b in range(B):
            for f in smask:
                S[i + 2, f, b, src[b, 1], src[b, 0]] += w[i]
        for c, f in enumerate(rchan):
            recs[i, :, c] = S[i + 2, f, bidx, rec[1], rec[0]]
    out_recs = np.split(recs, np.cumsum(counts)[:-1], axis=1)
    if not want_grad:
        return out_recs, {}
    gr = 2.0 * recs if grad_rec is None else grad_rec
    ngr = 4 if elastic else 7
    gacc = np.zeros((ngr, nz, nx), np.float32)
    lam = np.zeros((nt + 3, nf, B, nz, nx), np.float32)     # lam[i] = Lam_i ; lam[nt], lam[nt+1] = 0
    zero = np.zeros((nf, B, nz, nx), np.float32)
    for i in range(nt - 1, -1, -1):
        if elastic:
            if i == nt - 1:
                lam[i][:] = 0
            else:
                L.hc_elastic2d_adjoint(C.byref(P), _p(lam[i + 1]), _p(S[i + 2]), _p(pre_hist[i + 1]), _p(lam[i]), _p(gacc))
        else:
            L.hc_wave2d_adjoint(C.byref(P), _p(lam[i + 1]), _p(lam[i + 2]), _p(S[i + 2]), _p(S[i + 1]), _p(lam[i]), _p(gacc))
        for c, f in enumerate(rchan):
            np.add.at(lam[i][f], (bidx, rec[1], rec[0]), gr[i, :, c])
    # chain to the parameters through the coefficient maps
    if elastic:
        gmap = {1: 0, 2: 1, 3: 2, 4: 3}
        pairs = [(coefs[k], torch.from_numpy(gacc[g])) for k, g in gmap.items()]
    else:
        g_of = {0: 0, 2: 1, 3: 2, 4: 3, 5: 4, 6: 5, 7: 6}
        pairs = [(coefs[k], torch.from_numpy(gacc[g_of[s]])) for k, s in enumerate(slots) if s in g_of]
    pairs = [(c, g) for c, g in pairs if c.requires_grad]
    torch.autograd.backward([c for c, _ in pairs], [g for _, g in pairs])
    grads = {n: (p.grad.numpy() if p.grad is not None else None) for n, p in zip(names, params)}
    return out_recs, grads
