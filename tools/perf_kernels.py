"""Kernel-level timing helper (CUDA events, forward-only and forward+adjoint) used while
tuning; prints one line per configuration.  Not part of the product path.
    python tools/perf_kernels.py acoustic_habc 751 2301 8 200
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import seistorch_b200 as sb  # noqa: E402
from oracle import cases  # noqa: E402


def run(eq, nz, nx, B, nt, ny=None, grad=True):
    case = cases.make_case(eq, nz=nz, nx=nx, nshots=B, nt=nt, rec_step=2, ny=ny)
    cfg, model = sb.model_from_case(case, device="cuda", mode="inversion")
    x = torch.as_tensor(np.asarray(case["wavelet"]), device="cuda").unsqueeze(0)
    npts = int(np.prod(model.cell.geom.domain_shape))
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    with torch.no_grad():
        model(x)
        torch.cuda.synchronize()
        ev[0].record()
        model(x)
        ev[1].record()
    torch.cuda.synchronize()
    t_f = ev[0].elapsed_time(ev[1]) / nt
    msg = f"{eq} grid={model.cell.geom.domain_shape} B={B} nt={nt}: fwd {t_f*1e3:.1f} us/step {B*npts/t_f/1e6:.1f} Gpts/s"
    if grad:
        t_a = 1e30
        for _ in range(3):
            syn = model(x)
            loss = sum((s ** 2).sum() for s in syn)
            torch.cuda.synchronize()
            ev[2].record()
            loss.backward()
            ev[3].record()
            torch.cuda.synchronize()
            t_a = min(t_a, ev[2].elapsed_time(ev[3]) / nt)
        msg += f" | adj {t_a*1e3:.1f} us/step {B*npts/t_a/1e6:.1f} Gpts/s"
    print(msg, flush=True)


if __name__ == "__main__":
    a = sys.argv[1:]
    if a:
        run(a[0], int(a[1]), int(a[2]), int(a[3]), int(a[4]), ny=int(a[5]) if len(a) > 5 else None)
    else:
        run("acoustic", 751, 2301, 8, 200)
        run("acoustic_habc", 751, 2301, 8, 200)
        run("acoustic", 150, 300, 1, 2000)
        run("elastic", 400, 1000, 8, 200)
        run("vti_habc2", 500, 1200, 8, 100)
        run("tti_habc", 500, 1200, 8, 100)
        run("acoustic_lsrtm_habc", 500, 1200, 8, 100)
        run("acoustic_rho_habc", 500, 1200, 8, 100)
        run("acoustic_vti_lsrtm_habc", 500, 1200, 8, 100)
        run("acoustic_tti_lsrtm_habc", 500, 1200, 8, 100)
        run("acoustic_fwim_habc", 500, 1200, 8, 100)
        run("acoustic", 100, 200, 2, 50, ny=200)
