"""Gradient smoothing (seistorch/process.py:66-112 -> signal.py:247-319): the oracle restatement (CPU) and the CUDA
kernel (-m gpu) against golden vectors produced by the REAL reference's gaussian_filter (oracle/make_smooth_golden.py)."""
import os
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from conftest import rel

GOLD = os.path.join(os.path.dirname(__file__), "golden", "smooth.npz")


def _cfg(z, k):
    c = z[f"cfg_{k}"]
    return int(c[0]), {"z": float(c[1]), "x": float(c[2])}, {"z": int(c[3]), "x": int(c[4])}


@pytest.mark.parametrize("k", [0, 1])
def test_oracle_smoothing_matches_reference(k):
    from oracle import postproc
    z = np.load(GOLD)
    counts, sigma, radius = _cfg(z, k)
    assert rel(postproc.smooth_gradient(z["g"], counts, sigma, radius), z[f"y_{k}"]) < 1e-6
    with pytest.raises(ValueError):
        postproc.gaussian_filter(z["g"], 1.0, 3, 0)


@pytest.mark.gpu
@pytest.mark.parametrize("k", [0, 1])
def test_cuda_smoothing_matches_reference(k):
    """PostProcess.smooth_gradient on a stand-in model: the parameter gradient is smoothed in place, on the device."""
    from seistorch_b200.process import PostProcess, gaussian_filter
    z = np.load(GOLD)
    counts, sigma, radius = _cfg(z, k)
    par = torch.nn.Parameter(torch.zeros(z["g"].shape, device="cuda"))
    par.grad = torch.from_numpy(z["g"]).cuda()
    frozen = torch.nn.Parameter(torch.zeros(3, device="cuda"), requires_grad=False)
    model = SimpleNamespace(cell=SimpleNamespace(geom=SimpleNamespace(ndim=2)), parameters=lambda: [par, frozen])
    cfg = {"training": {"smooth": {"counts": counts, "sigma": sigma, "radius": radius}}, "geom": {"boundary": {"width": 50}, "multiple": False}}
    PostProcess(model, cfg, SimpleNamespace(grad_cut=False)).smooth_gradient()
    assert par.grad.is_cuda and rel(par.grad.cpu().numpy(), z[f"y_{k}"]) < 1e-6
    with pytest.raises(ValueError):
        gaussian_filter(par.grad, 1.0, 3, 0)
    with pytest.raises(RuntimeError):
        gaussian_filter(torch.zeros(8, 8, device="cuda"), 1.0, 8, 0)        # radius must be smaller than the axis
