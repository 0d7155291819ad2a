"""Record filter (seistorch/signal.py:49-101, backend='torch'): the oracle restatement (CPU) and the CUDA kernel
(-m gpu) against golden vectors generated from the REAL reference (oracle/make_misfit_golden.py filter): filtered
records and the gradient of a linear functional of them, low-pass and band-pass, ragged shots."""
import os

import numpy as np
import pytest
import torch

from conftest import rel

GOLD = os.path.join(os.path.dirname(__file__), "golden", "filter.npz")


@pytest.mark.parametrize("tag", ["low", "band"])
def test_oracle_filter_matches_reference(tag):
    from oracle import sigproc
    z = np.load(GOLD)
    b, a = sigproc.butter(int(z["order"]), z[f"{tag}_freqs"].tolist(), float(z["dt"]))
    for k in range(2):
        y = sigproc.filtfilt(z[f"x_{k}"], b, a)
        assert rel(y, z[f"{tag}_y_{k}"]) < 1e-6
        # self-adjoint operator: gradient of <w, F x> w.r.t. x is F w
        assert rel(sigproc.filtfilt(z[f"w_{k}"], b, a), z[f"{tag}_grad_{k}"]) < 1e-6


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["low", "band"])
def test_cuda_filter_matches_reference(tag):
    from seistorch_b200.signal import SeisSignal
    from seistorch_b200.type import TensorList
    z = np.load(GOLD)
    sig = SeisSignal({"geom": {"dt": float(z["dt"])}, "training": {"filter_ord": int(z["order"])}})
    xs = [torch.from_numpy(z[f"x_{k}"]).cuda().requires_grad_(True) for k in range(2)]
    out = sig.filter(TensorList([v * 1.0 for v in xs]), z[f"{tag}_freqs"].tolist(), backend="torch")
    loss = sum((o * torch.from_numpy(z[f"w_{k}"]).cuda()).sum() for k, o in enumerate(out.data))
    loss.backward()
    for k in range(2):
        assert rel(out.data[k].detach().cpu().numpy(), z[f"{tag}_y_{k}"]) < 1e-6
        assert rel(xs[k].grad.cpu().numpy(), z[f"{tag}_grad_{k}"]) < 1e-6
    assert sig.filter(out, "all") is out
