"""Trace-domain misfits: the oracle restatements (CPU) and the CUDA kernels (-m gpu) against golden vectors
generated from the REAL reference (oracle/make_misfit_golden.py: seistorch/loss.py L2, L1, CosineSimilarity,
Envelope) -- loss values and adjoint sources, shots with different receiver counts."""
import os

import numpy as np
import pytest
import torch

from conftest import rel

GOLD = os.path.join(os.path.dirname(__file__), "golden", "misfits.npz")
NAMES = ["l2", "l1", "sml1", "cs", "cc", "integration", "nim", "w1d", "envelope"]


def _load():
    z = np.load(GOLD)
    n = len([k for k in z.files if k.startswith("syn_")])
    return z, [z[f"syn_{k}"] for k in range(n)], [z[f"obs_{k}"] for k in range(n)]


@pytest.mark.parametrize("name", NAMES)
def test_oracle_misfit_matches_reference(name):
    from oracle import misfit
    fn = {"l2": misfit.l2, "l1": misfit.l1, "sml1": misfit.sml1, "cc": misfit.cc, "integration": misfit.integration, "cs": misfit.cs, "nim": misfit.nim, "w1d": misfit.w1d, "envelope": misfit.envelope_loss}[name]
    z, syn, obs = _load()
    xs = [torch.from_numpy(x).double().requires_grad_(True) for x in syn]
    loss = fn(xs, [torch.from_numpy(y).double() for y in obs])
    loss.backward()
    assert abs(float(loss) - float(z[f"{name}_loss"])) <= 1e-12 * abs(float(z[f"{name}_loss"]))
    for k, x in enumerate(xs):
        assert rel(x.grad.numpy(), z[f"{name}_grad_{k}"]) < 1e-12


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_cuda_misfit_matches_reference(name):
    import seistorch_b200 as sb
    z, syn, obs = _load()
    xs = [torch.from_numpy(x).cuda().requires_grad_(True) for x in syn]
    loss = sb.Loss(name).loss(None)(xs, [torch.from_numpy(y).cuda() for y in obs])
    loss.backward()
    assert abs(float(loss) - float(z[f"{name}_loss"])) <= 2e-5 * abs(float(z[f"{name}_loss"]))
    for k, x in enumerate(xs):
        assert rel(x.grad.cpu().numpy(), z[f"{name}_grad_{k}"]) < 2e-5
    # stacked form (equal receiver counts): one launch over all shots == the per-shot sum
    xs2 = torch.from_numpy(np.stack([syn[0], syn[0][::-1].copy()])).cuda().requires_grad_(True)
    ys2 = torch.from_numpy(np.stack([obs[0], obs[0][::-1].copy()])).cuda()
    l_stack = sb.Loss(name).loss(None)(xs2, ys2)
    l_list = sb.Loss(name).loss(None)([xs2[0], xs2[1]], [ys2[0], ys2[1]])
    assert abs(float(l_stack) - float(l_list)) <= 2e-5 * abs(float(l_list))


def test_oracle_traveltime_matches_reference():
    from oracle import misfit
    z = np.load(GOLD)
    xs = torch.from_numpy(z["tt_syn"]).double().requires_grad_(True)
    loss = misfit.traveltime(list(xs), list(torch.from_numpy(z["tt_obs"]).double()))
    loss.backward()
    assert abs(float(loss) - float(z["tt_loss"])) <= 1e-10 * abs(float(z["tt_loss"]))
    assert rel(xs.grad.numpy(), z["tt_grad"]) < 1e-9


@pytest.mark.gpu
def test_cuda_traveltime_matches_reference():
    import seistorch_b200 as sb
    z = np.load(GOLD)
    xs = torch.from_numpy(z["tt_syn"]).cuda().requires_grad_(True)
    loss = sb.Loss("traveltime").loss(None)(xs, torch.from_numpy(z["tt_obs"]).cuda())
    loss.backward()
    assert abs(float(loss) - float(z["tt_loss"])) <= 2e-5 * abs(float(z["tt_loss"]))
    assert rel(xs.grad.cpu().numpy(), z["tt_grad"]) < 2e-5
