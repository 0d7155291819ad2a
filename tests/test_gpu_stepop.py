"""GPU (-m gpu): the per-step plug-in surface `_time_step(*params, *fields, dt, h, d, habcs=)`
of every equation module against the oracle's step function (values and VJPs)."""
import numpy as np
import pytest
import torch

from conftest import rel

pytestmark = pytest.mark.gpu

EQS = ["acoustic", "acoustic_habc", "elastic", "vti_habc2", "tti_habc", "acoustic_lsrtm_habc", "acoustic_rho_habc", "acoustic_vti_lsrtm_habc",
       "acoustic_tti_lsrtm_habc", "acoustic_fwim_habc"]


@pytest.mark.parametrize("eq", EQS)
def test_single_step_value_and_vjp(eq):
    import importlib
    from oracle import cases, equations, loop
    from seistorch_b200.habc import bound_mask
    case = cases.make_case(eq, nz=21, nx=33, nshots=2, nt=4)
    names, params, d, *_ = loop.build_geometry(case, torch.float64)
    shape = tuple(params[0].shape)
    g = torch.Generator().manual_seed(3)
    nf = len(equations.WAVEFIELDS[eq])
    fields = [torch.randn((2,) + shape, generator=g, dtype=torch.float64) for _ in range(nf)]
    gouts = [torch.randn((2,) + shape, generator=g, dtype=torch.float64) for _ in range(nf)]
    dt, h = torch.tensor(1e-3, dtype=torch.float64), torch.tensor(10.0, dtype=torch.float64)
    P = [p.clone().requires_grad_(True) for p in params]
    F = [f.clone().requires_grad_(True) for f in fields]
    outs = equations.get_step(eq)(P, F, dt, h, d)
    torch.autograd.backward(list(outs), gouts)
    mod = importlib.import_module(f"seistorch_b200.equations2d.{eq}")
    Pc = [p.detach().float().cuda().requires_grad_(True) for p in params]
    Fc = [f.detach().float().cuda().requires_grad_(True) for f in fields]
    habcs = bound_mask(*shape, 50, "cuda", 2, return_idx=True) if "habc" in eq else None
    outs_c = mod._time_step(*Pc, *Fc, torch.tensor(1e-3).cuda(), torch.tensor(10.0).cuda(), d.float().cuda(), habcs=habcs)
    assert len(outs_c) == len(outs)
    for a, b in zip(outs_c, outs):
        assert rel(a.detach().cpu().numpy(), b.detach().numpy()) < 2e-6
    torch.autograd.backward(list(outs_c), [g_.float().cuda() for g_ in gouts])
    for a, b, n in zip(Fc, F, equations.WAVEFIELDS[eq]):
        assert rel(a.grad.cpu().numpy(), b.grad.numpy()) < 1e-5, n
    for a, b, n in zip(Pc, P, names):
        assert rel(a.grad.cpu().numpy(), b.grad.numpy()) < 1e-4, n
