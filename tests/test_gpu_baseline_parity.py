"""GPU (-m gpu): the CUDA path against the float64 oracle AT THE BASELINE CONFIGURATIONS (grid, acquisition and batch
sizes of bench.py's workloads -- the code path that earns the headline: full tile columns, full 8-shot tile groups, surface
acquisition across > 1000 receivers), on horizons the CPU oracle finishes in seconds.  Tolerances are the north star's:
seismograms <= 1e-5, gradients <= 1e-4 relative L2 against the reference's float64 arithmetic.

The oracle (oracle/loop.py) is pinned to the real reference by tests/golden (tests/test_oracle_golden.py)."""
import numpy as np
import pytest
import torch

from conftest import rel

pytestmark = pytest.mark.gpu
_CACHE = {}


def _run(case, mode="inversion", loss_shots=None, tma=None, monkeypatch=None):
    import seistorch_b200 as sb
    from seistorch_b200 import engine
    if monkeypatch is not None:
        if tma is None:
            monkeypatch.delenv("SEISTORCH_B200_TMA", raising=False)
        else:
            monkeypatch.setenv("SEISTORCH_B200_TMA", tma)
    cfg, model = sb.model_from_case(case, device="cuda", mode=mode)
    x = torch.as_tensor(np.asarray(case["wavelet"]), dtype=torch.float32, device="cuda").unsqueeze(0)
    if mode == "forward":
        with torch.no_grad():
            syn = model(x)
        return [s.cpu().numpy() for s in syn], None, dict(engine.KERNELS)
    syn = model(x)
    shots = range(len(syn)) if loss_shots is None else loss_shots
    sum((syn[k].double() ** 2).sum() for k in shots).backward()
    inv = [k for k, v in case["invlist"].items() if v]
    grads = {k: getattr(model.cell.geom, k).grad.cpu().numpy() for k in inv}
    return [s.detach().cpu().numpy() for s in syn], grads, dict(engine.KERNELS)


def _oracle(case, grad=True, dtype=torch.float64):
    from oracle import loop, misfit
    inv = [k for k, v in case["invlist"].items() if v] if grad else []
    if grad:
        recs, params = loop.simulate(case, dtype=dtype, requires_grad=inv)
        misfit.l2(recs, [torch.zeros_like(r) for r in recs]).backward()
        return [r.detach().numpy() for r in recs], {k: params[k].grad.numpy() for k in inv}
    with torch.no_grad():
        recs, _ = loop.simulate(case, dtype=dtype)
    return [r.numpy() for r in recs], None


def test_cfg1_full_three_number_protocol():
    """BASELINE configs[0] exactly: acoustic, PML, 300x150 layered model (padded 400x250), one shot, nt = 2000, forward
    modelling.  SURVEY 8c protocol: err(new32, ref64) <= 1e-5 and <= err(ref32, ref64); the three numbers are printed."""
    import bench
    true, _ = bench.WORKLOADS["cfg1"]["models"]()
    case = bench.make_case(1, workload="cfg1", models=true)
    assert case["nt"] == 2000 and len(case["sources"]) == 1
    new32, _, kern = _run(case, mode="forward")
    ref64, _ = _oracle(case, grad=False)
    ref32, _ = _oracle(case, grad=False, dtype=torch.float32)
    e_new, e_ref, e_nn = rel(new32[0], ref64[0]), rel(ref32[0], ref64[0]), rel(new32[0], ref32[0])
    print(f"cfg1 [{kern['forward']}]: err(new32,ref64)={e_new:.2e} err(ref32,ref64)={e_ref:.2e} err(new32,ref32)={e_nn:.2e}")
    assert e_new < 1e-5 and e_new <= e_ref


def test_cfg1_gradient_full_horizon():
    """cfg1 grid and horizon with a gradient (the FWI use of the same configuration), against float64 AD."""
    import bench
    true, _ = bench.WORKLOADS["cfg1"]["models"]()
    case = bench.make_case(1, workload="cfg1", models={"vp": (true["vp"] * 0.97).astype(np.float32)}, nt=600)
    case["invlist"] = {"vp": True}
    rec, g, _ = _run(case)
    orec, og = _oracle(case)
    assert rel(rec[0], orec[0]) < 1e-5
    assert rel(g["vp"], og["vp"]) < 1e-4


@pytest.mark.parametrize("B", [8, 16])
def test_cfg2_grid_full_shot_groups_against_oracle(B, monkeypatch):
    """BASELINE configs[1] grid (2301x751 -> padded 2401x851: 19 tile columns), bench geometry (sources at depth 1, 1151
    surface receivers per shot), B = 8 (one full tile group) and B = 16 (two), through the TMA kernels.  The oracle runs the
    first and the last shot of the first tile group forward in float64, and shot 0 with AD; the CUDA gradient is that of the
    misfit of shot 0's records, taken inside the full batch."""
    import bench
    nt = 100
    _true, init = bench.make_models()
    case = bench.make_case(B, total_shots=16, vp=init, nt=nt, delay=25)
    rec, g, kern = _run(case, loss_shots=[0], monkeypatch=monkeypatch)
    assert "tma" in kern["forward"] and "tma" in kern["adjoint"], kern
    pick = [0, 7]
    if "cfg2" not in _CACHE:              # the same two shots for B = 8 and B = 16: run the CPU oracle once
        sub = dict(case, sources=[case["sources"][k] for k in pick], receivers=[case["receivers"][k] for k in pick])
        one = dict(case, sources=case["sources"][:1], receivers=case["receivers"][:1])
        _CACHE["cfg2"] = (_oracle(sub, grad=False)[0], _oracle(one))
    orec, (orec0, og) = _CACHE["cfg2"]
    for k, o in zip(pick, orec):
        assert rel(rec[k], o) < 1e-5, (B, k)
    assert rel(rec[0], orec0[0]) < 1e-5
    assert rel(g["vp"], og["vp"]) < 1e-4
    # the other shots of the batch are the same computation: bit-equal to a B = 2 run of shots (0, B-1) on the same kernels
    pair = dict(case, sources=[case["sources"][0], case["sources"][B - 1]], receivers=[case["receivers"][0], case["receivers"][B - 1]])
    monkeypatch.setenv("SEISTORCH_B200_TMA", "1")
    rec2, _, _ = _run(pair, mode="forward")
    recB, _, _ = _run(case, mode="forward")
    assert np.array_equal(recB[0], rec2[0]) and np.array_equal(recB[B - 1], rec2[1])


def test_cfg3_grid_elastic_against_oracle():
    """BASELINE configs[2] grid: elastic, PML, 1000x400 (padded 1100x500), vp/vs/rho inverted, bench geometry, 2 shots."""
    import bench
    _true, init = bench.WORKLOADS["cfg3"]["models"]()
    case = bench.make_case(2, total_shots=8, workload="cfg3", models=init, nt=60, delay=20)
    rec, g, _ = _run(case, loss_shots=[0])
    one = dict(case, sources=case["sources"][:1], receivers=case["receivers"][:1])
    orec, og = _oracle(one)
    assert rel(rec[0], orec[0]) < 1e-5
    for k in ("vp", "vs", "rho"):
        assert rel(g[k], og[k]) < 1e-4, k


@pytest.mark.parametrize("wl", ["cfg4", "cfg4_tti", "cfg4_fwim"])
def test_cfg4_grid_against_oracle(wl):
    """BASELINE configs[3] grid: VTI / TTI LSRTM Born pair and the joint FWI-LSRTM equation, 1200x500 (padded 1300x600)."""
    import bench
    true, _init = bench.WORKLOADS[wl]["models"]()          # the true model has a non-zero reflectivity m / rx, rz
    case = bench.make_case(2, total_shots=12, workload=wl, models=true, nt=60, delay=20)
    rec, g, _ = _run(case, loss_shots=[0])
    one = dict(case, sources=case["sources"][:1], receivers=case["receivers"][:1])
    orec, og = _oracle(one)
    assert rel(rec[0], orec[0]) < 1e-5
    for k in g:
        assert rel(g[k], og[k]) < 1e-4, k


def test_cfg5_mid_size_3d_against_oracle():
    """BASELINE configs[4] at a size the float64 AD oracle can hold: model file (44, 30, 40) -> padded (144, 130, 140) =
    2.6 M cells (>= 128^3), 2 shots of the bench's acquisition pattern, K-step checkpoints forced."""
    import bench
    import seistorch_b200 as sb
    _true, init = bench._models_cfg5((44, 30, 40))
    case = bench.make_case(2, total_shots=2, workload="cfg5", models=init, nt=36, delay=12)
    cfg, model = sb.model_from_case(case, device="cuda", mode="inversion")
    model.segment = 11
    x = torch.as_tensor(np.asarray(case["wavelet"]), dtype=torch.float32, device="cuda").unsqueeze(0)
    syn = model(x)
    sum((s.double() ** 2).sum() for s in syn).backward()
    orec, og = _oracle(case)
    for a, b in zip(syn, orec):
        assert rel(a.detach().cpu().numpy(), b) < 1e-5
    assert rel(model.cell.geom.vp.grad.cpu().numpy(), og["vp"]) < 1e-4
