"""GPU (-m gpu): the persistent multi-timestep kernel (st_wave2d_persist.cu: one launch = the whole time loop of
rnn.py:178-205 for small acoustic-PML grids) against the per-step kernels -- bit for bit, the per-cell arithmetic
is the same expression -- and against the float64 oracle."""
import numpy as np
import pytest
import torch

from conftest import cat_records, rel

pytestmark = pytest.mark.gpu


def _run(case, monkeypatch, persist, mode="inversion", segment=None, variant=None, encoding=False, wav=None):
    import seistorch_b200 as sb
    from seistorch_b200 import engine
    monkeypatch.setenv("SEISTORCH_B200_PERSIST", "1" if persist else "0")
    if variant is None:
        monkeypatch.delenv("SEISTORCH_B200_PERSIST_VARIANT", raising=False)
    else:
        monkeypatch.setenv("SEISTORCH_B200_PERSIST_VARIANT", str(variant))
    cfg, model = sb.model_from_case(case, device="cuda", mode=mode, source_encoding=encoding)
    if encoding:
        model.reset_probes(model.probes[0])
    model.segment = segment
    w = np.asarray(case["wavelet"]) if wav is None else wav
    x = torch.as_tensor(w, dtype=torch.float32, device="cuda")
    if x.ndim == 1:
        x = x.unsqueeze(0)
    l0 = dict(engine.LAUNCHES)
    if mode == "forward":
        with torch.no_grad():
            syn = model(x)
        recs, g = [s.cpu().numpy() for s in syn], None
    else:
        x.requires_grad_(True)
        syn = model(x)
        sum((s ** 2).sum() for s in syn).backward()
        recs, g = [s.detach().cpu().numpy() for s in syn], model.cell.geom.vp.grad.cpu().numpy()
        _run.wavelet_grad = x.grad.cpu().numpy()
        assert (engine.KERNELS["adjoint"] == "wave2d_persist_adjoint_kernel") == bool(persist), engine.KERNELS
    assert (engine.KERNELS["forward"] == "wave2d_persist_forward_kernel") == bool(persist), engine.KERNELS
    return recs, g, engine.LAUNCHES["forward"] - l0["forward"]


@pytest.mark.parametrize("variant", [0, 1])
def test_cfg1_one_launch_bit_equal_to_per_step_kernels(variant, monkeypatch):
    """BASELINE configs[0]: 2000 time steps in ONE launch; records identical to 2000 single-step launches."""
    import bench
    true, _ = bench.WORKLOADS["cfg1"]["models"]()
    case = bench.make_case(1, workload="cfg1", models=true)
    r1, _, n1 = _run(case, monkeypatch, True, mode="forward", variant=variant)
    r0, _, n0 = _run(case, monkeypatch, False, mode="forward")
    assert n1 == 1 and n0 == 2000
    assert np.abs(r0[0]).max() > 0 and np.array_equal(r1[0], r0[0])


@pytest.mark.parametrize("nz,nx,nshots", [(150, 300, 3), (30, 44, 2), (100, 180, 2), (156, 412, 1)])
def test_gradient_runs_history_segments_and_oracle(nz, nx, nshots, monkeypatch):
    """Gradient runs: the persistent kernel writes the wavefield history (and restarts from K-step checkpoints);
    records and gradients identical to the per-step path, and within tolerance of the float64 oracle.  Grids with
    1..4 column strips, a partly filled last strip / last row strip, several shots (one cluster per shot)."""
    from oracle import cases, loop, misfit
    case = cases.make_case("acoustic", nz=nz, nx=nx, nshots=nshots, nt=150, rec_step=5)
    r1, g1, _ = _run(case, monkeypatch, True)
    w1 = _run.wavelet_grad
    r0, g0, _ = _run(case, monkeypatch, False)
    w0 = _run.wavelet_grad
    # forward: same expression -> identical records and stored states; the adjoint twin accumulates the gradient in
    # registers over the whole loop (the per-step kernels add one step at a time to the plane): rounding only
    assert all(np.array_equal(a, b) for a, b in zip(r1, r0))
    assert rel(g1, g0) < 2e-6 and rel(w1, w0) < 2e-6 and np.abs(w0).max() > 0
    # K-step checkpoints + recomputation through the persistent kernels: bit for bit the stored-history result (the
    # twin's gradient accumulator continues from the plane's running value)
    r2, g2, n2 = _run(case, monkeypatch, True, segment=41)
    assert all(np.array_equal(a, b) for a, b in zip(r2, r1)) and np.array_equal(g2, g1) and np.array_equal(_run.wavelet_grad, w1)
    orecs, params = loop.simulate(case, dtype=torch.float64, requires_grad=["vp"])
    misfit.l2(orecs, [torch.zeros_like(r) for r in orecs]).backward()
    assert rel(cat_records(r1), cat_records([r.detach().numpy() for r in orecs])) < 1e-5
    assert rel(g1, params["vp"].grad.numpy()) < 1e-4


def test_dense_receivers_ragged_shots_and_clustered_sources(monkeypatch):
    """More receivers per row strip than the kernel caches (walks the CSR instead), a different receiver count
    per shot, receivers on several rows; source encoding with four sources inside ONE thread's 4 x 4 cell patch
    (more than the two it tracks in registers) each with its own wavelet."""
    from oracle import cases
    case = cases.make_case("acoustic", nz=60, nx=200, nshots=2, nt=80)
    xs = list(range(0, 200))
    zs = list(range(2, 18))
    dense = [[x for z in zs for x in xs], [z for z in zs for x in xs]]          # 3200 receivers on 16 rows
    case["receivers"] = [dense, [xs[::7], [5] * len(xs[::7])]]
    r1, _, _ = _run(case, monkeypatch, True, mode="forward")
    r0, _, _ = _run(case, monkeypatch, False, mode="forward")
    assert [r.shape for r in r1] == [(80, 3200, 1), (80, len(xs[::7]), 1)]
    assert all(np.array_equal(a, b) for a, b in zip(r1, r0)) and np.abs(r1[0]).max() > 0
    enc = cases.make_case("acoustic", nz=60, nx=200, nshots=4, nt=80)
    enc["sources"] = [[100.0, 20.0], [101.0, 20.0], [102.0, 21.0], [103.0, 22.0]]
    enc["receivers"] = [enc["receivers"][0]] * 4
    w = np.stack([np.asarray(enc["wavelet"]) * s for s in (1.0, -0.5, 2.0, 0.25)]).astype(np.float32)
    e1, _, _ = _run(enc, monkeypatch, True, mode="forward", encoding=True, wav=w)
    e0, _, _ = _run(enc, monkeypatch, False, mode="forward", encoding=True, wav=w)
    assert np.abs(e0[0]).max() > 0 and rel(e1[0], e0[0]) < 1e-6        # atomics order of coincident adds may differ
    # gradient run of the dense / ragged acquisition (receiver staging beyond one record per thread and beyond the cache)
    # and of the clustered encoded sources (d loss / d wavelet of more than two sources per thread)
    r1, g1, _ = _run(case, monkeypatch, True)
    w1 = _run.wavelet_grad
    r0, g0, _ = _run(case, monkeypatch, False)
    assert rel(g1, g0) < 2e-6 and rel(w1, _run.wavelet_grad) < 2e-6
    x1 = _run(enc, monkeypatch, True, encoding=True, wav=w)
    w1 = _run.wavelet_grad
    x0 = _run(enc, monkeypatch, False, encoding=True, wav=w)
    assert rel(x1[1], x0[1]) < 2e-6 and rel(w1, _run.wavelet_grad) < 2e-6 and w1.shape == w.shape


def test_falls_back_when_the_grid_is_outside_its_class(monkeypatch):
    """HABC equations and grids that do not fit one cluster keep the per-step kernels (no error, same results)."""
    from oracle import cases
    from seistorch_b200 import engine
    import seistorch_b200 as sb
    monkeypatch.setenv("SEISTORCH_B200_PERSIST", "1")
    for eq, nz, nx in (("acoustic_habc", 40, 60), ("acoustic", 400, 300), ("acoustic", 100, 600)):
        case = cases.make_case(eq, nz=nz, nx=nx, nshots=1, nt=12)
        cfg, model = sb.model_from_case(case, device="cuda", mode="forward")
        with torch.no_grad():
            model(torch.as_tensor(np.asarray(case["wavelet"]), device="cuda").unsqueeze(0))
        assert engine.KERNELS["forward"] != "wave2d_persist_forward_kernel", (eq, nz, nx)
