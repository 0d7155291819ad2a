"""CPU: host-side logic of the product package (no kernels): index semantics, coefficient
generators and maps, acquisition tables, history planning, containers."""
import math

import numpy as np
import pytest
import torch

from conftest import load_golden, rel


def test_boundary_coefficients_equal_oracle():
    from oracle import boundary
    from seistorch_b200 import habc, pml
    for shape, mult in [((130, 144), False), ((80, 144), True), ((101, 203), False)]:
        assert torch.equal(pml.generate_pml_coefficients_2d(shape, multiple=mult), boundary.pml_coefficients_2d(shape, multiple=mult))
        assert torch.equal(habc.generate_habc_coefficients_2d(shape, multiple=mult), boundary.habc_coefficients_2d(shape, multiple=mult))
    assert torch.equal(pml.generate_pml_coefficients_3d((114, 112, 110)), boundary.pml_coefficients_3d((114, 112, 110)))
    tm, bm, lm, rm = habc.bound_mask(130, 144, 50, "cpu", batchsize=3, return_idx=True)
    otm, obm, olm, orm = boundary.habc_masks(130, 144, 50)
    assert tm.shape == (3, 50, 144) and np.array_equal(tm[0].numpy(), otm) and np.array_equal(rm[2].numpy(), orm)
    assert habc.bound_mask(80, 144, 50, "cpu", return_idx=True, multiple=True)[0] is None


def test_source_receiver_index_semantics():
    """SURVEY 8a probe of the reference: add bwidth, truncate toward zero; sources go through
    float32, receiver lists through float64."""
    from oracle import loop
    from seistorch_b200.setup import setup_rec_coords, setup_src_coords
    s = setup_src_coords([10.7, 3.2], 50)
    assert (int(s.x), int(s.y)) == (60, 53)
    r = setup_rec_coords([[10.7, 11.9, 9.0], [3.2, 3.0, 3.99]], 50)[0]
    assert r.x.tolist() == [60, 61, 59] and r.y.tolist() == [53, 53, 53]
    assert r.x.dtype == torch.int64
    s = setup_src_coords([10.7, 3.2], 50, multiple=True)
    assert (int(s.x), int(s.y)) == (60, 3)
    for v in [0.0, 0.999999, 7.5, 123.0000001]:
        assert int(setup_src_coords([v, 1], 50).x) == int(loop.source_indices([[v, 1]])[0, 0])
    s3 = setup_src_coords([3.3, 4.4, 1.0], 50)
    assert (int(s3.x), int(s3.y), int(s3.z)) == (53, 54, 51)


def test_acquisition_tables_on_cpu():
    from seistorch_b200.engine import Acquisition
    rec_b = torch.tensor([0, 0, 1, 1, 1])
    rec_idx = torch.tensor([[5, 9], [2, 3], [5, 1], [5, 0], [0, 7]])
    acq = Acquisition((6, 10), 2, torch.tensor([0, 1]), torch.tensor([[1, 2], [3, 4]]), rec_b, rec_idx, "cpu")
    rs = acq.row_start.tolist()
    assert len(rs) == 2 * 6 + 1 and rs[-1] == 5
    # every receiver is found in the row it belongs to, and rec_orig maps back
    for k in range(5):
        row = int(rec_b[k]) * 6 + int(rec_idx[k, 0])
        cols = acq.rec_col[rs[row]:rs[row + 1]].tolist()
        origs = acq.rec_orig[rs[row]:rs[row + 1]].tolist()
        assert int(rec_idx[k, 1]) in cols and k in origs
    with pytest.raises(IndexError):
        Acquisition((6, 10), 2, torch.tensor([0]), torch.tensor([[6, 2]]), rec_b, rec_idx, "cpu")
    with pytest.raises(IndexError):
        Acquisition((6, 10), 2, torch.tensor([2]), torch.tensor([[1, 2]]), rec_b, rec_idx, "cpu")
    empty = Acquisition((6, 10), 1, torch.zeros(0), torch.zeros(0, 2), torch.zeros(0), torch.zeros(0, 2), "cpu")
    assert empty.R == 0 and empty.ns == 0 and empty.row_start.tolist() == [0] * 7


def test_history_plan():
    from seistorch_b200.engine import Spec, _history_plan
    spec = Spec("wave2d", 5, (100, 200), 2, 1000, 1e-3)
    slot = spec.slot_elems * 4
    spec.history_budget_bytes = slot * 2000
    assert _history_plan(spec, "cpu") == (1000, 1)               # everything fits: no recompute
    spec.history_budget_bytes = slot * 100
    K, nseg = _history_plan(spec, "cpu")
    assert nseg == math.ceil(1000 / K) and K + 2 + 2 * nseg <= 100 and K > 20
    spec.segment = 7
    assert _history_plan(spec, "cpu") == (7, 143)
    spec.segment, spec.history_budget_bytes = None, slot * 3
    with pytest.raises(RuntimeError, match="does not fit"):
        _history_plan(spec, "cpu")


@pytest.mark.parametrize("eq", ["acoustic", "acoustic_habc", "vti_habc2", "tti_habc", "acoustic_fwim_habc"])
def test_coefficient_maps_reproduce_one_oracle_step(eq):
    """y = h1 + alpha (h1-h2) + A[h1] with our coefficient planes equals the oracle's (reference's)
    _time_step in the interior (fp64)."""
    from oracle import cases, equations, loop
    from seistorch_b200 import coefficients as cf
    case = cases.make_case(eq, nz=12, nx=16, nshots=1, nt=2)
    names, params, d, *_ = loop.build_geometry(case, torch.float64)
    g = torch.Generator().manual_seed(0)
    shape = tuple(params[0].shape)
    h1 = torch.randn((1,) + shape, generator=g, dtype=torch.float64)
    h2 = torch.randn((1,) + shape, generator=g, dtype=torch.float64)
    dt, h = torch.tensor(1e-3, dtype=torch.float64), torch.tensor(10.0, dtype=torch.float64)
    ref = equations.get_step(eq)(params, [h1, h2], dt, h, d)[0]
    coefs, slots = cf.wave2d_coefficients(eq, params, 1e-3, 10.0, d)
    c = {s: t.double() for s, t in zip(slots, coefs)}       # fp32-rounded planes
    pad = torch.nn.functional.pad(h1, (1, 1, 1, 1))
    C, N, S = pad[:, 1:-1, 1:-1], pad[:, :-2, 1:-1], pad[:, 2:, 1:-1]
    W, E = pad[:, 1:-1, :-2], pad[:, 1:-1, 2:]
    if eq in ("acoustic", "acoustic_habc", "acoustic_fwim_habc"):
        A = c[2] * (N + S + E + W - 4 * C)
    else:
        A = c[2] * (E + W - 2 * C) + c[3] * (N + S - 2 * C)
    if 4 in c:
        A = A + c[4] * ((pad[:, 2:, 2:] - pad[:, 2:, :-2]) - (pad[:, :-2, 2:] - pad[:, :-2, :-2]))
    if 5 in c:
        A = A + c[5] * (E - W) + c[6] * (S - N)
    alpha = c[3] if eq == "acoustic" else 1.0
    y = h1 + alpha * (h1 - h2) + A
    sl = (slice(None), slice(51, -51), slice(51, -51)) if "habc" in eq else (slice(None),) * 3
    assert rel(y[sl].numpy(), ref[sl].numpy()) < 5e-7          # fp32 rounding of the planes only


def test_tensorlist_and_loss_registry():
    import seistorch_b200 as sb
    tl = sb.TensorList([torch.ones(4, 3, 1), torch.ones(4, 3, 1)])
    assert tl.stack().shape == (2, 4, 3, 1)
    assert len(tl) == 2 and tl.has_nan() is False
    # reference quirk kept (type.py:56): F.pad pads the LAST dims, so ragged receiver counts
    # pad the channel axis and stacking fails exactly like in the reference
    with pytest.raises(RuntimeError):
        sb.TensorList([torch.ones(4, 3, 1), torch.ones(4, 2, 1)]).stack()
    bad = sb.TensorList([torch.tensor([float("nan")])])
    with pytest.raises(ValueError):
        bad.has_nan()
    assert sb.Loss("l2").loss(None).name == "l2" and sb.Loss("envelope").loss(None).name == "envelope"
    with pytest.raises(ValueError):
        sb.Loss("nope").loss(None)


def test_hilbert_kernel_matches_reference_transform():
    from oracle import misfit
    from seistorch_b200.loss import hilbert_kernel
    for nt in (16, 17, 120):
        x = torch.randn(nt, 3, 1, dtype=torch.float64)
        ana = misfit.hilbert(x)
        hk = hilbert_kernel(nt, "cpu").double()
        idx = (torch.arange(nt)[:, None] - torch.arange(nt)[None, :]) % nt
        Hx = torch.einsum("nm,mrc->nrc", hk[idx], x)
        assert rel(Hx.numpy(), ana.imag.numpy()) < 1e-6
        assert rel(ana.real.numpy(), x.numpy()) < 1e-12


def test_build_model_surface():
    import seistorch_b200 as sb
    z, case = load_golden("acoustic_habc")
    cfg, model = sb.model_from_case(case, device="cpu", mode="inversion")
    geom = model.cell.geom
    assert geom.domain_shape == (130, 144) and geom.use_habc and geom.bwidth == 50
    assert geom.pars_need_invert == ["vp"] and geom.vp.requires_grad
    assert [n for n, _ in model.named_parameters()] == ["vp"]
    assert model.second_order_equation and len(model.sources) == 2
    import seistorch_b200.equations2d.acoustic_habc as m
    assert model.cell.forward_func is m._time_step
    with pytest.raises(NotImplementedError):
        m._time_step_backward()


def test_signal_host_logic_and_loss_registry():
    """Host-side mirror of seistorch/signal.py:23-76 (filter type, Butterworth design) and of the loss registry
    (loss.py:23-50: classes found by their `name`); no device work."""
    import numpy as np
    import pytest
    from scipy import signal as sps
    import seistorch_b200 as sb
    from seistorch_b200.signal import SeisSignal
    sig = SeisSignal({"geom": {"dt": 0.002}, "training": {"filter_ord": 3}})
    assert sig.decide_filter_type(5.0) == "lowpass" and sig.decide_filter_type([5.0]) == "lowpass"
    assert sig.decide_filter_type([3.0, 8.0]) == "bandpass" and sig.decide_filter_type("all") == "all"
    b, a = sig.design([30.0])
    b0, a0 = sps.butter(3, Wn=2 * 30.0 * 0.002, btype="lowpass")
    assert np.array_equal(b, b0) and np.array_equal(a, a0)
    b, a = sig.design([8.0, 60.0])
    b0, a0 = sps.butter(3, Wn=[2 * 8.0 * 0.002, 2 * 60.0 * 0.002], btype="bandpass")
    assert np.array_equal(b, b0) and np.array_equal(a, a0)
    marker = object()
    assert sig.filter(marker, "all") is marker
    with pytest.raises(NotImplementedError):
        sig.filter(marker, [30.0], backend="scipy")
    for name in ("l2", "l1", "sml1", "cs", "cc", "integration", "nim", "w1d", "traveltime", "envelope"):
        assert sb.Loss(name).loss(None).name == name
    with pytest.raises(ValueError):
        sb.Loss("sinkhorn").loss(None)
    with pytest.raises(NotImplementedError):
        sb.Loss("nim").loss(None, criterion="l1")
