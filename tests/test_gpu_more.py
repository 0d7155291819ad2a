"""GPU (-m gpu): edge cases and size-independent properties of the product path."""
import numpy as np
import pytest
import torch

from conftest import cat_records, rel

pytestmark = pytest.mark.gpu


def _model(case, mode="inversion", **kw):
    import seistorch_b200 as sb
    cfg, model = sb.model_from_case(case, device="cuda", mode=mode, **kw)
    x = torch.as_tensor(np.asarray(case["wavelet"]), dtype=torch.float32, device="cuda").unsqueeze(0)
    return cfg, model, x


def test_3d_single_step_value_and_vjp():
    from oracle import cases, equations, loop
    import seistorch_b200.equations3d.acoustic as mod
    case = cases.make_case("acoustic", nz=6, nx=7, ny=5, nshots=2, nt=2)
    names, params, d, *_ = loop.build_geometry(case, torch.float64)
    shape = tuple(params[0].shape)
    g = torch.Generator().manual_seed(5)
    F = [torch.randn((2,) + shape, generator=g, dtype=torch.float64).requires_grad_(True) for _ in range(2)]
    G = [torch.randn((2,) + shape, generator=g, dtype=torch.float64) for _ in range(2)]
    P = [params[0].clone().requires_grad_(True)]
    dt, h = torch.tensor(1e-3, dtype=torch.float64), torch.tensor(10.0, dtype=torch.float64)
    outs = equations.get_step("acoustic3d")(P, F, dt, h, d)
    torch.autograd.backward(list(outs), G)
    Pc = [params[0].float().cuda().requires_grad_(True)]
    Fc = [f.detach().float().cuda().requires_grad_(True) for f in F]
    outs_c = mod._time_step(*Pc, *Fc, torch.tensor(1e-3).cuda(), torch.tensor(10.0).cuda(), d.float().cuda())
    for a, b in zip(outs_c, outs):
        assert rel(a.detach().cpu().numpy(), b.detach().numpy()) < 2e-6
    torch.autograd.backward(list(outs_c), [g_.float().cuda() for g_ in G])
    for a, b in zip(Fc, F):
        assert rel(a.grad.cpu().numpy(), b.grad.numpy()) < 1e-5
    assert rel(Pc[0].grad.cpu().numpy(), P[0].grad.numpy()) < 1e-4


def test_source_encoding_and_per_source_wavelets():
    """codingfwi.py mode: all sources fire into ONE wavefield, each with its own
    wavelet (source.py:54-55, rnn.py:113,162) == sum of the single-shot wavefields (linearity)."""
    from oracle import cases
    case = cases.make_case("acoustic_habc", nz=30, nx=50, nshots=3, nt=90)
    case["receivers"] = [case["receivers"][0]]          # one common receiver spread
    # encoded run: batch 1, 3 sources, per-source wavelets
    import seistorch_b200 as sb
    enc = dict(case)
    cfg, model = sb.model_from_case(dict(case, receivers=case["receivers"] * 3), device="cuda", mode="forward",
                                    source_encoding=True)
    model.reset_probes(model.probes[0])
    w = np.stack([np.asarray(case["wavelet"]) * s for s in (1.0, -0.5, 2.0)]).astype(np.float32)
    with torch.no_grad():
        rec_enc = model(torch.as_tensor(w, device="cuda"))[0].cpu().numpy()
    # reference by superposition of three ordinary single-shot runs
    tot = 0.0
    for k, s in enumerate((1.0, -0.5, 2.0)):
        one = dict(case, sources=[case["sources"][k]], wavelet=np.asarray(case["wavelet"]) * s)
        cfg1, m1, x1 = _model(one, mode="forward")
        with torch.no_grad():
            tot = tot + m1(x1)[0].cpu().numpy().astype(np.float64)
    assert rel(rec_enc, tot) < 2e-6


def test_source_encoding_against_reference_golden_and_oracle():
    """Encoded mode against the REAL reference (tests/golden/acoustic_habc_encoded.npz, generated through the call order of
    codingfwi.py) -- records and the gradient w.r.t. vp -- and against the float64 oracle's encoded mode on a mid-size
    grid with sources in different tile kinds."""
    import seistorch_b200 as sb
    from conftest import load_golden
    from oracle import cases, loop, misfit
    z, case = load_golden("acoustic_habc_encoded")
    for c, w in ((case, z["enc_wavelets"]), (None, None)):
        if c is None:
            c = cases.make_case("acoustic_habc", nz=150, nx=300, nshots=5, nt=80, rec_step=7)
            c["sources"] = [[s[0], 3.0 + 30 * k] for k, s in enumerate(c["sources"])]
            w = np.stack([np.asarray(c["wavelet"]) * s for s in (1.0, -1.0, 0.5, -2.0, 1.5)]).astype(np.float32)
        cfg, model = sb.model_from_case(c, device="cuda", mode="inversion", source_encoding=True)
        model.reset_probes(model.probes[0])
        syn = model(torch.as_tensor(w, device="cuda"))
        assert len(syn) == 1
        (syn[0].double() ** 2).sum().backward()
        g = model.cell.geom.vp.grad.cpu().numpy()
        if c is case:
            assert rel(syn[0].detach().cpu().numpy(), z["f64_rec_0"]) < 1e-5
            assert rel(g, z["f64_grad_vp"]) < 1e-4
        else:
            orecs, params = loop.simulate(c, dtype=torch.float64, requires_grad=["vp"], wavelet=torch.as_tensor(w), source_encoding=True)
            misfit.l2(orecs, [torch.zeros_like(r) for r in orecs]).backward()
            assert rel(syn[0].detach().cpu().numpy(), orecs[0].detach().numpy()) < 1e-5
            assert rel(g, params["vp"].grad.numpy()) < 1e-4


@pytest.mark.parametrize("eq,seg", [("acoustic_habc", None), ("acoustic_habc", 23), ("elastic", None), ("acoustic", None)])
def test_source_illumination_from_the_history(eq, seg):
    """rnn.py:127-128,204-205 + process.py:115-118: precondition = sum over time and shots of field[source_type]^2, used to
    scale the gradient.  Filled during backward() from the wavefield history (whole history or K-step segments; through the
    persistent kernels for the small acoustic grid); against the float64 oracle."""
    import seistorch_b200 as sb
    from seistorch_b200.process import PostProcess
    from types import SimpleNamespace
    from oracle import cases, loop
    case = cases.make_case(eq, nz=40, nx=64, nshots=2, nt=90)
    cfg, model = sb.model_from_case(case, device="cuda", mode="inversion")
    model.source_illumination = True
    model.segment = seg
    x = torch.as_tensor(np.asarray(case["wavelet"]), device="cuda").unsqueeze(0)
    syn = model(x)
    assert float(model.precondition.abs().max()) == 0.0          # reset at every forward call, filled by backward()
    sum((s ** 2).sum() for s in syn).backward()
    ill = []
    loop.simulate(case, dtype=torch.float64, illumination=ill)
    assert rel(model.precondition.cpu().numpy(), ill[0].numpy()) < 1e-5
    g0 = model.cell.geom.vp.grad.clone()
    PostProcess(model, cfg, SimpleNamespace(grad_cut=False)).precondition()
    lit = model.precondition > 0                                # cells the wavefield never reached divide 0 by 0, as in the reference
    assert bool(lit.any()) and torch.allclose((model.cell.geom.vp.grad * model.precondition)[lit], g0[lit], rtol=1e-5, atol=0)
    with pytest.raises(NotImplementedError):
        with torch.no_grad():
            model(x)


def test_empty_receivers_and_zero_wavelet():
    from oracle import cases
    case = cases.make_case("acoustic", nz=20, nx=30, nshots=2, nt=20)
    case["wavelet"] = np.zeros(20, np.float32)
    cfg, model, x = _model(case, mode="forward")
    with torch.no_grad():
        out = model(x)
    assert all(float(o.abs().max()) == 0.0 for o in out)


def test_out_of_domain_receiver_raises():
    from oracle import cases
    case = cases.make_case("acoustic", nz=20, nx=30, nshots=1, nt=5)
    case["receivers"] = [[[500], [2]]]
    cfg, model, x = _model(case, mode="forward")
    with pytest.raises(IndexError):
        model(x)


def test_nan_in_model_raises_value_error():
    """type.py:41-46: NaN in the records raises ValueError."""
    from oracle import cases
    case = cases.make_case("acoustic", nz=20, nx=30, nshots=1, nt=30)
    case["models"]["vp"] = case["models"]["vp"].copy()
    case["models"]["vp"][2, 10] = np.nan
    cfg, model, x = _model(case, mode="forward")
    with pytest.raises(ValueError):
        model(x)


@pytest.mark.parametrize("tma", ["1", "0", "auto"])
def test_baseline_size_properties(tma, monkeypatch):
    """BASELINE configs[1] grid (2301x751, padded 2401x851), short horizon: (i) checkpoint-recompute
    equals stored history bit for bit, (ii) linearity in the wavelet, (iii) shot independence:
    a 2-shot batch equals the two single-shot runs -- bit for bit when both go through the same kernels
    (TMA forced on / off), to rounding when the automatic choice sends the single shots to the register
    kernels and the batch to the TMA kernels."""
    import bench
    if tma != "auto":
        monkeypatch.setenv("SEISTORCH_B200_TMA", tma)
    else:
        monkeypatch.delenv("SEISTORCH_B200_TMA", raising=False)
    true, init = bench.make_models()
    case = bench.make_case(2, vp=init, nt=150)
    def grad_of(c, segment=None):
        cfg, model, x = _model(c)
        model.segment = segment
        syn = model(x)
        loss = sum((s ** 2).sum() for s in syn)
        loss.backward()
        return [s.detach().cpu().numpy() for s in syn], model.cell.geom.vp.grad.cpu().numpy()
    r0, g0 = grad_of(case)
    r1, g1 = grad_of(case, segment=37)
    assert all(np.array_equal(a, b) for a, b in zip(r0, r1)) and np.array_equal(g0, g1)
    ra, ga = grad_of(dict(case, sources=case["sources"][:1], receivers=case["receivers"][:1]))
    rb, gb = grad_of(dict(case, sources=case["sources"][1:], receivers=case["receivers"][1:]))
    if tma != "auto":
        assert np.array_equal(r0[0], ra[0]) and np.array_equal(r0[1], rb[0])
    else:
        assert rel(r0[0], ra[0]) < 2e-6 and rel(r0[1], rb[0]) < 2e-6
    assert rel(g0, ga + gb) < 1e-5
    r4, _ = grad_of(dict(case, wavelet=np.asarray(case["wavelet"]) * 4.0))
    assert rel(cat_records(r4), 4.0 * cat_records(r0)) < 1e-6


def test_l2_and_envelope_misfit_kernels_against_oracle():
    from oracle import misfit
    import seistorch_b200 as sb
    g = torch.Generator().manual_seed(1)
    syn = torch.randn(2, 64, 5, 2, generator=g)
    obs = torch.randn(2, 64, 5, 2, generator=g)
    for name, fn in (("l2", misfit.l2), ("envelope", misfit.envelope_loss)):
        s64 = syn.double().requires_grad_(True)
        lo = fn(list(s64), list(obs.double()))
        lo.backward()
        sc = syn.cuda().requires_grad_(True)
        lc = sb.Loss(name).loss(None)(sc, obs.cuda())
        lc.backward()
        assert abs(float(lc) - float(lo)) <= 2e-5 * abs(float(lo)), name
        assert rel(sc.grad.cpu().numpy(), s64.grad.numpy()) < 2e-5, name
        # list-of-shots form (ragged-capable)
        sc2 = [syn[k].cuda().requires_grad_(True) for k in range(2)]
        l2_ = sb.Loss(name).loss(None)(sc2, [obs[k].cuda() for k in range(2)])
        assert abs(float(l2_) - float(lo)) <= 2e-5 * abs(float(lo)), name


@pytest.mark.parametrize("eq", ["elastic", "acoustic", "vti_habc2", "tti_habc", "acoustic_fwim_habc", "acoustic_tti_lsrtm_habc"])
def test_mid_size_grid_against_oracle(eq):
    """250x400 padded grid: large enough that every kernel variant (vectorised interior tiles,
    border / frame tiles, tap-gather band) is exercised; compared with the oracle in float64."""
    from oracle import cases, loop, misfit
    case = cases.make_case(eq, nz=150, nx=300, nshots=2, nt=70, rec_step=9)
    cfg, model, x = _model(case)
    syn = model(x)
    loss = sum((s ** 2).sum() for s in syn)
    loss.backward()
    inv = [k for k, v in case["invlist"].items() if v]
    orecs, params = loop.simulate(case, dtype=torch.float64, requires_grad=inv)
    misfit.l2(orecs, [torch.zeros_like(r) for r in orecs]).backward()
    assert rel(cat_records([s.detach().cpu().numpy() for s in syn]), cat_records([r.detach().numpy() for r in orecs])) < 1e-5
    for k in inv:
        assert rel(getattr(model.cell.geom, k).grad.cpu().numpy(), params[k].grad.numpy()) < 1e-4, k


@pytest.mark.parametrize("acq", ["deep", "surface"])
@pytest.mark.parametrize("eq,multiple,nx", [("acoustic", False, 300), ("acoustic_habc", False, 300), ("acoustic_habc", True, 300),
                                            ("acoustic_habc", False, 216), ("acoustic_habc", True, 216)])
def test_tma_path_against_oracle_and_register_path(eq, multiple, nx, acq, monkeypatch):
    """The TMA-staged blocks (forced on: SEISTORCH_B200_TMA=1) against the float64 oracle and against the
    register/shuffle + tap-gather path (SEISTORCH_B200_TMA=0), 3 shots (ragged last shot group), sources
    and receivers inside TMA tiles.  Padded 250x400: frame-free + top/bottom frame tiles; padded 250x316:
    also the left/right frame tiles (one tile column per side)."""
    from oracle import cases, loop, misfit
    case = cases.make_case(eq, nz=150, nx=nx, nshots=3, nt=70, rec_step=9, multiple=multiple)
    # "deep": acquisition rows inside the side tiles / frame-free tiles; "surface" (the bench geometry): sources and
    # receivers just below the top frame, i.e. in the corner rows (generic corner tiles + masked TMA tiles) and in
    # the one-shot-per-block acquisition rows
    if acq == "deep":
        case["sources"] = [[s[0], 40.2] for s in case["sources"]]
        case["receivers"] = [[r[0], [30] * len(r[0])] for r in case["receivers"]]

    def run(mode):
        monkeypatch.setenv("SEISTORCH_B200_TMA", mode)
        cfg, model, x = _model(case)
        syn = model(x)
        loss = sum((s ** 2).sum() for s in syn)
        loss.backward()
        return cat_records([s.detach().cpu().numpy() for s in syn]), model.cell.geom.vp.grad.cpu().numpy()

    r_tma, g_tma = run("1")
    r_reg, g_reg = run("0")
    # two kernel families, same equations: fp32 rounding only (the gradient peaks at the source cells, where the
    # summation order of the frame terms differs most: 1.3e-6 measured)
    assert rel(r_tma, r_reg) < 2e-6 and rel(g_tma, g_reg) < 5e-6
    orecs, params = loop.simulate(case, dtype=torch.float64, requires_grad=["vp"])
    misfit.l2(orecs, [torch.zeros_like(r) for r in orecs]).backward()
    assert rel(r_tma, cat_records([r.detach().numpy() for r in orecs])) < 1e-5
    assert rel(g_tma, params["vp"].grad.numpy()) < 1e-4


def _props(case, inv, segment, scale=3.0, shot_tol=None):
    """Size-independent properties on one case: returns nothing, asserts
    (i) K-step checkpoint-recompute == stored history bit for bit (records and gradients),
    (ii) linearity of the records in the wavelet,
    (iii) shot independence: the 2-shot batch equals the two single-shot runs (records bit for bit unless
         shot_tol is given, gradients add up)."""
    def grad_of(c, seg=None):
        cfg, model, x = _model(c)
        model.segment = seg
        syn = model(x)
        loss = sum((s ** 2).sum() for s in syn)
        loss.backward()
        return ([s.detach().cpu().numpy() for s in syn],
                {k: getattr(model.cell.geom, k).grad.cpu().numpy().astype(np.float64) for k in inv})
    r0, g0 = grad_of(case)
    r1, g1 = grad_of(case, seg=segment)
    assert all(np.array_equal(a, b) for a, b in zip(r0, r1))
    assert all(np.array_equal(g0[k], g1[k]) for k in inv)
    assert all(np.isfinite(g0[k]).all() and np.abs(g0[k]).max() > 0 for k in inv)
    ra, ga = grad_of(dict(case, sources=case["sources"][:1], receivers=case["receivers"][:1]))
    rb, gb = grad_of(dict(case, sources=case["sources"][1:], receivers=case["receivers"][1:]))
    if shot_tol is None:
        assert np.array_equal(r0[0], ra[0]) and np.array_equal(r0[1], rb[0])
    else:
        assert rel(r0[0], ra[0]) < shot_tol and rel(r0[1], rb[0]) < shot_tol
    for k in inv:
        assert rel(g0[k], ga[k] + gb[k]) < 2e-5, k
    rs, _ = grad_of(dict(case, wavelet=np.asarray(case["wavelet"]) * scale))
    assert rel(cat_records(rs), scale * cat_records(r0)) < 2e-6


def test_baseline_size_properties_cfg3_elastic():
    """BASELINE configs[2]: elastic velocity-stress, PML, 1000x400 model (padded 1100x500), vp/vs/rho inverted,
    source vz, receivers vx+vz; short horizon."""
    from oracle import cases
    case = cases.make_case("elastic", nz=400, nx=1000, nshots=2, nt=120, rec_step=2)
    _props(case, ["vp", "vs", "rho"], segment=29)


@pytest.mark.parametrize("eq", ["acoustic_vti_lsrtm_habc", "acoustic_tti_lsrtm_habc", "acoustic_fwim_habc"])
def test_baseline_size_properties_cfg4(eq):
    """BASELINE configs[3]: VTI / TTI qP LSRTM (Born pair, receivers on the scattered field, m inverted jointly
    with vp) and the joint FWI-LSRTM equation, 1200x500 model (padded 1300x600); short horizon."""
    from oracle import cases
    case = cases.make_case(eq, nz=500, nx=1200, nshots=2, nt=120, rec_step=2)
    inv = [k for k, v in case["invlist"].items() if v]
    _props(case, inv, segment=31)


def test_baseline_size_properties_cfg5_3d():
    """BASELINE configs[4]: 3D acoustic, PML, model file (400,200,400) -> padded (500,300,500) tensor layout (x,z,y),
    K-step checkpointed wavefield reconstruction; 2 shots, short horizon (75 M cells per shot and state)."""
    from oracle import cases
    case = cases.make_case("acoustic", nz=200, nx=400, ny=400, nshots=2, nt=24, rec_step=8)
    _props(case, ["vp"], segment=7)


def test_tma_path_vertical_receiver_line_and_ragged_shots(monkeypatch):
    """TMA kernels with receivers on a vertical line (VSP-like: the acquisition spans too many rows for the
    one-shot-per-block split, so the source / receiver epilogue runs inside the multi-shot tile blocks) and a
    different receiver count per shot; against the register path and the float64 oracle."""
    from oracle import cases, loop, misfit
    case = cases.make_case("acoustic_habc", nz=150, nx=216, nshots=3, nt=70, rec_step=9)
    zs = list(range(3, 147, 4))
    case["receivers"] = [[[100 + 7 * k] * (len(zs) - 3 * k), zs[:len(zs) - 3 * k]] for k in range(3)]
    case["sources"] = [[s[0], 20.0 + 30 * k] for k, s in enumerate(case["sources"])]

    def run(mode):
        monkeypatch.setenv("SEISTORCH_B200_TMA", mode)
        cfg, model, x = _model(case)
        syn = model(x)
        assert [tuple(s.shape) for s in syn] == [(70, len(zs) - 3 * k, 1) for k in range(3)]
        loss = sum((s ** 2).sum() for s in syn)
        loss.backward()
        from seistorch_b200 import engine
        assert ("tma" in engine.KERNELS["adjoint"]) == (mode == "1")
        return cat_records([s.detach().cpu().numpy() for s in syn]), model.cell.geom.vp.grad.cpu().numpy()

    r_tma, g_tma = run("1")
    r_reg, g_reg = run("0")
    assert rel(r_tma, r_reg) < 2e-6 and rel(g_tma, g_reg) < 5e-6
    orecs, params = loop.simulate(case, dtype=torch.float64, requires_grad=["vp"])
    misfit.l2(orecs, [torch.zeros_like(r) for r in orecs]).backward()
    assert rel(r_tma, cat_records([r.detach().numpy() for r in orecs])) < 1e-5
    assert rel(g_tma, params["vp"].grad.numpy()) < 1e-4
