"""GPU (-m gpu): the time loops of the C-ABI calls replayed as CUDA graphs (csrc/st_graph.cuh).  An inversion makes the
same call every iteration: the first sighting runs the plain launch loop, the second is captured, later ones are replayed.
Every iteration must give bit-identical records and gradients (the graph holds the very launches of the loop), and the
counters exported by st_graph_counters must show the capture and the replays."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _iterate(eq, n_iter, **kw):
    import seistorch_b200 as sb
    from oracle import cases
    case = cases.make_case(eq, **kw)
    cfg, model = sb.model_from_case(case, device="cuda", mode="inversion")
    x = torch.as_tensor(np.asarray(case["wavelet"]), dtype=torch.float32, device="cuda").unsqueeze(0)
    params = [getattr(model.cell.geom, k) for k in model.cell.geom.model_parameters if getattr(model.cell.geom, k).requires_grad]
    outs = []
    for _ in range(n_iter):
        for p in params:
            p.grad = None
        syn = model(x)
        loss = sum((s ** 2).sum() for s in syn)
        loss.backward()
        outs.append(([s.detach().cpu().numpy() for s in syn], [p.grad.cpu().numpy() for p in params]))
        del syn, loss
    return outs


@pytest.mark.parametrize("eq,kw", [
    ("acoustic_habc", dict(nz=70, nx=300, nshots=4, nt=96, rec_step=3)),          # TMA kernels (tensor maps inside the graph)
    ("acoustic", dict(nz=40, nx=600, nshots=2, nt=80, rec_step=2)),               # register kernels, PML
    ("acoustic_vti_lsrtm_habc", dict(nz=60, nx=140, nshots=2, nt=72, rec_step=2)),
    ("elastic", dict(nz=50, nx=90, nshots=2, nt=80, rec_step=2)),
])
def test_replayed_time_loops_are_bit_identical(eq, kw, monkeypatch):
    from seistorch_b200 import _lib
    monkeypatch.setenv("SEISTORCH_B200_GRAPH", "1")          # opt-in (csrc/st_graph.cuh)
    monkeypatch.setenv("SEISTORCH_B200_PERSIST", "0")        # the persistent kernels are one launch already
    if eq == "acoustic_habc":
        monkeypatch.setenv("SEISTORCH_B200_TMA", "1")        # small grid: ask for the TMA kernels explicitly
    c0 = _lib.graph_counters()
    outs = _iterate(eq, 6, **kw)
    c1 = _lib.graph_counters()
    plain, captured, replayed = (b - a for a, b in zip(c0, c1))
    # torch's caching allocator hands out the same addresses from the second iteration on: forward and adjoint loop are each
    # captured once and replayed afterwards
    assert captured >= 2 and replayed >= 2, (plain, captured, replayed, _lib.graph_last_failure())
    r0, g0 = outs[0]
    assert max(np.abs(r).max() for r in r0) > 0 and max(np.abs(g).max() for g in g0) > 0
    for r, g in outs[1:]:
        assert all(np.array_equal(a, b) for a, b in zip(r, r0))
        assert all(np.array_equal(a, b) for a, b in zip(g, g0))


def test_short_loops_and_one_off_calls_stay_plain(monkeypatch):
    from seistorch_b200 import _lib
    monkeypatch.setenv("SEISTORCH_B200_GRAPH", "1")
    c0 = _lib.graph_counters()
    _iterate("acoustic_habc", 3, nz=60, nx=130, nshots=2, nt=40, rec_step=2)       # fewer steps than the graph threshold
    c1 = _lib.graph_counters()
    assert tuple(b - a for a, b in zip(c0, c1)) == (0, 0, 0)
    _iterate("acoustic_habc", 1, nz=60, nx=130, nshots=2, nt=70, rec_step=2)       # seen once: plain loop, nothing captured
    c2 = _lib.graph_counters()
    assert c2[1] == c1[1] and c2[2] == c1[2] and c2[0] - c1[0] == 2
    monkeypatch.setenv("SEISTORCH_B200_GRAPH", "0")          # default: the mechanism is off, nothing is even counted
    _iterate("acoustic_habc", 3, nz=60, nx=130, nshots=2, nt=70, rec_step=2)
    assert _lib.graph_counters() == c2
