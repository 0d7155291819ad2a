"""CPU: the oracle (oracle/*.py, torch restatement) against the committed golden vectors,
which were produced by the REAL reference (oracle/make_golden.py).  fp64 must agree to
rounding; fp32 is bit-close because the oracle mirrors the reference's op order."""
import numpy as np
import pytest
import torch

from conftest import cat_records, golden_records, load_golden, rel
from oracle import loop, misfit

SMALL = ["acoustic", "acoustic_habc", "elastic", "vti_habc2", "tti_habc", "acoustic_lsrtm_habc", "acoustic_rho_habc", "acoustic_vti_lsrtm_habc",
         "acoustic_tti_lsrtm_habc", "acoustic_fwim_habc", "acoustic_multiple", "acoustic_habc_multiple",
         "acoustic_habc_ragged", "acoustic_envelope", "elastic_l2_obs", "acoustic3d"]


@pytest.mark.parametrize("name", SMALL)
def test_oracle_matches_reference_fp64(name):
    z, case = load_golden(name)
    loss_name = bytes(z["loss_name"]).decode()
    inv = [k for k, v in case["invlist"].items() if v]
    if name == "acoustic3d":
        case = dict(case, nt=12)     # keep the CPU suite short: compare the first 12 samples only
    recs, params = loop.simulate(case, dtype=torch.float64, requires_grad=inv)
    ref = golden_records(z, "f64")
    if name == "acoustic3d":
        ref = [r[:12] for r in ref]
        assert rel(cat_records([r.detach().numpy() for r in recs]), cat_records(ref)) < 1e-12
        return
    assert rel(cat_records([r.detach().numpy() for r in recs]), cat_records(ref)) < 1e-12
    obs = case["obs"] or [np.zeros_like(r) for r in ref]
    obs = [torch.as_tensor(o, dtype=torch.float64) for o in obs]
    loss = (misfit.envelope_loss if loss_name == "envelope" else misfit.l2)(recs, obs)
    loss.backward()
    assert abs(float(loss) - float(z["f64_loss"])) <= 1e-10 * abs(float(z["f64_loss"]))
    for k in inv:
        assert rel(params[k].grad.numpy(), z[f"f64_grad_{k}"]) < 1e-9, k


@pytest.mark.parametrize("name", ["acoustic", "acoustic_habc", "elastic", "acoustic_fwim_habc"])
def test_oracle_matches_reference_fp32_bitwise(name):
    z, case = load_golden(name)
    recs, _ = loop.simulate(case, dtype=torch.float32)
    ref = golden_records(z, "f32")
    for a, b in zip(recs, ref):
        assert np.array_equal(a.numpy(), b)


def test_index_semantics_known_answers():
    """SURVEY 8a [probed on the reference]: source [10.7, 3.2] -> cell (x=60, y=53);
    receivers x=[10.7, 11.9, 9.0], z=[3.2, 3.0, 3.99] -> x=[60,61,59], y=[53,53,53]."""
    src = loop.source_indices([[10.7, 3.2]])
    assert src.tolist() == [[60, 53]]
    bidx, rec, counts = loop.receiver_indices([[[10.7, 11.9, 9.0], [3.2, 3.0, 3.99]]])
    assert rec.tolist() == [[60, 61, 59], [53, 53, 53]] and counts == [3] and bidx.tolist() == [0, 0, 0]
    assert loop.source_indices([[10.7, 3.2]], multiple=True).tolist() == [[60, 3]]


def test_impulse_response_timing():
    """rnn.py:183-202: the source is added after the step-0 update and before sampling, so a
    unit spike appears at sample 0 at the co-located receiver and (c dt/h)^2 at sample 1 next to it."""
    vp = np.full((20, 20), 1500.0, np.float32)
    w = np.zeros(5, np.float32)
    w[0] = 1.0
    case = dict(equation="acoustic", models={"vp": vp}, invlist={}, sources=[[10, 10]],
                receivers=[[[10, 11], [10, 10]]], nt=5, dt=1e-3, h=10.0, wavelet=w,
                source_type=["h1"], receiver_type=["h1"], boundary="pml")
    recs, _ = loop.simulate(case, dtype=torch.float64)
    r = recs[0].numpy()[:, :, 0]
    assert r[0, 0] == 1.0 and r[0, 1] == 0.0
    assert abs(r[1, 1] - 0.0225) < 1e-12


def test_oracle_source_encoding_matches_reference():
    """codingfwi.py mode (source.py:54-55, rnn.py:113,162): every source fires into ONE wavefield with its own wavelet.
    The golden vectors come from the real reference driven like codingfwi.py:88,129-132,240-261
    (oracle/ref_runner.run_reference_encoded); fp32 bit for bit, fp64 to rounding."""
    z, case = load_golden("acoustic_habc_encoded")
    w = torch.as_tensor(z["enc_wavelets"])
    recs, _ = loop.simulate(case, dtype=torch.float32, wavelet=w, source_encoding=True)
    assert len(recs) == 1 and np.array_equal(recs[0].numpy(), z["f32_rec_0"])
    recs, params = loop.simulate(case, dtype=torch.float64, requires_grad=["vp"], wavelet=w, source_encoding=True)
    assert rel(recs[0].detach().numpy(), z["f64_rec_0"]) < 1e-12
    loss = misfit.l2(recs, [torch.zeros_like(r) for r in recs])
    loss.backward()
    assert abs(float(loss.detach()) - float(z["f64_loss"])) <= 1e-10 * abs(float(z["f64_loss"]))
    assert rel(params["vp"].grad.numpy(), z["f64_grad_vp"]) < 1e-9
