"""GPU (-m gpu): the REFERENCE's own code running on the sm_100a kernels through seistorch_b200.overlay.

The reference package travels to the GPU box as the byte-for-byte copy oracle/_ref (oracle/make_ref.py); each case runs
in a subprocess (tests/ref_on_kernels.py) so the overlaid package never mixes with a plain import of the reference.

  * golden: the reference's build_model (model.py:23-93) + reset_geom + model(x) + Loss(name).loss(cfg) + backward --
    the call sequence of seistorch_dist.py:92-258 -- against the committed golden vectors of the un-overlaid reference;
  * forward: the worker body of fwi.py:146-162 (forward modelling, no_grad, TensorList.numpy());
  * dist: the UNMODIFIED seistorch_dist.py executed as __main__ under torchrun (DistributedDataParallel over NCCL,
    DistributedSampler + DataLoader over the reference's OBSDataset on an h5py stand-in, SeisSignal.filter(backend='torch'),
    Loss, optimizer step); the gradient it saves is compared with the float64 oracle.
"""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, rel

pytestmark = pytest.mark.gpu
DRIVER = os.path.join(ROOT, "tests", "ref_on_kernels.py")


def _have_reference():
    from oracle import ref_shim
    return ref_shim.reference_available()


def _run(args, timeout=600):
    r = subprocess.run([sys.executable, DRIVER] + args, capture_output=True, text=True, timeout=timeout, cwd=ROOT)
    assert r.returncode == 0, f"{args}: rc={r.returncode}\n{r.stdout[-3000:]}\n{r.stderr[-6000:]}"
    line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
    return json.loads(line)


@pytest.mark.parametrize("name", ["acoustic", "acoustic_habc", "acoustic_habc_multiple", "elastic", "acoustic_tti_lsrtm_habc",
                                  "acoustic3d", "acoustic_envelope", "acoustic_habc_ragged", "elastic_l2_obs"])
def test_reference_build_model_runs_on_kernels(name):
    if not _have_reference():
        pytest.fail("oracle/_ref is missing: run `python -m oracle.make_ref` where /root/reference exists")
    if not os.path.exists(os.path.join(ROOT, "tests", "golden", name + ".npz")):
        pytest.skip(f"no golden fixture {name}")
    out = _run(["golden", name])
    assert out["overlaid"] >= 20
    assert out["loss_class"].startswith("seistorch_b200.loss."), out          # fused misfit kernel, not MSELoss / cuFFT
    assert out["rec_err"] < 1e-5, out
    assert out["loss_err"] < 2e-5, out
    assert out["grad_err"] and all(v < 1e-4 for v in out["grad_err"].values()), out


def test_reference_fwi_forward_worker_body():
    out = _run(["forward", "acoustic_habc"])
    assert out["rec_err"] < 1e-5, out


def test_model_on_a_device_that_is_not_current():
    """seistorch_dist.py:89-94 builds the model on cuda:{rank} and never calls set_device."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    out = _run(["golden", "acoustic_habc", "cuda:last"])
    assert out["current_device"] == 0 and out["device"] != "cuda:0"
    assert out["rec_err"] < 1e-5 and all(v < 1e-4 for v in out["grad_err"].values()), out


def test_unmodified_seistorch_dist_driver(tmp_path):
    import torch
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import ref_on_kernels as rk
    world = min(2, torch.cuda.device_count())
    work = str(tmp_path / "dist")
    rk.prepare_dist(work)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29731", DRIVER, "dist", work]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, f"rc={r.returncode}\n{r.stdout[-3000:]}\n{r.stderr[-8000:]}"
    res = os.path.join(work, "results")
    g = torch.load(os.path.join(res, "grad_vp_nosm_0.pt"), map_location="cpu").numpy()
    expect, loss = rk.expected_dist_gradient(world)
    assert g.shape == expect.shape
    assert rel(g, expect) < 1e-4, rel(g, expect)
    assert os.path.exists(os.path.join(res, "model_0.pt"))
    # the records the driver dumps (seistorch_dist.py:252-253) are the filtered synthetics of rank 0's shots
    syn = np.load(os.path.join(res, "syn0.npy"))
    assert syn.shape[0] == 4 // world and np.isfinite(syn).all() and np.abs(syn).max() > 0
