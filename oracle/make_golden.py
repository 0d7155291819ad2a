"""TEST INFRASTRUCTURE ONLY.  Generates tests/golden/*.npz by running the REAL
reference (/root/reference, imported through oracle/ref_shim.py) on CPU.

    python -m oracle.make_golden            # all cases
    python -m oracle.make_golden acoustic   # cases whose name contains 'acoustic'

Each fixture stores the full case (models, geometry, wavelet) next to the
reference outputs: seismograms and pure-AD gradients (boundary_saving: false,
SURVEY.md 8c) in fp32 and in fp64.  The reference has no golden vectors of its
own; these files are what pins parity.  Recorded with torch
{torch.__version__}, 8 threads, CPU eager.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

from . import cases
from .ref_runner import run_reference

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def pack_case(case):
    meta = {k: case[k] for k in ("equation", "invlist", "sources", "receivers", "nt", "dt", "h",
                                 "source_type", "receiver_type", "boundary", "multiple")}
    arrs = {"wavelet": np.asarray(case["wavelet"], np.float32)}
    for k, v in case["models"].items():
        arrs["model_" + k] = np.asarray(v, np.float32)
    if case.get("obs") is not None:
        for i, o in enumerate(case["obs"]):
            arrs[f"obs_{i}"] = np.asarray(o, np.float32)
    arrs["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    return arrs


def unpack_case(npz):
    meta = json.loads(bytes(npz["meta"]).decode())
    case = dict(meta)
    case["wavelet"] = npz["wavelet"]
    case["models"] = {k[6:]: npz[k] for k in npz.files if k.startswith("model_")}
    obs = [npz[k] for k in sorted((k for k in npz.files if k.startswith("obs_")),
                                  key=lambda s: int(s[4:]))]
    case["obs"] = obs or None
    return case


def golden_cases():
    g = {}
    small = dict(nz=30, nx=44, nt=120)
    for eq in ["acoustic", "acoustic_habc", "elastic", "vti_habc2", "tti_habc",
               "acoustic_vti_lsrtm_habc", "acoustic_tti_lsrtm_habc", "acoustic_fwim_habc", "acoustic_lsrtm_habc", "acoustic_rho_habc"]:
        g[eq] = (cases.make_case(eq, **small), "l2")
    g["acoustic_multiple"] = (cases.make_case("acoustic", multiple=True, **small), "l2")
    g["acoustic_habc_multiple"] = (cases.make_case("acoustic_habc", multiple=True, **small), "l2")
    g["acoustic3d"] = (cases.make_case("acoustic", nz=12, nx=14, ny=10, nt=40, nshots=2), "l2")
    # ragged receivers + one shot with a single receiver
    c = cases.make_case("acoustic_habc", nz=24, nx=37, nt=90, nshots=3)
    c["receivers"] = [c["receivers"][0], [[3.99, 11.2], [2.0, 2.5]], [[5], [1]]]
    g["acoustic_habc_ragged"] = (c, "l2")
    # long horizon (reference fp32 noise dominates here)
    g["acoustic_long"] = (cases.make_case("acoustic", nz=40, nx=64, nt=1000, nshots=1, rec_step=4), "l2")
    g["acoustic_habc_long"] = (cases.make_case("acoustic_habc", nz=40, nx=64, nt=1000, nshots=1, rec_step=4), "l2")
    g["elastic_long"] = (cases.make_case("elastic", nz=40, nx=64, nt=1000, nshots=1, rec_step=4), "l2")
    return g


def main(argv):
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(8)
    gc = golden_cases()
    # envelope misfit case: observed data = perturbed synthetic
    env = cases.make_case("acoustic", nz=30, nx=44, nt=120)
    ref0 = run_reference(env, want_grad=False)
    rng = np.random.default_rng(7)
    env["obs"] = [(r * 0.8 + 0.01 * np.abs(r).max() * rng.standard_normal(r.shape)).astype(np.float32)
                  for r in ref0["records"]]
    gc["acoustic_envelope"] = (env, "envelope")
    l2o = cases.make_case("elastic", nz=30, nx=44, nt=120)
    ref0 = run_reference(l2o, want_grad=False)
    l2o["obs"] = [(r * 0.7).astype(np.float32) for r in ref0["records"]]
    gc["elastic_l2_obs"] = (l2o, "l2")
    # source encoding (codingfwi.py): 4 sources into one wavefield, one wavelet (random polarity / scale) per source
    if not argv or any(a in "acoustic_habc_encoded" for a in argv):
        from .ref_runner import run_reference_encoded
        enc = cases.make_case("acoustic_habc", nz=30, nx=44, nt=120, nshots=4)
        rngw = np.random.default_rng(11)
        wavs = np.stack([np.asarray(enc["wavelet"]) * s for s in (1.0, -1.0, 0.5, -2.0)]).astype(np.float32)
        arrs = pack_case(enc)
        arrs["enc_wavelets"] = wavs
        arrs["loss_name"] = np.frombuffer(b"l2", dtype=np.uint8)
        for dt in ("float32", "float64"):
            out = run_reference_encoded(enc, wavs, dtype=dt)
            tag = "f32" if dt == "float32" else "f64"
            arrs[f"{tag}_rec_0"] = out["records"][0]
            arrs[f"{tag}_loss"] = np.float64(out["loss"])
            for k, v in out["grads"].items():
                arrs[f"{tag}_grad_{k}"] = v
        np.savez_compressed(os.path.join(OUT, "acoustic_habc_encoded.npz"), **arrs)
        print("wrote acoustic_habc_encoded", flush=True)
    for name, (case, loss) in gc.items():
        if argv and not any(a in name for a in argv):
            continue
        arrs = pack_case(case)
        arrs["loss_name"] = np.frombuffer(loss.encode(), dtype=np.uint8)
        for dt in ("float32", "float64"):
            out = run_reference(case, dtype=dt, want_grad=True, boundary_saving=False, loss_name=loss)
            tag = "f32" if dt == "float32" else "f64"
            for i, r in enumerate(out["records"]):
                arrs[f"{tag}_rec_{i}"] = r
            arrs[f"{tag}_loss"] = np.float64(out["loss"])
            for k, v in out["grads"].items():
                arrs[f"{tag}_grad_{k}"] = v
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **arrs)
        print("wrote", name, {k: v.shape for k, v in arrs.items() if k.startswith("f64_grad")}, flush=True)


if __name__ == "__main__":
    main(sys.argv[1:])
