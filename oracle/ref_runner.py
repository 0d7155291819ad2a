"""TEST INFRASTRUCTURE ONLY -- drives the *real* reference (when /root/reference
is present) on CPU through its own public API, mirroring the call sequence of
seistorch_dist.py:92-258 (build_model -> reset_geom -> model(x) -> loss ->
backward).  Used to generate tests/golden/*.npz and to pin oracle/ against the
reference.  Never imported by the product path.
"""
from __future__ import annotations

import os
import pickle
import tempfile

import numpy as np
import torch
import yaml

from . import ref_shim

PARAM_KEYS = ["vp", "vs", "rho", "Q", "epsilon", "delta", "theta", "m", "rx", "rz"]


def build_reference(case: dict, dtype: str = "float32", want_grad: bool = True,
                    boundary_saving: bool = False, device: str = "cpu", source_encoding: bool = False):
    """Build the reference's own model for one case through its public API
    (seistorch/model.py:23-93 build_model + rnn.py reset_geom).  Returns (cfg, model, x)."""
    ref_shim.import_reference()
    from seistorch.model import build_model
    import seistorch.checkpoint_new as ckn
    import seistorch.checkpoint as ck

    tdtype = torch.float64 if dtype == "float64" else torch.float32
    with tempfile.TemporaryDirectory() as tmp:
        paths = {k: None for k in PARAM_KEYS}
        for name, arr in case["models"].items():
            p = os.path.join(tmp, f"{name}.npy")
            np.save(p, np.asarray(arr, dtype=np.float64 if dtype == "float64" else np.float32))
            paths[name] = p
        with open(os.path.join(tmp, "sources.pkl"), "wb") as f:
            pickle.dump(case["sources"], f)
        with open(os.path.join(tmp, "receivers.pkl"), "wb") as f:
            pickle.dump(case["receivers"], f)
        inv = {k: False for k in PARAM_KEYS}
        inv.update({k: bool(v) for k, v in case.get("invlist", {}).items()})
        cfg = {
            "seed": 20230503, "name": "oracle", "dtype": dtype,
            "equation": case["equation"],
            "training": {"implicit": {"use": False, "pretrained": None},
                         "minibatch": True, "batch_size": len(case["sources"]),
                         "N_epochs": 1, "lr": None, "scale_decay": 1.0,
                         "lr_decay": 1.0, "filter_ord": 3},
            "geom": {
                "obsPath": None, "truePath": dict(paths), "initPath": dict(paths),
                "sources": os.path.join(tmp, "sources.pkl"),
                "receivers": os.path.join(tmp, "receivers.pkl"),
                "wavelet": None, "multiple": bool(case.get("multiple", False)),
                "boundary_saving": bool(boundary_saving),
                "wavelet_delay": 0, "wavelet_inverse": False,
                "source_type": list(case["source_type"]),
                "receiver_type": list(case["receiver_type"]),
                "invlist": inv, "inv_savePath": None, "multiscale": ["all"],
                "dt": float(case["dt"]), "nt": int(case["nt"]), "fm": 10.0,
                "h": float(case["h"]), "Nshots": len(case["sources"]),
                "boundary": {"type": case["boundary"], "width": 50},
            },
        }
        cfg_path = os.path.join(tmp, "cfg.yml")
        with open(cfg_path, "w") as f:
            yaml.safe_dump(cfg, f)
        mode = "inversion" if want_grad else "forward"
        cfg2, model = build_model(cfg_path, device=device, mode=mode, source_encoding=source_encoding)
        if str(device) == "cpu":
            ref_shim.cast_module_kernels(tdtype, device="cpu")
        else:
            ref_shim.cast_module_kernels(tdtype)
        # reset class-level state of the BS checkpoint functions
        for mod in (ckn, ck):
            if hasattr(mod, "CheckpointFunction"):          # absent when seistorch_b200.overlay replaced the module
                mod.CheckpointFunction.counts = 0
                mod.CheckpointFunction.wavefields = []
        shots = list(range(len(case["sources"])))
        model.reset_geom(shots, case["sources"], case["receivers"], cfg2)
    x = torch.as_tensor(np.asarray(case["wavelet"]), dtype=tdtype).unsqueeze(0)
    return cfg2, model, x


def run_reference(case: dict, dtype: str = "float32", want_grad: bool = True,
                  boundary_saving: bool = False, loss_name: str = "l2",
                  threads: int | None = None):
    """Run one case through the reference.

    case keys: equation, models {name: ndarray (unpadded)}, invlist {name: bool},
    sources [[x, z] | [x, y, z]], receivers [[xs, zs(, ys)]] per shot, nt, dt, h,
    wavelet (nt,) ndarray, source_type [..], receiver_type [..], boundary 'pml'|'habc',
    multiple bool, obs: list of (nt, nrec, nchan) arrays or None (zeros).
    Returns dict(records=[...], loss=float, grads={name: ndarray}).
    """
    ref_shim.import_reference()
    from seistorch.loss import Loss

    if threads:
        torch.set_num_threads(threads)
    tdtype = torch.float64 if dtype == "float64" else torch.float32
    old_default = torch.get_default_dtype()
    try:
        cfg2, model, x = build_reference(case, dtype, want_grad, boundary_saving)
        if want_grad:
            model.train()
            syn = model(x)
        else:
            model.eval()
            with torch.no_grad():
                syn = model(x)
        out = {"records": [s.detach().numpy().copy() for s in syn]}
        if want_grad:
            obs = case.get("obs")
            if obs is None:
                obs = [np.zeros_like(r) for r in out["records"]]
            obs_t = [torch.as_tensor(o, dtype=tdtype) for o in obs]
            crit = Loss(loss_name).loss(cfg2)
            if len({tuple(t.shape) for t in syn}) == 1:
                loss = crit(torch.stack(list(syn), dim=0), torch.stack(obs_t, dim=0))
            else:  # ragged shots: L2.forward zips over shots (loss.py:417-421)
                loss = crit(list(syn), obs_t)
            loss.backward()
            out["loss"] = float(loss.item())
            out["grads"] = {}
            for name in model.cell.geom.model_parameters:
                p = getattr(model.cell.geom, name)
                if p.grad is not None:
                    out["grads"][name] = p.grad.detach().numpy().copy()
        return out
    finally:
        torch.set_default_dtype(old_default)


def run_reference_encoded(case: dict, wavelets, dtype: str = "float32", want_grad: bool = True, threads: int | None = None):
    """Source-encoded run through the reference, in the call order of codingfwi.py:88,129-132,240-261: model built with
    ``source_encoding=True``, probes set ONCE (first shot's receivers), all sources set, ``model(coding_wav)`` with one
    wavelet per source, L2 against zeros, backward.  Returns dict(records=[one array], loss, grads)."""
    ref_shim.import_reference()
    from seistorch.loss import Loss
    if threads:
        torch.set_num_threads(threads)
    tdtype = torch.float64 if dtype == "float64" else torch.float32
    old_default = torch.get_default_dtype()
    try:
        cfg2, model, _x = build_reference(case, dtype, want_grad, False, source_encoding=True)
        model.reset_probes(model.probes[0])
        x = torch.as_tensor(np.asarray(wavelets), dtype=tdtype)
        model.train() if want_grad else model.eval()
        with torch.set_grad_enabled(want_grad):
            syn = model(x)
        out = {"records": [s.detach().numpy().copy() for s in syn]}
        if want_grad:
            crit = Loss("l2").loss(cfg2)
            stacked = torch.stack(list(syn), dim=0)
            loss = crit(stacked, torch.zeros_like(stacked))
            loss.backward()
            out["loss"] = float(loss.item())
            out["grads"] = {n: getattr(model.cell.geom, n).grad.detach().numpy().copy()
                            for n in model.cell.geom.model_parameters if getattr(model.cell.geom, n).grad is not None}
        return out
    finally:
        torch.set_default_dtype(old_default)
