"""build_model -- same call contract as seistorch/model.py:23-93: YAML (or dict) config
-> geometry -> equation plug-in module selected by name -> WaveCell -> WaveRNN.

Differences: no torch.compile wrapping (the step functions are opaque custom ops; the
reference's ``SeisCompile`` would only wrap them lazily), models may be
passed in memory (``models=``) instead of ``.npy`` paths, and only the equations on the
accelerated path are accepted.
"""
from __future__ import annotations

import importlib
from typing import Dict, Optional

import numpy as np
import torch

from .cell import WaveCell
from .eqconfigure import Parameters, Wavefield
from .geom import WaveGeometryFreeForm
from .rnn import WaveRNN
from .utils import set_dtype


def update_cfg(cfg, models: Optional[Dict] = None, device="cuda"):
    """utils.py:282-329: padded Nx/Ny/Nz and ``domain_shape`` from the model shape."""
    g = cfg["geom"]
    arr = None
    if models:
        arr = np.asarray(next(iter(models.values())))
    else:
        for path in cfg["VEL_PATH"].values():
            if path is not None:
                arr = np.load(path)
                break
    if arr is None:
        raise ValueError("no model array/path given")
    pad, multiple = g["boundary"]["width"], g["multiple"]
    if arr.ndim == 2:
        ny, nx = arr.shape
        g["_oriNx"], g["_oriNy"], g["_oriNz"] = nx, ny, 0
        g["Nx"], g["Ny"], g["Nz"] = nx + 2 * pad, ny + (pad if multiple else 2 * pad), 0
        cfg["domain_shape"] = (g["Ny"], g["Nx"])
    else:
        nz, nx, ny = arr.shape          # utils.py:297: `nz, nx, ny = shape`
        g["_oriNx"], g["_oriNy"], g["_oriNz"] = nx, ny, nz
        g["Nx"], g["Ny"], g["Nz"] = nx + 2 * pad, ny + 2 * pad, nz + 2 * pad
        cfg["domain_shape"] = (g["Nz"], g["Nx"], g["Ny"])
    cfg["device"] = device
    return cfg


def check_config(cfg):
    """The hot-path subset of seistorch/default.py:3-140."""
    eq = cfg["equation"]
    if eq not in Parameters.valid_model_paras():
        raise ValueError(f"equation '{eq}' is not on the accelerated path")
    b = cfg["geom"]["boundary"]
    assert b["type"] in ["pml", "habc"], "boundary type should be 'pml' or 'habc'."
    assert b["width"] == 50, "Currently, the width of the boundary should be 50."       # default.py:45-46
    if b["type"] == "habc":
        assert "habc" in eq, "When boundary type is habc, the equation must be <...>_habc."
    names = Wavefield(eq).wavefields
    for st in cfg["geom"]["source_type"]:
        assert st in names, f"Valid source type are {names}, but got '{st}'."
    for rt in cfg["geom"]["receiver_type"]:
        if "lsrtm" in eq:
            assert rt.startswith("s"), "Receiver type should start with 's' in lsrtm equations."
        assert rt in names, f"Valid receiver type are {names}, but got '{rt}'."


def build_model(config, device="cuda", mode="forward", source_encoding=False, commands=None, logger=None,
                backend=None, models: Optional[Dict] = None):
    assert mode in ["forward", "inversion", "rtm"], f"No such mode {mode}!"
    if isinstance(config, str):
        from yaml import safe_load
        with open(config, "r") as f:
            cfg = safe_load(f)
    else:
        cfg = config
    cfg["VEL_PATH"] = cfg["geom"].get("initPath") if mode == "inversion" else cfg["geom"].get("truePath")
    cfg["geom"]["source_illumination"] = bool(getattr(commands, "source_illumination", False))
    check_config(cfg)
    if cfg.get("dtype", "float32") != "float32":
        raise ValueError("seistorch_b200: the sm_100a path computes in float32")
    set_dtype("float32")
    cfg = update_cfg(cfg, models, device=device)
    cfg["task"] = mode
    if cfg.get("seed") is not None:
        torch.manual_seed(cfg["seed"])
        np.random.seed(cfg["seed"])
    geom = WaveGeometryFreeForm(mode=mode, logger=logger, _models=models, **cfg)
    geom.inversion = mode == "inversion"
    module = importlib.import_module(f"{__package__}.equations{geom.ndim}d.{cfg['equation']}")
    backward_key = "_time_step_backward_multiple" if cfg["geom"]["multiple"] else "_time_step_backward"
    forward_func = getattr(module, "_time_step")
    backward_func = getattr(module, backward_key)
    cell = WaveCell(geom, forward_func, backward_func)
    model = WaveRNN(cell, source_encoding)
    model.to(device)
    return cfg, model


def model_from_case(case: Dict, device="cuda", mode="inversion", source_encoding=False):
    """Convenience used by tests/bench: build (cfg, WaveRNN) from an in-memory case dict
    (keys as in oracle/cases.py) and set its acquisition."""
    cfg = {
        "seed": 20230503, "dtype": "float32", "equation": case["equation"],
        "training": {"implicit": {"use": False}},
        "geom": {
            "multiple": bool(case.get("multiple", False)), "boundary_saving": False,
            "source_type": list(case["source_type"]), "receiver_type": list(case["receiver_type"]),
            "invlist": dict(case.get("invlist", {})), "dt": float(case["dt"]), "nt": int(case["nt"]),
            "h": float(case["h"]), "boundary": {"type": case["boundary"], "width": 50},
        },
    }
    cfg, model = build_model(cfg, device=device, mode=mode, source_encoding=source_encoding, models=case["models"])
    shots = list(range(len(case["sources"])))
    model.reset_geom(shots, case["sources"], case["receivers"], cfg)
    return cfg, model
