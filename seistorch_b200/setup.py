"""Acquisition set-up with the reference's index semantics
(seistorch/setup.py:261-274, 386-415, 452-483)."""
from __future__ import annotations

from .probe import WaveIntensityProbe
from .source import WaveSource


def setup_src_coords(coords, bwidth, multiple=False):
    """setup.py:452-483: add the boundary width (python float), truncated to int64 by
    WaveSource; 2D `multiple` removes the (absent) top pad from the depth."""
    keys = ["x", "y", "z"]
    kwargs = dict()
    for key, value in zip(keys, coords):
        assert type(value) in [int, float, type(None)], f"The source location must be a number, got {type(value)}."
        kwargs[key] = value + bwidth if isinstance(value, (int, float)) else value
    if "z" not in kwargs and multiple and bool(kwargs["y"]):
        kwargs["y"] -= bwidth
    if "z" in kwargs and multiple:
        raise NotImplementedError("Multiples in 3D case is not implemented yet.")
    return WaveSource(**kwargs)


def setup_rec_coords(coords, bwidth, multiple=False):
    """setup.py:386-415."""
    keys = ["x", "y", "z"]
    kwargs = dict()
    for key, value in zip(keys, coords):
        kwargs[key] = [v + bwidth if v is not None else None for v in value]
    if "z" not in kwargs and multiple:
        kwargs["y"] = [v - bwidth if v is not None else None for v in kwargs["y"]]
    if "z" in kwargs and multiple:
        raise NotImplementedError("Multiples in 3D case is not implemented yet.")
    return [WaveIntensityProbe(**kwargs)]


def setup_acquisition(shots, src_list, rec_list, cfg, *args, **kwargs):
    """setup.py:261-274."""
    bwidth = cfg["geom"]["boundary"]["width"]
    multiple = cfg["geom"]["multiple"]
    sources, receivers = [], []
    for shot in shots:
        sources.append(setup_src_coords(src_list[shot], bwidth, multiple))
        receivers.extend(setup_rec_coords(rec_list[shot], bwidth, multiple))
    return sources, receivers
