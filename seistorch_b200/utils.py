"""Helpers with reference semantics: seistorch/utils.py:235-280."""
from __future__ import annotations

import numpy as np
import torch


def ricker_wave(fm, dt, T, delay=80, dtype="tensor", inverse=False):
    """utils.py:235-249."""
    i = np.arange(T)
    c = np.pi * fm * (i * dt - delay * dt)
    w = (-1 if inverse else 1) * (1 - 2 * np.power(c, 2)) * np.exp(-np.power(c, 2))
    w = w.astype(np.float32)
    return w if dtype == "numpy" else torch.from_numpy(w)


def set_dtype(dtype=None):
    """utils.py:251-257."""
    if dtype in (None, "float32"):
        torch.set_default_dtype(torch.float32)
    elif dtype == "float64":
        torch.set_default_dtype(torch.float64)
    else:
        raise ValueError("Unsupported data type: %s; should be either float32 or float64" % dtype)


def to_tensor(x, dtype=None):
    """utils.py:259-280: python floats become a default-dtype tensor first, lists go
    through numpy (float64); ``.type(int64)`` then truncates toward zero."""
    dtype = dtype if dtype is not None else torch.get_default_dtype()
    if "numpy" in str(type(x)):
        x = np.asarray(x)
    if isinstance(x, np.ndarray):
        return torch.from_numpy(x).type(dtype)
    if isinstance(x, (float, int)):
        return torch.tensor(x).type(dtype)
    if isinstance(x, list):
        if None in x:
            return torch.Tensor([])
        items = [i.cpu().numpy() if hasattr(i, "device") else i for i in x]
        return torch.from_numpy(np.array(items)).type(dtype)
    return torch.from_numpy(x.cpu().numpy()).type(dtype)
