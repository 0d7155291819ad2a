"""Shared base of WaveSource / WaveProbe: a set of grid points held as int64 index buffers.

Both reference classes (seistorch/source.py, seistorch/probe.py) take their coordinates as keyword
arguments (``x=..., y=...[, z=...]``), register one int64 buffer per coordinate, expose
``coords()`` / ``ndim`` and pick ``forward2d`` or ``forward3d`` by the number of coordinates.
That contract lives here once.
"""
from __future__ import annotations

import torch

from .utils import to_tensor


class GridPoints(torch.nn.Module):
    def __init__(self, allow_none: bool, **coords):
        super().__init__()
        self.coord_labels = list(coords)
        self._ndim = len(self.coord_labels)
        for label, value in coords.items():
            index = None if (allow_none and value is None) else to_tensor(value, dtype=torch.int64)
            self.register_buffer(label, index)
        # reference code replaces .forward with the dimension-specific method at construction
        self.forward = self.get_forward_func()

    @property
    def ndim(self):
        return self._ndim

    def coords(self):
        """Mapping coordinate label -> int64 index tensor, in keyword order."""
        return {label: getattr(self, label) for label in self.coord_labels}

    def get_forward_func(self):
        return getattr(self, "forward%dd" % self._ndim)

    def _field_index(self, lead=None):
        """Index tuple into a field laid out (B, y, x) in 2D or (B, x, z, y) in 3D
        (rnn.py:164-166, probe.py:44,48)."""
        tail = (self.y, self.x) if self._ndim == 2 else (self.x, self.z, self.y)
        return tail if lead is None else (lead,) + tail
