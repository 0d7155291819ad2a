"""WaveSource -- same constructor, buffers and attributes as seistorch/source.py.

On the whole-loop path (WaveRNN.forward) the source add is fused into the step kernel
(csrc/st_wave2d.cu ...); ``forward2d/forward3d`` below back the per-step compatibility
surface only and keep the reference's semantics (source.py:47-70): the ``dt`` argument
is ignored on the non-encoded 2D path, the field is cloned for second-order equations.
"""
from __future__ import annotations

import torch

from .utils import to_tensor


class WaveSource(torch.nn.Module):
    def __init__(self, bidx=None, second_order_equation=False, **kwargs):
        super().__init__()
        self._ndim = len(kwargs)
        self.coord_labels = list(kwargs.keys())
        for key, value in kwargs.items():
            value = None if value is None else to_tensor(value, dtype=torch.int64)
            self.register_buffer(key, value)
        self.bidx = bidx
        self.second_order_equation = second_order_equation
        self.forward = self.get_forward_func()
        self._source_encoding = False
        self.smask = None

    @property
    def ndim(self):
        return self._ndim

    @property
    def source_encoding(self):
        return self._source_encoding

    @source_encoding.setter
    def source_encoding(self, value):
        self._source_encoding = value

    def coords(self):
        """{'x': ..., 'y': ...[, 'z': ...]} (source.py:32-42)."""
        return dict(zip(self.coord_labels, [getattr(self, key) for key in self.coord_labels]))

    def get_forward_func(self):
        return getattr(self, f"forward{self.ndim}d")

    def forward2d(self, Y, X, dt=1.0):
        Y_new = Y.clone() if self.second_order_equation else Y
        if not self.source_encoding:
            Y_new += self.smask * X
        else:
            Y_new[..., self.y, self.x] += dt * X
        return Y_new

    def forward3d(self, Y, X, dt=1.0):
        Y_new = Y.clone()
        if not self.source_encoding:
            for idx in range(self.x.size(0)):
                Y_new[idx:idx + 1, self.x[idx]:self.x[idx] + 1, self.z[idx], self.y[idx]] += dt * X
        return Y_new
