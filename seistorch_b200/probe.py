"""WaveProbe -- same constructor, buffers and attributes as seistorch/probe.py.

On the whole-loop path the receiver gather is fused into the step kernel; the methods
below back the per-step compatibility surface (probe.py:42-48)."""
from __future__ import annotations

import torch

from .utils import to_tensor


class WaveProbe(torch.nn.Module):
    def __init__(self, batchidx=None, reccounts=None, **kwargs):
        super().__init__()
        self._ndim = len(kwargs)
        self.coord_labels = list(kwargs.keys())
        for key, value in kwargs.items():
            self.register_buffer(key, to_tensor(value, dtype=torch.int64))
        self.forward = self.get_forward_func()
        self.batchsize = self.x.size(0) if self.x.ndim > 1 else 1
        self.bidx = batchidx
        self.reccounts = [] if reccounts is None else reccounts   # used by WaveRNN to split records

    @property
    def ndim(self):
        return self._ndim

    def coords(self):
        return dict(zip(self.coord_labels, [getattr(self, key) for key in self.coord_labels]))

    def get_forward_func(self):
        return getattr(self, f"forward{self.ndim}d")

    def forward2d(self, x):
        return x[self.bidx, self.y, self.x]

    def forward3d(self, x):
        return x[self.bidx, self.x, self.z, self.y]


class WaveIntensityProbe(WaveProbe):
    def __init__(self, **kwargs):
        super().__init__(**kwargs)
