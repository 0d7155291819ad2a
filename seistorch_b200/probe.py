"""WaveProbe -- constructor, buffers and attributes of seistorch/probe.py.

On the whole-loop path the receiver gather is fused into the step kernels; ``forward2d`` /
``forward3d`` serve step-by-step callers with the reference's indexing (probe.py:42-48):
``field[bidx, y, x]`` in 2D and ``field[bidx, x, z, y]`` in 3D, receivers of all shots
concatenated, ``bidx`` = shot of every receiver.
"""
from __future__ import annotations

from .points import GridPoints


class WaveProbe(GridPoints):
    def __init__(self, batchidx=None, reccounts=None, **kwargs):
        super().__init__(False, **kwargs)
        self.bidx = batchidx
        self.batchsize = self.x.size(0) if self.x.ndim > 1 else 1
        # receivers per shot: WaveRNN splits the stacked records with it (rnn.py:211)
        self.reccounts = list(reccounts) if reccounts is not None else []

    def forward2d(self, x):
        return x[self._field_index(self.bidx)]

    def forward3d(self, x):
        return x[self._field_index(self.bidx)]


class WaveIntensityProbe(WaveProbe):
    """Same sampling as WaveProbe (the reference's squared-intensity variant is commented out,
    probe.py:50-56)."""
