"""WaveSource -- constructor, buffers and attributes of seistorch/source.py.

The whole-loop path (WaveRNN.forward) fuses the source add into the step kernels; the two
methods below only serve code that drives the cell step by step.  Reference behaviour kept
(source.py:47-70): second-order equations get a fresh field (out of place), first-order ones are
updated in place; without encoding the wavelet sample is spread by the one-hot ``smask`` built in
rnn.py:160-166 and the ``dt`` argument is ignored in 2D; with encoding every source adds its own
sample (scaled by ``dt``) into the single shared wavefield.
"""
from __future__ import annotations

from .points import GridPoints


class WaveSource(GridPoints):
    def __init__(self, bidx=None, second_order_equation=False, **kwargs):
        super().__init__(True, **kwargs)
        self.bidx = bidx
        self.second_order_equation = second_order_equation
        self.smask = None
        self._source_encoding = False

    @property
    def source_encoding(self):
        return self._source_encoding

    @source_encoding.setter
    def source_encoding(self, value):
        self._source_encoding = value

    def forward2d(self, Y, X, dt=1.0):
        out = Y.clone() if self.second_order_equation else Y
        if self._source_encoding:
            where = (Ellipsis,) + self._field_index()
            out[where] += dt * X
        else:
            out += self.smask * X
        return out

    def forward3d(self, Y, X, dt=1.0):
        out = Y.clone()
        if self._source_encoding:
            return out                       # the reference adds nothing in this mode (source.py:66-68)
        for shot in range(self.x.size(0)):
            i0, i1, i2 = (int(c[shot]) for c in (self.x, self.z, self.y))
            out[shot:shot + 1, i0:i0 + 1, i1, i2] += dt * X
        return out
