"""Zero-phase Butterworth filtering of records on the device -- the `backend='torch'` branch of
seistorch/signal.py:49-101 (SeisSignal.filter), i.e. torchaudio's filtfilt in double precision with zero
initial state, applied along time to every shot of a TensorList between the forward modelling and the misfit.

The filter design stays on the host (scipy.signal.butter, as in the reference, signal.py:76); the two IIR
sweeps run in csrc/st_misfit.cu (one thread per trace).  The operator is A^T A (A = causal IIR matrix), so the
backward pass is the same kernel applied to the cotangent.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from .engine import _require_cuda, _stream_ptr, on_device_of


class _FiltFilt(torch.autograd.Function):
    @staticmethod
    @on_device_of(1)
    def forward(ctx, x, b, a):
        _require_cuda(x, "record")
        xs = x.detach().to(torch.float32).contiguous()
        nt = xs.shape[0]
        ntr = xs.numel() // max(nt, 1)
        y = torch.empty_like(xs)
        work = torch.empty(nt * ntr, dtype=torch.float64, device=xs.device)
        bb = np.ascontiguousarray(b, dtype=np.float64)
        aa = np.ascontiguousarray(a, dtype=np.float64)
        _lib.check(_lib.lib().st_filtfilt(xs.data_ptr(), y.data_ptr(), work.data_ptr(), nt, ntr,
                                          bb.ctypes.data_as(C.c_void_p), aa.ctypes.data_as(C.c_void_p), len(bb),
                                          _stream_ptr()), "filtfilt")
        ctx.b, ctx.a, ctx.dtype = bb, aa, x.dtype
        return y.to(x.dtype)

    @staticmethod
    def backward(ctx, g):
        return _FiltFilt.apply(g, ctx.b, ctx.a).to(ctx.dtype), None, None


def filtfilt(x: torch.Tensor, b, a) -> torch.Tensor:
    """Zero-phase filter of a record (nt, ...) along dim 0 with transfer function b / a."""
    return _FiltFilt.apply(x, b, a)


class SeisSignal:
    """signal.py:12-101, the filtering part."""

    def __init__(self, cfg=None, logger=None):
        self.cfg = cfg
        self.logger = logger
        self.dt = self.cfg["geom"]["dt"]
        self.forder = self.cfg["training"]["filter_ord"]

    def decide_filter_type(self, freq):
        """signal.py:23-37."""
        filter_mode = None
        if isinstance(freq, (int, float)):
            filter_mode = "lowpass"
        if isinstance(freq, list):
            if len(freq) == 1:
                filter_mode = "lowpass"
            if len(freq) == 2:
                filter_mode = "bandpass"
        if freq == "all":
            filter_mode = "all"
        return filter_mode

    def design(self, freqs):
        """signal.py:58-76: Butterworth coefficients for a cut-off (list of one) or a band (list of two), in Hz."""
        from scipy import signal
        filter_mode = self.decide_filter_type(freqs)
        assert filter_mode in ["lowpass", "highpass", "bandpass"], "mode must be lowpass, highpass or bandpass"
        if filter_mode in ["lowpass", "highpass"]:
            if isinstance(freqs, list):
                freqs = freqs[0]
            assert isinstance(freqs, (int, float)), "freqs must be a number for lowpass or highpass filter"
            freqs = [freqs]
        wn = [2 * f / (1 / self.dt) for f in list(freqs)]
        wn = wn[0] if len(wn) == 1 else wn
        return signal.butter(self.forder, Wn=wn, btype=filter_mode)

    def filter(self, d, freqs, axis=0, threads=1, backend="torch", **kwargs):
        """signal.py:49-101 with backend='torch': d is a TensorList of (nt, nrec, nchan) records (or one tensor
        stacked [shots, nt, nrec, nchan]); returns it filtered along time."""
        if self.logger is not None:
            self.logger.print(f"Data filtering (mode: {self.decide_filter_type(freqs)}): frequency:{freqs}")
        if freqs == "all":
            return d
        if backend != "torch":
            raise NotImplementedError("seistorch_b200: only the device ('torch') filtering backend is accelerated")
        b, a = self.design(freqs)
        if isinstance(d, torch.Tensor):
            return torch.stack([filtfilt(d[i], b, a) for i in range(d.shape[0])], 0)
        for i in range(d.shape[0]):
            d.data[i] = filtfilt(d.data[i], b, a)
        return d
