"""L2, L1, cosine-similarity and envelope misfits with fused adjoint-source kernels -- same class surface as
seistorch/loss.py (Loss wrapper :23-50, L2 :409-421, L1 :381-393, CosineSimilarity :52-85, Envelope :178-216).

The forward value and d loss / d syn come from csrc/st_misfit.cu in one pass; inputs are
stacked records [B, nt, nrec, nchan] (``TensorList.stack()``) or lists of per-shot
records, like the reference.
"""
from __future__ import annotations


import numpy as np
import torch

from . import _lib
from .engine import LAUNCHES, _require_cuda, _stream_ptr, on_device_of

_HKER = {}


def hilbert_kernel(nt: int, device) -> torch.Tensor:
    """Imaginary part of ifft(h), h = one-sided spectrum filter of transform.py:46-53
    (scipy convention, nfft = nt): the analytic signal is x + i (hker (*) x)."""
    key = (nt, str(device))
    if key not in _HKER:
        h = np.zeros(nt, dtype=np.float64)
        if nt % 2 == 0:
            h[0] = h[nt // 2] = 1
            h[1:nt // 2] = 2
        else:
            h[0] = 1
            h[1:(nt + 1) // 2] = 2
        _HKER[key] = torch.from_numpy(np.fft.ifft(h).imag.astype(np.float32)).to(device)
    return _HKER[key]


class _Misfit(torch.autograd.Function):
    """loss = misfit(syn, obs) on [nt, ntraces] data; saves d loss / d syn."""

    @staticmethod
    @on_device_of(2)
    def forward(ctx, kind, syn, obs, mean_over=1):
        _require_cuda(syn, "synthetic record")
        s = syn.detach().to(torch.float32).contiguous()
        o = obs.detach().to(device=s.device, dtype=torch.float32).contiguous()
        if s.shape != o.shape:
            raise ValueError(f"syn {tuple(s.shape)} and obs {tuple(o.shape)} differ in shape")
        loss = torch.zeros(1, dtype=torch.float64, device=s.device)
        adj = torch.empty_like(s)
        L = _lib.lib()
        nt = s.shape[0]
        ntr = s.numel() // max(nt, 1)
        if kind == "l2":
            _lib.check(L.st_misfit_l2(s.data_ptr(), o.data_ptr(), s.numel(), 1.0, loss.data_ptr(), adj.data_ptr(),
                                      _stream_ptr()), "misfit_l2")
            LAUNCHES["misfit"] += 1
        elif kind == "l1":
            _lib.check(L.st_misfit_l1(s.data_ptr(), o.data_ptr(), s.numel(), 1.0, loss.data_ptr(), adj.data_ptr(),
                                      _stream_ptr()), "misfit_l1")
            LAUNCHES["misfit"] += 1
        elif kind == "sml1":
            _lib.check(L.st_misfit_sml1(s.data_ptr(), o.data_ptr(), s.numel(), 0.001, 1.0, loss.data_ptr(), adj.data_ptr(),
                                        _stream_ptr()), "misfit_sml1")
            LAUNCHES["misfit"] += 1
        elif kind == "cc":
            _lib.check(L.st_misfit_cc(s.data_ptr(), o.data_ptr(), s.numel(), 1.0, loss.data_ptr(), adj.data_ptr(),
                                      _stream_ptr()), "misfit_cc")
            LAUNCHES["misfit"] += 1
        elif kind == "integration":
            _lib.check(L.st_misfit_integration(s.data_ptr(), o.data_ptr(), nt, ntr, nt * int(mean_over), 1.0, loss.data_ptr(),
                                               adj.data_ptr(), _stream_ptr()), "misfit_integration")
            LAUNCHES["misfit"] += 1
        elif kind == "nim":
            _lib.check(L.st_misfit_nim(s.data_ptr(), o.data_ptr(), nt, ntr, 1.0, loss.data_ptr(), adj.data_ptr(),
                                       _stream_ptr()), "misfit_nim")
            LAUNCHES["misfit"] += 1
        elif kind == "traveltime":
            n = 2 * nt - 1
            lag = ((n - 1) * torch.linspace(0, 1, n, dtype=torch.float32)).to(s.device)     # signal.py:206-207
            _lib.check(L.st_misfit_traveltime(s.data_ptr(), o.data_ptr(), nt, ntr, lag.data_ptr(), int(mean_over), 1.0,
                                              loss.data_ptr(), adj.data_ptr(), _stream_ptr()), "misfit_traveltime")
            LAUNCHES["misfit"] += 1
        elif kind == "w1d":
            # loss.py:919-922: shift by 1.1 * min(min x, min y, 0), a constant for the gradient; stays on the device
            shift = (1.1 * torch.minimum(s.min(), o.min()).clamp(max=0.0)).to(torch.float32).reshape(1)
            _lib.check(L.st_misfit_w1d(s.data_ptr(), o.data_ptr(), nt, ntr, shift.data_ptr(), 1.0, loss.data_ptr(),
                                       adj.data_ptr(), _stream_ptr()), "misfit_w1d")
            LAUNCHES["misfit"] += 1
        elif kind == "cs":
            _lib.check(L.st_misfit_cs(s.data_ptr(), o.data_ptr(), nt, ntr, int(mean_over), 1.0, loss.data_ptr(),
                                      adj.data_ptr(), _stream_ptr()), "misfit_cs")
            LAUNCHES["misfit"] += 1
        else:
            ws = torch.empty(L.st_misfit_envelope_workspace(nt, ntr), dtype=torch.float32, device=s.device)
            hk = hilbert_kernel(nt, s.device)
            _lib.check(L.st_misfit_envelope(s.data_ptr(), o.data_ptr(), nt, ntr, hk.data_ptr(), 1.0, loss.data_ptr(),
                                            adj.data_ptr(), ws.data_ptr(), _stream_ptr()), "misfit_envelope")
            LAUNCHES["misfit"] += 5
        ctx.save_for_backward(adj)
        ctx.dtype = syn.dtype
        return loss[0].to(syn.dtype)

    @staticmethod
    def backward(ctx, g):
        (adj,) = ctx.saved_tensors
        return None, (adj * g).to(ctx.dtype), None, None


def _per_shot(kind, x, y):
    """Sum of the misfit over shots; each shot record is (nt, nrec, nchan)."""
    if isinstance(x, torch.Tensor) and x.ndim == 4 and isinstance(y, torch.Tensor):
        # stacked [B, nt, nrec, nchan] -> one launch over [nt, B*nrec*nchan]
        B, nt = x.shape[0], x.shape[1]
        xs = x.permute(1, 0, 2, 3).reshape(nt, -1)
        ys = y.permute(1, 0, 2, 3).reshape(nt, -1)
        return _Misfit.apply(kind, xs, ys, xs.shape[1] // max(B, 1))      # per-shot trace count (mean of "cs")
    loss = 0.0
    for _x, _y in zip(x, y):
        xs = _x.reshape(_x.shape[0], -1)
        loss = loss + _Misfit.apply(kind, xs, torch.as_tensor(_y).reshape(_x.shape[0], -1), xs.shape[1])
    return loss


class L2(torch.nn.Module):
    """loss.py:409-421: sum over shots of MSELoss(reduction='sum')."""

    @property
    def name(self):
        return "l2"

    def forward(self, x, y):
        return _per_shot("l2", x, y)


class L1(torch.nn.Module):
    """loss.py:381-393: sum over shots of L1Loss(reduction='sum')."""

    @property
    def name(self):
        return "l1"

    def forward(self, x, y):
        return _per_shot("l1", x, y)


class CosineSimilarity(torch.nn.Module):
    """loss.py:52-85 ("cs", normalised cross-correlation): sum over shots of mean over traces of
    1 - cosine_similarity(x_trace, y_trace) along time (eps = 1e-10)."""

    @property
    def name(self):
        return "cs"

    def forward(self, x, y):
        return _per_shot("cs", x, y)


class NormalizedIntegrationMethod(torch.nn.Module):
    """loss.py:463-501 ("nim", Donno et al.): traces squared, normalised by their sum over time, integrated;
    sum of squared differences.  Only the reference's defaults (criterion 'l2', reduction 'sum', method 'square')."""

    def __init__(self, criterion="l2", reduction="sum", method="square"):
        super().__init__()
        if (criterion, reduction, method) != ("l2", "sum", "square"):
            raise NotImplementedError("seistorch_b200: only the default 'nim' misfit (l2 / sum / square) is accelerated")

    @property
    def name(self):
        return "nim"

    def forward(self, x, y):
        return _per_shot("nim", x, y)


class SML1(torch.nn.Module):
    """loss.py:395-407: sum over shots of SmoothL1Loss(reduction='sum', beta=0.001)."""

    @property
    def name(self):
        return "sml1"

    def forward(self, x, y):
        return _per_shot("sml1", x, y)


class Crosscorrelation(torch.nn.Module):
    """loss.py:126-176 ("cc"): minus the zero-lag cross-correlation of every trace (all channels), summed."""

    @property
    def name(self):
        return "cc"

    def forward(self, x, y, win=512, step=1):
        return _per_shot("cc", x, y)


class Integration(torch.nn.Module):
    """loss.py:366-379: sum over shots of MSELoss() (mean) between the time integrals (cumsum) of the records."""

    @property
    def name(self):
        return "integration"

    def forward(self, x, y):
        return _per_shot("integration", x, y)


class Traveltime(torch.nn.Module):
    """loss.py:674-728: mean over shots x receivers x channels of the squared soft-argmax lag of the cross-correlation
    of the max-normalised traces (signal.py:203-208).  Like the reference it takes stacked records
    [shots, nt, nrec, nchan] (a list of equally shaped records is stacked)."""

    @property
    def name(self):
        return "traveltime"

    def forward(self, x, y):
        if not isinstance(x, torch.Tensor):
            x, y = torch.stack(list(x), 0), torch.stack([torch.as_tensor(v) for v in y], 0)
        nb, nt = x.shape[0], x.shape[1]
        xs = x.permute(1, 0, 2, 3).reshape(nt, -1)
        ys = y.to(x.device).permute(1, 0, 2, 3).reshape(nt, -1)
        return _Misfit.apply("traveltime", xs, ys, xs.shape[1])


class Wasserstein1d(torch.nn.Module):
    """loss.py:900-955 ("w1d") with the default method 'linear': traces shifted to be positive, normalised by their
    sum over time, integrated; sum of squared differences of the two cumulative distributions.  The shift is taken
    per shot, so stacked input is processed shot by shot."""

    def __init__(self, method="linear"):
        super().__init__()
        if method != "linear":
            raise NotImplementedError("seistorch_b200: only the default 'linear' w1d misfit is accelerated")
        self.method = method

    @property
    def name(self):
        return "w1d"

    def forward(self, x, y):
        if isinstance(x, torch.Tensor) and x.ndim == 4:
            x, y = list(x), list(y)
        return _per_shot("w1d", x, y)


class Envelope(torch.nn.Module):
    """loss.py:178-216 with method='square' (the only working method there):
    sum over shots of 0.5 * sum((E(x)^2 - E(y)^2)^2), E = |analytic signal| along time."""

    def __init__(self, method="square"):
        super().__init__()
        if method != "square":
            raise NotImplementedError("seistorch_b200: only the 'square' envelope misfit is accelerated")
        self.method = method

    @property
    def name(self):
        return "envelope"

    def forward(self, x, y):
        return _per_shot("envelope", x, y)


class Loss:
    """loss.py:23-50: look a misfit up by its ``name`` property."""

    def __init__(self, loss="l2"):
        self.loss_name = loss

    def __repr__(self):
        return f"Loss(loss={self.loss_name})"

    def __call__(self, *args, **kwargs):
        return self.loss(*args, **kwargs)

    def loss(self, cfg=None, *args, **kwargs):
        for cls in (L2, L1, SML1, CosineSimilarity, Crosscorrelation, Integration, NormalizedIntegrationMethod, Wasserstein1d, Traveltime, Envelope):
            if cls().name == self.loss_name:
                obj = cls(**kwargs)
                obj.cfg = cfg
                return obj
        raise ValueError(f"Cannot find loss named {self.loss_name}")
