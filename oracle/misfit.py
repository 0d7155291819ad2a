"""TEST INFRASTRUCTURE ONLY (oracle).  CPU (torch) restatement of the reference's
L2 and envelope misfits.  Never imported by the product path.

Follows:
  * seistorch/loss.py:409-421      L2  (MSELoss(reduction='sum') summed over shots)
  * seistorch/loss.py:178-216      Envelope (method='square')
  * seistorch/loss.py:381-393      L1
  * seistorch/loss.py:52-85        CosineSimilarity ("cs")
  * seistorch/loss.py:463-501      NormalizedIntegrationMethod ("nim", defaults)
  * seistorch/loss.py:900-955      Wasserstein1d ("w1d", method 'linear')
  * seistorch/loss.py:674-728      Traveltime (+ signal.py:203-208)
  * seistorch/transform.py:24-66   envelope / hilbert (nfft = nt, scipy convention)
"""
from __future__ import annotations

import numpy as np
import torch


def l2(syn, obs):
    """loss.py:417-421."""
    loss = 0.0
    for x, y in zip(syn, obs):
        loss = loss + torch.sum((x - y) ** 2)
    return loss


def l1(syn, obs):
    """loss.py:389-393: L1Loss(reduction='sum') summed over shots."""
    loss = 0.0
    for x, y in zip(syn, obs):
        loss = loss + torch.sum(torch.abs(x - y))
    return loss


def cs(syn, obs):
    """loss.py:63-85: per shot mean over traces of 1 - cosine_similarity along time (eps = 1e-10)."""
    loss = 0.0
    for x, y in zip(syn, obs):
        nt = x.shape[0]
        xr, yr = x.reshape(nt, -1), y.reshape(nt, -1)
        nx = torch.clamp(torch.linalg.vector_norm(xr, dim=0), min=1e-10)
        ny = torch.clamp(torch.linalg.vector_norm(yr, dim=0), min=1e-10)
        sim = torch.sum((xr / nx) * (yr / ny), dim=0)
        loss = loss + torch.mean(1 - sim)
    return loss


def nim(syn, obs):
    """loss.py:476-501 with the defaults (criterion 'l2', reduction 'sum', method 'square')."""
    loss = 0.0
    for x, y in zip(syn, obs):
        x, y = x ** 2, y ** 2                                  # transform.both_nonnegative(type='square')
        x = x / torch.sum(x, dim=0, keepdim=True)
        y = y / torch.sum(y, dim=0, keepdim=True)
        x, y = torch.cumsum(x, dim=0), torch.cumsum(y, dim=0)
        loss = loss + torch.sum((x - y) ** 2)
    return loss


def w1d(syn, obs):
    """loss.py:940-955 with method 'linear' (:917-922)."""
    loss = 0.0
    for x, y in zip(syn, obs):
        m = torch.min(x.detach().min(), y.detach().min())
        m = m if m < 0 else 0
        x, y = x - 1.1 * m, y - 1.1 * m
        x = x / (torch.sum(x, dim=0, keepdim=True) + 1e-18)
        y = y / (torch.sum(y, dim=0, keepdim=True) + 1e-18)
        loss = loss + (torch.abs(torch.cumsum(x, dim=0) - torch.cumsum(y, dim=0)) ** 2).sum()
    return loss


def sml1(syn, obs):
    """loss.py:403-407."""
    loss = 0.0
    for x, y in zip(syn, obs):
        loss = loss + torch.nn.SmoothL1Loss(reduction="sum", beta=0.001)(x, y)
    return loss


def cc(syn, obs):
    """loss.py:148-161: conv1d of a trace with the full-length other trace = its zero-lag cross-correlation."""
    loss = 0.0
    for x, y in zip(syn, obs):
        loss = loss - torch.sum(x * y)
    return loss


def integration(syn, obs):
    """loss.py:375-379 with transform.integrate (cumsum along time), MSELoss() = mean."""
    loss = 0.0
    for x, y in zip(syn, obs):
        loss = loss + torch.mean((torch.cumsum(x, dim=0) - torch.cumsum(y, dim=0)) ** 2)
    return loss


def traveltime(syn, obs):
    """loss.py:683-728 with signal.py:203-208 (soft argmax of the cross-correlation); stacked [nb, nt, nr, nc]."""
    x, y = torch.stack(list(syn), 0), torch.stack(list(obs), 0)
    nb, nt, nr, nc = x.shape
    x = x / (torch.max(torch.abs(x), dim=1, keepdim=True).values + 1e-16)
    y = y / (torch.max(torch.abs(y), dim=1, keepdim=True).values + 1e-16)
    x = x.permute(0, 2, 3, 1).contiguous().view(nb * nr * nc, nt)
    y = y.permute(0, 2, 3, 1).contiguous().view(nb * nr * nc, 1, nt)
    cc = torch.nn.functional.conv1d(x, y, padding=nt - 1, groups=nb * nr * nc).view(nb, nr, nc, -1)
    n = cc.shape[-1]
    p = torch.softmax(1 * cc, dim=-1)
    idx = torch.linspace(0, 1, n)                           # fp32, as in the reference
    tau = torch.sum((n - 1) * p * idx, dim=-1) - nt + 1
    return (tau.view(nb, 1, nr, nc) ** 2).mean()


def hilbert(data):
    """transform.py:27-66: analytic signal along dim 0 of (nt, ntraces, nchan)."""
    nt = data.shape[0]
    spec = torch.fft.fft(data, n=nt, dim=0)
    hfilt = np.zeros(nt, dtype=np.float32)
    if nt % 2 == 0:
        hfilt[0] = hfilt[nt // 2] = 1
        hfilt[1:nt // 2] = 2
    else:
        hfilt[0] = 1
        hfilt[1:(nt + 1) // 2] = 2
    hfilt = torch.from_numpy(hfilt).view(-1, 1, 1)
    return torch.fft.ifft(spec * hfilt, dim=0)


def envelope(d):
    """transform.py:24-25."""
    return torch.abs(hilbert(d))


def envelope_loss(syn, obs):
    """loss.py:201-216 with method='square':  sum_shots 0.5*sum((E(x)^2-E(y)^2)^2)."""
    loss = 0.0
    for x, y in zip(syn, obs):
        loss = loss + 0.5 * torch.sum((envelope(x) ** 2 - envelope(y) ** 2) ** 2)
    return loss
