"""Gradient post-processing on the device -- same class surface as seistorch/process.py:9-117 (PostProcess).

The reference's ``smooth_gradient`` round-trips every parameter gradient through host numpy (process.py:66-112,
signal.py:247-319); here the truncated-Gaussian passes run in csrc/st_postproc.cu on the gradient where it lives.
``cut_gradient`` / ``precondition`` are element-wise products the reference already does on the device."""
from __future__ import annotations


import numpy as np
import torch
from torch.nn.parallel import DistributedDataParallel

from . import _lib
from .engine import LAUNCHES, _require_cuda, _stream_ptr


def gaussian_weights(sigma: float, radius: int) -> np.ndarray:
    """signal.py:263-275: kernel size 2*radius+1 for even radii (the even kernel the reference builds for odd radii
    returns an array one element longer than its input and cannot be assigned back to the parameter)."""
    radius = int(radius)
    if radius % 2 != 0:
        raise ValueError("seistorch_b200: gradient smoothing needs an even `radius` (the reference's kernel for odd radii "
                         "has an even length and changes the gradient's shape, signal.py:263-266)")
    k = 2 * radius + 1
    w = np.exp(-(np.arange(k) - k // 2) ** 2 / (2 * float(sigma) ** 2)).astype(np.float32)   # float32 like torch.tensor(..., float32)
    return (torch.from_numpy(w) / torch.from_numpy(w).sum()).numpy()


def gaussian_filter(grad: torch.Tensor, sigma: float, radius: int, axis: int) -> torch.Tensor:
    """One pass of signal.gaussian_filter (2D input, reflect boundaries) on the device."""
    _require_cuda(grad, "gradient")
    if grad.ndim != 2:
        raise NotImplementedError("3D smoothing not implemented")                       # process.py:93-95
    x = grad.detach().to(torch.float32).contiguous()
    w = torch.from_numpy(gaussian_weights(sigma, radius)).to(x.device)
    out = torch.empty_like(x)
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().st_gaussian_smooth2d(x.data_ptr(), out.data_ptr(), x.shape[0], x.shape[1], w.data_ptr(),
                                                   int(radius), int(axis), _stream_ptr()), "gaussian_smooth2d")
    LAUNCHES["misfit"] += 1
    return out.to(grad.dtype)


class PostProcess:
    """process.py:9-117."""

    def __init__(self, model, cfg, commands=None):
        self.model = model.module if isinstance(model, DistributedDataParallel) else model
        self.cfg = cfg
        self.commands = commands
        self.ndim = self.model.cell.geom.ndim
        if getattr(self.commands, "grad_cut", False):
            self.modelmask = self.load_seabed()

    def load_seabed(self):
        """process.py:27-41: the seabed mask padded like the model (zeros in the absorbing frame)."""
        padding = self.cfg["geom"]["boundary"]["width"]
        top = 0 if self.cfg["geom"]["multiple"] else padding
        seabed = torch.from_numpy(np.load(self.cfg["geom"]["seabed"])).float()
        pads = (padding, padding, top, padding) if self.ndim == 2 else (padding, padding, top, padding, padding, padding)
        return torch.nn.functional.pad(seabed, pads, mode="constant", value=0)

    def cut_gradient(self):
        for para in self.model.parameters():
            if para.requires_grad:
                para.grad = para.grad * self.modelmask.to(para.device)

    def smooth_gradient(self):
        """process.py:66-112: `counts` x (Gaussian along z, then along x), on the device."""
        smooth_cfg = self.cfg["training"]["smooth"]
        counts, sigma, radius = smooth_cfg["counts"], smooth_cfg["sigma"], smooth_cfg["radius"]
        for para in self.model.parameters():
            if para.requires_grad:
                grad = para.grad.detach()
                if getattr(self.commands, "grad_cut", False):
                    grad = grad * self.modelmask.to(grad.device)
                if para.ndim != 2:
                    raise NotImplementedError("3D smoothing not implemented")
                for _ in range(counts):
                    grad = gaussian_filter(grad, sigma["z"], radius["z"], axis=0)
                    grad = gaussian_filter(grad, sigma["x"], radius["x"], axis=1)
                para.grad.data = grad.to(para.device)

    def precondition(self):
        for para in self.model.parameters():
            if para.requires_grad:
                para.grad /= self.model.precondition
