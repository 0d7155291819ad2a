"""WaveCell -- the per-step dispatch object of seistorch/cell.py:9-76.

The whole-loop path in WaveRNN.forward bypasses ``WaveCell.forward``; it is kept (and
fully functional on the sm_100a per-step ops) because drivers and user code reach the
model parameters through ``model.cell.geom`` / ``model.cell.get_parameters``."""
from __future__ import annotations

import inspect

import torch

from .checkpoint import checkpoint as ckpt
from .checkpoint_new import checkpoint as ckpt_acoustic
from .eqconfigure import Parameters
from .habc import bound_mask
from .utils import to_tensor


class WaveCell(torch.nn.Module):
    def __init__(self, geometry, forward_func=None, backward_func=None):
        super().__init__()
        self.geom = geometry
        self.register_buffer("dt", to_tensor(self.geom.dt))
        self.forward_func = forward_func
        self.backward_func = backward_func
        func_name = inspect.getmodule(forward_func).__name__ if forward_func is not None else ""
        # cell.py:22-28: second-order equations use checkpoint_new
        self.ckpt = ckpt_acoustic if func_name.split(".")[-1] in Parameters.secondorder_equations() else ckpt
        self.habc_masks = None

    def setup_habc(self, batchsize):
        if self.geom.use_habc:
            self.habc_masks = bound_mask(*self.geom.domain_shape, self.geom.bwidth, self.geom.device, batchsize,
                                         return_idx=True, multiple=self.geom.multiple)

    def parameters(self, recursive=True):
        for param in self.geom.parameters():
            yield param

    def get_parameters(self, key=None, recursive=True, implicit=False):
        if implicit:
            for param in self.geom.nn[key].parameters():
                yield param
        else:
            yield getattr(self.geom, key)

    def forward(self, wavefields, model_vars, **kwargs):
        """cell.py:50-76."""
        save_condition = kwargs["is_last_frame"]
        source_term = kwargs["source"]
        geoms = self.dt, self.geom.h, self.geom.d
        habcs = self.habc_masks if self.geom.use_habc else None
        if self.geom.boundary_saving and self.geom.inversion:
            return self.ckpt(self.forward_func, self.backward_func, source_term, save_condition, len(model_vars),
                             *model_vars, *wavefields, *geoms, habcs=habcs)
        return self.forward_func(*model_vars, *wavefields, *geoms, habcs=habcs)
