"""WaveCell -- the per-step dispatch object of seistorch/cell.py:9-76.

The whole-loop path in WaveRNN.forward bypasses ``WaveCell.forward``; the class is kept (and
works, on the sm_100a per-step ops) because drivers and user code reach the model parameters
through ``model.cell.geom`` / ``model.cell.get_parameters``.
"""
from __future__ import annotations

import inspect

import torch

from . import checkpoint as _ckpt_first_order
from . import checkpoint_new as _ckpt_second_order
from .eqconfigure import Parameters
from .habc import bound_mask
from .utils import to_tensor


def _equation_of(func):
    """Last component of the plug-in module name = equation name (cell.py:22-23)."""
    module = inspect.getmodule(func) if func is not None else None
    return module.__name__.rsplit(".", 1)[-1] if module is not None else ""


class WaveCell(torch.nn.Module):
    def __init__(self, geometry, forward_func=None, backward_func=None):
        super().__init__()
        self.geom = geometry
        self.forward_func, self.backward_func = forward_func, backward_func
        self.register_buffer("dt", to_tensor(geometry.dt))
        second = _equation_of(forward_func) in Parameters.secondorder_equations()
        self.ckpt = (_ckpt_second_order if second else _ckpt_first_order).checkpoint      # cell.py:24-28
        self.habc_masks = None

    def setup_habc(self, batchsize):
        """cell.py:30-37: boolean side masks handed to ``_time_step(..., habcs=...)``."""
        g = self.geom
        if g.use_habc:
            self.habc_masks = bound_mask(*g.domain_shape, g.bwidth, g.device, batchsize,
                                         return_idx=True, multiple=g.multiple)

    def parameters(self, recursive=True):
        yield from self.geom.parameters()

    def get_parameters(self, key=None, recursive=True, implicit=False):
        if implicit:
            yield from self.geom.nn[key].parameters()
        else:
            yield getattr(self.geom, key)

    def forward(self, wavefields, model_vars, **kwargs):
        """One time step (cell.py:50-76): through ``checkpoint`` when boundary saving is requested in
        inversion mode, else a direct call of the equation's ``_time_step``."""
        g = self.geom
        step_args = (*model_vars, *wavefields, self.dt, g.h, g.d)
        habcs = self.habc_masks if g.use_habc else None
        if g.boundary_saving and g.inversion:
            return self.ckpt(self.forward_func, self.backward_func, kwargs["source"], kwargs["is_last_frame"],
                             len(model_vars), *step_args, habcs=habcs)
        return self.forward_func(*step_args, habcs=habcs)
