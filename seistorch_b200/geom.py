"""WaveGeometry -- holds the padded model parameters and the absorbing-boundary
coefficients, with the attribute surface WaveCell / WaveRNN and the drivers read from
``seistorch/geom.py`` (WaveGeometryFreeForm :95-239).  Host-side set-up, run once.

Unlike the reference it can also be built directly from in-memory arrays
(``WaveGeometry.from_arrays``), which is what the tests and bench.py use: file formats
are outside the accelerated path (component 22).
"""
from __future__ import annotations

import importlib
import os
from typing import Dict, Optional

import numpy as np
import torch

from .eqconfigure import Parameters
from .utils import to_tensor


class WaveGeometryFreeForm(torch.nn.Module):
    def __init__(self, mode="forward", logger=None, **kwargs):
        super().__init__()
        self.mode = mode
        self.autodiff = True
        self.kwargs = kwargs
        g = kwargs["geom"]
        self.unit = g.get("unit", 1.0)
        self.dt = g["dt"]
        self.dh = g["h"] * self.unit
        self.device = kwargs["device"]
        self.bwidth = g["boundary"]["width"]
        self.domain_shape = tuple(kwargs["domain_shape"])
        self.boundary_saving = g.get("boundary_saving", False)
        self.source_type = g["source_type"]
        self.receiver_type = g["receiver_type"]
        self.multiple = g["multiple"]
        self.model_parameters = []
        self.inversion = mode == "inversion"
        self.logger = logger
        self.source_illumination = g.get("source_illumination", False)
        self.ndim = len(self.domain_shape)
        self.equation = kwargs["equation"]
        self.use_implicit = kwargs.get("training", {}).get("implicit", {}).get("use", False)
        self.register_buffer("h", to_tensor(self.dh))
        self.setup_bc()
        self._init_model(kwargs.get("VEL_PATH", {}), g.get("invlist", {}), kwargs.get("_models"))

    # ---- boundary (geom.py:45-74)
    def setup_bc(self):
        btype = self.kwargs["geom"]["boundary"]["type"]
        self.use_pml = btype == "pml" and self.bwidth > 0
        self.use_random = btype == "random" and self.bwidth > 0
        self.use_habc = btype == "habc" and self.bwidth > 0
        assert btype in ["pml", "random", "habc"], "boundary type must be one of [pml, random, habc]"
        if self.use_random:
            raise NotImplementedError("seistorch_b200: random boundaries are not on the accelerated path")
        if self.use_pml or self.use_habc:
            module = importlib.import_module(f"{__package__}.{btype}")
            coes_func = getattr(module, f"generate_{btype}_coefficients_{self.ndim}d")
            if btype == "habc":
                self.bwidth = 50    # geom.py:62-65
            self.register_buffer("_d", coes_func(self.domain_shape, self.bwidth, multiple=self.multiple))
        else:
            self.register_buffer("_d", torch.zeros(self.domain_shape))

    @property
    def d(self):
        return self._d

    @property
    def padding_list(self):
        top = 0 if self.multiple else self.bwidth
        return [[top, self.bwidth]] + [[self.bwidth, self.bwidth]] * (self.ndim - 1)

    # ---- parameters (geom.py:180-239)
    def _init_model(self, model_path: Dict, invlist: Dict, arrays: Optional[Dict]):
        needed = Parameters.valid_model_paras()[self.equation]
        self.pars_need_invert = []
        self.true_models = dict()
        for name in needed:
            if arrays is not None and name in arrays:
                data = np.asarray(arrays[name])
            else:
                path = model_path[name]
                if path is None or not os.path.exists(path):
                    raise FileNotFoundError(f"Cannot find model file '{path}' needed by equation <{self.equation}>")
                data = np.load(path)
            if invlist.get(name):
                self.pars_need_invert.append(name)
            self.model_parameters.append(name)
            invert = False if self.mode == "forward" else bool(invlist.get(name))
            padded = np.pad(data * self.unit, self.padding_list, mode="edge")
            if tuple(padded.shape) != self.domain_shape:
                raise ValueError(f"model '{name}' pads to {padded.shape}, expected {self.domain_shape}")
            setattr(self, name, torch.nn.Parameter(to_tensor(padded), requires_grad=invert))

    def __repr__(self):
        return f"Paramters of {self.model_parameters} have been defined."

    def forward(self):
        raise NotImplementedError("WaveGeometry is a parameter container; forward() is never called.")
