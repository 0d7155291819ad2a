"""WaveRNN -- same public surface as seistorch/rnn.py:14-216, but ``forward`` hands the
whole time loop to the sm_100a propagators (seistorch_b200/engine.py) instead of
iterating ``nt`` times in Python.

Semantics preserved (rnn.py:100-216):
  * zero initial state; step i: fields <- step(fields); then ``field[st] += x[:, i]`` at
    the source cell for every ``source_type``; then ``field[rt]`` is sampled at the
    receivers for every ``receiver_type`` (channel order = receiver_type order);
  * returns a TensorList of per-shot records (nt, nrec_i, nchan); NaN -> ValueError;
  * ``super_source`` / ``super_probes`` (coords.single2batch) override the stored
    acquisition; source encoding puts all sources into one shot.
"""
from __future__ import annotations

from typing import Iterator, Tuple

import torch
from torch.nn.parameter import Parameter

from . import coefficients as _coef
from .engine import Acquisition, Spec, propagate
from .eqconfigure import Parameters, Wavefield, field_channels
from .probe import WaveProbe
from .setup import setup_acquisition
from .source import WaveSource
from .type import TensorList


class WaveRNN(torch.nn.Module):
    def __init__(self, cell, source_encoding=False):
        super().__init__()
        self.cell = cell
        self.source_encoding = source_encoding
        self.use_implicit = self.cell.geom.use_implicit
        self.second_order_equation = self.cell.geom.equation in Parameters.secondorder_equations()
        self.source_illumination = self.cell.geom.source_illumination
        # engine knobs (not in the reference): see engine._history_plan
        self.history_budget_bytes = None
        self.segment = None
        self._acq_cache = None

    def named_parameters(self, prefix: str = "", recurse: bool = True, remove_duplicate: bool = True) -> Iterator[Tuple[str, Parameter]]:
        if self.cell.geom.use_implicit:
            for key in self.cell.geom.nn:
                for name, param in self.cell.geom.nn[key].named_parameters(prefix, recurse, remove_duplicate):
                    yield name, param
        else:
            for name, param in self.cell.geom.named_parameters(prefix, recurse, remove_duplicate):
                yield name, param

    # ---- acquisition bookkeeping (rnn.py:35-97)
    def merge_sources_with_same_keys(self):
        super_source, batchindices = dict(), []
        for bidx, source in enumerate(self.sources):
            for key, v in source.coords().items():
                super_source.setdefault(key, []).append(v)
            batchindices.append(bidx * torch.ones(1, dtype=torch.int64))
        return batchindices, super_source

    def merge_receivers_with_same_keys(self):
        super_probes, batchindices, reccounts = dict(), [], []
        for bidx, probe in enumerate(self.probes):
            coords = probe.coords()
            for key, v in coords.items():
                super_probes.setdefault(key, []).append(v)
            n = len(coords[key])
            reccounts.append(n)
            batchindices.append(bidx * torch.ones(n, dtype=torch.int64))
        for key in super_probes:
            super_probes[key] = torch.concatenate(super_probes[key], dim=0)
        return reccounts, torch.concatenate(batchindices), super_probes

    def reset_sources(self, sources):
        self.sources = torch.nn.ModuleList(sources if isinstance(sources, list) else [sources])
        self._acq_cache = None

    def reset_probes(self, probes):
        self.probes = torch.nn.ModuleList(probes if isinstance(probes, list) else [probes])
        self._acq_cache = None

    def reset_geom(self, shots, src_list, rec_list, cfg):
        sources, receivers = setup_acquisition(shots, src_list, rec_list, cfg)
        self.reset_sources(sources)
        self.reset_probes(receivers)
        for module in self.probes:
            module.to(self.cell.geom.device)
        for module in self.sources:
            module.to(self.cell.geom.device)

    # ---- the hot path
    def forward(self, x, omega=10.0, super_source=None, super_probes=None, vp=None):
        geom = self.cell.geom
        device = torch.device(geom.device)
        if device.type != "cuda":
            raise RuntimeError("seistorch_b200: WaveRNN runs on CUDA devices only (no CPU fallback); "
                               f"geom.device = {geom.device}")
        ndim = len(geom.domain_shape)
        equation = geom.equation
        # The device-side index tables (validation, sort, CSR) are built once per acquisition and reused: the
        # model's own sources / probes until reset_*() is called, a caller-supplied (super_source, super_probes)
        # pair for as long as the same two objects are passed (they are immutable index buffers).
        own = super_source is None and super_probes is None
        key = (tuple(geom.domain_shape), bool(self.source_encoding), str(device))
        cached = getattr(self, "_acq_cache", None) if own else getattr(super_probes, "_st_acq", None)
        if cached is not None and cached[0] == key and (own or cached[1] is super_source):
            acq, reccounts, ns = cached[2], cached[3], cached[4]
        else:
            given_source = super_source
            if super_source is None:
                bidx_source, sourcekeys = self.merge_sources_with_same_keys()
                super_source = WaveSource(bidx_source, self.second_order_equation, **sourcekeys).to(device)
            if super_probes is None:
                reccounts, bidx_receivers, reckeys = self.merge_receivers_with_same_keys()
                super_probes = WaveProbe(bidx_receivers, **reckeys).to(device)
            else:
                reccounts = super_probes.reccounts
            super_source.source_encoding = self.source_encoding
            super_source.second_order_equation = self.second_order_equation

            ns = int(super_source.x.reshape(-1).shape[0])
            batchsize = 1 if self.source_encoding else ns            # rnn.py:112-116
            sx = super_source.x.reshape(-1).to(device)
            sy = super_source.y.reshape(-1).to(device)
            src_b = torch.zeros(ns, dtype=torch.int64, device=device) if self.source_encoding \
                else torch.arange(ns, dtype=torch.int64, device=device)
            rb = torch.as_tensor(super_probes.bidx, dtype=torch.int64, device=device).reshape(-1)
            rx = super_probes.x.reshape(-1).to(device)
            ry = super_probes.y.reshape(-1).to(device)
            if ndim == 2:
                src_idx = torch.stack([sy, sx], dim=1)               # smask[b, y, x]          rnn.py:164
                rec_idx = torch.stack([ry, rx], dim=1)               # field[bidx, y, x]       probe.py:44
            else:
                sz = super_source.z.reshape(-1).to(device)
                rz = super_probes.z.reshape(-1).to(device)
                src_idx = torch.stack([sx, sz, sy], dim=1)           # smask[b, x, z, y]       rnn.py:166
                rec_idx = torch.stack([rx, rz, ry], dim=1)           # field[bidx, x, z, y]    probe.py:48
            acq = Acquisition(geom.domain_shape, batchsize, src_b, src_idx, rb, rec_idx, device)
            reccounts = list(reccounts)
            if own:
                self._acq_cache = (key, None, acq, reccounts, ns)
            elif given_source is not None:
                super_probes._st_acq = (key, given_source, acq, reccounts, ns)
        batchsize = acq.B

        # wavelet -> per-source amplitudes amp[nt, ns]
        x = x.to(device)
        if x.ndim == 1:
            x = x.unsqueeze(0)
        nt = x.shape[1]
        if x.shape[0] == 1:
            amp = x[0].unsqueeze(1).expand(nt, ns)
        elif x.shape[0] == ns:
            amp = x.t()
        else:
            raise ValueError(f"wavelet batch {x.shape[0]} does not match {ns} sources")

        # model parameters -> kernel coefficients (differentiable torch ops, once per call)
        if self.use_implicit:
            params = [vp] + [getattr(geom, n) for n in geom.model_parameters[1:]]
        else:
            params = [getattr(geom, n) for n in geom.model_parameters]
        dt, h, d = float(self.cell.dt), float(geom.h), geom.d
        if ndim == 3:
            if equation != "acoustic":
                raise ValueError(f"seistorch_b200: 3D equation '{equation}' is not supported")
            family, flags, slots = "acoustic3d", 0, ()
            coefs = _coef.acoustic3d_coefficients(params, dt, h, d)
            chan = {"h1": 0}
        elif equation == "elastic":
            family, flags, slots = "elastic2d", 0, ()
            coefs = _coef.elastic_coefficients(params, dt, h, d)
            chan = field_channels(equation)
        else:
            family, flags = _coef.EQUATIONS[equation]
            coefs, slots = _coef.wave2d_coefficients(equation, params, dt, h, d)
            chan = field_channels(equation)
        for name in list(geom.source_type) + list(geom.receiver_type):
            if name not in chan:
                raise ValueError(f"seistorch_b200: wavefield '{name}' cannot be a source/receiver of '{equation}' "
                                 f"on the accelerated path (valid: {sorted(chan)})")
        fmask = 0
        for name in geom.source_type:
            fmask |= 1 << chan[name]
        spec = Spec(family=family, flags=flags, shape=tuple(geom.domain_shape), B=batchsize, nt=nt, dt=dt,
                    bw=int(geom.bwidth), multiple=bool(geom.multiple), src_fmask=fmask,
                    chan_f=tuple(chan[n] for n in geom.receiver_type), coef_slots=slots,
                    history_budget_bytes=self.history_budget_bytes, segment=self.segment)
        if self.source_illumination:
            # rnn.py:127-128,204-205: precondition = sum over time and shots of field[last source_type]^2, reset at every
            # forward call.  Here it is filled during backward() from the wavefield history (the drivers use it after
            # loss.backward(): seistorch_dist.py:258,279-280).
            spec.illum_chan = chan[list(geom.source_type)[-1]]
            illum = torch.zeros(tuple(geom.domain_shape[:-1]) + (spec.ld,), dtype=torch.float32, device=device)
            spec.illum_plane = illum
            self.precondition = illum[..., :geom.domain_shape[-1]]
        rec = propagate(spec, acq, amp, coefs)                   # [nt, sum(nrec), nchan]

        if bool(torch.isnan(rec).any()):                         # type.py:41-46 (one device sync, not one per shot)
            raise ValueError("The tensor list contains NaN values.")
        y = TensorList()
        y.data.extend(torch.split(rec, list(reccounts), dim=1))
        return y
