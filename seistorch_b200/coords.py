"""Batch coordinate packing -- seistorch/coords.py:82-155 (host glue used by the
torchrun driver between the DataLoader and WaveRNN.forward)."""
from __future__ import annotations

import torch

from .probe import WaveProbe
from .setup import setup_acquisition
from .source import WaveSource


def merge_sources_with_same_keys(sources):
    super_source, batchindices = dict(), []
    for bidx, source in enumerate(sources):
        for key, v in source.coords().items():
            super_source.setdefault(key, []).append(v)
        batchindices.append(bidx * torch.ones(1, dtype=torch.int64))
    return batchindices, super_source


def merge_receivers_with_same_keys(receivers):
    super_probes, batchindices, reccounts = dict(), [], []
    for bidx, probe in enumerate(receivers):
        coords = probe.coords()
        for key, v in coords.items():
            super_probes.setdefault(key, []).append(v)
        n = len(coords[key])
        reccounts.append(n)
        batchindices.append(bidx * torch.ones(n, dtype=torch.int64))
    for key in super_probes:
        super_probes[key] = torch.concatenate(super_probes[key], dim=0)
    return reccounts, torch.concatenate(batchindices), super_probes


def single2batch(src, rec, cfg, dev):
    """coords.py:127-155: DataLoader batch -> padded super source / super probes."""
    rec = rec.permute(2, 0, 1).cpu().numpy().tolist()
    src = torch.stack(src).cpu().numpy().T.tolist()
    padded_src, padded_rec = setup_acquisition(range(len(src)), src, rec, cfg)
    bidx_source, sourcekeys = merge_sources_with_same_keys(padded_src)
    super_source = WaveSource(bidx_source, **sourcekeys).to(dev)
    reccounts, bidx_receivers, reckeys = merge_receivers_with_same_keys(padded_rec)
    super_probes = WaveProbe(bidx_receivers, **reckeys).to(dev)
    super_probes.reccounts = reccounts
    return super_source, super_probes
