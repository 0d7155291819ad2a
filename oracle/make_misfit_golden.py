"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Golden vectors for the trace-domain misfits, generated from the REAL reference
(seistorch/loss.py: L2 :409-421, L1 :381-393, CosineSimilarity :52-85, NormalizedIntegrationMethod :463-501, Envelope :178-216) on seeded random
records with a different receiver count per shot:

    python -m oracle.make_misfit_golden        # writes tests/golden/misfits.npz
    python -m oracle.make_misfit_golden filter # writes tests/golden/filter.npz (record filter, see main_filter)

Stored: the records (fp32), and per misfit name the reference's loss and d loss / d syn in float64.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

from . import ref_shim

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "misfits.npz")
NAMES = ["l2", "l1", "sml1", "cs", "cc", "integration", "nim", "w1d", "envelope"]
SHAPES = [(64, 7, 2), (64, 5, 2), (64, 1, 2)]        # (nt, nrec, nchan) of each shot


def main():
    ref_shim.import_reference()
    from seistorch.loss import Loss
    rng = np.random.default_rng(20230503)
    syn = [rng.standard_normal(s).astype(np.float32) for s in SHAPES]
    obs = [(0.7 * x + 0.5 * rng.standard_normal(x.shape)).astype(np.float32) for x in syn]
    arrs = {}
    for k, (x, y) in enumerate(zip(syn, obs)):
        arrs[f"syn_{k}"], arrs[f"obs_{k}"] = x, y
    for name in NAMES:
        xs = [torch.from_numpy(x).double().requires_grad_(True) for x in syn]
        ys = [torch.from_numpy(y).double() for y in obs]
        crit = Loss(name).loss(None)
        if name == "envelope":      # loss.py:211-216 indexes a stacked [shots, nt, nrec, nchan] tensor: one shot at a time
            loss = sum(crit(x.unsqueeze(0), y.unsqueeze(0)) for x, y in zip(xs, ys))
        else:
            loss = crit(xs, ys)
        loss.backward()
        arrs[f"{name}_loss"] = np.float64(float(loss))
        for k, x in enumerate(xs):
            arrs[f"{name}_grad_{k}"] = x.grad.numpy()
        print(name, float(loss))
    # traveltime (loss.py:674-728) takes stacked records [shots, nt, nrec, nchan]
    xs = torch.from_numpy(np.stack([syn[0], syn[0][::-1].copy() * 0.5 + 0.1 * obs[0]])).double().requires_grad_(True)
    ys = torch.from_numpy(np.stack([obs[0], np.roll(obs[0], 3, axis=0)])).double()
    loss = Loss("traveltime").loss(None)(xs, ys)
    loss.backward()
    arrs["tt_syn"], arrs["tt_obs"] = xs.detach().numpy().astype(np.float32), ys.numpy().astype(np.float32)
    arrs["tt_loss"], arrs["tt_grad"] = np.float64(float(loss)), xs.grad.numpy()
    print("traveltime", float(loss))
    np.savez_compressed(OUT, **arrs)
    print("wrote", OUT)


def main_filter():
    """Golden vectors of the device-side record filter (seistorch/signal.py:49-101, backend='torch'):
    tests/golden/filter.npz -- low-pass and band-pass, two shots with different receiver counts, output and the
    gradient of a random linear functional of the output (autograd through torchaudio in the reference)."""
    ref_shim.import_reference()
    from seistorch.signal import SeisSignal
    from seistorch.type import TensorList
    rng = np.random.default_rng(20230503)
    dt, order = 0.002, 3
    shapes = [(300, 6, 2), (300, 3, 2)]
    x = [rng.standard_normal(s).astype(np.float32) for s in shapes]
    w = [rng.standard_normal(s).astype(np.float32) for s in shapes]
    arrs = {"dt": np.float64(dt), "order": np.int64(order)}
    for k in range(len(x)):
        arrs[f"x_{k}"], arrs[f"w_{k}"] = x[k], w[k]
    sig = SeisSignal({"geom": {"dt": dt}, "training": {"filter_ord": order}})
    for tag, freqs in (("low", [30.0]), ("band", [8.0, 60.0])):
        xs = [torch.from_numpy(v).clone().requires_grad_(True) for v in x]
        out = sig.filter(TensorList([v * 1.0 for v in xs]), list(freqs), backend="torch")
        loss = sum((o.double() * torch.from_numpy(wk).double()).sum() for o, wk in zip(out.data, w))
        loss.backward()
        arrs[f"{tag}_freqs"] = np.asarray(freqs, dtype=np.float64)
        for k in range(len(x)):
            arrs[f"{tag}_y_{k}"] = out.data[k].detach().numpy()
            arrs[f"{tag}_grad_{k}"] = xs[k].grad.numpy()
        print(tag, float(loss))
    out_path = os.path.join(os.path.dirname(OUT), "filter.npz")
    np.savez_compressed(out_path, **arrs)
    print("wrote", out_path)


if __name__ == "__main__":
    if "filter" in sys.argv:
        main_filter()
    else:
        main()
