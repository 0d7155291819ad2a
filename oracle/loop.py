"""TEST INFRASTRUCTURE ONLY (oracle).  CPU (torch) restatement of the reference's
time loop, source injection, receiver sampling and index semantics.  Never
imported by the product path.

Follows:
  * seistorch/rnn.py:100-216        WaveRNN.forward
  * seistorch/source.py:47-70       WaveSource.forward2d / forward3d
  * seistorch/probe.py:42-48        WaveProbe.forward2d / forward3d
  * seistorch/setup.py:386-483      setup_rec_coords / setup_src_coords (+bwidth)
  * seistorch/utils.py:259-280      to_tensor (float -> int64 truncation)
  * seistorch/geom.py:212-239       edge padding of model parameters
  * seistorch/utils.py:235-249      ricker_wave
"""
from __future__ import annotations

import numpy as np
import torch

from . import boundary, equations


def ricker_wave(fm, dt, T, delay=80):
    """utils.py:235-249 (dtype float32 result, built in float64)."""
    i = np.arange(T)
    c = np.pi * fm * (i * dt - delay * dt)
    return ((1 - 2 * np.power(c, 2)) * np.exp(-np.power(c, 2))).astype(np.float32)


def pad_model(arr, bwidth=50, multiple=False, ndim=2):
    """geom.py:212-239: np.pad(mode='edge'); top pad dropped for `multiple`."""
    top = 0 if multiple else bwidth
    pads = [[top, bwidth]] + [[bwidth, bwidth]] * (ndim - 1)
    return np.pad(np.asarray(arr), pads, mode="edge")


def source_indices(sources, bwidth=50, multiple=False):
    """setup.py:452-483 + utils.py:259-280: add bwidth then truncate toward zero
    (torch ``.type(int64)`` on a float tensor).  2D source = [x, z] -> (x, y);
    3D source = [x, y, z] keys ('x','y','z').  Returns int64 array (nshots, ncoord)
    in key order x, y[, z]."""
    out = []
    for s in sources:
        vals = [float(v) + bwidth for v in s]
        if len(s) == 2 and multiple and bool(vals[1]):
            vals[1] -= bwidth
        # reference: float32 default dtype tensor then .type(int64)
        out.append([int(np.trunc(np.float32(v))) for v in vals])
    return np.asarray(out, dtype=np.int64)


def receiver_indices(receivers, bwidth=50, multiple=False):
    """setup.py:386-415 + rnn.py:51-73: per shot lists -> concatenated index
    vectors and the shot id of every receiver.  Returns (bidx, coords[ncoord, R],
    reccounts)."""
    cols, bidx, counts = [], [], []
    for b, rec in enumerate(receivers):
        keys = []
        for k, vals in enumerate(rec):
            # lists go through np.array (float64) before .type(int64), utils.py:268-279
            v = np.asarray([float(x) + bwidth for x in vals], dtype=np.float64)
            if len(rec) == 2 and multiple and k == 1:
                v = v - bwidth
            keys.append(np.trunc(v).astype(np.int64))
        cols.append(np.stack(keys))
        counts.append(len(keys[0]))
        bidx.append(np.full(len(keys[0]), b, dtype=np.int64))
    return np.concatenate(bidx), np.concatenate(cols, axis=1), counts


def oracle_key(case):
    """The reference selects 2D vs 3D by the model file's ndim (utils.py:296-299)."""
    nd = np.asarray(next(iter(case["models"].values()))).ndim
    return case["equation"] + ("3d" if nd == 3 else "")


def build_geometry(case, dtype=torch.float32):
    """Padded parameters, damping array and index vectors for a case dict (same
    schema as oracle/ref_runner.run_reference)."""
    eq = oracle_key(case)
    ndim = 3 if eq.endswith("3d") else 2
    multiple = bool(case.get("multiple", False))
    names = equations.MODEL_PARAMS[eq]
    npdt = np.float64 if dtype == torch.float64 else np.float32
    params = [torch.from_numpy(pad_model(np.asarray(case["models"][n], dtype=npdt), 50, multiple, ndim).copy())
              for n in names]
    shape = tuple(params[0].shape)
    if case["boundary"] == "habc":
        d = boundary.habc_coefficients_2d(shape, 50, multiple, dtype)
    elif ndim == 2:
        d = boundary.pml_coefficients_2d(shape, 50, multiple, dtype)
    else:
        d = boundary.pml_coefficients_3d(shape, 50, dtype=dtype)
    src = source_indices(case["sources"], 50, multiple)
    bidx, rec, counts = receiver_indices(case["receivers"], 50, multiple)
    return names, params, d, src, bidx, rec, counts


def simulate(case, dtype=torch.float32, requires_grad=(), wavelet=None, source_encoding=False, illumination=None):
    """rnn.py:100-216 restated.  Returns (records [list of (nt, nrec, nchan)],
    params list).  Differentiable w.r.t. params named in requires_grad.

    ``source_encoding`` (codingfwi.py:129-132,240-241; rnn.py:113,162; source.py:54-55): ONE wavefield (batch 1)
    into which every source of the case fires with its own wavelet -- ``wavelet`` is then (nsources, nt) -- through
    ``Y[..., y, x] += X`` (an index_put WITHOUT accumulation: of two sources in the same cell only one counts, a
    reference quirk this restatement keeps by using the same torch indexing); the receivers are those of the first
    shot (the driver sets the probes once, codingfwi.py:129-132).

    ``illumination``: a list; the source illumination of rnn.py:127-128,204-205 (sum over time steps and shots of the
    squared field of the LAST source type, after the source was added) is appended to it."""
    eq = oracle_key(case)
    multiple = bool(case.get("multiple", False))
    names, params, d, src, bidx, rec, counts = build_geometry(case, dtype)
    for n, p in zip(names, params):
        p.requires_grad_(n in requires_grad)
    step = equations.get_step(eq, multiple)
    wf_names = equations.WAVEFIELDS[eq]
    B = 1 if source_encoding else len(case["sources"])
    if source_encoding:
        nrec0 = counts[0]
        bidx, rec, counts = bidx[:nrec0], rec[:, :nrec0], counts[:1]
    shape = tuple(params[0].shape)
    fields = [torch.zeros((B,) + shape, dtype=dtype) for _ in wf_names]
    dt = torch.tensor(float(case["dt"]), dtype=dtype)   # cell.py:18 0-dim tensor
    h = torch.tensor(float(case["h"]), dtype=dtype)     # geom.py:32
    if isinstance(wavelet, torch.Tensor):
        x = wavelet.to(dtype)
    else:
        x = torch.as_tensor(np.asarray(case["wavelet"] if wavelet is None else wavelet), dtype=dtype)
    nt = int(case["nt"])
    # one-hot source mask, rnn.py:160-166
    smask = torch.zeros((B,) + shape, dtype=dtype)
    st = torch.from_numpy(src)
    for b in range(0 if source_encoding else B):
        if len(shape) == 2:
            smask[b, src[b, 1], src[b, 0]] = 1.0
        else:  # 3D layout (B, x, z, y); source keys x, y, z
            smask[b, src[b, 0], src[b, 2], src[b, 1]] = 1.0
    bt = torch.from_numpy(bidx)
    rt = [torch.from_numpy(r) for r in rec]
    recs = {k: [] for k in case["receiver_type"]}
    precondition = torch.zeros(shape, dtype=dtype)
    for i in range(nt):
        fields = list(step(params, fields, dt, h, d))
        for stype in case["source_type"]:
            k = wf_names.index(stype)
            if source_encoding:                                    # source.py:54-55  Y_new[..., y, x] += dt * X, dt = 1.0
                y_new = fields[k].clone()
                if len(shape) == 2:
                    y_new[..., st[:, 1], st[:, 0]] += x[:, i].view(1, -1)
                else:
                    raise NotImplementedError("the reference has no encoded 3D injection (source.py:59-70)")
                fields[k] = y_new
            else:
                fields[k] = fields[k] + smask * x[i]
        if illumination is not None:
            precondition = precondition + torch.sum(fields[wf_names.index(case["source_type"][-1])].detach() ** 2, 0)
        for rtname in case["receiver_type"]:
            f = fields[wf_names.index(rtname)]
            if len(shape) == 2:
                recs[rtname].append(f[bt, rt[1], rt[0]])           # probe.py:44  x[bidx, y, x]
            else:
                recs[rtname].append(f[bt, rt[0], rt[2], rt[1]])    # probe.py:48  x[bidx, x, z, y]
    if illumination is not None:
        illumination.append(precondition)
    stacked = torch.stack([torch.stack(recs[k], dim=0) for k in recs], dim=2)
    return list(torch.split(stacked, counts, dim=1)), dict(zip(names, params))
