"""Whole-time-loop propagators on top of the C ABI (include/seistorch_b200.h).

This is the host side of the hot path: it replaces the Python time loop of
``WaveRNN.forward`` (seistorch/rnn.py:178-205), the per-step dispatch of
``WaveCell.forward`` (seistorch/cell.py:50-76) and torch autograd-through-time /
``CheckpointFunction`` (seistorch/checkpoint_new.py:109-217, checkpoint.py:108-230)
with ONE ``torch.autograd.Function`` per forward call:

  forward : nt fused step launches (stencil + boundary + source add + receiver gather)
  backward: exact discrete adjoint (transposed stencil + imaging condition), fed either
            by the stored wavefield history (when it fits the memory budget) or by
            K-step checkpoints + recomputation (memory O(nt/K + K) states).

PyTorch is used for device memory, streams and autograd plumbing only; all per-step
arithmetic is in the sm_100a kernels.  There is no CPU fallback: a missing library or a
CPU tensor raises.
"""
from __future__ import annotations

import ctypes as C
import math
import os
from dataclasses import dataclass
from typing import Optional, Sequence

import torch

from . import _lib

# equation-variant flags (include/seistorch_b200.h)
EQ_ISO, EQ_PML, EQ_HABC, EQ_XZ, EQ_G1, EQ_BORN = 1, 2, 4, 8, 16, 32

# (family, flags, field names in channel order, coefficient slots used -> gradient slot)
# wave2d coefficient order: r, b, cxx, czz, cxz, ax, az, m ; gradient order r,cxx,czz,cxz,ax,az,m
_W2_GRAD_OF_COEF = {0: 0, 2: 1, 3: 2, 4: 3, 5: 4, 6: 5, 7: 6}


# number of sm_100a kernel launches issued through the C ABI (bench.py reports it)
LAUNCHES = {"forward": 0, "adjoint": 0, "misfit": 0}
STEPS = {"forward": 0, "adjoint": 0}               # time steps advanced (a persistent launch advances many)
KERNELS = {"forward": None, "adjoint": None}      # kernel family of the most recent forward / adjoint call
LAUNCHES_LAST = {"adjoint": 0, "recompute": 0}    # time steps of the most recent backward(): adjoint, recomputed forward


def _round_up(n, m):
    return (n + m - 1) // m * m


def _stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _require_cuda(t: torch.Tensor, what: str):
    if not t.is_cuda:
        raise RuntimeError(f"seistorch_b200: {what} must be a CUDA tensor (no CPU fallback); got {t.device}")


def _one_device(tensors, what: str) -> torch.device:
    """All operands of one call must live on ONE CUDA device; returns it.  The C ABI launches on the
    current device's context, so every entry point runs under ``torch.cuda.device(<that device>)`` --
    the reference drivers build the model on ``cuda:{rank}`` without ever calling ``set_device``
    (seistorch_dist.py:89-94)."""
    dev = None
    for t in tensors:
        if t is None or not t.is_cuda:          # host operands (e.g. observed data still on the CPU) are moved by the op itself
            continue
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise RuntimeError(f"seistorch_b200: {what}: operands live on different devices ({dev} and {t.device})")
    if dev is None:
        raise RuntimeError(f"seistorch_b200: {what}: no CUDA operand (no CPU fallback)")
    return dev


def on_device_of(argpos: int = 1):
    """Decorator for autograd.Function.forward/backward bodies: run under the device of the first CUDA
    tensor found at/after positional argument ``argpos`` (backward: ctx.dev)."""
    def deco(fn):
        def wrapped(ctx, *args):
            dev = getattr(ctx, "dev", None)
            if dev is None:
                dev = _one_device([a for a in args[argpos - 1:] if isinstance(a, torch.Tensor)], fn.__qualname__)
                ctx.dev = dev
            with torch.cuda.device(dev):
                return fn(ctx, *args)
        wrapped.__name__, wrapped.__doc__ = fn.__name__, fn.__doc__
        return wrapped
    return deco


# ======================================================================== acquisition
class Acquisition:
    """Device-side source / receiver index tables (int32) in the layout of
    ``st_acquisition``.  Mirrors the index semantics of WaveSource / WaveProbe
    (source.py:47-70, probe.py:42-48): tensor-dimension order, one source row per point
    source, receivers of all shots concatenated (rnn.py:51-73)."""

    def __init__(self, shape: Sequence[int], B: int, src_b, src_idx, rec_b, rec_idx, device):
        self.shape = tuple(int(s) for s in shape)
        self.ndim = len(self.shape)
        self.B = int(B)
        dev = torch.device(device)
        if dev.type == "cuda" and dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        self.device = dev
        src_b = torch.as_tensor(src_b, dtype=torch.int64, device=dev).reshape(-1)
        src_idx = torch.as_tensor(src_idx, dtype=torch.int64, device=dev).reshape(-1, self.ndim)
        rec_b = torch.as_tensor(rec_b, dtype=torch.int64, device=dev).reshape(-1)
        rec_idx = torch.as_tensor(rec_idx, dtype=torch.int64, device=dev).reshape(-1, self.ndim)
        self.ns = int(src_b.numel())
        self.R = int(rec_b.numel())
        lim = torch.tensor(self.shape, dtype=torch.int64, device=dev)
        bad = False
        if self.ns:
            bad |= bool(((src_idx < 0) | (src_idx >= lim)).any() or (src_b < 0).any() or (src_b >= B).any())
        if self.R:
            bad |= bool(((rec_idx < 0) | (rec_idx >= lim)).any() or (rec_b < 0).any() or (rec_b >= B).any())
        if bad:
            raise IndexError("seistorch_b200: source/receiver index outside the padded domain")
        # row range of all sources and receivers (kernels skip the source/receiver epilogue elsewhere)
        rows = [t[:, 0] for t in (src_idx, rec_idx) if t.numel()]
        if rows:
            allrows = torch.cat(rows)
            self.row_lo, self.row_hi = int(allrows.min()), int(allrows.max())
        else:
            self.row_lo, self.row_hi = 1, 0
        self.src_b = src_b.to(torch.int32).contiguous()
        self.src_i = [src_idx[:, k].to(torch.int32).contiguous() for k in range(self.ndim)]
        # receivers: sort by row key, CSR over rows
        rows_per_shot = 1
        for s in self.shape[:-1]:
            rows_per_shot *= s
        nrows = self.B * rows_per_shot
        if nrows + 1 >= 2 ** 31:
            raise ValueError("seistorch_b200: too many rows for int32 receiver index")
        if self.R:
            key = rec_b
            for k in range(self.ndim - 1):
                key = key * self.shape[k] + rec_idx[:, k]
            order = torch.argsort(key, stable=True)
            counts = torch.bincount(key, minlength=nrows)
            self.rec_col = rec_idx[order, self.ndim - 1].to(torch.int32).contiguous()
            self.rec_orig = order.to(torch.int32).contiguous()
        else:
            counts = torch.zeros(nrows, dtype=torch.int64, device=dev)
            self.rec_col = torch.zeros(1, dtype=torch.int32, device=dev)
            self.rec_orig = torch.zeros(1, dtype=torch.int32, device=dev)
        rs = torch.zeros(nrows + 1, dtype=torch.int64, device=dev)
        rs[1:] = torch.cumsum(counts, 0)
        self.row_start = rs.to(torch.int32).contiguous()

    def fill(self, q: _lib.StAcquisition, amp, gamp, src_fmask, chan_f, rec_out, rec_adj):
        q.ns = self.ns
        q.src_b = _lib.ptr(self.src_b)
        if self.ndim == 3:
            q.src_i0, q.src_i1, q.src_i2 = (_lib.ptr(t) for t in self.src_i)
        else:
            q.src_i0 = None
            q.src_i1, q.src_i2 = (_lib.ptr(t) for t in self.src_i)
        q.amp = _lib.ptr(amp)
        q.gamp = _lib.ptr(gamp)
        q.src_fmask = int(src_fmask)
        q.R = self.R
        q.row_start = _lib.ptr(self.row_start)
        q.rec_col = _lib.ptr(self.rec_col)
        q.rec_orig = _lib.ptr(self.rec_orig)
        q.nchan = len(chan_f)
        for k in range(4):
            q.chan_f[k] = int(chan_f[k]) if k < len(chan_f) else 0
        q.rec_out = _lib.ptr(rec_out)
        q.rec_adj = _lib.ptr(rec_adj)
        q.row_lo, q.row_hi = self.row_lo, self.row_hi


# ======================================================================== plan
@dataclass
class Spec:
    """Static description of one propagation call."""
    family: str                     # 'wave2d' | 'elastic2d' | 'acoustic3d'
    flags: int
    shape: tuple                    # padded domain (nz,nx) or (n0,n1,n2)
    B: int
    nt: int
    dt: float
    bw: int = 50
    multiple: bool = False
    src_fmask: int = 1
    chan_f: tuple = (0,)
    coef_slots: tuple = ()          # for wave2d: slot index (0..7) of every coefficient tensor passed
    illum_chan: Optional[int] = None               # source illumination: field channel whose squares are summed (rnn.py:204-205)
    illum_plane: Optional[torch.Tensor] = None     # ... into this zeroed pitched plane, during backward()
    history_budget_bytes: Optional[int] = None     # None: derive from free memory
    segment: Optional[int] = None                  # force a checkpoint segment length (tests)

    @property
    def nf(self):
        if self.family == "wave2d":
            return 2 if self.flags & EQ_BORN else 1
        return 5 if self.family == "elastic2d" else 1

    @property
    def order(self):
        return 1 if self.family == "elastic2d" else 2

    @property
    def nlam(self):
        return 2 if self.family == "elastic2d" else 3

    @property
    def ngrad(self):
        return {"wave2d": 7, "elastic2d": 4, "acoustic3d": 1}[self.family]

    @property
    def ld(self):
        return _round_up(self.shape[-1], 4)

    @property
    def plane(self):
        n = self.ld
        for s in self.shape[:-1]:
            n *= s
        return n

    @property
    def slot_elems(self):
        return self.nf * self.B * self.plane


def _choose_bchunk(spec: Spec) -> int:
    """Shots handled by one adjoint block (gradient accumulated in registers across
    them); keep at least ~4 blocks per SM."""
    env = os.environ.get("SEISTORCH_B200_BCHUNK")
    if env:
        return max(1, min(int(env), spec.B))
    if spec.family == "elastic2d":
        # the vectorised adjoint keeps the gradient sums of a chunk's shots in shared memory (st_elastic2d.cu, fast path):
        # the longest chunk that still fills one wave of 2 blocks / SM (B200, 500x1100: 4 shots 73 -> 63 us with chunks of 2,
        # 97 us with one chunk of 4; 8 shots 137 -> 112 us with chunks of 4)
        tiles = math.ceil(spec.shape[1] / 128) * math.ceil(spec.shape[0] / 32)
        best = 1
        for c in range(1, spec.B + 1):
            if spec.B % c == 0 and tiles * (spec.B // c) >= 280:
                best = c
        return best
    if spec.family == "acoustic3d":
        tiles = math.ceil(spec.shape[2] / 64) * math.ceil(spec.shape[1] / 8) * math.ceil(spec.shape[0] / 16)
    elif spec.family == "wave2d":
        tiles = math.ceil(spec.shape[1] / 128) * math.ceil(spec.shape[0] / 32)     # pairs of fast tiles (tuned on B200: SEISTORCH_B200_BCHUNK sweep)
    else:
        tiles = math.ceil(spec.shape[1] / 64) * math.ceil(spec.shape[0] / 32)
    best = 1
    for c in range(1, spec.B + 1):
        # >= 2 waves of 3 blocks/SM on 148 SMs; the wave2d fast blocks keep their gradient sums in shared memory across the
        # chunk, so longer chunks pay (B200 sweep, 12 shots 600x1300: chunks of 2 / 4 / 6 shots: 207 / 197 / 193 us Born, 197 / 188 / 208 us tti_habc)
        if spec.B % c == 0 and tiles * (spec.B // c) >= (600 if spec.family == "wave2d" else 888):
            best = c
    return best


class _Problem:
    """Owns the ctypes problem struct + the torch buffers it points to."""

    def __init__(self, spec: Spec, coefp: torch.Tensor, acq: Acquisition, amp: torch.Tensor, nslots: int,
                 u: Optional[torch.Tensor] = None):
        self.spec = spec
        self.acq = acq
        self.coefp = coefp            # [ncoef, plane]  pitched coefficient planes
        self.amp = amp                # [nt, ns] fp32 contiguous
        dev = coefp.device
        self.nslots = nslots
        if u is None:
            # history slots are written in full by the kernels (incl. zero pitch padding), so only the
            # initial state needs clearing -- a 100+ GB memset per forward call would cost ~40 ms
            u = torch.empty(nslots * spec.slot_elems, dtype=torch.float32, device=dev)
            u[:min(nslots, spec.order + 1) * spec.slot_elems].zero_()
        self.u = u
        self.lam = None
        self.gacc = None
        self.bchunk = 1
        self.rec_out = None
        self.rec_adj = None
        self.gamp = None
        self.taps = None
        self._prepared = False
        L = _lib.lib()
        fam = spec.family
        if fam == "wave2d":
            self.p = _lib.StWave2dProblem()
            self.fwd, self.adj = L.st_wave2d_forward, L.st_wave2d_adjoint
        elif fam == "elastic2d":
            self.p = _lib.StElastic2dProblem()
            self.fwd, self.adj = L.st_elastic2d_forward, L.st_elastic2d_adjoint
        elif fam == "acoustic3d":
            self.p = _lib.StAcoustic3dProblem()
            self.fwd, self.adj = L.st_acoustic3d_forward, L.st_acoustic3d_adjoint
        else:
            raise ValueError(fam)

    def _sync_struct(self):
        s, p = self.spec, self.p
        if s.family == "wave2d":
            p.flags = s.flags
            p.B, p.nz, p.nx, p.ld = s.B, s.shape[0], s.shape[1], s.ld
            p.bw, p.multiple = s.bw, int(s.multiple)
            p.nt, p.dt = s.nt, float(s.dt)
            for k in range(8):
                p.coef[k] = None
            for row, slot in enumerate(s.coef_slots):
                p.coef[slot] = self.coefp[row].data_ptr()
        elif s.family == "elastic2d":
            p.B, p.nz, p.nx, p.ld, p.nt = s.B, s.shape[0], s.shape[1], s.ld, s.nt
            for k in range(5):
                p.coef[k] = self.coefp[k].data_ptr()
        else:
            p.B, p.n0, p.n1, p.n2, p.ld, p.nt = s.B, s.shape[0], s.shape[1], s.shape[2], s.ld, s.nt
            p.dt = float(s.dt)
            p.coef[0] = self.coefp[0].data_ptr()
            p.coef[1] = self.coefp[1].data_ptr()
        if s.family == "wave2d":
            if not self._prepared:
                p.taps = None
                n = int(_lib.lib().st_wave2d_taps_floats(C.byref(p)))
                if n > 0:
                    self.taps = torch.empty(n, dtype=torch.float32, device=self.coefp.device)
                    p.taps = self.taps.data_ptr()
                    p.u = self.u.data_ptr()
                    p.nslots = self.nslots
                    self.acq.fill(p.acq, self.amp, None, s.src_fmask, s.chan_f, None, None)
                    _lib.check(_lib.lib().st_wave2d_prepare(C.byref(p), _stream_ptr()), "wave2d_prepare")
                self._prepared = True
            p.taps = _lib.ptr(self.taps)
        p.u = self.u.data_ptr()
        p.nslots = self.nslots
        p.lam = _lib.ptr(self.lam)
        p.gacc = _lib.ptr(self.gacc)
        p.bchunk = self.bchunk
        self.acq.fill(p.acq, self.amp, self.gamp, s.src_fmask, s.chan_f, self.rec_out, self.rec_adj)

    def _note_kernel(self, which, nsteps=0):
        """Record which kernel family serves this call (bench.py / tests report it); returns the number of
        kernel launches the call will issue."""
        if which == "forward" and self.spec.family == "wave2d" and \
                _lib.lib().st_wave2d_uses_persist(C.byref(self.p), int(nsteps)):
            KERNELS[which] = "wave2d_persist_forward_kernel"
            return 1
        if which == "adjoint" and self.spec.family == "wave2d" and \
                _lib.lib().st_wave2d_adjoint_uses_persist(C.byref(self.p), int(nsteps)):
            KERNELS[which] = "wave2d_persist_adjoint_kernel"
            return 1
        if self.spec.family == "wave2d":
            tma = bool(_lib.lib().st_wave2d_uses_tma(C.byref(self.p), 1 if which == "adjoint" else 0))
            KERNELS[which] = f"wave2d_{which}_{'tma_' if tma else ''}kernel"
        else:
            KERNELS[which] = f"{self.spec.family}_{which}_kernel"
        return nsteps

    def forward(self, i0, nsteps, slot0, record=True):
        keep = self.rec_out
        if not record:
            self.rec_out = None
        self._sync_struct()
        nlaunch = self._note_kernel("forward", nsteps)
        self.rec_out = keep
        _lib.check(self.fwd(C.byref(self.p), i0, nsteps, slot0 % self.nslots, _stream_ptr()), f"{self.spec.family}_forward")
        LAUNCHES["forward"] += nlaunch
        STEPS["forward"] += nsteps

    def adjoint(self, i_hi, nsteps, slot_hi):
        self._sync_struct()
        nlaunch = self._note_kernel("adjoint", nsteps)
        _lib.check(self.adj(C.byref(self.p), i_hi, nsteps, slot_hi % self.nslots, _stream_ptr()), f"{self.spec.family}_adjoint")
        LAUNCHES["adjoint"] += nlaunch
        STEPS["adjoint"] += nsteps

    def slot_view(self, slot, count=1):
        e = self.spec.slot_elems
        slot %= self.nslots
        return self.u[slot * e:(slot + count) * e]


def _pack_coefs(spec: Spec, coefs: Sequence[torch.Tensor]) -> torch.Tensor:
    """[ncoef, *shape] fp32 -> pitched planes [ncoef, plane]."""
    dev = coefs[0].device
    nx, ld = spec.shape[-1], spec.ld
    out = torch.zeros((len(coefs),) + tuple(spec.shape[:-1]) + (ld,), dtype=torch.float32, device=dev)
    for k, c in enumerate(coefs):
        if tuple(c.shape) != tuple(spec.shape):
            raise ValueError(f"coefficient {k} has shape {tuple(c.shape)}, expected {spec.shape}")
        out[k, ..., :nx] = c.detach().to(torch.float32)
    return out.reshape(len(coefs), -1)


def _history_plan(spec: Spec, dev) -> tuple:
    """Return (K, nseg): K-step segments with stored history inside a segment.
    K == nt means the whole history is stored in the forward pass (no recomputation)."""
    nt, p = spec.nt, spec.order
    if spec.segment is not None:
        K = max(1, min(int(spec.segment), nt))
        return K, math.ceil(nt / K)
    slot_bytes = spec.slot_elems * 4
    if spec.history_budget_bytes is not None:
        budget = int(spec.history_budget_bytes)
    else:
        env = os.environ.get("SEISTORCH_B200_HISTORY_GB")
        if env:
            budget = int(float(env) * 2 ** 30)
        else:
            free, _total = torch.cuda.mem_get_info(dev)
            # blocks cached by torch's allocator (e.g. the previous call's history) are reusable
            free += torch.cuda.memory_reserved(dev) - torch.cuda.memory_allocated(dev) + _pooled_bytes(dev)
            reserve = (spec.nlam + 2) * slot_bytes + spec.ngrad * spec.plane * 4 * spec.B + (1 << 30)
            budget = int(0.85 * free) - reserve
    nslots_max = max(budget // slot_bytes, p + 2)
    if nt + p <= nslots_max:
        return nt, 1
    best = None
    for K in range(1, nt + 1):
        need = K + p + p * math.ceil(nt / K)
        if need <= nslots_max:
            best = K
    if best is None:
        raise RuntimeError(
            f"seistorch_b200: wavefield history does not fit: one state slot is {slot_bytes / 2**20:.1f} MiB and "
            f"the budget allows {nslots_max} slots; reduce the number of shots per call")
    return best, math.ceil(nt / best)


# ---- wavefield-history buffers are kept across calls -------------------------------------------------------------------
# A gradient call stores nt (or K) states of all shots: 100+ GB on the BASELINE grids.  Handing that block back to torch's
# caching allocator after every backward() lets later small allocations carve pieces out of it; the next forward then finds
# no block of its size, the allocator frees its whole cache (a device sync) and cudaMallocs again -- sporadic 30-500 ms
# stalls per step, measured on B200.  The engine therefore keeps ONE idle history buffer per (device, stream) and reuses it
# when it is large enough.  `release_buffers()` returns the memory; SEISTORCH_B200_HISTORY_POOL=0 turns the pool off.
_HISTORY_POOL: dict = {}


def _pool_on() -> bool:
    return os.environ.get("SEISTORCH_B200_HISTORY_POOL", "1") != "0"


def _pool_key(dev):
    return (dev.index if dev.index is not None else torch.cuda.current_device(), torch.cuda.current_stream(dev).cuda_stream)


def _pooled_bytes(dev) -> int:
    buf = _HISTORY_POOL.get(_pool_key(dev))
    return 0 if buf is None else buf.numel() * 4


def _take_history(nelem: int, dev) -> torch.Tensor:
    buf = _HISTORY_POOL.pop(_pool_key(dev), None) if _pool_on() else None
    if buf is not None and buf.numel() >= nelem:
        return buf
    del buf                                   # too small: give it back before asking for a bigger one
    return torch.empty(nelem, dtype=torch.float32, device=dev)


def _give_history(buf: torch.Tensor) -> None:
    if not _pool_on() or buf is None:
        return
    key = _pool_key(buf.device)
    cur = _HISTORY_POOL.get(key)
    if cur is None or cur.numel() < buf.numel():
        _HISTORY_POOL[key] = buf


def release_buffers() -> None:
    """Drop the idle wavefield-history buffers the engine keeps between calls (returns the memory to torch)."""
    _HISTORY_POOL.clear()


class _Propagate(torch.autograd.Function):
    """records[nt, R, nchan] = F(amp[nt, ns], *coefs)."""

    @staticmethod
    @on_device_of(3)
    def forward(ctx, spec: Spec, acq: Acquisition, amp: torch.Tensor, *coefs: torch.Tensor):
        dev = ctx.dev
        if acq.device != dev:
            raise RuntimeError(f"seistorch_b200: acquisition tables live on {acq.device}, the model on {dev}")
        amp32 = amp.detach().to(torch.float32).contiguous()
        if tuple(amp32.shape) != (spec.nt, acq.ns):
            raise ValueError(f"amp has shape {tuple(amp32.shape)}, expected {(spec.nt, acq.ns)}")
        coefp = _pack_coefs(spec, coefs)
        nchan = len(spec.chan_f)
        rec = torch.zeros((spec.nt, acq.R, nchan), dtype=torch.float32, device=dev)
        need_grad = any(ctx.needs_input_grad[2:])
        p = spec.order
        ctx.illum = spec.illum_plane if spec.illum_chan is not None else None
        if ctx.illum is not None and not need_grad:
            raise NotImplementedError("seistorch_b200: source illumination is accumulated from the wavefield history of a "
                                      "gradient run (mode 'inversion' with a parameter that requires grad); a forward-only "
                                      "run keeps no history")
        if not need_grad:
            prob = _Problem(spec, coefp, acq, amp32, p + 1)
            prob.rec_out = rec
            prob.forward(0, spec.nt, 0)
            return rec
        K, nseg = _history_plan(spec, dev)
        u = _take_history((K + p) * spec.slot_elems, dev)
        u[:(p + 1) * spec.slot_elems].zero_()            # initial state; every later slot is written in full by the kernels
        prob = _Problem(spec, coefp, acq, amp32, K + p, u=u)
        prob.rec_out = rec
        ckpts = []
        slot0_last = 0
        for s in range(nseg):
            a, b = s * K, min((s + 1) * K, spec.nt)
            slot0 = a % prob.nslots
            if nseg > 1 and s < nseg - 1:
                ckpts.append((prob.slot_view(slot0, 1).clone(), prob.slot_view(slot0 + 1, 1).clone() if p == 2 else None))
            prob.forward(a, b - a, slot0)
            slot0_last = slot0
        ctx.spec, ctx.acq, ctx.prob = spec, acq, prob
        ctx.K, ctx.nseg, ctx.ckpts, ctx.slot0_last = K, nseg, ckpts, slot0_last
        ctx.ncoef = len(coefs)
        ctx.coef_dtypes = [c.dtype for c in coefs]
        ctx.amp_dtype = amp.dtype
        return rec

    @staticmethod
    @on_device_of(1)
    def backward(ctx, grad_rec):
        if getattr(ctx, "prob", None) is None:
            raise RuntimeError("seistorch_b200: the wavefield history of this forward call was already consumed and "
                               "freed by a previous backward(); a second backward / retain_graph=True is not "
                               "supported on the whole-loop path -- call forward again")
        spec, acq, prob = ctx.spec, ctx.acq, ctx.prob
        dev = prob.u.device
        p = spec.order
        K, nseg = ctx.K, ctx.nseg
        prob.rec_adj = grad_rec.detach().to(torch.float32).contiguous()
        prob.lam = torch.zeros(spec.nlam * spec.slot_elems, dtype=torch.float32, device=dev)
        prob.bchunk = _choose_bchunk(spec)
        nchunk = math.ceil(spec.B / prob.bchunk)
        if spec.family == "wave2d":
            nchunk = max(nchunk, spec.B)      # frame blocks accumulate per shot (include/seistorch_b200.h)
        prob.gacc = torch.zeros(nchunk * spec.ngrad * spec.plane, dtype=torch.float32, device=dev)
        want_gamp = ctx.needs_input_grad[2]
        prob.gamp = torch.zeros((spec.nt, acq.ns), dtype=torch.float32, device=dev) if want_gamp else None
        l0 = dict(STEPS)
        for s in reversed(range(nseg)):
            a, b = s * K, min((s + 1) * K, spec.nt)
            if s == nseg - 1:
                slot0 = ctx.slot0_last           # history of the last segment is still in place
            else:
                slot0 = 0
                c0, c1 = ctx.ckpts[s]
                prob.slot_view(0, 1).copy_(c0)
                if p == 2:
                    prob.slot_view(1, 1).copy_(c1)
                prob.forward(a, b - a, 0, record=False)
            if ctx.illum is not None:
                # S_i, i = a .. b-1, sit in slots slot0 + p + (i - a): add their squares (all shots) to the illumination
                with torch.cuda.device(dev):
                    _lib.check(_lib.lib().st_illumination(
                        prob.u.data_ptr(), spec.slot_elems, prob.nslots, (slot0 + p) % prob.nslots, b - a,
                        int(spec.illum_chan) * spec.B * spec.plane, spec.B, spec.plane, ctx.illum.data_ptr(), _stream_ptr()),
                        "illumination")
                LAUNCHES["misfit"] += 1
            if p == 2:
                # Lam_i for i = b-1 .. a ; S_i sits in slot slot0 + 2 + (i - a)
                prob.adjoint(b - 1, b - a, slot0 + 2 + (b - 1 - a))
            else:
                # first-order: Lam_i for i = i_hi .. max(a-1,0); S_{i+1} in slot slot0 + 1 + (i+1-a)
                i_hi = b - 1 if s == nseg - 1 else b - 2
                i_lo = max(a - 1, 0)
                if i_hi >= i_lo:
                    prob.adjoint(i_hi, i_hi - i_lo + 1, slot0 + 1 + (i_hi + 1 - a))
        LAUNCHES_LAST["adjoint"] = STEPS["adjoint"] - l0["adjoint"]
        LAUNCHES_LAST["recompute"] = STEPS["forward"] - l0["forward"]
        g = prob.gacc.view(nchunk, spec.ngrad, *spec.shape[:-1], spec.ld).sum(0)[..., :spec.shape[-1]]
        grads = []
        for k in range(ctx.ncoef):
            if not ctx.needs_input_grad[3 + k]:
                grads.append(None)
                continue
            if spec.family == "wave2d":
                slot = spec.coef_slots[k]
                gi = _W2_GRAD_OF_COEF.get(slot)
                grads.append(None if gi is None else g[gi].to(ctx.coef_dtypes[k]))
            elif spec.family == "elastic2d":
                grads.append(None if k == 0 else g[k - 1].to(ctx.coef_dtypes[k]))
            else:
                if k == 0:
                    # acoustic3d: the kernel accumulates ciso * dL/dciso (st_acoustic3d.cu); divide once
                    ciso = prob.coefp[0].view(*spec.shape[:-1], spec.ld)[..., :spec.shape[-1]]
                    g0 = torch.where(ciso != 0, g[0] / ciso, torch.zeros_like(g[0]))
                    grads.append(g0.to(ctx.coef_dtypes[k]))
                else:
                    grads.append(None)
        gamp = prob.gamp.to(ctx.amp_dtype) if want_gamp else None
        # the history goes back to the engine's pool (next call), everything else is freed eagerly
        _give_history(prob.u)
        prob.u = None
        ctx.prob = None
        return (None, None, gamp, *grads)


def propagate(spec: Spec, acq: Acquisition, amp: torch.Tensor, coefs: Sequence[torch.Tensor]) -> torch.Tensor:
    """Run nt steps; returns seismograms [nt, R, nchan] (differentiable w.r.t. amp, coefs).  With ``spec.illum_chan`` /
    ``spec.illum_plane`` set, backward() also adds the source illumination to that plane (from the wavefield history)."""
    return _Propagate.apply(spec, acq, amp, *coefs)
