"""Per-step compatibility ops: the reference's plug-in convention is a module-level
``_time_step(*model_params, *wavefields, dt, h, d, habcs=None)`` per equation
(SURVEY.md 8b).  Each call below runs ONE fused sm_100a step through the same C ABI as
the whole-loop path, and carries an exact VJP (adjoint kernel), so code that drives the
cell step by step (WaveCell.forward, user loops) keeps working -- on the GPU only.

Returned tuple order = ``Wavefield(eq).wavefields`` (second-order pairs: (next, current)).
"""
from __future__ import annotations

import torch

from . import coefficients as _coef
from .engine import Acquisition, Spec, _Problem, _pack_coefs, _require_cuda, on_device_of
from .eqconfigure import Parameters, Wavefield


def _empty_acq(shape, B, dev):
    z = torch.zeros(0, dtype=torch.int64, device=dev)
    return Acquisition(shape, B, z, z.reshape(0, len(shape)), z, z.reshape(0, len(shape)), dev)


def _to_slots(fields, spec):
    """list of [B,*shape] tensors (one per channel) -> pitched [nf*B*plane] buffer."""
    B, ld, nx = spec.B, spec.ld, spec.shape[-1]
    out = torch.zeros((len(fields), B) + tuple(spec.shape[:-1]) + (ld,), dtype=torch.float32, device=fields[0].device)
    for k, f in enumerate(fields):
        out[k, ..., :nx] = f
    return out


class _Step(torch.autograd.Function):
    """fields_out = step(coefs, fields_in) for one time step (no source, no receivers)."""

    @staticmethod
    @on_device_of(3)
    def forward(ctx, spec: Spec, ncoef: int, *tensors):
        coefs, fields = tensors[:ncoef], tensors[ncoef:]
        for t in tensors:
            _require_cuda(t, "step operand")
        dev = fields[0].device
        acq = _empty_acq(spec.shape, spec.B, dev)
        coefp = _pack_coefs(spec, coefs)
        amp = torch.zeros((1, 0), dtype=torch.float32, device=dev)
        nx = spec.shape[-1]
        if spec.order == 2:
            # fields = (cur_0, prev_0[, cur_1, prev_1]); slots: 0 prev, 1 cur, 2 next
            cur = _to_slots([f.detach() for f in fields[0::2]], spec)
            prev = _to_slots([f.detach() for f in fields[1::2]], spec)
            u = torch.cat([prev.reshape(-1), cur.reshape(-1), torch.zeros_like(cur).reshape(-1)])
            prob = _Problem(spec, coefp, acq, amp, 3, u=u)
            prob.forward(0, 1, 0)
            nxt = prob.slot_view(2).view(spec.nf, spec.B, *spec.shape[:-1], spec.ld)[..., :nx]
            outs = []
            for k in range(spec.nf):
                outs += [nxt[k].clone(), fields[2 * k].detach().clone()]
        else:
            cur = _to_slots([f.detach() for f in fields], spec)
            u = torch.cat([cur.reshape(-1), torch.zeros_like(cur).reshape(-1)])
            prob = _Problem(spec, coefp, acq, amp, 2, u=u)
            prob.forward(0, 1, 0)
            nxt = prob.slot_view(1).view(spec.nf, spec.B, *spec.shape[:-1], spec.ld)[..., :nx]
            outs = [nxt[k].clone() for k in range(spec.nf)]
        ctx.spec, ctx.ncoef, ctx.prob = spec, ncoef, prob
        ctx.dtypes = [t.dtype for t in tensors]
        return tuple(outs)

    @staticmethod
    @on_device_of(1)
    def backward(ctx, *gouts):
        if getattr(ctx, "prob", None) is None:
            raise RuntimeError("seistorch_b200: step buffers already freed by a previous backward(); call the step again")
        spec, prob, ncoef = ctx.spec, ctx.prob, ctx.ncoef
        dev = prob.u.device
        nx, e = spec.shape[-1], spec.slot_elems
        gouts = [torch.zeros((spec.B,) + tuple(spec.shape), device=dev) if g is None else g for g in gouts]
        prob.bchunk = spec.B
        nplanes = spec.B if spec.family == "wave2d" else 1     # frame blocks accumulate per shot
        prob.gacc = torch.zeros(nplanes * spec.ngrad * spec.plane, dtype=torch.float32, device=dev)
        unpad = lambda buf: buf.view(spec.nf, spec.B, *spec.shape[:-1], spec.ld)[..., :nx]
        if spec.order == 2:
            gy = _to_slots(gouts[0::2], spec)
            scale = None
            if spec.family == "acoustic3d":
                # the 3D adjoint kernel works on the scaled cotangent w = ciso * Lam (include/seistorch_b200.h)
                scale = prob.coefp[0].view(*spec.shape[:-1], spec.ld)
                gy = gy * scale
            gy = gy.reshape(-1)
            zero = torch.zeros_like(gy)
            # call A: Lam_{i+1} = g_y, Lam_{i+2} = 0  ->  d/d cur (minus the pass-through) + coefficient grads
            prob.lam = torch.cat([zero, gy, zero])          # slots: 0 out, 1 lam1, 2 lam2
            prob.adjoint(0, 1, 1)                           # S_i in slot 1 (cur), S_{i-1} in slot 0 (prev)
            g_cur = unpad(prob.lam[:e]).clone()
            # call B: Lam_{i+1} = 0, Lam_{i+2} = g_y      ->  d/d prev
            gacc = prob.gacc
            prob.gacc = None
            prob.lam = torch.cat([zero, zero, gy])
            prob.adjoint(0, 1, 1)
            g_prev = unpad(prob.lam[:e]).clone()
            prob.gacc = gacc
            if scale is not None:
                inv = torch.where(scale != 0, 1.0 / scale, torch.zeros_like(scale))[..., :nx]
                g_cur, g_prev = g_cur * inv, g_prev * inv
            gfields = []
            for k in range(spec.nf):
                gfields += [g_cur[k] + gouts[2 * k + 1], g_prev[k]]
            g = prob.gacc.view(nplanes, spec.ngrad, *spec.shape[:-1], spec.ld).sum(0)[..., :nx]
            from .engine import _W2_GRAD_OF_COEF
            gcoefs = []
            for k in range(ncoef):
                if spec.family == "wave2d":
                    gi = _W2_GRAD_OF_COEF.get(spec.coef_slots[k])
                    gcoefs.append(None if gi is None else g[gi])
                else:
                    # acoustic3d: the kernel accumulates ciso * dL/dciso; divide once (st_acoustic3d.cu)
                    gcoefs.append(g[0] * inv if (k == 0 and scale is not None) else (g[0] if k == 0 else None))
        else:
            # first order: Lam_i from Lam_{i+1} = g_out; S_i in slot 0, S_{i+1} (no source added) in slot 1
            gl = _to_slots(gouts, spec).reshape(-1)
            prob.lam = torch.cat([torch.zeros_like(gl), gl])       # slot i mod 2: i=0 -> out in slot 0, lam1 in slot 1
            prob.p.nt = 2                                          # so that Lam_{i+1} (i+1 = 1) counts as live
            spec2 = spec
            prob.spec = Spec(**{**spec2.__dict__, "nt": 2})
            prob.amp = torch.zeros((2, 0), dtype=torch.float32, device=dev)
            prob.adjoint(0, 1, 1)
            gfields = list(unpad(prob.lam[:e]).clone())
            g = prob.gacc.view(spec.ngrad, *spec.shape[:-1], spec.ld)[..., :nx]
            gcoefs = [None] + [g[k] for k in range(4)]
        ctx.prob = None
        grads = [None if g is None else g.to(dt) for g, dt in zip(gcoefs + gfields, ctx.dtypes)]
        return (None, None, *grads)


def time_step(equation, ndim, *args, **kwargs):
    """Generic ``_time_step`` body shared by the equation modules."""
    multiple = False
    habcs = kwargs.get("habcs")
    names = Parameters.valid_model_paras()[equation]
    npar = len(names)
    nfld = len(Wavefield(equation).wavefields)
    params = list(args[:npar])
    fields = list(args[npar:npar + nfld])
    dt, h, d = args[npar + nfld:npar + nfld + 3]
    if habcs is not None and habcs[0] is None:
        multiple = True                      # acoustic_habc.py:157: `multiple = tmidx is None`
    shape = tuple(fields[0].shape[1:])
    B = fields[0].shape[0]
    dtf, hf = float(dt), float(h)
    if ndim == 3:
        coefs = _coef.acoustic3d_coefficients(params, dtf, hf, d)
        spec = Spec("acoustic3d", 0, shape, B, 1, dtf)
    elif equation == "elastic":
        coefs = _coef.elastic_coefficients(params, dtf, hf, d)
        spec = Spec("elastic2d", 0, shape, B, 1, dtf)
    else:
        family, flags = _coef.EQUATIONS[equation]
        coefs, slots = _coef.wave2d_coefficients(equation, params, dtf, hf, d)
        # absorbing width = depth of the one-way masks the caller built (cell.setup_habc: bottom mask is [B, bw, nx]);
        # the whole-loop path passes geom.bwidth, the two agree by construction
        bw = 50
        if habcs is not None and habcs[1] is not None:
            bw = int(habcs[1].shape[-2])
        spec = Spec(family, flags, shape, B, 1, dtf, bw=bw, multiple=multiple, coef_slots=slots)
    return _Step.apply(spec, len(coefs), *coefs, *fields)


def reverse_step_unavailable(equation):
    def _time_step_backward(*args, **kwargs):
        raise NotImplementedError(
            f"seistorch_b200: '{equation}._time_step_backward' (the reference's reverse-time boundary-saving "
            "reconstruction, used only by its CheckpointFunction) is superseded by the exact adjoint kernels; "
            "use WaveRNN.forward / loss.backward().")
    return _time_step_backward
