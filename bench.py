#!/usr/bin/env python
"""bench.py -- the hot path's headline metric on B200 (contract in the task brief).

Default workload (BASELINE.json configs[1], `--config cfg2`): 2D acoustic FWI gradient, HABC, synthetic
Marmousi-size 2301x751 model (padded 2401x851), nt = 2000, 128 shots shot-parallel on 8 GPUs = 16 shots per GPU
(weak scaling: per-GPU work fixed as N grows).  One "step" = one FWI-gradient evaluation of this rank's shots:
    forward modelling -> L2 misfit -> exact adjoint -> (N>1) one NCCL all-reduce of the model gradients.
Every other BASELINE configuration is selectable (`--config cfg1|cfg3|cfg4|cfg4_tti|cfg4_fwim|cfg5`, inputs as
SURVEY.md 8d specifies them); cfg1 is forward modelling of one shot (no gradient), as BASELINE words it.

    python bench.py --gpus 1 --steps 5 --warmup 3 [--config cfgK]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...      # CPU arm: the reference's own code (oracle/_ref) on the host cores

Prints ONE JSON line (rank 0).  At N = 1 the line also carries `parity` (CUDA path vs the float64 oracle on a short
horizon of the same workload, `--no-check` skips it), `cpu_baseline` and `roofline`.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

DT, H, FM, DELAY = 1e-3, 10.0, 10.0, 150
SEED = 20230503
NZ, NX, NT = 751, 2301, 2000          # cfg2 model grid (kept as module constants: tests import them)
SHOTS_PER_GPU = 16


def ricker(fm, dt, nt, delay):
    from seistorch_b200.utils import ricker_wave
    return ricker_wave(fm, dt, nt, delay, dtype="numpy")


# ----------------------------------------------------------------------------- workloads
def _vp_recipe(nz, nx, sigma_init=20):
    """vp = 1500 + 3000 z/nz + smoothed noise (sigma 8 cells, 200 m/s), clipped to [1500, 4700]; the initial model
    is a heavily smoothed copy."""
    rng = np.random.default_rng(SEED)
    from scipy.ndimage import gaussian_filter
    z = np.linspace(0.0, 1.0, nz, dtype=np.float32)[:, None]
    noise = gaussian_filter(rng.standard_normal((nz, nx)).astype(np.float32), sigma=8)
    noise *= 200.0 / max(float(np.abs(noise).max()), 1e-6)
    true = np.clip(1500.0 + 3000.0 * z + noise, 1500.0, 4700.0).astype(np.float32)
    init = gaussian_filter(true, sigma=sigma_init).astype(np.float32)
    return true, init


def make_models(nz=NZ, nx=NX):
    """cfg2 (true, initial) vp; kept under this name for the tests."""
    return _vp_recipe(nz, nx)


def _norm(a):
    return ((a - a.min()) / max(float(a.max() - a.min()), 1e-6)).astype(np.float32)


def _models_cfg1():
    vp = np.empty((150, 300), np.float32)
    vp[:50], vp[50:100], vp[100:] = 1500.0, 2000.0, 2500.0
    return {"vp": vp}, {"vp": vp}


def _models_cfg2():
    t, i = _vp_recipe(NZ, NX)
    return {"vp": t}, {"vp": i}


def _models_cfg3():
    t, i = _vp_recipe(400, 1000)
    mk = lambda v: {"vp": v, "vs": (v / 1.73).astype(np.float32), "rho": np.full_like(v, 2000.0)}
    return mk(t), mk(i)


def _models_cfg4(tti=False):
    t, i = _vp_recipe(500, 1200)
    def mk(v, with_m):
        vn = _norm(v)
        d = {"vp": v, "epsilon": (0.1 * vn).astype(np.float32), "delta": (0.05 * vn).astype(np.float32)}
        if tti:
            d["theta"] = np.full_like(v, 15.0)
        m = np.zeros_like(v)
        if with_m:
            m[1:] = v[1:] - v[:-1]
            m /= max(float(np.abs(m).max()), 1e-6)
        d["m"] = m.astype(np.float32)
        return d
    return mk(t, True), mk(i, False)


def _models_cfg4_fwim():
    t, i = _vp_recipe(500, 1200)
    def mk(v, refl):
        rz, rx = np.zeros_like(v), np.zeros_like(v)
        if refl:
            rz[1:] = (v[1:] - v[:-1]) / v[1:] * 0.02
            rx[:, 1:] = (v[:, 1:] - v[:, :-1]) / v[:, 1:] * 0.02
        return {"vp": v, "rx": rx.astype(np.float32), "rz": rz.astype(np.float32)}
    return mk(t, True), mk(i, False)


def _models_cfg5(shape=(400, 200, 400)):
    nx, nz, ny = shape
    vp = np.full((nx, nz, ny), 1500.0, np.float32)
    vp[:, nz // 2:, :] = 2000.0
    from scipy.ndimage import gaussian_filter1d
    init = gaussian_filter1d(vp, sigma=max(nz / 25.0, 1.0), axis=1).astype(np.float32)
    return {"vp": vp}, {"vp": init}


WORKLOADS = {
    # name: equation, model builder, boundary, nt, shots per GPU, total shots on 8 GPUs, mode, src / rec fields, inverted,
    #       algorithmic bytes per grid point per step (forward, adjoint: DESIGN.md 5), fields per state
    "cfg1": dict(index=0, equation="acoustic", models=_models_cfg1, boundary="pml", nt=2000, shots=1, mode="forward",
                 st=["h1"], rt=["h1"], inv=[], fwd_bytes=20, adj_bytes=32, nf=1,
                 desc="configs[0]: 2D scalar acoustic forward modelling, PML, 1 shot, layered 300x150 (padded 400x250), nt=2000"),
    "cfg2": dict(index=1, equation="acoustic_habc", models=_models_cfg2, boundary="habc", nt=2000, shots=16, mode="gradient",
                 st=["h1"], rt=["h1"], inv=["vp"], fwd_bytes=20, adj_bytes=32, nf=1,
                 desc="configs[1]: 2D acoustic FWI gradient, HABC, 2301x751 (padded 2401x851), nt=2000, 16 shots/GPU, L2 misfit"),
    "cfg3": dict(index=2, equation="elastic", models=_models_cfg3, boundary="pml", nt=2000, shots=8, mode="gradient",
                 st=["vz"], rt=["vx", "vz"], inv=["vp", "vs", "rho"], fwd_bytes=56, adj_bytes=96, nf=5,
                 desc="configs[2]: 2D elastic velocity-stress FWI (vp/vs/rho), PML, 1000x400 (padded 1100x500), nt=2000, 8 shots/GPU, L2"),
    "cfg4": dict(index=3, equation="acoustic_vti_lsrtm_habc", models=lambda: _models_cfg4(False), boundary="habc", nt=2000,
                 shots=12, mode="gradient", st=["p1"], rt=["sp1"], inv=["vp", "m"], fwd_bytes=44, adj_bytes=76, nf=2,
                 desc="configs[3]: 2D VTI qP LSRTM, joint gradient (vp, m), HABC, 1200x500 (padded 1300x600), nt=2000, 12 shots/GPU, L2"),
    "cfg4_tti": dict(index=3, equation="acoustic_tti_lsrtm_habc", models=lambda: _models_cfg4(True), boundary="habc", nt=2000,
                     shots=12, mode="gradient", st=["p1"], rt=["sp1"], inv=["vp", "m"], fwd_bytes=48, adj_bytes=80, nf=2,
                     desc="configs[3]: 2D TTI qP LSRTM, joint gradient (vp, m), HABC, 1200x500 (padded 1300x600), nt=2000, 12 shots/GPU, L2"),
    "cfg4_fwim": dict(index=3, equation="acoustic_fwim_habc", models=_models_cfg4_fwim, boundary="habc", nt=2000, shots=12,
                      mode="gradient", st=["h1"], rt=["h1"], inv=["vp", "rx", "rz"], fwd_bytes=28, adj_bytes=56, nf=1,
                      desc="configs[3]: joint FWI-LSRTM equation (vp, rx, rz), HABC, 1200x500 (padded 1300x600), nt=2000, 12 shots/GPU, L2"),
    "cfg5": dict(index=4, equation="acoustic", models=_models_cfg5, boundary="pml", nt=1000, shots=4, mode="gradient",
                 st=["h1"], rt=["h1"], inv=["vp"], fwd_bytes=20, adj_bytes=32, nf=1,
                 desc="configs[4]: 3D acoustic FWI gradient, PML, 400x400x200 (padded 500x300x500), nt=1000, 4 shots/GPU, "
                      "K-step checkpointed wavefield reconstruction, L2"),
}


def padded_shape(models, multiple=False):
    a = next(iter(models.values()))
    return tuple(int(s) + 100 for s in a.shape)


def make_case(nshots, first_shot=0, total_shots=None, nz=None, nx=None, nt=None, vp=None, workload="cfg2", models=None,
              delay=DELAY):
    """A case dict (schema of oracle/ref_runner.run_reference) for `nshots` shots of one workload: sources on a regular
    x-grid at depth index 1, receivers every 2nd x-cell at depth index 1 (3D: 8x4 source grid, 4-cell receiver lattice)."""
    w = WORKLOADS[workload]
    nt = nt or w["nt"]
    total = total_shots or nshots
    if models is None:
        if vp is None:
            raise ValueError("make_case needs models= (or vp= for the acoustic workloads)")
        models = {"vp": vp}
    arr = next(iter(models.values()))
    if arr.ndim == 2:
        mnz, mnx = arr.shape
        if workload == "cfg1":
            xs = np.array([mnx // 2], dtype=np.float64)
        else:
            xs = np.linspace(20, mnx - 21, total)
        sources = [[float(x), 1.0] for x in xs[first_shot:first_shot + nshots]]
        rx = list(range(0, mnx, 2))
        receivers = [[rx, [1] * len(rx)] for _ in range(nshots)]
    else:
        mnx, mnz, mny = arr.shape                      # model file (nx, nz, ny), doc/data_format.md:7
        gx, gy = (8, 4) if total >= 32 else (max(total, 1), 1)
        px, py = np.meshgrid(np.linspace(20, mnx - 21, gx), np.linspace(20, mny - 21, gy) if gy > 1 else [mny / 2.0], indexing="ij")
        pts = np.stack([px.ravel(), py.ravel()], 1)[first_shot:first_shot + nshots]
        sources = [[float(x), float(y), 1.0] for x, y in pts]
        lx, ly = np.meshgrid(np.arange(0, mnx, 4), np.arange(0, mny, 4), indexing="ij")
        receivers = [[lx.ravel().tolist(), ly.ravel().tolist(), [1] * lx.size] for _ in range(nshots)]
    return dict(equation=w["equation"], models=models, invlist={k: True for k in w["inv"]}, sources=sources,
                receivers=receivers, nt=nt, dt=DT, h=H, wavelet=ricker(FM, DT, nt, delay),
                source_type=list(w["st"]), receiver_type=list(w["rt"]), boundary=w["boundary"], multiple=False)


def workload_config(n, name="cfg2", shots=None, nt=None):
    w = WORKLOADS[name]
    shots = shots or w["shots"]
    t, _ = w["models"]() if name != "cfg2" else ({"vp": np.empty((NZ, NX), np.float32)}, None)
    pshape = list(padded_shape(t))
    if len(pshape) == 3:                                # tensor layout (x, z, y)
        pshape = [pshape[0], pshape[1], pshape[2]]
    return {"workload": w["desc"], "equation": w["equation"], "grid_padded": pshape, "nt": nt or w["nt"],
            "shots_per_gpu": shots, "shots_total": shots * n,
            "parallelism": f"shot-parallel x{n}" + (", one NCCL all-reduce of the model gradients per step" if w["mode"] == "gradient" else ""),
            "l2_flush": "not needed: per-step working set (wavefield history, tens of GB) >> 126 MB L2" if w["mode"] == "gradient"
                        else "state is meant to stay on chip (persistent multi-step kernel); records are rewritten every step"}


# ----------------------------------------------------------------------------- clocks
STEP_MS = {}          # per-step device times of the last timed() legs (reported as `step_ms`)


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        if os.environ.get("SEISTORCH_B200_BENCH_NO_SAMPLER"):        # diagnostics only: is the polling itself visible in the step times?
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", "200"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                smax = float(r[2])
                for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


# ----------------------------------------------------------------------------- CPU arm
CPU_STEPS = {"cfg1": 2000, "cfg2": 100, "cfg3": 100, "cfg4": 100, "cfg4_tti": 100, "cfg4_fwim": 100, "cfg5": 4}


class CpuArm:
    """The reference's OWN implementation of the path on the host cores: the unmodified package shipped as oracle/_ref
    (oracle/make_ref.py), driven through its public API exactly like seistorch_dist.py:92-258 -- build_model,
    reset_geom, model(x), Loss('l2'), backward (pure AD, boundary_saving off: the reference's numerically valid
    gradient path) -- on a bounded sample of the workload: ONE shot, `nt_cpu` time steps of the full grid
    (the per-step cost does not depend on nt), extrapolated linearly to the workload's nt.  If oracle/_ref is absent the
    oracle port (same torch ops, pinned bit-exact to the reference) is timed instead and `kind` says "port"."""

    def __init__(self, name, nt_cpu=None, threads=None):
        import torch
        from oracle import ref_shim
        self.name, self.w = name, WORKLOADS[name]
        self.nt_cpu = int(nt_cpu or CPU_STEPS[name])
        self.threads = threads or os.cpu_count() or 1
        torch.set_num_threads(self.threads)
        _true, init = self.w["models"]()
        self.case = make_case(1, workload=name, models=init, nt=self.nt_cpu,
                              delay=DELAY if self.nt_cpu > 400 else min(30, self.nt_cpu // 2))
        self.grad = self.w["mode"] == "gradient"
        self.kind = "reference" if ref_shim.reference_available() else "port"
        if self.kind == "reference":
            from oracle import ref_runner
            self.cfg, self.model, self.x = ref_runner.build_reference(self.case, "float32", want_grad=self.grad, device="cpu")
            from seistorch.loss import Loss
            self.crit = Loss("l2").loss(self.cfg)

    def sample(self):
        """One timed sample; returns seconds."""
        import torch
        t0 = time.perf_counter()
        if self.kind == "reference":
            if self.grad:
                self.model.train()
                for p in self.model.parameters():
                    p.grad = None
                syn = self.model(self.x)
                s = torch.stack(list(syn), 0)
                loss = self.crit(s, torch.zeros_like(s))
                loss.backward()
            else:
                with torch.no_grad():
                    self.model(self.x)
        else:
            from oracle import loop, misfit
            if self.grad:
                recs, _params = loop.simulate(self.case, dtype=torch.float32, requires_grad=list(self.w["inv"]))
                misfit.l2(recs, [torch.zeros_like(r) for r in recs]).backward()
            else:
                with torch.no_grad():
                    loop.simulate(self.case, dtype=torch.float32)
        return time.perf_counter() - t0

    def result(self, seconds):
        per_step = seconds / self.nt_cpu
        nt = self.w["nt"]
        value = 1.0 / (per_step * nt)
        pshape = "x".join(str(s) for s in padded_shape(self.case["models"]))
        what = "forward + loss + AD backward" if self.grad else "forward modelling"
        extra = "" if self.nt_cpu == nt else f", extrapolated linearly to nt = {nt}"
        return {"value": value, "unit": "shots/s", "cores": self.threads, "kind": self.kind,
                "sample": f"1 shot x {self.nt_cpu} of {nt} steps on the full {pshape} padded grid, {what}, "
                          f"{per_step * 1e3:.1f} ms/step{extra}"}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = WORKLOADS[args.config]
    arm = CpuArm(args.config, nt_cpu=args.cpu_steps)
    # keep the whole run within a few minutes: the first sample is the warm-up; later samples are skipped once
    # the time budget is spent (the per-sample cost is printed in `sample`)
    budget = args.cpu_budget
    t_start = time.perf_counter()
    secs = []
    for k in range(max(args.warmup, 1) + max(args.steps, 1)):
        if k > max(args.warmup, 1) and time.perf_counter() - t_start > budget:
            break
        s = arm.sample()
        if k >= min(max(args.warmup, 1), 1):
            secs.append(s)
    res = arm.result(float(np.mean(secs)))
    res["sample"] += f"; mean of {len(secs)} samples"
    value = res["value"]
    shots = args.shots or w["shots"]
    line = {"impl": "reference", "metric": metric_name(args.config), "value": value, "unit": "shots/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 / value * shots,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.gpus, args.config, shots, args.nt),
            "cpu_baseline": res,
            "e2e": {"value": value, "unit": "shots/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def metric_name(config):
    return "forward_shots_per_s" if WORKLOADS[config]["mode"] == "forward" else "fwi_gradient_shots_per_s"


# ----------------------------------------------------------------------------- parity check (checker, not product)
def parity_check(name, dev, mb):
    """CUDA path vs the float64 oracle (oracle/loop.py, pinned to the reference by tests/golden) on a short horizon of
    the SAME workload: same grid, same acquisition, the timed path's micro-batch size, so the same kernels and tile
    schedule run; records of shot 0 and the gradient of the L2 misfit of shot 0's records are compared.
    cfg5 (75 M cells/shot) cannot be held by the AD oracle: a (48,40,48) clone is checked instead and said so."""
    import torch
    import seistorch_b200 as sb
    from oracle import loop, misfit
    w = WORKLOADS[name]
    t0 = time.perf_counter()
    if name == "cfg5":
        _t, init = _models_cfg5((48, 40, 48))
        nt, note, nb = 60, "reduced clone (48,40,48), nt=60: the AD oracle cannot hold 500x300x500", 2
    elif name == "cfg1":
        _t, init = w["models"]()
        nt, note, nb = 2000, "full workload, nt=2000", 1
    else:
        _t, init = w["models"]()
        if "m" in init or "rx" in init:
            init = _t                 # LSRTM / FWIM: the initial reflectivity is zero (zero scattered field); check on the true model
        nt, note, nb = (100 if name == "cfg2" else 80), None, mb
    delay = DELAY if nt >= 1000 else 25
    case_b = make_case(nb, total_shots=max(w["shots"], nb), workload=name, models=init, nt=nt, delay=delay)
    cfg, model = sb.model_from_case(case_b, device=dev, mode="inversion" if w["inv"] else "forward")
    x = torch.as_tensor(case_b["wavelet"], device=dev).unsqueeze(0)
    out = {"nt": nt, "shots_in_batch": nb, "vs": "float64 oracle (oracle/loop.py)", "tol": {"rec": 1e-5, "grad": 1e-4}}
    if note:
        out["note"] = note
    rel = lambda a, b: float(np.linalg.norm(np.asarray(a, np.float64) - b) / max(np.linalg.norm(b), 1e-300))
    case_0 = dict(case_b, sources=case_b["sources"][:1], receivers=case_b["receivers"][:1])
    if w["inv"]:
        syn = model(x)
        (syn[0].double() ** 2).sum().backward()
        orecs, params = loop.simulate(case_0, dtype=torch.float64, requires_grad=list(w["inv"]))
        misfit.l2(orecs, [torch.zeros_like(r) for r in orecs]).backward()
        out["rec_err"] = rel(syn[0].detach().cpu().numpy(), orecs[0].detach().numpy())
        out["grad_err"] = {k: rel(getattr(model.cell.geom, k).grad.cpu().numpy(), params[k].grad.numpy()) for k in w["inv"]}
        out["ok"] = bool(out["rec_err"] < 1e-5 and all(v < 1e-4 for v in out["grad_err"].values()))
    else:
        with torch.no_grad():
            syn = model(x)
            orecs, _ = loop.simulate(case_0, dtype=torch.float64)
            # SURVEY 8c three-number protocol: new32 vs ref64, ref32 vs ref64, new32 vs ref32
            o32, _ = loop.simulate(case_0, dtype=torch.float32)
        a, b64, b32 = syn[0].cpu().numpy(), orecs[0].numpy(), o32[0].numpy().astype(np.float64)
        out["rec_err"] = rel(a, b64)
        out["ref32_vs_ref64"] = rel(b32, b64)
        out["new32_vs_ref32"] = rel(a, b32)
        out["grad_err"] = None
        out["ok"] = bool(out["rec_err"] < 1e-5)
    out["seconds"] = round(time.perf_counter() - t0, 1)
    del model
    torch.cuda.empty_cache()
    return out


# ----------------------------------------------------------------------------- GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--cpu-steps", type=int, default=0, help="time steps in the bounded CPU sample (0 = per-workload default)")
    ap.add_argument("--cpu-budget", type=float, default=150.0, help="seconds after which the CPU arm stops taking samples")
    ap.add_argument("--microbatch", type=int, default=0, help="shots per forward call (0 = auto)")
    ap.add_argument("--shots", type=int, default=0, help="shots per GPU (0 = workload)")
    ap.add_argument("--nt", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-check", action="store_true", help="skip the parity block (CUDA vs float64 oracle)")
    ap.add_argument("--check", action="store_true", help="(default at N=1) kept for explicitness")
    args = ap.parse_args()
    args.cpu_steps = args.cpu_steps or None
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    import torch.distributed as dist
    import seistorch_b200 as sb
    from seistorch_b200 import engine, parallel
    from seistorch_b200.coords import merge_receivers_with_same_keys, merge_sources_with_same_keys
    from seistorch_b200.probe import WaveProbe
    from seistorch_b200.source import WaveSource

    name = args.config
    w = WORKLOADS[name]
    gradient = w["mode"] == "gradient"
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    nshots, nt = args.shots or w["shots"], args.nt or w["nt"]
    true, init = w["models"]()
    pshape = padded_shape(true)
    npts = int(np.prod(pshape))
    ld = (pshape[-1] + 3) // 4 * 4
    state_bytes = w["nf"] * int(np.prod(pshape[:-1])) * ld * 4

    # micro-batch: largest divisor of nshots whose full wavefield history fits (no recompute); when not even one
    # shot's history fits (3D) the engine switches to K-step checkpoints + recomputation by itself
    free, total = torch.cuda.mem_get_info(dev)
    free += torch.cuda.memory_reserved(dev) - torch.cuda.memory_allocated(dev)
    if args.microbatch:
        mb = args.microbatch
    else:
        mb = 1
        for c in range(1, nshots + 1):
            if nshots % c == 0 and (not gradient or (nt + 8) * c * state_bytes < 0.80 * free):
                mb = c

    parity = None
    if rank == 0 and world == 1 and not args.no_check:
        try:
            parity = parity_check(name, dev, mb)
        except Exception as e:                                         # pragma: no cover
            parity = {"ok": False, "error": f"{type(e).__name__}: {e}"}
        sb.release_buffers()                                           # the check's (small) history buffer is not the workload's

    # ---- set-up (untimed): observed data = forward modelling of the true model with our path
    total_shots = nshots * world if name != "cfg1" else 1
    case_true = make_case(nshots, first_shot=rank * nshots if name != "cfg1" else 0, total_shots=total_shots, workload=name,
                          models=true, nt=nt)
    cfg, fwd_model = sb.model_from_case(case_true, device=dev, mode="forward")
    wav = torch.as_tensor(case_true["wavelet"], device=dev).unsqueeze(0)
    wav_host = torch.as_tensor(case_true["wavelet"]).unsqueeze(0).pin_memory()
    sources, probes = list(fwd_model.sources), list(fwd_model.probes)

    def batch_modules(lo, hi):
        bs, sk = merge_sources_with_same_keys(sources[lo:hi])
        ss = WaveSource(bs, True, **sk).to(dev)
        rc, br, rk = merge_receivers_with_same_keys(probes[lo:hi])
        pp = WaveProbe(br, **rk).to(dev)
        pp.reccounts = rc
        return ss, pp

    batches = [(lo, min(lo + mb, nshots)) + batch_modules(lo, min(lo + mb, nshots)) for lo in range(0, nshots, mb)]
    with torch.no_grad():
        obs_parts = [torch.stack(list(fwd_model(wav, None, ss, pp)), 0) for (_lo, _hi, ss, pp) in batches]
    obs_host = torch.cat(obs_parts, 0).cpu().pin_memory()                        # [B, nt, nrec, nchan]
    del obs_parts
    if gradient:
        del fwd_model
        case = dict(case_true, models=init)
        cfg, model = sb.model_from_case(case, device=dev, mode="inversion")
    else:
        model = fwd_model
    geom = model.cell.geom
    inv_params = [getattr(geom, k) for k in w["inv"]]
    all_params = [getattr(geom, k) for k in geom.model_parameters]
    par_host = [p.detach().cpu().pin_memory() for p in all_params]
    grad_host = [torch.empty_like(p.detach().cpu()).pin_memory() for p in inv_params]
    rec_host = torch.empty_like(obs_host).pin_memory() if not gradient else None
    crit = sb.Loss("l2").loss(cfg)

    # e2e leg: the observed data of every micro-batch travels host -> device on a copy stream while the
    # previous micro-batch is being propagated (still inside the timed region, every step)
    copy_stream = torch.cuda.Stream(device=dev)
    obs_stage = torch.empty_like(obs_host, device=dev) if gradient else None
    copy_done = [torch.cuda.Event() for _ in batches]

    def step(obs_dev, e2e=False):
        """one FWI-gradient evaluation (cfg1: one forward modelling) of this rank's shots."""
        x = wav
        if e2e:
            for p, ph in zip(all_params, par_host):
                p.data.copy_(ph, non_blocking=True)
            if not gradient:
                x = wav_host.to(dev, non_blocking=True)
            else:
                copy_stream.wait_stream(torch.cuda.current_stream(dev))      # the staging buffer is free again
                with torch.cuda.stream(copy_stream):
                    for k, (lo, hi, _, _) in enumerate(batches):
                        obs_stage[lo:hi].copy_(obs_host[lo:hi], non_blocking=True)
                        copy_done[k].record(copy_stream)
        if not gradient:
            with torch.no_grad():
                for k, (lo, hi, ss, pp) in enumerate(batches):
                    syn = model(x, None, ss, pp)
                    if e2e:
                        rec_host[lo:hi].copy_(torch.stack(list(syn), 0), non_blocking=True)
            if e2e:
                torch.cuda.current_stream(dev).synchronize()
                return float(rec_host[0, -1, 0, 0])
            return syn[0][-1, 0, 0]
        for p in inv_params:
            p.grad = None
        total_loss = torch.zeros((), device=dev)
        for k, (lo, hi, ss, pp) in enumerate(batches):
            if e2e:
                torch.cuda.current_stream(dev).wait_event(copy_done[k])
                ob = obs_stage[lo:hi]
            else:
                ob = obs_dev[lo:hi]
            syn = model(x, None, ss, pp)
            loss = crit(torch.stack(list(syn), 0), ob)
            loss.backward()
            total_loss = total_loss + loss.detach()
        if world > 1:
            parallel.allreduce_gradients(inv_params)
        if e2e:
            for gh, p in zip(grad_host, inv_params):
                gh.copy_(p.grad, non_blocking=True)
            return float(total_loss.item())
        return total_loss

    def timed(nsteps, e2e):
        obs_dev = None if (e2e or not gradient) else obs_host.to(dev)
        for _ in range(args.warmup):
            step(obs_dev, e2e)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        marks = [torch.cuda.Event(enable_timing=True) for _ in range(nsteps + 1)]     # one event per step boundary, no sync
        l0 = dict(engine.LAUNCHES)
        m0 = torch.cuda.memory_stats(dev)
        marks[0].record()
        for i in range(nsteps):
            out = step(obs_dev, e2e)
            marks[i + 1].record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = torch.tensor([marks[0].elapsed_time(marks[-1])], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        launches = sum(engine.LAUNCHES[k] - l0[k] for k in l0)
        STEP_MS["e2e" if e2e else "device"] = [round(marks[i].elapsed_time(marks[i + 1]), 3) for i in range(nsteps)]
        m1 = torch.cuda.memory_stats(dev)
        # cudaMalloc calls / allocator cache flushes inside the timed region (0 / 0 in steady state: the engine keeps its history buffer)
        STEP_MS[("e2e" if e2e else "device") + "_cudamallocs"] = int(m1.get("num_device_alloc", 0) - m0.get("num_device_alloc", 0))
        STEP_MS[("e2e" if e2e else "device") + "_alloc_retries"] = int(m1.get("num_alloc_retries", 0) - m0.get("num_alloc_retries", 0))
        return float(ms.item()), launches, out

    # the set-up above leaves millions of long-lived Python objects (acquisition lists, the CPU arm's modules); a full
    # garbage-collection pass over them inside the timed region stalls the launching thread for tens of ms right after the
    # forward's NaN check has drained the launch queue.  Park them in the permanent generation (nothing is skipped: the
    # collector keeps running on the objects the steps create).
    import gc
    gc.collect()
    gc.freeze()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_total, launches, last = timed(args.steps, e2e=False)
    clocks = sampler.stop() if rank == 0 else None
    ms_e2e, _, _ = timed(args.steps, e2e=True)
    ms_step = ms_total / args.steps
    value = nshots * world / (ms_step * 1e-3)
    e2e_value = nshots * world / (ms_e2e / args.steps * 1e-3)

    # ---- per-kernel timing for the roofline (CUDA events on the launching stream)
    roof, fd = None, None
    if rank == 0:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        lo, hi, ss, pp = batches[0]
        B = hi - lo
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
        with torch.no_grad():
            model(wav, None, ss, pp)                                   # warm
            torch.cuda.synchronize()
            ev[0].record()
            model(wav, None, ss, pp)                                   # forward-MODELLING mode (3 rolling slots, no history)
            ev[1].record()
        torch.cuda.synchronize()
        t_mod = ev[0].elapsed_time(ev[1]) / nt                         # ms per launch-equivalent time step
        k_fwd = engine.KERNELS["forward"]
        fd = B * npts / (t_mod * 1e-3) / 1e9
        kernels = {"forward_modelling": {"name": k_fwd, "ms_per_step": t_mod, "algorithmic_bytes_per_pt": w["fwd_bytes"],
                                         "GBps": w["fwd_bytes"] * fd, "frac": w["fwd_bytes"] * fd / peak, "shots_per_launch": B,
                                         "note": "no_grad: 3 rolling state slots, nothing kept"}}
        dom, ach, t_dom = k_fwd, w["fwd_bytes"] * fd, t_mod
        if gradient:
            ob0 = obs_host[lo:hi].to(dev)
            t_fwd, t_bwd = 1e30, 1e30
            for rep in range(3):                                       # pass 0 is untimed (history buffer allocation, plans);
                torch.cuda.synchronize()                               # best of the next two: one launch sequence each
                ev[2].record()
                syn = model(wav, None, ss, pp)                         # the forward of the timed step: writes the history
                ev[3].record()
                loss = crit(torch.stack(list(syn), 0), ob0)
                torch.cuda.synchronize()
                ev[4].record()
                loss.backward()
                ev[5].record()
                torch.cuda.synchronize()
                if rep > 0:
                    t_fwd = min(t_fwd, ev[2].elapsed_time(ev[3]) / nt)
                    t_bwd = min(t_bwd, ev[4].elapsed_time(ev[5]))
                del syn, loss
            nadj = max(engine.LAUNCHES_LAST.get("adjoint", nt), 1)
            nrec = engine.LAUNCHES_LAST.get("recompute", 0)
            t_adj = (t_bwd - nrec * t_fwd) / nadj                      # recomputed forward steps (checkpoints) taken out
            fwd_gbs = w["fwd_bytes"] * B * npts / (t_fwd * 1e-3) / 1e9
            adj_gbs = w["adj_bytes"] * B * npts / (t_adj * 1e-3) / 1e9
            k_adj = engine.KERNELS["adjoint"]
            kernels["forward_history"] = {"name": k_fwd, "ms_per_step": t_fwd, "algorithmic_bytes_per_pt": w["fwd_bytes"],
                                          "GBps": fwd_gbs, "frac": fwd_gbs / peak, "shots_per_launch": B,
                                          "note": "the forward inside the timed step: every state goes to the history buffer"}
            kernels["adjoint"] = {"name": k_adj, "ms_per_step": t_adj, "algorithmic_bytes_per_pt": w["adj_bytes"],
                                  "GBps": adj_gbs, "frac": adj_gbs / peak, "shots_per_launch": B,
                                  "recomputed_forward_steps": nrec}
            dom, ach, t_dom = (k_adj, adj_gbs, t_adj) if t_adj >= t_fwd else (k_fwd, fwd_gbs, t_fwd)
        roof = {"bound": "hbm", "kernel": dom + f"<{w['equation']}>", "achieved": ach, "peak": peak, "unit": "GB/s",
                "frac": ach / peak, "traffic": None, "peak_source": peak_src, "kernels": kernels}
        for prof in ("traffic_r02.json", "traffic_r01.json"):          # per-launch dram bytes from ncu --set full
            path = os.path.join(ROOT, "profiles", prof)
            if os.path.exists(path) and roof["traffic"] is None:
                try:
                    tr = json.load(open(path))
                    roof["traffic"] = tr.get(f"{name}:{dom}", tr.get(dom) if name == "cfg2" else None)
                    if roof["traffic"]:
                        # the same launch time against the MEASURED dram bytes of one launch (a lower bound of the
                        # steady-state traffic: an isolated ncu replay leaves part of the output dirty in L2)
                        roof["traffic_GBps"] = roof["traffic"] / (t_dom * 1e-3) / 1e9
                        roof["traffic_frac"] = roof["traffic_GBps"] / peak
                        roof["traffic_source"] = prof
                except Exception:
                    pass

    if world > 1:
        dist.barrier()
    if rank == 0:
        cpu = None
        if not args.no_cpu_baseline and world == 1:
            try:
                arm = CpuArm(name, nt_cpu=args.cpu_steps)
                arm.sample() if arm.nt_cpu <= 200 else None             # warm-up sample when it is cheap
                cpu = arm.result(arm.sample())
            except Exception as e:                                     # pragma: no cover
                cpu = {"value": None, "unit": "shots/s", "cores": os.cpu_count(), "kind": "reference", "sample": f"failed: {type(e).__name__}: {e}"}
        if gradient:
            h2d = int(obs_host.numel() * 4 + sum(p.numel() for p in par_host) * 4)
            d2h = int(sum(g.numel() for g in grad_host) * 4 + 4)
        else:
            h2d = int(wav_host.numel() * 4 + sum(p.numel() for p in par_host) * 4)
            d2h = int(rec_host.numel() * 4)
        line = {"metric": metric_name(name), "value": value, "unit": "shots/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": workload_config(world, name, nshots, nt), "microbatch_shots": mb, "step_ms": dict(STEP_MS),
                "ms_per_step_median": float(np.median(STEP_MS["device"])) if STEP_MS.get("device") else None,
                "graph": dict(zip(("plain_loops", "captured", "replayed"), sb._lib.graph_counters())),
                "e2e": {"value": e2e_value, "unit": "shots/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
                "gpu_launches": launches, "clocks": clocks, "roofline": roof, "cpu_baseline": cpu, "parity": parity,
                "fd_gpts_per_s": fd, "loss": float(last) if not isinstance(last, float) else last}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
