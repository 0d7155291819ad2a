#!/usr/bin/env python
"""bench.py -- the hot path's headline metric on B200 (contract in the task brief).

Workload (BASELINE.json configs[1]): 2D acoustic FWI gradient, HABC, synthetic
Marmousi-size 2301x751 model (padded 2401x851), nt = 2000, 128 shots shot-parallel on 8
GPUs = 16 shots per GPU (weak scaling: per-GPU work fixed as N grows).
One "step" = one FWI-gradient evaluation of this rank's 16 shots:
    forward modelling -> L2 misfit -> exact adjoint -> (N>1) one NCCL all-reduce of d/dvp.

    python bench.py --gpus 1 --steps 5 --warmup 3
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...      # CPU arm: the oracle port of the reference

Prints ONE JSON line (rank 0).
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

NZ, NX, NT = 751, 2301, 2000          # model grid (unpadded) and time samples
SHOTS_PER_GPU = 16
DT, H, FM, DELAY = 1e-3, 10.0, 10.0, 150
SEED = 20230503
FWD_BYTES_PER_PT = 20                  # read h1,h2,vp,d + write y
ADJ_BYTES_PER_PT = 32                  # read Lam1,Lam2,S_i,r,b + write Lam + gradient read-modify-write


# ----------------------------------------------------------------------------- workload
def make_models(nz=NZ, nx=NX):
    """cfg2 of SURVEY.md 8d: vp = 1500 + 3000 z/nz + smoothed noise, clipped; the initial
    model is a heavily smoothed copy."""
    rng = np.random.default_rng(SEED)
    from scipy.ndimage import gaussian_filter
    z = np.linspace(0.0, 1.0, nz, dtype=np.float32)[:, None]
    noise = gaussian_filter(rng.standard_normal((nz, nx)).astype(np.float32), sigma=8)
    noise *= 200.0 / max(float(np.abs(noise).max()), 1e-6)
    true = np.clip(1500.0 + 3000.0 * z + noise, 1500.0, 4700.0).astype(np.float32)
    init = gaussian_filter(true, sigma=20).astype(np.float32)
    return true, init


def make_case(nshots, first_shot=0, total_shots=None, nz=NZ, nx=NX, nt=NT, vp=None):
    from oracle.loop import ricker_wave
    total = total_shots or nshots
    xs = np.linspace(20, nx - 21, total)
    sources = [[float(x), 1.0] for x in xs[first_shot:first_shot + nshots]]
    rx = list(range(0, nx, 2))
    receivers = [[rx, [1] * len(rx)] for _ in range(nshots)]
    return dict(equation="acoustic_habc", models={"vp": vp}, invlist={"vp": True}, sources=sources,
                receivers=receivers, nt=nt, dt=DT, h=H, wavelet=ricker_wave(FM, DT, nt, DELAY),
                source_type=["h1"], receiver_type=["h1"], boundary="habc", multiple=False)


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
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
def cpu_reference(nt_cpu=12, threads=None):
    """The reference's own CPU algorithm (oracle port: same torch ops in the same order,
    pinned bit-exactly to the reference by tests/golden) on a bounded sample of the
    workload: ONE shot, nt_cpu time steps of the full 2401x851 grid, forward + pure-AD
    backward (the reference's gradient path), extrapolated linearly to nt = 2000."""
    import torch
    from oracle import loop, misfit
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    true, init = make_models()
    case = make_case(1, vp=init, nt=nt_cpu)
    loop.simulate(dict(case, nt=2), dtype=torch.float32)                 # warm-up (allocator, mkldnn)
    t0 = time.perf_counter()
    recs, params = loop.simulate(case, dtype=torch.float32, requires_grad=["vp"])
    loss = misfit.l2(recs, [torch.zeros_like(r) for r in recs])
    loss.backward()
    t = time.perf_counter() - t0
    per_step = t / nt_cpu
    shots_per_s = 1.0 / (per_step * NT)
    return {"value": shots_per_s, "unit": "shots/s", "cores": threads, "kind": "port",
            "sample": f"1 shot x {nt_cpu} of {NT} steps on the full 2401x851 grid, forward + AD backward, "
                      f"{per_step * 1e3:.1f} ms/step, extrapolated linearly in nt",
            "seconds": t}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    vals = []
    for _ in range(args.warmup):
        cpu_reference(nt_cpu=2)
    for _ in range(max(args.steps, 1)):
        vals.append(cpu_reference(nt_cpu=args.cpu_steps))
    best = max(vals, key=lambda v: v["value"])
    value = float(np.mean([v["value"] for v in vals]))
    ms = 1e3 / value
    line = {"impl": "reference", "metric": "fwi_gradient_shots_per_s", "value": value, "unit": "shots/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms * SHOTS_PER_GPU,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.gpus),
            "cpu_baseline": {"value": value, "unit": "shots/s", "cores": best["cores"], "kind": "port", "sample": best["sample"]},
            "e2e": {"value": value, "unit": "shots/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def workload_config(n):
    return {"workload": "configs[1]: 2D acoustic FWI gradient, HABC, 2301x751 (padded 2401x851), nt=2000, "
                        f"{SHOTS_PER_GPU} shots/GPU, L2 misfit",
            "equation": "acoustic_habc", "grid_padded": [851, 2401], "nt": NT, "shots_per_gpu": SHOTS_PER_GPU,
            "shots_total": SHOTS_PER_GPU * n, "parallelism": f"shot-parallel x{n}, one NCCL all-reduce of d/dvp per step",
            "l2_flush": "not needed: per-step working set (wavefield history, tens of GB) >> 126 MB L2"}


# ----------------------------------------------------------------------------- GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-steps", type=int, default=12, help="time steps in the bounded CPU sample")
    ap.add_argument("--microbatch", type=int, default=0, help="shots per forward call (0 = auto)")
    ap.add_argument("--shots", type=int, default=SHOTS_PER_GPU, help="shots per GPU (default = workload)")
    ap.add_argument("--nt", type=int, default=NT)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    import torch.distributed as dist
    import seistorch_b200 as sb
    from seistorch_b200 import engine, parallel
    from seistorch_b200.coords import merge_receivers_with_same_keys, merge_sources_with_same_keys
    from seistorch_b200.probe import WaveProbe
    from seistorch_b200.source import WaveSource

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    nshots, nt = args.shots, args.nt
    true, init = make_models()

    # ---- set-up (untimed): observed data = forward modelling of the true model with our path
    case_true = make_case(nshots, first_shot=rank * nshots, total_shots=nshots * world, vp=true, nt=nt)
    cfg, fwd_model = sb.model_from_case(case_true, device=dev, mode="forward")
    wav = torch.as_tensor(case_true["wavelet"], device=dev).unsqueeze(0)
    with torch.no_grad():
        obs_list = fwd_model(wav)
    obs_host = torch.stack([o for o in obs_list], 0).cpu().pin_memory()          # [B, nt, nrec, 1]
    del fwd_model, obs_list
    case = dict(case_true, models={"vp": init})
    cfg, model = sb.model_from_case(case, device=dev, mode="inversion")
    vp_param = model.cell.geom.vp
    vp_host = vp_param.detach().cpu().pin_memory()
    grad_host = torch.empty_like(vp_host).pin_memory()
    crit = sb.Loss("l2").loss(cfg)

    # micro-batch: largest divisor of nshots whose full wavefield history fits (no recompute)
    free, total = torch.cuda.mem_get_info(dev)
    free += torch.cuda.memory_reserved(dev) - torch.cuda.memory_allocated(dev)
    state_bytes = 851 * 2404 * 4
    if args.microbatch:
        mb = args.microbatch
    else:
        mb = 1
        for c in range(1, nshots + 1):
            if nshots % c == 0 and (nt + 8) * c * state_bytes < 0.80 * free:
                mb = c
    sources, probes = list(model.sources), list(model.probes)

    def batch_modules(lo, hi):
        bs, sk = merge_sources_with_same_keys(sources[lo:hi])
        ss = WaveSource(bs, True, **sk).to(dev)
        rc, br, rk = merge_receivers_with_same_keys(probes[lo:hi])
        pp = WaveProbe(br, **rk).to(dev)
        pp.reccounts = rc
        return ss, pp

    batches = [(lo, min(lo + mb, nshots)) + batch_modules(lo, min(lo + mb, nshots)) for lo in range(0, nshots, mb)]

    # e2e leg: the observed data of every micro-batch travels host -> device on a copy stream while the
    # previous micro-batch is being propagated (still inside the timed region, every step)
    copy_stream = torch.cuda.Stream(device=dev)
    obs_stage = torch.empty_like(obs_host, device=dev)
    copy_done = [torch.cuda.Event() for _ in batches]

    def step(obs_dev, e2e=False):
        """one FWI-gradient evaluation of this rank's shots."""
        if e2e:
            vp_param.data.copy_(vp_host, non_blocking=True)
            copy_stream.wait_stream(torch.cuda.current_stream(dev))      # the staging buffer is free again
            with torch.cuda.stream(copy_stream):
                for k, (lo, hi, _, _) in enumerate(batches):
                    obs_stage[lo:hi].copy_(obs_host[lo:hi], non_blocking=True)
                    copy_done[k].record(copy_stream)
        vp_param.grad = None
        total_loss = torch.zeros((), device=dev)
        for k, (lo, hi, ss, pp) in enumerate(batches):
            if e2e:
                torch.cuda.current_stream(dev).wait_event(copy_done[k])
                ob = obs_stage[lo:hi]
            else:
                ob = obs_dev[lo:hi]
            syn = model(wav, None, ss, pp)
            loss = crit(torch.stack(list(syn), 0), ob)
            loss.backward()
            total_loss = total_loss + loss.detach()
        if world > 1:
            parallel.allreduce_gradients([vp_param])
        if e2e:
            grad_host.copy_(vp_param.grad, non_blocking=True)
            return float(total_loss.item())
        return total_loss

    def timed(nsteps, e2e):
        obs_dev = None if e2e else obs_host.to(dev)
        for _ in range(args.warmup):
            step(obs_dev, e2e)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = dict(engine.LAUNCHES)
        e0.record()
        for _ in range(nsteps):
            out = step(obs_dev, e2e)
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        launches = sum(engine.LAUNCHES[k] - l0[k] for k in l0)
        return float(ms.item()), launches, out

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
        npts = 851 * 2401
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        with torch.no_grad():
            model(wav, None, ss, pp)                                   # warm
            torch.cuda.synchronize()
            ev[0].record()
            model(wav, None, ss, pp)                                   # forward modelling only (3 rolling slots)
            ev[1].record()
        torch.cuda.synchronize()
        t_fwd = ev[0].elapsed_time(ev[1]) / nt                         # ms per forward launch
        fd = B * npts / (t_fwd * 1e-3) / 1e9
        syn = model(wav, None, ss, pp)
        loss = crit(torch.stack(list(syn), 0), obs_host[lo:hi].to(dev))
        torch.cuda.synchronize()
        ev[2].record()
        loss.backward()
        ev[3].record()
        torch.cuda.synchronize()
        t_adj = ev[2].elapsed_time(ev[3]) / nt                         # ms per adjoint launch (misfit + reduction amortised)
        fwd_gbs = FWD_BYTES_PER_PT * B * npts / (t_fwd * 1e-3) / 1e9
        adj_gbs = ADJ_BYTES_PER_PT * B * npts / (t_adj * 1e-3) / 1e9
        from seistorch_b200 import engine as _engine
        dom = (_engine.KERNELS["adjoint"] if t_adj >= t_fwd else _engine.KERNELS["forward"]) + "<ISO|HABC>"
        ach = adj_gbs if t_adj >= t_fwd else fwd_gbs
        roof = {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                "traffic": None, "peak_source": peak_src,
                "forward_kernel": {"name": _engine.KERNELS["forward"], "ms_per_launch": t_fwd, "algorithmic_bytes_per_pt": FWD_BYTES_PER_PT, "GBps": fwd_gbs,
                                   "frac": fwd_gbs / peak, "shots_per_launch": B},
                "adjoint_kernel": {"name": _engine.KERNELS["adjoint"], "ms_per_launch": t_adj, "algorithmic_bytes_per_pt": ADJ_BYTES_PER_PT, "GBps": adj_gbs,
                                   "frac": adj_gbs / peak, "shots_per_launch": B}}
        prof = os.path.join(ROOT, "profiles", "traffic_r01.json")     # per-launch dram bytes from ncu --set full
        if os.path.exists(prof):
            try:
                roof["traffic"] = json.load(open(prof)).get(dom.split("<")[0])
                if roof["traffic"]:
                    # the same launch time against the MEASURED dram bytes of one launch (ncu --set full, profiles/):
                    # a lower bound of the real traffic -- an isolated ncu replay leaves part of the output dirty in L2
                    t_dom = t_adj if t_adj >= t_fwd else t_fwd
                    roof["traffic_GBps"] = roof["traffic"] / (t_dom * 1e-3) / 1e9
                    roof["traffic_frac"] = roof["traffic_GBps"] / peak
            except Exception:
                pass

    if world > 1:
        dist.barrier()
    if rank == 0:
        cpu = None
        if not args.no_cpu_baseline and world == 1:
            try:
                cpu = cpu_reference(nt_cpu=args.cpu_steps)
                cpu.pop("seconds", None)
            except Exception as e:                                     # pragma: no cover
                cpu = {"value": None, "unit": "shots/s", "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {e}"}
        h2d = int(obs_host.numel() * 4 + vp_host.numel() * 4)
        d2h = int(grad_host.numel() * 4 + 4)
        line = {"metric": "fwi_gradient_shots_per_s", "value": value, "unit": "shots/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": dict(workload_config(world), microbatch_shots=mb, nt=nt, shots_per_gpu=nshots),
                "e2e": {"value": e2e_value, "unit": "shots/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
                "gpu_launches": launches, "clocks": clocks, "roofline": roof, "cpu_baseline": cpu,
                "fd_gpts_per_s": fd, "loss": float(last) if not isinstance(last, float) else last}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
