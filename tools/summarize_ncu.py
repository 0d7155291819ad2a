"""Turns the `ncu --set full` captures of tools/profile_r02.sh (gpurun_out/r02/*.ncu-rep) into
profiles/ncu_r02_summary.md (one table row per kernel) and profiles/traffic_r02.json (dram bytes per launch, the
`roofline.traffic` source of bench.py).  Run in the authoring container:  python tools/summarize_ncu.py"""
import csv
import glob
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRCS = [os.path.join(ROOT, "gpurun_out", d) for d in ("r02", "r02b", "r02c", "r02d")]     # later passes override earlier captures
KEYS = {
    "gpu__time_duration.sum": "duration_us",
    "dram__bytes_read.sum": "dram_read_MB",
    "dram__bytes_write.sum": "dram_write_MB",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "launch__registers_per_thread": "regs",
    "launch__grid_size": "grid",
    "launch__waves_per_multiprocessor": "waves",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio": "stall_long_sb",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio": "stall_barrier",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio": "stall_wait",
    "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio": "stall_membar",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "smsp__inst_executed.sum": "warp_inst",
}
# workload of each capture: (bench config, kernel key for traffic_r02.json, cells per launch, algorithmic B/cell)
WORK = {
    "cfg2_adj": ("cfg2", "wave2d_adjoint_tma_kernel", 8 * 851 * 2401, 32),
    "cfg2_fwd": ("cfg2", "wave2d_forward_tma_kernel", 8 * 851 * 2401, 20),
    "cfg1_persist": ("cfg1", "wave2d_persist_forward_kernel", 400 * 250 * 400, 20),
    "cfg3_fwd": ("cfg3", "elastic2d_forward_kernel", 4 * 500 * 1100, 56),
    "cfg3_adj": ("cfg3", "elastic2d_adjoint_kernel", 4 * 500 * 1100, 96),
    "cfg4_fwd": ("cfg4", "wave2d_forward_kernel", 12 * 600 * 1300, 44),
    "cfg4_adj": ("cfg4", "wave2d_adjoint_kernel", 12 * 600 * 1300, 76),
    "cfg4tti_adj": ("cfg4_tti", "wave2d_adjoint_kernel", 12 * 600 * 1300, 80),
    "cfg5_fwd": ("cfg5", "acoustic3d_forward_kernel", 500 * 300 * 500, 20),
    "cfg5_adj": ("cfg5", "acoustic3d_adjoint_kernel", 500 * 300 * 500, 32),
}


def raw(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, val = rows[0], rows[1], rows[2]
    d = {h: (v, u) for h, u, v in zip(hdr, units, val)}
    name = d.get("Kernel Name", ("?", ""))[0]
    res = {"kernel": name}
    for k, short in KEYS.items():
        if k in d:
            v, u = d[k]
            try:
                f = float(v.replace(",", ""))
            except ValueError:
                continue
            if short.endswith("_MB"):
                f = f * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1.0)
            if short == "duration_us":
                f = f * {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3}.get(u, 1.0)
            res[short] = f
    return res


def main():
    rows, traffic = [], {}
    latest = {}
    for src in SRCS:
        for path in sorted(glob.glob(os.path.join(src, "*.ncu-rep"))):
            latest[os.path.basename(path)[:-8]] = path
    for tag, path in sorted(latest.items()):
        r = raw(path)
        r["capture"] = tag
        r["pass"] = os.path.basename(os.path.dirname(path))
        if tag in WORK:
            cfg, key, cells, bpc = WORK[tag]
            r["cells"] = cells
            dram = (r.get("dram_read_MB", 0) + r.get("dram_write_MB", 0)) * 1e6
            r["dram_B_per_cell"] = dram / cells
            r["algorithmic_B_per_cell"] = bpc
            if r.get("duration_us"):
                r["dram_TBps"] = dram / (r["duration_us"] * 1e-6) / 1e12
                r["Gpts_per_s"] = cells / (r["duration_us"] * 1e-6) / 1e9
            traffic[f"{cfg}:{key}"] = dram
            if cfg == "cfg2":
                traffic[key] = dram
        rows.append(r)
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    with open(os.path.join(ROOT, "profiles", "traffic_r02.json"), "w") as f:
        json.dump(traffic, f, indent=1, sort_keys=True)
    with open(os.path.join(ROOT, "profiles", "ncu_r02_raw.json"), "w") as f:
        json.dump(rows, f, indent=1)
    cols = ["capture", "pass", "duration_us", "Gpts_per_s", "dram_read_MB", "dram_write_MB", "dram_B_per_cell", "algorithmic_B_per_cell",
            "dram_TBps", "issue_active_pct", "warps_active_pct", "regs", "grid", "waves", "stall_long_sb", "stall_barrier",
            "stall_membar", "l2_hit_pct"]
    lines = ["| " + " | ".join(cols) + " |", "|" + "---|" * len(cols)]
    for r in rows:
        lines.append("| " + " | ".join((f"{r[c]:.3g}" if isinstance(r.get(c), float) else str(r.get(c, ""))) for c in cols) + " |")
    print("\n".join(lines))
    return lines


if __name__ == "__main__":
    main()
