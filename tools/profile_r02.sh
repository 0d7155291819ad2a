#!/bin/bash
# Round-2 profiling pass (run on the GPU box through gpurun): per-config bench lines, the ncu launch list of the default
# bench command, and one `ncu --set full` capture of the dominant kernels of every configuration.
# Outputs go to gpurun_out/r02/; tools/summarize_ncu.py turns them into profiles/ncu_r02_summary.md + traffic_r02.json.
set -u
O=gpurun_out/r02
mkdir -p $O
for c in cfg2 cfg1 cfg3 cfg4 cfg4_tti cfg4_fwim cfg5; do
  timeout 600 python bench.py --config $c --steps 3 --warmup 3 > $O/bench_$c.json 2> $O/bench_$c.err
  tail -c 400 $O/bench_$c.json | head -c 300; echo
done
# launch list of the default bench command (short horizon: every kernel of one step shows up; times are cold-cache)
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $O/launches_cfg2.csv \
    python bench.py --steps 1 --warmup 1 --nt 60 --no-cpu-baseline --no-check > $O/launches_cfg2.log 2>&1
NCU="ncu --set full --clock-control none --import-source on -f"
$NCU -k regex:wave2d_adjoint_tma -s 60 -c 1 -o $O/cfg2_adj python tools/perf_kernels.py acoustic_habc 751 2301 8 60 > $O/ncu.log 2>&1
$NCU -k regex:wave2d_forward_tma -s 200 -c 1 -o $O/cfg2_fwd python tools/perf_kernels.py acoustic_habc 751 2301 8 60 >> $O/ncu.log 2>&1
NT=400 $NCU -k regex:persist -c 1 -o $O/cfg1_persist python tools/run_cfg1_once.py >> $O/ncu.log 2>&1
$NCU -k regex:elastic2d_forward -s 20 -c 1 -o $O/cfg3_fwd python tools/perf_kernels.py elastic 400 1000 4 30 >> $O/ncu.log 2>&1
$NCU -k regex:elastic2d_adjoint_fast -s 20 -c 1 -o $O/cfg3_adj python tools/perf_kernels.py elastic 400 1000 4 30 >> $O/ncu.log 2>&1
$NCU -k regex:wave2d_forward_kernel -s 40 -c 1 -o $O/cfg4_fwd python tools/perf_kernels.py acoustic_vti_lsrtm_habc 500 1200 12 30 >> $O/ncu.log 2>&1
$NCU -k regex:wave2d_adjoint_kernel -s 20 -c 1 -o $O/cfg4_adj python tools/perf_kernels.py acoustic_vti_lsrtm_habc 500 1200 12 30 >> $O/ncu.log 2>&1
$NCU -k regex:acoustic3d_kernel -s 20 -c 1 -o $O/cfg5_fwd python tools/perf_kernels.py acoustic 200 400 1 20 400 >> $O/ncu.log 2>&1
$NCU -k regex:acoustic3d_kernel -s 70 -c 1 -o $O/cfg5_adj python tools/perf_kernels.py acoustic 200 400 1 20 400 >> $O/ncu.log 2>&1
tail -3 $O/ncu.log
ls -la $O
