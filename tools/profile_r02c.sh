#!/bin/bash
# Final profiling pass of round 2 (final build): bench lines of every configuration, the ncu launch list of the default bench
# command, ncu captures of the kernels that changed since the earlier passes.  Outputs: gpurun_out/r02c/.
set -u
O=gpurun_out/r02c
mkdir -p $O
for c in cfg2 cfg1 cfg3 cfg4 cfg4_tti cfg4_fwim cfg5; do
  timeout 600 python bench.py --config $c --steps 4 --warmup 3 > $O/bench_$c.json 2> $O/bench_$c.err
  tail -c 300 $O/bench_$c.json | head -c 200; echo
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $O/launches_cfg2.csv \
    python bench.py --steps 1 --warmup 1 --nt 60 --no-cpu-baseline --no-check > $O/launches_cfg2.log 2>&1
NCU="ncu --set full --clock-control none --import-source on -f"
$NCU -k regex:wave2d_adjoint_tma -s 60 -c 1 -o $O/cfg2_adj python tools/perf_kernels.py acoustic_habc 751 2301 8 60 > $O/ncu.log 2>&1
$NCU -k regex:wave2d_forward_tma -s 200 -c 1 -o $O/cfg2_fwd python tools/perf_kernels.py acoustic_habc 751 2301 8 60 >> $O/ncu.log 2>&1
$NCU -k regex:wave2d_forward_kernel -s 40 -c 1 -o $O/cfg4_fwd python tools/perf_kernels.py acoustic_vti_lsrtm_habc 500 1200 12 30 >> $O/ncu.log 2>&1
$NCU -k regex:wave2d_adjoint_kernel -s 20 -c 1 -o $O/cfg4_adj python tools/perf_kernels.py acoustic_vti_lsrtm_habc 500 1200 12 30 >> $O/ncu.log 2>&1
$NCU -k regex:wave2d_adjoint_kernel -s 20 -c 1 -o $O/cfg4tti_adj python tools/perf_kernels.py acoustic_tti_lsrtm_habc 500 1200 12 30 >> $O/ncu.log 2>&1
tail -2 $O/ncu.log
