#!/bin/bash
# Second profiling pass of round 2 (after the tap-gather corner tiles, the 13-tap TTI frames and the 3D gradient flush):
# bench lines of every configuration + ncu captures of the kernels that changed.  Outputs: gpurun_out/r02b/.
set -u
O=gpurun_out/r02b
mkdir -p $O
for c in cfg2 cfg1 cfg3 cfg4 cfg4_tti cfg4_fwim cfg5; do
  timeout 600 python bench.py --config $c --steps 4 --warmup 3 > $O/bench_$c.json 2> $O/bench_$c.err
  tail -c 300 $O/bench_$c.json | head -c 200; echo
done
NCU="ncu --set full --clock-control none --import-source on -f"
$NCU -k regex:wave2d_adjoint_tma -s 60 -c 1 -o $O/cfg2_adj python tools/perf_kernels.py acoustic_habc 751 2301 8 60 > $O/ncu.log 2>&1
$NCU -k regex:wave2d_forward_tma -s 200 -c 1 -o $O/cfg2_fwd python tools/perf_kernels.py acoustic_habc 751 2301 8 60 >> $O/ncu.log 2>&1
$NCU -k regex:elastic2d_forward -s 20 -c 1 -o $O/cfg3_fwd python tools/perf_kernels.py elastic 400 1000 4 30 >> $O/ncu.log 2>&1
$NCU -k regex:elastic2d_adjoint_fast -s 20 -c 1 -o $O/cfg3_adj python tools/perf_kernels.py elastic 400 1000 4 30 >> $O/ncu.log 2>&1
$NCU -k regex:wave2d_adjoint_kernel -s 20 -c 1 -o $O/cfg4_adj python tools/perf_kernels.py acoustic_vti_lsrtm_habc 500 1200 12 30 >> $O/ncu.log 2>&1
$NCU -k regex:wave2d_adjoint_kernel -s 20 -c 1 -o $O/cfg4tti_adj python tools/perf_kernels.py acoustic_tti_lsrtm_habc 500 1200 12 30 >> $O/ncu.log 2>&1
$NCU -k regex:acoustic3d_kernel -s 70 -c 1 -o $O/cfg5_adj python tools/perf_kernels.py acoustic 200 400 1 20 400 >> $O/ncu.log 2>&1
tail -2 $O/ncu.log
