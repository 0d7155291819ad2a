#!/bin/bash
# Profiling pass after the register-adjoint rework (fused Born rows, shared-memory gradient sums, adjoint strips for every
# flag set, elastic shot chunks): full GPU test suite, bench lines of every configuration, ncu captures of the kernels that
# changed.  Outputs: gpurun_out/r02d/.
set -u
O=gpurun_out/r02d
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; tail -2 $O/pytest_gpu.log
for c in cfg2 cfg1 cfg3 cfg4 cfg4_tti cfg4_fwim cfg5; do
  timeout 600 python bench.py --config $c --steps 4 --warmup 3 > $O/bench_$c.json 2> $O/bench_$c.err
  tail -c 300 $O/bench_$c.json | head -c 200; echo
done
NCU="ncu --set full --clock-control none --import-source on -f"
$NCU -k regex:elastic2d_adjoint_fast -s 20 -c 1 -o $O/cfg3_adj python tools/perf_kernels.py elastic 400 1000 4 30 > $O/ncu.log 2>&1
$NCU -k regex:wave2d_adjoint_kernel -s 20 -c 1 -o $O/cfg4_adj python tools/perf_kernels.py acoustic_vti_lsrtm_habc 500 1200 12 30 >> $O/ncu.log 2>&1
$NCU -k regex:wave2d_adjoint_kernel -s 20 -c 1 -o $O/cfg4tti_adj python tools/perf_kernels.py acoustic_tti_lsrtm_habc 500 1200 12 30 >> $O/ncu.log 2>&1
tail -2 $O/ncu.log
