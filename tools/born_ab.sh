# A/B timing of tuning builds / settings of the register kernels (tools only; not part of the product path)
mkdir -p gpurun_out/r02e
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02e/pytest_gpu_final.log 2>&1; tail -3 gpurun_out/r02e/pytest_gpu_final.log
for v in "" fwd0; do
  if [ -n "$v" ]; then export SEISTORCH_B200_LIB=$PWD/seistorch_b200/libseistorch_b200_$v.so; else unset SEISTORCH_B200_LIB; fi
  echo "== variant '$v'"
  for eq in acoustic_vti_lsrtm_habc acoustic_tti_lsrtm_habc tti_habc; do
    timeout 120 python tools/perf_kernels.py $eq 500 1200 12 60
  done
done 2>&1 | tee gpurun_out/r02e/fwd_strips_ab.log
