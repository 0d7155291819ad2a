# A/B timing of tuning builds / settings of the adjoint kernels (tools only; not part of the product path)
mkdir -p gpurun_out/r02d
timeout 900 python -m pytest tests -m gpu -x -q -k "elastic or golden or cfg3" 2>&1 | tail -3
for c in 1 2 4; do
  echo "== elastic bchunk $c"
  SEISTORCH_B200_BCHUNK=$c timeout 300 python tools/perf_kernels.py elastic 400 1000 4 200
  SEISTORCH_B200_BCHUNK=$c timeout 300 python tools/perf_kernels.py elastic 400 1000 8 200
done 2>&1 | tee gpurun_out/r02d/el_ab1.log
