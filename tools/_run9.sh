python -m pytest tests/test_gpu_more.py -x -q -m gpu -k "tma or baseline_size_properties" 2>&1 | tail -3
export SEISTORCH_B200_TMA=1
python tools/perf_kernels.py acoustic_habc 751 2301 8 1000 2>&1 | grep -v Warn
python tools/perf_kernels.py acoustic 751 2301 8 1000 2>&1 | grep -v Warn
