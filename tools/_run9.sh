export SEISTORCH_B200_TMA=1
export SEISTORCH_B200_LIB=$PWD/seistorch_b200/build/variants/occ3.so
python -m pytest tests/test_gpu_more.py -x -q -m gpu -k "tma" 2>&1 | tail -2
python tools/perf_kernels.py acoustic_habc 751 2301 8 1000 2>&1 | grep -v Warn
