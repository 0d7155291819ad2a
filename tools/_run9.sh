python -m pytest tests -x -q -m gpu 2>&1 | tail -3
export SEISTORCH_B200_TMA=1
python tools/perf_kernels.py acoustic_habc 751 2301 8 1000 2>&1 | grep -v Warn
