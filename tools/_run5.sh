export SEISTORCH_B200_TMA=1
python tools/perf_kernels.py acoustic_habc 751 2301 8 400 2>&1 | grep -v Warn
python tools/perf_kernels.py acoustic 751 2301 8 400 2>&1 | grep -v Warn
