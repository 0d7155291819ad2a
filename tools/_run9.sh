python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python tools/perf_kernels.py acoustic_vti_lsrtm_habc 500 1200 8 200 2>&1 | grep -v Warn
python tools/perf_kernels.py acoustic_tti_lsrtm_habc 500 1200 8 200 2>&1 | grep -v Warn
python tools/perf_kernels.py tti_habc 500 1200 8 200 2>&1 | grep -v Warn
python tools/perf_kernels.py acoustic_habc 751 2301 8 1000 2>&1 | grep -v Warn
