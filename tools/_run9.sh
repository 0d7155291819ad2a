python -m pytest tests -x -q -m gpu -k "lsrtm or cfg4 or parity or mid_size" 2>&1 | tail -4
python tools/perf_kernels.py acoustic_vti_lsrtm_habc 500 1200 8 200 2>&1 | grep -v Warn
python tools/perf_kernels.py acoustic_tti_lsrtm_habc 500 1200 8 200 2>&1 | grep -v Warn
