export SEISTORCH_B200_TMA=1
for k in 7 8 14; do
  export SEISTORCH_B200_LIB=$PWD/seistorch_b200/build/variants/skip$k.so
  echo "== skip $k (7 = TMA blocks only, 8 = everything but TMA blocks, 14 = old fast blocks only)"
  python tools/perf_kernels.py acoustic_habc 751 2301 8 400 2>&1 | grep -v Warn
done
