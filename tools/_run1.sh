for k in 0 1 2 4 6 7; do
  if [ $k = 0 ]; then unset SEISTORCH_B200_LIB; else export SEISTORCH_B200_LIB=$PWD/seistorch_b200/build/variants/skip$k.so; fi
  echo "== skip $k"; python tools/perf_kernels.py acoustic_habc 751 2301 8 400 2>&1 | grep -v Warn
done
unset SEISTORCH_B200_LIB
python tools/perf_kernels.py acoustic 751 2301 8 400 2>&1 | grep -v Warn
