for k in 2 4; do
  export SEISTORCH_B200_LIB=$PWD/seistorch_b200/build/variants/bsh$k.so
  for m in 0 1; do export SEISTORCH_B200_TMA=$m
  echo "== BSH $k TMA $m"
  python tools/perf_kernels.py acoustic_habc 751 2301 8 400 2>&1 | grep -v Warn
  done
done
unset SEISTORCH_B200_LIB
export SEISTORCH_B200_TMA=1
for t in 2 8; do export SEISTORCH_B200_TSH=$t; echo "== TSH $t"; python tools/perf_kernels.py acoustic_habc 751 2301 8 400 2>&1 | grep -v Warn; done
