# A/B timing of tuning builds of the register adjoint (tools only; not part of the product path)
mkdir -p gpurun_out/r02d
for v in "" nif3 nif4; do
  if [ -n "$v" ]; then export SEISTORCH_B200_LIB=$PWD/seistorch_b200/libseistorch_b200_$v.so; else unset SEISTORCH_B200_LIB; fi
  echo "== variant '$v'"
  for eq in acoustic_vti_lsrtm_habc acoustic_tti_lsrtm_habc tti_habc vti_habc2 acoustic_fwim_habc; do
    timeout 300 python tools/perf_kernels.py $eq 500 1200 12 100
  done
done 2>&1 | tee gpurun_out/r02d/born_ab9.log
