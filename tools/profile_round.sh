#!/bin/bash
# What was run on the B200 (under gpurun) for the round-1 evidence in profiles/:
#   smoke, the ncu launch list of the bench command, one ncu --set full capture of the two hot kernels.
# Outputs land in gpurun_out/; the summaries copied into profiles/ are described in profiles/ncu_r01_summary.md.
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 400 --csv --log-file gpurun_out/launches_r01c.csv python bench.py --steps 1 --warmup 1 --nt 120 --no-cpu-baseline > gpurun_out/b_ncu3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:wave2d -s 300 -c 4 -f -o gpurun_out/prof_r01c_wave2d python tools/perf_step.py 8 100 > gpurun_out/ncu_full3.log 2>&1
ls -la gpurun_out/prof_r01c_wave2d.ncu-rep gpurun_out/launches_r01c.csv
