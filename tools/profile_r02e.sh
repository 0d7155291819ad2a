#!/bin/bash
# Final pass of round 2 (final build: graph replay, history pool, chunked elastic adjoint, reworked register adjoint):
# full GPU test suite, bench lines of every configuration, launch list of the default bench command.  Outputs: gpurun_out/r02e/.
set -u
O=gpurun_out/r02e
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; tail -2 $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
for c in cfg2 cfg1 cfg3 cfg4 cfg4_tti cfg4_fwim cfg5; do
  timeout 600 python bench.py --config $c --steps 4 --warmup 3 > $O/bench_$c.json 2> $O/bench_$c.err
  python - <<PY
import json
try:
    d = json.loads(open("$O/bench_$c.json").read().strip().splitlines()[-1])
    print("$c", round(d["value"], 3), round(d["ms_per_step"], 1), round(d["e2e"]["value"], 3), d.get("step_ms", {}).get("device"), d.get("graph"))
except Exception as e:
    print("$c FAILED", e)
PY
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $O/launches_cfg2.csv \
    python bench.py --steps 1 --warmup 1 --nt 60 --no-cpu-baseline --no-check > $O/launches_cfg2.log 2>&1
tail -1 $O/launches_cfg2.log | head -c 300
