# graph replay: tests + repeatability of the default bench line (tools only)
mkdir -p gpurun_out/r02d
timeout 600 python -m pytest tests/test_gpu_graph.py -x -q 2>&1 | tail -8
for m in 1 0; do
  SEISTORCH_B200_GRAPH=$m timeout 600 python bench.py --config cfg2 --steps 12 --warmup 3 --no-cpu-baseline > gpurun_out/r02d/bench_cfg2_g$m.json 2>gpurun_out/r02d/bench_cfg2_g$m.err
  tail -2 gpurun_out/r02d/bench_cfg2_g$m.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/r02d/bench_cfg2_g$m.json").read().strip().splitlines()[-1]); print("graph=$m", round(d["value"],2), round(d["ms_per_step"],1), round(d["e2e"]["value"],2), d["step_ms"]["device"], d["step_ms"]["e2e"], d["parity"]["rec_err"], d["parity"]["grad_err"], d.get("graph"))
PY
done
