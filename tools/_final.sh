python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py 2>&1 | tail -1 > gpurun_out/bench_r01c.log; cut -c1-200 gpurun_out/bench_r01c.log
python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 | cut -c1-300
