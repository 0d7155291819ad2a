"""CPU: the host side of bench.py -- every BASELINE workload builds its synthetic case (SURVEY 8d inputs), the JSON
bookkeeping helpers agree with each other, and the CPU arm runs the reference's own code (oracle/_ref or /root/reference)
on a tiny sample."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT


def test_every_workload_builds_a_case():
    import bench
    expect = {"cfg1": ((150, 300), 1), "cfg2": ((751, 2301), 16), "cfg3": ((400, 1000), 8), "cfg4": ((500, 1200), 12),
              "cfg4_tti": ((500, 1200), 12), "cfg4_fwim": ((500, 1200), 12), "cfg5": ((400, 200, 400), 4)}
    assert sorted(bench.WORKLOADS) == sorted(expect)
    for name, (shape, shots) in expect.items():
        if name in ("cfg2", "cfg5"):
            continue                                     # the two large model builders are exercised on the GPU box
        w = bench.WORKLOADS[name]
        true, init = w["models"]()
        assert next(iter(true.values())).shape == shape and w["shots"] == shots
        case = bench.make_case(2 if name != "cfg1" else 1, total_shots=shots, workload=name, models=init, nt=8, delay=2)
        assert case["equation"] == w["equation"] and case["boundary"] == w["boundary"]
        assert set(case["invlist"]) == set(w["inv"]) and len(case["wavelet"]) == 8
        assert all(len(s) == 2 for s in case["sources"]) and len(case["receivers"][0][0]) == (shape[1] + 1) // 2
        assert bench.metric_name(name) == ("forward_shots_per_s" if name == "cfg1" else "fwi_gradient_shots_per_s")
        cfg = bench.workload_config(2, name)
        assert cfg["grid_padded"] == [s + 100 for s in shape] and cfg["shots_total"] == 2 * shots and cfg["nt"] == w["nt"]
    # 3D acquisition pattern on a small clone: 8 x 4 source grid when all 32 shots exist, a line otherwise
    _t, init = bench._models_cfg5((44, 30, 40))
    c32 = bench.make_case(32, total_shots=32, workload="cfg5", models=init, nt=4, delay=1)
    assert len({(s[0], s[1]) for s in c32["sources"]}) == 32 and all(s[2] == 1.0 for s in c32["sources"])
    c4 = bench.make_case(4, total_shots=4, workload="cfg5", models=init, nt=4, delay=1)
    assert len({s[1] for s in c4["sources"]}) == 1


def test_reference_arm_prints_the_contract_line():
    from oracle import ref_shim
    if not ref_shim.reference_available():
        pytest.skip("reference tree not present")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "cfg1", "--steps", "1",
                        "--warmup", "1", "--cpu-steps", "40"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert line["impl"] == "reference" and line["metric"] == "forward_shots_per_s" and line["higher_is_better"] is True
    assert line["cpu_baseline"]["kind"] == "reference" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "shots/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["config"] == json.loads(json.dumps(__import__("bench").workload_config(1, "cfg1")))


def test_adjoint_chunk_rule(monkeypatch):
    """engine._choose_bchunk: shots one adjoint block walks with its gradient sums on chip -- the longest divisor of B that
    keeps enough blocks in flight (wave2d: >= 600 fast blocks; elastic2d: >= 280 tiles of 128 x 32, one wave of 2 blocks / SM)."""
    from types import SimpleNamespace as NS
    from seistorch_b200 import engine
    monkeypatch.delenv("SEISTORCH_B200_BCHUNK", raising=False)
    assert engine._choose_bchunk(NS(family="wave2d", B=12, shape=(600, 1300))) == 4        # cfg4: 209 tile pairs x 3 chunks
    assert engine._choose_bchunk(NS(family="wave2d", B=8, shape=(851, 2401))) == 4         # cfg2 grid on the register path
    assert engine._choose_bchunk(NS(family="wave2d", B=1, shape=(250, 400))) == 1
    assert engine._choose_bchunk(NS(family="elastic2d", B=4, shape=(500, 1100))) == 2      # cfg3: 144 tiles x 2 chunks
    assert engine._choose_bchunk(NS(family="elastic2d", B=8, shape=(500, 1100))) == 4
    assert engine._choose_bchunk(NS(family="elastic2d", B=2, shape=(90, 150))) == 1
    monkeypatch.setenv("SEISTORCH_B200_BCHUNK", "3")
    assert engine._choose_bchunk(NS(family="elastic2d", B=2, shape=(90, 150))) == 2        # clamped to B
