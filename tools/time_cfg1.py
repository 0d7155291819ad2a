"""Times BASELINE configs[0] (forward modelling, 2000 steps, 1 shot) with the persistent kernel and per step."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench, seistorch_b200 as sb
from seistorch_b200 import engine
true, _ = bench.WORKLOADS["cfg1"]["models"]()
for nshots in (1, 9):
    case = bench.make_case(1, workload="cfg1", models=true)
    case["sources"] = case["sources"] * nshots
    case["receivers"] = case["receivers"] * nshots
    x = torch.as_tensor(case["wavelet"], device="cuda").unsqueeze(0)
    for env in ({"SEISTORCH_B200_PERSIST": "0"}, {"SEISTORCH_B200_PERSIST": "1", "SEISTORCH_B200_PERSIST_VARIANT": "0"},
                {"SEISTORCH_B200_PERSIST": "1", "SEISTORCH_B200_PERSIST_VARIANT": "1"}):
        os.environ.update(env)
        cfg, model = sb.model_from_case(case, device="cuda", mode="forward")
        with torch.no_grad():
            for _ in range(3):
                model(x)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                model(x)
            e1.record()
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print(f"shots={nshots} {env} {engine.KERNELS['forward']}: {ms:.3f} ms per 2000 steps = {ms / 2000 * 1e3:.3f} us/step, "
              f"{nshots * 250 * 400 * 2000 / ms / 1e6:.1f} Gpts/s", flush=True)
