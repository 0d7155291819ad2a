"""Times an FWI gradient on the BASELINE configs[0] grid (1 shot, nt 2000): forward + backward, persistent vs per step."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench, seistorch_b200 as sb
from seistorch_b200 import engine
true, _ = bench.WORKLOADS["cfg1"]["models"]()
case = bench.make_case(1, workload="cfg1", models={"vp": (true["vp"] * 0.97).astype(np.float32)})
case["invlist"] = {"vp": True}
x = torch.as_tensor(case["wavelet"], device="cuda").unsqueeze(0)
for env in ({"SEISTORCH_B200_PERSIST": "0"}, {"SEISTORCH_B200_PERSIST": "1", "SEISTORCH_B200_PERSIST_ADJ_VARIANT": "0"},
            {"SEISTORCH_B200_PERSIST": "1", "SEISTORCH_B200_PERSIST_ADJ_VARIANT": "2"}):
    os.environ.update(env)
    cfg, model = sb.model_from_case(case, device="cuda", mode="inversion")
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    tf, tb = 1e9, 1e9
    for rep in range(4):
        model.cell.geom.vp.grad = None
        torch.cuda.synchronize(); ev[0].record()
        syn = model(x)
        loss = (syn[0] ** 2).sum()
        ev[1].record()
        loss.backward()
        ev[2].record(); torch.cuda.synchronize()
        if rep: tf, tb = min(tf, ev[0].elapsed_time(ev[1])), min(tb, ev[1].elapsed_time(ev[2]))
    print(f"{env} [{engine.KERNELS['forward']}, {engine.KERNELS['adjoint']}]: forward {tf:.2f} ms, backward {tb:.2f} ms "
          f"({tb / 2000 * 1e3:.2f} us/step), gradient {1e3 / (tf + tb):.1f} shots/s", flush=True)
