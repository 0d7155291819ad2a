"""One FWI gradient on the BASELINE configs[0] grid (for ncu captures of the persistent adjoint kernel)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench, seistorch_b200 as sb
true, _ = bench.WORKLOADS["cfg1"]["models"]()
case = bench.make_case(1, workload="cfg1", models={"vp": (true["vp"] * 0.97).astype(np.float32)}, nt=int(os.environ.get("NT", "400")))
case["invlist"] = {"vp": True}
x = torch.as_tensor(case["wavelet"], device="cuda").unsqueeze(0)
cfg, model = sb.model_from_case(case, device="cuda", mode="inversion")
for _ in range(2):
    model.cell.geom.vp.grad = None
    (model(x)[0] ** 2).sum().backward()
torch.cuda.synchronize()
