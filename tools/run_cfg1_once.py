"""One forward modelling of BASELINE configs[0] (for ncu captures of the persistent kernel)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench, seistorch_b200 as sb
true, _ = bench.WORKLOADS["cfg1"]["models"]()
case = bench.make_case(1, workload="cfg1", models=true, nt=int(os.environ.get("NT", "2000")))
if os.environ.get("NOREC"):
    case["receivers"] = [[[], []]]
x = torch.as_tensor(case["wavelet"], device="cuda").unsqueeze(0)
cfg, model = sb.model_from_case(case, device="cuda", mode="forward")
with torch.no_grad():
    for _ in range(int(os.environ.get("REPS", "2"))):
        model(x)
torch.cuda.synchronize()
