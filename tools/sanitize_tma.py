"""compute-sanitizer driver for the TMA kernels (memcheck / racecheck / synccheck: 0 errors on B200, round 1):
    compute-sanitizer --tool memcheck python tools/sanitize_tma.py
Small grid with room for every tile kind (frame-free, four straight sides, masked corner rows, generic corners)."""
import sys, os
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np, torch
import seistorch_b200 as sb
from oracle import cases
os.environ["SEISTORCH_B200_TMA"] = "1"
for eq, mult in (("acoustic_habc", False), ("acoustic_habc", True), ("acoustic", False)):
    case = cases.make_case(eq, nz=150, nx=216, nshots=3, nt=12, rec_step=9, multiple=mult)
    cfg, model = sb.model_from_case(case, device="cuda", mode="inversion")
    x = torch.as_tensor(np.asarray(case["wavelet"]), dtype=torch.float32, device="cuda").unsqueeze(0)
    syn = model(x)
    loss = sum((s ** 2).sum() for s in syn)
    loss.backward()
    torch.cuda.synchronize()
    from seistorch_b200 import engine
    print(eq, mult, engine.KERNELS, float(loss))
