"""Persistent kernels (cluster barrier, distributed shared memory, receiver staging) for compute-sanitizer racecheck / synccheck."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import seistorch_b200 as sb
from seistorch_b200 import engine
from oracle import cases
case = cases.make_case("acoustic", nz=37, nx=70, nshots=2, nt=10)
cfg, model = sb.model_from_case(case, device="cuda", mode="inversion")
x = torch.as_tensor(np.asarray(case["wavelet"]), device="cuda").unsqueeze(0).requires_grad_(True)
syn = model(x)
sum((s ** 2).sum() for s in syn).backward()
torch.cuda.synchronize()
print(engine.KERNELS, "ok")
