import os, sys
sys.path.insert(0, '/root/repo')
os.environ["SEISTORCH_B200_PERSIST_DEBUG"] = "1"
import numpy as np, torch
import seistorch_b200 as sb
from seistorch_b200 import engine
from oracle import cases
for nz, nx, ns in ((30, 44, 2), (100, 180, 2), (156, 412, 1), (30, 44, 1)):
    case = cases.make_case("acoustic", nz=nz, nx=nx, nshots=ns, nt=20)
    cfg, model = sb.model_from_case(case, device="cuda", mode="forward")
    with torch.no_grad():
        model(torch.as_tensor(np.asarray(case["wavelet"]), device="cuda").unsqueeze(0))
    print(nz, nx, ns, engine.KERNELS["forward"], flush=True)
