"""Break one bench step (8 shots of the cfg2 workload) into host-visible phases."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import seistorch_b200 as sb
from seistorch_b200 import engine

dev = torch.device("cuda", 0)
true, init = bench.make_models()
nshots = int(sys.argv[1]) if len(sys.argv) > 1 else 8
nt = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
case = bench.make_case(nshots, vp=init, nt=nt)
cfg, model = sb.model_from_case(case, device=dev, mode="inversion")
wav = torch.as_tensor(case["wavelet"], device=dev).unsqueeze(0)
crit = sb.Loss("l2").loss(cfg)
obs = None
def sync():
    torch.cuda.synchronize(); return time.perf_counter()
for it in range(3):
    t0 = sync()
    syn = model(wav)
    t1 = sync()
    st = torch.stack(list(syn), 0)
    if obs is None: obs = (st.detach() * 0.9).clone()
    loss = crit(st, obs)
    t2 = sync()
    loss.backward()
    t3 = sync()
    print(f"iter {it}: forward {1e3*(t1-t0):.1f} ms ({1e6*(t1-t0)/nt:.1f} us/step)  loss {1e3*(t2-t1):.1f} ms  backward {1e3*(t3-t2):.1f} ms ({1e6*(t3-t2)/nt:.1f} us/step)  launches {engine.LAUNCHES}", flush=True)
# forward-only timing with explicit sub-phases inside the autograd function
import torch.autograd.profiler as prof
