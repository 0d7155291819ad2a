"""Small gradient runs through every kernel family touched in round 2, meant to be run under compute-sanitizer:
    compute-sanitizer --tool memcheck python tools/sanitize_r02.py
(persistent forward / adjoint, elastic fast forward / adjoint incl. edge tiles, tap-gather frames of the Born and TTI
equations, TMA kernels with tap corners, 3D, gradient smoothing)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import seistorch_b200 as sb
from seistorch_b200 import engine
from seistorch_b200.process import gaussian_filter
from oracle import cases

def run(eq, **kw):
    ny = kw.pop("ny", None)
    case = cases.make_case(eq, ny=ny, **kw)
    cfg, model = sb.model_from_case(case, device="cuda", mode="inversion")
    x = torch.as_tensor(np.asarray(case["wavelet"]), device="cuda").unsqueeze(0).requires_grad_(True)
    syn = model(x)
    sum((s ** 2).sum() for s in syn).backward()
    torch.cuda.synchronize()
    print(eq, kw, engine.KERNELS, "ok", flush=True)

run("acoustic", nz=37, nx=70, nshots=2, nt=12)                       # persistent kernels, partly filled strips
os.environ["SEISTORCH_B200_PERSIST"] = "0"
run("acoustic", nz=37, nx=70, nshots=2, nt=6)                        # per-step register kernels
os.environ.pop("SEISTORCH_B200_PERSIST")
run("elastic", nz=45, nx=77, nshots=2, nt=6)                         # edge tiles on every side
run("elastic", nz=150, nx=300, nshots=2, nt=4)                       # interior + edge tiles
for eq in ("acoustic_habc", "vti_habc2", "tti_habc", "acoustic_fwim_habc", "acoustic_rho_habc", "acoustic_lsrtm_habc",
           "acoustic_vti_lsrtm_habc", "acoustic_tti_lsrtm_habc"):
    run(eq, nz=33, nx=61, nshots=3, nt=5)
# second pass of round 2: frame-free tiles (unchecked-load rows of the Born pairs), strips of the tti / Born adjoints next to
# clean tiles, multi-shot chunks with shared-memory gradient sums (wave2d and elastic)
os.environ["SEISTORCH_B200_BCHUNK"] = "2"
for eq in ("acoustic_vti_lsrtm_habc", "acoustic_tti_lsrtm_habc", "tti_habc", "acoustic_fwim_habc"):
    run(eq, nz=60, nx=300, nshots=4, nt=4)
run("elastic", nz=150, nx=300, nshots=4, nt=4)
os.environ.pop("SEISTORCH_B200_BCHUNK")
os.environ["SEISTORCH_B200_TMA"] = "1"
run("acoustic_habc", nz=150, nx=216, nshots=3, nt=5)                 # TMA kernels, tap-gather corner tiles (adjoint)
run("acoustic", nz=150, nx=300, nshots=3, nt=5)
os.environ.pop("SEISTORCH_B200_TMA")
run("acoustic", nz=9, nx=11, ny=8, nshots=2, nt=4)                   # 3D
g = torch.randn(37, 53, device="cuda")
gaussian_filter(gaussian_filter(g, 2.0, 4, 0), 2.0, 6, 1)
torch.cuda.synchronize()
print("smoothing ok")
