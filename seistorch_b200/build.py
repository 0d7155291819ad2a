"""Build the sm_100a shared library in-tree (nvcc cross-compiles without a GPU).

    python -m seistorch_b200.build [--force]

Produces seistorch_b200/libseistorch_b200.so (git-ignored; ships to the GPU box with
the gpurun snapshot).  The 2D second-order family is instantiated once per flag set in
its own translation unit so the eight variants compile in parallel.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# tuning builds: SEISTORCH_B200_VARIANT=<tag> (+ SEISTORCH_B200_NVCC_EXTRA="-D...") writes libseistorch_b200_<tag>.so next to
# the product library; select it at run time with SEISTORCH_B200_LIB=<path>
_TAG = os.environ.get("SEISTORCH_B200_VARIANT", "")
OBJ = os.path.join(HERE, "build" + ("_" + _TAG if _TAG else ""))
LIB = os.path.join(HERE, "libseistorch_b200" + ("_" + _TAG if _TAG else "") + ".so")

_EXTRA = os.environ.get("SEISTORCH_B200_NVCC_EXTRA", "").split()
if any("ST_DBG" in f for f in _EXTRA) and not _TAG:
    # -DST_DBG_SKIP / -DST_DBG_TIMELINE builds skip work or add timers: never into the product library
    raise RuntimeError("seistorch_b200.build: ST_DBG_* defines are only allowed in a tagged tuning build (SEISTORCH_B200_VARIANT=...)")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-diag-suppress", "177"] + _EXTRA

W2_FLAG_SETS = [3, 5, 4, 12, 20, 21, 36, 44]


def _nvcc():
    return shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def _units():
    units = [("st_api", "st_api.cu", []), ("st_elastic2d", "st_elastic2d.cu", []),
             ("st_acoustic3d", "st_acoustic3d.cu", []), ("st_misfit", "st_misfit.cu", []),
             ("st_wave2d_band", "st_wave2d_band.cu", []), ("st_tma", "st_tma.cu", []),
             ("st_wave2d_persist", "st_wave2d_persist.cu", []), ("st_postproc", "st_postproc.cu", []),
             ("st_wave2d_dispatch", "st_wave2d.cu", ["-DST_W2_DISPATCH_ONLY"])]
    for fl in W2_FLAG_SETS:
        units.append((f"st_wave2d_{fl}", "st_wave2d.cu", [f"-DST_W2_INSTANCE={fl}"]))
    return [u for u in units if os.path.exists(os.path.join(CSRC, u[1]))]


def _digest():
    h = hashlib.sha256()
    for name in sorted(os.listdir(CSRC)):
        path = os.path.join(CSRC, name)
        if os.path.isfile(path):
            with open(path, "rb") as f:
                h.update(name.encode())
                h.update(f.read())
    with open(os.path.join(os.path.dirname(HERE), "include", "seistorch_b200.h"), "rb") as f:
        h.update(f.read())
    with open(os.path.abspath(__file__), "rb") as f:
        h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = True) -> str:
    os.makedirs(OBJ, exist_ok=True)
    stamp = os.path.join(OBJ, "stamp")
    dig = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == dig:
        return LIB

    def compile_one(unit):
        name, src, defs = unit
        obj = os.path.join(OBJ, name + ".o")
        cmd = [_nvcc()] + NVCC_FLAGS + defs + ["-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {name}:\n{r.stdout}\n{r.stderr}")
        return obj

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 4)) as ex:
        objs = list(ex.map(compile_one, _units()))
    cmd = [_nvcc(), "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    import ctypes
    ctypes.CDLL(LIB)        # fails here (not on the GPU box) if a symbol is unresolved
    with open(stamp, "w") as f:
        f.write(dig)
    if verbose:
        print(f"[seistorch_b200] built {LIB}")
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv)
