"""ctypes binding of include/seistorch_b200.h (the C-ABI boundary).

The library is loaded lazily; if it is missing the product path fails loudly --
there is no CPU or PyTorch fallback (BASELINE north star).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SEISTORCH_B200_LIB") or os.path.join(_HERE, "libseistorch_b200.so")

c_float_p = C.POINTER(C.c_float)
c_int_p = C.POINTER(C.c_int32)


class StAcquisition(C.Structure):
    _fields_ = [
        ("ns", C.c_int32),
        ("src_b", C.c_void_p), ("src_i0", C.c_void_p), ("src_i1", C.c_void_p), ("src_i2", C.c_void_p),
        ("amp", C.c_void_p), ("gamp", C.c_void_p),
        ("src_fmask", C.c_int32),
        ("R", C.c_int32),
        ("row_start", C.c_void_p), ("rec_col", C.c_void_p), ("rec_orig", C.c_void_p),
        ("nchan", C.c_int32),
        ("chan_f", C.c_int32 * 4),
        ("rec_out", C.c_void_p), ("rec_adj", C.c_void_p),
        ("row_lo", C.c_int32), ("row_hi", C.c_int32),
    ]


class StWave2dProblem(C.Structure):
    _fields_ = [
        ("flags", C.c_int32),
        ("B", C.c_int32), ("nz", C.c_int32), ("nx", C.c_int32), ("ld", C.c_int32),
        ("bw", C.c_int32), ("multiple", C.c_int32),
        ("nt", C.c_int32),
        ("dt", C.c_float),
        ("coef", C.c_void_p * 8),
        ("taps", C.c_void_p),
        ("u", C.c_void_p),
        ("nslots", C.c_int32),
        ("lam", C.c_void_p),
        ("gacc", C.c_void_p),
        ("bchunk", C.c_int32),
        ("acq", StAcquisition),
    ]


class StElastic2dProblem(C.Structure):
    _fields_ = [
        ("B", C.c_int32), ("nz", C.c_int32), ("nx", C.c_int32), ("ld", C.c_int32), ("nt", C.c_int32),
        ("coef", C.c_void_p * 5),
        ("u", C.c_void_p),
        ("nslots", C.c_int32),
        ("lam", C.c_void_p),
        ("gacc", C.c_void_p),
        ("bchunk", C.c_int32),
        ("acq", StAcquisition),
    ]


class StAcoustic3dProblem(C.Structure):
    _fields_ = [
        ("B", C.c_int32), ("n0", C.c_int32), ("n1", C.c_int32), ("n2", C.c_int32), ("ld", C.c_int32),
        ("nt", C.c_int32),
        ("dt", C.c_float),
        ("coef", C.c_void_p * 2),
        ("u", C.c_void_p),
        ("nslots", C.c_int32),
        ("lam", C.c_void_p),
        ("gacc", C.c_void_p),
        ("bchunk", C.c_int32),
        ("acq", StAcquisition),
    ]


# every symbol include/seistorch_b200.h declares
EXPORTS = [
    "st_version", "st_last_error", "st_graph_counters", "st_graph_last_failure",
    "st_wave2d_taps_floats", "st_wave2d_prepare", "st_wave2d_uses_tma", "st_wave2d_uses_persist", "st_wave2d_adjoint_uses_persist",
    "st_wave2d_forward", "st_wave2d_adjoint",
    "st_acoustic2d_forward", "st_acoustic2d_adjoint",
    "st_acoustic2d_habc_forward", "st_acoustic2d_habc_adjoint",
    "st_qp2d_forward", "st_qp2d_adjoint", "st_fwim2d_forward", "st_fwim2d_adjoint",
    "st_elastic2d_forward", "st_elastic2d_adjoint",
    "st_acoustic3d_forward", "st_acoustic3d_adjoint",
    "st_misfit_l2", "st_misfit_l1", "st_misfit_cs", "st_misfit_nim", "st_misfit_w1d", "st_misfit_traveltime", "st_misfit_sml1", "st_misfit_cc", "st_misfit_integration", "st_filtfilt", "st_misfit_envelope", "st_misfit_envelope_workspace", "st_gaussian_smooth2d", "st_illumination",
]

_lib = None


class LibraryMissing(RuntimeError):
    pass


def lib():
    """Load libseistorch_b200.so (once).  Raises LibraryMissing with build instructions."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise LibraryMissing(
            f"{LIB_PATH} not found: the CUDA extension is not built. Run "
            "`python -m seistorch_b200.build` (needs nvcc). There is no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    L.st_version.restype = C.c_int
    L.st_last_error.restype = C.c_char_p
    L.st_graph_last_failure.restype = C.c_char_p
    L.st_graph_counters.restype = None
    L.st_graph_counters.argtypes = [C.POINTER(C.c_int64)]
    step_args = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]
    L.st_wave2d_taps_floats.restype = C.c_int64
    L.st_wave2d_taps_floats.argtypes = [C.c_void_p]
    L.st_wave2d_prepare.restype = C.c_int
    L.st_wave2d_prepare.argtypes = [C.c_void_p, C.c_void_p]
    L.st_wave2d_uses_tma.restype = C.c_int
    L.st_wave2d_uses_tma.argtypes = [C.c_void_p, C.c_int32]
    L.st_wave2d_uses_persist.restype = C.c_int
    L.st_wave2d_uses_persist.argtypes = [C.c_void_p, C.c_int32]
    L.st_wave2d_adjoint_uses_persist.restype = C.c_int
    L.st_wave2d_adjoint_uses_persist.argtypes = [C.c_void_p, C.c_int32]
    for name in EXPORTS:
        if not (name.endswith("_forward") or name.endswith("_adjoint")):
            continue
        fn = getattr(L, name)
        fn.restype = C.c_int
        fn.argtypes = step_args
    L.st_misfit_l2.restype = C.c_int
    L.st_misfit_l2.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]
    L.st_misfit_l1.restype = C.c_int
    L.st_misfit_l1.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]
    L.st_misfit_cs.restype = C.c_int
    L.st_misfit_cs.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]
    L.st_misfit_nim.restype = C.c_int
    L.st_misfit_nim.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]
    L.st_misfit_w1d.restype = C.c_int
    L.st_misfit_w1d.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]
    L.st_misfit_sml1.restype = C.c_int
    L.st_misfit_sml1.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]
    L.st_misfit_cc.restype = C.c_int
    L.st_misfit_cc.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]
    L.st_misfit_integration.restype = C.c_int
    L.st_misfit_integration.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int64, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]
    L.st_misfit_traveltime.restype = C.c_int
    L.st_misfit_traveltime.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_float,
                                       C.c_void_p, C.c_void_p, C.c_void_p]
    L.st_filtfilt.restype = C.c_int
    L.st_filtfilt.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]
    L.st_misfit_envelope.restype = C.c_int
    L.st_misfit_envelope.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_float,
                                     C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.st_gaussian_smooth2d.restype = C.c_int
    L.st_gaussian_smooth2d.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]
    L.st_illumination.restype = C.c_int
    L.st_illumination.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int64, C.c_int32, C.c_int64,
                                  C.c_void_p, C.c_void_p]
    L.st_misfit_envelope_workspace.restype = C.c_int64
    L.st_misfit_envelope_workspace.argtypes = [C.c_int32, C.c_int32]
    _lib = L
    return L


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = lib().st_last_error().decode(errors="replace")
        raise RuntimeError(f"seistorch_b200 {what} failed (code {rc}): {msg}")


def ptr(t):
    """Device pointer of a tensor (or None -> NULL)."""
    return None if t is None else t.data_ptr()


def graph_counters():
    """(plain, captured, replayed) time-loop calls since process start (include/seistorch_b200.h: st_graph_counters)."""
    out = (C.c_int64 * 3)()
    lib().st_graph_counters(out)
    return tuple(int(v) for v in out)


def graph_last_failure() -> str:
    return lib().st_graph_last_failure().decode(errors="replace")
