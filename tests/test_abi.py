"""CPU: the C-ABI library loads (no GPU needed) and exports every symbol that
include/seistorch_b200.h declares; struct layouts in the ctypes binding match the header's
field order; the product path has no CPU fallback."""
import ctypes as C
import os
import re

import pytest
import torch

from conftest import ROOT


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "seistorch_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(st_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from seistorch_b200 import _lib
    lib = _lib.lib()
    declared = _declared_symbols()
    assert len(declared) >= 19
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert sorted(_lib.EXPORTS) == declared
    assert lib.st_version() == 1
    assert lib.st_misfit_envelope_workspace(2000, 10) == 3 * 2000 * 10


def test_struct_field_order_matches_header():
    from seistorch_b200 import _lib
    text = open(os.path.join(ROOT, "include", "seistorch_b200.h")).read()
    for cname, cls in [("st_acquisition", _lib.StAcquisition), ("st_wave2d_problem", _lib.StWave2dProblem),
                       ("st_elastic2d_problem", _lib.StElastic2dProblem), ("st_acoustic3d_problem", _lib.StAcoustic3dProblem)]:
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (cname, cname), text, flags=re.S).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        names = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            for part in decl.split(","):
                m = re.search(r"([A-Za-z_][A-Za-z0-9_]*)\s*(\[\d+\])?$", part.strip())
                names.append(m.group(1))
        assert names == [f[0] for f in cls._fields_], cname


def test_bad_arguments_return_error_codes_without_touching_the_gpu():
    from seistorch_b200 import _lib
    lib = _lib.lib()
    p = _lib.StWave2dProblem()           # all zero: bad shape
    rc = lib.st_wave2d_forward(C.byref(p), 0, 1, 0, None)
    assert rc == -1 and b"bad shape" in lib.st_last_error()
    rc = lib.st_acoustic2d_forward(C.byref(p), 0, 1, 0, None)
    assert rc == -1 and b"flags" in lib.st_last_error()
    e = _lib.StElastic2dProblem()
    assert lib.st_elastic2d_forward(C.byref(e), 0, 1, 0, None) == -1
    assert lib.st_misfit_l2(None, None, 10, 1.0, None, None, None) == -1


def test_no_cpu_fallback():
    import numpy as np
    import seistorch_b200 as sb
    from oracle import cases
    case = cases.make_case("acoustic", nz=10, nx=12, nshots=1, nt=4)
    cfg, model = sb.model_from_case(case, device="cpu", mode="forward")
    with pytest.raises(RuntimeError, match="CUDA"):
        model(torch.zeros(1, 4))
    with pytest.raises(RuntimeError, match="CUDA"):
        sb.L2()(torch.zeros(1, 4, 2, 1), torch.zeros(1, 4, 2, 1))
