"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Import shim for the *real* reference (GeophyAI/seistorch, mounted read-only at
/root/reference in the authoring container).  It is used by
``oracle/make_golden.py`` to generate the committed fixtures under
``tests/golden/``, by ``bench.py --impl reference`` (the CPU arm) and by the overlay tests.
The reference is looked for at ``$SEISTORCH_REFERENCE``, then ``/root/reference`` (authoring
container), then ``oracle/_ref`` -- the byte-for-byte copy that ``oracle/make_ref.py`` ships to
the GPU box (git-ignored).

What the shim does:
  1. stubs third-party modules the reference imports at module scope but never
     touches on the wave-propagation path (prettytable, segyio, obspy, h5py,
     geomloss, ot, matplotlib, mpi4py);
  2. redirects ``tensor.to("cuda")`` / ``tensor.to(Exception(...))`` to a no-op
     on GPU-less hosts so the equation modules (which hard-code device="cuda"
     at import, equations2d/acoustic.py:56) can be imported;
  3. sets ``seistorch.compile.force_compile = False`` so the eager functions
     are the oracle.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

def _find_reference():
    env = os.environ.get("SEISTORCH_REFERENCE")
    cands = [env] if env else []
    cands += ["/root/reference", os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")]
    for c in cands:
        if c and os.path.isdir(os.path.join(c, "seistorch")):
            return c
    return cands[0]


REFERENCE_ROOT = _find_reference()

_STUBS = [
    "prettytable", "segyio", "obspy", "h5py", "geomloss", "ot", "ot.utils",
    "matplotlib", "matplotlib.pyplot", "matplotlib.colors", "matplotlib.animation",
    "mpi4py", "mpi4py.util", "mpi4py.util.pkl5", "mpi4py.MPI",
]


class _Anything:
    """Attribute sink: any attribute / call returns another sink."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Anything()

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return _Anything()


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "seistorch"))


def _install_stubs():
    for name in _STUBS:
        if name in sys.modules:
            continue
        try:
            importlib.import_module(name)
            continue
        except Exception:
            pass
        mod = types.ModuleType(name)
        def _getattr(attr, _n=name):
            if attr.startswith("__") and attr.endswith("__"):
                raise AttributeError(attr)
            return _Anything()

        mod.__getattr__ = _getattr  # type: ignore
        mod.__path__ = []  # behave like a package
        sys.modules[name] = mod


_patched = False


def _patch_tensor_to():
    """Make ``.to('cuda')`` and ``.to(Exception)`` a no-op on CPU-only hosts."""
    global _patched
    import torch

    if _patched or torch.cuda.is_available():
        return
    orig_to = torch.Tensor.to

    def to(self, *args, **kwargs):
        if args and (isinstance(args[0], Exception) or
                     (isinstance(args[0], str) and args[0].startswith("cuda"))):
            args = args[1:]
            if not args and not kwargs:
                return self
        if isinstance(kwargs.get("device", None), str) and kwargs["device"].startswith("cuda"):
            kwargs = dict(kwargs)
            kwargs.pop("device")
        return orig_to(self, *args, **kwargs)

    torch.Tensor.to = to  # type: ignore
    _patched = True


def import_reference():
    """Return the reference ``seistorch`` package (imported from REFERENCE_ROOT)."""
    if not reference_available():
        raise ImportError(f"reference tree not found at {REFERENCE_ROOT}")
    _install_stubs()
    _patch_tensor_to()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import seistorch  # noqa: F401  (the reference package)
    import seistorch.compile as sc

    sc.force_compile = False
    return seistorch


def cast_module_kernels(dtype, device=None):
    """Reference quirk: module-level conv kernels are created in
    fp32 on "cuda" at import; cast them so the reference's own ``dtype: float64`` path runs,
    and (``device``) re-home them so its CPU path also runs on a host that has a GPU
    (equations3d/acoustic.py:67 never moves its kernel)."""
    import torch

    names = [
        "seistorch.equations2d.acoustic", "seistorch.equations2d.acoustic_habc",
        "seistorch.equations2d.convkernel", "seistorch.equations2d.vti_habc2",
        "seistorch.equations2d.tti_habc", "seistorch.equations2d.acoustic_vti_lsrtm_habc", "seistorch.equations2d.acoustic_lsrtm_habc", "seistorch.equations2d.acoustic_rho_habc",
        "seistorch.equations2d.acoustic_tti_lsrtm_habc", "seistorch.equations2d.acoustic_fwim_habc",
        "seistorch.equations3d.acoustic",
    ]
    for n in names:
        try:
            m = importlib.import_module(n)
        except Exception:
            continue
        for k, v in list(vars(m).items()):
            if k.startswith("kernel") and isinstance(v, torch.Tensor):
                setattr(m, k, v.to(dtype) if device is None else v.to(device=device, dtype=dtype))
