"""Overlay the accelerated modules onto an importable reference ``seistorch`` package.

    import seistorch_b200.overlay as ov; ov.install()      # before `import seistorch`
    from seistorch.model import build_model                # reference builder, our hot path

After ``install()`` the reference's own ``build_model`` (seistorch/model.py:23-93) finds
our modules under the names it imports -- ``seistorch.equations2d.<eq>``,
``seistorch.equations3d.acoustic``, ``seistorch.cell``, ``seistorch.rnn``,
``seistorch.source``, ``seistorch.probe``, ``seistorch.checkpoint[_new]`` -- so the drivers
(``seistorch_dist.py``, ``fwi.py --mode forward``, ``codingfwi.py``) run unchanged on the
sm_100a kernels.  ``seistorch.compile.force_compile`` is switched off so the step ops are
not wrapped by torch.compile.  See INTEGRATION.md.
"""
from __future__ import annotations

import importlib
import sys

_MODULES = ["cell", "rnn", "source", "probe", "checkpoint", "checkpoint_new", "type", "habc", "pml"]
_EQ2D = ["acoustic", "acoustic_habc", "vti_habc2", "tti_habc", "acoustic_fwim_habc",
         "acoustic_lsrtm_habc", "acoustic_rho_habc", "acoustic_vti_lsrtm_habc", "acoustic_tti_lsrtm_habc", "elastic"]


def install(reference_package: str = "seistorch"):
    """Register our modules under the reference package's names (idempotent)."""
    pairs = [(f"{reference_package}.{m}", f"seistorch_b200.{m}") for m in _MODULES]
    pairs += [(f"{reference_package}.equations2d.{e}", f"seistorch_b200.equations2d.{e}") for e in _EQ2D]
    pairs += [(f"{reference_package}.equations3d.acoustic", "seistorch_b200.equations3d.acoustic")]
    for ref_name, ours in pairs:
        sys.modules[ref_name] = importlib.import_module(ours)
    try:
        comp = importlib.import_module(f"{reference_package}.compile")
        comp.force_compile = False
    except Exception:
        pass
    return [p[0] for p in pairs]
