"""Equation registry -- mirrors seistorch/eqconfigure.py (Parameters :1-43,
Wavefield :45-129) for the equations on the accelerated path."""
from __future__ import annotations


class Parameters:
    """Which model parameters an equation needs (eqconfigure.py:5-30)."""

    @staticmethod
    def valid_model_paras():
        return {
            "acoustic": ["vp"],
            "acoustic_habc": ["vp"],
            "acoustic_fwim_habc": ["vp", "rx", "rz"],
            "acoustic_lsrtm_habc": ["vp", "m"],
            "acoustic_rho_habc": ["vp", "rho"],
            "acoustic_vti_lsrtm_habc": ["vp", "epsilon", "delta", "m"],
            "acoustic_tti_lsrtm_habc": ["vp", "epsilon", "delta", "theta", "m"],
            "elastic": ["vp", "vs", "rho"],
            "vti_habc2": ["vp", "epsilon", "delta"],
            "tti_habc": ["vp", "epsilon", "delta", "theta"],
        }

    @staticmethod
    def secondorder_equations():
        """eqconfigure.py:32-43 (restricted to the supported set)."""
        return ["acoustic", "acoustic_habc", "vti_habc2", "acoustic_fwim_habc", "acoustic_lsrtm_habc", "acoustic_rho_habc",
                "acoustic_vti_lsrtm_habc", "acoustic_tti_lsrtm_habc", "tti_habc"]


class Wavefield:
    """Wavefield names per equation, in the argument order of ``_time_step``
    (eqconfigure.py:45-129)."""

    _TABLE = {
        "acoustic": ["h1", "h2"],
        "acoustic_habc": ["h1", "h2"],
        "acoustic_fwim_habc": ["h1", "h2"],
        "acoustic_lsrtm_habc": ["h1", "h2", "sh1", "sh2"],
        "acoustic_rho_habc": ["h1", "h2"],
        "acoustic_vti_lsrtm_habc": ["p1", "p2", "sp1", "sp2"],
        "acoustic_tti_lsrtm_habc": ["p1", "p2", "sp1", "sp2"],
        "elastic": ["vx", "vz", "txx", "tzz", "txz"],
        "vti_habc2": ["p1", "p2"],
        "tti_habc": ["p1", "p2"],
    }

    def __init__(self, equation="acoustic"):
        if equation not in self._TABLE:
            raise ValueError(f"seistorch_b200: equation '{equation}' is not on the accelerated path "
                             f"(supported: {sorted(self._TABLE)})")
        self.wavefields = list(self._TABLE[equation])


def field_channels(equation):
    """wavefield name -> field channel of the kernel state (the 'current' member of a
    second-order pair carries the channel; the 'previous' member cannot be a source or
    receiver on the whole-loop path)."""
    names = Wavefield(equation).wavefields
    if equation == "elastic":
        return {n: i for i, n in enumerate(names)}
    return {names[2 * k]: k for k in range(len(names) // 2)}
