"""Plug-in module for the 3D acoustic equation -- same module-level surface as
seistorch/equations3d/acoustic.py (`_time_step`, `_time_step_backward`), backed by the
sm_100a kernels (csrc/st_acoustic3d.cu)."""
from ..stepop import reverse_step_unavailable, time_step

EQUATION = "acoustic"


def _time_step(*args, **kwargs):
    return time_step(EQUATION, 3, *args, **kwargs)


_time_step_backward = reverse_step_unavailable(EQUATION)
_time_step_backward_multiple = reverse_step_unavailable(EQUATION)
