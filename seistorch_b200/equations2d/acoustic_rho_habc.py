"""Plug-in module for equation 'acoustic_rho_habc' -- same module-level surface as
seistorch/equations2d/acoustic_rho_habc.py (`_time_step`, `_time_step_backward`,
`_time_step_backward_multiple`), backed by the sm_100a kernels.

`_time_step(*model_params, *wavefields, dt, h, d, habcs=None)` advances one step on the
GPU and returns the wavefields in the reference's order; the whole-time-loop path used by
WaveRNN.forward does not go through this function (see seistorch_b200/engine.py).
"""
from ..stepop import reverse_step_unavailable, time_step

EQUATION = "acoustic_rho_habc"


def _time_step(*args, **kwargs):
    return time_step(EQUATION, 2, *args, **kwargs)


_time_step_backward = reverse_step_unavailable(EQUATION)
_time_step_backward_multiple = reverse_step_unavailable(EQUATION)
