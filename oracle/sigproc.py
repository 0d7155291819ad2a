"""TEST INFRASTRUCTURE ONLY (oracle).  CPU restatement of the reference's device-side record filter:
seistorch/signal.py:49-101 with backend='torch' = torchaudio.functional.filtfilt(x.double(), a, b, clamp=False)
along time = causal lfilter, flip, causal lfilter, flip (zero initial state, no padding), cast to float32.
Never imported by the product path."""
from __future__ import annotations

import numpy as np
from scipy import signal


def butter(order, freqs, dt):
    """signal.py:58-76."""
    freqs = list(freqs) if isinstance(freqs, (list, tuple)) else [freqs]
    mode = "lowpass" if len(freqs) == 1 else "bandpass"
    wn = [2 * f / (1 / dt) for f in freqs]
    return signal.butter(order, Wn=wn[0] if len(wn) == 1 else wn, btype=mode)


def filtfilt(x, b, a):
    """x: (nt, ...) array; returns float32."""
    y = signal.lfilter(b, a, np.asarray(x, dtype=np.float64), axis=0)
    y = signal.lfilter(b, a, y[::-1], axis=0)[::-1]
    return y.astype(np.float32)
