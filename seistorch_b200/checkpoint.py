"""``checkpoint`` entry point of seistorch/checkpoint.py:232 and checkpoint_new.py:220
(first- and second-order equations share this implementation).

The reference's CheckpointFunction trades memory for an *approximate* gradient: it saves
1-cell boundary strips and reconstructs the wavefield backwards in time, which is exact
only where the damping is zero (BS vs AD differs by ~2 %).  The
accelerated path does not need it: ``WaveRNN.forward`` runs the whole time loop in one
autograd.Function with the exact discrete adjoint and K-step checkpoint/recompute
(seistorch_b200/engine.py).  This per-step entry point therefore simply evaluates the
(sm_100a) step op, which carries its own exact VJP; state is kept on the autograd graph,
never in class attributes (fixes checkpoint_new.py:111-112).
"""
from __future__ import annotations


def checkpoint(function, backfunction, source_function, save_condition, para_counts, *args,
               use_reentrant: bool = True, habcs=None, **kwargs):
    if kwargs:
        raise ValueError("Unexpected keyword arguments: " + ",".join(arg for arg in kwargs))
    return function(*args, habcs=habcs)
