"""Shot-parallel data parallelism: one process per GPU, shots sharded across ranks, one
all-reduce of the model gradient per iteration.

Mirrors what the reference gets from DistributedSampler + DistributedDataParallel
(seistorch_dist.py:123,151-156) or from its MPI allreduce(SUM) (fwi.py:365), without
DDP's per-forward buffer broadcast: the path has exactly one exchange step.
"""
from __future__ import annotations

from typing import List, Sequence

import torch
import torch.distributed as dist


def shard_shots(nshots: int, rank: int, world_size: int) -> List[int]:
    """Disjoint strided shard of the shot list (DistributedSampler's layout without
    shuffling or padding: rank r takes shots r, r+W, r+2W, ...)."""
    return list(range(rank, nshots, world_size))


def allreduce_gradients(params: Sequence[torch.Tensor], average: bool = False) -> None:
    """Sum (fwi.py:365) or average (DDP, seistorch_dist.py:123) the ``.grad`` of the model
    parameters over all ranks with ONE flattened all-reduce."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    if average:
        flat /= dist.get_world_size()
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off:off + n].view_as(g))
        off += n
