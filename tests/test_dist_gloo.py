"""CPU, world_size 2 over gloo: the N>1 host logic -- disjoint shot shards and ONE flattened
all-reduce of the parameter gradients giving the same result as the single-process sum."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from seistorch_b200 import parallel
    shots = parallel.shard_shots(7, rank, world)
    # stand-in for the per-shot gradient kernel: a deterministic function of the shot id
    p1 = torch.nn.Parameter(torch.zeros(5, 6))
    p2 = torch.nn.Parameter(torch.zeros(3))
    p1.grad = sum(torch.full((5, 6), float(s + 1)) for s in shots)
    p2.grad = sum(torch.arange(3.0) * (s + 1) for s in shots)
    parallel.allreduce_gradients([p1, p2])
    out[rank] = (shots, p1.grad.clone().numpy(), p2.grad.clone().numpy())
    dist.destroy_process_group()


def test_shard_and_allreduce_world2():
    from seistorch_b200 import parallel
    assert parallel.shard_shots(7, 0, 2) == [0, 2, 4, 6] and parallel.shard_shots(7, 1, 2) == [1, 3, 5]
    assert sorted(sum((parallel.shard_shots(128, r, 8) for r in range(8)), [])) == list(range(128))
    mgr = mp.Manager()
    out = mgr.dict()
    port = _free_port()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    total = sum(range(1, 8))
    for r in range(2):
        shots, g1, g2 = out[r]
        assert np.array_equal(g1, np.full((5, 6), float(total)))
        assert np.array_equal(g2, np.arange(3.0) * total)
    assert sorted(out[0][0] + out[1][0]) == list(range(7))


def test_allreduce_is_noop_without_process_group():
    from seistorch_b200 import parallel
    p = torch.nn.Parameter(torch.zeros(2))
    p.grad = torch.ones(2)
    parallel.allreduce_gradients([p])
    assert torch.equal(p.grad, torch.ones(2))
