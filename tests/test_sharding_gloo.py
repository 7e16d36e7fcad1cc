"""The N > 1 path on CPU: two gloo ranks shard an ensemble by instance index and sum a per-rank quantity
(the ensemble log-marginal-likelihood reduction is the only collective on the path)."""

import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from probdiffeq_b200 import sharding


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, B, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = sharding.shard_bounds(B, rank, world)
    values = torch.arange(B, dtype=torch.float64)[lo:hi]  # stand-in for per-instance log-likelihoods
    total = sharding.allreduce_sum(values.sum().reshape(1))
    counts = sharding.allreduce_sum(torch.tensor([float(hi - lo)], dtype=torch.float64))
    if rank == 0:
        out.put((float(total), float(counts)))
    dist.destroy_process_group()


def test_two_ranks_cover_the_ensemble_exactly_once():
    B, world = 1001, 2
    ctx = mp.get_context("spawn")
    out = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, B, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    total, count = out.get()
    assert count == B and total == B * (B - 1) / 2


def test_shard_bounds_partition():
    for B in (0, 1, 7, 1 << 20):
        for world in (1, 2, 3, 8):
            spans = [sharding.shard_bounds(B, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    assert np.array_equal(sharding.permutation(10, seed=0), sharding.permutation(10, seed=0))
