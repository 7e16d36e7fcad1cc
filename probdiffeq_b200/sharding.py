"""Ensemble sharding across GPUs: static split by instance index, no data-path collective.

The reference batches IVP instances with ``jax.vmap`` only; instances never
exchange data, so one process per GPU solves a contiguous slice of a (optionally permuted) instance
index.  The single collective on the path is the sum of per-rank log-marginal-likelihoods in the
parameter-estimation configuration.
"""

from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def shard_bounds(num_instances: int, rank: int, world_size: int) -> tuple[int, int]:
    """Half-open range of instance indices owned by `rank`; sizes differ by at most one."""
    base, rem = divmod(num_instances, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def permutation(num_instances: int, seed: int = 0) -> np.ndarray:
    """Instance permutation that decorrelates step counts from the index (load balance across ranks)."""
    return np.random.Generator(np.random.PCG64(seed)).permutation(num_instances)


def allreduce_sum(x: torch.Tensor) -> torch.Tensor:
    """Sum over ranks (NCCL on GPUs, gloo on CPU); identity when not distributed."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(x, op=dist.ReduceOp.SUM)
    return x
