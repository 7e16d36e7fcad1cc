"""Ensemble sharding across GPUs: one process per GPU, a static split of a permuted instance index, no data-path
collective.

The reference batches IVP instances with ``jax.vmap`` only (probdiffeq/backend/func.py:9-10):
instances never exchange data.  Step counts correlate with the parameters, so the instance index is permuted before it
is cut into contiguous per-rank slices (`permutation`, `shard_bounds`, `shard`); within a GPU the persistent kernels
balance dynamically.  The single collective on the path is the sum of the per-rank log-marginal-likelihoods in the
parameter-estimation configuration: `Communicator.allreduce_sum` issues it through the C ABI
(`pdeq_allreduce_sum_f64` = ``ncclAllReduce``) on an NCCL communicator created through the C ABI as well, whose
128-byte unique id travels over whatever process group `torch.distributed` already has.
"""

from __future__ import annotations

import ctypes as C

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


def shard(arrays, rank: int, world_size: int, seed: int | None = 0):
    """This rank's rows of every array in `arrays` (all with the ensemble axis leading) plus their global instance
    indices: rows `perm[lo:hi]` of the permuted ensemble (`seed=None`: no permutation). Works on numpy arrays and
    torch tensors alike; `unshard` puts per-rank results back in global order."""
    arrays = list(arrays)
    B = arrays[0].shape[0]
    lo, hi = shard_bounds(B, rank, world_size)
    idx = np.arange(lo, hi) if seed is None else permutation(B, seed)[lo:hi]
    out = []
    for a in arrays:
        if a.shape[0] != B:
            raise ValueError("all arrays must share the ensemble axis")
        out.append(a[torch.as_tensor(idx, device=a.device)] if isinstance(a, torch.Tensor) else a[idx])
    return out, idx


def unshard(per_rank_values, per_rank_indices):
    """Inverse of `shard` on the host: concatenate per-rank result arrays and restore the global instance order."""
    vals = np.concatenate([np.asarray(v) for v in per_rank_values], axis=0)
    idx = np.concatenate([np.asarray(i) for i in per_rank_indices], axis=0)
    out = np.empty_like(vals)
    out[idx] = vals
    return out


class QuadraticCostModel:
    """A cheap predictor of how expensive each ensemble member is, for `solve(..., cost_hint=...)`.

    The number of step attempts of an adaptive solve varies smoothly with the parameters and initial values of the
    ensemble member. `fit` regresses the attempt counts of a small pilot solve on a quadratic polynomial of the
    (standardised, by default log-transformed) inputs on the host; `predict` evaluates it for a whole ensemble on the
    device (a handful of small torch kernels). Only the ORDER of the predictions is used (longest first), so the
    model needs to rank well, not to be calibrated.
    """

    def __init__(self, mu, sd, w0, w1, W2, log):
        self.mu, self.sd, self.w0, self.w1, self.W2, self.log = mu, sd, float(w0), w1, W2, bool(log)
        self._dev = {}

    @classmethod
    def fit(cls, inputs: np.ndarray, cost: np.ndarray, log: bool = True) -> "QuadraticCostModel":
        x = np.asarray(inputs, dtype=np.float64)
        log = bool(log and np.all(x > 0))
        z = np.log(x) if log else x
        mu, sd = z.mean(axis=0), z.std(axis=0)
        sd = np.where(sd > 0, sd, 1.0)
        z = (z - mu) / sd
        F = z.shape[1]
        iu = np.triu_indices(F)
        feats = np.concatenate([np.ones((len(z), 1)), z, z[:, iu[0]] * z[:, iu[1]]], axis=1)
        w, *_ = np.linalg.lstsq(feats, np.asarray(cost, dtype=np.float64), rcond=None)
        W2 = np.zeros((F, F))
        W2[iu] = w[1 + F :]
        W2 = 0.5 * (W2 + W2.T)  # z^T W2 z reproduces the upper-triangular products
        return cls(mu, sd, w[0], w[1 : 1 + F], W2, log)

    def _on(self, device):
        key = str(device)
        if key not in self._dev:
            f64 = dict(dtype=torch.float64, device=device)
            self._dev[key] = (torch.as_tensor(self.mu, **f64), torch.as_tensor(1.0 / self.sd, **f64),
                              torch.as_tensor(self.w1, **f64), torch.as_tensor(self.W2, **f64))  # fmt: skip
        return self._dev[key]

    def predict(self, inputs: torch.Tensor) -> torch.Tensor:
        mu, isd, w1, W2 = self._on(inputs.device)
        z = ((torch.log(inputs) if self.log else inputs) - mu) * isd
        return torch.addmv(torch.sum((z @ W2) * z, dim=1), z, w1) + self.w0


def allreduce_sum(x: torch.Tensor) -> torch.Tensor:
    """Sum over ranks through `torch.distributed` (NCCL on GPUs, gloo on CPU); identity when not distributed."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(x, op=dist.ReduceOp.SUM)
    return x


class Communicator:
    """An NCCL communicator owned by the library's C ABI (pdeq_nccl_* in include/probdiffeq_b200.h).

    `Communicator.from_torch_distributed()` creates one over the ranks of the default process group: rank 0 draws the
    unique id, `dist.broadcast` carries it, every rank initialises with its current CUDA device.  With one rank it is
    a one-rank communicator (the all-reduce is then a copy), so callers need no special case.
    """

    def __init__(self, handle: int, rank: int, world_size: int):
        self._handle = C.c_void_p(handle)
        self.rank, self.world_size = rank, world_size

    @classmethod
    def from_torch_distributed(cls, device: torch.device | None = None) -> "Communicator":
        from probdiffeq_b200 import _lib

        lib = _lib.load()
        distributed = dist.is_available() and dist.is_initialized()
        rank = dist.get_rank() if distributed else 0
        world = dist.get_world_size() if distributed else 1
        device = device or torch.device("cuda", torch.cuda.current_device())
        uid = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            buf = (C.c_char * 128)()
            _lib.check(lib.pdeq_nccl_unique_id(buf), "pdeq_nccl_unique_id")
            uid = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).clone()
        if world > 1:
            on_gpu = dist.get_backend() == "nccl"
            t = uid.to(device) if on_gpu else uid
            dist.broadcast(t, src=0)
            uid = t.cpu()
        raw = (C.c_char * 128).from_buffer_copy(bytes(uid.numpy().tobytes()))
        handle = C.c_void_p()
        with torch.cuda.device(device):
            _lib.check(lib.pdeq_nccl_comm_init_rank(C.byref(handle), world, raw, rank), "pdeq_nccl_comm_init_rank")
        return cls(handle.value, rank, world)

    def count(self) -> int:
        from probdiffeq_b200 import _lib

        return int(_lib.load().pdeq_nccl_comm_count(self._handle))

    def allreduce_sum(self, x: torch.Tensor) -> torch.Tensor:
        """In-place sum of a contiguous float64 CUDA tensor over the communicator's ranks (ncclAllReduce on torch's
        current stream, through `pdeq_allreduce_sum_f64`)."""
        from probdiffeq_b200 import _lib

        if x.dtype != torch.float64 or not x.is_cuda or not x.is_contiguous():
            raise ValueError("allreduce_sum needs a contiguous float64 CUDA tensor")
        stream = C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)
        with torch.cuda.device(x.device):
            rc = _lib.load().pdeq_allreduce_sum_f64(self._handle, C.c_void_p(x.data_ptr()), x.numel(), stream)
        _lib.check(rc, "pdeq_allreduce_sum_f64")
        return x

    def destroy(self) -> None:
        from probdiffeq_b200 import _lib

        if self._handle:
            _lib.check(_lib.load().pdeq_nccl_comm_destroy(self._handle), "pdeq_nccl_comm_destroy")
            self._handle = C.c_void_p()
