"""Multi-GPU layer: one process per GPU (torchrun), contiguous row shards, NCCL only for the
cross-batch sum / mean.

Every batched operation of the reference is row-independent (`for row in 0..n`,
/root/reference/src/batch.rs:384), so a batch of N rows is split into `world` contiguous blocks
[g*N/G, (g+1)*N/G) and each rank runs the single-GPU kernels on its block: no halo, no data-path
collective.  Broadcast operands (one-row batches, rotors) are replicated.  The only exchange is

    sum / mean  =  per-GPU partial (lcc_batch_sum, num_blades values)  ->  all_reduce(SUM)  [-> / N_total]

`torch.distributed` is the plumbing (NCCL on GPUs; gloo in the CPU tests of the host logic).
"""
from __future__ import annotations

import os
from dataclasses import dataclass

__all__ = ["ShardInfo", "shard_rows", "init_process_group", "allreduce_sum", "allreduce_mean", "ShardedBatch"]


@dataclass(frozen=True)
class ShardInfo:
    rank: int
    world: int
    start: int
    end: int

    @property
    def rows(self) -> int:
        return self.end - self.start


def shard_rows(n_total: int, rank: int, world: int) -> ShardInfo:
    """Contiguous row block of `rank`: [rank*N//G, (rank+1)*N//G).  Blocks tile [0, N) exactly,
    differ in size by at most one row, and are empty only when N < G."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world: {rank}/{world}")
    if n_total < 0:
        raise ValueError("negative batch size")
    return ShardInfo(rank, world, rank * n_total // world, (rank + 1) * n_total // world)


def init_process_group(backend: str | None = None):
    """Join the torchrun rendezvous (RANK / WORLD_SIZE / LOCAL_RANK / MASTER_*), bind this process to
    its GPU and return (rank, world, local_rank).  A no-op group of one when launched plainly."""
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    cuda = torch.cuda.is_available()
    if cuda:
        torch.cuda.set_device(local)
        from . import cabi

        cabi.check(cabi.lib().lcc_set_device(local))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        backend = backend or ("nccl" if cuda else "gloo")
        kwargs = {"device_id": torch.device("cuda", local)} if backend == "nccl" else {}
        dist.init_process_group(backend, rank=rank, world_size=world, **kwargs)
    return rank, world, local


def _world():
    import torch.distributed as dist

    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def allreduce_sum(partial):
    """Sum per-rank partial column sums (a tensor of num_blades values, CPU or CUDA) in place."""
    import torch.distributed as dist

    if _world() > 1:
        dist.all_reduce(partial, op=dist.ReduceOp.SUM)
    return partial


def allreduce_mean(partial, n_local: int):
    """Global mean from per-rank partial sums and per-rank row counts: sum(partials) / sum(n_local)."""
    import torch
    import torch.distributed as dist

    total = torch.tensor([float(n_local)], dtype=torch.float64, device=partial.device)
    if _world() > 1:
        dist.all_reduce(partial, op=dist.ReduceOp.SUM)
        dist.all_reduce(total, op=dist.ReduceOp.SUM)
    n = total.item()
    return partial / n if n > 0 else partial


class ShardedBatch:
    """This rank's row block of a global batch: a device-resident MultivectorBatch plus the global
    row count.  Row-wise operations are forwarded to the local batch unchanged; sum / mean add the
    NCCL all-reduce."""

    def __init__(self, local_batch, n_total: int, info: ShardInfo):
        self.local, self.n_total, self.info = local_batch, int(n_total), info

    @classmethod
    def from_numpy(cls, algebra, coeffs_global, rank=None, world=None):
        """Each rank keeps rows [start, end) of the same global host array."""
        import largecrimsoncanine as L

        rank = int(os.environ.get("RANK", "0")) if rank is None else rank
        world = _world() if world is None else world
        info = shard_rows(len(coeffs_global), rank, world)
        return cls(L.MultivectorBatch.from_numpy(algebra, coeffs_global[info.start:info.end]), len(coeffs_global), info)

    def _wrap(self, local):
        return ShardedBatch(local, self.n_total, self.info)

    def __len__(self):
        return self.n_total

    # row-wise operations: no communication
    def geometric_product(self, other):
        return self._wrap(self.local.geometric_product(other.local if isinstance(other, ShardedBatch) else other))

    def outer_product(self, other):
        return self._wrap(self.local.outer_product(other.local if isinstance(other, ShardedBatch) else other))

    def left_contraction(self, other):
        return self._wrap(self.local.left_contraction(other.local if isinstance(other, ShardedBatch) else other))

    def sandwich(self, rotor):
        return self._wrap(self.local.sandwich(rotor))

    def grade(self, k):
        return self._wrap(self.local.grade(k))

    def reverse(self):
        return self._wrap(self.local.reverse())

    def to_numpy_local(self):
        return self.local.to_numpy()

    # the one collective of the path
    def _partial(self):
        import ctypes as C

        import numpy as np
        import torch

        from . import cabi

        nb = self.local.num_blades()
        tdt = torch.float32 if self.local.dtype == np.float32 else torch.float64
        partial = torch.zeros(nb, dtype=tdt, device="cuda")
        code = cabi.LCC_F32 if tdt == torch.float32 else cabi.LCC_F64
        view = cabi.view(self.local.data_ptr, self.local.len(), self.local.ld, code, cabi.LCC_SOA)
        alg = self.local.algebra()
        handle = cabi.Algebra(alg.p, alg.q, alg.r)
        stream = C.c_void_p(torch.cuda.current_stream().cuda_stream or 1)  # 0 -> cudaStreamLegacy
        import largecrimsoncanine as L

        L.synchronize()  # the batch was produced on the library's stream
        cabi.check(cabi.lib().lcc_batch_sum(handle.h, C.byref(view), C.c_void_p(partial.data_ptr()), stream))
        return partial

    def sum(self):
        return allreduce_sum(self._partial()).cpu().numpy()

    def mean(self):
        return allreduce_mean(self._partial(), self.local.len()).cpu().numpy()
