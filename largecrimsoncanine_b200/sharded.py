"""Multi-GPU layer: thin ctypes binding of the native sharding API (`lcc_shard_*`, include/lcc_b200.h,
csrc/shard_group.cu).  Nothing here computes or communicates on its own: row partition, per-GPU streams,
launches and the NCCL all-reduce of the batch sum / mean all live behind the C ABI.

Every batched operation of the reference is row-independent (`for row in 0..n`,
/root/reference/src/batch.rs:384), so a batch of N rows is split into contiguous blocks
[k*N/W, (k+1)*N/W), one per GPU; broadcast operands (one-row batches, rotors) are replicated; the only
exchange is

    sum / mean  =  per-GPU fused partial (num_blades values)  ->  ncclAllReduce(SUM)  [-> / N_total]

Two ways to form a group:
  * `ShardGroup(devices=None)`      one process drives all (or the listed) GPUs of the box (ncclCommInitAll);
  * `ShardGroup.from_env()`         one rank per GPU (torchrun: RANK / WORLD_SIZE / LOCAL_RANK); rank 0's NCCL unique
                                    id travels through torch.distributed when a process group exists, else through a
                                    TCPStore on MASTER_ADDR:MASTER_PORT+17 -- 128 bytes of plumbing, once.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

from . import cabi

__all__ = ["ShardInfo", "shard_rows", "ShardGroup", "ShardedBatch", "distribute_unique_id"]

_OPS = {"gp": cabi.LCC_OP_GP, "outer": cabi.LCC_OP_OUTER, "inner": cabi.LCC_OP_LC, "lc": cabi.LCC_OP_LC}


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
    """Contiguous row block of `rank` (lcc_shard_rows): [rank*N//W, (rank+1)*N//W).  Blocks tile [0, N) exactly,
    differ in size by at most one row, and are empty only when N < W."""
    b, e = C.c_int64(), C.c_int64()
    cabi.check(cabi.lib().lcc_shard_rows(int(n_total), int(rank), int(world), C.byref(b), C.byref(e)))
    return ShardInfo(int(rank), int(world), b.value, e.value)


def distribute_unique_id(make_id, rank: int, world: int) -> bytes:
    """Rank 0 calls `make_id()` (128 bytes); every rank returns the same bytes."""
    if world == 1:
        return make_id()
    try:
        import torch.distributed as dist

        if dist.is_available() and dist.is_initialized():
            box = [make_id() if rank == 0 else None]
            dist.broadcast_object_list(box, src=0)
            return bytes(box[0])
    except ImportError:
        pass
    from torch.distributed import TCPStore

    host = os.environ.get("MASTER_ADDR", "127.0.0.1")
    port = int(os.environ.get("MASTER_PORT", "29500")) + 17
    import datetime

    store = TCPStore(host, port, world, is_master=(rank == 0), timeout=datetime.timedelta(seconds=120))
    if rank == 0:
        store.set("lcc_shard_unique_id", make_id())
    return bytes(store.get("lcc_shard_unique_id"))


def _dtype_code(dt) -> int:
    dt = np.dtype(dt)
    if dt == np.float32:
        return cabi.LCC_F32
    if dt == np.float64:
        return cabi.LCC_F64
    raise TypeError(f"expected float64 or float32, got {dt}")


class ShardGroup:
    """Owning handle of an lcc_shard_group."""

    def __init__(self, devices=None, _handle=None, _rank=0):
        self._lib = cabi.lib()
        if _handle is not None:
            self.h = _handle
        else:
            self.h = C.c_void_p()
            if devices is None:
                cabi.check(self._lib.lcc_shard_group_create(0, None, C.byref(self.h)))
            else:
                arr = (C.c_int * len(devices))(*devices)
                cabi.check(self._lib.lcc_shard_group_create(len(devices), arr, C.byref(self.h)))
        self.local_size = self._lib.lcc_shard_group_local_size(self.h)
        self.world = self._lib.lcc_shard_group_world_size(self.h)
        self.first_rank = self._lib.lcc_shard_group_first_rank(self.h)

    @classmethod
    def from_env(cls, device=None):
        """One rank per GPU: RANK / WORLD_SIZE / LOCAL_RANK of the launcher (torchrun)."""
        rank = int(os.environ.get("RANK", "0"))
        world = int(os.environ.get("WORLD_SIZE", "1"))
        local = int(os.environ.get("LOCAL_RANK", "0")) if device is None else device
        lib = cabi.lib()

        def make_id():
            buf = C.create_string_buffer(128)
            cabi.check(lib.lcc_shard_unique_id(buf))
            return buf.raw

        uid = distribute_unique_id(make_id, rank, world) if world > 1 else None
        h = C.c_void_p()
        cabi.check(lib.lcc_shard_group_create_rank(local, world, rank, uid, C.byref(h)))
        return cls(_handle=h)

    def views(self):
        return (cabi.LccView * self.local_size)()

    def sync(self):
        cabi.check(self._lib.lcc_shard_group_sync(self.h))

    def stream(self, local_index=0) -> int:
        s = C.c_void_p()
        cabi.check(self._lib.lcc_shard_group_stream(self.h, local_index, C.byref(s)))
        return s.value

    def close(self):
        h, self.h = getattr(self, "h", None), None
        if h:
            self._lib.lcc_shard_group_destroy(h)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class ShardedBatch:
    """The local shards of a global batch of `n_total` rows: one device-resident blade-major batch per local GPU.
    Row-wise operations launch once per GPU with no communication; sum / mean add the NCCL all-reduce."""

    def __init__(self, group: ShardGroup, algebra: cabi.Algebra, views, n_total: int, dtype, owns=True):
        self.group, self.algebra, self.views, self.n_total, self.dtype, self._owns = group, algebra, views, int(n_total), np.dtype(dtype), owns

    # ---- construction / export
    @classmethod
    def empty(cls, group, algebra, n_total, dtype=np.float64):
        views = group.views()
        cabi.check(group._lib.lcc_shard_batch_alloc(group.h, algebra.h, _dtype_code(dtype), int(n_total), views, None))
        return cls(group, algebra, views, n_total, dtype)

    @classmethod
    def from_numpy(cls, group, algebra, coeffs_global):
        """Every rank passes the same global (N, num_blades) host array; each GPU uploads its own row block."""
        x = np.ascontiguousarray(coeffs_global)
        if x.ndim != 2 or x.shape[1] != algebra.num_blades:
            raise ValueError(f"coeffs must have {algebra.num_blades} columns but got {x.shape[1] if x.ndim == 2 else x.shape}")
        out = cls.empty(group, algebra, len(x), x.dtype)
        cabi.check(group._lib.lcc_shard_upload(group.h, algebra.h, x.ctypes.data, len(x), out.views))
        return out

    @classmethod
    def broadcast(cls, group, algebra, row):
        """A one-row operand replicated on every local GPU."""
        r = np.ascontiguousarray(row).reshape(-1)
        views = group.views()
        cabi.check(group._lib.lcc_shard_broadcast_alloc(group.h, algebra.h, _dtype_code(r.dtype), r.ctypes.data, views))
        return cls(group, algebra, views, 1, r.dtype)

    def to_numpy(self, out=None):
        """The global array; rows owned by other processes are left untouched (single-process groups fill all of it)."""
        nb = self.algebra.num_blades
        if out is None:
            out = np.zeros((self.n_total, nb), self.dtype)
        cabi.check(self.group._lib.lcc_shard_download(self.group.h, self.algebra.h, self.views, out.ctypes.data, self.n_total))
        return out

    def local_rows(self):
        """[(start, end)] of each local shard."""
        return [(lambda i: (i.start, i.end))(shard_rows(self.n_total, self.group.first_rank + k, self.group.world)) for k in range(self.group.local_size)]

    def __len__(self):
        return self.n_total

    def _like(self):
        return ShardedBatch.empty(self.group, self.algebra, self.n_total, self.dtype)

    # ---- row-wise operations: no communication
    def _product(self, other, op, flags=0):
        out = self._like() if self.n_total >= other.n_total else other._like()
        cabi.check(self.group._lib.lcc_shard_product(self.group.h, self.algebra.h, op, self.views, other.views, out.views, flags))
        return out

    def geometric_product(self, other, flags=0):
        return self._product(other, cabi.LCC_OP_GP, flags)

    def outer_product(self, other, flags=0):
        return self._product(other, cabi.LCC_OP_OUTER, flags)

    def left_contraction(self, other, flags=0):
        return self._product(other, cabi.LCC_OP_LC, flags)

    def sandwich(self, rotor, flags=0):
        r = np.ascontiguousarray(rotor, dtype=np.float64).reshape(-1)
        out = self._like()
        cabi.check(self.group._lib.lcc_shard_sandwich(self.group.h, self.algebra.h, r.ctypes.data_as(C.POINTER(C.c_double)), self.views, out.views, flags))
        return out

    def _unary(self, op, arg=0):
        out = self._like()
        cabi.check(self.group._lib.lcc_shard_unary(self.group.h, self.algebra.h, op, arg, self.views, out.views))
        return out

    def grade(self, k):
        return self._unary(cabi.LCC_UN_GRADE, k)

    def reverse(self):
        return self._unary(cabi.LCC_UN_REVERSE)

    # ---- the one collective of the path
    def _reduce(self, divisor):
        res = np.empty(self.algebra.num_blades, self.dtype)
        cabi.check(self.group._lib.lcc_shard_sum(self.group.h, self.algebra.h, self.views, divisor, res.ctypes.data))
        return res

    def sum(self):
        return self._reduce(0)

    def mean(self):
        return self._reduce(self.n_total)

    def product_sum(self, other, op="inner", grade=None, mean=False, flags=0):
        """sum (or mean) over all rows of (<self>_grade (op) other): the fused config-4 pipeline + all-reduce."""
        res = np.empty(self.algebra.num_blades, self.dtype)
        n = max(self.n_total, other.n_total)
        cabi.check(self.group._lib.lcc_shard_product_sum(self.group.h, self.algebra.h, _OPS[op], self.views, -1 if grade is None else int(grade),
                                                         other.views, n if mean else 0, res.ctypes.data, flags))
        return res

    def free(self):
        if self._owns and self.views is not None and getattr(self.group, "h", None):
            self.group._lib.lcc_shard_batch_free(self.group.h, self.views)
        self.views = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass
