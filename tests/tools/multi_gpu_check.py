#!/usr/bin/env python3
"""Run under torchrun on N GPUs: the same global batch sharded N ways must give the same rows as the
single-GPU result and a sum / mean within tolerance of the oracle (plan item 5).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tests/tools/multi_gpu_check.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import largecrimsoncanine as L  # noqa: E402
import oracle  # noqa: E402
from largecrimsoncanine_b200 import sharded  # noqa: E402


def main():
    rank, world, local = sharded.init_process_group()
    rng = np.random.default_rng(2024)  # identical global arrays on every rank
    n = 1_000_003
    x, w = rng.standard_normal((n, 16)), rng.standard_normal((n, 16))
    sta, o = L.Algebra.sta(), oracle.OracleAlgebra(1, 3, 0)
    X = sharded.ShardedBatch.from_numpy(sta, x, rank, world)
    W = sharded.ShardedBatch.from_numpy(sta, w, rank, world)
    prod = X.grade(1).left_contraction(W)  # BASELINE config 4 pipeline: grade projection + inner product
    info = X.info
    local_ref = o.left_contraction(o.grade(x[info.start:info.end], 1), w[info.start:info.end])
    rows_ok = bool(np.max(np.abs(prod.to_numpy_local() - local_ref)) <= 1e-12 * np.abs(local_ref).max())
    total, mean = prod.sum(), prod.mean()
    full_ref = o.left_contraction(o.grade(x, 1), w)
    want, mag = o.sum(full_ref)
    sum_ok = bool(np.all(np.abs(total - want) <= 1e-12 * mag))
    mean_ok = bool(np.all(np.abs(mean - want / n) <= 1e-12 * mag / n))
    print(f"rank {rank}/{world} rows [{info.start},{info.end}) rows_ok={rows_ok} sum_ok={sum_ok} mean_ok={mean_ok}", flush=True)
    import torch.distributed as dist

    if dist.is_initialized():
        dist.barrier()
        dist.destroy_process_group()
    return 0 if rows_ok and sum_ok and mean_ok else 1


if __name__ == "__main__":
    sys.exit(main())
