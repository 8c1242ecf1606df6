#!/usr/bin/env python3
"""The native sharding API on N GPUs, against the oracle: the same global batch sharded N ways must give bit-identical
rows (strict mode) and a sum / mean within tolerance, in BOTH group modes:

    python tests/tools/multi_gpu_check.py                       one process, all GPUs of the box (ncclCommInitAll)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        tests/tools/multi_gpu_check.py                          one rank per GPU (ncclCommInitRank)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import oracle  # noqa: E402
from largecrimsoncanine_b200 import cabi, sharded  # noqa: E402


def main():
    multi_process = int(os.environ.get("WORLD_SIZE", "1")) > 1
    if multi_process:
        cabi.check(cabi.lib().lcc_set_device(int(os.environ.get("LOCAL_RANK", "0"))))
        group = sharded.ShardGroup.from_env()
    else:
        group = sharded.ShardGroup()
    rng = np.random.default_rng(2024)  # identical global arrays on every rank
    n = 1_000_003
    x, w = rng.standard_normal((n, 16)), rng.standard_normal((n, 16))
    alg, o = cabi.Algebra(1, 3, 0), oracle.OracleAlgebra(1, 3, 0)
    X, W = sharded.ShardedBatch.from_numpy(group, alg, x), sharded.ShardedBatch.from_numpy(group, alg, w)
    prod = X.grade(1).left_contraction(W, cabi.LCC_FLAG_STRICT_FP)  # BASELINE config 4 pipeline: grade projection + inner product
    full_ref = o.left_contraction(o.grade(x, 1), w)
    got = prod.to_numpy()
    rows_ok = all(np.array_equal(got[b:e], full_ref[b:e]) for b, e in prod.local_rows())
    want, mag = o.sum(full_ref)
    total, mean, fused = prod.sum(), prod.mean(), X.product_sum(W, "inner", grade=1)
    sum_ok = bool(np.all(np.abs(total - want) <= 1e-12 * mag) and np.all(np.abs(fused - want) <= 1e-12 * mag))
    mean_ok = bool(np.all(np.abs(mean - want / n) <= 1e-12 * mag / n))
    # the pipelined form: 12 asynchronous reductions through the 8-slot ring, fetched afterwards
    lib = cabi.lib()
    for step in range(12):
        cabi.check(lib.lcc_shard_product_sum_async(group.h, alg.h, cabi.LCC_OP_LC, X.views, 1, W.views, 0, step % 8, 0))
    res = np.empty(16)
    async_ok = True
    for slot in range(8):
        cabi.check(lib.lcc_shard_reduce_fetch(group.h, alg.h, cabi.LCC_F64, slot, res.ctypes.data))
        async_ok = async_ok and bool(np.array_equal(res, fused))  # fixed reduction tree: bit-identical every time
    print(f"mode={'ranks' if multi_process else 'single-process'} rank {group.first_rank}/{group.world} local_gpus={group.local_size} "
          f"nccl={lib.lcc_shard_nccl_version()} rows {prod.local_rows()} rows_ok={rows_ok} sum_ok={sum_ok} mean_ok={mean_ok} async_ok={async_ok}", flush=True)
    group.sync()
    return 0 if rows_ok and sum_ok and mean_ok and async_ok else 1


if __name__ == "__main__":
    sys.exit(main())
