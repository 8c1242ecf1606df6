"""Host logic of the N > 1 path on CPU: world_size-2 gloo process group (no GPU).

Covers the row partition (tiles [0, N) exactly for ragged N) and the all-reduce that combines
per-rank partial column sums into the global sum / mean -- checked against the oracle's column
sums of the whole array.  The per-rank partial itself is a CUDA kernel (tests/test_gpu_parity.py)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_rows_tiles_the_batch():
    from largecrimsoncanine_b200.sharded import shard_rows

    for n in (0, 1, 2, 7, 8, 1000, 2**28 + 3):
        for world in (1, 2, 3, 4, 8):
            blocks = [shard_rows(n, r, world) for r in range(world)]
            assert blocks[0].start == 0 and blocks[-1].end == n
            assert all(a.end == b.start for a, b in zip(blocks, blocks[1:]))
            sizes = [b.rows for b in blocks]
            assert max(sizes) - min(sizes) <= 1 and sum(sizes) == n
    with pytest.raises(ValueError):
        shard_rows(10, 2, 2)


def _worker(rank, world, port, n, out):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    import oracle
    from largecrimsoncanine_b200 import sharded

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(5)
        x = rng.standard_normal((n, 16))  # same global array on every rank (seeded)
        info = sharded.shard_rows(n, rank, world)
        o = oracle.OracleAlgebra(1, 3, 0)
        local = x[info.start:info.end]
        partial = torch.tensor(local.sum(axis=0)) if len(local) else torch.zeros(16, dtype=torch.float64)
        total = sharded.allreduce_sum(partial.clone()).numpy()
        mean = sharded.allreduce_mean(partial.clone(), info.rows).numpy()
        want, mag = o.sum(x)
        ok = bool(np.all(np.abs(total - want) <= 1e-13 * mag) and np.all(np.abs(mean - want / n) <= 1e-13 * mag / n))
        out.put((rank, ok, info.rows))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [1001, 1])
def test_two_rank_sum_and_mean_over_gloo(n):
    import torch.multiprocessing as mp

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in results)
    assert sum(rows for _, _, rows in results) == n
