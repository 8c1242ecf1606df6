"""Host logic of the N > 1 path on CPU (no GPU): the row partition through the C ABI (lcc_shard_rows), argument
validation of the sharding entry points, and -- in a world_size-2 gloo process group -- the one piece of plumbing that
is Python: getting rank 0's 128-byte NCCL unique id to every rank.  The partials, the launches and the all-reduce are
native (csrc/shard_group.cu) and run on GPUs (tests/test_gpu_parity.py::test_sharded_batch_single_rank,
tests/tools/multi_gpu_check.py)."""
import ctypes as C
import os
import socket
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_rows_tiles_the_batch():
    from largecrimsoncanine_b200.sharded import shard_rows

    for n in (0, 1, 2, 7, 8, 1000, 2**28 + 3, 2**40 + 11):
        for world in (1, 2, 3, 4, 8):
            blocks = [shard_rows(n, r, world) for r in range(world)]
            assert blocks[0].start == 0 and blocks[-1].end == n
            assert all(a.end == b.start for a, b in zip(blocks, blocks[1:]))
            sizes = [b.rows for b in blocks]
            assert max(sizes) - min(sizes) <= 1 and sum(sizes) == n
            assert [(b.start, b.end) for b in blocks] == [(r * n // world, (r + 1) * n // world) for r in range(world)]
    with pytest.raises(ValueError):
        shard_rows(10, 2, 2)
    with pytest.raises(ValueError):
        shard_rows(-1, 0, 2)


def test_shard_entry_points_validate_and_refuse_without_a_gpu():
    from largecrimsoncanine_b200 import cabi

    lib = cabi.lib()
    g = C.c_void_p()
    assert lib.lcc_shard_group_create_rank(0, 2, 2, None, C.byref(g)) == cabi.LCC_ERR_VALUE  # rank outside the world
    assert lib.lcc_shard_group_create_rank(0, 2, 1, None, C.byref(g)) == cabi.LCC_ERR_VALUE  # multi-rank group without an id
    assert b"unique id" in lib.lcc_last_error()
    assert lib.lcc_shard_group_local_size(None) == 0 and lib.lcc_shard_group_world_size(None) == 0
    if not os.path.exists("/dev/nvidiactl"):
        assert lib.lcc_shard_group_create(0, None, C.byref(g)) == cabi.LCC_ERR_NO_DEVICE  # never emulated on the host
        assert b"no CPU fallback" in lib.lcc_last_error()


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    import datetime

    import torch.distributed as dist

    from largecrimsoncanine_b200 import sharded

    # loopback only: gloo otherwise resolves the container's hostname to pick an interface, which may not resolve (or hang)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank),
                      GLOO_SOCKET_IFNAME="lo")
    token = bytes(range(128))
    calls = []

    def make_id():  # stands in for lcc_shard_unique_id (NCCL), which only rank 0 may call
        calls.append(rank)
        return token

    try:
        via_store = sharded.distribute_unique_id(make_id, rank, world)  # no process group yet: TCPStore on MASTER_PORT+17
        dist.init_process_group("gloo", rank=rank, world_size=world, timeout=datetime.timedelta(seconds=90))
        via_group = sharded.distribute_unique_id(make_id, rank, world)  # process group present: one broadcast
        info = sharded.shard_rows(1001, rank, world)
        out.put((rank, via_store == token and via_group == token, calls == ([0, 0] if rank == 0 else []), info.rows))
    except BaseException as exc:  # a worker that dies silently would leave the parent waiting for its timeout
        import traceback

        out.put((rank, "error", f"{type(exc).__name__}: {exc}\n{traceback.format_exc()}", 0))
        raise
    finally:
        if dist.is_initialized():
            dist.destroy_process_group()


def _run_two_ranks():
    import queue

    import torch.multiprocessing as mp

    port = None
    for _ in range(64):  # a free port whose +17 neighbour (the TCPStore of distribute_unique_id) is free as well
        with socket.socket() as s, socket.socket() as s17:
            s.bind(("127.0.0.1", 0))
            cand = s.getsockname()[1]
            try:
                s17.bind(("127.0.0.1", cand + 17))
            except OSError:
                continue
            port = cand
            break
    assert port is not None
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results, problem = [], None
    try:
        for _ in procs:
            results.append(q.get(timeout=240))  # (a cold `import torch` in a spawned worker can take a minute right after a build)
    except queue.Empty:
        problem = "a rank did not answer within 240 s"
    errors = [r[2] for r in results if r[1] == "error"]
    if errors:
        problem = "; ".join(errors)
    for p in procs:
        p.join(timeout=5 if problem else 120)
        if p.is_alive():
            p.terminate()
            p.join(timeout=10)
        elif not problem and p.exitcode != 0:
            problem = f"rank exited with code {p.exitcode}"
    return results, problem


def test_two_ranks_share_the_unique_id_over_gloo():
    # the rendezvous (two TCP stores on loopback ports picked here) is retried on a fresh pair of ports: a port can be taken
    # between the probe and the bind
    for attempt in range(3):
        results, problem = _run_two_ranks()
        if problem is None:
            break
    assert problem is None, problem
    assert all(same and only_rank0 for _, same, only_rank0, _ in results)
    assert sum(rows for _, _, _, rows in results) == 1001
