"""Where does the multi-GPU step of the fused STA workload lose time?  Per rank: (A) the fused kernel alone,
(B) kernel + asynchronous all-reduce into a ring (what bench.py does), (C) the all-reduce alone, (D) kernel + one
all-reduce of the whole ring every 4 steps.  Prints max-over-ranks ms per step.  Launch with torchrun."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch
import torch.distributed as dist

from largecrimsoncanine_b200 import cabi

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
lib = cabi.lib()
cabi.check(lib.lcc_set_device(local))
alg = cabi.Algebra(1, 3, 0)
n, nb = 1 << 27, 16
stream = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(stream)
sptr = C.c_void_p(stream.cuda_stream)
a = torch.randn((nb, n), device=dev, dtype=torch.float64)
b = torch.randn((nb, n), device=dev, dtype=torch.float64)
va, vb = (cabi.view(t.data_ptr(), n, n, cabi.LCC_F64, cabi.LCC_SOA) for t in (a, b))
ring = [torch.zeros(nb, device=dev, dtype=torch.float64) for _ in range(4)]
big = torch.zeros(4 * nb, device=dev, dtype=torch.float64)


def kernel(dst):
    cabi.check(lib.lcc_batch_product_sum(alg.h, cabi.LCC_OP_LC, C.byref(va), 1, C.byref(vb), C.c_void_p(dst.data_ptr()), 0, sptr))


def run(variant, steps=20):
    pending = [None] * 4

    def step(i):
        j = i % 4
        if variant == "A":
            kernel(ring[j])
        elif variant == "B":
            if pending[j] is not None:
                pending[j].wait()
            kernel(ring[j])
            pending[j] = dist.all_reduce(ring[j], async_op=True)
        elif variant == "C":
            if pending[j] is not None:
                pending[j].wait()
            pending[j] = dist.all_reduce(ring[j], async_op=True)
        elif variant == "D":
            kernel(big[j * nb:(j + 1) * nb])
            if j == 3:
                if pending[0] is not None:
                    pending[0].wait()
                pending[0] = dist.all_reduce(big, async_op=True)

    for i in range(4):
        step(i)
    for h in pending:
        if h is not None:
            h.wait()
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(steps):
        step(i)
    for h in pending:
        if h is not None:
            h.wait()
    e1.record(stream)
    dist.barrier(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / steps], device=dev, dtype=torch.float64)
    lo = t.clone()
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"variant {variant}: max {t.item():.4f} ms  min {lo.item():.4f} ms per step", flush=True)


for v in ("A", "B", "C", "D", "A"):
    run(v)
dist.destroy_process_group()
