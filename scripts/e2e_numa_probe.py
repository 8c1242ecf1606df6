"""8-rank host<->device bandwidth probe: what the PCIe / host-memory side of the e2e path can do when all GPUs
copy at once, with and without NUMA-local pinned buffers.  Launch with torchrun, one rank per GPU."""
import ctypes
import os
import time

import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
libc = ctypes.CDLL(None, use_errno=True)


def gpu_numa_node():
    bdf = torch.cuda.get_device_properties(local).pci_bus_id if hasattr(torch.cuda.get_device_properties(local), "pci_bus_id") else None
    try:
        import pynvml

        pynvml.nvmlInit()
        bdf = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(local)).busId
        bdf = (bdf.decode() if isinstance(bdf, bytes) else bdf).lower()
        if len(bdf.split(":")[0]) == 8:
            bdf = bdf[4:]
        return int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read()), bdf
    except Exception as e:  # noqa: BLE001
        return -1, f"{bdf} ({e})"


def set_mempolicy_bind(node):
    MPOL_DEFAULT, MPOL_BIND = 0, 2
    if node < 0:
        return libc.syscall(238, MPOL_DEFAULT, None, 0)
    mask = ctypes.c_ulong(1 << node)
    return libc.syscall(238, MPOL_BIND, ctypes.byref(mask), 64)


def status(key):
    for ln in open("/proc/self/status"):
        if ln.startswith(key):
            return ln.split(":", 1)[1].strip()


def copy_rate(nbytes, iters=6):
    h_in = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    h_out = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    h_in.zero_(); h_out.zero_()
    d_in = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    d_out = torch.zeros(nbytes, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    res = {}
    for mode in ("h2d", "d2h", "both"):
        dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(iters):
            if mode in ("h2d", "both"):
                with torch.cuda.stream(s1):
                    d_in.copy_(h_in, non_blocking=True)
            if mode in ("d2h", "both"):
                with torch.cuda.stream(s2):
                    h_out.copy_(d_out, non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        res[mode] = nbytes * iters / float(t.item()) / 1e9
    del h_in, h_out
    return res


node, bdf = gpu_numa_node()
info = f"rank {rank} gpu {local} bdf {bdf} numa {node} cpus_allowed {status('Cpus_allowed_list')} mems_allowed {status('Mems_allowed_list')} affinity {len(os.sched_getaffinity(0))}"
for r in range(world):
    if r == rank:
        print(info, flush=True)
    dist.barrier()
nbytes = 2 << 30
base = copy_rate(nbytes)
rc = set_mempolicy_bind(node)
bound = copy_rate(nbytes) if rc == 0 else None
set_mempolicy_bind(-1)
for r in range(world):
    if r == rank:
        print(f"rank {rank}: per-GPU GB/s (max-over-ranks time) default {base}  numa-bound(rc={rc}) {bound}", flush=True)
    dist.barrier()
dist.destroy_process_group()
