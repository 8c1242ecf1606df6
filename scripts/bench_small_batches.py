"""Per-call cost of the batched API on small batches (launch-bound regime): microseconds per call of the geometric
product through the raw C ABI (device-resident views) and through the `largecrimsoncanine` extension, host-timed over
many back-to-back calls with one synchronisation at the end.  One JSON line per case."""
import ctypes as C
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch

import largecrimsoncanine as L
from largecrimsoncanine_b200 import cabi

lib = cabi.lib()
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(stream)
sptr = C.c_void_p(stream.cuda_stream)


def per_call_us(fn, calls):
    for _ in range(20):
        fn()
    torch.cuda.synchronize()
    lib.lcc_stream_sync(None)
    t0 = time.perf_counter()
    for _ in range(calls):
        fn()
    torch.cuda.synchronize()
    lib.lcc_stream_sync(None)
    return (time.perf_counter() - t0) / calls * 1e6


for sig, dt in (((3, 0, 0), "f64"), ((3, 0, 1), "f32"), ((4, 1, 0), "f64")):
    alg = cabi.Algebra(*sig)
    ealg = L.Algebra(*sig)
    nb = alg.num_blades
    tdt, code, ndt = (torch.float32, cabi.LCC_F32, np.float32) if dt == "f32" else (torch.float64, cabi.LCC_F64, np.float64)
    for n in (1, 1000, 100_000, 1_000_000):
        a, b, out = (torch.randn((n, nb), device=dev, dtype=tdt) for _ in range(3))
        va, vb, vo = (cabi.view(t.data_ptr(), n, nb, code, cabi.LCC_AOS) for t in (a, b, out))
        us_abi = per_call_us(lambda: cabi.check(lib.lcc_batch_product(alg.h, cabi.LCC_OP_GP, C.byref(va), C.byref(vb), C.byref(vo), 0, sptr)), 500)
        row = {"sig": sig, "dtype": dt, "rows": n, "c_abi_us_per_call": round(us_abi, 2)}
        if dt == "f64":
            A = L.MultivectorBatch.from_numpy(ealg, a.cpu().numpy())
            B = L.MultivectorBatch.from_numpy(ealg, b.cpu().numpy())
            row["extension_us_per_call"] = round(per_call_us(lambda: A.geometric_product(B), 300), 2)
        row["stream_floor_us"] = round(3 * nb * a.element_size() * n / 7.0e6, 2)  # at 7 TB/s
        print(json.dumps(row), flush=True)
