"""Where do the TMA-staged kernels start to pay?  PGA f32 sandwich and CGA f64 gp at 2^12 .. 2^26 rows, both layouts,
with the staged path forced on and forced off through lcc_set_option.  One JSON line per (case, layout, rows)."""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch

from largecrimsoncanine_b200 import cabi

lib = cabi.lib()
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(stream)
sptr = C.c_void_p(stream.cuda_stream)
HUGE = 1 << 60


def timed(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(iters):
        fn()
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3  # us


def run(case, sig, dt, max_log2):
    alg = cabi.Algebra(*sig)
    nb = alg.num_blades
    tdt, code, es = (torch.float32, cabi.LCC_F32, 4) if dt == "f32" else (torch.float64, cabi.LCC_F64, 8)
    motor = np.zeros(16)
    motor[[0, 3, 5, 6, 9, 10, 12]] = [0.9, 0.1, -0.2, 0.3, 0.5, -0.4, 0.2]
    mptr = motor.ctypes.data_as(C.POINTER(C.c_double))
    for layout_name, layout in (("soa", cabi.LCC_SOA), ("aos", cabi.LCC_AOS)):
        for log2n in range(12, max_log2 + 1, 2):
            n = 1 << log2n
            shape = (nb, n) if layout == cabi.LCC_SOA else (n, nb)
            a, b, out = (torch.randn(shape, device=dev, dtype=tdt) for _ in range(3))
            va, vb, vo = (cabi.view(t.data_ptr(), n, t.stride(0), code, layout) for t in (a, b, out))
            if case == "pga_sandwich":
                fn = lambda: cabi.check(lib.lcc_batch_sandwich(alg.h, mptr, C.byref(va), C.byref(vo), 0, sptr))  # noqa: E731
                bytes_per = 2 * nb * es
            else:
                fn = lambda: cabi.check(lib.lcc_batch_product(alg.h, cabi.LCC_OP_GP, C.byref(va), C.byref(vb), C.byref(vo), 0, sptr))  # noqa: E731
                bytes_per = 3 * nb * es
            res = {}
            for mode, rows_opt, bytes_opt in (("direct", HUGE, HUGE), ("tma", 0, 0)):
                cabi.check(lib.lcc_set_option(b"aos_tma_min_rows", rows_opt))
                cabi.check(lib.lcc_set_option(b"soa_tma_min_bytes", bytes_opt))
                res[mode] = timed(fn, 50 if log2n <= 20 else 10)
            print(json.dumps({"case": case, "dtype": dt, "layout": layout_name, "log2_rows": log2n, "operand_MiB": round(n * nb * es / 2**20, 2),
                              "direct_us": round(res["direct"], 2), "tma_us": round(res["tma"], 2),
                              "direct_GBps": round(bytes_per * n / res["direct"] / 1e3, 1), "tma_GBps": round(bytes_per * n / res["tma"] / 1e3, 1)}), flush=True)
            del a, b, out


if __name__ == "__main__":
    run("pga_sandwich", (3, 0, 1), "f32", 26)
    run("cga_gp", (4, 1, 0), "f64", 24)
