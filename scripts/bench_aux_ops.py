"""Bandwidth of the auxiliary streaming kernels (K3 unary / add / scale, K4 norms, K5 sum, K7 layout conversion)
through the C ABI, both layouts.  One JSON line per case: algorithmic bytes / CUDA-event time."""
import ctypes as C
import json
import sys

import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

from largecrimsoncanine_b200 import cabi

lib = cabi.lib()
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(stream)
sptr = C.c_void_p(stream.cuda_stream)


def timed(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(iters):
        fn()
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


ONLY = sys.argv[1] if len(sys.argv) > 1 else None  # substring filter on the op name (used for ncu captures)


def run(sig, dt, log2n):
    alg = cabi.Algebra(*sig)
    nb, n = alg.num_blades, 1 << log2n
    tdt, code, es = (torch.float32, cabi.LCC_F32, 4) if dt == "f32" else (torch.float64, cabi.LCC_F64, 8)
    for layout_name, layout in (("soa", cabi.LCC_SOA), ("aos", cabi.LCC_AOS)):
        shape = (nb, n) if layout == cabi.LCC_SOA else (n, nb)
        a, b, out = (torch.randn(shape, device=dev, dtype=tdt) for _ in range(3))
        norms = torch.empty(n, device=dev, dtype=tdt)
        sums = torch.empty(nb, device=dev, dtype=tdt)

        def v(t):
            return cabi.view(t.data_ptr(), n, t.stride(0), code, layout)

        va, vb, vo = v(a), v(b), v(out)
        cases = {
            "reverse": (2 * nb * es, lambda: cabi.check(lib.lcc_batch_unary(alg.h, cabi.LCC_UN_REVERSE, 0, C.byref(va), C.byref(vo), sptr))),
            "grade1": (2 * nb * es, lambda: cabi.check(lib.lcc_batch_unary(alg.h, cabi.LCC_UN_GRADE, 1, C.byref(va), C.byref(vo), sptr))),
            "scale": (2 * nb * es, lambda: cabi.check(lib.lcc_batch_scale(alg.h, C.byref(va), 0.5, C.byref(vo), sptr))),
            "add": (3 * nb * es, lambda: cabi.check(lib.lcc_batch_elementwise(alg.h, cabi.LCC_EW_ADD, C.byref(va), C.byref(vb), C.byref(vo), sptr))),
            "norm": (nb * es + es, lambda: cabi.check(lib.lcc_batch_norm(alg.h, cabi.LCC_NORM, C.byref(va), C.c_void_p(norms.data_ptr()), 0, sptr))),
            "normalized": (2 * nb * es, lambda: cabi.check(lib.lcc_batch_normalized(alg.h, C.byref(va), C.byref(vo), 0, sptr))),
            "sum": (nb * es, lambda: cabi.check(lib.lcc_batch_sum(alg.h, C.byref(va), C.c_void_p(sums.data_ptr()), sptr))),
        }
        if nb <= 16:  # K2r: one rotor per row (3 tiles), and one device-resident rotor row for the whole batch (2 tiles)
            r1 = cabi.view(b.data_ptr(), 1, b.stride(0) if layout == cabi.LCC_SOA else nb, code, layout)
            cases["sandwich_rows"] = (3 * nb * es, lambda: cabi.check(lib.lcc_batch_sandwich_rows(alg.h, C.byref(vb), C.byref(va), C.byref(vo), 0, sptr)))
            cases["sandwich_rows_bcast"] = (2 * nb * es, lambda: cabi.check(lib.lcc_batch_sandwich_rows(alg.h, C.byref(r1), C.byref(va), C.byref(vo), 0, sptr)))
            # the usual case: the one rotor is an even versor (rotor / motor) -- its odd-grade coefficients are zero, which the
            # kernel tests once per thread and then runs the pruned term lists
            rot = torch.randn(nb, device=dev, dtype=tdt)
            rot[[k for k in range(nb) if bin(k).count("1") % 2]] = 0
            keep_rot = rot.reshape(nb, 1).repeat(1, 4).contiguous() if layout == cabi.LCC_SOA else rot.reshape(1, nb).contiguous()
            r2 = cabi.view(keep_rot.data_ptr(), 1, 4 if layout == cabi.LCC_SOA else nb, code, layout)
            # one EVEN rotor per row (a rotor / motor field): the warps vote and run the pruned term lists
            rows_even = b.clone()
            odd = [k for k in range(nb) if bin(k).count("1") % 2]
            if layout == cabi.LCC_SOA:
                rows_even[odd, :] = 0
            else:
                rows_even[:, odd] = 0
            vre = v(rows_even)
            cases["sandwich_rows_even"] = (3 * nb * es, lambda: cabi.check(lib.lcc_batch_sandwich_rows(alg.h, C.byref(vre), C.byref(va), C.byref(vo), 0, sptr)))
            cases["sandwich_rows_bcast_even"] = (2 * nb * es, lambda: cabi.check(lib.lcc_batch_sandwich_rows(alg.h, C.byref(r2), C.byref(va), C.byref(vo), 0, sptr)))
        if alg.r == 0:
            cases["dual"] = (2 * nb * es, lambda: cabi.check(lib.lcc_batch_unary(alg.h, cabi.LCC_UN_DUAL, 0, C.byref(va), C.byref(vo), sptr)))
        for name, (bytes_per, fn) in cases.items():
            if ONLY and ONLY not in name:
                continue
            ms = timed(fn)
            print(json.dumps({"sig": sig, "dtype": dt, "rows": n, "layout": layout_name, "op": name, "ms": round(ms, 4),
                              "GBps": round(bytes_per * n / ms / 1e6, 1)}), flush=True)
    # K7: AoS <-> SoA
    x_aos = torch.randn((n, nb), device=dev, dtype=tdt)
    x_soa = torch.empty((nb, n), device=dev, dtype=tdt)
    vA = cabi.view(x_aos.data_ptr(), n, nb, code, cabi.LCC_AOS)
    vS = cabi.view(x_soa.data_ptr(), n, n, code, cabi.LCC_SOA)
    for name, src, dst in (("convert_aos_to_soa", vA, vS), ("convert_soa_to_aos", vS, vA)):
        if ONLY and ONLY not in name:
            continue
        ms = timed(lambda: cabi.check(lib.lcc_batch_convert(alg.h, C.byref(src), C.byref(dst), sptr)))
        print(json.dumps({"sig": sig, "dtype": dt, "rows": n, "layout": "both", "op": name, "ms": round(ms, 4),
                          "GBps": round(2 * nb * es * n / ms / 1e6, 1)}), flush=True)


if __name__ == "__main__":
    run((3, 0, 1), "f32", 26)
    if not ONLY or os.environ.get("AUX_ALL_SIGS"):
        run((4, 1, 0), "f64", 24)
        run((3, 0, 0), "f64", 26)
