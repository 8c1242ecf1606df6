#!/usr/bin/env python3
"""Bring-up check of the tcgen05 fixed-operand product (K6) against the oracle, then a timing at the
BASELINE config-5 shape (Cl(8,0,0), 2^24 rows, f32).  Run on the GPU box under `timeout`."""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import largecrimsoncanine as L  # noqa: E402
import oracle  # noqa: E402


def check(sig, n, side):
    alg, o = L.Algebra(*sig), oracle.OracleAlgebra(*sig)
    nb = o.num_blades
    rng = np.random.default_rng(n + nb + side)
    m = rng.standard_normal((1, nb)).astype(np.float32)
    x = rng.standard_normal((n, nb)).astype(np.float32)
    M, X = L.MultivectorBatch.from_numpy(alg, m), L.MultivectorBatch.from_numpy(alg, x)
    got = (M.geometric_product(X) if side == 0 else X.geometric_product(M)).to_numpy()
    a64, b64 = (m.astype(np.float64), x.astype(np.float64)) if side == 0 else (x.astype(np.float64), m.astype(np.float64))
    ref = o.geometric_product(a64, b64)
    bound = o.product_abs_bound("gp", np.abs(a64), np.abs(b64))
    err = np.abs(got - ref) / np.maximum(np.abs(ref), bound)
    rel_row = np.abs(got - ref).max(axis=1) / np.abs(ref).max(axis=1)
    print(f"sig={sig} n={n} side={side}: max err/bound {err.max():.3e}  max row-rel {rel_row.max():.3e}", flush=True)
    return err.max() <= 1e-5 and rel_row.max() <= 1e-5


def timing():
    import torch
    from largecrimsoncanine_b200 import cabi

    lib, alg = cabi.lib(), cabi.Algebra(8, 0, 0)
    n, nb = 1 << 24, 256
    dev = torch.device("cuda", 0)
    x = torch.randn((nb, n), device=dev, dtype=torch.float32)
    out = torch.empty_like(x)
    m = torch.randn((nb, 4), device=dev, dtype=torch.float32)  # one-row SoA view with ld = 4
    vx = cabi.view(x.data_ptr(), n, n, cabi.LCC_F32, cabi.LCC_SOA)
    vo = cabi.view(out.data_ptr(), n, n, cabi.LCC_F32, cabi.LCC_SOA)
    vm = cabi.view(m.data_ptr(), 1, 4, cabi.LCC_F32, cabi.LCC_SOA)
    st = torch.cuda.Stream()
    torch.cuda.synchronize()
    with torch.cuda.stream(st):
        sp = C.c_void_p(st.cuda_stream)
        for _ in range(3):
            cabi.check(lib.lcc_batch_product(alg.h, 0, C.byref(vm), C.byref(vx), C.byref(vo), 0, sp))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(10):
            cabi.check(lib.lcc_batch_product(alg.h, 0, C.byref(vm), C.byref(vx), C.byref(vo), 0, sp))
        e1.record(st)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    flops = 3 * 2 * nb * nb * n
    print(f"cfg5 Cl(8,0,0) 2^24 rows f32: {ms:.3f} ms/step  {n / ms / 1e6:.2f} G mv/s  HBM {2 * nb * 4 * n / ms / 1e6:.0f} GB/s  "
          f"tensor {flops / ms / 1e9:.0f} TFLOP/s (tf32 x3)  useful {2 * nb * nb * n / ms / 1e9:.0f} TFLOP/s", flush=True)
    # spot check a few rows against the oracle
    o = oracle.OracleAlgebra(8, 0, 0)
    idx = torch.tensor([0, 1, 127, 128, n // 2, n - 1], device=dev)
    xs = x[:, idx].T.contiguous().cpu().numpy().astype(np.float64)
    ref = o.geometric_product(m[:, 0].cpu().numpy().astype(np.float64)[None, :], xs)
    got = out[:, idx].T.contiguous().cpu().numpy()
    print("spot rel err", np.abs(got - ref).max() / np.abs(ref).max(), flush=True)


if __name__ == "__main__":
    ok = True
    for sig in [(6, 0, 0), (4, 2, 1), (8, 0, 0)]:
        for n in (128, 1000, 70001):
            for side in (0, 1):
                ok &= check(sig, n, side)
    print("ALL OK" if ok else "FAILED", flush=True)
    if ok and "--time" in sys.argv:
        timing()
    sys.exit(0 if ok else 1)
