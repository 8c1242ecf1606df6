#!/usr/bin/env python3
"""lcc.torch (B200 kernels on array-of-structures tensors) against the reference's own GPU story:
stock PyTorch eager ops that materialise an (N, nb, nb) product tensor and scatter-add it
(the algorithm of the reference's lcc/torch/multivector.py:839-879, restated here as the baseline)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lcc.torch import TorchAlgebra, TorchMultivector  # noqa: E402


def eager_gp(a, b, signs, blades):
    products = a.unsqueeze(2) * b.unsqueeze(1) * signs.to(a.dtype).unsqueeze(0)
    out = torch.zeros_like(a)
    out.scatter_add_(1, blades.reshape(-1).expand(a.shape[0], -1), products.reshape(a.shape[0], -1))
    return out


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    rows = []
    for name, sig, dt, log2n in (("PGA Cl(3,0,1) gp", (3, 0, 1), torch.float32, 22), ("CGA Cl(4,1,0) gp", (4, 1, 0), torch.float64, 20),
                                 ("R3 Cl(3,0,0) gp", (3, 0, 0), torch.float64, 22)):
        alg = TorchAlgebra._from_signature(*sig, device="cuda")
        n, nb = 1 << log2n, alg.num_blades
        a = torch.randn(n, nb, device="cuda", dtype=dt)
        b = torch.randn(n, nb, device="cuda", dtype=dt)
        A, B = TorchMultivector(alg, a), TorchMultivector(alg, b)
        ms_ours = timeit(lambda: A.geometric_product(B))
        ms_eager = timeit(lambda: eager_gp(a, b, alg.cayley_signs, alg.cayley_blades), iters=3)
        err = (A.geometric_product(B).coeffs - eager_gp(a, b, alg.cayley_signs, alg.cayley_blades)).abs().max().item()
        es = a.element_size()
        ar = a.clone().requires_grad_(True)
        def fwd_bwd():
            out = TorchMultivector(alg, ar).geometric_product(B).coeffs
            out.backward(torch.ones_like(out))
        ms_fb = timeit(fwd_bwd, iters=5)
        rows.append({"case": name, "rows": n, "dtype": str(dt), "lcc_torch_ms": ms_ours, "torch_eager_ms": ms_eager,
                     "speedup": ms_eager / ms_ours, "lcc_torch_GBps": 3 * nb * es * n / ms_ours / 1e6, "max_abs_diff": err,
                     "fwd_bwd_ms": ms_fb})
        print(json.dumps(rows[-1]), flush=True)
    # sandwich with a fixed motor on AoS tensors (fused K2 kernel)
    alg = TorchAlgebra.pga(3, device="cuda")
    n = 1 << 26
    x = torch.randn(n, 16, device="cuda", dtype=torch.float32)
    r = torch.zeros(1, 16, device="cuda", dtype=torch.float32)
    r[0, [0, 3, 5, 6, 9, 10, 12]] = torch.randn(7, device="cuda")
    X, R = TorchMultivector(alg, x), TorchMultivector(alg, r)
    with torch.no_grad():
        ms = timeit(lambda: X.sandwich(R))
    print(json.dumps({"case": "PGA sandwich AoS f32 (lcc.torch)", "rows": n, "ms": ms, "GBps": 128 * n / ms / 1e6}), flush=True)
    # Cl(8,0,0): one DEVICE-resident operand times a batch of (N, 256) rows -- K8 on units of whole rows, the 16 x 16 factor
    # built on the device (no host copy of the operand); the dense-GEMM form of the same call beside it (matrix_rep = 0)
    from largecrimsoncanine_b200 import cabi
    alg = TorchAlgebra._from_signature(8, 0, 0, device="cuda")
    for dt, log2n in ((torch.float64, 22), (torch.float32, 23)):
        n = 1 << log2n
        x = torch.randn(n, 256, device="cuda", dtype=dt)
        m = torch.randn(1, 256, device="cuda", dtype=dt)
        X, M = TorchMultivector(alg, x), TorchMultivector(alg, m)
        with torch.no_grad():
            ms = timeit(lambda: M.geometric_product(X))
            cabi.check(cabi.lib().lcc_set_option(b"matrix_rep", 0))
            ms_gemm = timeit(lambda: M.geometric_product(X), iters=3)
            cabi.check(cabi.lib().lcc_set_option(b"matrix_rep", 2))
        es = x.element_size()
        print(json.dumps({"case": "Cl(8,0,0) device operand * batch (lcc.torch)", "rows": n, "dtype": str(dt), "ms": ms, "GBps": 2 * 256 * es * n / ms / 1e6,
                          "dense_gemm_ms": ms_gemm}), flush=True)


if __name__ == "__main__":
    main()
