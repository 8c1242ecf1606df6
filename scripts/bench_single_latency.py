#!/usr/bin/env python3
"""Per-call latency of single `Multivector` operations through the drop-in extension (the reference quotes ~1 us per
product and ~2 us per sandwich for its Rust loops, benches/README.md:191-197).  One JSON line per (algebra, operation):
median and best of 5 repeats of 2000 calls, host wall clock.  `LCC_SINGLE_PAIR=device` re-measures the round-1 behaviour
(one-row batches through the GPU) for comparison."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np

import largecrimsoncanine as L


def timeit(fn, calls=2000, repeats=5):
    fn()
    best = []
    for _ in range(repeats):
        t0 = time.perf_counter()
        for _ in range(calls):
            fn()
        best.append((time.perf_counter() - t0) / calls)
    return float(np.median(best)) * 1e6, float(min(best)) * 1e6


rng = np.random.default_rng(3)
mode = os.environ.get("LCC_SINGLE_PAIR", "host")
for name, alg in (("R3 Cl(3,0,0)", L.Algebra.euclidean(3)), ("PGA Cl(3,0,1)", L.Algebra.pga(3)), ("CGA Cl(4,1,0)", L.Algebra.cga(3)), ("Cl(8,0,0)", L.Algebra(8))):
    nb = alg.num_blades
    a = L.Multivector.from_list_in(list(rng.standard_normal(nb)), alg)
    b = L.Multivector.from_list_in(list(rng.standard_normal(nb)), alg)
    biv = a.grade(2).scale(0.3) if hasattr(a, "grade") else a
    ops = {"geometric_product": lambda: a.geometric_product(b), "outer_product": lambda: a.outer_product(b),
           "left_contraction": lambda: a.left_contraction(b), "sandwich": lambda: a.sandwich(b), "reverse": lambda: a.reverse()}
    if nb <= 32:
        ops["exp(bivector)"] = lambda: biv.exp()
        small = a.scale(0.1)
        ops["exp(general multivector)"] = lambda: small.exp()
    calls = 200 if (mode == "device" or nb > 32) else 2000
    for op, fn in ops.items():
        try:
            med, best = timeit(fn, calls=calls)
        except Exception as exc:
            print(json.dumps({"mode": mode, "algebra": name, "op": op, "error": f"{type(exc).__name__}: {exc}"}), flush=True)
            continue
        print(json.dumps({"mode": mode, "algebra": name, "op": op, "us_per_call_median": round(med, 3), "us_per_call_best": round(best, 3), "calls": calls}), flush=True)
