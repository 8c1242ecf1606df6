"""Generate tests/golden/ref_torch_golden.npz from the REFERENCE's own Python implementation.

Run in the build container only (it reads /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

The reference's Rust crate cannot be built here, but its pure-PyTorch twin of the same math,
/root/reference/lcc/torch/multivector.py, imports by file path and runs on
the CPU.  Its outputs on seeded inputs are committed as golden vectors; the oracle
(oracle/lcc_oracle.c) must reproduce them bit-for-bit in f64, and the CUDA path is then checked
against both.  Two further independent table builders of the reference are captured as well:
TorchAlgebra._from_signature (lcc/torch/multivector.py:183-286) and the pure-Python
lcc/codegen Signature/blade_product (lcc/codegen/__init__.py:29-257).
"""
from __future__ import annotations

import importlib.util
import os
import sys

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_torch_golden.npz")

SIGNATURES = [(2, 0, 0), (3, 0, 0), (3, 0, 1), (4, 1, 0), (1, 3, 0), (2, 1, 1), (6, 0, 0)]
ROWS = 12


def load_ref_torch():
    spec = importlib.util.spec_from_file_location("ref_torch_mv", f"{REF}/lcc/torch/multivector.py")
    mod = importlib.util.module_from_spec(spec)
    sys.modules["ref_torch_mv"] = mod
    spec.loader.exec_module(mod)
    return mod


def load_ref_codegen_tables():
    """exec the table-building head of lcc/codegen/__init__.py (its tail imports missing modules)."""
    src = open(f"{REF}/lcc/codegen/__init__.py").read().splitlines()
    head = []
    for line in src:
        if line.startswith("from .") or line.startswith("from lcc"):
            break
        head.append(line)
    import types

    mod = types.ModuleType("ref_codegen_head")
    sys.modules["ref_codegen_head"] = mod  # dataclasses looks the module up by name
    exec(compile("\n".join(head), "ref_codegen_head", "exec"), mod.__dict__)
    return mod.__dict__


def main():
    torch.set_num_threads(1)
    m = load_ref_torch()
    out = {}
    rng = np.random.default_rng(42)
    for sig in SIGNATURES:
        tag = "cl%d%d%d" % sig
        alg = m.TorchAlgebra._from_signature(*sig)
        nb = alg.num_blades
        out[f"{tag}.signs"] = alg.cayley_signs.numpy().astype(np.int8)
        out[f"{tag}.blades"] = alg.cayley_blades.numpy().astype(np.uint16)
        if nb > 32:
            continue  # tables only for the large algebra
        for dt, tdt in (("f64", torch.float64), ("f32", torch.float32)):
            a = rng.standard_normal((ROWS, nb))
            b = rng.standard_normal((ROWS, nb))
            rotor = rng.standard_normal((1, nb))
            if dt == "f32":
                a, b = a.astype(np.float32), b.astype(np.float32)
            A = m.TorchMultivector(alg, torch.tensor(a, dtype=tdt))
            B = m.TorchMultivector(alg, torch.tensor(b, dtype=tdt))
            R = m.TorchMultivector(alg, torch.tensor(rotor, dtype=tdt))
            k = f"{tag}.{dt}"
            out[f"{k}.a"], out[f"{k}.b"], out[f"{k}.rotor"] = a, b, rotor
            out[f"{k}.gp"] = A.geometric_product(B).coeffs.numpy()
            out[f"{k}.gp_bcast"] = A.geometric_product(m.TorchMultivector(alg, B.coeffs[:1])).coeffs.numpy()
            out[f"{k}.outer"] = A.outer_product(B).coeffs.numpy()
            out[f"{k}.inner"] = A.inner_product(B).coeffs.numpy()
            out[f"{k}.sandwich"] = A.sandwich(R).coeffs.numpy()
            out[f"{k}.reverse"] = A.reverse().coeffs.numpy()
            out[f"{k}.involution"] = A.grade_involution().coeffs.numpy()
            out[f"{k}.conjugate"] = A.conjugate().coeffs.numpy()
            out[f"{k}.norm_squared"] = A.norm_squared().numpy()
            out[f"{k}.norm"] = A.norm().numpy()
            out[f"{k}.normalized"] = A.normalized().coeffs.numpy()
            for g in range(alg.dimension + 1):
                out[f"{k}.grade{g}"] = A.grade(g).coeffs.numpy()
    # third table builder: lcc/codegen
    ns = load_ref_codegen_tables()
    try:
        for sig in SIGNATURES:
            if sum(sig) > 5:
                continue
            s = ns["Signature"](*sig)
            nb = 1 << sum(sig)
            signs = np.zeros((nb, nb), np.int8)
            blades = np.zeros((nb, nb), np.uint16)
            for x in range(nb):
                for y in range(nb):
                    bl, sg = ns["blade_product"](x, y, s)
                    blades[x, y], signs[x, y] = bl, int(sg)
            out["cl%d%d%d.codegen_signs" % sig] = signs
            out["cl%d%d%d.codegen_blades" % sig] = blades
    except Exception as exc:  # the codegen head is best-effort; say so rather than hide it
        print("codegen table builder not captured:", repr(exc))
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
