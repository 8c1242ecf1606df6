"""Build recipe for the CPU oracle (TEST INFRASTRUCTURE ONLY -- see lcc_oracle.c).

    python oracle/build.py          # -> oracle/_build/liblcc_oracle.so

The Rust reference never contracts a*b+c, so the restatement is compiled with
-ffp-contract=off and without -ffast-math.  There is no oracle/_ref: the reference is a
Rust/PyO3 crate and this image has no cargo/rustc/maturin nor the vendored crates, so it
cannot be compiled here ("Oracle").
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "liblcc_oracle.so")
SOURCES = [os.path.join(HERE, "lcc_oracle.c")]
DEPS = SOURCES + [os.path.join(HERE, "lcc_oracle_batch.inc")]


def build(force: bool = False) -> str:
    os.makedirs(OUT_DIR, exist_ok=True)
    if not force and os.path.exists(LIB):
        if all(os.path.getmtime(LIB) >= os.path.getmtime(d) for d in DEPS):
            return LIB
    cmd = [
        "gcc", "-O3", "-march=x86-64-v2", "-ffp-contract=off", "-fno-fast-math", "-fopenmp",
        "-fvisibility=hidden", "-shared", "-fPIC", "-std=c11", "-Wall", "-Wextra",
        "-o", LIB, *SOURCES, "-lm",
    ]
    subprocess.run(cmd, check=True, cwd=HERE)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
