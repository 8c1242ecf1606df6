"""largecrimsoncanine_b200 -- B200-native engine for the batched Clifford-algebra path of
LargeCrimsonCanine (src/batch.rs + src/simd.rs and the table machinery they read).

Layout
    csrc/            hand-written sm_100a kernels, the build-time Cayley generator, the C ABI
    lib/             liblcc_b200.so (built in-tree by ``python -m largecrimsoncanine_b200._build``)
    cabi.py          ctypes declaration of include/lcc_b200.h
    sharded.py       one-process-per-GPU sharding + the NCCL batch sum / mean
The drop-in names live at the repo root: the ``largecrimsoncanine`` extension (Algebra,
Multivector, MultivectorBatch) and the ``lcc`` package (``lcc.torch``).

There is no CPU fallback: importing works without a GPU (the algebra tables are host side), every
batched operation raises when no CUDA device is usable.
"""
from . import cabi  # noqa: F401

__all__ = ["cabi", "Algebra", "Multivector", "MultivectorBatch", "native_libraries"]


def native_libraries():
    """Paths of the native pieces this package runs on (all in-tree)."""
    import largecrimsoncanine as _ext

    return {"c_abi": cabi.LIB_PATH, "extension": _ext.__file__}


def __getattr__(name):
    if name in ("Algebra", "Multivector", "MultivectorBatch"):
        import largecrimsoncanine as _ext  # the CPython extension at the repo root

        return getattr(_ext, name)
    raise AttributeError(name)
