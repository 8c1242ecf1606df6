"""lcc -- the reference's Python package name (lcc/__init__.py:14-25 of the reference): re-exports
the three classes of the `largecrimsoncanine` extension.  Here that extension is the B200-native
engine; `lcc.torch` is the tensor interop over the same C ABI.  The reference's `lcc.viz`
(`show`, `Graph`), `lcc.jax`, `lcc.symbolic` and `lcc.codegen` are outside the rebuilt path.
"""
from largecrimsoncanine import Algebra, Multivector, MultivectorBatch, __version__  # noqa: F401

__all__ = ["Algebra", "Multivector", "MultivectorBatch", "__version__"]
