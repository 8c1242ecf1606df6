"""PyTorch tensor interop of the B200-native engine (same public names as the reference's
lcc/torch/__init__.py:31-34: TorchMultivector, TorchAlgebra, TORCH_AVAILABLE).

    >>> from lcc.torch import TorchAlgebra, TorchMultivector
    >>> alg = TorchAlgebra.pga(3, device="cuda")
    >>> x = TorchMultivector(alg, torch.randn(1 << 20, 16, device="cuda", requires_grad=True))
    >>> (x.geometric_product(x)).coeffs.sum().backward()     # hand-written sm_100a kernels, autograd included
"""
try:
    import torch  # noqa: F401

    TORCH_AVAILABLE = True
except ImportError:  # pragma: no cover - torch is part of this image
    TORCH_AVAILABLE = False

if TORCH_AVAILABLE:
    from lcc.torch.multivector import TorchAlgebra, TorchMultivector, set_strict_fp

    __all__ = ["TorchMultivector", "TorchAlgebra", "TORCH_AVAILABLE", "set_strict_fp"]
else:  # pragma: no cover
    __all__ = ["TORCH_AVAILABLE"]

    class TorchMultivector:
        def __init__(self, *args, **kwargs):
            raise ImportError("PyTorch is not installed. Install it with: pip install torch")

    class TorchAlgebra:
        def __init__(self, *args, **kwargs):
            raise ImportError("PyTorch is not installed. Install it with: pip install torch")
