"""lcc.torch -- tensor interop of the B200-native engine (drop-in for the reference's
lcc/torch/multivector.py: same class names, methods, argument meaning and error messages).

The reference implements every operation with stock PyTorch ops that materialise an
(N, num_blades, num_blades) product tensor and scatter-add it (lcc/torch/multivector.py:839-1019 of
the reference).  Here `TorchMultivector` borrows the caller's CUDA tensor zero-copy as an
array-of-structures view (`LCC_AOS`, the (N, num_blades) row-major layout the reference uses) and
runs the hand-written sm_100a kernels behind include/lcc_b200.h on torch's current stream.
PyTorch is only the owner of the memory and of the stream.

Autograd: the three products are bilinear, so their vector-Jacobian products are products again
under transposed sign tables (`lcc_batch_product_vjp`); they are wired in through
`torch.autograd.Function`.  Linear maps (reverse, grade, ...) differentiate through plain tensor ops.

There is no CPU path: tensors must live on a CUDA device (`.cuda()`); CPU tensors raise.
"""
from __future__ import annotations

import ctypes as C
import weakref
from typing import Optional, Tuple, Union

import numpy as np
import torch
from torch import Tensor

from largecrimsoncanine_b200 import cabi

_OPS = {"gp": cabi.LCC_OP_GP, "outer": cabi.LCC_OP_OUTER, "inner": cabi.LCC_OP_LC}
_HANDLES: dict = {}
_STRICT = False


def set_strict_fp(on: bool) -> None:
    """Reproduce the reference's rounding sequence bit for bit (no FMA) in lcc.torch calls."""
    global _STRICT
    _STRICT = bool(on)


def _flags() -> int:
    return cabi.LCC_FLAG_STRICT_FP if _STRICT else 0


def _handle(signature: Tuple[int, int, int]) -> cabi.Algebra:
    h = _HANDLES.get(signature)
    if h is None:
        h = _HANDLES[signature] = cabi.Algebra(*signature)
    return h


_ROTOR_CACHE: list = [None]  # (weakref to the rotor tensor, its version counter, host float64 copy)


def _rotor_on_host(t: Tensor) -> np.ndarray:
    """Host float64 copy of a one-row rotor tensor for the fixed-versor kernel (its coefficients travel as kernel
    parameters).  The device->host read is a synchronisation point, so the copy is cached per tensor object and
    version counter: applying the same motor to many batches syncs once, not once per call."""
    hit = _ROTOR_CACHE[0]
    if hit is not None and hit[0]() is t and hit[1] == t._version:
        return hit[2]
    r = t.detach().to(torch.float64).cpu().numpy().reshape(-1).copy()
    try:
        _ROTOR_CACHE[0] = (weakref.ref(t), t._version, r)
    except TypeError:
        _ROTOR_CACHE[0] = None
    return r


def _rotor_if_on_host(t: Tensor) -> Optional[np.ndarray]:
    """Host float64 copy of a one-row rotor tensor WITHOUT ever blocking: the first call with a given tensor (object and
    version counter) starts an asynchronous copy into pinned memory and returns None -- the caller then uses the
    device-rotor kernel; later calls return the copy once its event has completed."""
    hit = _ROTOR_CACHE[0]
    if hit is not None and hit[0]() is t and hit[1] == t._version:
        if isinstance(hit[2], np.ndarray):
            return hit[2]
        pinned, event = hit[2]
        if event.query():
            r = pinned.numpy().reshape(-1).copy()
            _ROTOR_CACHE[0] = (hit[0], hit[1], r)
            return r
        return None
    try:
        pinned = torch.empty(t.shape, dtype=torch.float64, pin_memory=True)
        with torch.cuda.device(t.device):
            pinned.copy_(t.detach().to(torch.float64), non_blocking=True)
            event = torch.cuda.Event()
            event.record()
        _ROTOR_CACHE[0] = (weakref.ref(t), t._version, (pinned, event))
    except (TypeError, RuntimeError):
        _ROTOR_CACHE[0] = None
    return None


def _blade_grade(blade: int) -> int:
    return bin(blade).count("1")


def _reverse_sign(grade: int) -> float:
    return 1.0 if (grade * (grade - 1) // 2) % 2 == 0 else -1.0


def _grade_involution_sign(grade: int) -> float:
    return 1.0 if grade % 2 == 0 else -1.0


def _dtype_code(t: Tensor) -> int:
    if t.dtype == torch.float32:
        return cabi.LCC_F32
    if t.dtype == torch.float64:
        return cabi.LCC_F64
    raise TypeError(f"lcc.torch supports float32 and float64 tensors, got {t.dtype}")


def _require_cuda(t: Tensor) -> None:
    if not t.is_cuda:
        raise RuntimeError(
            "lcc.torch runs on CUDA tensors only (B200 engine, no CPU fallback); move the data with .cuda()"
        )


def _view(t: Tensor) -> cabi.LccView:
    """Borrow a contiguous (N, num_blades) CUDA tensor as an LCC_AOS view (zero copy)."""
    return cabi.view(t.data_ptr(), t.shape[0], t.stride(0) if t.shape[0] > 1 else t.shape[1], _dtype_code(t), cabi.LCC_AOS)


def _stream() -> C.c_void_p:
    s = torch.cuda.current_stream().cuda_stream
    return C.c_void_p(s if s else 1)  # 0 is torch's legacy default stream == cudaStreamLegacy (0x1) in the C ABI


def _prep(t: Tensor) -> Tensor:
    _require_cuda(t)
    return t if t.is_contiguous() else t.contiguous()


def _broadcast_rows(na: int, nb: int) -> int:
    if na == nb or nb == 1:
        return na
    if na == 1:
        return nb
    raise ValueError(f"batch sizes must match or be 1 for broadcasting: {na} vs {nb}")


def _product_forward(sig, op: int, a: Tensor, b: Tensor) -> Tensor:
    a, b = _prep(a), _prep(b)
    if a.dtype != b.dtype:
        b = b.to(a.dtype)
    n = _broadcast_rows(a.shape[0], b.shape[0])
    out = torch.empty((n, a.shape[1]), dtype=a.dtype, device=a.device)
    if n:
        with torch.cuda.device(a.device):
            cabi.check(cabi.lib().lcc_batch_product(_handle(sig).h, op, C.byref(_view(a)), C.byref(_view(b)), C.byref(_view(out)), _flags(), _stream()))
    return out


def _product_vjp(sig, op: int, wrt: int, grad_out: Tensor, other: Tensor) -> Tensor:
    grad_out, other = _prep(grad_out), _prep(other)
    out = torch.empty_like(grad_out)
    if grad_out.shape[0]:
        with torch.cuda.device(grad_out.device):
            cabi.check(cabi.lib().lcc_batch_product_vjp(_handle(sig).h, op, wrt, C.byref(_view(grad_out)), C.byref(_view(other)), C.byref(_view(out)), _flags(), _stream()))
    return out


class _Product(torch.autograd.Function):
    """out = a (op) b row by row, with the reference's one-row broadcast."""

    @staticmethod
    def forward(ctx, a: Tensor, b: Tensor, sig, op: int):
        ctx.sig, ctx.op = sig, op
        ctx.b_dtype = b.dtype
        if b.dtype != a.dtype:  # the product runs in a's dtype; save what the kernels actually saw
            b = b.to(a.dtype)
        ctx.save_for_backward(a, b)
        return _product_forward(sig, op, a, b)

    @staticmethod
    def backward(ctx, grad_out: Tensor):
        a, b = ctx.saved_tensors
        ga = gb = None
        if ctx.needs_input_grad[0]:
            ga = _product_vjp(ctx.sig, ctx.op, 0, grad_out, b)
            if a.shape[0] == 1 and ga.shape[0] > 1:
                ga = ga.sum(dim=0, keepdim=True)
        if ctx.needs_input_grad[1]:
            gb = _product_vjp(ctx.sig, ctx.op, 1, grad_out, a)
            if b.shape[0] == 1 and gb.shape[0] > 1:
                gb = gb.sum(dim=0, keepdim=True)
            if gb.dtype != ctx.b_dtype:
                gb = gb.to(ctx.b_dtype)
        return ga, gb, None, None


class _SandwichFixed(torch.autograd.Function):
    """out = R x ~R with ONE rotor that needs no gradient: forward is the fused sandwich kernel (one pass over x instead of
    two products and a full-size intermediate); x -> R x ~R is linear, so the cotangent of x is two vector-Jacobian
    passes with broadcast operands -- g_rx = vjp_left(g, ~R), g_x = vjp_right(g_rx, R) -- and nothing has to be saved
    but the rotor."""

    @staticmethod
    def forward(ctx, x: Tensor, r: Tensor, sig):
        ctx.sig = sig
        x, r = _prep(x), _prep(r.detach().to(x.dtype))
        ctx.save_for_backward(r)
        out = torch.empty_like(x)
        if x.shape[0]:
            host = _rotor_if_on_host(r)
            with torch.cuda.device(x.device):
                if host is not None:
                    cabi.check(cabi.lib().lcc_batch_sandwich(_handle(sig).h, host.ctypes.data_as(C.POINTER(C.c_double)),
                                                             C.byref(_view(x)), C.byref(_view(out)), _flags(), _stream()))
                else:
                    cabi.check(cabi.lib().lcc_batch_sandwich_rows(_handle(sig).h, C.byref(_view(r)), C.byref(_view(x)),
                                                                  C.byref(_view(out)), _flags(), _stream()))
        return out

    @staticmethod
    def backward(ctx, grad_out: Tensor):
        (r,) = ctx.saved_tensors
        nb = r.shape[1]
        rev = torch.tensor([_reverse_sign(_blade_grade(i)) for i in range(nb)], dtype=r.dtype, device=r.device)
        g_rx = _product_vjp(ctx.sig, cabi.LCC_OP_GP, 0, grad_out, r * rev)
        return _product_vjp(ctx.sig, cabi.LCC_OP_GP, 1, g_rx, r), None, None


class _SandwichRows(torch.autograd.Function):
    """out = R x ~R with gradients to x and / or R (one rotor per row, or one rotor for all rows): forward is the fused
    per-row kernel (K2r), backward ONE fused cotangent kernel (lcc_batch_sandwich_rows_vjp: R, x and the incoming
    cotangent in registers, 5 * num_blades values of traffic per row) instead of autograd through two products, a
    reverse and a saved intermediate (reference: lcc/torch/multivector.py:988-1019)."""

    @staticmethod
    def forward(ctx, x: Tensor, r: Tensor, sig):
        ctx.sig, ctx.r_dtype = sig, r.dtype
        x = _prep(x)
        r = _prep(r if r.dtype == x.dtype else r.to(x.dtype))
        ctx.save_for_backward(x, r)
        out = torch.empty_like(x)
        if x.shape[0]:
            with torch.cuda.device(x.device):
                cabi.check(cabi.lib().lcc_batch_sandwich_rows(_handle(sig).h, C.byref(_view(r)), C.byref(_view(x)),
                                                              C.byref(_view(out)), _flags(), _stream()))
        return out

    @staticmethod
    def backward(ctx, grad_out: Tensor):
        x, r = ctx.saved_tensors
        grad_out = _prep(grad_out)
        gx, gr = torch.empty_like(x), torch.empty_like(x)
        if x.shape[0]:
            with torch.cuda.device(x.device):
                cabi.check(cabi.lib().lcc_batch_sandwich_rows_vjp(_handle(ctx.sig).h, C.byref(_view(r)), C.byref(_view(x)), C.byref(_view(grad_out)),
                                                                  C.byref(_view(gx)), C.byref(_view(gr)), _flags(), _stream()))
        else:
            gr = torch.zeros_like(r)
        if r.shape[0] == 1 and gr.shape[0] != 1:
            gr = gr.sum(dim=0, keepdim=True)
        if gr.dtype != ctx.r_dtype:
            gr = gr.to(ctx.r_dtype)
        return (gx if ctx.needs_input_grad[0] else None), (gr if ctx.needs_input_grad[1] else None), None


class TorchAlgebra:
    """An algebra's Cayley tables as tensors plus the handle of the native engine.

    Mirrors the reference dataclass (lcc/torch/multivector.py:24-220): `signature`, `num_blades`,
    `dimension`, `cayley_signs` (num_blades x num_blades, float64), `cayley_blades` (long), `device`.
    """

    def __init__(self, signature: Tuple[int, int, int], device: Union[str, torch.device] = "cpu"):
        self.signature = tuple(int(v) for v in signature)
        h = _handle(self.signature)
        self.num_blades = h.num_blades
        self.dimension = h.dimension
        self.device = torch.device(device)
        nb = self.num_blades
        self.cayley_signs = torch.tensor(h.cayley_signs().reshape(nb, nb).astype(np.float64), device=self.device)
        self.cayley_blades = torch.tensor(h.cayley_blades().reshape(nb, nb).astype(np.int64), device=self.device)

    def to(self, device: Union[str, torch.device]) -> "TorchAlgebra":
        return TorchAlgebra(self.signature, device)

    def cuda(self, device: Optional[int] = None) -> "TorchAlgebra":
        return self.to("cuda" if device is None else f"cuda:{device}")

    def cpu(self) -> "TorchAlgebra":
        return self.to("cpu")

    @classmethod
    def from_rust_algebra(cls, algebra, device: Union[str, torch.device] = "cpu") -> "TorchAlgebra":
        """From a `largecrimsoncanine.Algebra` (works with its p/q/r getters; the reference's version
        calls methods the PyO3 class does not have)."""
        return cls((algebra.p, algebra.q, algebra.r), device)

    @classmethod
    def euclidean(cls, n: int, device: Union[str, torch.device] = "cpu") -> "TorchAlgebra":
        return cls((n, 0, 0), device)

    @classmethod
    def pga(cls, n: int, device: Union[str, torch.device] = "cpu") -> "TorchAlgebra":
        return cls((n, 0, 1), device)

    @classmethod
    def cga(cls, n: int, device: Union[str, torch.device] = "cpu") -> "TorchAlgebra":
        return cls((n + 1, 1, 0), device)

    @classmethod
    def sta(cls, device: Union[str, torch.device] = "cpu") -> "TorchAlgebra":
        return cls((1, 3, 0), device)

    @classmethod
    def _from_signature(cls, p: int, q: int, r: int, device: Union[str, torch.device] = "cpu") -> "TorchAlgebra":
        return cls((p, q, r), device)

    def __repr__(self) -> str:
        return f"TorchAlgebra(signature={self.signature}, device='{self.device}')"


class TorchMultivector:
    """A batch of multivectors backed by a (batch_size, num_blades) tensor
    (reference: lcc/torch/multivector.py:305-832)."""

    def __init__(self, algebra: TorchAlgebra, coeffs: Tensor):
        self.algebra = algebra
        if coeffs.ndim == 1:
            coeffs = coeffs.unsqueeze(0)
        if coeffs.shape[-1] != algebra.num_blades:
            raise ValueError(
                f"coeffs must have {algebra.num_blades} columns for "
                f"algebra with signature {algebra.signature}, got {coeffs.shape[-1]}"
            )
        if coeffs.device != algebra.device:
            coeffs = coeffs.to(algebra.device)
        self.coeffs = coeffs

    # ------------------------------------------------------------------ properties / movement
    @property
    def batch_size(self) -> int:
        return self.coeffs.shape[0]

    @property
    def num_blades(self) -> int:
        return self.algebra.num_blades

    @property
    def shape(self) -> Tuple[int, int]:
        return tuple(self.coeffs.shape)

    @property
    def device(self) -> torch.device:
        return self.coeffs.device

    @property
    def requires_grad(self) -> bool:
        return self.coeffs.requires_grad

    def to(self, device: Union[str, torch.device]) -> "TorchMultivector":
        device = torch.device(device)
        return TorchMultivector(self.algebra.to(device), self.coeffs.to(device))

    def cuda(self, device: Optional[int] = None) -> "TorchMultivector":
        return self.to("cuda" if device is None else f"cuda:{device}")

    def cpu(self) -> "TorchMultivector":
        return self.to("cpu")

    def detach(self) -> "TorchMultivector":
        return TorchMultivector(self.algebra, self.coeffs.detach())

    def requires_grad_(self, requires_grad: bool = True) -> "TorchMultivector":
        self.coeffs.requires_grad_(requires_grad)
        return self

    # ------------------------------------------------------------------ construction
    @classmethod
    def from_numpy(cls, algebra: TorchAlgebra, coeffs: np.ndarray) -> "TorchMultivector":
        return cls(algebra, torch.tensor(coeffs, dtype=torch.float64, device=algebra.device))

    def to_numpy(self) -> np.ndarray:
        return self.coeffs.detach().cpu().numpy()

    @classmethod
    def from_vectors(cls, algebra: TorchAlgebra, coords: Tensor) -> "TorchMultivector":
        if coords.ndim == 1:
            coords = coords.unsqueeze(0)
        if coords.shape[1] != algebra.dimension:
            raise ValueError(
                f"coords must have {algebra.dimension} columns for "
                f"algebra with signature {algebra.signature}, got {coords.shape[1]}"
            )
        coeffs = torch.zeros(coords.shape[0], algebra.num_blades, dtype=coords.dtype, device=algebra.device)
        idx = torch.tensor([1 << i for i in range(algebra.dimension)], device=algebra.device)
        coeffs[:, idx] = coords.to(algebra.device)
        return cls(algebra, coeffs)

    @classmethod
    def from_scalars(cls, algebra: TorchAlgebra, values: Tensor) -> "TorchMultivector":
        if values.ndim == 0:
            values = values.unsqueeze(0)
        coeffs = torch.zeros(values.shape[0], algebra.num_blades, dtype=values.dtype, device=algebra.device)
        coeffs[:, 0] = values.to(algebra.device)
        return cls(algebra, coeffs)

    @classmethod
    def from_bivectors(cls, algebra: TorchAlgebra, biv_coeffs: Tensor) -> "TorchMultivector":
        if biv_coeffs.ndim == 1:
            biv_coeffs = biv_coeffs.unsqueeze(0)
        dim = algebra.dimension
        num_bivectors = dim * (dim - 1) // 2
        if biv_coeffs.shape[1] != num_bivectors:
            raise ValueError(
                f"biv_coeffs must have {num_bivectors} columns for "
                f"algebra with signature {algebra.signature}, got {biv_coeffs.shape[1]}"
            )
        idx = torch.tensor([i for i in range(algebra.num_blades) if _blade_grade(i) == 2], device=algebra.device, dtype=torch.long)
        coeffs = torch.zeros(biv_coeffs.shape[0], algebra.num_blades, dtype=biv_coeffs.dtype, device=algebra.device)
        coeffs[:, idx] = biv_coeffs.to(algebra.device)
        return cls(algebra, coeffs)

    def to_vector_coords(self) -> Tensor:
        idx = torch.tensor([1 << i for i in range(self.algebra.dimension)], device=self.device)
        return self.coeffs[:, idx]

    def scalar(self) -> Tensor:
        return self.coeffs[:, 0]

    # ------------------------------------------------------------------ products (K1 kernels + autograd)
    def _product(self, other: "TorchMultivector", name: str) -> "TorchMultivector":
        if tuple(other.algebra.signature) != tuple(self.algebra.signature):
            raise ValueError(f"algebra mismatch: left is {self.algebra.signature} but right is {other.algebra.signature}")
        result = _Product.apply(self.coeffs, other.coeffs, self.algebra.signature, _OPS[name])
        return TorchMultivector(self.algebra, result)

    def geometric_product(self, other: "TorchMultivector") -> "TorchMultivector":
        return self._product(other, "gp")

    gp = geometric_product

    def outer_product(self, other: "TorchMultivector") -> "TorchMultivector":
        return self._product(other, "outer")

    wedge = outer_product

    def inner_product(self, other: "TorchMultivector") -> "TorchMultivector":
        return self._product(other, "inner")

    lc = inner_product

    def sandwich(self, rotor: "TorchMultivector") -> "TorchMultivector":
        """rotor * self * ~rotor (reference: lcc/torch/multivector.py:635-659, 988-1019).
        One rotor and no gradients -> the fused fixed-versor kernel (K2, one pass over HBM); gradients to x only -> the
        same forward with two broadcast cotangent passes; gradients to the rotor(s) -> the fused per-row kernel with ONE
        fused cotangent kernel; what is left (a one-row x against many rotors) composes two differentiable products as
        the reference does."""
        needs_grad = torch.is_grad_enabled() and (self.coeffs.requires_grad or rotor.coeffs.requires_grad)
        if rotor.batch_size == 1 and not needs_grad:
            x = _prep(self.coeffs)
            out = torch.empty_like(x)
            if x.shape[0]:
                r = _rotor_if_on_host(rotor.coeffs)
                with torch.cuda.device(x.device):
                    if r is not None:  # host copy known: the fixed-versor kernel (rotor in the constant bank, even-versor pruning)
                        cabi.check(cabi.lib().lcc_batch_sandwich(_handle(self.algebra.signature).h, r.ctypes.data_as(C.POINTER(C.c_double)),
                                                                 C.byref(_view(x)), C.byref(_view(out)), _flags(), _stream()))
                    else:  # first sight of this rotor tensor: read it on the DEVICE (one broadcast row), no host synchronisation
                        rd = _prep(rotor.coeffs.detach().to(x.dtype))
                        cabi.check(cabi.lib().lcc_batch_sandwich_rows(_handle(self.algebra.signature).h, C.byref(_view(rd)), C.byref(_view(x)),
                                                                      C.byref(_view(out)), _flags(), _stream()))
            return TorchMultivector(self.algebra, out)
        if rotor.batch_size == 1 and not (torch.is_grad_enabled() and rotor.coeffs.requires_grad):
            # gradients flow to x only: fused forward, two broadcast VJP passes backward
            return TorchMultivector(self.algebra, _SandwichFixed.apply(self.coeffs, rotor.coeffs, self.algebra.signature))
        if not needs_grad and rotor.batch_size == self.batch_size and self.batch_size > 0:
            # one rotor per row (reference :988-1019): lcc_batch_sandwich_rows on the borrowed tensors
            x, r = _prep(self.coeffs), _prep(rotor.coeffs.to(self.coeffs.dtype))
            out = torch.empty_like(x)
            with torch.cuda.device(x.device):
                cabi.check(cabi.lib().lcc_batch_sandwich_rows(_handle(self.algebra.signature).h, C.byref(_view(r)), C.byref(_view(x)),
                                                              C.byref(_view(out)), _flags(), _stream()))
            return TorchMultivector(self.algebra, out)
        if needs_grad and self.batch_size > 0 and rotor.batch_size in (1, self.batch_size):
            # gradients to the rotor(s) too: fused forward, ONE fused cotangent kernel backward
            return TorchMultivector(self.algebra, _SandwichRows.apply(self.coeffs, rotor.coeffs, self.algebra.signature))
        return rotor.geometric_product(self).geometric_product(rotor.reverse())

    apply = sandwich

    # ------------------------------------------------------------------ involutions / grades (linear maps)
    def _scaled(self, factors) -> "TorchMultivector":
        f = torch.tensor(factors, dtype=self.coeffs.dtype, device=self.device)
        return TorchMultivector(self.algebra, self.coeffs * f)

    def reverse(self) -> "TorchMultivector":
        return self._scaled([_reverse_sign(_blade_grade(i)) for i in range(self.num_blades)])

    def grade_involution(self) -> "TorchMultivector":
        return self._scaled([_grade_involution_sign(_blade_grade(i)) for i in range(self.num_blades)])

    def conjugate(self) -> "TorchMultivector":
        return self.reverse().grade_involution()

    def grade(self, k: int) -> "TorchMultivector":
        return self._scaled([1.0 if _blade_grade(i) == k else 0.0 for i in range(self.num_blades)])

    def even(self) -> "TorchMultivector":
        return self._scaled([1.0 if _blade_grade(i) % 2 == 0 else 0.0 for i in range(self.num_blades)])

    def odd(self) -> "TorchMultivector":
        return self._scaled([1.0 if _blade_grade(i) % 2 == 1 else 0.0 for i in range(self.num_blades)])

    # ------------------------------------------------------------------ norms
    def norm_squared(self) -> Tensor:
        """Scalar part of self * ~self (reference :711-721).  Without gradients: the K4 kernel, one read
        of the batch; with gradients: the differentiable product composition."""
        if torch.is_grad_enabled() and self.coeffs.requires_grad:
            return self.geometric_product(self.reverse()).scalar()
        x = _prep(self.coeffs)
        out = torch.empty(x.shape[0], dtype=x.dtype, device=x.device)
        if x.shape[0]:
            with torch.cuda.device(x.device):
                cabi.check(cabi.lib().lcc_batch_norm(_handle(self.algebra.signature).h, cabi.LCC_NORM_SQUARED, C.byref(_view(x)),
                                                     C.c_void_p(out.data_ptr()), _flags(), _stream()))
        return out

    def norm(self) -> Tensor:
        return torch.sqrt(torch.abs(self.norm_squared()))

    def normalized(self) -> "TorchMultivector":
        norms = self.norm()
        safe = torch.where(norms > 1e-14, norms, torch.ones_like(norms))
        return TorchMultivector(self.algebra, self.coeffs / safe.unsqueeze(-1))

    # ------------------------------------------------------------------ arithmetic / indexing
    def __add__(self, other: "TorchMultivector") -> "TorchMultivector":
        return TorchMultivector(self.algebra, self.coeffs + other.coeffs)

    def __sub__(self, other: "TorchMultivector") -> "TorchMultivector":
        return TorchMultivector(self.algebra, self.coeffs - other.coeffs)

    def __neg__(self) -> "TorchMultivector":
        return TorchMultivector(self.algebra, -self.coeffs)

    def __mul__(self, scalar: Union[float, Tensor]) -> "TorchMultivector":
        return TorchMultivector(self.algebra, self.coeffs * scalar)

    def __rmul__(self, scalar: Union[float, Tensor]) -> "TorchMultivector":
        return self.__mul__(scalar)

    def __truediv__(self, scalar: Union[float, Tensor]) -> "TorchMultivector":
        return TorchMultivector(self.algebra, self.coeffs / scalar)

    def __getitem__(self, idx) -> "TorchMultivector":
        result = self.coeffs[idx]
        if result.ndim == 1:
            result = result.unsqueeze(0)
        return TorchMultivector(self.algebra, result)

    def __len__(self) -> int:
        return self.batch_size

    def sum(self) -> Tensor:
        """Column sums over the batch (K5, fixed reduction tree); new surface."""
        if torch.is_grad_enabled() and self.coeffs.requires_grad:
            return self.coeffs.sum(dim=0)
        x = _prep(self.coeffs)
        out = torch.zeros(self.num_blades, dtype=x.dtype, device=x.device)
        with torch.cuda.device(x.device):
            cabi.check(cabi.lib().lcc_batch_sum(_handle(self.algebra.signature).h, C.byref(_view(x)), C.c_void_p(out.data_ptr()), _stream()))
        return out

    def __repr__(self) -> str:
        device_str = f", device='{self.device}'" if self.device.type != "cpu" else ""
        grad_str = ", requires_grad=True" if self.requires_grad else ""
        return f"TorchMultivector(batch_size={self.batch_size}, signature={self.algebra.signature}{device_str}{grad_str})"
