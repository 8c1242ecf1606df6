"""ctypes front end of oracle/lcc_oracle.c -- TEST INFRASTRUCTURE ONLY.

Every function takes and returns NumPy arrays in the reference's layout: array-of-structures,
row-major ``(n, num_blades)``, float64 (the reference's only dtype) or float32 ("same loops in
float").  Citations are in the C file; this module adds nothing numerical.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build as _build

__all__ = [
    "OracleAlgebra", "basis_square", "reorder_sign", "blade_product", "blade_grade", "reverse_sign",
    "grade_involution_sign", "clifford_conjugate_sign", "grade_mask", "max_threads", "lib_path",
    "sta_bivectors", "sta_boosts", "pga_points_embed", "pga_points_extract", "pga_transform_points",
]

_LIB = None


def lib_path() -> str:
    return _build.LIB


def _lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    path = _build.LIB
    if not os.path.exists(path) or any(
        os.path.exists(d) and os.path.getmtime(d) > os.path.getmtime(path) for d in _build.DEPS
    ):
        path = _build.build()
    lib = C.CDLL(path)
    i64, u64, dbl, vp, ci = C.c_int64, C.c_uint64, C.c_double, C.c_void_p, C.c_int
    lib.oracle_basis_square.restype = dbl
    lib.oracle_basis_square.argtypes = [ci, ci, ci, ci]
    lib.oracle_reorder_sign.restype = dbl
    lib.oracle_reorder_sign.argtypes = [u64, u64]
    lib.oracle_blade_product.restype = None
    lib.oracle_blade_product.argtypes = [u64, u64, ci, ci, ci, C.POINTER(u64), C.POINTER(dbl)]
    lib.oracle_blade_grade.restype = ci
    lib.oracle_blade_grade.argtypes = [u64]
    for name in ("oracle_reverse_sign", "oracle_grade_involution_sign", "oracle_clifford_conjugate_sign"):
        getattr(lib, name).restype = dbl
        getattr(lib, name).argtypes = [ci]
    lib.oracle_grade_mask.restype = u64
    lib.oracle_grade_mask.argtypes = [ci, ci, C.POINTER(ci)]
    lib.oracle_algebra_new.restype = vp
    lib.oracle_algebra_new.argtypes = [ci, ci, ci]
    lib.oracle_algebra_free.restype = None
    lib.oracle_algebra_free.argtypes = [vp]
    lib.oracle_algebra_num_blades.restype = i64
    lib.oracle_algebra_num_blades.argtypes = [vp]
    lib.oracle_algebra_blades.restype = C.POINTER(u64)
    lib.oracle_algebra_blades.argtypes = [vp]
    lib.oracle_algebra_signs.restype = C.POINTER(dbl)
    lib.oracle_algebra_signs.argtypes = [vp]
    for sfx in ("f64", "f32"):
        f = getattr(lib, f"oracle_batch_product_{sfx}")
        f.restype = i64
        f.argtypes = [vp, ci, ci, vp, i64, vp, i64, vp, ci]
        f = getattr(lib, f"oracle_batch_sandwich_{sfx}")
        f.restype = None
        f.argtypes = [vp, vp, vp, i64, vp, ci]
        f = getattr(lib, f"oracle_batch_norm_{sfx}")
        f.restype = None
        f.argtypes = [vp, ci, vp, i64, vp]
        f = getattr(lib, f"oracle_batch_normalized_{sfx}")
        f.restype = None
        f.argtypes = [vp, vp, i64, vp]
        f = getattr(lib, f"oracle_batch_involution_{sfx}")
        f.restype = None
        f.argtypes = [vp, ci, vp, i64, vp]
        f = getattr(lib, f"oracle_batch_grade_{sfx}")
        f.restype = ci
        f.argtypes = [vp, ci, vp, i64, vp]
        f = getattr(lib, f"oracle_batch_addsub_{sfx}")
        f.restype = i64
        f.argtypes = [vp, ci, vp, i64, vp, i64, vp]
        f = getattr(lib, f"oracle_batch_scale_{sfx}")
        f.restype = None
        f.argtypes = [vp, vp, i64, dbl, vp]
        for nm in ("oracle_from_vectors", "oracle_to_vector_coords", "oracle_from_bivectors"):
            f = getattr(lib, f"{nm}_{sfx}")
            f.restype = None
            f.argtypes = [vp, vp, i64, vp]
        f = getattr(lib, f"oracle_sum_rows_{sfx}")
        f.restype = None
        f.argtypes = [vp, i64, i64, vp, vp]
    lib.oracle_simd_geometric_product.restype = None
    lib.oracle_simd_geometric_product.argtypes = [vp, vp, vp, vp]
    lib.oracle_dual_rows.restype = ci
    lib.oracle_dual_rows.argtypes = [vp, ci, vp, i64, vp]
    lib.oracle_pga_dual_rows.restype = None
    lib.oracle_pga_dual_rows.argtypes = [vp, i64, vp]
    lib.oracle_max_threads.restype = ci
    lib.oracle_max_threads.argtypes = []
    _LIB = lib
    return lib


# ----------------------------------------------------------------------------- integer primitives

def basis_square(p, q, r, i):
    return _lib().oracle_basis_square(p, q, r, i)


def reorder_sign(a, b):
    return _lib().oracle_reorder_sign(a, b)


def blade_product(a, b, p, q=0, r=0):
    blade, sign = C.c_uint64(), C.c_double()
    _lib().oracle_blade_product(a, b, p, q, r, C.byref(blade), C.byref(sign))
    return int(blade.value), float(sign.value)


def blade_grade(i):
    return _lib().oracle_blade_grade(i)


def reverse_sign(g):
    return _lib().oracle_reverse_sign(g)


def grade_involution_sign(g):
    return _lib().oracle_grade_involution_sign(g)


def clifford_conjugate_sign(g):
    return _lib().oracle_clifford_conjugate_sign(g)


def grade_mask(dimension, grade):
    """One-word mask as in blades.rs:211-220; raises for dimension > 6 (undefined in the reference)."""
    ok = C.c_int()
    m = _lib().oracle_grade_mask(dimension, grade, C.byref(ok))
    if not ok.value:
        raise OverflowError("grade_mask is undefined in the reference for dimension > 6")
    return int(m)


def max_threads():
    return _lib().oracle_max_threads()


# ----------------------------------------------------------------------------- algebra + batched loops

def _sfx(dtype):
    dtype = np.dtype(dtype)
    if dtype == np.float64:
        return "f64"
    if dtype == np.float32:
        return "f32"
    raise TypeError(f"oracle supports float64/float32, got {dtype}")


def _arr(x, dtype=None, ndim=2):
    x = np.ascontiguousarray(x, dtype=dtype)
    if x.ndim != ndim:
        raise ValueError(f"expected a {ndim}-D array, got shape {x.shape}")
    return x


class OracleAlgebra:
    """Cl(p,q,r) with the reference's flat Cayley tables (cayley.rs:123-144)."""

    def __init__(self, p, q=0, r=0):
        self.p, self.q, self.r = int(p), int(q), int(r)
        self.dimension = self.p + self.q + self.r
        self.num_blades = 1 << self.dimension
        self._h = _lib().oracle_algebra_new(self.p, self.q, self.r)
        nb = self.num_blades
        self.cayley_blades = np.ctypeslib.as_array(_lib().oracle_algebra_blades(self._h), shape=(nb * nb,)).copy()
        self.cayley_signs = np.ctypeslib.as_array(_lib().oracle_algebra_signs(self._h), shape=(nb * nb,)).copy()

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h and _LIB is not None:
            _LIB.oracle_algebra_free(h)

    @property
    def signature(self):
        return (self.p, self.q, self.r)

    # -- tables in the digest form used by SURVEY.md section 8c
    def signs_int8(self):
        return self.cayley_signs.astype(np.int8)

    def blades_u16(self):
        return self.cayley_blades.astype("<u2")

    def grade_masks(self):
        return [grade_mask(self.dimension, g) for g in range(self.dimension + 1)]

    # -- products
    def _product(self, op, a, b, absmode=False, threads=1):
        a = _arr(a)
        b = _arr(b, dtype=a.dtype)
        nb = self.num_blades
        if a.shape[1] != nb or b.shape[1] != nb:
            raise ValueError(f"coeffs must have {nb} columns")
        na, nbr = a.shape[0], b.shape[0]
        n = na if na == nbr or nbr == 1 else (nbr if na == 1 else -1)
        if n < 0:
            raise ValueError(f"batch sizes must match or be 1 for broadcasting: {na} vs {nbr}")
        out = np.empty((n, nb), dtype=a.dtype)
        f = getattr(_lib(), f"oracle_batch_product_{_sfx(a.dtype)}")
        got = f(self._h, op, int(absmode), a.ctypes.data, na, b.ctypes.data, nbr, out.ctypes.data, threads)
        assert got == n
        return out

    def geometric_product(self, a, b, **kw):
        return self._product(0, a, b, **kw)

    def outer_product(self, a, b, **kw):
        return self._product(1, a, b, **kw)

    def left_contraction(self, a, b, **kw):
        return self._product(2, a, b, **kw)

    def right_contraction(self, a, b, **kw):
        """Row-wise Multivector::right_contraction, src/multivector.rs:1616-1661."""
        return self._product(3, a, b, **kw)

    def scalar_product(self, a, b, **kw):
        """Row-wise Multivector::scalar_product, src/multivector.rs:1673-1715 (blade 0 only, others 0)."""
        return self._product(4, a, b, **kw)

    def commutator(self, a, b):
        """(a*b - b*a).scale(0.5), src/multivector.rs:1732-1745 (each step rounded separately)."""
        return self.scale(self.sub(self.geometric_product(a, b), self.geometric_product(b, a)), 0.5)

    def anticommutator(self, a, b):
        """(a*b + b*a).scale(0.5), src/multivector.rs:1755-1768."""
        return self.scale(self.add(self.geometric_product(a, b), self.geometric_product(b, a)), 0.5)

    def regressive(self, a, b):
        """undual(dual(a) ^ dual(b)), src/multivector.rs:1781-1794; ValueError when the pseudoscalar is null."""
        a, b = _arr(a), _arr(b)
        dt = a.dtype
        w = self.outer_product(self.dual(a).astype(dt), self.dual(b).astype(dt))
        return self.undual(w).astype(dt)

    def product_abs_bound(self, op, a, b):
        """sum_i |a_i||b_j| over the op's term set: the scale tolerance tests divide errors by."""
        return self._product({"gp": 0, "outer": 1, "lc": 2, "rc": 3, "scalar": 4}[op], a, b, absmode=True)

    def sandwich(self, rotor, x, threads=1):
        x = _arr(x)
        rotor = np.ascontiguousarray(rotor, dtype=np.float64).reshape(-1)
        if rotor.size != self.num_blades or x.shape[1] != self.num_blades:
            raise ValueError("rotor / batch do not match the algebra")
        out = np.empty_like(x)
        f = getattr(_lib(), f"oracle_batch_sandwich_{_sfx(x.dtype)}")
        f(self._h, rotor.ctypes.data, x.ctypes.data, x.shape[0], out.ctypes.data, threads)
        return out

    def _norm(self, which, x):
        x = _arr(x)
        out = np.empty(x.shape[0], dtype=x.dtype)
        getattr(_lib(), f"oracle_batch_norm_{_sfx(x.dtype)}")(self._h, which, x.ctypes.data, x.shape[0], out.ctypes.data)
        return out

    def norm_squared(self, x):
        return self._norm(0, x)

    def norm(self, x):
        return self._norm(1, x)

    def normalized(self, x):
        x = _arr(x)
        out = np.empty_like(x)
        getattr(_lib(), f"oracle_batch_normalized_{_sfx(x.dtype)}")(self._h, x.ctypes.data, x.shape[0], out.ctypes.data)
        return out

    def _involution(self, which, x):
        x = _arr(x)
        out = np.empty_like(x)
        getattr(_lib(), f"oracle_batch_involution_{_sfx(x.dtype)}")(self._h, which, x.ctypes.data, x.shape[0], out.ctypes.data)
        return out

    def reverse(self, x):
        return self._involution(0, x)

    def grade_involution(self, x):
        return self._involution(1, x)

    def clifford_conjugate(self, x):
        return self._involution(2, x)

    def grade(self, x, k):
        x = _arr(x)
        out = np.empty_like(x)
        rc = getattr(_lib(), f"oracle_batch_grade_{_sfx(x.dtype)}")(self._h, k, x.ctypes.data, x.shape[0], out.ctypes.data)
        if rc != 0:
            raise ValueError(f"grade {k} exceeds dimension {self.dimension} of algebra")
        return out

    def _addsub(self, which, a, b):
        a = _arr(a)
        b = _arr(b, dtype=a.dtype)
        na, nbr = a.shape[0], b.shape[0]
        n = na if na == nbr or nbr == 1 else (nbr if na == 1 else -1)
        if n < 0:
            raise ValueError(f"batch sizes must match or be 1 for broadcasting: {na} vs {nbr}")
        out = np.empty((n, self.num_blades), dtype=a.dtype)
        getattr(_lib(), f"oracle_batch_addsub_{_sfx(a.dtype)}")(self._h, which, a.ctypes.data, na, b.ctypes.data, nbr, out.ctypes.data)
        return out

    def add(self, a, b):
        return self._addsub(0, a, b)

    def sub(self, a, b):
        return self._addsub(1, a, b)

    def scale(self, x, factor):
        x = _arr(x)
        out = np.empty_like(x)
        getattr(_lib(), f"oracle_batch_scale_{_sfx(x.dtype)}")(self._h, x.ctypes.data, x.shape[0], float(factor), out.ctypes.data)
        return out

    def div(self, x, s):
        if s == 0.0:
            raise ZeroDivisionError("division by zero")
        return self.scale(x, 1.0 / s)

    def from_vectors(self, coords):
        coords = _arr(coords)
        if coords.shape[1] != self.dimension:
            raise ValueError(f"coords must have {self.dimension} columns")
        out = np.empty((coords.shape[0], self.num_blades), dtype=coords.dtype)
        getattr(_lib(), f"oracle_from_vectors_{_sfx(coords.dtype)}")(self._h, coords.ctypes.data, coords.shape[0], out.ctypes.data)
        return out

    def to_vector_coords(self, x):
        x = _arr(x)
        out = np.empty((x.shape[0], self.dimension), dtype=x.dtype)
        getattr(_lib(), f"oracle_to_vector_coords_{_sfx(x.dtype)}")(self._h, x.ctypes.data, x.shape[0], out.ctypes.data)
        return out

    def from_bivectors(self, biv):
        biv = _arr(biv)
        nbiv = self.dimension * (self.dimension - 1) // 2
        if biv.shape[1] != nbiv:
            raise ValueError(f"bivector_coeffs must have {nbiv} columns")
        out = np.empty((biv.shape[0], self.num_blades), dtype=biv.dtype)
        getattr(_lib(), f"oracle_from_bivectors_{_sfx(biv.dtype)}")(self._h, biv.ctypes.data, biv.shape[0], out.ctypes.data)
        return out

    # -- scalar twins
    def simd_geometric_product(self, a, b):
        a = np.ascontiguousarray(a, dtype=np.float64).reshape(-1)
        b = np.ascontiguousarray(b, dtype=np.float64).reshape(-1)
        out = np.empty(self.num_blades)
        _lib().oracle_simd_geometric_product(self._h, a.ctypes.data, b.ctypes.data, out.ctypes.data)
        return out

    def _dual(self, which, x):
        x = _arr(x, dtype=np.float64)
        out = np.empty_like(x)
        ok = _lib().oracle_dual_rows(self._h, which, x.ctypes.data, x.shape[0], out.ctypes.data)
        if not ok:
            raise ValueError("cannot invert multivector with zero norm")
        return out

    def dual(self, x):
        return self._dual(0, x)

    def undual(self, x):
        return self._dual(1, x)

    def right_dual(self, x):
        return self._dual(2, x)

    def pga_dual(self, x):
        if self.signature != (3, 0, 1):
            raise ValueError("pga_dual requires a PGA3D element")
        x = _arr(x, dtype=np.float64)
        out = np.empty_like(x)
        _lib().oracle_pga_dual_rows(x.ctypes.data, x.shape[0], out.ctypes.data)
        return out

    def sum(self, x):
        """(sum, sum of magnitudes) per blade in long double -- NEW surface, no reference counterpart."""
        x = _arr(x)
        out = np.empty(self.num_blades)
        mag = np.empty(self.num_blades)
        getattr(_lib(), f"oracle_sum_rows_{_sfx(x.dtype)}")(x.ctypes.data, x.shape[0], self.num_blades, out.ctypes.data, mag.ctypes.data)
        return out, mag


# ----------------------------------------------------------------------------- STA batched constructors
def sta_bivectors(e, b):
    """Row-wise Multivector::sta_bivector, src/sta.rs:202-212: c[3,5,9] = E, c[12] = Bx, c[10] = -By, c[6] = Bz."""
    e, b = _arr(e), _arr(b)
    out = np.zeros((e.shape[0], 16), dtype=e.dtype)
    out[:, 3], out[:, 5], out[:, 9] = e[:, 0], e[:, 1], e[:, 2]
    out[:, 12], out[:, 10], out[:, 6] = b[:, 0], -b[:, 1], b[:, 2]
    return out


def sta_boosts(v):
    """Row-wise Multivector::sta_boost, src/sta.rs:253-296 (float64): identity below 1e-14, NaN row where the
    reference raises (|v| >= 1)."""
    v = _arr(v, dtype=np.float64)
    out = np.zeros((v.shape[0], 16))
    for r in range(v.shape[0]):  # small test batches only
        vx, vy, vz = (float(t) for t in v[r])
        vmag = np.sqrt(vx * vx + vy * vy + vz * vz)
        if vmag < 1e-14:
            out[r, 0] = 1.0
        elif vmag >= 1.0:
            out[r, [0, 3, 5, 9]] = np.nan
        else:
            half = np.arctanh(vmag) / 2.0
            ch, sh = np.cosh(half), np.sinh(half)
            out[r, 0] = ch
            out[r, 3], out[r, 5], out[r, 9] = -sh * (vx / vmag), -sh * (vy / vmag), -sh * (vz / vmag)
    return out


# ----------------------------------------------------------------------------- PGA points (compact form)
def pga_points_embed(xyz):
    """Row-wise Multivector::pga_point, src/pga.rs:131-134: c[7]=1, c[14]=-x, c[13]=y, c[11]=-z."""
    xyz = _arr(xyz)
    out = np.zeros((xyz.shape[0], 16), dtype=xyz.dtype)
    out[:, 7] = 1.0
    out[:, 14], out[:, 13], out[:, 11] = -xyz[:, 0], xyz[:, 1], -xyz[:, 2]
    return out


def pga_points_extract(c):
    """Row-wise Multivector::pga_point_coords, src/pga.rs:579-599; NaN where the reference raises (|w| < 1e-10)."""
    c = _arr(c)
    w = c[:, 7]
    ok = np.abs(w) >= 1e-10
    safe = np.where(ok, w, 1)
    out = np.stack([-c[:, 14] / safe, c[:, 13] / safe, -c[:, 11] / safe], axis=1).astype(c.dtype)
    out[~ok] = np.nan
    return out


def pga_transform_points(motor, xyz, threads=1):
    """pga_point_coords(M * pga_point(xyz) * ~M): embedding, batch.rs:434-503 sandwich, extraction."""
    return pga_points_extract(OracleAlgebra(3, 0, 1).sandwich(motor, pga_points_embed(xyz), threads=threads))
