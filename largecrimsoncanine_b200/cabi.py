"""ctypes binding of include/lcc_b200.h (the C ABI in lib/liblcc_b200.so).

This is the same boundary a Rust maintainer would bind with an ``extern "C"`` block (see
INTEGRATION.md); Python users normally go through the ``largecrimsoncanine`` extension or
``lcc.torch`` instead.  Nothing here computes: it only declares signatures and turns a failing
``lcc_status`` into the exception type the reference raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "liblcc_b200.so")

LCC_OK, LCC_ERR_VALUE, LCC_ERR_INDEX, LCC_ERR_ZERODIV, LCC_ERR_CUDA, LCC_ERR_NO_DEVICE, LCC_ERR_UNSUPPORTED, LCC_ERR_ALLOC = range(8)
LCC_F32, LCC_F64 = 0, 1
LCC_SOA, LCC_AOS = 0, 1
LCC_OP_GP, LCC_OP_OUTER, LCC_OP_LC = 0, 1, 2
LCC_OP_RC, LCC_OP_SCALAR, LCC_OP_COMMUTATOR, LCC_OP_ANTICOMMUTATOR, LCC_OP_REGRESSIVE = 4, 5, 6, 7, 8
(LCC_UN_REVERSE, LCC_UN_INVOLUTE, LCC_UN_CONJUGATE, LCC_UN_GRADE, LCC_UN_DUAL, LCC_UN_UNDUAL,
 LCC_UN_RIGHT_DUAL, LCC_UN_PGA_DUAL, LCC_UN_EVEN, LCC_UN_ODD, LCC_UN_COPY) = range(11)
LCC_EW_ADD, LCC_EW_SUB = 0, 1
LCC_NORM_SQUARED, LCC_NORM = 0, 1
LCC_FLAG_STRICT_FP = 1


class LccView(C.Structure):
    """lcc_view: one device batch (include/lcc_b200.h)."""

    _fields_ = [("data", C.c_void_p), ("n", C.c_int64), ("ld", C.c_int64), ("dtype", C.c_int32), ("layout", C.c_int32)]


class LccError(RuntimeError):
    pass


_lib = None


def lib():
    """Load liblcc_b200.so (fails loudly when the native build is missing -- there is no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m largecrimsoncanine_b200._build` "
            "(the batched operations have no Python or CPU fallback)"
        )
    L = C.CDLL(LIB_PATH)
    vp, i32, i64, u32, dbl, sz = C.c_void_p, C.c_int, C.c_int64, C.c_uint, C.c_double, C.c_size_t
    V = C.POINTER(LccView)
    sig = {
        "lcc_version": (C.c_char_p, []),
        "lcc_last_error": (C.c_char_p, []),
        "lcc_device_count": (i32, [C.POINTER(i32)]),
        "lcc_set_device": (i32, [i32]),
        "lcc_get_device": (i32, [C.POINTER(i32)]),
        "lcc_stream": (i32, [C.POINTER(vp)]),
        "lcc_stream_sync": (i32, [vp]),
        "lcc_algebra_create": (i32, [i32, i32, i32, C.POINTER(vp)]),
        "lcc_algebra_destroy": (None, [vp]),
        "lcc_algebra_p": (i32, [vp]),
        "lcc_algebra_q": (i32, [vp]),
        "lcc_algebra_r": (i32, [vp]),
        "lcc_algebra_dimension": (i32, [vp]),
        "lcc_algebra_num_blades": (i32, [vp]),
        "lcc_algebra_cayley_signs": (C.POINTER(C.c_int8), [vp]),
        "lcc_algebra_cayley_blades": (C.POINTER(C.c_uint16), [vp]),
        "lcc_algebra_grade_mask": (i32, [vp, i32, C.POINTER(C.c_uint64), i32]),
        "lcc_algebra_blade_name": (i32, [vp, i32, C.c_char_p, i32]),
        "lcc_algebra_is_specialised": (i32, [vp]),
        "lcc_algebra_matrix_rep": (i32, [vp, C.POINTER(C.c_uint8), C.POINTER(C.c_int8)]),
        "lcc_algebra_matrix_rep_layout": (i32, [vp, C.POINTER(C.c_uint32)]),
        "lcc_algebra_matrix_rep_layout_aos": (i32, [vp, i32, C.POINTER(C.c_uint32)]),
        "lcc_reverse_sign": (i32, [i32]),
        "lcc_grade_involution_sign": (i32, [i32]),
        "lcc_clifford_conjugate_sign": (i32, [i32]),
        "lcc_device_alloc": (i32, [C.POINTER(vp), sz, vp]),
        "lcc_device_free": (i32, [vp, vp]),
        "lcc_host_alloc_pinned": (i32, [C.POINTER(vp), sz]),
        "lcc_host_free_pinned": (i32, [vp]),
        "lcc_memcpy_h2d": (i32, [vp, vp, sz, vp]),
        "lcc_memcpy_d2h": (i32, [vp, vp, sz, vp]),
        "lcc_memcpy_d2d": (i32, [vp, vp, sz, vp]),
        "lcc_memset_zero": (i32, [vp, sz, vp]),
        "lcc_batch_alloc": (i32, [vp, i32, i64, V, vp]),
        "lcc_batch_free": (i32, [V, vp]),
        "lcc_batch_convert": (i32, [vp, V, V, vp]),
        "lcc_batch_scatter_columns": (i32, [vp, vp, i32, C.POINTER(C.c_int32), V, vp]),
        "lcc_batch_gather_columns": (i32, [vp, V, i32, C.POINTER(C.c_int32), vp, vp]),
        "lcc_batch_copy_rows": (i32, [vp, V, i64, i64, V, i64, vp]),
        "lcc_batch_product": (i32, [vp, i32, V, V, V, u32, vp]),
        "lcc_batch_gp_outer": (i32, [vp, V, V, V, V, u32, vp]),
        "lcc_batch_fixed_product": (i32, [vp, i32, i32, C.POINTER(dbl), V, V, u32, vp]),
        "lcc_batch_product_sum": (i32, [vp, i32, V, i32, V, vp, u32, vp]),
        "lcc_batch_product_vjp": (i32, [vp, i32, i32, V, V, V, u32, vp]),
        "lcc_batch_sandwich": (i32, [vp, C.POINTER(dbl), V, V, u32, vp]),
        "lcc_batch_sandwich_rows": (i32, [vp, V, V, V, u32, vp]),
        "lcc_batch_sandwich_rows_vjp": (i32, [vp, V, V, V, V, V, u32, vp]),
        "lcc_batch_unary": (i32, [vp, i32, i32, V, V, vp]),
        "lcc_batch_scale": (i32, [vp, V, dbl, V, vp]),
        "lcc_batch_elementwise": (i32, [vp, i32, V, V, V, vp]),
        "lcc_batch_norm": (i32, [vp, i32, V, vp, u32, vp]),
        "lcc_batch_normalized": (i32, [vp, V, V, u32, vp]),
        "lcc_batch_sum": (i32, [vp, V, vp, vp]),
        "lcc_pga_points_embed": (i32, [vp, vp, V, vp]),
        "lcc_pga_points_extract": (i32, [vp, V, vp, vp]),
        "lcc_cga_points_embed": (i32, [vp, vp, V, vp]),
        "lcc_cga_points_extract": (i32, [vp, V, vp, vp]),
        "lcc_pga_points_sandwich": (i32, [vp, i32, C.POINTER(dbl), vp, vp, i64, u32, vp]),
        "lcc_sta_bivectors_embed": (i32, [vp, vp, V, vp]),
        "lcc_sta_boosts_embed": (i32, [vp, vp, V, vp]),
        "lcc_batch_upload": (i32, [vp, vp, i64, V, vp]),
        "lcc_batch_download": (i32, [vp, V, vp, vp]),
        "lcc_host_cache_alloc": (i32, [C.POINTER(vp), sz]),
        "lcc_host_cache_free": (i32, [vp]),
        "lcc_host_cache_trim": (i32, []),
        "lcc_host_cache_stats": (i32, [C.POINTER(sz), C.POINTER(sz)]),
        "lcc_host_is_pinned": (i32, [vp, sz]),
        "lcc_host_sandwich": (i32, [vp, i32, C.POINTER(dbl), vp, vp, i64, u32]),
        "lcc_host_product": (i32, [vp, i32, i32, vp, i64, vp, i64, vp, u32]),
        "lcc_host_pga_points_sandwich": (i32, [vp, i32, C.POINTER(dbl), vp, vp, i64, u32]),
        "lcc_shard_rows": (i32, [i64, i32, i32, C.POINTER(i64), C.POINTER(i64)]),
        "lcc_shard_unique_id": (i32, [vp]),
        "lcc_shard_group_create": (i32, [i32, C.POINTER(i32), C.POINTER(vp)]),
        "lcc_shard_group_create_rank": (i32, [i32, i32, i32, vp, C.POINTER(vp)]),
        "lcc_shard_group_destroy": (None, [vp]),
        "lcc_shard_group_local_size": (i32, [vp]),
        "lcc_shard_group_world_size": (i32, [vp]),
        "lcc_shard_group_first_rank": (i32, [vp]),
        "lcc_shard_group_device": (i32, [vp, i32]),
        "lcc_shard_nccl_version": (i32, []),
        "lcc_shard_group_stream": (i32, [vp, i32, C.POINTER(vp)]),
        "lcc_shard_group_sync": (i32, [vp]),
        "lcc_shard_batch_alloc": (i32, [vp, vp, i32, i64, V, C.POINTER(i64)]),
        "lcc_shard_broadcast_alloc": (i32, [vp, vp, i32, vp, V]),
        "lcc_shard_batch_free": (i32, [vp, V]),
        "lcc_shard_upload": (i32, [vp, vp, vp, i64, V]),
        "lcc_shard_download": (i32, [vp, vp, V, vp, i64]),
        "lcc_shard_product": (i32, [vp, vp, i32, V, V, V, u32]),
        "lcc_shard_sandwich": (i32, [vp, vp, C.POINTER(dbl), V, V, u32]),
        "lcc_shard_unary": (i32, [vp, vp, i32, i32, V, V]),
        "lcc_shard_sum": (i32, [vp, vp, V, i64, vp]),
        "lcc_shard_product_sum": (i32, [vp, vp, i32, V, i32, V, i64, vp, u32]),
        "lcc_shard_product_sum_async": (i32, [vp, vp, i32, V, i32, V, i64, i32, u32]),
        "lcc_shard_reduce_fetch": (i32, [vp, vp, i32, i32, vp]),
        "lcc_set_option": (i32, [C.c_char_p, i64]),
        "lcc_get_option": (i32, [C.c_char_p, C.POINTER(i64)]),
        "lcc_kernel_launch_count": (C.c_uint64, []),
        "lcc_kernel_launch_count_reset": (None, []),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype, fn.argtypes = res, args
    L._lcc_signatures = sig
    _lib = L
    return L


_EXC = {LCC_ERR_VALUE: ValueError, LCC_ERR_INDEX: IndexError, LCC_ERR_ZERODIV: ZeroDivisionError, LCC_ERR_ALLOC: MemoryError}


def check(status: int) -> None:
    """Raise the reference's exception type for a failing lcc_status."""
    if status == LCC_OK:
        return
    msg = lib().lcc_last_error().decode(errors="replace")
    raise _EXC.get(status, LccError)(msg)


class Algebra:
    """Owning handle of an lcc_algebra (host tables; no GPU needed)."""

    def __init__(self, p: int, q: int = 0, r: int = 0):
        h = C.c_void_p()
        check(lib().lcc_algebra_create(p, q, r, C.byref(h)))
        self.h = h
        self.p, self.q, self.r = p, q, r
        self.dimension = p + q + r
        self.num_blades = 1 << self.dimension

    def __del__(self):
        h, self.h = getattr(self, "h", None), None
        if h and _lib is not None:
            _lib.lcc_algebra_destroy(h)

    def cayley_signs(self):
        import numpy as np

        nb = self.num_blades
        return np.ctypeslib.as_array(lib().lcc_algebra_cayley_signs(self.h), shape=(nb * nb,)).copy()

    def cayley_blades(self):
        import numpy as np

        nb = self.num_blades
        return np.ctypeslib.as_array(lib().lcc_algebra_cayley_blades(self.h), shape=(nb * nb,)).copy()

    def grade_mask(self, grade: int) -> int:
        words = (self.num_blades + 63) // 64
        buf = (C.c_uint64 * words)()
        lib().lcc_algebra_grade_mask(self.h, grade, buf, words)
        return sum(int(buf[i]) << (64 * i) for i in range(words))

    def blade_name(self, index: int) -> str:
        buf = C.create_string_buffer(32)
        if lib().lcc_algebra_blade_name(self.h, index, buf, 32) < 0:
            raise IndexError(index)
        return buf.value.decode()


def view(ptr: int, n: int, ld: int, dtype: int, layout: int) -> LccView:
    return LccView(C.c_void_p(ptr), n, ld, dtype, layout)
