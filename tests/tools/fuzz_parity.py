"""Randomised strict-mode parity sweep through the raw C ABI: random signatures, dtypes, layouts, row counts, row
pitches, operators and broadcast patterns, every result compared bit for bit with the CPU oracle.  The TMA thresholds
are randomised too, so both kernel forms see every shape.  Prints one summary line; exits non-zero on a mismatch."""
import ctypes as C
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))

import numpy as np

import oracle
from largecrimsoncanine_b200 import cabi

lib = cabi.lib()
rng = np.random.default_rng(int(os.environ.get("FUZZ_SEED", "2024")))
SIGS = [(3, 0, 0), (3, 0, 1), (4, 1, 0), (1, 3, 0), (2, 0, 0), (2, 1, 1), (1, 0, 0), (4, 0, 0), (0, 0, 2), (3, 1, 1), (5, 0, 0), (1, 1, 1)]
S = cabi.LCC_FLAG_STRICT_FP


def dev_put(host):
    p = C.c_void_p()
    cabi.check(lib.lcc_device_alloc(C.byref(p), max(host.nbytes, 16), None))
    cabi.check(lib.lcc_memcpy_h2d(p, host.ctypes.data, host.nbytes, None))
    return p


def dev_get(p, like):
    out = np.empty_like(like)
    cabi.check(lib.lcc_memcpy_d2h(out.ctypes.data, p, out.nbytes, None))
    cabi.check(lib.lcc_stream_sync(None))
    return out


class Operand:
    """A batch in a random layout / pitch with canary values in the padding."""

    def __init__(self, x, layout, dt):
        n, nb = x.shape
        self.n, self.nb, self.layout, self.dt = n, nb, layout, dt
        es = np.dtype(dt).itemsize
        unit = 16 // es
        if layout == cabi.LCC_AOS:
            self.ld = nb + int(rng.choice([0, 0, unit, 1]))
            img = np.full((n, self.ld), 7.0, dt)
            img[:, :nb] = x
        else:
            base = ((n + unit - 1) // unit) * unit  # SoA views need a pitch that keeps blade rows 16-byte aligned
            self.ld = base + unit * int(rng.choice([0, 0, 3, 16]))
            self.base = base  # rows n..base-1 complete the last 16-byte vector: the contract lets kernels write them
            img = np.full((nb, self.ld), 7.0, dt)
            img[:, :n] = x.T
        self.img = img
        self.ptr = dev_put(img)
        code = cabi.LCC_F64 if dt == np.float64 else cabi.LCC_F32
        self.view = cabi.view(self.ptr.value, n, self.ld, code, layout)

    def read(self):
        got = dev_get(self.ptr, self.img)
        if self.layout == cabi.LCC_AOS:
            assert np.all(got[:, self.nb:] == 7.0), "row padding overwritten"
            return got[:, :self.nb]
        assert np.all(got[:, self.base:] == 7.0), "blade-row padding beyond the last vector overwritten"
        return got[:, :self.n].T

    def free(self):
        cabi.check(lib.lcc_device_free(self.ptr, None))


def main(cases):
    t0, done = time.time(), 0
    for case in range(cases):
        sig = SIGS[rng.integers(len(SIGS))]
        o, alg = oracle.OracleAlgebra(*sig), cabi.Algebra(*sig)
        nb = o.num_blades
        dt = np.float64 if rng.random() < 0.5 else np.float32
        layout = cabi.LCC_AOS if rng.random() < 0.5 else cabi.LCC_SOA
        n = int(rng.choice([1, 2, 3, 31, 64, 65, 255, 1000, 4096, 4097, 6000, 20011, 70000, 70001, 131072]))
        cabi.check(lib.lcc_set_option(b"aos_tma_min_rows", int(rng.choice([0, 128, 4096, 1 << 40]))))
        cabi.check(lib.lcc_set_option(b"soa_tma_min_bytes", int(rng.choice([0, 1 << 20, 1 << 40]))))
        a = rng.standard_normal((n, nb)).astype(dt)
        b = rng.standard_normal((n, nb)).astype(dt)
        bc = int(rng.integers(3)) if n > 1 else 0  # 0 none, 1 left operand is one row, 2 right operand is one row
        a_used, b_used = (a[:1] if bc == 1 else a), (b[:1] if bc == 2 else b)
        A, B = Operand(a_used, layout, dt), Operand(b_used, layout, dt)
        out, out2 = Operand(np.zeros((n, nb), dt), layout, dt), Operand(np.zeros((n, nb), dt), layout, dt)
        kind = rng.integers(10)
        cabi.check(lib.lcc_set_option(b"host_chunk_bytes", int(rng.choice([1 << 16, 1 << 18, 64 << 20]))))
        label = f"case {case}: sig {sig} {np.dtype(dt).name} layout {layout} n {n} bc {bc} kind {kind}"
        try:
            if kind == 0:
                op, ref = [(cabi.LCC_OP_GP, o.geometric_product), (cabi.LCC_OP_OUTER, o.outer_product), (cabi.LCC_OP_LC, o.left_contraction),
                           (cabi.LCC_OP_RC, o.right_contraction), (cabi.LCC_OP_SCALAR, o.scalar_product),
                           (cabi.LCC_OP_COMMUTATOR, o.commutator), (cabi.LCC_OP_ANTICOMMUTATOR, o.anticommutator)][rng.integers(7)]
                cabi.check(lib.lcc_batch_product(alg.h, op, C.byref(A.view), C.byref(B.view), C.byref(out.view), S, None))
                assert np.array_equal(out.read(), ref(a_used, b_used)), label + f" op {op}"
            elif kind == 1:
                cabi.check(lib.lcc_batch_gp_outer(alg.h, C.byref(A.view), C.byref(B.view), C.byref(out.view), C.byref(out2.view), S, None))
                assert np.array_equal(out.read(), o.geometric_product(a_used, b_used)), label
                assert np.array_equal(out2.read(), o.outer_product(a_used, b_used)), label
            elif kind == 2 and bc != 1:
                rotor = rng.standard_normal(nb)
                if rng.random() < 0.5:
                    rotor *= np.array([1.0 if bin(i).count("1") % 2 == 0 else 0.0 for i in range(nb)])
                A2 = A if bc == 0 else Operand(a, layout, dt)
                cabi.check(lib.lcc_batch_sandwich(alg.h, rotor.ctypes.data_as(C.POINTER(C.c_double)), C.byref(A2.view), C.byref(out.view), S, None))
                assert np.array_equal(out.read(), o.sandwich(rotor, a)), label
                if A2 is not A:
                    A2.free()
            elif kind == 3 and bc == 0:
                op, arg, ref = [(cabi.LCC_UN_REVERSE, 0, o.reverse), (cabi.LCC_UN_INVOLUTE, 0, o.grade_involution),
                                (cabi.LCC_UN_CONJUGATE, 0, o.clifford_conjugate), (cabi.LCC_UN_GRADE, int(rng.integers(o.dimension + 1)), None)][rng.integers(4)]
                cabi.check(lib.lcc_batch_unary(alg.h, op, arg, C.byref(A.view), C.byref(out.view), None))
                want = o.grade(a, arg) if ref is None else ref(a)
                assert np.array_equal(out.read(), want), label
            elif kind == 4:
                which = int(rng.integers(2))
                cabi.check(lib.lcc_batch_elementwise(alg.h, which, C.byref(A.view), C.byref(B.view), C.byref(out.view), None))
                assert np.array_equal(out.read(), (o.sub if which else o.add)(a_used, b_used)), label
            elif kind == 5 and bc == 0:
                cabi.check(lib.lcc_batch_normalized(alg.h, C.byref(A.view), C.byref(out.view), S, None))
                assert np.array_equal(out.read(), o.normalized(a)), label
                norms = np.zeros(n, dt)
                pn = dev_put(norms)
                cabi.check(lib.lcc_batch_norm(alg.h, cabi.LCC_NORM_SQUARED, C.byref(A.view), pn, S, None))
                assert np.array_equal(dev_get(pn, norms), o.norm_squared(a)), label
                cabi.check(lib.lcc_device_free(pn, None))
            elif kind == 6 and bc == 0:  # layout conversion both ways (TMA transpose from 4096 rows on, scalar tail and fallbacks)
                other = Operand(np.zeros((n, nb), dt), cabi.LCC_SOA if layout == cabi.LCC_AOS else cabi.LCC_AOS, dt)
                cabi.check(lib.lcc_batch_convert(alg.h, C.byref(A.view), C.byref(other.view), None))
                assert np.array_equal(other.read(), a), label + " (convert)"
                cabi.check(lib.lcc_batch_convert(alg.h, C.byref(other.view), C.byref(out.view), None))
                assert np.array_equal(out.read(), a), label + " (convert back)"
                sums = np.zeros(nb, dt)
                ps = dev_put(sums)
                cabi.check(lib.lcc_batch_sum(alg.h, C.byref(A.view), ps, None))
                want, mag = o.sum(a)
                assert np.all(np.abs(dev_get(ps, sums) - want) <= (1e-12 if dt == np.float64 else 2e-6) * np.maximum(mag, 1.0)), label + " (sum)"
                cabi.check(lib.lcc_device_free(ps, None))
                other.free()
            elif kind == 7 and bc != 2:  # one rotor per row, or (bc == 1) ONE device-resident rotor row for the whole batch (K2r / composed)
                cabi.check(lib.lcc_batch_sandwich_rows(alg.h, C.byref(A.view), C.byref(B.view), C.byref(out.view), S, None))
                rev = np.array([1.0 if bin(i).count("1") % 4 < 2 else -1.0 for i in range(nb)], dt)
                assert np.array_equal(out.read(), o.geometric_product(o.geometric_product(a_used, b), a_used * rev)), label
            elif kind == 8 and bc == 0:  # host array -> batch -> host array through the chunked pipelines (pageable and pinned)
                host = a
                if rng.random() < 0.5:
                    pp = C.c_void_p()
                    cabi.check(lib.lcc_host_cache_alloc(C.byref(pp), a.nbytes))
                    host = np.ctypeslib.as_array(C.cast(pp, C.POINTER(C.c_double if dt == np.float64 else C.c_float)), shape=a.shape)
                    host[...] = a
                cabi.check(lib.lcc_batch_upload(alg.h, host.ctypes.data, n, C.byref(out.view), None))
                assert np.array_equal(out.read(), a), label + " (upload)"
                back = np.zeros_like(a)
                cabi.check(lib.lcc_batch_download(alg.h, C.byref(out.view), back.ctypes.data, None))
                assert np.array_equal(back, a), label + " (download)"
                if host is not a:
                    cabi.check(lib.lcc_host_cache_free(pp))
            elif kind == 9 and bc == 0:  # broadcast product with the fixed operand on the host
                side = int(rng.integers(2))
                op, ref = [(cabi.LCC_OP_GP, o.geometric_product), (cabi.LCC_OP_OUTER, o.outer_product), (cabi.LCC_OP_LC, o.left_contraction)][rng.integers(3)]
                fixed = b[0].astype(np.float64)
                cabi.check(lib.lcc_batch_fixed_product(alg.h, op, side, fixed.ctypes.data_as(C.POINTER(C.c_double)), C.byref(A.view), C.byref(out.view), S, None))
                want = ref(b[:1], a) if side == 0 else ref(a, b[:1])
                assert np.array_equal(out.read(), want), label + f" (fixed product op {op} side {side})"
            else:
                continue
            assert np.array_equal(A.read(), a_used) and np.array_equal(B.read(), b_used), label + " (inputs modified)"
            done += 1
        finally:
            for x in (A, B, out, out2):
                x.free()
    print(f"fuzz parity: {done} of {cases} random cases executed, all bit-exact in strict mode ({time.time() - t0:.1f} s)")


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 400)
