"""CPU-side checks of the C ABI and the host logic (no compute call is made without a GPU):
the library loads and exports every symbol include/lcc_b200.h declares; the host algebra tables
are bit-exact against the oracle, the SURVEY.md digests and the build-time generator; error
behaviour of the host-only entry points."""
import ctypes as C
import hashlib
import os
import re

import numpy as np
import pytest

import oracle
from largecrimsoncanine_b200 import cabi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SIGS = [(1, 0, 0), (2, 0, 0), (3, 0, 0), (3, 0, 1), (4, 1, 0), (1, 3, 0), (2, 1, 1), (0, 2, 0), (2, 0, 2), (6, 0, 0), (4, 2, 1), (8, 0, 0)]


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "lcc_b200.h")).read()
    return re.findall(r"LCC_API[^;]*?\b(lcc_[a-z0-9_]+)\s*\(", text)


def test_library_exports_every_declared_symbol():
    L = cabi.lib()
    names = declared_symbols()
    assert len(names) >= 50
    for name in names:
        assert hasattr(L, name), f"{name} declared in include/lcc_b200.h but not exported"
    assert set(names) == set(L._lcc_signatures), "cabi.py and the header disagree"
    assert L.lcc_version().decode().startswith("0.3.0")


@pytest.mark.parametrize("sig", SIGS)
def test_host_tables_bit_exact(sig):
    a, o = cabi.Algebra(*sig), oracle.OracleAlgebra(*sig)
    assert a.num_blades == o.num_blades
    assert np.array_equal(a.cayley_signs(), o.signs_int8())
    assert np.array_equal(a.cayley_blades(), o.blades_u16().astype(np.uint16))
    if a.dimension <= 6:  # the reference's one-word masks (blades.rs:211-220)
        assert [a.grade_mask(g) for g in range(a.dimension + 1)] == o.grade_masks()
    else:  # multi-word extension: same rule, popcount(i) == g
        for g in range(a.dimension + 1):
            m = a.grade_mask(g)
            assert all(((m >> i) & 1) == (bin(i).count("1") == g) for i in range(a.num_blades))
    assert a.grade_mask(a.dimension + 1) == 0  # registry.rs:224-232


def test_digests_and_names():
    want = {
        (3, 0, 0): "40bc89779d69516cb07eb06fbac3ad3977c2993fa25d30549b8567d059478f9a",
        (3, 0, 1): "28398685f6263fa53dd7fe21123656b681ccac926c017907649b14a9b0de7390",
        (4, 1, 0): "9c0bc8a9265f229dec3ad5743e29060b3c6034f1049545a51b9bad13a9e881e4",
        (1, 3, 0): "365efcddbaab4cb448d229369d7e213ee566ac8583e6c8382844b0a76e248c26",
        (8, 0, 0): "b9593cd2fdc94cdee10612bc8911e0d7bf4045a354132ddf10b82bcd6474a660",
    }
    for sig, digest in want.items():
        assert hashlib.sha256(cabi.Algebra(*sig).cayley_signs().tobytes()).hexdigest() == digest
    r3 = cabi.Algebra(3)
    assert [r3.blade_name(i) for i in range(8)] == ["1", "e1", "e2", "e12", "e3", "e13", "e23", "e123"]  # registry.rs:436-448
    assert cabi.Algebra(2, 0, 1).blade_name(4) == "e3"  # registry.rs:450-458
    L = cabi.lib()
    assert [L.lcc_reverse_sign(i) for i in (0, 1, 3, 7, 15)] == [1, 1, -1, -1, 1]
    assert [L.lcc_grade_involution_sign(i) for i in (0, 1, 3, 7)] == [1, -1, 1, -1]
    assert [L.lcc_clifford_conjugate_sign(i) for i in (0, 1, 3, 7, 15)] == [1, -1, -1, 1, 1]


def test_generated_term_lists_match_oracle():
    """The build-time generator's gather-form lists reproduce the oracle's tables and term order."""
    gen = os.path.join(ROOT, "largecrimsoncanine_b200", "csrc", "generated")
    for sig in [(3, 0, 0), (3, 0, 1), (4, 1, 0), (1, 3, 0)]:
        o = oracle.OracleAlgebra(*sig)
        nb = o.num_blades
        signs = o.signs_int8().reshape(nb, nb)
        seen = np.zeros((nb, nb), np.int8)
        last_i = {}
        for line in open(os.path.join(gen, "cayley_%d_%d_%d.inc" % sig)):
            m = re.match(r"LCC_TERM\((\d+), (\d+), (\d+), (\d), (\d)\)", line)
            if not m:
                continue
            k, i, j, neg, flags = map(int, m.groups())
            assert j == i ^ k and signs[i, j] == (-1 if neg else 1)
            assert last_i.get(k, -1) < i  # ascending left blade inside each output group
            last_i[k] = i
            assert bool(flags & 1) == ((i & j) == 0) and bool(flags & 2) == ((i & j) == i) and bool(flags & 4) == ((i & j) == j)
            seen[i, j] = signs[i, j]
        assert np.array_equal(seen, signs)  # every non-zero term emitted exactly where it belongs
        # vector-Jacobian lists (autograd of lcc.torch): the same table transposed onto either operand
        for wrt in (0, 1):
            seen = np.zeros((nb, nb), np.int8)
            for line in open(os.path.join(gen, "cayley_%d_%d_%d_vjp%d.inc" % (sig + (wrt,)))):
                m = re.match(r"LCC_VTERM\((\d+), (\d+), (\d+), (\d), (\d), (\d+)\)", line)
                if not m:
                    continue
                out, x, y, neg, flags, mi = map(int, m.groups())
                i, j = (out, y) if wrt == 0 else (x, out)  # forward term (i, j) this cotangent term comes from
                assert (x == i ^ j and y == j) if wrt == 0 else (x == i and y == i ^ j)
                assert signs[i, j] == (-1 if neg else 1) and mi == i & j and seen[i, j] == 0
                assert bool(flags & 1) == ((i & j) == 0) and bool(flags & 2) == ((i & j) == i) and bool(flags & 4) == ((i & j) == j)
                seen[i, j] = signs[i, j]
            assert np.array_equal(seen, signs)
    assert cabi.lib().lcc_algebra_is_specialised(cabi.Algebra(3, 0, 1).h) == 1
    assert cabi.lib().lcc_algebra_is_specialised(cabi.Algebra(2, 1, 1).h) == 0


@pytest.mark.parametrize("sig,has_rep", [((8, 0, 0), True), ((4, 4, 0), True), ((5, 3, 0), True), ((0, 8, 0), True), ((1, 7, 0), True),
                                         ((7, 0, 0), True), ((4, 3, 0), True), ((0, 7, 0), True), ((7, 1, 0), False), ((6, 1, 0), False),
                                         ((6, 2, 0), False), ((7, 0, 1), False), ((6, 0, 0), False), ((3, 0, 1), False)])
def test_matrix_representation_is_an_algebra_homomorphism(sig, has_rep):
    """csrc/matrix_rep.hpp: the 16 x 16 signed-permutation matrices found for Cl(p,q) reproduce the algebra's own Cayley
    table -- rho(e_A) rho(e_B) == sign[A][B] rho(e_{A^B}) -- and are linearly independent (distinct strings), so the
    block-diagonalised product of K8 is the reference product in another basis.  Checked with dense NumPy matrices."""
    L = cabi.lib()
    a = cabi.Algebra(*sig)
    nb = a.num_blades
    strings, signs = (C.c_uint8 * nb)(), (C.c_int8 * nb)()
    d = L.lcc_algebra_matrix_rep(a.h, strings, signs)
    assert (d == 16) == has_rep
    if not has_rep:
        return
    strings, signs = np.array(strings[:]), np.array(signs[:])
    assert len(set(strings.tolist())) == nb and set(np.abs(signs).tolist()) == {1}
    j = np.arange(16)

    def mat(A):
        x, z = int(strings[A]) >> 4, int(strings[A]) & 15
        m = np.zeros((16, 16))
        m[j ^ x, j] = np.array([(-1.0) ** bin(z & int(t)).count("1") for t in j])
        return signs[A] * m

    table = a.cayley_signs().reshape(nb, nb)
    assert np.array_equal(mat(0), np.eye(16))
    rng = np.random.default_rng(nb + sig[1])
    pairs = [(1 << i, 1 << k) for i in range(a.dimension) for k in range(a.dimension)] + [tuple(rng.integers(0, nb, 2)) for _ in range(300)]
    for A, B in pairs:
        assert np.array_equal(mat(A) @ mat(B), table[A, B] * mat(A ^ B)), (A, B)


def _rep_matrix(strings, signs, coeffs):
    """rho(M) from coefficients (csrc/matrix_rep.hpp, rep_matrix)."""
    m = np.zeros((16, 16))
    j = np.arange(16)
    for A, c in enumerate(coeffs):
        x, z = int(strings[A]) >> 4, int(strings[A]) & 15
        m[j ^ x, j] += c * signs[A] * np.array([(-1.0) ** bin(z & int(t)).count("1") for t in j])
    return m


def _wht16(v):
    v = np.array(v, dtype=np.float64)
    h = 1
    while h < 16:
        for i in range(16):
            if not i & h:
                v[i], v[i | h] = v[i] + v[i | h], v[i] - v[i | h]
        h <<= 1
    return v


@pytest.mark.parametrize("sig", [(8, 0, 0), (4, 4, 0), (5, 3, 0), (0, 8, 0), (1, 7, 0), (7, 0, 0), (4, 3, 0), (0, 7, 0), (3, 4, 0)])
def test_k8_shared_memory_layout(sig):
    """K8 v7 (kernels/matrix_rep_kernel.cu) never moves a value to another slot: matrix entry (f, s) lives in the slot of the
    coefficient with string (f ^ s, s), the Walsh-Hadamard passes and both GEMM passes work in place, and every address is
    F(slot) ^ row bytes from the host tables of csrc/matrix_rep.hpp (RepLayout).  This replays the kernel's address
    arithmetic on the host for one row -- stage 1, column pass and / or row pass, stage 3 -- against the algebra's own
    Cayley table for the left product, the right product and the sandwich, and checks that every warp-wide fragment
    access the kernel makes touches each bank once (the property the layout's basis is chosen for)."""
    L = cabi.lib()
    a = cabi.Algebra(*sig)
    nb = a.num_blades
    strings, signs = (C.c_uint8 * nb)(), (C.c_int8 * nb)()
    assert L.lcc_algebra_matrix_rep(a.h, strings, signs) == 16
    strings, signs = np.array(strings[:]), np.array(signs[:], dtype=np.float64)
    tables = (C.c_uint32 * 128)()
    assert L.lcc_algebra_matrix_rep_layout(a.h, tables) == 1
    t = np.array(tables[:], dtype=np.int64).reshape(8, 16)
    f_lo, f_hi = t[0], t[1]
    passes = {"column": (t[2], t[3], t[4]), "row": (t[5], t[6], t[7])}
    sign_at = np.zeros(256)
    for A in range(nb):
        sign_at[strings[A]] = signs[A]
    # every offset is a swizzle-consistent slot offset: F(s) = s << 7 | (s & 7) << 4
    for off in list(f_lo) + list(f_hi) + [o for p in passes.values() for o in list(p[0]) + list(p[1])]:
        assert off == ((off >> 7) << 7 | ((off >> 7) & 7) << 4)
    # the slots of string row a are the blades of its strings (the TMA engine delivers slot = blade)
    for A in range(nb):
        sa, sb = int(strings[A]) >> 4, int(strings[A]) & 15
        assert (f_hi[sa] ^ f_lo[sb]) >> 7 == A
    table = a.cayley_signs().reshape(nb, nb).astype(np.float64)
    rng = np.random.default_rng(sum(sig) * 7 + sig[1])
    x, m, r = rng.standard_normal(nb), rng.standard_normal(nb), rng.standard_normal(nb)
    idx = np.arange(nb)

    def gp(u, v):
        out = np.zeros(nb)
        for A in range(nb):
            np.add.at(out, A ^ idx, u[A] * table[A] * v)
        return out

    def run(plan):
        slots = np.full(256, np.nan)
        slots[:nb] = x
        for sa in range(16):  # stage 1
            where = [(f_hi[sa] ^ f_lo[b]) >> 7 for b in range(16)]
            v = [slots[where[b]] * sign_at[sa << 4 | b] if sign_at[sa << 4 | b] else 0.0 for b in range(16)]
            slots[where] = _wht16(v)
        for kind, fac in plan:  # stage 2: D[o] = sum_k fac[o][k] Z[k] in physical order
            fk, cw, kmap = passes[kind]
            phys = fac[np.ix_(kmap, kmap)]
            for w in range(16):
                where = [(fk[i] ^ cw[w]) >> 7 for i in range(16)]
                slots[where] = phys @ slots[where]
        for sa in range(16):  # stage 3
            where = [(f_hi[sa] ^ f_lo[b]) >> 7 for b in range(16)]
            v = _wht16(slots[where])
            slots[where] = [v[b] * sign_at[sa << 4 | b] for b in range(16)]
        return slots[:nb]

    rho_m = _rep_matrix(strings, signs, m)
    assert np.allclose(run([("column", rho_m / 16)]), gp(m, x), rtol=0, atol=1e-11)
    assert np.allclose(run([("row", rho_m.T / 16)]), gp(x, m), rtol=0, atol=1e-11)
    rev = np.array([1.0 if (bin(A).count("1") % 4) < 2 else -1.0 for A in range(nb)])
    rho_r, rho_rr = _rep_matrix(strings, signs, r), _rep_matrix(strings, signs, r * rev)
    assert np.allclose(run([("column", rho_r / 16), ("row", rho_rr.T)]), gp(gp(r, x), r * rev), rtol=0, atol=1e-9)
    # bank checks: lane = 4 g + t
    for fk, cw, _ in passes.values():
        for w in range(16):
            for q in range(4):  # f32, m16n8k8: B loads (32 lanes x 4 bytes), D stores (half-warps x 8 bytes)
                for ks in range(2):
                    for h in range(2):
                        banks = {((fk[8 * ks + 4 * h + tt] ^ (g << 2) ^ cw[w] ^ (q << 5)) >> 2) & 31 for g in range(8) for tt in range(4)}
                        assert len(banks) == 32
                for mb in range(2):
                    for half in range(2):
                        banks = {((fk[g + 8 * mb] ^ (tt << 3) ^ cw[w] ^ (q << 5)) >> 3) & 15 for g in range(4 * half, 4 * half + 4) for tt in range(4)}
                        assert len(banks) == 16
            for q in range(2):  # f64, m8n8k4: B loads (half-warps x 8 bytes), D stores (quarter-warps x 16 bytes)
                for ks in range(4):
                    for half in range(2):
                        banks = {((fk[4 * ks + tt] ^ (g << 3) ^ cw[w] ^ (q << 6)) >> 3) & 15 for g in range(4 * half, 4 * half + 4) for tt in range(4)}
                        assert len(banks) == 16
                for mb in range(2):
                    for quarter in range(4):
                        banks = {((fk[g + 8 * mb] ^ (tt << 4) ^ cw[w] ^ (q << 6)) >> 4) & 7 for g in range(2 * quarter, 2 * quarter + 2) for tt in range(4)}
                        assert len(banks) == 8


@pytest.mark.parametrize("es", [4, 8])
@pytest.mark.parametrize("sig", [(8, 0, 0), (4, 4, 0), (5, 3, 0), (0, 8, 0), (1, 7, 0), (7, 0, 0), (4, 3, 0), (0, 7, 0), (3, 4, 0)])
def test_k8_reference_layout_units(sig, es):
    """K8 on reference-layout (AoS) batches (kernels/matrix_rep_kernel.cu, AOS = true): a unit is 32 (f32) / 16 (f64) whole
    rows, delivered by one 3-D TMA box as [chunk][row][128 bytes swizzled]; element (row r, blade A) sits at G(A) ^ R(r)
    (csrc/matrix_rep.hpp, RepLayoutAos).  Replays the kernel's address arithmetic for one row against the algebra's Cayley
    table -- left product, right product, sandwich -- and checks every warp-wide access against the bank rule."""
    L = cabi.lib()
    a = cabi.Algebra(*sig)
    nb = a.num_blades
    strings, signs = (C.c_uint8 * nb)(), (C.c_int8 * nb)()
    assert L.lcc_algebra_matrix_rep(a.h, strings, signs) == 16
    strings, signs = np.array(strings[:]), np.array(signs[:], dtype=np.float64)
    tables = (C.c_uint32 * (13 * 16))()
    assert L.lcc_algebra_matrix_rep_layout_aos(a.h, es, tables) == 1
    t = np.array(tables[:], dtype=np.int64).reshape(13, 16)
    g_lo, g_hi, amap = t[0], t[1], t[2]
    passes = {"column": tuple(t[3:8]), "row": tuple(t[8:13])}  # fk, fo, cw, kmap, omap

    def G(A):
        return ((A >> 5) << 12 | ((A >> 2) & 7) << 4 | (A & 3) << 2) if es == 4 else ((A >> 4) << 11 | ((A >> 1) & 7) << 4 | (A & 1) << 3)

    def R(r):
        return r << 7 | (r & 7) << 4

    blade_of_offset = {G(A): A for A in range(256)}
    assert sorted(amap) == list(range(16))
    sign_at = np.zeros(256)
    for A in range(nb):
        sign_at[strings[A]] = signs[A]
        sa, sb = int(strings[A]) >> 4, int(strings[A]) & 15
        assert g_hi[list(amap).index(sa)] ^ g_lo[sb] == G(A)
    table = a.cayley_signs().reshape(nb, nb).astype(np.float64)
    rng = np.random.default_rng(sum(sig) * 11 + sig[1] + es)
    x, m, r = rng.standard_normal(nb), rng.standard_normal(nb), rng.standard_normal(nb)
    idx = np.arange(nb)

    def gp(u, v):
        out = np.zeros(nb)
        for A in range(nb):
            np.add.at(out, A ^ idx, u[A] * table[A] * v)
        return out

    def run(plan):
        slots = np.full(256, np.nan)
        slots[:nb] = x
        for i in range(16):  # stage 1, physical string row i
            sa = amap[i]
            where = [blade_of_offset[g_hi[i] ^ g_lo[b]] for b in range(16)]
            v = [slots[where[b]] * sign_at[sa << 4 | b] if sign_at[sa << 4 | b] else 0.0 for b in range(16)]
            slots[where] = _wht16(v)
        for kind, fac in plan:  # stage 2: D[m] = sum_k fac[omap[m]][kmap[k]] Z[k]
            fk, fo, cw, kmap, omap = passes[kind]
            phys = fac[np.ix_(omap, kmap)]
            for w in range(16):
                src = [blade_of_offset[fk[i] ^ cw[w]] for i in range(16)]
                dst = [blade_of_offset[fo[i] ^ cw[w]] for i in range(16)]
                assert sorted(src) == sorted(dst)  # in place: the pass writes the slots it read
                slots[dst] = phys @ slots[src]
        for i in range(16):  # stage 3
            sa = amap[i]
            where = [blade_of_offset[g_hi[i] ^ g_lo[b]] for b in range(16)]
            v = _wht16(slots[where])
            slots[where] = [v[b] * sign_at[sa << 4 | b] for b in range(16)]
        return slots[:nb]

    rho_m = _rep_matrix(strings, signs, m)
    assert np.allclose(run([("column", rho_m / 16)]), gp(m, x), rtol=0, atol=1e-11)
    assert np.allclose(run([("row", rho_m.T / 16)]), gp(x, m), rtol=0, atol=1e-11)
    rev = np.array([1.0 if (bin(A).count("1") % 4) < 2 else -1.0 for A in range(nb)])
    rho_r, rho_rr = _rep_matrix(strings, signs, r), _rep_matrix(strings, signs, r * rev)
    assert np.allclose(run([("column", rho_r / 16), ("row", rho_rr.T)]), gp(gp(r, x), r * rev), rtol=0, atol=1e-9)

    def conflict_free(offsets):  # offsets in lane order; 4-byte accesses: the warp, 8-byte accesses: each half-warp
        if es == 4:
            return len({(o >> 2) & 31 for o in offsets}) == 32
        return all(len({(o >> 3) & 15 for o in offsets[h * 16:h * 16 + 16]}) == 16 for h in range(2))

    rows = 128 // es
    for wl in range(8):  # transforms: lane = 8 alpha + rho -> physical string row 4 (wl >> 1) + alpha at row 8 (wl & 1) + rho
        for b in range(16):
            for extra in ([0, 16] if es == 4 else [0]):
                assert conflict_free([g_hi[4 * (wl >> 1) + (lane >> 3)] ^ R(8 * (wl & 1) + (lane & 7) + extra) ^ g_lo[b] for lane in range(32)])
    for fk, fo, cw, _, _ in passes.values():  # MMA fragments: lane = 4 g + t
        for w in range(16):
            for q in range(rows // 8):
                for j in range(4):
                    assert conflict_free([fk[4 * j + (lane & 3)] ^ R(8 * q + (lane >> 2)) ^ cw[w] for lane in range(32)])
                for mb in range(2):
                    for e in range(2):
                        assert conflict_free([fo[(lane >> 2) + 8 * mb] ^ R(8 * q + 2 * (lane & 3) + e) ^ cw[w] for lane in range(32)])


def test_host_side_errors():
    L = cabi.lib()
    h = C.c_void_p()
    assert L.lcc_algebra_create(5, 3, 1, C.byref(h)) == cabi.LCC_ERR_UNSUPPORTED
    assert b"at most 8" in L.lcc_last_error()
    assert L.lcc_algebra_create(-1, 0, 0, C.byref(h)) == cabi.LCC_ERR_VALUE
    with pytest.raises(ValueError):
        cabi.check(cabi.LCC_ERR_VALUE)


def test_extension_host_logic():
    """Algebra / Multivector host behaviour of the `largecrimsoncanine` extension (src/pyalgebra.rs)."""
    import largecrimsoncanine as L

    a = L.Algebra(3, 0, 1)
    assert (a.p, a.q, a.r, a.dimension, a.num_blades) == (3, 0, 1, 4, 16)
    assert str(a) == repr(a) == "Cl(3,0,1)" and a == L.Algebra.pga(3) and hash(a) == hash(L.Algebra.pga(3))
    assert a.is_pga() and not a.is_euclidean() and L.Algebra.euclidean(3).is_euclidean()
    assert str(L.Algebra.cga(3)) == "Cl(4,1,0)" and str(L.Algebra.sta()) == "Cl(1,3,0)" and str(L.Algebra(2)) == "Cl(2,0,0)"
    assert a.grade_masks == [0x1, 0x116, 0x1668, 0x6880, 0x8000]
    assert np.array_equal(a.cayley_signs().ravel(), oracle.OracleAlgebra(3, 0, 1).signs_int8())
    v = L.Multivector.from_vector([3.0, 4.0, 0.0])
    assert v.norm() == 5.0 and v.to_vector_coords() == [3.0, 4.0, 0.0] and v.dims == 3 and v.algebra() is None
    assert L.Multivector.from_scalar(2.5, 3).coefficients() == [2.5, 0, 0, 0, 0, 0, 0, 0]
    with pytest.raises(ValueError, match="power of 2"):
        L.Multivector([1.0, 2.0, 3.0])
    r = L.Multivector.from_axis_angle(L.Multivector.from_vector([0.0, 0.0, 2.0]), np.pi / 2)
    np.testing.assert_allclose(r.coefficients(), [np.cos(np.pi / 4), 0, 0, -np.sin(np.pi / 4), 0, 0, 0, 0], atol=1e-15)
    p = L.Multivector.pga_point(a, 1.0, 2.0, 3.0)
    assert p.pga_point_coords() == (1.0, 2.0, 3.0) and p.algebra() == a
    with pytest.raises(ValueError, match="PGA3D"):
        L.Multivector.pga_point(L.Algebra.euclidean(3), 1.0, 2.0, 3.0)


def test_batched_calls_refuse_to_run_without_a_gpu():
    if os.path.exists("/dev/nvidiactl"):
        pytest.skip("a GPU is present")
    import largecrimsoncanine as L

    with pytest.raises(RuntimeError, match="no CPU fallback"):
        L.MultivectorBatch.from_numpy(L.Algebra.euclidean(3), np.zeros((4, 8)))
    # ONE-pair Multivector products are host C++ (the reference's per-object path, src/multivector.rs:1438-1481 /
    # src/simd.rs:55-135 -- SURVEY.md 2.1); they are not a fallback of anything batched, which still refuses above
    e12 = L.Multivector.from_vector([1.0, 0.0, 0.0]).geometric_product(L.Multivector.from_vector([0.0, 1.0, 0.0]))
    assert e12.coefficients() == [0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 0.0]
    # the host-buffer entry points validate their arguments first and then fail the same way (no host arithmetic anywhere)
    lib = cabi.lib()
    xyz, motor = np.zeros((8, 3), np.float32), np.zeros(16)
    motor[0] = 1.0
    mptr = motor.ctypes.data_as(C.POINTER(C.c_double))
    assert lib.lcc_host_pga_points_sandwich(cabi.Algebra(3, 0, 0).h, cabi.LCC_F32, mptr, xyz.ctypes.data, xyz.ctypes.data, 8, 0) == cabi.LCC_ERR_VALUE
    assert b"PGA3D" in lib.lcc_last_error()
    assert lib.lcc_host_pga_points_sandwich(cabi.Algebra(3, 0, 1).h, cabi.LCC_F32, mptr, xyz.ctypes.data, xyz.ctypes.data, 0, 0) == cabi.LCC_OK
    assert lib.lcc_host_pga_points_sandwich(cabi.Algebra(3, 0, 1).h, cabi.LCC_F32, mptr, xyz.ctypes.data, xyz.ctypes.data, 8, 0) == cabi.LCC_ERR_NO_DEVICE
    assert b"no CPU fallback" in lib.lcc_last_error()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        L.pga_transform_points(L.Algebra.pga(3), L.Multivector.from_scalar(1.0, 4), np.zeros((4, 3)))


@pytest.mark.parametrize("sig", [(3, 0, 0), (3, 0, 1), (4, 1, 0), (1, 3, 0), (2, 1, 1), (6, 0, 0)])
def test_single_pair_products_follow_the_reference_loops(sig):
    """Multivector (op) Multivector on the host == the oracle's restatement of the same reference loops, bit for bit,
    for the whole product family; 16 blades and more take the src/simd.rs order ((a*b)*sign)."""
    import largecrimsoncanine as L

    alg, o = L.Algebra(*sig), oracle.OracleAlgebra(*sig)
    nb = o.num_blades
    rng = np.random.default_rng(sum(sig) * 7 + sig[1])
    for trial in range(5):
        a, b = rng.standard_normal(nb), rng.standard_normal(nb)
        if trial == 3:
            a[rng.integers(0, nb, nb // 2)] = 0.0  # zero-skips
            b[rng.integers(0, nb, nb // 2)] = 0.0
        A, B = L.Multivector.from_list_in(list(a), alg), L.Multivector.from_list_in(list(b), alg)
        assert np.array_equal(A.geometric_product(B).coefficients(), o.geometric_product(a[None], b[None])[0])
        assert np.array_equal(A.outer_product(B).coefficients(), o.outer_product(a[None], b[None])[0])
        assert np.array_equal(A.left_contraction(B).coefficients(), o.left_contraction(a[None], b[None])[0])
        assert np.array_equal((A * B).coefficients(), (A.geometric_product(B)).coefficients())
        ab, ba = o.geometric_product(a[None], b[None])[0], o.geometric_product(b[None], a[None])[0]
        assert np.array_equal(A.commutator(B).coefficients(), (ab - ba) * 0.5)
        assert np.array_equal(A.anticommutator(B).coefficients(), (ab + ba) * 0.5)
        rx = o.geometric_product(a[None], b[None])
        rev = a * np.array([1.0 if (bin(i).count("1") % 4) < 2 else -1.0 for i in range(nb)])
        assert np.array_equal(A.sandwich(B).coefficients(), o.geometric_product(rx, rev[None])[0])  # R x ~R, two rounded stages


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` runs on host cores only (the oracle port of the reference loops) and prints ONE
    JSON line with the keys the driver reads; checked here on the small CPU-runnable configuration."""
    import json
    import subprocess
    import sys

    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "r3_gp",
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=300, check=True).stdout
    lines = [ln for ln in out.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "products/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] == 1
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert set(d["config"]) == {"workload", "rows_per_gpu", "l2_policy"}  # the workload definition, identical to the GPU arm's
    assert d["impl_config"]["strict_fp"] is True and "host memory" in d["impl_config"]["layout"]
