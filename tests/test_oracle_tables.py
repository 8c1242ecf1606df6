"""Pin the oracle's integer layer against every table KAT the reference's own unit tests hold.

Citations are /root/reference paths (file:line).  Bit-exact: blades, signs, grade masks.
"""
import hashlib

import numpy as np
import pytest

import oracle

EUCLID3 = (3, 0, 0)
PGA3 = (3, 0, 1)
CGA3 = (4, 1, 0)
STA = (1, 3, 0)


def bp(a, b, sig):
    return oracle.blade_product(a, b, *sig)


def test_scalar_and_basis_squares():  # src/algebra/cayley.rs:175-193
    assert bp(0, 0, EUCLID3) == (0, 1.0)
    assert bp(0, 1, EUCLID3) == (1, 1.0)
    assert bp(0, 7, EUCLID3) == (7, 1.0)
    assert bp(1, 0, EUCLID3) == (1, 1.0)
    for e in (1, 2, 4):
        assert bp(e, e, EUCLID3) == (0, 1.0)


def test_orthogonal_products_anticommute():  # cayley.rs:195-220
    assert bp(1, 2, EUCLID3) == (3, 1.0)
    assert bp(2, 1, EUCLID3) == (3, -1.0)
    for a, b in ((1, 2), (1, 4), (2, 4)):
        (b1, s1), (b2, s2) = bp(a, b, EUCLID3), bp(b, a, EUCLID3)
        assert b1 == b2 and s1 == -s2


def test_pga_degenerate():  # cayley.rs:226-256, registry.rs:389-400
    e0 = 8
    assert bp(e0, e0, PGA3) == (0, 0.0)
    for e in (1, 2, 4):
        assert bp(e, e, PGA3) == (0, 1.0)
    assert bp(1, e0, PGA3) == (1 | e0, 1.0)
    assert bp(e0, 1, PGA3) == (1 | e0, -1.0)
    # PGA2D: e0 is bit 2 (cayley.rs:327-345)
    assert bp(4, 4, (2, 0, 1)) == (0, 0.0)
    assert bp(1, 1, (2, 0, 1)) == (0, 1.0) and bp(2, 2, (2, 0, 1)) == (0, 1.0)


def test_sta_and_cga_squares():  # cayley.rs:262-305, registry.rs:402-428
    assert bp(1, 1, STA) == (0, 1.0)
    for e in (2, 4, 8):
        assert bp(e, e, STA) == (0, -1.0)
    (b1, s1), (b2, s2) = bp(1, 2, STA), bp(2, 1, STA)
    assert b1 == b2 and s1 == -s2
    for e in (1, 2, 4, 8):
        assert bp(e, e, CGA3) == (0, 1.0)
    assert bp(16, 16, CGA3) == (0, -1.0)


def test_cayley_table_r2():  # cayley.rs:311-326 and the hand-written table in src/simd.rs:389-408
    alg = oracle.OracleAlgebra(2)
    assert alg.num_blades == 4
    assert alg.cayley_blades.tolist() == [0, 1, 2, 3, 1, 0, 3, 2, 2, 3, 0, 1, 3, 2, 1, 0]
    assert alg.cayley_signs.tolist() == [1, 1, 1, 1, 1, 1, 1, 1, 1, -1, 1, -1, 1, -1, 1, -1]


@pytest.mark.parametrize("sig", [EUCLID3, STA])
def test_associativity(sig):  # cayley.rs:366-404
    alg = oracle.OracleAlgebra(*sig)
    nb = alg.num_blades
    prod = lambda a, b: (int(alg.cayley_blades[a * nb + b]), alg.cayley_signs[a * nb + b])
    ab, s_ab = prod(1, 2)
    l, s_l = prod(ab, 4)
    bc, s_bc = prod(2, 4)
    r, s_r = prod(1, bc)
    assert l == r and s_ab * s_l == s_bc * s_r


@pytest.mark.parametrize(
    "sig,squares",
    [((4, 0, 0), [1, 1, 1, 1]), (PGA3, [1, 1, 1, 0]), (STA, [1, -1, -1, -1]), ((3, 1, 0), [1, 1, 1, -1]),
     ((2, 1, 1), [1, 1, -1, 0]), ((2, 2, 1), [1, 1, -1, -1, 0])],
)
def test_all_basis_squares_and_anticommutativity(sig, squares):  # cayley.rs:410-475, registry.rs:574-630
    alg = oracle.OracleAlgebra(*sig)
    nb = alg.num_blades
    for i, sq in enumerate(squares):
        e = 1 << i
        assert alg.cayley_blades[e * nb + e] == 0
        assert alg.cayley_signs[e * nb + e] == sq
        assert oracle.basis_square(*sig, i) == sq  # signature.rs:256-299
    for i in range(alg.dimension):
        for j in range(i + 1, alg.dimension):
            ei, ej = 1 << i, 1 << j
            assert alg.cayley_blades[ei * nb + ej] == alg.cayley_blades[ej * nb + ei]
            assert alg.cayley_signs[ei * nb + ej] == -alg.cayley_signs[ej * nb + ei]


def test_pseudoscalar_squares():  # registry.rs:534-548
    for n, expect in ((2, -1.0), (3, -1.0), (4, 1.0)):
        alg = oracle.OracleAlgebra(n)
        ps = alg.num_blades - 1
        assert alg.cayley_signs[ps * alg.num_blades + ps] == expect


def test_involution_signs():  # blades.rs:277-309
    assert [oracle.reverse_sign(g) for g in range(8)] == [1, 1, -1, -1, 1, 1, -1, -1]
    assert [oracle.grade_involution_sign(g) for g in range(6)] == [1, -1, 1, -1, 1, -1]
    assert [oracle.clifford_conjugate_sign(g) for g in range(6)] == [1, -1, -1, 1, 1, -1]
    assert [oracle.blade_grade(i) for i in (0, 1, 2, 3, 4, 5, 6, 7, 15)] == [0, 1, 1, 2, 1, 2, 2, 3, 4]  # :245-255


def test_grade_masks():  # blades.rs:312-332, registry.rs:464-481
    assert [oracle.grade_mask(3, g) for g in range(4)] == [0b00000001, 0b00010110, 0b01101000, 0b10000000]
    assert oracle.OracleAlgebra(*PGA3).grade_masks() == [0x1, 0x116, 0x1668, 0x6880, 0x8000]
    assert oracle.OracleAlgebra(*CGA3).grade_masks() == [0x1, 0x10116, 0x1161668, 0x16686880, 0x68808000, 0x80000000]
    with pytest.raises(OverflowError):  # `1 << i` overflows usize in the reference for n >= 7
        oracle.grade_mask(7, 1)


DIGESTS = {  # sha256 of row-major int8 signs / little-endian u16 blades
    EUCLID3: ("40bc89779d69516cb07eb06fbac3ad3977c2993fa25d30549b8567d059478f9a", "7ee74828", 16),
    PGA3: ("28398685f6263fa53dd7fe21123656b681ccac926c017907649b14a9b0de7390", "4c3ae65e", 32),
    CGA3: ("9c0bc8a9265f229dec3ad5743e29060b3c6034f1049545a51b9bad13a9e881e4", "434a61cd", 64),
    STA: ("365efcddbaab4cb448d229369d7e213ee566ac8583e6c8382844b0a76e248c26", "4c3ae65e", 32),
    (8, 0, 0): ("b9593cd2fdc94cdee10612bc8911e0d7bf4045a354132ddf10b82bcd6474a660", "dd54e4d1", 512),
}


@pytest.mark.parametrize("sig", list(DIGESTS))
def test_table_digests(sig):
    alg = oracle.OracleAlgebra(*sig)
    signs_hex, blades_prefix, total = DIGESTS[sig]
    assert hashlib.sha256(alg.signs_int8().tobytes()).hexdigest() == signs_hex
    assert hashlib.sha256(alg.blades_u16().tobytes()).hexdigest().startswith(blades_prefix)
    assert int(alg.cayley_signs.sum()) == total
    nb = alg.num_blades
    idx = np.arange(nb)
    assert np.array_equal(alg.cayley_blades.reshape(nb, nb), idx[:, None] ^ idx[None, :])


def test_tables_match_reference_python_builders(golden):
    """Three independent builders of the reference agree with the oracle (tests/golden/make_golden.py)."""
    for sig in [(2, 0, 0), EUCLID3, PGA3, CGA3, STA, (2, 1, 1), (6, 0, 0)]:
        tag = "cl%d%d%d" % sig
        alg = oracle.OracleAlgebra(*sig)
        assert np.array_equal(golden[tag + ".signs"].ravel(), alg.signs_int8())
        assert np.array_equal(golden[tag + ".blades"].ravel(), alg.blades_u16())
        if tag + ".codegen_signs" in golden:
            assert np.array_equal(golden[tag + ".codegen_signs"].ravel(), alg.signs_int8())
            assert np.array_equal(golden[tag + ".codegen_blades"].ravel(), alg.blades_u16())


# ----------------------------------------------------------------------------- rest of the product family
def _vec3(x, y, z):
    v = np.zeros((1, 8))
    v[0, [1, 2, 4]] = (x, y, z)
    return v


def _blade(nb, index, value=1.0):
    v = np.zeros((1, nb))
    v[0, index] = value
    return v


def test_product_family_known_answers_of_the_reference_tests():
    """Pins oracle ops 3/4 and the composed commutator / anticommutator / regressive to the known answers
    of the reference's own scalar tests (tests/test_basic.py:2275-2316 right contraction, :2345-2402 scalar
    product, :2407-2520 commutator / anticommutator, :2528-2557 regressive)."""
    o = oracle.OracleAlgebra(3, 0, 0)
    a, b = _vec3(1, 2, 3), _vec3(4, 5, 6)
    rc = o.right_contraction(a, b)
    assert rc[0, 0] == 32.0 and not rc[0, 1:].any()
    e12, e13, e1, e2 = _blade(8, 3), _blade(8, 5), _blade(8, 1), _blade(8, 2)
    assert np.array_equal(o.right_contraction(e12, e1), -e2)              # e12 |_ e1 = -e2
    assert not o.right_contraction(e1, e12).any()                         # grade(A) < grade(B) => 0
    assert np.array_equal(o.left_contraction(e1, e12), e2)
    sp = o.scalar_product(e12, e12)
    assert sp[0, 0] == -1.0 and not sp[0, 1:].any()
    assert o.scalar_product(a, b)[0, 0] == 32.0
    assert not o.scalar_product(e1, e12).any()
    o2 = oracle.OracleAlgebra(2, 0, 0)
    f1, f2 = _blade(4, 1), _blade(4, 2)
    assert np.array_equal(o2.commutator(f1, f2), _blade(4, 3))            # [e1, e2] = e12
    assert not o2.commutator(f1, f1).any()
    assert not o2.anticommutator(f1, f2).any()
    assert np.array_equal(o2.anticommutator(f1, f1), _blade(4, 0))
    rng = np.random.default_rng(5)
    x, y = rng.standard_normal((7, 8)), rng.standard_normal((7, 8))
    assert np.allclose(o.commutator(x, y), -o.commutator(y, x), atol=1e-15)
    assert np.allclose(o.commutator(x, y) + o.anticommutator(x, y), o.geometric_product(x, y), atol=1e-14)
    ps = _blade(8, 7)
    assert np.allclose(o.regressive(ps, a), a, atol=1e-15)                # I v A = A
    meet = o.regressive(e12, e13)                                         # two planes meet in the e1 axis
    assert abs(meet[0, 1]) == 1.0 and not np.delete(meet[0], 1).any()
    with pytest.raises(ValueError):                                       # null pseudoscalar: dual() raises
        oracle.OracleAlgebra(3, 0, 1).regressive(np.ones((1, 16)), np.ones((1, 16)))
