"""Pin the oracle's floating-point loops against the reference's own outputs and KATs.

1. tests/golden/ref_torch_golden.npz -- produced by importing the reference's
   lcc/torch/multivector.py (tests/golden/make_golden.py); products must match BIT FOR BIT.
2. The derived Criterion-vector golden of SURVEY.md section 8c.
3. Closed-form KATs of /root/reference/tests/test_batch.py and test_pga.py replayed on the oracle.
"""
import math

import numpy as np
import pytest

import oracle

SIGS = [(2, 0, 0), (3, 0, 0), (3, 0, 1), (4, 1, 0), (1, 3, 0), (2, 1, 1)]


def ulp_close(a, b, ulps):
    a, b = np.asarray(a), np.asarray(b)
    eps = np.finfo(a.dtype).eps
    return np.all(np.abs(a - b) <= ulps * eps * np.maximum(np.abs(a), np.abs(b)))


@pytest.mark.parametrize("sig", SIGS)
@pytest.mark.parametrize("dt", ["f64", "f32"])
def test_bit_exact_against_reference_torch(golden, sig, dt):
    alg = oracle.OracleAlgebra(*sig)
    k = "cl%d%d%d.%s" % (*sig, dt)
    a, b, rotor = golden[k + ".a"], golden[k + ".b"], golden[k + ".rotor"]
    exact = {
        "gp": alg.geometric_product(a, b),
        "gp_bcast": alg.geometric_product(a, b[:1]),
        "outer": alg.outer_product(a, b),
        "inner": alg.left_contraction(a, b),
        "sandwich": alg.sandwich(rotor[0].astype(a.dtype), a),
        "reverse": alg.reverse(a),
        "involution": alg.grade_involution(a),
        "conjugate": alg.clifford_conjugate(a),
        "norm_squared": alg.norm_squared(a),
    }
    for g in range(alg.dimension + 1):
        exact[f"grade{g}"] = alg.grade(a, g)
    for name, mine in exact.items():
        assert np.array_equal(golden[f"{k}.{name}"], mine), name
    # torch's vectorised CPU sqrt is not correctly rounded (1 ulp off on some inputs); Rust's
    # f64::sqrt and C's sqrt are.  norm / normalized are therefore pinned to 2 ulp here and
    # bit-exactly to IEEE sqrt below.
    assert ulp_close(golden[k + ".norm"], alg.norm(a), 2)
    assert ulp_close(golden[k + ".normalized"], alg.normalized(a), 4)
    n2 = alg.norm_squared(a)
    assert np.array_equal(alg.norm(a), np.sqrt(np.abs(n2)))


def criterion_vector(dims, seed):  # benches/geometric_product.rs:21-32 (usize arithmetic, then as f64)
    out = []
    for i in range(1 << dims):
        v = ((i + seed) * 2654435761) & 0xFFFFFFFFFFFFFFFF
        out.append(math.fmod(float(v), 100.0) / 50.0 - 1.0)
    return np.array(out)


def test_criterion_vectors_r3():
    alg = oracle.OracleAlgebra(3)
    got = alg.geometric_product(criterion_vector(3, 42)[None], criterion_vector(3, 137)[None])[0]
    want = [
        "0x1.886594af4f0d8p-1", "0x1.c9eecbfb15b55p-2", "0x1.4e3bcd35a8587p-1", "-0x1.398c7e28240b7p+0",
        "-0x1.4467381d7dbf2p-3", "-0x1.c154c985f06f7p-1", "-0x1.41f212d77318ep+0", "0x1.84b5dcc63f13fp-1",
    ]
    assert [float(x).hex() for x in got] == want


def axis_angle_rotor(axis, angle):  # src/multivector.rs:955-999
    ax = np.asarray(axis, float)
    ax = ax / np.linalg.norm(ax)
    c, s = math.cos(angle / 2), math.sin(angle / 2)
    r = np.zeros(8)
    r[0], r[3], r[5], r[6] = c, -s * ax[2], s * ax[1], -s * ax[0]
    return r


def test_batch_kats_r3():
    """Closed forms of /root/reference/tests/test_batch.py replayed on the oracle."""
    alg = oracle.OracleAlgebra(3)
    rng = np.random.default_rng(0)
    a, b = rng.standard_normal((50, 3)), rng.standard_normal((50, 3))
    A, B = alg.from_vectors(a), alg.from_vectors(b)
    assert np.array_equal(alg.to_vector_coords(A), a)  # :36-45
    ident = np.zeros(8)
    ident[0] = 1.0
    np.testing.assert_allclose(alg.to_vector_coords(alg.sandwich(ident, A)), a, rtol=1e-10)  # :166-178
    rot = alg.sandwich(axis_angle_rotor([0, 0, 1], math.pi / 4), A)
    np.testing.assert_allclose(alg.norm(rot), alg.norm(A), rtol=1e-10)  # :180-195
    e1 = alg.from_vectors(np.array([[1.0, 0.0, 0.0]]))
    r90 = alg.sandwich(axis_angle_rotor([0, 0, 1], math.pi / 2), e1)
    np.testing.assert_allclose(alg.to_vector_coords(r90), [[0.0, 1.0, 0.0]], atol=1e-10)  # :197-212
    two = np.zeros((1, 8))
    two[0, 0] = 2.0
    np.testing.assert_allclose(alg.to_vector_coords(alg.geometric_product(A, two)), a * 2.0, rtol=1e-10)  # :267-281
    with pytest.raises(ValueError, match="batch sizes must match"):  # :283-290
        alg.geometric_product(A[:10], B[:5])
    w = alg.outer_product(e1, alg.from_vectors(np.array([[0.0, 1.0, 0.0]])))
    assert abs(w[0, 3] - 1.0) < 1e-10  # :300-320
    np.testing.assert_allclose(alg.outer_product(A, B), -alg.outer_product(B, A), rtol=1e-10)  # :322-336
    np.testing.assert_allclose(alg.left_contraction(A, B)[:, 0], np.sum(a * b, axis=1), rtol=1e-10)  # :345-365
    np.testing.assert_allclose(alg.norm(A), np.linalg.norm(a, axis=1), rtol=1e-10)  # :375-386
    np.testing.assert_allclose(alg.norm_squared(A), np.sum(a**2, axis=1), rtol=1e-10)  # :388-398
    np.testing.assert_allclose(alg.norm(alg.normalized(A)), np.ones(50), rtol=1e-10)  # :400-410
    np.testing.assert_allclose(alg.to_vector_coords(alg.reverse(A)), a, rtol=1e-10)  # :420-430
    biv = alg.from_bivectors(rng.standard_normal((20, 3)))
    np.testing.assert_allclose(alg.reverse(biv)[:, [3, 5, 6]], -biv[:, [3, 5, 6]], rtol=1e-10)  # :432-447
    np.testing.assert_allclose(alg.to_vector_coords(alg.grade_involution(A)), -a, rtol=1e-10)  # :449-460
    np.testing.assert_allclose(alg.to_vector_coords(alg.add(A, B)), a + b, rtol=1e-10)  # :470-495
    np.testing.assert_allclose(alg.to_vector_coords(alg.sub(A, B)), a - b, rtol=1e-10)
    np.testing.assert_allclose(alg.to_vector_coords(alg.div(A, 2.0)), a / 2.0, rtol=1e-10)  # :520-533
    with pytest.raises(ZeroDivisionError):
        alg.div(A, 0.0)
    mixed = rng.standard_normal((10, 8))
    g1 = alg.grade(mixed, 1)
    assert np.all(g1[:, [0, 3, 5, 6, 7]] == 0) and np.array_equal(g1[:, [1, 2, 4]], mixed[:, [1, 2, 4]])  # :579-597
    with pytest.raises(ValueError, match="exceeds dimension"):  # src/batch.rs:1029-1033
        alg.grade(mixed, 4)


def pga_point(x, y, z):  # src/pga.rs:131-134
    c = np.zeros(16)
    c[7], c[14], c[13], c[11] = 1.0, -x, y, -z
    return c


def pga_point_coords(c):  # src/pga.rs:579-599
    w = c[7]
    return (-c[14] / w, c[13] / w, -c[11] / w)


def pga_translator(dx, dy, dz):  # src/pga.rs:386-389
    c = np.zeros(16)
    c[0], c[9], c[10], c[12] = 1.0, dx / 2, dy / 2, dz / 2
    return c


def pga_motor(axis, point, angle):  # src/pga.rs:287-345
    ax = np.asarray(axis, float) / np.linalg.norm(axis)
    p = np.asarray(point, float)
    c, s = math.cos(angle / 2), math.sin(angle / 2)
    m = np.zeros(16)
    m[0], m[6], m[5], m[3] = c, -s * ax[0], s * ax[1], -s * ax[2]
    mom = np.cross(p, ax)
    m[9], m[10], m[12] = s * mom[0], s * mom[1], s * mom[2]
    return m


def test_pga_kats():
    """/root/reference/tests/test_pga.py: translate (1,2,3) by (4,5,6) -> (5,7,9) (:245-255);
    90 degree motor about z maps (1,0,0) -> (0,1,0) (:287-300)."""
    alg = oracle.OracleAlgebra(3, 0, 1)
    moved = alg.sandwich(pga_translator(4, 5, 6), pga_point(1, 2, 3)[None])[0]
    np.testing.assert_allclose(pga_point_coords(moved), (5.0, 7.0, 9.0), atol=1e-12)
    rot = alg.sandwich(pga_motor([0, 0, 1], [0, 0, 0], math.pi / 2), pga_point(1, 0, 0)[None])[0]
    np.testing.assert_allclose(pga_point_coords(rot), (0.0, 1.0, 0.0), atol=1e-12)
    # pga_dual is an involution up to sign and swaps complements (src/pga.rs:534-549)
    x = np.random.default_rng(1).standard_normal((4, 16))
    d = alg.pga_dual(x)
    assert np.array_equal(np.abs(d[:, ::-1]), np.abs(x))
    with pytest.raises(ValueError):  # I^2 = 0: Multivector::dual raises (multivector.rs:2991-2996)
        alg.dual(x)


def test_dual_is_signed_permutation():
    """out[k^ps] = S[k][ps]*S[ps][ps]*a[k]; R3 per-k signs [-1,-1,1,1,-1,-1,1,1]."""
    alg = oracle.OracleAlgebra(3)
    x = np.random.default_rng(2).standard_normal((5, 8))
    d = alg.dual(x)
    signs = np.array([-1, -1, 1, 1, -1, -1, 1, 1.0])
    want = np.zeros_like(x)
    for k in range(8):
        want[:, k ^ 7] = signs[k] * x[:, k]
    assert np.array_equal(d, want)
    assert np.array_equal(alg.undual(d), x)
    for sig in [(4, 1, 0), (1, 3, 0), (4, 0, 0)]:
        alg = oracle.OracleAlgebra(*sig)
        x = np.random.default_rng(3).standard_normal((3, alg.num_blades))
        assert np.array_equal(alg.undual(alg.dual(x)), x)


def test_simd_path_equals_scalar_path():  # src/simd.rs:55-135 vs src/batch.rs:384-404
    for sig in [(3, 0, 1), (4, 1, 0), (1, 3, 0)]:
        alg = oracle.OracleAlgebra(*sig)
        rng = np.random.default_rng(5)
        a, b = rng.standard_normal(alg.num_blades), rng.standard_normal(alg.num_blades)
        assert np.array_equal(alg.simd_geometric_product(a, b), alg.geometric_product(a[None], b[None])[0])


def test_threads_do_not_change_results():
    alg = oracle.OracleAlgebra(4, 1, 0)
    rng = np.random.default_rng(7)
    a, b = rng.standard_normal((257, 32)), rng.standard_normal((257, 32))
    assert np.array_equal(alg.geometric_product(a, b, threads=1), alg.geometric_product(a, b, threads=4))
    r = rng.standard_normal(32)
    assert np.array_equal(alg.sandwich(r, a, threads=1), alg.sandwich(r, a, threads=3))


# ----------------------------------------------------------------------------------------------- CGA / STA known answers
def _cga_point(x, y, z):
    """Multivector::cga_point, src/cga.rs:171-205: x e1 + y e2 + z e3 + (|x|^2/2 - 1/2) e+ + (|x|^2/2 + 1/2) e-,
    e+ = blade 8, e- = blade 16 (cga.rs:45-53)."""
    c = np.zeros(32)
    c[1], c[2], c[4] = x, y, z
    n2 = x * x + y * y + z * z
    c[8], c[16] = 0.5 * n2 - 0.5, 0.5 * n2 + 0.5
    return c


def test_cga_kats():
    """Known answers of the reference's tests/test_cga.py (35-61 null basis, 249-319 incidence and distances) and
    src/cga.rs:734-851, evaluated with the oracle's products on the reference's coefficient layouts."""
    o = oracle.OracleAlgebra(4, 1, 0)
    gp = lambda a, b: o.geometric_product(a[None, :], b[None, :])[0]  # noqa: E731
    dot = lambda a, b: gp(a, b)[0]  # noqa: E731  (two vectors: the scalar part of the product is their inner product)
    eo, einf = np.zeros(32), np.zeros(32)
    eo[8], eo[16] = -0.5, 0.5      # e_o = (e- - e+) / 2, cga.rs:92-97
    einf[8], einf[16] = 1.0, 1.0   # e_inf = e- + e+, cga.rs:135-140
    assert dot(eo, eo) == 0.0 and dot(einf, einf) == 0.0 and dot(eo, einf) == -1.0
    assert np.array_equal(_cga_point(0.0, 0.0, 0.0), eo)
    for p in ((1.0, 2.0, 3.0), (-0.5, 4.0, 0.25), (3.0, 4.0, 0.0)):
        assert abs(dot(_cga_point(*p), _cga_point(*p))) < 1e-12            # conformal points are null
        assert abs(dot(_cga_point(*p), einf) + 1.0) < 1e-12                # normalised: P . e_inf = -1
    dist = lambda p, q: np.sqrt(-2.0 * dot(_cga_point(*p), _cga_point(*q)))  # noqa: E731  (cga_distance, cga.rs)
    assert abs(dist((0.0, 0.0, 0.0), (3.0, 4.0, 0.0)) - 5.0) < 1e-10       # test_cga.py:283-290
    assert abs(dist((1.0, 2.0, 3.0), (4.0, 6.0, 3.0)) - 5.0) < 1e-10       # :292-300
    assert abs(dist((1.0, 2.0, 3.0), (4.0, 5.0, 6.0)) - dist((4.0, 5.0, 6.0), (1.0, 2.0, 3.0))) < 1e-12
    sphere = _cga_point(0.0, 0.0, 0.0) - 0.5 * einf                        # unit sphere, cga.rs:294-309 (S = P - r^2/2 e_inf)
    assert abs(dot(_cga_point(1.0, 0.0, 0.0), sphere)) < 1e-12 and abs(dot(_cga_point(0.0, 1.0, 0.0), sphere)) < 1e-12
    assert dot(_cga_point(0.5, 0.0, 0.0), sphere) > 1e-3 and dot(_cga_point(2.0, 0.0, 0.0), sphere) < -1e-3
    # a batch of points against one sphere through the broadcast rule of batch.rs:361-376: P . S = (r^2 - |p - c|^2) / 2
    rng = np.random.default_rng(5)
    pts = rng.standard_normal((257, 3))
    P = np.stack([_cga_point(*p) for p in pts])
    S = (_cga_point(1.0, 2.0, 3.0) - 0.5 * 25.0 * einf)[None, :]
    got = o.geometric_product(P, S)[:, 0]
    want = 0.5 * (25.0 - ((pts - np.array([1.0, 2.0, 3.0])) ** 2).sum(axis=1))
    np.testing.assert_allclose(got, want, rtol=0, atol=1e-12)


def test_sta_kats():
    """Known answers of the reference's tests/test_sta.py (54-160 metric, anticommutation, pseudoscalar; 189-436
    intervals, boosts, rotations, field bivectors) and src/sta.rs:675-934 with the oracle's products, sandwich and
    the row-wise boost / bivector constructors.  gamma_mu = blades 1, 2, 4, 8 (sta.rs:34-37)."""
    o = oracle.OracleAlgebra(1, 3, 0)
    gp = lambda a, b: o.geometric_product(a[None, :], b[None, :])[0]  # noqa: E731
    g = [np.eye(16)[i] for i in (1, 2, 4, 8)]
    eta = np.diag([1.0, -1.0, -1.0, -1.0])
    for mu in range(4):
        for nu in range(4):
            anti = gp(g[mu], g[nu]) + gp(g[nu], g[mu])
            assert anti[0] == 2.0 * eta[mu, nu] and not anti[1:].any()       # {gamma_mu, gamma_nu} = 2 eta, test_sta.py:109-127
    I = gp(gp(g[0], g[1]), gp(g[2], g[3]))
    assert I[15] == 1.0 and np.count_nonzero(I) == 1 and gp(I, I)[0] == -1.0  # test_sta.py:133-141

    def vec(t, x, y, z):  # sta_vector, sta.rs:105-109
        c = np.zeros(16)
        c[1], c[2], c[4], c[8] = t, x, y, z
        return c

    interval = lambda v: gp(v, v)[0]  # noqa: E731
    assert interval(vec(5.0, 3.0, 0.0, 0.0)) == 16.0 and interval(vec(1.0, 5.0, 0.0, 0.0)) == -24.0   # test_sta.py:190-213
    assert interval(vec(1.0, 1.0, 0.0, 0.0)) == 0.0 and abs(interval(vec(np.sqrt(3.0), 1.0, 1.0, 1.0))) < 1e-15
    # boosts (sta.rs:253-296): gamma = 2 s^2 - 1 (sta.rs:641-653), intervals are preserved by R x ~R
    R = oracle.sta_boosts(np.array([[0.6, 0.0, 0.0], [0.8, 0.0, 0.0], [0.0, 0.6, 0.0], [0.0, 0.0, 0.0], [0.5, 0.0, 0.0]]))
    np.testing.assert_allclose(2.0 * R[:4, 0] ** 2 - 1.0, [1.25, 5.0 / 3.0, 1.25, 1.0], rtol=1e-14)    # test_sta.py:247-291
    assert np.array_equal(R[3], np.eye(16)[0])
    x = vec(5.0, 3.0, 0.0, 0.0)
    xb = o.sandwich(R[4], x[None, :])[0]
    assert abs(interval(xb) - 16.0) < 1e-12                                                             # test_sta.py:230-245
    np.testing.assert_allclose(np.delete(xb, [1, 2, 4, 8]), 0.0, atol=1e-14)                           # stays a vector
    # a boost of the rest observer gamma_0 by v has time component gamma(v)
    np.testing.assert_allclose(o.sandwich(R[0], g[0][None, :])[0][1], 1.25, rtol=1e-14)
    # composition of two collinear boosts: rapidities add, so gamma grows (test_sta.py:262-271); (0.3 + 0.3)/(1 + 0.09)
    R1 = oracle.sta_boosts(np.array([[0.3, 0.0, 0.0]]))[0]
    R12 = gp(R1, R1)
    v12 = 0.6 / 1.09
    np.testing.assert_allclose(2.0 * R12[0] ** 2 - 1.0, 1.0 / np.sqrt(1.0 - v12 * v12), rtol=1e-13)
    # spatial rotations (sta.rs:329-376): R = cos(a/2) + sin(a/2) (nx g23 - ny g13 + nz g12), blades 12, 10, 6
    def rotation(axis, angle):
        n = np.asarray(axis, float) / np.linalg.norm(axis)
        c = np.zeros(16)
        c[0], c[12], c[10], c[6] = np.cos(angle / 2), np.sin(angle / 2) * n[0], -np.sin(angle / 2) * n[1], np.sin(angle / 2) * n[2]
        return c
    xr = o.sandwich(rotation([0.0, 0.0, 1.0], np.pi / 2), vec(0.0, 1.0, 0.0, 0.0)[None, :])[0]
    np.testing.assert_allclose(xr[[1, 2, 4, 8]], [0.0, 0.0, 1.0, 0.0], atol=1e-15)                     # x -> y, test_sta.py:343-353
    xr = o.sandwich(rotation([1.0, 0.0, 0.0], np.pi / 2), vec(0.0, 0.0, 1.0, 0.0)[None, :])[0]
    np.testing.assert_allclose(xr[[1, 2, 4, 8]], [0.0, 0.0, 0.0, 1.0], atol=1e-15)                     # y -> z, :355-364
    xr = o.sandwich(rotation([0.0, 0.0, 1.0], np.pi / 4), vec(5.0, 3.0, 4.0, 0.0)[None, :])[0]
    assert abs(interval(xr) - 0.0) < 1e-12 and abs(xr[1] - 5.0) < 1e-14                                # interval and time are invariant
    # field bivector F = E + I B (sta.rs:202-212): the two Lorentz invariants F^2 = (E^2 - B^2) + 2 (E.B) I
    E, B = np.array([1.0, 2.0, 3.0]), np.array([0.5, -0.5, 1.0])
    F = oracle.sta_bivectors(E[None, :], B[None, :])[0]
    F2 = gp(F, F)
    np.testing.assert_allclose([F2[0], F2[15]], [E @ E - B @ B, 2.0 * (E @ B)], rtol=1e-14)
    np.testing.assert_allclose(np.delete(F2, [0, 15]), 0.0, atol=1e-15)
