"""PGA / CGA / STA specialisations of the `largecrimsoncanine` extension: the geometric known answers
the reference pins in tests/test_pga.py, test_cga.py, test_sta.py and test_algebras.py (cited per
test), written against the same public names.  Every product below is executed by the CUDA kernels
(one-row device calls), so these are GPU tests; the pure constructor / extractor checks run on CPU.
"""
import math

import numpy as np
import pytest

import oracle


@pytest.fixture(scope="module")
def api():
    import largecrimsoncanine as L

    return L


# ----------------------------------------------------------------------------- CPU: constructors and layouts
def test_pga_layout_constants_and_constructors(api):  # src/pga.rs:50-78,131-134,176-179,386-389; tests/test_pga.py:35-120
    A, M = api.Algebra, api.Multivector
    pga = A.pga(3)
    p = M.pga_point(pga, 1.0, 2.0, 3.0).coefficients()
    assert (p[7], p[14], p[13], p[11]) == (1.0, -1.0, 2.0, -3.0) and sum(abs(c) for c in p) == 7.0
    pl = M.pga_plane(pga, 1.0, 2.0, 3.0, 4.0).coefficients()
    assert (pl[1], pl[2], pl[4], pl[8]) == (1.0, 2.0, 3.0, 4.0)
    t = M.pga_translator(pga, 2.0, 4.0, 6.0).coefficients()
    assert (t[0], t[9], t[10], t[12]) == (1.0, 1.0, 2.0, 3.0)
    m = M.pga_motor_from_axis_angle(pga, [0.0, 0.0, 2.0], [1.0, 0.0, 0.0], math.pi / 2).coefficients()
    s = math.sin(math.pi / 4)
    np.testing.assert_allclose([m[0], m[3], m[5], m[6], m[9], m[10], m[12]], [math.cos(math.pi / 4), -s, 0, 0, 0, -s, 0], atol=1e-15)
    for x, y, z in [(0.0, 0.0, 0.0), (-1.0, 2.5, -3.7), (100.0, -200.0, 300.0), (0.001, 0.002, 0.003)]:  # test_pga.py:517-528
        assert M.pga_point(pga, x, y, z).pga_point_coords() == pytest.approx((x, y, z))
    scaled = M.pga_point(pga, 1.0, 2.0, 3.0) * 5.0
    assert scaled.pga_normalize().pga_point_coords() == pytest.approx((1.0, 2.0, 3.0))
    plane = M.pga_plane(pga, 1.0, 0.0, 0.0, 2.0)
    assert plane.pga_ideal_part().coefficients()[8] == 2.0 and plane.pga_ideal_part().coefficients()[1] == 0.0
    assert plane.pga_euclidean_part().coefficients()[1] == 1.0 and plane.pga_euclidean_part().coefficients()[8] == 0.0
    assert M.pga_point(pga, 1.0, 2.0, 3.0).pga_point_plane_distance(M.pga_plane(pga, 0.0, 0.0, 1.0, -1.0)) == pytest.approx(2.0)
    with pytest.raises(ValueError, match="ideal point"):
        M.zero_in(pga).pga_point_coords()
    with pytest.raises(ValueError, match="PGA3D"):
        M.pga_translator(A.euclidean(3), 1.0, 0.0, 0.0)
    d = oracle.OracleAlgebra(3, 0, 1).pga_dual(np.array([M.pga_point(pga, 1.0, 2.0, 3.0).coefficients()]))[0]
    assert M.pga_point(pga, 1.0, 2.0, 3.0).pga_dual().coefficients() == list(d)


def test_cga_and_sta_constructors(api):  # src/cga.rs:92-97,135-140,189-206,304-306,351-360; src/sta.rs:105-109,202-212,253-296,361-369
    A, M = api.Algebra, api.Multivector
    cga, sta = A.cga(3), A.sta()
    eo, einf = M.cga_eo(cga).coefficients(), M.cga_einf(cga).coefficients()
    assert (eo[8], eo[16], einf[8], einf[16]) == (-0.5, 0.5, 1.0, 1.0)
    p = M.cga_point(cga, 1.0, 2.0, 3.0).coefficients()
    assert (p[1], p[2], p[4], p[8], p[16]) == (1.0, 2.0, 3.0, 6.5, 7.5)
    assert M.cga_point_from_coords(cga, [1.0, 2.0, 3.0]) == M.cga_point(cga, 1.0, 2.0, 3.0)
    assert M.cga_point_from_coords(A.cga(2), [3.0, 4.0]).coefficients()[4] == 12.0  # e+ of CGA2D is blade 4
    s = M.cga_sphere(cga, 1.0, 2.0, 3.0, 2.0).coefficients()
    assert (s[8], s[16]) == (4.5, 5.5)
    pl = M.cga_plane(cga, 0.0, 0.0, 1.0, 5.0).coefficients()
    assert (pl[4], pl[8], pl[16]) == (1.0, 5.0, 5.0)
    assert M.cga_sphere(cga, 1.0, 2.0, 3.0, 5.0).cga_extract_center() == pytest.approx((1.0, 2.0, 3.0))
    with pytest.raises(ValueError, match="flat object"):
        M.cga_plane(cga, 0.0, 0.0, 1.0, 5.0).cga_extract_center()
    with pytest.raises(ValueError, match="CGA"):
        M.cga_point(A.pga(3), 1.0, 2.0, 3.0)
    v = M.sta_vector(sta, 5.0, 3.0, 4.0, 0.0)
    assert v.sta_interval_squared() == 0.0 and v.sta_is_lightlike() and not v.sta_is_timelike()  # src/sta.rs:500-568
    assert M.sta_vector(sta, 2.0, 1.0, 0.0, 0.0).sta_is_timelike() and M.sta_vector(sta, 1.0, 2.0, 0.0, 0.0).sta_is_spacelike()
    assert M.sta_vector(sta, 5.0, 3.0, 0.0, 0.0).sta_proper_time() == pytest.approx(4.0)
    assert v.sta_spacetime_split() == (5.0, [3.0, 4.0, 0.0])
    assert [M.sta_gamma(sta, i).coefficients().index(1.0) for i in range(4)] == [1, 2, 4, 8]
    with pytest.raises(ValueError, match="0-3"):
        M.sta_gamma(sta, 4)
    F = M.sta_bivector(sta, [1.0, 2.0, 3.0], [0.5, -0.5, 1.0])
    assert F.sta_field_decompose() == ([1.0, 2.0, 3.0], [0.5, -0.5, 1.0]) and F.pure_grade() == 2
    assert M.sta_boost(sta, [0.6, 0.0, 0.0]).sta_lorentz_gamma() == pytest.approx(1.25)  # tests/test_sta.py:254-266
    assert M.sta_boost(sta, [0.0, 0.8, 0.0]).sta_lorentz_gamma() == pytest.approx(5.0 / 3.0)
    assert M.sta_boost(sta, [0.0, 0.0, 0.0]).coefficients() == [1.0] + [0.0] * 15
    with pytest.raises(ValueError, match="speed of light"):
        M.sta_boost(sta, [1.0, 0.0, 0.0])
    r = M.sta_rotation(sta, [0.0, 0.0, 2.0], math.pi / 2).coefficients()
    assert r[0] == pytest.approx(math.cos(math.pi / 4)) and r[6] == pytest.approx(math.sin(math.pi / 4))
    assert M.sta_I(sta).coefficients()[15] == 1.0
    with pytest.raises(ValueError, match=r"Cl\(1,3,0\)"):
        M.sta_vector(A.euclidean(4), 1.0, 0.0, 0.0, 0.0)


# ----------------------------------------------------------------------------- GPU: products behind the KATs
gpu = pytest.mark.gpu


@gpu
def test_pga_motors(api):  # tests/test_pga.py:225-425, 558-574
    A, M = api.Algebra, api.Multivector
    pga = A.pga(3)
    T = M.pga_translator(pga, 4.0, 5.0, 6.0)
    assert T.is_pga_motor() and not M.pga_plane(pga, 1.0, 0.0, 0.0, 0.0).is_pga_motor() and not M.zero_in(pga).is_pga_motor()
    assert T.sandwich(M.pga_point(pga, 1.0, 2.0, 3.0)).pga_point_coords() == pytest.approx((5.0, 7.0, 9.0))
    z90 = M.pga_motor_from_axis_angle(pga, [0.0, 0.0, 1.0], [0.0, 0.0, 0.0], math.pi / 2)
    assert z90.is_pga_motor()
    assert z90.sandwich(M.pga_point(pga, 1.0, 0.0, 0.0)).pga_point_coords() == pytest.approx((0.0, 1.0, 0.0), abs=1e-10)
    z180 = M.pga_motor_from_axis_angle(pga, [0.0, 0.0, 1.0], [0.0, 0.0, 0.0], math.pi)
    assert z180.sandwich(M.pga_point(pga, 1.0, 0.0, 0.0)).pga_point_coords() == pytest.approx((-1.0, 0.0, 0.0), abs=1e-10)
    x90 = M.pga_motor_from_axis_angle(pga, [1.0, 0.0, 0.0], [0.0, 0.0, 0.0], math.pi / 2)
    assert x90.sandwich(M.pga_point(pga, 0.0, 1.0, 0.0)).pga_point_coords() == pytest.approx((0.0, 0.0, 1.0), abs=1e-10)
    ident = M.pga_motor_from_axis_angle(pga, [1.0, 0.0, 0.0], [0.0, 0.0, 0.0], 0.0)
    assert ident.sandwich(M.pga_point(pga, 5.0, 6.0, 7.0)).pga_point_coords() == pytest.approx((5.0, 6.0, 7.0))
    # composition: products keep the algebra (multivector.rs:1476-1480)
    T1, T2 = M.pga_translator(pga, 1.0, 0.0, 0.0), M.pga_translator(pga, 0.0, 2.0, 0.0)
    assert (T2 * T1).algebra() == pga
    assert (T2 * T1).sandwich(M.pga_point(pga, 0.0, 0.0, 0.0)).pga_point_coords() == pytest.approx((1.0, 2.0, 0.0))
    assert (z90 * z90).sandwich(M.pga_point(pga, 1.0, 0.0, 0.0)).pga_point_coords() == pytest.approx((-1.0, 0.0, 0.0), abs=1e-10)
    TR = M.pga_translator(pga, 1.0, 0.0, 0.0) * z90
    assert TR.sandwich(M.pga_point(pga, 1.0, 0.0, 0.0)).pga_point_coords() == pytest.approx((1.0, 1.0, 0.0), abs=1e-10)
    # rotation about an offset axis keeps points on the axis fixed
    off = M.pga_motor_from_axis_angle(pga, [0.0, 0.0, 1.0], [1.0, 1.0, 0.0], 1.234)
    assert off.sandwich(M.pga_point(pga, 1.0, 1.0, 5.0)).pga_point_coords() == pytest.approx((1.0, 1.0, 5.0), abs=1e-10)
    # join of two points / meet of two planes describe the same line (tests/test_pga.py:530-556)
    l1 = M.pga_line_from_points(M.pga_point(pga, 0.0, 0.0, 0.0), M.pga_point(pga, 1.0, 0.0, 0.0))
    l2 = M.pga_line_from_planes(M.pga_plane(pga, 0.0, 1.0, 0.0, 0.0), M.pga_plane(pga, 0.0, 0.0, 1.0, 0.0))
    assert l1.pure_grade() == 2 and l2.pure_grade() == 2
    a, b = l1 * (1.0 / l1.norm()), l2 * (1.0 / l2.norm())
    assert min((a - b).norm(), (a + b).norm()) < 0.1
    # double reflection in a plane is the identity (tests/test_pga.py:558-574)
    p, xy = M.pga_point(pga, 1.0, 2.0, 3.0), M.pga_plane(pga, 0.0, 0.0, 1.0, 0.0)
    twice = -xy * (-xy * p * xy) * xy
    assert twice.pga_normalize().pga_point_coords() == pytest.approx((1.0, 2.0, 3.0), abs=1e-10)


@gpu
def test_cga_geometry(api):  # tests/test_cga.py:35-61,125-141,249-319,386-406
    A, M = api.Algebra, api.Multivector
    cga = A.cga(3)
    eo, einf = M.cga_eo(cga), M.cga_einf(cga)
    assert (eo * eo).scalar() == pytest.approx(0.0) and (einf * einf).scalar() == pytest.approx(0.0)  # null vectors
    assert eo.inner(einf).scalar() == pytest.approx(-1.0)
    P = M.cga_point(cga, 1.0, 2.0, 3.0)
    assert (P * P).scalar() == pytest.approx(0.0, abs=1e-12)
    S = M.cga_sphere(cga, 0.0, 0.0, 0.0, 1.0)
    assert M.cga_point(cga, 1.0, 0.0, 0.0).cga_point_on_sphere(S) and M.cga_point(cga, 0.0, 1.0, 0.0).cga_point_on_sphere(S)
    assert not M.cga_point(cga, 0.5, 0.0, 0.0).cga_point_on_sphere(S) and not M.cga_point(cga, 2.0, 0.0, 0.0).cga_point_on_sphere(S)
    assert M.cga_sphere(cga, 1.0, 2.0, 3.0, 2.5).cga_extract_radius() == pytest.approx(2.5)
    assert S.cga_is_round() and not S.cga_is_flat() and M.cga_plane(cga, 0.0, 0.0, 1.0, 0.0).cga_is_flat()
    assert abs(M.cga_point(cga, 0.0, 0.0, 0.0).cga_distance(M.cga_point(cga, 3.0, 4.0, 0.0)) - 5.0) < 1e-10
    assert abs(M.cga_point(cga, 1.0, 2.0, 3.0).cga_distance(M.cga_point(cga, 4.0, 6.0, 3.0)) - 5.0) < 1e-10
    assert abs(P.cga_distance(P)) < 1e-10
    line = M.cga_line(M.cga_plane(cga, 0.0, 0.0, 1.0, 0.0), M.cga_plane(cga, 0.0, 1.0, 0.0, 0.0))
    assert line.max_grade() == 2
    circle = M.cga_circle(S, M.cga_sphere(cga, 1.0, 0.0, 0.0, 1.0))
    assert circle.max_grade() == 2 and M.cga_point_pair(P, M.cga_point(cga, 0.0, 0.0, 0.0)).max_grade() == 2
    four = M.cga_point(cga, 1.0, 0.0, 0.0) ^ M.cga_point(cga, 0.0, 1.0, 0.0) ^ M.cga_point(cga, 0.0, 0.0, 1.0) ^ M.cga_point(cga, -1.0, 0.0, 0.0)
    assert four.max_grade() == 4
    reflected = M.cga_plane(cga, 0.0, 0.0, 1.0, 0.0).sandwich(P)  # plane * P * ~plane
    assert reflected.cga_extract_center() == pytest.approx((1.0, 2.0, -3.0), abs=1e-10)


@gpu
def test_sta_physics(api):  # tests/test_sta.py:237-436, src/sta.rs:675-934
    A, M = api.Algebra, api.Multivector
    sta = A.sta()
    g = [M.sta_gamma(sta, i) for i in range(4)]
    assert (g[0] * g[0]).scalar() == 1.0 and all((g[i] * g[i]).scalar() == -1.0 for i in (1, 2, 3))  # metric (+,-,-,-)
    assert (g[0] * g[1]) == -(g[1] * g[0])
    x = M.sta_vector(sta, 5.0, 3.0, 0.0, 0.0)
    R = M.sta_boost(sta, [0.5, 0.0, 0.0])
    boosted = R * x * R.reverse()
    assert boosted.sta_interval_squared() == pytest.approx(x.sta_interval_squared(), rel=1e-10)
    assert boosted.pure_grade() == 1
    gamma = 1.0 / math.sqrt(1 - 0.25)  # boost of the event (t, x): t' = gamma (t + v x) for this rotor's sign convention or (t - v x)
    t, r = boosted.sta_spacetime_split()
    assert t == pytest.approx(gamma * (5.0 - 0.5 * 3.0)) or t == pytest.approx(gamma * (5.0 + 0.5 * 3.0))
    R1 = M.sta_boost(sta, [0.3, 0.0, 0.0])
    assert (R1 * R1).sta_lorentz_gamma() > R1.sta_lorentz_gamma()  # tests/test_sta.py:268-279
    rot = M.sta_rotation(sta, [0.0, 0.0, 1.0], math.pi / 2)
    y = M.sta_vector(sta, 5.0, 3.0, 4.0, 0.0)
    assert (rot * y * rot.reverse()).sta_spacetime_split()[0] == pytest.approx(5.0)
    t, r = (rot * M.sta_vector(sta, 0.0, 1.0, 0.0, 0.0) * rot.reverse()).sta_spacetime_split()
    assert t == pytest.approx(0.0) and r == pytest.approx([0.0, 1.0, 0.0], abs=1e-10)
    rx = M.sta_rotation(sta, [1.0, 0.0, 0.0], math.pi / 2)
    assert (rx * M.sta_vector(sta, 0.0, 0.0, 1.0, 0.0) * rx.reverse()).sta_spacetime_split()[1] == pytest.approx([0.0, 0.0, 1.0], abs=1e-10)
    F = M.sta_bivector(sta, [1.0, 0.0, 0.0], [0.0, 0.0, 0.0])
    E_dual, B_dual = F.sta_em_dual().sta_field_decompose()
    assert abs(B_dual[0]) == pytest.approx(1.0, abs=1e-10) and E_dual == pytest.approx([0.0, 0.0, 0.0], abs=1e-12)
    assert (M.sta_I(sta) * M.sta_I(sta)).scalar() == -1.0  # I^2 = -1 in Cl(1,3)


@gpu
def test_algebra_metric_squares(api):  # tests/test_algebras.py:171-246
    A, M = api.Algebra, api.Multivector
    for alg, squares in ((A.euclidean(3), [1, 1, 1]), (A.pga(3), [1, 1, 1, 0]), (A.cga(3), [1, 1, 1, 1, -1]), (A.sta(), [1, -1, -1, -1]), (A(2, 1, 1), [1, 1, -1, 0])):
        for i, sq in enumerate(squares):
            e = M.basis_in(i + 1, alg)
            assert (e * e).scalar() == float(sq) and (e * e).algebra() == alg
        e1, e2 = M.basis_in(1, alg), M.basis_in(2, alg)
        assert (e1 * e2) == -(e2 * e1) and (e1 ^ e2) == (e1 * e2)
    with pytest.raises(ValueError, match="algebra mismatch"):
        M.basis_in(1, A.pga(3)).geometric_product(M.basis_in(1, A.sta()))
    v = M.vector_in([1.0, 2.0, 3.0, 0.0], A.pga(3))
    assert v.is_vector() and v.algebra() == A.pga(3) and M.pseudoscalar_in(A.pga(3)).max_grade() == 4
    legacy, explicit = M.from_vector([0.0, 1.0, 0.0]), M.vector_in([1.0, 0.0, 0.0], A.euclidean(3))
    assert (explicit * legacy).coefficients()[3] == 1.0  # explicit Euclidean mixes with dims-only (multivector.rs:48-55)
    assert M.scalar_in(2.5, A.cga(3)).scalar() == 2.5


@pytest.mark.gpu
def test_sta_batched_constructors(api):
    """Batched sta_vector / sta_bivector / sta_boost (src/sta.rs:105-109,202-212,253-296): the rows equal what
    the per-object constructors (pinned by the reference's tests/test_sta.py) give, and a batch of events
    boosted row by row keeps its interval t^2 - x^2 - y^2 - z^2 (tests/test_sta.py:237-290)."""
    A, M, B = api.Algebra, api.Multivector, api.MultivectorBatch
    sta = A.sta()
    rng = np.random.default_rng(3)
    n = 257
    for dt in (np.float64, np.float32):
        txyz = rng.standard_normal((n, 4)).astype(dt)
        ev = B.from_events_sta(sta, txyz).to_numpy()
        assert ev.dtype == dt and np.array_equal(ev[:, [1, 2, 4, 8]], txyz) and np.count_nonzero(ev) == np.count_nonzero(txyz)
        e, b = rng.standard_normal((n, 3)).astype(dt), rng.standard_normal((n, 3)).astype(dt)
        f = B.from_fields_sta(sta, e, b).to_numpy()
        assert np.array_equal(f, oracle.sta_bivectors(e, b))
    for r in (0, 100, 256):  # against the per-object constructor
        want = M.sta_bivector(sta, [float(v) for v in e[r]], [float(v) for v in b[r]]).coefficients()
        assert np.array_equal(f[r].astype(np.float64), np.asarray(want, np.float64).astype(np.float32))
    v = rng.uniform(-0.55, 0.55, (n, 3))
    v[0] = 0.0            # identity rotor
    v[1] = (0.9, 0.9, 0.0)  # |v| >= 1: no rotor
    boosts = B.from_boosts_sta(sta, v)
    got = boosts.to_numpy()
    want = oracle.sta_boosts(v)
    assert np.all(np.isnan(got[1, [0, 3, 5, 9]])) and got[0, 0] == 1.0 and not got[0, 1:].any()
    ok = np.ones(n, bool); ok[1] = False
    np.testing.assert_allclose(got[ok], want[ok], rtol=1e-14, atol=1e-16)  # libm vs CUDA math: a few ulp
    for r in (2, 77):
        np.testing.assert_allclose(got[r], M.sta_boost(sta, [float(t) for t in v[r]]).coefficients(), rtol=1e-14, atol=1e-16)
    with pytest.raises(ValueError, match="STA"):
        B.from_boosts_sta(A.pga(3), v)
