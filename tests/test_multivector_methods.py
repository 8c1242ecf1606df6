"""Per-object `Multivector` surface beyond the hot path: constructors, rotor conversions,
exp / log / slerp, reflections and projections, coefficient statistics and the Python protocol methods, written
against the reference's public names (src/multivector.rs, cited per block).  Every product inside them runs on the
CUDA kernels with n = 1, so these are GPU tests.  The reference's own tests/test_basic.py (795 tests) is the
acceptance suite for this surface; this file keeps a representative known-answer subset in the repository."""
import copy
import math
import pickle

import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def M():
    import largecrimsoncanine as L

    return L.Multivector


def test_basis_constructors_and_indexing(M):  # :433-754, :1228-1258, :2233-2273
    assert M.e1(3).coefficients() == [0, 1, 0, 0, 0, 0, 0, 0] and M.e3(3)[4] == 1.0 and M.e12(2)[3] == 1.0
    assert M.e31(3)[5] == -1.0 and M.e23(3)[6] == 1.0 and M.e123()[7] == 1.0 and M.pseudoscalar(4)[15] == 1.0
    with pytest.raises(ValueError, match="e3 requires dimension >= 3"):
        M.e3(2)
    B = M.from_bivector([1.0, 2.0, 3.0, 4.0, 5.0, 6.0], dims=4)  # lexicographic pairs: e12 e13 e14 e23 e24 e34
    assert [B[i] for i in (3, 5, 9, 6, 10, 12)] == [1.0, 2.0, 3.0, 4.0, 5.0, 6.0]
    v = M.from_vector([1.0, 2.0, 3.0])
    assert v.coefficient(2) == 2.0 and v.set_coefficient(0, 5.0).scalar() == 5.0 and v.n_coeffs == 8 and v.dimension == 3
    with pytest.raises(ValueError, match="out of bounds"):
        v.coefficient(8)
    with pytest.raises(ValueError, match="not a power of 2"):
        M.from_list([1.0, 2.0, 3.0])
    assert M.from_dict(v.to_dict()) == v and list(v) == v.coefficients() and bool(v) and not bool(M.zero(2))


def test_rotors_from_vectors_and_representations(M):  # :887-943, :1013-1213
    a, b = M.from_vector([1.0, 0.0, 0.0]), M.from_vector([0.0, 1.0, 0.0])
    R = M.rotor_from_vectors(a, b)
    assert R.is_rotor() and R.sandwich(a).approx_eq(b) and abs(R.rotation_angle() - math.pi / 2) < 1e-12
    half_turn = M.rotor_from_vectors(a, M.from_vector([-1.0, 0.0, 0.0]))  # anti-parallel branch
    assert half_turn.sandwich(a).approx_eq(a.scale(-1.0))
    q = M.from_quaternion(0.5, 0.5, 0.5, 0.5)
    assert q.to_quaternion() == (0.5, 0.5, 0.5, 0.5) and M.from_rotation_matrix(q.to_rotation_matrix()).same_rotation(q)
    e = M.from_euler_angles(0.3, -0.2, 0.1)
    assert e.to_euler_angles() == pytest.approx((0.3, -0.2, 0.1), abs=1e-12) and e.is_spinor()
    axis, angle = M.from_axis_angle(M.from_vector([0.0, 0.0, 2.0]), 0.8).axis_angle()
    # the reference takes the dual of the rotor's bivector part, which points along -axis for from_axis_angle's
    # sign convention (:981-997 vs :3694-3698); the drop-in reproduces that, sign included
    assert angle == pytest.approx(0.8) and axis.approx_eq(M.from_vector([0.0, 0.0, -1.0]))
    n1, n2 = M.from_vector([1.0, 0.0, 0.0]), M.from_vector([1.0, 1.0, 0.0])
    assert M.rotor_from_reflections(n1, n2).rotation_angle() == pytest.approx(math.pi / 2)
    assert M.random_rotor(3).is_rotor(1e-9) and M.random_vector(4).is_unit(1e-12) and M.random(2).n_coeffs == 4


def test_exp_log_and_interpolation(M):  # :3492-3568, :4321-4486
    B = M.from_bivector([0.4, 0.0, 0.0], dims=3)
    R = B.exp()
    assert R.scalar() == pytest.approx(math.cos(0.4)) and R[3] == pytest.approx(math.sin(0.4)) and R.log().approx_eq(B)
    assert R.sqrt().squared().approx_eq(R) and R.powf(2.0).approx_eq(R * R) and (R ** 3).approx_eq(R * R * R)
    assert (R ** -1).approx_eq(R.inverse()) and (R ** 0).approx_eq(M.from_scalar(1.0, 3))
    assert M.from_scalar(1.5, 2).exp().scalar() == pytest.approx(math.exp(1.5))
    general = M.from_list([0.1, 0.2, 0.0, 0.0]).exp()  # scalar + vector: power series, e^0.1 (cosh 0.2 + sinh 0.2 e1)
    assert general.scalar() == pytest.approx(math.exp(0.1) * math.cosh(0.2)) and general[1] == pytest.approx(math.exp(0.1) * math.sinh(0.2))
    one = M.from_scalar(1.0, 3)
    assert one.slerp(R, 0.0).approx_eq(one) and one.slerp(R, 1.0).approx_eq(R) and one.slerp(R, 0.5).approx_eq(R.sqrt())
    assert one.lerp(R, 0.25).scalar() == pytest.approx(0.75 + 0.25 * math.cos(0.4)) and one.nlerp(R, 0.5).is_unit()
    with pytest.raises(ValueError, match="unit rotor"):
        (R * 2.0).log()


def test_reflection_projection_and_angles(M):  # :3717-4204, :4510-4831
    v, n = M.from_vector([1.0, 2.0, 3.0]), M.from_vector([0.0, 0.0, 1.0])
    assert v.reflect(n).approx_eq(M.from_vector([1.0, 2.0, -3.0])) and v.double_reflection(n, n).approx_eq(v)
    e12 = M.e12(3)
    assert v.project(e12).approx_eq(M.from_vector([1.0, 2.0, 0.0])) and v.reject(e12).approx_eq(M.from_vector([0.0, 0.0, 3.0]))
    assert v.reflect_in_plane(e12).approx_eq(M.from_vector([1.0, 2.0, -3.0]))  # B x B with unit B: the in-plane part stays, the normal part flips
    a, b = M.from_vector([1.0, 0.0, 0.0]), M.from_vector([1.0, 1.0, 0.0])
    assert a.angle_between(b) == pytest.approx(math.pi / 4) and a.cos_angle(b) == pytest.approx(math.sqrt(0.5)) and a.sin_angle(b) == pytest.approx(math.sqrt(0.5))
    assert a.signed_angle(b, n) == pytest.approx(math.pi / 4) and b.signed_angle(a, n) == pytest.approx(-math.pi / 4)
    assert a.cross(b).approx_eq(M.from_vector([0.0, 0.0, 1.0])) and a.triple_product(b, n) == pytest.approx(1.0)
    assert a.is_parallel(a * 3.0) and a.is_antiparallel(a * -2.0) and a.is_orthogonal(n) and not a.is_same_direction(a * -1.0)
    assert a.rotate_in_plane(math.pi / 2, e12).approx_eq(M.from_vector([0.0, 1.0, 0.0]))
    assert a.area(b) == pytest.approx(1.0) and a.volume(b, n) == pytest.approx(1.0) and v.distance(a) == pytest.approx(math.sqrt(13.0))
    assert v.angle_to_plane(e12) == pytest.approx(math.asin(3.0 / math.sqrt(14.0))) and v.distance_to_plane(e12, M.zero(3)) == pytest.approx(3.0)
    assert a.lies_in_plane(e12) and n.is_perpendicular_to_plane(e12) and v.midpoint(a).approx_eq(M.from_vector([1.0, 1.0, 1.5]))
    with pytest.raises(ValueError, match="only defined in 3D"):
        M.from_vector([1.0, 0.0]).cross(M.from_vector([0.0, 1.0]))


def test_grade_surgery_predicates_and_statistics(M):  # :1293-1428, :1906-1966, :2285-2487, :2856-2880, :3087-3475
    m = M.from_list([1.0, 2.0, -3.0, 4.0])
    assert m.negate_grade(1).coefficients() == [1.0, -2.0, 3.0, 4.0] and m.clear_grade(2)[3] == 0.0 and m.scale_grade(0, 3.0).scalar() == 3.0
    assert m.add_scalar(1.0).scalar() == 2.0 and m.with_scalar(9.0).scalar() == 9.0 and m.filter_grades([0, 2]).coefficients() == [1.0, 0.0, 0.0, 4.0]
    assert m.even().coefficients() == [1.0, 0.0, 0.0, 4.0] and m.odd().coefficients() == [0.0, 2.0, -3.0, 0.0] and m.hat().coefficients() == [1.0, -2.0, 3.0, 4.0]
    assert m.conjugate().coefficients() == [1.0, -2.0, 3.0, -4.0] and m.dagger() == m.reverse() and m.split_even_odd()[1] == m.odd()
    assert m.sign().coefficients() == [1.0, 1.0, -1.0, 1.0] and m.positive_part()[2] == 0.0 and m.negative_part()[2] == -3.0 and m.abs_coefficients()[2] == 3.0
    assert m.clamp_coefficients(-1.0, 2.0).coefficients() == [1.0, 2.0, -1.0, 2.0] and M.from_list([1e-12, 1.0]).clean().coefficients() == [0.0, 1.0]
    assert M.from_list([1.2345, 2.5]).round_coefficients(2)[0] == 1.23 and round(M.from_list([1.6, -0.4]))[0] == 2.0 and math.floor(M.from_list([1.6, -0.4]))[1] == -1.0
    assert m.dominant_blade() == 3 and m.dominant_grade() == 1 and m.dominant_coefficient() == 4.0 and m.sorted_blades()[0] == (3, 4.0, 2)
    assert m.sum_coefficients() == 4.0 and m.max_coefficient() == 4.0 and m.min_coefficient() == -3.0 and m.mean_coefficient() == 1.0 and m.median_coefficient() == 1.5
    assert m.l1_norm() == 10.0 and m.linf_norm() == 4.0 and m.nonzero_count() == 4 and m.sparsity() == 0.0 and m.variance_coefficient() == 6.5
    assert m.grade_count() == 3 and m.min_grade() == 0 and m.has_grade(2) and not m.is_homogeneous() and m.homogeneous_grade() is None and m.max_grade_part()[3] == 4.0
    v = M.from_vector([3.0, 4.0])
    assert v.is_blade() and v.is_versor() and not v.is_null() and v.is_odd() and v.blade_square() == 25.0 and v.unit().is_unit() and v.is_reflection() is False
    assert M.e12(3).is_bivector() and M.e123().is_pseudoscalar() and M.e123().is_trivector() and M.zero(3).is_zero() and M.from_list([1e-12, 0.0]).almost_zero()
    assert v.normalize_dominant()[2] == 1.0 and v.components() == [(1, 3.0), (2, 4.0)] and v.blade_indices() == [1, 2] and v.grade_norm(1) == 5.0
    assert v.allclose(v * (1.0 + 1e-7)) and v.max_abs_diff(v * 2.0) == 4.0 and v.relative_error(v * 2.0) == pytest.approx(1.0) and v.coefficient_distance(v * 2.0) == 5.0
    e1, e2 = M.e1(2), M.e2(2)
    assert e1.anticommutes_with(e2) and e1.commutes_with(e1) and e1.symmetric_product(e2).is_zero() and e1.fat_dot(M.e12(2)).approx_eq(e2)
    assert e1.complement().approx_eq(e2) and e1.join(e2).approx_eq(M.e12(2)) and v.div(v).approx_eq(M.from_scalar(1.0, 2))


def test_python_protocols(M):  # :1824-1882, :2490-2603, :5832-5898
    v = M.from_vector([1.234, -2.5])
    assert str(v) == "1.234*e1 + -2.5*e2" and str(M.zero(2)) == "0" and str(M.from_list([2.0, 0.0])) == "2" and format(v) == str(v)
    assert format(v, ".2f") == "1.23*e1 + -2.50*e2" and f"{v:+.1f}" == "+1.2*e1 + -2.5*e2"
    assert format(M.from_vector([1234.0, 0.00567]), ".2e") == "1.23e3*e1 + 5.67e-3*e2"  # Rust-style compact exponent
    with pytest.raises(ValueError):
        format(v, "invalid")
    assert abs(M.from_vector([3.0, 4.0])) == 5.0 and float(M.from_scalar(2.5, 2)) == 2.5 and int(M.from_scalar(2.5, 2)) == 2
    with pytest.raises(ValueError, match="non-scalar"):
        float(v)
    assert copy.copy(v) == v and copy.deepcopy(v) == v and pickle.loads(pickle.dumps(v)) == v and v.copy() == v
