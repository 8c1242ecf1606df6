"""Drop-in behaviour of `largecrimsoncanine.MultivectorBatch` on the GPU: the same checks the
reference's own acceptance suite makes (/root/reference/tests/test_batch.py, cited per test),
written against the same public names so they read like the reference's tests.  Inputs are seeded
(the reference's are not) and every closed form is also compared with the oracle.
"""
import numpy as np
import pytest
from numpy.testing import assert_allclose

import oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def api():
    import largecrimsoncanine as L

    return L


@pytest.fixture()
def rng():
    return np.random.default_rng(1234)


def test_creation(api, rng):  # test_batch.py:22-88
    Algebra, Batch = api.Algebra, api.MultivectorBatch
    R3 = Algebra.euclidean(3)
    batch = Batch.from_numpy(R3, rng.standard_normal((100, 8)))
    assert batch.len() == 100 and batch.num_blades() == 8 and len(batch) == 100 and not batch.is_empty()
    with pytest.raises(ValueError, match="must have 8 columns"):
        Batch.from_numpy(R3, rng.standard_normal((100, 4)))
    coords = rng.standard_normal((100, 3))
    vecs = Batch.from_vectors(R3, coords)
    assert_allclose(vecs.to_vector_coords(), coords)
    assert np.array_equal(vecs.to_numpy(), oracle.OracleAlgebra(3).from_vectors(coords))
    with pytest.raises(ValueError, match="must have 3 columns"):
        Batch.from_vectors(R3, rng.standard_normal((100, 2)))
    values = np.array([1.0, 2.0, 3.0, 4.0, 5.0])
    assert_allclose(Batch.from_scalars(R3, values).scalar(), values)
    biv = rng.standard_normal((50, 3))
    B = Batch.from_bivectors(R3, biv)
    assert B.len() == 50 and np.array_equal(B.to_numpy(), oracle.OracleAlgebra(3).from_bivectors(biv))
    with pytest.raises(ValueError, match="must have 3 columns"):
        Batch.from_bivectors(R3, rng.standard_normal((5, 2)))
    assert Batch.from_vectors(R3, rng.standard_normal((20, 3))).to_numpy().shape == (20, 8)
    empty = Batch.from_numpy(R3, np.zeros((0, 8)))
    assert empty.len() == 0 and empty.is_empty() and empty.to_numpy().shape == (0, 8)
    assert empty.geometric_product(empty).len() == 0 and empty.reverse().len() == 0 and empty.norm().shape == (0,)
    with pytest.raises(TypeError):  # rust-numpy does not cast: non-float arrays are a TypeError (batch.rs:79)
        Batch.from_numpy(R3, np.zeros((4, 8), dtype=np.int64))
    # non-contiguous and float32 inputs (additive)
    wide = rng.standard_normal((10, 16))
    assert np.array_equal(Batch.from_numpy(R3, wide[:, ::2]).to_numpy(), wide[:, ::2])
    f32 = rng.standard_normal((10, 8)).astype(np.float32)
    out = Batch.from_numpy(R3, f32).to_numpy()
    assert out.dtype == np.float32 and np.array_equal(out, f32)


def test_element_access(api, rng):  # test_batch.py:95-158
    Algebra, Multivector, Batch = api.Algebra, api.Multivector, api.MultivectorBatch
    R3 = Algebra.euclidean(3)
    batch = Batch.from_vectors(R3, np.eye(3))
    assert_allclose(batch.get(0).to_vector_coords(), [1.0, 0.0, 0.0])
    assert_allclose(batch.get(1).to_vector_coords(), [0.0, 1.0, 0.0])
    assert batch.get(2).algebra() == R3  # batch.rs:297-302: elements carry the batch's algebra
    two = Batch.from_vectors(R3, np.array([[1.0, 2.0, 3.0], [4.0, 5.0, 6.0]]))
    assert_allclose(two[0].to_vector_coords(), [1.0, 2.0, 3.0])
    assert_allclose(two[-1].to_vector_coords(), [4.0, 5.0, 6.0])
    five = Batch.from_vectors(R3, rng.standard_normal((5, 3)))
    with pytest.raises(IndexError):
        five.get(10)
    with pytest.raises(IndexError):
        five[100]
    with pytest.raises(IndexError):
        five[-6]
    zeros = Batch.from_vectors(R3, np.zeros((3, 3)))
    zeros.set(1, Multivector.from_vector([1.0, 2.0, 3.0]))
    assert_allclose(zeros.get(1).to_vector_coords(), [1.0, 2.0, 3.0])
    assert_allclose(zeros.to_numpy()[[0, 2]], 0.0)
    with pytest.raises(ValueError, match="blades"):  # batch.rs:323-329
        zeros.set(0, Multivector.from_vector([1.0, 2.0]))
    with pytest.raises(IndexError):
        zeros.set(3, Multivector.from_vector([1.0, 2.0, 3.0]))


def test_sandwich(api, rng):  # test_batch.py:165-235, 623-638
    Algebra, Multivector, Batch = api.Algebra, api.Multivector, api.MultivectorBatch
    R3 = Algebra.euclidean(3)
    coords = rng.standard_normal((1000, 3))
    batch = Batch.from_vectors(R3, coords)
    identity = Multivector.from_scalar(1.0, 3)
    assert_allclose(batch.sandwich(identity).to_vector_coords(), coords, rtol=1e-10)
    rotor = Multivector.from_axis_angle(Multivector.from_vector([0.0, 0.0, 1.0]), np.pi / 4)
    assert_allclose(batch.sandwich(rotor).norm(), batch.norm(), rtol=1e-10)
    e1 = Batch.from_vectors(R3, np.array([[1.0, 0.0, 0.0]]))
    quarter = Multivector.from_axis_angle(Multivector.from_vector([0.0, 0.0, 1.0]), np.pi / 2)
    assert_allclose(e1.sandwich(quarter).to_vector_coords(), [[0.0, 1.0, 0.0]], atol=1e-10)
    assert_allclose(e1.apply(quarter).to_vector_coords(), [[0.0, 1.0, 0.0]], atol=1e-10)
    # batch vs a loop of scalar Multivector sandwiches (each of those is a one-row device call)
    rotor = Multivector.from_axis_angle(Multivector.from_vector([1.0, 1.0, 1.0]), 0.7)
    got = batch.slice(0, 50).sandwich(rotor).to_vector_coords()
    loop = np.array([rotor.sandwich(Multivector.from_vector(c.tolist())).to_vector_coords() for c in coords[:50]])
    assert_allclose(got, loop, rtol=1e-10)
    big = Batch.from_vectors(R3, rng.standard_normal((10000, 3)))
    r6 = Multivector.from_axis_angle(Multivector.from_vector([0.0, 0.0, 1.0]), np.pi / 6)
    assert_allclose(big.sandwich(r6).norm(), big.norm(), rtol=1e-9)


def test_products(api, rng):  # test_batch.py:242-365
    Algebra, Multivector, Batch = api.Algebra, api.Multivector, api.MultivectorBatch
    R3 = Algebra.euclidean(3)
    a, b = rng.standard_normal((10, 3)), rng.standard_normal((10, 3))
    A, B = Batch.from_vectors(R3, a), Batch.from_vectors(R3, b)
    result = A.geometric_product(B)
    assert result.len() == 10
    for i in range(10):
        expected = Multivector.from_vector(a[i].tolist()).geometric_product(Multivector.from_vector(b[i].tolist()))
        assert np.max(np.abs(np.array(result.get(i).coefficients()) - np.array(expected.coefficients()))) < 1e-10
    assert np.array_equal(A.gp(B).to_numpy(), result.to_numpy())
    coords = rng.standard_normal((20, 3))
    scaled = Batch.from_vectors(R3, coords).geometric_product(Batch.from_scalars(R3, np.array([2.0])))
    assert scaled.len() == 20
    assert_allclose(scaled.to_vector_coords(), coords * 2.0, rtol=1e-10)
    with pytest.raises(ValueError, match="batch sizes must match"):
        A.geometric_product(Batch.from_vectors(R3, rng.standard_normal((5, 3))))
    with pytest.raises(ValueError, match="algebra mismatch"):  # batch.rs:354-359
        A.geometric_product(Batch.from_vectors(Algebra.euclidean(2), rng.standard_normal((10, 2))))
    with pytest.raises(TypeError):  # a Multivector is not accepted despite the docstring
        A.geometric_product(Multivector.from_vector([1.0, 0.0, 0.0]))
    e1, e2 = Batch.from_vectors(R3, np.array([[1.0, 0.0, 0.0]])), Batch.from_vectors(R3, np.array([[0.0, 1.0, 0.0]]))
    assert abs(e1.outer_product(e2).grade(2).get(0).coefficients()[3] - 1.0) < 1e-10
    assert_allclose(A.wedge(B).to_numpy(), -B.wedge(A).to_numpy(), rtol=1e-10)
    assert_allclose(A.left_contraction(B).scalar(), np.sum(a * b, axis=1), rtol=1e-10)
    assert np.array_equal(A.inner(B).to_numpy(), A.lc(B).to_numpy())


def test_norms_involutions_arithmetic(api, rng):  # test_batch.py:372-533
    Algebra, Batch = api.Algebra, api.MultivectorBatch
    R3 = Algebra.euclidean(3)
    coords = rng.standard_normal((100, 3))
    batch = Batch.from_vectors(R3, coords)
    assert_allclose(batch.norm(), np.linalg.norm(coords, axis=1), rtol=1e-10)
    assert_allclose(batch.norm_squared(), np.sum(coords**2, axis=1), rtol=1e-10)
    assert_allclose(batch.normalized().norm(), np.ones(100), rtol=1e-10)
    zero_rows = Batch.from_numpy(R3, np.zeros((3, 8)))
    assert np.array_equal(zero_rows.normalized().to_numpy(), np.zeros((3, 8)))  # zero norm: left unchanged
    assert_allclose(batch.reverse().to_vector_coords(), coords, rtol=1e-10)
    biv = Batch.from_bivectors(R3, rng.standard_normal((20, 3)))
    assert_allclose(biv.reverse().to_numpy()[:, [3, 5, 6]], -biv.to_numpy()[:, [3, 5, 6]], rtol=1e-10)
    assert_allclose(batch.grade_involution().to_vector_coords(), -coords, rtol=1e-10)
    other = rng.standard_normal((100, 3))
    B = Batch.from_vectors(R3, other)
    assert_allclose((batch + B).to_vector_coords(), coords + other, rtol=1e-10)
    assert_allclose((batch - B).to_vector_coords(), coords - other, rtol=1e-10)
    assert_allclose((batch * 2.5).to_vector_coords(), coords * 2.5, rtol=1e-10)
    assert_allclose((2.5 * batch).to_vector_coords(), coords * 2.5, rtol=1e-10)
    assert_allclose(batch.scale(2.5).to_vector_coords(), coords * 2.5, rtol=1e-10)
    assert_allclose((-batch).to_vector_coords(), -coords, rtol=1e-10)
    assert_allclose((batch / 2.0).to_vector_coords(), coords / 2.0, rtol=1e-10)
    with pytest.raises(ZeroDivisionError):
        batch / 0.0
    with pytest.raises(ValueError, match="batch sizes must match"):
        batch + Batch.from_vectors(R3, rng.standard_normal((7, 3)))


def test_slice_concat_grade_repr(api, rng):  # test_batch.py:540-612
    Algebra, Batch = api.Algebra, api.MultivectorBatch
    R3 = Algebra.euclidean(3)
    coords = rng.standard_normal((100, 3))
    batch = Batch.from_vectors(R3, coords)
    assert batch.slice(10, 20).len() == 10
    assert_allclose(batch.slice(10, 20).to_vector_coords(), coords[10:20], rtol=1e-10)
    assert batch.slice(5, 5).len() == 0 and batch.slice(100, 100).len() == 0
    for bad in ((0, 101), (50, 40), (101, 101)):
        with pytest.raises(IndexError, match="invalid slice"):
            batch.slice(*bad)
    a, b = rng.standard_normal((30, 3)), rng.standard_normal((20, 3))
    both = Batch.from_vectors(R3, a).concat(Batch.from_vectors(R3, b))
    assert both.len() == 50
    assert_allclose(both.to_vector_coords(), np.vstack([a, b]), rtol=1e-10)
    mixed = rng.standard_normal((10, 8))
    vec = Batch.from_numpy(R3, mixed).grade(1).to_numpy()
    assert np.all(vec[:, [0, 3, 5, 6, 7]] == 0) and np.array_equal(vec[:, [1, 2, 4]], mixed[:, [1, 2, 4]])
    s = repr(Batch.from_vectors(R3, rng.standard_normal((42, 3))))
    assert "42" in s and "Cl(3,0,0)" in s
