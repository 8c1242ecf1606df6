"""GPU parity tests proper: the CUDA path (through the `largecrimsoncanine` extension and through the
raw C ABI) against the CPU oracle on the same seeded inputs.

Bar (BASELINE.json north_star): bit-exact for tables / masks / indexing (tests/test_cabi_host.py);
products within 1e-12 (f64) / 1e-5 (f32) of the reference arithmetic.  On top of that the STRICT
mode (separate multiply and add in the reference's term order) must be BIT-EXACT.

Tolerance definition for the FMA mode (hard part 3): element-wise relative error
is ill-posed where cancellation drives an output towards zero, so each element is compared with
    |gpu - ref| <= tol * max(|ref|, sum_terms |a_i||b_j|)
and each row with  max|gpu - ref| <= tol * max|ref|.
"""
import ctypes as C
import math

import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu

SIGS = [(3, 0, 0), (3, 0, 1), (4, 1, 0), (1, 3, 0), (2, 0, 0), (2, 1, 1), (1, 0, 0), (4, 0, 0), (0, 0, 2), (3, 1, 1)]
TOL = {np.float64: 1e-12, np.float32: 1e-5}
SIZES = [1, 3, 64, 257, 1000]


@pytest.fixture(scope="module")
def L():
    import largecrimsoncanine as ext

    assert ext.device_count() >= 1
    return ext


@pytest.fixture()
def strict(L):
    L.set_strict_fp(True)
    yield
    L.set_strict_fp(False)


def rand(rng, n, nb, dt):
    return rng.standard_normal((n, nb)).astype(dt)


def assert_close(got, ref, bound, dt):
    tol = TOL[dt]
    got, ref = np.asarray(got, np.float64), np.asarray(ref, np.float64)
    scale = np.maximum(np.abs(ref), np.asarray(bound, np.float64))
    bad = np.abs(got - ref) > tol * scale + np.finfo(np.float64).tiny
    assert not bad.any(), f"{bad.sum()} elements off; worst {np.abs(got - ref).max():.3e}"
    if got.ndim == 2 and got.size:
        row_err = np.abs(got - ref).max(axis=1)
        row_ref = np.abs(ref).max(axis=1)
        assert np.all(row_err <= tol * np.maximum(row_ref, np.asarray(bound, np.float64).max(axis=1)))


def algebras(L, sig):
    return L.Algebra(*sig), oracle.OracleAlgebra(*sig)


# ----------------------------------------------------------------------------------------------- products
@pytest.mark.parametrize("sig", SIGS)
@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_products_strict_mode_is_bit_exact(L, strict, sig, dt):
    alg, o = algebras(L, sig)
    rng = np.random.default_rng(hash(sig) % 1000)
    for n in SIZES:
        a, b = rand(rng, n, o.num_blades, dt), rand(rng, n, o.num_blades, dt)
        A, B = L.MultivectorBatch.from_numpy(alg, a), L.MultivectorBatch.from_numpy(alg, b)
        assert np.array_equal(A.to_numpy(), a)
        assert np.array_equal(A.geometric_product(B).to_numpy(), o.geometric_product(a, b))
        assert np.array_equal(A.outer_product(B).to_numpy(), o.outer_product(a, b))
        assert np.array_equal(A.left_contraction(B).to_numpy(), o.left_contraction(a, b))
        g, w = A.gp_and_outer(B)
        assert np.array_equal(g.to_numpy(), o.geometric_product(a, b)) and np.array_equal(w.to_numpy(), o.outer_product(a, b))
        assert np.array_equal(A.norm_squared(), o.norm_squared(a))
        assert np.array_equal(A.norm(), o.norm(a))
        assert np.array_equal(A.normalized().to_numpy(), o.normalized(a))
    # broadcast in both positions (batch.rs:361-376)
    a, b = rand(rng, 130, o.num_blades, dt), rand(rng, 1, o.num_blades, dt)
    A, B = L.MultivectorBatch.from_numpy(alg, a), L.MultivectorBatch.from_numpy(alg, b)
    assert np.array_equal(A.geometric_product(B).to_numpy(), o.geometric_product(a, b))
    assert np.array_equal(B.geometric_product(A).to_numpy(), o.geometric_product(b, a))
    assert np.array_equal(B.left_contraction(A).to_numpy(), o.left_contraction(b, a))


@pytest.mark.parametrize("sig", SIGS)
@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_products_fma_mode_within_tolerance(L, sig, dt):
    alg, o = algebras(L, sig)
    rng = np.random.default_rng(7)
    n = 2051
    a, b = rand(rng, n, o.num_blades, dt), rand(rng, n, o.num_blades, dt)
    A, B = L.MultivectorBatch.from_numpy(alg, a), L.MultivectorBatch.from_numpy(alg, b)
    for name, op, ref in (("gp", A.geometric_product, o.geometric_product), ("outer", A.outer_product, o.outer_product),
                          ("lc", A.left_contraction, o.left_contraction)):
        assert_close(op(B).to_numpy(), ref(a, b), o.product_abs_bound(name, a, b), dt)
    g, w = A.gp_and_outer(B)
    assert_close(g.to_numpy(), o.geometric_product(a, b), o.product_abs_bound("gp", a, b), dt)
    assert_close(w.to_numpy(), o.outer_product(a, b), o.product_abs_bound("outer", a, b), dt)


@pytest.mark.parametrize("sig", SIGS)
@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_sandwich(L, sig, dt):
    """batch.rs:434-503: one rotor for the whole batch, two rounded stages.  Dense rotors take the
    general kernel, even ones (rotors / motors) the kernel with odd blades pruned at compile time."""
    alg, o = algebras(L, sig)
    rng = np.random.default_rng(11)
    nb = o.num_blades
    dense = rng.standard_normal(nb)
    even = dense * np.array([1.0 if bin(i).count("1") % 2 == 0 else 0.0 for i in range(nb)])
    for rotor in (dense, even):
        R = rotor_mv(L, alg, rotor)
        for n in (1, 5, 258, 1027):
            x = rand(rng, n, nb, dt)
            X = L.MultivectorBatch.from_numpy(alg, x)
            ref = o.sandwich(rotor, x)
            L.set_strict_fp(True)
            try:
                assert np.array_equal(X.sandwich(R).to_numpy(), ref)
            finally:
                L.set_strict_fp(False)
            assert_close(X.sandwich(R).to_numpy(), ref, sandwich_bound(o, rotor, x), dt)
    with pytest.raises(ValueError, match="algebra mismatch"):  # batch.rs:437-442
        X.sandwich(L.Multivector.from_scalar(1.0, o.dimension + 1))


def rotor_mv(L, alg, coeffs):
    """An algebra-carrying Multivector with arbitrary coefficients (additive constructor)."""
    return L.Multivector.from_list_in([float(c) for c in coeffs], alg)


def sandwich_bound(o, rotor, x):
    ar = np.abs(np.asarray(rotor, np.float64))[None, :].astype(x.dtype)
    rx = o.product_abs_bound("gp", ar, np.abs(x))
    return o.product_abs_bound("gp", rx, ar)


# ----------------------------------------------------------------------------------------------- unary family
@pytest.mark.parametrize("sig", SIGS)
@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_unary_family_exact(L, sig, dt):
    alg, o = algebras(L, sig)
    rng = np.random.default_rng(3)
    for n in (1, 2, 255, 1030):
        x = rand(rng, n, o.num_blades, dt)
        X = L.MultivectorBatch.from_numpy(alg, x)
        assert np.array_equal(X.reverse().to_numpy(), o.reverse(x))
        assert np.array_equal(X.grade_involution().to_numpy(), o.grade_involution(x))
        assert np.array_equal(X.clifford_conjugate().to_numpy(), o.clifford_conjugate(x))
        for k in range(o.dimension + 1):
            assert np.array_equal(X.grade(k).to_numpy(), o.grade(x, k))
        assert np.array_equal(X.scale(2.5).to_numpy(), o.scale(x, 2.5))
        assert np.array_equal((X / 3.0).to_numpy(), o.div(x, 3.0))
        assert np.array_equal((-X).to_numpy(), o.scale(x, -1.0))
        y = rand(rng, n, o.num_blades, dt)
        Y = L.MultivectorBatch.from_numpy(alg, y)
        assert np.array_equal((X + Y).to_numpy(), o.add(x, y)) and np.array_equal((X - Y).to_numpy(), o.sub(x, y))
        one = L.MultivectorBatch.from_numpy(alg, y[:1])
        assert np.array_equal((X + one).to_numpy(), o.add(x, y[:1])) and np.array_equal((one - X).to_numpy(), o.sub(y[:1], x))
    with pytest.raises(ValueError, match="exceeds dimension"):
        X.grade(o.dimension + 1)


def test_duals(L):
    rng = np.random.default_rng(5)
    for sig in [(3, 0, 0), (4, 1, 0), (1, 3, 0), (4, 0, 0)]:
        alg, o = algebras(L, sig)
        x = rng.standard_normal((77, o.num_blades))
        X = L.MultivectorBatch.from_numpy(alg, x)
        assert np.array_equal(X.dual().to_numpy(), o.dual(x))
        assert np.array_equal(X.undual().to_numpy(), o.undual(x))
        assert np.array_equal(X.right_dual().to_numpy(), o.right_dual(x))
        assert np.array_equal(X.dual().undual().to_numpy(), x)
    alg, o = algebras(L, (3, 0, 1))
    x = rng.standard_normal((33, 16))
    X = L.MultivectorBatch.from_numpy(alg, x)
    assert np.array_equal(X.pga_dual().to_numpy(), o.pga_dual(x))
    with pytest.raises(ValueError):  # I*I = 0 in PGA: Multivector::dual raises (multivector.rs:2991-2996)
        X.dual()


# ----------------------------------------------------------------------------------------------- golden fixtures
@pytest.mark.parametrize("sig", [(2, 0, 0), (3, 0, 0), (3, 0, 1), (4, 1, 0), (1, 3, 0), (2, 1, 1)])
@pytest.mark.parametrize("dt", ["f64", "f32"])
def test_cuda_path_reproduces_reference_golden_vectors(L, strict, golden, sig, dt):
    """tests/golden/ref_torch_golden.npz comes from the reference's own lcc/torch implementation."""
    alg = L.Algebra(*sig)
    k = "cl%d%d%d.%s" % (*sig, dt)
    a, b, rotor = golden[k + ".a"], golden[k + ".b"], golden[k + ".rotor"]
    A, B = L.MultivectorBatch.from_numpy(alg, a), L.MultivectorBatch.from_numpy(alg, b)
    assert np.array_equal(A.geometric_product(B).to_numpy(), golden[k + ".gp"])
    assert np.array_equal(A.geometric_product(L.MultivectorBatch.from_numpy(alg, b[:1])).to_numpy(), golden[k + ".gp_bcast"])
    assert np.array_equal(A.outer_product(B).to_numpy(), golden[k + ".outer"])
    assert np.array_equal(A.left_contraction(B).to_numpy(), golden[k + ".inner"])
    R = rotor_mv(L, alg, rotor[0].astype(a.dtype).astype(np.float64))
    assert np.array_equal(A.sandwich(R).to_numpy(), golden[k + ".sandwich"])
    assert np.array_equal(A.reverse().to_numpy(), golden[k + ".reverse"])
    assert np.array_equal(A.grade_involution().to_numpy(), golden[k + ".involution"])
    assert np.array_equal(A.clifford_conjugate().to_numpy(), golden[k + ".conjugate"])
    assert np.array_equal(A.norm_squared(), golden[k + ".norm_squared"])
    for g in range(alg.dimension + 1):
        assert np.array_equal(A.grade(g).to_numpy(), golden[f"{k}.grade{g}"])


# ----------------------------------------------------------------------------------------------- big algebras
@pytest.mark.parametrize("sig", [(6, 0, 0), (4, 2, 1), (8, 0, 0)])
def test_algebras_above_32_blades(L, strict, sig):
    alg, o = algebras(L, sig)
    rng = np.random.default_rng(9)
    n = 37
    a, b = rng.standard_normal((n, o.num_blades)), rng.standard_normal((n, o.num_blades))
    A, B = L.MultivectorBatch.from_numpy(alg, a), L.MultivectorBatch.from_numpy(alg, b)
    assert np.array_equal(A.geometric_product(B).to_numpy(), o.geometric_product(a, b))
    assert np.array_equal(A.outer_product(B).to_numpy(), o.outer_product(a, b))
    assert np.array_equal(A.left_contraction(B).to_numpy(), o.left_contraction(a, b))
    # fixed-operand left product = the broadcast branch (BASELINE config 5 shape, small N)
    M = L.MultivectorBatch.from_numpy(alg, a[:1])
    assert np.array_equal(M.geometric_product(B).to_numpy(), o.geometric_product(a[:1], b))
    rotor = rng.standard_normal(o.num_blades)
    R = rotor_mv(L, alg, rotor)
    assert np.array_equal(A.sandwich(R).to_numpy(), o.sandwich(rotor, a))
    assert np.array_equal(A.reverse().to_numpy(), o.reverse(a))
    assert np.array_equal(A.norm_squared(), o.norm_squared(a))


# ----------------------------------------------------------------------------------------------- sums
@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_batch_sum_and_mean(L, dt):
    alg, o = algebras(L, (1, 3, 0))
    rng = np.random.default_rng(13)
    for n in (1, 1000, 100_003):
        x = rand(rng, n, 16, dt)
        X = L.MultivectorBatch.from_numpy(alg, x)
        want, mag = o.sum(x)
        tol = 1e-13 if dt == np.float64 else 1e-6
        assert np.all(np.abs(X.sum().astype(np.float64) - want) <= tol * mag)
        assert np.all(np.abs(X.mean().astype(np.float64) - want / n) <= tol * mag / n)
        assert np.array_equal(X.sum(), X.sum())  # fixed reduction tree: run-to-run identical


# ----------------------------------------------------------------------------------------------- PGA / CGA
def test_pga_kats_through_the_batched_path(L):
    """/root/reference/tests/test_pga.py:245-255 and :287-300 replayed with N = 1 and N = large."""
    pga = L.Algebra.pga(3)
    T = L.Multivector.pga_translator(pga, 4.0, 5.0, 6.0)
    pts = np.array([[1.0, 2.0, 3.0]] + np.random.default_rng(1).standard_normal((4096, 3)).tolist())
    P = L.MultivectorBatch.from_points_pga(pga, pts)
    moved = P.sandwich(T).pga_point_coords()
    np.testing.assert_allclose(moved[0], (5.0, 7.0, 9.0), atol=1e-12)
    np.testing.assert_allclose(moved, pts + np.array([4.0, 5.0, 6.0]), atol=1e-12)
    M = L.Multivector.pga_motor_from_axis_angle(pga, [0.0, 0.0, 1.0], [0.0, 0.0, 0.0], math.pi / 2)
    rot = L.MultivectorBatch.from_points_pga(pga, np.array([[1.0, 0.0, 0.0]])).sandwich(M).pga_point_coords()
    np.testing.assert_allclose(rot, [[0.0, 1.0, 0.0]], atol=1e-12)
    # embedding equals the scalar constructor (src/pga.rs:131-134), bit for bit
    emb = P.to_numpy()
    one = L.Multivector.pga_point(pga, *pts[5])
    assert np.array_equal(emb[5], np.array(one.coefficients()))
    # scalar Multivector.sandwich (device, one row) agrees with the batch
    assert np.allclose(T.sandwich(one).pga_point_coords(), moved[5], atol=1e-12)


def test_cga_points(L):
    cga = L.Algebra.cga(3)
    pts = np.random.default_rng(2).standard_normal((513, 3))
    P = L.MultivectorBatch.from_points_cga(cga, pts)
    emb = P.to_numpy()
    nsq = pts[:, 0] * pts[:, 0] + pts[:, 1] * pts[:, 1] + pts[:, 2] * pts[:, 2]  # src/cga.rs:196
    want = np.zeros((513, 32))
    want[:, 1], want[:, 2], want[:, 4] = pts[:, 0], pts[:, 1], pts[:, 2]
    want[:, 8], want[:, 16] = 0.5 * nsq - 0.5, 0.5 * nsq + 0.5
    assert np.array_equal(emb, want)
    np.testing.assert_allclose(P.cga_extract_center(), pts, rtol=1e-14)
    # null vectors: P.P = 0 and distance formula sqrt(-2 P.Q) (src/cga.rs:701-720), 3-4-5 KAT (tests/test_cga.py:284-291)
    a = L.MultivectorBatch.from_points_cga(cga, np.array([[0.0, 0.0, 0.0]]))
    b = L.MultivectorBatch.from_points_cga(cga, np.array([[3.0, 4.0, 0.0]]))
    assert abs(math.sqrt(-2.0 * a.left_contraction(b).scalar()[0]) - 5.0) < 1e-12
    assert np.all(np.abs(P.left_contraction(P).scalar()) < 1e-12)


# ----------------------------------------------------------------------------------------------- raw C ABI
def test_c_abi_with_raw_views_aos_and_soa(L):
    """Same products through include/lcc_b200.h with caller-owned device memory in both layouts."""
    from largecrimsoncanine_b200 import cabi

    lib = cabi.lib()
    o = oracle.OracleAlgebra(3, 0, 1)
    alg = cabi.Algebra(3, 0, 1)
    rng = np.random.default_rng(17)
    n, nb = 777, 16
    for dt, code in ((np.float64, cabi.LCC_F64), (np.float32, cabi.LCC_F32)):
        a, b = rand(rng, n, nb, dt), rand(rng, n, nb, dt)
        es = a.itemsize
        for layout in (cabi.LCC_AOS, cabi.LCC_SOA):
            ld = nb if layout == cabi.LCC_AOS else ((n + 63) // 64) * 64
            elems = n * nb if layout == cabi.LCC_AOS else nb * ld
            ptrs = []
            for _ in range(3):
                p = C.c_void_p()
                cabi.check(lib.lcc_device_alloc(C.byref(p), elems * es, None))
                ptrs.append(p)
            def host_image(x):
                if layout == cabi.LCC_AOS:
                    return np.ascontiguousarray(x)
                img = np.zeros((nb, ld), dt)
                img[:, :n] = x.T
                return img
            ia, ib = host_image(a), host_image(b)
            cabi.check(lib.lcc_memcpy_h2d(ptrs[0], ia.ctypes.data, ia.nbytes, None))
            cabi.check(lib.lcc_memcpy_h2d(ptrs[1], ib.ctypes.data, ib.nbytes, None))
            va, vb, vo = (cabi.view(p.value, n, ld, code, layout) for p in ptrs)
            for op, ref in ((cabi.LCC_OP_GP, o.geometric_product), (cabi.LCC_OP_OUTER, o.outer_product), (cabi.LCC_OP_LC, o.left_contraction)):
                cabi.check(lib.lcc_batch_product(alg.h, op, C.byref(va), C.byref(vb), C.byref(vo), cabi.LCC_FLAG_STRICT_FP, None))
                out = np.empty_like(ia)
                cabi.check(lib.lcc_memcpy_d2h(out.ctypes.data, ptrs[2], out.nbytes, None))
                cabi.check(lib.lcc_stream_sync(None))
                got = out if layout == cabi.LCC_AOS else out[:, :n].T
                assert np.array_equal(got, ref(a, b))
            rotor = rng.standard_normal(nb)
            cabi.check(lib.lcc_batch_sandwich(alg.h, rotor.ctypes.data_as(C.POINTER(C.c_double)), C.byref(va), C.byref(vo), cabi.LCC_FLAG_STRICT_FP, None))
            out = np.empty_like(ia)
            cabi.check(lib.lcc_memcpy_d2h(out.ctypes.data, ptrs[2], out.nbytes, None))
            cabi.check(lib.lcc_stream_sync(None))
            got = out if layout == cabi.LCC_AOS else out[:, :n].T
            assert np.array_equal(got, o.sandwich(rotor, a))
            for p in ptrs:
                cabi.check(lib.lcc_device_free(p, None))
    # argument validation: wrong column count, mismatched rows (batch.rs:83-88,372-375)
    bad = cabi.view(1 << 20, 4, 8, cabi.LCC_F64, cabi.LCC_AOS)
    assert lib.lcc_batch_product(alg.h, 0, C.byref(bad), C.byref(bad), C.byref(bad), 0, None) == cabi.LCC_ERR_VALUE
    assert b"must have 16 columns" in lib.lcc_last_error()


def test_host_buffer_entry_points(L):
    """lcc_host_sandwich / lcc_host_product: chunked H2D -> kernel -> D2H pipeline over several chunks."""
    from largecrimsoncanine_b200 import cabi

    lib = cabi.lib()
    alg, o = cabi.Algebra(3, 0, 1), oracle.OracleAlgebra(3, 0, 1)
    rng = np.random.default_rng(19)
    n = 2_300_000  # > 2 chunks of 64 MiB at 64 B/row
    x = rng.standard_normal((n, 16), dtype=np.float32)
    motor = np.zeros(16)
    motor[[0, 3, 5, 6, 9, 10, 12]] = rng.standard_normal(7)
    out = np.empty_like(x)
    cabi.check(lib.lcc_host_sandwich(alg.h, cabi.LCC_F32, motor.ctypes.data_as(C.POINTER(C.c_double)), x.ctypes.data, out.ctypes.data, n, cabi.LCC_FLAG_STRICT_FP))
    sample = np.r_[0:4096, n // 2:n // 2 + 4096, n - 4096:n]
    assert np.array_equal(out[sample], o.sandwich(motor, x[sample]))
    a = rng.standard_normal((300_001, 16))
    b = rng.standard_normal((1, 16))
    res = np.empty_like(a)
    cabi.check(lib.lcc_host_product(alg.h, cabi.LCC_F64, cabi.LCC_OP_GP, a.ctypes.data, a.shape[0], b.ctypes.data, 1, res.ctypes.data, cabi.LCC_FLAG_STRICT_FP))
    assert np.array_equal(res[:5000], o.geometric_product(a[:5000], b))


def test_sharded_batch_single_rank(L):
    """The native sharding API (lcc_shard_*) with a group of one GPU -- no NCCL involved; with two or more GPUs in the
    box the same calls run as a single-process multi-GPU group (ncclCommInitAll) and must agree with the oracle too.
    (Host logic on CPU: tests/test_sharding_gloo.py; both group modes at N >= 2: tests/tools/multi_gpu_check.py.)"""
    from largecrimsoncanine_b200 import cabi, sharded

    rng = np.random.default_rng(23)
    x, w = rng.standard_normal((5001, 16)), rng.standard_normal((5001, 16))
    o, alg = oracle.OracleAlgebra(1, 3, 0), cabi.Algebra(1, 3, 0)
    ref = o.left_contraction(o.grade(x, 1), w)
    want, mag = o.sum(ref)
    for devices in ([0], None):
        group = sharded.ShardGroup(devices)
        assert group.world == group.local_size == (1 if devices else L.device_count())
        X, W = sharded.ShardedBatch.from_numpy(group, alg, x), sharded.ShardedBatch.from_numpy(group, alg, w)
        prod = X.grade(1).left_contraction(W, cabi.LCC_FLAG_STRICT_FP)
        assert np.array_equal(prod.to_numpy(), ref)  # every shard, gathered back in row order
        assert np.all(np.abs(prod.sum() - want) <= 1e-12 * mag)
        assert np.all(np.abs(prod.mean() - want / 5001) <= 1e-12 * mag / 5001)
        fused = X.product_sum(W, "inner", grade=1)
        assert np.all(np.abs(fused - want) <= 1e-12 * mag)
        assert np.all(np.abs(X.product_sum(W, "inner", grade=1, mean=True) - want / 5001) <= 1e-12 * mag / 5001)
        m = rng.standard_normal(16)
        assert np.array_equal(X.sandwich(m, cabi.LCC_FLAG_STRICT_FP).to_numpy(), o.sandwich(m, x))
        B = sharded.ShardedBatch.broadcast(group, alg, w[0])
        assert np.array_equal(B.geometric_product(X, cabi.LCC_FLAG_STRICT_FP).to_numpy(), o.geometric_product(w[:1], x))
        assert len(prod) == 5001 and sum(e - b for b, e in prod.local_rows()) == 5001
        group.sync()
        for t in (X, W, prod, B):
            t.free()
        group.close()


def test_fixed_operand_product_with_a_pitch_the_tensor_map_cannot_cover(L):
    """A raw SoA f32 view whose pitch is not a multiple of 32 rows (a torch (nb, n) tensor with n = 1000): the K6 tensor
    map walks rows in strips of 32, so the dispatcher must take the exact kernel instead of returning zeros for rows
    992..999 (round-1 advisor finding)."""
    from largecrimsoncanine_b200 import cabi

    lib, alg, o = cabi.lib(), cabi.Algebra(6, 0, 0), oracle.OracleAlgebra(6, 0, 0)
    nb, n, ld = 64, 1000, 1000
    rng = np.random.default_rng(77)
    m, x = rng.standard_normal((1, nb)).astype(np.float32), rng.standard_normal((n, nb)).astype(np.float32)
    px, pm, po = C.c_void_p(), C.c_void_p(), C.c_void_p()
    for p, bytes_ in ((px, nb * ld * 4), (pm, nb * 4 * 4), (po, nb * ld * 4)):
        cabi.check(lib.lcc_device_alloc(C.byref(p), bytes_, None))
    xi = np.ascontiguousarray(x.T)
    mi = np.zeros((nb, 4), np.float32)
    mi[:, 0] = m[0]
    cabi.check(lib.lcc_memcpy_h2d(px, xi.ctypes.data, xi.nbytes, None))
    cabi.check(lib.lcc_memcpy_h2d(pm, mi.ctypes.data, mi.nbytes, None))
    vx, vm, vo = cabi.view(px.value, n, ld, cabi.LCC_F32, cabi.LCC_SOA), cabi.view(pm.value, 1, 4, cabi.LCC_F32, cabi.LCC_SOA), cabi.view(po.value, n, ld, cabi.LCC_F32, cabi.LCC_SOA)
    cabi.check(lib.lcc_batch_product(alg.h, cabi.LCC_OP_GP, C.byref(vm), C.byref(vx), C.byref(vo), 0, None))
    out = np.empty_like(xi)
    cabi.check(lib.lcc_memcpy_d2h(out.ctypes.data, po, out.nbytes, None))
    cabi.check(lib.lcc_stream_sync(None))
    ref = o.geometric_product(m.astype(np.float64), x.astype(np.float64))
    assert np.all(np.abs(out.T - ref).max(axis=1) <= 1e-5 * np.abs(ref).max(axis=1))  # every row, including the last 8
    for p in (px, pm, po):
        cabi.check(lib.lcc_device_free(p, None))


@pytest.mark.parametrize("sig", [(6, 0, 0), (4, 2, 1), (8, 0, 0)])
def test_fixed_operand_product_f64_on_the_fp64_tensor_pipe(L, sig):
    """K6d: the same broadcast product in the reference's own precision -> DMMA GEMM (mma.sync m8n8k4.f64).
    Tensor-core summation order, so the bar is the f64 tolerance of the path (1e-12 against the magnitude of the
    terms); ragged sizes cover partial 128-row tiles and the zero-filled rows past the padded pitch."""
    alg, o = algebras(L, sig)
    nb = o.num_blades
    rng = np.random.default_rng(53)
    for n in (2, 127, 128, 129, 1000, 20001):
        m = rng.standard_normal((1, nb))
        x = rng.standard_normal((n, nb))
        M, X = L.MultivectorBatch.from_numpy(alg, m), L.MultivectorBatch.from_numpy(alg, x)
        for got, a, b in ((M.geometric_product(X).to_numpy(), m, x), (X.geometric_product(M).to_numpy(), x, m)):
            ref = o.geometric_product(a, b)
            bound = o.product_abs_bound("gp", np.abs(a), np.abs(b))
            assert np.all(np.abs(got - ref) <= 1e-12 * np.maximum(np.abs(ref), bound))


@pytest.mark.parametrize("sig", [(3, 0, 0), (3, 0, 1), (1, 3, 0), (2, 1, 1), (4, 1, 0)])
@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_per_row_sandwich_through_the_c_abi(sig, dt):
    """lcc_batch_sandwich_rows: one rotor per row, or ONE device-resident rotor row for all (the no-read-back path of
    lcc.torch), both layouts.  Up to 16 blades it is one fused kernel (K2r: direct loads, TMA-staged AoS tiles from 4096 /
    131072 rows on); strict mode is bit-exact against the reference composition gp(gp(R, x), ~R)
    (/root/reference/lcc/torch/multivector.py:988-1019); 32 blades take the composed products -- same bits."""
    from largecrimsoncanine_b200 import cabi

    lib, alg, o = cabi.lib(), cabi.Algebra(*sig), oracle.OracleAlgebra(*sig)
    nb = o.num_blades
    code = cabi.LCC_F32 if dt == np.float32 else cabi.LCC_F64
    rev = np.array([1.0 if bin(i).count("1") % 4 < 2 else -1.0 for i in range(nb)], dt)
    rng = np.random.default_rng(4 + nb)
    big = 131072 + 37 if nb * np.dtype(dt).itemsize <= 64 else 4096 + 37  # past the AoS TMA threshold of this row size
    for n in (1, 63, 1000, big):
        x, r = rand(rng, n, nb, dt), rand(rng, n, nb, dt)
        # even versors (rotors / motors: odd-grade coefficients zero) take the pruned term lists -- per row when the whole
        # warp agrees (row-per-thread kernels vote), for the one device rotor always; a batch that mixes both kinds must
        # come out the same as well
        odd = [k for k in range(nb) if bin(k).count("1") % 2]
        r_even, r_mixed = r.copy(), r.copy()
        r_even[:, odd] = 0
        r_mixed[::3, odd] = 0
        r_mixed[n // 2:, odd] = 0
        for rr in (r, r[:1], r_even, r_even[:1], r_mixed):
            want = o.geometric_product(o.geometric_product(rr, x), rr * rev)
            for layout in (cabi.LCC_AOS, cabi.LCC_SOA):
                def put(a):
                    rows = a.shape[0]
                    ld = nb if layout == cabi.LCC_AOS else ((rows + 63) // 64) * 64
                    img = np.ascontiguousarray(a) if layout == cabi.LCC_AOS else np.zeros((nb, ld), dt)
                    if layout == cabi.LCC_SOA:
                        img[:, :rows] = a.T
                    p = C.c_void_p()
                    cabi.check(lib.lcc_device_alloc(C.byref(p), max(img.nbytes, 16), None))
                    cabi.check(lib.lcc_memcpy_h2d(p, img.ctypes.data, img.nbytes, None))
                    return p, cabi.view(p.value, rows, ld, code, layout), img
                px, vx, _ = put(x)
                pr, vr, _ = put(rr)
                po, vo, img = put(np.zeros_like(x))
                cabi.check(lib.lcc_batch_sandwich_rows(alg.h, C.byref(vr), C.byref(vx), C.byref(vo), cabi.LCC_FLAG_STRICT_FP, None))
                cabi.check(lib.lcc_memcpy_d2h(img.ctypes.data, po, img.nbytes, None))
                cabi.check(lib.lcc_stream_sync(None))
                got = img if layout == cabi.LCC_AOS else img[:, :n].T
                assert np.array_equal(got, want), (n, rr.shape[0], layout)
                cabi.check(lib.lcc_batch_sandwich_rows(alg.h, C.byref(vr), C.byref(vx), C.byref(vo), 0, None))  # FMA mode: tolerance
                cabi.check(lib.lcc_memcpy_d2h(img.ctypes.data, po, img.nbytes, None))
                cabi.check(lib.lcc_stream_sync(None))
                got = img if layout == cabi.LCC_AOS else img[:, :n].T
                assert np.all(np.abs(got - want).max(axis=1) <= 10 * TOL[dt] * np.abs(want).max(axis=1))
                for p in (px, pr, po):
                    cabi.check(lib.lcc_device_free(p, None))


@pytest.mark.parametrize("sig", [(6, 0, 0), (4, 2, 1), (8, 0, 0)])
@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_fixed_rotor_sandwich_and_masked_products_on_tensor_cores(L, sig, dt):
    """64..256 blades, FMA mode: x -> R x ~R with one rotor is ONE linear map -> one tensor-core GEMM against the
    sandwich matrix (reference: two rounded products, /root/reference/src/batch.rs:452-497); fixed-operand outer
    product / left contraction -> the same GEMM with the op's term mask folded into the matrix (batch.rs:514-574,585-655).
    Bar: the path's tolerance against the reference arithmetic in f64; strict mode stays bit-exact (table kernel)."""
    alg, o = algebras(L, sig)
    nb = o.num_blades
    tol = TOL[dt]
    rng = np.random.default_rng(97 + nb)
    for n in (1, 255, 257, 40001):
        x = rng.standard_normal((n, nb)).astype(dt)
        r = rng.standard_normal(nb) / np.sqrt(nb)
        X = L.MultivectorBatch.from_numpy(alg, x)
        got = X.sandwich(rotor_mv(L, alg, r)).to_numpy()
        ref = o.sandwich(r, x.astype(np.float64))
        assert np.all(np.abs(got - ref).max(axis=1) <= tol * np.abs(ref).max(axis=1)), f"sandwich n={n}"
        m = rng.standard_normal((1, nb)).astype(dt)
        M = L.MultivectorBatch.from_numpy(alg, m)
        for name, fn in (("outer", o.outer_product), ("lc", o.left_contraction)):
            for got, a, b in ((getattr(M, name if name == "lc" else "outer_product")(X).to_numpy(), m, x),
                              (getattr(X, name if name == "lc" else "outer_product")(M).to_numpy(), x, m)):
                ref = fn(a.astype(np.float64), b.astype(np.float64))
                bound = o.product_abs_bound("outer" if name == "outer" else "lc", np.abs(a.astype(np.float64)), np.abs(b.astype(np.float64)))
                assert np.all(np.abs(got - ref) <= tol * np.maximum(np.abs(ref), bound) + 1e-300), f"{name} n={n}"
        gp, outer = M.gp_and_outer(X)
        ref_gp = o.geometric_product(m.astype(np.float64), x.astype(np.float64))
        assert np.all(np.abs(gp.to_numpy() - ref_gp).max(axis=1) <= tol * np.abs(ref_gp).max(axis=1))
    x = rng.standard_normal((300, nb)).astype(dt)
    r = rng.standard_normal(nb)
    L.set_strict_fp(True)
    try:
        assert np.array_equal(L.MultivectorBatch.from_numpy(alg, x).sandwich(rotor_mv(L, alg, r)).to_numpy(), o.sandwich(r, x))
    finally:
        L.set_strict_fp(False)


@pytest.mark.parametrize("sig", [(8, 0, 0), (4, 4, 0), (5, 3, 0), (0, 8, 0), (7, 0, 0), (4, 3, 0), (1, 6, 0)])
@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_block_diagonalised_fixed_operand_product_and_sandwich(L, sig, dt):
    """K8: M * x, x * M and R x ~R in the algebra's 16 x 16 matrix representation (two Walsh-Hadamard passes + one 16^3
    product per row) against the reference arithmetic (/root/reference/src/batch.rs:365-404,452-497) in f64, at ragged
    sizes (the last 1-3 rows of a blade-major batch go through the exact kernels), and against the dense GEMM path."""
    alg, o = algebras(L, sig)
    nb, tol = o.num_blades, TOL[dt]
    rng = np.random.default_rng(1000 + nb + sig[1])
    assert L.get_option("matrix_rep") == 2  # default: K8 for both precisions
    for n in (64, 65, 4099, 40003):
        x = rng.standard_normal((n, nb)).astype(dt)
        m = rng.standard_normal((1, nb)).astype(dt)
        X, M = L.MultivectorBatch.from_numpy(alg, x), L.MultivectorBatch.from_numpy(alg, m)
        x64, m64 = x.astype(np.float64), m.astype(np.float64)
        before = L.kernel_launch_count()
        left, right = M.geometric_product(X).to_numpy(), X.geometric_product(M).to_numpy()
        for got, ref, bound in ((left, o.geometric_product(m64, x64), o.product_abs_bound("gp", np.abs(m64), np.abs(x64))),
                                (right, o.geometric_product(x64, m64), o.product_abs_bound("gp", np.abs(x64), np.abs(m64)))):
            assert np.all(np.abs(got - ref) <= tol * np.maximum(np.abs(ref), bound))
            assert np.all(np.abs(got - ref).max(axis=1) <= tol * np.abs(ref).max(axis=1))
        r = rng.standard_normal(nb) / np.sqrt(nb)
        got = X.sandwich(rotor_mv(L, alg, r)).to_numpy()
        ref = o.sandwich(r, x64)
        assert np.all(np.abs(got - ref).max(axis=1) <= tol * np.abs(ref).max(axis=1))
        L.set_option("matrix_rep", 0)  # the dense GEMM of the same map must agree within the same bar
        try:
            dense = M.geometric_product(X).to_numpy()
        finally:
            L.set_option("matrix_rep", 2)
        assert np.all(np.abs(dense - left).max(axis=1) <= 2 * tol * np.abs(left).max(axis=1))
    # a basis-blade operand is a signed permutation: exact in both paths
    e = np.zeros((1, nb), dt)
    e[0, nb - 1] = 1.0
    got = L.MultivectorBatch.from_numpy(alg, e).geometric_product(X).to_numpy()
    assert np.allclose(got, o.geometric_product(e.astype(np.float64), x64), rtol=0, atol=tol * np.abs(x64).max())


@pytest.mark.parametrize("sig", [(6, 0, 0), (8, 0, 0)])
def test_reference_layout_views_reach_the_tensor_cores(L, sig):
    """Raw C ABI on LCC_AOS views (what lcc.torch and lcc_host_product pass) in 64 / 256-blade algebras: fixed-operand
    product and fixed-rotor sandwich go through the tensor-core GEMMs (transposed in and out), not the table kernel --
    same tolerance, padded row pitch left untouched."""
    from largecrimsoncanine_b200 import cabi

    lib, alg, o = cabi.lib(), cabi.Algebra(*sig), oracle.OracleAlgebra(*sig)
    nb, n, ld = o.num_blades, 5003, o.num_blades + 8
    rng = np.random.default_rng(5)
    for dt, code, tol in ((np.float32, cabi.LCC_F32, 1e-5), (np.float64, cabi.LCC_F64, 1e-12)):
        x = rng.standard_normal((n, nb)).astype(dt)
        m = rng.standard_normal((1, nb)).astype(dt)
        img = np.full((n, ld), -3.5, dt)
        img[:, :nb] = x
        es = img.itemsize
        px, pm, po = C.c_void_p(), C.c_void_p(), C.c_void_p()
        cabi.check(lib.lcc_device_alloc(C.byref(px), img.nbytes, None))
        cabi.check(lib.lcc_device_alloc(C.byref(po), img.nbytes, None))
        cabi.check(lib.lcc_device_alloc(C.byref(pm), nb * es, None))
        cabi.check(lib.lcc_memcpy_h2d(px, img.ctypes.data, img.nbytes, None))
        cabi.check(lib.lcc_memcpy_h2d(po, img.ctypes.data, img.nbytes, None))
        cabi.check(lib.lcc_memcpy_h2d(pm, m.ctypes.data, m.nbytes, None))
        vx, vo, vm = cabi.view(px.value, n, ld, code, cabi.LCC_AOS), cabi.view(po.value, n, ld, code, cabi.LCC_AOS), cabi.view(pm.value, 1, nb, code, cabi.LCC_AOS)
        before = lib.lcc_kernel_launch_count()
        cabi.check(lib.lcc_batch_product(alg.h, cabi.LCC_OP_GP, C.byref(vm), C.byref(vx), C.byref(vo), 0, None))
        out = np.empty_like(img)
        cabi.check(lib.lcc_memcpy_d2h(out.ctypes.data, po, out.nbytes, None))
        cabi.check(lib.lcc_stream_sync(None))
        ref = o.geometric_product(m.astype(np.float64), x.astype(np.float64))
        assert np.all(np.abs(out[:, :nb] - ref).max(axis=1) <= tol * np.abs(ref).max(axis=1)) and np.all(out[:, nb:] == -3.5)
        r = rng.standard_normal(nb) / np.sqrt(nb)
        cabi.check(lib.lcc_batch_sandwich(alg.h, r.ctypes.data_as(C.POINTER(C.c_double)), C.byref(vx), C.byref(vo), 0, None))
        cabi.check(lib.lcc_memcpy_d2h(out.ctypes.data, po, out.nbytes, None))
        cabi.check(lib.lcc_stream_sync(None))
        ref = o.sandwich(r, x.astype(np.float64))
        assert np.all(np.abs(out[:, :nb] - ref).max(axis=1) <= tol * np.abs(ref).max(axis=1)) and np.all(out[:, nb:] == -3.5)
        for p in (px, pm, po):
            cabi.check(lib.lcc_device_free(p, None))


@pytest.mark.parametrize("sig", [(8, 0, 0), (4, 4, 0), (5, 3, 0), (0, 8, 0), (1, 7, 0), (7, 0, 0), (4, 3, 0), (0, 7, 0), (3, 4, 0)])
@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_block_diagonalised_product_on_reference_layout_units(L, sig, dt):
    """K8 on LCC_AOS views (the reference's (N, num_blades) row-major layout, /root/reference/src/batch.rs:41-47): units of
    whole rows, no transposes -- M * x, x * M (batch.rs:365-404) and R x ~R (:452-497) at ragged sizes, with a tight and a
    padded row pitch (padding must stay untouched), in place (out == x), against the reference arithmetic in f64."""
    from largecrimsoncanine_b200 import cabi

    lib, alg, o = cabi.lib(), cabi.Algebra(*sig), oracle.OracleAlgebra(*sig)
    nb, tol = o.num_blades, TOL[dt]
    code = cabi.LCC_F32 if dt == np.float32 else cabi.LCC_F64
    tables = (C.c_uint32 * (13 * 16))()
    assert lib.lcc_algebra_matrix_rep_layout_aos(alg.h, np.dtype(dt).itemsize, tables) == 1
    rng = np.random.default_rng(2000 + nb + sig[1])
    L.set_option("matrix_rep", 2)
    try:
        for n, pad in ((64, 0), (65, 8), (4099, 0), (12003, 4)):
            ld = nb + pad
            x = rng.standard_normal((n, nb)).astype(dt)
            m = rng.standard_normal((1, nb)).astype(dt)
            x64, m64 = x.astype(np.float64), m.astype(np.float64)
            img = np.full((n, ld), -3.5, dt)
            img[:, :nb] = x
            px, po = C.c_void_p(), C.c_void_p()
            for ptr in (px, po):
                cabi.check(lib.lcc_device_alloc(C.byref(ptr), img.nbytes, None))
                cabi.check(lib.lcc_memcpy_h2d(ptr, img.ctypes.data, img.nbytes, None))
            vx, vo = cabi.view(px.value, n, ld, code, cabi.LCC_AOS), cabi.view(po.value, n, ld, code, cabi.LCC_AOS)
            out = np.empty_like(img)

            def fetch(ptr=po):
                cabi.check(lib.lcc_memcpy_d2h(out.ctypes.data, ptr, out.nbytes, None))
                cabi.check(lib.lcc_stream_sync(None))
                assert np.all(out[:, nb:] == -3.5)
                return out[:, :nb]

            fixed = m64[0].copy()  # the fixed operand is known on the host (MultivectorBatch keeps the row it was built from)
            fptr = fixed.ctypes.data_as(C.POINTER(C.c_double))
            before = lib.lcc_kernel_launch_count()
            cabi.check(lib.lcc_batch_fixed_product(alg.h, cabi.LCC_OP_GP, 0, fptr, C.byref(vx), C.byref(vo), 0, None))
            assert lib.lcc_kernel_launch_count() - before == 1  # one K8 launch: no transposes, no tail kernel
            ref = o.geometric_product(m64, x64)
            assert np.all(np.abs(fetch() - ref).max(axis=1) <= tol * np.abs(ref).max(axis=1)), f"left n={n}"
            cabi.check(lib.lcc_batch_fixed_product(alg.h, cabi.LCC_OP_GP, 1, fptr, C.byref(vx), C.byref(vo), 0, None))
            ref = o.geometric_product(x64, m64)
            assert np.all(np.abs(fetch() - ref).max(axis=1) <= tol * np.abs(ref).max(axis=1)), f"right n={n}"
            # the same two products with the fixed operand resident on the DEVICE (what lcc.torch passes): K8 builds its
            # 16 x 16 factor from the device row with a one-block kernel -- two launches, no host copy of the operand
            pm = C.c_void_p()
            cabi.check(lib.lcc_device_alloc(C.byref(pm), m.nbytes, None))
            cabi.check(lib.lcc_memcpy_h2d(pm, m.ctypes.data, m.nbytes, None))
            vm = cabi.view(pm.value, 1, nb, code, cabi.LCC_AOS)
            before = lib.lcc_kernel_launch_count()
            cabi.check(lib.lcc_batch_product(alg.h, cabi.LCC_OP_GP, C.byref(vm), C.byref(vx), C.byref(vo), 0, None))
            assert lib.lcc_kernel_launch_count() - before == 2
            ref = o.geometric_product(m64, x64)
            assert np.all(np.abs(fetch() - ref).max(axis=1) <= tol * np.abs(ref).max(axis=1)), f"device operand, left n={n}"
            cabi.check(lib.lcc_batch_product(alg.h, cabi.LCC_OP_GP, C.byref(vx), C.byref(vm), C.byref(vo), 0, None))
            ref = o.geometric_product(x64, m64)
            assert np.all(np.abs(fetch() - ref).max(axis=1) <= tol * np.abs(ref).max(axis=1)), f"device operand, right n={n}"
            if n == 4099:  # blade-major views of the same data: 128-byte pieces per blade, the last 4099 % (16 / itemsize) rows on the exact kernels
                ps, pso, pms = C.c_void_p(), C.c_void_p(), C.c_void_p()
                lds = n + (-n) % 4
                soa = np.zeros((nb, lds), dt)
                soa[:, :n] = x.T
                msoa = np.zeros((nb, 4), dt)
                msoa[:, 0] = m[0]
                for ptr, arr in ((ps, soa), (pso, soa), (pms, msoa)):
                    cabi.check(lib.lcc_device_alloc(C.byref(ptr), arr.nbytes, None))
                    cabi.check(lib.lcc_memcpy_h2d(ptr, arr.ctypes.data, arr.nbytes, None))
                vs, vso, vms = cabi.view(ps.value, n, lds, code, cabi.LCC_SOA), cabi.view(pso.value, n, lds, code, cabi.LCC_SOA), cabi.view(pms.value, 1, 4, code, cabi.LCC_SOA)
                cabi.check(lib.lcc_batch_product(alg.h, cabi.LCC_OP_GP, C.byref(vms), C.byref(vs), C.byref(vso), 0, None))
                got = np.empty_like(soa)
                cabi.check(lib.lcc_memcpy_d2h(got.ctypes.data, pso, got.nbytes, None))
                cabi.check(lib.lcc_stream_sync(None))
                ref = o.geometric_product(m64, x64)
                assert np.all(np.abs(got[:, :n].T - ref).max(axis=1) <= tol * np.abs(ref).max(axis=1)), "device operand, blade-major"
                # and one device rotor for the blade-major batch (two-pass K8 + the exact composition on the last rows)
                rs = rng.standard_normal(nb).astype(dt) / np.sqrt(nb).astype(dt)
                rsoa = np.zeros((nb, 4), dt)
                rsoa[:, 0] = rs
                prs = C.c_void_p()
                cabi.check(lib.lcc_device_alloc(C.byref(prs), rsoa.nbytes, None))
                cabi.check(lib.lcc_memcpy_h2d(prs, rsoa.ctypes.data, rsoa.nbytes, None))
                vrs = cabi.view(prs.value, 1, 4, code, cabi.LCC_SOA)
                cabi.check(lib.lcc_batch_sandwich_rows(alg.h, C.byref(vrs), C.byref(vs), C.byref(vso), 0, None))
                cabi.check(lib.lcc_memcpy_d2h(got.ctypes.data, pso, got.nbytes, None))
                cabi.check(lib.lcc_stream_sync(None))
                ref = o.sandwich(rs.astype(np.float64), x64)
                assert np.all(np.abs(got[:, :n].T - ref).max(axis=1) <= tol * np.abs(ref).max(axis=1)), "device rotor, blade-major"
                cabi.check(lib.lcc_device_free(prs, None))
                for ptr in (ps, pso, pms):
                    cabi.check(lib.lcc_device_free(ptr, None))
            cabi.check(lib.lcc_device_free(pm, None))
            r = rng.standard_normal(nb) / np.sqrt(nb)
            # one DEVICE rotor for the batch (lcc_batch_sandwich_rows, one-row rotors view): K8's two-pass form, both factors
            # built on the device -- three launches instead of reverse + two products
            rdev = r.astype(dt)[None, :].copy()
            pr = C.c_void_p()
            cabi.check(lib.lcc_device_alloc(C.byref(pr), rdev.nbytes, None))
            cabi.check(lib.lcc_memcpy_h2d(pr, rdev.ctypes.data, rdev.nbytes, None))
            vr = cabi.view(pr.value, 1, nb, code, cabi.LCC_AOS)
            before = lib.lcc_kernel_launch_count()
            cabi.check(lib.lcc_batch_sandwich_rows(alg.h, C.byref(vr), C.byref(vx), C.byref(vo), 0, None))
            assert lib.lcc_kernel_launch_count() - before == 3
            ref = o.sandwich(rdev[0].astype(np.float64), x64)
            assert np.all(np.abs(fetch() - ref).max(axis=1) <= tol * np.abs(ref).max(axis=1)), f"device rotor n={n}"
            cabi.check(lib.lcc_device_free(pr, None))
            cabi.check(lib.lcc_batch_sandwich(alg.h, r.ctypes.data_as(C.POINTER(C.c_double)), C.byref(vx), C.byref(vx), 0, None))  # in place
            ref = o.sandwich(r, x64)
            assert np.all(np.abs(fetch(px) - ref).max(axis=1) <= tol * np.abs(ref).max(axis=1)), f"sandwich n={n}"
            for ptr in (px, po):
                cabi.check(lib.lcc_device_free(ptr, None))
    finally:
        L.set_option("matrix_rep", 2)


@pytest.mark.parametrize("sig", [(1, 3, 0), (3, 0, 1), (4, 1, 0), (2, 1, 1)])
@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_fused_grade_product_sum(L, sig, dt):
    """lcc_batch_product_sum: sum_rows(<a>_g (op) b) in one pass equals the three separate operations
    (batch.rs:1027-1052 + 353-655 + column sums) within summation tolerance; BASELINE config 4 pipeline."""
    alg, o = algebras(L, sig)
    rng = np.random.default_rng(41)
    for n in (1, 257, 50_001):
        a, b = rand(rng, n, o.num_blades, dt), rand(rng, n, o.num_blades, dt)
        A, B = L.MultivectorBatch.from_numpy(alg, a), L.MultivectorBatch.from_numpy(alg, b)
        tol = 1e-12 if dt == np.float64 else 2e-5
        for op, ref_fn in (("lc", o.left_contraction), ("gp", o.geometric_product), ("outer", o.outer_product)):
            for g in (None, 1, 2):
                ag = a if g is None else o.grade(a, g)
                want, _ = o.sum(ref_fn(ag, b))
                _, mag = o.sum(o.product_abs_bound(op, np.abs(ag), np.abs(b)))
                got = A.product_sum(B, op=op, grade=g).astype(np.float64)
                assert np.all(np.abs(got - want) <= tol * np.maximum(mag, 1e-300)), (op, g, n)
        one = L.MultivectorBatch.from_numpy(alg, b[:1])
        want, _ = o.sum(o.left_contraction(o.grade(a, 1), b[:1]))
        _, mag = o.sum(o.product_abs_bound("lc", np.abs(o.grade(a, 1)), np.abs(b[:1])))
        assert np.all(np.abs(A.product_sum(one, op="lc", grade=1).astype(np.float64) - want) <= tol * np.maximum(mag, 1e-300))
    with pytest.raises(ValueError, match="exceeds dimension"):
        A.product_sum(B, op="lc", grade=o.dimension + 1)


# ----------------------------------------------------------------------------------------------- AoS tiles through TMA
class _Dev:
    """Caller-owned device arrays for raw C-ABI calls."""

    def __init__(self, lib):
        self.lib, self.ptrs = lib, []

    def put(self, host):
        from largecrimsoncanine_b200 import cabi

        p = C.c_void_p()
        cabi.check(self.lib.lcc_device_alloc(C.byref(p), max(host.nbytes, 16), None))
        cabi.check(self.lib.lcc_memcpy_h2d(p, host.ctypes.data, host.nbytes, None))
        self.ptrs.append(p)
        return p

    def get(self, p, like):
        from largecrimsoncanine_b200 import cabi

        out = np.empty_like(like)
        cabi.check(self.lib.lcc_memcpy_d2h(out.ctypes.data, p, out.nbytes, None))
        cabi.check(self.lib.lcc_stream_sync(None))
        return out

    def free(self):
        from largecrimsoncanine_b200 import cabi

        for p in self.ptrs:
            cabi.check(self.lib.lcc_device_free(p, None))
        self.ptrs = []


@pytest.mark.parametrize("sig", [(3, 0, 0), (3, 0, 1), (4, 1, 0), (1, 3, 0), (2, 1, 1), (2, 0, 0), (5, 0, 0)])
@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_aos_batches_through_tma_tiles(L, sig, dt):
    """Reference-layout (N, num_blades) batches of >= 4096 rows take the TMA-staged kernels
    (kernels/tma_tile.cuh): ragged last tile, padded row pitch, broadcast operands, the fused dual-output
    form and the sandwich, all bit-exact in strict mode; a pitch TMA cannot describe (not a multiple of
    16 bytes) must silently take the direct-load kernels and give the same bits."""
    from largecrimsoncanine_b200 import cabi

    lib = cabi.lib()
    o, alg = oracle.OracleAlgebra(*sig), cabi.Algebra(*sig)
    nb = o.num_blades
    code = cabi.LCC_F64 if dt == np.float64 else cabi.LCC_F32
    rng = np.random.default_rng(23)
    S = cabi.LCC_FLAG_STRICT_FP
    dev = _Dev(lib)
    cabi.check(lib.lcc_set_option(b"aos_tma_min_rows", 128))  # default 4096 (x32 for rows under 128 bytes): force the tiles at test sizes
    try:
        _aos_tma_cases(lib, alg, o, nb, code, rng, S, dev, dt)
    finally:
        cabi.check(lib.lcc_set_option(b"aos_tma_min_rows", 4096))


def _aos_tma_cases(lib, alg, o, nb, code, rng, S, dev, dt):
    from largecrimsoncanine_b200 import cabi

    for n, pad in ((4096, 0), (4096 + 37, 0), (10007, 0), (5000, 16 // np.dtype(dt).itemsize), (4500, 1)):
        ld = nb + pad
        a, b = rand(rng, n, nb, dt), rand(rng, n, nb, dt)

        def image(x):
            img = np.full((x.shape[0], ld), 7.0, dt)  # pitch padding must be neither read as data nor written
            img[:, :nb] = x
            return img

        ia, ib = image(a), image(b)
        pa, pb, po, po2 = dev.put(ia), dev.put(ib), dev.put(image(np.zeros_like(a))), dev.put(image(np.zeros_like(a)))
        va, vb, vo, vo2 = (cabi.view(p.value, n, ld, code, cabi.LCC_AOS) for p in (pa, pb, po, po2))
        for op, ref in ((cabi.LCC_OP_GP, o.geometric_product), (cabi.LCC_OP_OUTER, o.outer_product), (cabi.LCC_OP_LC, o.left_contraction)):
            cabi.check(lib.lcc_batch_product(alg.h, op, C.byref(va), C.byref(vb), C.byref(vo), S, None))
            got = dev.get(po, ia)
            assert np.array_equal(got[:, :nb], ref(a, b)), (n, pad, op)
            assert np.all(got[:, nb:] == 7.0)
        cabi.check(lib.lcc_batch_gp_outer(alg.h, C.byref(va), C.byref(vb), C.byref(vo), C.byref(vo2), S, None))
        assert np.array_equal(dev.get(po, ia)[:, :nb], o.geometric_product(a, b))
        assert np.array_equal(dev.get(po2, ia)[:, :nb], o.outer_product(a, b))
        # FMA mode stays within the stated tolerance
        cabi.check(lib.lcc_batch_product(alg.h, cabi.LCC_OP_GP, C.byref(va), C.byref(vb), C.byref(vo), 0, None))
        assert_close(dev.get(po, ia)[:, :nb], o.geometric_product(a, b), o.product_abs_bound("gp", a, b), dt)
        # broadcast on either side (batch.rs:361-376)
        one = rand(rng, 1, nb, dt)
        p1 = dev.put(np.ascontiguousarray(one))
        v1 = cabi.view(p1.value, 1, nb, code, cabi.LCC_AOS)
        cabi.check(lib.lcc_batch_product(alg.h, cabi.LCC_OP_GP, C.byref(v1), C.byref(vb), C.byref(vo), S, None))
        assert np.array_equal(dev.get(po, ia)[:, :nb], o.geometric_product(one, b))
        cabi.check(lib.lcc_batch_product(alg.h, cabi.LCC_OP_LC, C.byref(va), C.byref(v1), C.byref(vo), S, None))
        assert np.array_equal(dev.get(po, ia)[:, :nb], o.left_contraction(a, one))
        # sandwich: dense rotor and even rotor (the pruned kernel)
        dense = rng.standard_normal(nb)
        even = dense * np.array([1.0 if bin(i).count("1") % 2 == 0 else 0.0 for i in range(nb)])
        for rotor in (dense, even):
            cabi.check(lib.lcc_batch_sandwich(alg.h, rotor.ctypes.data_as(C.POINTER(C.c_double)), C.byref(va), C.byref(vo), S, None))
            got = dev.get(po, ia)
            assert np.array_equal(got[:, :nb], o.sandwich(rotor, a)), (n, pad)
            assert np.all(got[:, nb:] == 7.0)
        dev.free()


@pytest.mark.parametrize("sig", [(3, 0, 0), (3, 0, 1), (4, 1, 0), (1, 3, 0), (2, 1, 1), (1, 0, 0), (5, 0, 0)])
@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_soa_batches_through_tma_tiles(L, strict, sig, dt):
    """Device-resident (blade-major) batches of >= 4096 rows take the TMA-staged SoA kernels: ragged last
    tile, broadcast operands, fused dual output and sandwich, bit-exact in strict mode."""
    alg, o = algebras(L, sig)
    nb = o.num_blades
    rng = np.random.default_rng(29)
    before = L.get_option("soa_tma_min_bytes")
    L.set_option("soa_tma_min_bytes", 0)  # the default (128 MiB per operand) would send these sizes to the direct kernels
    try:
        _soa_tma_cases(L, alg, o, nb, rng, dt)
    finally:
        L.set_option("soa_tma_min_bytes", before)


def _soa_tma_cases(L, alg, o, nb, rng, dt):
    for n in (4096, 4099, 9001):
        a, b = rand(rng, n, nb, dt), rand(rng, n, nb, dt)
        A, B = L.MultivectorBatch.from_numpy(alg, a), L.MultivectorBatch.from_numpy(alg, b)
        assert np.array_equal(A.geometric_product(B).to_numpy(), o.geometric_product(a, b))
        assert np.array_equal(A.outer_product(B).to_numpy(), o.outer_product(a, b))
        assert np.array_equal(A.left_contraction(B).to_numpy(), o.left_contraction(a, b))
        g, w = A.gp_and_outer(B)
        assert np.array_equal(g.to_numpy(), o.geometric_product(a, b)) and np.array_equal(w.to_numpy(), o.outer_product(a, b))
        one = rand(rng, 1, nb, dt)
        One = L.MultivectorBatch.from_numpy(alg, one)
        assert np.array_equal(One.geometric_product(B).to_numpy(), o.geometric_product(one, b))
        assert np.array_equal(A.outer_product(One).to_numpy(), o.outer_product(a, one))
        dense = rng.standard_normal(nb)
        even = dense * np.array([1.0 if bin(i).count("1") % 2 == 0 else 0.0 for i in range(nb)])
        for rotor in (dense, even):
            assert np.array_equal(A.sandwich(rotor_mv(L, alg, rotor)).to_numpy(), o.sandwich(rotor, a))
        # inputs must be left untouched by the in-place tile reuse
        assert np.array_equal(A.to_numpy(), a) and np.array_equal(B.to_numpy(), b)


# ----------------------------------------------------------------------------------------------- rest of the product family
@pytest.mark.parametrize("sig", [(3, 0, 0), (3, 0, 1), (4, 1, 0), (1, 3, 0), (2, 1, 1), (2, 0, 0), (6, 0, 0)])
@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_product_family_batched(L, sig, dt):
    """Batched forms of Multivector::{right_contraction, scalar_product, commutator, anticommutator,
    regressive} (src/multivector.rs:1616-1797): bit-exact in strict mode (direct and TMA-tile kernels),
    within tolerance in FMA mode, broadcast rule, and the reference's error for a null pseudoscalar."""
    alg, o = algebras(L, sig)
    nb = o.num_blades
    rng = np.random.default_rng(37)
    degenerate = sig[2] > 0
    for n in ((3, 300) if nb > 32 else (1, 77, 1030, 4100)):
        a, b = rand(rng, n, nb, dt), rand(rng, n, nb, dt)
        A, B = L.MultivectorBatch.from_numpy(alg, a), L.MultivectorBatch.from_numpy(alg, b)
        L.set_strict_fp(True)
        try:
            assert np.array_equal(A.right_contraction(B).to_numpy(), o.right_contraction(a, b))
            assert np.array_equal(A.rc(B).to_numpy(), o.right_contraction(a, b))
            assert np.array_equal(A.scalar_product(B).to_numpy(), o.scalar_product(a, b))
            assert np.array_equal(A.commutator(B).to_numpy(), o.commutator(a, b))
            assert np.array_equal(A.anticommutator(B).to_numpy(), o.anticommutator(a, b))
            if degenerate:
                with pytest.raises(ValueError):
                    A.regressive(B)
            else:
                assert np.array_equal(A.regressive(B).to_numpy(), o.regressive(a, b))
                assert np.array_equal(A.meet(B).to_numpy(), o.regressive(a, b))
        finally:
            L.set_strict_fp(False)
        assert_close(A.right_contraction(B).to_numpy(), o.right_contraction(a, b), o.product_abs_bound("rc", a, b), dt)
        assert_close(A.scalar_product(B).to_numpy(), o.scalar_product(a, b), o.product_abs_bound("scalar", a, b), dt)
        gpb = o.product_abs_bound("gp", a, b)
        assert_close(A.commutator(B).to_numpy(), o.commutator(a, b), gpb, dt)
        assert_close(A.anticommutator(B).to_numpy(), o.anticommutator(a, b), gpb, dt)
    one = rand(rng, 1, nb, dt)
    One = L.MultivectorBatch.from_numpy(alg, one)
    L.set_strict_fp(True)
    try:
        assert np.array_equal(One.commutator(B).to_numpy(), o.commutator(one, b))
        assert np.array_equal(A.right_contraction(One).to_numpy(), o.right_contraction(a, one))
    finally:
        L.set_strict_fp(False)


def test_product_family_scalar_multivectors(L):
    """The per-object methods go through the same kernels with n = 1 (known answers of
    /root/reference/tests/test_basic.py:2275-2316,2369-2378,2407-2418,2528-2557)."""
    M = L.Multivector
    a, b = M.from_vector([1.0, 2.0, 3.0]), M.from_vector([4.0, 5.0, 6.0])
    assert a.right_contraction(b).scalar() == 32.0
    e12, e13, e1 = M.from_bivector([1.0, 0.0, 0.0], dims=3), M.from_bivector([0.0, 1.0, 0.0], dims=3), M.from_vector([1.0, 0.0, 0.0])
    assert e12.rc(e1).approx_eq(M.from_vector([0.0, -1.0, 0.0]))
    assert e12.scalar_product(e12).scalar() == -1.0 and e12.dot(e12).scalar() == -1.0
    f1, f2 = M.from_vector([1.0, 0.0]), M.from_vector([0.0, 1.0])
    assert f1.commutator(f2).approx_eq(M.from_bivector([1.0], dims=2)) and f1.x(f2).approx_eq(f1.commutator(f2))
    assert f1.anticommutator(f1).scalar() == 1.0
    meet = e12.regressive(e13)
    assert meet.grade(1).norm() > 0.1 and meet.approx_eq(meet.grade(1)) and e12.meet(e13).approx_eq(meet) and e12.vee(e13).approx_eq(meet)
    with pytest.raises(ValueError, match="dimension mismatch"):
        a.commutator(f1)
    with pytest.raises(ValueError, match="algebra mismatch"):
        a.right_contraction(f1)


# ----------------------------------------------------------------------------------------------- K3 / K4 / K5 on reference-layout views
@pytest.mark.parametrize("sig", [(3, 0, 0), (3, 0, 1), (4, 1, 0), (2, 0, 0), (1, 0, 0), (6, 0, 0), (8, 0, 0)])
@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_unary_norm_sum_on_aos_views(L, sig, dt):
    """Row-wise maps, add / sub, norms, normalized and the batch sum through the raw C ABI on (N, num_blades)
    row-major views (what lcc.torch passes): the 128-bit chunk kernels, the lane-group norm (which keeps the
    reference's ascending-blade summation order, so it is bit-exact in strict mode) and their scalar fallbacks
    (row pitch not a multiple of 16 bytes)."""
    from largecrimsoncanine_b200 import cabi

    lib = cabi.lib()
    o, alg = oracle.OracleAlgebra(*sig), cabi.Algebra(*sig)
    nb = o.num_blades
    code = cabi.LCC_F64 if dt == np.float64 else cabi.LCC_F32
    rng = np.random.default_rng(41)
    S = cabi.LCC_FLAG_STRICT_FP
    dev = _Dev(lib)
    for n, pad in ((1, 0), (37, 0), (1000, 0), (513, 1), (5003, 0)):  # from 4096 rows on: the TMA-tiled norm kernels (rows of 32-256 bytes)
        ld = nb + pad
        a, b = rand(rng, n, nb, dt), rand(rng, n, nb, dt)
        a[n // 2] = 0.0  # a zero row: normalized leaves it unchanged (batch.rs:805-811)

        def image(x):
            img = np.full((x.shape[0], ld), 7.0, dt)
            img[:, :nb] = x
            return img

        ia = image(a)
        pa, pb, po = dev.put(ia), dev.put(image(b)), dev.put(image(np.zeros_like(a)))
        va, vb, vo = (cabi.view(p.value, n, ld, code, cabi.LCC_AOS) for p in (pa, pb, po))

        def result():
            got = dev.get(po, ia)
            assert np.all(got[:, nb:] == 7.0)
            return got[:, :nb]

        for op, ref in ((cabi.LCC_UN_REVERSE, o.reverse), (cabi.LCC_UN_INVOLUTE, o.grade_involution), (cabi.LCC_UN_CONJUGATE, o.clifford_conjugate)):
            cabi.check(lib.lcc_batch_unary(alg.h, op, 0, C.byref(va), C.byref(vo), None))
            assert np.array_equal(result(), ref(a))
        for g in range(0, o.dimension + 1, max(1, o.dimension // 2)):
            cabi.check(lib.lcc_batch_unary(alg.h, cabi.LCC_UN_GRADE, g, C.byref(va), C.byref(vo), None))
            assert np.array_equal(result(), o.grade(a, g))
        if dt == np.float64 and sig[2] == 0:  # the permuting (non-diagonal) map
            cabi.check(lib.lcc_batch_unary(alg.h, cabi.LCC_UN_DUAL, 0, C.byref(va), C.byref(vo), None))
            assert np.array_equal(result(), o.dual(a))
        cabi.check(lib.lcc_batch_scale(alg.h, C.byref(va), 0.375, C.byref(vo), None))
        assert np.array_equal(result(), o.scale(a, 0.375))
        for op, ref in ((cabi.LCC_EW_ADD, o.add), (cabi.LCC_EW_SUB, o.sub)):
            cabi.check(lib.lcc_batch_elementwise(alg.h, op, C.byref(va), C.byref(vb), C.byref(vo), None))
            assert np.array_equal(result(), ref(a, b))
        one = rand(rng, 1, nb, dt)
        p1 = dev.put(np.ascontiguousarray(one))
        v1 = cabi.view(p1.value, 1, nb, code, cabi.LCC_AOS)
        if n > 1:
            cabi.check(lib.lcc_batch_elementwise(alg.h, cabi.LCC_EW_SUB, C.byref(va), C.byref(v1), C.byref(vo), None))
            assert np.array_equal(result(), o.sub(a, one))
        norms = np.zeros(n, dt)
        pn = dev.put(norms)
        for kind, ref in ((cabi.LCC_NORM_SQUARED, o.norm_squared), (cabi.LCC_NORM, o.norm)):
            cabi.check(lib.lcc_batch_norm(alg.h, kind, C.byref(va), pn, S, None))
            assert np.array_equal(dev.get(pn, norms), ref(a))
            cabi.check(lib.lcc_batch_norm(alg.h, kind, C.byref(va), pn, 0, None))
            if kind == cabi.LCC_NORM_SQUARED:  # |error| <= tol * sum_k a_k^2 (the cancellation-free bound)
                assert np.all(np.abs(dev.get(pn, norms).astype(np.float64) - ref(a)) <= TOL[dt] * (a.astype(np.float64) ** 2).sum(axis=1) + 1e-300)
        cabi.check(lib.lcc_batch_normalized(alg.h, C.byref(va), C.byref(vo), S, None))
        assert np.array_equal(result(), o.normalized(a))
        if sig[1] == 0:  # definite signature: no cancellation inside norm^2
            cabi.check(lib.lcc_batch_normalized(alg.h, C.byref(va), C.byref(vo), 0, None))
            np.testing.assert_allclose(result(), o.normalized(a), rtol=10 * TOL[dt], atol=TOL[dt])
        sums = np.zeros(nb, dt)
        ps = dev.put(sums)
        cabi.check(lib.lcc_batch_sum(alg.h, C.byref(va), ps, None))
        want, mag = o.sum(a)
        assert np.all(np.abs(dev.get(ps, sums) - want) <= (1e-12 if dt == np.float64 else 2e-6) * np.maximum(mag, 1.0))
        dev.free()


@pytest.mark.parametrize("sig", [(3, 0, 0), (3, 0, 1), (4, 1, 0), (2, 0, 0)])
@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_normalized_soa_register_kernel(L, strict, sig, dt):
    """`normalized` on device-resident batches: the compile-time-NB kernel (rows in registers) for row counts
    that are whole vectors, the generic kernel for the rest; both bit-exact in strict mode."""
    alg, o = algebras(L, sig)
    rng = np.random.default_rng(43)
    for n in (4, 64, 1000, 1001, 4096, 70_000, 65_537):  # from 65536 rows on the TMA-tiled form may run (whole 16-byte row counts)
        a = rand(rng, n, o.num_blades, dt)
        a[0] = 0.0
        assert np.array_equal(L.MultivectorBatch.from_numpy(alg, a).normalized().to_numpy(), o.normalized(a))
    if sig[1] == 0:  # definite signature: norm^2 has no cancellation, so the FMA / one-division-per-row form stays close
        L.set_strict_fp(False)
        np.testing.assert_allclose(L.MultivectorBatch.from_numpy(alg, a).normalized().to_numpy(), o.normalized(a), rtol=10 * TOL[dt], atol=TOL[dt])


# ----------------------------------------------------------------------------------------------- compact PGA points
@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_pga_points_fused_sandwich(L, dt):
    """(N,3) xyz -> motor -> (N,3) xyz in one kernel: bit-exact in strict mode against embed + sandwich + extract
    of the oracle (src/pga.rs:131-134, batch.rs:434-503, pga.rs:579-599), full bulk-copied blocks and the ragged
    tail, even motors (pruned kernel) and dense versors, and it agrees with the unfused batched path."""
    pga = L.Algebra.pga(3)
    rng = np.random.default_rng(47)
    motor = L.Multivector.pga_motor_from_axis_angle(pga, [1.0, 2.0, 3.0], [0.5, -0.25, 1.0], 0.7)
    dense = L.Multivector.from_list_in([float(v) for v in rng.standard_normal(16)], pga)
    for n in (1, 5, 1024, 1027, 5000):
        xyz = rng.standard_normal((n, 3)).astype(dt)
        for M in (motor, dense):
            m = np.array(M.coefficients())
            L.set_strict_fp(True)
            try:
                got = L.pga_transform_points(pga, M, xyz)
            finally:
                L.set_strict_fp(False)
            want = oracle.pga_transform_points(m, xyz)
            assert got.dtype == dt and np.array_equal(got, want, equal_nan=True), (n, np.abs(got - want).max())
            fast = L.pga_transform_points(pga, M, xyz)
            np.testing.assert_allclose(fast, want, rtol=10 * TOL[dt], atol=10 * TOL[dt])
        unfused = L.MultivectorBatch.from_points_pga(pga, xyz).sandwich(motor).pga_point_coords()
        np.testing.assert_allclose(L.pga_transform_points(pga, motor, xyz), unfused, rtol=10 * TOL[dt], atol=10 * TOL[dt])
    with pytest.raises(ValueError):
        L.pga_transform_points(L.Algebra.cga(3), motor, xyz)


def test_second_device_in_the_same_process(L):
    """One process driving two GPUs (lcc_set_device): kernels that need more than 48 KB of dynamic shared memory
    (the TMA-staged SoA products, K6 / K6d) must get their launch attribute on every device, not only on the first
    one that ran them."""
    if L.device_count() < 2:
        pytest.skip("needs two GPUs")
    alg, o = algebras(L, (4, 1, 0))
    rng = np.random.default_rng(59)
    a, b = rand(rng, 5000, 32, np.float64), rand(rng, 5000, 32, np.float64)
    before = L.get_option("soa_tma_min_bytes")
    L.set_option("soa_tma_min_bytes", 0)
    L.set_strict_fp(True)
    try:
        for dev in (0, 1, 0):
            L.set_device(dev)
            A, B = L.MultivectorBatch.from_numpy(alg, a), L.MultivectorBatch.from_numpy(alg, b)
            assert np.array_equal(A.geometric_product(B).to_numpy(), o.geometric_product(a, b))
            assert np.array_equal(A.sandwich(rotor_mv(L, alg, a[0].astype(np.float64))).to_numpy(), o.sandwich(a[0].astype(np.float64), a))
    finally:
        L.set_strict_fp(False)
        L.set_option("soa_tma_min_bytes", before)
        L.set_device(0)
    alg8, o8 = algebras(L, (8, 0, 0))
    m, x = rng.standard_normal((1, 256)), rng.standard_normal((300, 256))
    L.set_device(1)
    try:
        for dt, tol in ((np.float64, 1e-12), (np.float32, 1e-5)):
            got = L.MultivectorBatch.from_numpy(alg8, m.astype(dt)).geometric_product(L.MultivectorBatch.from_numpy(alg8, x.astype(dt))).to_numpy()
            ref = o8.geometric_product(m.astype(dt).astype(np.float64), x.astype(dt).astype(np.float64))
            assert np.abs(got - ref).max() <= tol * np.abs(ref).max()
    finally:
        L.set_device(0)


@pytest.mark.parametrize("sig", [(3, 0, 0), (3, 0, 1), (4, 1, 0), (2, 1, 1), (6, 0, 0)])
@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_sandwich_with_one_rotor_per_row(L, sig, dt):
    """lcc_batch_sandwich_rows: out[r] = R[r] * x[r] * ~R[r] (lcc/torch/multivector.py:988-1019 of the reference,
    two products in the reference's term order), both layouts, strict mode bit-exact against the oracle's composition."""
    from largecrimsoncanine_b200 import cabi

    lib = cabi.lib()
    o, alg = oracle.OracleAlgebra(*sig), cabi.Algebra(*sig)
    nb = o.num_blades
    code = cabi.LCC_F64 if dt == np.float64 else cabi.LCC_F32
    rng = np.random.default_rng(61)
    dev = _Dev(lib)
    for n in ((3, 200) if nb > 32 else (1, 77, 5000)):
        x, r = rand(rng, n, nb, dt), rand(rng, n, nb, dt)
        want = o.geometric_product(o.geometric_product(r, x), o.reverse(r))
        for layout in (cabi.LCC_AOS, cabi.LCC_SOA):
            if layout == cabi.LCC_AOS:
                ix, ir, ld = np.ascontiguousarray(x), np.ascontiguousarray(r), nb
            else:
                ld = ((n + 63) // 64) * 64
                ix, ir = np.zeros((nb, ld), dt), np.zeros((nb, ld), dt)
                ix[:, :n], ir[:, :n] = x.T, r.T
            px, pr, po = dev.put(ix), dev.put(ir), dev.put(np.zeros_like(ix))
            vx, vr, vo = (cabi.view(p.value, n, ld, code, layout) for p in (px, pr, po))
            cabi.check(lib.lcc_batch_sandwich_rows(alg.h, C.byref(vr), C.byref(vx), C.byref(vo), cabi.LCC_FLAG_STRICT_FP, None))
            got = dev.get(po, ix)
            got = got if layout == cabi.LCC_AOS else got[:, :n].T
            assert np.array_equal(got, want), (sig, dt, n, layout)
            dev.free()


@pytest.mark.parametrize("sig", [(3, 0, 0), (3, 0, 1), (4, 1, 0), (2, 0, 0), (6, 0, 0)])
@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_layout_conversion_through_tma_tiles(L, sig, dt):
    """lcc_batch_convert (what from_numpy / to_numpy run, batch.rs:79-94,233-235): (N, num_blades) row-major <->
    blade-major, bit-exact copies in both directions.  From 4096 rows on, rows of 32-256 bytes go through the
    TMA-staged transpose; ragged last tiles, the padding of the blade rows (canaries) and views the engine cannot
    describe (pitch not a multiple of 16 bytes) are covered, as is the scalar kernel below the threshold."""
    from largecrimsoncanine_b200 import cabi

    lib = cabi.lib()
    alg = cabi.Algebra(*sig)
    nb = 1 << sum(sig)
    code = cabi.LCC_F64 if dt == np.float64 else cabi.LCC_F32
    rng = np.random.default_rng(53)
    dev = _Dev(lib)
    for n, aos_pad in ((1000, 0), (4096, 0), (4097, 0), (70_001, 0), (33_333, 4), (5000, 1)):
        x = rand(rng, n, nb, dt)
        ld_aos = nb + aos_pad
        ld_soa = ((n + 63) // 64) * 64
        aos_img = np.full((n, ld_aos), 7.0, dt)
        aos_img[:, :nb] = x
        soa_canary = np.full((nb, ld_soa), -3.0, dt)
        pa, ps, pb = dev.put(aos_img), dev.put(soa_canary), dev.put(np.full_like(aos_img, 7.0))
        va = cabi.view(pa.value, n, ld_aos, code, cabi.LCC_AOS)
        vs = cabi.view(ps.value, n, ld_soa, code, cabi.LCC_SOA)
        vb = cabi.view(pb.value, n, ld_aos, code, cabi.LCC_AOS)
        cabi.check(lib.lcc_batch_convert(alg.h, C.byref(va), C.byref(vs), None))
        soa = dev.get(ps, soa_canary)
        assert np.array_equal(soa[:, :n], x.T)
        assert np.all(soa[:, n:] == -3.0)  # the tail of every blade row is not touched
        cabi.check(lib.lcc_batch_convert(alg.h, C.byref(vs), C.byref(vb), None))
        back = dev.get(pb, aos_img)
        assert np.array_equal(back[:, :nb], x)
        assert np.all(back[:, nb:] == 7.0)
        dev.free()


def test_host_pga_points_entry_point(L):
    """lcc_host_pga_points_sandwich: (N,3) host arrays through the chunked H2D -> fused kernel -> D2H pipeline, several
    chunks of 64 MiB and a ragged last one, out of place and in place; strict mode is bit-exact against the oracle's
    embed + sandwich + extract (src/pga.rs:131-134, batch.rs:434-503, pga.rs:579-599)."""
    from largecrimsoncanine_b200 import cabi

    lib = cabi.lib()
    alg = cabi.Algebra(3, 0, 1)
    rng = np.random.default_rng(59)
    n = 12_000_003  # > 2 chunks at 12 B/point
    xyz = rng.standard_normal((n, 3), dtype=np.float32)
    motor = np.zeros(16)
    motor[[0, 3, 5, 6, 9, 10, 12]] = rng.standard_normal(7)
    mptr = motor.ctypes.data_as(C.POINTER(C.c_double))
    out = np.empty_like(xyz)
    cabi.check(lib.lcc_host_pga_points_sandwich(alg.h, cabi.LCC_F32, mptr, xyz.ctypes.data, out.ctypes.data, n, cabi.LCC_FLAG_STRICT_FP))
    sample = np.r_[0:4096, n // 2:n // 2 + 4096, n - 4096:n]
    want = oracle.pga_transform_points(motor, xyz[sample])
    assert np.array_equal(out[sample], want, equal_nan=True)
    inplace = xyz.copy()
    cabi.check(lib.lcc_host_pga_points_sandwich(alg.h, cabi.LCC_F32, mptr, inplace.ctypes.data, inplace.ctypes.data, n, cabi.LCC_FLAG_STRICT_FP))
    assert np.array_equal(inplace, out, equal_nan=True)
    small = rng.standard_normal((5, 3))
    res = np.empty_like(small)
    cabi.check(lib.lcc_host_pga_points_sandwich(alg.h, cabi.LCC_F64, mptr, small.ctypes.data, res.ctypes.data, 5, cabi.LCC_FLAG_STRICT_FP))
    assert np.array_equal(res, oracle.pga_transform_points(motor, small), equal_nan=True)
    assert lib.lcc_host_pga_points_sandwich(cabi.Algebra(3, 0, 0).h, cabi.LCC_F32, mptr, xyz.ctypes.data, out.ctypes.data, 8, 0) == cabi.LCC_ERR_VALUE
