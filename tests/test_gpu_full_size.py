"""Parity at BASELINE.json's FULL sizes, through properties that do not need a CPU pass over 2^24..2^28 rows.

The oracle finishes a few thousand rows per millisecond, so the row-by-row comparisons of tests/test_gpu_parity.py
stop at ~10^5 rows.  Here every configuration of BASELINE.json runs at its quoted size through the C ABI and is
checked over ALL rows on the device with identities of the algebra that are independent of the kernels under test
(torch elementwise arithmetic and reductions are the checker here, never the product), plus an oracle comparison of
the first, the last and a random sample of rows:

  cfg1  Cl(3,0,0) gp, 2^20 pairs f64         every row against the oracle, bit-exact in strict mode
  cfg2  PGA motor sandwich, 2^28 points f32  rigid motion: weights stay 1, non-point blades stay 0, consecutive
                                             distances are preserved, the reverse motor undoes it
  cfg3  CGA gp + outer, 2^26 pairs f64       scaling an operand by 2 scales the result by 2 BIT-EXACTLY; the scalar
                                             part equals sum_k sign(k,k) a_k b_k; <a ^ b>_0 == a_0 b_0 exactly;
                                             fused gp+outer == the two separate products, bit for bit
  cfg4  STA grade + inner + batch sum, 2^27  fused == unfused pipeline; doubling an operand doubles the sums exactly;
                                             the scalar column against a torch reduction
  cfg5  Cl(8,0,0) fixed operand, 2^24 rows   a basis blade as the fixed operand is a signed permutation of the blades
                                             (tensor-core path, tolerance 1e-5); dense operand on sampled rows

Tolerances are BASELINE's: 1e-12 relative (f64), 1e-5 relative (f32), against the a-priori magnitude of the terms.
"""
import ctypes as C

import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env():
    import torch

    from largecrimsoncanine_b200 import cabi

    lib = cabi.lib()
    assert torch.cuda.is_available()
    dev = torch.device("cuda", 0)
    cabi.check(lib.lcc_set_device(0))
    stream = torch.cuda.Stream(device=dev)  # non-default: a NULL handle would mean "the library's own stream"
    assert stream.cuda_stream != 0

    class Env:
        pass

    e = Env()
    e.torch, e.cabi, e.lib, e.dev, e.stream, e.sptr = torch, cabi, lib, dev, stream, C.c_void_p(stream.cuda_stream)
    with torch.cuda.stream(stream):
        yield e
    torch.cuda.synchronize()
    torch.cuda.empty_cache()


def need_gb(env, gb):
    env.torch.cuda.empty_cache()
    free, _ = env.torch.cuda.mem_get_info()
    if free < gb * 2**30:
        pytest.skip(f"needs {gb} GiB of free device memory, {free / 2**30:.0f} GiB available")


def soa_view(env, t, code):  # torch (nb, n) tensor == blade-major batch with ld = row stride
    return env.cabi.view(t.data_ptr(), t.shape[1], t.stride(0), code, env.cabi.LCC_SOA)


def sample_rows(n, seed):
    rng = np.random.default_rng(seed)
    return np.unique(np.r_[0:2048, n - 2048:n, rng.integers(0, n, 4096)])


def rows_to_host(env, t, idx):
    return t[:, env.torch.as_tensor(idx, device=env.dev)].T.contiguous().cpu().numpy()


def pga_motor():
    """A unit motor: rotation by 0.7 about the axis (1,2,3) through (0.5,-0.25,1), then translation by (4,5,6)
    (the bench's headline motor; src/pga.rs:287-345,386-389)."""
    axis = np.array([1.0, 2.0, 3.0])
    axis /= np.linalg.norm(axis)
    p = np.array([0.5, -0.25, 1.0])
    c, s = np.cos(0.35), np.sin(0.35)
    R = np.zeros(16)
    R[0], R[6], R[5], R[3] = c, -s * axis[0], s * axis[1], -s * axis[2]
    mom = np.cross(p, axis)
    R[9], R[10], R[12] = s * mom[0], s * mom[1], s * mom[2]
    T = np.zeros(16)
    T[0], T[9], T[10], T[12] = 1.0, 2.0, 2.5, 3.0
    return oracle.OracleAlgebra(3, 0, 1).geometric_product(T[None, :], R[None, :])[0]


# ----------------------------------------------------------------------------------------------- cfg1
def test_cfg1_r3_gp_every_row_against_the_oracle():
    """BASELINE configs[0]: Cl(3,0,0) gp on 2^20 pairs, f64, through the reference's batch API."""
    import largecrimsoncanine as L

    n = 1 << 20
    rng = np.random.default_rng(42)
    a, b = rng.standard_normal((n, 8)), rng.standard_normal((n, 8))
    alg, o = L.Algebra(3), oracle.OracleAlgebra(3, 0, 0)
    A, B = L.MultivectorBatch.from_numpy(alg, a), L.MultivectorBatch.from_numpy(alg, b)
    want = o.geometric_product(a, b)
    L.set_strict_fp(True)
    try:
        assert np.array_equal(A.geometric_product(B).to_numpy(), want)
    finally:
        L.set_strict_fp(False)
    got = A.geometric_product(B).to_numpy()
    bound = o.product_abs_bound("gp", np.abs(a), np.abs(b))
    assert np.all(np.abs(got - want) <= 1e-12 * np.maximum(np.abs(want), bound))


# ----------------------------------------------------------------------------------------------- cfg2
def test_cfg2_pga_motor_sandwich_is_a_rigid_motion_on_2p28_points(env):
    """BASELINE configs[1] at full size: 2^28 points, f32, 16 blades (3 x 17.2 GB resident)."""
    need_gb(env, 75)
    torch, cabi, lib = env.torch, env.cabi, env.lib
    n, nb = 1 << 28, 16
    alg = cabi.Algebra(3, 0, 1)
    gen = torch.Generator(device=env.dev)
    gen.manual_seed(42)
    x = torch.zeros((nb, n), device=env.dev, dtype=torch.float32)
    pts = torch.randn((3, n), device=env.dev, dtype=torch.float32, generator=gen)
    x[7].fill_(1.0)  # src/pga.rs:131-134
    x[14].copy_(-pts[0]); x[13].copy_(pts[1]); x[11].copy_(-pts[2])
    del pts
    out, back = torch.empty_like(x), torch.empty_like(x)
    M = pga_motor()
    rev = np.array([1.0 if (bin(i).count("1") % 4) < 2 else -1.0 for i in range(16)])
    Mrev = M * rev
    vx, vo, vb = (soa_view(env, t, cabi.LCC_F32) for t in (x, out, back))
    dptr = lambda v: v.ctypes.data_as(C.POINTER(C.c_double))  # noqa: E731
    cabi.check(lib.lcc_batch_sandwich(alg.h, dptr(M), C.byref(vx), C.byref(vo), 0, env.sptr))
    cabi.check(lib.lcc_batch_sandwich(alg.h, dptr(Mrev), C.byref(vo), C.byref(vb), 0, env.sptr))
    tol = 1e-5
    m1 = float(np.abs(M).sum())
    # (a) a motor maps points to points: weight e123 stays 1, every non-trivector blade stays 0
    coord_mag = float(torch.stack([x[k].abs().max() for k in (11, 13, 14)]).max()) + 1.0
    assert float((out[7] - 1.0).abs().max()) <= tol * m1 * m1
    for k in range(nb):
        if k not in (7, 11, 13, 14):
            assert float(out[k].abs().max()) <= tol * m1 * m1 * coord_mag, k
    # (b) distances between consecutive points are preserved (x = -e032, y = e013, z = -e021 coefficients)
    def dist2(t):
        return sum((t[k][1:] - t[k][:-1]) ** 2 for k in (14, 13, 11))
    d_in, d_out = dist2(x), dist2(out)
    out_mag = float(torch.stack([out[k].abs().max() for k in (11, 13, 14)]).max()) + 1.0
    assert float((d_out - d_in).abs().max()) <= 4 * tol * m1 * m1 * out_mag  # 2 |d| * (coordinate error), |d| <= 2 out_mag
    del d_in, d_out
    # (c) the reverse motor undoes it: ~M (M x ~M) M == x, every row (two sandwiches, each within tol * |M|_1^2 * |row|)
    assert float((back - x).abs().max()) <= 2 * tol * m1 * m1 * out_mag
    # (d) sampled rows against the oracle
    idx = sample_rows(n, 1)
    want = oracle.OracleAlgebra(3, 0, 1).sandwich(M, rows_to_host(env, x, idx))
    got = rows_to_host(env, out, idx)
    assert np.all(np.abs(got - want) <= tol * np.maximum(np.abs(want), m1 * m1 * coord_mag))


# ----------------------------------------------------------------------------------------------- cfg3
def test_cfg3_cga_gp_and_outer_on_2p26_pairs(env):
    """BASELINE configs[2] at full size: 2^26 pairs, f64, 32 blades (5 x 17.2 GB resident)."""
    need_gb(env, 95)
    torch, cabi, lib = env.torch, env.cabi, env.lib
    n, nb = 1 << 26, 32
    alg, o = cabi.Algebra(4, 1, 0), oracle.OracleAlgebra(4, 1, 0)
    gen = torch.Generator(device=env.dev)
    gen.manual_seed(43)
    a = torch.randn((nb, n), device=env.dev, dtype=torch.float64, generator=gen)
    b = torch.randn((nb, n), device=env.dev, dtype=torch.float64, generator=gen)
    gp, outer, gp2 = torch.empty_like(a), torch.empty_like(a), torch.empty_like(a)
    va, vb, vg, vo, vg2 = (soa_view(env, t, cabi.LCC_F64) for t in (a, b, gp, outer, gp2))
    cabi.check(lib.lcc_batch_gp_outer(alg.h, C.byref(va), C.byref(vb), C.byref(vg), C.byref(vo), 0, env.sptr))
    # (a) fused gp+outer == the separate products, bit for bit (same terms in the same order)
    cabi.check(lib.lcc_batch_product(alg.h, cabi.LCC_OP_GP, C.byref(va), C.byref(vb), C.byref(vg2), 0, env.sptr))
    assert torch.equal(gp2, gp)
    cabi.check(lib.lcc_batch_product(alg.h, cabi.LCC_OP_OUTER, C.byref(va), C.byref(vb), C.byref(vg2), 0, env.sptr))
    assert torch.equal(gp2, outer)
    # (b) scalar part of a*b = sum_k sign(k,k) a_k b_k, every row, against a torch reduction in another order
    signs = np.ctypeslib.as_array(lib.lcc_algebra_cayley_signs(alg.h), shape=(nb, nb))
    ref0, mag0 = torch.zeros(n, device=env.dev, dtype=torch.float64), torch.zeros(n, device=env.dev, dtype=torch.float64)
    for k in range(nb):
        t = a[k] * b[k]
        ref0.add_(t, alpha=float(signs[k, k]))
        mag0.add_(t.abs_())
    assert bool(((gp[0] - ref0).abs() <= 1e-12 * mag0).all())
    del ref0, mag0
    # (c) <a ^ b>_0 has the single term a_0 b_0: exact
    assert torch.equal(outer[0], a[0] * b[0])
    # (d) sampled rows against the oracle (FMA mode: 1e-12 of the term magnitudes)
    idx = sample_rows(n, 2)
    ha, hb = rows_to_host(env, a, idx), rows_to_host(env, b, idx)
    for got_t, op, ref in ((gp, "gp", o.geometric_product), (outer, "outer", o.outer_product)):
        want, got = ref(ha, hb), rows_to_host(env, got_t, idx)
        bound = o.product_abs_bound(op, np.abs(ha), np.abs(hb))
        assert np.all(np.abs(got - want) <= 1e-12 * np.maximum(np.abs(want), bound)), op
    # (e) linearity with an exactly representable factor: (2a) b == 2 (a b), bit for bit, every element
    a.mul_(2.0)
    cabi.check(lib.lcc_batch_product(alg.h, cabi.LCC_OP_GP, C.byref(va), C.byref(vb), C.byref(vg2), 0, env.sptr))
    gp.mul_(2.0)
    assert torch.equal(gp2, gp)


# ----------------------------------------------------------------------------------------------- cfg4
def test_cfg4_sta_grade_inner_sum_on_2p27_multivectors(env):
    """BASELINE configs[3] at full size: <x>_1 . w summed over 2^27 rows, f64, fused kernel vs the three separate ones."""
    need_gb(env, 95)
    torch, cabi, lib = env.torch, env.cabi, env.lib
    n, nb = 1 << 27, 16
    alg = cabi.Algebra(1, 3, 0)
    gen = torch.Generator(device=env.dev)
    gen.manual_seed(44)
    a = torch.randn((nb, n), device=env.dev, dtype=torch.float64, generator=gen)
    b = torch.randn((nb, n), device=env.dev, dtype=torch.float64, generator=gen)
    g1, prod = torch.empty_like(a), torch.empty_like(a)
    fused, fused2, unfused = (torch.zeros(nb, device=env.dev, dtype=torch.float64) for _ in range(3))
    va, vb, vg, vp = (soa_view(env, t, cabi.LCC_F64) for t in (a, b, g1, prod))
    ptr = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
    cabi.check(lib.lcc_batch_product_sum(alg.h, cabi.LCC_OP_LC, C.byref(va), 1, C.byref(vb), ptr(fused), 0, env.sptr))
    cabi.check(lib.lcc_batch_unary(alg.h, cabi.LCC_UN_GRADE, 1, C.byref(va), C.byref(vg), env.sptr))
    cabi.check(lib.lcc_batch_product(alg.h, cabi.LCC_OP_LC, C.byref(vg), C.byref(vb), C.byref(vp), 0, env.sptr))
    cabi.check(lib.lcc_batch_sum(alg.h, C.byref(vp), ptr(unfused), env.sptr))
    mag = prod.abs().sum(dim=1)  # a-priori magnitude of each column sum
    assert bool(((fused - unfused).abs() <= 1e-12 * mag).all())
    assert bool(((unfused - prod.sum(dim=1)).abs() <= 1e-12 * mag).all())  # the batch sum itself, against torch
    # the scalar column in closed form: <x>_1 . w has scalar part sum over the four vectors e_i of e_i^2 x_i w_i
    signs = np.ctypeslib.as_array(lib.lcc_algebra_cayley_signs(alg.h), shape=(nb, nb))
    ref0 = sum(float(signs[k, k]) * (a[k] * b[k]).sum() for k in (1, 2, 4, 8))
    assert abs(float(fused[0]) - float(ref0)) <= 1e-12 * float(mag[0])
    # doubling an operand doubles every partial sum exactly
    b.mul_(2.0)
    cabi.check(lib.lcc_batch_product_sum(alg.h, cabi.LCC_OP_LC, C.byref(va), 1, C.byref(vb), ptr(fused2), 0, env.sptr))
    assert torch.equal(fused2, fused * 2.0)


# ----------------------------------------------------------------------------------------------- cfg5
def test_cfg5_cl8_fixed_operand_on_2p24_rows(env):
    """BASELINE configs[4] at full size: one fixed left operand (device-resident) times 2^24 multivectors of Cl(8,0,0), f32:
    K8, block-diagonalised in the 16 x 16 matrix representation, stage 2 on the tensor cores."""
    need_gb(env, 45)
    torch, cabi, lib = env.torch, env.cabi, env.lib
    n, nb = 1 << 24, 256
    alg = cabi.Algebra(8, 0, 0)
    gen = torch.Generator(device=env.dev)
    gen.manual_seed(45)
    x = torch.randn((nb, n), device=env.dev, dtype=torch.float32, generator=gen)
    out = torch.empty_like(x)
    vx, vo = soa_view(env, x, cabi.LCC_F32), soa_view(env, out, cabi.LCC_F32)
    signs = np.ctypeslib.as_array(lib.lcc_algebra_cayley_signs(alg.h), shape=(nb, nb))
    # (a) a basis blade as the fixed operand permutes the blades with signs: out[i ^ k] = sign(i, k) x[k], every row
    i = 0b10110101
    m = torch.zeros((nb, 4), device=env.dev, dtype=torch.float32)  # one-row SoA view, ld = 4
    m[i, 0] = 1.0
    vm = cabi.view(m.data_ptr(), 1, 4, cabi.LCC_F32, cabi.LCC_SOA)
    cabi.check(lib.lcc_batch_product(alg.h, cabi.LCC_OP_GP, C.byref(vm), C.byref(vx), C.byref(vo), 0, env.sptr))
    # (K8 transforms each row through two Walsh-Hadamard passes, so the permutation is exact to 1e-5 of the ROW's magnitude --
    # the path's stated tolerance -- not of every single coefficient)
    row_mag = x.abs().amax(dim=0)
    for k in range(nb):
        ref = x[k] if signs[i, k] > 0 else -x[k]
        assert bool(((out[i ^ k] - ref).abs() <= 1e-5 * row_mag).all()), k
    # (b) a dense fixed operand on sampled rows against the oracle (1e-5 of the row magnitude: tensor-core summation order)
    dense = torch.randn((nb, 4), device=env.dev, dtype=torch.float32, generator=gen)
    vd = cabi.view(dense.data_ptr(), 1, 4, cabi.LCC_F32, cabi.LCC_SOA)
    cabi.check(lib.lcc_batch_product(alg.h, cabi.LCC_OP_GP, C.byref(vd), C.byref(vx), C.byref(vo), 0, env.sptr))
    idx = sample_rows(n, 3)[:3000]
    hx, hm = rows_to_host(env, x, idx).astype(np.float64), dense[:, 0].cpu().numpy().astype(np.float64)[None, :]
    o = oracle.OracleAlgebra(8, 0, 0)
    want, got = o.geometric_product(hm, hx), rows_to_host(env, out, idx).astype(np.float64)
    bound = o.product_abs_bound("gp", np.abs(hm), np.abs(hx))
    assert np.all(np.abs(got - want) <= 1e-5 * np.maximum(np.abs(want), bound))
