"""lcc.torch on the GPU: forward parity against the oracle, gradients against plain PyTorch autograd
of the same bilinear map (a dense sign-tensor einsum in float64 on the CPU)."""
import numpy as np
import pytest

import oracle

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

SIGS = [(3, 0, 0), (3, 0, 1), (4, 1, 0), (1, 3, 0), (2, 1, 1)]
MASK = {"gp": lambda i, j: True, "outer": lambda i, j: (i & j) == 0, "inner": lambda i, j: (i & j) == i}


def sign_tensor(o, name):
    """T[i, j, k] = sign(i, j) if k == i^j and the term belongs to the product, else 0."""
    nb = o.num_blades
    T = np.zeros((nb, nb, nb))
    s = o.cayley_signs.reshape(nb, nb)
    for i in range(nb):
        for j in range(nb):
            if MASK[name](i, j):
                T[i, j, i ^ j] = s[i, j]
    return torch.tensor(T)


@pytest.fixture(scope="module")
def lt():
    import lcc.torch as m

    return m


@pytest.mark.parametrize("sig", SIGS)
@pytest.mark.parametrize("dt", [torch.float64, torch.float32])
def test_forward_matches_oracle_bit_for_bit_in_strict_mode(lt, sig, dt):
    o = oracle.OracleAlgebra(*sig)
    alg = lt.TorchAlgebra._from_signature(*sig, device="cuda")
    g = torch.Generator().manual_seed(3)
    a = torch.randn(301, o.num_blades, generator=g, dtype=dt)
    b = torch.randn(301, o.num_blades, generator=g, dtype=dt)
    A, B = lt.TorchMultivector(alg, a.cuda()), lt.TorchMultivector(alg, b.cuda())
    lt.set_strict_fp(True)
    try:
        assert np.array_equal(A.geometric_product(B).to_numpy(), o.geometric_product(a.numpy(), b.numpy()))
        assert np.array_equal(A.outer_product(B).to_numpy(), o.outer_product(a.numpy(), b.numpy()))
        assert np.array_equal(A.inner_product(B).to_numpy(), o.left_contraction(a.numpy(), b.numpy()))
        assert np.array_equal(A.gp(B[0]).to_numpy(), o.geometric_product(a.numpy(), b.numpy()[:1]))
        rotor = B[5]
        assert np.array_equal(A.sandwich(rotor).to_numpy(), o.sandwich(b.numpy()[5].astype(np.float64), a.numpy()))
        assert np.array_equal(A.norm_squared().cpu().numpy(), o.norm_squared(a.numpy()))
    finally:
        lt.set_strict_fp(False)
    assert np.array_equal(A.reverse().to_numpy(), o.reverse(a.numpy()))
    assert np.array_equal(A.grade_involution().to_numpy(), o.grade_involution(a.numpy()))
    assert np.array_equal(A.conjugate().to_numpy(), o.clifford_conjugate(a.numpy()))
    assert np.array_equal(A.grade(1).to_numpy(), o.grade(a.numpy(), 1))
    want, mag = o.sum(a.numpy())
    assert np.all(np.abs(A.sum().cpu().numpy().astype(np.float64) - want) <= (1e-13 if dt == torch.float64 else 1e-6) * mag)


@pytest.mark.parametrize("sig", SIGS)
@pytest.mark.parametrize("name", ["gp", "outer", "inner"])
def test_gradients_match_pytorch_autograd_of_the_same_bilinear_map(lt, sig, name):
    o = oracle.OracleAlgebra(*sig)
    T = sign_tensor(o, name)
    alg = lt.TorchAlgebra._from_signature(*sig, device="cuda")
    g = torch.Generator().manual_seed(11)
    from largecrimsoncanine_b200 import cabi

    # full batch and broadcast right operand; 64 rows take the direct-load VJP kernels, 4500 rows the TMA-tile ones
    # (the default threshold is 4096 rows, x32 for rows under 128 bytes: lowered so every signature takes the tiles)
    cabi.check(cabi.lib().lcc_set_option(b"aos_tma_min_rows", 128))
    try:
        _gradient_cases(lt, o, T, alg, g, name)
    finally:
        cabi.check(cabi.lib().lcc_set_option(b"aos_tma_min_rows", 4096))


def _gradient_cases(lt, o, T, alg, g, name):
    for n, nb_rows in ((64, 64), (64, 1), (4500, 4500), (4500, 1)):
        a = torch.randn(n, o.num_blades, generator=g, dtype=torch.float64)
        b = torch.randn(nb_rows, o.num_blades, generator=g, dtype=torch.float64)
        w = torch.randn(n, o.num_blades, generator=g, dtype=torch.float64)
        ar, br = a.clone().requires_grad_(True), b.clone().requires_grad_(True)
        ref = torch.einsum("ni,nj,ijk->nk", ar, br.expand(n, -1), T)
        (ref * w).sum().backward()
        ac, bc = a.cuda().requires_grad_(True), b.cuda().requires_grad_(True)
        A, B = lt.TorchMultivector(alg, ac), lt.TorchMultivector(alg, bc)
        out = {"gp": A.geometric_product, "outer": A.outer_product, "inner": A.inner_product}[name](B)
        (out.coeffs * w.cuda()).sum().backward()
        assert torch.allclose(out.coeffs.detach().cpu(), ref.detach(), rtol=1e-12, atol=1e-12)
        assert torch.allclose(ac.grad.cpu(), ar.grad, rtol=1e-11, atol=1e-11)
        assert torch.allclose(bc.grad.cpu(), br.grad, rtol=1e-11, atol=1e-11)


def test_sandwich_with_gradients_and_error_behaviour(lt):
    alg = lt.TorchAlgebra.pga(3, device="cuda")
    o = oracle.OracleAlgebra(3, 0, 1)
    g = torch.Generator().manual_seed(5)
    x = torch.randn(33, 16, generator=g, dtype=torch.float64)
    r = torch.randn(1, 16, generator=g, dtype=torch.float64)
    xc = x.cuda().requires_grad_(True)
    out = lt.TorchMultivector(alg, xc).sandwich(lt.TorchMultivector(alg, r.cuda()))
    assert torch.allclose(out.coeffs.detach().cpu(), torch.tensor(o.sandwich(r.numpy()[0], x.numpy())), rtol=1e-12, atol=1e-12)
    out.coeffs.sum().backward()
    # sandwich is linear in x: grad = M^T 1, check by finite structure: apply to basis
    basis = torch.eye(16, dtype=torch.float64)
    M = torch.tensor(o.sandwich(r.numpy()[0], basis.numpy()))  # rows: image of each basis blade
    assert torch.allclose(xc.grad.cpu(), M.sum(dim=1).expand(33, -1), rtol=1e-11, atol=1e-11)
    with pytest.raises(ValueError, match="must have 16 columns"):
        lt.TorchMultivector(alg, torch.zeros(4, 8))
    cpu_alg = lt.TorchAlgebra.pga(3)
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        lt.TorchMultivector(cpu_alg, torch.zeros(4, 16)).geometric_product(lt.TorchMultivector(cpu_alg, torch.zeros(4, 16)))
    with pytest.raises(ValueError, match="batch sizes must match"):
        lt.TorchMultivector(alg, torch.zeros(4, 16)).gp(lt.TorchMultivector(alg, torch.zeros(3, 16)))
    # from_rust_algebra works with the extension's Algebra (the reference's version is broken)
    import largecrimsoncanine as L

    ta = lt.TorchAlgebra.from_rust_algebra(L.Algebra.cga(3), device="cuda")
    assert ta.signature == (4, 1, 0) and ta.num_blades == 32 and ta.cayley_signs.shape == (32, 32)
    vecs = lt.TorchMultivector.from_vectors(ta, torch.randn(7, 5, dtype=torch.float64))
    assert vecs.to_vector_coords().shape == (7, 5) and len(vecs) == 7


@pytest.mark.parametrize("sig,dt", [((3, 0, 1), torch.float64), ((3, 0, 1), torch.float32), ((4, 1, 0), torch.float64), ((3, 0, 0), torch.float64)])
def test_fixed_rotor_sandwich_gradient_matches_the_composed_products(lt, sig, dt):
    """x.requires_grad with one rotor that needs no gradient: fused forward (no host read-back on the first call with a
    rotor tensor, the constant-bank kernel once its host copy has arrived) and two broadcast VJP passes backward, against
    the reference composition gp(gp(R, x), ~R) differentiated by autograd -- for a random cotangent."""
    alg = lt.TorchAlgebra(sig, device="cuda")
    nb = alg.num_blades
    g = torch.Generator().manual_seed(11)
    x = torch.randn(257, nb, generator=g, dtype=dt).cuda()
    r = torch.randn(1, nb, generator=g, dtype=dt).cuda()
    cot = torch.randn(257, nb, generator=g, dtype=dt).cuda()
    tol = 1e-11 if dt == torch.float64 else 2e-4
    R = lt.TorchMultivector(alg, r)
    for attempt in range(3):  # first call: device-rotor kernel; later calls: host copy + fixed-versor kernel
        xa = x.clone().requires_grad_(True)
        out = lt.TorchMultivector(alg, xa).sandwich(R)
        (out.coeffs * cot).sum().backward()
        xb = x.clone().requires_grad_(True)
        rb = r.clone().requires_grad_(True)
        Rb = lt.TorchMultivector(alg, rb)  # the reference's composition, differentiated by autograd
        ref = Rb.geometric_product(lt.TorchMultivector(alg, xb)).geometric_product(Rb.reverse())
        (ref.coeffs * cot).sum().backward()
        scale = float(ref.coeffs.abs().max())
        assert float((out.coeffs - ref.coeffs).abs().max()) <= tol * scale
        assert float((xa.grad - xb.grad).abs().max()) <= tol * float(xb.grad.abs().max())
        assert rb.grad is not None and rb.grad.shape == (1, nb)
        torch.cuda.synchronize()


@pytest.mark.parametrize("sig,dt", [((3, 0, 1), torch.float64), ((3, 0, 1), torch.float32), ((3, 0, 0), torch.float64), ((1, 3, 0), torch.float64),
                                    ((2, 0, 1), torch.float32), ((2, 1, 1), torch.float64), ((4, 1, 0), torch.float64)])
@pytest.mark.parametrize("per_row", [True, False])
def test_fused_sandwich_cotangents_match_autograd_of_the_composed_products(lt, sig, dt, per_row):
    """Gradients to the rotor(s): forward = the fused per-row sandwich, backward = lcc_batch_sandwich_rows_vjp (ONE kernel up
    to 16 blades -- specialised and run-time-metric signatures; the composed five passes behind the same entry point for
    Cl(2,1,1) in f64 and for 32 blades), against torch autograd through gp(gp(R, x), ~R)
    (/root/reference/lcc/torch/multivector.py:988-1019) for a random cotangent, at a size with a ragged last block."""
    alg = lt.TorchAlgebra(sig, device="cuda")
    nb = alg.num_blades
    g = torch.Generator().manual_seed(17 + nb)
    n = 4099
    x = torch.randn(n, nb, generator=g, dtype=dt).cuda()
    r = torch.randn(n if per_row else 1, nb, generator=g, dtype=dt).cuda()
    cot = torch.randn(n, nb, generator=g, dtype=dt).cuda()
    tol = 1e-11 if dt == torch.float64 else 2e-4
    xa, ra = x.clone().requires_grad_(True), r.clone().requires_grad_(True)
    out = lt.TorchMultivector(alg, xa).sandwich(lt.TorchMultivector(alg, ra))
    (out.coeffs * cot).sum().backward()
    xb, rb = x.clone().requires_grad_(True), r.clone().requires_grad_(True)
    Rb = lt.TorchMultivector(alg, rb)
    ref = Rb.geometric_product(lt.TorchMultivector(alg, xb)).geometric_product(Rb.reverse())
    (ref.coeffs * cot).sum().backward()
    assert float((out.coeffs - ref.coeffs).abs().max()) <= tol * float(ref.coeffs.abs().max())
    assert xa.grad.shape == x.shape and ra.grad.shape == r.shape
    assert float((xa.grad - xb.grad).abs().max()) <= tol * float(xb.grad.abs().max())
    assert float((ra.grad - rb.grad).abs().max()) <= (tol if per_row else 20 * tol) * float(rb.grad.abs().max())
    # only the rotor needs a gradient
    rc = r.clone().requires_grad_(True)
    out = lt.TorchMultivector(alg, x).sandwich(lt.TorchMultivector(alg, rc))
    (out.coeffs * cot).sum().backward()
    assert float((rc.grad - rb.grad).abs().max()) <= (tol if per_row else 20 * tol) * float(rb.grad.abs().max())


def test_sandwich_with_a_batch_of_rotors(lt):
    """One rotor per row, no gradients: the C-ABI per-row sandwich on the borrowed tensors; with gradients: the same forward
    kernel inside an autograd function.  Both agree with the oracle."""
    alg = lt.TorchAlgebra.pga(3, device="cuda")
    o = oracle.OracleAlgebra(3, 0, 1)
    g = torch.Generator().manual_seed(13)
    x = torch.randn(300, 16, generator=g, dtype=torch.float64)
    r = torch.randn(300, 16, generator=g, dtype=torch.float64)
    want = torch.tensor(o.geometric_product(o.geometric_product(r.numpy(), x.numpy()), o.reverse(r.numpy())))
    with torch.no_grad():
        out = lt.TorchMultivector(alg, x.cuda()).sandwich(lt.TorchMultivector(alg, r.cuda()))
    assert torch.allclose(out.coeffs.cpu(), want, rtol=1e-12, atol=1e-12)
    xg = x.cuda().requires_grad_(True)
    out2 = lt.TorchMultivector(alg, xg).sandwich(lt.TorchMultivector(alg, r.cuda()))
    assert torch.allclose(out2.coeffs.detach().cpu(), want, rtol=1e-12, atol=1e-12)
    out2.coeffs.sum().backward()
    assert xg.grad is not None and torch.isfinite(xg.grad).all()
