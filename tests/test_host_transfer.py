"""from_numpy / to_numpy / sandwich_numpy: the chunked host<->device pipelines of host_transfer.cu
(reference semantics: /root/reference/src/batch.rs:79-94 copies the array in, :233-235 copies it out).

GPU tests force tiny chunks so a few thousand rows already run as several chunks with a ragged tail, on pageable
and on pinned arrays, and compare bit for bit (a copy has no tolerance).  CPU tests cover argument validation and
the host-side copy pool's refusal to work without a device."""
import ctypes as C

import numpy as np
import pytest

import oracle
from largecrimsoncanine_b200 import cabi


@pytest.fixture(scope="module")
def L():
    import largecrimsoncanine as ext

    return ext


@pytest.fixture()
def small_chunks(L):
    old = L.get_option("host_chunk_bytes")
    L.set_option("host_chunk_bytes", 1 << 16)  # 64 KiB: 1024 PGA f32 rows / 256 CGA f64 rows per chunk
    yield
    L.set_option("host_chunk_bytes", old)


def test_upload_validates_before_touching_the_device():
    lib, alg = cabi.lib(), cabi.Algebra(3, 0, 1)
    host = np.zeros((8, 16))
    v = cabi.view(0x1000, 9, 64, cabi.LCC_F64, cabi.LCC_SOA)
    assert lib.lcc_batch_upload(alg.h, host.ctypes.data, 8, C.byref(v), None) == cabi.LCC_ERR_VALUE
    assert b"8 rows" in lib.lcc_last_error()
    v = cabi.view(0x1000, 8, 64, cabi.LCC_F64, cabi.LCC_SOA)
    assert lib.lcc_batch_upload(alg.h, None, 8, C.byref(v), None) == cabi.LCC_ERR_VALUE
    assert lib.lcc_batch_download(alg.h, C.byref(v), None, None) == cabi.LCC_ERR_VALUE
    empty = cabi.view(None, 0, 64, cabi.LCC_F64, cabi.LCC_SOA)
    assert lib.lcc_batch_upload(alg.h, host.ctypes.data, 0, C.byref(empty), None) == cabi.LCC_OK
    assert lib.lcc_host_cache_free(C.c_void_p(0x1234)) == cabi.LCC_ERR_VALUE  # not one of ours
    assert lib.lcc_host_is_pinned(host.ctypes.data, host.nbytes) == 0
    # lcc_batch_copy_rows: negative bounds are an IndexError at the ABI boundary (the reference's usize cannot be negative)
    src = cabi.view(0x1000, 8, 64, cabi.LCC_F64, cabi.LCC_SOA)
    assert lib.lcc_batch_copy_rows(alg.h, C.byref(src), -5, 3, C.byref(v), 0, None) == cabi.LCC_ERR_INDEX
    assert lib.lcc_batch_copy_rows(alg.h, C.byref(src), 0, -1, C.byref(v), 0, None) == cabi.LCC_ERR_INDEX


@pytest.mark.gpu
@pytest.mark.parametrize("sig,dt", [((3, 0, 1), np.float32), ((4, 1, 0), np.float64), ((3, 0, 0), np.float64), ((6, 0, 0), np.float32)])
def test_round_trip_in_several_ragged_chunks(L, small_chunks, sig, dt):
    alg = L.Algebra(*sig)
    nb = alg.num_blades
    rows_per_chunk = (1 << 16) // (nb * np.dtype(dt).itemsize)
    rows_per_chunk -= rows_per_chunk % 64
    rng = np.random.default_rng(5)
    for n in (1, 63, rows_per_chunk, 3 * rows_per_chunk + 37, 7 * rows_per_chunk + 1):
        x = rng.standard_normal((n, nb)).astype(dt)
        pinned = L.pinned_empty((n, nb), dt)
        pinned[...] = x
        assert L.is_pinned(pinned) and not L.is_pinned(x)
        for src in (x, pinned):
            b = L.MultivectorBatch.from_numpy(alg, src)
            assert np.array_equal(b.to_numpy(), x)  # pinned-cache or pageable result, depending on its size
            out = np.empty_like(x)
            assert b.to_numpy(out=out) is out and np.array_equal(out, x)
            pout = L.pinned_empty((n, nb), dt)
            b.to_numpy(out=pout)
            assert np.array_equal(pout, x)
    with pytest.raises(ValueError, match="out must be"):
        L.MultivectorBatch.from_numpy(alg, x).to_numpy(out=np.empty((3, nb), dt))


@pytest.mark.gpu
def test_upload_download_through_the_c_abi_any_layout(small_chunks):
    """Raw C ABI: AoS and SoA device views, host array read once, canary rows around the destination untouched."""
    lib, alg = cabi.lib(), cabi.Algebra(1, 3, 0)
    nb, n = 16, 5 * 512 + 19
    rng = np.random.default_rng(9)
    x = rng.standard_normal((n, nb))
    for layout in (cabi.LCC_SOA, cabi.LCC_AOS):
        ld = 3008 if layout == cabi.LCC_SOA else nb + 4  # padded pitch in both layouts
        elems = nb * ld if layout == cabi.LCC_SOA else n * ld
        dev = C.c_void_p()
        cabi.check(lib.lcc_device_alloc(C.byref(dev), elems * 8, None))
        canary = np.full(elems, -7.25)
        cabi.check(lib.lcc_memcpy_h2d(dev, canary.ctypes.data, canary.nbytes, None))
        v = cabi.view(dev.value, n, ld, cabi.LCC_F64, layout)
        cabi.check(lib.lcc_batch_upload(alg.h, x.ctypes.data, n, C.byref(v), None))
        raw = np.empty(elems)
        cabi.check(lib.lcc_memcpy_d2h(raw.ctypes.data, dev, raw.nbytes, None))
        cabi.check(lib.lcc_stream_sync(None))
        if layout == cabi.LCC_SOA:
            raw = raw.reshape(nb, ld)
            assert np.array_equal(raw[:, :n].T, x) and np.all(raw[:, n:] == -7.25)
        else:
            raw = raw.reshape(n, ld)
            assert np.array_equal(raw[:, :nb], x) and np.all(raw[:, nb:] == -7.25)
        back = np.zeros_like(x)
        cabi.check(lib.lcc_batch_download(alg.h, C.byref(v), back.ctypes.data, None))
        assert np.array_equal(back, x)
        cabi.check(lib.lcc_device_free(dev, None))


@pytest.mark.gpu
@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_sandwich_numpy_equals_the_three_call_sequence(L, small_chunks, dt):
    """One fused host call == from_numpy -> sandwich -> to_numpy, bit for bit, pageable and pinned, and == the oracle."""
    alg, o = L.Algebra.pga(3), oracle.OracleAlgebra(3, 0, 1)
    motor = L.Multivector.pga_motor_from_axis_angle(alg, [1.0, 2.0, 3.0], [0.5, -0.25, 1.0], 0.7)
    rng = np.random.default_rng(11)
    n = 3 * 1024 + 333
    x = rng.standard_normal((n, 16)).astype(dt)
    L.set_strict_fp(True)
    try:
        three = L.MultivectorBatch.from_numpy(alg, x).sandwich(motor).to_numpy()
        fused = L.MultivectorBatch.sandwich_numpy(alg, motor, x)
        px = L.pinned_empty((n, 16), dt)
        px[...] = x
        pout = L.pinned_empty((n, 16), dt)
        assert L.MultivectorBatch.sandwich_numpy(alg, motor, px, out=pout) is pout
    finally:
        L.set_strict_fp(False)
    want = o.sandwich(np.array(motor.coefficients()), x)
    assert np.array_equal(three, want) and np.array_equal(fused, want) and np.array_equal(pout, want)
    with pytest.raises(ValueError, match="must have 16 columns"):
        L.MultivectorBatch.sandwich_numpy(alg, motor, np.zeros((4, 8), dt))
    with pytest.raises(ValueError, match="algebra mismatch"):
        L.MultivectorBatch.sandwich_numpy(L.Algebra.cga(3), motor, np.zeros((4, 32), dt))


@pytest.mark.gpu
def test_large_results_come_from_the_pinned_cache_and_return_to_it(L):
    alg = L.Algebra.pga(3)
    L.host_cache_trim()
    x = np.random.default_rng(2).standard_normal((1 << 16, 16)).astype(np.float32)  # 4 MiB >= pinned_output_min_bytes
    b = L.MultivectorBatch.from_numpy(alg, x)
    out = b.to_numpy()
    assert L.is_pinned(out) and np.array_equal(out, x)
    live, idle = L.host_cache_stats()
    assert live >= out.nbytes
    del out
    live2, idle2 = L.host_cache_stats()
    assert live2 == live - (4 << 20) and idle2 >= (4 << 20)  # back in the cache ...
    again = b.to_numpy()
    assert L.is_pinned(again) and L.host_cache_stats()[1] == idle2 - (4 << 20)  # ... and handed out again
    small = L.MultivectorBatch.from_numpy(alg, x[:64]).to_numpy()
    assert not L.is_pinned(small)  # below the threshold: ordinary NumPy memory
    del again
    L.host_cache_trim()
    assert L.host_cache_stats()[1] == 0
