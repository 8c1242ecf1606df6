// product_impl.cuh -- K1 (binary product family) and K2 (fixed-versor sandwich) for ONE signature
// and ONE precision.  The including translation unit defines
//
//   LCC_TERMS_FILE      the generated term list (csrc/generated/cayley_*.inc)
//   LCC_NB              number of blades (2, 4, 8, 16 or 32)
//   LCC_RUNTIME_METRIC  0: signs are the signature's own, folded into FMA operand negation
//                       1: the list holds the metric-free reorder signs; the metric factor
//                          mu[i & j] in {+1,-1,0} arrives as a kernel argument (any Cl(p,q,r))
//   LCC_REAL            float or double
//   LCC_TU              token used to name the two exported launchers
//   LCC_VJP0_FILE / LCC_VJP1_FILE   term lists of the vector-Jacobian products (gen_cayley.py, emit_vjp_terms);
//                       the product kernels take a LIST template argument: 0 forward, 1 / 2 cotangent of the
//                       left / right operand (autograd of lcc.torch)
//
// Reference semantics (file:line in /root/reference): out[i^j] += sign[i][j]*a[i]*b[j], i outer,
// j inner, src/batch.rs:384-404 (geometric), :544-568 (outer, (i&j)==0), :615-649 (left
// contraction, (i&j)==i); sandwich = two such stages with one rotor, :452-497.  In gather form
// every output blade k receives its terms in ascending i, which is the order of the generated
// list, so STRICT kernels reproduce the reference bit for bit and FAST kernels differ only by
// the single rounding an FMA saves per term.
//
// Roofline: these kernels are HBM-bound streaming kernels.  Algorithmic bytes per row are
// 3*NB*sizeof(T) for a product (2 reads + 1 write), 4*NB*sizeof(T) for the fused gp+outer form
// and 2*NB*sizeof(T) for the sandwich; each byte is touched exactly once by a full-width
// coalesced access (common.cuh).
#pragma once
#include <cstdlib>

#include "common.cuh"
#include "launch_args.h"
#include "tma_tile.cuh"

#ifndef LCC_TERMS_FILE
#error "define LCC_TERMS_FILE, LCC_NB, LCC_RUNTIME_METRIC, LCC_REAL and LCC_TU before including product_impl.cuh"
#endif

namespace lcc {
namespace LCC_TU {

constexpr int NB = LCC_NB;
constexpr bool kRuntimeMetric = LCC_RUNTIME_METRIC != 0;
using real = LCC_REAL;
constexpr int kRowsSoA = RowsPerThread<real, NB, kRuntimeMetric ? 256 : 512>::value;  // rows per thread, SoA

struct Coeffs { real v[NB]; int stream_hint; };  // a uniform multivector (rotor, metric factors) in the constant bank (+ the TMA L2 hint switch)

__host__ __device__ constexpr bool even_grade(int blade) {
    int c = 0;
    for (int b = blade; b; b >>= 1) c += b & 1;
    return (c & 1) == 0;
}
// flags of a generated term: 1 = (i & j) == 0 (outer), 2 = (i & j) == i (left contraction),
// 4 = (i & j) == j (right contraction); k = output blade (the scalar product keeps k == 0 only)
__host__ __device__ constexpr bool term_in_op(int op, int flags, int k) {
    return op == kOpGP || op == kOpGPOuter || op == kOpCommutator || op == kOpAnticommutator ||
           (op == kOpOuter && (flags & 1)) || (op == kOpLC && (flags & 2)) || (op == kOpRC && (flags & 4)) ||
           (op == kOpScalar && k == 0);
}
__host__ __device__ constexpr bool two_sided(int op) { return op == kOpCommutator || op == kOpAnticommutator; }

// (a*b -/+ b*a) * 0.5 with every step rounded, as Multivector::commutator / anticommutator compose them
// (/root/reference/src/multivector.rs:1732-1768: geometric_product twice, __sub__ / __add__, scale(0.5)).
template <int OP, class T, int V> static __device__ __forceinline__ void finish_two_sided(T (&acc)[V], const T (&acc2)[V]) {
    if constexpr (two_sided(OP)) {
#pragma unroll
        for (int l = 0; l < V; ++l) acc[l] = mul_rn(add_rn(acc[l], OP == kOpCommutator ? -acc2[l] : acc2[l]), T(0.5));
    }
}

// Output sink: SoA stores V rows of one blade per call; AoS collects 16 bytes of consecutive
// blades of the thread's row before storing.
template <int LAYOUT, int V> struct Sink {
    static constexpr int W = 16 / (int)sizeof(real);
    real *base; int64_t ld; int64_t r0; bool vec_ok;
    real buf[W];
    template <int K> __device__ __forceinline__ void put(const real (&acc)[V]) {
        if constexpr (LAYOUT == kSoA) {
            stg_stream(base + (int64_t)K * ld + r0, acc);
        } else {
            if constexpr (NB < W) {
                base[r0 * ld + K] = acc[0];
            } else {
                buf[K % W] = acc[0];
                if constexpr ((K % W) == W - 1) {
                    real *p = base + r0 * ld + (K - (W - 1));
                    if (vec_ok) {
                        stg_stream(p, buf);
                    } else {
#pragma unroll
                        for (int l = 0; l < W; ++l) p[l] = buf[l];
                    }
                }
            }
        }
    }
};

// every blade of a one-row tile through a sink, in ascending order (the AoS sink collects 16-byte groups)
template <int K, int LAYOUT> static __device__ __forceinline__ void put_all(Sink<LAYOUT, 1> &sink, const Tile<real, NB, 1> &t) {
    if constexpr (K < NB) {
        sink.template put<K>(t.v[K]);
        put_all<K + 1>(sink, t);
    }
}

// ------------------------------------------------------------------------------------------ K1
template <int LAYOUT, int OP, bool STRICT, int LIST = 0>
__global__ void __launch_bounds__(kThreads)
product_kernel(const real *__restrict__ a, int64_t lda, int a_bcast, const real *__restrict__ b, int64_t ldb,
               int b_bcast, real *__restrict__ out, int64_t ldo, real *__restrict__ out2, int64_t ldo2,
               int64_t n, int vec_ok, const Coeffs mu) {
    constexpr int V = LAYOUT == kSoA ? kRowsSoA : 1;
    const int64_t r0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * V;
    if (r0 >= n) return;

    Tile<real, NB, V> A, B;
    if constexpr (LAYOUT == kSoA) {
        load_soa(A, a, lda, r0, a_bcast != 0);
        load_soa(B, b, ldb, r0, b_bcast != 0);
    } else {
        load_aos(A, a, lda, r0, a_bcast != 0, vec_ok != 0);
        load_aos(B, b, ldb, r0, b_bcast != 0, vec_ok != 0);
    }
    Sink<LAYOUT, V> sink{out, ldo, r0, vec_ok != 0, {}};
    Sink<LAYOUT, V> sink2{out2, ldo2, r0, vec_ok != 0, {}};
    (void)sink2;
    (void)mu;

#define LCC_OUT_BEGIN(k) { real acc[V]; real acc2[V]; _Pragma("unroll") for (int l = 0; l < V; ++l) { acc[l] = real(0); acc2[l] = real(0); }
#define LCC_TERM6(k, i, j, neg, flags, mi)                                                                   \
    if constexpr (term_in_op(OP, flags, k)) {                                                             \
        _Pragma("unroll") for (int l = 0; l < V; ++l) {                                                \
            real av = (neg) ? -A.v[i][l] : A.v[i][l];                                                  \
            if constexpr (kRuntimeMetric) av = mul_rn(av, mu.v[mi]);                            \
            acc[l] = Arith<STRICT>::madd(av, B.v[j][l], acc[l]);                                       \
            if constexpr (OP == kOpGPOuter && ((flags) & 1)) acc2[l] = Arith<STRICT>::madd(av, B.v[j][l], acc2[l]); \
            if constexpr (two_sided(OP)) {                                                             \
                real bv = (neg) ? -B.v[i][l] : B.v[i][l];                                              \
                if constexpr (kRuntimeMetric) bv = mul_rn(bv, mu.v[mi]);                        \
                acc2[l] = Arith<STRICT>::madd(bv, A.v[j][l], acc2[l]);                                 \
            }                                                                                          \
        }                                                                                              \
    }
#define LCC_OUT_END(k) finish_two_sided<OP>(acc, acc2); sink.template put<k>(acc); if constexpr (OP == kOpGPOuter) sink2.template put<k>(acc2); }
#define LCC_TERM(k, i, j, neg, flags) LCC_TERM6(k, i, j, neg, flags, (i) & (j))
#define LCC_VTERM(k, i, j, neg, flags, mi) LCC_TERM6(k, i, j, neg, flags, mi)
    if constexpr (LIST == 0) {
#include LCC_TERMS_FILE
    } else if constexpr (LIST == 1) {
#include LCC_VJP0_FILE
    } else {
#include LCC_VJP1_FILE
    }
#undef LCC_VTERM
#undef LCC_TERM6
#undef LCC_OUT_BEGIN
#undef LCC_TERM
#undef LCC_OUT_END
}

// ------------------------------------------------------------------------------------------ K1 + K5
// sums[k] = sum over rows of (<a>_g (op) b)[k] without ever writing the product: the grade projection is a
// load-time mask (masked blade rows of `a` are not read at all), the product terms run in the reference order,
// and every thread adds its rows' results into its own f64 column of a shared-memory tile; one partial per
// blade and block goes to global memory and sum_final (misc_kernels.cu) finishes the fixed tree.
// Traffic for STA <x>_1 . w: 4 + 16 blade rows read per multivector (160 B in f64) instead of the 672 B of
// the three separate kernels.
__host__ __device__ constexpr int blade_grade_of(int blade) {
    int c = 0;
    for (int b = blade; b; b >>= 1) c += b & 1;
    return c;
}

template <int LAYOUT, int OP, bool STRICT>
__global__ void __launch_bounds__(kThreads)
product_sum_kernel(const real *__restrict__ a, int64_t lda, int a_bcast, int a_grade, const real *__restrict__ b, int64_t ldb,
                   int b_bcast, int64_t n, int vec_ok, const Coeffs mu, double *__restrict__ partial) {
    constexpr int V = LAYOUT == kSoA ? kRowsSoA : 1;
    extern __shared__ double ssum[];  // [NB][blockDim.x]
    const int bd = blockDim.x, tid = threadIdx.x;
#pragma unroll
    for (int k = 0; k < NB; ++k) ssum[k * bd + tid] = 0.0;
    (void)mu;
    for (int64_t r0 = ((int64_t)blockIdx.x * bd + tid) * V; r0 < n; r0 += (int64_t)gridDim.x * bd * V) {
        Tile<real, NB, V> A, B;
        if constexpr (LAYOUT == kSoA) {
#pragma unroll
            for (int k = 0; k < NB; ++k) {
                if (a_grade >= 0 && blade_grade_of(k) != a_grade) {  // warp-uniform: this blade row is never read
#pragma unroll
                    for (int l = 0; l < V; ++l) A.v[k][l] = real(0);
                } else if (a_bcast) {
                    const real sv = __ldg(a + (int64_t)k * lda);
#pragma unroll
                    for (int l = 0; l < V; ++l) A.v[k][l] = sv;
                } else {
                    ldg_stream(a + (int64_t)k * lda + r0, A.v[k]);
                }
            }
            load_soa(B, b, ldb, r0, b_bcast != 0);
        } else {
            load_aos(A, a, lda, r0, a_bcast != 0, vec_ok != 0);
            load_aos(B, b, ldb, r0, b_bcast != 0, vec_ok != 0);
            if (a_grade >= 0) {
#pragma unroll
                for (int k = 0; k < NB; ++k)
                    if (blade_grade_of(k) != a_grade) A.v[k][0] = real(0);
            }
        }
#define LCC_OUT_BEGIN(k) { real acc[V]; _Pragma("unroll") for (int l = 0; l < V; ++l) acc[l] = real(0);
#define LCC_TERM(k, i, j, neg, flags)                                                                   \
    if constexpr (term_in_op(OP, flags, k)) {                                                             \
        _Pragma("unroll") for (int l = 0; l < V; ++l) {                                                \
            real av = (neg) ? -A.v[i][l] : A.v[i][l];                                                  \
            if constexpr (kRuntimeMetric) av = mul_rn(av, mu.v[(i) & (j)]);                            \
            acc[l] = Arith<STRICT>::madd(av, B.v[j][l], acc[l]);                                       \
        }                                                                                              \
    }
#define LCC_OUT_END(k) { double rowsum = 0.0; _Pragma("unroll") for (int l = 0; l < V; ++l) if (r0 + l < n) rowsum += (double)acc[l]; \
                         ssum[(k) * bd + tid] += rowsum; } }
#include LCC_TERMS_FILE
#undef LCC_OUT_BEGIN
#undef LCC_TERM
#undef LCC_OUT_END
    }
    __syncthreads();
    if (tid < NB) {  // fixed order over the block's threads
        double total = 0.0;
        for (int t = 0; t < bd; ++t) total += ssum[tid * bd + t];
        partial[(int64_t)tid * gridDim.x + blockIdx.x] = total;
    }
}

// ------------------------------------------------------------------------------------------ K2
// out = R * x * ~R, two stages, the intermediate R*x rounded to storage precision exactly as the
// reference's `rx` vector (batch.rs:462-478).  EVEN prunes the odd-grade blades of R at compile
// time (rotors and motors are even); pruned terms would only have added +-0.
template <int LAYOUT, bool STRICT, bool EVEN>
__global__ void __launch_bounds__(kThreads)
sandwich_kernel(const real *__restrict__ x, int64_t ldx, real *__restrict__ out, int64_t ldo, int64_t n,
                int vec_ok, const Coeffs R, const Coeffs Rrev, const Coeffs mu) {
    constexpr int V = LAYOUT == kSoA ? kRowsSoA : 1;
    const int64_t r0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * V;
    if (r0 >= n) return;

    Tile<real, NB, V> X, RX;
    if constexpr (LAYOUT == kSoA) load_soa(X, x, ldx, r0, false);
    else load_aos(X, x, ldx, r0, false, vec_ok != 0);
    (void)mu;

    // stage 1: rx[k] = sum_i sign(i, i^k) * R[i] * x[i^k]
#define LCC_OUT_BEGIN(k) { real acc[V]; _Pragma("unroll") for (int l = 0; l < V; ++l) acc[l] = real(0);
#define LCC_TERM(k, i, j, neg, flags)                                                                   \
    if constexpr (!EVEN || even_grade(i)) {                                                            \
        real rv = (neg) ? -R.v[i] : R.v[i];                                                            \
        if constexpr (kRuntimeMetric) rv = mul_rn(rv, mu.v[(i) & (j)]);                                \
        _Pragma("unroll") for (int l = 0; l < V; ++l) acc[l] = Arith<STRICT>::madd(rv, X.v[j][l], acc[l]); \
    }
#define LCC_OUT_END(k) _Pragma("unroll") for (int l = 0; l < V; ++l) RX.v[k][l] = acc[l]; }
#include LCC_TERMS_FILE
#undef LCC_TERM
#undef LCC_OUT_END

    // stage 2: out[k] = sum_i sign(i, i^k) * rx[i] * ~R[i^k]
    Sink<LAYOUT, V> sink{out, ldo, r0, vec_ok != 0, {}};
#define LCC_TERM(k, i, j, neg, flags)                                                                   \
    if constexpr (!EVEN || even_grade(j)) {                                                            \
        real rv = (neg) ? -Rrev.v[j] : Rrev.v[j];                                                      \
        if constexpr (kRuntimeMetric) rv = mul_rn(rv, mu.v[(i) & (j)]);                                \
        _Pragma("unroll") for (int l = 0; l < V; ++l) acc[l] = Arith<STRICT>::madd(RX.v[i][l], rv, acc[l]); \
    }
#define LCC_OUT_END(k) sink.template put<k>(acc); }
#include LCC_TERMS_FILE
#undef LCC_OUT_BEGIN
#undef LCC_TERM
#undef LCC_OUT_END
}

// ------------------------------------------------------------------------------------------ K1 / K2 on TMA-staged AoS tiles
// Same arithmetic as product_kernel<kAoS> / sandwich_kernel<kAoS>, different data movement: the block's
// kRows consecutive rows arrive in shared memory through cp.async.bulk.tensor (one elected thread, mbarrier
// completion), each thread reads its own row from the swizzled tile, writes its result row over it, and the
// tile leaves through a TMA store.  A thread only ever touches the shared-memory bytes of its own row, so the
// in-place overwrite needs no barrier; the only block-wide sync is the one that precedes the store.
constexpr int kRowBytes = NB * (int)sizeof(real);
constexpr bool kHasTmaPath = kRowBytes >= 32;
using TmaTile = AosTmaTile<(kRowBytes >= 32 ? kRowBytes : 32)>;
constexpr int kChunkElems = 16 / (int)sizeof(real);

struct alignas(16) Chunk { real v[kChunkElems]; };

static __device__ __forceinline__ void load_tile_row(Tile<real, NB, 1> &t, const uint8_t *tile, int row) {
#pragma unroll
    for (int j = 0; j < NB / kChunkElems; ++j) {
        const Chunk c = *reinterpret_cast<const Chunk *>(tile + TmaTile::chunk_offset(row, j));
#pragma unroll
        for (int l = 0; l < kChunkElems; ++l) t.v[j * kChunkElems + l][0] = c.v[l];
    }
}
static __device__ __forceinline__ void load_bcast_row(Tile<real, NB, 1> &t, const real *__restrict__ row) {
#pragma unroll
    for (int k = 0; k < NB; ++k) t.v[k][0] = __ldg(row + k);
}

struct TileSink {  // collects 16 bytes of consecutive blades, then writes them to the thread's row of the tile
    uint8_t *tile; int row;
    Chunk buf;
    template <int K> __device__ __forceinline__ void put(const real (&acc)[1]) {
        buf.v[K % kChunkElems] = acc[0];
        if constexpr ((K % kChunkElems) == kChunkElems - 1)
            *reinterpret_cast<Chunk *>(tile + TmaTile::chunk_offset(row, K / kChunkElems)) = buf;
    }
};

static __device__ __forceinline__ uint8_t *tma_smem_base(uint8_t *raw) {
    return (uint8_t *)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);  // swizzle patterns are functions of the address bits
}
static __device__ __forceinline__ void tma_load_tile(uint8_t *dst, const CUtensorMap *map, uint64_t *bar, int row0, bool hint) {
#pragma unroll
    for (int sub = 0; sub < TmaTile::kSub; ++sub)
        tma::load_2d(dst + sub * TmaTile::kRows * TmaTile::kSpan, map, bar, sub * (TmaTile::kSpan / (int)sizeof(real)), row0, hint);
}
static __device__ __forceinline__ void tma_store_tile(const CUtensorMap *map, const uint8_t *src, int row0, bool hint) {
#pragma unroll
    for (int sub = 0; sub < TmaTile::kSub; ++sub)
        tma::store_2d(map, src + sub * TmaTile::kRows * TmaTile::kSpan, sub * (TmaTile::kSpan / (int)sizeof(real)), row0, hint);
}

template <int OP, bool STRICT, int LIST = 0>
__global__ void __launch_bounds__(TmaTile::kRows)
product_kernel_aos_tma(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                       const __grid_constant__ CUtensorMap map_out, const __grid_constant__ CUtensorMap map_out2,
                       const real *__restrict__ a_row, const real *__restrict__ b_row, const Coeffs mu) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *tile_a = tma_smem_base(smem_raw), *tile_b = tile_a + TmaTile::kBytes;
    uint64_t *bar = reinterpret_cast<uint64_t *>(tile_b + TmaTile::kBytes);
    const int t = threadIdx.x, row0 = blockIdx.x * TmaTile::kRows;
    if (t == 0) {
        tma::mbar_init(bar, 1);
        tma::mbar_expect_tx(bar, (a_row ? 0u : (uint32_t)TmaTile::kBytes) + (b_row ? 0u : (uint32_t)TmaTile::kBytes));
        if (!a_row) tma_load_tile(tile_a, &map_a, bar, row0, mu.stream_hint != 0);
        if (!b_row) tma_load_tile(tile_b, &map_b, bar, row0, mu.stream_hint != 0);
    }
    __syncthreads();  // the barrier word is initialised before anyone polls it
    Tile<real, NB, 1> A, B;
    if (a_row) load_bcast_row(A, a_row);  // one-row operand (reference broadcast rule, batch.rs:361-376): no tile
    if (b_row) load_bcast_row(B, b_row);
    tma::mbar_wait(bar, 0);
    if (!a_row) load_tile_row(A, tile_a, t);
    if (!b_row) load_tile_row(B, tile_b, t);
    TileSink sink{tile_a, t, {}};
    TileSink sink2{tile_b, t, {}};
    (void)sink2;
    (void)mu;
    constexpr int V = 1;
#define LCC_OUT_BEGIN(k) { real acc[V]; real acc2[V]; acc[0] = real(0); acc2[0] = real(0);
#define LCC_TERM6(k, i, j, neg, flags, mi)                                                                   \
    if constexpr (term_in_op(OP, flags, k)) {                                                             \
        real av = (neg) ? -A.v[i][0] : A.v[i][0];                                                      \
        if constexpr (kRuntimeMetric) av = mul_rn(av, mu.v[mi]);                                \
        acc[0] = Arith<STRICT>::madd(av, B.v[j][0], acc[0]);                                           \
        if constexpr (OP == kOpGPOuter && ((flags) & 1)) acc2[0] = Arith<STRICT>::madd(av, B.v[j][0], acc2[0]); \
        if constexpr (two_sided(OP)) {                                                                 \
            real bv = (neg) ? -B.v[i][0] : B.v[i][0];                                                  \
            if constexpr (kRuntimeMetric) bv = mul_rn(bv, mu.v[mi]);                            \
            acc2[0] = Arith<STRICT>::madd(bv, A.v[j][0], acc2[0]);                                     \
        }                                                                                              \
    }
#define LCC_OUT_END(k) finish_two_sided<OP>(acc, acc2); sink.template put<k>(acc); if constexpr (OP == kOpGPOuter) sink2.template put<k>(acc2); }
#define LCC_TERM(k, i, j, neg, flags) LCC_TERM6(k, i, j, neg, flags, (i) & (j))
#define LCC_VTERM(k, i, j, neg, flags, mi) LCC_TERM6(k, i, j, neg, flags, mi)
    if constexpr (LIST == 0) {
#include LCC_TERMS_FILE
    } else if constexpr (LIST == 1) {
#include LCC_VJP0_FILE
    } else {
#include LCC_VJP1_FILE
    }
#undef LCC_VTERM
#undef LCC_TERM6
#undef LCC_OUT_BEGIN
#undef LCC_TERM
#undef LCC_OUT_END
    tma::fence_proxy_async();
    __syncthreads();
    if (t == 0) {
        tma_store_tile(&map_out, tile_a, row0, mu.stream_hint != 0);
        if constexpr (OP == kOpGPOuter) tma_store_tile(&map_out2, tile_b, row0, mu.stream_hint != 0);
        tma::store_commit_and_wait_read();  // the tile must outlive the engine's read of it
    }
}

template <bool STRICT, bool EVEN>
__global__ void __launch_bounds__(TmaTile::kRows)
sandwich_kernel_aos_tma(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_out,
                        const Coeffs R, const Coeffs Rrev, const Coeffs mu) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *tile = tma_smem_base(smem_raw);
    uint64_t *bar = reinterpret_cast<uint64_t *>(tile + TmaTile::kBytes);
    const int t = threadIdx.x, row0 = blockIdx.x * TmaTile::kRows;
    if (t == 0) {
        tma::mbar_init(bar, 1);
        tma::mbar_expect_tx(bar, (uint32_t)TmaTile::kBytes);
        tma_load_tile(tile, &map_x, bar, row0, mu.stream_hint != 0);
    }
    __syncthreads();
    tma::mbar_wait(bar, 0);
    constexpr int V = 1;
    Tile<real, NB, 1> X, RX;
    load_tile_row(X, tile, t);
    (void)mu;
#define LCC_OUT_BEGIN(k) { real acc[V]; acc[0] = real(0);
#define LCC_TERM(k, i, j, neg, flags)                                                                   \
    if constexpr (!EVEN || even_grade(i)) {                                                            \
        real rv = (neg) ? -R.v[i] : R.v[i];                                                            \
        if constexpr (kRuntimeMetric) rv = mul_rn(rv, mu.v[(i) & (j)]);                                \
        acc[0] = Arith<STRICT>::madd(rv, X.v[j][0], acc[0]);                                           \
    }
#define LCC_OUT_END(k) RX.v[k][0] = acc[0]; }
#include LCC_TERMS_FILE
#undef LCC_TERM
#undef LCC_OUT_END
    TileSink sink{tile, t, {}};
#define LCC_TERM(k, i, j, neg, flags)                                                                   \
    if constexpr (!EVEN || even_grade(j)) {                                                            \
        real rv = (neg) ? -Rrev.v[j] : Rrev.v[j];                                                      \
        if constexpr (kRuntimeMetric) rv = mul_rn(rv, mu.v[(i) & (j)]);                                \
        acc[0] = Arith<STRICT>::madd(RX.v[i][0], rv, acc[0]);                                          \
    }
#define LCC_OUT_END(k) sink.template put<k>(acc); }
#include LCC_TERMS_FILE
#undef LCC_OUT_BEGIN
#undef LCC_TERM
#undef LCC_OUT_END
    tma::fence_proxy_async();
    __syncthreads();
    if (t == 0) {
        tma_store_tile(&map_out, tile, row0, mu.stream_hint != 0);
        tma::store_commit_and_wait_read();
    }
}

// ------------------------------------------------------------------------------------------ host
static inline Coeffs to_coeffs(const double *src, double fill) {
    static const int hint = [] { const char *e = std::getenv("LCC_TMA_L2_HINT"); return (e && e[0] == '1') ? 1 : 0; }();
    Coeffs c;
    for (int i = 0; i < NB; ++i) c.v[i] = src ? (real)src[i] : (real)fill;
    c.stream_hint = hint;
    return c;
}

// Threads per block.  These kernels hold 100-170 registers per thread, so a 256-thread block fits
// once or twice per SM and all its warps move through load / compute / store in lock step; small
// blocks (128 threads at 32 blades, 64 below) fit 3-8 times and their phases interleave, which is
// what keeps HBM busy: CGA gp f64 5263 -> 6804 GB/s, PGA sandwich f32 6370 -> 6766 GB/s
// (gpurun_out/blocksize.log, profiles/r01_blocksize_sweep.txt).
// LCC_BLOCK_THREADS (32, 64, 128 or 256) overrides the choice for experiments.
static inline int block_threads() {
    static const int chosen = [] {
        const char *e = std::getenv("LCC_BLOCK_THREADS");
        if (e) { int v = std::atoi(e); if (v == 32 || v == 64 || v == 128 || v == 256) return v; }
        return NB >= 32 ? 128 : 64;
    }();
    return chosen;
}

template <int LAYOUT> static inline unsigned grid_for(int64_t n) {
    constexpr int V = LAYOUT == kSoA ? kRowsSoA : 1;
    const int64_t threads = (n + V - 1) / V, block = block_threads();
    return (unsigned)((threads + block - 1) / block);
}

template <int LAYOUT, int OP, bool STRICT> static cudaError_t launch_product_sum_t(const ProductArgs &g) {
    constexpr int W = 16 / (int)sizeof(real);
    int vec_ok = 1;
    if (LAYOUT == kAoS) vec_ok = ((uintptr_t)g.a % 16 == 0) && (g.lda % W == 0) && ((uintptr_t)g.b % 16 == 0) && (g.ldb % W == 0);
    const int bd = block_threads() < NB ? NB : block_threads();
    const size_t smem = (size_t)NB * bd * sizeof(double);
    auto kernel = product_sum_kernel<LAYOUT, OP, STRICT>;
    if (smem > 48 * 1024) cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    kernel<<<g.sum_blocks, bd, smem, g.stream>>>((const real *)g.a, g.lda, g.a_bcast, g.a_grade, (const real *)g.b, g.ldb,
                                               g.b_bcast, g.n, vec_ok, to_coeffs(g.mu, 1.0), g.sum_partial);
    note_kernel_launch();
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------ K1 / K2 on TMA-staged SoA tiles
// Engine-native blade-major batches: the tile is NB blade rows x SoaTileCfg::kRows batch rows (NB segments of
// contiguous HBM), a thread owns V consecutive batch rows exactly as in the direct-load kernels and reads
// them from shared memory with one 16-byte (or 8-byte) access per blade.
// Tile size per operand: two-operand products keep 2 x 32 KB per block (3 blocks per SM), the sandwich one
// tile of LCC_SOA_SANDWICH_TILE_BYTES.  Longer row segments per blade mean longer HBM bursts: CGA gp f64 at
// 2^26 rows runs at 6837 GB/s with 16 KB tiles, 7012 GB/s with 32 KB and 4657 GB/s with 64 KB (one block per SM);
// PGA sandwich f32 at 2^28 rows: 6836 / 6855 / 6889 GB/s (profiles/r01_soa_tile_sweep.json).
#ifndef LCC_SOA_TILE_BYTES
#define LCC_SOA_TILE_BYTES 32768
#endif
#ifndef LCC_SOA_SANDWICH_TILE_BYTES
#define LCC_SOA_SANDWICH_TILE_BYTES 65536
#endif
template <int TILE_BYTES> struct SoaTileCfg {
    static constexpr int kRows = []() constexpr {
        int r = TILE_BYTES / (NB * (int)sizeof(real));
        r = r > 1024 ? 1024 : r;
        return r < 32 * kRowsSoA ? 32 * kRowsSoA : r;  // at least one warp
    }();
    static constexpr int kBoxRows = kRows > 256 ? 256 : kRows;  // TMA boxes are at most 256 elements wide
    static constexpr int kBoxes = kRows / kBoxRows;
    static constexpr int kBoxBytes = kBoxRows * NB * (int)sizeof(real);
    static constexpr int kThreads = kRows / kRowsSoA;
    static constexpr int kBytes = kRows * NB * (int)sizeof(real);
    // byte offset of (blade k, tile row r): boxes one after the other, each [NB][kBoxRows]
    static __device__ __forceinline__ size_t offset(int k, int r) {
        return ((size_t)(r / kBoxRows) * NB * kBoxRows + (size_t)k * kBoxRows + (r % kBoxRows)) * sizeof(real);
    }
    static __device__ __forceinline__ void tma_load(uint8_t *dst, const CUtensorMap *map, uint64_t *bar, int row0, bool hint) {
#pragma unroll
        for (int box = 0; box < kBoxes; ++box) tma::load_2d(dst + box * kBoxBytes, map, bar, row0 + box * kBoxRows, 0, hint);
    }
    static __device__ __forceinline__ void tma_store(const CUtensorMap *map, const uint8_t *src, int row0, bool hint) {
#pragma unroll
        for (int box = 0; box < kBoxes; ++box) tma::store_2d(map, src + box * kBoxBytes, row0 + box * kBoxRows, 0, hint);
    }
};
using SoaProductTile = SoaTileCfg<LCC_SOA_TILE_BYTES>;
using SoaSandwichTile = SoaTileCfg<LCC_SOA_SANDWICH_TILE_BYTES>;

template <int V> struct alignas(V * sizeof(real)) RowVec { real v[V]; };

template <class CFG, int V>
static __device__ __forceinline__ void load_soa_tile(Tile<real, NB, V> &t, const uint8_t *tile, int first_row) {
#pragma unroll
    for (int k = 0; k < NB; ++k) {
        const RowVec<V> c = *reinterpret_cast<const RowVec<V> *>(tile + CFG::offset(k, first_row));
#pragma unroll
        for (int l = 0; l < V; ++l) t.v[k][l] = c.v[l];
    }
}
template <class CFG, int V> struct SoaTileSink {
    uint8_t *tile; int first_row;
    template <int K> __device__ __forceinline__ void put(const real (&acc)[V]) {
        RowVec<V> c;
#pragma unroll
        for (int l = 0; l < V; ++l) c.v[l] = acc[l];
        *reinterpret_cast<RowVec<V> *>(tile + CFG::offset(K, first_row)) = c;
    }
};
template <int V>
static __device__ __forceinline__ void load_bcast_soa(Tile<real, NB, V> &t, const real *__restrict__ base, int64_t ld) {
#pragma unroll
    for (int k = 0; k < NB; ++k) {
        const real sv = __ldg(base + (int64_t)k * ld);
#pragma unroll
        for (int l = 0; l < V; ++l) t.v[k][l] = sv;
    }
}

template <int OP, bool STRICT, int LIST = 0>
__global__ void __launch_bounds__(SoaProductTile::kThreads)
product_kernel_soa_tma(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                       const __grid_constant__ CUtensorMap map_out, const __grid_constant__ CUtensorMap map_out2,
                       const real *__restrict__ a_row, int64_t lda, const real *__restrict__ b_row, int64_t ldb, const Coeffs mu) {
    constexpr int V = kRowsSoA;
    using T_ = SoaProductTile;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *tile_a = tma_smem_base(smem_raw), *tile_b = tile_a + T_::kBytes;
    uint64_t *bar = reinterpret_cast<uint64_t *>(tile_b + T_::kBytes);
    const int t = threadIdx.x, row0 = blockIdx.x * T_::kRows;
    if (t == 0) {
        tma::mbar_init(bar, 1);
        tma::mbar_expect_tx(bar, (a_row ? 0u : (uint32_t)T_::kBytes) + (b_row ? 0u : (uint32_t)T_::kBytes));
        if (!a_row) T_::tma_load(tile_a, &map_a, bar, row0, mu.stream_hint != 0);
        if (!b_row) T_::tma_load(tile_b, &map_b, bar, row0, mu.stream_hint != 0);
    }
    __syncthreads();
    Tile<real, NB, V> A, B;
    if (a_row) load_bcast_soa(A, a_row, lda);
    if (b_row) load_bcast_soa(B, b_row, ldb);
    tma::mbar_wait(bar, 0);
    if (!a_row) load_soa_tile<T_>(A, tile_a, t * V);
    if (!b_row) load_soa_tile<T_>(B, tile_b, t * V);
    SoaTileSink<T_, V> sink{tile_a, t * V};
    SoaTileSink<T_, V> sink2{tile_b, t * V};
    (void)sink2;
    (void)mu;
#define LCC_OUT_BEGIN(k) { real acc[V]; real acc2[V]; _Pragma("unroll") for (int l = 0; l < V; ++l) { acc[l] = real(0); acc2[l] = real(0); }
#define LCC_TERM6(k, i, j, neg, flags, mi)                                                                   \
    if constexpr (term_in_op(OP, flags, k)) {                                                             \
        _Pragma("unroll") for (int l = 0; l < V; ++l) {                                                \
            real av = (neg) ? -A.v[i][l] : A.v[i][l];                                                  \
            if constexpr (kRuntimeMetric) av = mul_rn(av, mu.v[mi]);                            \
            acc[l] = Arith<STRICT>::madd(av, B.v[j][l], acc[l]);                                       \
            if constexpr (OP == kOpGPOuter && ((flags) & 1)) acc2[l] = Arith<STRICT>::madd(av, B.v[j][l], acc2[l]); \
            if constexpr (two_sided(OP)) {                                                             \
                real bv = (neg) ? -B.v[i][l] : B.v[i][l];                                              \
                if constexpr (kRuntimeMetric) bv = mul_rn(bv, mu.v[mi]);                        \
                acc2[l] = Arith<STRICT>::madd(bv, A.v[j][l], acc2[l]);                                 \
            }                                                                                          \
        }                                                                                              \
    }
#define LCC_OUT_END(k) finish_two_sided<OP>(acc, acc2); sink.template put<k>(acc); if constexpr (OP == kOpGPOuter) sink2.template put<k>(acc2); }
#define LCC_TERM(k, i, j, neg, flags) LCC_TERM6(k, i, j, neg, flags, (i) & (j))
#define LCC_VTERM(k, i, j, neg, flags, mi) LCC_TERM6(k, i, j, neg, flags, mi)
    if constexpr (LIST == 0) {
#include LCC_TERMS_FILE
    } else if constexpr (LIST == 1) {
#include LCC_VJP0_FILE
    } else {
#include LCC_VJP1_FILE
    }
#undef LCC_VTERM
#undef LCC_TERM6
#undef LCC_OUT_BEGIN
#undef LCC_TERM
#undef LCC_OUT_END
    tma::fence_proxy_async();
    __syncthreads();
    if (t == 0) {
        T_::tma_store(&map_out, tile_a, row0, mu.stream_hint != 0);
        if constexpr (OP == kOpGPOuter) T_::tma_store(&map_out2, tile_b, row0, mu.stream_hint != 0);
        tma::store_commit_and_wait_read();
    }
}

template <bool STRICT, bool EVEN>
__global__ void __launch_bounds__(SoaSandwichTile::kThreads)
sandwich_kernel_soa_tma(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_out,
                        const Coeffs R, const Coeffs Rrev, const Coeffs mu) {
    constexpr int V = kRowsSoA;
    using T_ = SoaSandwichTile;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *tile = tma_smem_base(smem_raw);
    uint64_t *bar = reinterpret_cast<uint64_t *>(tile + T_::kBytes);
    const int t = threadIdx.x, row0 = blockIdx.x * T_::kRows;
    if (t == 0) {
        tma::mbar_init(bar, 1);
        tma::mbar_expect_tx(bar, (uint32_t)T_::kBytes);
        T_::tma_load(tile, &map_x, bar, row0, mu.stream_hint != 0);
    }
    __syncthreads();
    tma::mbar_wait(bar, 0);
    Tile<real, NB, V> X, RX;
    load_soa_tile<T_>(X, tile, t * V);
    (void)mu;
#define LCC_OUT_BEGIN(k) { real acc[V]; _Pragma("unroll") for (int l = 0; l < V; ++l) acc[l] = real(0);
#define LCC_TERM(k, i, j, neg, flags)                                                                   \
    if constexpr (!EVEN || even_grade(i)) {                                                            \
        real rv = (neg) ? -R.v[i] : R.v[i];                                                            \
        if constexpr (kRuntimeMetric) rv = mul_rn(rv, mu.v[(i) & (j)]);                                \
        _Pragma("unroll") for (int l = 0; l < V; ++l) acc[l] = Arith<STRICT>::madd(rv, X.v[j][l], acc[l]); \
    }
#define LCC_OUT_END(k) _Pragma("unroll") for (int l = 0; l < V; ++l) RX.v[k][l] = acc[l]; }
#include LCC_TERMS_FILE
#undef LCC_TERM
#undef LCC_OUT_END
    SoaTileSink<T_, V> sink{tile, t * V};
#define LCC_TERM(k, i, j, neg, flags)                                                                   \
    if constexpr (!EVEN || even_grade(j)) {                                                            \
        real rv = (neg) ? -Rrev.v[j] : Rrev.v[j];                                                      \
        if constexpr (kRuntimeMetric) rv = mul_rn(rv, mu.v[(i) & (j)]);                                \
        _Pragma("unroll") for (int l = 0; l < V; ++l) acc[l] = Arith<STRICT>::madd(RX.v[i][l], rv, acc[l]); \
    }
#define LCC_OUT_END(k) sink.template put<k>(acc); }
#include LCC_TERMS_FILE
#undef LCC_OUT_BEGIN
#undef LCC_TERM
#undef LCC_OUT_END
    tma::fence_proxy_async();
    __syncthreads();
    if (t == 0) {
        T_::tma_store(&map_out, tile, row0, mu.stream_hint != 0);
        tma::store_commit_and_wait_read();
    }
}

// Which batches go through the TMA-staged kernels (lcc_set_option, include/lcc_b200.h; size sweep in
// profiles/r01_tma_threshold_sweep.jsonl):
//   AoS  rows of 128 bytes or more: at least aos_tma_min_rows rows (default 4096; CGA f64 gp already wins at 4096 rows,
//        6.3 vs 8.3 us, and runs 2.2-2.5x faster from 2^16 rows on).  Narrower rows (PGA f32, R3 f64: 64 bytes) waste
//        less per direct access and need 32x as many rows (131072) before the tiles win (2^18 rows: 8.3 vs 10.4 us).
//        LCC_AOS_TMA=0 forces the direct-load kernels.
//   SoA  at least soa_tma_min_bytes per operand (default 8 GiB): the direct-load SoA kernels are already fully
//        coalesced and are as fast or faster up to 4 GiB per operand (PGA f32 2^26 rows: 6854 vs 6832 GB/s; 2^22 rows:
//        84 vs 91 us); the tiles only win on the very largest batches (17 GB operands: 6759 -> 6889 GB/s, CGA f64
//        6773 -> 7012 GB/s), where the direct kernels' many widely spaced streams start to miss in the TLB.
constexpr int64_t kAosNarrowRowFactor = kRowBytes >= 128 ? 1 : 32;
static inline bool tma_path_enabled() {
    static const bool on = [] { const char *e = std::getenv("LCC_AOS_TMA"); return !(e && e[0] == '0'); }();
    return on;
}
static inline bool aos_map(CUtensorMap *m, const void *base, int64_t ld, int64_t n) {
    return make_aos_tensor_map(m, base, (int)sizeof(real), NB, ld, n, TmaTile::kSpan, TmaTile::kRows);
}

// returns cudaErrorNotSupported when the operands do not qualify (caller falls back to the direct kernel)
template <int OP, bool STRICT, int LIST = 0> static cudaError_t launch_product_tma(const ProductArgs &g) {
    if constexpr (!kHasTmaPath) {
        return cudaErrorNotSupported;
    } else {
        if (!tma_path_enabled() || g.n < option_aos_tma_min_rows() * kAosNarrowRowFactor) return cudaErrorNotSupported;
        CUtensorMap ma, mb, mo, mo2;
        if (!aos_map(&mo, g.out, g.ldo, g.n)) return cudaErrorNotSupported;
        ma = mb = mo2 = mo;  // placeholders for operands that do not use a tile
        if (!g.a_bcast && !aos_map(&ma, g.a, g.lda, g.n)) return cudaErrorNotSupported;
        if (!g.b_bcast && !aos_map(&mb, g.b, g.ldb, g.n)) return cudaErrorNotSupported;
        if (OP == kOpGPOuter && !aos_map(&mo2, g.out2, g.ldo2, g.n)) return cudaErrorNotSupported;
        auto kernel = product_kernel_aos_tma<OP, STRICT, LIST>;
        constexpr int smem = 2 * TmaTile::kBytes + 1024 + 16;
        // (set on every call: the attribute belongs to the (function, device) pair and one process may drive several devices)
        const cudaError_t attr = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (attr != cudaSuccess) return attr;
        const unsigned grid = (unsigned)((g.n + TmaTile::kRows - 1) / TmaTile::kRows);
        kernel<<<grid, TmaTile::kRows, smem, g.stream>>>(ma, mb, mo, mo2, g.a_bcast ? (const real *)g.a : nullptr,
                                                        g.b_bcast ? (const real *)g.b : nullptr, to_coeffs(g.mu, 1.0));
        note_kernel_launch();
        return cudaGetLastError();
    }
}

template <bool STRICT, bool EVEN> static cudaError_t launch_sandwich_tma(const SandwichArgs &g, const Coeffs &R, const Coeffs &Rrev) {
    if constexpr (!kHasTmaPath) {
        return cudaErrorNotSupported;
    } else {
        if (!tma_path_enabled() || g.n < option_aos_tma_min_rows() * kAosNarrowRowFactor) return cudaErrorNotSupported;
        CUtensorMap mx, mo;
        if (!aos_map(&mx, g.x, g.ldx, g.n) || !aos_map(&mo, g.out, g.ldo, g.n)) return cudaErrorNotSupported;
        auto kernel = sandwich_kernel_aos_tma<STRICT, EVEN>;
        constexpr int smem = TmaTile::kBytes + 1024 + 16;
        const unsigned grid = (unsigned)((g.n + TmaTile::kRows - 1) / TmaTile::kRows);
        kernel<<<grid, TmaTile::kRows, smem, g.stream>>>(mx, mo, R, Rrev, to_coeffs(g.mu, 1.0));
        note_kernel_launch();
        return cudaGetLastError();
    }
}

static inline bool soa_tma_enabled() {  // LCC_SOA_TMA=0 forces the direct-load SoA kernels
    static const bool on = [] { const char *e = std::getenv("LCC_SOA_TMA"); return !(e && e[0] == '0'); }();
    return on;
}
template <class CFG> static inline bool soa_map(CUtensorMap *m, const void *base, int64_t ld, int64_t n) {
    return make_soa_tensor_map(m, base, (int)sizeof(real), NB, ld, n, CFG::kBoxRows);
}

template <int OP, bool STRICT, int LIST = 0> static cudaError_t launch_product_soa_tma(const ProductArgs &g) {
    using T_ = SoaProductTile;
    if (!soa_tma_enabled() || g.n * (int64_t)kRowBytes < option_soa_tma_min_bytes()) return cudaErrorNotSupported;
    CUtensorMap ma, mb, mo, mo2;
    if (!soa_map<T_>(&mo, g.out, g.ldo, g.n)) return cudaErrorNotSupported;
    ma = mb = mo2 = mo;
    if (!g.a_bcast && !soa_map<T_>(&ma, g.a, g.lda, g.n)) return cudaErrorNotSupported;
    if (!g.b_bcast && !soa_map<T_>(&mb, g.b, g.ldb, g.n)) return cudaErrorNotSupported;
    if (OP == kOpGPOuter && !soa_map<T_>(&mo2, g.out2, g.ldo2, g.n)) return cudaErrorNotSupported;
    auto kernel = product_kernel_soa_tma<OP, STRICT, LIST>;
    constexpr int smem = 2 * T_::kBytes + 1024 + 16;
    const cudaError_t attr = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (attr != cudaSuccess) return attr;
    const unsigned grid = (unsigned)((g.n + T_::kRows - 1) / T_::kRows);
    kernel<<<grid, T_::kThreads, smem, g.stream>>>(ma, mb, mo, mo2, g.a_bcast ? (const real *)g.a : nullptr, g.lda,
                                                  g.b_bcast ? (const real *)g.b : nullptr, g.ldb, to_coeffs(g.mu, 1.0));
    note_kernel_launch();
    return cudaGetLastError();
}

template <bool STRICT, bool EVEN> static cudaError_t launch_sandwich_soa_tma(const SandwichArgs &g, const Coeffs &R, const Coeffs &Rrev) {
    using T_ = SoaSandwichTile;
    if (!soa_tma_enabled() || g.n * (int64_t)kRowBytes < option_soa_tma_min_bytes()) return cudaErrorNotSupported;
    CUtensorMap mx, mo;
    if (!soa_map<T_>(&mx, g.x, g.ldx, g.n) || !soa_map<T_>(&mo, g.out, g.ldo, g.n)) return cudaErrorNotSupported;
    auto kernel = sandwich_kernel_soa_tma<STRICT, EVEN>;
    constexpr int smem = T_::kBytes + 1024 + 16;
    const cudaError_t attr = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (attr != cudaSuccess) return attr;
    const unsigned grid = (unsigned)((g.n + T_::kRows - 1) / T_::kRows);
    kernel<<<grid, T_::kThreads, smem, g.stream>>>(mx, mo, R, Rrev, to_coeffs(g.mu, 1.0));
    note_kernel_launch();
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------ K2 on compact PGA points
// out_xyz = pga_point_coords(M * pga_point(xyz) * ~M) without ever materialising the 16-blade rows: 24 bytes per
// point (f32) instead of the 128 of the dense layout.  Embedding and extraction are the reference's
// (src/pga.rs:131-134: c[7]=1, c[14]=-x, c[13]=y, c[11]=-z;  :579-599: x=-c[14]/w, y=c[13]/w, z=-c[11]/w, w=c[7]),
// the two product stages are the sandwich of batch.rs:452-497 with the terms whose x blade is structurally zero
// left out (the reference skips them at run time: `if b == 0.0 { continue }`) and only the four blades the
// extraction reads computed in stage two -- same roundings in the same order, so strict mode is bit-exact.
// Rows whose weight |w| < 1e-10 have no finite point: NaN, as lcc_pga_points_extract writes them.
// Data movement: a block's 2048 points (f32; 1024 in f64) are one contiguous 24 KB span; it is copied to shared memory
// with one cp.async.bulk (1-D TMA), each thread handles 8 points (stride 256: bank-conflict free at 3 words per point),
// the results overwrite the tile and leave with one bulk store.  The last, partial block uses plain accesses.
// Default mode divides once per point (reciprocal of the weight, then three multiplies; <= 1 ulp from the quotients).
#if LCC_NB == 16 && !LCC_RUNTIME_METRIC && defined(LCC_SIG_PGA3D)
#ifndef LCC_POINTS_SHAPE_DEFAULT
#define LCC_POINTS_SHAPE_DEFAULT 1
#endif
// Block shape: THREADS threads, PPT points per thread.  More bytes per resident block keep more of the HBM pipe
// busy (Little's law: ~70 KB must be in flight per SM); LCC_POINTS_SHAPE selects among the compiled shapes.
// 2^28 f32 points: 256 x 4 (12 KB tiles) 6201 GB/s, 256 x 8 (24 KB) 6888 GB/s, 128 x 8 (12 KB, 16 blocks per SM)
// 6874 GB/s (profiles/r01_points_shape_sweep.txt); 256 x 8 is the default.
__host__ __device__ constexpr bool point_blade(int k) { return k == 7 || k == 11 || k == 13 || k == 14; }

template <bool STRICT, bool EVEN>
static __device__ __forceinline__ void transform_point(const Coeffs &R, const Coeffs &Rrev, real x, real y, real z, real (&o)[3]) {
    constexpr int V = 1;
    Tile<real, NB, 1> X, RX;
#pragma unroll
    for (int k = 0; k < NB; ++k) { X.v[k][0] = real(0); RX.v[k][0] = real(0); }
    X.v[7][0] = real(1); X.v[14][0] = -x; X.v[13][0] = y; X.v[11][0] = -z;
#define LCC_OUT_BEGIN(k) { real acc[V]; acc[0] = real(0);
#define LCC_TERM(k, i, j, neg, flags)                                                                   \
    if constexpr (point_blade(j) && (!EVEN || even_grade(i))) {                                        \
        const real rv = (neg) ? -R.v[i] : R.v[i];                                                      \
        acc[0] = Arith<STRICT>::madd(rv, X.v[j][0], acc[0]);                                           \
    }
#define LCC_OUT_END(k) RX.v[k][0] = acc[0]; }
#include LCC_TERMS_FILE
#undef LCC_TERM
#undef LCC_OUT_END
    real res[NB];
#define LCC_TERM(k, i, j, neg, flags)                                                                   \
    if constexpr (point_blade(k) && (!EVEN || even_grade(j))) {                                        \
        const real rv = (neg) ? -Rrev.v[j] : Rrev.v[j];                                                \
        acc[0] = Arith<STRICT>::madd(RX.v[i][0], rv, acc[0]);                                          \
    }
#define LCC_OUT_END(k) res[k] = acc[0]; }
#include LCC_TERMS_FILE
#undef LCC_OUT_BEGIN
#undef LCC_TERM
#undef LCC_OUT_END
    const real w = res[7];
    const bool ok = (w < real(0) ? -w : w) >= (real)1e-10;
    const real nanv = (real)__int_as_float(0x7fc00000);
    if constexpr (STRICT) {  // the reference's three divisions (pga.rs:579-599)
        o[0] = ok ? div_rn(-res[14], w) : nanv;
        o[1] = ok ? div_rn(res[13], w) : nanv;
        o[2] = ok ? div_rn(-res[11], w) : nanv;
    } else {                 // default mode: one correctly rounded reciprocal per point, <= 1 ulp from the quotient
        const real rw = div_rn(real(1), w);
        o[0] = ok ? mul_rn(-res[14], rw) : nanv;
        o[1] = ok ? mul_rn(res[13], rw) : nanv;
        o[2] = ok ? mul_rn(-res[11], rw) : nanv;
    }
}

template <bool STRICT, bool EVEN, int kPointThreads, int PPT>
__global__ void __launch_bounds__(kPointThreads)
pga_points_sandwich_kernel(const real *__restrict__ xyz_in, real *__restrict__ xyz_out, int64_t n, int bulk_ok,
                           const Coeffs R, const Coeffs Rrev) {
    constexpr int kPointsPerBlock = kPointThreads * PPT;
    __shared__ alignas(128) real tile[kPointsPerBlock * 3];
    __shared__ uint64_t bar;
    const int t = threadIdx.x;
    const int64_t p0 = (int64_t)blockIdx.x * kPointsPerBlock;
    const bool full = bulk_ok && p0 + kPointsPerBlock <= n;
    constexpr uint32_t kTileBytes = kPointsPerBlock * 3 * sizeof(real);
    if (full) {
        if (t == 0) {
            tma::mbar_init(&bar, 1);
            tma::mbar_expect_tx(&bar, kTileBytes);
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(tma::smem_u32(tile)), "l"(xyz_in + p0 * 3), "r"(kTileBytes), "r"(tma::smem_u32(&bar)) : "memory");
        }
        __syncthreads();
        tma::mbar_wait(&bar, 0);
    }
#pragma unroll
    for (int q = 0; q < kPointsPerBlock / kPointThreads; ++q) {
        const int lp = t + q * kPointThreads;
        const int64_t p = p0 + lp;
        if (p >= n) break;
        real x, y, z, o[3];
        if (full) { x = tile[lp * 3]; y = tile[lp * 3 + 1]; z = tile[lp * 3 + 2]; }
        else { x = __ldg(xyz_in + p * 3); y = __ldg(xyz_in + p * 3 + 1); z = __ldg(xyz_in + p * 3 + 2); }
        transform_point<STRICT, EVEN>(R, Rrev, x, y, z, o);
        if (full) { tile[lp * 3] = o[0]; tile[lp * 3 + 1] = o[1]; tile[lp * 3 + 2] = o[2]; }
        else { xyz_out[p * 3] = o[0]; xyz_out[p * 3 + 1] = o[1]; xyz_out[p * 3 + 2] = o[2]; }
    }
    if (full) {
        tma::fence_proxy_async();
        __syncthreads();
        if (t == 0) {
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                         ::"l"(xyz_out + p0 * 3), "r"(tma::smem_u32(tile)), "r"(kTileBytes) : "memory");
            tma::store_commit_and_wait_read();
        }
    }
}

cudaError_t pga_points_sandwich(const SandwichArgs &g);  // x / out = dense (n,3) arrays here
#if LCC_PART == 3 && LCC_PART_STRICT == 0
template <bool STRICT, bool EVEN, int THREADS, int PPT>
static cudaError_t launch_points_shape(const SandwichArgs &g, const Coeffs &R, const Coeffs &Rrev) {
    const int bulk_ok = ((uintptr_t)g.x % 16 == 0) && ((uintptr_t)g.out % 16 == 0);
    constexpr int kPointsPerBlock = THREADS * PPT;
    const unsigned grid = (unsigned)((g.n + kPointsPerBlock - 1) / kPointsPerBlock);
    pga_points_sandwich_kernel<STRICT, EVEN, THREADS, PPT><<<grid, THREADS, 0, g.stream>>>((const real *)g.x, (real *)g.out, g.n, bulk_ok, R, Rrev);
    note_kernel_launch();
    return cudaGetLastError();
}
static inline int points_shape() {  // 0: 256 threads x 4 points (12 KB f32 tiles), 1: 256 x 8, 2: 128 x 8
    static const int shape = [] { const char *e = std::getenv("LCC_POINTS_SHAPE"); return e ? std::atoi(e) : LCC_POINTS_SHAPE_DEFAULT; }();
    return shape;
}
template <bool STRICT, bool EVEN> static cudaError_t launch_points_t(const SandwichArgs &g, const Coeffs &R, const Coeffs &Rrev) {
    switch (points_shape()) {
        case 1: return launch_points_shape<STRICT, EVEN, 256, (sizeof(real) == 4 ? 8 : 4)>(g, R, Rrev);  // 24 KB tiles
        case 2: return launch_points_shape<STRICT, EVEN, 128, 8>(g, R, Rrev);
        default: return launch_points_shape<STRICT, EVEN, 256, 4>(g, R, Rrev);
    }
}
cudaError_t pga_points_sandwich(const SandwichArgs &g) {
    if (g.n <= 0) return cudaSuccess;
    Coeffs R = to_coeffs(g.rotor, 0.0), Rrev;
    Rrev.stream_hint = R.stream_hint;
    for (int i = 0; i < NB; ++i) {
        int grade = 0;
        for (int b = i; b; b >>= 1) grade += b & 1;
        Rrev.v[i] = R.v[i] * ((grade % 4 < 2) ? real(1) : real(-1));  // src/batch.rs:452-457
    }
    if (g.strict) return g.even ? launch_points_t<true, true>(g, R, Rrev) : launch_points_t<true, false>(g, R, Rrev);
    return g.even ? launch_points_t<false, true>(g, R, Rrev) : launch_points_t<false, false>(g, R, Rrev);
}
#endif
#endif  // PGA3D

// ------------------------------------------------------------------------------------------ K2r: one rotor PER ROW
// out[row] = R[row] * x[row] * ~R[row]  (reference: lcc/torch/multivector.py:988-1019 -- rx = gp(R, x), then gp(rx, ~R)
// with ~R = R * reverse signs).  One pass: R and x in registers, the intermediate rx rounded to storage precision as the
// reference's tensor, ~R never materialised (its sign is folded into the term's sign at compile time).  3 * nb * s bytes
// per row instead of the 8 * nb * s of reverse + two products through two full-size temporaries.  `r_bcast`: one rotor
// row on the DEVICE for the whole batch (what lcc.torch has when the rotor is a CUDA tensor: no host read-back).
__host__ __device__ constexpr bool reverse_negative(int blade) {
    int g = 0;
    for (int b = blade; b; b >>= 1) g += b & 1;
    return (g % 4) >= 2;  // blades.rs:128-135
}

// The two stages on register tiles.  EVEN: the odd-grade coefficients of R are known to be zero (rotors, motors), their
// terms are pruned at compile time -- half of both products; pruned terms would only have added +-0.
template <bool STRICT, bool EVEN, int V, class SinkT>
static __device__ __forceinline__ void sandwich_rows_body(const Tile<real, NB, V> &R, const Tile<real, NB, V> &X, SinkT &sink, const Coeffs &mu) {
    Tile<real, NB, V> RX;
    (void)mu;
#define LCC_OUT_BEGIN(k) { real acc[V]; _Pragma("unroll") for (int l = 0; l < V; ++l) acc[l] = real(0);
#define LCC_TERM(k, i, j, neg, flags)                                                                   \
    if constexpr (!EVEN || even_grade(i)) {                                                            \
        _Pragma("unroll") for (int l = 0; l < V; ++l) {                                                \
            real rv = (neg) ? -R.v[i][l] : R.v[i][l];                                                  \
            if constexpr (kRuntimeMetric) rv = mul_rn(rv, mu.v[(i) & (j)]);                            \
            acc[l] = Arith<STRICT>::madd(rv, X.v[j][l], acc[l]);                                       \
        }                                                                                              \
    }
#define LCC_OUT_END(k) _Pragma("unroll") for (int l = 0; l < V; ++l) RX.v[k][l] = acc[l]; }
#include LCC_TERMS_FILE
#undef LCC_TERM
#undef LCC_OUT_END
#define LCC_TERM(k, i, j, neg, flags)                                                                   \
    if constexpr (!EVEN || even_grade(j)) {                                                            \
        _Pragma("unroll") for (int l = 0; l < V; ++l) {                                                \
            real av = (neg) ? -RX.v[i][l] : RX.v[i][l];                                                \
            if constexpr (kRuntimeMetric) av = mul_rn(av, mu.v[(i) & (j)]);                            \
            const real rv = reverse_negative(j) ? -R.v[j][l] : R.v[j][l];                              \
            acc[l] = Arith<STRICT>::madd(av, rv, acc[l]);                                              \
        }                                                                                              \
    }
#define LCC_OUT_END(k) sink.template put<k>(acc); }
#include LCC_TERMS_FILE
#undef LCC_OUT_BEGIN
#undef LCC_TERM
#undef LCC_OUT_END
}
// Are the rotors this WARP holds all even (odd-grade coefficients zero)?  One device rotor for the whole batch: every
// thread holds the same row; one rotor per row: rotor fields are even versors in practice, and the vote keeps the branch
// warp-uniform when a batch mixes both kinds.
template <int V> static __device__ __forceinline__ bool warp_rotors_are_even(const Tile<real, NB, V> &R) {
    bool even = true;
#pragma unroll
    for (int i = 0; i < NB; ++i)
        if (!even_grade(i)) {
#pragma unroll
            for (int l = 0; l < V; ++l) even = even && (R.v[i][l] == real(0));
        }
    return __all_sync(__activemask(), even) != 0;
}

template <int LAYOUT, bool STRICT>
__global__ void __launch_bounds__(kThreads)
sandwich_rows_kernel(const real *__restrict__ r, int64_t ldr, int r_bcast, const real *__restrict__ x, int64_t ldx,
                     real *__restrict__ out, int64_t ldo, int64_t n, int vec_ok, const Coeffs mu) {
    constexpr int V = LAYOUT == kSoA ? kRowsSoA : 1;
    const int64_t r0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * V;
    if (r0 >= n) return;
    Tile<real, NB, V> R, X;
    if constexpr (LAYOUT == kSoA) { load_soa(R, r, ldr, r0, r_bcast != 0); load_soa(X, x, ldx, r0, false); }
    else { load_aos(R, r, ldr, r0, r_bcast != 0, vec_ok != 0); load_aos(X, x, ldx, r0, false, vec_ok != 0); }
    Sink<LAYOUT, V> sink{out, ldo, r0, vec_ok != 0, {}};
    // (blade-major threads hold V rows: the vote over per-row rotors costs them registers and occupancy -- measured
    // 6.4 -> 4.8 TB/s on PGA f32 -- so there only the one-device-rotor case is tested)
    bool even = false;
    if constexpr (V == 1) even = warp_rotors_are_even(R);
    else if (r_bcast) {
        even = true;
#pragma unroll
        for (int i = 0; i < NB; ++i)
            if (!even_grade(i)) even = even && (R.v[i][0] == real(0));
    }
    if (even) sandwich_rows_body<STRICT, true, V>(R, X, sink, mu);
    else sandwich_rows_body<STRICT, false, V>(R, X, sink, mu);
}

// the same on TMA-staged AoS tiles (two input tiles; the result overwrites the x tile and leaves by TMA store)
template <bool STRICT>
__global__ void __launch_bounds__(TmaTile::kRows)
sandwich_rows_kernel_aos_tma(const __grid_constant__ CUtensorMap map_r, const __grid_constant__ CUtensorMap map_x,
                             const __grid_constant__ CUtensorMap map_out, const real *__restrict__ r_row, const Coeffs mu) {
    extern __shared__ uint8_t smem_raw[];
    // one device rotor for the batch: no rotor tile, the block is launched with one tile of shared memory (twice the
    // blocks, and bytes in flight, per SM)
    uint8_t *tile_r = tma_smem_base(smem_raw), *tile_x = r_row ? tile_r : tile_r + TmaTile::kBytes;
    uint64_t *bar = reinterpret_cast<uint64_t *>(tile_x + TmaTile::kBytes);
    const int t = threadIdx.x, row0 = blockIdx.x * TmaTile::kRows;
    if (t == 0) {
        tma::mbar_init(bar, 1);
        tma::mbar_expect_tx(bar, (r_row ? 0u : (uint32_t)TmaTile::kBytes) + (uint32_t)TmaTile::kBytes);
        if (!r_row) tma_load_tile(tile_r, &map_r, bar, row0, mu.stream_hint != 0);
        tma_load_tile(tile_x, &map_x, bar, row0, mu.stream_hint != 0);
    }
    __syncthreads();
    Tile<real, NB, 1> R, X;
    if (r_row) load_bcast_row(R, r_row);
    tma::mbar_wait(bar, 0);
    if (!r_row) load_tile_row(R, tile_r, t);
    load_tile_row(X, tile_x, t);
    TileSink sink{tile_x, t, {}};
    if (warp_rotors_are_even(R)) sandwich_rows_body<STRICT, true, 1>(R, X, sink, mu);
    else sandwich_rows_body<STRICT, false, 1>(R, X, sink, mu);
    tma::fence_proxy_async();
    __syncthreads();
    if (t == 0) {
        tma_store_tile(&map_out, tile_x, row0, mu.stream_hint != 0);
        tma::store_commit_and_wait_read();
    }
}

// Cotangents of the per-row sandwich  y = R x ~R  (autograd of lcc.torch, /root/reference/lcc/torch/multivector.py:988-1019:
// there the backward pass is torch autograd through two products and a reverse).  ONE kernel: R, x and the incoming
// cotangent g in registers per row; with t = R x and y = t ~R every generated term  out[k] += s a[i] b[j]  gives its two
// transposed terms  ga[i] += s g[k] b[j],  gb[j] += s a[i] g[k]:
//     g_t = dy/dt^T g,  g_(~R) = dy/d(~R)^T g      (second product)
//     g_R = dt/dR^T g_t + rev . g_(~R),  g_x = dt/dx^T g_t      (first product; ~R[j] = rev(j) R[j])
// 3 reads + 2 writes of nb values per row (5 * nb * s bytes) instead of five product-shaped passes (15 * nb * s) plus
// the saved intermediate of the composed form.  FMA arithmetic: gradients have no reference rounding order.
template <int LAYOUT>
__global__ void __launch_bounds__(kThreads)
sandwich_rows_vjp_kernel(const real *__restrict__ r, int64_t ldr, int r_bcast, const real *__restrict__ x, int64_t ldx,
                         const real *__restrict__ g, int64_t ldg, real *__restrict__ gx, int64_t ldgx,
                         real *__restrict__ gr, int64_t ldgr, int64_t n, int vec_ok, const Coeffs mu) {
    const int64_t r0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // one row per thread in both layouts
    if (r0 >= n) return;
    Tile<real, NB, 1> R, X, G, T, GT, GR, GX;
    if constexpr (LAYOUT == kSoA) { load_soa(R, r, ldr, r0, r_bcast != 0); load_soa(X, x, ldx, r0, false); load_soa(G, g, ldg, r0, false); }
    else { load_aos(R, r, ldr, r0, r_bcast != 0, vec_ok != 0); load_aos(X, x, ldx, r0, false, vec_ok != 0); load_aos(G, g, ldg, r0, false, vec_ok != 0); }
    (void)mu;
#pragma unroll
    for (int k = 0; k < NB; ++k) { T.v[k][0] = real(0); GT.v[k][0] = real(0); GR.v[k][0] = real(0); GX.v[k][0] = real(0); }
#define LCC_OUT_BEGIN(k) {
#define LCC_OUT_END(k) }
    // t = R x
#define LCC_TERM(k, i, j, neg, flags)                                                                   \
    {                                                                                                  \
        real rv = (neg) ? -R.v[i][0] : R.v[i][0];                                                      \
        if constexpr (kRuntimeMetric) rv = mul_rn(rv, mu.v[(i) & (j)]);                                \
        T.v[k][0] = Arith<false>::madd(rv, X.v[j][0], T.v[k][0]);                                      \
    }
#include LCC_TERMS_FILE
#undef LCC_TERM
    // y = t ~R:  g_t[i] += s g[k] ~R[j];  g_(~R)[j] += s t[i] g[k], accumulated into g_R[j] with the reverse sign
#define LCC_TERM(k, i, j, neg, flags)                                                                   \
    {                                                                                                  \
        real gv = (neg) ? -G.v[k][0] : G.v[k][0];                                                      \
        if constexpr (kRuntimeMetric) gv = mul_rn(gv, mu.v[(i) & (j)]);                                \
        const real gs = reverse_negative(j) ? -gv : gv;                                                \
        GT.v[i][0] = Arith<false>::madd(gs, R.v[j][0], GT.v[i][0]);                                    \
        GR.v[j][0] = Arith<false>::madd(gs, T.v[i][0], GR.v[j][0]);                                    \
    }
#include LCC_TERMS_FILE
#undef LCC_TERM
    // t = R x:  g_R[i] += s g_t[k] x[j];  g_x[j] += s R[i] g_t[k]
#define LCC_TERM(k, i, j, neg, flags)                                                                   \
    {                                                                                                  \
        real gv = (neg) ? -GT.v[k][0] : GT.v[k][0];                                                    \
        if constexpr (kRuntimeMetric) gv = mul_rn(gv, mu.v[(i) & (j)]);                                \
        GR.v[i][0] = Arith<false>::madd(gv, X.v[j][0], GR.v[i][0]);                                    \
        GX.v[j][0] = Arith<false>::madd(gv, R.v[i][0], GX.v[j][0]);                                    \
    }
#include LCC_TERMS_FILE
#undef LCC_TERM
#undef LCC_OUT_BEGIN
#undef LCC_OUT_END
    Sink<LAYOUT, 1> sx{gx, ldgx, r0, vec_ok != 0, {}}, sr{gr, ldgr, r0, vec_ok != 0, {}};
    put_all<0>(sx, GX);
    put_all<0>(sr, GR);
}

// ------------------------------------------------------------------------------------------ entry points, by part
// One (signature, precision) is compiled as several object files so that ptxas runs in parallel
// (_build.py writes one stub per part):
//   LCC_PART 1  product kernels with direct loads            (LCC_PART_STRICT = 0 / 1)
//   LCC_PART 2  product kernels on TMA tiles, AoS and SoA    (LCC_PART_STRICT = 0 / 1)
//   LCC_PART 3  sandwich kernels, direct and TMA             (LCC_PART_STRICT = 0 / 1)
//   LCC_PART 4  fused product + batch-sum kernels and the two routers the dispatcher calls
//   LCC_PART 5  vector-Jacobian products (direct and TMA forms)
//   LCC_PART 6  per-row sandwich (direct and AoS-TMA forms, both rounding modes) and its fused cotangent kernel
cudaError_t product_direct_s0(const ProductArgs &g);
cudaError_t product_direct_s1(const ProductArgs &g);
cudaError_t product_tma_s0(const ProductArgs &g);  // cudaErrorNotSupported: operands do not qualify, use the direct kernels
cudaError_t product_tma_s1(const ProductArgs &g);
cudaError_t sandwich_s0(const SandwichArgs &g);
cudaError_t sandwich_s1(const SandwichArgs &g);
cudaError_t product_vjp(const ProductArgs &g);      // g.list = 1 / 2
cudaError_t sandwich_rows(const SandwichArgs &g);   // g.rotors set: one rotor per row (or one device row for all)
cudaError_t sandwich_rows_vjp(const SandwichArgs &g);  // g.grad set as well: cotangents of x and of the rotor rows
cudaError_t launch_product(const ProductArgs &g);
cudaError_t launch_sandwich(const SandwichArgs &g);

#ifndef LCC_PART
#error "define LCC_PART (1-6) before including product_impl.cuh"
#endif
#ifndef LCC_PART_STRICT
#define LCC_PART_STRICT 0
#endif
#define LCC_CAT_(a, b) a##b
#define LCC_CAT(a, b) LCC_CAT_(a, b)
constexpr bool kPartStrict = LCC_PART_STRICT != 0;

#define LCC_FOR_EACH_PRODUCT_OP(X) \
    X(kOpGP) X(kOpOuter) X(kOpLC) X(kOpGPOuter) X(kOpRC) X(kOpScalar) X(kOpCommutator) X(kOpAnticommutator)

#if LCC_PART == 1
template <int LAYOUT, int OP> static cudaError_t launch_product_direct_t(const ProductArgs &g) {
    constexpr int W = 16 / (int)sizeof(real);
    int vec_ok = 1;
    if (LAYOUT == kAoS) {
        auto ok = [&](const void *p, int64_t ld) { return ((uintptr_t)p % 16 == 0) && (ld % W == 0); };
        vec_ok = ok(g.a, g.lda) && ok(g.b, g.ldb) && ok(g.out, g.ldo) && (OP != kOpGPOuter || ok(g.out2, g.ldo2));
    }
    product_kernel<LAYOUT, OP, kPartStrict><<<grid_for<LAYOUT>(g.n), block_threads(), 0, g.stream>>>(
        (const real *)g.a, g.lda, g.a_bcast, (const real *)g.b, g.ldb, g.b_bcast, (real *)g.out, g.ldo,
        (real *)g.out2, g.ldo2, g.n, vec_ok, to_coeffs(g.mu, 1.0));
    note_kernel_launch();
    return cudaGetLastError();
}
cudaError_t LCC_CAT(product_direct_s, LCC_PART_STRICT)(const ProductArgs &g) {
    switch (g.op) {
#define LCC_X(OPV) case OPV: return g.layout == kSoA ? launch_product_direct_t<kSoA, OPV>(g) : launch_product_direct_t<kAoS, OPV>(g);
        LCC_FOR_EACH_PRODUCT_OP(LCC_X)
#undef LCC_X
    }
    return cudaErrorInvalidValue;
}
#endif  // LCC_PART == 1

#if LCC_PART == 2
cudaError_t LCC_CAT(product_tma_s, LCC_PART_STRICT)(const ProductArgs &g) {
    switch (g.op) {
#define LCC_X(OPV) case OPV: return g.layout == kSoA ? launch_product_soa_tma<OPV, kPartStrict>(g) : launch_product_tma<OPV, kPartStrict>(g);
        LCC_FOR_EACH_PRODUCT_OP(LCC_X)
#undef LCC_X
    }
    return cudaErrorInvalidValue;
}
#endif  // LCC_PART == 2

#if LCC_PART == 5
// cotangents of gp / outer / left contraction: the same kernels over the VJP term lists (FMA arithmetic; gradients
// have no reference rounding order).  a = first operand of the list, b = second (lcc_abi.cu, lcc_batch_product_vjp).
template <int LAYOUT, int OP, int LIST> static cudaError_t launch_vjp_t(const ProductArgs &g) {
    if (g.n <= 0) return cudaSuccess;
    cudaError_t e = LAYOUT == kSoA ? launch_product_soa_tma<OP, false, LIST>(g) : launch_product_tma<OP, false, LIST>(g);
    if (e != cudaErrorNotSupported) return e;
    constexpr int W = 16 / (int)sizeof(real);
    int vec_ok = 1;
    if (LAYOUT == kAoS) {
        auto ok = [&](const void *p, int64_t ld) { return ((uintptr_t)p % 16 == 0) && (ld % W == 0); };
        vec_ok = ok(g.a, g.lda) && ok(g.b, g.ldb) && ok(g.out, g.ldo);
    }
    product_kernel<LAYOUT, OP, false, LIST><<<grid_for<LAYOUT>(g.n), block_threads(), 0, g.stream>>>(
        (const real *)g.a, g.lda, g.a_bcast, (const real *)g.b, g.ldb, g.b_bcast, (real *)g.out, g.ldo,
        nullptr, 0, g.n, vec_ok, to_coeffs(g.mu, 1.0));
    note_kernel_launch();
    return cudaGetLastError();
}
template <int LAYOUT, int LIST> static cudaError_t launch_vjp_op(const ProductArgs &g) {
    switch (g.op) {
        case kOpGP: return launch_vjp_t<LAYOUT, kOpGP, LIST>(g);
        case kOpOuter: return launch_vjp_t<LAYOUT, kOpOuter, LIST>(g);
        case kOpLC: return launch_vjp_t<LAYOUT, kOpLC, LIST>(g);
    }
    return cudaErrorInvalidValue;
}
cudaError_t product_vjp(const ProductArgs &g) {
    if (g.list == 1) return g.layout == kSoA ? launch_vjp_op<kSoA, 1>(g) : launch_vjp_op<kAoS, 1>(g);
    if (g.list == 2) return g.layout == kSoA ? launch_vjp_op<kSoA, 2>(g) : launch_vjp_op<kAoS, 2>(g);
    return cudaErrorInvalidValue;
}
#endif  // LCC_PART == 5

#if LCC_PART == 6
template <int LAYOUT, bool STRICT> static cudaError_t launch_sandwich_rows_t(const SandwichArgs &g) {
    if (g.n <= 0) return cudaSuccess;
    if constexpr (NB > 16) {
        return cudaErrorNotSupported;  // three 32-blade tiles do not fit the register file: the caller composes two products
    } else {
        constexpr int W = 16 / (int)sizeof(real);
        const real *r = (const real *)g.rotors;
        if constexpr (LAYOUT == kAoS && kHasTmaPath) {
            if (tma_path_enabled() && g.n >= option_aos_tma_min_rows() * kAosNarrowRowFactor) {
                CUtensorMap mr, mx, mo;
                if (aos_map(&mx, g.x, g.ldx, g.n) && aos_map(&mo, g.out, g.ldo, g.n) && (g.r_bcast || aos_map(&mr, g.rotors, g.ldr, g.n))) {
                    if (g.r_bcast) mr = mx;
                    auto kernel = sandwich_rows_kernel_aos_tma<STRICT>;
                    constexpr int smem = 2 * TmaTile::kBytes + 1024 + 16;
                    const cudaError_t attr = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
                    if (attr != cudaSuccess) return attr;
                    const unsigned grid = (unsigned)((g.n + TmaTile::kRows - 1) / TmaTile::kRows);
                    kernel<<<grid, TmaTile::kRows, g.r_bcast ? smem - TmaTile::kBytes : smem, g.stream>>>(mr, mx, mo, g.r_bcast ? r : nullptr, to_coeffs(g.mu, 1.0));
                    note_kernel_launch();
                    return cudaGetLastError();
                }
            }
        }
        int vec_ok = 1;
        if (LAYOUT == kAoS)
            vec_ok = ((uintptr_t)g.x % 16 == 0) && (g.ldx % W == 0) && ((uintptr_t)g.out % 16 == 0) && (g.ldo % W == 0) &&
                     ((uintptr_t)g.rotors % 16 == 0) && (g.r_bcast || g.ldr % W == 0);
        sandwich_rows_kernel<LAYOUT, STRICT><<<grid_for<LAYOUT>(g.n), block_threads(), 0, g.stream>>>(
            r, g.ldr, g.r_bcast ? 1 : 0, (const real *)g.x, g.ldx, (real *)g.out, g.ldo, g.n, vec_ok, to_coeffs(g.mu, 1.0));
        note_kernel_launch();
        return cudaGetLastError();
    }
}
cudaError_t sandwich_rows(const SandwichArgs &g) {
    if (g.layout == kSoA) return g.strict ? launch_sandwich_rows_t<kSoA, true>(g) : launch_sandwich_rows_t<kSoA, false>(g);
    return g.strict ? launch_sandwich_rows_t<kAoS, true>(g) : launch_sandwich_rows_t<kAoS, false>(g);
}
template <int LAYOUT> static cudaError_t launch_sandwich_rows_vjp_t(const SandwichArgs &g) {
    if (g.n <= 0) return cudaSuccess;
    if constexpr (NB > 16 || (NB > 8 && sizeof(real) == 8 && kRuntimeMetric)) {
        return cudaErrorNotSupported;  // seven tiles do not fit the register file: the caller composes products
    } else {
        constexpr int W = 16 / (int)sizeof(real);
        int vec_ok = 1;
        if (LAYOUT == kAoS) {
            auto ok = [&](const void *p, int64_t ld) { return ((uintptr_t)p % 16 == 0) && (ld % W == 0); };
            vec_ok = ok(g.x, g.ldx) && ok(g.grad, g.ldg) && ok(g.gx, g.ldgx) && ok(g.gr, g.ldgr) && ((uintptr_t)g.rotors % 16 == 0) && (g.r_bcast || g.ldr % W == 0);
        }
        const int block = 64;  // 7 tiles of registers per thread: small blocks keep several resident and out of phase
        const unsigned grid = (unsigned)((g.n + block - 1) / block);
        sandwich_rows_vjp_kernel<LAYOUT><<<grid, block, 0, g.stream>>>(
            (const real *)g.rotors, g.ldr, g.r_bcast ? 1 : 0, (const real *)g.x, g.ldx, (const real *)g.grad, g.ldg,
            (real *)g.gx, g.ldgx, (real *)g.gr, g.ldgr, g.n, vec_ok, to_coeffs(g.mu, 1.0));
        note_kernel_launch();
        return cudaGetLastError();
    }
}
cudaError_t sandwich_rows_vjp(const SandwichArgs &g) {
    return g.layout == kSoA ? launch_sandwich_rows_vjp_t<kSoA>(g) : launch_sandwich_rows_vjp_t<kAoS>(g);
}
#endif  // LCC_PART == 6

#if LCC_PART == 3
template <int LAYOUT, bool EVEN> static cudaError_t launch_sandwich_t(const SandwichArgs &g) {
    if (g.n <= 0) return cudaSuccess;
    constexpr int W = 16 / (int)sizeof(real);
    int vec_ok = 1;
    if (LAYOUT == kAoS)
        vec_ok = ((uintptr_t)g.x % 16 == 0) && (g.ldx % W == 0) && ((uintptr_t)g.out % 16 == 0) && (g.ldo % W == 0);
    Coeffs R = to_coeffs(g.rotor, 0.0), Rrev;
    Rrev.stream_hint = R.stream_hint;
    for (int i = 0; i < NB; ++i) {
        int grade = 0;
        for (int b = i; b; b >>= 1) grade += b & 1;
        // ~R[i] = R[i] * reverse_sign(grade): src/batch.rs:452-457, src/algebra/blades.rs:128-135
        Rrev.v[i] = R.v[i] * ((grade % 4 < 2) ? real(1) : real(-1));
    }
    if constexpr (LAYOUT == kAoS) {
        const cudaError_t e = launch_sandwich_tma<kPartStrict, EVEN>(g, R, Rrev);
        if (e != cudaErrorNotSupported) return e;
    } else {
        const cudaError_t e = launch_sandwich_soa_tma<kPartStrict, EVEN>(g, R, Rrev);
        if (e != cudaErrorNotSupported) return e;
    }
    sandwich_kernel<LAYOUT, kPartStrict, EVEN><<<grid_for<LAYOUT>(g.n), block_threads(), 0, g.stream>>>(
        (const real *)g.x, g.ldx, (real *)g.out, g.ldo, g.n, vec_ok, R, Rrev, to_coeffs(g.mu, 1.0));
    note_kernel_launch();
    return cudaGetLastError();
}
cudaError_t LCC_CAT(sandwich_s, LCC_PART_STRICT)(const SandwichArgs &g) {
    if (g.layout == kSoA) return g.even ? launch_sandwich_t<kSoA, true>(g) : launch_sandwich_t<kSoA, false>(g);
    return g.even ? launch_sandwich_t<kAoS, true>(g) : launch_sandwich_t<kAoS, false>(g);
}
#endif  // LCC_PART == 3

#if LCC_PART == 4
template <int LAYOUT, bool STRICT> static cudaError_t launch_product_sum_op(const ProductArgs &g) {
    switch (g.op) {
        case kOpGP: return launch_product_sum_t<LAYOUT, kOpGP, STRICT>(g);
        case kOpOuter: return launch_product_sum_t<LAYOUT, kOpOuter, STRICT>(g);
        case kOpLC: return launch_product_sum_t<LAYOUT, kOpLC, STRICT>(g);
    }
    return cudaErrorInvalidValue;
}
cudaError_t launch_product(const ProductArgs &g) {
    if (g.list != 0) return product_vjp(g);
    if (g.sum_partial) {
        if (g.layout == kSoA) return g.strict ? launch_product_sum_op<kSoA, true>(g) : launch_product_sum_op<kSoA, false>(g);
        return g.strict ? launch_product_sum_op<kAoS, true>(g) : launch_product_sum_op<kAoS, false>(g);
    }
    if (g.n <= 0) return cudaSuccess;
    const cudaError_t e = g.strict ? product_tma_s1(g) : product_tma_s0(g);
    if (e != cudaErrorNotSupported) return e;
    return g.strict ? product_direct_s1(g) : product_direct_s0(g);
}
cudaError_t launch_sandwich(const SandwichArgs &g) {
    if (g.rotors && g.grad) return sandwich_rows_vjp(g);
    if (g.rotors) return sandwich_rows(g);
    return g.strict ? sandwich_s1(g) : sandwich_s0(g);
}
#endif  // LCC_PART == 4

}  // namespace LCC_TU
}  // namespace lcc
