// misc_kernels.cu -- run-time-nb kernels: K3 signed permutations / masks / scale / add / sub,
// K4 norms, K5 batch sums, K7 layout conversion, column movers, the looped product for algebras
// above 32 blades, and the PGA/CGA point embeddings.  All are streaming, HBM-bound kernels: one
// read and one write per element (norms: one read, 1/nb write), coalesced along rows for SoA.
#include <cstdlib>

#include "common.cuh"
#include "launch_args.h"
#include "misc_kernels.cuh"
#include "tma_tile.cuh"

namespace lcc {

namespace {

template <class T> struct VecW { static constexpr int value = 16 / (int)sizeof(T); };

// AoS element index -> (row, blade); nb is a power of two
__device__ __forceinline__ void split_index(int64_t e, int log2nb, int64_t &r, int &k) {
    r = e >> log2nb;
    k = (int)(e & ((1 << log2nb) - 1));
}
inline int ilog2(int v) { int l = 0; while ((1 << l) < v) ++l; return l; }
inline unsigned blocks_for(int64_t items, int per_block = kThreads) { return (unsigned)((items + per_block - 1) / per_block); }

// ------------------------------------------------------------------------------------------ K3
// out[k][r] = factor[k] == 0 ? +0 : x[source[k]][r] * (factor[k] * scale)
// reverse / grade_involution multiply by +-1.0 exactly like /root/reference/src/batch.rs:823-860;
// grade(k) writes literal zeros like batch.rs:1038-1046; scale is batch.rs:957-962.
// Each thread moves kMapChunks 16-byte chunks, kThreads chunks apart, and issues their loads back to back: with a
// single chunk per thread the map kernels sat at 6.2-6.6 TB/s while `add` (two loads per thread) reached 7.1 TB/s.
constexpr int kMapChunks = 2;

template <class T>
__global__ void __launch_bounds__(kThreads)
blade_map_soa(const T *__restrict__ x, int64_t ldx, T *__restrict__ out, int64_t ldo, int64_t n, const BladeMap map, T scale) {
    constexpr int W = VecW<T>::value, U = kMapChunks;
    const int k = blockIdx.y;
    const int64_t r0 = ((int64_t)blockIdx.x * (kThreads * U) + threadIdx.x) * W;
    const int f = map.factor[k];
    const T m = (f > 0 ? scale : -scale);
    T v[U][W];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const int64_t r = r0 + (int64_t)u * kThreads * W;
#pragma unroll
        for (int l = 0; l < W; ++l) v[u][l] = T(0);
        if (f != 0 && r < n) ldg_stream(x + (int64_t)map.source[k] * ldx + r, v[u]);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const int64_t r = r0 + (int64_t)u * kThreads * W;
        if (r >= n) break;
        if (f != 0) {
#pragma unroll
            for (int l = 0; l < W; ++l) v[u][l] = mul_rn(v[u][l], m);
        }
        stg_stream(out + (int64_t)k * ldo + r, v[u]);
    }
}

template <class T>
__global__ void __launch_bounds__(kThreads)
blade_map_aos(const T *__restrict__ x, int64_t ldx, T *__restrict__ out, int64_t ldo, int64_t n, int log2nb,
              const BladeMap map, T scale) {
    const int64_t e = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    int64_t r; int k;
    split_index(e, log2nb, r, k);
    if (r >= n) return;
    const int f = map.factor[k];
    T v = T(0);
    if (f != 0) v = mul_rn(__ldg(x + r * ldx + map.source[k]), f > 0 ? scale : -scale);
    out[r * ldo + k] = v;
}

// out = a +- b with the one-row broadcast of batch.rs:863-954
template <class T>
__global__ void __launch_bounds__(kThreads)
addsub_soa(const T *__restrict__ a, int64_t lda, int a_bc, const T *__restrict__ b, int64_t ldb, int b_bc,
           T *__restrict__ out, int64_t ldo, int64_t n, int sub) {
    constexpr int W = VecW<T>::value;
    const int k = blockIdx.y;
    const int64_t r0 = ((int64_t)blockIdx.x * kThreads + threadIdx.x) * W;
    if (r0 >= n) return;
    T va[W], vb[W];
    if (a_bc) { T s = __ldg(a + (int64_t)k * lda);
#pragma unroll
        for (int l = 0; l < W; ++l) va[l] = s;
    } else ldg_stream(a + (int64_t)k * lda + r0, va);
    if (b_bc) { T s = __ldg(b + (int64_t)k * ldb);
#pragma unroll
        for (int l = 0; l < W; ++l) vb[l] = s;
    } else ldg_stream(b + (int64_t)k * ldb + r0, vb);
#pragma unroll
    for (int l = 0; l < W; ++l) va[l] = sub ? va[l] - vb[l] : va[l] + vb[l];
    stg_stream(out + (int64_t)k * ldo + r0, va);
}

template <class T>
__global__ void __launch_bounds__(kThreads)
addsub_aos(const T *__restrict__ a, int64_t lda, int a_bc, const T *__restrict__ b, int64_t ldb, int b_bc,
           T *__restrict__ out, int64_t ldo, int64_t n, int log2nb, int sub) {
    const int64_t e = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    int64_t r; int k;
    split_index(e, log2nb, r, k);
    if (r >= n) return;
    T va = __ldg(a + (a_bc ? 0 : r * lda) + k), vb = __ldg(b + (b_bc ? 0 : r * ldb) + k);
    out[r * ldo + k] = sub ? va - vb : va + vb;
}

// ------------------------------------------------------------------------------------------ K3 on AoS rows, 128-bit
// One thread per 16-byte chunk (W consecutive blades of one row); consecutive threads take consecutive chunks,
// so every warp instruction moves 512 contiguous bytes.  The per-blade tables sit in shared memory (indexing the
// by-value BladeMap with a lane-dependent blade serialises on the constant bank: the scalar kernel above runs at
// 0.95 TB/s, this one at copy speed).  DIAG: source[k] == k for every blade (reverse, involutions, grade, scale,
// even/odd), so the input chunk is one vector load too -- skipped when all its factors are zero (grade projection
// then reads only the blades it keeps).  Otherwise (duals) the W inputs are gathered from the same row.
template <class T, bool DIAG>
__global__ void __launch_bounds__(kThreads)
blade_map_aos_vec(const T *__restrict__ x, int64_t ldx, T *__restrict__ out, int64_t ldo, int64_t n, int nb, int log2cpr,
                  const BladeMap map, T scale) {
    constexpr int W = VecW<T>::value;
    __shared__ T sfac[256];
    __shared__ uint16_t ssrc[256];
    for (int k = threadIdx.x; k < nb; k += kThreads) {
        const int f = map.factor[k];
        sfac[k] = f == 0 ? T(0) : (f > 0 ? scale : -scale);
        ssrc[k] = map.source[k];
    }
    __syncthreads();
    constexpr int U = kMapChunks;
    const int64_t c0 = (int64_t)blockIdx.x * (kThreads * U) + threadIdx.x;
    const int k0 = (int)(c0 & ((1 << log2cpr) - 1)) * W;  // the same chunk of the row for every u (kThreads % cpr == 0)
    T f[W], v[U][W];
    bool any = false;
#pragma unroll
    for (int l = 0; l < W; ++l) { f[l] = sfac[k0 + l]; any |= f[l] != T(0); }
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const int64_t r = (c0 + (int64_t)u * kThreads) >> log2cpr;
        if (r >= n) continue;
        if (DIAG) {
            if (any) ldg_stream(x + r * ldx + k0, v[u]);
        } else {
#pragma unroll
            for (int l = 0; l < W; ++l) v[u][l] = f[l] != T(0) ? __ldg(x + r * ldx + ssrc[k0 + l]) : T(0);
        }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const int64_t r = (c0 + (int64_t)u * kThreads) >> log2cpr;
        if (r >= n) break;
#pragma unroll
        for (int l = 0; l < W; ++l) v[u][l] = f[l] != T(0) ? mul_rn(v[u][l], f[l]) : T(0);
        stg_stream(out + r * ldo + k0, v[u]);
    }
}

template <class T>
__global__ void __launch_bounds__(kThreads)
addsub_aos_vec(const T *__restrict__ a, int64_t lda, int a_bc, const T *__restrict__ b, int64_t ldb, int b_bc,
               T *__restrict__ out, int64_t ldo, int64_t n, int log2cpr, int sub) {
    constexpr int W = VecW<T>::value;
    const int64_t c = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    const int64_t r = c >> log2cpr;
    if (r >= n) return;
    const int k0 = (int)(c & ((1 << log2cpr) - 1)) * W;
    T va[W], vb[W];
    if (a_bc) {
#pragma unroll
        for (int l = 0; l < W; ++l) va[l] = __ldg(a + k0 + l);
    } else ldg_stream(a + r * lda + k0, va);
    if (b_bc) {
#pragma unroll
        for (int l = 0; l < W; ++l) vb[l] = __ldg(b + k0 + l);
    } else ldg_stream(b + r * ldb + k0, vb);
#pragma unroll
    for (int l = 0; l < W; ++l) va[l] = sub ? va[l] - vb[l] : va[l] + vb[l];
    stg_stream(out + r * ldo + k0, va);
}

static inline bool aos_vec_ok(const ViewArg &v, int nb, int W) {
    return nb >= W && ((uintptr_t)v.data & 15) == 0 && (v.ld % W) == 0;
}

// ------------------------------------------------------------------------------------------ K4
// norm^2[r] = sum_k w[k] * x[r][k]^2, k ascending, w[k] = sign[k][k]*reverse_sign(k): the scalar part
// of x*~x that batch.rs:685-711 reaches through its double loop (only j == i lands on blade 0).
struct Weights { int8_t w[256]; };

template <class T, bool STRICT, int LAYOUT>
__global__ void __launch_bounds__(kThreads)
norm_kernel(const T *__restrict__ x, int64_t ldx, int64_t n, int nb, const Weights wt, int mode,
            T *__restrict__ norms, T *__restrict__ out, int64_t ldo) {
    const int64_t r = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (r >= n) return;
    T acc = T(0);
    for (int k = 0; k < nb; ++k) {
        const int w = wt.w[k];
        if (w == 0) continue;
        const T v = LAYOUT == kSoA ? __ldg(x + (int64_t)k * ldx + r) : __ldg(x + r * ldx + k);
        acc = Arith<STRICT>::madd(w > 0 ? v : -v, v, acc);
    }
    if (mode == 0) { norms[r] = acc; return; }
    const T nrm = sqrt_rn(acc < T(0) ? -acc : acc);
    if (mode == 1) { norms[r] = nrm; return; }
    const bool divide = nrm > (T)1e-14;  // batch.rs:805-811
    for (int k = 0; k < nb; ++k) {
        const T v = LAYOUT == kSoA ? __ldg(x + (int64_t)k * ldx + r) : __ldg(x + r * ldx + k);
        const T o = divide ? div_rn(v, nrm) : v;
        if (LAYOUT == kSoA) out[(int64_t)k * ldo + r] = o; else out[r * ldo + k] = o;
    }
}

// AoS rows by lane groups: a row of nb blades is nb/W consecutive 16-byte chunks = cpr consecutive lanes of one
// warp (cpr <= 32), every lane loads one chunk (fully coalesced), the running sum travels through the group with
// warp shuffles in ascending blade order -- the reference's summation order, so STRICT stays bit-exact -- and for
// `normalized` each lane scales and stores the chunk it already holds: one read and one write per element.
// Every thread takes kNormRowsInFlight chunks, kThreads chunks apart, and issues all their loads before the first
// reduction: with one 16-byte load per thread the kernel is latency-bound (3.5-4.5 TB/s, profiles/r01_aux_ops_after.jsonl).
constexpr int kNormRowsInFlight = 4;
#ifndef LCC_NORMALIZED_MODE_DEFAULT
#define LCC_NORMALIZED_MODE_DEFAULT 2
#endif

template <class T, bool STRICT>
__global__ void __launch_bounds__(kThreads)
norm_aos_lanes(const T *__restrict__ x, int64_t ldx, int64_t n, int log2cpr, const Weights wt, int mode,
               T *__restrict__ norms, T *__restrict__ out, int64_t ldo) {
    constexpr int W = VecW<T>::value, U = kNormRowsInFlight;
    const int cpr = 1 << log2cpr;
    __shared__ int8_t sw[256];                 // lane-dependent blade index: shared memory, not the constant bank
    if (threadIdx.x < cpr * W) sw[threadIdx.x] = wt.w[threadIdx.x];
    __syncthreads();
    const int64_t c0 = (int64_t)blockIdx.x * (kThreads * U) + threadIdx.x;
    const int j = (int)(c0 & (cpr - 1));      // chunk inside the row == position inside the lane group (kThreads % cpr == 0)
    T v[U][W];
    int64_t r[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        r[u] = (c0 + (int64_t)u * kThreads) >> log2cpr;  // whole lane groups are live or not
#pragma unroll
        for (int l = 0; l < W; ++l) v[u][l] = T(0);
        if (r[u] < n) ldg_stream(x + r[u] * ldx + j * W, v[u]);
    }
    int w[W];
#pragma unroll
    for (int l = 0; l < W; ++l) w[l] = sw[j * W + l];
    T acc[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        acc[u] = T(0);
        if (STRICT) {
            for (int step = 0; step < cpr; ++step) {   // lane `step` of each group extends the running sum
                const T incoming = __shfl_up_sync(0xffffffffu, acc[u], 1, cpr);
                if (j == step) {
                    if (step > 0) acc[u] = incoming;
#pragma unroll
                    for (int l = 0; l < W; ++l)
                        if (w[l] != 0) acc[u] = Arith<true>::madd(w[l] > 0 ? v[u][l] : -v[u][l], v[u][l], acc[u]);
                }
            }
            acc[u] = __shfl_sync(0xffffffffu, acc[u], cpr - 1, cpr);  // the last lane holds the row's total
        } else {                                       // default mode: per-lane partial, then a butterfly over the group
#pragma unroll
            for (int l = 0; l < W; ++l)
                if (w[l] != 0) acc[u] = Arith<false>::madd(w[l] > 0 ? v[u][l] : -v[u][l], v[u][l], acc[u]);
            for (int o = cpr >> 1; o > 0; o >>= 1) acc[u] += __shfl_xor_sync(0xffffffffu, acc[u], o, cpr);
        }
    }
    if (mode != 2 && cpr >= U) {
        // Norms only: every lane of a group holds its row's total, but one square root and one store per row are
        // enough.  The U * (32 / cpr) <= 32 rows of the warp are gathered one per lane (lane = u * G + g), so the
        // warp issues ONE square-root sequence instead of U and writes runs of G consecutive norms.
        const int lane = threadIdx.x & 31, G = 32 >> log2cpr;
        const int g = lane & (G - 1), mu = lane >> (5 - log2cpr);  // G is a power of two: lane / G
        T mine = T(0);
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const T other = __shfl_sync(0xffffffffu, acc[u], g << log2cpr);
            if (mu == u) mine = other;
        }
        if (mu < U) {
            const int64_t row = (c0 - lane + ((int64_t)g << log2cpr) + (int64_t)mu * kThreads) >> log2cpr;
            if (row < n) norms[row] = mode == 0 ? mine : sqrt_rn(mine < T(0) ? -mine : mine);
        }
        return;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
        if (r[u] >= n) continue;
        if (mode == 0) { if (j == 0) norms[r[u]] = acc[u]; continue; }
        const T nrm = sqrt_rn(acc[u] < T(0) ? -acc[u] : acc[u]);
        if (mode == 1) { if (j == 0) norms[r[u]] = nrm; continue; }
        if (nrm > (T)1e-14) {                      // batch.rs:805-811
            if (STRICT) {
#pragma unroll
                for (int l = 0; l < W; ++l) v[u][l] = div_rn(v[u][l], nrm);
            } else {                               // one division per row instead of one per element (<= 1 ulp apart)
                const T inv = div_rn(T(1), nrm);
#pragma unroll
                for (int l = 0; l < W; ++l) v[u][l] = mul_rn(v[u][l], inv);
            }
        }
        stg_stream(out + r[u] * ldo + j * W, v[u]);
    }
}

// SoA `normalized` with the row held in registers (NB known at compile time): one read and one write per element.
template <class T, bool STRICT, int NB, int V>
__global__ void __launch_bounds__(kThreads)
normalized_soa_regs(const T *__restrict__ x, int64_t ldx, int64_t n, const Weights wt, T *__restrict__ out, int64_t ldo) {
    const int64_t r0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * V;
    if (r0 >= n) return;
    T v[NB][V], acc[V];
#pragma unroll
    for (int l = 0; l < V; ++l) acc[l] = T(0);
#pragma unroll
    for (int k = 0; k < NB; ++k) {
        ldg_stream(x + (int64_t)k * ldx + r0, v[k]);
        const int w = wt.w[k];
        if (w != 0) {
#pragma unroll
            for (int l = 0; l < V; ++l) acc[l] = Arith<STRICT>::madd(w > 0 ? v[k][l] : -v[k][l], v[k][l], acc[l]);
        }
    }
    T nrm[V], inv[V];
#pragma unroll
    for (int l = 0; l < V; ++l) {
        nrm[l] = sqrt_rn(acc[l] < T(0) ? -acc[l] : acc[l]);
        inv[l] = nrm[l] > (T)1e-14 ? div_rn(T(1), nrm[l]) : T(1);  // default mode: one division per row
    }
#pragma unroll
    for (int k = 0; k < NB; ++k) {
#pragma unroll
        for (int l = 0; l < V; ++l) {
            if (STRICT) v[k][l] = nrm[l] > (T)1e-14 ? div_rn(v[k][l], nrm[l]) : v[k][l];
            else if (nrm[l] > (T)1e-14) v[k][l] = mul_rn(v[k][l], inv[l]);
        }
        stg_stream(out + (int64_t)k * ldo + r0, v[k]);
    }
}

// SoA `normalized` on TMA-staged tiles: the (NB blades x kRows rows) tile arrives through cp.async.bulk.tensor, a thread
// owns V consecutive rows (one 16-byte shared-memory access per blade), scales them in place and the tile leaves through a
// TMA store.  Same arithmetic as normalized_soa_regs; the register kernel is bound by the latency of its own loads
// (long-scoreboard stalls, DRAM 70 % busy), here the copy engine keeps whole tiles in flight.
template <class T, int NB> struct NormTile {
    static constexpr int V = VecW<T>::value;
    static constexpr int kRowsRaw = 32768 / (NB * (int)sizeof(T));
    static constexpr int kRows = kRowsRaw > 1024 ? 1024 : (kRowsRaw < 32 * V ? 32 * V : kRowsRaw);
    static constexpr int kBoxRows = kRows > 256 ? 256 : kRows;   // TMA boxes are at most 256 elements wide
    static constexpr int kBoxes = kRows / kBoxRows;
    static constexpr int kBoxBytes = kBoxRows * NB * (int)sizeof(T);
    static constexpr int kBytes = kRows * NB * (int)sizeof(T);
    static constexpr int kThreadsTile = kRows / V;
    static __device__ __forceinline__ size_t offset(int k, int r) {
        return ((size_t)(r / kBoxRows) * NB * kBoxRows + (size_t)k * kBoxRows + (r % kBoxRows)) * sizeof(T);
    }
};

template <class T, bool STRICT, int NB>
__global__ void __launch_bounds__(NormTile<T, NB>::kThreadsTile)
normalized_soa_tma(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_out, const Weights wt) {
    using NT = NormTile<T, NB>;
    constexpr int V = NT::V;
    struct alignas(16) RowVec { T v[V]; };
    extern __shared__ uint8_t nt_smem_raw[];
    uint8_t *tile = (uint8_t *)(((uintptr_t)nt_smem_raw + 127) & ~(uintptr_t)127);
    uint64_t *bar = reinterpret_cast<uint64_t *>(tile + NT::kBytes);
    const int t = threadIdx.x, row0 = blockIdx.x * NT::kRows;
    if (t == 0) {
        tma::mbar_init(bar, 1);
        tma::mbar_expect_tx(bar, (uint32_t)NT::kBytes);
#pragma unroll
        for (int box = 0; box < NT::kBoxes; ++box) tma::load_2d(tile + box * NT::kBoxBytes, &map_x, bar, row0 + box * NT::kBoxRows, 0);
    }
    __syncthreads();
    tma::mbar_wait(bar, 0);
    T v[NB][V], acc[V];
#pragma unroll
    for (int l = 0; l < V; ++l) acc[l] = T(0);
#pragma unroll
    for (int k = 0; k < NB; ++k) {
        const RowVec c = *reinterpret_cast<const RowVec *>(tile + NT::offset(k, t * V));
        const int w = wt.w[k];
#pragma unroll
        for (int l = 0; l < V; ++l) {
            v[k][l] = c.v[l];
            if (w != 0) acc[l] = Arith<STRICT>::madd(w > 0 ? v[k][l] : -v[k][l], v[k][l], acc[l]);
        }
    }
    T nrm[V], inv[V];
#pragma unroll
    for (int l = 0; l < V; ++l) {
        nrm[l] = sqrt_rn(acc[l] < T(0) ? -acc[l] : acc[l]);
        inv[l] = nrm[l] > (T)1e-14 ? div_rn(T(1), nrm[l]) : T(1);
    }
#pragma unroll
    for (int k = 0; k < NB; ++k) {
        RowVec c;
#pragma unroll
        for (int l = 0; l < V; ++l) {
            T e = v[k][l];
            if (STRICT) e = nrm[l] > (T)1e-14 ? div_rn(e, nrm[l]) : e;
            else if (nrm[l] > (T)1e-14) e = mul_rn(e, inv[l]);
            c.v[l] = e;
        }
        *reinterpret_cast<RowVec *>(tile + NT::offset(k, t * V)) = c;
    }
    tma::fence_proxy_async();
    __syncthreads();
    if (t == 0) {
#pragma unroll
        for (int box = 0; box < NT::kBoxes; ++box) tma::store_2d(&map_out, tile + box * NT::kBoxBytes, row0 + box * NT::kBoxRows, 0);
        tma::store_commit_and_wait_read();
    }
}

// ------------------------------------------------------------------------------------------ K5
// Column sums with a fixed tree: thread-sequential over a strided chunk, warp shuffle, shared
// memory across warps, then a second pass over the per-block partials in block order.  f32 data
// is accumulated in f64 and rounded once at the end.
constexpr int kSumBlocksMax = 1024;       // blade-major: the grid is (blocks, blades)
constexpr int kSumBlocksMaxAos = 148 * 32;  // row-major: one block spans all blades; at 64 registers 4 blocks fit per SM, and 1024
                                           // blocks were 1.7 waves whose tail idled 13 % of the kernel (profiles/r01_aux_ops_r7.jsonl: 5.4 TB/s)

__device__ __forceinline__ double block_reduce_sum(double v) {
    __shared__ double warp_part[kThreads / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) warp_part[threadIdx.x >> 5] = v;
    __syncthreads();
    double total = 0.0;
    if (threadIdx.x == 0)
        for (int w = 0; w < kThreads / 32; ++w) total += warp_part[w];
    __syncthreads();
    return total;  // valid on thread 0
}

// One blade row per blockIdx.y.  Each thread keeps four independent 128-bit loads in flight (a single scalar load
// per trip leaves only 2048 x 4-8 bytes per SM outstanding, which is latency-bound far below the HBM rate) and
// four accumulators that are folded in a fixed order, so the result does not depend on timing.
template <class T>
__global__ void __launch_bounds__(kThreads)
sum_partial_soa(const T *__restrict__ x, int64_t ldx, int64_t n, int64_t rows_per_block, int vec_ok, double *__restrict__ partial) {
    constexpr int W = VecW<T>::value;
    const int k = blockIdx.y;
    const T *__restrict__ row = x + (int64_t)k * ldx;
    const int64_t begin = (int64_t)blockIdx.x * rows_per_block;
    const int64_t end = begin + rows_per_block < n ? begin + rows_per_block : n;
    double acc = 0.0;
    if (vec_ok) {  // begin and the row base are multiples of W elements
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
        constexpr int64_t kStride = (int64_t)kThreads * W;
        int64_t r = begin + (int64_t)threadIdx.x * W;
        for (; r + 3 * kStride + W <= end; r += 4 * kStride) {
            T v0[W], v1[W], v2[W], v3[W];
            ldg_stream(row + r, v0);
            ldg_stream(row + r + kStride, v1);
            ldg_stream(row + r + 2 * kStride, v2);
            ldg_stream(row + r + 3 * kStride, v3);
#pragma unroll
            for (int l = 0; l < W; ++l) { a0 += (double)v0[l]; a1 += (double)v1[l]; a2 += (double)v2[l]; a3 += (double)v3[l]; }
        }
        for (; r + W <= end; r += kStride) {
            T v0[W];
            ldg_stream(row + r, v0);
#pragma unroll
            for (int l = 0; l < W; ++l) a0 += (double)v0[l];
        }
        for (int64_t rr = r; rr < end && rr < r + W; ++rr) a1 += (double)__ldg(row + rr);  // the one partial vector
        acc = (a0 + a1) + (a2 + a3);
    } else {
        for (int64_t r = begin + threadIdx.x; r < end; r += kThreads) acc += (double)__ldg(row + r);
    }
    const double total = block_reduce_sum(acc);
    if (threadIdx.x == 0) partial[(int64_t)k * gridDim.x + blockIdx.x] = total;
}

template <class T>
__global__ void __launch_bounds__(kThreads)
sum_partial_aos(const T *__restrict__ x, int64_t ldx, int64_t n, int nb, int64_t rows_per_block,
                double *__restrict__ partial) {
    __shared__ double lane_part[kThreads];
    const int lanes = kThreads / nb;       // row lanes per blade (nb <= 256)
    const int k = threadIdx.x % nb, lane = threadIdx.x / nb;
    const int64_t begin = (int64_t)blockIdx.x * rows_per_block;
    const int64_t end = begin + rows_per_block < n ? begin + rows_per_block : n;
    double acc = 0.0;
    if (lane < lanes)
        for (int64_t r = begin + lane; r < end; r += lanes) acc += (double)__ldg(x + r * ldx + k);
    lane_part[threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.x < nb) {
        double total = 0.0;
        for (int l = 0; l < lanes; ++l) total += lane_part[l * nb + threadIdx.x];
        partial[(int64_t)threadIdx.x * gridDim.x + blockIdx.x] = total;
    }
}

// 128-bit form: thread (row lane, chunk j) adds W blades of every (kThreads / cpr)-th row of the block's range in
// f64, then the row lanes are folded in a fixed order.  Same fixed-tree guarantee, 16-byte coalesced loads.
template <class T>
__global__ void __launch_bounds__(kThreads)
sum_partial_aos_vec(const T *__restrict__ x, int64_t ldx, int64_t n, int nb, int log2cpr, int64_t rows_per_block,
                    double *__restrict__ partial) {
    constexpr int W = VecW<T>::value;
    __shared__ double part[kThreads * W];
    const int cpr = 1 << log2cpr, lanes = kThreads >> log2cpr;
    const int j = threadIdx.x & (cpr - 1), lane = threadIdx.x >> log2cpr;
    const int64_t begin = (int64_t)blockIdx.x * rows_per_block;
    const int64_t end = begin + rows_per_block < n ? begin + rows_per_block : n;
    double acc[W], acc1[W], acc2[W], acc3[W];
#pragma unroll
    for (int l = 0; l < W; ++l) acc[l] = acc1[l] = acc2[l] = acc3[l] = 0.0;
    int64_t r = begin + lane;
    for (; r + 3 * (int64_t)lanes < end; r += 4 * (int64_t)lanes) {  // four loads in flight per thread, folded in a fixed order
        T v0[W], v1[W], v2[W], v3[W];
        ldg_stream(x + r * ldx + j * W, v0);
        ldg_stream(x + (r + lanes) * ldx + j * W, v1);
        ldg_stream(x + (r + 2 * (int64_t)lanes) * ldx + j * W, v2);
        ldg_stream(x + (r + 3 * (int64_t)lanes) * ldx + j * W, v3);
        if constexpr (sizeof(T) == 4) {
            // f32: the four rows are added pairwise in f32 first (fixed order), then once in f64 -- a quarter of the
            // f32 -> f64 conversions, which run at a fraction of the FP32 rate and were what held this kernel at 6.0 TB/s;
            // the extra rounding is 2 f32 ulps of a 4-term partial, far inside the f32 result's own precision
#pragma unroll
            for (int l = 0; l < W; ++l) acc[l] += (double)((v0[l] + v1[l]) + (v2[l] + v3[l]));
        } else {
#pragma unroll
            for (int l = 0; l < W; ++l) { acc[l] += (double)v0[l]; acc1[l] += (double)v1[l]; acc2[l] += (double)v2[l]; acc3[l] += (double)v3[l]; }
        }
    }
    for (; r < end; r += lanes) {
        T v[W];
        ldg_stream(x + r * ldx + j * W, v);
#pragma unroll
        for (int l = 0; l < W; ++l) acc[l] += (double)v[l];
    }
#pragma unroll
    for (int l = 0; l < W; ++l) acc[l] = (acc[l] + acc1[l]) + (acc2[l] + acc3[l]);
#pragma unroll
    for (int l = 0; l < W; ++l) part[(lane * cpr + j) * W + l] = acc[l];  // [lane][blade]
    __syncthreads();
    if (threadIdx.x < nb) {
        double total = 0.0;
        for (int l = 0; l < lanes; ++l) total += part[l * nb + threadIdx.x];
        partial[(int64_t)threadIdx.x * gridDim.x + blockIdx.x] = total;
    }
}

template <class T>
__global__ void __launch_bounds__(kThreads)
sum_final(const double *__restrict__ partial, int nblocks, T *__restrict__ sums) {
    const int k = blockIdx.x;
    double acc = 0.0;
    for (int i = threadIdx.x; i < nblocks; i += kThreads) acc += partial[(int64_t)k * nblocks + i];
    const double total = block_reduce_sum(acc);
    if (threadIdx.x == 0) sums[k] = (T)total;
}

// ------------------------------------------------------------------------------------------ K7
// AoS (n, nb) <-> SoA (nb, ld) through a padded shared-memory tile of kTileRows rows x C blades:
// both the global read and the global write are contiguous runs.
constexpr int kTileRows = 128;

template <class T, bool TO_SOA>
__global__ void __launch_bounds__(kThreads)
transpose_kernel(const T *__restrict__ src, int64_t lds, T *__restrict__ dst, int64_t ldd, int64_t n, int C) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *tile = reinterpret_cast<T *>(smem_raw);  // [kTileRows][C+1]
    const int64_t r0 = (int64_t)blockIdx.x * kTileRows;
    const int c0 = blockIdx.y * C;
    const int rows = (int)((n - r0) < kTileRows ? (n - r0) : kTileRows);
    const int pitch = C + 1;
    if (TO_SOA) {
        for (int e = threadIdx.x; e < rows * C; e += kThreads) {
            const int r = e / C, c = e % C;
            tile[r * pitch + c] = __ldg(src + (r0 + r) * lds + c0 + c);
        }
        __syncthreads();
        for (int e = threadIdx.x; e < kTileRows * C; e += kThreads) {
            const int c = e / kTileRows, r = e % kTileRows;
            if (r < rows) dst[(int64_t)(c0 + c) * ldd + r0 + r] = tile[r * pitch + c];
        }
    } else {
        for (int e = threadIdx.x; e < kTileRows * C; e += kThreads) {
            const int c = e / kTileRows, r = e % kTileRows;
            if (r < rows) tile[r * pitch + c] = __ldg(src + (int64_t)(c0 + c) * lds + r0 + r);
        }
        __syncthreads();
        for (int e = threadIdx.x; e < rows * C; e += kThreads) {
            const int r = e / C, c = e % C;
            dst[(r0 + r) * ldd + c0 + c] = tile[r * pitch + c];
        }
    }
}

// TMA form of K7 for rows of 32-256 bytes: the AoS side of the tile is one contiguous span of HBM and moves as
// swizzled 2-D boxes (a thread reads / writes 16-byte chunks of its own row without bank conflicts), the SoA side is
// one (rows x blades) box whose blade segments are 1-2 KB runs; neither side issues a single per-thread global access.
// The scalar kernel above moves 4 bytes per thread and instruction and stops at 4.0-4.1 TB/s for f32 rows.
template <int RB> struct TransposeTile {
    static constexpr int kSpan = RB >= 128 ? 128 : RB;      // bytes per AoS box row == swizzle span
    static constexpr int kSub = RB / kSpan;                 // AoS boxes per tile
    static constexpr int kRows = 32768 / RB > 256 ? 256 : 32768 / RB;  // rows per tile == threads per block (TMA boxes: <= 256)
    static constexpr int kBytes = kRows * RB;               // 8-32 KB per side
    static constexpr uint32_t kSwzMask = kSpan == 128 ? 7u : (kSpan == 64 ? 3u : 1u);
    static __device__ __forceinline__ uint32_t chunk_offset(int t, int j) {  // 16-byte chunk j of tile row t (tile base 1024-aligned)
        const uint32_t byte = (uint32_t)j * 16u;
        const uint32_t off = (byte / kSpan) * (uint32_t)(kRows * kSpan) + (uint32_t)t * kSpan + (byte % kSpan);
        return off ^ (((off >> 7) & kSwzMask) << 4);
    }
};

template <class T, int NB, bool TO_SOA>
__global__ void __launch_bounds__(TransposeTile<NB * (int)sizeof(T)>::kRows)
transpose_tma_kernel(const __grid_constant__ CUtensorMap map_aos, const __grid_constant__ CUtensorMap map_soa) {
    using TT = TransposeTile<NB * (int)sizeof(T)>;
    constexpr int W = VecW<T>::value;
    struct alignas(16) Chunk { T v[W]; };
    extern __shared__ uint8_t tr_smem_raw[];
    uint8_t *tile_in = (uint8_t *)(((uintptr_t)tr_smem_raw + 1023) & ~(uintptr_t)1023), *tile_out = tile_in + TT::kBytes;
    uint64_t *bar = reinterpret_cast<uint64_t *>(tile_out + TT::kBytes);
    const int t = threadIdx.x, row0 = blockIdx.x * TT::kRows;
    if (t == 0) {
        tma::mbar_init(bar, 1);
        tma::mbar_expect_tx(bar, (uint32_t)TT::kBytes);
        if (TO_SOA) {
#pragma unroll
            for (int sub = 0; sub < TT::kSub; ++sub)
                tma::load_2d(tile_in + sub * TT::kRows * TT::kSpan, &map_aos, bar, sub * (TT::kSpan / (int)sizeof(T)), row0);
        } else {
            tma::load_2d(tile_in, &map_soa, bar, row0, 0);
        }
    }
    __syncthreads();
    tma::mbar_wait(bar, 0);
    if (TO_SOA) {
        T *soa = reinterpret_cast<T *>(tile_out);  // [NB][kRows]
#pragma unroll
        for (int j = 0; j < NB / W; ++j) {
            const Chunk c = *reinterpret_cast<const Chunk *>(tile_in + TT::chunk_offset(t, j));
#pragma unroll
            for (int l = 0; l < W; ++l) soa[(j * W + l) * TT::kRows + t] = c.v[l];
        }
    } else {
        const T *soa = reinterpret_cast<const T *>(tile_in);
#pragma unroll
        for (int j = 0; j < NB / W; ++j) {
            Chunk c;
#pragma unroll
            for (int l = 0; l < W; ++l) c.v[l] = soa[(j * W + l) * TT::kRows + t];
            *reinterpret_cast<Chunk *>(tile_out + TT::chunk_offset(t, j)) = c;
        }
    }
    tma::fence_proxy_async();
    __syncthreads();
    if (t == 0) {  // rows beyond n are clipped by the engine (and arrived as zeros)
        if (TO_SOA) {
            tma::store_2d(&map_soa, tile_out, row0, 0);
        } else {
#pragma unroll
            for (int sub = 0; sub < TT::kSub; ++sub)
                tma::store_2d(&map_aos, tile_out + sub * TT::kRows * TT::kSpan, sub * (TT::kSpan / (int)sizeof(T)), row0);
        }
        tma::store_commit_and_wait_read();
    }
}

// K4 on TMA-staged AoS tiles (rows of 32-256 bytes, from 4096 rows on): the tile of kRows consecutive rows is one
// contiguous span of HBM, a thread owns ONE row -- it reads it from the swizzled tile in 16-byte chunks, adds the weighted
// squares in ascending blade order (the reference's order, batch.rs:685-711, so the strict mode is bit-exact without any
// lane chain), takes one square root, and for `normalized` scales the row in place before the tile leaves through a TMA
// store.  No shuffles and one square root per row: the lane-group kernel above spends 71 % of its issue slots on them.
template <class T, int NB, bool STRICT>
__global__ void __launch_bounds__(TransposeTile<NB * (int)sizeof(T)>::kRows)
norm_aos_tma(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_out, const Weights wt, int mode,
             T *__restrict__ norms, int64_t n) {
    using TT = TransposeTile<NB * (int)sizeof(T)>;
    constexpr int W = VecW<T>::value;
    struct alignas(16) Chunk { T v[W]; };
    extern __shared__ uint8_t na_smem_raw[];
    uint8_t *tile = (uint8_t *)(((uintptr_t)na_smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t *bar = reinterpret_cast<uint64_t *>(tile + TT::kBytes);
    const int t = threadIdx.x, row0 = blockIdx.x * TT::kRows;
    if (t == 0) {
        tma::mbar_init(bar, 1);
        tma::mbar_expect_tx(bar, (uint32_t)TT::kBytes);
#pragma unroll
        for (int sub = 0; sub < TT::kSub; ++sub)
            tma::load_2d(tile + sub * TT::kRows * TT::kSpan, &map_x, bar, sub * (TT::kSpan / (int)sizeof(T)), row0);
    }
    __syncthreads();
    tma::mbar_wait(bar, 0);
    T v[NB];
#pragma unroll
    for (int j = 0; j < NB / W; ++j) {
        const Chunk c = *reinterpret_cast<const Chunk *>(tile + TT::chunk_offset(t, j));
#pragma unroll
        for (int l = 0; l < W; ++l) v[j * W + l] = c.v[l];
    }
    T acc = T(0);
#pragma unroll
    for (int k = 0; k < NB; ++k) {
        const int w = wt.w[k];
        if (w != 0) acc = Arith<STRICT>::madd(w > 0 ? v[k] : -v[k], v[k], acc);
    }
    const int64_t row = (int64_t)row0 + t;
    if (mode == 0) { if (row < n) norms[row] = acc; return; }
    const T nrm = sqrt_rn(acc < T(0) ? -acc : acc);
    if (mode == 1) { if (row < n) norms[row] = nrm; return; }
    if (nrm > (T)1e-14) {  // batch.rs:805-811
        const T inv = STRICT ? T(1) : div_rn(T(1), nrm);
#pragma unroll
        for (int k = 0; k < NB; ++k) v[k] = STRICT ? div_rn(v[k], nrm) : mul_rn(v[k], inv);
    }
#pragma unroll
    for (int j = 0; j < NB / W; ++j) {
        Chunk c;
#pragma unroll
        for (int l = 0; l < W; ++l) c.v[l] = v[j * W + l];
        *reinterpret_cast<Chunk *>(tile + TT::chunk_offset(t, j)) = c;
    }
    tma::fence_proxy_async();
    __syncthreads();
    if (t == 0) {
#pragma unroll
        for (int sub = 0; sub < TT::kSub; ++sub)
            tma::store_2d(&map_out, tile + sub * TT::kRows * TT::kSpan, sub * (TT::kSpan / (int)sizeof(T)), row0);
        tma::store_commit_and_wait_read();
    }
}

// ------------------------------------------------------------------------------------------ column movers
struct ColumnMap { int16_t col_of_blade[256]; };  // -1: blade stays zero

template <class T, int LAYOUT>
__global__ void __launch_bounds__(kThreads)
scatter_columns_kernel(const T *__restrict__ cols, int ncols, T *__restrict__ dst, int64_t ldd, int64_t n, int log2nb,
                       const ColumnMap map) {
    int64_t r; int k;
    if (LAYOUT == kSoA) { r = (int64_t)blockIdx.x * kThreads + threadIdx.x; k = blockIdx.y; }
    else split_index((int64_t)blockIdx.x * kThreads + threadIdx.x, log2nb, r, k);
    if (r >= n) return;
    const int c = map.col_of_blade[k];
    const T v = c >= 0 ? __ldg(cols + r * ncols + c) : T(0);
    if (LAYOUT == kSoA) dst[(int64_t)k * ldd + r] = v; else dst[r * ldd + k] = v;
}

struct BladeList { int32_t blade_of_col[256]; };

template <class T, int LAYOUT>
__global__ void __launch_bounds__(kThreads)
gather_columns_kernel(const T *__restrict__ src, int64_t lds, int64_t n, int ncols, const BladeList list,
                      T *__restrict__ cols) {
    const int64_t e = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    const int64_t r = e / ncols;
    const int c = (int)(e % ncols);
    if (r >= n) return;
    const int k = list.blade_of_col[c];
    cols[e] = LAYOUT == kSoA ? __ldg(src + (int64_t)k * lds + r) : __ldg(src + r * lds + k);
}

// ------------------------------------------------------------------------------------------ table-driven product
// out[k] = sum_i G[k][i] * x[i] * y[i^k], i ascending, with an arbitrary sign table G (int8, on the
// device).  Serves (1) the products of algebras above 32 blades, where G[k][i] = sign(i, i^k) keeps
// the reference's ascending-left-blade order per output blade (batch.rs:384-404), and (2) the
// vector-Jacobian products behind lcc.torch autograd for every algebra (tables in lcc_abi.cu).
// One block owns 32 rows; both operand tiles sit in shared memory as [blade][row]; warp w produces
// output blades w, w+8, ...  Not a bandwidth-bound kernel: 2*nb^2 shared loads per row.
template <class T> __device__ __forceinline__ T view_load(const T *p, int64_t ld, int layout, int64_t r, int k) {
    return layout == kSoA ? __ldg(p + (int64_t)k * ld + r) : __ldg(p + r * ld + k);
}

template <class T, bool STRICT>
__global__ void __launch_bounds__(kThreads)
table_product_kernel(const T *__restrict__ x, int64_t ldx, int x_bc, const T *__restrict__ y, int64_t ldy, int y_bc,
                     T *__restrict__ out, int64_t ldo, int64_t n, int nb, int layout, const int8_t *__restrict__ gsign) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *sx = reinterpret_cast<T *>(smem_raw);
    T *sy = sx + (size_t)nb * 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t r = (int64_t)blockIdx.x * 32 + lane;
    const bool live = r < n;
    for (int i = warp; i < nb; i += kThreads / 32) {
        sx[i * 32 + lane] = live ? view_load(x, ldx, layout, x_bc ? 0 : r, i) : T(0);
        sy[i * 32 + lane] = live ? view_load(y, ldy, layout, y_bc ? 0 : r, i) : T(0);
    }
    __syncthreads();
    for (int k = warp; k < nb; k += kThreads / 32) {
        T acc = T(0);
        const int8_t *g = gsign + (size_t)k * nb;
        for (int i = 0; i < nb; ++i) {
            const int s = g[i];
            if (s == 0) continue;
            const T xv = sx[i * 32 + lane];
            acc = Arith<STRICT>::madd(s > 0 ? xv : -xv, sy[(i ^ k) * 32 + lane], acc);
        }
        if (live) {
            if (layout == kSoA) out[(int64_t)k * ldo + r] = acc; else out[r * ldo + k] = acc;
        }
    }
}

// ------------------------------------------------------------------------------------------ PGA / CGA points
// which 0: PGA3D point  c[7]=1, c[14]=-x, c[13]=y, c[11]=-z        (src/pga.rs:131-134)
// which 1: CGA3D point  c[1,2,4]=x,y,z, c[8]=0.5|x|^2-0.5, c[16]=0.5|x|^2+0.5   (src/cga.rs:189-206)
// which 2: STA bivector from (E, B), 6 columns: c[3,5,9]=E, c[12]=Bx, c[10]=-By, c[6]=Bz   (src/sta.rs:202-212)
// which 3: STA boost rotor from a velocity, 3 columns: c[0]=cosh(phi/2), c[3,5,9]=-sinh(phi/2)*v/|v|,
//          phi=atanh|v|; |v|<1e-14 gives the identity; |v|>=1 has no rotor (the scalar API raises ValueError,
//          the batched form writes NaN for that row)   (src/sta.rs:253-296)
template <class T, int LAYOUT>
__global__ void __launch_bounds__(kThreads)
points_embed_kernel(int which, const T *__restrict__ cols, T *__restrict__ dst, int64_t ldd, int64_t n) {
    const int64_t r = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (r >= n) return;
    const int ncols = which == 2 ? 6 : 3;
    T c[6];
    for (int i = 0; i < ncols; ++i) c[i] = __ldg(cols + r * ncols + i);
    const T x = c[0], y = c[1], z = c[2];
    const int nb = which == 1 ? 32 : 16;
    T nsq = T(0);
    if (which == 1 || which == 3) {  // x*x + y*y + z*z, each operation rounded on its own (no FMA in the reference)
        const T xx = mul_rn(x, x), yy = mul_rn(y, y), zz = mul_rn(z, z);
        nsq = Arith<true>::madd(T(1), zz, Arith<true>::madd(T(1), yy, xx));
    }
    T b0 = T(1), bx = T(0), by = T(0), bz = T(0);
    if (which == 3) {
        const T vmag = sqrt_rn(nsq);
        if (vmag >= T(1)) {
            b0 = bx = by = bz = (T)__int_as_float(0x7fc00000);
        } else if (!(vmag < (T)1e-14)) {
            const T half_phi = div_rn(atanh(vmag), T(2));
            const T ch = cosh(half_phi), sh = sinh(half_phi);
            b0 = ch;
            bx = mul_rn(-sh, div_rn(x, vmag)); by = mul_rn(-sh, div_rn(y, vmag)); bz = mul_rn(-sh, div_rn(z, vmag));
        }
    }
    for (int k = 0; k < nb; ++k) {
        T v = T(0);
        if (which == 0) {
            if (k == 7) v = T(1); else if (k == 14) v = -x; else if (k == 13) v = y; else if (k == 11) v = -z;
        } else if (which == 1) {
            if (k == 1) v = x; else if (k == 2) v = y; else if (k == 4) v = z;
            else if (k == 8) v = mul_rn(T(0.5), nsq) - T(0.5);
            else if (k == 16) v = mul_rn(T(0.5), nsq) + T(0.5);
        } else if (which == 2) {
            if (k == 3) v = c[0]; else if (k == 5) v = c[1]; else if (k == 9) v = c[2];
            else if (k == 12) v = c[3]; else if (k == 10) v = -c[4]; else if (k == 6) v = c[5];
        } else {
            if (k == 0) v = b0; else if (k == 3) v = bx; else if (k == 5) v = by; else if (k == 9) v = bz;
        }
        if (LAYOUT == kSoA) dst[(int64_t)k * ldd + r] = v; else dst[r * ldd + k] = v;
    }
}

// which 0: x=-c[14]/w, y=c[13]/w, z=-c[11]/w, w=c[7]   (src/pga.rs:579-599)
// which 1: (c[1],c[2],c[4]) / (c[16]-c[8])             (src/cga.rs:492-508)
// Rows whose weight is below the reference's threshold (1e-10 / 1e-12) have no finite point: the
// scalar API raises ValueError there; the batched form writes NaN for that row.
template <class T, int LAYOUT>
__global__ void __launch_bounds__(kThreads)
points_extract_kernel(int which, const T *__restrict__ src, int64_t lds, int64_t n, T *__restrict__ xyz) {
    const int64_t r = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (r >= n) return;
    auto at = [&](int k) { return LAYOUT == kSoA ? __ldg(src + (int64_t)k * lds + r) : __ldg(src + r * lds + k); };
    T w, x, y, z, thr;
    if (which == 0) { w = at(7); x = -at(14); y = at(13); z = -at(11); thr = (T)1e-10; }
    else { w = at(16) - at(8); x = at(1); y = at(2); z = at(4); thr = (T)1e-12; }
    const bool ok = (w < T(0) ? -w : w) >= thr;
    const T nanv = (T)__int_as_float(0x7fc00000);
    xyz[r * 3 + 0] = ok ? div_rn(x, w) : nanv;
    xyz[r * 3 + 1] = ok ? div_rn(y, w) : nanv;
    xyz[r * 3 + 2] = ok ? div_rn(z, w) : nanv;
}

template <class F32, class F64> cudaError_t by_dtype(int dtype, F32 f32, F64 f64) {
    if (dtype == kF32) f32(); else if (dtype == kF64) f64(); else return cudaErrorInvalidValue;
    note_kernel_launch();
    return cudaGetLastError();
}

}  // namespace

// ============================================================================================ launchers

cudaError_t launch_blade_map(int nb, const ViewArg &x, const ViewArg &out, const BladeMap &map, double scale, cudaStream_t s) {
    const int64_t n = out.n;
    if (n <= 0) return cudaSuccess;
    auto run = [&](auto tag) {
        using T = decltype(tag);
        if (x.layout == kSoA) {
            constexpr int W = VecW<T>::value;
            dim3 grid(blocks_for((n + W - 1) / W, kThreads * kMapChunks), nb);
            blade_map_soa<T><<<grid, kThreads, 0, s>>>((const T *)x.data, x.ld, (T *)out.data, out.ld, n, map, (T)scale);
        } else if (aos_vec_ok(x, nb, VecW<T>::value) && aos_vec_ok(out, nb, VecW<T>::value)) {
            constexpr int W = VecW<T>::value;
            const int cpr = nb / W;
            bool diag = true;
            for (int k = 0; k < nb; ++k) diag = diag && (map.factor[k] == 0 || map.source[k] == k);
            const unsigned grid = blocks_for(n * cpr, kThreads * kMapChunks);
            if (diag) blade_map_aos_vec<T, true><<<grid, kThreads, 0, s>>>((const T *)x.data, x.ld, (T *)out.data, out.ld, n, nb, ilog2(cpr), map, (T)scale);
            else blade_map_aos_vec<T, false><<<grid, kThreads, 0, s>>>((const T *)x.data, x.ld, (T *)out.data, out.ld, n, nb, ilog2(cpr), map, (T)scale);
        } else {
            blade_map_aos<T><<<blocks_for(n * nb), kThreads, 0, s>>>((const T *)x.data, x.ld, (T *)out.data, out.ld, n, ilog2(nb), map, (T)scale);
        }
    };
    return by_dtype(x.dtype, [&] { run(float()); }, [&] { run(double()); });
}

cudaError_t launch_addsub(int nb, int op, const ViewArg &a, bool a_bc, const ViewArg &b, bool b_bc, const ViewArg &out, cudaStream_t s) {
    const int64_t n = out.n;
    if (n <= 0) return cudaSuccess;
    auto run = [&](auto tag) {
        using T = decltype(tag);
        if (out.layout == kSoA) {
            constexpr int W = VecW<T>::value;
            dim3 grid(blocks_for((n + W - 1) / W), nb);
            addsub_soa<T><<<grid, kThreads, 0, s>>>((const T *)a.data, a.ld, a_bc, (const T *)b.data, b.ld, b_bc, (T *)out.data, out.ld, n, op);
        } else if (aos_vec_ok(a, nb, VecW<T>::value) && aos_vec_ok(b, nb, VecW<T>::value) && aos_vec_ok(out, nb, VecW<T>::value)) {
            const int cpr = nb / VecW<T>::value;
            addsub_aos_vec<T><<<blocks_for(n * cpr), kThreads, 0, s>>>((const T *)a.data, a.ld, a_bc, (const T *)b.data, b.ld, b_bc, (T *)out.data, out.ld, n, ilog2(cpr), op);
        } else {
            addsub_aos<T><<<blocks_for(n * nb), kThreads, 0, s>>>((const T *)a.data, a.ld, a_bc, (const T *)b.data, b.ld, b_bc, (T *)out.data, out.ld, n, ilog2(nb), op);
        }
    };
    return by_dtype(out.dtype, [&] { run(float()); }, [&] { run(double()); });
}

template <class T, int NB> static bool launch_norm_aos_tma_nb(int mode, bool strict, const ViewArg &x, const Weights &wt, void *norms, const ViewArg &out, cudaStream_t s) {
    using TT = TransposeTile<NB * (int)sizeof(T)>;
    CUtensorMap mx, mo;
    if (!make_aos_tensor_map(&mx, x.data, (int)sizeof(T), NB, x.ld, x.n, TT::kSpan, TT::kRows)) return false;
    mo = mx;
    if (mode == 2 && !make_aos_tensor_map(&mo, out.data, (int)sizeof(T), NB, out.ld, x.n, TT::kSpan, TT::kRows)) return false;
    constexpr int smem = TT::kBytes + 1024 + 16;
    const unsigned grid = (unsigned)((x.n + TT::kRows - 1) / TT::kRows);
    auto go = [&](auto kernel) {
        if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return false;
        kernel<<<grid, TT::kRows, smem, s>>>(mx, mo, wt, mode, (T *)norms, x.n);
        return true;
    };
    return strict ? go(norm_aos_tma<T, NB, true>) : go(norm_aos_tma<T, NB, false>);
}
// false: the view does not qualify (row size, alignment, batch size) -- the lane-group / scalar kernels take it
template <class T> static bool launch_norm_aos_tma(int nb, int mode, bool strict, const ViewArg &x, const Weights &wt, void *norms, const ViewArg &out, cudaStream_t s) {
    static const bool on = [] { const char *e = std::getenv("LCC_NORM_AOS_TMA"); return !(e && e[0] == '0'); }();
    if (!on || x.n < 4096) return false;
    switch (nb * (int)sizeof(T)) {
        case 32: return launch_norm_aos_tma_nb<T, 32 / (int)sizeof(T)>(mode, strict, x, wt, norms, out, s);
        case 64: return launch_norm_aos_tma_nb<T, 64 / (int)sizeof(T)>(mode, strict, x, wt, norms, out, s);
        case 128: return launch_norm_aos_tma_nb<T, 128 / (int)sizeof(T)>(mode, strict, x, wt, norms, out, s);
        case 256: return launch_norm_aos_tma_nb<T, 256 / (int)sizeof(T)>(mode, strict, x, wt, norms, out, s);
    }
    return false;
}

// SoA `normalized` (LCC_NORMALIZED_MODE): 2 = TMA tiles from 65536 rows on, thin register kernel below (default);
// 1 = register kernel with 8-byte accesses; 0 = with 16-byte accesses.  2^26 PGA f32 / 2^24 CGA f64 / 2^26 R3 f64 rows:
// mode 0 5891 / 6633 / 5736 GB/s, mode 1 6015 / 6638 / 6758, mode 2 6838 / 6794 / 6901 (profiles/r01_normalized_modes.txt).
static int normalized_mode() {
    static const int m = [] { const char *e = std::getenv("LCC_NORMALIZED_MODE"); return e ? std::atoi(e) : LCC_NORMALIZED_MODE_DEFAULT; }();
    return m;
}
template <class T, int NB> static bool launch_normalized_tma(const ViewArg &x, const ViewArg &out, const Weights &wt, bool strict, cudaStream_t s) {
    using NT = NormTile<T, NB>;
    CUtensorMap mx, mo;
    if (!make_soa_tensor_map(&mx, x.data, (int)sizeof(T), NB, x.ld, x.n, NT::kBoxRows)) return false;
    if (!make_soa_tensor_map(&mo, out.data, (int)sizeof(T), NB, out.ld, x.n, NT::kBoxRows)) return false;
    constexpr int smem = NT::kBytes + 128 + 16;
    const unsigned grid = (unsigned)((x.n + NT::kRows - 1) / NT::kRows);
    auto go = [&](auto kernel) {
        if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return false;
        kernel<<<grid, NT::kThreadsTile, smem, s>>>(mx, mo, wt);
        return true;
    };
    return strict ? go(normalized_soa_tma<T, true, NB>) : go(normalized_soa_tma<T, false, NB>);
}

cudaError_t launch_norm(int nb, int mode, bool strict, const ViewArg &x, const int8_t *weights, void *norms, const ViewArg &out, cudaStream_t s) {
    const int64_t n = x.n;
    if (n <= 0) return cudaSuccess;
    Weights wt;
    for (int k = 0; k < 256; ++k) wt.w[k] = k < nb ? weights[k] : 0;
    auto run = [&](auto tag) {
        using T = decltype(tag);
        const unsigned g = blocks_for(n);
        constexpr int W = VecW<T>::value;
        if (x.layout == kAoS && launch_norm_aos_tma<T>(nb, mode, strict, x, wt, norms, out, s)) return;
        // strict norms (no output row to write) stay on the row-per-thread kernel: the bit-exact lane chain is
        // latency-bound and only pays off when it saves the second pass of `normalized`
        if (x.layout == kAoS && nb / W <= 32 && aos_vec_ok(x, nb, W) && (mode != 2 || aos_vec_ok(out, nb, W)) && !(strict && mode != 2)) {
            const int cpr = nb / W;
            const unsigned gl = blocks_for(n * cpr, kThreads * kNormRowsInFlight);
            if (strict) norm_aos_lanes<T, true><<<gl, kThreads, 0, s>>>((const T *)x.data, x.ld, n, ilog2(cpr), wt, mode, (T *)norms, (T *)out.data, out.ld);
            else norm_aos_lanes<T, false><<<gl, kThreads, 0, s>>>((const T *)x.data, x.ld, n, ilog2(cpr), wt, mode, (T *)norms, (T *)out.data, out.ld);
            return;
        }
        if (x.layout == kSoA && mode == 2 && (nb == 4 || nb == 8 || nb == 16 || nb == 32) && ((uintptr_t)x.data & 15) == 0 &&
            ((uintptr_t)out.data & 15) == 0 && x.ld % W == 0 && out.ld % W == 0) {
#define LCC_NORMALIZED_CASE(NBV)                                                                                                  \
    case NBV: {                                                                                                                   \
        if (normalized_mode() == 2 && n >= 65536 && launch_normalized_tma<T, NBV>(x, out, wt, strict, s)) return;               \
        /* 64-thread blocks: the row lives in 64-130 registers, so 256-thread blocks fit twice per SM and move through  \
         * load / compute / store in lock step (5.3 TB/s); small blocks interleave their phases, as in K1 / K2 */         \
        constexpr int kBlock = 64;                                                                                                \
        constexpr int VW = NBV >= 32 ? 1 : (NBV >= 16 && sizeof(T) == 8 ? 1 : W);                                                 \
        constexpr int V8 = 8 / (int)sizeof(T) < VW ? 8 / (int)sizeof(T) : VW;  /* mode 1: 8-byte accesses, thinner threads */   \
        if (normalized_mode() >= 1 && n % V8 == 0) {                                                                              \
            const unsigned gr = blocks_for((n + V8 - 1) / V8, kBlock);                                                            \
            if (strict) normalized_soa_regs<T, true, NBV, V8><<<gr, kBlock, 0, s>>>((const T *)x.data, x.ld, n, wt, (T *)out.data, out.ld); \
            else normalized_soa_regs<T, false, NBV, V8><<<gr, kBlock, 0, s>>>((const T *)x.data, x.ld, n, wt, (T *)out.data, out.ld);   \
            return;                                                                                                               \
        }                                                                                                                         \
        constexpr int V = VW;                                                                                                     \
        const unsigned gr = blocks_for((n + V - 1) / V, kBlock);                                                                  \
        if (n % V != 0) break; /* padded rows are addressable (ld >= n rounded up), but keep the tail on the generic kernel */  \
        if (strict) normalized_soa_regs<T, true, NBV, V><<<gr, kBlock, 0, s>>>((const T *)x.data, x.ld, n, wt, (T *)out.data, out.ld); \
        else normalized_soa_regs<T, false, NBV, V><<<gr, kBlock, 0, s>>>((const T *)x.data, x.ld, n, wt, (T *)out.data, out.ld);   \
        return;                                                                                                                   \
    }
            switch (nb) {
                LCC_NORMALIZED_CASE(4)
                LCC_NORMALIZED_CASE(8)
                LCC_NORMALIZED_CASE(16)
                LCC_NORMALIZED_CASE(32)
            }
#undef LCC_NORMALIZED_CASE
        }
#define LCC_NORM_LAUNCH(STRICT, LAYOUT) \
    norm_kernel<T, STRICT, LAYOUT><<<g, kThreads, 0, s>>>((const T *)x.data, x.ld, n, nb, wt, mode, (T *)norms, (T *)out.data, out.ld)
        if (x.layout == kSoA) { if (strict) LCC_NORM_LAUNCH(true, kSoA); else LCC_NORM_LAUNCH(false, kSoA); }
        else { if (strict) LCC_NORM_LAUNCH(true, kAoS); else LCC_NORM_LAUNCH(false, kAoS); }
#undef LCC_NORM_LAUNCH
    };
    return by_dtype(x.dtype, [&] { run(float()); }, [&] { run(double()); });
}

static int sum_blocks_cap_aos() {  // LCC_SUM_AOS_BLOCKS: experiment knob for the row-major grid
    static const int cap = [] { const char *e = std::getenv("LCC_SUM_AOS_BLOCKS"); const int v = e ? std::atoi(e) : 0; return v > 0 ? v : kSumBlocksMaxAos; }();
    return cap;
}
static int sum_blocks(int64_t n, int layout) {
    int64_t b = (n + (int64_t)kThreads * 16 - 1) / ((int64_t)kThreads * 16);
    const int64_t cap = layout == kSoA ? kSumBlocksMax : sum_blocks_cap_aos();
    if (b < 1) b = 1;
    if (b > cap) b = cap;
    return (int)b;
}
int64_t sum_scratch_elems(int nb, int64_t n) { return (int64_t)nb * sum_blocks(n, kAoS); }  // the larger of the two grids

cudaError_t launch_sum(int nb, const ViewArg &x, void *sums_dev, void *scratch_dev, int64_t scratch_elems, cudaStream_t s) {
    const int64_t n = x.n;
    const int nblocks = sum_blocks(n, x.layout);
    if (scratch_elems < (int64_t)nb * nblocks) return cudaErrorInvalidValue;
    int64_t rows_per_block = (n + nblocks - 1) / nblocks;
    rows_per_block = (rows_per_block + 3) & ~(int64_t)3;  // block ranges start on 16-byte boundaries
    auto run = [&](auto tag) {
        using T = decltype(tag);
        if (x.layout == kSoA) {
            const int vec_ok = ((uintptr_t)x.data & 15) == 0 && (x.ld % VecW<T>::value) == 0;
            sum_partial_soa<T><<<dim3(nblocks, nb), kThreads, 0, s>>>((const T *)x.data, x.ld, n, rows_per_block, vec_ok, (double *)scratch_dev);
        } else if (aos_vec_ok(x, nb, VecW<T>::value)) {
            const int cpr = nb / VecW<T>::value;
            // Whole waves only: at 34 (f32) / 42 (f64) registers 6 / 5 blocks are resident per SM, so the full grid of
            // 148 * 32 blocks was 5.3 / 6.4 waves and the last, partly filled wave left half of the SMs idle for a
            // block's lifetime (5.98-6.4 TB/s where the sibling kernels stream at 6.9+).  The largest multiple of
            // (SMs x resident blocks) below the cap keeps every wave full; the partial sums stay a fixed tree.
            static const int wave = [] {
                int dev = 0, sms = 148, occ = 0;
                cudaGetDevice(&dev);
                cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
                if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, sum_partial_aos_vec<T>, kThreads, 0) != cudaSuccess || occ < 1) occ = 4;
                return sms * occ;
            }();
            int grid = nblocks;
            if (nblocks == sum_blocks_cap_aos() && wave <= nblocks) {
                grid = nblocks / wave * wave;
                rows_per_block = ((n + grid - 1) / grid + 3) & ~(int64_t)3;
            }
            sum_partial_aos_vec<T><<<grid, kThreads, 0, s>>>((const T *)x.data, x.ld, n, nb, ilog2(cpr), rows_per_block, (double *)scratch_dev);
            note_kernel_launch();
            sum_final<T><<<nb, kThreads, 0, s>>>((const double *)scratch_dev, grid, (T *)sums_dev);
            return;
        } else {
            sum_partial_aos<T><<<nblocks, kThreads, 0, s>>>((const T *)x.data, x.ld, n, nb, rows_per_block, (double *)scratch_dev);
        }
        note_kernel_launch();
        sum_final<T><<<nb, kThreads, 0, s>>>((const double *)scratch_dev, nblocks, (T *)sums_dev);
    };
    return by_dtype(x.dtype, [&] { run(float()); }, [&] { run(double()); });
}

cudaError_t launch_sum_final(int nb, int nblocks, const double *partial, void *sums_dev, int dtype, cudaStream_t s) {
    return by_dtype(dtype, [&] { sum_final<float><<<nb, kThreads, 0, s>>>(partial, nblocks, (float *)sums_dev); },
                    [&] { sum_final<double><<<nb, kThreads, 0, s>>>(partial, nblocks, (double *)sums_dev); });
}

// cudaErrorNotSupported: this pair of views does not qualify (row size, alignment, batch too small) -- scalar kernel
template <class T, int NB> static cudaError_t launch_convert_tma_nb(const ViewArg &src, const ViewArg &dst, cudaStream_t s) {
    using TT = TransposeTile<NB * (int)sizeof(T)>;
    const ViewArg &aos = src.layout == kAoS ? src : dst, &soa = src.layout == kAoS ? dst : src;
    const int64_t n = src.n;
    CUtensorMap ma, ms;
    if (!make_aos_tensor_map(&ma, aos.data, (int)sizeof(T), NB, aos.ld, n, TT::kSpan, TT::kRows)) return cudaErrorNotSupported;
    if (!make_soa_tensor_map(&ms, soa.data, (int)sizeof(T), NB, soa.ld, n, TT::kRows)) return cudaErrorNotSupported;
    constexpr int smem = 2 * TT::kBytes + 1024 + 16;
    const unsigned grid = (unsigned)((n + TT::kRows - 1) / TT::kRows);
    auto go = [&](auto kernel) {
        const cudaError_t attr = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (attr != cudaSuccess) return attr;
        kernel<<<grid, TT::kRows, smem, s>>>(ma, ms);
        note_kernel_launch();
        return cudaGetLastError();
    };
    return dst.layout == kSoA ? go(transpose_tma_kernel<T, NB, true>) : go(transpose_tma_kernel<T, NB, false>);
}
template <class T> static cudaError_t launch_convert_tma(int nb, const ViewArg &src, const ViewArg &dst, cudaStream_t s) {
    static const bool on = [] { const char *e = std::getenv("LCC_TMA_TRANSPOSE"); return !(e && e[0] == '0'); }();
    if (!on || src.n < 4096) return cudaErrorNotSupported;  // small batches: launch-bound either way
    switch (nb * (int)sizeof(T)) {
        case 32: return launch_convert_tma_nb<T, 32 / (int)sizeof(T)>(src, dst, s);
        case 64: return launch_convert_tma_nb<T, 64 / (int)sizeof(T)>(src, dst, s);
        case 128: return launch_convert_tma_nb<T, 128 / (int)sizeof(T)>(src, dst, s);
        case 256: return launch_convert_tma_nb<T, 256 / (int)sizeof(T)>(src, dst, s);
    }
    return cudaErrorNotSupported;
}

cudaError_t launch_convert(int nb, const ViewArg &src, const ViewArg &dst, cudaStream_t s) {
    const int64_t n = src.n;
    if (n <= 0) return cudaSuccess;
    const size_t es = src.dtype == kF32 ? 4 : 8;
    if (src.layout == dst.layout) {  // same layout, possibly different leading dimensions: strided copy
        if (src.layout == kSoA) {
            // cudaMemcpy2DAsync rejects pitches above cudaDeviceProp::memPitch (2^31 - 1): a 2^28-row f64 batch has a
            // 2 GiB pitch, so such batches are copied one blade row at a time (at most 256 plain copies)
            if ((uint64_t)dst.ld * es <= 0x7fffffffull && (uint64_t)src.ld * es <= 0x7fffffffull)
                return cudaMemcpy2DAsync(dst.data, dst.ld * es, src.data, src.ld * es, n * es, nb, cudaMemcpyDeviceToDevice, s);
            for (int k = 0; k < nb; ++k) {
                cudaError_t e = cudaMemcpyAsync((char *)dst.data + (size_t)k * dst.ld * es, (const char *)src.data + (size_t)k * src.ld * es,
                                                (size_t)n * es, cudaMemcpyDeviceToDevice, s);
                if (e != cudaSuccess) return e;
            }
            return cudaSuccess;
        }
        return cudaMemcpy2DAsync(dst.data, dst.ld * es, src.data, src.ld * es, nb * es, n, cudaMemcpyDeviceToDevice, s);
    }
    // TMA stores clip at 16-byte granularity along the contiguous dimension, so the tiles take the rows up to the last
    // multiple of 16 bytes of a blade row and the scalar kernel below the 1-3 that remain (a blade-major view may be a
    // row slice of a larger batch: the element after its last row belongs to somebody else).
    const int64_t main_rows = n - n % (int64_t)(16 / es);
    int64_t done = 0;
    if (main_rows > 0) {
        ViewArg ms = src, md = dst;
        ms.n = md.n = main_rows;
        const cudaError_t e = src.dtype == kF32 ? launch_convert_tma<float>(nb, ms, md, s) : launch_convert_tma<double>(nb, ms, md, s);
        if (e == cudaSuccess) done = main_rows;
        else if (e != cudaErrorNotSupported) return e;
    }
    if (done == n) return cudaSuccess;
    auto advanced = [&](const ViewArg &v) {
        ViewArg r = v;
        r.data = (char *)v.data + (size_t)done * es * (size_t)(v.layout == kAoS ? v.ld : 1);
        r.n = n - done;
        return r;
    };
    const ViewArg rs = advanced(src), rd = advanced(dst);
    const int C = nb < 32 ? nb : 32;
    dim3 grid(blocks_for(rs.n, kTileRows), nb / C);
    const size_t smem = (size_t)kTileRows * (C + 1) * es;
    auto run = [&](auto tag) {
        using T = decltype(tag);
        if (dst.layout == kSoA) transpose_kernel<T, true><<<grid, kThreads, smem, s>>>((const T *)rs.data, rs.ld, (T *)rd.data, rd.ld, rs.n, C);
        else transpose_kernel<T, false><<<grid, kThreads, smem, s>>>((const T *)rs.data, rs.ld, (T *)rd.data, rd.ld, rs.n, C);
    };
    return by_dtype(src.dtype, [&] { run(float()); }, [&] { run(double()); });
}

cudaError_t launch_scatter_columns(int nb, const void *cols, int ncols, const int32_t *blade_of_col, const ViewArg &dst, cudaStream_t s) {
    const int64_t n = dst.n;
    if (n <= 0) return cudaSuccess;
    ColumnMap map;
    for (int k = 0; k < 256; ++k) map.col_of_blade[k] = -1;
    for (int c = 0; c < ncols; ++c) map.col_of_blade[blade_of_col[c]] = (int16_t)c;
    auto run = [&](auto tag) {
        using T = decltype(tag);
        if (dst.layout == kSoA) scatter_columns_kernel<T, kSoA><<<dim3(blocks_for(n), nb), kThreads, 0, s>>>((const T *)cols, ncols, (T *)dst.data, dst.ld, n, 0, map);
        else scatter_columns_kernel<T, kAoS><<<blocks_for(n * nb), kThreads, 0, s>>>((const T *)cols, ncols, (T *)dst.data, dst.ld, n, ilog2(nb), map);
    };
    return by_dtype(dst.dtype, [&] { run(float()); }, [&] { run(double()); });
}

cudaError_t launch_gather_columns(int nb, const ViewArg &src, int ncols, const int32_t *blade_of_col, void *cols, cudaStream_t s) {
    (void)nb;
    const int64_t n = src.n;
    if (n <= 0 || ncols <= 0) return cudaSuccess;
    BladeList list;
    for (int c = 0; c < 256; ++c) list.blade_of_col[c] = c < ncols ? blade_of_col[c] : 0;
    auto run = [&](auto tag) {
        using T = decltype(tag);
        if (src.layout == kSoA) gather_columns_kernel<T, kSoA><<<blocks_for(n * ncols), kThreads, 0, s>>>((const T *)src.data, src.ld, n, ncols, list, (T *)cols);
        else gather_columns_kernel<T, kAoS><<<blocks_for(n * ncols), kThreads, 0, s>>>((const T *)src.data, src.ld, n, ncols, list, (T *)cols);
    };
    return by_dtype(src.dtype, [&] { run(float()); }, [&] { run(double()); });
}

cudaError_t launch_copy_rows(int nb, const ViewArg &src, int64_t start, int64_t end, const ViewArg &dst, int64_t dst_start, cudaStream_t s) {
    const size_t es = src.dtype == kF32 ? 4 : 8;
    ViewArg a = src, b = dst;
    a.n = b.n = end - start;
    a.data = (char *)src.data + (src.layout == kSoA ? start : start * src.ld) * es;
    b.data = (char *)dst.data + (dst.layout == kSoA ? dst_start : dst_start * dst.ld) * es;
    return launch_convert(nb, a, b, s);
}

cudaError_t launch_table_product(int nb, bool strict, const ViewArg &x, bool x_bc, const ViewArg &y, bool y_bc, const ViewArg &out, const int8_t *gsign, cudaStream_t s) {
    const int64_t n = out.n;
    if (n <= 0) return cudaSuccess;
    const size_t es = out.dtype == kF32 ? 4 : 8;
    const size_t smem = (size_t)2 * nb * 32 * es;
    auto run = [&](auto tag) {
        using T = decltype(tag);
        auto go = [&](auto kernel) {
            cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            kernel<<<blocks_for(n, 32), kThreads, smem, s>>>((const T *)x.data, x.ld, x_bc, (const T *)y.data, y.ld, y_bc, (T *)out.data, out.ld, n, nb, out.layout, gsign);
        };
        if (strict) go(table_product_kernel<T, true>); else go(table_product_kernel<T, false>);
    };
    return by_dtype(out.dtype, [&] { run(float()); }, [&] { run(double()); });
}

cudaError_t launch_points_embed(int which, const void *xyz, const ViewArg &dst, cudaStream_t s) {
    const int64_t n = dst.n;
    if (n <= 0) return cudaSuccess;
    auto run = [&](auto tag) {
        using T = decltype(tag);
        if (dst.layout == kSoA) points_embed_kernel<T, kSoA><<<blocks_for(n), kThreads, 0, s>>>(which, (const T *)xyz, (T *)dst.data, dst.ld, n);
        else points_embed_kernel<T, kAoS><<<blocks_for(n), kThreads, 0, s>>>(which, (const T *)xyz, (T *)dst.data, dst.ld, n);
    };
    return by_dtype(dst.dtype, [&] { run(float()); }, [&] { run(double()); });
}

cudaError_t launch_points_extract(int which, const ViewArg &src, void *xyz, cudaStream_t s) {
    const int64_t n = src.n;
    if (n <= 0) return cudaSuccess;
    auto run = [&](auto tag) {
        using T = decltype(tag);
        if (src.layout == kSoA) points_extract_kernel<T, kSoA><<<blocks_for(n), kThreads, 0, s>>>(which, (const T *)src.data, src.ld, n, (T *)xyz);
        else points_extract_kernel<T, kAoS><<<blocks_for(n), kThreads, 0, s>>>(which, (const T *)src.data, src.ld, n, (T *)xyz);
    };
    return by_dtype(src.dtype, [&] { run(float()); }, [&] { run(double()); });
}

// ------------------------------------------------------------------------------------------ TMA tensor maps
// cuTensorMapEncodeTiled is reached through the runtime's driver entry point query, so the library
// does not link libcuda directly.
namespace {
using EncodeFn = CUresult (*)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                              const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                              CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeFn tensor_map_encoder() {
    static EncodeFn enc = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
        return (EncodeFn)p;
    }();
    return enc;
}
}  // namespace

bool make_soa_tensor_map(CUtensorMap *map, const void *base, int elem_bytes, int nb, int64_t ld, int64_t n, int rows, bool swizzle128, int l2_promotion) {
    EncodeFn enc = tensor_map_encoder();
    if (!enc || n <= 0 || n >= (int64_t)1 << 31) return false;
    if (((uintptr_t)base & 15) != 0 || ((ld * elem_bytes) & 15) != 0 || ld < n) return false;
    // stores are clipped at 16-byte granularity along the row dimension: with n * elem_bytes not a multiple of 16 the
    // engine would write the element after the last row (zero), which a row-slice view does not own
    if (((n * elem_bytes) & 15) != 0) return false;
    const cuuint64_t dims[2] = {(cuuint64_t)n, (cuuint64_t)nb};
    const cuuint64_t strides[1] = {(cuuint64_t)ld * (cuuint64_t)elem_bytes};
    const cuuint32_t box[2] = {(cuuint32_t)rows, (cuuint32_t)nb};
    const cuuint32_t estr[2] = {1, 1};
    const CUtensorMapDataType dt = elem_bytes == 8 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
    if (swizzle128 && rows * elem_bytes != 128) return false;  // the swizzle span is one 128-byte box row
    return enc(map, dt, 2, const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
               l2_promotion >= 256 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : (l2_promotion >= 128 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : (l2_promotion >= 64 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B : CU_TENSOR_MAP_L2_PROMOTION_NONE)),
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

bool make_aos_tensor_map(CUtensorMap *map, const void *base, int elem_bytes, int nb, int64_t ld, int64_t n,
                         int span_bytes, int rows) {
    EncodeFn enc = tensor_map_encoder();
    if (!enc || n <= 0 || n >= (int64_t)1 << 31) return false;
    if (((uintptr_t)base & 15) != 0 || ((ld * elem_bytes) & 15) != 0 || ld < nb) return false;
    const cuuint64_t dims[2] = {(cuuint64_t)nb, (cuuint64_t)n};
    const cuuint64_t strides[1] = {(cuuint64_t)ld * (cuuint64_t)elem_bytes};
    const cuuint32_t box[2] = {(cuuint32_t)(span_bytes / elem_bytes), (cuuint32_t)rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUtensorMapSwizzle swz = span_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                   : span_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
    const CUtensorMapDataType dt = elem_bytes == 8 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
    return enc(map, dt, 2, const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
               CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

bool make_aos_chunked_tensor_map(CUtensorMap *map, const void *base, int elem_bytes, int nb, int64_t ld, int64_t n, int rows) {
    EncodeFn enc = tensor_map_encoder();
    if (!enc || n <= 0 || n >= (int64_t)1 << 31 || rows > 256) return false;
    if (((uintptr_t)base & 15) != 0 || ((ld * elem_bytes) & 15) != 0 || ld < nb || (nb * elem_bytes) % 128 != 0) return false;
    const cuuint32_t per_line = (cuuint32_t)(128 / elem_bytes), chunks = (cuuint32_t)(nb * elem_bytes / 128);
    const cuuint64_t dims[3] = {per_line, (cuuint64_t)n, chunks};
    const cuuint64_t strides[2] = {(cuuint64_t)ld * (cuuint64_t)elem_bytes, 128};
    const cuuint32_t box[3] = {per_line, (cuuint32_t)rows, chunks};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUtensorMapDataType dt = elem_bytes == 8 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
    return enc(map, dt, 3, const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
               CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace lcc
