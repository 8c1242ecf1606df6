// matrix_rep_kernel.cu -- K8: fixed-operand geometric product and fixed-rotor sandwich in 128 / 256-blade algebras
// through the algebra's 16 x 16 real matrix representation (csrc/matrix_rep.hpp).
//
// Reference semantics: MultivectorBatch(1 x nb).geometric_product(batch), batch.geometric_product(1 x nb)
// (/root/reference/src/batch.rs:365-404, broadcast branch) and batch.sandwich(rotor) (:452-497).  The dense form of
// these maps is an nb x nb matrix (K6 / K6d: 2*nb^2 = 131 072 flops per row at 256 blades, tensor-pipe bound).  In the
// representation basis the same matrix is I_16 (x) rho(M): per row
//     stage 1   y[a][b] = sign_A x_A ;  rho(x)[j^a][j] = WHT_b(y[a][.])[j]            16 Walsh-Hadamard transforms of 16
//     stage 2   C = rho(M) rho(x)   (or rho(x) rho(M); sandwich: rho(R) rho(x) rho(~R))  16 x 16 x 16 FMAs (x2)
//     stage 3   z[c][.] = WHT_j(C[j^c][j]) ;  out_A = sign_A z[c][b]                  (the 1/16 is folded into rho(M))
// = 8192 + 2048 flops instead of 131 072: the product becomes HBM-bound (2 * nb * sizeof(T) bytes per row).
//
// Shape: persistent, one CTA of 16 warps per SM (512 threads, so each may hold 128 registers), a ring of three 64 KB
// stages (64-row tiles in f32, 32-row tiles in f64: 256-byte TMA segments per blade row).  The warps form TWO groups of
// eight; group g transforms the tiles g, g + 2, ... of its CTA IN PLACE, each behind its own named barrier, so the groups
// sit in different phases instead of sixteen warps marching through the five barriers of a tile in lockstep.  The first
// thread of a group is its TMA thread: it stores the finished stage and, one barrier into the group's next tile (by then
// the engine has read the stage), refills it with the tile three ahead.  Inside a stage, slot s holds the rows of one
// blade / matrix entry; in the transform stages a warp works on ONE slot per instruction with its 32 lanes on 32 rows
// (conflict free), and the three stages only re-index the 256 slots (blade index -> matrix entry -> blade index; the
// blade of a string is linear over F_2, so an address costs one XOR and one IMAD).  Stage 2 runs on the tensor cores with
// the fixed 16 x 16 factor as MMA A-fragments in shared memory: mma.sync m16n8k8 TF32 x 3 (hi/lo split, the same error
// budget as K6) for f32, mma.sync m8n8k4 on the FP64 tensor pipe for f64.
// History (profiles/r02_prof_k8_*): v1 did stage 2 with FFMA / DFMA against constant-bank operands --
// issue-bound at 67 % for f32, bound by lane-indexed constant loads (ADU pipe 75 % busy) for f64: 10.1 / 25.9 ms at 2^24
// rows; v2 kept the MMA fragments in registers and ptxas re-materialised them with indexed LDC inside the tile loop.
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdlib>
#include <type_traits>

#include "launch_args.h"
#include "misc_kernels.cuh"
#include "tma_tile.cuh"

namespace lcc {
namespace {

constexpr int kThreadsRep = 512;  // 16 warps: with no 17th warp every thread may hold 128 registers

template <typename T> struct RepParams {
    T p1[256];           // first factor, [o * 16 + k]:  acc[o] += p1[o][k] * x[k]
    T p2[256];           // second factor (sandwich only), applied from the right
    uint8_t blade[256];  // string (a << 4 | b) -> blade index
    int8_t sign[256];    // string -> +1 / -1, 0 when no blade maps to the string (128-blade algebras)
    // the string -> blade map is LINEAR over F_2 (the string of a product of blades is the XOR of their strings), so
    // blade(a << 4 | b) = lin_hi(a) ^ lin_lo(b): the kernel needs 8 bytes, not a table lookup per element
    uint8_t lin[8];      // image of string bit i (bits 0-3 = b, bits 4-7 = a)
};

template <typename T> __device__ __forceinline__ void wht16(T (&v)[16]) {
#pragma unroll
    for (int h = 1; h < 16; h <<= 1)
#pragma unroll
        for (int i = 0; i < 16; ++i)
            if (!(i & h)) { const T x = v[i], y = v[i | h]; v[i] = x + y; v[i | h] = x - y; }
}


// Matrix slots: entry (f, s) of a 16 x 16 matrix lives in slot f * 16 + s; inside a slot row r sits at position
// r ^ swz(f).  The XOR spreads the four k-lanes of an MMA B-fragment load and the m-lanes of a D-fragment store over
// different banks (derivation in DESIGN.md section 4, K8); in the transform stages f is warp-uniform per instruction, so
// there the swizzle is a permutation of the 32 lanes and costs nothing.
template <typename T> struct Swz;
template <> struct Swz<float> { static __device__ __forceinline__ int of(int f) { return (f & 3) << 3; } };
template <> struct Swz<double> { static __device__ __forceinline__ int of(int f) { return ((f & 1) << 3) | ((f & 2) << 1); } };

__device__ __forceinline__ float tf32_rna(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
// (not volatile: pure functions of their operands, so the compiler may interleave independent accumulation chains)
__device__ __forceinline__ void mma_tf32(float (&d)[4], const float (&a)[4], float b0, float b1) {
    asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(__float_as_uint(a[0])), "r"(__float_as_uint(a[1])), "r"(__float_as_uint(a[2])), "r"(__float_as_uint(a[3])),
          "r"(__float_as_uint(b0)), "r"(__float_as_uint(b1)));
}
__device__ __forceinline__ void mma_f64(double (&d)[2], double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d[0]), "+d"(d[1]) : "d"(a), "d"(b));
}

// The fixed 16 x 16 factor as MMA A-fragments.  They depend on the lane only, so warp 0 builds them once per kernel from
// the parameter block into shared memory ([chunk][lane], 16 bytes per lane and chunk: conflict-free LDS.128) and every
// warp re-reads the chunk it needs right before its MMAs.  (Held in registers for the whole kernel they did not survive
// the 96-register budget: ptxas re-materialised them inside the tile loop with lane-indexed constant loads, and the
// MMAs waited on those -- 40 % of all stall samples in profiles/r02_prof_k8_v2_f64.)
template <typename T> struct Factor;
template <> struct Factor<float> {  // m16n8k8, 3xTF32: chunks hi[ks=0], lo[0], hi[1], lo[1]
    static constexpr int kChunks = 4;
    const float4 *frag;
    static __device__ __forceinline__ void build(float4 *dst, const float *p, int lane) {
        const int g = lane >> 2, t = lane & 3;
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
            const float a[4] = {p[g * 16 + 8 * ks + t], p[(g + 8) * 16 + 8 * ks + t], p[g * 16 + 8 * ks + t + 4], p[(g + 8) * 16 + 8 * ks + t + 4]};
            float hi[4], lo[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) { hi[e] = tf32_rna(a[e]); lo[e] = tf32_rna(a[e] - hi[e]); }
            dst[(2 * ks) * 32 + lane] = make_float4(hi[0], hi[1], hi[2], hi[3]);
            dst[(2 * ks + 1) * 32 + lane] = make_float4(lo[0], lo[1], lo[2], lo[3]);
        }
    }
};
template <> struct Factor<double> {  // m8n8k4: chunk ks = {a[mb = 0][ks], a[mb = 1][ks]}
    static constexpr int kChunks = 4;
    const double2 *frag;
    static __device__ __forceinline__ void build(double2 *dst, const double *p, int lane) {
        const int g = lane >> 2, t = lane & 3;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) dst[ks * 32 + lane] = make_double2(p[g * 16 + 4 * ks + t], p[(8 + g) * 16 + 4 * ks + t]);
    }
};

// D = F * Z for all rows of one column (second index = w) of the matrix in `st`, in place: Z[k][w] -> D[o][w].
// n-blocks of 8 rows, kQ of them side by side so that their accumulation chains (6 dependent MMAs for f32, 4 for f64)
// overlap; every B fragment of a group of n-blocks is read before any of its D fragments is written.
template <int TR>
__device__ __forceinline__ void column_gemm(float *st, const Factor<float> &F, int w, int lane) {
    const int g = lane >> 2, t = lane & 3;
    constexpr int kQ = 4;
#pragma unroll
    for (int q0 = 0; q0 < TR / 8; q0 += kQ) {
        float d[kQ][4], b[kQ][2][2];
#pragma unroll
        for (int qq = 0; qq < kQ; ++qq) {
            const int q = q0 + qq;
#pragma unroll
            for (int e = 0; e < 4; ++e) d[qq][e] = 0.f;
#pragma unroll
            for (int ks = 0; ks < 2; ++ks)
#pragma unroll
                for (int h = 0; h < 2; ++h)  // row (8q + g) of the half it is in, XOR (t << 3)
                    b[qq][ks][h] = st[((8 * ks + 4 * h + t) * 16 + w) * TR + (q >> 2) * 32 + (8 * ((q & 3) ^ t) + g)];
        }
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
            const float4 fh = F.frag[(2 * ks) * 32 + lane], fl = F.frag[(2 * ks + 1) * 32 + lane];
            const float hi[4] = {fh.x, fh.y, fh.z, fh.w}, lo[4] = {fl.x, fl.y, fl.z, fl.w};
            float bh[kQ][2], bl[kQ][2];
#pragma unroll
            for (int qq = 0; qq < kQ; ++qq)
#pragma unroll
                for (int h = 0; h < 2; ++h) { bh[qq][h] = tf32_rna(b[qq][ks][h]); bl[qq][h] = tf32_rna(b[qq][ks][h] - bh[qq][h]); }
#pragma unroll
            for (int qq = 0; qq < kQ; ++qq) mma_tf32(d[qq], lo, bh[qq][0], bh[qq][1]);  // small terms first
#pragma unroll
            for (int qq = 0; qq < kQ; ++qq) mma_tf32(d[qq], hi, bl[qq][0], bl[qq][1]);
#pragma unroll
            for (int qq = 0; qq < kQ; ++qq) mma_tf32(d[qq], hi, bh[qq][0], bh[qq][1]);
        }
#pragma unroll
        for (int qq = 0; qq < kQ; ++qq) {
            const int q = q0 + qq;
            const int pos = (q >> 2) * 32 + 8 * ((q & 3) ^ (g & 3)) + 2 * t;  // (8q + 2t) ^ ((g & 3) << 3); rows 2t, 2t+1 stay adjacent
            *reinterpret_cast<float2 *>(st + (g * 16 + w) * TR + pos) = make_float2(d[qq][0], d[qq][1]);
            *reinterpret_cast<float2 *>(st + ((g + 8) * 16 + w) * TR + pos) = make_float2(d[qq][2], d[qq][3]);
        }
    }
}
template <int TR>
__device__ __forceinline__ void column_gemm(double *st, const Factor<double> &F, int w, int lane) {
    const int g = lane >> 2, t = lane & 3;
    constexpr int kQ = 2;
    const int sb = ((t & 1) << 3) | ((t & 2) << 1), sd = ((g & 1) << 3) | ((g & 2) << 1);
#pragma unroll
    for (int q0 = 0; q0 < TR / 8; q0 += kQ) {
        double d[kQ][2][2], b[kQ][4];
#pragma unroll
        for (int qq = 0; qq < kQ; ++qq) {
            const int q = q0 + qq;
#pragma unroll
            for (int mb = 0; mb < 2; ++mb) d[qq][mb][0] = d[qq][mb][1] = 0.0;
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) b[qq][ks] = st[((4 * ks + t) * 16 + w) * TR + (q >> 2) * 32 + (((8 * (q & 3)) + g) ^ sb)];
        }
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            const double2 f = F.frag[ks * 32 + lane];
#pragma unroll
            for (int qq = 0; qq < kQ; ++qq) {
                mma_f64(d[qq][0], f.x, b[qq][ks]);
                mma_f64(d[qq][1], f.y, b[qq][ks]);
            }
        }
#pragma unroll
        for (int qq = 0; qq < kQ; ++qq) {
            const int q = q0 + qq;
            const int pos = (q >> 2) * 32 + ((8 * (q & 3) + 2 * t) ^ sd);
#pragma unroll
            for (int mb = 0; mb < 2; ++mb)
                *reinterpret_cast<double2 *>(st + ((8 * mb + g) * 16 + w) * TR + pos) = make_double2(d[qq][mb][0], d[qq][mb][1]);
        }
    }
}

template <typename T> struct RepShape;  // rows per tile (32 per lane-row) and ring depth
template <> struct RepShape<float> { static constexpr int kRowsPerLane = 2, kStages = 3; };   // 64-row tiles: 256-byte TMA segments, 64 KB stages
template <> struct RepShape<double> { static constexpr int kRowsPerLane = 1, kStages = 3; };  // 32-row tiles: 256-byte TMA segments, 64 KB stages

// MODE 0: out = M * x      C = P1 X            matrix slots hold X / C as (row index, column index)
// MODE 1: out = x * M      C^T = P1 X^T        the same column GEMM on the TRANSPOSED slots (p1 arrives as [j][k] = rho(M)[k][j])
// MODE 2: out = R x ~R     T = P1 X, transpose pass, C^T = P2 T^T   (p1 = rho(R) / 16, p2[j][k] = rho(~R)[k][j])
template <typename T, int MODE>
__global__ void __launch_bounds__(kThreadsRep, 1)
matrix_rep_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_out,
                  const __grid_constant__ RepParams<T> P, int nb, int64_t n, int dbg) {
    constexpr int RPL = RepShape<T>::kRowsPerLane, TR = 32 * RPL, NS = RepShape<T>::kStages;
    constexpr uint32_t kStageBytes = 256u * TR * sizeof(T);
    extern __shared__ __align__(1024) uint8_t smem[];  // no static shared memory in this kernel: the window starts aligned
    uint64_t *full = (uint64_t *)(smem + (size_t)NS * kStageBytes);
    using Frag = typename std::remove_const<typename std::remove_pointer<decltype(Factor<T>::frag)>::type>::type;
    Frag *frag1 = (Frag *)(smem + (size_t)NS * kStageBytes + 64), *frag2 = frag1 + Factor<T>::kChunks * 32;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t tiles = (n + TR - 1) / TR;
    const uint32_t box_bytes = (uint32_t)nb * TR * sizeof(T);
    const int64_t mine = tiles > blockIdx.x ? (tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    auto tile_row = [&](int64_t k) { return (int)((blockIdx.x + k * gridDim.x) * TR); };
    if (threadIdx.x == 0) {
        for (int s = 0; s < NS; ++s) tma::mbar_init(&full[s], 1);
        tma::prefetch_map(&map_x);
        tma::prefetch_map(&map_out);
    }
    if (warp == 0) {
        Factor<T>::build(frag1, P.p1, lane);
        if (MODE == 2) Factor<T>::build(frag2, P.p2, lane);
    }
    __syncthreads();
    if (threadIdx.x == 0)
        for (int64_t k = 0; k < mine && k < NS; ++k) {  // fill the ring
            tma::mbar_expect_tx(&full[k], box_bytes);
            tma::load_2d(smem + (size_t)k * kStageBytes, &map_x, &full[k], tile_row(k), 0);
        }
    {
        // ---- 16 compute warps in TWO groups of 8: group g transforms the tiles g, g + 2, g + 4, ... of this CTA, each on
        // its own stage and behind its own named barrier, so the two groups sit in different phases (one in a
        // shared-memory pass while the other is in its MMAs) instead of all 16 warps marching through the five barriers
        // of a tile in lockstep.  A warp owns TWO indices (wl and wl + 8) of its group's tile; lane = row (and row + 32).
        const int grp = warp >> 3, wl = warp & 7;
        const int gtid = threadIdx.x & 255;
        auto group_sync = [&]() {
            if (grp == 0) asm volatile("bar.sync 1, 256;" ::: "memory");
            else asm volatile("bar.sync 2, 256;" ::: "memory");
        };
        // per-index constants (two indices per warp): the warp-dependent part of the blade index, sign-flip masks of the
        // 16 blades of string row w (applied to the sign bit: one XOR per element), validity mask (128-blade algebras)
        int blade_hi[2];
        uint32_t neg[2], valid[2];
        int lsw[2];  // lane ^ swz(w): with swz linear over XOR, lane ^ swz(j ^ w) = lsw ^ swz(j)
#pragma unroll
        for (int ii = 0; ii < 2; ++ii) {
            const int w = wl + 8 * ii;
            blade_hi[ii] = 0; neg[ii] = 0; valid[ii] = 0;
#pragma unroll
            for (int i = 0; i < 4; ++i) blade_hi[ii] ^= ((w >> i) & 1) ? (int)P.lin[4 + i] : 0;
#pragma unroll
            for (int b = 0; b < 16; ++b) {
                const int sg = P.sign[(w << 4) | b];
                neg[ii] |= (sg < 0 ? 1u : 0u) << b;
                valid[ii] |= (sg != 0 ? 1u : 0u) << b;
            }
            lsw[ii] = lane ^ Swz<T>::of(w);
        }
        int blade_lo[16];   // blade of string b: the same for every thread of the grid
#pragma unroll
        for (int b = 0; b < 16; ++b)
            blade_lo[b] = ((b & 1) ? (int)P.lin[0] : 0) ^ ((b & 2) ? (int)P.lin[1] : 0) ^ ((b & 4) ? (int)P.lin[2] : 0) ^ ((b & 8) ? (int)P.lin[3] : 0);
        const bool full_algebra = nb == 256;
        // matrix-slot offset of entry j of index w: normal slots X[j^w][j] or transposed X^T[j][j^w]
        auto slot_off = [&](int j, int w, int ls, bool transposed) {
            const int f = transposed ? j : (j ^ w), sec = transposed ? (j ^ w) : j;
            const int pos = transposed ? (lane ^ Swz<T>::of(j)) : (ls ^ Swz<T>::of(j));
            return (f * 16 + sec) * TR + pos;
        };
        auto flip = [](T x, bool negative) { return negative ? -x : x; };
        Factor<T> F1, F2;
        F1.frag = frag1;
        F2.frag = frag2;
        // The group's first thread is also its TMA thread: it stores the group's finished tile and, one barrier into the
        // NEXT tile (by then the engine has read the stage), refills that stage with the tile NS ahead.
        int64_t stored = -1;  // tile whose store this thread issued last (its stage is refilled once the store has been read)
        auto refill = [&]() {
            if (gtid == 0 && stored >= 0) {
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                if (stored + NS < mine) {
                    const int rs = (int)(stored % NS);
                    tma::mbar_expect_tx(&full[rs], box_bytes);
                    tma::load_2d(smem + (size_t)rs * kStageBytes, &map_x, &full[rs], tile_row(stored + NS), 0);
                }
                stored = -1;
            }
        };
        auto store_tile = [&](int64_t it, int s) {
            if (gtid == 0) {
                tma::store_2d(&map_out, smem + (size_t)s * kStageBytes, tile_row(it), 0);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                stored = it;
            }
        };
        for (int64_t it = grp; it < mine; it += 2) {
            const int s = (int)(it % NS);
            const uint32_t ph = (uint32_t)((it / NS) & 1);
            T *st = (T *)(smem + (size_t)s * kStageBytes);
            tma::mbar_wait(&full[s], ph);
            if (dbg == 1) {  // LCC_K8_DEBUG=1: tiles pass through untouched (memory pipeline alone)
                refill();
                group_sync();
                store_tile(it, s);
                continue;
            }
            T v[2][RPL][16];
            // stage 1: gather y[w][.] from the blade slots, transform, scatter to the matrix slots
#pragma unroll
            for (int ii = 0; ii < 2; ++ii)
#pragma unroll
                for (int h = 0; h < RPL; ++h) {
#pragma unroll
                    for (int b = 0; b < 16; ++b) {
                        const T x = st[(blade_hi[ii] ^ blade_lo[b]) * TR + lane + 32 * h];
                        v[ii][h][b] = flip(x, (neg[ii] >> b) & 1u);
                        if (!full_algebra && !((valid[ii] >> b) & 1u)) v[ii][h][b] = T(0);
                    }
                    wht16(v[ii][h]);
                }
            refill();      // the previous tile's store has long been read: put the next load in flight
            group_sync();  // every blade slot of the tile has been read
#pragma unroll
            for (int ii = 0; ii < 2; ++ii)
#pragma unroll
                for (int h = 0; h < RPL; ++h)
#pragma unroll
                    for (int j = 0; j < 16; ++j) st[slot_off(j, wl + 8 * ii, lsw[ii], MODE == 1) + 32 * h] = v[ii][h][j];
            group_sync();
            // stage 2 on the tensor cores: the columns with second index w, in place
            if (dbg != 2) {  // LCC_K8_DEBUG=2: transforms only
                column_gemm<TR>(st, F1, wl, lane);
                column_gemm<TR>(st, F1, wl + 8, lane);
            }
            if (MODE == 2) {
                group_sync();
                // transpose pass: T[i][w] -> T^T[w][i] (the first index is warp-uniform on both sides: conflict free)
#pragma unroll
                for (int ii = 0; ii < 2; ++ii)
#pragma unroll
                    for (int h = 0; h < RPL; ++h)
#pragma unroll
                        for (int i = 0; i < 16; ++i) v[ii][h][i] = st[(i * 16 + wl + 8 * ii) * TR + 32 * h + (lane ^ Swz<T>::of(i))];
                group_sync();
#pragma unroll
                for (int ii = 0; ii < 2; ++ii)
#pragma unroll
                    for (int h = 0; h < RPL; ++h)
#pragma unroll
                        for (int i = 0; i < 16; ++i) st[((wl + 8 * ii) * 16 + i) * TR + 32 * h + lsw[ii]] = v[ii][h][i];
                group_sync();
                column_gemm<TR>(st, F2, wl, lane);
                column_gemm<TR>(st, F2, wl + 8, lane);
            }
            group_sync();
            // stage 3: gather the 16 matrix entries of string row w, transform back, scatter to the blade slots
#pragma unroll
            for (int ii = 0; ii < 2; ++ii)
#pragma unroll
                for (int h = 0; h < RPL; ++h) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[ii][h][j] = st[slot_off(j, wl + 8 * ii, lsw[ii], MODE != 0) + 32 * h];
                    wht16(v[ii][h]);
                }
            group_sync();  // every matrix slot has been read
#pragma unroll
            for (int ii = 0; ii < 2; ++ii)
#pragma unroll
                for (int h = 0; h < RPL; ++h)
#pragma unroll
                    for (int b = 0; b < 16; ++b)
                        if (full_algebra || ((valid[ii] >> b) & 1u))
                            st[(blade_hi[ii] ^ blade_lo[b]) * TR + lane + 32 * h] = flip(v[ii][h][b], (neg[ii] >> b) & 1u);
            tma::fence_proxy_async();
            group_sync();
            store_tile(it, s);
        }
        refill();  // (waits for the read of the last store; nothing left to load)
        if (gtid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // global writes complete before the CTA exits
    }
}

template <typename T, int MODE>
cudaError_t launch_rep_t(int nb, const void *x, int64_t ldx, void *out, int64_t ldo, int64_t n, const RepParams<T> &P, cudaStream_t s) {
    constexpr int TR = 32 * RepShape<T>::kRowsPerLane, NS = RepShape<T>::kStages;
    CUtensorMap mx, mo;
    if (!make_soa_tensor_map(&mx, x, (int)sizeof(T), nb, ldx, n, TR) || !make_soa_tensor_map(&mo, out, (int)sizeof(T), nb, ldo, n, TR))
        return cudaErrorNotSupported;
    auto kernel = matrix_rep_kernel<T, MODE>;
    constexpr int smem = NS * 256 * TR * (int)sizeof(T) + 64 + 2 * Factor<T>::kChunks * 32 * 16;  // stages, barriers, two factors
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    static int sm_count = [] { int d = 0, c = 0; cudaGetDevice(&d); cudaDeviceGetAttribute(&c, cudaDevAttrMultiProcessorCount, d); return c > 0 ? c : 148; }();
    const int64_t tiles = (n + TR - 1) / TR;
    const int grid = (int)(tiles < sm_count ? tiles : sm_count);
    static const int dbg = [] { const char *e = getenv("LCC_K8_DEBUG"); return e ? atoi(e) : 0; }();
    kernel<<<grid, kThreadsRep, smem, s>>>(mx, mo, P, nb, n, dbg);
    note_kernel_launch();
    return cudaGetLastError();
}

template <typename T>
cudaError_t launch_rep(int nb, int mode, const double *p1, const double *p2, const uint8_t *blade_at, const int8_t *sign_at,
                       const ViewArg &x, const ViewArg &out, cudaStream_t s) {
    RepParams<T> P;
    for (int i = 0; i < 256; ++i) {
        P.p1[i] = (T)p1[i];
        P.p2[i] = p2 ? (T)p2[i] : T(0);
        P.blade[i] = blade_at[i];
        P.sign[i] = sign_at[i];
    }
    // linear extension of string -> blade to all 256 strings: close the valid (string, blade) pairs of the basis blades
    // under XOR; a 128-blade algebra spans only half the strings, the other half gets blade indices 128..255 (slots that
    // exist in a stage but hold no data: such strings are masked out by sign == 0 anyway)
    {
        int lin[256], have[256];
        for (int i = 0; i < 256; ++i) { lin[i] = 0; have[i] = 0; }
        have[0] = 1;
        int count = 1, next_blade = 1;
        auto add = [&](int str, int blade) {
            if (have[str]) return;
            for (int t = 0; t < 256; ++t)
                if (have[t] == 1) { lin[t ^ str] = lin[t] ^ blade; have[t ^ str] = 2; }
            for (int t = 0; t < 256; ++t) if (have[t] == 2) have[t] = 1;
            count *= 2;
        };
        for (int str = 1; str < 256; ++str)
            if (sign_at[str] != 0 && (blade_at[str] & (blade_at[str] - 1)) == 0) { add(str, blade_at[str]); if (blade_at[str] >= next_blade) next_blade = blade_at[str] << 1; }
        for (int str = 1; str < 256 && count < 256; ++str)
            if (!have[str]) { add(str, next_blade); next_blade <<= 1; }
        for (int str = 0; str < 256; ++str)
            if (sign_at[str] != 0 && lin[str] != blade_at[str]) return cudaErrorNotSupported;  // cannot happen for a homomorphism
        for (int i = 0; i < 8; ++i) P.lin[i] = (uint8_t)lin[1 << i];
    }
    switch (mode) {
        case 0: return launch_rep_t<T, 0>(nb, x.data, x.ld, out.data, out.ld, x.n, P, s);
        case 1: return launch_rep_t<T, 1>(nb, x.data, x.ld, out.data, out.ld, x.n, P, s);
        case 2: return launch_rep_t<T, 2>(nb, x.data, x.ld, out.data, out.ld, x.n, P, s);
    }
    return cudaErrorInvalidValue;
}

}  // namespace

// K8 launcher.  mode 0 / 1 / 2 = left product, right product, sandwich; p1 (and p2 for the sandwich) are row-major 16 x 16
// f64 factors already in the orientation the kernel indexes ([o][k], see RepParams) with the 1/16 of the inverse
// transform folded into p1; blade_at / sign_at map a string to its blade and sign.  x and out are blade-major (SoA) views
// of nb = 128 or 256 blades whose row count is a multiple of 16 bytes (TMA stores clip at that granularity).
cudaError_t launch_matrix_rep_product(int nb, int mode, const double *p1, const double *p2, const uint8_t *blade_at, const int8_t *sign_at,
                                      const ViewArg &x, const ViewArg &out, cudaStream_t s) {
    if (x.n <= 0) return cudaSuccess;
    if (x.layout != kSoA || out.layout != kSoA || (nb != 128 && nb != 256)) return cudaErrorNotSupported;
    return x.dtype == kF32 ? launch_rep<float>(nb, mode, p1, p2, blade_at, sign_at, x, out, s)
                           : launch_rep<double>(nb, mode, p1, p2, blade_at, sign_at, x, out, s);
}

}  // namespace lcc
