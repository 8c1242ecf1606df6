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
// v7 layout (matrix_rep.hpp, RepLayout): a unit = 128 bytes of rows (32 f32 / 16 f64) x 256 slots = 32 KB, delivered by
// TMA with the hardware 128-byte swizzle, slot = blade.  Matrix entry (f, s) lives IN the slot of the coefficient with
// string (a = f ^ s, b = s) -- nothing is ever scattered to another place:
//   * stage 1 / 3: a thread owns one 8-byte element (a row pair in f32, a row in f64) of the 16 slots of ONE string row a
//     and transforms them in place, in registers: no barrier inside the stage, 32 live data registers, packed f32x2
//     arithmetic (FADD2) in f32; signs and the validity mask (128-blade algebras) are one AND-XOR per word from a table
//     in shared memory;
//   * stage 2: mma.sync on the slots of one column (left factor) or one row (right factor) -- the slot of index k is
//     Lambda(k) ^ Gamma(w), both linear, so a fragment address is two XORs; the sandwich is a column pass followed by a
//     row pass, no transpose pass in between;
//   * 3 barriers per unit (4 for the sandwich), each over one group of 8 warps; 2 or 3 groups per CTA work on different
//     units of a ring of six, so the groups sit in different stages while TMA refills behind them.
// History (profiles/r02_prof_k8_*): v1 FFMA / DFMA against constant-bank operands (issue-bound / ADU-bound), v2-v5 stage 2
// on the tensor cores, v6 two 8-warp groups: 9.2 ms f32 / 18.5 ms f64 at 2^24 rows -- ncu showed 2.9 (f32) / 7.2 (f64)
// warps per issue stalled on long_scoreboard: the per-thread slot tables had spilled, and with 200 KB of the SM carved out
// as shared memory every reload missed L1 (31 M local loads per 2^22 rows); plus five barriers and ~2300 instructions per
// thread and tile.  v7 has no per-thread tables (offsets are constant-bank operands), ~1000 instructions, no spills.
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <type_traits>
#include <vector>

#include "../matrix_rep.hpp"
#include "launch_args.h"
#include "misc_kernels.cuh"
#include "tma_tile.cuh"

namespace lcc {
namespace {

constexpr int kGroupThreads = 256;   // 8 warps: 16 half-warps = the 16 string rows of a unit
constexpr int kUnitBytes = 32768;    // 256 slots x 128 bytes
constexpr int kRing = 6;             // units in flight per CTA (192 KB)

template <typename T> struct RepParams {
    T factor[2][256];     // per pass: the fixed 16 x 16 factor in MMA "A" orientation, PHYSICAL order [m * 16 + k]
    uint32_t f_lo[16];    // F(blade of string b)        F(s) = s << 7 | (s & 7) << 4: byte offset of slot s, swizzle folded in
    uint32_t f_hi[16];    // F(blade of string a << 4)
    uint32_t fk[2][16];   // per pass: F(Lambda(logical index at physical MMA k position i))
    uint32_t cw[2][16];   // per pass: F(Gamma(w))
    uint32_t fo[2][16];   // per pass: F(Lambda(logical index at physical MMA m position i)) (blade-major units: == fk)
    uint32_t amap[16];    // reference-layout units: physical string-row index -> string row a (f_hi is indexed physically there)
    uint32_t neg[8];      // bit s: string s carries sign -1
    uint32_t valid[8];    // bit s: a blade maps to string s
};

// signs and validity of the 16 strings of one string row, as bits (a thread works on one string row for the whole kernel)
struct RowMask { uint32_t neg, valid; };

// ---- packed arithmetic: an 8-byte element is a pair of f32 rows or one f64 row
template <typename T> struct Elem;
template <> struct Elem<float> {
    static __device__ __forceinline__ uint64_t add(uint64_t a, uint64_t b) { uint64_t d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
    static __device__ __forceinline__ uint64_t sub(uint64_t a, uint64_t b) { uint64_t d; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
    // sign (bit b of row.neg) / validity (bit b of row.valid, 128-blade algebras only): both words are values
    template <bool HALF> static __device__ __forceinline__ uint64_t mask(uint64_t v, RowMask row, int b) {
        const uint32_t m = (row.neg << (31 - b)) & 0x80000000u;
        uint32_t lo = (uint32_t)v ^ m, hi = (uint32_t)(v >> 32) ^ m;
        if (HALF && !((row.valid >> b) & 1u)) lo = hi = 0u;
        return ((uint64_t)hi << 32) | lo;
    }
};
template <> struct Elem<double> {
    static __device__ __forceinline__ uint64_t add(uint64_t a, uint64_t b) { return (uint64_t)__double_as_longlong(__longlong_as_double((long long)a) + __longlong_as_double((long long)b)); }
    static __device__ __forceinline__ uint64_t sub(uint64_t a, uint64_t b) { return (uint64_t)__double_as_longlong(__longlong_as_double((long long)a) - __longlong_as_double((long long)b)); }
    template <bool HALF> static __device__ __forceinline__ uint64_t mask(uint64_t v, RowMask row, int b) {  // the sign bit sits in the high word
        uint32_t lo = (uint32_t)v, hi = (uint32_t)(v >> 32) ^ ((row.neg << (31 - b)) & 0x80000000u);
        if (HALF && !((row.valid >> b) & 1u)) lo = hi = 0u;
        return ((uint64_t)hi << 32) | lo;
    }
};

template <typename T> __device__ __forceinline__ void wht16(uint64_t (&v)[16]) {
#pragma unroll
    for (int h = 1; h < 16; h <<= 1)
#pragma unroll
        for (int i = 0; i < 16; ++i)
            if (!(i & h)) { const uint64_t x = v[i], y = v[i | h]; v[i] = Elem<T>::add(x, y); v[i | h] = Elem<T>::sub(x, y); }
}

// Stage 1 (INPUT) / stage 3: the 16 slots of string row a at this thread's element, in place.  t0 = F(lin_hi(a)) ^ byte
// offset of the element in its line; the slot of string (a, b) is at unit + (t0 ^ f_lo[b]) -- f_lo[b] is a constant-bank
// operand, so an address is one XOR.  tbl_a = sign / validity bits of the row's 16 strings.
template <typename T, bool INPUT, bool HALF>
__device__ __forceinline__ void transform_row(uint8_t *unit, uint32_t t0, RowMask tbl_a, const RepParams<T> &P) {
    uint64_t v[16];
    asm volatile("" : "+r"(tbl_a.neg), "+r"(tbl_a.valid), "+r"(t0));  // keeps the 16 per-string masks and offsets from being hoisted out of the unit loop (registers)
#pragma unroll
    for (int b = 0; b < 16; ++b) {
        v[b] = *reinterpret_cast<const uint64_t *>(unit + (t0 ^ P.f_lo[b]));
        if (INPUT) v[b] = Elem<T>::template mask<HALF>(v[b], tbl_a, b);
    }
    wht16<T>(v);
#pragma unroll
    for (int b = 0; b < 16; ++b) {
        if (!INPUT) v[b] = Elem<T>::template mask<HALF>(v[b], tbl_a, b);
        *reinterpret_cast<uint64_t *>(unit + (t0 ^ P.f_lo[b])) = v[b];
    }
}

__device__ __forceinline__ float tf32_rna(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
// x = hi + lo with hi = x rounded to tf32 (nearest, ties away) by integer arithmetic: cvt.rna.tf32.f32 is a five-instruction
// sequence on sm_100a (it tests for Inf / NaN), and the MMA ignores the 13 low mantissa bits of its operands anyway, so lo
// needs no rounding of its own (it is truncated at 2^-10 |lo| <= 2^-21 |x|).
__device__ __forceinline__ void split_tf32(float x, float &hi, float &lo) {
    hi = __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
    lo = x - hi;
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

// The fixed 16 x 16 factor as MMA A-fragments ([chunk][lane], 16 bytes per lane and chunk: conflict-free LDS.128), built
// once per kernel from the parameter block (already in physical m / k order) and re-read right before the MMAs.
template <typename T> struct Factor;
template <> struct Factor<float> {  // m16n8k8, 3xTF32: chunks hi[ks=0], lo[0], hi[1], lo[1]
    using Frag = float4;
    static __device__ __forceinline__ void build(Frag *dst, const float *p, int lane) {
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
    using Frag = double2;
    static __device__ __forceinline__ void build(Frag *dst, const double *p, int lane) {
        const int g = lane >> 2, t = lane & 3;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) dst[ks * 32 + lane] = make_double2(p[g * 16 + 4 * ks + t], p[(8 + g) * 16 + 4 * ks + t]);
    }
};

// Stage 2, one pass: D[o] = sum_k A[o][k] Z[k] on the 16 slots {Lambda(k) ^ Gamma(w)} of index w, all rows, in place, for
// w = wl and wl + 8.  tabs = fk[16] then cw[16] of the pass.  Row bytes r of slot offset F sit at F ^ r.  Every B fragment
// of a group of n-blocks is read before any of its D fragments is written (the MMAs are warp-synchronous in between).
__device__ __forceinline__ void gemm_pass(uint8_t *unit, const float4 *frag, const uint32_t *tabs, int wl, int lane, float) {
    const int g = lane >> 2, t = lane & 3;
    uint32_t lk[2][2], lo[2];
#pragma unroll
    for (int ks = 0; ks < 2; ++ks)
#pragma unroll
        for (int h = 0; h < 2; ++h) lk[ks][h] = tabs[8 * ks + 4 * h + t] ^ (uint32_t)(g << 2);  // B[k = 8ks + 4h + t][n = g]: row 8q + g
#pragma unroll
    for (int mb = 0; mb < 2; ++mb) lo[mb] = tabs[g + 8 * mb] ^ (uint32_t)(t << 3);              // D[m = g + 8mb][n = 2t, 2t + 1]: rows 8q + 2t, + 1
#pragma unroll 1
    for (int wi = 0; wi < 2; ++wi) {
        const uint32_t cw = tabs[16 + wl + 8 * wi];
        float d[4][4], b[4][2][2];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
#pragma unroll
            for (int e = 0; e < 4; ++e) d[q][e] = 0.f;
#pragma unroll
            for (int ks = 0; ks < 2; ++ks)
#pragma unroll
                for (int h = 0; h < 2; ++h) b[q][ks][h] = *reinterpret_cast<const float *>(unit + (lk[ks][h] ^ cw ^ (uint32_t)(q << 5)));
        }
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
            const float4 fh = frag[(2 * ks) * 32 + lane], fl = frag[(2 * ks + 1) * 32 + lane];
            const float hi[4] = {fh.x, fh.y, fh.z, fh.w}, lw[4] = {fl.x, fl.y, fl.z, fl.w};
            float bh[4][2], bl[4][2];
#pragma unroll
            for (int q = 0; q < 4; ++q)
#pragma unroll
                for (int h = 0; h < 2; ++h) split_tf32(b[q][ks][h], bh[q][h], bl[q][h]);
#pragma unroll
            for (int q = 0; q < 4; ++q) mma_tf32(d[q], lw, bh[q][0], bh[q][1]);  // small terms first
#pragma unroll
            for (int q = 0; q < 4; ++q) mma_tf32(d[q], hi, bl[q][0], bl[q][1]);
#pragma unroll
            for (int q = 0; q < 4; ++q) mma_tf32(d[q], hi, bh[q][0], bh[q][1]);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int mb = 0; mb < 2; ++mb)
                *reinterpret_cast<float2 *>(unit + (lo[mb] ^ cw ^ (uint32_t)(q << 5))) = make_float2(d[q][2 * mb], d[q][2 * mb + 1]);
    }
}
__device__ __forceinline__ void gemm_pass(uint8_t *unit, const double2 *frag, const uint32_t *tabs, int wl, int lane, double) {
    const int g = lane >> 2, t = lane & 3;
    uint32_t lk[4], lo[2];
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) lk[ks] = tabs[4 * ks + t] ^ (uint32_t)(g << 3);              // B[k = 4ks + t][n = g]: row 8q + g
#pragma unroll
    for (int mb = 0; mb < 2; ++mb) lo[mb] = tabs[g + 8 * mb] ^ (uint32_t)(t << 4);              // D[m = g + 8mb][n = 2t, 2t + 1]
    const uint32_t cw[2] = {tabs[16 + wl], tabs[16 + wl + 8]};
    double d[2][2][2][2], b[2][2][4];  // [wi][q]...: four independent chains of four DMMAs per m-block
#pragma unroll
    for (int wi = 0; wi < 2; ++wi)
#pragma unroll
        for (int q = 0; q < 2; ++q) {
#pragma unroll
            for (int mb = 0; mb < 2; ++mb) d[wi][q][mb][0] = d[wi][q][mb][1] = 0.0;
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) b[wi][q][ks] = *reinterpret_cast<const double *>(unit + (lk[ks] ^ cw[wi] ^ (uint32_t)(q << 6)));
        }
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
        const double2 f = frag[ks * 32 + lane];
#pragma unroll
        for (int wi = 0; wi < 2; ++wi)
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                mma_f64(d[wi][q][0], f.x, b[wi][q][ks]);
                mma_f64(d[wi][q][1], f.y, b[wi][q][ks]);
            }
    }
#pragma unroll
    for (int wi = 0; wi < 2; ++wi)
#pragma unroll
        for (int q = 0; q < 2; ++q)
#pragma unroll
            for (int mb = 0; mb < 2; ++mb)
                *reinterpret_cast<double2 *>(unit + (lo[mb] ^ cw[wi] ^ (uint32_t)(q << 6))) = make_double2(d[wi][q][mb][0], d[wi][q][mb][1]);
}

// ---- reference-layout (AoS) units: [chunk][row][128 bytes swizzled], element (row r, blade A) at G(A) ^ R(r) with
// R(r) = r << 7 | (r & 7) << 4 (matrix_rep.hpp, RepLayoutAos).  A warp's lanes are 8 rows x 4 indices.
__device__ __forceinline__ uint32_t aos_row_term(int r) { return ((uint32_t)r << 7) | (((uint32_t)r & 7u) << 4); }

// Stage 1 / 3: the 16 slots of string row a at row r -- f32: packed with row r + 16 (same a, same masks; R(r + 16) =
// R(r) + 2048 for r < 16), f64: one row.  t0 = G(lin_hi(a)) ^ R(r).
template <typename T, bool INPUT, bool HALF>
__device__ __forceinline__ void transform_row_aos(uint8_t *unit, uint32_t t0, RowMask tbl_a, const RepParams<T> &P) {
    uint64_t v[16];
    asm volatile("" : "+r"(tbl_a.neg), "+r"(tbl_a.valid), "+r"(t0));  // keeps the 16 per-string masks and offsets from being hoisted out of the unit loop (registers)
#pragma unroll
    for (int b = 0; b < 16; ++b) {
        const uint8_t *p = unit + (t0 ^ P.f_lo[b]);
        if constexpr (sizeof(T) == 4) v[b] = (uint64_t) * reinterpret_cast<const uint32_t *>(p) | ((uint64_t) * reinterpret_cast<const uint32_t *>(p + 2048) << 32);
        else v[b] = *reinterpret_cast<const uint64_t *>(p);
        if (INPUT) v[b] = Elem<T>::template mask<HALF>(v[b], tbl_a, b);
    }
    wht16<T>(v);
#pragma unroll
    for (int b = 0; b < 16; ++b) {
        if (!INPUT) v[b] = Elem<T>::template mask<HALF>(v[b], tbl_a, b);
        uint8_t *p = unit + (t0 ^ P.f_lo[b]);
        if constexpr (sizeof(T) == 4) {
            *reinterpret_cast<uint32_t *>(p) = (uint32_t)v[b];
            *reinterpret_cast<uint32_t *>(p + 2048) = (uint32_t)(v[b] >> 32);
        } else {
            *reinterpret_cast<uint64_t *>(p) = v[b];
        }
    }
}

// Stage 2 on a reference-layout unit.  tabs = fk[16], cw[16], fo[16] of the pass.  B[k][n = g] is row 8 q + g of the slot
// at k position; D[m][n = 2 t, 2 t + 1] are two rows (128 bytes apart): two stores.
__device__ __forceinline__ void gemm_pass_aos(uint8_t *unit, const float4 *frag, const uint32_t *tabs, int wl, int lane, float) {
    const int g = lane >> 2, t = lane & 3;
    uint32_t lk[2][2], lo[2][2];
#pragma unroll
    for (int ks = 0; ks < 2; ++ks)
#pragma unroll
        for (int h = 0; h < 2; ++h) lk[ks][h] = tabs[8 * ks + 4 * h + t] ^ aos_row_term(g);
#pragma unroll
    for (int mb = 0; mb < 2; ++mb)
#pragma unroll
        for (int e = 0; e < 2; ++e) lo[mb][e] = tabs[32 + g + 8 * mb] ^ aos_row_term(2 * t + e);
#pragma unroll 1
    for (int wi = 0; wi < 2; ++wi) {
        const uint32_t cw = tabs[16 + wl + 8 * wi];
        float d[4][4], b[4][2][2];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
#pragma unroll
            for (int e = 0; e < 4; ++e) d[q][e] = 0.f;
#pragma unroll
            for (int ks = 0; ks < 2; ++ks)
#pragma unroll
                for (int h = 0; h < 2; ++h) b[q][ks][h] = *reinterpret_cast<const float *>(unit + ((lk[ks][h] ^ cw) + (uint32_t)(q << 10)));
        }
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
            const float4 fh = frag[(2 * ks) * 32 + lane], fl = frag[(2 * ks + 1) * 32 + lane];
            const float hi[4] = {fh.x, fh.y, fh.z, fh.w}, lw[4] = {fl.x, fl.y, fl.z, fl.w};
            float bh[4][2], bl[4][2];
#pragma unroll
            for (int q = 0; q < 4; ++q)
#pragma unroll
                for (int h = 0; h < 2; ++h) split_tf32(b[q][ks][h], bh[q][h], bl[q][h]);
#pragma unroll
            for (int q = 0; q < 4; ++q) mma_tf32(d[q], lw, bh[q][0], bh[q][1]);  // small terms first
#pragma unroll
            for (int q = 0; q < 4; ++q) mma_tf32(d[q], hi, bl[q][0], bl[q][1]);
#pragma unroll
            for (int q = 0; q < 4; ++q) mma_tf32(d[q], hi, bh[q][0], bh[q][1]);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int mb = 0; mb < 2; ++mb)
#pragma unroll
                for (int e = 0; e < 2; ++e) *reinterpret_cast<float *>(unit + ((lo[mb][e] ^ cw) + (uint32_t)(q << 10))) = d[q][2 * mb + e];
    }
}
__device__ __forceinline__ void gemm_pass_aos(uint8_t *unit, const double2 *frag, const uint32_t *tabs, int wl, int lane, double) {
    const int g = lane >> 2, t = lane & 3;
    uint32_t lk[4], lo[2][2];
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) lk[ks] = tabs[4 * ks + t] ^ aos_row_term(g);
#pragma unroll
    for (int mb = 0; mb < 2; ++mb)
#pragma unroll
        for (int e = 0; e < 2; ++e) lo[mb][e] = tabs[32 + g + 8 * mb] ^ aos_row_term(2 * t + e);
    const uint32_t cw[2] = {tabs[16 + wl], tabs[16 + wl + 8]};
    double d[2][2][2][2], b[2][2][4];
#pragma unroll
    for (int wi = 0; wi < 2; ++wi)
#pragma unroll
        for (int q = 0; q < 2; ++q) {
#pragma unroll
            for (int mb = 0; mb < 2; ++mb) d[wi][q][mb][0] = d[wi][q][mb][1] = 0.0;
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) b[wi][q][ks] = *reinterpret_cast<const double *>(unit + ((lk[ks] ^ cw[wi]) + (uint32_t)(q << 10)));
        }
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
        const double2 f = frag[ks * 32 + lane];
#pragma unroll
        for (int wi = 0; wi < 2; ++wi)
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                mma_f64(d[wi][q][0], f.x, b[wi][q][ks]);
                mma_f64(d[wi][q][1], f.y, b[wi][q][ks]);
            }
    }
#pragma unroll
    for (int wi = 0; wi < 2; ++wi)
#pragma unroll
        for (int q = 0; q < 2; ++q)
#pragma unroll
            for (int mb = 0; mb < 2; ++mb)
#pragma unroll
                for (int e = 0; e < 2; ++e) *reinterpret_cast<double *>(unit + ((lo[mb][e] ^ cw[wi]) + (uint32_t)(q << 10))) = d[wi][q][mb][e];
}

// shared memory behind the ring: barriers, sign table, pass tables, factor fragments
constexpr int kAuxBar = 0, kAuxSlot = 64, kAuxTabs = 128, kAuxFrag = kAuxTabs + 2 * 48 * 4, kAuxBytes = kAuxFrag + 2 * 4 * 32 * 16;

// NPASS 1: out = M x (column pass) or x M (row pass) -- the host chooses the tables; NPASS 2: out = R x ~R (column pass
// with rho(R) / 16, row pass with rho(~R)).  NG groups of 8 warps.
template <typename T, int NPASS, int NG, bool AOS, bool HALF>
__global__ void __launch_bounds__(NG *kGroupThreads, 1)
matrix_rep_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_out,
                  const __grid_constant__ RepParams<T> P, int nb, int64_t n, int dbg, unsigned int *counter, const T *__restrict__ factor_dev) {
    constexpr int TR = 128 / (int)sizeof(T);
    using Frag = typename Factor<T>::Frag;
    extern __shared__ __align__(1024) uint8_t smem[];  // no static shared memory in this kernel: the window starts aligned
    uint8_t *aux = smem + (size_t)kRing * kUnitBytes;
    uint64_t *full = reinterpret_cast<uint64_t *>(aux + kAuxBar);
    int *slot_row = reinterpret_cast<int *>(aux + kAuxSlot);  // first row of the unit in each ring slot, -1: no unit left
    uint32_t *tabs = reinterpret_cast<uint32_t *>(aux + kAuxTabs);
    Frag *frag = reinterpret_cast<Frag *>(aux + kAuxFrag);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int units = (int)((n + TR - 1) / TR);
    const uint32_t box_bytes = (uint32_t)nb * 128u;
    // Which units a CTA works on: its first ring is static (unit blockIdx.x + k * gridDim.x in slot k); after that a TMA
    // thread takes the next unit from a global counter whenever it refills a slot.  SMs do not get the same share of the HBM
    // bandwidth (two dies, different distances to the memory partitions), and with equal static shares the slowest SM sets
    // the time: a pass-through ring moves 6.2 TB/s with static shares and 6.9-7.0 with the counter
    // (scripts/tma_passthrough_probe.cu, profiles/r02_tma_passthrough_probe.jsonl).  The thread asks one refill ahead, so
    // the round trip of the atomic is off its group's critical path.  counter == nullptr (batches of at most one ring per
    // CTA, LCC_K8_DYNAMIC=0): the static interleaved sequence goes on.
    int pending = 0;
    auto ask = [&]() { if (counter) pending = (int)gridDim.x * kRing + (int)atomicAdd(counter, 1u); };
    auto next_row = [&](int k) -> int {  // first row of local unit k, or -1 when the batch is exhausted
        int u = (int)blockIdx.x + k * (int)gridDim.x;
        if (counter && k >= kRing) { u = pending; ask(); }
        return u < units ? u * TR : -1;
    };
    if (threadIdx.x == 0) {
        for (int s = 0; s < kRing; ++s) tma::mbar_init(&full[s], 1);
        tma::prefetch_map(&map_x);
        tma::prefetch_map(&map_out);
    }
    if (threadIdx.x < 96) {
        const int p = threadIdx.x / 48, i = threadIdx.x % 48;
        tabs[threadIdx.x] = i < 16 ? P.fk[p][i] : (i < 32 ? P.cw[p][i - 16] : P.fo[p][i - 32]);
    }
    // the fixed factor comes with the parameters (operand known on the host) or from a device buffer a builder kernel
    // filled from a device-resident operand (no host read-back)
    if (warp < NPASS) Factor<T>::build(frag + warp * 4 * 32, factor_dev ? factor_dev + warp * 256 : P.factor[warp], lane);
    __syncthreads();
    auto load_slot = [&](int k) {  // local unit k into ring slot k % kRing (TMA threads only)
        const int rs = k % kRing, row = next_row(k);
        slot_row[rs] = row;
        if (row >= 0) {
            tma::mbar_expect_tx(&full[rs], box_bytes);
            if (AOS) tma::load_3d(smem + (size_t)rs * kUnitBytes, &map_x, &full[rs], 0, row, 0);
            else tma::load_2d(smem + (size_t)rs * kUnitBytes, &map_x, &full[rs], row, 0);
        }
    };
    if (threadIdx.x == 0)
        for (int k = 0; k < kRing; ++k) load_slot(k);  // fill the ring
    __syncthreads();  // slot_row of the first ring is visible to every group
    if ((threadIdx.x & (kGroupThreads - 1)) == 0) ask();
    const int grp = warp >> 3, wl = warp & 7;
    const int gtid = threadIdx.x & (kGroupThreads - 1);
    auto group_sync = [&]() { asm volatile("bar.sync %0, %1;" ::"r"(grp + 1), "n"(kGroupThreads) : "memory"); };
    // stage 1 / 3 role.  Blade-major units: string row a = 2 * wl + (lane >> 4), 8-byte element lane & 15 of the 128-byte
    // line.  Reference-layout units: physical string row 4 * (wl >> 1) + (lane >> 3) at row 8 * (wl & 1) + (lane & 7) (f32:
    // and that row + 16).
    const int a_phys = AOS ? 4 * (wl >> 1) + (lane >> 3) : 2 * wl + (lane >> 4);
    const int a_row = AOS ? (int)P.amap[a_phys] : a_phys;
    const uint32_t t0 = P.f_hi[a_phys] ^ (AOS ? aos_row_term(8 * (wl & 1) + (lane & 7)) : (uint32_t)((lane & 15) << 3));
    const RowMask tbl_a = {(P.neg[a_row >> 1] >> ((a_row & 1) * 16)) & 0xffffu, (P.valid[a_row >> 1] >> ((a_row & 1) * 16)) & 0xffffu};
    // The group's first thread is also its TMA thread: it stores the group's finished unit and, one barrier into the
    // group's NEXT unit (by then the engine has read the slot), refills that slot with the unit kRing ahead.
    int stored = -1;
    auto refill = [&]() {
        if (gtid == 0 && stored >= 0) {
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            load_slot(stored + kRing);  // (published to the group by the barriers between here and its use, two units on)
            stored = -1;
        }
    };
    auto store_unit = [&](int it, int s) {
        if (gtid == 0) {
            if (AOS) tma::store_3d(&map_out, smem + (size_t)s * kUnitBytes, 0, slot_row[s], 0);
            else tma::store_2d(&map_out, smem + (size_t)s * kUnitBytes, slot_row[s], 0);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            stored = it;
        }
    };
    for (int it = grp;; it += NG) {
        const int s = it % kRing;
        if (slot_row[s] < 0) break;  // units are dealt in increasing order: nothing comes after an empty slot
        const uint32_t ph = (uint32_t)((it / kRing) & 1);
        uint8_t *unit = smem + (size_t)s * kUnitBytes;
        tma::mbar_wait(&full[s], ph);
        if (dbg == 1) {  // LCC_K8_DEBUG=1: units pass through untouched (memory pipeline alone)
            refill();
            group_sync();
            store_unit(it, s);
            continue;
        }
        if (AOS) transform_row_aos<T, true, HALF>(unit, t0, tbl_a, P);
        else transform_row<T, true, HALF>(unit, t0, tbl_a, P);
        refill();      // the previous unit's store has long been read: put the next load in flight
        group_sync();
        if (dbg != 2) {  // LCC_K8_DEBUG=2: transforms only
#pragma unroll
            for (int p = 0; p < NPASS; ++p) {
                if (AOS) gemm_pass_aos(unit, frag + p * 4 * 32, tabs + p * 48, wl, lane, T());
                else gemm_pass(unit, frag + p * 4 * 32, tabs + p * 48, wl, lane, T());
                group_sync();
            }
        }
        if (AOS) transform_row_aos<T, false, HALF>(unit, t0, tbl_a, P);
        else transform_row<T, false, HALF>(unit, t0, tbl_a, P);
        tma::fence_proxy_async();
        group_sync();
        store_unit(it, s);
    }
    refill();  // (waits for the read of the last store; nothing left to load)
    if (gtid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // global writes complete before the CTA exits
}

int rep_groups() {
    static const int g = [] { const char *e = getenv("LCC_K8_GROUPS"); const int v = e ? atoi(e) : 0; return v == 2 || v == 3 ? v : 0; }();
    return g;
}

template <typename T, int NPASS, int NG, bool AOS, bool HALF>
cudaError_t launch_rep_h(int nb, const void *x, int64_t ldx, void *out, int64_t ldo, int64_t n, const RepParams<T> &P, const T *factor_dev, cudaStream_t s) {
    constexpr int TR = 128 / (int)sizeof(T);
    CUtensorMap mx, mo;
    static const int promo = [] { const char *e = getenv("LCC_K8_L2PROMO"); return e ? atoi(e) : 256; }();
    static const bool swz = [] { const char *e = getenv("LCC_K8_SWIZZLE"); return !(e && e[0] == '0'); }();  // 0: only meaningful with LCC_K8_DEBUG=1
    if (AOS) {
        if (!make_aos_chunked_tensor_map(&mx, x, (int)sizeof(T), nb, ldx, n, TR) || !make_aos_chunked_tensor_map(&mo, out, (int)sizeof(T), nb, ldo, n, TR))
            return cudaErrorNotSupported;
    } else if (!make_soa_tensor_map(&mx, x, (int)sizeof(T), nb, ldx, n, TR, swz, promo) || !make_soa_tensor_map(&mo, out, (int)sizeof(T), nb, ldo, n, TR, swz, promo)) {
        return cudaErrorNotSupported;
    }
    auto kernel = matrix_rep_kernel<T, NPASS, NG, AOS, HALF>;
    constexpr int smem = kRing * kUnitBytes + kAuxBytes;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    static int sm_count = [] { int d = 0, c = 0; cudaGetDevice(&d); cudaDeviceGetAttribute(&c, cudaDevAttrMultiProcessorCount, d); return c > 0 ? c : 148; }();
    const int64_t units = (n + TR - 1) / TR;
    const int grid = (int)(units < sm_count ? units : sm_count);
    static const int dbg = [] { const char *e = getenv("LCC_K8_DEBUG"); return e ? atoi(e) : 0; }();
    // more units than one ring per CTA: deal the rest out through a counter (stream-ordered scratch word, zeroed before the launch)
    static const bool dyn = [] { const char *e = getenv("LCC_K8_DYNAMIC"); return !(e && e[0] == '0'); }();
    unsigned int *counter = nullptr;
    if (dyn && units > (int64_t)grid * kRing) {
        e = cudaMallocAsync((void **)&counter, sizeof(unsigned int), s);
        if (e != cudaSuccess) return e;
        e = cudaMemsetAsync(counter, 0, sizeof(unsigned int), s);
        if (e != cudaSuccess) { cudaFreeAsync(counter, s); return e; }
    }
    kernel<<<grid, NG * kGroupThreads, smem, s>>>(mx, mo, P, nb, n, dbg, counter, factor_dev);
    note_kernel_launch();
    e = cudaGetLastError();
    if (counter) cudaFreeAsync(counter, s);
    return e;
}

template <typename T, int NPASS, int NG, bool AOS>
cudaError_t launch_rep_t(int nb, const void *x, int64_t ldx, void *out, int64_t ldo, int64_t n, const RepParams<T> &P, const T *factor_dev, cudaStream_t s) {
    return nb == 128 ? launch_rep_h<T, NPASS, NG, AOS, true>(nb, x, ldx, out, ldo, n, P, factor_dev, s)
                     : launch_rep_h<T, NPASS, NG, AOS, false>(nb, x, ldx, out, ldo, n, P, factor_dev, s);
}

// groups of 8 warps per CTA, measured at 2^24 rows of Cl(8,0,0) (profiles/r02_k8_v8_*): three groups are capped at 80
// registers and spill a little; f32 5.78 ms with two groups against 5.93 with three (reference layout), 6.39 / 6.99
// (blade-major); f64 11.37 / 11.02 (reference layout), 12.36 / 12.51 (blade-major)
template <typename T, bool AOS> constexpr int default_groups() { return sizeof(T) == 8 && AOS ? 3 : 2; }

template <typename T, bool AOS>
cudaError_t launch_rep_p(int nb, int mode, const RepParams<T> &P, const T *factor_dev, const ViewArg &x, const ViewArg &out, cudaStream_t s) {
    const int ng = rep_groups() ? rep_groups() : default_groups<T, AOS>();
    if (mode == 2)
        return ng == 2 ? launch_rep_t<T, 2, 2, AOS>(nb, x.data, x.ld, out.data, out.ld, x.n, P, factor_dev, s)
                       : launch_rep_t<T, 2, 3, AOS>(nb, x.data, x.ld, out.data, out.ld, x.n, P, factor_dev, s);
    if (mode == 0 || mode == 1)
        return ng == 2 ? launch_rep_t<T, 1, 2, AOS>(nb, x.data, x.ld, out.data, out.ld, x.n, P, factor_dev, s)
                       : launch_rep_t<T, 1, 3, AOS>(nb, x.data, x.ld, out.data, out.ld, x.n, P, factor_dev, s);
    return cudaErrorInvalidValue;
}

// rho(M) / 16 (or its transpose) of a DEVICE-resident operand M, written in the MMA's physical m / k order:
// rho(M)[j ^ a][j] = sum_b (-1)^popcount(b & j) sign(a, b) M[blade(a, b)]   (csrc/matrix_rep.hpp).  One thread per entry.
// `pass` selects the index maps (sandwich: 0 = column pass with rho(R) / 16, 1 = row pass with rho(~R)^T); `reverse`
// multiplies every coefficient by its reverse sign (blades.rs:128-135): ~R without materialising it.
struct RepTables { uint8_t blade_at[256]; int8_t sign_at[256]; uint8_t omap[2][16], kmap[2][16]; };
template <typename T>
__global__ void build_rep_factor_kernel(const T *__restrict__ m, int64_t m_stride, const __grid_constant__ RepTables tb, int pass, int transpose,
                                        int reverse, double scale, T *__restrict__ factor) {
    const int mi = threadIdx.x >> 4, ki = threadIdx.x & 15;
    const int o = tb.omap[pass][mi], k = tb.kmap[pass][ki];
    const int i = transpose ? k : o, j = transpose ? o : k;
    const int a = i ^ j;
    double acc = 0.0;
    for (int b = 0; b < 16; ++b) {
        const int str = (a << 4) | b, sg = tb.sign_at[str];
        if (sg == 0) continue;
        const int blade = tb.blade_at[str];
        double v = (double)m[(int64_t)blade * m_stride];
        if (reverse && (__popc(blade) & 3) >= 2) v = -v;
        acc += ((__popc(b & j) & 1) != 0) == (sg > 0) ? -v : v;
    }
    factor[mi * 16 + ki] = (T)(acc * scale);
}

// The shared-memory layouts of a representation (index bases found by search: 6-25 us of host time) are computed once and
// kept, keyed by the representation's tables.
struct LayoutSet { uint8_t blade_at[256]; int8_t sign_at[256]; RepLayout L; RepLayoutAos A4, A8; };
LayoutSet layouts_of(const uint8_t *blade_at, const int8_t *sign_at) {
    static std::mutex mu;
    static std::vector<std::unique_ptr<LayoutSet>> cache;
    std::lock_guard<std::mutex> lock(mu);
    for (const auto &e : cache)
        if (!memcmp(e->blade_at, blade_at, 256) && !memcmp(e->sign_at, sign_at, 256)) return *e;
    if (cache.size() >= 64) cache.erase(cache.begin());
    auto e = std::make_unique<LayoutSet>();
    memcpy(e->blade_at, blade_at, 256);
    memcpy(e->sign_at, sign_at, 256);
    e->L = make_rep_layout(blade_at, sign_at);
    e->A4 = make_rep_layout_aos(e->L, 4);
    e->A8 = make_rep_layout_aos(e->L, 8);
    cache.push_back(std::move(e));
    return *cache.back();
}

template <typename T>
cudaError_t launch_rep(int nb, int mode, const double *p1, const double *p2, const void *m_dev, int64_t m_stride, const uint8_t *blade_at,
                       const int8_t *sign_at, const ViewArg &x, const ViewArg &out, cudaStream_t s) {
    const LayoutSet layouts = layouts_of(blade_at, sign_at);  // (a copy: 3 KB, and no lifetime question)
    const RepLayout &L = layouts.L;
    const bool aos = x.layout == kAoS;
    RepParams<T> P;
    const double *fac[2] = {p1, p2};
    RepTables tb;
    for (int i = 0; i < 256; ++i) { tb.blade_at[i] = blade_at[i]; tb.sign_at[i] = sign_at[i]; }
    // mode 0: column pass with p1; mode 1: row pass with p1; mode 2: column pass with p1, row pass with p2
    if (aos) {
        const RepLayoutAos &A = sizeof(T) == 4 ? layouts.A4 : layouts.A8;
        if (!A.ok) return cudaErrorNotSupported;
        const RepPassAos *pass[2] = {mode == 1 ? &A.row : &A.column, &A.row};
        for (int p = 0; p < 2; ++p)
            for (int i = 0; i < 16; ++i) {
                P.fk[p][i] = pass[p]->fk[i];
                P.cw[p][i] = pass[p]->cw[i];
                P.fo[p][i] = pass[p]->fo[i];
                for (int k = 0; k < 16; ++k) P.factor[p][i * 16 + k] = fac[p] ? (T)fac[p][pass[p]->omap[i] * 16 + pass[p]->kmap[k]] : T(0);
            }
        for (int i = 0; i < 16; ++i) { P.f_lo[i] = A.g_lo[i]; P.f_hi[i] = A.g_hi[i]; P.amap[i] = A.amap[i]; for (int p = 0; p < 2; ++p) { tb.omap[p][i] = pass[p]->omap[i]; tb.kmap[p][i] = pass[p]->kmap[i]; } }
    } else {
        if (!L.ok) return cudaErrorNotSupported;
        const RepPass *pass[2] = {mode == 1 ? &L.row : &L.column, &L.row};
        for (int p = 0; p < 2; ++p)
            for (int i = 0; i < 16; ++i) {
                P.fk[p][i] = P.fo[p][i] = pass[p]->fk[i];
                P.cw[p][i] = pass[p]->cw[i];
                for (int k = 0; k < 16; ++k) P.factor[p][i * 16 + k] = fac[p] ? (T)fac[p][pass[p]->kmap[i] * 16 + pass[p]->kmap[k]] : T(0);
            }
        for (int i = 0; i < 16; ++i) { P.f_lo[i] = L.f_lo[i]; P.f_hi[i] = L.f_hi[i]; P.amap[i] = (uint32_t)i; for (int p = 0; p < 2; ++p) tb.omap[p][i] = tb.kmap[p][i] = pass[p]->kmap[i]; }
    }
    for (int i = 0; i < 8; ++i) P.neg[i] = P.valid[i] = 0;
    for (int str = 0; str < 256; ++str) {
        if (sign_at[str] != 0) P.valid[str >> 5] |= 1u << (str & 31);
        if (sign_at[str] < 0) P.neg[str >> 5] |= 1u << (str & 31);
    }
    T *factor_dev = nullptr;
    if (m_dev) {  // device-resident operand (mode 0: fixed * x, mode 1: x * fixed, mode 2: R x ~R): the factors are built on the device
        cudaError_t e = cudaMallocAsync((void **)&factor_dev, 512 * sizeof(T), s);
        if (e != cudaSuccess) return e;
        build_rep_factor_kernel<T><<<1, 256, 0, s>>>((const T *)m_dev, m_stride, tb, 0, mode == 1 ? 1 : 0, 0, 1.0 / 16.0, factor_dev);
        note_kernel_launch();
        if (mode == 2) {
            build_rep_factor_kernel<T><<<1, 256, 0, s>>>((const T *)m_dev, m_stride, tb, 1, 1, 1, 1.0, factor_dev + 256);
            note_kernel_launch();
        }
    }
    const cudaError_t e = aos ? launch_rep_p<T, true>(nb, mode, P, factor_dev, x, out, s) : launch_rep_p<T, false>(nb, mode, P, factor_dev, x, out, s);
    if (factor_dev) cudaFreeAsync(factor_dev, s);
    return e;
}

}  // namespace

// K8 launcher.  mode 0 / 1 / 2 = left product, right product, sandwich; p1 (and p2 for the sandwich) are row-major 16 x 16
// f64 factors in the orientation the MMA contracts ([o][k]: mode 0 rho(M) / 16, mode 1 rho(M)^T / 16, mode 2 rho(R) / 16 and
// rho(~R)^T) -- or m_dev: the fixed operand / rotor itself as a device row of the batch dtype (element k at
// m_dev[k * m_stride]), from which a one-block kernel builds the factor; blade_at / sign_at map a string to its blade and sign.  x and out are views of nb = 128 or 256 blades in ONE
// layout: blade-major (SoA) with a row count that is a multiple of 16 bytes (TMA stores clip at that granularity), or the
// reference's row-major (AoS) layout with any row count -- whole rows are contiguous there, and a unit is one 16-32 KB span.
cudaError_t launch_matrix_rep_product(int nb, int mode, const double *p1, const double *p2, const uint8_t *blade_at, const int8_t *sign_at,
                                      const ViewArg &x, const ViewArg &out, cudaStream_t s, const void *m_dev, int64_t m_stride) {
    if (x.n <= 0) return cudaSuccess;
    if (x.layout != out.layout || (nb != 128 && nb != 256)) return cudaErrorNotSupported;
    return x.dtype == kF32 ? launch_rep<float>(nb, mode, p1, p2, m_dev, m_stride, blade_at, sign_at, x, out, s)
                           : launch_rep<double>(nb, mode, p1, p2, m_dev, m_stride, blade_at, sign_at, x, out, s);
}

}  // namespace lcc
