// common.cuh -- shared device helpers for the sm_100a batched Clifford-algebra kernels.
//
// Storage model (DESIGN.md "Data layout in HBM"):
//   SoA  element (row r, blade k) at base[k*ld + r]; a thread owns V consecutive rows and moves them
//        with one 128-bit (or 64-bit) access per blade, so a warp touches 512 contiguous bytes per
//        blade per instruction -- every HBM sector it opens is fully used.
//   AoS  element (row r, blade k) at base[r*ld + k] (the reference's Array2<f64> layout,
//        /root/reference/src/batch.rs:41-47); a thread owns one row and reads it in 128-bit pieces.
// Streaming data is read through the non-coherent path without L1 allocation and written with
// the streaming (evict-first) hint: nothing is reused, so it must not displace anything.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace lcc {

constexpr int kThreads = 256;

template <class T> struct Vec2 { T x, y; };

// ---------------------------------------------------------------- arithmetic policies
// FAST: one FMA per term.  STRICT: separately rounded multiply and add, the reference's sequence
// (Rust never contracts a*b+c; /root/reference/src/batch.rs:398-401).
template <bool STRICT> struct Arith;
template <> struct Arith<false> {
    static __device__ __forceinline__ float madd(float a, float b, float c) { return fmaf(a, b, c); }
    static __device__ __forceinline__ double madd(double a, double b, double c) { return fma(a, b, c); }
};
template <> struct Arith<true> {
    static __device__ __forceinline__ float madd(float a, float b, float c) { return __fadd_rn(c, __fmul_rn(a, b)); }
    static __device__ __forceinline__ double madd(double a, double b, double c) { return __dadd_rn(c, __dmul_rn(a, b)); }
};
static __device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
static __device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
static __device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
static __device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
static __device__ __forceinline__ float div_rn(float a, float b) { return __fdiv_rn(a, b); }
static __device__ __forceinline__ double div_rn(double a, double b) { return __ddiv_rn(a, b); }
static __device__ __forceinline__ float sqrt_rn(float a) { return __fsqrt_rn(a); }
static __device__ __forceinline__ double sqrt_rn(double a) { return __dsqrt_rn(a); }

// ---------------------------------------------------------------- streaming loads / stores
// 128-bit
static __device__ __forceinline__ void ldg_stream(const float *p, float (&v)[4]) {
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]) : "l"(p));
}
static __device__ __forceinline__ void ldg_stream(const double *p, double (&v)[2]) {
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(v[0]), "=d"(v[1]) : "l"(p));
}
// 64-bit
static __device__ __forceinline__ void ldg_stream(const float *p, float (&v)[2]) {
    asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0,%1}, [%2];" : "=f"(v[0]), "=f"(v[1]) : "l"(p));
}
static __device__ __forceinline__ void ldg_stream(const double *p, double (&v)[1]) {
    asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v[0]) : "l"(p));
}
// 32-bit
static __device__ __forceinline__ void ldg_stream(const float *p, float (&v)[1]) {
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v[0]) : "l"(p));
}

static __device__ __forceinline__ void stg_stream(float *p, const float (&v)[4]) {
    asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" :: "l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]) : "memory");
}
static __device__ __forceinline__ void stg_stream(double *p, const double (&v)[2]) {
    asm volatile("st.global.cs.v2.f64 [%0], {%1,%2};" :: "l"(p), "d"(v[0]), "d"(v[1]) : "memory");
}
static __device__ __forceinline__ void stg_stream(float *p, const float (&v)[2]) {
    asm volatile("st.global.cs.v2.f32 [%0], {%1,%2};" :: "l"(p), "f"(v[0]), "f"(v[1]) : "memory");
}
static __device__ __forceinline__ void stg_stream(double *p, const double (&v)[1]) {
    asm volatile("st.global.cs.f64 [%0], %1;" :: "l"(p), "d"(v[0]) : "memory");
}
static __device__ __forceinline__ void stg_stream(float *p, const float (&v)[1]) {
    asm volatile("st.global.cs.f32 [%0], %1;" :: "l"(p), "f"(v[0]) : "memory");
}

// ---------------------------------------------------------------- rows-per-thread choice
// Widest access that keeps both operands of a binary product (2*NB*V values) within BUDGET bytes of
// registers per thread.  512 bytes (specialised kernels): 128-bit up to 16 blades, 64-bit at 32
// blades.  The run-time-metric kernels carry extra temporaries and use 256 bytes.
template <class T, int NB, int BUDGET = 512> struct RowsPerThread {
    static constexpr int kFull = 16 / (int)sizeof(T);
    static constexpr int kBudget = BUDGET / (2 * NB * (int)sizeof(T));
    static constexpr int value = kBudget >= kFull ? kFull : (kBudget >= 1 ? kBudget : 1);
};

// ---------------------------------------------------------------- operand tiles in registers
// Tile<T,NB,V>: NB blades x V rows held by one thread.
template <class T, int NB, int V> struct Tile { T v[NB][V]; };

// SoA: one vector access per blade (rows r0..r0+V-1).  `bcast` (warp-uniform) loads a one-row
// operand and replicates it (reference broadcast rule, batch.rs:361-376).
template <class T, int NB, int V>
static __device__ __forceinline__ void load_soa(Tile<T, NB, V> &t, const T *__restrict__ base, int64_t ld,
                                                int64_t r0, bool bcast) {
    if (bcast) {
#pragma unroll
        for (int k = 0; k < NB; ++k) {
            T s = __ldg(base + (int64_t)k * ld);
#pragma unroll
            for (int l = 0; l < V; ++l) t.v[k][l] = s;
        }
    } else {
#pragma unroll
        for (int k = 0; k < NB; ++k) ldg_stream(base + (int64_t)k * ld + r0, t.v[k]);
    }
}

// AoS: V == 1; the thread's row is contiguous (NB values), read in 128-bit pieces when aligned.
template <class T, int NB>
static __device__ __forceinline__ void load_aos(Tile<T, NB, 1> &t, const T *__restrict__ base, int64_t ld,
                                                int64_t r, bool bcast, bool vec_ok) {
    const T *row = base + (bcast ? 0 : r * ld);
    constexpr int W = 16 / (int)sizeof(T);
    if (NB >= W && vec_ok) {
#pragma unroll
        for (int k = 0; k < NB; k += W) {
            T tmp[W];
            if (bcast) {
#pragma unroll
                for (int l = 0; l < W; ++l) tmp[l] = __ldg(row + k + l);
            } else {
                ldg_stream(row + k, tmp);
            }
#pragma unroll
            for (int l = 0; l < W; ++l) t.v[k + l][0] = tmp[l];
        }
    } else {
#pragma unroll
        for (int k = 0; k < NB; ++k) t.v[k][0] = __ldg(row + k);
    }
}

template <class T, int V>
static __device__ __forceinline__ void store_soa(T *__restrict__ base, int64_t ld, int k, int64_t r0, const T (&val)[V]) {
    stg_stream(base + (int64_t)k * ld + r0, val);
}

}  // namespace lcc
