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
// Shape: persistent, one CTA per SM.  Warp 16 streams 32-row tiles (32 x nb box of the blade-major batch) into a ring of
// shared-memory stages with TMA, warps 0-15 transform a stage IN PLACE, warp 17 hands it back to the TMA engine for the
// store and frees the stage once the engine has read it -- loads run several tiles ahead of the arithmetic, stores
// trail it.  Inside a stage, element (slot s, row r) lives at [s * 32 + r]: a warp always works on ONE slot for 32
// consecutive rows, so every shared-memory access is conflict free, and the three stages only re-index the 256 slots
// (blade index -> matrix entry -> blade index).  The 16 x 16 factor(s) travel as kernel parameters and are compile-time
// indexed, so they enter the FMAs as constant-bank operands: no loads for the fixed operand at all.
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>

#include "launch_args.h"
#include "misc_kernels.cuh"
#include "tma_tile.cuh"

namespace lcc {
namespace {

constexpr int kTileRows = 32;
constexpr int kConsumerWarps = 16, kConsumers = kConsumerWarps * 32, kThreadsRep = kConsumers + 64;

template <typename T> struct RepParams {
    T p1[256];           // first factor, [o * 16 + k]:  acc[o] += p1[o][k] * x[k]
    T p2[256];           // second factor (sandwich only), applied from the right
    uint8_t blade[256];  // string (a << 4 | b) -> blade index
    int8_t sign[256];    // string -> +1 / -1, 0 when no blade maps to the string (128-blade algebras)
};

template <typename T> __device__ __forceinline__ void wht16(T (&v)[16]) {
#pragma unroll
    for (int h = 1; h < 16; h <<= 1)
#pragma unroll
        for (int i = 0; i < 16; ++i)
            if (!(i & h)) { const T x = v[i], y = v[i | h]; v[i] = x + y; v[i | h] = x - y; }
}

__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kConsumers) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tma::smem_u32(bar)) : "memory");
}

// MODE 0: out = M * x   (C = P1 * X, a thread owns column j)
// MODE 1: out = x * M   (C = X * P1^T-indexed, a thread owns row i)
// MODE 2: out = R x ~R  (MODE 0 with p1 = rho(R) / 16, then MODE 1 with p2 = rho(~R))
template <typename T, int MODE, int NS>
__global__ void __launch_bounds__(kThreadsRep, 1)
matrix_rep_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_out,
                  const __grid_constant__ RepParams<T> P, int nb, int64_t n) {
    constexpr int kG = 4;                                   // outputs accumulated side by side
    constexpr int kOuterUnroll = sizeof(T) == 4 ? 4 : 1;   // of the 4 groups of kG outputs
    constexpr uint32_t kStageElems = 256 * kTileRows;
    constexpr uint32_t kStageBytes = kStageElems * sizeof(T);
    extern __shared__ __align__(1024) uint8_t smem[];  // no static shared memory in this kernel: the window starts aligned
    uint64_t *bars = (uint64_t *)(smem + (size_t)NS * kStageBytes);
    uint64_t *full = bars, *done = bars + NS, *empty = bars + 2 * NS;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < NS; ++s) { tma::mbar_init(&full[s], 1); tma::mbar_init(&done[s], 1); tma::mbar_init(&empty[s], 1); }
    }
    __syncthreads();
    const int64_t tiles = (n + kTileRows - 1) / kTileRows;
    const uint32_t box_bytes = (uint32_t)nb * kTileRows * sizeof(T);

    if (warp == kConsumerWarps) {            // ---- loader
        if (lane == 0) {
            tma::prefetch_map(&map_x);
            int s = 0; uint32_t ph = 0;
            for (int64_t t = blockIdx.x; t < tiles; t += gridDim.x) {
                tma::mbar_wait(&empty[s], ph ^ 1);
                tma::mbar_expect_tx(&full[s], box_bytes);
                tma::load_2d(smem + (size_t)s * kStageBytes, &map_x, &full[s], (int)(t * kTileRows), 0);
                if (++s == NS) { s = 0; ph ^= 1; }
            }
        }
    } else if (warp == kConsumerWarps + 1) {  // ---- storer
        if (lane == 0) {
            tma::prefetch_map(&map_out);
            int s = 0; uint32_t ph = 0;
            for (int64_t t = blockIdx.x; t < tiles; t += gridDim.x) {
                tma::mbar_wait(&done[s], ph);
                tma::store_2d(&map_out, smem + (size_t)s * kStageBytes, (int)(t * kTileRows), 0);
                tma::store_commit_and_wait_read();  // the engine has read the stage: it may be refilled
                mbar_arrive(&empty[s]);
                if (++s == NS) { s = 0; ph ^= 1; }
            }
            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // global writes complete before the CTA exits
        }
    } else {                                  // ---- 16 consumer warps: warp w = one index (a, j, i or c), lane = row
        const int w = warp;
        int s = 0; uint32_t ph = 0;
        for (int64_t t = blockIdx.x; t < tiles; t += gridDim.x) {
            T *st = (T *)(smem + (size_t)s * kStageBytes) + lane;
            tma::mbar_wait(&full[s], ph);
            T v[16];
            // stage 1: gather y[a][.] (a = w) from the blade slots, transform, scatter to the matrix slots (j^a, j)
#pragma unroll
            for (int b = 0; b < 16; ++b) {
                const int str = (w << 4) | b;
                const int sg = P.sign[str];
                const T x = st[(int)P.blade[str] * kTileRows];
                v[b] = sg == 0 ? T(0) : (sg > 0 ? x : -x);
            }
            wht16(v);
            consumer_sync();  // every blade slot has been read
#pragma unroll
            for (int j = 0; j < 16; ++j) st[(((j ^ w) << 4) | j) * kTileRows] = v[j];
            consumer_sync();
            // stage 2: 16 x 16 x 16 in place.  The column (row) is in registers, so each result goes straight back to its
            // slot; kG outputs are accumulated side by side so that the FMA chains (16 dependent FMAs per output) overlap.
            // f32: fully unrolled, the factor enters the FFMAs as uniform-register operands.  f64: the group loop stays
            // rolled (unrolled, ptxas hoists all 256 doubles out of the tile loop and spills them to local memory).
            if (MODE == 0 || MODE == 2) {  // column j = w:  C[i][j] = sum_k P1[i][k] X[k][j]
#pragma unroll
                for (int k = 0; k < 16; ++k) v[k] = st[((k << 4) | w) * kTileRows];
#pragma unroll kOuterUnroll
                for (int i0 = 0; i0 < 16; i0 += kG) {
                    T a0[kG];
#pragma unroll
                    for (int g = 0; g < kG; ++g) a0[g] = T(0);
#pragma unroll
                    for (int k = 0; k < 16; ++k)
#pragma unroll
                        for (int g = 0; g < kG; ++g) a0[g] = fma(P.p1[(i0 + g) * 16 + k], v[k], a0[g]);
#pragma unroll
                    for (int g = 0; g < kG; ++g) st[(((i0 + g) << 4) | w) * kTileRows] = a0[g];
                }
                if (MODE == 2) consumer_sync();
            }
            if (MODE == 1 || MODE == 2) {  // row i = w:  C[i][j] = sum_k X[i][k] P[k][j],  factor stored as [j][k]
#pragma unroll
                for (int k = 0; k < 16; ++k) v[k] = st[((w << 4) | k) * kTileRows];
#pragma unroll kOuterUnroll
                for (int j0 = 0; j0 < 16; j0 += kG) {
                    T a0[kG];
#pragma unroll
                    for (int g = 0; g < kG; ++g) a0[g] = T(0);
#pragma unroll
                    for (int k = 0; k < 16; ++k)
#pragma unroll
                        for (int g = 0; g < kG; ++g) a0[g] = fma(MODE == 2 ? P.p2[(j0 + g) * 16 + k] : P.p1[(j0 + g) * 16 + k], v[k], a0[g]);
#pragma unroll
                    for (int g = 0; g < kG; ++g) st[((w << 4) | (j0 + g)) * kTileRows] = a0[g];
                }
            }
            consumer_sync();
            // stage 3: gather C[j^c][j] (c = w), transform back, scatter to the blade slots
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = st[(((j ^ w) << 4) | j) * kTileRows];
            wht16(v);
            consumer_sync();  // every matrix slot has been read
#pragma unroll
            for (int b = 0; b < 16; ++b) {
                const int str = (w << 4) | b;
                const int sg = P.sign[str];
                if (sg != 0) st[(int)P.blade[str] * kTileRows] = sg > 0 ? v[b] : -v[b];
            }
            tma::fence_proxy_async();
            consumer_sync();
            if (threadIdx.x == 0) mbar_arrive(&done[s]);
            if (++s == NS) { s = 0; ph ^= 1; }
        }
    }
}

template <typename T> struct RepStages { static constexpr int value = sizeof(T) == 4 ? 5 : 3; };  // 160 / 192 KB of stages

template <typename T, int MODE>
cudaError_t launch_rep_t(int nb, const void *x, int64_t ldx, void *out, int64_t ldo, int64_t n, const RepParams<T> &P, cudaStream_t s) {
    constexpr int NS = RepStages<T>::value;
    CUtensorMap mx, mo;
    if (!make_soa_tensor_map(&mx, x, (int)sizeof(T), nb, ldx, n, kTileRows) || !make_soa_tensor_map(&mo, out, (int)sizeof(T), nb, ldo, n, kTileRows))
        return cudaErrorNotSupported;
    auto kernel = matrix_rep_kernel<T, MODE, NS>;
    constexpr int smem = NS * 256 * kTileRows * (int)sizeof(T) + 3 * NS * 8;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    static int sm_count = [] { int d = 0, c = 0; cudaGetDevice(&d); cudaDeviceGetAttribute(&c, cudaDevAttrMultiProcessorCount, d); return c > 0 ? c : 148; }();
    const int64_t tiles = (n + kTileRows - 1) / kTileRows;
    const int grid = (int)(tiles < sm_count ? tiles : sm_count);
    kernel<<<grid, kThreadsRep, smem, s>>>(mx, mo, P, nb, n);
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
