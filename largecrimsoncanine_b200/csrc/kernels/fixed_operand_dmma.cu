// fixed_operand_dmma.cu -- K6d: the f64 form of K6 (fixed_operand_gemm.cu): product of a large batch with ONE fixed
// multivector in a 64 / 128 / 256-blade algebra, cast as a GEMM against the operand's multiplication matrix and run
// on the FP64 tensor pipe (mma.sync.m8n8k4.f64, "DMMA").
//
// Reference semantics: MultivectorBatch(1 x nb).geometric_product(batch) / batch.geometric_product(1 x nb), the
// broadcast branch of /root/reference/src/batch.rs:365-404, in the reference's own precision (f64):
//   left  : out[n] = sum_k sign[n^k][k] * M[n^k] * x[k]      ->  Out = X * Wt^T,  Wt[n][k] = sign[n^k][k] * M[n^k]
//   right : out[n] = sum_k sign[k][k^n] * x[k] * M[k^n]      ->                   Wt[n][k] = sign[k][k^n] * M[k^n]
// 2*nb^2 flops for 2*nb*8 bytes per row (AI = 32 flop/B at 256 blades): FP64-math bound, not HBM bound.  tcgen05
// has no f64 kind, so this is the warp-level MMA.  Sums run in tensor-core order; used only in the default (FMA) mode,
// where the bar is 1e-12 relative to the magnitude of the terms -- the strict mode keeps the exact-order table kernel.
//
// Tiling: CTA = 128 batch rows x BN output blades (BN = min(nb, 128); 256 blades = two column blocks, blockIdx.y),
// 8 warps, K chunks of 32 blades double-buffered with cp.async (16-blade chunks: 79.5 % DMMA-pipe active, two block barriers per chunk):
//   Xs[32][128+4]  rows of one blade are contiguous in the SoA batch -> 16-byte coalesced copies; the +4 pad makes the
//                  A-fragment reads (lane -> blade k0 + lane%4, row r0 + lane/4) bank-conflict free
//   Ws[BN][32+4]   Wt rows, k contiguous; same pad argument for the B fragments (lane -> n0 + lane/4, k0 + lane%4)
// A warp owns a (128*BN/8)-element block of the tile as 8x8 accumulator fragments (64 f64 per lane).
#include <cuda_runtime.h>

#include <cstdint>

#include "launch_args.h"
#include "misc_kernels.cuh"

namespace lcc {
namespace {

constexpr int kBM = 128, kBK = 32, kThreadsDmma = 256;
constexpr int kXPitch = kBM + 4, kWPitch = kBK + 4;

__device__ __forceinline__ void cp_async16(void *dst, const void *src, bool valid) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
    const int bytes = valid ? 16 : 0;  // src-size 0: the 16 destination bytes are zero-filled
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void dmma_m8n8k4(double (&c)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

template <int NB>
__global__ void __launch_bounds__(kThreadsDmma, 1)
fixed_product_f64_dmma_kernel(const double *__restrict__ x, int64_t ldx, double *__restrict__ out, int64_t ldo, int64_t n,
                              const double *__restrict__ wt) {
    constexpr int BN = NB < 128 ? NB : 128;          // output blades per CTA
    constexpr int WARPS_N = BN / 32;                  // warps across the blade dimension (32 blades each)
    constexpr int WARPS_M = 8 / WARPS_N;              // warps across the rows
    constexpr int WM = kBM / WARPS_M;                 // rows per warp: 64 (BN = 128) or 32 (BN = 64)
    constexpr int MT = WM / 8, NT = 4;                // 8x8 accumulator tiles per warp
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *xs = reinterpret_cast<double *>(smem_raw);                  // [2][kBK][kXPitch]
    double *ws = xs + 2 * kBK * kXPitch;                                // [2][BN][kWPitch]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t4 = lane & 3;
    const int64_t row0 = (int64_t)blockIdx.x * kBM;
    const int n0 = blockIdx.y * BN;
    const int wm0 = (warp / WARPS_N) * WM, wn0 = (warp % WARPS_N) * 32;

    auto load_chunk = [&](int buf, int k0) {
        double *xd = xs + buf * kBK * kXPitch, *wd = ws + buf * BN * kWPitch;
        // X: kBK blades x 128 rows, two rows per 16-byte copy; rows beyond the padded allocation are zero-filled
        for (int c = tid; c < kBK * (kBM / 2); c += kThreadsDmma) {
            const int k = c / (kBM / 2), m = (c % (kBM / 2)) * 2;
            const int64_t r = row0 + m;
            cp_async16(xd + k * kXPitch + m, x + (int64_t)(k0 + k) * ldx + (r < ldx ? r : 0), r + 1 < ldx);
        }
        // W: BN rows of Wt, kBK consecutive k each
        for (int c = tid; c < BN * (kBK / 2); c += kThreadsDmma) {
            const int nn = c / (kBK / 2), k = (c % (kBK / 2)) * 2;
            cp_async16(wd + nn * kWPitch + k, wt + (int64_t)(n0 + nn) * NB + k0 + k, true);
        }
        cp_async_commit();
    };

    double acc[MT][NT][2];
#pragma unroll
    for (int mi = 0; mi < MT; ++mi)
#pragma unroll
        for (int ni = 0; ni < NT; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;

    constexpr int kChunks = NB / kBK;
    load_chunk(0, 0);
    for (int c = 0; c < kChunks; ++c) {
        if (c + 1 < kChunks) { load_chunk((c + 1) & 1, (c + 1) * kBK); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
        __syncthreads();
        const double *xd = xs + (c & 1) * kBK * kXPitch, *wd = ws + (c & 1) * BN * kWPitch;
#pragma unroll
        for (int kk = 0; kk < kBK; kk += 4) {
            double a[MT], b[NT];
#pragma unroll
            for (int mi = 0; mi < MT; ++mi) a[mi] = xd[(kk + t4) * kXPitch + wm0 + mi * 8 + g];   // A[row g][col t4]
#pragma unroll
            for (int ni = 0; ni < NT; ++ni) b[ni] = wd[(wn0 + ni * 8 + g) * kWPitch + kk + t4];   // B[k t4][col g]
#pragma unroll
            for (int mi = 0; mi < MT; ++mi)
#pragma unroll
                for (int ni = 0; ni < NT; ++ni) dmma_m8n8k4(acc[mi][ni], a[mi], b[ni]);
        }
        __syncthreads();  // the buffer is refilled by the load issued at the top of the next iteration
    }

    // accumulator fragment: rows g, columns 2*t4 and 2*t4 + 1 of each 8x8 tile -> SoA stores (8 consecutive rows per column)
#pragma unroll
    for (int mi = 0; mi < MT; ++mi) {
        const int64_t r = row0 + wm0 + mi * 8 + g;
        if (r >= n) continue;
#pragma unroll
        for (int ni = 0; ni < NT; ++ni) {
            const int col = n0 + wn0 + ni * 8 + 2 * t4;
            __stcs(out + (int64_t)col * ldo + r, acc[mi][ni][0]);
            __stcs(out + (int64_t)(col + 1) * ldo + r, acc[mi][ni][1]);
        }
    }
}

// A persistent variant (one CTA per SM, (tile, chunk) flattened behind a three-stage cp.async ring, one barrier per
// chunk, next tile's loads in flight during the epilogue) was measured and does not pay: 30.71 vs 30.53 TFLOP/s at 2^24
// rows of Cl(8,0,0).  The loop is bound by the DMMA pipe itself -- 128 flop/clk/SM x 148 SMs x 1.965 GHz = 37.2 TFLOP/s,
// of which the kernel sustains 82 % -- not by its prologue or barriers, so the simpler form stays.
template <int NB> cudaError_t launch_dmma_t(const double *x, int64_t ldx, double *out, int64_t ldo, int64_t n, const double *wt, cudaStream_t s) {
    constexpr int BN = NB < 128 ? NB : 128;
    constexpr int smem = (2 * kBK * kXPitch + 2 * BN * kWPitch) * (int)sizeof(double);
    auto kernel = fixed_product_f64_dmma_kernel<NB>;
    const cudaError_t attr = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (attr != cudaSuccess) return attr;
    dim3 grid((unsigned)((n + kBM - 1) / kBM), NB / BN);
    kernel<<<grid, kThreadsDmma, smem, s>>>(x, ldx, out, ldo, n, wt);
    note_kernel_launch();
    return cudaGetLastError();
}

}  // namespace

// out (n x nb, SoA f64) = X * Wt^T for a given nb x nb f64 matrix (fixed_operand_matrix.cu) on the FP64 tensor pipe.
// cudaErrorNotSupported when the batch view is not 16-byte aligned with an even row pitch.
cudaError_t launch_matrix_product_f64_dmma(int nb, const double *wt, const ViewArg &x, const ViewArg &out, cudaStream_t s) {
    if (x.n <= 0) return cudaSuccess;
    if (((uintptr_t)x.data & 15) != 0 || (x.ld & 1) != 0) return cudaErrorNotSupported;  // 16-byte row pairs
    switch (nb) {
        case 64: return launch_dmma_t<64>((const double *)x.data, x.ld, (double *)out.data, out.ld, x.n, wt, s);
        case 128: return launch_dmma_t<128>((const double *)x.data, x.ld, (double *)out.data, out.ld, x.n, wt, s);
        case 256: return launch_dmma_t<256>((const double *)x.data, x.ld, (double *)out.data, out.ld, x.n, wt, s);
    }
    return cudaErrorNotSupported;
}

}  // namespace lcc
