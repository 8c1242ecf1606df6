// fixed_operand_gemm.cu -- K6: product of a large batch with ONE fixed multivector in a large algebra
// (64, 128 or 256 blades), cast as a GEMM on the 5th-generation tensor cores (tcgen05 + TMEM + TMA).
//
// Reference semantics: MultivectorBatch(1 x nb).geometric_product(batch) / batch.geometric_product(1 x nb),
// the broadcast branch of /root/reference/src/batch.rs:365-404 (BASELINE config 5: Cl(8,0,0), 2^24 rows).
//   left  : out[n] = sum_k sign[n^k][k] * M[n^k] * x[k]      ->  Out = X * W,  Wt[n][k] = sign[n^k][k] * M[n^k]
//   right : out[n] = sum_k sign[k][k^n] * x[k] * M[k^n]      ->                Wt[n][k] = sign[k][k^n] * M[k^n]
// This is the one genuinely dense contraction of the path (2*nb^2 flops per 2*nb*4 bytes, AI = 64
// flop/B at 256 blades), so it goes to the tensor pipe; everything else stays a streaming kernel.
//
// Precision: TF32 alone (10-bit mantissa) misses the 1e-5 bar, so every operand is split
// x = x_hi + x_lo (both rounded to tf32), and three MMAs accumulate in fp32 in TMEM:
//   x_hi*w_hi + x_lo*w_hi + x_hi*w_lo        (the dropped x_lo*w_lo term is ~2^-22 relative)
//
// Structure (one persistent CTA per SM, 384 threads, warp specialised):
//   warp 0      TMA producer: X chunk (128 rows x 32 blades, MN-major because batches are blade-major
//               SoA in HBM) and the W_hi / W_lo chunks (nb x 32, K-major), 128B-swizzled, mbarrier tx
//   warps 8-11  splitters: rewrite the X chunk in place as x_hi and write x_lo beside it (element-wise,
//               so the swizzled layout is untouched), fence.proxy.async, arrive
//   warp 1      one elected thread issues 3 tcgen05.mma (M128, N=nb, K8, kind::tf32) per k-step;
//               tcgen05.commit frees the smem stage and, after the last chunk, publishes the accumulator
//   warps 4-7   epilogue: tcgen05.ld 32x32b -> registers -> coalesced SoA stores (lane = row)
// TMEM holds two accumulators (2 x nb columns) so the epilogue of tile t overlaps the MMAs of t+1.
//
// Roofline: algorithmic 2*nb*4 bytes and 3 * 2*nb^2 tensor flops per row (tf32 x3).
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>

#include "launch_args.h"
#include "misc_kernels.cuh"

namespace lcc {
namespace {

constexpr int kBlockM = 128;   // rows per tile = TMEM lanes
constexpr int kThreadsGemm = 384;

// CK = blades per pipeline stage: 32 (W rows are one 128-byte swizzle span) or 16 (64-byte span, twice the stages)
template <int NB, int CK> struct Cfg {
    static constexpr uint32_t kXChunkBytes = kBlockM * CK * 4;
    static constexpr uint32_t kWChunkBytes = NB * CK * 4;
    static constexpr uint32_t kStageBytes = 2 * kXChunkBytes + 2 * kWChunkBytes;
    static constexpr int kStagesFit = (int)((227u * 1024 - 2048) / kStageBytes);
    static constexpr int kStages = kStagesFit > 8 ? 8 : kStagesFit;
    static constexpr uint32_t kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;
    static constexpr int kTmemCols = 2 * NB < 32 ? 32 : 2 * NB;  // two accumulators; power of two >= 32
};

// ------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tcgen05_mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(0u), "r"(0u), "r"(0u), "r"(0u) : "memory");
}
__device__ __forceinline__ void tmem_load_32x32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// UMMA shared-memory matrix descriptor (sm_100): start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46),
// version = 1 [46,48), layout type [61,64) (2 = 128-byte swizzle).
// layout: 2 = 128-byte swizzle of 16-byte units (K-major operands); 1 = 128-byte swizzle of 32-byte
// units, the only layout tcgen05 accepts for MN-major tf32 operands (its atom is 32 MN x 4 K).
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout = 2) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)layout << 61;
    return d;
}
// Instruction descriptor, kind::tf32: D = f32 [4,6)=1, A = tf32 [7,10)=2, B = tf32 [10,13)=2,
// A MN-major [15]=1, B K-major [16]=0, N>>3 [17,23), M>>4 [24,29).
__host__ __device__ constexpr uint32_t instr_desc_tf32(int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (0u << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kBlockM >> 4) << 24);
}

// x = hi + lo with both parts rounded to tf32 (round-to-nearest): |x - hi - lo| <= 2^-22 |x|
__device__ __forceinline__ float tf32_rn(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

// ------------------------------------------------------------------ W split
// Wt given in f64 (fixed_operand_matrix.cu) -> tf32 hi + lo
__global__ void split_w_kernel(const double *__restrict__ wt, int count, float *__restrict__ w_hi, float *__restrict__ w_lo) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= count) return;
    const double w = wt[idx];
    const float hi = tf32_rn((float)w);
    w_hi[idx] = hi;
    w_lo[idx] = tf32_rn((float)(w - (double)hi));
}

// ------------------------------------------------------------------ the GEMM kernel
template <int NB, int CK>
__global__ void __launch_bounds__(kThreadsGemm, 1)
fixed_product_tf32x3_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_whi,
                            const __grid_constant__ CUtensorMap map_wlo, float *__restrict__ out, int64_t ldo, int64_t n) {
    using C = Cfg<NB, CK>;
    constexpr int kChunkK = CK;
    constexpr uint32_t kXChunkBytes = C::kXChunkBytes;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);  // 128B swizzle atoms need 1024-byte alignment
    auto stage_ptr = [&](int s) { return smem + (size_t)s * C::kStageBytes; };
    uint64_t *bars = (uint64_t *)(smem + (size_t)C::kStages * C::kStageBytes);
    uint64_t *full = bars, *ready = bars + C::kStages, *empty = bars + 2 * C::kStages;
    uint64_t *acc_full = bars + 3 * C::kStages, *acc_empty = acc_full + 2;
    uint32_t *tmem_slot = (uint32_t *)(acc_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t num_tiles = (n + kBlockM - 1) / kBlockM;
    constexpr int kChunks = NB / kChunkK;

    if (threadIdx.x == 0) {
        for (int s = 0; s < C::kStages; ++s) { mbar_init(&full[s], 1); mbar_init(&ready[s], 128); mbar_init(&empty[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&acc_full[a], 1); mbar_init(&acc_empty[a], 128); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {  // TMEM allocation is a whole-warp operation; the same warp frees it at the end
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(C::kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            int s = 0; uint32_t ph = 0;
            for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                for (int c = 0; c < kChunks; ++c) {
                    mbar_wait(&empty[s], ph ^ 1);
                    uint8_t *st = stage_ptr(s);
                    mbar_expect_tx(&full[s], kXChunkBytes + 2 * C::kWChunkBytes);
                    // X: dims (row-in-strip 32, blade, strip); box (32, 32, 4) lands as [strip][blade][32 rows]
                    tma_load_3d(st, &map_x, &full[s], 0, c * kChunkK, (int)(tile * (kBlockM / 32)));
                    tma_load_2d(st + 2 * kXChunkBytes, &map_whi, &full[s], c * kChunkK, 0);
                    tma_load_2d(st + 2 * kXChunkBytes + C::kWChunkBytes, &map_wlo, &full[s], c * kChunkK, 0);
                    if (++s == C::kStages) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (one thread) =====
        if (lane == 0) {
            constexpr uint32_t idesc = instr_desc_tf32(NB);
            int s = 0; uint32_t ph = 0; int a = 0; uint32_t pa = 0;
            for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                mbar_wait(&acc_empty[a], pa ^ 1);
                tcgen05_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(a * NB);
                for (int c = 0; c < kChunks; ++c) {
                    mbar_wait(&ready[s], ph);
                    tcgen05_fence_after();
                    const uint32_t xhi = smem_u32(stage_ptr(s)), xlo = xhi + kXChunkBytes;
                    const uint32_t whi = xhi + 2 * kXChunkBytes, wlo = whi + C::kWChunkBytes;
#pragma unroll
                    for (int kk = 0; kk < kChunkK / 8; ++kk) {
                        // A (MN-major, SW128 with 32-byte units): 4-blade K-atoms of 512 bytes (SBO), 8 blades = 1024 bytes
                        // per k-step; the four 32-row strips of the tile are 4096 bytes apart (LBO)
                        constexpr uint32_t kStrip = CK * 128;  // one 32-row strip of the chunk
                        const uint64_t a_hi = umma_desc(xhi + kk * 1024, kStrip, 512, 1), a_lo = umma_desc(xlo + kk * 1024, kStrip, 512, 1);
                        // B (K-major, SW128): 8 tf32 = 32 bytes per k-step inside the 128-byte row; 8-row groups 1024 bytes apart (SBO)
                        constexpr uint32_t kBLayout = CK == 32 ? 2 : 4, kBGroup = CK * 4 * 8;  // 128B / 64B swizzle; 8-row group pitch
                        const uint64_t b_hi = umma_desc(whi + kk * 32, 16, kBGroup, kBLayout), b_lo = umma_desc(wlo + kk * 32, 16, kBGroup, kBLayout);
                        tcgen05_mma_tf32(d_tmem, a_hi, b_hi, idesc, (c | kk) != 0);
                        tcgen05_mma_tf32(d_tmem, a_lo, b_hi, idesc, 1);
                        tcgen05_mma_tf32(d_tmem, a_hi, b_lo, idesc, 1);
                    }
                    tcgen05_commit(&empty[s]);  // stage reusable once these MMAs have read it
                    if (++s == C::kStages) { s = 0; ph ^= 1; }
                }
                tcgen05_commit(&acc_full[a]);
                if (++a == 2) { a = 0; pa ^= 1; }
            }
        }
    } else if (warp >= 4 && warp < 8) {
        // ===== epilogue: TMEM -> registers -> SoA global (lane = row, coalesced) =====
        const int q = warp & 3;
        int a = 0; uint32_t pa = 0;
        for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            mbar_wait(&acc_full[a], pa);
            tcgen05_fence_after();
            const int64_t row = tile * kBlockM + q * 32 + lane;
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * NB);
#pragma unroll 1
            for (int col0 = 0; col0 < NB; col0 += 32) {
                uint32_t v[32];
                tmem_load_32x32(taddr + col0, v);
                if (row < n) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) __stcs(out + (int64_t)(col0 + j) * ldo + row, __uint_as_float(v[j]));
                }
            }
            tcgen05_fence_before();
            mbar_arrive(&acc_empty[a]);
            if (++a == 2) { a = 0; pa ^= 1; }
        }
    } else if (warp >= 8) {
        // ===== splitters: x -> (x_hi in place, x_lo beside it), element-wise on the swizzled tile =====
        const int t = threadIdx.x - 256;  // 0..127
        int s = 0; uint32_t ph = 0;
        for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            for (int c = 0; c < kChunks; ++c) {
                mbar_wait(&full[s], ph);
                float4 *hi = (float4 *)stage_ptr(s);
                float4 *lo = (float4 *)(stage_ptr(s) + kXChunkBytes);
#pragma unroll
                for (int j = 0; j < (int)(kXChunkBytes / 16 / 128); ++j) {
                    const int idx = t + 128 * j;
                    float4 v = hi[idx], h;
                    h.x = tf32_rn(v.x); h.y = tf32_rn(v.y); h.z = tf32_rn(v.z); h.w = tf32_rn(v.w);
                    hi[idx] = h;
                    lo[idx] = make_float4(tf32_rn(v.x - h.x), tf32_rn(v.y - h.y), tf32_rn(v.z - h.z), tf32_rn(v.w - h.w));
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the tensor core
                mbar_arrive(&ready[s]);
                if (++s == C::kStages) { s = 0; ph ^= 1; }
            }
        }
    }

    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(C::kTmemCols) : "memory");
    }
}

// ------------------------------------------------------------------ the GEMM kernel on CTA pairs (cta_group::2)
// With one CTA per tile every M128 x N256 x K8 tf32 MMA pulls 4 KB of A and 8 KB of B out of shared memory in
// 128 cycles; together with the TMA fills and the hi/lo split that is 177 B/clk against a 128 B/clk port, and
// the tensor pipe idles 30 % of the time (profiles/r01_prof_cl8_fixed.summary.json).  Here two CTAs of one
// cluster (the two SMs of a TPC) share each MMA: the tile is 256 rows, every CTA stages ITS 128 rows of X and
// only HALF of the W chunk (its 128 output blades), the leader CTA's elected thread issues
// tcgen05.mma.cta_group::2 (M256 x N=nb x K8) which reads A from both CTAs and one half of B from each, and
// every CTA's TMEM receives the accumulator of its own 128 rows.  Per SM: 125 B/clk of shared-memory traffic
// and half the W traffic from L2.
//
// Barriers (all in each CTA's own shared memory at identical offsets):
//   full[s]      local   TMA bytes of the stage have landed                       (count 1 + tx)
//   ready[s]     LEADER  both CTAs' splitter warps have rewritten their X chunk   (count 8, remote arrives)
//   empty[s]     local   the MMAs that read the stage have completed              (multicast tcgen05.commit)
//   acc_full[a]  local   accumulator a is complete                                (multicast tcgen05.commit)
//   acc_empty[a] LEADER  both CTAs' epilogue warps have drained accumulator a     (count 8, remote arrives)
constexpr int kCK2 = 16;  // blades per stage (64-byte swizzle span)

template <int NB> struct Cfg2 {
    static constexpr uint32_t kXChunkBytes = kBlockM * kCK2 * 4;        // this CTA's 128 rows
    static constexpr uint32_t kWHalfBytes = (NB / 2) * kCK2 * 4;        // this CTA's half of the output blades
    static constexpr uint32_t kStageBytes = 2 * kXChunkBytes + 2 * kWHalfBytes;
    static constexpr int kStagesFit = (int)((227u * 1024 - 2048) / kStageBytes);
    static constexpr int kStages = kStagesFit > 8 ? 8 : kStagesFit;
    static constexpr uint32_t kSmemBytes = kStages * kStageBytes + 1024 + 256;
    static constexpr int kTmemCols = 2 * NB < 32 ? 32 : 2 * NB;
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the barrier at the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint64_t *bar, uint32_t rank) {
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(bar)), "r"(rank));
    // default (CTA-scope release) semantics: a cluster-scope release compiles to MEMBAR.ALL.GPU + ERRBAR per
    // arrival, which cost the first version of this kernel half of its time (profiles/r01_prof_cl8_pair_v1).
    // What the arrivals publish is already ordered without it: the splitters' shared-memory writes are made
    // visible to the tensor core by fence.proxy.async before the arrive, the epilogue's TMEM reads are retired
    // by tcgen05.wait::ld + tcgen05.fence::before_thread_sync.
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP_C:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_C;\n\t"
        "bra WAIT_LOOP_C;\n\t"
        "DONE_C:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tcgen05_commit_pair(uint64_t *bar) {  // one arrival on `bar` in BOTH CTAs of the pair
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void tcgen05_mma_tf32_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(0u) : "memory");
}
// kind::tf32, D f32, A MN-major, B K-major, N = n, M = 256 (two CTAs x 128 lanes)
__host__ __device__ constexpr uint32_t instr_desc_tf32_pair(int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (0u << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
}

template <int NB>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreadsGemm, 1)
fixed_product_tf32x3_pair_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_whi,
                                 const __grid_constant__ CUtensorMap map_wlo, float *__restrict__ out, int64_t ldo, int64_t n) {
    using C = Cfg2<NB>;
    constexpr uint32_t kXChunkBytes = C::kXChunkBytes;
    extern __shared__ uint8_t smem_raw[];
    // identical offsets in both CTAs: the dynamic shared window starts at the same address in every CTA of a kernel
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    auto stage_ptr = [&](int s) { return smem + (size_t)s * C::kStageBytes; };
    uint64_t *bars = (uint64_t *)(smem + (size_t)C::kStages * C::kStageBytes);
    uint64_t *full = bars, *ready = bars + C::kStages, *empty = bars + 2 * C::kStages;
    uint64_t *acc_full = bars + 3 * C::kStages, *acc_empty = acc_full + 2;
    uint32_t *tmem_slot = (uint32_t *)(acc_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int64_t num_tiles = (n + 2 * kBlockM - 1) / (2 * kBlockM);  // 256-row tiles, one per cluster pass
    const int64_t cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
    constexpr int kChunks = NB / kCK2;

    if (threadIdx.x == 0) {
        for (int s = 0; s < C::kStages; ++s) { mbar_init(&full[s], 1); mbar_init(&ready[s], 8); mbar_init(&empty[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&acc_full[a], 1); mbar_init(&acc_empty[a], 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {  // both CTAs of the pair allocate, same warp, same slot address
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(C::kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    cluster_sync_all();  // barriers of BOTH CTAs are initialised before anyone arrives remotely
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer (every CTA: its own rows of X, its own half of W) =====
        if (lane == 0) {
            int s = 0; uint32_t ph = 0;
            for (int64_t tile = cluster_id; tile < num_tiles; tile += num_clusters) {
                const int strip0 = (int)((tile * 2 + rank) * (kBlockM / 32));
                for (int c = 0; c < kChunks; ++c) {
                    mbar_wait(&empty[s], ph ^ 1);
                    uint8_t *st = stage_ptr(s);
                    mbar_expect_tx(&full[s], kXChunkBytes + 2 * C::kWHalfBytes);
                    tma_load_3d(st, &map_x, &full[s], 0, c * kCK2, strip0);
                    tma_load_2d(st + 2 * kXChunkBytes, &map_whi, &full[s], c * kCK2, (int)rank * (NB / 2));
                    tma_load_2d(st + 2 * kXChunkBytes + C::kWHalfBytes, &map_wlo, &full[s], c * kCK2, (int)rank * (NB / 2));
                    if (++s == C::kStages) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: one thread of the LEADER CTA drives the tensor cores of both SMs =====
        if (rank == 0 && lane == 0) {
            constexpr uint32_t idesc = instr_desc_tf32_pair(NB);
            int s = 0; uint32_t ph = 0; int a = 0; uint32_t pa = 0;
            for (int64_t tile = cluster_id; tile < num_tiles; tile += num_clusters) {
                mbar_wait_cluster(&acc_empty[a], pa ^ 1);
                tcgen05_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(a * NB);
                for (int c = 0; c < kChunks; ++c) {
                    mbar_wait_cluster(&ready[s], ph);
                    tcgen05_fence_after();
                    const uint32_t xhi = smem_u32(stage_ptr(s)), xlo = xhi + kXChunkBytes;
                    const uint32_t whi = xhi + 2 * kXChunkBytes, wlo = whi + C::kWHalfBytes;
#pragma unroll
                    for (int kk = 0; kk < kCK2 / 8; ++kk) {
                        constexpr uint32_t kStrip = kCK2 * 128;           // one 32-row strip of the chunk
                        constexpr uint32_t kBGroup = kCK2 * 4 * 8;        // 8-row group pitch of the K-major W chunk
                        const uint64_t a_hi = umma_desc(xhi + kk * 1024, kStrip, 512, 1), a_lo = umma_desc(xlo + kk * 1024, kStrip, 512, 1);
                        const uint64_t b_hi = umma_desc(whi + kk * 32, 16, kBGroup, 4), b_lo = umma_desc(wlo + kk * 32, 16, kBGroup, 4);
                        tcgen05_mma_tf32_pair(d_tmem, a_hi, b_hi, idesc, (c | kk) != 0);
                        tcgen05_mma_tf32_pair(d_tmem, a_lo, b_hi, idesc, 1);
                        tcgen05_mma_tf32_pair(d_tmem, a_hi, b_lo, idesc, 1);
                    }
                    tcgen05_commit_pair(&empty[s]);
                    if (++s == C::kStages) { s = 0; ph ^= 1; }
                }
                tcgen05_commit_pair(&acc_full[a]);
                if (++a == 2) { a = 0; pa ^= 1; }
            }
        }
    } else if (warp >= 4 && warp < 8) {
        // ===== epilogue: this CTA's 128 rows, TMEM -> registers -> SoA global =====
        const int q = warp & 3;
        int a = 0; uint32_t pa = 0;
        for (int64_t tile = cluster_id; tile < num_tiles; tile += num_clusters) {
            mbar_wait(&acc_full[a], pa);
            tcgen05_fence_after();
            const int64_t row = (tile * 2 + rank) * kBlockM + q * 32 + lane;
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * NB);
#pragma unroll 1
            for (int col0 = 0; col0 < NB; col0 += 32) {
                uint32_t v[32];
                tmem_load_32x32(taddr + col0, v);
                if (row < n) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) __stcs(out + (int64_t)(col0 + j) * ldo + row, __uint_as_float(v[j]));
                }
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_remote(&acc_empty[a], 0);
            if (++a == 2) { a = 0; pa ^= 1; }
        }
    } else if (warp >= 8) {
        // ===== splitters: this CTA's X chunk -> (x_hi in place, x_lo beside it) =====
        const int t = threadIdx.x - 256;
        int s = 0; uint32_t ph = 0;
        for (int64_t tile = cluster_id; tile < num_tiles; tile += num_clusters) {
            for (int c = 0; c < kChunks; ++c) {
                mbar_wait(&full[s], ph);
                float4 *hi = (float4 *)stage_ptr(s);
                float4 *lo = (float4 *)(stage_ptr(s) + kXChunkBytes);
#pragma unroll
                for (int j = 0; j < (int)(kXChunkBytes / 16 / 128); ++j) {
                    const int idx = t + 128 * j;
                    float4 v = hi[idx], h;
                    h.x = tf32_rn(v.x); h.y = tf32_rn(v.y); h.z = tf32_rn(v.z); h.w = tf32_rn(v.w);
                    hi[idx] = h;
                    lo[idx] = make_float4(tf32_rn(v.x - h.x), tf32_rn(v.y - h.y), tf32_rn(v.z - h.z), tf32_rn(v.w - h.w));
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive_remote(&ready[s], 0);
                if (++s == C::kStages) { s = 0; ph ^= 1; }
            }
        }
    }

    tcgen05_fence_before();
    cluster_sync_all();  // the leader's last MMAs (which write the peer's TMEM) are committed before anyone frees
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(C::kTmemCols) : "memory");
    }
}

// ------------------------------------------------------------------ host side
using EncodeTiledFn = CUresult (*)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                   const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled() {
    static EncodeTiledFn fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
        return (EncodeTiledFn)p;
    }();
    return fn;
}

template <int NB, int CK>
cudaError_t launch_t(const float *x, int64_t ldx, float *out, int64_t ldo, int64_t n, const float *w_hi, const float *w_lo, cudaStream_t s) {
    EncodeTiledFn enc = encode_tiled();
    if (!enc) return cudaErrorNotSupported;
    CUtensorMap mx, mhi, mlo;
    {   // X (SoA): element (row r = 32*strip + i, blade k) at x[k*ldx + r]
        cuuint64_t dims[3] = {32, (cuuint64_t)NB, (cuuint64_t)(ldx / 32)};
        cuuint64_t strides[2] = {(cuuint64_t)ldx * 4, 128};
        cuuint32_t box[3] = {32, (cuuint32_t)CK, (cuuint32_t)(kBlockM / 32)};
        cuuint32_t estr[3] = {1, 1, 1};
        if (enc(&mx, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void *)x, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return cudaErrorInvalidValue;
    }
    for (int h = 0; h < 2; ++h) {  // Wt[n][k], k contiguous
        cuuint64_t dims[2] = {(cuuint64_t)NB, (cuuint64_t)NB};
        cuuint64_t strides[1] = {(cuuint64_t)NB * 4};
        cuuint32_t box[2] = {(cuuint32_t)CK, (cuuint32_t)NB};
        cuuint32_t estr[2] = {1, 1};
        if (enc(h ? &mlo : &mhi, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void *)(h ? w_lo : w_hi), dims, strides, box, estr,
                CU_TENSOR_MAP_INTERLEAVE_NONE, CK == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return cudaErrorInvalidValue;
    }
    static int sm_count = [] { int d = 0, c = 0; cudaGetDevice(&d); cudaDeviceGetAttribute(&c, cudaDevAttrMultiProcessorCount, d); return c > 0 ? c : 148; }();
    const int64_t tiles = (n + kBlockM - 1) / kBlockM;
    const int grid = (int)(tiles < sm_count ? tiles : sm_count);
    auto kernel = fixed_product_tf32x3_kernel<NB, CK>;
    using C = Cfg<NB, CK>;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::kSmemBytes);
    if (e != cudaSuccess) return e;
    kernel<<<grid, kThreadsGemm, C::kSmemBytes, s>>>(mx, mhi, mlo, out, ldo, n);
    note_kernel_launch();
    return cudaGetLastError();
}

template <int NB>
cudaError_t launch_pair_t(const float *x, int64_t ldx, float *out, int64_t ldo, int64_t n, const float *w_hi, const float *w_lo, cudaStream_t s) {
    EncodeTiledFn enc = encode_tiled();
    if (!enc) return cudaErrorNotSupported;
    CUtensorMap mx, mhi, mlo;
    {
        cuuint64_t dims[3] = {32, (cuuint64_t)NB, (cuuint64_t)(ldx / 32)};
        cuuint64_t strides[2] = {(cuuint64_t)ldx * 4, 128};
        cuuint32_t box[3] = {32, (cuuint32_t)kCK2, (cuuint32_t)(kBlockM / 32)};
        cuuint32_t estr[3] = {1, 1, 1};
        if (enc(&mx, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void *)x, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return cudaErrorInvalidValue;
    }
    for (int h = 0; h < 2; ++h) {  // Wt[n][k], k contiguous; each CTA of a pair takes NB/2 rows
        cuuint64_t dims[2] = {(cuuint64_t)NB, (cuuint64_t)NB};
        cuuint64_t strides[1] = {(cuuint64_t)NB * 4};
        cuuint32_t box[2] = {(cuuint32_t)kCK2, (cuuint32_t)(NB / 2)};
        cuuint32_t estr[2] = {1, 1};
        if (enc(h ? &mlo : &mhi, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void *)(h ? w_lo : w_hi), dims, strides, box, estr,
                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return cudaErrorInvalidValue;
    }
    auto kernel = fixed_product_tf32x3_pair_kernel<NB>;
    using C = Cfg2<NB>;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::kSmemBytes);
    if (e != cudaSuccess) return e;
    // The kernel is persistent with a static round-robin over tiles, so the grid must be ONE wave: ask the
    // runtime how many CTA pairs can be resident at once (GPCs with an odd number of free SMs strand one).
    static const int resident_pairs = [&] {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(2 * 74); cfg.blockDim = dim3(kThreadsGemm); cfg.dynamicSmemBytes = C::kSmemBytes;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        int pairs = 0;
        if (cudaOccupancyMaxActiveClusters(&pairs, kernel, &cfg) != cudaSuccess || pairs <= 0) { cudaGetLastError(); pairs = 64; }
        if (const char *env = getenv("LCC_GEMM_PAIRS")) { const int v = atoi(env); if (v > 0) pairs = v; }
        return pairs;
    }();
    const int64_t tiles = (n + 2 * kBlockM - 1) / (2 * kBlockM);
    const int clusters = (int)(tiles < resident_pairs ? tiles : resident_pairs);
    kernel<<<2 * clusters, kThreadsGemm, C::kSmemBytes, s>>>(mx, mhi, mlo, out, ldo, n);
    note_kernel_launch();
    return cudaGetLastError();
}

}  // namespace

// out (n x nb, SoA f32) = X * Wt^T for a given nb x nb matrix Wt (f64 on the device: a fixed operand's multiplication
// matrix, a masked one, or a whole sandwich -- fixed_operand_matrix.cu), through the tcgen05 tensor cores with the
// 3xTF32 split.  scratch_w: 2*nb*nb floats.  cudaErrorNotSupported: no tensor-map encoder, or a view the map cannot cover.
cudaError_t launch_matrix_product_tf32x3(int nb, const double *wt64, const ViewArg &x, const ViewArg &out, float *scratch_w, cudaStream_t s) {
    if (x.n <= 0) return cudaSuccess;
    if (!encode_tiled() || x.ld % 32 != 0 || out.ld % 4 != 0 || ((uintptr_t)x.data & 15) != 0 || ((uintptr_t)out.data & 15) != 0) return cudaErrorNotSupported;
    float *w_hi = scratch_w, *w_lo = scratch_w + (size_t)nb * nb;
    split_w_kernel<<<(nb * nb + 255) / 256, 256, 0, s>>>(wt64, nb * nb, w_hi, w_lo);
    note_kernel_launch();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    // CTA pairs (cta_group::2) by default; LCC_GEMM_PAIR=0 selects the one-CTA-per-tile kernel
    static const bool pair = [] { const char *e = getenv("LCC_GEMM_PAIR"); return !(e && e[0] == '0'); }();
    if (pair) {
        switch (nb) {
            case 64: return launch_pair_t<64>((const float *)x.data, x.ld, (float *)out.data, out.ld, x.n, w_hi, w_lo, s);
            case 128: return launch_pair_t<128>((const float *)x.data, x.ld, (float *)out.data, out.ld, x.n, w_hi, w_lo, s);
            case 256: return launch_pair_t<256>((const float *)x.data, x.ld, (float *)out.data, out.ld, x.n, w_hi, w_lo, s);
        }
        return cudaErrorNotSupported;
    }
    static const int chunk = [] { const char *e = getenv("LCC_GEMM_CHUNK"); return e && atoi(e) == 32 ? 32 : 16; }();
#define LCC_GEMM_CASE(NBV)                                                                                                   \
    case NBV:                                                                                                                \
        return chunk == 32 ? launch_t<NBV, 32>((const float *)x.data, x.ld, (float *)out.data, out.ld, x.n, w_hi, w_lo, s)   \
                           : launch_t<NBV, 16>((const float *)x.data, x.ld, (float *)out.data, out.ld, x.n, w_hi, w_lo, s);
    switch (nb) {
        LCC_GEMM_CASE(64)
        LCC_GEMM_CASE(128)
        LCC_GEMM_CASE(256)
    }
#undef LCC_GEMM_CASE
    return cudaErrorNotSupported;
}

}  // namespace lcc
