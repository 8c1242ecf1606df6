// tma_passthrough_probe.cu -- which tile shape lets a persistent TMA ring stream a 256-blade batch at the HBM rate?
//
// K8 (matrix_rep_kernel.cu) keeps whole 256-blade rows in shared memory, so with the blade-major (SoA) layout a unit of
// 128 bytes of rows is 256 separate 128-byte pieces, one per blade row (64 MB apart at 2^24 rows).  Its pass-through mode
// (LCC_K8_DEBUG=1) reaches only 4.5 TB/s where the 16-blade kernels (512 contiguous bytes per blade row) reach 6.9.
// This probe moves the same 2^24 x 256 f32 batch through shared memory untouched with one elected thread per CTA and a
// ring of stages, and varies only the shape of a stage:
//   soa   B adjacent boxes of (TR rows x 256 blades) per stage      -> B * TR * 4 contiguous bytes per blade row
//   aos   row-major (n, 256) batch, one 3-D box of (32 blades x ROWS rows x 8 chunks) -> ROWS KB contiguous
//   bulk  1-D cp.async.bulk of contiguous stage-sized pieces (upper bound of the ring itself)
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o build/tma_probe scripts/tma_passthrough_probe.cu
// Run (GPU box): build/tma_probe  -> one JSON line per shape on stdout.
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "PROBE_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra PROBE_DONE;\n\t"
        "bra PROBE_WAIT;\n\t"
        "PROBE_DONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

struct Shape {
    int mode;          // 0 soa (2-D boxes), 1 aos (3-D box), 2 bulk (1-D)
    int boxes;         // soa: adjacent boxes per stage
    int tr;            // soa: rows (elements) per box; aos: rows per stage
    int ring;          // stages in flight
    int order;         // 0 interleaved over CTAs, 1 contiguous range per CTA
    int hint;          // 1: L2 evict_first policy on loads and stores (aos / 3-D mode only)
    int dynamic;       // 1: persistent CTAs take their next stage from a global counter (order is ignored)
    unsigned long long *counter;
    uint32_t stage_bytes, box_bytes;
    long long stages;  // total
};

__global__ void __launch_bounds__(32) pass_kernel(const __grid_constant__ CUtensorMap mx, const __grid_constant__ CUtensorMap mo,
                                                     const float *x, float *out, Shape p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + (size_t)p.ring * p.stage_bytes);
    if (threadIdx.x != 0) return;
    for (int s = 0; s < p.ring; ++s) mbar_init(&full[s], 1);
    uint64_t policy = 0;
    if (p.hint) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
    const long long per_cta = (p.stages + gridDim.x - 1) / gridDim.x;
    const long long first = p.order ? (long long)blockIdx.x * per_cta : (long long)blockIdx.x;
    const long long step = p.order ? 1 : (long long)gridDim.x;
    const long long mine = p.order ? (first >= p.stages ? 0 : (p.stages - first < per_cta ? p.stages - first : per_cta))
                                   : (p.stages > blockIdx.x ? (p.stages - blockIdx.x + gridDim.x - 1) / gridDim.x : 0);
    auto load = [&](long long st, int s) {
        uint8_t *dst = smem + (size_t)s * p.stage_bytes;
        mbar_expect_tx(&full[s], p.stage_bytes);
        if (p.mode == 0) {
            for (int j = 0; j < p.boxes; ++j)
                asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                             ::"r"(smem_u32(dst + (size_t)j * p.box_bytes)), "l"(&mx), "r"(smem_u32(&full[s])), "r"((int)((st * p.boxes + j) * p.tr)), "r"(0) : "memory");
        } else if (p.mode == 1 && p.hint) {
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4, %5}], [%2], %6;"
                         ::"r"(smem_u32(dst)), "l"(&mx), "r"(smem_u32(&full[s])), "r"(0), "r"((int)(st * p.tr)), "r"(0), "l"(policy) : "memory");
        } else if (p.mode == 1) {
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                         ::"r"(smem_u32(dst)), "l"(&mx), "r"(smem_u32(&full[s])), "r"(0), "r"((int)(st * p.tr)), "r"(0) : "memory");
        } else {
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(smem_u32(dst)), "l"((const uint8_t *)x + (size_t)st * p.stage_bytes), "r"(p.stage_bytes), "r"(smem_u32(&full[s])) : "memory");
        }
    };
    auto store = [&](long long st, int s) {
        const uint8_t *src = smem + (size_t)s * p.stage_bytes;
        if (p.mode == 0) {
            for (int j = 0; j < p.boxes; ++j)
                asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                             ::"l"(&mo), "r"(smem_u32(src + (size_t)j * p.box_bytes)), "r"((int)((st * p.boxes + j) * p.tr)), "r"(0) : "memory");
        } else if (p.mode == 1 && p.hint) {
            asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%2, %3, %4}], [%1], %5;"
                         ::"l"(&mo), "r"(smem_u32(src)), "r"(0), "r"((int)(st * p.tr)), "r"(0), "l"(policy) : "memory");
        } else if (p.mode == 1) {
            asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                         ::"l"(&mo), "r"(smem_u32(src)), "r"(0), "r"((int)(st * p.tr)), "r"(0) : "memory");
        } else {
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                         ::"l"((uint8_t *)out + (size_t)st * p.stage_bytes), "r"(smem_u32(src)), "r"(p.stage_bytes) : "memory");
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    };
    if (p.dynamic) {
        long long in_slot[16];
        auto next = [&]() { return (long long)atomicAdd(p.counter, 1ull); };
        for (int s = 0; s < p.ring; ++s) {
            in_slot[s] = next();
            if (in_slot[s] < p.stages) load(in_slot[s], s);
        }
        for (long long it = 0;; ++it) {
            const int s = (int)(it % p.ring);
            if (in_slot[s] >= p.stages) break;  // stage numbers are handed out in increasing order: nothing after this slot either
            mbar_wait(&full[s], (uint32_t)((it / p.ring) & 1));
            store(in_slot[s], s);
            if (it >= 1) {
                const int ps = (int)((it - 1) % p.ring);
                asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                in_slot[ps] = next();
                if (in_slot[ps] < p.stages) load(in_slot[ps], ps);
            }
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        return;
    }
    for (long long k = 0; k < mine && k < p.ring; ++k) load(first + k * step, (int)k);
    for (long long it = 0; it < mine; ++it) {
        const int s = (int)(it % p.ring);
        mbar_wait(&full[s], (uint32_t)((it / p.ring) & 1));
        store(first + it * step, s);
        if (p.ring == 1) {  // no second slot: wait for this store's read, then refill
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            if (it + 1 < mine) load(first + (it + 1) * step, 0);
        } else if (it >= 1) {  // the previous stage's store has been read out of shared memory: refill its slot
            asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            if (it - 1 + p.ring < mine) load(first + (it - 1 + p.ring) * step, (int)((it - 1) % p.ring));
        }
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

using EncodeFn = CUresult (*)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                              const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                              CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char **argv) {
    const long long n = argc > 1 ? atoll(argv[1]) : (1ll << 24);
    const int nb = 256;
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    EncodeFn enc = (EncodeFn)fn;
    float *x, *out;
    const size_t bytes = (size_t)n * nb * 4;
    CK(cudaMalloc(&x, bytes));
    CK(cudaMalloc(&out, bytes));
    CK(cudaMemset(x, 1, bytes));
    CK(cudaMemset(out, 0, bytes));
    int sms = 148;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    CK(cudaFuncSetAttribute(pass_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    unsigned long long *counter;
    CK(cudaMalloc(&counter, 8));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    struct Case { const char *name; int mode, boxes, tr, ring, order, swz, promo, per_sm, hint, dynamic; };  // per_sm 0: one CTA per stage (not persistent)
    const Case cases[] = {
        {"aos_32rows_ring6_persistent_static", 1, 1, 32, 6, 0, 1, 128, 1, 0, 0},
        {"aos_32rows_ring6_persistent_dynamic", 1, 1, 32, 6, 0, 1, 128, 1, 0, 1},
        {"aos_32rows_ring4_persistent_dynamic", 1, 1, 32, 4, 0, 1, 128, 1, 0, 1},
        {"aos_32rows_ring3_persistent_dynamic", 1, 1, 32, 3, 0, 1, 128, 1, 0, 1},
        {"aos_16rows_ring12_persistent_dynamic", 1, 1, 16, 12, 0, 1, 128, 1, 0, 1},
        {"aos_32rows_one_cta_per_stage", 1, 1, 32, 1, 0, 1, 128, 0, 0, 0},
        {"soa_128Bx256_ring6_static", 0, 1, 32, 6, 1, 1, 256, 1, 0, 0},
        {"soa_128Bx256_ring6_dynamic", 0, 1, 32, 6, 1, 1, 256, 1, 0, 1},
        {"soa_128Bx256_one_cta_per_stage", 0, 1, 32, 1, 0, 1, 256, 0, 0, 0},
        {"soa_256Bx256_noswz_ring3_dynamic", 0, 1, 64, 3, 0, 0, 256, 1, 0, 1},
        {"soa_256Bx256_noswz_one_cta_per_stage", 0, 1, 64, 1, 0, 0, 256, 0, 0, 0},
        {"bulk_32KB_ring6_persistent_dynamic", 2, 1, 32, 6, 0, 0, 0, 1, 0, 1},
    };
    for (const Case &c : cases) {
        CUtensorMap mx, mo;
        memset(&mx, 0, sizeof mx);
        memset(&mo, 0, sizeof mo);
        Shape p{};
        p.mode = c.mode; p.boxes = c.boxes; p.tr = c.tr; p.ring = c.ring; p.order = c.order; p.hint = c.hint; p.dynamic = c.dynamic; p.counter = counter;
        const CUtensorMapL2promotion promo = c.promo >= 256 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_L2_128B;
        bool ok = true;
        if (c.mode == 0) {
            const cuuint64_t dims[2] = {(cuuint64_t)n, (cuuint64_t)nb};
            const cuuint64_t strides[1] = {(cuuint64_t)n * 4};
            const cuuint32_t box[2] = {(cuuint32_t)c.tr, (cuuint32_t)nb}, es[2] = {1, 1};
            const CUtensorMapSwizzle swz = c.swz == 1 ? CU_TENSOR_MAP_SWIZZLE_128B : (c.swz == 2 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE);
            for (CUtensorMap *m : {&mx, &mo})
                ok &= enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, m == &mx ? (void *)x : (void *)out, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, swz, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
            p.box_bytes = (uint32_t)c.tr * 4 * nb;
            p.stage_bytes = p.box_bytes * c.boxes;
            p.stages = n / ((long long)c.tr * c.boxes);
        } else if (c.mode == 1) {
            // (n, 256) row-major as (32 blades, n rows, 8 chunks of 32 blades): shared-memory image [chunk][row][128 B swizzled]
            const cuuint64_t dims[3] = {32, (cuuint64_t)n, 8};
            const cuuint64_t strides[2] = {(cuuint64_t)nb * 4, 128};
            const cuuint32_t box[3] = {32, (cuuint32_t)c.tr, 8}, es[3] = {1, 1, 1};
            for (CUtensorMap *m : {&mx, &mo})
                ok &= enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, m == &mx ? (void *)x : (void *)out, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
            p.stage_bytes = p.box_bytes = (uint32_t)c.tr * nb * 4;
            p.stages = n / c.tr;
        } else {
            p.stage_bytes = p.box_bytes = (uint32_t)c.tr * 1024;
            p.stages = (long long)(bytes / p.stage_bytes);
        }
        if (!ok) { printf("{\"case\": \"%s\", \"error\": \"tensor map\"}\n", c.name); continue; }
        const size_t smem = (size_t)p.ring * p.stage_bytes + 256;
        float best = 1e30f, sum = 0;
        const int reps = 5;
        for (int r = 0; r < reps + 2; ++r) {
            CK(cudaMemsetAsync(counter, 0, 8));
            CK(cudaEventRecord(e0));
            const long long grid = c.per_sm ? (long long)sms * c.per_sm : p.stages;
            pass_kernel<<<(unsigned)grid, 32, smem>>>(mx, mo, x, out, p);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            CK(cudaGetLastError());
            float ms;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            if (r >= 2) { sum += ms; if (ms < best) best = ms; }
        }
        // spot check: the output equals the input
        uint32_t hx[4], ho[4];
        const size_t off = bytes / 3 / 16 * 16;
        CK(cudaMemcpy(hx, (uint8_t *)x + off, 16, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(ho, (uint8_t *)out + off, 16, cudaMemcpyDeviceToHost));
        const bool same = memcmp(hx, ho, 16) == 0;
        CK(cudaMemset(out, 0, bytes));
        printf("{\"case\": \"%s\", \"stage_bytes\": %u, \"ring\": %d, \"ms_mean\": %.3f, \"ms_min\": %.3f, \"GBps\": %.1f, \"copied\": %s}\n", c.name,
               p.stage_bytes, p.ring, sum / reps, best, 2.0 * bytes / (sum / reps) / 1e6, same ? "true" : "false");
        fflush(stdout);
    }
    return 0;
}
