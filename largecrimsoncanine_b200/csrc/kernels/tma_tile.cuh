// tma_tile.cuh -- TMA staging of array-of-structures tiles for the streaming kernels (sm_100a).
//
// The reference's batch layout is AoS: row r of an (N, num_blades) row-major array holds one
// multivector (/root/reference/src/batch.rs:41-47).  A thread that owns one row reads 32-256
// contiguous bytes, so a warp-wide 128-bit access touches 32 different 16-byte pieces that are a
// whole row apart: every instruction opens 32 sectors and uses half (or less) of each.  Here the
// tile of kRows consecutive rows -- one contiguous span of HBM -- is moved by the TMA engine
// (cp.async.bulk.tensor) into shared memory as full-line bursts, the threads read their rows from
// shared memory, write the result rows back over them, and one thread hands the tile to the TMA
// engine again for the store.  The hardware swizzle of the tensor map (XOR of the 16-byte chunk
// index with the 128-byte line index) makes the row-per-thread shared-memory accesses conflict free.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>

namespace lcc {

// ---------------------------------------------------------------- tile geometry for a row of RB bytes
template <int RB> struct AosTmaTile {
    static_assert(RB >= 32 && (RB & (RB - 1)) == 0, "rows of 32..256 bytes");
    static constexpr int kSpan = RB >= 128 ? 128 : RB;  // bytes per TMA box row == swizzle span
    static constexpr int kSub = RB / kSpan;             // boxes per tile (a 256-byte row is two 128-byte halves)
    static constexpr int kRows = RB >= 256 ? 64 : (RB >= 128 ? 128 : 256);  // rows per tile == threads per block
    static constexpr int kBytes = kRows * RB;           // 8-16 KB per operand tile
    static constexpr uint32_t kSwzMask = kSpan == 128 ? 7u : (kSpan == 64 ? 3u : 1u);
    // Shared-memory byte offset of 16-byte chunk j of row t: boxes are stored one after the other, each
    // [kRows][kSpan] with the tensor map's swizzle applied to the address bits (tile base 1024-aligned).
    static __device__ __forceinline__ uint32_t chunk_offset(int t, int j) {
        const uint32_t byte = (uint32_t)j * 16u;
        const uint32_t off = (byte / kSpan) * (uint32_t)(kRows * kSpan) + (uint32_t)t * kSpan + (byte % kSpan);
        return off ^ (((off >> 7) & kSwzMask) << 4);
    }
};

// ---------------------------------------------------------------- PTX wrappers
namespace tma {
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
        "LCC_TMA_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra LCC_TMA_DONE;\n\t"
        "bra LCC_TMA_WAIT;\n\t"
        "LCC_TMA_DONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// Streaming data is touched exactly once: with `stream_hint` the copies carry an L2 evict-first policy so the
// 17-68 GB that pass through do not displace each other's lines before they are written back / consumed.
__device__ __forceinline__ uint64_t evict_first_policy() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
// global -> shared, completion on the mbarrier; c0 = element coordinate inside the row, c1 = row
__device__ __forceinline__ void load_2d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, bool stream_hint = false) {
    if (stream_hint) {
        asm volatile(
            "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;"
            ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "l"(evict_first_policy()) : "memory");
    } else {
        asm volatile(
            "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
            ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
    }
}
// shared -> global (rows beyond the tensor's extent are clipped by the hardware)
__device__ __forceinline__ void store_2d(const CUtensorMap *map, const void *src, int c0, int c1, bool stream_hint = false) {
    if (stream_hint) {
        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%2, %3}], [%1], %4;"
                     ::"l"(map), "r"(smem_u32(src)), "r"(c0), "r"(c1), "l"(evict_first_policy()) : "memory");
    } else {
        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                     ::"l"(map), "r"(smem_u32(src)), "r"(c0), "r"(c1) : "memory");
    }
}
// 3-D forms (K8 on reference-layout batches: one box = {128 bytes of blades, rows, chunks})
__device__ __forceinline__ void load_3d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void store_3d(const CUtensorMap *map, const void *src, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                 ::"l"(map), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void store_commit_and_wait_read() {
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
// generic-proxy writes to shared memory -> visible to the TMA engine
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void prefetch_map(const CUtensorMap *map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
}  // namespace tma

// ---------------------------------------------------------------- host: tensor map of an AoS batch
// Describes an (n, nb) row-major array with row pitch ld elements as a 2-D tensor and a box of
// (span bytes, rows) with the matching swizzle.  Returns false when the driver entry point is missing
// or the array does not meet TMA's alignment rules (16-byte base and pitch) -- callers then use the
// direct-load kernels.
bool make_aos_tensor_map(CUtensorMap *map, const void *base, int elem_bytes, int nb, int64_t ld, int64_t n,
                         int span_bytes, int rows);

// Blade-major (SoA) batch: element (row r, blade k) at base[k*ld + r].  Tensor = (n rows, nb blades) with
// the rows contiguous; box = (rows, nb), no swizzle (a thread reads consecutive rows of one blade, so
// neighbouring threads already hit neighbouring banks).  Rows beyond n are zero-filled / clipped.
// swizzle128 (K8): box rows of exactly 128 bytes with the hardware 128-byte swizzle -- 16-byte chunk c of blade k lands at
// chunk c ^ (k & 7) of its 128-byte line, so that accesses ACROSS blades at fixed rows spread over the banks.
// l2_promotion: bytes the L2 fetches from DRAM per request (64 / 128 / 256; 0 = none).
bool make_soa_tensor_map(CUtensorMap *map, const void *base, int elem_bytes, int nb, int64_t ld, int64_t n, int rows,
                         bool swizzle128 = false, int l2_promotion = 128);

// Reference-layout (AoS) batch seen as (128 bytes of blades, n rows, nb * elem_bytes / 128 chunks): a box of
// (128 bytes, rows, all chunks) is `rows` whole multivectors -- one contiguous span of HBM when ld == nb -- and lands in
// shared memory as [chunk][row][128 bytes] with the hardware 128-byte swizzle (K8, matrix_rep_kernel.cu).  Rows of at least
// 128 bytes only; any row count (whole rows are clipped / zero-filled at the end).
bool make_aos_chunked_tensor_map(CUtensorMap *map, const void *base, int elem_bytes, int nb, int64_t ld, int64_t n, int rows);

}  // namespace lcc
