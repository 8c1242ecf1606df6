// abi_internal.hpp -- helpers shared by the translation units that implement include/lcc_b200.h
// (lcc_abi.cu: validation + dispatch; host_transfer.cu: host<->device pipelines and the pinned host
// cache; shard_group.cu: multi-GPU row sharding + NCCL).  Internal; nothing here crosses the C ABI.
#pragma once
#include <cstdint>
#include <string>

#include <cuda_runtime.h>

#include "../../include/lcc_b200.h"
#include "algebra_host.hpp"
#include "kernels/misc_kernels.cuh"

namespace lcc {
namespace abi {

// thread-local error message + status (lcc_last_error)
lcc_status fail(lcc_status code, const char *fmt, ...) __attribute__((format(printf, 2, 3)));
lcc_status cuda_fail(cudaError_t e, const char *what);

#define LCC_CUDA(call)                                                  \
    do {                                                                \
        cudaError_t e__ = (call);                                       \
        if (e__ != cudaSuccess) return ::lcc::abi::cuda_fail(e__, #call); \
    } while (0)
#define LCC_TRY(expr)                               \
    do {                                            \
        lcc_status s__ = (expr);                    \
        if (s__ != LCC_OK) return s__;              \
    } while (0)

struct TransferEngine;  // host_transfer.cu

struct DeviceState {
    int device = 0;
    cudaStream_t stream = nullptr;  // compute
    cudaStream_t copy_in = nullptr, copy_out = nullptr;
    TransferEngine *transfer = nullptr;  // created on first host<->device pipeline call
    bool ready = false;
};
lcc_status device_state(DeviceState **out);                    // of the current device
lcc_status resolve_stream(void *stream, cudaStream_t *out);    // NULL -> the library's stream

inline size_t elem_size(int dtype) { return dtype == LCC_F32 ? 4 : 8; }
lcc_status check_view(const lcc_algebra *alg, const lcc_view *v, const char *name, bool allow_null_data = false);
inline ViewArg arg_of(const lcc_view *v) { return ViewArg{v->data, v->n, v->ld, v->dtype, v->layout}; }
lcc_status broadcast_rows(int64_t na, int64_t nbr, int64_t *n);
std::string sig_str(const lcc_algebra *a);
int64_t padded_ld(int64_t n);

// RAII for stream-ordered temporaries
struct PoolBuffer {
    void *ptr = nullptr; cudaStream_t stream = nullptr;
    ~PoolBuffer() { if (ptr) cudaFreeAsync(ptr, stream); }
    lcc_status alloc(size_t bytes, cudaStream_t s) {
        stream = s;
        cudaError_t e = cudaMallocAsync(&ptr, bytes ? bytes : 16, s);
        if (e != cudaSuccess) { ptr = nullptr; return cuda_fail(e, "cudaMallocAsync"); }
        return LCC_OK;
    }
};

// host_transfer.cu options (lcc_set_option)
int64_t option_host_chunk_bytes();
void set_option_host_chunk_bytes(int64_t v);
int64_t option_pinned_output_min_bytes();
void set_option_pinned_output_min_bytes(int64_t v);

// LCC_LOG=1: one line per C-ABI compute entry on stderr (the reference has no logging; this is the
// engine's own trace).  NVTX ranges are pushed around the same entries when libnvToolsExt / nvtx3 is usable.
bool log_enabled();
void log_line(const char *fmt, ...) __attribute__((format(printf, 1, 2)));
struct TraceRange {  // NVTX range + optional log line for one entry point
    explicit TraceRange(const char *name);
    ~TraceRange();
};

}  // namespace abi
}  // namespace lcc
