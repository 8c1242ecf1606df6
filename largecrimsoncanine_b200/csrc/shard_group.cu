// shard_group.cu -- multi-GPU row sharding behind the C ABI.
//
// Every batched operation of the reference is row-independent (`for row in 0..n`, /root/reference/src/batch.rs:384),
// so a batch of N rows is split into contiguous row blocks, one per GPU, and each block runs the single-GPU kernels
// on its own non-blocking stream: no halo, no data-path collective.  The ONE exchange of the path is the batch
// sum / mean (new surface, no reference counterpart): per-GPU fused partial (num_blades values) ->
// ncclAllReduce(count = num_blades) enqueued right behind it, no host synchronisation in between.
//
// A group is either
//   * single-process: all local devices joined with ncclCommInitAll, one host thread issuing the asynchronous launches
//     of every GPU (what a Rust host crate binding this header gets), or
//   * one member of a multi-process job (one rank per GPU, e.g. torchrun): ncclCommInitRank with a unique id that
//     the launcher distributes.  Same entry points, local_size == 1.
// NCCL is loaded with dlopen at first use (libnccl.so.2: the copy already in the process if there is one), so the
// library itself has no link-time dependency on it and one-device groups work without it.
#include <dlfcn.h>
#include <nccl.h>

#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "abi_internal.hpp"

using namespace lcc;
using namespace lcc::abi;

namespace {

// ------------------------------------------------------------------ NCCL through dlopen
struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GetVersion)(int *) = nullptr;
    std::string error;
};

NcclApi *nccl_api() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char *env = getenv("LCC_NCCL_LIB");
        const char *names[] = {env && *env ? env : "libnccl.so.2", "libnccl.so.2", "libnccl.so"};
        for (const char *n : names) {
            api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (api.handle) break;
        }
        if (!api.handle) { api.error = std::string("cannot load libnccl.so.2: ") + (dlerror() ? dlerror() : "not found"); return; }
        auto sym = [&](const char *name) { void *p = dlsym(api.handle, name); if (!p && api.error.empty()) api.error = std::string("libnccl lacks ") + name; return p; };
        api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
        api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
        api.CommInitAll = (decltype(api.CommInitAll))sym("ncclCommInitAll");
        api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
        api.AllReduce = (decltype(api.AllReduce))sym("ncclAllReduce");
        api.AllGather = (decltype(api.AllGather))sym("ncclAllGather");
        api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart");
        api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
        api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
        api.GetVersion = (decltype(api.GetVersion))sym("ncclGetVersion");
    });
    return &api;
}

lcc_status nccl_fail(ncclResult_t r, const char *what) {
    NcclApi *n = nccl_api();
    return fail(LCC_ERR_CUDA, "%s: NCCL error %d (%s)", what, (int)r, n->GetErrorString ? n->GetErrorString(r) : "?");
}
#define LCC_NCCL(call)                                                  \
    do {                                                                \
        ncclResult_t r__ = (call);                                      \
        if (r__ != ncclSuccess) return nccl_fail(r__, #call);           \
    } while (0)

lcc_status need_nccl(NcclApi **out) {
    NcclApi *n = nccl_api();
    if (!n->handle || !n->error.empty()) return fail(LCC_ERR_UNSUPPORTED, "multi-GPU reduction needs NCCL: %s", n->error.c_str());
    *out = n;
    return LCC_OK;
}

struct DeviceGuard {  // kernels launch on the CURRENT device: select the shard's GPU, restore on exit
    int prev = -1;
    explicit DeviceGuard(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); }
    ~DeviceGuard() { int cur = -1; cudaGetDevice(&cur); if (prev >= 0 && cur != prev) cudaSetDevice(prev); }
};

constexpr int kSlots = 8;         // ring of asynchronous reductions in flight
constexpr int kSlotElems = 256;   // num_blades <= 256

struct ShardDevice {
    int device = 0;
    cudaStream_t compute = nullptr, comm = nullptr;
    ncclComm_t nccl = nullptr;
    double *slot[kSlots + 1] = {};                  // [kSlots] = the synchronous entry points' buffer
    cudaEvent_t produced[kSlots + 1] = {}, reduced[kSlots + 1] = {};
    bool reduced_valid[kSlots + 1] = {};
};

__global__ void divide_kernel(void *buf, int nb, int is_f32, double divisor) {
    const int k = threadIdx.x;
    if (k >= nb) return;
    if (is_f32) ((float *)buf)[k] = (float)((double)((float *)buf)[k] / divisor);
    else ((double *)buf)[k] = ((double *)buf)[k] / divisor;
}

}  // namespace

struct lcc_shard_group {
    std::vector<ShardDevice> dev;
    int world = 1, first_rank = 0;  // ranks of the local devices: first_rank .. first_rank + dev.size() - 1
};

namespace {

lcc_status init_device(ShardDevice &d, int device) {
    d.device = device;
    DeviceGuard g(device);
    int lo = 0, hi = 0;
    LCC_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    LCC_CUDA(cudaStreamCreateWithFlags(&d.compute, cudaStreamNonBlocking));
    LCC_CUDA(cudaStreamCreateWithPriority(&d.comm, cudaStreamNonBlocking, hi));  // the tiny collective must never queue behind a 3 ms kernel
    for (int s = 0; s <= kSlots; ++s) {
        LCC_CUDA(cudaMalloc(&d.slot[s], kSlotElems * sizeof(double)));
        LCC_CUDA(cudaEventCreateWithFlags(&d.produced[s], cudaEventDisableTiming));
        LCC_CUDA(cudaEventCreateWithFlags(&d.reduced[s], cudaEventDisableTiming));
    }
    DeviceState *st = nullptr;  // creates the library's per-device state (pool threshold, copy streams) for this GPU
    LCC_TRY(device_state(&st));
    return LCC_OK;
}

lcc_status check_views(const lcc_shard_group *g, const lcc_algebra *alg, const lcc_view *views, const char *name) {
    if (!g) return fail(LCC_ERR_VALUE, "shard group is NULL");
    if (!views) return fail(LCC_ERR_VALUE, "%s: view array is NULL", name);
    for (size_t i = 0; i < g->dev.size(); ++i) LCC_TRY(check_view(alg, &views[i], name));
    return LCC_OK;
}

// sums of one shard into buf (num_blades elements of the batch dtype), zero rows -> zeros
lcc_status shard_partial_sum(const lcc_algebra *alg, const lcc_view *x, void *buf, cudaStream_t s) {
    if (x->n == 0) { LCC_CUDA(cudaMemsetAsync(buf, 0, (size_t)alg->nb * elem_size(x->dtype), s)); return LCC_OK; }
    return lcc_batch_sum(alg, x, buf, (void *)s);
}

// all-reduce slot `s` of every local device in place on `stream_of(device)`, then the mean's division
lcc_status all_reduce_slot(lcc_shard_group *g, int nb, int dtype, int s, bool on_comm, int64_t mean_divisor) {
    if (g->world > 1) {
        NcclApi *n = nullptr;
        LCC_TRY(need_nccl(&n));
        LCC_NCCL(n->GroupStart());
        for (auto &d : g->dev) {
            ncclResult_t r = n->AllReduce(d.slot[s], d.slot[s], (size_t)nb, dtype == LCC_F32 ? ncclFloat32 : ncclFloat64, ncclSum, d.nccl,
                                          on_comm ? d.comm : d.compute);
            if (r != ncclSuccess) { n->GroupEnd(); return nccl_fail(r, "ncclAllReduce"); }
        }
        LCC_NCCL(n->GroupEnd());
    }
    if (mean_divisor > 0)
        for (auto &d : g->dev) {
            DeviceGuard guard(d.device);
            divide_kernel<<<1, 256, 0, on_comm ? d.comm : d.compute>>>(d.slot[s], nb, dtype == LCC_F32, (double)mean_divisor);
            note_kernel_launch();
            LCC_CUDA(cudaGetLastError());
        }
    return LCC_OK;
}

}  // namespace

// ================================================================== C ABI
extern "C" {

lcc_status lcc_shard_rows(int64_t n_total, int rank, int world, int64_t *begin, int64_t *end) {
    if (world <= 0 || rank < 0 || rank >= world) return fail(LCC_ERR_VALUE, "bad rank/world: %d/%d", rank, world);
    if (n_total < 0) return fail(LCC_ERR_VALUE, "negative batch size");
    // contiguous blocks [rank*N/W, (rank+1)*N/W): they tile [0, N) exactly and differ by at most one row
    const __int128 n = n_total;
    if (begin) *begin = (int64_t)(n * rank / world);
    if (end) *end = (int64_t)(n * (rank + 1) / world);
    return LCC_OK;
}

lcc_status lcc_shard_unique_id(void *id128) {
    if (!id128) return fail(LCC_ERR_VALUE, "id buffer is NULL");
    NcclApi *n = nullptr;
    LCC_TRY(need_nccl(&n));
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    LCC_NCCL(n->GetUniqueId(&id));
    memcpy(id128, &id, sizeof id);
    return LCC_OK;
}

lcc_status lcc_shard_group_create(int ndev, const int *devices, lcc_shard_group **out) {
    if (!out) return fail(LCC_ERR_VALUE, "out is NULL");
    int count = 0;
    LCC_TRY(lcc_device_count(&count));
    if (ndev <= 0) ndev = count;
    if (ndev > count) return fail(LCC_ERR_VALUE, "%d devices requested, %d present", ndev, count);
    std::vector<int> list(ndev);
    for (int i = 0; i < ndev; ++i) {
        list[i] = devices ? devices[i] : i;
        if (list[i] < 0 || list[i] >= count) return fail(LCC_ERR_VALUE, "device index %d out of range", list[i]);
    }
    auto *g = new lcc_shard_group();
    g->dev.resize(ndev);
    g->world = ndev;
    for (int i = 0; i < ndev; ++i) {
        lcc_status st = init_device(g->dev[i], list[i]);
        if (st != LCC_OK) { lcc_shard_group_destroy(g); return st; }
    }
    if (ndev > 1) {
        NcclApi *n = nullptr;
        lcc_status st = need_nccl(&n);
        if (st != LCC_OK) { lcc_shard_group_destroy(g); return st; }
        std::vector<ncclComm_t> comms(ndev);
        ncclResult_t r = n->CommInitAll(comms.data(), ndev, list.data());
        if (r != ncclSuccess) { lcc_shard_group_destroy(g); return nccl_fail(r, "ncclCommInitAll"); }
        for (int i = 0; i < ndev; ++i) g->dev[i].nccl = comms[i];
    }
    *out = g;
    return LCC_OK;
}

lcc_status lcc_shard_group_create_rank(int device, int world, int rank, const void *id128, lcc_shard_group **out) {
    if (!out) return fail(LCC_ERR_VALUE, "out is NULL");
    if (world <= 0 || rank < 0 || rank >= world) return fail(LCC_ERR_VALUE, "bad rank/world: %d/%d", rank, world);
    if (world > 1 && !id128) return fail(LCC_ERR_VALUE, "a multi-rank group needs the unique id of rank 0 (lcc_shard_unique_id)");
    int count = 0;
    LCC_TRY(lcc_device_count(&count));
    if (device < 0 || device >= count) return fail(LCC_ERR_VALUE, "device index %d out of range", device);
    auto *g = new lcc_shard_group();
    g->dev.resize(1);
    g->world = world;
    g->first_rank = rank;
    lcc_status st = init_device(g->dev[0], device);
    if (st != LCC_OK) { lcc_shard_group_destroy(g); return st; }
    if (world > 1) {
        NcclApi *n = nullptr;
        st = need_nccl(&n);
        if (st != LCC_OK) { lcc_shard_group_destroy(g); return st; }
        ncclUniqueId id;
        memcpy(&id, id128, sizeof id);
        DeviceGuard guard(device);
        ncclResult_t r = n->CommInitRank(&g->dev[0].nccl, world, id, rank);
        if (r != ncclSuccess) { lcc_shard_group_destroy(g); return nccl_fail(r, "ncclCommInitRank"); }
    }
    *out = g;
    return LCC_OK;
}

void lcc_shard_group_destroy(lcc_shard_group *g) {
    if (!g) return;
    NcclApi *n = nccl_api();
    for (auto &d : g->dev) {
        DeviceGuard guard(d.device);
        if (d.compute) cudaStreamSynchronize(d.compute);
        if (d.comm) cudaStreamSynchronize(d.comm);
        if (d.nccl && n->CommDestroy) n->CommDestroy(d.nccl);
        for (int s = 0; s <= kSlots; ++s) {
            if (d.slot[s]) cudaFree(d.slot[s]);
            if (d.produced[s]) cudaEventDestroy(d.produced[s]);
            if (d.reduced[s]) cudaEventDestroy(d.reduced[s]);
        }
        if (d.compute) cudaStreamDestroy(d.compute);
        if (d.comm) cudaStreamDestroy(d.comm);
    }
    delete g;
}

int lcc_shard_group_local_size(const lcc_shard_group *g) { return g ? (int)g->dev.size() : 0; }
int lcc_shard_group_world_size(const lcc_shard_group *g) { return g ? g->world : 0; }
int lcc_shard_group_first_rank(const lcc_shard_group *g) { return g ? g->first_rank : 0; }
int lcc_shard_group_device(const lcc_shard_group *g, int local_index) {
    return g && local_index >= 0 && local_index < (int)g->dev.size() ? g->dev[local_index].device : -1;
}
int lcc_shard_nccl_version(void) {
    NcclApi *n = nccl_api();
    int v = 0;
    if (n->handle && n->GetVersion) n->GetVersion(&v);
    return v;
}

lcc_status lcc_shard_group_stream(const lcc_shard_group *g, int local_index, void **stream) {
    if (!g || !stream || local_index < 0 || local_index >= (int)g->dev.size()) return fail(LCC_ERR_VALUE, "bad shard index");
    *stream = (void *)g->dev[local_index].compute;
    return LCC_OK;
}

lcc_status lcc_shard_group_sync(lcc_shard_group *g) {
    if (!g) return fail(LCC_ERR_VALUE, "shard group is NULL");
    for (auto &d : g->dev) {
        DeviceGuard guard(d.device);
        LCC_CUDA(cudaStreamSynchronize(d.compute));
        LCC_CUDA(cudaStreamSynchronize(d.comm));
    }
    return LCC_OK;
}

lcc_status lcc_shard_batch_alloc(lcc_shard_group *g, const lcc_algebra *alg, int dtype, int64_t n_total, lcc_view *views, int64_t *row_begin) {
    if (!g || !alg || !views) return fail(LCC_ERR_VALUE, "NULL argument");
    for (size_t i = 0; i < g->dev.size(); ++i) views[i] = lcc_view{nullptr, 0, 0, dtype, LCC_SOA};
    for (size_t i = 0; i < g->dev.size(); ++i) {
        int64_t b = 0, e = 0;
        LCC_TRY(lcc_shard_rows(n_total, g->first_rank + (int)i, g->world, &b, &e));
        DeviceGuard guard(g->dev[i].device);
        LCC_TRY(lcc_batch_alloc(alg, dtype, e - b, &views[i], (void *)g->dev[i].compute));
        if (row_begin) row_begin[i] = b;
    }
    return LCC_OK;
}

lcc_status lcc_shard_broadcast_alloc(lcc_shard_group *g, const lcc_algebra *alg, int dtype, const void *host_row, lcc_view *views) {
    if (!g || !alg || !views || !host_row) return fail(LCC_ERR_VALUE, "NULL argument");
    for (size_t i = 0; i < g->dev.size(); ++i) {
        DeviceGuard guard(g->dev[i].device);
        LCC_TRY(lcc_batch_alloc(alg, dtype, 1, &views[i], (void *)g->dev[i].compute));
        LCC_TRY(lcc_batch_upload(alg, host_row, 1, &views[i], (void *)g->dev[i].compute));
    }
    return LCC_OK;
}

lcc_status lcc_shard_batch_free(lcc_shard_group *g, lcc_view *views) {
    if (!g || !views) return LCC_OK;
    for (size_t i = 0; i < g->dev.size(); ++i) {
        DeviceGuard guard(g->dev[i].device);
        LCC_TRY(lcc_batch_free(&views[i], (void *)g->dev[i].compute));
    }
    return LCC_OK;
}

// host (n_total, num_blades) array <-> the shards of this group; every local GPU runs its own chunked pipeline on its own
// host thread, so all PCIe links are busy at once
static lcc_status shard_transfer(lcc_shard_group *g, const lcc_algebra *alg, const void *host_in, void *host_out, int64_t n_total, const lcc_view *views) {
    const size_t nd = g->dev.size();
    std::vector<lcc_status> st(nd, LCC_OK);
    std::vector<std::string> msg(nd);
    auto work = [&](size_t i) {
        cudaSetDevice(g->dev[i].device);
        int64_t b = 0, e = 0;
        st[i] = lcc_shard_rows(n_total, g->first_rank + (int)i, g->world, &b, &e);
        if (st[i] == LCC_OK && e - b != views[i].n) st[i] = fail(LCC_ERR_VALUE, "shard %zu has %lld rows, expected %lld", i, (long long)views[i].n, (long long)(e - b));
        const size_t row_bytes = (size_t)alg->nb * elem_size(views[i].dtype);
        if (st[i] == LCC_OK)
            st[i] = host_in ? lcc_batch_upload(alg, (const char *)host_in + (size_t)b * row_bytes, e - b, &views[i], (void *)g->dev[i].compute)
                            : lcc_batch_download(alg, &views[i], (char *)host_out + (size_t)b * row_bytes, (void *)g->dev[i].compute);
        if (st[i] != LCC_OK) msg[i] = lcc_last_error();
    };
    if (nd == 1) {
        DeviceGuard guard(g->dev[0].device);
        work(0);
    } else {
        std::vector<std::thread> th;
        for (size_t i = 0; i < nd; ++i) th.emplace_back(work, i);
        for (auto &t : th) t.join();
    }
    for (size_t i = 0; i < nd; ++i)
        if (st[i] != LCC_OK) return fail(st[i], "%s", msg[i].c_str());
    return LCC_OK;
}

lcc_status lcc_shard_upload(lcc_shard_group *g, const lcc_algebra *alg, const void *host, int64_t n_total, const lcc_view *views) {
    TraceRange tr("lcc_shard_upload");
    LCC_TRY(check_views(g, alg, views, "dst"));
    if (n_total > 0 && !host) return fail(LCC_ERR_VALUE, "host pointer is NULL");
    return shard_transfer(g, alg, host, nullptr, n_total, views);
}

lcc_status lcc_shard_download(lcc_shard_group *g, const lcc_algebra *alg, const lcc_view *views, void *host, int64_t n_total) {
    TraceRange tr("lcc_shard_download");
    LCC_TRY(check_views(g, alg, views, "src"));
    if (n_total > 0 && !host) return fail(LCC_ERR_VALUE, "host pointer is NULL");
    return shard_transfer(g, alg, nullptr, host, n_total, views);
}

lcc_status lcc_shard_product(lcc_shard_group *g, const lcc_algebra *alg, int op, const lcc_view *a, const lcc_view *b, const lcc_view *out, unsigned flags) {
    TraceRange tr("lcc_shard_product");
    LCC_TRY(check_views(g, alg, a, "a"));
    LCC_TRY(check_views(g, alg, b, "b"));
    LCC_TRY(check_views(g, alg, out, "out"));
    for (size_t i = 0; i < g->dev.size(); ++i) {
        DeviceGuard guard(g->dev[i].device);
        LCC_TRY(lcc_batch_product(alg, op, &a[i], &b[i], &out[i], flags, (void *)g->dev[i].compute));
    }
    return LCC_OK;
}

lcc_status lcc_shard_sandwich(lcc_shard_group *g, const lcc_algebra *alg, const double *rotor, const lcc_view *x, const lcc_view *out, unsigned flags) {
    TraceRange tr("lcc_shard_sandwich");
    LCC_TRY(check_views(g, alg, x, "x"));
    LCC_TRY(check_views(g, alg, out, "out"));
    for (size_t i = 0; i < g->dev.size(); ++i) {
        DeviceGuard guard(g->dev[i].device);
        LCC_TRY(lcc_batch_sandwich(alg, rotor, &x[i], &out[i], flags, (void *)g->dev[i].compute));
    }
    return LCC_OK;
}

lcc_status lcc_shard_unary(lcc_shard_group *g, const lcc_algebra *alg, int op, int arg, const lcc_view *x, const lcc_view *out) {
    TraceRange tr("lcc_shard_unary");
    LCC_TRY(check_views(g, alg, x, "x"));
    LCC_TRY(check_views(g, alg, out, "out"));
    for (size_t i = 0; i < g->dev.size(); ++i) {
        DeviceGuard guard(g->dev[i].device);
        LCC_TRY(lcc_batch_unary(alg, op, arg, &x[i], &out[i], (void *)g->dev[i].compute));
    }
    return LCC_OK;
}

// ---- the one collective.  Synchronous forms: partial -> all-reduce -> (mean) -> host, all on the compute streams.
static lcc_status finish_to_host(lcc_shard_group *g, const lcc_algebra *alg, int dtype, int slot, bool on_comm, void *result_host) {
    ShardDevice &d0 = g->dev[0];
    DeviceGuard guard(d0.device);
    cudaStream_t s = on_comm ? d0.comm : d0.compute;
    if (result_host) LCC_CUDA(cudaMemcpyAsync(result_host, d0.slot[slot], (size_t)alg->nb * elem_size(dtype), cudaMemcpyDeviceToHost, s));
    LCC_CUDA(cudaStreamSynchronize(s));
    return LCC_OK;
}

lcc_status lcc_shard_sum(lcc_shard_group *g, const lcc_algebra *alg, const lcc_view *x, int64_t mean_divisor, void *result_host) {
    TraceRange tr("lcc_shard_sum");
    LCC_TRY(check_views(g, alg, x, "x"));
    const int dtype = x[0].dtype;
    for (size_t i = 0; i < g->dev.size(); ++i) {
        DeviceGuard guard(g->dev[i].device);
        LCC_TRY(shard_partial_sum(alg, &x[i], g->dev[i].slot[kSlots], g->dev[i].compute));
    }
    LCC_TRY(all_reduce_slot(g, alg->nb, dtype, kSlots, false, mean_divisor));
    return finish_to_host(g, alg, dtype, kSlots, false, result_host);
}

lcc_status lcc_shard_product_sum(lcc_shard_group *g, const lcc_algebra *alg, int op, const lcc_view *a, int a_grade, const lcc_view *b,
                                 int64_t mean_divisor, void *result_host, unsigned flags) {
    TraceRange tr("lcc_shard_product_sum");
    LCC_TRY(check_views(g, alg, a, "a"));
    LCC_TRY(check_views(g, alg, b, "b"));
    const int dtype = a[0].dtype;
    for (size_t i = 0; i < g->dev.size(); ++i) {
        DeviceGuard guard(g->dev[i].device);
        LCC_TRY(lcc_batch_product_sum(alg, op, &a[i], a_grade, &b[i], g->dev[i].slot[kSlots], flags, (void *)g->dev[i].compute));
    }
    LCC_TRY(all_reduce_slot(g, alg->nb, dtype, kSlots, false, mean_divisor));
    return finish_to_host(g, alg, dtype, kSlots, false, result_host);
}

// Asynchronous form for pipelines: the fused partial runs on the compute stream, the all-reduce of `slot` on the
// high-priority communication stream behind an event, so the collective of step i overlaps the kernel of step i+1.
// lcc_shard_reduce_fetch(slot) waits for it.  A slot may be reused once fetched (or kSlots steps later: the compute
// stream waits for the slot's previous reduction before overwriting its buffer).
lcc_status lcc_shard_product_sum_async(lcc_shard_group *g, const lcc_algebra *alg, int op, const lcc_view *a, int a_grade, const lcc_view *b,
                                       int64_t mean_divisor, int slot, unsigned flags) {
    TraceRange tr("lcc_shard_product_sum_async");
    if (slot < 0 || slot >= kSlots) return fail(LCC_ERR_VALUE, "slot must be in [0, %d)", kSlots);
    LCC_TRY(check_views(g, alg, a, "a"));
    LCC_TRY(check_views(g, alg, b, "b"));
    const int dtype = a[0].dtype;
    for (size_t i = 0; i < g->dev.size(); ++i) {
        ShardDevice &d = g->dev[i];
        DeviceGuard guard(d.device);
        if (d.reduced_valid[slot]) LCC_CUDA(cudaStreamWaitEvent(d.compute, d.reduced[slot], 0));
        LCC_TRY(lcc_batch_product_sum(alg, op, &a[i], a_grade, &b[i], d.slot[slot], flags, (void *)d.compute));
        LCC_CUDA(cudaEventRecord(d.produced[slot], d.compute));
        LCC_CUDA(cudaStreamWaitEvent(d.comm, d.produced[slot], 0));
    }
    LCC_TRY(all_reduce_slot(g, alg->nb, dtype, slot, true, mean_divisor));
    for (auto &d : g->dev) {
        DeviceGuard guard(d.device);
        LCC_CUDA(cudaEventRecord(d.reduced[slot], d.comm));
        d.reduced_valid[slot] = true;
    }
    return LCC_OK;
}

lcc_status lcc_shard_reduce_fetch(lcc_shard_group *g, const lcc_algebra *alg, int dtype, int slot, void *result_host) {
    if (!g || !alg) return fail(LCC_ERR_VALUE, "NULL argument");
    if (slot < 0 || slot >= kSlots) return fail(LCC_ERR_VALUE, "slot must be in [0, %d)", kSlots);
    if (dtype != LCC_F32 && dtype != LCC_F64) return fail(LCC_ERR_VALUE, "dtype must be LCC_F32 or LCC_F64");
    for (size_t i = 1; i < g->dev.size(); ++i) {  // every local replica of the result is complete when this returns
        DeviceGuard guard(g->dev[i].device);
        LCC_CUDA(cudaStreamSynchronize(g->dev[i].comm));
    }
    return finish_to_host(g, alg, dtype, slot, true, result_host);
}

}  // extern "C"
