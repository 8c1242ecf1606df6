// host_transfer.cu -- everything that moves a batch between HOST memory and HBM:
//
//   lcc_batch_upload / lcc_batch_download   the copies behind MultivectorBatch.from_numpy / to_numpy
//                                           (/root/reference/src/batch.rs:79-94,233-235), chunked and pipelined
//   lcc_host_sandwich / _product / _pga_points_sandwich   whole operations on host arrays, H2D copy of chunk c+1,
//                                           kernel of chunk c and D2H copy of chunk c-1 overlapped on three streams
//   lcc_host_cache_*                        a cache of pinned host blocks (what to_numpy hands out for large results,
//                                           so steady-state exports never pay cudaHostAlloc or page faults)
//
// One generic pipeline serves all of them.  Per device it owns (a) a few device staging buffers in the reference's
// (rows, num_blades) layout, (b) for PAGEABLE host arrays a ring of pinned bounce slots that a small pool of host
// threads fills / drains with memcpy while the DMA engines work on the neighbouring slots, and (c) the events that
// order the reuse of both.  Pinned host arrays (cudaHostAlloc / cudaHostRegister'ed memory, detected with
// cudaPointerGetAttributes) are DMA'd directly.  There is no host arithmetic here: every batch operation is a kernel.
#include <atomic>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <thread>
#include <unistd.h>
#include <vector>

#include "abi_internal.hpp"

using namespace lcc;
using namespace lcc::abi;

// implemented in lcc_abi.cu (the public entry points the per-chunk kernels go through)
extern "C" {
lcc_status lcc_batch_convert(const lcc_algebra *alg, const lcc_view *src, const lcc_view *dst, void *stream);
}

namespace lcc {
namespace abi {

// ------------------------------------------------------------------ tuning knobs (lcc_set_option)
static std::atomic<int64_t> g_chunk_bytes{[] { const char *e = getenv("LCC_HOST_CHUNK_BYTES"); return e && *e ? atoll(e) : (int64_t)(64ll << 20); }()};
static std::atomic<int64_t> g_pinned_output_min{[] { const char *e = getenv("LCC_PINNED_OUTPUT_MIN_BYTES"); return e && *e ? atoll(e) : (int64_t)(1ll << 20); }()};
int64_t option_host_chunk_bytes() { return g_chunk_bytes.load(std::memory_order_relaxed); }
void set_option_host_chunk_bytes(int64_t v) { g_chunk_bytes = v < (1 << 16) ? (1 << 16) : v; }
int64_t option_pinned_output_min_bytes() { return g_pinned_output_min.load(std::memory_order_relaxed); }
void set_option_pinned_output_min_bytes(int64_t v) { g_pinned_output_min = v; }

// ------------------------------------------------------------------ host copy pool
// memcpy between pageable arrays and the pinned bounce slots, split into 1 MiB pieces that the pool's threads and the
// calling thread pull from a shared counter.  One thread moves 6-12 GB/s on the boxes measured; PCIe gen5 wants ~50.
class CopyPool {
public:
    static CopyPool &get() { static CopyPool *p = new CopyPool(); return *p; }  // leaked on purpose: no join at exit
    void copy(void *dst, const void *src, size_t bytes) {
        if (bytes < (4u << 20) || nthreads_ == 0) { memcpy(dst, src, bytes); return; }
        auto job = std::make_shared<Job>();
        job->dst = (char *)dst; job->src = (const char *)src; job->bytes = bytes;
        job->pieces = (bytes + kPiece - 1) / kPiece;
        {
            std::lock_guard<std::mutex> lk(mu_);
            current_ = job;
            ++generation_;
        }
        cv_.notify_all();
        work(*job);
        std::unique_lock<std::mutex> lk(mu_);
        done_cv_.wait(lk, [&] { return job->done.load() == job->pieces; });
        if (current_ == job) current_.reset();
    }
    int threads() const { return nthreads_ + 1; }

private:
    static constexpr size_t kPiece = 1u << 20;
    struct Job {  // each copy has its own counters, so a worker that wakes up late can only touch the job it was woken for
        char *dst = nullptr; const char *src = nullptr; size_t bytes = 0, pieces = 0;
        std::atomic<size_t> next{0}, done{0};
    };
    CopyPool() {
        int n = 0;
        if (const char *e = getenv("LCC_HOST_COPY_THREADS")) n = atoi(e) - 1;
        else { const int hw = (int)std::thread::hardware_concurrency(); n = hw / 2 - 1; if (n > 7) n = 7; }
        if (n < 0) n = 0;
        nthreads_ = n;
        for (int i = 0; i < n; ++i) std::thread([this] { loop(); }).detach();
    }
    void work(Job &job) {
        for (;;) {
            const size_t i = job.next.fetch_add(1);
            if (i >= job.pieces) return;
            const size_t off = i * kPiece, len = off + kPiece <= job.bytes ? kPiece : job.bytes - off;
            memcpy(job.dst + off, job.src + off, len);
            if (job.done.fetch_add(1) + 1 == job.pieces) { std::lock_guard<std::mutex> lk(mu_); done_cv_.notify_all(); }
        }
    }
    void loop() {
        uint64_t seen = 0;
        for (;;) {
            std::shared_ptr<Job> job;
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] { return generation_ != seen; });
                seen = generation_;
                job = current_;
            }
            if (job) work(*job);
        }
    }
    int nthreads_ = 0;
    std::mutex mu_;
    std::condition_variable cv_, done_cv_;
    uint64_t generation_ = 0;
    std::shared_ptr<Job> current_;
};

// ------------------------------------------------------------------ pinned host cache
namespace {
struct HostCache {
    std::mutex mu;
    std::multimap<size_t, void *> idle;     // size -> block
    std::map<void *, size_t> live;          // handed out
    size_t idle_bytes = 0, live_bytes = 0;
    size_t max_idle = [] {
        if (const char *e = getenv("LCC_PINNED_CACHE_MAX_BYTES")) return (size_t)atoll(e);
        const long pages = sysconf(_SC_PHYS_PAGES), psz = sysconf(_SC_PAGE_SIZE);
        return pages > 0 && psz > 0 ? (size_t)pages * (size_t)psz / 4 : (size_t)8 << 30;  // a quarter of the box's RAM
    }();
};
HostCache &host_cache() { static HostCache *c = new HostCache(); return *c; }
constexpr size_t kCacheGranule = 2u << 20;
}  // namespace

// ------------------------------------------------------------------ the per-device transfer engine
struct Stage {  // one device staging buffer and the events that order its reuse
    void *ptr = nullptr; size_t bytes = 0;
    cudaEvent_t filled = nullptr, drained = nullptr;
    bool drained_valid = false;
};
struct BounceSlot {  // one pinned host slot; `free_at` fires when the DMA that used it last has finished
    void *ptr = nullptr;
    cudaEvent_t free_at = nullptr;
    bool in_flight = false;
};
struct BounceRing {
    static constexpr int kSlots = 4;
    BounceSlot slot[kSlots];
    size_t slot_bytes = 0;
    int next = 0;
};

struct TransferEngine {
    std::mutex mu;  // one host pipeline per device at a time
    std::deque<Stage> stages;  // deque: growing it keeps the Stage pointers handed out earlier valid
    BounceRing in_ring, out_ring;

    lcc_status stage(size_t index, size_t bytes, Stage **out) {
        if (stages.size() <= index) stages.resize(index + 1);
        Stage &s = stages[index];
        if (!s.filled) {
            LCC_CUDA(cudaEventCreateWithFlags(&s.filled, cudaEventDisableTiming));
            LCC_CUDA(cudaEventCreateWithFlags(&s.drained, cudaEventDisableTiming));
        }
        if (s.bytes < bytes) {
            if (s.ptr) { LCC_CUDA(cudaDeviceSynchronize()); LCC_CUDA(cudaFree(s.ptr)); s.ptr = nullptr; s.bytes = 0; s.drained_valid = false; }
            cudaError_t e = cudaMalloc(&s.ptr, bytes);
            if (e == cudaErrorMemoryAllocation) { cudaGetLastError(); return fail(LCC_ERR_ALLOC, "out of device memory allocating a %zu-byte staging buffer", bytes); }
            LCC_CUDA(e);
            s.bytes = bytes;
        }
        *out = &s;
        return LCC_OK;
    }
    lcc_status ring(BounceRing &r, size_t bytes) {
        if (r.slot_bytes >= bytes) return LCC_OK;
        for (auto &s : r.slot) {
            if (s.in_flight) { LCC_CUDA(cudaEventSynchronize(s.free_at)); s.in_flight = false; }
            if (s.ptr) { LCC_CUDA(cudaFreeHost(s.ptr)); s.ptr = nullptr; }
            if (!s.free_at) LCC_CUDA(cudaEventCreateWithFlags(&s.free_at, cudaEventDisableTiming | cudaEventBlockingSync));
            LCC_CUDA(cudaHostAlloc(&s.ptr, bytes, cudaHostAllocPortable));
        }
        r.slot_bytes = bytes;
        return LCC_OK;
    }
    // next slot of the ring, after the DMA that used it last has completed
    lcc_status acquire(BounceRing &r, BounceSlot **out) {
        BounceSlot &s = r.slot[r.next];
        r.next = (r.next + 1) % BounceRing::kSlots;
        if (s.in_flight) { LCC_CUDA(cudaEventSynchronize(s.free_at)); s.in_flight = false; }
        *out = &s;
        return LCC_OK;
    }
};

static lcc_status engine_of(DeviceState *st, TransferEngine **out) {
    static std::mutex mu;
    std::lock_guard<std::mutex> lk(mu);
    if (!st->transfer) st->transfer = new TransferEngine();
    *out = st->transfer;
    return LCC_OK;
}

// true when [p, p + bytes) is page-locked memory the DMA engines can read or write directly
static bool is_pinned(const void *p, size_t bytes) {
    if (!p || bytes == 0) return false;
    cudaPointerAttributes a0, a1;
    if (cudaPointerGetAttributes(&a0, p) != cudaSuccess) { cudaGetLastError(); return false; }
    if (cudaPointerGetAttributes(&a1, (const char *)p + bytes - 1) != cudaSuccess) { cudaGetLastError(); return false; }
    return a0.type == cudaMemoryTypeHost && a1.type == cudaMemoryTypeHost;
}

// ------------------------------------------------------------------ the generic chunked pipeline
struct HostInput {
    const void *host = nullptr;
    size_t row_bytes = 0;
    bool broadcast = false;  // one row for the whole batch: uploaded once
};
struct HostOutput {
    void *host = nullptr;
    size_t row_bytes = 0;
};
// kernel(rows, first_row, dev_in[], dev_out, compute_stream): enqueue the chunk's work on the compute stream
using ChunkKernel = std::function<lcc_status(int64_t rows, int64_t first_row, void *const *dev_in, void *dev_out, cudaStream_t s)>;

static lcc_status run_pipeline(int64_t n, int64_t chunk_rows, const std::vector<HostInput> &ins, const HostOutput *outp,
                               const ChunkKernel &kernel, cudaStream_t compute_or_null) {
    DeviceState *st = nullptr;
    LCC_TRY(device_state(&st));
    TransferEngine *eng = nullptr;
    LCC_TRY(engine_of(st, &eng));
    std::lock_guard<std::mutex> lock(eng->mu);
    cudaStream_t compute = compute_or_null ? compute_or_null : st->stream;
    if (chunk_rows > n) chunk_rows = n;
    const size_t nin = ins.size();
    // device staging: 2 buffers per streamed input, 1 per broadcast input, 2 for the output
    std::vector<Stage *> in_stage(2 * nin, nullptr);
    Stage *out_stage[2] = {nullptr, nullptr};
    size_t sidx = 0;
    for (size_t i = 0; i < nin; ++i) {
        const size_t bytes = (ins[i].broadcast ? 1 : (size_t)chunk_rows) * ins[i].row_bytes;
        LCC_TRY(eng->stage(sidx++, bytes, &in_stage[2 * i]));
        if (ins[i].broadcast) in_stage[2 * i + 1] = in_stage[2 * i];
        else LCC_TRY(eng->stage(sidx++, bytes, &in_stage[2 * i + 1]));
    }
    if (outp) for (int b = 0; b < 2; ++b) LCC_TRY(eng->stage(sidx++, (size_t)chunk_rows * outp->row_bytes, &out_stage[b]));
    // pageable arrays go through the pinned bounce rings
    std::vector<char> in_pinned(nin, 0);
    size_t in_slot = 0;
    for (size_t i = 0; i < nin; ++i) {
        in_pinned[i] = is_pinned(ins[i].host, (ins[i].broadcast ? 1 : (size_t)n) * ins[i].row_bytes);
        if (!in_pinned[i]) in_slot = std::max(in_slot, (ins[i].broadcast ? 1 : (size_t)chunk_rows) * ins[i].row_bytes);
    }
    if (in_slot) LCC_TRY(eng->ring(eng->in_ring, in_slot));
    const bool out_pinned = outp && is_pinned(outp->host, (size_t)n * outp->row_bytes);
    if (outp && !out_pinned) LCC_TRY(eng->ring(eng->out_ring, (size_t)chunk_rows * outp->row_bytes));
    CopyPool &pool = CopyPool::get();

    struct PendingOut { BounceSlot *slot; char *dst; size_t bytes; };
    std::vector<PendingOut> pending;
    auto drain_one = [&](const PendingOut &p) -> lcc_status {
        LCC_CUDA(cudaEventSynchronize(p.slot->free_at));
        p.slot->in_flight = false;
        pool.copy(p.dst, p.slot->ptr, p.bytes);
        return LCC_OK;
    };

    // upload one host span into a staging buffer on the copy-in stream
    auto h2d = [&](size_t i, Stage *dst, const char *src, size_t bytes) -> lcc_status {
        if (in_pinned[i]) { LCC_CUDA(cudaMemcpyAsync(dst->ptr, src, bytes, cudaMemcpyHostToDevice, st->copy_in)); return LCC_OK; }
        BounceSlot *slot = nullptr;
        LCC_TRY(eng->acquire(eng->in_ring, &slot));
        pool.copy(slot->ptr, src, bytes);
        LCC_CUDA(cudaMemcpyAsync(dst->ptr, slot->ptr, bytes, cudaMemcpyHostToDevice, st->copy_in));
        LCC_CUDA(cudaEventRecord(slot->free_at, st->copy_in));
        slot->in_flight = true;
        return LCC_OK;
    };

    for (size_t i = 0; i < nin; ++i)
        if (ins[i].broadcast) {
            Stage *sg = in_stage[2 * i];
            if (sg->drained_valid) LCC_CUDA(cudaStreamWaitEvent(st->copy_in, sg->drained, 0));
            LCC_TRY(h2d(i, sg, (const char *)ins[i].host, ins[i].row_bytes));
        }

    int64_t done = 0;
    std::vector<void *> dev_in(nin);
    for (int64_t c = 0; done < n; ++c, done += chunk_rows) {
        const int b = (int)(c & 1);
        const int64_t rows = (n - done) < chunk_rows ? (n - done) : chunk_rows;
        for (size_t i = 0; i < nin; ++i) {
            Stage *sg = in_stage[2 * i + b];
            dev_in[i] = sg->ptr;
            if (ins[i].broadcast) continue;
            if (sg->drained_valid) LCC_CUDA(cudaStreamWaitEvent(st->copy_in, sg->drained, 0));  // the kernel that read it last
            LCC_TRY(h2d(i, sg, (const char *)ins[i].host + (size_t)done * ins[i].row_bytes, (size_t)rows * ins[i].row_bytes));
        }
        if (nin) {
            Stage *first = in_stage[b];  // one event per chunk is enough: all its uploads sit on copy_in before it
            LCC_CUDA(cudaEventRecord(first->filled, st->copy_in));
            LCC_CUDA(cudaStreamWaitEvent(compute, first->filled, 0));
        }
        Stage *og = outp ? out_stage[b] : nullptr;
        if (og && og->drained_valid) LCC_CUDA(cudaStreamWaitEvent(compute, og->drained, 0));  // its previous D2H copy
        LCC_TRY(kernel(rows, done, dev_in.data(), og ? og->ptr : nullptr, compute));
        for (size_t i = 0; i < nin; ++i) {
            Stage *sg = in_stage[2 * i + b];
            LCC_CUDA(cudaEventRecord(sg->drained, compute));
            sg->drained_valid = true;
        }
        if (og) {
            LCC_CUDA(cudaEventRecord(og->filled, compute));
            LCC_CUDA(cudaStreamWaitEvent(st->copy_out, og->filled, 0));
            char *dst = (char *)outp->host + (size_t)done * outp->row_bytes;
            const size_t bytes = (size_t)rows * outp->row_bytes;
            if (out_pinned) {
                LCC_CUDA(cudaMemcpyAsync(dst, og->ptr, bytes, cudaMemcpyDeviceToHost, st->copy_out));
            } else {
                // keep at most kSlots - 1 chunks in flight: drain the oldest before its slot comes round again
                while (pending.size() >= (size_t)BounceRing::kSlots - 1) { LCC_TRY(drain_one(pending.front())); pending.erase(pending.begin()); }
                BounceSlot *slot = nullptr;
                LCC_TRY(eng->acquire(eng->out_ring, &slot));
                LCC_CUDA(cudaMemcpyAsync(slot->ptr, og->ptr, bytes, cudaMemcpyDeviceToHost, st->copy_out));
                LCC_CUDA(cudaEventRecord(slot->free_at, st->copy_out));
                slot->in_flight = true;
                pending.push_back({slot, dst, bytes});
            }
            LCC_CUDA(cudaEventRecord(og->drained, st->copy_out));
            og->drained_valid = true;
        }
    }
    for (const PendingOut &p : pending) LCC_TRY(drain_one(p));
    // every host INPUT byte has been read once the copy-in stream is idle; host OUTPUT is complete once copy-out is.
    // The compute stream is left running when there is no host output (upload): the result is stream-ordered.
    if (nin) LCC_CUDA(cudaStreamSynchronize(st->copy_in));
    if (outp) LCC_CUDA(cudaStreamSynchronize(st->copy_out));
    return LCC_OK;
}

static int64_t chunk_rows_for(size_t row_bytes, int64_t granule) {
    int64_t rows = option_host_chunk_bytes() / (int64_t)row_bytes;
    rows -= rows % granule;
    return rows < granule ? granule : rows;
}

}  // namespace abi
}  // namespace lcc

// ================================================================== C ABI
extern "C" {

lcc_status lcc_host_cache_alloc(void **ptr, size_t bytes) {
    if (!ptr) return fail(LCC_ERR_VALUE, "ptr is NULL");
    HostCache &c = host_cache();
    const size_t want = ((bytes ? bytes : 1) + kCacheGranule - 1) / kCacheGranule * kCacheGranule;
    {
        std::lock_guard<std::mutex> lk(c.mu);
        auto it = c.idle.lower_bound(want);
        if (it != c.idle.end() && it->first <= want + want / 2) {  // close enough in size: reuse
            *ptr = it->second;
            c.live[it->second] = it->first; c.live_bytes += it->first; c.idle_bytes -= it->first;
            c.idle.erase(it);
            return LCC_OK;
        }
    }
    void *p = nullptr;
    cudaError_t e = cudaHostAlloc(&p, want, cudaHostAllocPortable);
    if (e != cudaSuccess) {  // make room and retry once
        cudaGetLastError();
        lcc_host_cache_trim();
        e = cudaHostAlloc(&p, want, cudaHostAllocPortable);
    }
    if (e == cudaErrorMemoryAllocation) { cudaGetLastError(); return fail(LCC_ERR_ALLOC, "out of pinned host memory allocating %zu bytes", want); }
    LCC_CUDA(e);
    std::lock_guard<std::mutex> lk(c.mu);
    c.live[p] = want; c.live_bytes += want;
    *ptr = p;
    return LCC_OK;
}

lcc_status lcc_host_cache_free(void *ptr) {
    if (!ptr) return LCC_OK;
    HostCache &c = host_cache();
    size_t size = 0;
    bool release = false;
    {
        std::lock_guard<std::mutex> lk(c.mu);
        auto it = c.live.find(ptr);
        if (it == c.live.end()) return fail(LCC_ERR_VALUE, "lcc_host_cache_free: pointer was not handed out by lcc_host_cache_alloc");
        size = it->second;
        c.live.erase(it); c.live_bytes -= size;
        if (c.idle_bytes + size > c.max_idle) release = true;
        else { c.idle.emplace(size, ptr); c.idle_bytes += size; }
    }
    if (release) LCC_CUDA(cudaFreeHost(ptr));
    return LCC_OK;
}

lcc_status lcc_host_cache_trim(void) {
    HostCache &c = host_cache();
    std::vector<void *> blocks;
    {
        std::lock_guard<std::mutex> lk(c.mu);
        for (auto &kv : c.idle) blocks.push_back(kv.second);
        c.idle.clear(); c.idle_bytes = 0;
    }
    for (void *p : blocks) LCC_CUDA(cudaFreeHost(p));
    return LCC_OK;
}

lcc_status lcc_host_cache_stats(size_t *live_bytes, size_t *idle_bytes) {
    HostCache &c = host_cache();
    std::lock_guard<std::mutex> lk(c.mu);
    if (live_bytes) *live_bytes = c.live_bytes;
    if (idle_bytes) *idle_bytes = c.idle_bytes;
    return LCC_OK;
}

int lcc_host_is_pinned(const void *ptr, size_t bytes) { return is_pinned(ptr, bytes) ? 1 : 0; }

lcc_status lcc_batch_upload(const lcc_algebra *alg, const void *host, int64_t n, const lcc_view *dst, void *stream) {
    TraceRange tr("lcc_batch_upload");
    LCC_TRY(check_view(alg, dst, "dst"));
    if (n != dst->n) return fail(LCC_ERR_VALUE, "upload: host array has %lld rows, the batch %lld", (long long)n, (long long)dst->n);
    if (n == 0) return LCC_OK;
    if (!host) return fail(LCC_ERR_VALUE, "host pointer is NULL");
    const size_t es = elem_size(dst->dtype), row_bytes = (size_t)alg->nb * es;
    const lcc_view d = *dst;
    const int nb = alg->nb;
    const int dtype = dst->dtype;
    std::vector<HostInput> ins(1);
    ins[0].host = host; ins[0].row_bytes = row_bytes;
    // chunks start at multiples of 64 rows so every blade-major sub-view stays 16-byte aligned (TMA transposes)
    return run_pipeline(n, chunk_rows_for(row_bytes, 64), ins, nullptr,
        [&](int64_t rows, int64_t first, void *const *dev_in, void *, cudaStream_t s) -> lcc_status {
            lcc_view src{dev_in[0], rows, nb, dtype, LCC_AOS};
            lcc_view sub = d;
            sub.n = rows;
            sub.data = (char *)d.data + (size_t)(d.layout == LCC_SOA ? first : first * d.ld) * es;
            return lcc_batch_convert(alg, &src, &sub, (void *)s);
        }, (cudaStream_t)stream);
}

lcc_status lcc_batch_download(const lcc_algebra *alg, const lcc_view *src, void *host, void *stream) {
    TraceRange tr("lcc_batch_download");
    LCC_TRY(check_view(alg, src, "src"));
    if (src->n == 0) return LCC_OK;
    if (!host) return fail(LCC_ERR_VALUE, "host pointer is NULL");
    const size_t es = elem_size(src->dtype), row_bytes = (size_t)alg->nb * es;
    const lcc_view v = *src;
    const int nb = alg->nb;
    HostOutput out;
    out.host = host; out.row_bytes = row_bytes;
    return run_pipeline(v.n, chunk_rows_for(row_bytes, 64), {}, &out,
        [&](int64_t rows, int64_t first, void *const *, void *dev_out, cudaStream_t s) -> lcc_status {
            lcc_view sub = v;
            sub.n = rows;
            sub.data = (char *)v.data + (size_t)(v.layout == LCC_SOA ? first : first * v.ld) * es;
            lcc_view dst{dev_out, rows, nb, v.dtype, LCC_AOS};
            return lcc_batch_convert(alg, &sub, &dst, (void *)s);
        }, (cudaStream_t)stream);
}

lcc_status lcc_host_sandwich(const lcc_algebra *alg, int dtype, const double *rotor, const void *x_host, void *out_host,
                             int64_t n, unsigned flags) {
    TraceRange tr("lcc_host_sandwich");
    if (!alg) return fail(LCC_ERR_VALUE, "algebra handle is NULL");
    if (dtype != LCC_F32 && dtype != LCC_F64) return fail(LCC_ERR_VALUE, "dtype must be LCC_F32 or LCC_F64");
    if (n < 0) return fail(LCC_ERR_VALUE, "negative batch size");
    if (n == 0) return LCC_OK;
    if (!rotor || !x_host || !out_host) return fail(LCC_ERR_VALUE, "NULL buffer");
    const size_t row_bytes = (size_t)alg->nb * elem_size(dtype);
    const int nb = alg->nb;
    std::vector<HostInput> ins(1);
    ins[0].host = x_host; ins[0].row_bytes = row_bytes;
    HostOutput out;
    out.host = out_host; out.row_bytes = row_bytes;
    return run_pipeline(n, chunk_rows_for(row_bytes, 4), ins, &out,
        [&](int64_t rows, int64_t, void *const *dev_in, void *dev_out, cudaStream_t s) -> lcc_status {
            lcc_view vx{dev_in[0], rows, nb, dtype, LCC_AOS}, vo{dev_out, rows, nb, dtype, LCC_AOS};
            return lcc_batch_sandwich(alg, rotor, &vx, &vo, flags, (void *)s);
        }, nullptr);
}

lcc_status lcc_host_pga_points_sandwich(const lcc_algebra *alg, int dtype, const double *motor, const void *xyz_host_in,
                                        void *xyz_host_out, int64_t n, unsigned flags) {
    TraceRange tr("lcc_host_pga_points_sandwich");
    if (!alg) return fail(LCC_ERR_VALUE, "algebra handle is NULL");
    if (!(alg->p == 3 && alg->q == 0 && alg->r == 1)) return fail(LCC_ERR_VALUE, "pga_point requires a PGA3D algebra Cl(3,0,1)");  // src/pga.rs:110-114
    if (dtype != LCC_F32 && dtype != LCC_F64) return fail(LCC_ERR_VALUE, "dtype must be LCC_F32 or LCC_F64");
    if (n < 0) return fail(LCC_ERR_VALUE, "negative batch size");
    if (n == 0) return LCC_OK;
    if (!motor || !xyz_host_in || !xyz_host_out) return fail(LCC_ERR_VALUE, "NULL buffer");
    const size_t row_bytes = 3 * elem_size(dtype);
    // a quarter of the usual chunk: the whole batch is only 12-24 bytes per point, so the un-overlapped first H2D and
    // last D2H copy weigh more
    int64_t chunk = ((option_host_chunk_bytes() / 4) / (int64_t)row_bytes) & ~(int64_t)3;
    if (chunk < 4) chunk = 4;
    std::vector<HostInput> ins(1);
    ins[0].host = xyz_host_in; ins[0].row_bytes = row_bytes;
    HostOutput out;
    out.host = xyz_host_out; out.row_bytes = row_bytes;
    return run_pipeline(n, chunk, ins, &out,
        [&](int64_t rows, int64_t, void *const *dev_in, void *dev_out, cudaStream_t s) -> lcc_status {
            return lcc_pga_points_sandwich(alg, dtype, motor, dev_in[0], dev_out, rows, flags, (void *)s);
        }, nullptr);
}

lcc_status lcc_host_product(const lcc_algebra *alg, int dtype, int op, const void *a_host, int64_t na, const void *b_host,
                            int64_t nb_rows, void *out_host, unsigned flags) {
    TraceRange tr("lcc_host_product");
    if (!alg) return fail(LCC_ERR_VALUE, "algebra handle is NULL");
    if (dtype != LCC_F32 && dtype != LCC_F64) return fail(LCC_ERR_VALUE, "dtype must be LCC_F32 or LCC_F64");
    if (op != LCC_OP_GP && op != LCC_OP_OUTER && op != LCC_OP_LC) return fail(LCC_ERR_VALUE, "unknown product op %d", op);
    int64_t n = 0;
    LCC_TRY(broadcast_rows(na, nb_rows, &n));
    if (n == 0) return LCC_OK;
    if (!a_host || !b_host || !out_host) return fail(LCC_ERR_VALUE, "NULL buffer");
    const size_t row_bytes = (size_t)alg->nb * elem_size(dtype);
    const int nb = alg->nb;
    const bool a_bc = na == 1 && n > 1, b_bc = nb_rows == 1 && n > 1;
    std::vector<HostInput> ins(2);
    ins[0].host = a_host; ins[0].row_bytes = row_bytes; ins[0].broadcast = a_bc;
    ins[1].host = b_host; ins[1].row_bytes = row_bytes; ins[1].broadcast = b_bc;
    HostOutput out;
    out.host = out_host; out.row_bytes = row_bytes;
    return run_pipeline(n, chunk_rows_for(row_bytes, 4), ins, &out,
        [&](int64_t rows, int64_t, void *const *dev_in, void *dev_out, cudaStream_t s) -> lcc_status {
            lcc_view va{dev_in[0], a_bc ? 1 : rows, nb, dtype, LCC_AOS};
            lcc_view vb{dev_in[1], b_bc ? 1 : rows, nb, dtype, LCC_AOS};
            lcc_view vo{dev_out, rows, nb, dtype, LCC_AOS};
            return lcc_batch_product(alg, op, &va, &vb, &vo, flags, (void *)s);
        }, nullptr);
}

}  // extern "C"
