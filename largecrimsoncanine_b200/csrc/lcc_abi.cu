// lcc_abi.cu -- the C ABI of include/lcc_b200.h: argument validation with the reference's error
// behaviour, the algebra registry, the stream-ordered device memory pool, and dispatch to the
// sm_100a kernels (per-signature generated translation units + misc_kernels.cu).
//
// There is deliberately no host implementation of any batched operation in this file: without a
// usable CUDA device every compute entry point returns LCC_ERR_NO_DEVICE.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include <nvtx3/nvToolsExt.h>

#include "abi_internal.hpp"
#include "matrix_rep.hpp"
#include "generated/cayley_tables.h"
#include "kernels/launch_args.h"

// ------------------------------------------------------------------ generated kernel registry
namespace lcc {
#define LCC_SPEC_ENTRY(p, q, r, tu32, tu64)                         \
    cudaError_t launch_product_##tu32(const ProductArgs &);         \
    cudaError_t launch_sandwich_##tu32(const SandwichArgs &);       \
    cudaError_t launch_product_##tu64(const ProductArgs &);         \
    cudaError_t launch_sandwich_##tu64(const SandwichArgs &);
#define LCC_GENERIC_ENTRY(dim, tu32, tu64) LCC_SPEC_ENTRY(dim, 0, 0, tu32, tu64)
#include "generated/dispatch_table.inc"
#undef LCC_SPEC_ENTRY
#undef LCC_GENERIC_ENTRY

struct SpecEntry { int p, q, r; ProductLauncher prod[2]; SandwichLauncher sand[2]; };
struct GenericEntry { int dim; ProductLauncher prod[2]; SandwichLauncher sand[2]; };
static const SpecEntry kSpecEntries[] = {
#define LCC_SPEC_ENTRY(p, q, r, tu32, tu64) \
    {p, q, r, {launch_product_##tu32, launch_product_##tu64}, {launch_sandwich_##tu32, launch_sandwich_##tu64}},
#define LCC_GENERIC_ENTRY(dim, tu32, tu64)
#include "generated/dispatch_table.inc"
#undef LCC_SPEC_ENTRY
#undef LCC_GENERIC_ENTRY
    {-1, -1, -1, {nullptr, nullptr}, {nullptr, nullptr}}};
static const GenericEntry kGenericEntries[] = {
#define LCC_SPEC_ENTRY(p, q, r, tu32, tu64)
#define LCC_GENERIC_ENTRY(dim, tu32, tu64) \
    {dim, {launch_product_##tu32, launch_product_##tu64}, {launch_sandwich_##tu32, launch_sandwich_##tu64}},
#include "generated/dispatch_table.inc"
#undef LCC_SPEC_ENTRY
#undef LCC_GENERIC_ENTRY
    {-1, {nullptr, nullptr}, {nullptr, nullptr}}};

cudaError_t launch_pga_points_sandwich_float(const SandwichArgs &);   // the Cl(3,0,1) translation units
cudaError_t launch_pga_points_sandwich_double(const SandwichArgs &);

static thread_local uint64_t t_launches = 0;
void note_kernel_launch() { ++t_launches; }

static int64_t env_or(const char *name, int64_t fallback) {
    const char *e = getenv(name);
    return e && *e ? atoll(e) : fallback;
}
static std::atomic<int64_t> g_aos_tma_min_rows{env_or("LCC_AOS_TMA_MIN_ROWS", 4096)};
static std::atomic<int64_t> g_soa_tma_min_bytes{env_or("LCC_SOA_TMA_MIN_BYTES", 8ll << 30)};
static std::atomic<int64_t> g_matrix_rep{env_or("LCC_MATRIX_REP", 2)};
// grid of the fused product + batch-sum kernel (kernels/launch_args.h); LCC_FUSED_SUM_BLOCKS overrides it for sweeps
static int fused_sum_blocks() {
    static const int v = [] { const int64_t e = env_or("LCC_FUSED_SUM_BLOCKS", kFusedSumBlocks); return e >= 1 && e <= (1 << 20) ? (int)e : kFusedSumBlocks; }();
    return v;
}
int64_t option_aos_tma_min_rows() { return g_aos_tma_min_rows.load(std::memory_order_relaxed); }
int64_t option_soa_tma_min_bytes() { return g_soa_tma_min_bytes.load(std::memory_order_relaxed); }

// ------------------------------------------------------------------ algebra builder
void build_algebra(lcc_algebra &alg, int p, int q, int r) {
    alg.p = p; alg.q = q; alg.r = r; alg.dim = p + q + r; alg.nb = 1 << alg.dim;
    const int nb = alg.nb;
    alg.signs.resize((size_t)nb * nb);
    alg.blades.resize((size_t)nb * nb);
    for (int a = 0; a < nb; ++a)
        for (int b = 0; b < nb; ++b) {
            alg.signs[(size_t)a * nb + b] = (int8_t)blade_sign((unsigned)a, (unsigned)b, p, q, r);
            alg.blades[(size_t)a * nb + b] = (uint16_t)(a ^ b);
        }
    alg.names.resize(nb);
    for (int i = 0; i < nb; ++i) {  // blades.rs:48-61: "1", then "e" + 1-based vector indices
        if (i == 0) { alg.names[i] = "1"; continue; }
        std::string s = "e";
        for (int b = 0; b < alg.dim; ++b)
            if ((i >> b) & 1) s += std::to_string(b + 1);
        alg.names[i] = s;
    }
    const int words = (nb + 63) / 64;
    alg.grade_masks.assign(alg.dim + 1, std::vector<uint64_t>(words, 0));
    for (int i = 0; i < nb; ++i) alg.grade_masks[popcount32(i)][i / 64] |= 1ull << (i % 64);
    alg.mu.resize(nb);
    for (int c = 0; c < nb; ++c) {
        int m = 1;
        for (int b = 0; b < alg.dim; ++b)
            if ((c >> b) & 1) m *= basis_square(p, q, r, b);
        alg.mu[c] = (double)m;
    }
    alg.norm_weights.resize(nb);
    for (int k = 0; k < nb; ++k)
        alg.norm_weights[k] = (int8_t)(alg.signs[(size_t)k * nb + k] * reverse_sign_of_grade(popcount32(k)));
    alg.pseudoscalar_invertible = alg.signs[(size_t)(nb - 1) * nb + (nb - 1)] != 0;
    for (const SpecEntry *e = kSpecEntries; e->p >= 0; ++e)
        if (e->p == p && e->q == q && e->r == r) {
            alg.specialised = true;
            for (int d = 0; d < 2; ++d) { alg.product[d] = e->prod[d]; alg.sandwich[d] = e->sand[d]; }
        }
    if (!alg.specialised)
        for (const GenericEntry *e = kGenericEntries; e->dim >= 0; ++e)
            if (e->dim == alg.dim)
                for (int d = 0; d < 2; ++d) { alg.product[d] = e->prod[d]; alg.sandwich[d] = e->sand[d]; }
}
}  // namespace lcc

using namespace lcc;

// ------------------------------------------------------------------ errors, tracing
static thread_local std::string t_error;

namespace lcc {
namespace abi {
lcc_status fail(lcc_status code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    t_error = buf;
    if (log_enabled()) log_line("error %d: %s", (int)code, buf);
    return code;
}
lcc_status cuda_fail(cudaError_t e, const char *what) {
    if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver)
        return fail(LCC_ERR_NO_DEVICE, "%s: no usable CUDA device (%s); batched operations have no CPU fallback",
                    what, cudaGetErrorString(e));
    return fail(LCC_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
}
std::string sig_str(const lcc_algebra *a) {
    char b[64];
    snprintf(b, sizeof b, "Cl(%d,%d,%d)", a->p, a->q, a->r);
    return b;
}
bool log_enabled() {
    static const bool on = [] { const char *e = getenv("LCC_LOG"); return e && e[0] && e[0] != '0'; }();
    return on;
}
void log_line(const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    fprintf(stderr, "[lcc] %s\n", buf);
}
static bool nvtx_enabled() {  // LCC_NVTX=0 switches the ranges off (they cost ~20 ns each without a profiler attached)
    static const bool on = [] { const char *e = getenv("LCC_NVTX"); return !(e && e[0] == '0'); }();
    return on;
}
TraceRange::TraceRange(const char *name) {
    if (nvtx_enabled()) nvtxRangePushA(name);
    if (log_enabled()) log_line("%s", name);
}
TraceRange::~TraceRange() {
    if (nvtx_enabled()) nvtxRangePop();
}
}  // namespace abi
}  // namespace lcc
using namespace lcc::abi;

// ------------------------------------------------------------------ device state
namespace lcc {
namespace abi {
static std::mutex g_dev_mutex;
static DeviceState g_devices[64];

lcc_status device_state(DeviceState **out) {
    int dev = 0;
    LCC_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return fail(LCC_ERR_CUDA, "device index %d out of range", dev);
    std::lock_guard<std::mutex> lock(g_dev_mutex);
    DeviceState &st = g_devices[dev];
    if (!st.ready) {
        st.device = dev;
        LCC_CUDA(cudaStreamCreateWithFlags(&st.stream, cudaStreamNonBlocking));
        LCC_CUDA(cudaStreamCreateWithFlags(&st.copy_in, cudaStreamNonBlocking));
        LCC_CUDA(cudaStreamCreateWithFlags(&st.copy_out, cudaStreamNonBlocking));
        cudaMemPool_t pool;
        LCC_CUDA(cudaDeviceGetDefaultMemPool(&pool, dev));
        uint64_t keep = UINT64_MAX;  // keep freed blocks in the pool: steady-state calls never hit cudaMalloc
        LCC_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
        st.ready = true;
    }
    *out = &st;
    return LCC_OK;
}

lcc_status resolve_stream(void *stream, cudaStream_t *out) {
    // the device state is set up on EVERY path, also when the caller brings its own stream: it configures the stream-ordered
    // pool to keep freed blocks (release threshold), without which the scratch buffers of the entry points (batch sums,
    // K8's unit counter, matrix builders) go back to the driver at every synchronisation and are re-mapped on the next call
    DeviceState *st = nullptr;
    LCC_TRY(device_state(&st));
    *out = stream ? (cudaStream_t)stream : st->stream;
    return LCC_OK;
}

lcc_status check_view(const lcc_algebra *alg, const lcc_view *v, const char *name, bool allow_null_data) {
    if (!alg) return fail(LCC_ERR_VALUE, "algebra handle is NULL");
    if (!v) return fail(LCC_ERR_VALUE, "%s: view is NULL", name);
    if (v->dtype != LCC_F32 && v->dtype != LCC_F64) return fail(LCC_ERR_VALUE, "%s: dtype must be LCC_F32 or LCC_F64", name);
    if (v->layout != LCC_SOA && v->layout != LCC_AOS) return fail(LCC_ERR_VALUE, "%s: layout must be LCC_SOA or LCC_AOS", name);
    if (v->n < 0) return fail(LCC_ERR_VALUE, "%s: negative batch size", name);
    if (v->n > 0 && !v->data && !allow_null_data) return fail(LCC_ERR_VALUE, "%s: data pointer is NULL", name);
    const int64_t w = 16 / (int64_t)elem_size(v->dtype);
    if (v->layout == LCC_SOA) {
        if (v->ld < v->n || v->ld % w != 0 || ((uintptr_t)v->data % 16) != 0)
            return fail(LCC_ERR_VALUE, "%s: SoA views need ld >= n, ld a multiple of %lld and 16-byte aligned data "
                        "(use lcc_batch_alloc)", name, (long long)w);
    } else if (v->ld < alg->nb) {
        return fail(LCC_ERR_VALUE, "coeffs must have %d columns for %s but got %lld", alg->nb, sig_str(alg).c_str(), (long long)v->ld);
    }
    return LCC_OK;
}

// batch.rs:361-376
lcc_status broadcast_rows(int64_t na, int64_t nbr, int64_t *n) {
    if (na == nbr || nbr == 1) { *n = na; return LCC_OK; }
    if (na == 1) { *n = nbr; return LCC_OK; }
    return fail(LCC_ERR_VALUE, "batch sizes must match or be 1 for broadcasting: %lld vs %lld", (long long)na, (long long)nbr);
}

int64_t padded_ld(int64_t n) { return n <= 0 ? 64 : ((n + 63) / 64) * 64; }
}  // namespace abi
}  // namespace lcc

namespace {
lcc_status same_kind(const lcc_view *a, const lcc_view *b, const char *what) {
    if (a->dtype != b->dtype) return fail(LCC_ERR_VALUE, "%s: operands must share one dtype", what);
    if (a->layout != b->layout) return fail(LCC_ERR_VALUE, "%s: operands must share one layout", what);
    return LCC_OK;
}

// Device copy of a gather-form sign table for the table-driven kernel.
//   forward  (kind = op):        out[k] = sum_i S[i][i^k] a[i] b[i^k]              G[k][i] = S[i][i^k]
//   VJP wrt a (kind = 16 + op):  ga[i]  = sum_k' S[i][i^k'] g[k'] b[k'^i]           G[i][k'] = S[i][i^k']  (x = g, y = b)
//   VJP wrt b (kind = 32 + op):  gb[j]  = sum_i  S[i][j]    a[i]  g[i^j]            G[j][i]  = S[i][j]     (x = a, y = g)
//   kind = 48: the plain sign table sign[a*nb+b]
// each restricted to the op's term set ((i&j)==0 outer, (i&j)==i left contraction, (i&j)==j right contraction,
// i==j scalar product; batch.rs:555,636, multivector.rs:1647,1697).
lcc_status gather_signs_on_device(const lcc_algebra *calg, int kind, const int8_t **out) {
    lcc_algebra *alg = const_cast<lcc_algebra *>(calg);
    int dev = 0;
    LCC_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(alg->mu_dev);
    auto key = std::make_pair(dev, kind);
    auto it = alg->gather_signs_dev.find(key);
    if (it != alg->gather_signs_dev.end()) { *out = it->second; return LCC_OK; }
    const int nb = alg->nb, op = kind % 16, which = kind / 16;  // which == 3: the plain sign table sign[a*nb+b]
    auto in_op = [&](int i, int j) {
        return op == kOpGP || (op == kOpOuter && (i & j) == 0) || (op == kOpLC && (i & j) == i) || (op == kOpRC && (i & j) == j) ||
               (op == kOpScalar && i == j);
    };
    auto S = [&](int i, int j) -> int8_t { return in_op(i, j) ? alg->signs[(size_t)i * nb + j] : (int8_t)0; };
    std::vector<int8_t> g((size_t)nb * nb);
    for (int o = 0; o < nb; ++o)
        for (int x = 0; x < nb; ++x) {
            int8_t v;
            if (which == 0) v = S(x, x ^ o);        // o = output blade k, x = left blade i, right blade i^k
            else if (which == 1) v = S(o, o ^ x);   // o = left blade i, x = grad blade k', right blade i^k'
            else if (which == 2) v = S(x, o);       // o = right blade j, x = left blade i, grad blade i^j
            else v = alg->signs[(size_t)o * nb + x]; // plain table (tensor-core path builds W from it)
            g[(size_t)o * nb + x] = v;
        }
    int8_t *d = nullptr;
    LCC_CUDA(cudaMalloc(&d, g.size()));
    LCC_CUDA(cudaMemcpy(d, g.data(), g.size(), cudaMemcpyHostToDevice));
    alg->gather_signs_dev[key] = d;
    *out = d;
    return LCC_OK;
}

lcc_status make_soa_temp(const lcc_algebra *alg, int dtype, int64_t n, PoolBuffer &buf, lcc_view *v, cudaStream_t s) {
    v->n = n; v->ld = padded_ld(n); v->dtype = dtype; v->layout = LCC_SOA;
    LCC_TRY(buf.alloc((size_t)alg->nb * v->ld * elem_size(dtype), s));
    v->data = buf.ptr;
    return LCC_OK;
}
}  // namespace

// ================================================================== library / device
extern "C" {

const char *lcc_version(void) { return "0.3.0+b200.1"; }
const char *lcc_last_error(void) { return t_error.c_str(); }

lcc_status lcc_device_count(int *count) {
    int c = 0;
    cudaError_t e = cudaGetDeviceCount(&c);
    if (e != cudaSuccess || c == 0) {
        if (count) *count = 0;
        return cuda_fail(e == cudaSuccess ? cudaErrorNoDevice : e, "cudaGetDeviceCount");
    }
    if (count) *count = c;
    return LCC_OK;
}
lcc_status lcc_set_device(int device) { LCC_CUDA(cudaSetDevice(device)); return LCC_OK; }
lcc_status lcc_get_device(int *device) { LCC_CUDA(cudaGetDevice(device)); return LCC_OK; }
lcc_status lcc_stream(void **stream) {
    DeviceState *st = nullptr;
    LCC_TRY(device_state(&st));
    *stream = (void *)st->stream;
    return LCC_OK;
}
lcc_status lcc_stream_sync(void *stream) {
    cudaStream_t s;
    LCC_TRY(resolve_stream(stream, &s));
    LCC_CUDA(cudaStreamSynchronize(s));
    return LCC_OK;
}

// ================================================================== algebra
lcc_status lcc_algebra_create(int p, int q, int r, lcc_algebra **out) {
    if (!out) return fail(LCC_ERR_VALUE, "out is NULL");
    if (p < 0 || q < 0 || r < 0) return fail(LCC_ERR_VALUE, "signature entries must be non-negative");
    if (p + q + r > 8) return fail(LCC_ERR_UNSUPPORTED, "Cl(%d,%d,%d): at most 8 basis vectors (256 blades) are supported", p, q, r);
    lcc_algebra *alg = new lcc_algebra();
    build_algebra(*alg, p, q, r);
    *out = alg;
    return LCC_OK;
}
void lcc_algebra_destroy(lcc_algebra *alg) {
    if (!alg) return;
    for (auto &kv : alg->gather_signs_dev) cudaFree(kv.second);
    delete alg;
}
int lcc_algebra_p(const lcc_algebra *a) { return a->p; }
int lcc_algebra_q(const lcc_algebra *a) { return a->q; }
int lcc_algebra_r(const lcc_algebra *a) { return a->r; }
int lcc_algebra_dimension(const lcc_algebra *a) { return a->dim; }
int lcc_algebra_num_blades(const lcc_algebra *a) { return a->nb; }
const int8_t *lcc_algebra_cayley_signs(const lcc_algebra *a) { return a->signs.data(); }
const uint16_t *lcc_algebra_cayley_blades(const lcc_algebra *a) { return a->blades.data(); }
int lcc_algebra_grade_mask(const lcc_algebra *a, int grade, uint64_t *words, int capacity) {
    const int nwords = (a->nb + 63) / 64;
    if (words)
        for (int w = 0; w < nwords && w < capacity; ++w)
            words[w] = (grade >= 0 && grade <= a->dim) ? a->grade_masks[grade][w] : 0;  // registry.rs:224-232
    return nwords;
}
int lcc_algebra_blade_name(const lcc_algebra *a, int index, char *buf, int capacity) {
    if (index < 0 || index >= a->nb || !buf || capacity <= 0) return -1;
    const std::string &s = a->names[index];
    snprintf(buf, capacity, "%s", s.c_str());
    return (int)s.size();
}
int lcc_algebra_is_specialised(const lcc_algebra *a) { return a->specialised ? 1 : 0; }
int lcc_reverse_sign(int blade) { return reverse_sign_of_grade(popcount32((unsigned)blade)); }
int lcc_grade_involution_sign(int blade) { return involution_sign_of_grade(popcount32((unsigned)blade)); }
int lcc_clifford_conjugate_sign(int blade) { return conjugate_sign_of_grade(popcount32((unsigned)blade)); }

// ================================================================== memory
lcc_status lcc_device_alloc(void **ptr, size_t bytes, void *stream) {
    cudaStream_t s;
    LCC_TRY(resolve_stream(stream, &s));
    cudaError_t e = cudaMallocAsync(ptr, bytes ? bytes : 16, s);
    if (e == cudaErrorMemoryAllocation) { cudaGetLastError(); return fail(LCC_ERR_ALLOC, "out of device memory allocating %zu bytes", bytes); }
    LCC_CUDA(e);
    return LCC_OK;
}
lcc_status lcc_device_free(void *ptr, void *stream) {
    if (!ptr) return LCC_OK;
    cudaStream_t s;
    LCC_TRY(resolve_stream(stream, &s));
    LCC_CUDA(cudaFreeAsync(ptr, s));
    return LCC_OK;
}
lcc_status lcc_host_alloc_pinned(void **ptr, size_t bytes) { LCC_CUDA(cudaHostAlloc(ptr, bytes ? bytes : 16, cudaHostAllocDefault)); return LCC_OK; }
lcc_status lcc_host_free_pinned(void *ptr) { if (ptr) LCC_CUDA(cudaFreeHost(ptr)); return LCC_OK; }
static lcc_status copy_async(void *dst, const void *src, size_t bytes, cudaMemcpyKind kind, void *stream) {
    if (bytes == 0) return LCC_OK;
    cudaStream_t s;
    LCC_TRY(resolve_stream(stream, &s));
    LCC_CUDA(cudaMemcpyAsync(dst, src, bytes, kind, s));
    return LCC_OK;
}
lcc_status lcc_memcpy_h2d(void *dst, const void *src, size_t bytes, void *stream) { return copy_async(dst, src, bytes, cudaMemcpyHostToDevice, stream); }
lcc_status lcc_memcpy_d2h(void *dst, const void *src, size_t bytes, void *stream) { return copy_async(dst, src, bytes, cudaMemcpyDeviceToHost, stream); }
lcc_status lcc_memcpy_d2d(void *dst, const void *src, size_t bytes, void *stream) { return copy_async(dst, src, bytes, cudaMemcpyDeviceToDevice, stream); }
lcc_status lcc_memset_zero(void *dst, size_t bytes, void *stream) {
    if (bytes == 0) return LCC_OK;
    cudaStream_t s;
    LCC_TRY(resolve_stream(stream, &s));
    LCC_CUDA(cudaMemsetAsync(dst, 0, bytes, s));
    return LCC_OK;
}
lcc_status lcc_batch_alloc(const lcc_algebra *alg, int dtype, int64_t n, lcc_view *view, void *stream) {
    if (!alg || !view) return fail(LCC_ERR_VALUE, "algebra / view is NULL");
    if (dtype != LCC_F32 && dtype != LCC_F64) return fail(LCC_ERR_VALUE, "dtype must be LCC_F32 or LCC_F64");
    if (n < 0) return fail(LCC_ERR_VALUE, "negative batch size");
    view->n = n; view->ld = padded_ld(n); view->dtype = dtype; view->layout = LCC_SOA; view->data = nullptr;
    return lcc_device_alloc(&view->data, (size_t)alg->nb * view->ld * elem_size(dtype), stream);
}
lcc_status lcc_batch_free(lcc_view *view, void *stream) {
    if (!view || !view->data) return LCC_OK;
    lcc_status s = lcc_device_free(view->data, stream);
    view->data = nullptr;
    return s;
}

// ================================================================== layout conversion / movers
lcc_status lcc_batch_convert(const lcc_algebra *alg, const lcc_view *src, const lcc_view *dst, void *stream) {
    TraceRange tr("lcc_batch_convert");
    LCC_TRY(check_view(alg, src, "src"));
    LCC_TRY(check_view(alg, dst, "dst"));
    if (src->dtype != dst->dtype) return fail(LCC_ERR_VALUE, "convert: dtype mismatch");
    if (src->n != dst->n) return fail(LCC_ERR_VALUE, "convert: batch sizes differ: %lld vs %lld", (long long)src->n, (long long)dst->n);
    cudaStream_t s;
    LCC_TRY(resolve_stream(stream, &s));
    LCC_CUDA(launch_convert(alg->nb, arg_of(src), arg_of(dst), s));
    return LCC_OK;
}

lcc_status lcc_batch_scatter_columns(const lcc_algebra *alg, const void *cols, int ncols, const int32_t *blade_of_col,
                                     const lcc_view *dst, void *stream) {
    TraceRange tr("lcc_batch_scatter_columns");
    LCC_TRY(check_view(alg, dst, "dst"));
    if (ncols < 0 || ncols > alg->nb) return fail(LCC_ERR_VALUE, "scatter: bad column count %d", ncols);
    for (int c = 0; c < ncols; ++c)
        if (blade_of_col[c] < 0 || blade_of_col[c] >= alg->nb) return fail(LCC_ERR_INDEX, "scatter: blade index %d out of range", blade_of_col[c]);
    cudaStream_t s;
    LCC_TRY(resolve_stream(stream, &s));
    LCC_CUDA(launch_scatter_columns(alg->nb, cols, ncols, blade_of_col, arg_of(dst), s));
    return LCC_OK;
}

lcc_status lcc_batch_gather_columns(const lcc_algebra *alg, const lcc_view *src, int ncols, const int32_t *blade_of_col,
                                    void *cols, void *stream) {
    TraceRange tr("lcc_batch_gather_columns");
    LCC_TRY(check_view(alg, src, "src"));
    if (ncols < 0 || ncols > alg->nb) return fail(LCC_ERR_VALUE, "gather: bad column count %d", ncols);
    for (int c = 0; c < ncols; ++c)
        if (blade_of_col[c] < 0 || blade_of_col[c] >= alg->nb) return fail(LCC_ERR_INDEX, "gather: blade index %d out of range", blade_of_col[c]);
    cudaStream_t s;
    LCC_TRY(resolve_stream(stream, &s));
    LCC_CUDA(launch_gather_columns(alg->nb, arg_of(src), ncols, blade_of_col, cols, s));
    return LCC_OK;
}

lcc_status lcc_batch_copy_rows(const lcc_algebra *alg, const lcc_view *src, int64_t start, int64_t end,
                               const lcc_view *dst, int64_t dst_start, void *stream) {
    TraceRange tr("lcc_batch_copy_rows");
    LCC_TRY(check_view(alg, src, "src"));
    LCC_TRY(check_view(alg, dst, "dst"));
    if (src->dtype != dst->dtype) return fail(LCC_ERR_VALUE, "copy_rows: dtype mismatch");
    if (start < 0 || end < 0 || start > src->n || end > src->n || start > end)  // batch.rs:1085-1090 (usize there: negatives cannot occur)
        return fail(LCC_ERR_INDEX, "invalid slice [%lld:%lld] for batch of size %lld", (long long)start, (long long)end, (long long)src->n);
    if (dst_start < 0 || dst_start + (end - start) > dst->n)
        return fail(LCC_ERR_INDEX, "index %lld out of bounds for batch of size %lld", (long long)dst_start, (long long)dst->n);
    if (end == start) return LCC_OK;
    // sub-views of SoA batches start mid-row: only the strided-copy / transpose paths touch them, no alignment needed
    cudaStream_t s;
    LCC_TRY(resolve_stream(stream, &s));
    LCC_CUDA(launch_copy_rows(alg->nb, arg_of(src), start, end, arg_of(dst), dst_start, s));
    return LCC_OK;
}

// ================================================================== the hot path
static bool tensor_path_off() {
    static const bool off = [] { const char *e = getenv("LCC_DISABLE_TENSOR_PATH"); return e && e[0] == '1'; }();
    return off;
}

// Out = X * Wt^T on the tensor cores for a 64 / 128 / 256-blade batch in either layout and dtype (K6 / K6d).  The kernels
// read and write blade-major tiles, so a reference-layout (AoS) batch is transposed in and out through two pool
// temporaries (K7, 6.9 TB/s).  *handled stays false when the kernels cannot take the views (caller falls back).
static lcc_status tensor_matrix_product(const lcc_algebra *alg, const double *wt64, const lcc_view *x, const lcc_view *out, cudaStream_t s, bool *handled) {
    *handled = false;
    const int nb = alg->nb;
    PoolBuffer tin, tout, wsplit;
    lcc_view sx = *x, so = *out;
    if (x->layout == LCC_AOS) {
        LCC_TRY(make_soa_temp(alg, x->dtype, x->n, tin, &sx, s));
        LCC_TRY(make_soa_temp(alg, x->dtype, x->n, tout, &so, s));
        LCC_CUDA(launch_convert(nb, arg_of(x), arg_of(&sx), s));
    }
    cudaError_t e;
    if (x->dtype == LCC_F32) {
        LCC_TRY(wsplit.alloc((size_t)2 * nb * nb * sizeof(float), s));
        e = launch_matrix_product_tf32x3(nb, wt64, arg_of(&sx), arg_of(&so), (float *)wsplit.ptr, s);
    } else {
        e = launch_matrix_product_f64_dmma(nb, wt64, arg_of(&sx), arg_of(&so), s);
    }
    if (e == cudaErrorNotSupported) return LCC_OK;
    LCC_CUDA(e);
    if (x->layout == LCC_AOS) LCC_CUDA(launch_convert(nb, arg_of(&so), arg_of(out), s));
    *handled = true;
    return LCC_OK;
}

// The algebra's 16 x 16 matrix representation (matrix_rep.hpp), searched once.
static bool matrix_rep_of(const lcc_algebra *calg) {
    lcc_algebra *alg = const_cast<lcc_algebra *>(calg);
    std::lock_guard<std::mutex> lock(alg->mu_dev);
    if (alg->rep_state == 0) {
        // Equivalent representations differ in their string -> blade map; K8 wants one whose map gives its MMA fragment
        // accesses distinct shared-memory banks (make_rep_layout).  Variant 0 is the search result itself; the others
        // relabel it by Clifford automorphisms of the Pauli group (a few hundred cheap candidates, tried once per algebra).
        auto install = [&](const MatrixRep &rep) {
            alg->rep_string_of.assign(rep.string_of, rep.string_of + 256);
            alg->rep_sign_of.assign(rep.sign_of, rep.sign_of + 256);
            alg->rep_blade_at.assign(256, 0);
            alg->rep_sign_at.assign(256, 0);
            for (int sidx = 0; sidx < 256; ++sidx)
                if (rep.blade_at[sidx] >= 0) { alg->rep_blade_at[sidx] = (uint8_t)rep.blade_at[sidx]; alg->rep_sign_at[sidx] = rep.sign_of[rep.blade_at[sidx]]; }
        };
        alg->rep_state = 2;
        // K8 has three shared-memory layouts -- blade-major units (make_rep_layout) and reference-layout units in f32 and
        // f64 (make_rep_layout_aos) -- each with its own bank conditions on the string -> blade map.  Take the first
        // variant that satisfies all three; failing that, the first that satisfies the blade-major one.
        std::vector<unsigned> gens;
        int best = -1, best_score = -1;
        for (int variant = 0; variant < 1 + 384 * 13 && best_score < 3; ++variant) {
            const MatrixRep rep = find_matrix_rep(*alg, variant, &gens, variant == 0);
            if (!rep.ok) { if (variant == 0) break; else continue; }
            if (variant == 0) { install(rep); alg->rep_state = 1; }
            uint8_t blade_at[256]; int8_t sign_at[256];
            for (int sidx = 0; sidx < 256; ++sidx) {
                blade_at[sidx] = rep.blade_at[sidx] >= 0 ? (uint8_t)rep.blade_at[sidx] : 0;
                sign_at[sidx] = rep.blade_at[sidx] >= 0 ? rep.sign_of[rep.blade_at[sidx]] : 0;
            }
            const RepLayout L = make_rep_layout(blade_at, sign_at);
            if (!L.ok) continue;
            const int score = 1 + (make_rep_layout_aos(L, 4).ok ? 1 : 0) + (make_rep_layout_aos(L, 8).ok ? 1 : 0);
            if (score > best_score) { best_score = score; best = variant; }
        }
        if (best > 0) {
            const MatrixRep rep = find_matrix_rep(*alg, best, &gens, true);  // verified against the Cayley table before use
            if (rep.ok) install(rep);
        }
    }
    return alg->rep_state == 1;
}

// rho(M) (row-major 16 x 16) from num_blades coefficients; scale folds the 1/16 of the inverse transform in
static void rep_matrix_of(const lcc_algebra *alg, const double *coeffs, double scale, bool transpose, double *out) {
    double m[256];
    for (int i = 0; i < 256; ++i) m[i] = 0.0;
    for (int A = 0; A < alg->nb; ++A) {
        if (coeffs[A] == 0.0) continue;
        const unsigned a = alg->rep_string_of[A] >> 4, b = alg->rep_string_of[A] & 15;
        const double v = coeffs[A] * alg->rep_sign_of[A];
        for (unsigned j = 0; j < 16; ++j) m[(j ^ a) * 16 + j] += (__builtin_popcount(b & j) & 1) ? -v : v;
    }
    for (int i = 0; i < 16; ++i)
        for (int j = 0; j < 16; ++j) out[i * 16 + j] = scale * (transpose ? m[j * 16 + i] : m[i * 16 + j]);
}

// option "matrix_rep": 0 never, 1 f64 batches only, 2 (default) f32 batches too.  Measured at 2^24 rows of Cl(8,0,0) (DESIGN.md
// section 5): f64 K8 11.0-12.4 ms against 72 ms for the DMMA GEMM; f32 K8 5.8-6.4 ms against 9.0-10.2 ms for the tcgen05 GEMM.
static bool matrix_rep_usable(const lcc_algebra *alg, const lcc_view *x, bool strict) {
    const int64_t mode = g_matrix_rep.load(std::memory_order_relaxed);
    if (strict || mode == 0 || (x->dtype == LCC_F32 && mode < 2)) return false;
    return (alg->nb == 128 || alg->nb == 256) && x->n >= 64 && matrix_rep_of(alg);
}

// K8 on x -> out for the rows it can take: all of them for reference-layout (AoS) views, which are transposed in and out
// through pool temporaries; for blade-major views the rows up to the last multiple of 16 bytes (TMA stores clip there) --
// *rows_done tells the caller where its exact kernels have to pick up the 1-3 that remain.
static lcc_status matrix_rep_apply(const lcc_algebra *alg, int mode, const double *p1, const double *p2, const lcc_view *x,
                                   const lcc_view *out, cudaStream_t s, int64_t *rows_done, const lcc_view *fixed_dev = nullptr) {
    *rows_done = 0;
    const int nb = alg->nb;
    // fixed_dev: the fixed operand as a one-row DEVICE view (modes 0 / 1) -- K8 builds its factor on the device, no read-back
    const void *m_dev = fixed_dev ? fixed_dev->data : nullptr;
    const int64_t m_stride = fixed_dev && fixed_dev->layout == LCC_SOA ? fixed_dev->ld : 1;
    const int64_t w = 16 / (int64_t)elem_size(x->dtype);
    PoolBuffer tin, tout;
    lcc_view sx = *x, so = *out;
    // reference-layout views run in place on units of whole rows (one contiguous 16-32 KB span each: 6.3 TB/s of TMA
    // traffic against 4.6 for the 128-byte pieces of a blade-major unit, profiles/r02_tma_passthrough_probe.jsonl)
    static const bool aos_direct = [] { const char *e = getenv("LCC_K8_AOS"); return !(e && e[0] == '0'); }();
    if (aos_direct && x->layout == LCC_AOS && out->layout == LCC_AOS) {
        const cudaError_t e = launch_matrix_rep_product(nb, mode, p1, p2, alg->rep_blade_at.data(), alg->rep_sign_at.data(), arg_of(x), arg_of(out), s, m_dev, m_stride);
        if (e != cudaErrorNotSupported) {
            LCC_CUDA(e);
            *rows_done = x->n;
            return LCC_OK;
        }
    }
    if (x->layout == LCC_AOS) {
        LCC_TRY(make_soa_temp(alg, x->dtype, x->n, tin, &sx, s));
        LCC_TRY(make_soa_temp(alg, x->dtype, x->n, tout, &so, s));
        LCC_CUDA(launch_convert(nb, arg_of(x), arg_of(&sx), s));
        sx.n = so.n = (x->n + w - 1) / w * w;  // the temporaries are padded: whole 16-byte groups, the extra rows are never copied back
    } else {
        sx.n = so.n = x->n - x->n % w;
    }
    if (sx.n > 0) {
        const cudaError_t e = launch_matrix_rep_product(nb, mode, p1, p2, alg->rep_blade_at.data(), alg->rep_sign_at.data(), arg_of(&sx), arg_of(&so), s, m_dev, m_stride);
        if (e == cudaErrorNotSupported) return LCC_OK;
        LCC_CUDA(e);
    }
    if (x->layout == LCC_AOS) {
        so.n = x->n;
        LCC_CUDA(launch_convert(nb, arg_of(&so), arg_of(out), s));
        *rows_done = x->n;
    } else {
        *rows_done = sx.n;
    }
    return LCC_OK;
}

static lcc_view tail_view(const lcc_view *v, int64_t first) {  // rows [first, n) of a view
    lcc_view t = *v;
    t.n = v->n - first;
    t.data = (char *)v->data + (size_t)(v->layout == LCC_SOA ? first : first * v->ld) * elem_size(v->dtype);
    return t;
}

static lcc_status run_product(const lcc_algebra *alg, int op, const lcc_view *a, const lcc_view *b, const lcc_view *out,
                              const lcc_view *out2, unsigned flags, void *stream) {
    LCC_TRY(check_view(alg, a, "a"));
    LCC_TRY(check_view(alg, b, "b"));
    LCC_TRY(check_view(alg, out, "out"));
    LCC_TRY(same_kind(a, b, "product"));
    LCC_TRY(same_kind(a, out, "product"));
    if (out2) { LCC_TRY(check_view(alg, out2, "outer_out")); LCC_TRY(same_kind(a, out2, "product")); }
    int64_t n = 0;
    LCC_TRY(broadcast_rows(a->n, b->n, &n));
    if (out->n != n || (out2 && out2->n != n)) return fail(LCC_ERR_VALUE, "product: output has %lld rows, expected %lld", (long long)out->n, (long long)n);
    if (n == 0) return LCC_OK;
    cudaStream_t s;
    LCC_TRY(resolve_stream(stream, &s));
    const bool strict = (flags & LCC_FLAG_STRICT_FP) != 0;
    const bool a_bc = a->n == 1 && n > 1, b_bc = b->n == 1 && n > 1;
    if (alg->nb <= 32) {
        ProductLauncher launch = alg->product[a->dtype];
        if (!launch) return fail(LCC_ERR_UNSUPPORTED, "no product kernel built for %s", sig_str(alg).c_str());
        ProductArgs g;
        g.op = op; g.dtype = a->dtype; g.layout = a->layout; g.strict = strict; g.n = n;
        g.a = a->data; g.lda = a->ld; g.a_bcast = a_bc;
        g.b = b->data; g.ldb = b->ld; g.b_bcast = b_bc;
        g.out = out->data; g.ldo = out->ld;
        if (out2) { g.out2 = out2->data; g.ldo2 = out2->ld; }
        g.mu = alg->specialised ? nullptr : alg->mu.data();
        g.stream = s;
        LCC_CUDA(launch(g));
        return LCC_OK;
    }
    // 64..256 blades with ONE fixed operand (the broadcast branch, batch.rs:365-376), FMA mode: the map x -> fixed (op) x is
    // linear, so it is a GEMM against its nb x nb matrix (term mask of outer / contractions folded in) on the tensor
    // cores -- K6 (tcgen05, 3xTF32) for f32, K6d (FP64 tensor pipe) for f64; AoS views are transposed through pool
    // temporaries.  Strict mode, and views the kernels cannot take, keep the exact-order table kernel below.
    const bool fixed_operand = (a_bc != b_bc) && (alg->nb == 64 || alg->nb == 128 || alg->nb == 256);
    // geometric product against ONE device-resident operand in an algebra with a real 16 x 16 matrix representation: K8,
    // its 16 x 16 factor built from the device row by a one-block kernel (what lcc.torch passes: no host copy of the operand)
    if (op == kOpGP && fixed_operand && !tensor_path_off() && a->layout == out->layout && b->layout == out->layout) {
        const lcc_view *fixed = a_bc ? a : b, *batch = a_bc ? b : a;
        if (matrix_rep_usable(alg, batch, strict)) {
            int64_t rows_done = 0;
            LCC_TRY(matrix_rep_apply(alg, a_bc ? 0 : 1, nullptr, nullptr, batch, out, s, &rows_done, fixed));
            if (rows_done == n) return LCC_OK;
            if (rows_done > 0) {  // blade-major views: the last 1-3 rows (TMA stores clip at 16 bytes) take the exact kernels
                const lcc_view bt = tail_view(batch, rows_done), ot = tail_view(out, rows_done);
                return run_product(alg, op, a_bc ? a : &bt, a_bc ? &bt : b, &ot, nullptr, flags, stream);
            }
        }
    }
    auto tensor_one = [&](int single_op, const lcc_view *dst, bool *handled) -> lcc_status {
        *handled = false;
        if (tensor_path_off() || strict || !fixed_operand) return LCC_OK;
        if (single_op != kOpGP && single_op != kOpOuter && single_op != kOpLC && single_op != kOpRC && single_op != kOpScalar) return LCC_OK;
        const int8_t *signs_dev = nullptr;
        LCC_TRY(gather_signs_on_device(alg, 48, &signs_dev));
        PoolBuffer w;
        LCC_TRY(w.alloc((size_t)alg->nb * alg->nb * sizeof(double), s));
        const lcc_view *fixed = a_bc ? a : b, *batch = a_bc ? b : a;
        LCC_CUDA(launch_build_product_matrix(alg->nb, single_op, a_bc ? 0 : 1, signs_dev, fixed->data, fixed->layout == LCC_SOA ? fixed->ld : 1,
                                             fixed->dtype, (double *)w.ptr, s));
        return tensor_matrix_product(alg, (const double *)w.ptr, batch, dst, s, handled);
    };
    {
        bool handled = false;
        if (op == kOpGPOuter) {
            LCC_TRY(tensor_one(kOpGP, out, &handled));
            if (handled) {
                LCC_TRY(tensor_one(kOpOuter, out2, &handled));
                if (handled) return LCC_OK;
            }
        } else {
            LCC_TRY(tensor_one(op, out, &handled));
            if (handled) return LCC_OK;
        }
    }
    // 64..256 blades: table-driven shared-memory kernel (any layout)
    auto one = [&](int single_op, const lcc_view *dst) -> lcc_status {
        const int8_t *gs = nullptr;
        LCC_TRY(gather_signs_on_device(alg, single_op, &gs));
        LCC_CUDA(launch_table_product(alg->nb, strict, arg_of(a), a_bc, arg_of(b), b_bc, arg_of(dst), gs, s));
        return LCC_OK;
    };
    if (op == kOpGPOuter) { LCC_TRY(one(kOpGP, out)); return one(kOpOuter, out2); }
    if (op == kOpCommutator || op == kOpAnticommutator) {
        // composed exactly as the reference does (multivector.rs:1732-1768): a*b, b*a, -/+, scale(0.5)
        PoolBuffer tmp;
        lcc_view ba = *out;
        LCC_TRY(tmp.alloc((size_t)(out->layout == LCC_SOA ? alg->nb * out->ld : out->n * out->ld) * elem_size(out->dtype), s));
        ba.data = tmp.ptr;
        const int8_t *gs = nullptr;
        LCC_TRY(gather_signs_on_device(alg, kOpGP, &gs));
        LCC_CUDA(launch_table_product(alg->nb, strict, arg_of(a), a_bc, arg_of(b), b_bc, arg_of(out), gs, s));
        LCC_CUDA(launch_table_product(alg->nb, strict, arg_of(b), b_bc, arg_of(a), a_bc, arg_of(&ba), gs, s));
        LCC_CUDA(launch_addsub(alg->nb, op == kOpCommutator ? LCC_EW_SUB : LCC_EW_ADD, arg_of(out), false, arg_of(&ba), false, arg_of(out), s));
        return lcc_batch_scale(alg, out, 0.5, out, (void *)s);
    }
    return one(op, out);
}

// A v B = undual(dual(A) ^ dual(B)), multivector.rs:1781-1794; LCC_ERR_VALUE when the pseudoscalar is null
// (every PGA), which is what Multivector::dual raises.
static lcc_status run_regressive(const lcc_algebra *alg, const lcc_view *a, const lcc_view *b, const lcc_view *out,
                                 unsigned flags, void *stream) {
    LCC_TRY(check_view(alg, a, "a"));
    LCC_TRY(check_view(alg, b, "b"));
    LCC_TRY(check_view(alg, out, "out"));
    cudaStream_t s;
    LCC_TRY(resolve_stream(stream, &s));
    auto bytes_of = [&](const lcc_view *v) { return (size_t)(v->layout == LCC_SOA ? alg->nb * v->ld : v->n * v->ld) * elem_size(v->dtype); };
    PoolBuffer ta, tb, tw;
    lcc_view da = *a, db = *b, w = *out;
    LCC_TRY(ta.alloc(bytes_of(a), s));
    LCC_TRY(tb.alloc(bytes_of(b), s));
    LCC_TRY(tw.alloc(bytes_of(out), s));
    da.data = ta.ptr; db.data = tb.ptr; w.data = tw.ptr;
    LCC_TRY(lcc_batch_unary(alg, LCC_UN_DUAL, 0, a, &da, (void *)s));
    LCC_TRY(lcc_batch_unary(alg, LCC_UN_DUAL, 0, b, &db, (void *)s));
    LCC_TRY(run_product(alg, kOpOuter, &da, &db, &w, nullptr, flags, (void *)s));
    return lcc_batch_unary(alg, LCC_UN_UNDUAL, 0, &w, out, (void *)s);
}

lcc_status lcc_batch_product(const lcc_algebra *alg, int op, const lcc_view *a, const lcc_view *b, const lcc_view *out,
                             unsigned flags, void *stream) {
    TraceRange tr("lcc_batch_product");
    switch (op) {
        case LCC_OP_GP: case LCC_OP_OUTER: case LCC_OP_LC: case LCC_OP_RC: case LCC_OP_SCALAR:
        case LCC_OP_COMMUTATOR: case LCC_OP_ANTICOMMUTATOR:
            return run_product(alg, op, a, b, out, nullptr, flags, stream);
        case LCC_OP_REGRESSIVE:
            return run_regressive(alg, a, b, out, flags, stream);
    }
    return fail(LCC_ERR_VALUE, "unknown product op %d", op);
}

lcc_status lcc_batch_gp_outer(const lcc_algebra *alg, const lcc_view *a, const lcc_view *b, const lcc_view *gp_out,
                              const lcc_view *outer_out, unsigned flags, void *stream) {
    TraceRange tr("lcc_batch_gp_outer");
    if (!outer_out) return fail(LCC_ERR_VALUE, "outer_out is NULL");
    return run_product(alg, kOpGPOuter, a, b, gp_out, outer_out, flags, stream);
}

int lcc_algebra_matrix_rep(const lcc_algebra *alg, uint8_t *string_of_blade, int8_t *sign_of_blade) {
    if (!alg || !matrix_rep_of(alg)) return 0;
    for (int A = 0; A < alg->nb; ++A) {
        if (string_of_blade) string_of_blade[A] = alg->rep_string_of[A];
        if (sign_of_blade) sign_of_blade[A] = alg->rep_sign_of[A];
    }
    return 16;
}

int lcc_algebra_matrix_rep_layout(const lcc_algebra *alg, uint32_t *tables) {
    if (!alg || !matrix_rep_of(alg)) return 0;
    const RepLayout L = make_rep_layout(alg->rep_blade_at.data(), alg->rep_sign_at.data());
    if (!L.ok) return 0;
    if (tables)
        for (int i = 0; i < 16; ++i) {
            tables[0 * 16 + i] = L.f_lo[i];        tables[1 * 16 + i] = L.f_hi[i];
            tables[2 * 16 + i] = L.column.fk[i];   tables[3 * 16 + i] = L.column.cw[i];   tables[4 * 16 + i] = L.column.kmap[i];
            tables[5 * 16 + i] = L.row.fk[i];      tables[6 * 16 + i] = L.row.cw[i];      tables[7 * 16 + i] = L.row.kmap[i];
        }
    return 1;
}

int lcc_algebra_matrix_rep_layout_aos(const lcc_algebra *alg, int elem_bytes, uint32_t *tables) {
    if (!alg || (elem_bytes != 4 && elem_bytes != 8) || !matrix_rep_of(alg)) return 0;
    const RepLayout L = make_rep_layout(alg->rep_blade_at.data(), alg->rep_sign_at.data());
    const RepLayoutAos A = make_rep_layout_aos(L, elem_bytes);
    if (!A.ok) return 0;
    if (tables)
        for (int i = 0; i < 16; ++i) {
            tables[0 * 16 + i] = A.g_lo[i];        tables[1 * 16 + i] = A.g_hi[i];        tables[2 * 16 + i] = A.amap[i];
            const RepPassAos *pass[2] = {&A.column, &A.row};
            for (int p = 0; p < 2; ++p) {
                tables[(3 + 5 * p) * 16 + i] = pass[p]->fk[i];   tables[(4 + 5 * p) * 16 + i] = pass[p]->fo[i];   tables[(5 + 5 * p) * 16 + i] = pass[p]->cw[i];
                tables[(6 + 5 * p) * 16 + i] = pass[p]->kmap[i]; tables[(7 + 5 * p) * 16 + i] = pass[p]->omap[i];
            }
        }
    return 1;
}

lcc_status lcc_batch_fixed_product(const lcc_algebra *alg, int op, int side, const double *fixed_host, const lcc_view *x, const lcc_view *out,
                                   unsigned flags, void *stream) {
    TraceRange tr("lcc_batch_fixed_product");
    LCC_TRY(check_view(alg, x, "x"));
    LCC_TRY(check_view(alg, out, "out"));
    LCC_TRY(same_kind(x, out, "fixed_product"));
    if (!fixed_host) return fail(LCC_ERR_VALUE, "fixed operand is NULL");
    if (side != 0 && side != 1) return fail(LCC_ERR_VALUE, "side must be 0 (fixed * x) or 1 (x * fixed)");
    if (x->n != out->n) return fail(LCC_ERR_VALUE, "fixed_product: output has %lld rows, expected %lld", (long long)out->n, (long long)x->n);
    if (x->n == 0) return LCC_OK;
    cudaStream_t s;
    LCC_TRY(resolve_stream(stream, &s));
    const bool strict = (flags & LCC_FLAG_STRICT_FP) != 0;
    const int nb = alg->nb;
    lcc_view xt = *x, ot = *out;
    if (op == LCC_OP_GP && matrix_rep_usable(alg, x, strict)) {
        double p1[256];
        rep_matrix_of(alg, fixed_host, 1.0 / 16.0, side == 1, p1);
        int64_t rows_done = 0;
        LCC_TRY(matrix_rep_apply(alg, side, p1, nullptr, x, out, s, &rows_done));
        if (rows_done == x->n) return LCC_OK;
        xt = tail_view(x, rows_done); ot = tail_view(out, rows_done);
    }
    // everything else: the operand becomes a one-row device batch (written by a kernel from its parameters: no host
    // staging, no synchronisation) and the broadcast branch of lcc_batch_product does the rest
    PoolBuffer row;
    LCC_TRY(row.alloc((size_t)2 * 256 * sizeof(double), s));
    double padded[256];
    for (int i = 0; i < 256; ++i) padded[i] = i < nb ? fixed_host[i] : 0.0;
    LCC_CUDA(launch_write_rotor_rows(nb, x->dtype, padded, row.ptr, s));
    lcc_view one{row.ptr, 1, x->layout == LCC_SOA ? 1 : nb, x->dtype, x->layout};
    if (x->layout == LCC_SOA) one.ld = 16 / (int64_t)elem_size(x->dtype);  // check_view wants a 16-byte multiple; blade k must sit at data[k * ld]
    if (x->layout == LCC_SOA) {
        // re-space the row so that blade k sits at data[k * ld]: write it as an AoS row and convert
        PoolBuffer spaced;
        LCC_TRY(spaced.alloc((size_t)nb * one.ld * elem_size(x->dtype), s));
        lcc_view src{row.ptr, 1, nb, x->dtype, LCC_AOS}, dst{spaced.ptr, 1, one.ld, x->dtype, LCC_SOA};
        LCC_CUDA(launch_convert(nb, arg_of(&src), arg_of(&dst), s));
        one.data = spaced.ptr;
        return side == 0 ? lcc_batch_product(alg, op, &one, &xt, &ot, flags, (void *)s) : lcc_batch_product(alg, op, &xt, &one, &ot, flags, (void *)s);
    }
    return side == 0 ? lcc_batch_product(alg, op, &one, &xt, &ot, flags, (void *)s) : lcc_batch_product(alg, op, &xt, &one, &ot, flags, (void *)s);
}

lcc_status lcc_batch_product_sum(const lcc_algebra *alg, int op, const lcc_view *a, int a_grade, const lcc_view *b, void *sums_dev,
                                 unsigned flags, void *stream) {
    TraceRange tr("lcc_batch_product_sum");
    if (op != LCC_OP_GP && op != LCC_OP_OUTER && op != LCC_OP_LC) return fail(LCC_ERR_VALUE, "unknown product op %d", op);
    LCC_TRY(check_view(alg, a, "a"));
    LCC_TRY(check_view(alg, b, "b"));
    LCC_TRY(same_kind(a, b, "product_sum"));
    if (!sums_dev) return fail(LCC_ERR_VALUE, "sums_dev is NULL");
    if (a_grade > alg->dim) return fail(LCC_ERR_VALUE, "grade %d exceeds dimension %d of algebra", a_grade, alg->dim);  // batch.rs:1029-1033
    int64_t n = 0;
    LCC_TRY(broadcast_rows(a->n, b->n, &n));
    cudaStream_t s;
    LCC_TRY(resolve_stream(stream, &s));
    const bool strict = (flags & LCC_FLAG_STRICT_FP) != 0;
    const bool a_bc = a->n == 1 && n > 1, b_bc = b->n == 1 && n > 1;
    if (alg->nb <= 32 && alg->product[a->dtype]) {
        PoolBuffer partial;
        LCC_TRY(partial.alloc((size_t)alg->nb * fused_sum_blocks() * sizeof(double), s));
        ProductArgs g;
        g.op = op; g.dtype = a->dtype; g.layout = a->layout; g.strict = strict; g.n = n;
        g.a = a->data; g.lda = a->ld; g.a_bcast = a_bc;
        g.b = b->data; g.ldb = b->ld; g.b_bcast = b_bc;
        g.mu = alg->specialised ? nullptr : alg->mu.data();
        g.sum_partial = (double *)partial.ptr; g.sum_blocks = fused_sum_blocks(); g.a_grade = a_grade < 0 ? -1 : a_grade;
        g.stream = s;
        LCC_CUDA(alg->product[a->dtype](g));
        LCC_CUDA(launch_sum_final(alg->nb, fused_sum_blocks(), (const double *)partial.ptr, sums_dev, a->dtype, s));
        return LCC_OK;
    }
    // large algebras: unfused composition through pool temporaries (grade -> product -> sum)
    PoolBuffer bg, bp;
    const size_t es = elem_size(a->dtype);
    lcc_view ga = *a, prod = (a->n >= b->n) ? *a : *b;
    prod.n = n;
    if (a_grade >= 0) {
        LCC_TRY(bg.alloc((size_t)(a->layout == LCC_SOA ? alg->nb * a->ld : a->n * a->ld) * es, s));
        ga.data = bg.ptr;
        LCC_TRY(lcc_batch_unary(alg, LCC_UN_GRADE, a_grade, a, &ga, (void *)s));
    }
    LCC_TRY(bp.alloc((size_t)(prod.layout == LCC_SOA ? alg->nb * prod.ld : prod.n * prod.ld) * es, s));
    prod.data = bp.ptr;
    LCC_TRY(run_product(alg, op, &ga, b, &prod, nullptr, flags, (void *)s));
    return lcc_batch_sum(alg, &prod, sums_dev, (void *)s);
}

lcc_status lcc_batch_product_vjp(const lcc_algebra *alg, int op, int wrt, const lcc_view *grad_out, const lcc_view *other,
                                 const lcc_view *grad_in, unsigned flags, void *stream) {
    TraceRange tr("lcc_batch_product_vjp");
    if (op != LCC_OP_GP && op != LCC_OP_OUTER && op != LCC_OP_LC) return fail(LCC_ERR_VALUE, "unknown product op %d", op);
    if (wrt != 0 && wrt != 1) return fail(LCC_ERR_VALUE, "wrt must be 0 (left operand) or 1 (right operand)");
    LCC_TRY(check_view(alg, grad_out, "grad_out"));
    LCC_TRY(check_view(alg, other, "other"));
    LCC_TRY(check_view(alg, grad_in, "grad_in"));
    LCC_TRY(same_kind(grad_out, other, "product_vjp"));
    LCC_TRY(same_kind(grad_out, grad_in, "product_vjp"));
    if (other->n != grad_out->n && other->n != 1)
        return fail(LCC_ERR_VALUE, "batch sizes must match or be 1 for broadcasting: %lld vs %lld", (long long)grad_out->n, (long long)other->n);
    if (grad_in->n != grad_out->n) return fail(LCC_ERR_VALUE, "product_vjp: grad_in has %lld rows, expected %lld", (long long)grad_in->n, (long long)grad_out->n);
    if (grad_out->n == 0) return LCC_OK;
    cudaStream_t s;
    LCC_TRY(resolve_stream(stream, &s));
    const bool strict = (flags & LCC_FLAG_STRICT_FP) != 0, other_bc = other->n == 1 && grad_out->n > 1;
    if (alg->nb <= 32 && alg->product[grad_out->dtype]) {
        // register kernels over the build-time VJP term lists (kernels/product_impl.cuh, LIST = 1 / 2): the cotangent
        // streams at the same HBM rate as the forward product
        ProductArgs g;
        g.op = op; g.dtype = grad_out->dtype; g.layout = grad_out->layout; g.strict = false; g.n = grad_out->n;  // cotangents always use FMA (no reference rounding order exists for them)
        g.list = 1 + wrt;
        const lcc_view *first = wrt == 0 ? grad_out : other, *second = wrt == 0 ? other : grad_out;
        g.a = first->data; g.lda = first->ld; g.a_bcast = wrt == 1 && other_bc;
        g.b = second->data; g.ldb = second->ld; g.b_bcast = wrt == 0 && other_bc;
        g.out = grad_in->data; g.ldo = grad_in->ld;
        g.mu = alg->specialised ? nullptr : alg->mu.data();
        g.stream = s;
        LCC_CUDA(alg->product[grad_out->dtype](g));
        return LCC_OK;
    }
    const int8_t *gs = nullptr;
    LCC_TRY(gather_signs_on_device(alg, 16 + 16 * wrt + op, &gs));
    if (wrt == 0) LCC_CUDA(launch_table_product(alg->nb, strict, arg_of(grad_out), false, arg_of(other), other_bc, arg_of(grad_in), gs, s));
    else LCC_CUDA(launch_table_product(alg->nb, strict, arg_of(other), other_bc, arg_of(grad_out), false, arg_of(grad_in), gs, s));
    return LCC_OK;
}

lcc_status lcc_batch_sandwich(const lcc_algebra *alg, const double *rotor, const lcc_view *x, const lcc_view *out,
                              unsigned flags, void *stream) {
    TraceRange tr("lcc_batch_sandwich");
    LCC_TRY(check_view(alg, x, "x"));
    LCC_TRY(check_view(alg, out, "out"));
    LCC_TRY(same_kind(x, out, "sandwich"));
    if (!rotor) return fail(LCC_ERR_VALUE, "rotor is NULL");
    if (x->n != out->n) return fail(LCC_ERR_VALUE, "sandwich: output has %lld rows, expected %lld", (long long)out->n, (long long)x->n);
    if (x->n == 0) return LCC_OK;
    cudaStream_t s;
    LCC_TRY(resolve_stream(stream, &s));
    const bool strict = (flags & LCC_FLAG_STRICT_FP) != 0;
    const int nb = alg->nb;
    if (nb <= 32) {
        SandwichLauncher launch = alg->sandwich[x->dtype];
        if (!launch) return fail(LCC_ERR_UNSUPPORTED, "no sandwich kernel built for %s", sig_str(alg).c_str());
        SandwichArgs g;
        g.dtype = x->dtype; g.layout = x->layout; g.strict = strict; g.n = x->n;
        g.even = true;
        for (int i = 0; i < nb; ++i)
            if ((popcount32((unsigned)i) & 1) && rotor[i] != 0.0) g.even = false;
        g.rotor = rotor; g.x = x->data; g.ldx = x->ld; g.out = out->data; g.ldo = out->ld;
        g.mu = alg->specialised ? nullptr : alg->mu.data();
        g.stream = s;
        LCC_CUDA(launch(g));
        return LCC_OK;
    }
    // 128 / 256 blades with a 16 x 16 matrix representation, FMA mode: R x ~R = rho(R) rho(x) rho(~R) per row (K8, HBM-bound)
    lcc_view xt = *x, ot = *out;  // what is left for the kernels below (all of it, or the 1-3 rows K8 cannot store)
    if (matrix_rep_usable(alg, x, strict)) {
        double p1[256], p2[256], rrev[256];
        for (int i = 0; i < nb; ++i) rrev[i] = rotor[i] * reverse_sign_of_grade(popcount32((unsigned)i));
        rep_matrix_of(alg, rotor, 1.0 / 16.0, false, p1);
        rep_matrix_of(alg, rrev, 1.0, true, p2);
        int64_t rows_done = 0;
        LCC_TRY(matrix_rep_apply(alg, 2, p1, p2, x, out, s, &rows_done));
        if (rows_done == x->n) return LCC_OK;
        xt = tail_view(x, rows_done); ot = tail_view(out, rows_done);
        x = &xt; out = &ot;
    }
    const int8_t *gs = nullptr;
    // > 32 blades, FMA mode: x -> R x ~R is ONE linear map; its nb x nb matrix is built on the device from the rotor
    // (kernel parameter, no host staging, no synchronisation) and the whole sandwich is one tensor-core GEMM
    if (!strict && !tensor_path_off() && (nb == 64 || nb == 128 || nb == 256)) {
        LCC_TRY(gather_signs_on_device(alg, 48, &gs));
        PoolBuffer w;
        LCC_TRY(w.alloc((size_t)nb * nb * sizeof(double), s));
        LCC_CUDA(launch_build_sandwich_matrix(nb, gs, rotor, (double *)w.ptr, s));
        bool handled = false;
        LCC_TRY(tensor_matrix_product(alg, (const double *)w.ptr, x, out, s, &handled));
        if (handled) return LCC_OK;
    }
    // strict mode (or views the GEMM cannot take): two exact-order table products with the rotor as a one-row batch
    // (batch.rs:462-496); R and ~R are written to the device by a kernel from its parameters
    const size_t es = elem_size(x->dtype);
    PoolBuffer br, brx;
    LCC_TRY(br.alloc((size_t)2 * nb * es, s));
    LCC_CUDA(launch_write_rotor_rows(nb, x->dtype, rotor, br.ptr, s));
    // one-row views of the rotor in x's layout: SoA with ld = 1 or AoS with ld = nb both put blade k at data[k]
    lcc_view r1{br.ptr, 1, x->layout == LCC_SOA ? 1 : nb, x->dtype, x->layout};
    lcc_view r2{(char *)br.ptr + nb * es, 1, r1.ld, x->dtype, x->layout};
    lcc_view rx = *x;  // intermediate R*x with x's shape
    LCC_TRY(brx.alloc((size_t)(x->layout == LCC_SOA ? nb * x->ld : x->n * x->ld) * es, s));
    rx.data = brx.ptr;
    LCC_TRY(gather_signs_on_device(alg, kOpGP, &gs));
    LCC_CUDA(launch_table_product(nb, strict, arg_of(&r1), true, arg_of(x), false, arg_of(&rx), gs, s));
    LCC_CUDA(launch_table_product(nb, strict, arg_of(&rx), false, arg_of(&r2), true, arg_of(out), gs, s));
    return LCC_OK;
}

lcc_status lcc_batch_sandwich_rows(const lcc_algebra *alg, const lcc_view *rotors, const lcc_view *x, const lcc_view *out,
                                   unsigned flags, void *stream) {
    TraceRange tr("lcc_batch_sandwich_rows");
    LCC_TRY(check_view(alg, rotors, "rotors"));
    LCC_TRY(check_view(alg, x, "x"));
    LCC_TRY(check_view(alg, out, "out"));
    if (rotors->n != x->n && rotors->n != 1) return fail(LCC_ERR_VALUE, "batch sizes must match or be 1 for broadcasting: %lld vs %lld", (long long)rotors->n, (long long)x->n);
    cudaStream_t s;
    LCC_TRY(resolve_stream(stream, &s));
    LCC_TRY(same_kind(rotors, x, "sandwich_rows"));
    LCC_TRY(same_kind(x, out, "sandwich_rows"));
    if (x->n != out->n) return fail(LCC_ERR_VALUE, "sandwich_rows: output has %lld rows, expected %lld", (long long)out->n, (long long)x->n);
    if (x->n == 0) return LCC_OK;
    // up to 16 blades: ONE kernel, R and x in registers per row, rx rounded in registers, ~R folded into the signs
    // (K2r; 3 * nb * s bytes per row).  cudaErrorNotSupported (32 blades and more): the composition below.
    if (alg->nb <= 32 && alg->sandwich[x->dtype]) {
        SandwichArgs g;
        g.dtype = x->dtype; g.layout = x->layout; g.strict = (flags & LCC_FLAG_STRICT_FP) != 0; g.n = x->n;
        g.rotors = rotors->data; g.ldr = rotors->ld; g.r_bcast = rotors->n == 1 && x->n > 1;
        g.x = x->data; g.ldx = x->ld; g.out = out->data; g.ldo = out->ld;
        g.mu = alg->specialised ? nullptr : alg->mu.data();
        g.stream = s;
        const cudaError_t e = alg->sandwich[x->dtype](g);
        if (e != cudaErrorNotSupported) { LCC_CUDA(e); return LCC_OK; }
    }
    // 128 / 256 blades, ONE device rotor for the batch, an algebra with a real 16 x 16 representation: K8's two-pass form,
    // both factors (rho(R) / 16, rho(~R)^T) built from the device row by two one-block kernels -- one pass over x instead
    // of reverse + two products through a full-size temporary
    lcc_view x_tail, out_tail;
    if (rotors->n == 1 && x->n > 1 && x->layout == out->layout && rotors->layout == x->layout &&
        matrix_rep_usable(alg, x, (flags & LCC_FLAG_STRICT_FP) != 0)) {
        int64_t rows_done = 0;
        LCC_TRY(matrix_rep_apply(alg, 2, nullptr, nullptr, x, out, s, &rows_done, rotors));
        if (rows_done == x->n) return LCC_OK;
        if (rows_done > 0) {  // blade-major views: the last 1-3 rows take the composition below
            x_tail = tail_view(x, rows_done); out_tail = tail_view(out, rows_done);
            x = &x_tail; out = &out_tail;
        }
    }
    // rx = R * x ; out = rx * ~R   (lcc/torch/multivector.py:1006-1017), composed from K1 and K3
    PoolBuffer brev, brx;
    lcc_view rev = *rotors, rx = *x;
    const size_t es = elem_size(x->dtype);
    const int64_t rows_r = rotors->layout == LCC_SOA ? rotors->ld : rotors->n;
    const int64_t rows_x = x->layout == LCC_SOA ? x->ld : x->n;
    LCC_TRY(brev.alloc((size_t)(rotors->layout == LCC_SOA ? alg->nb * rows_r : rows_r * rotors->ld) * es, s));
    LCC_TRY(brx.alloc((size_t)(x->layout == LCC_SOA ? alg->nb * rows_x : rows_x * x->ld) * es, s));
    rev.data = brev.ptr; rx.data = brx.ptr;
    LCC_TRY(lcc_batch_unary(alg, LCC_UN_REVERSE, 0, rotors, &rev, (void *)s));
    LCC_TRY(run_product(alg, kOpGP, rotors, x, &rx, nullptr, flags, (void *)s));
    return run_product(alg, kOpGP, &rx, &rev, out, nullptr, flags, (void *)s);
}

lcc_status lcc_batch_sandwich_rows_vjp(const lcc_algebra *alg, const lcc_view *rotors, const lcc_view *x, const lcc_view *grad_out,
                                       const lcc_view *grad_x, const lcc_view *grad_rotors, unsigned flags, void *stream) {
    TraceRange tr("lcc_batch_sandwich_rows_vjp");
    LCC_TRY(check_view(alg, rotors, "rotors"));
    LCC_TRY(check_view(alg, x, "x"));
    LCC_TRY(check_view(alg, grad_out, "grad_out"));
    LCC_TRY(check_view(alg, grad_x, "grad_x"));
    LCC_TRY(check_view(alg, grad_rotors, "grad_rotors"));
    if (rotors->n != x->n && rotors->n != 1) return fail(LCC_ERR_VALUE, "batch sizes must match or be 1 for broadcasting: %lld vs %lld", (long long)rotors->n, (long long)x->n);
    LCC_TRY(same_kind(rotors, x, "sandwich_rows_vjp"));
    LCC_TRY(same_kind(x, grad_out, "sandwich_rows_vjp"));
    LCC_TRY(same_kind(x, grad_x, "sandwich_rows_vjp"));
    LCC_TRY(same_kind(x, grad_rotors, "sandwich_rows_vjp"));
    if (grad_out->n != x->n || grad_x->n != x->n || grad_rotors->n != x->n)
        return fail(LCC_ERR_VALUE, "sandwich_rows_vjp: grad_out, grad_x and grad_rotors must have %lld rows", (long long)x->n);
    if (x->n == 0) return LCC_OK;
    cudaStream_t s;
    LCC_TRY(resolve_stream(stream, &s));
    // up to 16 blades: ONE kernel (product_impl.cuh, sandwich_rows_vjp_kernel): 5 * nb * s bytes per row
    if (alg->nb <= 32 && alg->sandwich[x->dtype]) {
        SandwichArgs g;
        g.dtype = x->dtype; g.layout = x->layout; g.strict = false; g.n = x->n;
        g.rotors = rotors->data; g.ldr = rotors->ld; g.r_bcast = rotors->n == 1 && x->n > 1;
        g.x = x->data; g.ldx = x->ld;
        g.grad = grad_out->data; g.ldg = grad_out->ld;
        g.gx = grad_x->data; g.ldgx = grad_x->ld; g.gr = grad_rotors->data; g.ldgr = grad_rotors->ld;
        g.mu = alg->specialised ? nullptr : alg->mu.data();
        g.stream = s;
        const cudaError_t e = alg->sandwich[x->dtype](g);
        if (e != cudaErrorNotSupported) { LCC_CUDA(e); return LCC_OK; }
    }
    // larger algebras: the same five bilinear passes composed from the product and cotangent entry points, with
    // t = R x recomputed (nothing was saved by the forward call) and three pool temporaries
    const size_t es = elem_size(x->dtype);
    auto full_bytes = [&](const lcc_view *v) { return (size_t)(v->layout == LCC_SOA ? (int64_t)alg->nb * v->ld : v->n * v->ld) * es; };
    PoolBuffer brev, bt, bgt;
    lcc_view rev = *rotors, t = *x, gt = *x;
    LCC_TRY(brev.alloc(full_bytes(rotors), s));
    LCC_TRY(bt.alloc(full_bytes(x), s));
    LCC_TRY(bgt.alloc(full_bytes(x), s));
    rev.data = brev.ptr; t.data = bt.ptr; gt.data = bgt.ptr;
    void *vs = (void *)s;
    LCC_TRY(lcc_batch_unary(alg, LCC_UN_REVERSE, 0, rotors, &rev, vs));
    LCC_TRY(run_product(alg, kOpGP, rotors, x, &t, nullptr, flags & ~LCC_FLAG_STRICT_FP, vs));         // t = R x
    LCC_TRY(lcc_batch_product_vjp(alg, LCC_OP_GP, 0, grad_out, &rev, &gt, flags, vs));                  // g_t = d(t ~R)/dt ^T g
    LCC_TRY(lcc_batch_product_vjp(alg, LCC_OP_GP, 1, grad_out, &t, grad_rotors, flags, vs));            // g_(~R), held in grad_rotors
    LCC_TRY(lcc_batch_unary(alg, LCC_UN_REVERSE, 0, grad_rotors, grad_rotors, vs));                     // rev . g_(~R)
    LCC_TRY(lcc_batch_product_vjp(alg, LCC_OP_GP, 0, &gt, x, &t, flags, vs));                           // d(R x)/dR ^T g_t (t is free again)
    LCC_TRY(lcc_batch_elementwise(alg, LCC_EW_ADD, grad_rotors, &t, grad_rotors, vs));
    return lcc_batch_product_vjp(alg, LCC_OP_GP, 1, &gt, rotors, grad_x, flags, vs);                    // d(R x)/dx ^T g_t
}

lcc_status lcc_batch_unary(const lcc_algebra *alg, int op, int arg, const lcc_view *x, const lcc_view *out, void *stream) {
    TraceRange tr("lcc_batch_unary");
    LCC_TRY(check_view(alg, x, "x"));
    LCC_TRY(check_view(alg, out, "out"));
    LCC_TRY(same_kind(x, out, "unary"));
    if (x->n != out->n) return fail(LCC_ERR_VALUE, "unary: output has %lld rows, expected %lld", (long long)out->n, (long long)x->n);
    const int nb = alg->nb, ps = nb - 1;
    auto S = [&](int a, int b) { return (int)alg->signs[(size_t)a * nb + b]; };
    BladeMap map;
    for (int k = 0; k < 256; ++k) { map.factor[k] = 0; map.source[k] = (uint16_t)(k < nb ? k : 0); }
    for (int k = 0; k < nb; ++k) {
        const int g = popcount32((unsigned)k);
        switch (op) {
            case LCC_UN_REVERSE: map.factor[k] = (int8_t)reverse_sign_of_grade(g); break;
            case LCC_UN_INVOLUTE: map.factor[k] = (int8_t)involution_sign_of_grade(g); break;
            case LCC_UN_CONJUGATE: map.factor[k] = (int8_t)conjugate_sign_of_grade(g); break;
            case LCC_UN_GRADE: map.factor[k] = g == arg ? 1 : 0; break;
            case LCC_UN_EVEN: map.factor[k] = (g % 2 == 0) ? 1 : 0; break;
            case LCC_UN_ODD: map.factor[k] = (g % 2 == 1) ? 1 : 0; break;
            case LCC_UN_COPY: map.factor[k] = 1; break;
            // out[i^ps] = S[i][ps]*S[ps][ps]*a[i]  (A * I^-1, I^-1 = S[ps][ps]*e_ps; multivector.rs:4249-4255,2989-2999)
            case LCC_UN_DUAL: map.source[k] = (uint16_t)(k ^ ps); map.factor[k] = (int8_t)(S(k ^ ps, ps) * S(ps, ps)); break;
            case LCC_UN_UNDUAL: map.source[k] = (uint16_t)(k ^ ps); map.factor[k] = (int8_t)S(k ^ ps, ps); break;  // A * I, :4270-4274
            case LCC_UN_RIGHT_DUAL: map.source[k] = (uint16_t)(k ^ ps); map.factor[k] = (int8_t)(S(ps, k ^ ps) * S(ps, ps)); break;  // I^-1 * A, :4285-4290
            // out[15^i] = reverse_sign(grade i) * a[i]  (src/pga.rs:534-549)
            case LCC_UN_PGA_DUAL: map.source[k] = (uint16_t)(k ^ ps); map.factor[k] = (int8_t)reverse_sign_of_grade(popcount32((unsigned)(k ^ ps))); break;
            default: return fail(LCC_ERR_VALUE, "unknown unary op %d", op);
        }
    }
    if (op == LCC_UN_GRADE && (arg < 0 || arg > alg->dim))  // batch.rs:1029-1033
        return fail(LCC_ERR_VALUE, "grade %d exceeds dimension %d of algebra", arg, alg->dim);
    if ((op == LCC_UN_DUAL || op == LCC_UN_RIGHT_DUAL) && !alg->pseudoscalar_invertible)  // multivector.rs:2991-2996
        return fail(LCC_ERR_VALUE, "cannot invert multivector with zero norm; the pseudoscalar of %s squares to 0 "
                    "(use pga_dual for PGA)", sig_str(alg).c_str());
    if (op == LCC_UN_PGA_DUAL && !(alg->p == 3 && alg->q == 0 && alg->r == 1))
        return fail(LCC_ERR_VALUE, "pga_dual requires a PGA3D element");
    if (x->n == 0) return LCC_OK;
    cudaStream_t s;
    LCC_TRY(resolve_stream(stream, &s));
    LCC_CUDA(launch_blade_map(nb, arg_of(x), arg_of(out), map, 1.0, s));
    return LCC_OK;
}

lcc_status lcc_batch_scale(const lcc_algebra *alg, const lcc_view *x, double factor, const lcc_view *out, void *stream) {
    TraceRange tr("lcc_batch_scale");
    LCC_TRY(check_view(alg, x, "x"));
    LCC_TRY(check_view(alg, out, "out"));
    LCC_TRY(same_kind(x, out, "scale"));
    if (x->n != out->n) return fail(LCC_ERR_VALUE, "scale: output has %lld rows, expected %lld", (long long)out->n, (long long)x->n);
    if (x->n == 0) return LCC_OK;
    BladeMap map;
    for (int k = 0; k < 256; ++k) { map.factor[k] = k < alg->nb ? 1 : 0; map.source[k] = (uint16_t)(k < alg->nb ? k : 0); }
    cudaStream_t s;
    LCC_TRY(resolve_stream(stream, &s));
    LCC_CUDA(launch_blade_map(alg->nb, arg_of(x), arg_of(out), map, factor, s));
    return LCC_OK;
}

lcc_status lcc_batch_elementwise(const lcc_algebra *alg, int op, const lcc_view *a, const lcc_view *b, const lcc_view *out, void *stream) {
    TraceRange tr("lcc_batch_elementwise");
    LCC_TRY(check_view(alg, a, "a"));
    LCC_TRY(check_view(alg, b, "b"));
    LCC_TRY(check_view(alg, out, "out"));
    LCC_TRY(same_kind(a, b, "elementwise"));
    LCC_TRY(same_kind(a, out, "elementwise"));
    if (op != LCC_EW_ADD && op != LCC_EW_SUB) return fail(LCC_ERR_VALUE, "unknown elementwise op %d", op);
    int64_t n = 0;
    LCC_TRY(broadcast_rows(a->n, b->n, &n));
    if (out->n != n) return fail(LCC_ERR_VALUE, "elementwise: output has %lld rows, expected %lld", (long long)out->n, (long long)n);
    if (n == 0) return LCC_OK;
    cudaStream_t s;
    LCC_TRY(resolve_stream(stream, &s));
    LCC_CUDA(launch_addsub(alg->nb, op, arg_of(a), a->n == 1 && n > 1, arg_of(b), b->n == 1 && n > 1, arg_of(out), s));
    return LCC_OK;
}

lcc_status lcc_batch_norm(const lcc_algebra *alg, int kind, const lcc_view *x, void *norms, unsigned flags, void *stream) {
    TraceRange tr("lcc_batch_norm");
    LCC_TRY(check_view(alg, x, "x"));
    if (kind != LCC_NORM_SQUARED && kind != LCC_NORM) return fail(LCC_ERR_VALUE, "unknown norm kind %d", kind);
    if (x->n == 0) return LCC_OK;
    if (!norms) return fail(LCC_ERR_VALUE, "norms is NULL");
    cudaStream_t s;
    LCC_TRY(resolve_stream(stream, &s));
    ViewArg none{nullptr, 0, 0, x->dtype, x->layout};
    LCC_CUDA(launch_norm(alg->nb, kind == LCC_NORM_SQUARED ? 0 : 1, (flags & LCC_FLAG_STRICT_FP) != 0, arg_of(x),
                         alg->norm_weights.data(), norms, none, s));
    return LCC_OK;
}

lcc_status lcc_batch_normalized(const lcc_algebra *alg, const lcc_view *x, const lcc_view *out, unsigned flags, void *stream) {
    TraceRange tr("lcc_batch_normalized");
    LCC_TRY(check_view(alg, x, "x"));
    LCC_TRY(check_view(alg, out, "out"));
    LCC_TRY(same_kind(x, out, "normalized"));
    if (x->n != out->n) return fail(LCC_ERR_VALUE, "normalized: output has %lld rows, expected %lld", (long long)out->n, (long long)x->n);
    if (x->n == 0) return LCC_OK;
    cudaStream_t s;
    LCC_TRY(resolve_stream(stream, &s));
    LCC_CUDA(launch_norm(alg->nb, 2, (flags & LCC_FLAG_STRICT_FP) != 0, arg_of(x), alg->norm_weights.data(), nullptr, arg_of(out), s));
    return LCC_OK;
}

lcc_status lcc_batch_sum(const lcc_algebra *alg, const lcc_view *x, void *sums_dev, void *stream) {
    TraceRange tr("lcc_batch_sum");
    LCC_TRY(check_view(alg, x, "x"));
    if (!sums_dev) return fail(LCC_ERR_VALUE, "sums_dev is NULL");
    cudaStream_t s;
    LCC_TRY(resolve_stream(stream, &s));
    PoolBuffer scratch;
    const int64_t elems = sum_scratch_elems(alg->nb, x->n);
    LCC_TRY(scratch.alloc((size_t)elems * sizeof(double), s));
    LCC_CUDA(launch_sum(alg->nb, arg_of(x), sums_dev, scratch.ptr, elems, s));
    return LCC_OK;
}

// ================================================================== PGA / CGA points
static lcc_status points_check(const lcc_algebra *alg, int which) {
    if (which == 0 && !(alg->p == 3 && alg->q == 0 && alg->r == 1))
        return fail(LCC_ERR_VALUE, "pga_point requires a PGA3D algebra Cl(3,0,1)");  // src/pga.rs:110-114
    if (which == 1 && !(alg->q == 1 && alg->r == 0 && alg->dim == 5))
        return fail(LCC_ERR_VALUE, "cga_point(x,y,z) requires CGA3D (dimension 5); use Algebra.cga(3)");  // src/cga.rs:172-184
    if ((which == 2 || which == 3) && !(alg->p == 1 && alg->q == 3 && alg->r == 0))
        return fail(LCC_ERR_VALUE, "requires an STA algebra Cl(1,3,0); use Algebra.sta()");  // src/sta.rs:60-70
    return LCC_OK;
}
static lcc_status points_embed(const lcc_algebra *alg, int which, const void *xyz, const lcc_view *dst, void *stream) {
    LCC_TRY(check_view(alg, dst, "dst"));
    LCC_TRY(points_check(alg, which));
    if (dst->n == 0) return LCC_OK;
    if (!xyz) return fail(LCC_ERR_VALUE, "xyz is NULL");
    cudaStream_t s;
    LCC_TRY(resolve_stream(stream, &s));
    LCC_CUDA(launch_points_embed(which, xyz, arg_of(dst), s));
    return LCC_OK;
}
static lcc_status points_extract(const lcc_algebra *alg, int which, const lcc_view *src, void *xyz, void *stream) {
    LCC_TRY(check_view(alg, src, "src"));
    LCC_TRY(points_check(alg, which));
    if (src->n == 0) return LCC_OK;
    if (!xyz) return fail(LCC_ERR_VALUE, "xyz is NULL");
    cudaStream_t s;
    LCC_TRY(resolve_stream(stream, &s));
    LCC_CUDA(launch_points_extract(which, arg_of(src), xyz, s));
    return LCC_OK;
}
lcc_status lcc_pga_points_embed(const lcc_algebra *alg, const void *xyz, const lcc_view *dst, void *stream) { return points_embed(alg, 0, xyz, dst, stream); }
lcc_status lcc_pga_points_extract(const lcc_algebra *alg, const lcc_view *src, void *xyz, void *stream) { return points_extract(alg, 0, src, xyz, stream); }
lcc_status lcc_cga_points_embed(const lcc_algebra *alg, const void *xyz, const lcc_view *dst, void *stream) { return points_embed(alg, 1, xyz, dst, stream); }
lcc_status lcc_cga_points_extract(const lcc_algebra *alg, const lcc_view *src, void *xyz, void *stream) { return points_extract(alg, 1, src, xyz, stream); }
lcc_status lcc_pga_points_sandwich(const lcc_algebra *alg, int dtype, const double *motor, const void *xyz_in, void *xyz_out,
                                   int64_t n, unsigned flags, void *stream) {
    TraceRange tr("lcc_pga_points_sandwich");
    if (!alg) return fail(LCC_ERR_VALUE, "algebra handle is NULL");
    LCC_TRY(points_check(alg, 0));
    if (dtype != LCC_F32 && dtype != LCC_F64) return fail(LCC_ERR_VALUE, "dtype must be LCC_F32 or LCC_F64");
    if (!motor) return fail(LCC_ERR_VALUE, "motor is NULL");
    if (n < 0) return fail(LCC_ERR_VALUE, "n must be >= 0");
    if (n == 0) return LCC_OK;
    if (!xyz_in || !xyz_out) return fail(LCC_ERR_VALUE, "xyz is NULL");
    cudaStream_t s;
    LCC_TRY(resolve_stream(stream, &s));
    SandwichArgs g;
    g.dtype = dtype; g.strict = (flags & LCC_FLAG_STRICT_FP) != 0; g.n = n;
    g.even = true;
    for (int i = 0; i < 16; ++i)
        if ((popcount32((unsigned)i) & 1) && motor[i] != 0.0) g.even = false;
    g.rotor = motor; g.x = xyz_in; g.out = xyz_out; g.stream = s;
    LCC_CUDA(dtype == LCC_F32 ? launch_pga_points_sandwich_float(g) : launch_pga_points_sandwich_double(g));
    return LCC_OK;
}
lcc_status lcc_sta_bivectors_embed(const lcc_algebra *alg, const void *eb, const lcc_view *dst, void *stream) { return points_embed(alg, 2, eb, dst, stream); }
lcc_status lcc_sta_boosts_embed(const lcc_algebra *alg, const void *velocity, const lcc_view *dst, void *stream) { return points_embed(alg, 3, velocity, dst, stream); }

// (host-buffer entry points, lcc_batch_upload / download and the pinned host cache: host_transfer.cu;
//  multi-GPU row sharding + NCCL: shard_group.cu)

// ================================================================== tuning
lcc_status lcc_set_option(const char *name, int64_t value) {
    if (!name) return fail(LCC_ERR_VALUE, "option name is NULL");
    if (value < 0) return fail(LCC_ERR_VALUE, "option %s: value must be >= 0", name);
    if (!strcmp(name, "aos_tma_min_rows")) { g_aos_tma_min_rows = value; return LCC_OK; }
    if (!strcmp(name, "soa_tma_min_bytes")) { g_soa_tma_min_bytes = value; return LCC_OK; }
    if (!strcmp(name, "matrix_rep")) { g_matrix_rep = value; return LCC_OK; }
    if (!strcmp(name, "host_chunk_bytes")) { set_option_host_chunk_bytes(value); return LCC_OK; }
    if (!strcmp(name, "pinned_output_min_bytes")) { set_option_pinned_output_min_bytes(value); return LCC_OK; }
    return fail(LCC_ERR_VALUE, "unknown option %s", name);
}
lcc_status lcc_get_option(const char *name, int64_t *value) {
    if (!name || !value) return fail(LCC_ERR_VALUE, "option name / value is NULL");
    if (!strcmp(name, "aos_tma_min_rows")) { *value = g_aos_tma_min_rows; return LCC_OK; }
    if (!strcmp(name, "soa_tma_min_bytes")) { *value = g_soa_tma_min_bytes; return LCC_OK; }
    if (!strcmp(name, "matrix_rep")) { *value = g_matrix_rep; return LCC_OK; }
    if (!strcmp(name, "host_chunk_bytes")) { *value = option_host_chunk_bytes(); return LCC_OK; }
    if (!strcmp(name, "pinned_output_min_bytes")) { *value = option_pinned_output_min_bytes(); return LCC_OK; }
    return fail(LCC_ERR_VALUE, "unknown option %s", name);
}

// ================================================================== instrumentation
uint64_t lcc_kernel_launch_count(void) { return t_launches; }
void lcc_kernel_launch_count_reset(void) { t_launches = 0; }

}  // extern "C"
