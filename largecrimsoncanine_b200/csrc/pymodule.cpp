// pymodule.cpp -- CPython extension `largecrimsoncanine`: the host-side mirror of the reference's
// PyO3 boundary (/root/reference/src/lib.rs:37-45) over the C ABI of include/lcc_b200.h.
//
//   Algebra          <- src/pyalgebra.rs:30-200
//   Multivector      <- the host value type of src/multivector.rs.  ONE-pair products run in a host C++ loop that
//                       restates the reference's per-object product (host_pair_product below) --
//                       a single pair is ~1 us of arithmetic, not GPU work.  Nothing batched ever takes that loop.
//   MultivectorBatch <- src/batch.rs:39-1098, device resident (structure-of-arrays in HBM)
//
// Same names, argument meaning, exception types and message substrings as the reference so its own
// tests run unmodified (tests/test_reference_suites.py).  Written in C++ because the reference's
// host layer is compiled Rust and this image has no Rust toolchain.
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <algorithm>
#include <array>
#include <chrono>
#include <cmath>
#include <limits>
#include <optional>
#include <cstdlib>
#include <map>
#include <memory>
#include <mutex>
#include <cstring>
#include <sstream>
#include <string>
#include <vector>

#include "lcc_b200.h"

namespace py = pybind11;

namespace {

// ------------------------------------------------------------------ error translation
[[noreturn]] void raise_status(lcc_status st) {
    const std::string msg = lcc_last_error();
    switch (st) {
        case LCC_ERR_VALUE: throw py::value_error(msg);
        case LCC_ERR_INDEX: throw py::index_error(msg);
        case LCC_ERR_ZERODIV: { py::gil_scoped_acquire gil; PyErr_SetString(PyExc_ZeroDivisionError, msg.c_str()); throw py::error_already_set(); }
        case LCC_ERR_ALLOC: throw std::bad_alloc();  // pybind11 maps this to MemoryError
        default: throw std::runtime_error(msg);
    }
}
inline void check(lcc_status st) { if (st != LCC_OK) raise_status(st); }

bool g_strict_fp = [] { const char *e = std::getenv("LCC_STRICT_FP"); return e && e[0] && e[0] != '0'; }();
inline unsigned fp_flags() { return g_strict_fp ? LCC_FLAG_STRICT_FP : 0u; }

inline int popcount(unsigned v) { return __builtin_popcount(v); }
inline double reverse_sign(int grade) { return (grade % 4 < 2) ? 1.0 : -1.0; }

// ------------------------------------------------------------------ Algebra
struct AlgebraHandle {
    lcc_algebra *h = nullptr;
    int p, q, r;
    AlgebraHandle(int p_, int q_, int r_) : p(p_), q(q_), r(r_) { check(lcc_algebra_create(p_, q_, r_, &h)); }
    ~AlgebraHandle() { lcc_algebra_destroy(h); }
    AlgebraHandle(const AlgebraHandle &) = delete;
    int dim() const { return p + q + r; }
    int nb() const { return 1 << dim(); }
    bool same_signature(const AlgebraHandle &o) const { return p == o.p && q == o.q && r == o.r; }
    bool is_euclidean() const { return q == 0 && r == 0; }
    bool is_pga() const { return q == 0 && r == 1; }
    std::string str() const { std::ostringstream s; s << "Cl(" << p << "," << q << "," << r << ")"; return s.str(); }
};
using AlgebraPtr = std::shared_ptr<AlgebraHandle>;

struct PyAlgebra { AlgebraPtr inner; };

PyAlgebra make_algebra(long p, long q, long r) {
    if (p < 0 || q < 0 || r < 0) throw py::type_error("signature entries must be non-negative integers");
    return PyAlgebra{std::make_shared<AlgebraHandle>((int)p, (int)q, (int)r)};
}

// get_euclidean, src/algebra/registry.rs:275-303
AlgebraPtr cached_euclidean(int n) {
    static std::mutex m;
    static std::map<int, AlgebraPtr> cache;
    std::lock_guard<std::mutex> lock(m);
    auto it = cache.find(n);
    if (it != cache.end()) return it->second;
    auto a = std::make_shared<AlgebraHandle>(n, 0, 0);
    cache[n] = a;
    return a;
}

// ------------------------------------------------------------------ one-row device products (Multivector)
// Runs op(a, b) for ONE pair through lcc_batch_product on the GPU (array-of-structures views, n = 1).
std::vector<double> device_pair_product(const AlgebraPtr &alg, int op, const std::vector<double> &a, const std::vector<double> &b) {
    const int nb = alg->nb();
    const size_t bytes = (size_t)nb * sizeof(double);
    void *buf = nullptr;
    check(lcc_device_alloc(&buf, 3 * bytes, nullptr));
    std::vector<double> out(nb);
    char *base = (char *)buf;
    lcc_status st = lcc_memcpy_h2d(base, a.data(), bytes, nullptr);
    if (st == LCC_OK) st = lcc_memcpy_h2d(base + bytes, b.data(), bytes, nullptr);
    lcc_view va{base, 1, nb, LCC_F64, LCC_AOS}, vb{base + bytes, 1, nb, LCC_F64, LCC_AOS}, vo{base + 2 * bytes, 1, nb, LCC_F64, LCC_AOS};
    if (st == LCC_OK) st = lcc_batch_product(alg->h, op, &va, &vb, &vo, fp_flags(), nullptr);
    if (st == LCC_OK) st = lcc_memcpy_d2h(out.data(), base + 2 * bytes, bytes, nullptr);
    if (st == LCC_OK) st = lcc_stream_sync(nullptr);
    lcc_device_free(buf, nullptr);
    check(st);
    return out;
}

// ------------------------------------------------------------------ single-pair products on the host
// The reference's per-object product (src/multivector.rs:1438-1715) for ONE pair, restated in C++ (SURVEY.md section 2.1:
// "the single-Multivector CPU product keeps a scalar C++ loop").  A one-pair product is ~1 us of arithmetic; routing it
// through the GPU costs an allocation, two uploads, a launch, a download and a synchronisation (tens of microseconds).
// This is NOT a fallback of the batched path: MultivectorBatch / lcc_batch_* never come here and still refuse to run
// without a CUDA device.  Term order and rounding follow the reference exactly (compiled with -ffp-contract=off):
//   * geometric product, fewer than 16 blades: i outer, j inner, zero operands skipped, (sign*a)*b, sequential adds
//   * geometric product, 16 blades and more: the f64x4 path of src/simd.rs:55-135 -- (a*b)*sign, b == 0 is NOT skipped
//     and sign-0 terms are added as (+/-)0; identical to the scalar order except for (+/-)0 / NaN / Inf operands
//   * outer / left / right contraction / scalar product: the masked scalar loops of :1497-1715
// LCC_SINGLE_PAIR=device keeps the old behaviour (n = 1 through the CUDA kernels) for cross-checks.
bool single_pair_on_device() {
    static const bool dev = [] { const char *e = std::getenv("LCC_SINGLE_PAIR"); return e && std::string(e) == "device"; }();
    return dev;
}

std::vector<double> host_pair_product(const AlgebraPtr &alg, int op, const std::vector<double> &a, const std::vector<double> &b) {
    const int nb = alg->nb();
    const int8_t *signs = lcc_algebra_cayley_signs(alg->h);
    std::vector<double> out(nb, 0.0);
    if (op == LCC_OP_GP && nb >= 16) {  // src/simd.rs:21-30 (SIMD_THRESHOLD = 16), :55-135; nb is a power of two: no remainder loop
        for (int i = 0; i < nb; ++i) {
            const double ai = a[i];
            if (ai == 0.0) continue;
            const int8_t *row = signs + (size_t)i * nb;
            for (int j = 0; j < nb; ++j) out[i ^ j] += (ai * b[j]) * (double)row[j];
        }
        return out;
    }
    for (int i = 0; i < nb; ++i) {
        const double ai = a[i];
        if (ai == 0.0) continue;
        const int8_t *row = signs + (size_t)i * nb;
        for (int j = 0; j < nb; ++j) {
            const double bj = b[j];
            if (bj == 0.0) continue;
            switch (op) {
                case LCC_OP_GP: break;
                case LCC_OP_OUTER: if ((i & j) != 0) continue; break;                  // :1510-1513
                case LCC_OP_LC: if ((i & j) != i) continue; break;                     // :1570-1583 (the grade tests are implied by the subset test)
                case LCC_OP_RC: if ((i & j) != j) continue; break;                     // :1637-1650
                case LCC_OP_SCALAR: if (i != j) continue; break;                       // :1697-1700 (blade == 0 iff i == j)
                default: throw std::runtime_error("host_pair_product: unexpected op");
            }
            out[i ^ j] += ((double)row[j] * ai) * bj;
        }
    }
    return out;
}

// ------------------------------------------------------------------ Multivector (host value type)
struct Multivector {
    std::vector<double> coeffs;
    int dims = 0;
    AlgebraPtr algebra_opt;  // may be null: "legacy" Euclidean multivector (src/multivector.rs:24-45)

    AlgebraPtr get_algebra() const { return algebra_opt ? algebra_opt : cached_euclidean(dims); }

    // src/multivector.rs:48-55
    bool same_algebra(const Multivector &o) const {
        if (algebra_opt && o.algebra_opt) return algebra_opt.get() == o.algebra_opt.get() || algebra_opt->same_signature(*o.algebra_opt);
        if (!algebra_opt && !o.algebra_opt) return dims == o.dims;
        const AlgebraPtr &a = algebra_opt ? algebra_opt : o.algebra_opt;
        const int other_dims = algebra_opt ? o.dims : dims;
        return a->is_euclidean() && a->dim() == other_dims;
    }
    Multivector like(std::vector<double> c, bool keep_algebra = true) const { return Multivector{std::move(c), dims, keep_algebra ? algebra_opt : nullptr}; }

    static Multivector from_coeffs(std::vector<double> c) {  // src/multivector.rs:123-142
        const size_t len = c.size();
        if (len == 0) throw py::value_error("coeffs must not be empty; provide a list of coefficients (e.g., [1.0] for a scalar, [0.0, 1.0, 0.0, 0.0] for e1 in 2D)");
        if (len & (len - 1)) {
            size_t lower = 1; while (lower * 2 < len) lower *= 2;
            std::ostringstream s; s << "coeffs length must be a power of 2 (got " << len << ", expected " << lower << " or " << lower * 2 << ")";
            throw py::value_error(s.str());
        }
        int dims = 0; while ((size_t(1) << dims) < len) ++dims;
        return Multivector{std::move(c), dims, nullptr};
    }
    Multivector product(int op, const Multivector &o) const {  // src/multivector.rs:1438-1593
        if (!same_algebra(o)) {
            std::ostringstream s;
            s << "algebra mismatch: left operand is Cl(" << dims << ") but right operand is Cl(" << o.dims << "); both multivectors must be in the same algebra";
            throw py::value_error(s.str());
        }
        AlgebraPtr alg = get_algebra();
        if (single_pair_on_device()) return like(device_pair_product(alg, op, coeffs, o.coeffs));
        switch (op) {
            case LCC_OP_COMMUTATOR: case LCC_OP_ANTICOMMUTATOR: {  // src/multivector.rs:1732-1768: a*b, b*a, -/+, scale(0.5)
                std::vector<double> ab = host_pair_product(alg, LCC_OP_GP, coeffs, o.coeffs), ba = host_pair_product(alg, LCC_OP_GP, o.coeffs, coeffs);
                for (size_t k = 0; k < ab.size(); ++k) ab[k] = (op == LCC_OP_COMMUTATOR ? ab[k] - ba[k] : ab[k] + ba[k]) * 0.5;
                return like(std::move(ab));
            }
            case LCC_OP_REGRESSIVE: {  // :1781-1794: undual(dual(A) ^ dual(B)); inverse() raises when the pseudoscalar is null
                const Multivector ps = pseudoscalar_like(), ps_inv = ps.inverse();
                std::vector<double> da = host_pair_product(alg, LCC_OP_GP, coeffs, ps_inv.coeffs), db = host_pair_product(alg, LCC_OP_GP, o.coeffs, ps_inv.coeffs);
                std::vector<double> w = host_pair_product(alg, LCC_OP_OUTER, da, db);
                return like(host_pair_product(alg, LCC_OP_GP, w, ps.coeffs));
            }
            default: return like(host_pair_product(alg, op, coeffs, o.coeffs));
        }
    }
    // commutator / anticommutator / regressive test the dimension only (src/multivector.rs:1733-1739,1782-1788)
    Multivector product_same_dims(int op, const Multivector &o) const {
        if (dims != o.dims) {
            std::ostringstream s;
            s << "dimension mismatch: left operand is Cl(" << dims << ") but right operand is Cl(" << o.dims << "); both multivectors must have the same dimension";
            throw py::value_error(s.str());
        }
        return product(op, o);
    }
    Multivector reverse() const {  // src/multivector.rs:2711-2726
        std::vector<double> c(coeffs.size());
        for (size_t i = 0; i < c.size(); ++i) c[i] = coeffs[i] * reverse_sign(popcount((unsigned)i));
        return like(std::move(c));
    }
    Multivector scale(double s) const {  // src/multivector.rs:2153-2160
        std::vector<double> c(coeffs.size());
        for (size_t i = 0; i < c.size(); ++i) c[i] = coeffs[i] * s;
        return like(std::move(c));
    }
    // scalar part of A*~A, i ascending (src/multivector.rs:2890-2913); only j == i reaches blade 0
    double norm_squared() const {
        AlgebraPtr alg = get_algebra();
        const int8_t *signs = lcc_algebra_cayley_signs(alg->h);
        const int nb = alg->nb();
        double acc = 0.0;
        for (int i = 0; i < nb; ++i) {
            const double a = coeffs[i];
            if (a == 0.0) continue;
            const double b = a * reverse_sign(popcount((unsigned)i));
            const double t = (double)signs[(size_t)i * nb + i] * a;  // exact; (sign*a)*b then one add, no FMA (-ffp-contract=off)
            acc += t * b;
        }
        return acc;
    }
    double norm() const { return std::sqrt(std::fabs(norm_squared())); }
    Multivector grade(int k) const {  // src/multivector.rs:1269-1287 (drops the explicit algebra)
        if (k < 0 || k > dims) {
            std::ostringstream s; s << "grade " << k << " exceeds algebra dimension " << dims << " (max grade is " << dims << ")";
            throw py::value_error(s.str());
        }
        std::vector<double> c(coeffs.size(), 0.0);
        for (size_t i = 0; i < c.size(); ++i) if (popcount((unsigned)i) == k) c[i] = coeffs[i];
        return like(std::move(c), false);
    }
    std::vector<int> grades() const {  // src/multivector.rs:3153-3171
        std::vector<int> g;
        for (int k = 0; k <= dims; ++k)
            for (size_t i = 0; i < coeffs.size(); ++i)
                if (coeffs[i] != 0.0 && popcount((unsigned)i) == k) { g.push_back(k); break; }
        return g;
    }
    Multivector sandwich(const Multivector &x) const {  // src/multivector.rs:3040-3052
        if (dims != x.dims) {
            std::ostringstream s; s << "dimension mismatch: rotor is Cl(" << dims << ") but operand is Cl(" << x.dims << "); both must have the same dimension";
            throw py::value_error(s.str());
        }
        return product(LCC_OP_GP, x).product(LCC_OP_GP, reverse());
    }
    Multivector inverse() const {  // src/multivector.rs:2989-2999
        const double n2 = norm_squared();
        if (n2 == 0.0) throw py::value_error("cannot invert multivector with zero norm; only non-null blades and versors are invertible");
        return reverse().scale(1.0 / n2);
    }
    Multivector pseudoscalar_like() const { std::vector<double> c(coeffs.size(), 0.0); c.back() = 1.0; return like(std::move(c)); }
};

Multivector mv_in(const PyAlgebra &alg, std::vector<double> c) { return Multivector{std::move(c), alg.inner->dim(), alg.inner}; }

void require_pga3d(const PyAlgebra &a, const char *who) {
    if (!a.inner->is_pga() || a.inner->dim() != 4) throw py::value_error(std::string(who) + " requires a PGA3D algebra Cl(3,0,1)");
}
void require_pga3d_element(const Multivector &m, const char *who) {
    AlgebraPtr alg = m.get_algebra();
    if (!alg->is_pga() || m.dims != 4) throw py::value_error(std::string(who) + " requires a PGA3D element");
}

// ------------------------------------------------------------------ MultivectorBatch (device resident)
struct Batch {
    AlgebraPtr algebra;
    lcc_view view{nullptr, 0, 0, LCC_F64, LCC_SOA};
    int device = 0;
    // host copy of a ONE-row batch that came from host data (from_numpy / set): lets a broadcast product hand the fixed
    // operand to lcc_batch_fixed_product without reading device memory back (empty when unknown)
    std::vector<double> host_row;

    Batch(AlgebraPtr alg, int dtype, int64_t n) : algebra(std::move(alg)) {
        check(lcc_get_device(&device));
        check(lcc_batch_alloc(algebra->h, dtype, n, &view, nullptr));
    }
    ~Batch() { lcc_batch_free(&view, nullptr); }
    Batch(const Batch &) = delete;
    int nb() const { return algebra->nb(); }
    int64_t n() const { return view.n; }
    size_t es() const { return view.dtype == LCC_F32 ? 4 : 8; }
    std::shared_ptr<Batch> empty_like(int64_t rows) const { return std::make_shared<Batch>(algebra, view.dtype, rows); }
};
using BatchPtr = std::shared_ptr<Batch>;

struct DeviceTemp {  // stream-ordered scratch
    void *ptr = nullptr;
    explicit DeviceTemp(size_t bytes) { check(lcc_device_alloc(&ptr, bytes, nullptr)); }
    ~DeviceTemp() { lcc_device_free(ptr, nullptr); }
};

int dtype_of(const py::array &arr) {
    if (arr.dtype().is(py::dtype::of<double>())) return LCC_F64;
    if (arr.dtype().is(py::dtype::of<float>())) return LCC_F32;
    throw py::type_error("expected a float64 (or float32) NumPy array, got dtype " + std::string(py::str(arr.dtype())));
}
py::array contiguous(const py::array &arr) {
    return py::array::ensure(arr, py::array::c_style);
}
py::dtype np_dtype(int dtype) { return dtype == LCC_F32 ? py::dtype::of<float>() : py::dtype::of<double>(); }

// upload an (n, cols) host array as a dense device AoS block; caller frees
// (pointer and size are taken while the GIL is held; the copy itself runs without it)
void upload(const void *host, size_t bytes, void *dst) {
    check(lcc_memcpy_h2d(dst, host, bytes, nullptr));
    check(lcc_stream_sync(nullptr));  // the NumPy buffer may be pageable and may change after we return
}

BatchPtr batch_from_numpy(const PyAlgebra &alg, const py::array &coeffs_in) {  // src/batch.rs:79-94
    py::array coeffs = contiguous(coeffs_in);
    const int dtype = dtype_of(coeffs);
    if (coeffs.ndim() != 2) throw py::type_error("coeffs must be a 2-D array of shape (N, num_blades)");
    const int nb = alg.inner->nb();
    if (coeffs.shape(1) != nb) {
        std::ostringstream s; s << "coeffs must have " << nb << " columns for " << alg.inner->str() << " but got " << coeffs.shape(1);
        throw py::value_error(s.str());
    }
    const int64_t n = coeffs.shape(0);
    auto b = std::make_shared<Batch>(alg.inner, dtype, n);
    if (n == 0) return b;
    const void *host = coeffs.data();
    if (n == 1) {
        b->host_row.resize(nb);
        for (int k = 0; k < nb; ++k) b->host_row[k] = dtype == LCC_F64 ? ((const double *)host)[k] : (double)((const float *)host)[k];
    }
    py::gil_scoped_release nogil;
    // chunked pipeline (host_transfer.cu): the array is read exactly once -- directly by the DMA engine when it is
    // pinned, through the pinned bounce ring otherwise -- and each chunk is transposed into the blade-major batch while
    // the next one crosses PCIe.  Returns once the NumPy buffer has been consumed (it may change afterwards).
    check(lcc_batch_upload(alg.inner->h, host, n, &b->view, nullptr));
    return b;
}

BatchPtr batch_from_columns(const PyAlgebra &alg, const py::array &cols_in, const std::vector<int32_t> &blades, const char *what) {
    py::array cols = contiguous(cols_in);
    const int dtype = dtype_of(cols);
    const int64_t n = cols.shape(0);
    const int ncols = (int)blades.size();
    auto b = std::make_shared<Batch>(alg.inner, dtype, n);
    if (n == 0) return b;
    (void)what;
    const void *host = cols.data();
    const size_t bytes = (size_t)cols.nbytes();
    py::gil_scoped_release nogil;
    DeviceTemp stage(std::max<size_t>(bytes, 16));
    if (bytes) upload(host, bytes, stage.ptr);
    check(lcc_batch_scatter_columns(alg.inner->h, stage.ptr, ncols, blades.data(), &b->view, nullptr));
    return b;
}

py::array batch_gather(const Batch &b, const std::vector<int32_t> &blades, bool flat) {
    const int ncols = (int)blades.size();
    std::vector<py::ssize_t> shape;
    if (flat) shape = {(py::ssize_t)b.n()}; else shape = {(py::ssize_t)b.n(), (py::ssize_t)ncols};
    py::array out(np_dtype(b.view.dtype), shape);
    void *out_ptr = out.mutable_data();
    if (b.n() == 0 || ncols == 0) return out;
    py::gil_scoped_release nogil;
    DeviceTemp stage((size_t)b.n() * ncols * b.es());
    check(lcc_batch_gather_columns(b.algebra->h, &b.view, ncols, blades.data(), stage.ptr, nullptr));
    check(lcc_memcpy_d2h(out_ptr, stage.ptr, (size_t)b.n() * ncols * b.es(), nullptr));
    check(lcc_stream_sync(nullptr));
    return out;
}

// A NumPy array over a block of the pinned host cache (lcc_host_cache_*); the block returns to the cache when the
// array is garbage collected.  Used for large results so the D2H copy is a direct DMA into the array's memory.
py::array pinned_array(int dtype, const std::vector<py::ssize_t> &shape) {
    size_t bytes = dtype == LCC_F32 ? 4 : 8;
    for (auto d : shape) bytes *= (size_t)d;
    void *p = nullptr;
    check(lcc_host_cache_alloc(&p, bytes));
    py::capsule owner(p, [](void *q) { lcc_host_cache_free(q); });
    return py::array(np_dtype(dtype), shape, p, owner);
}
py::array output_array(int dtype, const std::vector<py::ssize_t> &shape) {
    size_t bytes = dtype == LCC_F32 ? 4 : 8;
    for (auto d : shape) bytes *= (size_t)d;
    int64_t min_bytes = 0;
    lcc_get_option("pinned_output_min_bytes", &min_bytes);
    if (min_bytes >= 0 && (int64_t)bytes >= min_bytes && bytes > 0) {
        try { return pinned_array(dtype, shape); } catch (const std::exception &) { /* no pinned memory left: pageable */ }
    }
    return py::array(np_dtype(dtype), shape);
}

py::array batch_to_numpy(const Batch &b, py::object out_obj = py::none()) {  // src/batch.rs:233-235
    const std::vector<py::ssize_t> shape{(py::ssize_t)b.n(), (py::ssize_t)b.nb()};
    py::array out;
    if (out_obj.is_none()) {
        out = output_array(b.view.dtype, shape);
    } else {  // additive: export into a caller-provided (e.g. pinned) array
        out = py::array::ensure(out_obj);
        if (!out || dtype_of(out) != b.view.dtype || out.ndim() != 2 || out.shape(0) != shape[0] || out.shape(1) != shape[1] ||
            !(out.flags() & py::array::c_style) || !out.writeable())
            throw py::value_error("out must be a writeable C-contiguous (N, num_blades) array of the batch's dtype");
    }
    void *out_ptr = out.mutable_data();
    if (b.n() == 0) return out;
    py::gil_scoped_release nogil;
    check(lcc_batch_download(b.algebra->h, &b.view, out_ptr, nullptr));
    return out;
}

void require_same_algebra(const Batch &a, const Batch &b) {  // src/batch.rs:354-359
    if (!a.algebra->same_signature(*b.algebra))
        throw py::value_error("algebra mismatch: left is " + a.algebra->str() + " but right is " + b.algebra->str());
    if (a.view.dtype != b.view.dtype) throw py::type_error("batches must share one dtype (float64 or float32)");
}
int64_t broadcast_n(const Batch &a, const Batch &b) {  // src/batch.rs:361-376
    if (a.n() == b.n() || b.n() == 1) return a.n();
    if (a.n() == 1) return b.n();
    std::ostringstream s; s << "batch sizes must match or be 1 for broadcasting: " << a.n() << " vs " << b.n();
    throw py::value_error(s.str());
}

BatchPtr batch_product(const Batch &a, const Batch &b, int op) {
    require_same_algebra(a, b);
    auto out = a.empty_like(broadcast_n(a, b));
    py::gil_scoped_release nogil;
    // broadcast branch (src/batch.rs:365-376) in a 128 / 256-blade algebra with the one-row operand known on the host:
    // lcc_batch_fixed_product can run it block-diagonalised (K8) instead of as the dense GEMM
    if (a.nb() >= 128 && (a.n() == 1) != (b.n() == 1)) {
        const Batch &fixed = a.n() == 1 ? a : b, &batch = a.n() == 1 ? b : a;
        if (!fixed.host_row.empty() && batch.n() > 1) {
            check(lcc_batch_fixed_product(a.algebra->h, op, a.n() == 1 ? 0 : 1, fixed.host_row.data(), &batch.view, &out->view, fp_flags(), nullptr));
            return out;
        }
    }
    check(lcc_batch_product(a.algebra->h, op, &a.view, &b.view, &out->view, fp_flags(), nullptr));
    return out;
}
BatchPtr batch_unary(const Batch &a, int op, int arg = 0) {
    auto out = a.empty_like(a.n());
    py::gil_scoped_release nogil;
    check(lcc_batch_unary(a.algebra->h, op, arg, &a.view, &out->view, nullptr));
    return out;
}
BatchPtr batch_scale(const Batch &a, double f) {  // src/batch.rs:957-962
    auto out = a.empty_like(a.n());
    py::gil_scoped_release nogil;
    check(lcc_batch_scale(a.algebra->h, &a.view, f, &out->view, nullptr));
    return out;
}
BatchPtr batch_addsub(const Batch &a, const Batch &b, int op) {
    require_same_algebra(a, b);
    auto out = a.empty_like(broadcast_n(a, b));
    py::gil_scoped_release nogil;
    check(lcc_batch_elementwise(a.algebra->h, op, &a.view, &b.view, &out->view, nullptr));
    return out;
}
py::array batch_norm(const Batch &a, int kind) {
    py::array out(np_dtype(a.view.dtype), std::vector<py::ssize_t>{(py::ssize_t)a.n()});
    void *out_ptr = out.mutable_data();
    if (a.n() == 0) return out;
    py::gil_scoped_release nogil;
    DeviceTemp stage((size_t)a.n() * a.es());
    check(lcc_batch_norm(a.algebra->h, kind, &a.view, stage.ptr, fp_flags(), nullptr));
    check(lcc_memcpy_d2h(out_ptr, stage.ptr, (size_t)a.n() * a.es(), nullptr));
    check(lcc_stream_sync(nullptr));
    return out;
}
BatchPtr batch_sandwich(const Batch &a, const Multivector &rotor) {  // src/batch.rs:434-503
    AlgebraPtr ralg = rotor.get_algebra();
    if (!a.algebra->same_signature(*ralg))
        throw py::value_error("algebra mismatch: batch is " + a.algebra->str() + " but rotor is " + ralg->str());
    auto out = a.empty_like(a.n());
    py::gil_scoped_release nogil;
    check(lcc_batch_sandwich(a.algebra->h, rotor.coeffs.data(), &a.view, &out->view, fp_flags(), nullptr));
    return out;
}
Multivector batch_get(const Batch &b, int64_t index) {  // src/batch.rs:288-303
    if (index < 0 || index >= b.n()) {
        std::ostringstream s; s << "index " << index << " out of bounds for batch of size " << b.n();
        throw py::index_error(s.str());
    }
    const int nb = b.nb();
    std::vector<double> c(nb);
    {
        py::gil_scoped_release nogil;
        DeviceTemp stage((size_t)nb * b.es());
        lcc_view row{stage.ptr, 1, nb, b.view.dtype, LCC_AOS};
        check(lcc_batch_copy_rows(b.algebra->h, &b.view, index, index + 1, &row, 0, nullptr));
        if (b.view.dtype == LCC_F64) {
            check(lcc_memcpy_d2h(c.data(), stage.ptr, (size_t)nb * 8, nullptr));
            check(lcc_stream_sync(nullptr));
        } else {
            std::vector<float> f(nb);
            check(lcc_memcpy_d2h(f.data(), stage.ptr, (size_t)nb * 4, nullptr));
            check(lcc_stream_sync(nullptr));
            for (int i = 0; i < nb; ++i) c[i] = f[i];
        }
    }
    return Multivector{std::move(c), b.algebra->dim(), b.algebra};
}
void batch_set(Batch &b, int64_t index, const Multivector &mv) {  // src/batch.rs:314-335
    if (index < 0 || index >= b.n()) {
        std::ostringstream s; s << "index " << index << " out of bounds for batch of size " << b.n();
        throw py::index_error(s.str());
    }
    const int nb = b.nb();
    if ((int)mv.coeffs.size() != nb) {
        std::ostringstream s; s << "multivector has " << mv.coeffs.size() << " blades but batch expects " << nb;
        throw py::value_error(s.str());
    }
    py::gil_scoped_release nogil;
    DeviceTemp stage((size_t)nb * b.es());
    std::vector<float> f;
    const void *src = mv.coeffs.data();
    if (b.view.dtype == LCC_F32) { f.assign(mv.coeffs.begin(), mv.coeffs.end()); src = f.data(); }
    check(lcc_memcpy_h2d(stage.ptr, src, (size_t)nb * b.es(), nullptr));
    lcc_view row{stage.ptr, 1, nb, b.view.dtype, LCC_AOS};
    check(lcc_batch_copy_rows(b.algebra->h, &row, 0, 1, &b.view, index, nullptr));
    check(lcc_stream_sync(nullptr));
    if (b.n() == 1) {
        b.host_row.resize(nb);
        for (int k = 0; k < nb; ++k) b.host_row[k] = b.view.dtype == LCC_F64 ? mv.coeffs[k] : (double)(float)mv.coeffs[k];
    }
}
BatchPtr batch_slice(const Batch &b, int64_t start, int64_t end) {  // src/batch.rs:1083-1097
    if (start < 0 || end < 0 || start > b.n() || end > b.n() || start > end) {
        std::ostringstream s; s << "invalid slice [" << start << ":" << end << "] for batch of size " << b.n();
        throw py::index_error(s.str());
    }
    auto out = b.empty_like(end - start);
    py::gil_scoped_release nogil;
    check(lcc_batch_copy_rows(b.algebra->h, &b.view, start, end, &out->view, 0, nullptr));
    return out;
}
BatchPtr batch_concat(const Batch &a, const Batch &b) {  // src/batch.rs:1055-1073
    require_same_algebra(a, b);
    auto out = a.empty_like(a.n() + b.n());
    py::gil_scoped_release nogil;
    check(lcc_batch_copy_rows(a.algebra->h, &a.view, 0, a.n(), &out->view, 0, nullptr));
    check(lcc_batch_copy_rows(a.algebra->h, &b.view, 0, b.n(), &out->view, a.n(), nullptr));
    return out;
}
py::array batch_sum(const Batch &a, bool mean) {
    const int nb = a.nb();
    py::array out(np_dtype(a.view.dtype), std::vector<py::ssize_t>{(py::ssize_t)nb});
    void *out_ptr = out.mutable_data();
    {
        py::gil_scoped_release nogil;
        DeviceTemp stage((size_t)nb * a.es());
        check(lcc_batch_sum(a.algebra->h, &a.view, stage.ptr, nullptr));
        check(lcc_memcpy_d2h(out_ptr, stage.ptr, (size_t)nb * a.es(), nullptr));
        check(lcc_stream_sync(nullptr));
    }
    if (mean && a.n() > 0) {
        if (a.view.dtype == LCC_F64) { double *p = (double *)out.mutable_data(); for (int k = 0; k < nb; ++k) p[k] /= (double)a.n(); }
        else { float *p = (float *)out.mutable_data(); for (int k = 0; k < nb; ++k) p[k] = (float)((double)p[k] / (double)a.n()); }
    }
    return out;
}

std::vector<int32_t> vector_blades(int dim) { std::vector<int32_t> b; for (int i = 0; i < dim; ++i) b.push_back(1 << i); return b; }
std::vector<int32_t> blades_of_grade(int nb, int g) { std::vector<int32_t> b; for (int i = 0; i < nb; ++i) if (popcount((unsigned)i) == g) b.push_back(i); return b; }

std::string repr_float(double v) { return py::cast<std::string>(py::repr(py::float_(v))); }

}  // namespace

#include "multivector_methods.inc"

PYBIND11_MODULE(largecrimsoncanine, m) {
    m.doc() = "B200-native drop-in for the batched path of LargeCrimsonCanine (Algebra, Multivector, MultivectorBatch)";
    m.attr("__version__") = lcc_version();
    m.def("set_strict_fp", [](bool on) { g_strict_fp = on; },
          "Reproduce the reference's rounding sequence bit for bit (separate multiply and add) instead of FMA.");
    m.def("get_strict_fp", [] { return g_strict_fp; });
    m.def("kernel_launch_count", [] { return lcc_kernel_launch_count(); });
    m.def("pga_transform_points", [](const PyAlgebra &a, const Multivector &motor, const py::array &xyz_in) {
        // pga_point_coords(M * pga_point(x,y,z) * ~M) for an (N,3) array, fused on the device (include/lcc_b200.h)
        require_pga3d(a, "pga_transform_points");
        if (motor.coeffs.size() != 16) throw py::value_error("motor must be a PGA3D multivector (16 blades)");
        py::array xyz = contiguous(xyz_in);
        const int dtype = dtype_of(xyz);
        if (xyz.ndim() != 2 || xyz.shape(1) != 3) throw py::value_error("points must have 3 columns");
        py::array out = output_array(dtype, std::vector<py::ssize_t>{xyz.shape(0), 3});
        const size_t bytes = (size_t)xyz.nbytes();
        if (bytes == 0) return out;
        const void *host = xyz.data();
        void *out_ptr = out.mutable_data();
        const int64_t n = xyz.shape(0);
        const std::vector<double> m = motor.coeffs;
        const unsigned flags = fp_flags();
        py::gil_scoped_release nogil;
        // chunked H2D / kernel / D2H pipeline on host arrays (include/lcc_b200.h): 24 bytes per f32 point cross PCIe
        check(lcc_host_pga_points_sandwich(a.inner->h, dtype, m.data(), host, out_ptr, n, flags));
        return out;
    }, "Apply a PGA3D motor to an (N,3) array of points: embed, sandwich and extract fused in one kernel.");
    m.def("pinned_empty", [](std::vector<py::ssize_t> shape, const py::object &dtype) {
        const py::dtype dt = py::dtype::from_args(dtype);
        int code;
        if (dt.is(py::dtype::of<double>())) code = LCC_F64;
        else if (dt.is(py::dtype::of<float>())) code = LCC_F32;
        else throw py::type_error("pinned_empty supports float64 and float32");
        return pinned_array(code, shape);
    }, py::arg("shape"), py::arg("dtype") = py::dtype::of<double>(),
          "Uninitialised NumPy array in page-locked host memory (from the engine's pinned cache): from_numpy / sandwich_numpy "
          "DMA such arrays directly instead of staging them.");
    m.def("is_pinned", [](const py::array &a) { return lcc_host_is_pinned(a.data(), (size_t)a.nbytes()) != 0; });
    m.def("host_cache_trim", [] { check(lcc_host_cache_trim()); });
    m.def("host_cache_stats", [] { size_t live = 0, idle = 0; check(lcc_host_cache_stats(&live, &idle)); return py::make_tuple(live, idle); });
    m.def("set_option", [](const std::string &name, int64_t value) { check(lcc_set_option(name.c_str(), value)); },
          "Tuning knobs of the engine (include/lcc_b200.h, lcc_set_option); results never depend on them.");
    m.def("get_option", [](const std::string &name) { int64_t v = 0; check(lcc_get_option(name.c_str(), &v)); return v; });
    m.def("device_count", [] { int c = 0; check(lcc_device_count(&c)); return c; });
    m.def("set_device", [](int d) { check(lcc_set_device(d)); });
    m.def("synchronize", [] { check(lcc_stream_sync(nullptr)); });

    // ------------------------------------------------------------------ Algebra (src/pyalgebra.rs:30-200)
    py::class_<PyAlgebra>(m, "Algebra")
        .def(py::init([](long p, long q, long r) { return make_algebra(p, q, r); }), py::arg("p"), py::arg("q") = 0, py::arg("r") = 0)
        .def_static("euclidean", [](long n) { return make_algebra(n, 0, 0); })
        .def_static("pga", [](long n) { return make_algebra(n, 0, 1); })
        .def_static("cga", [](long n) { return make_algebra(n + 1, 1, 0); })
        .def_static("sta", [] { return make_algebra(1, 3, 0); })
        .def_property_readonly("dimension", [](const PyAlgebra &a) { return a.inner->dim(); })
        .def_property_readonly("num_blades", [](const PyAlgebra &a) { return a.inner->nb(); })
        .def_property_readonly("p", [](const PyAlgebra &a) { return a.inner->p; })
        .def_property_readonly("q", [](const PyAlgebra &a) { return a.inner->q; })
        .def_property_readonly("r", [](const PyAlgebra &a) { return a.inner->r; })
        .def("is_euclidean", [](const PyAlgebra &a) { return a.inner->is_euclidean(); })
        .def("is_pga", [](const PyAlgebra &a) { return a.inner->is_pga(); })
        .def("blade_name", [](const PyAlgebra &a, long index) {
            if (index < 0 || index >= a.inner->nb()) throw py::index_error("blade index out of range");
            char buf[32];
            lcc_algebra_blade_name(a.inner->h, (int)index, buf, sizeof buf);
            return std::string(buf);
        })
        .def("__repr__", [](const PyAlgebra &a) { return a.inner->str(); })
        .def("__str__", [](const PyAlgebra &a) { return a.inner->str(); })
        .def("__eq__", [](const PyAlgebra &a, const PyAlgebra &b) { return a.inner->same_signature(*b.inner); })
        .def("__hash__", [](const PyAlgebra &a) { return py::hash(py::make_tuple(a.inner->p, a.inner->q, a.inner->r)); })
        // additive surface: lets table parity be tested from Python and makes lcc.torch interop work
        .def_property_readonly("signature", [](const PyAlgebra &a) { return py::make_tuple(a.inner->p, a.inner->q, a.inner->r); })
        .def("cayley_signs", [](const PyAlgebra &a) {
            const int nb = a.inner->nb();
            py::array_t<int8_t> out({nb, nb});
            std::memcpy(out.mutable_data(), lcc_algebra_cayley_signs(a.inner->h), (size_t)nb * nb);
            return out;
        })
        .def("cayley_blades", [](const PyAlgebra &a) {
            const int nb = a.inner->nb();
            py::array_t<uint16_t> out({nb, nb});
            std::memcpy(out.mutable_data(), lcc_algebra_cayley_blades(a.inner->h), (size_t)nb * nb * 2);
            return out;
        })
        .def_property_readonly("grade_masks", [](const PyAlgebra &a) {
            py::list out;
            const int words = (a.inner->nb() + 63) / 64;
            std::vector<uint64_t> w(words);
            for (int g = 0; g <= a.inner->dim(); ++g) {
                lcc_algebra_grade_mask(a.inner->h, g, w.data(), words);
                py::object v = py::int_(0);
                for (int i = words - 1; i >= 0; --i) {
                    py::object shifted = py::reinterpret_steal<py::object>(PyNumber_Lshift(v.ptr(), py::int_(64).ptr()));
                    py::object word = py::reinterpret_steal<py::object>(PyLong_FromUnsignedLongLong(w[i]));
                    v = py::reinterpret_steal<py::object>(PyNumber_Or(shifted.ptr(), word.ptr()));
                }
                out.append(v);
            }
            return out;
        })
        .def_property_readonly("is_specialised", [](const PyAlgebra &a) { return lcc_algebra_is_specialised(a.inner->h) != 0; });

    // ------------------------------------------------------------------ Multivector
    py::class_<Multivector> mv(m, "Multivector");
    mv.def(py::init([](std::vector<double> c) { return Multivector::from_coeffs(std::move(c)); }))
        .def_static("from_list", [](std::vector<double> c) {  // src/multivector.rs:351-369 (its own messages)
            const size_t len = c.size();
            if (len == 0) throw py::value_error("coefficient list must not be empty");
            if (len & (len - 1)) {
                std::ostringstream s; s << "coefficient list length " << len << " is not a power of 2; expected 2^dims (e.g., 2, 4, 8, 16, ...)";
                throw py::value_error(s.str());
            }
            int dims = 0; while ((size_t(1) << dims) < len) ++dims;
            return Multivector{std::move(c), dims, nullptr};
        })
        .def_static("zero", [](int dims) {
            if (dims <= 0) throw py::value_error("dimension must be at least 1");
            return Multivector{std::vector<double>(size_t(1) << dims, 0.0), dims, nullptr};
        })
        .def_static("from_scalar", [](double value, int dims) {  // src/multivector.rs:318-329
            if (dims <= 0) throw py::value_error("dimension must be at least 1");
            std::vector<double> c(size_t(1) << dims, 0.0);
            c[0] = value;
            return Multivector{std::move(c), dims, nullptr};
        }, py::arg("value"), py::arg("dims"))
        .def_static("from_vector", [](std::vector<double> coords) {  // src/multivector.rs:380-395
            const int dims = (int)coords.size();
            if (dims == 0) throw py::value_error("coords must not be empty; provide at least one coordinate");
            std::vector<double> c(size_t(1) << dims, 0.0);
            for (int i = 0; i < dims; ++i) c[size_t(1) << i] = coords[i];
            return Multivector{std::move(c), dims, nullptr};
        })
        .def_static("basis", [](int index, int dims) {
            if (dims <= 0) throw py::value_error("dimension must be at least 1");
            if (index <= 0 || index > dims) throw py::value_error("basis index out of range");
            std::vector<double> c(size_t(1) << dims, 0.0);
            c[size_t(1) << (index - 1)] = 1.0;
            return Multivector{std::move(c), dims, nullptr};
        }, py::arg("index"), py::arg("dims"))
        .def_static("from_bivector", [](std::vector<double> comps, int dims) {  // src/multivector.rs:1228-1258
            if (dims < 2) throw py::value_error("dimension must be at least 2 for bivectors");
            const size_t expected = (size_t)dims * (dims - 1) / 2;
            if (comps.size() != expected) {
                std::ostringstream s; s << "expected " << expected << " bivector components for Cl(" << dims << ") (got " << comps.size() << "); components are [e12, e13, e23, ...] in lexicographic order";
                throw py::value_error(s.str());
            }
            std::vector<double> c(size_t(1) << dims, 0.0);
            size_t next = 0;
            for (int i = 0; i < dims; ++i)            // e_i ^ e_j, i < j, pairs in lexicographic order:
                for (int j = i + 1; j < dims; ++j)    // e12, e13, e14, e23, ... (NOT ascending blade index from 4-D on)
                    c[(size_t(1) << i) | (size_t(1) << j)] = comps[next++];
            return Multivector{std::move(c), dims, nullptr};
        }, py::arg("components"), py::arg("dims"))
        .def_static("from_list_in", [](std::vector<double> c, const PyAlgebra &a) {  // additive: coefficients + explicit algebra
            if ((int)c.size() != a.inner->nb()) {
                std::ostringstream s; s << "coefficient list has " << c.size() << " entries but " << a.inner->str() << " has " << a.inner->nb() << " blades";
                throw py::value_error(s.str());
            }
            return mv_in(a, std::move(c));
        }, py::arg("coeffs"), py::arg("algebra"))
        .def_static("zero_in", [](const PyAlgebra &a) { return mv_in(a, std::vector<double>(a.inner->nb(), 0.0)); })
        .def_static("scalar_in", [](double v, const PyAlgebra &a) { std::vector<double> c(a.inner->nb(), 0.0); c[0] = v; return mv_in(a, std::move(c)); })
        .def_static("vector_in", [](std::vector<double> coords, const PyAlgebra &a) {  // src/multivector.rs:498-517
            if ((int)coords.size() != a.inner->dim()) {
                std::ostringstream s; s << "coords length " << coords.size() << " doesn't match algebra dimension " << a.inner->dim();
                throw py::value_error(s.str());
            }
            std::vector<double> c(a.inner->nb(), 0.0);
            for (size_t i = 0; i < coords.size(); ++i) c[size_t(1) << i] = coords[i];
            return mv_in(a, std::move(c));
        })
        .def_static("basis_in", [](int index, const PyAlgebra &a) {  // src/multivector.rs:530-547
            const int dim = a.inner->dim();
            if (index <= 0 || index > dim) {
                std::ostringstream s; s << "basis index " << index << " out of range for " << dim << "-dimensional algebra (use 1 to " << dim << ")";
                throw py::value_error(s.str());
            }
            std::vector<double> c(a.inner->nb(), 0.0);
            c[size_t(1) << (index - 1)] = 1.0;
            return mv_in(a, std::move(c));
        })
        .def_static("pseudoscalar_in", [](const PyAlgebra &a) { std::vector<double> c(a.inner->nb(), 0.0); c.back() = 1.0; return mv_in(a, std::move(c)); })
        .def_static("from_axis_angle", [](const Multivector &axis, double angle) {  // src/multivector.rs:955-999
            if (axis.dims != 3) { std::ostringstream s; s << "from_axis_angle requires 3D vectors, got Cl(" << axis.dims << ")"; throw py::value_error(s.str()); }
            auto g = axis.grades();
            if (g.size() != 1 || g[0] != 1) throw py::value_error("axis must be a vector (grade 1)");
            const double n = axis.norm();
            if (n == 0.0) throw py::value_error("axis has zero norm; cannot create rotor");
            const double inv = 1.0 / n;
            const double ax = axis.coeffs[1] * inv, ay = axis.coeffs[2] * inv, az = axis.coeffs[4] * inv;
            const double half = angle / 2.0, c = std::cos(half), s = std::sin(half);
            std::vector<double> r(8, 0.0);
            r[0] = c; r[3] = -s * az; r[5] = s * ay; r[6] = -s * ax;
            return Multivector{std::move(r), 3, nullptr};
        })
        .def("algebra", [](const Multivector &a) -> py::object { if (!a.algebra_opt) return py::none(); return py::cast(PyAlgebra{a.algebra_opt}); })
        .def_property_readonly("dims", [](const Multivector &a) { return a.dims; })
        .def_property_readonly("dimension", [](const Multivector &a) { return a.dims; })  // #[getter], src/multivector.rs:1994-1997
        .def("scalar", [](const Multivector &a) { return a.coeffs[0]; })
        .def("coefficients", [](const Multivector &a) { return a.coeffs; })
        .def("to_list", [](const Multivector &a) { return a.coeffs; })
        .def("to_vector_coords", [](const Multivector &a) { std::vector<double> v; for (int i = 0; i < a.dims; ++i) v.push_back(a.coeffs[size_t(1) << i]); return v; })
        .def("to_bivector_coords", [](const Multivector &a) { std::vector<double> v; for (size_t i = 0; i < a.coeffs.size(); ++i) if (popcount((unsigned)i) == 2) v.push_back(a.coeffs[i]); return v; })
        .def("blades", [](const Multivector &a) {
            std::vector<std::tuple<size_t, double, int>> out;
            for (size_t i = 0; i < a.coeffs.size(); ++i) if (a.coeffs[i] != 0.0) out.emplace_back(i, a.coeffs[i], popcount((unsigned)i));
            return out;
        })
        .def("grade", &Multivector::grade)
        .def("vector_part", [](const Multivector &a) { return a.grade(1); })
        .def("bivector_part", [](const Multivector &a) { return a.dims >= 2 ? a.grade(2) : a; })
        .def("grades", &Multivector::grades)
        .def("max_grade", [](const Multivector &a) -> py::object { auto g = a.grades(); if (g.empty()) return py::none(); return py::int_(g.back()); })
        .def("pure_grade", [](const Multivector &a) -> py::object { auto g = a.grades(); if (g.size() != 1) return py::none(); return py::int_(g[0]); })
        .def("is_vector", [](const Multivector &a) { auto g = a.grades(); return g.size() == 1 && g[0] == 1; })
        .def("is_scalar", [](const Multivector &a) { auto g = a.grades(); return g.size() == 1 && g[0] == 0; })
        .def("geometric_product", [](const Multivector &a, const Multivector &b) { return a.product(LCC_OP_GP, b); })
        .def("gp", [](const Multivector &a, const Multivector &b) { return a.product(LCC_OP_GP, b); })
        .def("outer_product", [](const Multivector &a, const Multivector &b) { return a.product(LCC_OP_OUTER, b); })
        .def("wedge", [](const Multivector &a, const Multivector &b) { return a.product(LCC_OP_OUTER, b); })
        .def("left_contraction", [](const Multivector &a, const Multivector &b) { return a.product(LCC_OP_LC, b); })
        .def("inner", [](const Multivector &a, const Multivector &b) { return a.product(LCC_OP_LC, b); })
        .def("lc", [](const Multivector &a, const Multivector &b) { return a.product(LCC_OP_LC, b); })
        // the rest of the product family, src/multivector.rs:1616-1797
        .def("right_contraction", [](const Multivector &a, const Multivector &b) { return a.product(LCC_OP_RC, b); })
        .def("rc", [](const Multivector &a, const Multivector &b) { return a.product(LCC_OP_RC, b); })
        .def("scalar_product", [](const Multivector &a, const Multivector &b) { return a.product(LCC_OP_SCALAR, b); })
        .def("dot", [](const Multivector &a, const Multivector &b) { return a.product(LCC_OP_SCALAR, b); })
        .def("commutator", [](const Multivector &a, const Multivector &b) { return a.product_same_dims(LCC_OP_COMMUTATOR, b); })
        .def("x", [](const Multivector &a, const Multivector &b) { return a.product_same_dims(LCC_OP_COMMUTATOR, b); })
        .def("anticommutator", [](const Multivector &a, const Multivector &b) { return a.product_same_dims(LCC_OP_ANTICOMMUTATOR, b); })
        .def("regressive", [](const Multivector &a, const Multivector &b) { return a.product_same_dims(LCC_OP_REGRESSIVE, b); })
        .def("meet", [](const Multivector &a, const Multivector &b) { return a.product_same_dims(LCC_OP_REGRESSIVE, b); })
        .def("vee", [](const Multivector &a, const Multivector &b) { return a.product_same_dims(LCC_OP_REGRESSIVE, b); })
        .def("reverse", &Multivector::reverse)
        .def("tilde", &Multivector::reverse)
        .def("__invert__", &Multivector::reverse)
        .def("norm_squared", &Multivector::norm_squared)
        .def("norm", &Multivector::norm)
        .def("magnitude", &Multivector::norm)
        .def("scale", &Multivector::scale)
        .def("normalized", [](const Multivector &a) {  // src/multivector.rs:2957-2967
            const double n = a.norm();
            if (n == 0.0) throw py::value_error("cannot normalize zero multivector; normalization requires non-zero norm (check that your multivector has at least one non-zero coefficient)");
            return a.scale(1.0 / n);
        })
        .def("normalize", [](const Multivector &a) -> py::object { const double n = a.norm(); if (n == 0.0) return py::none(); return py::cast(a.scale(1.0 / n)); })
        .def("inverse", &Multivector::inverse)
        .def("sandwich", &Multivector::sandwich)
        .def("apply", &Multivector::sandwich)
        .def("dual", [](const Multivector &a) { return a.product(LCC_OP_GP, a.pseudoscalar_like().inverse()); })      // src/multivector.rs:4249-4255
        .def("undual", [](const Multivector &a) { return a.product(LCC_OP_GP, a.pseudoscalar_like()); })               // :4270-4274
        .def("right_dual", [](const Multivector &a) { return a.pseudoscalar_like().inverse().product(LCC_OP_GP, a); })  // :4285-4290
        .def("approx_eq", [](const Multivector &a, const Multivector &b, double tol) {
            if (a.coeffs.size() != b.coeffs.size()) return false;
            for (size_t i = 0; i < a.coeffs.size(); ++i) if (std::fabs(a.coeffs[i] - b.coeffs[i]) > tol) return false;
            return true;
        }, py::arg("other"), py::arg("tol") = 1e-10)
        .def("__mul__", [](const Multivector &a, const py::object &o) -> Multivector {  // src/multivector.rs:2030-2041
            if (py::isinstance<Multivector>(o)) return a.product(LCC_OP_GP, o.cast<const Multivector &>());
            if (py::isinstance<py::float_>(o) || py::isinstance<py::int_>(o)) return a.scale(o.cast<double>());
            throw py::type_error("unsupported operand type for *: expected Multivector or float");
        })
        .def("__rmul__", [](const Multivector &a, const py::object &o) -> Multivector {
            if (py::isinstance<py::float_>(o) || py::isinstance<py::int_>(o)) return a.scale(o.cast<double>());
            throw py::type_error("unsupported operand type for *: expected float");
        })
        .def("__truediv__", [](const Multivector &a, const py::object &o) -> Multivector {  // src/multivector.rs:2054-2070
            if (py::isinstance<Multivector>(o)) return a.product(LCC_OP_GP, o.cast<const Multivector &>().inverse());
            if (py::isinstance<py::float_>(o) || py::isinstance<py::int_>(o)) {
                const double s = o.cast<double>();
                if (s == 0.0) { PyErr_SetString(PyExc_ZeroDivisionError, "cannot divide multivector by zero"); throw py::error_already_set(); }
                return a.scale(1.0 / s);
            }
            throw py::type_error("unsupported operand type for /: expected Multivector or float");
        })
        .def("__xor__", [](const Multivector &a, const Multivector &b) { return a.product(LCC_OP_OUTER, b); })
        .def("__or__", [](const Multivector &a, const Multivector &b) { return a.product(LCC_OP_LC, b); })
        .def("__add__", [](const Multivector &a, const Multivector &b) {  // src/multivector.rs:2084-2103 (drops the algebra)
            if (a.dims != b.dims) { std::ostringstream s; s << "dimension mismatch: left operand is Cl(" << a.dims << ") but right operand is Cl(" << b.dims << "); both multivectors must have the same dimension"; throw py::value_error(s.str()); }
            std::vector<double> c(a.coeffs.size());
            for (size_t i = 0; i < c.size(); ++i) c[i] = a.coeffs[i] + b.coeffs[i];
            return Multivector{std::move(c), a.dims, nullptr};
        })
        .def("__sub__", [](const Multivector &a, const Multivector &b) {
            if (a.dims != b.dims) { std::ostringstream s; s << "dimension mismatch: left operand is Cl(" << a.dims << ") but right operand is Cl(" << b.dims << "); both multivectors must have the same dimension"; throw py::value_error(s.str()); }
            std::vector<double> c(a.coeffs.size());
            for (size_t i = 0; i < c.size(); ++i) c[i] = a.coeffs[i] - b.coeffs[i];
            return Multivector{std::move(c), a.dims, nullptr};
        })
        .def("__neg__", [](const Multivector &a) { std::vector<double> c(a.coeffs); for (auto &v : c) v = -v; return a.like(std::move(c)); })
        .def("__pos__", [](const Multivector &a) { return a; })
        .def("__eq__", [](const Multivector &a, const py::object &o) {
            if (!py::isinstance<Multivector>(o)) return false;
            const Multivector &b = o.cast<const Multivector &>();
            return a.dims == b.dims && a.coeffs == b.coeffs;
        })
        .def("__hash__", [](const Multivector &a) { return py::hash(py::make_tuple(a.dims, py::tuple(py::cast(a.coeffs)))); })
        .def("__len__", [](const Multivector &a) { return a.coeffs.size(); })
        .def("__getitem__", [](const Multivector &a, long i) {
            const long n = (long)a.coeffs.size();
            const long idx = i < 0 ? n + i : i;
            if (idx < 0 || idx >= n) throw py::index_error("coefficient index out of range");
            return a.coeffs[idx];
        })
        .def("__repr__", [](const Multivector &a) {  // src/multivector.rs:1819-1821
            std::ostringstream s;
            s << "Multivector([";
            for (size_t i = 0; i < a.coeffs.size(); ++i) s << (i ? ", " : "") << repr_float(a.coeffs[i]);
            s << "], dims=" << a.dims << ")";
            return s.str();
        });

    // ---- PGA3D (src/pga.rs)
    mv.def_static("pga_point", [](const PyAlgebra &a, double x, double y, double z) {  // :109-141
          require_pga3d(a, "pga_point");
          std::vector<double> c(16, 0.0);
          c[7] = 1.0; c[14] = -x; c[13] = y; c[11] = -z;
          return mv_in(a, std::move(c));
      })
        .def_static("pga_plane", [](const PyAlgebra &a, double pa, double pb, double pc, double pd) {  // :165-186
            require_pga3d(a, "pga_plane");
            std::vector<double> c(16, 0.0);
            c[1] = pa; c[2] = pb; c[4] = pc; c[8] = pd;
            return mv_in(a, std::move(c));
        })
        .def_static("pga_translator", [](const PyAlgebra &a, double dx, double dy, double dz) {  // :371-396
            require_pga3d(a, "pga_translator");
            std::vector<double> c(16, 0.0);
            c[0] = 1.0; c[9] = dx / 2.0; c[10] = dy / 2.0; c[12] = dz / 2.0;
            return mv_in(a, std::move(c));
        })
        .def_static("pga_motor_from_axis_angle", [](const PyAlgebra &a, std::vector<double> axis, std::vector<double> point, double angle) {  // :263-345
            require_pga3d(a, "pga_motor_from_axis_angle");
            if (axis.size() != 3) throw py::value_error("axis must have 3 components");
            if (point.size() != 3) throw py::value_error("point must have 3 components");
            const double len = std::sqrt(axis[0] * axis[0] + axis[1] * axis[1] + axis[2] * axis[2]);
            if (len < 1e-10) throw py::value_error("axis has zero length");
            const double ax = axis[0] / len, ay = axis[1] / len, az = axis[2] / len;
            const double px = point[0], py_ = point[1], pz = point[2];
            const double half = angle / 2.0, c = std::cos(half), s = std::sin(half);
            std::vector<double> m(16, 0.0);
            m[0] = c; m[6] = -s * ax; m[5] = s * ay; m[3] = -s * az;
            const double mx = py_ * az - pz * ay, my = pz * ax - px * az, mz = px * ay - py_ * ax;
            m[9] = s * mx; m[10] = s * my; m[12] = s * mz;
            return mv_in(a, std::move(m));
        })
        .def("pga_point_coords", [](const Multivector &p) {  // :571-599
            require_pga3d_element(p, "pga_point_coords");
            const double w = p.coeffs[7];
            if (std::fabs(w) < 1e-10) throw py::value_error("cannot extract coordinates from ideal point (e123 = 0)");
            return py::make_tuple(-p.coeffs[14] / w, p.coeffs[13] / w, -p.coeffs[11] / w);
        })
        .def("pga_normalize", [](const Multivector &p) {  // :413-439
            require_pga3d_element(p, "pga_normalize");
            const double w = p.coeffs[7];
            if (std::fabs(w) < 1e-10) throw py::value_error("cannot normalize ideal point (e123 coefficient is zero)");
            std::vector<double> c(p.coeffs);
            for (auto &v : c) v = v / w;
            return p.like(std::move(c));
        })
        .def("pga_dual", [](const Multivector &p) {  // :521-556
            require_pga3d_element(p, "pga_dual");
            std::vector<double> c(16, 0.0);
            for (int i = 0; i < 16; ++i)
                if (p.coeffs[i] != 0.0) c[15 ^ i] += reverse_sign(popcount((unsigned)i)) * p.coeffs[i];
            return p.like(std::move(c));
        });

    mv.def_static("pga_line_from_points", [](const Multivector &p1, const Multivector &p2) {  // src/pga.rs:204-216: join = dual(dual(p1) ^ dual(p2))
          auto pga_dual = [](const Multivector &p) {
              require_pga3d_element(p, "pga_dual");
              std::vector<double> c(16, 0.0);
              for (int i = 0; i < 16; ++i)
                  if (p.coeffs[i] != 0.0) c[15 ^ i] += reverse_sign(popcount((unsigned)i)) * p.coeffs[i];
              return p.like(std::move(c));
          };
          return pga_dual(pga_dual(p1).product(LCC_OP_OUTER, pga_dual(p2)));
      })
        .def_static("pga_line_from_planes", [](const Multivector &a, const Multivector &b) { return a.product(LCC_OP_OUTER, b); })  // :234-240
        .def("pga_ideal_part", [](const Multivector &p) {  // :452-476: blades containing e0 (bit 3)
            require_pga3d_element(p, "pga_ideal_part");
            std::vector<double> c(16, 0.0);
            for (int i = 0; i < 16; ++i) if (i & 8) c[i] = p.coeffs[i];
            return p.like(std::move(c));
        })
        .def("pga_euclidean_part", [](const Multivector &p) {  // :489-512
            require_pga3d_element(p, "pga_euclidean_part");
            std::vector<double> c(16, 0.0);
            for (int i = 0; i < 16; ++i) if (!(i & 8)) c[i] = p.coeffs[i];
            return p.like(std::move(c));
        })
        .def("pga_point_plane_distance", [](const Multivector &p, const Multivector &plane) {  // :621-658
            AlgebraPtr alg = p.get_algebra();
            if (!alg->is_pga() || p.dims != 4) throw py::value_error("pga_point_plane_distance requires PGA3D elements");
            const double w = p.coeffs[7];
            if (std::fabs(w) < 1e-10) throw py::value_error("cannot compute distance from ideal point");
            const double a = plane.coeffs[1], b = plane.coeffs[2], c = plane.coeffs[4], d = plane.coeffs[8];
            const double len = std::sqrt(a * a + b * b + c * c);
            if (len < 1e-10) throw py::value_error("plane has zero normal");
            const double x = -p.coeffs[14] / w, y = p.coeffs[13] / w, z = -p.coeffs[11] / w;
            return (a * x + b * y + c * z + d) / len;
        })
        .def("is_pga_motor", [](const Multivector &m, double tol) {  // :670-703: even and M*~M = 1
            AlgebraPtr alg = m.get_algebra();
            if (!alg->is_pga() || m.dims != 4) return false;
            for (int i = 0; i < 16; ++i) if ((popcount((unsigned)i) & 1) && std::fabs(m.coeffs[i]) > tol) return false;
            Multivector prod = m.product(LCC_OP_GP, m.reverse());
            if (std::fabs(prod.coeffs[0] - 1.0) > tol) return false;
            for (int i = 1; i < 16; ++i) if (std::fabs(prod.coeffs[i]) > tol) return false;
            return true;
        }, py::arg("tol") = 1e-10);

    // ---- CGA (src/cga.rs): e+ = blade 2^n, e- = blade 2^(n+1) (:45-53)
    auto require_cga = [](const PyAlgebra &a, const char *msg) { if (a.inner->q != 1 || a.inner->r != 0) throw py::value_error(msg); };
    auto cga_algebra_of = [](const Multivector &m, const char *who) -> AlgebraPtr {
        if (!m.algebra_opt) throw py::value_error(std::string(who) + " requires a CGA multivector with explicit algebra");
        if (m.algebra_opt->q != 1 || m.algebra_opt->r != 0) throw py::value_error(std::string(who) + " requires a CGA multivector");
        return m.algebra_opt;
    };
    auto squared_norm_inner = [](const Multivector &m) { return m.product(LCC_OP_GP, m.reverse()).coeffs[0]; };  // :650-654
    mv.def_static("cga_eo", [require_cga](const PyAlgebra &a) {  // :79-104: (e- - e+)/2
          require_cga(a, "cga_eo requires a CGA algebra (q=1, r=0); use Algebra.cga(n)");
          const int n = a.inner->dim() - 2;
          std::vector<double> c(a.inner->nb(), 0.0);
          c[size_t(1) << n] = -0.5; c[size_t(1) << (n + 1)] = 0.5;
          return mv_in(a, std::move(c));
      })
        .def_static("cga_einf", [require_cga](const PyAlgebra &a) {  // :122-147: e- + e+
            require_cga(a, "cga_einf requires a CGA algebra (q=1, r=0); use Algebra.cga(n)");
            const int n = a.inner->dim() - 2;
            std::vector<double> c(a.inner->nb(), 0.0);
            c[size_t(1) << n] = 1.0; c[size_t(1) << (n + 1)] = 1.0;
            return mv_in(a, std::move(c));
        })
        .def_static("cga_point", [require_cga](const PyAlgebra &a, double x, double y, double z) {  // :171-212
            require_cga(a, "cga_point requires a CGA algebra; use Algebra.cga(3)");
            if (a.inner->dim() != 5) throw py::value_error("cga_point(x,y,z) requires CGA3D (dimension 5); use Algebra.cga(3)");
            std::vector<double> c(32, 0.0);
            c[1] = x; c[2] = y; c[4] = z;
            const double nsq = x * x + y * y + z * z;
            c[8] = 0.5 * nsq - 0.5; c[16] = 0.5 * nsq + 0.5;
            return mv_in(a, std::move(c));
        })
        .def_static("cga_point_from_coords", [require_cga](const PyAlgebra &a, std::vector<double> coords) {  // :230-272
            require_cga(a, "cga_point_from_coords requires a CGA algebra; use Algebra.cga(n)");
            const int n = a.inner->dim() - 2;
            if ((int)coords.size() != n) {
                std::ostringstream s; s << "coords length " << coords.size() << " doesn't match Euclidean dimension " << n << "; for CGA" << n << " use " << n << " coordinates";
                throw py::value_error(s.str());
            }
            std::vector<double> c(a.inner->nb(), 0.0);
            double nsq = 0.0;
            for (int i = 0; i < n; ++i) { c[size_t(1) << i] = coords[i]; nsq += coords[i] * coords[i]; }
            c[size_t(1) << n] = 0.5 * nsq - 0.5; c[size_t(1) << (n + 1)] = 0.5 * nsq + 0.5;
            return mv_in(a, std::move(c));
        })
        .def_static("cga_sphere", [require_cga](const PyAlgebra &a, double cx, double cy, double cz, double r) {  // :294-309
            require_cga(a, "cga_point requires a CGA algebra; use Algebra.cga(3)");
            if (a.inner->dim() != 5) throw py::value_error("cga_point(x,y,z) requires CGA3D (dimension 5); use Algebra.cga(3)");
            std::vector<double> c(32, 0.0);
            c[1] = cx; c[2] = cy; c[4] = cz;
            const double nsq = cx * cx + cy * cy + cz * cz;
            c[8] = 0.5 * nsq - 0.5; c[16] = 0.5 * nsq + 0.5;
            const double half = 0.5 * r * r;
            c[8] -= half; c[16] -= half;
            return mv_in(a, std::move(c));
        })
        .def_static("cga_plane", [require_cga](const PyAlgebra &a, double pa, double pb, double pc, double pd) {  // :332-366: n + d*einf
            require_cga(a, "cga_plane requires a CGA algebra; use Algebra.cga(3)");
            if (a.inner->dim() != 5) throw py::value_error("cga_plane requires CGA3D (dimension 5); use Algebra.cga(3)");
            std::vector<double> c(32, 0.0);
            c[1] = pa; c[2] = pb; c[4] = pc; c[8] = pd; c[16] = pd;
            return mv_in(a, std::move(c));
        })
        .def_static("cga_circle", [](const Multivector &s1, const Multivector &s2) { return s1.product(LCC_OP_OUTER, s2); })       // :387-389
        .def_static("cga_point_pair", [](const Multivector &p1, const Multivector &p2) { return p1.product(LCC_OP_OUTER, p2); })  // :408-410
        .def_static("cga_line", [](const Multivector &a, const Multivector &b) { return a.product(LCC_OP_OUTER, b); })            // :428-430
        .def("cga_extract_center", [cga_algebra_of](const Multivector &m) {  // :451-511
            AlgebraPtr alg = cga_algebra_of(m, "cga_extract_center");
            if (alg->dim() != 5) throw py::value_error("cga_extract_center currently only supports CGA3D (dimension 5)");
            const double w = m.coeffs[16] - m.coeffs[8];
            if (std::fabs(w) < 1e-12) throw py::value_error("cannot extract center from a flat object (plane/line); use cga_is_flat() to check first");
            return py::make_tuple(m.coeffs[1] / w, m.coeffs[2] / w, m.coeffs[4] / w);
        })
        .def("cga_extract_radius", [cga_algebra_of, squared_norm_inner](const Multivector &m) {  // :527-574
            AlgebraPtr alg = cga_algebra_of(m, "cga_extract_radius");
            if (alg->dim() != 5) throw py::value_error("cga_extract_radius currently only supports CGA3D (dimension 5)");
            const double nsq = squared_norm_inner(m);
            const double w = m.coeffs[16] - m.coeffs[8], w2 = w * w;
            if (std::fabs(w2) < 1e-24) throw py::value_error("cannot extract radius from a flat object (plane/line)");
            const double r2 = nsq / w2;
            return r2 >= 0.0 ? std::sqrt(r2) : -std::sqrt(-r2);
        })
        .def("cga_is_round", [squared_norm_inner](const Multivector &m) { return squared_norm_inner(m) > 1e-10; })  // :591-594
        .def("cga_is_flat", [cga_algebra_of, squared_norm_inner](const Multivector &m) {  // :611-648
            AlgebraPtr alg = cga_algebra_of(m, "cga_is_flat");
            const int n = alg->dim() - 2;
            const double w = std::fabs(m.coeffs[size_t(1) << (n + 1)] - m.coeffs[size_t(1) << n]);
            return w < 1e-10 || squared_norm_inner(m) < -1e-10;
        })
        .def("cga_point_on_sphere", [](const Multivector &p, const Multivector &sphere) {  // :674-680
            Multivector in = p.product(LCC_OP_LC, sphere);
            for (double c : in.coeffs) if (std::fabs(c) >= 1e-10) return false;
            return true;
        })
        .def("cga_distance", [](const Multivector &p, const Multivector &q) {  // :701-720: sqrt(-2 (P . Q))
            const double d2 = -2.0 * p.product(LCC_OP_LC, q).coeffs[0];
            if (d2 < 0.0) {
                if (d2 > -1e-10) return 0.0;
                std::ostringstream s; s << "invalid distance calculation: d² = " << d2 << " (should be >= 0)";
                throw py::value_error(s.str());
            }
            return std::sqrt(d2);
        });

    // ---- STA Cl(1,3,0) (src/sta.rs): gamma_mu = blade 1<<mu (:34-51)
    auto require_sta = [](const AlgebraPtr &a) {
        if (!(a->p == 1 && a->q == 3 && a->r == 0)) {
            std::ostringstream s; s << "STA methods require Cl(1,3,0) algebra, got Cl(" << a->p << "," << a->q << "," << a->r << ")";
            throw py::value_error(s.str());
        }
    };
    mv.def_static("sta_vector", [require_sta](const PyAlgebra &a, double t, double x, double y, double z) {  // :102-114
          require_sta(a.inner);
          std::vector<double> c(16, 0.0);
          c[1] = t; c[2] = x; c[4] = y; c[8] = z;
          return mv_in(a, std::move(c));
      })
        .def_static("sta_gamma", [require_sta](const PyAlgebra &a, long i) {  // :139-156
            require_sta(a.inner);
            if (i < 0) throw py::type_error("gamma index must be a non-negative integer");
            if (i > 3) { std::ostringstream s; s << "gamma index must be 0-3, got " << i; throw py::value_error(s.str()); }
            std::vector<double> c(16, 0.0);
            c[size_t(1) << i] = 1.0;
            return mv_in(a, std::move(c));
        })
        .def_static("sta_bivector", [require_sta](const PyAlgebra &a, std::vector<double> e, std::vector<double> b) {  // :183-216
            require_sta(a.inner);
            if (e.size() != 3) { std::ostringstream s; s << "E field must have 3 components, got " << e.size(); throw py::value_error(s.str()); }
            if (b.size() != 3) { std::ostringstream s; s << "B field must have 3 components, got " << b.size(); throw py::value_error(s.str()); }
            std::vector<double> c(16, 0.0);
            c[3] = e[0]; c[5] = e[1]; c[9] = e[2];
            c[12] = b[0]; c[10] = -b[1]; c[6] = b[2];
            return mv_in(a, std::move(c));
        })
        .def_static("sta_boost", [require_sta](const PyAlgebra &a, std::vector<double> v) {  // :243-300
            require_sta(a.inner);
            if (v.size() != 3) { std::ostringstream s; s << "velocity must have 3 components, got " << v.size(); throw py::value_error(s.str()); }
            const double mag = std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
            std::vector<double> c(16, 0.0);
            if (mag < 1e-14) { c[0] = 1.0; return mv_in(a, std::move(c)); }
            if (mag >= 1.0) { std::ostringstream s; s << "velocity magnitude " << mag << " >= c (speed of light = 1 in natural units)"; throw py::value_error(s.str()); }
            const double phi = std::atanh(mag), nx = v[0] / mag, ny = v[1] / mag, nz = v[2] / mag;
            const double ch = std::cosh(phi / 2.0), sh = std::sinh(phi / 2.0);
            c[0] = ch; c[3] = -sh * nx; c[5] = -sh * ny; c[9] = -sh * nz;
            return mv_in(a, std::move(c));
        })
        .def_static("sta_rotation", [require_sta](const PyAlgebra &a, std::vector<double> axis, double angle) {  // :329-373
            require_sta(a.inner);
            if (axis.size() != 3) { std::ostringstream s; s << "axis must have 3 components, got " << axis.size(); throw py::value_error(s.str()); }
            const double mag = std::sqrt(axis[0] * axis[0] + axis[1] * axis[1] + axis[2] * axis[2]);
            if (mag < 1e-14) throw py::value_error("axis has zero magnitude");
            const double nx = axis[0] / mag, ny = axis[1] / mag, nz = axis[2] / mag;
            const double ch = std::cos(angle / 2.0), sh = std::sin(angle / 2.0);
            std::vector<double> c(16, 0.0);
            c[0] = ch; c[12] = sh * nx; c[10] = -sh * ny; c[6] = sh * nz;
            return mv_in(a, std::move(c));
        })
        .def_static("sta_I", [require_sta](const PyAlgebra &a) {  // :393-404
            require_sta(a.inner);
            std::vector<double> c(16, 0.0);
            c[15] = 1.0;
            return mv_in(a, std::move(c));
        })
        .def("sta_spacetime_split", [require_sta](const Multivector &m) {  // :428-437
            require_sta(m.get_algebra());
            return py::make_tuple(m.coeffs[1], std::vector<double>{m.coeffs[2], m.coeffs[4], m.coeffs[8]});
        })
        .def("sta_interval_squared", [require_sta](const Multivector &m) {  // :500-510
            require_sta(m.get_algebra());
            const double t = m.coeffs[1], x = m.coeffs[2], y = m.coeffs[4], z = m.coeffs[8];
            return t * t - x * x - y * y - z * z;
        })
        .def("sta_proper_time", [require_sta](const Multivector &m) {  // :458-474
            require_sta(m.get_algebra());
            const double t = m.coeffs[1], x = m.coeffs[2], y = m.coeffs[4], z = m.coeffs[8];
            const double i2 = t * t - x * x - y * y - z * z;
            return i2 >= 0.0 ? std::sqrt(i2) : -std::sqrt(-i2);
        })
        .def("sta_is_timelike", [require_sta](const Multivector &m, py::object tol) {  // :523-528
            require_sta(m.get_algebra());
            const double t = m.coeffs[1], x = m.coeffs[2], y = m.coeffs[4], z = m.coeffs[8];
            return (t * t - x * x - y * y - z * z) > (tol.is_none() ? 1e-10 : tol.cast<double>());
        }, py::arg("tol") = py::none())
        .def("sta_is_spacelike", [require_sta](const Multivector &m, py::object tol) {  // :541-546
            require_sta(m.get_algebra());
            const double t = m.coeffs[1], x = m.coeffs[2], y = m.coeffs[4], z = m.coeffs[8];
            return (t * t - x * x - y * y - z * z) < -(tol.is_none() ? 1e-10 : tol.cast<double>());
        }, py::arg("tol") = py::none())
        .def("sta_is_lightlike", [require_sta](const Multivector &m, py::object tol) {  // :563-568
            require_sta(m.get_algebra());
            const double t = m.coeffs[1], x = m.coeffs[2], y = m.coeffs[4], z = m.coeffs[8];
            return std::fabs(t * t - x * x - y * y - z * z) <= (tol.is_none() ? 1e-10 : tol.cast<double>());
        }, py::arg("tol") = py::none())
        .def("sta_field_decompose", [require_sta](const Multivector &m) {  // :584-596
            require_sta(m.get_algebra());
            return py::make_tuple(std::vector<double>{m.coeffs[3], m.coeffs[5], m.coeffs[9]},
                                  std::vector<double>{m.coeffs[12], -m.coeffs[10], m.coeffs[6]});
        })
        .def("sta_em_dual", [require_sta](const Multivector &m) {  // :613-623: I * F
            AlgebraPtr alg = m.get_algebra();
            require_sta(alg);
            std::vector<double> c(16, 0.0);
            c[15] = 1.0;
            return Multivector{std::move(c), 4, alg}.product(LCC_OP_GP, m);
        })
        .def("sta_lorentz_gamma", [require_sta](const Multivector &m) {  // :641-647
            require_sta(m.get_algebra());
            return 2.0 * m.coeffs[0] * m.coeffs[0] - 1.0;
        });

    // ------------------------------------------------------------------ MultivectorBatch (src/batch.rs)
    bind_multivector_methods(m, mv);

    py::class_<Batch, BatchPtr>(m, "MultivectorBatch")
        .def_static("from_numpy", &batch_from_numpy, py::arg("algebra"), py::arg("coeffs"))
        .def_static("from_vectors", [](const PyAlgebra &a, const py::array &coords) {  // :112-138
            if (coords.ndim() != 2) throw py::type_error("coords must be a 2-D array of shape (N, dim)");
            if (coords.shape(1) != a.inner->dim()) {
                std::ostringstream s; s << "coords must have " << a.inner->dim() << " columns for " << a.inner->str() << " but got " << coords.shape(1);
                throw py::value_error(s.str());
            }
            return batch_from_columns(a, coords, vector_blades(a.inner->dim()), "coords");
        }, py::arg("algebra"), py::arg("coords"))
        .def_static("from_scalars", [](const PyAlgebra &a, const py::array &values) {  // :156-171
            if (values.ndim() != 1) throw py::type_error("values must be a 1-D array of shape (N,)");
            return batch_from_columns(a, values, std::vector<int32_t>{0}, "values");
        }, py::arg("algebra"), py::arg("values"))
        .def_static("from_bivectors", [](const PyAlgebra &a, const py::array &biv) {  // :191-221
            if (biv.ndim() != 2) throw py::type_error("bivector_coeffs must be a 2-D array");
            const int dim = a.inner->dim(), nbiv = dim * (dim - 1) / 2;
            if (biv.shape(1) != nbiv) {
                std::ostringstream s; s << "bivector_coeffs must have " << nbiv << " columns for " << a.inner->str() << " but got " << biv.shape(1);
                throw py::value_error(s.str());
            }
            return batch_from_columns(a, biv, blades_of_grade(a.inner->nb(), 2), "bivector_coeffs");
        }, py::arg("algebra"), py::arg("bivector_coeffs"))
        .def("to_numpy", [](const Batch &b, py::object out) { return batch_to_numpy(b, out); }, py::arg("out") = py::none())
        .def("to_vector_coords", [](const Batch &b) { return batch_gather(b, vector_blades(b.algebra->dim()), false); })
        .def("scalar", [](const Batch &b) { return batch_gather(b, std::vector<int32_t>{0}, true); })
        .def("len", &Batch::n)
        .def("__len__", &Batch::n)
        .def("is_empty", [](const Batch &b) { return b.n() == 0; })
        .def("num_blades", &Batch::nb)
        .def("get", [](const Batch &b, int64_t i) { return batch_get(b, i); })
        .def("set", [](Batch &b, int64_t i, const Multivector &mv) { batch_set(b, i, mv); })
        .def("__getitem__", [](const Batch &b, int64_t index) {  // :1004-1014
            const int64_t idx = index < 0 ? b.n() + index : index;
            if (idx < 0 || idx >= b.n()) {
                std::ostringstream s; s << "index " << index << " out of bounds for batch of size " << b.n();
                throw py::index_error(s.str());
            }
            return batch_get(b, idx);
        })
        .def("geometric_product", [](const Batch &a, const Batch &b) { return batch_product(a, b, LCC_OP_GP); })
        .def("gp", [](const Batch &a, const Batch &b) { return batch_product(a, b, LCC_OP_GP); })
        .def("outer_product", [](const Batch &a, const Batch &b) { return batch_product(a, b, LCC_OP_OUTER); })
        .def("wedge", [](const Batch &a, const Batch &b) { return batch_product(a, b, LCC_OP_OUTER); })
        .def("left_contraction", [](const Batch &a, const Batch &b) { return batch_product(a, b, LCC_OP_LC); })
        .def("inner", [](const Batch &a, const Batch &b) { return batch_product(a, b, LCC_OP_LC); })
        .def("lc", [](const Batch &a, const Batch &b) { return batch_product(a, b, LCC_OP_LC); })
        // batched forms of the rest of the scalar product family (src/multivector.rs:1616-1797); new surface
        .def("right_contraction", [](const Batch &a, const Batch &b) { return batch_product(a, b, LCC_OP_RC); })
        .def("rc", [](const Batch &a, const Batch &b) { return batch_product(a, b, LCC_OP_RC); })
        .def("scalar_product", [](const Batch &a, const Batch &b) { return batch_product(a, b, LCC_OP_SCALAR); })
        .def("dot", [](const Batch &a, const Batch &b) { return batch_product(a, b, LCC_OP_SCALAR); })
        .def("commutator", [](const Batch &a, const Batch &b) { return batch_product(a, b, LCC_OP_COMMUTATOR); })
        .def("anticommutator", [](const Batch &a, const Batch &b) { return batch_product(a, b, LCC_OP_ANTICOMMUTATOR); })
        .def("regressive", [](const Batch &a, const Batch &b) { return batch_product(a, b, LCC_OP_REGRESSIVE); })
        .def("meet", [](const Batch &a, const Batch &b) { return batch_product(a, b, LCC_OP_REGRESSIVE); })
        .def("sandwich", &batch_sandwich)
        .def("apply", &batch_sandwich)
        .def_static("sandwich_numpy", [](const PyAlgebra &alg, const Multivector &rotor, const py::array &x_in, py::object out_obj) {
            // from_numpy(x).sandwich(rotor).to_numpy() as ONE host-array call (lcc_host_sandwich): the upload of chunk
            // c+1, the kernel of chunk c and the download of chunk c-1 overlap, so both PCIe directions are busy at once
            // -- the three-call sequence cannot do that, each of its calls has to finish before the next starts
            AlgebraPtr ralg = rotor.get_algebra();
            if (!alg.inner->same_signature(*ralg))
                throw py::value_error("algebra mismatch: batch is " + alg.inner->str() + " but rotor is " + ralg->str());
            py::array x = contiguous(x_in);
            const int dtype = dtype_of(x);
            const int nb = alg.inner->nb();
            if (x.ndim() != 2) throw py::type_error("coeffs must be a 2-D array of shape (N, num_blades)");
            if (x.shape(1) != nb) {
                std::ostringstream s; s << "coeffs must have " << nb << " columns for " << alg.inner->str() << " but got " << x.shape(1);
                throw py::value_error(s.str());
            }
            const std::vector<py::ssize_t> shape{x.shape(0), (py::ssize_t)nb};
            py::array out;
            if (out_obj.is_none()) out = output_array(dtype, shape);
            else {
                out = py::array::ensure(out_obj);
                if (!out || dtype_of(out) != dtype || out.ndim() != 2 || out.shape(0) != shape[0] || out.shape(1) != shape[1] ||
                    !(out.flags() & py::array::c_style) || !out.writeable())
                    throw py::value_error("out must be a writeable C-contiguous (N, num_blades) array of x's dtype");
            }
            const int64_t n = x.shape(0);
            if (n == 0) return out;
            const void *host = x.data();
            void *out_ptr = out.mutable_data();
            const std::vector<double> r = rotor.coeffs;
            const unsigned flags = fp_flags();
            py::gil_scoped_release nogil;
            check(lcc_host_sandwich(alg.inner->h, dtype, r.data(), host, out_ptr, n, flags));
            return out;
        }, py::arg("algebra"), py::arg("rotor"), py::arg("x"), py::arg("out") = py::none())
        .def("norm", [](const Batch &a) { return batch_norm(a, LCC_NORM); })
        .def("norm_squared", [](const Batch &a) { return batch_norm(a, LCC_NORM_SQUARED); })
        .def("normalized", [](const Batch &a) {
            auto out = a.empty_like(a.n());
            py::gil_scoped_release nogil;
            check(lcc_batch_normalized(a.algebra->h, &a.view, &out->view, fp_flags(), nullptr));
            return out;
        })
        .def("reverse", [](const Batch &a) { return batch_unary(a, LCC_UN_REVERSE); })
        .def("grade_involution", [](const Batch &a) { return batch_unary(a, LCC_UN_INVOLUTE); })
        .def("grade", [](const Batch &a, long k) {
            if (k < 0) throw py::type_error("grade must be a non-negative integer");
            return batch_unary(a, LCC_UN_GRADE, (int)std::min<long>(k, 1 << 20));
        })
        .def("__add__", [](const Batch &a, const Batch &b) { return batch_addsub(a, b, LCC_EW_ADD); })
        .def("__sub__", [](const Batch &a, const Batch &b) { return batch_addsub(a, b, LCC_EW_SUB); })
        .def("scale", [](const Batch &a, double f) { return batch_scale(a, f); })
        .def("__neg__", [](const Batch &a) { return batch_scale(a, -1.0); })
        .def("__mul__", [](const Batch &a, double f) { return batch_scale(a, f); })
        .def("__rmul__", [](const Batch &a, double f) { return batch_scale(a, f); })
        .def("__truediv__", [](const Batch &a, double s) {  // :980-987
            if (s == 0.0) { PyErr_SetString(PyExc_ZeroDivisionError, "division by zero"); throw py::error_already_set(); }
            return batch_scale(a, 1.0 / s);
        })
        .def("__repr__", [](const Batch &b) {
            std::ostringstream s; s << "MultivectorBatch(size=" << b.n() << ", algebra=" << b.algebra->str() << ")";
            return s.str();
        })
        .def("concat", [](const Batch &a, const Batch &b) { return batch_concat(a, b); })
        .def("slice", [](const Batch &b, int64_t s, int64_t e) { return batch_slice(b, s, e); })
        // ---- additive surface
        .def("clifford_conjugate", [](const Batch &a) { return batch_unary(a, LCC_UN_CONJUGATE); })
        .def("even", [](const Batch &a) { return batch_unary(a, LCC_UN_EVEN); })
        .def("odd", [](const Batch &a) { return batch_unary(a, LCC_UN_ODD); })
        .def("dual", [](const Batch &a) { return batch_unary(a, LCC_UN_DUAL); })
        .def("undual", [](const Batch &a) { return batch_unary(a, LCC_UN_UNDUAL); })
        .def("right_dual", [](const Batch &a) { return batch_unary(a, LCC_UN_RIGHT_DUAL); })
        .def("pga_dual", [](const Batch &a) { return batch_unary(a, LCC_UN_PGA_DUAL); })
        .def("gp_and_outer", [](const Batch &a, const Batch &b) {
            require_same_algebra(a, b);
            const int64_t n = broadcast_n(a, b);
            auto g = a.empty_like(n), w = a.empty_like(n);
            py::gil_scoped_release nogil;
            check(lcc_batch_gp_outer(a.algebra->h, &a.view, &b.view, &g->view, &w->view, fp_flags(), nullptr));
            return std::make_pair(g, w);
        })
        .def("product_sum", [](const Batch &a, const Batch &b, const std::string &op, py::object grade) {
            // sums over the batch of <a>_grade (op) b in one fused pass (lcc_batch_product_sum)
            require_same_algebra(a, b);
            const int64_t n = broadcast_n(a, b);
            (void)n;
            int opc;
            if (op == "gp" || op == "geometric_product") opc = LCC_OP_GP;
            else if (op == "outer" || op == "wedge" || op == "outer_product") opc = LCC_OP_OUTER;
            else if (op == "lc" || op == "inner" || op == "left_contraction") opc = LCC_OP_LC;
            else throw py::value_error("op must be one of 'gp', 'outer', 'lc'");
            const int g = grade.is_none() ? -1 : grade.cast<int>();
            if (g < -1) throw py::value_error("grade must be non-negative");
            const int nb = a.nb();
            py::array out(np_dtype(a.view.dtype), std::vector<py::ssize_t>{(py::ssize_t)nb});
            void *out_ptr = out.mutable_data();
            py::gil_scoped_release nogil;
            DeviceTemp stage((size_t)nb * a.es());
            check(lcc_batch_product_sum(a.algebra->h, opc, &a.view, g, &b.view, stage.ptr, fp_flags(), nullptr));
            check(lcc_memcpy_d2h(out_ptr, stage.ptr, (size_t)nb * a.es(), nullptr));
            check(lcc_stream_sync(nullptr));
            return out;
        }, py::arg("other"), py::arg("op") = "gp", py::arg("grade") = py::none())
        .def("sum", [](const Batch &a) { return batch_sum(a, false); })
        .def("mean", [](const Batch &a) { return batch_sum(a, true); })
        .def("algebra", [](const Batch &b) { return PyAlgebra{b.algebra}; })
        .def_property_readonly("dtype", [](const Batch &b) { return np_dtype(b.view.dtype); })
        .def_property_readonly("device", [](const Batch &b) { return b.device; })
        .def_property_readonly("data_ptr", [](const Batch &b) { return (uintptr_t)b.view.data; })
        .def_property_readonly("ld", [](const Batch &b) { return b.view.ld; })
        .def_static("from_points_pga", [](const PyAlgebra &a, const py::array &xyz_in) {
            py::array xyz = contiguous(xyz_in);
            const int dtype = dtype_of(xyz);
            if (xyz.ndim() != 2 || xyz.shape(1) != 3) throw py::value_error("points must have 3 columns");
            auto b = std::make_shared<Batch>(a.inner, dtype, xyz.shape(0));
            if (xyz.shape(0) == 0) { require_pga3d(a, "pga_point"); return b; }
            const void *host = xyz.data();
            const size_t bytes = (size_t)xyz.nbytes();
            py::gil_scoped_release nogil;
            DeviceTemp stage(bytes);
            upload(host, bytes, stage.ptr);
            check(lcc_pga_points_embed(a.inner->h, stage.ptr, &b->view, nullptr));
            return b;
        })
        .def_static("from_points_cga", [](const PyAlgebra &a, const py::array &xyz_in) {
            py::array xyz = contiguous(xyz_in);
            const int dtype = dtype_of(xyz);
            if (xyz.ndim() != 2 || xyz.shape(1) != 3) throw py::value_error("points must have 3 columns");
            auto b = std::make_shared<Batch>(a.inner, dtype, xyz.shape(0));
            if (xyz.shape(0) == 0) return b;
            const void *host = xyz.data();
            const size_t bytes = (size_t)xyz.nbytes();
            py::gil_scoped_release nogil;
            DeviceTemp stage(bytes);
            upload(host, bytes, stage.ptr);
            check(lcc_cga_points_embed(a.inner->h, stage.ptr, &b->view, nullptr));
            return b;
        })
        // batched sta_vector / sta_bivector / sta_boost (src/sta.rs:105-109,202-212,253-296)
        .def_static("from_events_sta", [](const PyAlgebra &a, const py::array &txyz_in) {
            py::array txyz = contiguous(txyz_in);
            const int dtype = dtype_of(txyz);
            if (!(a.inner->p == 1 && a.inner->q == 3 && a.inner->r == 0)) throw py::value_error("requires an STA algebra Cl(1,3,0); use Algebra.sta()");
            if (txyz.ndim() != 2 || txyz.shape(1) != 4) throw py::value_error("events must have 4 columns (t, x, y, z)");
            auto b = std::make_shared<Batch>(a.inner, dtype, txyz.shape(0));
            if (txyz.shape(0) == 0) return b;
            const void *host = txyz.data();
            const size_t bytes = (size_t)txyz.nbytes();
            static const int32_t blades[4] = {1, 2, 4, 8};
            py::gil_scoped_release nogil;
            DeviceTemp stage(bytes);
            upload(host, bytes, stage.ptr);
            check(lcc_batch_scatter_columns(a.inner->h, stage.ptr, 4, blades, &b->view, nullptr));
            return b;
        })
        .def_static("from_fields_sta", [](const PyAlgebra &a, const py::array &e_in, const py::array &b_in) {
            py::array e = contiguous(e_in), bf = contiguous(b_in);
            const int dtype = dtype_of(e);
            if (dtype_of(bf) != dtype) throw py::type_error("E and B must share one dtype");
            if (e.ndim() != 2 || e.shape(1) != 3 || bf.ndim() != 2 || bf.shape(1) != 3 || bf.shape(0) != e.shape(0))
                throw py::value_error("E and B must both be (n, 3) arrays");
            const py::ssize_t n = e.shape(0);
            const size_t es = dtype == LCC_F32 ? 4 : 8;
            std::vector<char> packed((size_t)n * 6 * es);  // (Ex,Ey,Ez,Bx,By,Bz) rows
            for (py::ssize_t r = 0; r < n; ++r) {
                std::memcpy(packed.data() + (size_t)r * 6 * es, (const char *)e.data() + (size_t)r * 3 * es, 3 * es);
                std::memcpy(packed.data() + ((size_t)r * 6 + 3) * es, (const char *)bf.data() + (size_t)r * 3 * es, 3 * es);
            }
            auto b = std::make_shared<Batch>(a.inner, dtype, n);
            py::gil_scoped_release nogil;
            DeviceTemp stage(packed.size() ? packed.size() : 16);
            if (n) upload(packed.data(), packed.size(), stage.ptr);
            check(lcc_sta_bivectors_embed(a.inner->h, stage.ptr, &b->view, nullptr));
            return b;
        })
        .def_static("from_boosts_sta", [](const PyAlgebra &a, const py::array &v_in) {
            py::array v = contiguous(v_in);
            const int dtype = dtype_of(v);
            if (v.ndim() != 2 || v.shape(1) != 3) throw py::value_error("velocity must have 3 components");
            auto b = std::make_shared<Batch>(a.inner, dtype, v.shape(0));
            const void *host = v.data();
            const size_t bytes = (size_t)v.nbytes();
            py::gil_scoped_release nogil;
            DeviceTemp stage(bytes ? bytes : 16);
            if (bytes) upload(host, bytes, stage.ptr);
            check(lcc_sta_boosts_embed(a.inner->h, stage.ptr, &b->view, nullptr));
            return b;
        })
        .def("pga_point_coords", [](const Batch &b) {
            py::array out(np_dtype(b.view.dtype), std::vector<py::ssize_t>{(py::ssize_t)b.n(), 3});
    void *out_ptr = out.mutable_data();
            if (b.n() == 0) return out;
            py::gil_scoped_release nogil;
            DeviceTemp stage((size_t)b.n() * 3 * b.es());
            check(lcc_pga_points_extract(b.algebra->h, &b.view, stage.ptr, nullptr));
            check(lcc_memcpy_d2h(out_ptr, stage.ptr, (size_t)b.n() * 3 * b.es(), nullptr));
            check(lcc_stream_sync(nullptr));
            return out;
        })
        .def("cga_extract_center", [](const Batch &b) {
            py::array out(np_dtype(b.view.dtype), std::vector<py::ssize_t>{(py::ssize_t)b.n(), 3});
    void *out_ptr = out.mutable_data();
            if (b.n() == 0) return out;
            py::gil_scoped_release nogil;
            DeviceTemp stage((size_t)b.n() * 3 * b.es());
            check(lcc_cga_points_extract(b.algebra->h, &b.view, stage.ptr, nullptr));
            check(lcc_memcpy_d2h(out_ptr, stage.ptr, (size_t)b.n() * 3 * b.es(), nullptr));
            check(lcc_stream_sync(nullptr));
            return out;
        });
}
