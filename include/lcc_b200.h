/*
 * lcc_b200.h -- C ABI of the B200-native batched Clifford-algebra engine.
 *
 * This is the drop-in boundary for ONE path of LargeCrimsonCanine/largecrimsoncanine: the batched
 * products of src/batch.rs (+ the table machinery of src/algebra/ they read and the single-pair
 * kernel of src/simd.rs).  The reference has no FFI of its own -- its boundary is the PyO3 module
 * `largecrimsoncanine` (src/lib.rs:37-45) -- so every entry point below names the reference
 * method it replaces (file:line relative to the reference root).  INTEGRATION.md shows the
 * binding a maintainer would add on the Rust side (`extern "C"` block + safe wrappers).
 *
 * Conventions
 *   - Plain C: pointers, sizes, enums.  No torch / numpy / C++ types cross this boundary.
 *   - Every function returns an lcc_status; lcc_last_error() gives the thread-local message.
 *     LCC_ERR_VALUE / LCC_ERR_INDEX / LCC_ERR_ZERODIV map to the reference's ValueError /
 *     IndexError / ZeroDivisionError (src/batch.rs:83-88,289-295,981-985); everything CUDA is
 *     LCC_ERR_CUDA.  There is NO CPU fallback: without a usable CUDA device every compute entry
 *     point fails with LCC_ERR_NO_DEVICE.
 *   - `stream` is a cudaStream_t passed as void*.  NULL means the library's own non-blocking
 *     stream of the current device (lcc_stream()); pass cudaStreamLegacy / cudaStreamPerThread
 *     explicitly if the default stream is wanted.  All compute calls are stream-ordered and
 *     never synchronise the host unless the name says `_sync` or the data lands in host memory.
 *   - Device data is described by lcc_view.  Two layouts:
 *       LCC_SOA  structure-of-arrays, element (row r, blade k) at data[k*ld + r]   (engine native)
 *       LCC_AOS  array-of-structures, element (row r, blade k) at data[r*ld + k]   (the reference's
 *                `Array2<f64>` (N, num_blades) row-major layout, src/batch.rs:41-47; ld >= num_blades)
 *     SoA buffers allocated by lcc_batch_alloc have ld rounded up so every blade row starts
 *     16-byte aligned; rows n..ld-1 are padding and are never read by reductions.
 *   - A view with n == 1 used as a binary operand broadcasts (src/batch.rs:361-376).
 */
#ifndef LCC_B200_H
#define LCC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define LCC_API
#else
#define LCC_API __attribute__((visibility("default")))
#endif

typedef enum {
    LCC_OK = 0,
    LCC_ERR_VALUE = 1,       /* ValueError in the reference */
    LCC_ERR_INDEX = 2,       /* IndexError */
    LCC_ERR_ZERODIV = 3,     /* ZeroDivisionError */
    LCC_ERR_CUDA = 4,        /* a CUDA runtime call failed */
    LCC_ERR_NO_DEVICE = 5,   /* no CUDA device / driver: compute is refused, never emulated */
    LCC_ERR_UNSUPPORTED = 6, /* valid request this build has no kernel for */
    LCC_ERR_ALLOC = 7
} lcc_status;

typedef enum { LCC_F32 = 0, LCC_F64 = 1 } lcc_dtype;
typedef enum { LCC_SOA = 0, LCC_AOS = 1 } lcc_layout;

/* Binary products, src/batch.rs:353-410 (geometric), 514-574 (outer), 585-655 (left contraction).
 * The rest of the row-wise product family has no batched form in the reference; these are the batched
 * forms of Multivector::right_contraction (src/multivector.rs:1616-1661), scalar_product (:1673-1715),
 * commutator (:1732-1745), anticommutator (:1755-1768) and regressive (:1781-1794, LCC_ERR_VALUE when the
 * pseudoscalar is null).  Code 3 is reserved. */
typedef enum {
    LCC_OP_GP = 0, LCC_OP_OUTER = 1, LCC_OP_LC = 2, LCC_OP_RC = 4, LCC_OP_SCALAR = 5,
    LCC_OP_COMMUTATOR = 6, LCC_OP_ANTICOMMUTATOR = 7, LCC_OP_REGRESSIVE = 8
} lcc_product_op;

/* Row-wise unary maps.  REVERSE / INVOLUTE: src/batch.rs:823-860.  GRADE: batch.rs:1027-1052 (arg = k).
 * CONJUGATE: src/algebra/blades.rs:182-188.  DUAL / UNDUAL / RIGHT_DUAL: src/multivector.rs:4249-4290
 * (batched form is new surface; LCC_ERR_VALUE when the pseudoscalar is null, as the reference raises).
 * PGA_DUAL: src/pga.rs:521-556.  EVEN / ODD: lcc/torch/multivector.py:766-780.  COPY: identity. */
typedef enum {
    LCC_UN_REVERSE = 0, LCC_UN_INVOLUTE = 1, LCC_UN_CONJUGATE = 2, LCC_UN_GRADE = 3,
    LCC_UN_DUAL = 4, LCC_UN_UNDUAL = 5, LCC_UN_RIGHT_DUAL = 6, LCC_UN_PGA_DUAL = 7,
    LCC_UN_EVEN = 8, LCC_UN_ODD = 9, LCC_UN_COPY = 10
} lcc_unary_op;

/* src/batch.rs:863-954 */
typedef enum { LCC_EW_ADD = 0, LCC_EW_SUB = 1 } lcc_elementwise_op;

/* src/batch.rs:677-817 */
typedef enum { LCC_NORM_SQUARED = 0, LCC_NORM = 1 } lcc_norm_kind;

/* Flags for compute calls. */
enum {
    /* Reproduce the reference's rounding sequence exactly: separate multiply and add, left blade
     * ascending (Rust never contracts to FMA).  Default (0) uses FMA in the same term order. */
    LCC_FLAG_STRICT_FP = 1
};

typedef struct lcc_algebra lcc_algebra; /* opaque; replaces Arc<Algebra>, src/algebra/registry.rs:36-49 */

typedef struct {
    void *data;     /* device pointer */
    int64_t n;      /* rows (multivectors) */
    int64_t ld;     /* SoA: elements between consecutive blades; AoS: elements between consecutive rows */
    int32_t dtype;  /* lcc_dtype */
    int32_t layout; /* lcc_layout */
} lcc_view;

/* ---------------------------------------------------------------- library / device */

LCC_API const char *lcc_version(void);            /* src/lib.rs:40 (__version__) */
LCC_API const char *lcc_last_error(void);         /* thread-local message of the last failing call */
LCC_API lcc_status lcc_device_count(int *count);  /* LCC_ERR_NO_DEVICE when there is none */
LCC_API lcc_status lcc_set_device(int device);
LCC_API lcc_status lcc_get_device(int *device);
LCC_API lcc_status lcc_stream(void **stream);     /* the library's stream on the current device */
LCC_API lcc_status lcc_stream_sync(void *stream);

/* ---------------------------------------------------------------- algebra (host side, no GPU needed) */

/* Algebra::new, src/algebra/registry.rs:70-81; PyAlgebra::new, src/pyalgebra.rs:53-60.  n = p+q+r <= 8. */
LCC_API lcc_status lcc_algebra_create(int p, int q, int r, lcc_algebra **out);
LCC_API void lcc_algebra_destroy(lcc_algebra *alg);
LCC_API int lcc_algebra_p(const lcc_algebra *alg);          /* src/pyalgebra.rs:141-157 */
LCC_API int lcc_algebra_q(const lcc_algebra *alg);
LCC_API int lcc_algebra_r(const lcc_algebra *alg);
LCC_API int lcc_algebra_dimension(const lcc_algebra *alg);  /* src/pyalgebra.rs:129-133 */
LCC_API int lcc_algebra_num_blades(const lcc_algebra *alg); /* src/pyalgebra.rs:135-139 */
/* Flat Cayley tables [a*nb+b]: registry.rs:192-202 (signs as int8 in {-1,0,+1}; blades as uint16). */
LCC_API const int8_t *lcc_algebra_cayley_signs(const lcc_algebra *alg);
LCC_API const uint16_t *lcc_algebra_cayley_blades(const lcc_algebra *alg);
/* Grade masks, blades.rs:211-229 / registry.rs:224-232: one bit per blade, ceil(nb/64) words per
 * grade (the reference's single usize is words==1, i.e. dimension <= 6).  Returns words per grade. */
LCC_API int lcc_algebra_grade_mask(const lcc_algebra *alg, int grade, uint64_t *words, int capacity);
/* blade_name, blades.rs:48-61; writes a NUL-terminated name, returns its length (or -1). */
LCC_API int lcc_algebra_blade_name(const lcc_algebra *alg, int index, char *buf, int capacity);
/* 1 when this build carries kernels specialised at build time for the signature (generated Cayley
 * term lists); 0 when it will run the generic run-time-metric kernels. */
LCC_API int lcc_algebra_is_specialised(const lcc_algebra *alg);
/* blades.rs:128-188 -- +1 / -1 for a blade index */
LCC_API int lcc_reverse_sign(int blade_index);
LCC_API int lcc_grade_involution_sign(int blade_index);
LCC_API int lcc_clifford_conjugate_sign(int blade_index);

/* Real matrix representation of a non-degenerate 7- or 8-dimensional algebra by 16 x 16 signed permutation matrices
 * (csrc/matrix_rep.hpp; host side, no GPU needed): blade A = sign_of_blade[A] * X^a Z^b with string_of_blade[A] = a << 4 | b,
 * X^a Z^b [i][j] = [i == j ^ a] * (-1)^popcount(b & j).  Returns the matrix size (16), or 0 when the algebra has no such
 * representation (then fixed-operand products stay on the dense GEMM).  Buffers hold num_blades entries; either may be NULL. */
LCC_API int lcc_algebra_matrix_rep(const lcc_algebra *alg, uint8_t *string_of_blade, int8_t *sign_of_blade);
/* Diagnostic (host side): the shared-memory layout K8 derives from that representation (csrc/matrix_rep.hpp, RepLayout).
 * tables receives 8 x 16 words: byte offsets F(blade of string b), F(blade of string a << 4); then for the column pass
 * (fixed factor on the left) and the row pass (on the right) each: slot offsets of the contraction index at the MMA's
 * physical positions, slot offsets Gamma(w) of the free index, and the physical -> logical index map.  Returns 1, or 0
 * when the algebra has no representation / no conflict-free layout (K8 is then not used).  tables may be NULL. */
LCC_API int lcc_algebra_matrix_rep_layout(const lcc_algebra *alg, uint32_t *tables);
/* The same for K8's units of whole reference-layout rows (RepLayoutAos; element (row r, blade A) of a unit at G(A) ^ R(r)):
 * elem_bytes 4 or 8; tables receives 13 x 16 words: G(blade of string b), G(blade of string row amap[i] << 4), amap, then per
 * pass (column, row): G of the contraction index at the MMA's k positions, G of the output index at its m positions,
 * G(Gamma(w)), and the two physical -> logical maps.  Returns 1, or 0 when there is none (reference-layout batches are then
 * transposed through temporaries).  tables may be NULL. */
LCC_API int lcc_algebra_matrix_rep_layout_aos(const lcc_algebra *alg, int elem_bytes, uint32_t *tables);

/* ---------------------------------------------------------------- device memory (stream-ordered pool) */

LCC_API lcc_status lcc_device_alloc(void **ptr, size_t bytes, void *stream);
LCC_API lcc_status lcc_device_free(void *ptr, void *stream);
LCC_API lcc_status lcc_host_alloc_pinned(void **ptr, size_t bytes);
LCC_API lcc_status lcc_host_free_pinned(void *ptr);
LCC_API lcc_status lcc_memcpy_h2d(void *dst, const void *src, size_t bytes, void *stream);
LCC_API lcc_status lcc_memcpy_d2h(void *dst, const void *src, size_t bytes, void *stream);
LCC_API lcc_status lcc_memcpy_d2d(void *dst, const void *src, size_t bytes, void *stream);
LCC_API lcc_status lcc_memset_zero(void *dst, size_t bytes, void *stream);
/* Allocate an engine-native SoA batch of n rows for `alg`; fills view (ld is padded). */
LCC_API lcc_status lcc_batch_alloc(const lcc_algebra *alg, int dtype, int64_t n, lcc_view *view, void *stream);
LCC_API lcc_status lcc_batch_free(lcc_view *view, void *stream);

/* ---------------------------------------------------------------- layout conversion (K7) */

/* Copy between any two views of equal n and dtype (AoS<->SoA transposes, SoA->SoA, AoS->AoS).
 * Replaces the host copies of from_numpy / to_numpy, src/batch.rs:79-94,233-235. */
LCC_API lcc_status lcc_batch_convert(const lcc_algebra *alg, const lcc_view *src, const lcc_view *dst, void *stream);
/* Scatter `ncols` dense columns (AoS (n, ncols) device array, row stride = ncols) to the listed
 * blades of a zero-filled batch: from_vectors / from_scalars / from_bivectors, batch.rs:112-221. */
LCC_API lcc_status lcc_batch_scatter_columns(const lcc_algebra *alg, const void *cols, int ncols,
                                             const int32_t *blade_of_col, const lcc_view *dst, void *stream);
/* Inverse gather: to_vector_coords / scalar, batch.rs:248-261,1017-1024. */
LCC_API lcc_status lcc_batch_gather_columns(const lcc_algebra *alg, const lcc_view *src, int ncols,
                                            const int32_t *blade_of_col, void *cols, void *stream);
/* Rows [start, end) of src into dst rows [dst_start, ...): slice / concat / get / set,
 * batch.rs:288-335,1055-1097. */
LCC_API lcc_status lcc_batch_copy_rows(const lcc_algebra *alg, const lcc_view *src, int64_t start, int64_t end,
                                       const lcc_view *dst, int64_t dst_start, void *stream);

/* ---------------------------------------------------------------- the hot path */

/* out = a (op) b row by row; a.n == b.n or either == 1 (broadcast).  batch.rs:353-410,514-574,585-655.
 * All three views share dtype and layout.  out.n must equal max(a.n, b.n). */
LCC_API lcc_status lcc_batch_product(const lcc_algebra *alg, int op, const lcc_view *a, const lcc_view *b,
                                     const lcc_view *out, unsigned flags, void *stream);
/* Fused dual-output: gp_out = a*b and outer_out = a^b reading a and b once (BASELINE config 3). */
LCC_API lcc_status lcc_batch_gp_outer(const lcc_algebra *alg, const lcc_view *a, const lcc_view *b,
                                      const lcc_view *gp_out, const lcc_view *outer_out, unsigned flags, void *stream);
/* The broadcast branch with the fixed operand on the HOST (num_blades doubles): out = fixed (op) x (side 0) or
 * x (op) fixed (side 1), batch.rs:365-404 with a one-row left / right batch.  Knowing the operand on the host lets the
 * engine pick the cheapest form without reading device memory back: in 128 / 256-blade non-degenerate algebras that have
 * a real 16 x 16 matrix representation (Cl(8,0), Cl(4,4), Cl(7,0), ... -- csrc/matrix_rep.hpp) the geometric product runs
 * block-diagonalised (K8: two Walsh-Hadamard passes + a 16 x 16 x 16 product per row) instead of as the nb x nb
 * tensor-core GEMM (f64: 6x faster, f32: 1.5x); option "matrix_rep" (below) narrows or disables it.  Views in the
 * reference's row-major layout are processed in place as units of whole rows, blade-major views as units of 128 bytes of
 * rows per blade.  Strict mode always takes the exact-order kernel. */
LCC_API lcc_status lcc_batch_fixed_product(const lcc_algebra *alg, int op, int side, const double *fixed_host, const lcc_view *x,
                                           const lcc_view *out, unsigned flags, void *stream);
/* Fused pipeline  sums[k] = sum over rows of ( <a>_g (op) b )[k]  -- grade projection of the left operand
 * (a_grade < 0: none; batch.rs:1027-1052), product (batch.rs:353-655) and batch sum in ONE pass that reads
 * only the needed blade rows and writes num_blades values (BASELINE config 4; the per-GPU partial that
 * precedes the NCCL all-reduce).  sums_dev: num_blades elements of the operands' dtype on the device. */
LCC_API lcc_status lcc_batch_product_sum(const lcc_algebra *alg, int op, const lcc_view *a, int a_grade, const lcc_view *b,
                                         void *sums_dev, unsigned flags, void *stream);
/* Vector-Jacobian products of the three bilinear products (autograd of lcc.torch; new surface):
 *   wrt = 0  grad_in[i] = sum_j sign[i][j] * grad_out[i^j] * other[j]   (other = the right operand b)
 *   wrt = 1  grad_in[j] = sum_i sign[i][j] * other[i] * grad_out[i^j]   (other = the left operand a)
 * restricted to the op's term set.  `other` may be a one-row batch (broadcast); the caller sums
 * grad_in over rows (lcc_batch_sum) when the differentiated operand itself was broadcast.
 * Cotangents have no reference rounding order: up to 32 blades they are always computed with FMA in
 * ascending-term order and LCC_FLAG_STRICT_FP is ignored; above 32 blades the table kernel honours it. */
LCC_API lcc_status lcc_batch_product_vjp(const lcc_algebra *alg, int op, int wrt, const lcc_view *grad_out,
                                         const lcc_view *other, const lcc_view *grad_in, unsigned flags, void *stream);
/* out[row] = R * x[row] * ~R with one rotor/motor (host array of num_blades doubles) for the whole
 * batch, two rounded stages as in batch.rs:434-503. */
LCC_API lcc_status lcc_batch_sandwich(const lcc_algebra *alg, const double *rotor, const lcc_view *x,
                                      const lcc_view *out, unsigned flags, void *stream);
/* Per-row rotors (rotors.n == x.n), or ONE device-resident rotor row for the whole batch (rotors.n == 1: no host copy
 * of the rotor is needed): lcc/torch/multivector.py:988-1019.  Up to 16 blades one fused kernel (even rotors take pruned
 * term lists); 128 / 256-blade algebras with a real 16 x 16 matrix representation and one rotor: the block-diagonalised
 * two-pass form (K8) with both factors built on the device; otherwise reverse + two products. */
LCC_API lcc_status lcc_batch_sandwich_rows(const lcc_algebra *alg, const lcc_view *rotors, const lcc_view *x,
                                           const lcc_view *out, unsigned flags, void *stream);
/* Cotangents of lcc_batch_sandwich_rows (autograd of lcc.torch; the reference's backward pass is torch autograd through
 * the two products and the reverse of lcc/torch/multivector.py:988-1019): given grad_out = d loss / d out,
 * grad_x = d loss / d x and grad_rotors = d loss / d rotors, one row each per row of x -- for a broadcast rotor
 * (rotors.n == 1) the caller sums grad_rotors over the rows.  Up to 16 blades this is ONE kernel with R, x and grad_out in
 * registers (5 * num_blades * s bytes per row); larger algebras compose the five bilinear passes with R x recomputed.
 * FMA arithmetic (LCC_FLAG_STRICT_FP is ignored up to 16 blades, as for lcc_batch_product_vjp). */
LCC_API lcc_status lcc_batch_sandwich_rows_vjp(const lcc_algebra *alg, const lcc_view *rotors, const lcc_view *x,
                                               const lcc_view *grad_out, const lcc_view *grad_x, const lcc_view *grad_rotors,
                                               unsigned flags, void *stream);
/* Signed permutations / masks, see lcc_unary_op.  `arg` is the grade for LCC_UN_GRADE. */
LCC_API lcc_status lcc_batch_unary(const lcc_algebra *alg, int op, int arg, const lcc_view *x,
                                   const lcc_view *out, void *stream);
/* out = x * factor (batch.rs:957-962); division is the caller's scale(1.0/s) (batch.rs:980-987). */
LCC_API lcc_status lcc_batch_scale(const lcc_algebra *alg, const lcc_view *x, double factor,
                                   const lcc_view *out, void *stream);
/* out = a +/- b with the size-1 broadcast, batch.rs:863-954. */
LCC_API lcc_status lcc_batch_elementwise(const lcc_algebra *alg, int op, const lcc_view *a, const lcc_view *b,
                                         const lcc_view *out, void *stream);
/* norms[row] (device array of n elements of x's dtype): batch.rs:677-762. */
LCC_API lcc_status lcc_batch_norm(const lcc_algebra *alg, int kind, const lcc_view *x, void *norms,
                                  unsigned flags, void *stream);
/* Rows with norm > 1e-14 divided by their norm, others copied: batch.rs:770-817. */
LCC_API lcc_status lcc_batch_normalized(const lcc_algebra *alg, const lcc_view *x, const lcc_view *out,
                                        unsigned flags, void *stream);
/* Column sums over the batch: sums_dev[k] = sum_r x[r][k] (num_blades elements of x's dtype on the
 * device; NEW surface -- the per-GPU partial that precedes the NCCL all-reduce).  Fixed reduction
 * tree: the result does not depend on timing. */
LCC_API lcc_status lcc_batch_sum(const lcc_algebra *alg, const lcc_view *x, void *sums_dev, void *stream);

/* ---------------------------------------------------------------- PGA / CGA / STA specialisations */

/* Batched embeddings of src/pga.rs:131-134 (pga_point), src/cga.rs:189-206 (cga_point),
 * src/sta.rs:105-109 (sta_vector) and the matching extractors pga.rs:579-599, cga.rs:492-508.
 * xyz / txyz are dense AoS device arrays (n,3) or (n,4) of dst's dtype. */
LCC_API lcc_status lcc_pga_points_embed(const lcc_algebra *alg, const void *xyz, const lcc_view *dst, void *stream);
LCC_API lcc_status lcc_pga_points_extract(const lcc_algebra *alg, const lcc_view *src, void *xyz, void *stream);
LCC_API lcc_status lcc_cga_points_embed(const lcc_algebra *alg, const void *xyz, const lcc_view *dst, void *stream);
LCC_API lcc_status lcc_cga_points_extract(const lcc_algebra *alg, const lcc_view *src, void *xyz, void *stream);
/* Fused  xyz_out = pga_point_coords( M * pga_point(xyz_in) * ~M )  on dense (n,3) device arrays: embedding
 * (pga.rs:131-134), the two-stage sandwich (batch.rs:452-497) and extraction (pga.rs:579-599) in one pass that
 * moves 6 elements per point instead of 32.  motor: 16 doubles on the host.  Points whose transformed weight is
 * below 1e-10 come back as NaN (the scalar API raises ValueError there).  In-place (xyz_out == xyz_in) is allowed. */
LCC_API lcc_status lcc_pga_points_sandwich(const lcc_algebra *alg, int dtype, const double *motor, const void *xyz_in,
                                           void *xyz_out, int64_t n, unsigned flags, void *stream);
/* STA: sta_vector(t,x,y,z) is lcc_batch_scatter_columns onto blades 1,2,4,8 (src/sta.rs:105-109).
 * sta_bivector (src/sta.rs:202-212): eb = dense (n,6) rows (Ex,Ey,Ez,Bx,By,Bz).
 * sta_boost (src/sta.rs:253-296): velocity = dense (n,3); rows with |v| >= 1 (no rotor; ValueError in the
 * scalar API) come back as NaN. */
LCC_API lcc_status lcc_sta_bivectors_embed(const lcc_algebra *alg, const void *eb, const lcc_view *dst, void *stream);
LCC_API lcc_status lcc_sta_boosts_embed(const lcc_algebra *alg, const void *velocity, const lcc_view *dst, void *stream);
/* ---------------------------------------------------------------- host <-> device (from_numpy / to_numpy) */

/* from_numpy / to_numpy, src/batch.rs:79-94,233-235: copy a HOST array in the reference's (n, num_blades) row-major
 * layout into / out of a device batch of the same dtype (any layout; rows must match).  The copy is chunked and
 * pipelined: pinned host memory (lcc_host_alloc_pinned, lcc_host_cache_alloc, cudaHostRegister'ed arrays -- detected,
 * not declared) is DMA'd directly, pageable memory goes through a ring of pinned bounce slots filled by a small pool of
 * host threads; the layout transpose of chunk c runs while chunk c+1 crosses PCIe, and the only extra device memory is
 * two chunk-sized staging buffers per device (option "host_chunk_bytes", default 64 MiB).
 * lcc_batch_upload returns when the host array has been read completely (it may be reused); the batch itself is
 * complete in `stream` order.  lcc_batch_download returns when `host` holds the whole batch. */
LCC_API lcc_status lcc_batch_upload(const lcc_algebra *alg, const void *host, int64_t n, const lcc_view *dst, void *stream);
LCC_API lcc_status lcc_batch_download(const lcc_algebra *alg, const lcc_view *src, void *host, void *stream);
/* Cache of pinned host blocks: what to_numpy() hands out for results of at least "pinned_output_min_bytes" (default
 * 1 MiB), so that repeated exports pay neither cudaHostAlloc nor first-touch page faults.  Freed blocks stay cached up
 * to LCC_PINNED_CACHE_MAX_BYTES (default: a quarter of the box's RAM); lcc_host_cache_trim releases the idle ones. */
LCC_API lcc_status lcc_host_cache_alloc(void **ptr, size_t bytes);
LCC_API lcc_status lcc_host_cache_free(void *ptr);
LCC_API lcc_status lcc_host_cache_trim(void);
LCC_API lcc_status lcc_host_cache_stats(size_t *live_bytes, size_t *idle_bytes);
LCC_API int lcc_host_is_pinned(const void *ptr, size_t bytes); /* 1 when the DMA engines can address the range directly */

/* ---------------------------------------------------------------- host-buffer (end-to-end) entry points */

/* Same operations with HOST buffers in the reference's (n, num_blades) row-major layout: the
 * library chunks the batch, overlaps H2D copy / kernel / D2H copy on three streams and returns
 * when `out` is complete.  Pinned buffers (lcc_host_alloc_pinned) run at PCIe speed in both directions at once;
 * pageable ones are staged through the same pinned bounce rings as lcc_batch_upload / download.  Unlike the
 * three-call sequence from_numpy -> op -> to_numpy, whose upload and download cannot overlap (each call must
 * finish before the next starts), these keep both PCIe directions busy at the same time. */
LCC_API lcc_status lcc_host_sandwich(const lcc_algebra *alg, int dtype, const double *rotor, const void *x_host,
                                     void *out_host, int64_t n, unsigned flags);
LCC_API lcc_status lcc_host_product(const lcc_algebra *alg, int dtype, int op, const void *a_host, int64_t na,
                                    const void *b_host, int64_t nb_rows, void *out_host, unsigned flags);
/* lcc_pga_points_sandwich with HOST (n,3) arrays: 2 x 3 elements per point cross PCIe instead of the 2 x 16 of the
 * dense path (the reference has no batched form: a Python loop over Multivector.pga_point / sandwich /
 * pga_point_coords, src/pga.rs:110-135,579-599).  xyz_host_out may equal xyz_host_in. */
LCC_API lcc_status lcc_host_pga_points_sandwich(const lcc_algebra *alg, int dtype, const double *motor, const void *xyz_host_in,
                                                void *xyz_host_out, int64_t n, unsigned flags);

/* ---------------------------------------------------------------- multi-GPU: row shards + the NCCL sum / mean */

/* Every batched operation is row-independent (`for row in 0..n`, src/batch.rs:384), so N rows are split into
 * contiguous blocks, shard k of W owning rows [k*N/W, (k+1)*N/W), each resident on its own GPU and driven on its own
 * non-blocking stream; one-row (broadcast) operands and rotors are replicated.  No data-path collective exists.  The
 * only exchange is the batch sum / mean (new surface): per-GPU fused partial of num_blades values, then
 * ncclAllReduce(count = num_blades, ncclSum) enqueued right behind it on the same stream, no host synchronisation in
 * between; the mean's division runs as a one-block kernel after the all-reduce.
 *
 * A group is either all GPUs of ONE process (lcc_shard_group_create: ncclCommInitAll, one host thread issues the
 * asynchronous launches of every GPU) or one member of a multi-process job with one rank per GPU
 * (lcc_shard_group_create_rank: ncclCommInitRank; rank 0 makes the 128-byte id with lcc_shard_unique_id and the
 * launcher distributes it).  Functions taking `views` expect one lcc_view per LOCAL device, in local order.
 * NCCL is loaded with dlopen on first use; one-device groups never need it. */
typedef struct lcc_shard_group lcc_shard_group;

LCC_API lcc_status lcc_shard_rows(int64_t n_total, int rank, int world, int64_t *begin, int64_t *end); /* host only */
LCC_API lcc_status lcc_shard_unique_id(void *id128);
LCC_API lcc_status lcc_shard_group_create(int ndev, const int *devices /* NULL: 0..ndev-1; ndev <= 0: all */, lcc_shard_group **out);
LCC_API lcc_status lcc_shard_group_create_rank(int device, int world, int rank, const void *id128, lcc_shard_group **out);
LCC_API void lcc_shard_group_destroy(lcc_shard_group *group);
LCC_API int lcc_shard_group_local_size(const lcc_shard_group *group);
LCC_API int lcc_shard_group_world_size(const lcc_shard_group *group);
LCC_API int lcc_shard_group_first_rank(const lcc_shard_group *group);
LCC_API int lcc_shard_group_device(const lcc_shard_group *group, int local_index);
LCC_API int lcc_shard_nccl_version(void); /* 0 when NCCL could not be loaded */
LCC_API lcc_status lcc_shard_group_stream(const lcc_shard_group *group, int local_index, void **stream);
LCC_API lcc_status lcc_shard_group_sync(lcc_shard_group *group);
/* One engine-native batch per local device holding that shard's rows of an n_total-row batch. */
LCC_API lcc_status lcc_shard_batch_alloc(lcc_shard_group *group, const lcc_algebra *alg, int dtype, int64_t n_total,
                                         lcc_view *views, int64_t *row_begin /* may be NULL */);
/* A one-row (broadcast) operand replicated on every local device from a host row of num_blades elements. */
LCC_API lcc_status lcc_shard_broadcast_alloc(lcc_shard_group *group, const lcc_algebra *alg, int dtype, const void *host_row, lcc_view *views);
LCC_API lcc_status lcc_shard_batch_free(lcc_shard_group *group, lcc_view *views);
/* Host (n_total, num_blades) array <-> the local shards (each GPU runs lcc_batch_upload / download on its own rows,
 * all local GPUs concurrently). */
LCC_API lcc_status lcc_shard_upload(lcc_shard_group *group, const lcc_algebra *alg, const void *host, int64_t n_total, const lcc_view *views);
LCC_API lcc_status lcc_shard_download(lcc_shard_group *group, const lcc_algebra *alg, const lcc_view *views, void *host, int64_t n_total);
/* Row-wise operations: one asynchronous launch per local GPU on its own stream (lcc_batch_product / _sandwich / _unary). */
LCC_API lcc_status lcc_shard_product(lcc_shard_group *group, const lcc_algebra *alg, int op, const lcc_view *a, const lcc_view *b,
                                     const lcc_view *out, unsigned flags);
LCC_API lcc_status lcc_shard_sandwich(lcc_shard_group *group, const lcc_algebra *alg, const double *rotor, const lcc_view *x,
                                      const lcc_view *out, unsigned flags);
LCC_API lcc_status lcc_shard_unary(lcc_shard_group *group, const lcc_algebra *alg, int op, int arg, const lcc_view *x, const lcc_view *out);
/* The collective: result_host[k] = sum over ALL rows of ALL shards of x[.][k] (num_blades elements of the batch dtype;
 * may be NULL), divided by mean_divisor when mean_divisor > 0 (pass the global row count for the mean).  Every rank /
 * local device ends up with the same values.  lcc_shard_product_sum is the fused config-4 form (lcc_batch_product_sum
 * per GPU, then the all-reduce). */
LCC_API lcc_status lcc_shard_sum(lcc_shard_group *group, const lcc_algebra *alg, const lcc_view *x, int64_t mean_divisor, void *result_host);
LCC_API lcc_status lcc_shard_product_sum(lcc_shard_group *group, const lcc_algebra *alg, int op, const lcc_view *a, int a_grade,
                                         const lcc_view *b, int64_t mean_divisor, void *result_host, unsigned flags);
/* Pipelined form: the partial runs on the compute stream and the all-reduce of `slot` (0..7) on a high-priority
 * communication stream behind an event, so the collective of step i overlaps the kernel of step i+1;
 * lcc_shard_reduce_fetch waits for that slot and copies its result to the host. */
LCC_API lcc_status lcc_shard_product_sum_async(lcc_shard_group *group, const lcc_algebra *alg, int op, const lcc_view *a, int a_grade,
                                               const lcc_view *b, int64_t mean_divisor, int slot, unsigned flags);
LCC_API lcc_status lcc_shard_reduce_fetch(lcc_shard_group *group, const lcc_algebra *alg, int dtype, int slot, void *result_host);

/* ---------------------------------------------------------------- tuning */

/* Process-wide integer options (no reference counterpart):
 *   "aos_tma_min_rows"   AoS batches with at least this many rows use the TMA-staged kernels (default 4096; rows
 *                        narrower than 128 bytes need 32x as many)
 *   "soa_tma_min_bytes"  SoA batches with at least this many bytes per operand do (default 8 GiB)
 *   "matrix_rep"         fixed-operand products / sandwiches of 128- and 256-blade algebras in the 16 x 16 matrix
 *                        representation (K8) where one exists: 2 (default) for f64 and f32 batches, 1 for f64 batches
 *                        only, 0 never (always the nb x nb tensor-core GEMM, K6 / K6d)
 *   "host_chunk_bytes"   chunk size of the host <-> device pipelines (default 64 MiB)
 *   "pinned_output_min_bytes"  to_numpy() results of at least this size come from the pinned host cache (default 1 MiB;
 *                        a very large value switches the cache off)
 * Results do not depend on either: the staged and the direct-load kernels run the same arithmetic. */
LCC_API lcc_status lcc_set_option(const char *name, int64_t value);
LCC_API lcc_status lcc_get_option(const char *name, int64_t *value);

/* ---------------------------------------------------------------- instrumentation */

/* Number of kernels this library has launched on this thread's behalf since the last reset
 * (bench.py reports it as gpu_launches). */
LCC_API uint64_t lcc_kernel_launch_count(void);
LCC_API void lcc_kernel_launch_count_reset(void);

#ifdef __cplusplus
}
#endif
#endif /* LCC_B200_H */
