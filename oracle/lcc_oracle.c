/*
 * lcc_oracle.c -- CPU restatement of the reference's batched Clifford-algebra path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under largecrimsoncanine_b200/, lcc/ or the
 * `largecrimsoncanine` extension may import, link or execute this file.  The only
 * permitted callers are tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` legs, and there only as the checker or the timed CPU baseline.
 *
 * What it restates (file:line relative to /root/reference):
 *   src/algebra/signature.rs:173-182   basis_square
 *   src/algebra/cayley.rs:63-144       blade_product_signed, compute_reorder_sign, compute_cayley_table
 *   src/algebra/blades.rs:24-26,128-188,211-229   grade, involution signs, grade masks
 *   src/batch.rs:112-261,353-1052      the batched loops (array-of-structures f64, i outer / j inner,
 *                                      zero skips, separate multiply and add -- never fused)
 *   src/simd.rs:55-135                 the 4-wide single-pair product
 *   src/multivector.rs:4249-4290,2989-2999   dual / undual / right_dual of one multivector
 *   src/pga.rs:521-556                 pga_dual
 *
 * Rounding contract: Rust never contracts a*b+c into an FMA, so this file must be
 * compiled with  -O3 -ffp-contract=off  and without -ffast-math (see oracle/build.py).
 *
 * Parity pinning: the Rust crate cannot be built here (no cargo/rustc, no vendored
 * crates), so there is no oracle/_ref.  This restatement is pinned against
 *   (1) every table KAT the reference's own unit tests hold (tests/test_oracle_tables.py),
 *   (2) golden vectors produced by importing the reference's own Python implementation
 *       /root/reference/lcc/torch/multivector.py (tests/golden/make_golden.py), and
 *   (3) the closed-form KATs of /root/reference/tests/test_batch.py, test_pga.py, ...
 * The f32 variants and batch sum/mean have no counterpart in the Rust crate (f64 only);
 * for those the oracle is "same loops in float" / a long-double column sum, and DESIGN.md
 * says "parity unpinned" for them.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORACLE_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------ L1: integer primitives */

/* signature.rs:173-182 -- first p vectors square to +1, next q to -1, last r to 0 */
ORACLE_API double oracle_basis_square(int p, int q, int r, int i) {
    (void)r;
    if (i < p) return 1.0;
    if (i < p + q) return -1.0;
    return 0.0;
}

/* cayley.rs:89-111 -- for every bit of b (low to high) count the bits of a strictly above it */
ORACLE_API double oracle_reorder_sign(uint64_t a, uint64_t b) {
    double sign = 1.0;
    uint64_t rest = b;
    while (rest != 0) {
        uint64_t low = rest & (~rest + 1);
        int pos = __builtin_ctzll(low);
        uint64_t above = a >> (pos + 1);
        if (__builtin_popcountll(above) % 2 == 1) sign = -sign;
        rest &= ~low;
    }
    return sign;
}

/* cayley.rs:63-80 -- blade = a^b, sign = reorder * product of metric over the shared vectors */
ORACLE_API void oracle_blade_product(uint64_t a, uint64_t b, int p, int q, int r,
                                     uint64_t *blade, double *sign) {
    double s = oracle_reorder_sign(a, b);
    uint64_t common = a & b;
    int n = p + q + r;
    for (int i = 0; i < n; ++i)
        if ((common >> i) & 1) s *= oracle_basis_square(p, q, r, i);
    *blade = a ^ b;
    *sign = s;
}

/* cayley.rs:123-144 -- flat tables [a*nb+b]; usize -> uint64_t, f64 -> double */
ORACLE_API void oracle_cayley_table(int p, int q, int r, uint64_t *blades, double *signs) {
    uint64_t nb = 1ull << (p + q + r);
    for (uint64_t a = 0; a < nb; ++a)
        for (uint64_t b = 0; b < nb; ++b)
            oracle_blade_product(a, b, p, q, r, &blades[a * nb + b], &signs[a * nb + b]);
}

/* blades.rs:24-26 */
ORACLE_API int oracle_blade_grade(uint64_t index) { return __builtin_popcountll(index); }
/* blades.rs:128-135 */
ORACLE_API double oracle_reverse_sign(int grade) { return (grade % 4 < 2) ? 1.0 : -1.0; }
/* blades.rs:154-160 */
ORACLE_API double oracle_grade_involution_sign(int grade) { return (grade % 2 == 0) ? 1.0 : -1.0; }
/* blades.rs:182-188 */
ORACLE_API double oracle_clifford_conjugate_sign(int grade) {
    int m = grade % 4;
    return (m == 0 || m == 3) ? 1.0 : -1.0;
}
/* blades.rs:211-220 -- one machine word per grade; defined only while 2^dimension <= 64
 * (dimension <= 6): the Rust `1 << i` overflows beyond that.  Returns 0 and sets *ok=0 there. */
ORACLE_API uint64_t oracle_grade_mask(int dimension, int grade, int *ok) {
    if (ok) *ok = dimension <= 6;
    if (dimension > 6) return 0;
    uint64_t nb = 1ull << dimension, mask = 0;
    for (uint64_t i = 0; i < nb; ++i)
        if (oracle_blade_grade(i) == grade) mask |= 1ull << i;
    return mask;
}

/* ------------------------------------------------------------------ algebra handle */

typedef struct {
    int p, q, r, dim;
    int64_t nb;
    uint64_t *blades; /* nb*nb */
    double *signs;    /* nb*nb */
} oracle_algebra;

ORACLE_API oracle_algebra *oracle_algebra_new(int p, int q, int r) {
    oracle_algebra *alg = (oracle_algebra *)calloc(1, sizeof(*alg));
    alg->p = p; alg->q = q; alg->r = r; alg->dim = p + q + r;
    alg->nb = (int64_t)1 << alg->dim;
    alg->blades = (uint64_t *)malloc(sizeof(uint64_t) * alg->nb * alg->nb);
    alg->signs = (double *)malloc(sizeof(double) * alg->nb * alg->nb);
    oracle_cayley_table(p, q, r, alg->blades, alg->signs);
    return alg;
}
ORACLE_API void oracle_algebra_free(oracle_algebra *alg) {
    if (!alg) return;
    free(alg->blades); free(alg->signs); free(alg);
}
ORACLE_API int64_t oracle_algebra_num_blades(const oracle_algebra *alg) { return alg->nb; }
ORACLE_API const uint64_t *oracle_algebra_blades(const oracle_algebra *alg) { return alg->blades; }
ORACLE_API const double *oracle_algebra_signs(const oracle_algebra *alg) { return alg->signs; }

/* ------------------------------------------------------------------ L3: batched loops (two precisions) */

#define REAL double
#define SUFFIX(name) name##_f64
#include "lcc_oracle_batch.inc"
#undef REAL
#undef SUFFIX

#define REAL float
#define SUFFIX(name) name##_f32
#include "lcc_oracle_batch.inc"
#undef REAL
#undef SUFFIX

/* ------------------------------------------------------------------ scalar twins used as semantic spec */

/* simd.rs:55-135 -- one pair, 4-wide chunks: (a*b)*sign, zero skip on a only (chunks never skip b);
 * nb is a power of two >= 4 so the scalar remainder loop of the reference never runs. */
ORACLE_API void oracle_simd_geometric_product(const oracle_algebra *alg, const double *a,
                                              const double *b, double *result) {
    int64_t nb = alg->nb;
    for (int64_t k = 0; k < nb; ++k) result[k] = 0.0;
    for (int64_t i = 0; i < nb; ++i) {
        double ai = a[i];
        if (ai == 0.0) continue;
        int64_t row = i * nb;
        int64_t chunks = nb / 4, rem = nb % 4;
        for (int64_t c = 0; c < chunks; ++c) {
            int64_t j0 = c * 4;
            double prod[4];
            for (int l = 0; l < 4; ++l) prod[l] = (ai * b[j0 + l]) * alg->signs[row + j0 + l];
            for (int l = 0; l < 4; ++l) result[alg->blades[row + j0 + l]] += prod[l];
        }
        for (int64_t j = chunks * 4; j < chunks * 4 + rem; ++j) {
            double bj = b[j];
            if (bj == 0.0) continue;
            result[alg->blades[row + j]] += alg->signs[row + j] * ai * bj;
        }
    }
}

/* multivector.rs:2890-2913 -- scalar part of A*~A for one multivector */
static double scalar_norm_squared(const oracle_algebra *alg, const double *a) {
    int64_t nb = alg->nb;
    double acc = 0.0;
    for (int64_t i = 0; i < nb; ++i) {
        if (a[i] == 0.0) continue;
        for (int64_t j = 0; j < nb; ++j) {
            double bj = a[j] * oracle_reverse_sign(oracle_blade_grade((uint64_t)j));
            if (bj == 0.0) continue;
            if (alg->blades[i * nb + j] == 0) acc += alg->signs[i * nb + j] * a[i] * bj;
        }
    }
    return acc;
}

/* multivector.rs:2989-2999 -- inverse = reverse scaled by 1/norm_squared; returns 0 when not invertible */
static int scalar_inverse(const oracle_algebra *alg, const double *a, double *inv) {
    double n2 = scalar_norm_squared(alg, a);
    if (n2 == 0.0) return 0;
    double f = 1.0 / n2;
    for (int64_t i = 0; i < alg->nb; ++i)
        inv[i] = (a[i] * oracle_reverse_sign(oracle_blade_grade((uint64_t)i))) * f;
    return 1;
}

/* multivector.rs:4249-4290 -- which: 0 dual = A*I^-1, 1 undual = A*I, 2 right_dual = I^-1*A.
 * Applied row by row to an (n, nb) array so tests can compare whole batches.
 * Returns 0 if the pseudoscalar is not invertible (ValueError in the reference). */
ORACLE_API int oracle_dual_rows(const oracle_algebra *alg, int which, const double *x, int64_t n,
                                double *out) {
    int64_t nb = alg->nb;
    double *ps = (double *)calloc(nb, sizeof(double));
    double *op = (double *)calloc(nb, sizeof(double));
    ps[nb - 1] = 1.0;
    if (which == 1) {
        memcpy(op, ps, sizeof(double) * nb);
    } else if (!scalar_inverse(alg, ps, op)) {
        free(ps); free(op);
        return 0;
    }
    for (int64_t row = 0; row < n; ++row) {
        const double *a = which == 2 ? op : x + row * nb;
        const double *b = which == 2 ? x + row * nb : op;
        double *res = out + row * nb;
        for (int64_t k = 0; k < nb; ++k) res[k] = 0.0;
        /* Multivector::geometric_product: scalar loop below 16 blades, simd path from 16 blades on
         * (multivector.rs:1451-1473); both give the same values for finite inputs. */
        if (nb >= 16) {
            oracle_simd_geometric_product(alg, a, b, res);
        } else {
            for (int64_t i = 0; i < nb; ++i) {
                if (a[i] == 0.0) continue;
                for (int64_t j = 0; j < nb; ++j) {
                    if (b[j] == 0.0) continue;
                    res[alg->blades[i * nb + j]] += alg->signs[i * nb + j] * a[i] * b[j];
                }
            }
        }
    }
    free(ps); free(op);
    return 1;
}

/* pga.rs:521-556 -- complement dual of PGA3D, sign = (-1)^(g(g-1)/2) with the usize wrap at g=0 -> +1 */
ORACLE_API void oracle_pga_dual_rows(const double *x, int64_t n, double *out) {
    for (int64_t row = 0; row < n; ++row) {
        const double *a = x + row * 16;
        double *res = out + row * 16;
        for (int k = 0; k < 16; ++k) res[k] = 0.0;
        for (int i = 0; i < 16; ++i) {
            double c = a[i];
            if (c != 0.0) {
                uint64_t g = (uint64_t)oracle_blade_grade((uint64_t)i);
                uint64_t e = (g * (g - 1)) / 2; /* wraps like release-mode usize */
                double sign = (e % 2 == 0) ? 1.0 : -1.0;
                res[15 ^ i] += sign * c;
            }
        }
    }
}

/* Batch sum / mean are NEW surface (no reference counterpart): long-double column sums. */
ORACLE_API void oracle_sum_rows_f64(const double *x, int64_t n, int64_t nb, double *out, double *abs_out) {
    for (int64_t k = 0; k < nb; ++k) {
        long double s = 0.0L, sa = 0.0L;
        for (int64_t row = 0; row < n; ++row) { s += x[row * nb + k]; sa += fabsl((long double)x[row * nb + k]); }
        out[k] = (double)s;
        if (abs_out) abs_out[k] = (double)sa;
    }
}
ORACLE_API void oracle_sum_rows_f32(const float *x, int64_t n, int64_t nb, double *out, double *abs_out) {
    for (int64_t k = 0; k < nb; ++k) {
        long double s = 0.0L, sa = 0.0L;
        for (int64_t row = 0; row < n; ++row) { s += x[row * nb + k]; sa += fabsl((long double)x[row * nb + k]); }
        out[k] = (double)s;
        if (abs_out) abs_out[k] = (double)sa;
    }
}

ORACLE_API int oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
