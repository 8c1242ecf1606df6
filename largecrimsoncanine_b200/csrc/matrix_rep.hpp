// matrix_rep.hpp -- a faithful real MATRIX representation of a non-degenerate Clifford algebra Cl(p,q), found on the
// host once per algebra, that block-diagonalises the multiplication matrix of a fixed operand.
//
// North star, config 5: "products of a large batch against a fixed operand in high-dimensional algebras, cast as a GEMM
// with that operand's left-multiplication matrix" -- L_M is nb x nb (2*nb^2 flops per row).  For Cl(8,0) ~ M_16(R) (and
// every signature of dimension 7 or 8 that fits real 16 x 16 matrices) there is a basis in which L_M = I_16 (x) rho(M):
// sixteen copies of one 16 x 16 block.  In that basis the product costs 2*16^3 = 8192 flops per row plus two signed
// Walsh-Hadamard transforms of 256 values -- 16x fewer flops than the dense GEMM, which makes the product HBM-bound.
//
// Construction: real Pauli strings on m = 4 "qubits".  P(a, b) = X^a Z^b (a, b in F_2^4; X, Z the real Pauli matrices,
// Kronecker-wise) is the signed permutation matrix  P(a,b)[i][j] = [i == j ^ a] * (-1)^popcount(b & j), with
//     P(a,b) P(a',b') = (-1)^popcount(b & a') P(a^a', b^b'),     P(a,b)^2 = (-1)^popcount(a & b) I.
// The 256 strings are a basis of M_16(R).  A depth-first search picks n strings g_1..g_n that mutually anticommute, square
// to the metric (+1 for the first p, -1 for the next q) and are independent over F_2; every blade e_A (ascending product
// of generators, /root/reference/src/algebra/cayley.rs:63-111) is then  sign_A * P(string_A), distinct blades get
// distinct strings, and the map is checked against the algebra's own Cayley sign table for ALL blade pairs before it is
// used (verify()).  With y[a][b] = sign_A x_A (string_A = (a,b)):
//     rho(x)[j^a][j] = sum_b (-1)^popcount(b & j) y[a][b]                         (16 Walsh-Hadamard transforms of 16)
//     x_A = sign_A / 16 * sum_j (-1)^popcount(b & j) rho(x)[j^a][j]              (the inverse: the same transforms)
#pragma once
#include <cstdint>
#include <vector>

#include "algebra_host.hpp"

namespace lcc {

struct MatrixRep {
    bool ok = false;
    int dim = 0;                 // matrix size d (16)
    uint8_t string_of[256];      // blade -> (a << 4) | b
    int8_t sign_of[256];         // blade -> +1 / -1
    int16_t blade_at[256];       // string -> blade, -1 when no blade maps there (algebras with fewer than d^2 blades)
};

namespace matrix_rep_detail {
inline int par(unsigned v) { return __builtin_popcount(v) & 1; }
struct PS { unsigned a, b; int sign; };
inline PS mul(const PS &x, const PS &y) { return PS{x.a ^ y.a, x.b ^ y.b, x.sign * y.sign * (par(x.b & y.a) ? -1 : 1)}; }
inline bool anticommute(unsigned s, unsigned t) { return par(((s & 15) & (t >> 4)) ^ ((t & 15) & (s >> 4))) != 0; }  // s = (a << 4) | b
inline int square(unsigned s) { return par((s >> 4) & (s & 15)) ? -1 : 1; }

// rank test over F_2: is `s` independent of the span of basis (kept in echelon form by leading bit)?
inline bool add_independent(std::vector<unsigned> &echelon, unsigned s) {
    for (unsigned e : echelon) {
        const unsigned lead = 1u << (31 - __builtin_clz(e));
        if (s & lead) s ^= e;
    }
    if (!s) return false;
    echelon.push_back(s);
    return true;
}

inline bool search(int p, int q, int k, std::vector<unsigned> &gens, std::vector<unsigned> &echelon) {
    const int n = p + q;
    if (k == n) return true;
    const int want = k < p ? 1 : -1;
    for (unsigned s = 1; s < 256; ++s) {
        if (square(s) != want) continue;
        bool good = true;
        for (unsigned g : gens) good = good && anticommute(g, s);
        if (!good) continue;
        std::vector<unsigned> saved = echelon;
        if (!add_independent(echelon, s)) { echelon = saved; continue; }
        gens.push_back(s);
        if (search(p, q, k + 1, gens, echelon)) return true;
        gens.pop_back();
        echelon = saved;
    }
    return false;
}
}  // namespace matrix_rep_detail

// Representation of `alg` by real 16 x 16 signed permutation matrices, or ok == false when none exists in that size
// (degenerate metrics, fewer than 7 dimensions -- the dense GEMM is already HBM-bound there -- and the signatures whose
// algebra is quaternionic or a direct sum, e.g. Cl(7,1) ~ M_8(H)).
//
// `variant` > 0 relabels the strings of the generators by a real Clifford automorphism of the Pauli group (a qubit
// permutation, X <-> Z swaps on a subset of the qubits, optionally a CNOT): commutation relations and squares are
// preserved, so the images still generate a faithful representation -- equivalent as a representation, but with a different
// string -> blade map.  K8 needs one whose map gives it conflict-free MMA fragment accesses (make_rep_layout below).
inline unsigned relabel_string(unsigned s, int variant) {
    if (variant <= 0) return s;
    const int v = variant - 1;
    const int perm_id = v % 24, hmask = (v / 24) % 16, cnot = (v / 384) % 13;  // cnot: 0 none, else control/target pair
    static const int perms[24][4] = {{0,1,2,3},{0,1,3,2},{0,2,1,3},{0,2,3,1},{0,3,1,2},{0,3,2,1},{1,0,2,3},{1,0,3,2},{1,2,0,3},{1,2,3,0},{1,3,0,2},{1,3,2,0},
                                     {2,0,1,3},{2,0,3,1},{2,1,0,3},{2,1,3,0},{2,3,0,1},{2,3,1,0},{3,0,1,2},{3,0,2,1},{3,1,0,2},{3,1,2,0},{3,2,0,1},{3,2,1,0}};
    unsigned a = s >> 4, b = s & 15, na = 0, nb = 0;
    for (int qb = 0; qb < 4; ++qb) {
        unsigned aq = (a >> qb) & 1, bq = (b >> qb) & 1;
        if ((hmask >> qb) & 1) { const unsigned t = aq; aq = bq; bq = t; }
        na |= aq << perms[perm_id][qb];
        nb |= bq << perms[perm_id][qb];
    }
    if (cnot) {
        const int c = (cnot - 1) / 3, t0 = (cnot - 1) % 3, t = t0 >= c ? t0 + 1 : t0;  // 12 ordered pairs
        na ^= ((na >> c) & 1) << t;   // a_t ^= a_c
        nb ^= ((nb >> t) & 1) << c;   // b_c ^= b_t
    }
    return (na << 4) | nb;
}

// `base_gens` (optional): generator strings of an earlier variant-0 search, so that relabelled variants skip the search;
// `verify` = false skips the all-pairs check (the caller verifies the variant it finally installs).
inline MatrixRep find_matrix_rep(const lcc_algebra &alg, int variant = 0, std::vector<unsigned> *base_gens = nullptr, bool verify = true) {
    using namespace matrix_rep_detail;
    MatrixRep rep;
    for (int i = 0; i < 256; ++i) { rep.string_of[i] = 0; rep.sign_of[i] = 0; rep.blade_at[i] = -1; }
    if (alg.r != 0 || alg.dim < 7 || alg.dim > 8) return rep;
    // Classification of real Clifford algebras by (p - q) mod 8: real 16 x 16 matrices carry a faithful representation of
    // Cl(p,q) exactly for  n = 8: 0, 2 (M_16(R))  and  n = 7: 1 (M_8(R)+M_8(R)), 3, 7 (M_8(C)); the other classes are
    // quaternionic and need 32 x 32.  Gate on it: an exhaustive search that must FAIL takes minutes.
    const int cls = (((alg.p - alg.q) % 8) + 8) % 8;
    if (alg.dim == 8 ? !(cls == 0 || cls == 2) : !(cls == 1 || cls == 3 || cls == 7)) return rep;
    std::vector<unsigned> gens, echelon;
    if (base_gens && !base_gens->empty()) gens = *base_gens;
    else {
        if (!search(alg.p, alg.q, 0, gens, echelon)) return rep;
        if (base_gens) *base_gens = gens;
    }
    for (unsigned &g : gens) g = relabel_string(g, variant);
    const int nb = alg.nb;
    for (int A = 0; A < nb; ++A) {
        PS acc{0, 0, 1};
        for (int i = 0; i < alg.dim; ++i)
            if ((A >> i) & 1) acc = mul(acc, PS{gens[i] >> 4, gens[i] & 15, 1});  // ascending product of generators
        const unsigned s = (acc.a << 4) | acc.b;
        if (rep.blade_at[s] >= 0) return rep;  // two blades on one string: not faithful
        rep.string_of[A] = (uint8_t)s;
        rep.sign_of[A] = (int8_t)acc.sign;
        rep.blade_at[s] = (int16_t)A;
    }
    // verify against the algebra's own sign table for every pair of blades: rho(e_A) rho(e_B) == sign[A][B] rho(e_{A^B})
    for (int A = 0; verify && A < nb; ++A)
        for (int B = 0; B < nb; ++B) {
            const PS x{(unsigned)rep.string_of[A] >> 4, (unsigned)rep.string_of[A] & 15, rep.sign_of[A]};
            const PS y{(unsigned)rep.string_of[B] >> 4, (unsigned)rep.string_of[B] & 15, rep.sign_of[B]};
            const PS z = mul(x, y);
            const int C = A ^ B;
            if (((z.a << 4) | z.b) != rep.string_of[C] || z.sign != alg.signs[(size_t)A * nb + B] * rep.sign_of[C]) return rep;
        }
    rep.dim = 16;
    rep.ok = true;
    return rep;
}

// rho(M) as a row-major d x d matrix (f64) from the coefficients of M:  rho(M)[j^a][j] += sign_A M_A (-1)^popcount(b & j)
inline void rep_matrix(const MatrixRep &rep, int nb, const double *coeffs, double *out /* 256 */) {
    for (int i = 0; i < 256; ++i) out[i] = 0.0;
    for (int A = 0; A < nb; ++A) {
        if (coeffs[A] == 0.0) continue;
        const unsigned a = rep.string_of[A] >> 4, b = rep.string_of[A] & 15;
        for (unsigned j = 0; j < 16; ++j) {
            const double v = coeffs[A] * rep.sign_of[A];
            out[(j ^ a) * 16 + j] += matrix_rep_detail::par(b & j) ? -v : v;
        }
    }
}


// ---------------------------------------------------------------------------------------------------------------------
// Shared-memory layout of K8 (kernels/matrix_rep_kernel.cu), computed on the host once per launch.
//
// The string -> blade map is LINEAR over F_2 (the string of a product of blades is the XOR of their strings), so
// blade(a << 4 | b) = lin_hi(a) ^ lin_lo(b).  A 32 KB unit of the kernel holds 256 slots of 128 bytes (slot = blade, as the
// TMA engine delivers them, hardware 128-byte swizzle: 16-byte chunk c of slot s sits at chunk c ^ (s & 7)).  Matrix entry
// (f, s) of rho(x) lives IN the slot of the coefficient it is built from first, blade(f ^ s, s); then
//   * the Walsh-Hadamard transforms (string row a fixed) touch the 16 slots blade(a, .) -- the same 16 on the way in and
//     on the way out, so a thread transforms its rows in place without any barrier;
//   * a product with a fixed factor from the LEFT contracts the first index at fixed second index w (column pass): entry
//     (k, w) sits in slot Lambda(k) ^ Gamma(w) with Lambda(k) = lin_hi(k), Gamma(w) = lin_hi(w) ^ lin_lo(w);
//   * a product from the RIGHT contracts the second index at fixed first index w (row pass): entry (w, k) sits in slot
//     Lambda(k) ^ Gamma(w) with Lambda(k) = lin_hi(k) ^ lin_lo(k), Gamma(w) = lin_hi(w).
// Byte offset of slot s with the swizzle pre-applied: F(s) = s << 7 | (s & 7) << 4 (linear too); the byte offset of row
// bytes r inside the slot is F(s) ^ r.  An MMA B-fragment load reads four slots (the k-lanes of a quad) at eight rows:
// conflict free iff those four slots differ in blade bits 1-2, so the contraction index is laid onto the MMA's physical k
// (and m) positions through a basis v0..v3 of F_2^4 whose first two vectors have independent (Lambda >> 1) & 3 and whose
// first has blade bit 2 set (f64 D-fragment stores); pick_basis() finds one or reports that this representation has none.
struct RepPass {
    uint32_t fk[16];   // F(Lambda(kmap[physical index]))
    uint32_t cw[16];   // F(Gamma(w))
    uint8_t kmap[16];  // physical MMA index (k or m) -> logical matrix index
};
struct RepLayout {
    bool ok = false;
    bool lin_ok = false;  // lin is filled (the passes may still have failed)
    uint8_t lin[256];     // string -> blade, extended linearly to strings no blade maps to (128-blade algebras: blades 128..255)
    uint32_t f_lo[16];    // F(lin_lo(b))
    uint32_t f_hi[16];    // F(lin_hi(a))
    RepPass column, row;
};

namespace matrix_rep_detail {
inline uint32_t slot_offset(unsigned s) { return (s << 7) | ((s & 7u) << 4); }
inline bool pick_basis(const unsigned lam[16], unsigned v[4]) {
    auto phi = [&](unsigned x) { return (lam[x] >> 1) & 3u; };
    for (unsigned v0 = 1; v0 < 16; ++v0) {
        if (!((lam[v0] >> 2) & 1u)) continue;
        for (unsigned v1 = 1; v1 < 16; ++v1) {
            if (v1 == v0 || phi(v1) == 0 || phi(v1) == phi(v0)) continue;
            for (unsigned v2 = 1; v2 < 16; ++v2) {
                if (v2 == v0 || v2 == v1 || v2 == (v0 ^ v1)) continue;
                for (unsigned v3 = 1; v3 < 16; ++v3) {
                    bool in_span = false;
                    for (unsigned m = 0; m < 8; ++m)
                        if (v3 == (((m & 1) ? v0 : 0) ^ ((m & 2) ? v1 : 0) ^ ((m & 4) ? v2 : 0))) in_span = true;
                    if (in_span) continue;
                    v[0] = v0; v[1] = v1; v[2] = v2; v[3] = v3;
                    return true;
                }
            }
        }
    }
    return false;
}
inline bool make_pass(const unsigned lam[16], const unsigned gam[16], RepPass &p) {
    unsigned v[4];
    if (!pick_basis(lam, v)) return false;
    for (unsigned i = 0; i < 16; ++i) {
        unsigned k = 0;
        for (int bit = 0; bit < 4; ++bit) if ((i >> bit) & 1) k ^= v[bit];
        p.kmap[i] = (uint8_t)k;
        p.fk[i] = slot_offset(lam[k]);
        p.cw[i] = slot_offset(gam[i]);
    }
    return true;
}
}  // namespace matrix_rep_detail

// blade_at / sign_at: string -> blade / sign (0 where no blade maps to the string)
inline RepLayout make_rep_layout(const uint8_t *blade_at, const int8_t *sign_at) {
    using namespace matrix_rep_detail;
    RepLayout L;
    // linear extension of string -> blade to all 256 strings: close the (string, blade) pairs of the basis blades under
    // XOR; a 128-blade algebra spans only half the strings, the other half gets blades 128..255 (slots that exist in a
    // unit but hold no data: such strings are masked out by sign == 0)
    int lin[256], have[256];
    for (int i = 0; i < 256; ++i) { lin[i] = 0; have[i] = 0; }
    have[0] = 1;
    int count = 1, next_blade = 1;
    auto add = [&](int str, int blade) {
        if (have[str]) return;
        for (int t = 0; t < 256; ++t)
            if (have[t] == 1) { lin[t ^ str] = lin[t] ^ blade; have[t ^ str] = 2; }
        for (int t = 0; t < 256; ++t) if (have[t] == 2) have[t] = 1;
        count *= 2;
    };
    for (int str = 1; str < 256; ++str)
        if (sign_at[str] != 0 && (blade_at[str] & (blade_at[str] - 1)) == 0) { add(str, blade_at[str]); if (blade_at[str] >= next_blade) next_blade = blade_at[str] << 1; }
    for (int str = 1; str < 256 && count < 256; ++str)
        if (!have[str]) { add(str, next_blade); next_blade <<= 1; }
    for (int str = 0; str < 256; ++str) {
        if (sign_at[str] != 0 && lin[str] != blade_at[str]) return L;  // cannot happen for a homomorphism
        if (lin[str] < 0 || lin[str] > 255) return L;
        L.lin[str] = (uint8_t)lin[str];
    }
    L.lin_ok = true;
    unsigned lam_c[16], gam_c[16], lam_r[16], gam_r[16];
    for (unsigned i = 0; i < 16; ++i) {
        L.f_lo[i] = slot_offset(L.lin[i]);
        L.f_hi[i] = slot_offset(L.lin[i << 4]);
        lam_c[i] = L.lin[i << 4];            gam_c[i] = L.lin[(i << 4) | i];
        lam_r[i] = L.lin[(i << 4) | i];      gam_r[i] = L.lin[i << 4];
    }
    L.ok = make_pass(lam_c, gam_c, L.column) && make_pass(lam_r, gam_r, L.row);
    return L;
}

// ---------------------------------------------------------------------------------------------------------------------
// K8 on reference-layout (AoS) batches: a unit is 128 / sizeof(T) consecutive ROWS (32 f32, 16 f64) of the (n, nb) array --
// one contiguous 16-32 KB span of HBM -- delivered by ONE 3-D TMA box {128 bytes of blades, rows, nb * sizeof(T) / 128
// chunks} with the 128-byte swizzle, so its shared-memory image is [chunk][row][128 bytes]:
//     offset(row r, blade A) = G(A) ^ R(r),   R(r) = r << 7 | (r & 7) << 4,
//     G(A) = (A >> 5) << 12 | ((A >> 2) & 7) << 4 | (A & 3) << 2        (f32: 32 blades per line, 4 KB per chunk)
//     G(A) = (A >> 4) << 11 | ((A >> 1) & 7) << 4 | (A & 1) << 3        (f64: 16 blades per line, 2 KB per chunk)
// -- the same "linear slot offset XOR row term" algebra as the blade-major layout above, with other constants.  A warp's
// lanes are 8 rows x 4 indices (string rows a in the transforms, contraction / output indices in the MMA fragments); the 8
// rows always fall on 8 different 16-byte columns (the swizzle), so an access is conflict free iff the 4 (f64 half-warps:
// 2 to 4) indices differ in the low blade bits that select the word inside the 16-byte column (or, for the f64 B
// fragments whose half-warps hold only 4 rows, in blade bit 3).  Three index bases are searched for that:
//     amap   string rows of the transforms          f32: lin_hi & 3 independent on (u0, u1); f64: lin_hi(u0) & 1
//     kmap   contraction index at the MMA's k       f32: Lambda & 3 independent on (v0, v1); f64: bits {0, 3} of Lambda
//     omap   output index at the MMA's m            f32: Lambda & 7 independent on (w0, w1, w2); f64: Lambda & 3 on (w0, w1)
// and every warp-wide access the kernel makes is then replayed against the bank rule before the layout is accepted.
struct RepPassAos {
    uint32_t fk[16];   // G(Lambda(kmap[physical k]))
    uint32_t fo[16];   // G(Lambda(omap[physical m]))
    uint32_t cw[16];   // G(Gamma(w))
    uint8_t kmap[16], omap[16];
};
struct RepLayoutAos {
    bool ok = false;
    uint32_t g_lo[16];   // G(lin_lo(b))
    uint32_t g_hi[16];   // G(lin_hi(amap[i])): physical string-row index i
    uint8_t amap[16];    // physical string-row index -> string row a
    RepPassAos column, row;
};

namespace matrix_rep_detail {
inline uint32_t aos_offset(unsigned A, int elem_bytes) {
    return elem_bytes == 4 ? ((A >> 5) << 12) | (((A >> 2) & 7u) << 4) | ((A & 3u) << 2) : ((A >> 4) << 11) | (((A >> 1) & 7u) << 4) | ((A & 1u) << 3);
}
inline uint32_t aos_row_term(unsigned r) { return (r << 7) | ((r & 7u) << 4); }
// do the `lanes` byte offsets of one access of `bytes` bytes per lane fall on distinct banks (4-byte accesses: 32 lanes;
// 8-byte accesses: the hardware serves half-warps)?
inline bool banks_distinct(const uint32_t *off, int lanes, int bytes) {
    uint32_t seen = 0;
    for (int i = 0; i < lanes; ++i) {
        const uint32_t bank = bytes == 4 ? (off[i] >> 2) & 31u : (off[i] >> 3) & 15u;
        if (seen & (1u << bank)) return false;
        seen |= 1u << bank;
    }
    return true;
}
// first basis (ascending search) of F_2^4 whose index map i -> XOR of the basis vectors of the set bits of i passes `good`
template <class F> inline bool pick_map(uint8_t map[16], F good) {
    for (unsigned v0 = 1; v0 < 16; ++v0)
        for (unsigned v1 = 1; v1 < 16; ++v1) {
            if (v1 == v0) continue;
            for (unsigned v2 = 1; v2 < 16; ++v2) {
                if (v2 == v0 || v2 == v1 || v2 == (v0 ^ v1)) continue;
                for (unsigned v3 = 1; v3 < 16; ++v3) {
                    bool in_span = false;
                    for (unsigned m = 0; m < 8; ++m)
                        if (v3 == (((m & 1) ? v0 : 0) ^ ((m & 2) ? v1 : 0) ^ ((m & 4) ? v2 : 0))) in_span = true;
                    if (in_span) continue;
                    const unsigned v[4] = {v0, v1, v2, v3};
                    for (unsigned i = 0; i < 16; ++i) {
                        unsigned k = 0;
                        for (int bit = 0; bit < 4; ++bit) if ((i >> bit) & 1) k ^= v[bit];
                        map[i] = (uint8_t)k;
                    }
                    if (good(map)) return true;
                    break;  // the conditions only involve v0..v2: another v3 cannot help
                }
            }
        }
    return false;
}
inline bool make_pass_aos(const unsigned lam[16], const unsigned gam[16], int es, RepPassAos &p) {
    // B fragments: lane = 4 g + t reads index position (4 j + t) at row g;  D fragments: lane holds output position g (+ 8)
    // at rows 2 t and 2 t + 1 (two separate stores)
    auto k_ok = [&](const uint8_t *map) {
        for (int j = 0; j < 4; ++j) {
            uint32_t off[32];
            for (int g = 0; g < 8; ++g)
                for (int t = 0; t < 4; ++t) off[4 * g + t] = aos_offset(lam[map[4 * j + t]], es) ^ aos_row_term(g);
            if (es == 4 ? !banks_distinct(off, 32, 4) : !(banks_distinct(off, 16, 8) && banks_distinct(off + 16, 16, 8))) return false;
        }
        return true;
    };
    auto o_ok = [&](const uint8_t *map) {
        for (int mb = 0; mb < 2; ++mb)
            for (int e = 0; e < 2; ++e) {
                uint32_t off[32];
                for (int g = 0; g < 8; ++g)
                    for (int t = 0; t < 4; ++t) off[4 * g + t] = aos_offset(lam[map[g + 8 * mb]], es) ^ aos_row_term(2 * t + e);
                if (es == 4 ? !banks_distinct(off, 32, 4) : !(banks_distinct(off, 16, 8) && banks_distinct(off + 16, 16, 8))) return false;
            }
        return true;
    };
    if (!pick_map(p.kmap, k_ok) || !pick_map(p.omap, o_ok)) return false;
    for (unsigned i = 0; i < 16; ++i) {
        p.fk[i] = aos_offset(lam[p.kmap[i]], es);
        p.fo[i] = aos_offset(lam[p.omap[i]], es);
        p.cw[i] = aos_offset(gam[i], es);
    }
    return true;
}
}  // namespace matrix_rep_detail

inline RepLayoutAos make_rep_layout_aos(const RepLayout &base, int elem_bytes) {
    using namespace matrix_rep_detail;
    RepLayoutAos L;
    if (!base.lin_ok) return L;
    unsigned lam_c[16], gam_c[16], lam_r[16], gam_r[16];
    for (unsigned i = 0; i < 16; ++i) {
        L.g_lo[i] = aos_offset(base.lin[i], elem_bytes);
        lam_c[i] = base.lin[i << 4];              gam_c[i] = base.lin[(i << 4) | i];
        lam_r[i] = base.lin[(i << 4) | i];        gam_r[i] = base.lin[i << 4];
    }
    // transforms: lane = 8 alpha + rho handles string row amap[4 * group + alpha] at row rho (+ 8, + 16, ...)
    auto a_ok = [&](const uint8_t *map) {
        for (int grp = 0; grp < 4; ++grp) {
            uint32_t off[32];
            for (int al = 0; al < 4; ++al)
                for (int rho = 0; rho < 8; ++rho) off[8 * al + rho] = aos_offset(base.lin[(unsigned)map[4 * grp + al] << 4], elem_bytes) ^ aos_row_term(rho);
            if (elem_bytes == 4 ? !banks_distinct(off, 32, 4) : !(banks_distinct(off, 16, 8) && banks_distinct(off + 16, 16, 8))) return false;
        }
        return true;
    };
    if (!pick_map(L.amap, a_ok)) return L;
    for (unsigned i = 0; i < 16; ++i) L.g_hi[i] = aos_offset(base.lin[(unsigned)L.amap[i] << 4], elem_bytes);
    L.ok = make_pass_aos(lam_c, gam_c, elem_bytes, L.column) && make_pass_aos(lam_r, gam_r, elem_bytes, L.row);
    return L;
}

}  // namespace lcc
