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
inline MatrixRep find_matrix_rep(const lcc_algebra &alg) {
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
    if (!search(alg.p, alg.q, 0, gens, echelon)) return rep;
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
    for (int A = 0; A < nb; ++A)
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

}  // namespace lcc
