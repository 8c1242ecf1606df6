// algebra_host.hpp -- host-side Cl(p,q,r) registry entry: signature, flat Cayley tables, blade
// names and grade masks.  The B200-native counterpart of Arc<Algebra>
// (/root/reference/src/algebra/registry.rs:36-81) with its table builder
// (/root/reference/src/algebra/cayley.rs:63-144, signature.rs:173-182, blades.rs:24-61,128-229).
// Integer work only; runs without a GPU.  Everything here is bit-exact contract surface.
#pragma once
#include <cstdint>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "kernels/launch_args.h"

struct lcc_algebra {
    int p = 0, q = 0, r = 0, dim = 0, nb = 1;
    std::vector<int8_t> signs;     // [a*nb+b] in {-1,0,+1}
    std::vector<uint16_t> blades;  // [a*nb+b] = a^b
    std::vector<std::string> names;
    std::vector<std::vector<uint64_t>> grade_masks;  // [grade][word]
    std::vector<double> mu;        // mu[c] = product of basis squares over the bits of c (metric factor of i&j == c)
    std::vector<int8_t> norm_weights;  // sign[k][k] * reverse_sign(k)
    bool pseudoscalar_invertible = false;
    bool specialised = false;      // build-time generated kernels exist for this signature
    lcc::ProductLauncher product[2] = {nullptr, nullptr};   // [dtype]
    lcc::SandwichLauncher sandwich[2] = {nullptr, nullptr};
    // gather-form sign tables G[k*nb+i] for the table-driven kernel, uploaded lazily per (device, kind):
    // kinds 0..2 forward gp / outer / lc, 3..5 their VJP wrt the left operand, 6..8 wrt the right one
    std::mutex mu_dev;
    std::map<std::pair<int, int>, int8_t *> gather_signs_dev;
    // 16 x 16 real matrix representation (matrix_rep.hpp), searched on first use; rep_state: 0 not tried, 1 found, 2 none
    int rep_state = 0;
    std::vector<uint8_t> rep_string_of, rep_blade_at;  // blade -> string, string -> blade (0 when none)
    std::vector<int8_t> rep_sign_of, rep_sign_at;      // blade -> sign, string -> sign (0 when none)
};

namespace lcc {

inline int basis_square(int p, int q, int /*r*/, int i) { return i < p ? 1 : (i < p + q ? -1 : 0); }

inline int popcount32(unsigned v) { return __builtin_popcount(v); }

// cayley.rs:89-111 -- for each bit of b, the parity of the bits of a strictly above it
inline int reorder_sign(unsigned a, unsigned b) {
    int sign = 1;
    while (b) {
        unsigned low = b & (~b + 1u);
        int pos = __builtin_ctz(low);
        if (popcount32(a >> (pos + 1)) & 1) sign = -sign;
        b &= ~low;
    }
    return sign;
}

// cayley.rs:63-80
inline int blade_sign(unsigned a, unsigned b, int p, int q, int r) {
    int sign = reorder_sign(a, b);
    unsigned common = a & b;
    for (int i = 0; i < p + q + r; ++i)
        if ((common >> i) & 1u) sign *= basis_square(p, q, r, i);
    return sign;
}

inline int reverse_sign_of_grade(int g) { return (g % 4 < 2) ? 1 : -1; }            // blades.rs:128-135
inline int involution_sign_of_grade(int g) { return (g % 2 == 0) ? 1 : -1; }        // blades.rs:154-160
inline int conjugate_sign_of_grade(int g) { int m = g % 4; return (m == 0 || m == 3) ? 1 : -1; }  // blades.rs:182-188

void build_algebra(lcc_algebra &alg, int p, int q, int r);

}  // namespace lcc
