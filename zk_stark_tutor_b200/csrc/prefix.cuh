// prefix.cuh - host arithmetic behind the subgroup-prefix interpolation of stark.cu (pure host code on the verified
// host path of fe128.cuh, so tests/test_air_host.py can check it against the oracle without a GPU).
#pragma once
#include <stdint.h>
#include <vector>
#include "fe128.cuh"

namespace zkb {

// Z(X) = prod_{i < L} (X - root^i): L + 1 canonical coefficients, ascending (monic).  O(L^2) schoolbook, run once per shape.
inline std::vector<fe> prefix_zerofier(const fe& root, uint64_t L) {
    std::vector<fe> Z(L + 1, fe_zero());                       // Z[k] = 0 above the current degree
    Z[0] = fe_from_u32(1);
    fe x = fe_from_u32(1);
    for (uint64_t i = 0; i < L; i++) {                          // Z <- Z * (X - root^i): Z'[k] = Z[k-1] - root^i * Z[k]
        const fe x_m = fe_to_mont(x);
        for (int64_t k = (int64_t)i + 1; k >= 0; k--)
            Z[k] = fe_sub(k ? Z[k - 1] : fe_zero(), fe_montmul(x_m, Z[k]));
        x = fe_mul(x, root);
    }
    return Z;
}

// g = rev(Z)^-1 mod X^m for a monic Z (so rev(Z)[0] = 1): g[0] = 1, g[k] = -sum_{j=1..min(k, deg Z)} rev(Z)[j] * g[k-j].  O(m * deg Z).
inline std::vector<fe> reversed_series_inverse(const std::vector<fe>& Z, uint64_t m) {
    const uint64_t L = Z.size() - 1;
    std::vector<fe> g(m, fe_zero());
    if (m == 0) return g;
    std::vector<fe> f_m(L + 1);
    for (uint64_t j = 0; j <= L; j++) f_m[j] = fe_to_mont(Z[L - j]);
    g[0] = fe_from_u32(1);
    for (uint64_t k = 1; k < m; k++) {
        fe acc = fe_zero();
        const uint64_t jmax = k < L ? k : L;
        for (uint64_t j = 1; j <= jmax; j++) acc = fe_add(acc, fe_montmul(f_m[j], g[k - j]));
        g[k] = fe_neg(acc);
    }
    return g;
}

}  // namespace zkb
