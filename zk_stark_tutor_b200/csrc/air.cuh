// air.cuh - evaluation-form AIR quotients and the nonlinear combination, one point of the FRI coset at a time.
//
// What it replaces (reference file:line): the middle of Stark::prove between the committed codewords and
// FRI::prove - MPolynomial::evaluate_symbolic (src/m_polynomial.rs:128-142, schoolbook polynomial products),
// the transition quotients' fast_coset_divide (src/stark/stark.rs:388-420), the x^shift products
// (fast_multiply, stark.rs:466-498), the weighted sum (stark.rs:500-512) and the final LDE of the combination
// (stark.rs:514-519).  All of these are polynomial identities, so on the coset x_i = offset*omega^i the
// combined codeword is, value for value,
//     comb[i] = w0*rand[i] + sum_j q_j(x_i)*(w_{1+2j} + w_{2+2j}*x_i^shift_j) + sum_s bq_s[i]*(w_.. + w_..*x_i^shift_s)
//     q_j(x)  = tc_j(x, t(x), t(omicron*x)) / Z_T(x),        t_s(x) = bq_s(x)*Z_B,s(x) + I_s(x)
// - exactly the per-index formula the reference's verifier evaluates (stark.rs:679-769) - computed here for
// every index from the committed codewords, without leaving HBM.  Field arithmetic is exact, so the result is
// bit-identical to the reference's coefficient-form route as long as every degree stays below the domain length.
//
// The point body is __host__ __device__ so tests/test_air_host.py can run it on the CPU against the oracle.
#pragma once
#include <stdint.h>
#include <map>
#include <vector>
#include "fe128.cuh"

namespace zkb {

constexpr int AIR_MAX_STATE = 16;      // 2 * num_registers (this row, next row)

// One group of terms of a constraint that share the exponents of the state variables: the inner factor is a dense
// univariate polynomial in x (Horner), the outer factor a product of small powers of the state variables.
struct AirGroup {
    uint32_t coef_off, ncoef;          // coefficients coefs[coef_off .. +ncoef), ascending powers of x, Montgomery form
    uint8_t e[AIR_MAX_STATE];
};

// Batched: instance b = blockIdx.y.  What depends on the AIR's shape only (zerofier codewords, grouped terms, shifts) is
// shared; the committed codewords, the boundary interpolants (they carry the instance's public values) and the weights
// (from the instance's transcript) have an instance stride.
struct AirView {
    uint64_t n, rot;                   // domain length; index distance of the next trace row (= expansion factor)
    uint32_t nr, nc;                   // registers, constraints
    const fe* bq; uint64_t bq_stride, bq_inst;   // boundary-quotient codewords: register s of instance b at bq + s*bq_stride + b*bq_inst
    const fe* rnd; uint64_t rnd_inst;  // randomizer codewords
    const fe* zb;                      // boundary zerofier codewords, register s at + s*n (shared)
    const fe* ib; uint64_t ib_inst;    // boundary interpolant codewords, register s of instance b at + s*n + b*ib_inst
    const fe* tz_inv_m;                // 1 / Z_T(x_i), Montgomery form (shared)
    const AirGroup* groups;            // all constraints' groups
    const uint32_t* group_begin;       // nc + 1 offsets into groups
    const fe* coefs;
    const fe* weights; uint32_t nw;    // per instance 1 + 2*nc + 2*nr canonical values, stark.rs:447-450 order
    const uint32_t* shifts;            // nc + nr exponents of x (stark.rs:476, 489)
    fe* tq_out; uint64_t tq_inst;      // nullptr, or per instance nc codewords of n values: the transition quotients
};

#if defined(__CUDA_ARCH__)
#define ZKB_AIR_LD(p) fe_ldg(p)
#else
#define ZKB_AIR_LD(p) (*(p))
#endif

// a^-1 * R for a*R given (Fermat: p-2 = 0xCB7FFFFF_FFFFFFFF_FFFFFFFF_FFFFFFFF); 0 -> 0
ZKB_HD fe fe_mont_inv(const fe& a_m) {
    fe acc = fe_mont_one();
    for (int bit = 127; bit >= 0; bit--) {
        acc = fe_montmul(acc, acc);
        uint32_t word = (bit >= 96) ? 0xCB7FFFFFu : 0xFFFFFFFFu;
        if ((word >> (bit & 31)) & 1u) acc = fe_montmul(acc, a_m);
    }
    return acc;
}

// t_s(x_i) = bq_s(x_i) * Z_B,s(x_i) + I_s(x_i)  (stark.rs:722-731) for instance b
ZKB_HD fe air_trace_value(const AirView& a, uint32_t b, uint32_t s, uint64_t i) {
    const fe q = ZKB_AIR_LD(a.bq + s * a.bq_stride + b * a.bq_inst + i);
    return fe_add(fe_montmul(fe_to_mont(q), ZKB_AIR_LD(a.zb + s * a.n + i)), ZKB_AIR_LD(a.ib + s * a.n + b * a.ib_inst + i));
}

// comb[i] of instance b.  x_m = x_i * R.
ZKB_HD fe air_point(const AirView& a, uint32_t b, uint64_t i, const fe& x_m) {
    const uint64_t i2 = (i + a.rot) % a.n;
    fe v[AIR_MAX_STATE];                                     // state variables, Montgomery form
    for (uint32_t s = 0; s < a.nr; s++) {
        v[s] = fe_to_mont(air_trace_value(a, b, s, i));
        v[a.nr + s] = fe_to_mont(air_trace_value(a, b, s, i2));
    }
    const fe tz_inv_m = ZKB_AIR_LD(a.tz_inv_m + i);
    const fe* w = a.weights + (uint64_t)b * a.nw;
    fe comb = fe_montmul(fe_to_mont(ZKB_AIR_LD(a.rnd + b * a.rnd_inst + i)), ZKB_AIR_LD(w));      // w0 * randomizer
    for (uint32_t j = 0; j < a.nc; j++) {
        fe acc = fe_zero();                                                                    // tc_j(point) * R
        for (uint32_t g = a.group_begin[j]; g < a.group_begin[j + 1]; g++) {
            const AirGroup& gr = a.groups[g];
            const fe* cf = a.coefs + gr.coef_off;
            fe h = ZKB_AIR_LD(cf + gr.ncoef - 1);
            for (uint32_t k = gr.ncoef - 1; k-- > 0;) h = fe_add(fe_montmul(h, x_m), ZKB_AIR_LD(cf + k));
            for (uint32_t s = 0; s < 2 * a.nr; s++) {
                uint32_t e = gr.e[s];
                if (e == 0) continue;
                fe pw = v[s];
                for (uint32_t r = 1; r < e; r++) pw = fe_montmul(pw, v[s]);                   // exponents are tiny (<= the AIR degree)
                h = fe_montmul(h, pw);
            }
            acc = fe_add(acc, h);
        }
        const fe q_m = fe_montmul(acc, tz_inv_m);                                              // transition quotient (stark.rs:744-746)
        if (a.tq_out) a.tq_out[b * a.tq_inst + (uint64_t)j * a.n + i] = fe_from_mont(q_m);
        const fe xs_m = fe_mont_pow(x_m, a.shifts[j]);
        const fe cw = fe_add(ZKB_AIR_LD(w + 1 + 2 * j), fe_montmul(xs_m, ZKB_AIR_LD(w + 2 + 2 * j)));
        comb = fe_add(comb, fe_montmul(q_m, cw));
    }
    for (uint32_t s = 0; s < a.nr; s++) {
        const fe xs_m = fe_mont_pow(x_m, a.shifts[a.nc + s]);
        const uint32_t k = 1 + 2 * a.nc + 2 * s;
        const fe cw = fe_add(ZKB_AIR_LD(w + k), fe_montmul(xs_m, ZKB_AIR_LD(w + k + 1)));
        comb = fe_add(comb, fe_montmul(fe_to_mont(ZKB_AIR_LD(a.bq + s * a.bq_stride + b * a.bq_inst + i)), cw));
    }
    return comb;
}

// boundary quotient in evaluation form (stark.rs:331-360 on the coset): bq_s(x_i) = (t_s(x_i) - I_s(x_i)) / Z_B,s(x_i);
// zb_inv_m = 1 / Z_B,s(x_i) in Montgomery form
ZKB_HD fe air_boundary_quotient(const fe& t, const fe& interpolant, const fe& zb_inv_m) {
    return fe_montmul(fe_sub(t, interpolant), zb_inv_m);
}

// ---- host: MPolynomial dictionaries (flattened) -> groups ----------------------------------------
struct AirTables {
    std::vector<AirGroup> groups;
    std::vector<uint32_t> group_begin;
    std::vector<fe> coefs;             // Montgomery form
};
// term t of constraint j: coefficient coefs[t], exponents exps[t*nvars .. +nvars) over (x, registers now, registers next).
// Returns 0, or -1 if an exponent of a state variable exceeds 255 / nvars is inconsistent.
inline int air_group_terms(uint32_t nc, uint32_t nr, const uint32_t* term_counts, const fe* coefs, const uint32_t* exps, AirTables* out) {
    const uint32_t nvars = 1 + 2 * nr;
    if (2 * nr > (uint32_t)AIR_MAX_STATE) return -1;
    size_t t = 0;
    out->group_begin.assign(1, 0);
    for (uint32_t j = 0; j < nc; j++) {
        std::map<std::vector<uint8_t>, std::vector<fe>> by_state;              // state exponents -> coefficients by power of x
        for (uint32_t k = 0; k < term_counts[j]; k++, t++) {
            const uint32_t* e = exps + t * nvars;
            std::vector<uint8_t> key(2 * nr);
            for (uint32_t s = 0; s < 2 * nr; s++) { if (e[1 + s] > 255) return -1; key[s] = (uint8_t)e[1 + s]; }
            if (e[0] > (1u << 20)) return -1;
            std::vector<fe>& cf = by_state[key];
            if (cf.size() <= e[0]) cf.resize(e[0] + 1, fe_zero());
            cf[e[0]] = fe_add(cf[e[0]], coefs[t]);                               // the dictionary has unique keys; adding is harmless
        }
        for (auto& kv : by_state) {
            AirGroup g;
            g.coef_off = (uint32_t)out->coefs.size();
            g.ncoef = (uint32_t)kv.second.size();
            for (int s = 0; s < AIR_MAX_STATE; s++) g.e[s] = s < (int)kv.first.size() ? kv.first[s] : 0;
            for (const fe& c : kv.second) out->coefs.push_back(fe_to_mont(c));
            out->groups.push_back(g);
        }
        out->group_begin.push_back((uint32_t)out->groups.size());
    }
    return 0;
}

}  // namespace zkb
