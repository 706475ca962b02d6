// TEST INFRASTRUCTURE: csrc/air.cuh (the __host__ __device__ point body of k_air_combine and the host-side
// grouping of constraint terms) compiled for the CPU, so tests/test_air_host.py can check the evaluation-form
// combination against the oracle's coefficient-form Stark::prove without a GPU.  The device build of the very
// same functions is covered by the -m gpu tests.
#include <cstring>
#include "../../zk_stark_tutor_b200/csrc/air.cuh"
#include "../../zk_stark_tutor_b200/csrc/prefix.cuh"

using namespace zkb;

// instances: `batch` of them, instance-major buffers (bq: batch x nr x n, rnd: batch x n, ib: batch x nr x n, weights: batch x nw,
// out: batch x n, tq_out: batch x nc x n); zb: nr x n and tz: n are shared.  Returns the number of term groups, or -7 if a zerofier
// value is 0 (the device path reports ZKB_ERR_DIV_ZERO there).
extern "C" int air_host_combination(uint64_t n, uint64_t rot, uint32_t nr, uint32_t nc, uint32_t batch, const uint32_t* term_counts, const fe* coefs,
                                    const uint32_t* exps, const fe* bq, const fe* rnd, const fe* zb, const fe* ib, const fe* tz,
                                    const fe* weights, const uint32_t* shifts, const fe* offset, const fe* omega, fe* out, fe* tq_out) {
    AirTables tab;
    if (air_group_terms(nc, nr, term_counts, coefs, exps, &tab) != 0) return -1;
    std::vector<fe> tz_inv_m(n);
    bool dz = false;
    for (uint64_t i = 0; i < n; i++) { dz = dz || fe_is_zero(tz[i]); tz_inv_m[i] = fe_mont_inv(fe_to_mont(tz[i])); }
    AirView v;
    v.n = n; v.rot = rot; v.nr = nr; v.nc = nc;
    v.bq = bq; v.bq_stride = n; v.bq_inst = (uint64_t)nr * n;
    v.rnd = rnd; v.rnd_inst = n;
    v.zb = zb; v.ib = ib; v.ib_inst = (uint64_t)nr * n; v.tz_inv_m = tz_inv_m.data();
    v.groups = tab.groups.data(); v.group_begin = tab.group_begin.data(); v.coefs = tab.coefs.data();
    v.weights = weights; v.nw = 1 + 2 * nc + 2 * nr; v.shifts = shifts;
    v.tq_out = tq_out; v.tq_inst = (uint64_t)nc * n;
    const fe w_m = fe_to_mont(*omega);
    for (uint32_t b = 0; b < batch; b++) {
        fe x_m = fe_to_mont(*offset);
        for (uint64_t i = 0; i < n; i++) {
            out[(uint64_t)b * n + i] = air_point(v, b, i, x_m);
            x_m = fe_montmul(x_m, w_m);
        }
    }
    return dz ? -7 : (int)tab.groups.size();
}

// bq = (t - I) / Z_B pointwise (k_boundary_quotient's body), one register
extern "C" void air_host_boundary_quotient(uint64_t n, const fe* t, const fe* interpolant, const fe* zb, fe* out) {
    for (uint64_t i = 0; i < n; i++) out[i] = air_boundary_quotient(t[i], interpolant[i], fe_mont_inv(fe_to_mont(zb[i])));
}

// the host tables of the subgroup-prefix interpolation (csrc/prefix.cuh): Z (L + 1 values) and rev(Z)^-1 mod x^m (m values)
extern "C" void prefix_host_tables(const fe* root, uint64_t L, uint64_t m, fe* Z_out, fe* g_out) {
    std::vector<fe> Z = prefix_zerofier(*root, L), g = reversed_series_inverse(Z, m);
    memcpy(Z_out, Z.data(), (L + 1) * sizeof(fe));
    if (m) memcpy(g_out, g.data(), m * sizeof(fe));
}
