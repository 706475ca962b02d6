// TEST INFRASTRUCTURE: csrc/air.cuh (the __host__ __device__ point body of k_air_combine and the host-side
// grouping of constraint terms) compiled for the CPU, so tests/test_air_host.py can check the evaluation-form
// combination against the oracle's coefficient-form Stark::prove without a GPU.  The device build of the very
// same functions is covered by the -m gpu tests.
#include <cstring>
#include "../../zk_stark_tutor_b200/csrc/air.cuh"

using namespace zkb;

extern "C" int air_host_combination(uint64_t n, uint64_t rot, uint32_t nr, uint32_t nc, const uint32_t* term_counts, const fe* coefs,
                                    const uint32_t* exps, const fe* bq, const fe* rnd, const fe* zb, const fe* ib, const fe* tz,
                                    const fe* weights, const uint32_t* shifts, const fe* offset, const fe* omega, fe* out, fe* tq_out) {
    AirTables tab;
    if (air_group_terms(nc, nr, term_counts, coefs, exps, &tab) != 0) return -1;
    AirView v;
    v.n = n; v.rot = rot; v.nr = nr; v.nc = nc;
    v.bq = bq; v.bq_stride = n; v.rnd = rnd; v.zb = zb; v.ib = ib; v.tz = tz;
    v.groups = tab.groups.data(); v.group_begin = tab.group_begin.data(); v.coefs = tab.coefs.data();
    v.weights = weights; v.shifts = shifts; v.tq_out = tq_out;
    fe x_m = fe_to_mont(*offset);
    const fe w_m = fe_to_mont(*omega);
    bool dz = false;
    for (uint64_t i = 0; i < n; i++) {
        out[i] = air_point(v, i, x_m, &dz);
        x_m = fe_montmul(x_m, w_m);
    }
    return dz ? -7 : (int)tab.groups.size();
}
