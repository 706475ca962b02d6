// air.cu - the Stark prover's boundary quotients, transition quotients and nonlinear combination in evaluation
// form on the FRI coset (see air.cuh for the identity and the reference lines it replaces), for one instance or a
// batch of instances of the same AIR advancing in lockstep (blockIdx.y = instance).
//   zkb_air_create              what depends on the AIR's shape only: grouped constraint terms, the boundary / transition
//                               zerofiers evaluated on the coset and inverted there (once per AIR, not per proof)
//   zkb_air_set_interpolants    the instances' boundary interpolants -> codewords (they carry the public values)
//   zkb_air_boundary_quotients  bq = (t - I) / Z_B pointwise from the trace codewords          (stark.rs:331-360)
//   zkb_air_combine             transition quotients + x^shift products + weighted sum        (stark.rs:388-519)
//   zkb_air_combination         the one-call form for a single instance (create + set + combine)
// One thread per domain point; every input is a codeword already in HBM or a small table broadcast through L1/L2.
#include <string.h>
#include "air.cuh"
#include "ctx.hpp"
#include "ntt.cuh"

struct zkb_air {
    zkb_ctx* ctx = nullptr;
    uint64_t n = 0, rot = 0;
    uint32_t nr = 0, nc = 0, nw = 0;
    zkb::fe omega, offset;
    void* tables = nullptr;          // one device allocation: coefs | groups | group_begin | shifts | flag
    size_t o_coef = 0, o_groups = 0, o_begin = 0, o_shift = 0, o_flag = 0;
    zkb::fe* zb = nullptr;           // nr x n   boundary zerofier codewords
    zkb::fe* zb_inv_m = nullptr;     // nr x n   their inverses, Montgomery form
    zkb::fe* tz_inv_m = nullptr;     // n        inverse transition zerofier, Montgomery form
    zkb::fe* ib = nullptr;           // batch x nr x n interpolant codewords (zkb_air_set_interpolants)
    size_t ib_batch = 0;
    zkb::fe* weights = nullptr;      // device staging for the instances' weights
    size_t weights_cap = 0;
};

namespace zkb {

struct AirLaunch {
    AirView v;
    DevPow omega_pow;     // omega^i * R
    fe offset_m;          // offset * R
    fe* out; uint64_t out_inst;
};

__device__ __forceinline__ fe air_x_m(const DevPow& t, const fe& offset_m, uint64_t i) {
    const fe w_m = fe_montmul(fe_ldg(t.hi + (i >> t.lo_bits)), fe_ldg(t.lo + (i & ((1ull << t.lo_bits) - 1))));
    return fe_montmul(w_m, offset_m);                         // offset * omega^i * R
}

__global__ void __launch_bounds__(128) k_air_combine(AirLaunch a) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.v.n) return;
    const uint32_t b = blockIdx.y;
    fe_store(a.out + b * a.out_inst + i, air_point(a.v, b, i, air_x_m(a.omega_pow, a.offset_m, i)));
}

// out[i] = (1 / in[i]) * R; *flag |= 1 if some in[i] == 0 (the reference's division panics there, field_element.rs:85)
__global__ void __launch_bounds__(128) k_invert_m(const fe* __restrict__ in, fe* __restrict__ out, uint64_t count, uint32_t* flag) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const fe v = fe_ldg(in + i);
    if (fe_is_zero(v)) atomicOr(flag, 1u);
    fe_store(out + i, fe_mont_inv(fe_to_mont(v)));
}

struct BqLaunch {
    uint64_t n; uint32_t nr;
    const fe* t; uint64_t t_stride, t_inst;
    const fe* ib; uint64_t ib_inst;
    const fe* zb_inv_m;
    fe* bq; uint64_t bq_stride, bq_inst;
};
__global__ void __launch_bounds__(128) k_boundary_quotient(BqLaunch a) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    const uint32_t b = blockIdx.y;
    for (uint32_t s = 0; s < a.nr; s++)
        fe_store(a.bq + s * a.bq_stride + b * a.bq_inst + i,
                 air_boundary_quotient(fe_ldg(a.t + s * a.t_stride + b * a.t_inst + i), fe_ldg(a.ib + s * a.n + b * a.ib_inst + i),
                                       fe_ldg(a.zb_inv_m + s * a.n + i)));
}

static int air_check_flag(zkb_air* a, const char* what) {
    zkb_ctx* c = a->ctx;
    uint32_t flag = 0;
    uint32_t* d_flag = (uint32_t*)((uint8_t*)a->tables + a->o_flag);
    ZKB_CUDA(c, cudaMemcpyAsync(&flag, d_flag, 4, cudaMemcpyDeviceToHost, c->stream));
    ZKB_CUDA(c, ctx_stream_sync(c));
    if (flag) {
        cudaMemsetAsync(d_flag, 0, 4, c->stream);
        return set_err(c, ZKB_ERR_DIV_ZERO, "%s vanishes on the FRI domain (divide by zero)", what);
    }
    return 0;
}

}  // namespace zkb

using namespace zkb;

extern "C" {

void zkb_air_free(zkb_air* a) {
    if (!a) return;
    zkb_ctx* c = a->ctx;
    dev_free(c, a->tables);
    dev_free(c, a->zb);
    dev_free(c, a->ib);
    dev_free(c, a->weights);
    delete a;
}

int zkb_air_create(zkb_ctx* c, const zkb_air_shape* d, zkb_air** out) {
    if (!c || !d || !out) return ZKB_ERR_ARG;
    *out = nullptr;
    const uint64_t n = d->domain_length;
    const uint32_t nr = d->num_registers, nc = d->num_constraints;
    if (n < 2 || (n & (n - 1))) return set_err(c, ZKB_ERR_ARG, "air: domain length %llu is not a power of two", (unsigned long long)n);
    if (nr == 0 || 2 * nr > (uint32_t)AIR_MAX_STATE) return set_err(c, ZKB_ERR_ARG, "air: 1..%d registers supported", AIR_MAX_STATE / 2);
    if (d->expansion_factor == 0 || d->expansion_factor >= n) return set_err(c, ZKB_ERR_ARG, "air: bad expansion factor");
    if ((nc && (!d->term_counts || !d->coefs || !d->exps)) || !d->boundary_zerofiers || !d->boundary_zerofier_lens || !d->transition_zerofier || !d->shifts)
        return ZKB_ERR_ARG;
    ZKB_CUDA(c, cudaSetDevice(c->device));
    AirTables tab;
    if (air_group_terms(nc, nr, d->term_counts, (const fe*)d->coefs, d->exps, &tab) != 0)
        return set_err(c, ZKB_ERR_ARG, "air: constraint exponents out of range");
    const uint32_t ns = nc + nr;
    std::vector<uint32_t> shifts(ns);
    for (uint32_t k = 0; k < ns; k++) {
        if (d->shifts[k] > 0xFFFFFFFFull) return set_err(c, ZKB_ERR_ARG, "air: shift out of range");
        shifts[k] = (uint32_t)d->shifts[k];
    }
    size_t plen = d->transition_zerofier_len;
    for (uint32_t s = 0; s < nr; s++) plen = plen > d->boundary_zerofier_lens[s] ? plen : d->boundary_zerofier_lens[s];
    if (plen == 0 || plen > n) return set_err(c, ZKB_ERR_TOO_LONG, "air: zerofier longer than the domain");

    zkb_air* a = new zkb_air();
    a->ctx = c; a->n = n; a->rot = d->expansion_factor; a->nr = nr; a->nc = nc; a->nw = 1 + 2 * nc + 2 * nr;
    a->omega = h_load(d->omega); a->offset = h_load(d->offset);
    auto align16 = [](size_t x) { return (x + 15) & ~(size_t)15; };
    const uint32_t ncols = nr + 1;
    const size_t o_poly = 0;
    a->o_coef = o_poly + (size_t)ncols * plen * sizeof(fe);
    a->o_groups = a->o_coef + tab.coefs.size() * sizeof(fe);
    a->o_begin = align16(a->o_groups + tab.groups.size() * sizeof(AirGroup));
    a->o_shift = align16(a->o_begin + tab.group_begin.size() * 4);
    a->o_flag = align16(a->o_shift + ns * 4);
    const size_t total = align16(a->o_flag + 4);
    std::vector<uint8_t> stage(total, 0);
    for (uint32_t s = 0; s < nr; s++)
        if (d->boundary_zerofier_lens[s]) memcpy(&stage[o_poly + (size_t)s * plen * sizeof(fe)], d->boundary_zerofiers[s], d->boundary_zerofier_lens[s] * sizeof(fe));
    memcpy(&stage[o_poly + (size_t)nr * plen * sizeof(fe)], d->transition_zerofier, d->transition_zerofier_len * sizeof(fe));
    if (!tab.coefs.empty()) memcpy(&stage[a->o_coef], tab.coefs.data(), tab.coefs.size() * sizeof(fe));
    if (!tab.groups.empty()) memcpy(&stage[a->o_groups], tab.groups.data(), tab.groups.size() * sizeof(AirGroup));
    memcpy(&stage[a->o_begin], tab.group_begin.data(), tab.group_begin.size() * 4);
    memcpy(&stage[a->o_shift], shifts.data(), ns * 4);
    int rc = 0;
    do {
        if (dev_alloc(c, &a->tables, total) != cudaSuccess || dev_alloc(c, (void**)&a->zb, (size_t)(3 * nr + 2) * n * sizeof(fe)) != cudaSuccess) {
            rc = set_err(c, ZKB_ERR_CUDA, "air: device allocation failed");
            break;
        }
        // layout of the codeword block: zb (nr) | tz (1) | zb_inv_m (nr) | tz_inv_m (1) | spare (nr, unused)
        fe* tz = a->zb + (size_t)nr * n;
        a->zb_inv_m = a->zb + (size_t)(nr + 1) * n;
        a->tz_inv_m = a->zb + (size_t)(2 * nr + 1) * n;
        if (cudaMemcpyAsync(a->tables, stage.data(), total, cudaMemcpyHostToDevice, c->stream) != cudaSuccess) { rc = set_err(c, ZKB_ERR_CUDA, "air: upload failed"); break; }
        NttOpts o;
        o.has_scale = true;
        o.scale_base = a->offset;
        rc = ntt_exec(c, a->omega, (const fe*)((const uint8_t*)a->tables + o_poly), plen, plen, a->zb, n, ncols, ilog2_u64(n), o);
        if (rc) break;
        (void)tz;
        {
            LaunchScope ls(c, K_ELEMENTWISE);
            const uint64_t count = (uint64_t)ncols * n;
            k_invert_m<<<(unsigned)((count + 127) / 128), 128, 0, c->stream>>>(a->zb, a->zb_inv_m, count, (uint32_t*)((uint8_t*)a->tables + a->o_flag));
        }
        if (cudaGetLastError() != cudaSuccess) { rc = set_err(c, ZKB_ERR_CUDA, "air: k_invert_m launch failed"); break; }
        rc = air_check_flag(a, "air: a boundary or the transition zerofier");       // also makes `stage` safe to drop
    } while (0);
    if (rc) { zkb_air_free(a); return rc; }
    *out = a;
    return 0;
}

int zkb_air_set_interpolants(zkb_air* a, size_t batch, const void* interpolants, size_t interp_len) {
    if (!a || batch == 0 || !interpolants || interp_len == 0) return ZKB_ERR_ARG;
    zkb_ctx* c = a->ctx;
    if (interp_len > a->n) return set_err(c, ZKB_ERR_TOO_LONG, "air: interpolant longer than the domain");
    if (batch > 4096) return set_err(c, ZKB_ERR_ARG, "air: at most 4096 instances per call");
    ZKB_CUDA(c, cudaSetDevice(c->device));
    const size_t cols = batch * a->nr;
    if (batch > a->ib_batch) {
        dev_free(c, a->ib);
        a->ib = nullptr; a->ib_batch = 0;
        ZKB_CUDA(c, dev_alloc(c, (void**)&a->ib, cols * a->n * sizeof(fe)));
        a->ib_batch = batch;
    }
    DevBuf in;
    const void* d_in = nullptr;
    ZKB_TRY(stage_in(c, interpolants, cols * interp_len * sizeof(fe), in, &d_in));
    NttOpts o;
    o.has_scale = true;
    o.scale_base = a->offset;
    ZKB_TRY(ntt_exec(c, a->omega, (const fe*)d_in, interp_len, interp_len, a->ib, a->n, cols, ilog2_u64(a->n), o));
    if (in.p) ZKB_CUDA(c, ctx_stream_sync(c));        // the caller's host buffer has been consumed
    return 0;
}

int zkb_air_boundary_quotients(zkb_air* a, size_t batch, const void* trace_codewords, size_t trace_stride, size_t trace_inst,
                               void* bq_out, size_t bq_stride, size_t bq_inst) {
    if (!a || batch == 0 || !trace_codewords || !bq_out) return ZKB_ERR_ARG;
    zkb_ctx* c = a->ctx;
    if (batch > a->ib_batch) return set_err(c, ZKB_ERR_ARG, "air: zkb_air_set_interpolants must cover the batch first");
    if (!is_device_ptr(trace_codewords) || !is_device_ptr(bq_out)) return set_err(c, ZKB_ERR_ARG, "air: boundary_quotients takes device codewords");
    ZKB_CUDA(c, cudaSetDevice(c->device));
    BqLaunch l;
    l.n = a->n; l.nr = a->nr;
    l.t = (const fe*)trace_codewords; l.t_stride = trace_stride; l.t_inst = trace_inst;
    l.ib = a->ib; l.ib_inst = (uint64_t)a->nr * a->n;
    l.zb_inv_m = a->zb_inv_m;
    l.bq = (fe*)bq_out; l.bq_stride = bq_stride; l.bq_inst = bq_inst;
    {
        LaunchScope ls(c, K_ELEMENTWISE);
        k_boundary_quotient<<<dim3((unsigned)((a->n + 127) / 128), (unsigned)batch), 128, 0, c->stream>>>(l);
    }
    ZKB_CUDA(c, cudaGetLastError());
    return 0;
}

int zkb_air_combine(zkb_air* a, size_t batch, const uint8_t* weights, const void* bq, size_t bq_stride, size_t bq_inst,
                    const void* rnd, size_t rnd_inst, void* out, size_t out_inst, void* tq_out, size_t tq_inst) {
    if (!a || batch == 0 || !weights || !bq || !rnd || !out) return ZKB_ERR_ARG;
    zkb_ctx* c = a->ctx;
    if (batch > a->ib_batch) return set_err(c, ZKB_ERR_ARG, "air: zkb_air_set_interpolants must cover the batch first");
    if (!is_device_ptr(bq) || !is_device_ptr(rnd) || !is_device_ptr(out) || (tq_out && !is_device_ptr(tq_out)))
        return set_err(c, ZKB_ERR_ARG, "air: combine takes device codewords (they are the outputs of the LDEs on the device)");
    ZKB_CUDA(c, cudaSetDevice(c->device));
    const size_t wbytes = batch * a->nw * sizeof(fe);
    if (wbytes > a->weights_cap) {
        dev_free(c, a->weights);
        a->weights = nullptr; a->weights_cap = 0;
        ZKB_CUDA(c, dev_alloc(c, (void**)&a->weights, wbytes));
        a->weights_cap = wbytes;
    }
    ZKB_CUDA(c, cudaMemcpyAsync(a->weights, weights, wbytes, cudaMemcpyHostToDevice, c->stream));
    AirLaunch l;
    const uint8_t* dp = (const uint8_t*)a->tables;
    l.v.n = a->n; l.v.rot = a->rot; l.v.nr = a->nr; l.v.nc = a->nc;
    l.v.bq = (const fe*)bq; l.v.bq_stride = bq_stride; l.v.bq_inst = bq_inst;
    l.v.rnd = (const fe*)rnd; l.v.rnd_inst = rnd_inst;
    l.v.zb = a->zb; l.v.ib = a->ib; l.v.ib_inst = (uint64_t)a->nr * a->n;
    l.v.tz_inv_m = a->tz_inv_m;
    l.v.groups = (const AirGroup*)(dp + a->o_groups); l.v.group_begin = (const uint32_t*)(dp + a->o_begin);
    l.v.coefs = (const fe*)(dp + a->o_coef); l.v.weights = a->weights; l.v.nw = a->nw;
    l.v.shifts = (const uint32_t*)(dp + a->o_shift);
    l.v.tq_out = (fe*)tq_out; l.v.tq_inst = tq_inst;
    ZKB_TRY(get_pow_table(c, a->omega, ilog2_u64(a->n), &l.omega_pow));
    l.offset_m = fe_to_mont(a->offset);
    l.out = (fe*)out; l.out_inst = out_inst;
    {
        LaunchScope ls(c, K_ELEMENTWISE);
        k_air_combine<<<dim3((unsigned)((a->n + 127) / 128), (unsigned)batch), 128, 0, c->stream>>>(l);
    }
    ZKB_CUDA(c, cudaGetLastError());
    ZKB_CUDA(c, ctx_stream_sync(c));           // `weights` is the caller's (pageable) host memory
    return 0;
}

int zkb_air_combination(zkb_ctx* c, const zkb_air_desc* d, const void* bq_codewords, size_t bq_stride,
                        const void* randomizer_codeword, void* combined_out, void* tq_out) {
    if (!c || !d || !bq_codewords || !randomizer_codeword || !combined_out) return ZKB_ERR_ARG;
    if (!d->boundary_interpolants || !d->boundary_interpolant_lens || !d->weights) return ZKB_ERR_ARG;
    const uint32_t nr = d->num_registers;
    if (nr > 1 && bq_stride < d->domain_length) return set_err(c, ZKB_ERR_ARG, "air_combination: bq_stride shorter than the codewords");
    if (!is_device_ptr(bq_codewords) || !is_device_ptr(randomizer_codeword) || !is_device_ptr(combined_out) || (tq_out && !is_device_ptr(tq_out)))
        return set_err(c, ZKB_ERR_ARG, "air_combination takes device codewords (they are the outputs of zkb_coset_lde on the device)");
    zkb_air_shape sh;
    memcpy(sh.offset, d->offset, 16); memcpy(sh.omega, d->omega, 16);
    sh.domain_length = d->domain_length; sh.expansion_factor = d->expansion_factor;
    sh.num_registers = nr; sh.num_constraints = d->num_constraints;
    sh.term_counts = d->term_counts; sh.coefs = d->coefs; sh.exps = d->exps;
    sh.boundary_zerofiers = d->boundary_zerofiers; sh.boundary_zerofier_lens = d->boundary_zerofier_lens;
    sh.transition_zerofier = d->transition_zerofier; sh.transition_zerofier_len = d->transition_zerofier_len;
    sh.shifts = d->shifts;
    zkb_air* a = nullptr;
    ZKB_TRY(zkb_air_create(c, &sh, &a));
    size_t ilen = 1;
    for (uint32_t s = 0; s < nr; s++) ilen = ilen > d->boundary_interpolant_lens[s] ? ilen : d->boundary_interpolant_lens[s];
    std::vector<fe> interp((size_t)nr * ilen, fe_zero());
    for (uint32_t s = 0; s < nr; s++)
        if (d->boundary_interpolant_lens[s]) memcpy(&interp[(size_t)s * ilen], d->boundary_interpolants[s], d->boundary_interpolant_lens[s] * sizeof(fe));
    int rc = zkb_air_set_interpolants(a, 1, interp.data(), ilen);
    if (rc == 0) rc = zkb_air_combine(a, 1, d->weights, bq_codewords, bq_stride, 0, randomizer_codeword, 0, combined_out, 0, tq_out, 0);
    zkb_air_free(a);
    return rc;
}

}  // extern "C"
