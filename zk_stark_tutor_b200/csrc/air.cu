// air.cu - zkb_air_combination: the Stark prover's transition quotients and nonlinear combination in evaluation
// form on the FRI coset (see air.cuh for the identity and the reference lines it replaces).
// One thread per domain point; every input is a codeword already in HBM (the committed boundary-quotient and
// randomizer codewords) or a small table (grouped constraint coefficients, broadcast through L1/L2).
#include "air.cuh"
#include "ctx.hpp"
#include "ntt.cuh"

namespace zkb {

struct AirLaunch {
    AirView v;
    DevPow omega_pow;     // omega^i * R
    fe offset_m;          // offset * R
    uint32_t* flag;       // set to 1 if a division by zero was met
};

__global__ void __launch_bounds__(128) k_air_combine(AirLaunch a, fe* __restrict__ out) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.v.n) return;
    const fe w_m = fe_montmul(fe_ldg(a.omega_pow.hi + (i >> a.omega_pow.lo_bits)), fe_ldg(a.omega_pow.lo + (i & ((1ull << a.omega_pow.lo_bits) - 1))));
    const fe x_m = fe_montmul(w_m, a.offset_m);               // offset * omega^i * R
    bool dz = false;
    fe_store(out + i, air_point(a.v, i, x_m, &dz));
    if (dz) atomicOr(a.flag, 1u);
}

}  // namespace zkb

using namespace zkb;

extern "C" int zkb_air_combination(zkb_ctx* c, const zkb_air_desc* d, const void* bq_codewords, size_t bq_stride,
                                   const void* randomizer_codeword, void* combined_out, void* tq_out) {
    if (!c || !d || !bq_codewords || !randomizer_codeword || !combined_out) return ZKB_ERR_ARG;
    const uint64_t n = d->domain_length;
    const uint32_t nr = d->num_registers, nc = d->num_constraints;
    if (n < 2 || (n & (n - 1))) return set_err(c, ZKB_ERR_ARG, "air_combination: domain length %llu is not a power of two", (unsigned long long)n);
    if (nr == 0 || 2 * nr > (uint32_t)AIR_MAX_STATE) return set_err(c, ZKB_ERR_ARG, "air_combination: 1..%d registers supported", AIR_MAX_STATE / 2);
    if (d->expansion_factor == 0 || d->expansion_factor >= n) return set_err(c, ZKB_ERR_ARG, "air_combination: bad expansion factor");
    if ((nc && (!d->term_counts || !d->coefs || !d->exps)) || !d->boundary_zerofiers || !d->boundary_zerofier_lens || !d->boundary_interpolants ||
        !d->boundary_interpolant_lens || !d->transition_zerofier || !d->weights || !d->shifts)
        return ZKB_ERR_ARG;
    if (nr > 1 && bq_stride < n) return set_err(c, ZKB_ERR_ARG, "air_combination: bq_stride shorter than the codewords");
    if (!is_device_ptr(bq_codewords) || !is_device_ptr(randomizer_codeword) || !is_device_ptr(combined_out) || (tq_out && !is_device_ptr(tq_out)))
        return set_err(c, ZKB_ERR_ARG, "air_combination takes device codewords (they are the outputs of zkb_coset_lde on the device)");
    ZKB_CUDA(c, cudaSetDevice(c->device));

    // constraint terms -> groups (host), uploaded with the weights and shifts in one staging buffer
    AirTables tab;
    if (air_group_terms(nc, nr, d->term_counts, (const fe*)d->coefs, d->exps, &tab) != 0)
        return set_err(c, ZKB_ERR_ARG, "air_combination: constraint exponents out of range");
    const uint32_t nw = 1 + 2 * nc + 2 * nr, ns = nc + nr;
    std::vector<uint32_t> shifts(ns);
    for (uint32_t k = 0; k < ns; k++) {
        if (d->shifts[k] > 0xFFFFFFFFull) return set_err(c, ZKB_ERR_ARG, "air_combination: shift out of range");
        shifts[k] = (uint32_t)d->shifts[k];
    }
    // small polynomials -> codewords on the coset: one batched LDE of 2*nr + 1 zero-padded columns
    size_t plen = d->transition_zerofier_len;
    for (uint32_t s = 0; s < nr; s++) {
        plen = plen > d->boundary_zerofier_lens[s] ? plen : d->boundary_zerofier_lens[s];
        plen = plen > d->boundary_interpolant_lens[s] ? plen : d->boundary_interpolant_lens[s];
    }
    if (plen == 0 || plen > n) return set_err(c, ZKB_ERR_TOO_LONG, "air_combination: zerofier / interpolant longer than the domain");
    const uint32_t ncols = 2 * nr + 1;
    auto align16 = [](size_t x) { return (x + 15) & ~(size_t)15; };
    const size_t o_poly = 0, o_coef = o_poly + ncols * plen * sizeof(fe), o_w = o_coef + tab.coefs.size() * sizeof(fe),
                 o_groups = o_w + nw * sizeof(fe), o_begin = align16(o_groups + tab.groups.size() * sizeof(AirGroup)),
                 o_shift = align16(o_begin + tab.group_begin.size() * 4), o_flag = align16(o_shift + ns * 4), total = align16(o_flag + 4);
    std::vector<uint8_t> stage(total, 0);
    for (uint32_t s = 0; s < nr; s++) {                       // columns: zerofiers, interpolants, transition zerofier
        if (d->boundary_zerofier_lens[s]) memcpy(&stage[o_poly + (size_t)s * plen * sizeof(fe)], d->boundary_zerofiers[s], d->boundary_zerofier_lens[s] * sizeof(fe));
        if (d->boundary_interpolant_lens[s]) memcpy(&stage[o_poly + (size_t)(nr + s) * plen * sizeof(fe)], d->boundary_interpolants[s], d->boundary_interpolant_lens[s] * sizeof(fe));
    }
    memcpy(&stage[o_poly + (size_t)(2 * nr) * plen * sizeof(fe)], d->transition_zerofier, d->transition_zerofier_len * sizeof(fe));
    if (!tab.coefs.empty()) memcpy(&stage[o_coef], tab.coefs.data(), tab.coefs.size() * sizeof(fe));
    memcpy(&stage[o_w], d->weights, nw * sizeof(fe));
    if (!tab.groups.empty()) memcpy(&stage[o_groups], tab.groups.data(), tab.groups.size() * sizeof(AirGroup));
    memcpy(&stage[o_begin], tab.group_begin.data(), tab.group_begin.size() * 4);
    memcpy(&stage[o_shift], shifts.data(), ns * 4);

    DevBuf dstage, dcw;
    ZKB_TRY(dstage.alloc(c, total));
    ZKB_TRY(dcw.alloc(c, (size_t)ncols * n * sizeof(fe)));
    ZKB_CUDA(c, cudaMemcpyAsync(dstage.p, stage.data(), total, cudaMemcpyHostToDevice, c->stream));
    const uint8_t* dp = (const uint8_t*)dstage.p;
    NttOpts o;
    o.has_scale = true;
    o.scale_base = h_load(d->offset);
    ZKB_TRY(ntt_exec(c, h_load(d->omega), (const fe*)(dp + o_poly), plen, plen, (fe*)dcw.p, n, ncols, ilog2_u64(n), o));

    AirLaunch a;
    a.v.n = n; a.v.rot = d->expansion_factor; a.v.nr = nr; a.v.nc = nc;
    a.v.bq = (const fe*)bq_codewords; a.v.bq_stride = bq_stride;
    a.v.rnd = (const fe*)randomizer_codeword;
    a.v.zb = (const fe*)dcw.p; a.v.ib = (const fe*)dcw.p + (size_t)nr * n; a.v.tz = (const fe*)dcw.p + (size_t)(2 * nr) * n;
    a.v.groups = (const AirGroup*)(dp + o_groups); a.v.group_begin = (const uint32_t*)(dp + o_begin);
    a.v.coefs = (const fe*)(dp + o_coef); a.v.weights = (const fe*)(dp + o_w); a.v.shifts = (const uint32_t*)(dp + o_shift);
    a.v.tq_out = (fe*)tq_out;
    ZKB_TRY(get_pow_table(c, h_load(d->omega), ilog2_u64(n), &a.omega_pow));
    a.offset_m = fe_to_mont(h_load(d->offset));
    a.flag = (uint32_t*)(dstage.p) + o_flag / 4;
    {
        LaunchScope ls(c, K_ELEMENTWISE);
        k_air_combine<<<(unsigned)((n + 127) / 128), 128, 0, c->stream>>>(a, (fe*)combined_out);
    }
    ZKB_CUDA(c, cudaGetLastError());
    uint32_t flag = 0;
    ZKB_CUDA(c, cudaMemcpyAsync(&flag, a.flag, 4, cudaMemcpyDeviceToHost, c->stream));
    ZKB_CUDA(c, cudaStreamSynchronize(c->stream));           // `stage` (pageable) and `flag` are host locals
    if (flag) return set_err(c, ZKB_ERR_DIV_ZERO, "air_combination: the transition zerofier vanishes on the FRI domain (divide by zero)");
    return 0;
}
