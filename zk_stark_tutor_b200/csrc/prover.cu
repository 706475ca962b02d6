// prover.cu - Stark::prove for a batch of instances of ONE AIR as a single C call (host orchestration only; every
// stage is one of the batched device entry points of air.cu / stark.cu / batch.cu).
//
// Reference: src/stark/stark.rs:276-563 (`prove`): randomize the trace (:286-301), interpolate it (:303-326), boundary
// quotients (:331-360), commit them (:366-386), randomizer polynomial (:424-445), weights from the transcript
// (:447-450), transition quotients + nonlinear combination (:388-519) with the degree check (:451-464), FRI::prove
// (:522), the quadrupled indices (:524-543) and the Value / Path openings (:546-560).
//
// zk_stark_tutor_b200/stark.py `prove_batch` does the same with one ctypes call per stage; what it does in between
// (packing the traces register-major, drawing the randomizers, Lagrange sums for the boundary interpolants, weights,
// index arithmetic) is ~3.4 ms of Python per batch of 32 RPSSS signatures under the GIL against 1.25 ms of kernel
// time, and it is what limits several batches in flight on host threads.  Here that glue is C++ and the GIL is released
// for the whole proof.
#include <stdlib.h>
#include <string.h>
#include <sys/random.h>
#include <stdio.h>
#include <algorithm>
#include <chrono>
#include <thread>
#include <vector>
#include "ctx.hpp"

using namespace zkb;

namespace {

// count field elements from OS entropy: 16 random bytes each, reduced once (p > 2^127) - as random as Field::sample of 17 bytes
int os_random_elements(fe* out, size_t count) {
    uint8_t* p = reinterpret_cast<uint8_t*>(out);
    size_t left = count * sizeof(fe);
    while (left) {
        ssize_t got = getrandom(p, left, 0);
        if (got < 0) return -1;
        p += got; left -= (size_t)got;
    }
    for (size_t i = 0; i < count; i++)
        if (fe_ge_p(out[i])) {
            fe pm; pm.v[0] = P0; pm.v[1] = 0; pm.v[2] = 0; pm.v[3] = P3;
            out[i] = fe_sub(out[i], pm);               // plain difference: out[i] >= p, no wrap
        }
    return 0;
}

// fn(lo, hi) over [0, n) split across up to `threads` host threads (the context's zkb_ctx_assembly_threads setting)
template <typename F>
bool split_over_threads(size_t n, size_t threads, F fn) {
    threads = std::max<size_t>(1, std::min(threads, n));
    if (threads == 1) return fn(0, n);
    std::vector<std::thread> pool;
    std::vector<char> ok(threads, 0);
    const size_t per = (n + threads - 1) / threads;
    for (size_t t = 0; t < threads; t++)
        pool.emplace_back([&, t]() { const size_t lo = t * per, hi = std::min(n, lo + per); try { ok[t] = lo >= hi || fn(lo, hi); } catch (...) { ok[t] = 0; } });
    for (auto& th : pool) th.join();
    return std::all_of(ok.begin(), ok.end(), [](char x) { return x != 0; });
}

struct TreeGuard {
    std::vector<zkb_tree*> t;
    ~TreeGuard() { for (size_t i = t.size(); i-- > 0;) if (t[i]) zkb_merkle_free(t[i]); }    // tree 0 owns the shared arena: free it last
};

}  // namespace

extern "C" {

static int zkb_stark_prove_batch_impl(zkb_ctx* c, zkb_air* air, const zkb_stark_shape* sh, size_t B, const void* traces_v,
                                      const void* boundary_values_v, const void* randomness_v, zkb_ps* const* ps, uint64_t* proof_len_out) {
    if (!c || !air || !sh || !traces_v || !ps || B == 0 || (sh->num_boundary && (!boundary_values_v || !sh->boundary_register || !sh->lagrange)))
        return ZKB_ERR_ARG;
    const size_t nr = sh->num_registers, nc = sh->num_constraints, t0 = sh->trace_length, nrand = sh->num_randomizers;
    const size_t L = t0 + nrand, n = sh->fri.domain_length, K = nr + 1, n_tr = nrand * nr, n_rp = sh->rnd_poly_len, nb = sh->num_boundary;
    const size_t ncc = sh->fri.num_colinearity_tests, ef = sh->fri.expansion_factor;
    if (nr == 0 || n == 0 || (n & (n - 1)) || L > sh->omicron_order || n_rp == 0 || n_rp > n)
        return set_err(c, ZKB_ERR_ARG, "stark_prove_batch: inconsistent shape");
    for (size_t b = 0; b < B; b++) if (!ps[b]) return ZKB_ERR_ARG;
    const fe* traces = (const fe*)traces_v;
    const fe* bvals = (const fe*)boundary_values_v;
    // ZKB_PROVER_TIMING=1: host wall time between the stages (where the calling thread waits), printed to stderr per call
    const bool timing = getenv("ZKB_PROVER_TIMING") != nullptr;
    std::vector<std::pair<const char*, double>> marks;
    auto t_prev = std::chrono::steady_clock::now();
    auto mark = [&](const char* name) {
        if (!timing) return;
        auto now = std::chrono::steady_clock::now();
        marks.push_back({name, std::chrono::duration<double, std::milli>(now - t_prev).count()});
        t_prev = now;
    };

    // ---- randomness (stark.rs:286-301, 424-432): per instance n_tr randomizer values (row by row, register by register) + n_rp coefficients
    std::vector<fe> rnd_own;
    const fe* rnd = (const fe*)randomness_v;
    if (!rnd) {
        rnd_own.resize(B * (n_tr + n_rp));
        fe* dst = rnd_own.data();                                     // ~44 KB of OS entropy per instance (0.85 GB/s per thread)
        if (!split_over_threads(B, c->assembly_threads, [&](size_t lo, size_t hi) { return os_random_elements(dst + lo * (n_tr + n_rp), (hi - lo) * (n_tr + n_rp)) == 0; }))
            return set_err(c, ZKB_ERR_ARG, "stark_prove_batch: getrandom failed");
        rnd = rnd_own.data();
    }
    mark("randomness");
    // ---- randomized traces, register-major: column s * B + b of the batched calls is register s of instance b
    std::vector<fe> values(nr * B * L);
    for (size_t b = 0; b < B; b++) {
        const fe* tr = traces + b * t0 * nr;
        const fe* dr = rnd + b * (n_tr + n_rp);
        for (size_t s = 0; s < nr; s++) {
            fe* col = values.data() + (s * B + b) * L;
            for (size_t r = 0; r < t0; r++) col[r] = tr[r * nr + s];
            for (size_t r = 0; r < nrand; r++) col[t0 + r] = dr[r * nr + s];
        }
    }
    DevBuf tcw, cws, tq;
    ZKB_TRY(tcw.alloc(c, nr * B * n * sizeof(fe)));                  // trace codewords
    ZKB_TRY(cws.alloc(c, (nr + 2) * B * n * sizeof(fe)));            // planes: boundary quotients | randomizer | combination
    fe* d_cws = (fe*)cws.p;
    ZKB_TRY(zkb_trace_lde_batch(c, sh->omicron, sh->omicron_order, L, sh->fri.omega, n, sh->fri.offset, values.data(), L, nr * B, tcw.p, n, nullptr));

    mark("pack + trace_lde");
    // ---- boundary interpolants (stark.rs:215-243): sum_i value_i * basis_i, the unique interpolant through a handful of points
    std::vector<size_t> m(nr, 0), lag_off(nr, 0);
    for (size_t j = 0; j < nb; j++) { if (sh->boundary_register[j] >= nr) return set_err(c, ZKB_ERR_ARG, "stark_prove_batch: boundary register out of range"); m[sh->boundary_register[j]]++; }
    size_t ilen = 1, off = 0;
    for (size_t s = 0; s < nr; s++) { ilen = std::max(ilen, m[s]); lag_off[s] = off; off += m[s] * m[s]; }
    const fe* lag = (const fe*)sh->lagrange;
    std::vector<fe> interp(B * nr * ilen, fe_zero());
    for (size_t b = 0; b < B; b++) {
        std::vector<size_t> seen(nr, 0);
        for (size_t j = 0; j < nb; j++) {
            const size_t s = sh->boundary_register[j], i = seen[s]++;
            const fe v = bvals[b * nb + j];
            fe* poly = interp.data() + (b * nr + s) * ilen;
            const fe* basis = lag + lag_off[s] + i * m[s];
            for (size_t k = 0; k < m[s]; k++) poly[k] = fe_add(poly[k], h_mul(v, basis[k]));
        }
    }
    ZKB_TRY(zkb_air_set_interpolants(air, B, interp.data(), ilen));
    ZKB_TRY(zkb_air_boundary_quotients(air, B, tcw.p, B * n, n, d_cws, B * n, n));                       // stark.rs:331-360
    mark("interpolants + boundary quotients");
    // ---- randomizer polynomial (stark.rs:424-436)
    std::vector<fe> rpoly(B * n_rp);
    for (size_t b = 0; b < B; b++) memcpy(rpoly.data() + b * n_rp, rnd + b * (n_tr + n_rp) + n_tr, n_rp * sizeof(fe));
    ZKB_TRY(zkb_coset_lde_batch(c, sh->fri.omega, n, sh->fri.offset, rpoly.data(), n_rp, n_rp, d_cws + nr * B * n, n, B));
    mark("randomizer lde");
    // ---- commits (stark.rs:366-386, 441-445): tree t * B + b belongs to proof b
    TreeGuard trees;
    trees.t.assign(K * B, nullptr);
    std::vector<zkb_ps*> ps_of_tree(K * B);
    for (size_t t = 0; t < K; t++) for (size_t b = 0; b < B; b++) ps_of_tree[t * B + b] = ps[b];
    ZKB_TRY(zkb_merkle_build_batch(c, d_cws, n, n, K * B, trees.t.data(), ps_of_tree.data()));
    mark("merkle_build_batch");
    // ---- weights (stark.rs:447-450; all weights of a proof are equal)
    const size_t nw = 1 + 2 * nc + 2 * nr;
    std::vector<uint8_t> weights(B * nw * 16), fs(sh->proof_bytes ? sh->proof_bytes : 32);
    for (size_t b = 0; b < B; b++) {
        ZKB_TRY(zkb_ps_fiat_shamir(ps[b], fs.size(), fs.data()));
        uint8_t w[16];
        zkb_field_sample(fs.data(), fs.size(), w);
        for (size_t i = 0; i < nw; i++) memcpy(weights.data() + (b * nw + i) * 16, w, 16);
    }
    mark("weights");
    const bool check = sh->tq_degree_bounds != nullptr;
    if (check) ZKB_TRY(tq.alloc(c, B * nc * n * sizeof(fe)));
    ZKB_TRY(zkb_air_combine(air, B, weights.data(), d_cws, B * n, n, d_cws + nr * B * n, n, d_cws + (nr + 1) * B * n, n, check ? tq.p : nullptr, nc * n));
    if (check) {                                                                                        // stark.rs:451-464
        std::vector<int64_t> degs(B * nc);
        ZKB_TRY(zkb_coset_degree_batch(c, sh->fri.omega, tq.p, n, n, B * nc, degs.data()));
        for (size_t b = 0; b < B; b++)
            for (size_t k = 0; k < nc; k++) {
                if (degs[b * nc + k] < 0) return set_err(c, ZKB_ERR_DEGREE, "Failed to get degree of transition quotient");
                if (degs[b * nc + k] != sh->tq_degree_bounds[k]) return set_err(c, ZKB_ERR_DEGREE, "transition quotient degrees do not match with expectation");
            }
    }
    mark("air_combine + degree check");
    // ---- FRI (stark.rs:522) and the quadrupled, sorted indices (stark.rs:524-543)
    std::vector<uint64_t> top(B * ncc);
    ZKB_TRY(zkb_fri_prove_batch(c, &sh->fri, d_cws + (nr + 1) * B * n, n, n, B, ps, top.data()));
    mark("fri_prove_batch");
    const size_t k = 4 * ncc;
    std::vector<uint64_t> idx(K * B * k);
    for (size_t b = 0; b < B; b++) {
        uint64_t* q = idx.data() + b * k;
        for (size_t i = 0; i < ncc; i++) { q[i] = top[b * ncc + i]; q[ncc + i] = (top[b * ncc + i] + ef) % n; }
        for (size_t i = 0; i < 2 * ncc; i++) q[2 * ncc + i] = (q[i] + n / 2) % n;
        std::sort(q, q + k);
    }
    for (size_t t = 1; t < K; t++) memcpy(idx.data() + t * B * k, idx.data(), B * k * sizeof(uint64_t));
    ZKB_TRY(zkb_merkle_open_ps_batch(trees.t.data(), K * B, idx.data(), k, ps_of_tree.data()));         // stark.rs:546-560
    mark("indices + merkle_open_ps_batch");
    if (timing) { for (auto& m : marks) fprintf(stderr, "  %-36s %8.3f ms\n", m.first, m.second); }
    if (proof_len_out) for (size_t b = 0; b < B; b++) proof_len_out[b] = zkb_ps_digest(ps[b], nullptr, 0);
    // the device buffers go back to the context's stream-ordered pool; everything queued on the stream has been waited for by the openings
    return 0;
}

int zkb_stark_prove_batch(zkb_ctx* c, zkb_air* air, const zkb_stark_shape* sh, size_t batch, const void* traces, const void* boundary_values,
                          const void* randomness, zkb_ps* const* ps, uint64_t* proof_len_out) {
    if (!c) return ZKB_ERR_ARG;
    ZKB_ABI_GUARD(c, return zkb_stark_prove_batch_impl(c, air, sh, batch, traces, boundary_values, randomness, ps, proof_len_out);)
}

}  // extern "C"
