// batch.cu - batches of small independent instances (RPSSS-shaped proofs: 4096-point FRI domain).
//
// At these sizes one Merkle tree / FRI layer is a few dozen warps: a single instance leaves the GPU
// idle and every kernel is latency-bound.  The reference runs one proof after the other
// (src/rpsss.rs:70-87 -> src/stark/stark.rs:276-563); independent proofs are independent units
// (SURVEY.md 8e.1), so here `batch` of them advance in lockstep: every kernel launch carries all
// instances (blockIdx.y = instance, identical buffer layouts at a fixed stride), the per-round
// Fiat-Shamir hop handles all roots at once, and the proof-stream objects of the instances are
// assembled by a few host threads.  Results per instance are byte-identical to the single-instance
// entry points (zkb_merkle_build, zkb_fri_prove, zkb_merkle_open_ps): same kernels, same order.
//
//   zkb_merkle_build_batch    MerkleRoot::commit per codeword   (stark.rs:373-381, 431-436)
//   zkb_fri_prove_batch       FRI::prove per codeword           (fri.rs:210-248)
//   zkb_merkle_open_ps_batch  the Value + Path opening loop     (stark.rs:546-560)
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <atomic>
#include <thread>
#include <vector>
#include "merkle.cuh"
#include "hosthash.hpp"
#include "keccak.cuh"
#include "fri_tail.cuh"

namespace zkb {

// Leafs triples (fri.rs:189-195) for every instance: out[b][s] = (cur[a], cur[a + half], nxt[a]), a = idx[b][s]
__global__ void k_gather3_batch(const fe* cur, uint64_t cur_stride, const fe* nxt, uint64_t nxt_stride, uint64_t half,
                                const uint64_t* idx, uint32_t k, fe* out) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= k) return;
    const uint32_t b = blockIdx.y;
    cur += (uint64_t)b * cur_stride;
    nxt += (uint64_t)b * nxt_stride;
    const uint64_t a = idx[(uint64_t)b * k + s];
    fe* o = out + ((uint64_t)b * k + s) * 3;
    fe_store(o, fe_ldg(cur + a));
    fe_store(o + 1, fe_ldg(cur + a + half));
    fe_store(o + 2, fe_ldg(nxt + a));
}
__global__ void k_gather_vals_batch(const fe* vals, uint64_t vals_stride, const uint64_t* idx, uint32_t k, fe* out) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= k) return;
    const uint32_t b = blockIdx.y;
    fe_store(out + (uint64_t)b * k + s, fe_ldg(vals + (uint64_t)b * vals_stride + idx[(uint64_t)b * k + s]));
}

// run fn(i) for i in [0, n) on up to `threads` host threads (per-instance proof-stream assembly)
// cap = the context's zkb_ctx_assembly_threads setting: a caller that already keeps several batches in flight on its own threads
// wants few (measured, 8 batches of 32 signatures in flight on a 16-core host: 1 thread 9,724/s, 4: 9,666/s, 16: 8,339/s)
template <typename F>
static bool parallel_for(size_t n, size_t cap, F fn) {
    size_t threads = std::min<size_t>(std::min<size_t>(n, cap ? cap : 1), std::max(1u, std::thread::hardware_concurrency()));
    std::atomic<bool> ok(true);
    auto guarded = [&](size_t i) { try { fn(i); } catch (...) { ok = false; } };     // an exception must not leave a worker thread (std::terminate)
    if (threads <= 1) { for (size_t i = 0; i < n; i++) guarded(i); return ok; }
    std::vector<std::thread> pool;
    try {
        for (size_t t = 0; t < threads; t++)
            pool.emplace_back([=, &guarded]() { for (size_t i = t; i < n; i += threads) guarded(i); });
    } catch (...) { ok = false; }                                                      // thread creation failed: the started ones still finish
    for (auto& th : pool) th.join();
    return ok;
}

}  // namespace zkb

using namespace zkb;

extern "C" {

static int zkb_merkle_build_batch_impl(zkb_ctx* c, const void* vals, size_t n, size_t stride, size_t batch, zkb_tree** trees_out, zkb_ps* const* ps) {
    if (!c || !vals || !trees_out || batch == 0) return ZKB_ERR_ARG;
    for (size_t b = 0; b < batch; b++) trees_out[b] = nullptr;
    if (n == 0 || (n & (n - 1))) return set_err(c, ZKB_ERR_NOT_POW2, "Leafs len must be power of two (got %zu)", n);
    if (!is_device_ptr(vals)) return set_err(c, ZKB_ERR_ARG, "merkle_build_batch takes a device pointer");
    if (stride < n) return set_err(c, ZKB_ERR_ARG, "merkle_build_batch: stride shorter than a codeword");
    if (batch > ZKB_MAX_BATCH) return set_err(c, ZKB_ERR_ARG, "merkle_build_batch: at most %d trees per call", ZKB_MAX_BATCH);
    ZKB_CUDA(c, cudaSetDevice(c->device));
    TreeLayout L;
    L.init(ilog2_u64(n));
    if (L.top != 0) return set_err(c, ZKB_ERR_ARG, "merkle_build_batch: trees of more than 2^%u leaves fill the GPU on their own; use zkb_merkle_build", (unsigned)ZKB_TREE_LEAF_LOG);
    const size_t tree_bytes = (L.total_nodes * 64 + 255) & ~(size_t)255;
    uint8_t* arena = nullptr;
    ZKB_CUDA(c, dev_alloc(c, (void**)&arena, tree_bytes * batch));
    BatchArgs ba;
    ba.batch = (uint32_t)batch; ba.vals_stride = stride; ba.nodes_stride = tree_bytes;
    int rc = merkle_build_levels_batch(c, (const fe*)vals, nullptr, n, L, arena, ba);
    std::vector<uint8_t> roots(batch * 64);
    if (rc == 0) rc = merkle_batch_roots(c, L, arena, ba, roots.data());
    if (rc != 0) { dev_free(c, arena); return rc; }
    for (size_t b = 0; b < batch; b++) {
        zkb_tree* t = new zkb_tree();
        t->ctx = c; t->n = n; t->layout = L;
        t->nodes = arena + b * tree_bytes;
        t->owns_nodes = b == 0;                       // the arena is released with the FIRST tree: free it last
        t->vals = (const fe*)vals + b * stride;
        memcpy(t->root, roots.data() + 64 * b, 64);
        trees_out[b] = t;
        if (ps && ps[b]) zkb_ps_push_root(ps[b], t->root, 64);       // stark.rs:380-381, 436: push Root(commit(codeword))
    }
    return 0;
}

int zkb_merkle_build_batch(zkb_ctx* c, const void* vals, size_t n, size_t stride, size_t batch, zkb_tree** trees_out, zkb_ps* const* ps) {
    if (!c) return ZKB_ERR_ARG;
    ZKB_ABI_GUARD(c, return zkb_merkle_build_batch_impl(c, vals, n, stride, batch, trees_out, ps);)
}

static int zkb_fri_prove_batch_impl(zkb_ctx* c, const zkb_fri_params* p, const void* codewords, size_t n, size_t stride, size_t batch,
                        zkb_ps* const* ps, uint64_t* top_indices_out) {
    if (!c || !p || !codewords || !ps || !top_indices_out || batch == 0) return ZKB_ERR_ARG;
    if (p->domain_length != n) return set_err(c, ZKB_ERR_LENGTH, "Length of the domain doesnt match the length of initial codeword");
    if (n < 2 || (n & (n - 1))) return set_err(c, ZKB_ERR_NOT_POW2, "Leafs len must be power of two (got %zu)", n);
    if (!is_device_ptr(codewords)) return set_err(c, ZKB_ERR_ARG, "fri_prove_batch takes a device pointer");
    if (stride < n || batch > ZKB_MAX_BATCH) return set_err(c, ZKB_ERR_ARG, "fri_prove_batch: bad stride / batch");
    const uint64_t R = zkb_fri_num_rounds(p), ncc = p->num_colinearity_tests;
    if (R < 2) return set_err(c, ZKB_ERR_ROUNDS, "FRI::prove needs at least two rounds (fri.rs:225)");
    ZKB_CUDA(c, cudaSetDevice(c->device));
    const fe omega = h_load(p->omega), offset = h_load(p->offset);
    if (!fe_eq(h_pow(omega, n), fe_from_u32(1)) || fe_is_zero(omega))
        return set_err(c, ZKB_ERR_ROOT_ORDER, "error in commit: omega does not have the right order!");
    if (fe_is_zero(offset)) return set_err(c, ZKB_ERR_DIV_ZERO, "divide by zero");
    // ---- per-instance arena: folded codewords (rounds 1..R-1) and all trees; identical for every instance
    std::vector<uint64_t> len(R);
    std::vector<TreeLayout> lay(R);
    std::vector<size_t> cw_off(R, 0), node_off(R);
    size_t inst_bytes = 0;
    for (uint64_t r = 0; r < R; r++) {
        len[r] = n >> r;
        lay[r].init(ilog2_u64(len[r]));
        if (lay[r].top != 0) return set_err(c, ZKB_ERR_ARG, "fri_prove_batch: codewords of more than 2^%u values fill the GPU on their own; use zkb_fri_prove", (unsigned)ZKB_TREE_LEAF_LOG);
        if (r > 0) { cw_off[r] = inst_bytes; inst_bytes += len[r] * sizeof(fe); }
        node_off[r] = inst_bytes;
        inst_bytes += lay[r].total_nodes * 64;
    }
    if (ncc > len[R - 1]) return set_err(c, ZKB_ERR_ARG, "Cannot sample more indices than available in the last codeword");
    inst_bytes = (inst_bytes + 255) & ~(size_t)255;
    DevBuf arena, kk_dev;
    ZKB_TRY(arena.alloc(c, inst_bytes * batch));
    ZKB_TRY(kk_dev.alloc(c, batch * sizeof(fe)));
    uint8_t* A = (uint8_t*)arena.p;
    auto cw_ptr = [&](uint64_t r) -> const fe* { return r == 0 ? (const fe*)codewords : (const fe*)(A + cw_off[r]); };
    auto cw_stride = [&](uint64_t r) -> uint64_t { return r == 0 ? (uint64_t)stride : (uint64_t)(inst_bytes / sizeof(fe)); };
    DevPow winv_tab;
    const fe omega_inv0 = h_inv(omega);
    ZKB_TRY(get_pow_table(c, omega_inv0, ilog2_u64(n), &winv_tab));
    std::vector<fe> inv_offset(R);
    { fe io = h_inv(offset); for (uint64_t r = 0; r < R; r++) { inv_offset[r] = io; io = h_mul(io, io); } }

    // ---- commit phase (fri.rs:115-172), all instances in lockstep
    const bool host_path = getenv("ZKB_HOST_ASSEMBLY") != nullptr;     // the round-1 path: host Fiat-Shamir hop per round, host object framing
    std::vector<uint8_t> roots(batch * 64);
    std::vector<fe> alpha(batch, fe_zero()), kk(batch);
    fe omega_inv_r = omega_inv0;
    const bool dev_fs = !host_path && R <= ZKB_FS_MAX_ROUNDS;
    if (dev_fs) {
        // device Fiat-Shamir: one sponge per instance (FsDev[b]); the tree kernel of round r appends Root, draws alpha and leaves
        // alpha / offset_r for the fold of round r + 1 - no copy, no synchronisation and no host hashing between the rounds
        ZKB_TRY(ensure_fs_dev(c, batch));
        FsDev* fs = (FsDev*)c->fs_dev;
        const size_t head = offsetof(FsDev, inv_off_m2) + R * sizeof(fe);
        uint8_t* stage = nullptr;
        ZKB_TRY(host_scratch_reserve(c, 0, batch * head, &stage));
        const fe r2 = ZKB_FE_R2;
        std::vector<fe> inv_m2(R);
        for (uint64_t r = 0; r < R; r++) inv_m2[r] = fe_montmul(fe_to_mont(inv_offset[r]), r2);
        parallel_for(batch, c->assembly_threads, [&](size_t b) {
            FsDev* h = (FsDev*)(stage + b * head);                       // only the first `head` bytes of each FsDev are staged
            ps_export_sponge(ps[b], &h->sp);
            h->kk_m = fe_zero(); h->alpha = fe_zero();
            memcpy(h->inv_off_m2, inv_m2.data(), R * sizeof(fe));
        });
        ZKB_CUDA(c, cudaMemcpy2DAsync(fs, sizeof(FsDev), stage, head, head, batch, cudaMemcpyHostToDevice, c->stream));
        for (uint64_t r = 0; r < R; r++) {
            BatchArgs ba;
            ba.batch = (uint32_t)batch; ba.nodes_stride = inst_bytes;
            FsHook hook;
            hook.fs = fs; hook.round = (uint32_t)r; hook.want_alpha = r + 1 < R;
            if (r == 0) {
                ba.vals_stride = stride;
                ZKB_TRY(merkle_build_levels_batch(c, (const fe*)codewords, nullptr, len[0], lay[0], A + node_off[0], ba, &hook));
            } else {
                FoldArgs f;
                f.cw = cw_ptr(r - 1); f.next = (fe*)(A + cw_off[r]); f.half = len[r];
                f.winv = winv_tab; f.exp_mul = 1ull << (r - 1);
                f.kk_m = fe_zero();
                f.kk_dev = (const uint8_t*)&fs->kk_m; f.kk_stride = sizeof(FsDev);
                f.wr_inv_m = fe_to_mont(omega_inv_r);
                ba.vals_stride = cw_stride(r - 1); ba.next_stride = inst_bytes / sizeof(fe);
                ZKB_TRY(merkle_build_levels_batch(c, nullptr, &f, len[r], lay[r], A + node_off[r], ba, &hook));
                omega_inv_r = h_mul(omega_inv_r, omega_inv_r);
            }
        }
        roots.resize(batch * R * 64);
        ZKB_CUDA(c, cudaMemcpy2DAsync(roots.data(), R * 64, fs->roots, sizeof(FsDev), R * 64, batch, cudaMemcpyDeviceToHost, c->stream));
    } else {
    for (uint64_t r = 0; r < R; r++) {
        BatchArgs ba;
        ba.batch = (uint32_t)batch; ba.nodes_stride = inst_bytes;
        if (r == 0) {
            ba.vals_stride = stride;
            ZKB_TRY(merkle_build_levels_batch(c, (const fe*)codewords, nullptr, len[0], lay[0], A + node_off[0], ba));
        } else {
            for (size_t b = 0; b < batch; b++) kk[b] = fe_to_mont(h_mul(alpha[b], inv_offset[r - 1]));
            ZKB_CUDA(c, cudaMemcpyAsync(kk_dev.p, kk.data(), batch * sizeof(fe), cudaMemcpyHostToDevice, c->stream));
            FoldArgs f;
            f.cw = cw_ptr(r - 1); f.next = (fe*)(A + cw_off[r]); f.half = len[r];
            f.winv = winv_tab; f.exp_mul = 1ull << (r - 1);
            f.kk_m = fe_zero();
            f.wr_inv_m = fe_to_mont(omega_inv_r);
            ba.vals_stride = cw_stride(r - 1); ba.next_stride = inst_bytes / sizeof(fe); ba.kk_m = (const fe*)kk_dev.p;
            ZKB_TRY(merkle_build_levels_batch(c, nullptr, &f, len[r], lay[r], A + node_off[r], ba));   // fold fused with leaf hashing
            omega_inv_r = h_mul(omega_inv_r, omega_inv_r);
        }
        ZKB_TRY(merkle_batch_roots(c, lay[r], A + node_off[r], ba, roots.data()));   // also orders the kk upload before its reuse
        const bool want_alpha = r + 1 < R;
        for (size_t b = 0; b < batch; b++) {
            zkb_ps_push_root(ps[b], roots.data() + 64 * b, 64);             // fri.rs:136-137
            if (want_alpha) {
                uint8_t ch[32], a16[16];
                zkb_ps_fiat_shamir(ps[b], 32, ch);                          // fri.rs:145
                zkb_field_sample(ch, 32, a16);                              // fri.rs:146
                alpha[b] = h_load(a16);
            }
        }
    }
    }
    // ---- last codewords (fri.rs:166), top-level indices (fri.rs:223-228)
    const uint64_t last_len = len[R - 1];
    std::vector<uint8_t> last(batch * last_len * 16);
    ZKB_CUDA(c, cudaMemcpy2DAsync(last.data(), last_len * 16, cw_ptr(R - 1), cw_stride(R - 1) * sizeof(fe), last_len * 16, batch,
                                  cudaMemcpyDeviceToHost, c->stream));
    ZKB_CUDA(c, ctx_stream_sync(c));
    std::vector<uint64_t> idx(batch * ncc);
    int bad = 0;
    for (size_t b = 0; b < batch; b++) {
        if (dev_fs)
            for (uint64_t r = 0; r < R; r++) zkb_ps_push_root(ps[b], roots.data() + (b * R + r) * 64, 64);      // fri.rs:136-137, in round order
        zkb_ps_push_codeword(ps[b], last.data() + b * last_len * 16, last_len);
        uint8_t seed[32];
        zkb_ps_fiat_shamir(ps[b], 32, seed);
        if (zkb_fri_sample_indices(seed, 32, len[1], last_len, ncc, top_indices_out + b * ncc) != 0) bad = 1;
        for (uint64_t s = 0; s < ncc; s++) idx[b * ncc + s] = top_indices_out[b * ncc + s];
    }
    if (bad) return set_err(c, ZKB_ERR_ARG, "sample_indices failed");
    // ---- query phase (fri.rs:174-208, 234-246)
    bool wire = !host_path;
    for (uint64_t r = 0; r < R; r++) if (lay[r].log_n < 1 || lay[r].log_n > 30) wire = false;
    if (wire) {
        // The device writes the finished objects: per instance ONE contiguous segment [round 0: ncc Leafs, ncc x (Path a, Path b, Path c)]
        // [round 1: ...] ... in wire format (k_leafs_wire, k_open_wire), one D2H for the whole batch, one append per proof stream.
        std::vector<uint64_t> base(R, 0);
        uint64_t seg = 0;
        for (uint64_t r = 0; r + 1 < R; r++) {
            base[r] = seg;
            seg += ncc * 57 + ncc * (2 * (9 + 72ull * lay[r].log_n) + (9 + 72ull * lay[r + 1].log_n));
        }
        const uint64_t seg_pad = (seg + 15) & ~15ull;
        // indices of every round, computed up front (fri.rs:234-237): per round [ab: batch x 2 ncc][c: batch x ncc]
        const size_t per_round = batch * ncc * 3;
        std::vector<uint64_t> hidx((R - 1) * per_round + batch);
        for (uint64_t r = 0; r + 1 < R; r++) {
            const uint64_t half = len[r] / 2;
            for (auto& i : idx) i %= half;
            uint64_t* ab = hidx.data() + r * per_round;
            uint64_t* cc = ab + 2 * batch * ncc;
            for (size_t q = 0; q < batch * ncc; q++) { ab[2 * q] = idx[q]; ab[2 * q + 1] = idx[q] + half; cc[q] = idx[q]; }
        }
        uint64_t* yoff = hidx.data() + (R - 1) * per_round;
        for (size_t bi = 0; bi < batch; bi++) yoff[bi] = bi * seg_pad;
        DevBuf q;
        const size_t idx_bytes = (hidx.size() * 8 + 255) & ~(size_t)255;
        ZKB_TRY(q.alloc(c, idx_bytes + batch * seg_pad));
        uint64_t* d_idx = (uint64_t*)q.p;
        uint8_t* d_wire = (uint8_t*)q.p + idx_bytes;
        const uint64_t* d_yoff = d_idx + (R - 1) * per_round;
        ZKB_CUDA(c, cudaMemcpyAsync(d_idx, hidx.data(), hidx.size() * 8, cudaMemcpyHostToDevice, c->stream));
        for (uint64_t r = 0; r + 1 < R; r++) {
            const uint64_t half = len[r] / 2;
            const uint64_t pc = 9 + 72ull * lay[r].log_n, pn = 9 + 72ull * lay[r + 1].log_n, trip = 2 * pc + pn;
            const uint64_t* d_ab = d_idx + r * per_round;
            const uint64_t* d_c = d_ab + 2 * batch * ncc;
            ZKB_TRY(fri_leafs_wire_batch(c, cw_ptr(r), cw_stride(r), cw_ptr(r + 1), cw_stride(r + 1), half, d_c, ncc, (uint32_t)batch, d_wire, d_yoff, base[r]));
            ZKB_TRY(merkle_open_wire_batch(c, cw_ptr(r), lay[r], A + node_off[r], d_ab, 2 * ncc, (uint32_t)batch, cw_stride(r), inst_bytes,
                                           d_wire, d_yoff, 2, base[r] + ncc * 57, trip, pc, false));
            ZKB_TRY(merkle_open_wire_batch(c, cw_ptr(r + 1), lay[r + 1], A + node_off[r + 1], d_c, ncc, (uint32_t)batch, cw_stride(r + 1), inst_bytes,
                                           d_wire, d_yoff, 1, base[r] + ncc * 57 + 2 * pc, trip, 0, false));
        }
        // every segment is copied straight to its final address: the tail of its proof stream's body, which lives in pinned host
        // memory (hosthash.hpp PsBody) - one DMA per proof, no staging buffer and no host memcpy of the ~0.4 MB
        std::vector<uint8_t*> dst(batch, nullptr);
        const bool ok = parallel_for(batch, c->assembly_threads, [&](size_t bi) { dst[bi] = ps_body_extend(ps[bi], seg); });
        if (!ok) return set_err(c, ZKB_ERR_NOMEM, "out of host memory while appending to the proof streams (they are unusable now)");
        for (size_t bi = 0; bi < batch; bi++)
            ZKB_CUDA(c, cudaMemcpyAsync(dst[bi], d_wire + bi * seg_pad, seg, cudaMemcpyDeviceToHost, c->stream));
        ZKB_CUDA(c, ctx_stream_sync(c));
        for (size_t bi = 0; bi < batch; bi++) ps[bi]->has_field = true;      // Leafs carry field elements (proof_stream_enum.rs:105-112)
        return 0;
    }
    // (round-1 path, kept for ZKB_HOST_ASSEMBLY=1 and for layer shapes the wire kernels do not take: raw paths to the host,
    // objects framed by host threads)
    DevBuf q;
    const size_t max_path = (size_t)lay[0].log_n * 64;
    const size_t idx_bytes = (batch * ncc * 8 * 3 + 255) & ~(size_t)255, leaf_bytes = (batch * ncc * 48 + 255) & ~(size_t)255;
    ZKB_TRY(q.alloc(c, idx_bytes + leaf_bytes + batch * ncc * 3 * max_path));
    uint64_t* d_ab = (uint64_t*)q.p;
    uint64_t* d_c = d_ab + 2 * batch * ncc;
    fe* d_leafs = (fe*)((uint8_t*)q.p + idx_bytes);
    uint8_t* d_pab = (uint8_t*)q.p + idx_bytes + leaf_bytes;
    std::vector<uint64_t> ab(2 * batch * ncc);
    uint8_t *leafs = nullptr, *hab = nullptr, *hc = nullptr;                 // pinned: the D2H copies run at PCIe rate
    ZKB_TRY(host_scratch_reserve(c, 0, batch * ncc * 48, &leafs));
    ZKB_TRY(host_scratch_reserve(c, 1, 2 * batch * ncc * max_path, &hab));
    ZKB_TRY(host_scratch_reserve(c, 2, batch * ncc * max_path, &hc));
    for (uint64_t r = 0; r + 1 < R; r++) {
        const uint64_t half = len[r] / 2;
        const size_t d_cur = lay[r].log_n, d_nxt = lay[r + 1].log_n, pb_cur = d_cur * 64, pb_nxt = d_nxt * 64;
        for (auto& i : idx) i %= half;                                       // fri.rs:234-237
        for (size_t b = 0; b < batch; b++)
            for (uint64_t s = 0; s < ncc; s++) { ab[(b * ncc + s) * 2] = idx[b * ncc + s]; ab[(b * ncc + s) * 2 + 1] = idx[b * ncc + s] + half; }
        uint8_t* d_pc = d_pab + 2 * batch * ncc * pb_cur;
        ZKB_CUDA(c, cudaMemcpyAsync(d_ab, ab.data(), ab.size() * 8, cudaMemcpyHostToDevice, c->stream));
        ZKB_CUDA(c, cudaMemcpyAsync(d_c, idx.data(), idx.size() * 8, cudaMemcpyHostToDevice, c->stream));
        {
            LaunchScope ls(c, K_GATHER);
            k_gather3_batch<<<dim3((unsigned)((ncc + 127) / 128), (unsigned)batch), 128, 0, c->stream>>>(
                cw_ptr(r), cw_stride(r), cw_ptr(r + 1), cw_stride(r + 1), half, d_c, (uint32_t)ncc, d_leafs);
        }
        ZKB_CUDA(c, cudaGetLastError());
        ZKB_TRY(merkle_open_device_batch(c, cw_ptr(r), lay[r], A + node_off[r], d_ab, 2 * ncc, d_pab, (uint32_t)batch, cw_stride(r), inst_bytes));
        if (d_nxt > 0) ZKB_TRY(merkle_open_device_batch(c, cw_ptr(r + 1), lay[r + 1], A + node_off[r + 1], d_c, ncc, d_pc, (uint32_t)batch, cw_stride(r + 1), inst_bytes));
        ZKB_CUDA(c, cudaMemcpyAsync(leafs, d_leafs, batch * ncc * 48, cudaMemcpyDeviceToHost, c->stream));
        ZKB_CUDA(c, cudaMemcpyAsync(hab, d_pab, 2 * batch * ncc * pb_cur, cudaMemcpyDeviceToHost, c->stream));
        if (pb_nxt) ZKB_CUDA(c, cudaMemcpyAsync(hc, d_pc, batch * ncc * pb_nxt, cudaMemcpyDeviceToHost, c->stream));
        ZKB_CUDA(c, ctx_stream_sync(c));
        parallel_for(batch, c->assembly_threads, [&](size_t b) {
            const uint8_t* lf = leafs + b * ncc * 48;
            for (uint64_t s = 0; s < ncc; s++)                               // fri.rs:189-195
                zkb_ps_push_leafs(ps[b], lf + 48 * s, lf + 48 * s + 16, lf + 48 * s + 32);
            const uint8_t* pa = hab + b * 2 * ncc * pb_cur;
            const uint8_t* pc = hc + b * ncc * pb_nxt;
            for (uint64_t s = 0; s < ncc; s++) {                             // fri.rs:198-202
                zkb_ps_push_path(ps[b], pa + (2 * s) * pb_cur, d_cur);
                zkb_ps_push_path(ps[b], pa + (2 * s + 1) * pb_cur, d_cur);
                zkb_ps_push_path(ps[b], pc + s * pb_nxt, d_nxt);
            }
        });
    }
    return 0;
}

int zkb_fri_prove_batch(zkb_ctx* c, const zkb_fri_params* p, const void* codewords, size_t n, size_t stride, size_t batch, zkb_ps* const* ps, uint64_t* top_indices_out) {
    if (!c) return ZKB_ERR_ARG;
    ZKB_ABI_GUARD(c, return zkb_fri_prove_batch_impl(c, p, codewords, n, stride, batch, ps, top_indices_out);)
}

static int zkb_merkle_open_ps_batch_impl(zkb_tree* const* trees, size_t count, const uint64_t* idx, size_t k, zkb_ps* const* ps) {
    if (!trees || !ps || count == 0 || (k && !idx)) return ZKB_ERR_ARG;
    for (size_t i = 0; i < count; i++) if (!trees[i] || !ps[i]) return ZKB_ERR_ARG;
    zkb_ctx* c = trees[0]->ctx;
    const uint64_t n = trees[0]->n;
    if (n < 2) return set_err(c, ZKB_ERR_INDEX, "open on a 1-leaf tree (the reference recurses forever)");
    if (k == 0) return 0;
    // the trees must be one uniform batch (zkb_merkle_build_batch): same size, constant strides
    const ptrdiff_t nstride = count > 1 ? trees[1]->nodes - trees[0]->nodes : 0;
    const ptrdiff_t vstride = count > 1 ? trees[1]->vals - trees[0]->vals : 0;
    for (size_t i = 0; i < count; i++) {
        if (trees[i]->ctx != c || trees[i]->n != n || trees[i]->layout.top != 0 ||
            trees[i]->nodes - trees[0]->nodes != (ptrdiff_t)i * nstride || trees[i]->vals - trees[0]->vals != (ptrdiff_t)i * vstride)
            return set_err(c, ZKB_ERR_ARG, "merkle_open_ps_batch: the trees are not one batch of zkb_merkle_build_batch");
    }
    if (count > 1 && (nstride <= 0 || vstride <= 0)) return set_err(c, ZKB_ERR_ARG, "merkle_open_ps_batch: the trees are not one batch of zkb_merkle_build_batch");
    for (size_t i = 0; i < count * k; i++)
        if (idx[i] >= n) return set_err(c, ZKB_ERR_INDEX, "cannot open invalid index %llu", (unsigned long long)idx[i]);
    ZKB_CUDA(c, cudaSetDevice(c->device));
    if (getenv("ZKB_HOST_ASSEMBLY") == nullptr && trees[0]->layout.log_n >= 1 && trees[0]->layout.log_n <= 30) {
        // device-side framing: record (Value, Path) s of tree i at its final position inside its proof's segment; trees that share a
        // proof stream append in increasing tree order (stark.rs:546-560 loops index-major per codeword in the order of the commits)
        const uint64_t rec = 25 + 9 + 72ull * trees[0]->layout.log_n;
        std::vector<zkb_ps*> streams;
        std::vector<uint64_t> members;                                       // trees per stream
        std::vector<uint64_t> hbuf(count * k + count);
        uint64_t* yoff = hbuf.data() + count * k;
        memcpy(hbuf.data(), idx, count * k * 8);
        std::vector<size_t> group(count);
        for (size_t i = 0; i < count; i++) {
            size_t g = std::find(streams.begin(), streams.end(), ps[i]) - streams.begin();
            if (g == streams.size()) { streams.push_back(ps[i]); members.push_back(0); }
            group[i] = g;
            members[g]++;
        }
        std::vector<uint64_t> seg_off(streams.size() + 1, 0), fill(streams.size(), 0);
        for (size_t g = 0; g < streams.size(); g++) seg_off[g + 1] = seg_off[g] + ((members[g] * k * rec + 15) & ~15ull);
        for (size_t i = 0; i < count; i++) { yoff[i] = seg_off[group[i]] + fill[group[i]] * k * rec; fill[group[i]]++; }
        DevBuf buf;
        const size_t hb = (hbuf.size() * 8 + 255) & ~(size_t)255;
        ZKB_TRY(buf.alloc(c, hb + seg_off.back()));
        uint64_t* d_idx = (uint64_t*)buf.p;
        uint8_t* d_wire = (uint8_t*)buf.p + hb;
        ZKB_CUDA(c, cudaMemcpyAsync(d_idx, hbuf.data(), hbuf.size() * 8, cudaMemcpyHostToDevice, c->stream));
        ZKB_TRY(merkle_open_wire_batch(c, trees[0]->vals, trees[0]->layout, trees[0]->nodes, d_idx, k, (uint32_t)count, (uint64_t)vstride, (uint64_t)nstride,
                                       d_wire, d_idx + count * k, 1, 0, rec, 0, true));
        // as in zkb_fri_prove_batch: each stream's records go straight from device memory to the tail of its (pinned) body
        std::vector<uint8_t*> dst(streams.size(), nullptr);
        const bool ok = parallel_for(streams.size(), c->assembly_threads, [&](size_t g) { dst[g] = ps_body_extend(streams[g], members[g] * k * rec); });
        if (!ok) return set_err(c, ZKB_ERR_NOMEM, "out of host memory while appending to the proof streams (they are unusable now)");
        for (size_t g = 0; g < streams.size(); g++)
            ZKB_CUDA(c, cudaMemcpyAsync(dst[g], d_wire + seg_off[g], members[g] * k * rec, cudaMemcpyDeviceToHost, c->stream));
        ZKB_CUDA(c, ctx_stream_sync(c));
        for (size_t g = 0; g < streams.size(); g++) streams[g]->has_field = true;   // Value objects carry field elements
        return 0;
    }
    const size_t depth = trees[0]->layout.log_n, path_bytes = depth * 64;
    const size_t idx_bytes = (count * k * 8 + 255) & ~(size_t)255, val_bytes = (count * k * 16 + 255) & ~(size_t)255;
    DevBuf buf;
    ZKB_TRY(buf.alloc(c, idx_bytes + val_bytes + count * k * path_bytes));
    uint64_t* d_idx = (uint64_t*)buf.p;
    fe* d_vals = (fe*)((uint8_t*)buf.p + idx_bytes);
    uint8_t* d_paths = (uint8_t*)buf.p + idx_bytes + val_bytes;
    ZKB_CUDA(c, cudaMemcpyAsync(d_idx, idx, count * k * 8, cudaMemcpyHostToDevice, c->stream));
    {
        LaunchScope ls(c, K_GATHER);
        k_gather_vals_batch<<<dim3((unsigned)((k + 127) / 128), (unsigned)count), 128, 0, c->stream>>>(trees[0]->vals, (uint64_t)vstride, d_idx, (uint32_t)k, d_vals);
    }
    ZKB_CUDA(c, cudaGetLastError());
    ZKB_TRY(merkle_open_device_batch(c, trees[0]->vals, trees[0]->layout, trees[0]->nodes, d_idx, k, d_paths, (uint32_t)count, (uint64_t)vstride, (uint64_t)nstride));
    uint8_t* host = nullptr;                                                  // pinned
    ZKB_TRY(host_scratch_reserve(c, 1, val_bytes + count * k * path_bytes, &host));
    ZKB_CUDA(c, cudaMemcpyAsync(host, d_vals, val_bytes + count * k * path_bytes, cudaMemcpyDeviceToHost, c->stream));
    ZKB_CUDA(c, ctx_stream_sync(c));
    // trees that share a proof stream append in increasing tree order; one worker per distinct stream
    std::vector<zkb_ps*> streams;
    std::vector<std::vector<size_t>> members;
    for (size_t i = 0; i < count; i++) {
        size_t g = std::find(streams.begin(), streams.end(), ps[i]) - streams.begin();
        if (g == streams.size()) { streams.push_back(ps[i]); members.emplace_back(); }
        members[g].push_back(i);
    }
    parallel_for(streams.size(), c->assembly_threads, [&](size_t g) {
        for (size_t i : members[g])
            for (size_t s = 0; s < k; s++) {                                 // stark.rs:546-560
                zkb_ps_push_value(streams[g], host + (i * k + s) * 16);
                zkb_ps_push_path(streams[g], host + val_bytes + (i * k + s) * path_bytes, depth);
            }
    });
    return 0;
}

int zkb_merkle_open_ps_batch(zkb_tree* const* trees, size_t count, const uint64_t* idx, size_t k, zkb_ps* const* ps) {
    if (!trees || count == 0 || !trees[0]) return ZKB_ERR_ARG;
    ZKB_ABI_GUARD(trees[0]->ctx, return zkb_merkle_open_ps_batch_impl(trees, count, idx, k, ps);)
}

}  // extern "C"
