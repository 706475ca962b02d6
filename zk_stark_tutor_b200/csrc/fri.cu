// fri.cu - FRI commit / query / prove on the device.
//
// Replaces the bodies of (reference file:line):
//   FRI::num_rounds   src/fri.rs:40-50
//   FRI::commit       src/fri.rs:115-172   (Merkle root -> Fiat-Shamir alpha -> split-and-fold)
//   FRI::query        src/fri.rs:174-208
//   FRI::prove        src/fri.rs:210-248
//
// The fold  c'[i] = 2^-1((1 + a/x_i) c[i] + (1 - a/x_i) c[i+n/2]),  x_i = offset*omega^i
// is evaluated as  half(c[i] + c[i+n/2] + k_i (c[i] - c[i+n/2])),  k_i = (alpha/offset) omega^-i
// - the same field element (exact arithmetic), 2.25 multiplications per output instead of the
// reference's pow + xgcd inverse + 5 products.  omega^-i comes from one two-level power
// table of omega_0^-1 shared by all rounds (round r uses exponent i*2^r).  The fold runs inside
// the leaf-hash kernel of the NEXT layer (merkle.cu k_leaf8<true> / k_leaf1<true>): the folded
// value is written once to HBM and hashed while still in registers.  Every layer (codeword + pruned tree) stays on the device for the
// query phase; only the 64-byte root goes to the host each round, where the Fiat-Shamir
// callback turns it into alpha (SHAKE256 over a transcript of < 1.5 KB).
#include <stddef.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "merkle.cuh"
#include "ntt.cuh"
#include "hosthash.hpp"
#include "keccak.cuh"
#include "fri_tail.cuh"

namespace zkb {

__device__ __forceinline__ fe pow2lvl_f(const DevPow& t, uint64_t e) {
    fe lo = fe_ldg(t.lo + (e & ((1ull << t.lo_bits) - 1)));
    fe hi = fe_ldg(t.hi + (e >> t.lo_bits));
    return fe_montmul(hi, lo);
}

// stand-alone fold (small layers, and the zkb_fri_fold entry point)
__global__ void k_fold(FoldArgs f) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= f.half) return;
    fe k_m = fe_montmul(f.kk_m, pow2lvl_f(f.winv, i * f.exp_mul));
    fe a = fe_ldg(f.cw + i), b = fe_ldg(f.cw + f.half + i);
    fe s = fe_add(a, b), d = fe_sub(a, b);
    fe_store(f.next + i, fe_half(fe_add(s, fe_montmul(k_m, d))));
}

__global__ void k_gather3(const fe* cur, const fe* nxt, uint64_t half, const uint64_t* idx, uint32_t k, fe* out) {
    uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= k) return;
    uint64_t a = idx[s];
    fe_store(out + 3 * s, fe_ldg(cur + a));
    fe_store(out + 3 * s + 1, fe_ldg(cur + a + half));
    fe_store(out + 3 * s + 2, fe_ldg(nxt + a));
}

static int launch_fold(zkb_ctx* c, const FoldArgs& f) {
    { LaunchScope ls(c, K_FOLD); k_fold<<<(unsigned)((f.half + 255) / 256), 256, 0, c->stream>>>(f); }
    ZKB_CUDA(c, cudaGetLastError());
    return 0;
}

}  // namespace zkb

using namespace zkb;

struct zkb_fri_layers {
    zkb_ctx* ctx = nullptr;
    uint64_t rounds = 0;
    std::vector<uint64_t> len;           // codeword length per round
    std::vector<const fe*> cw;           // device codeword per round (round 0 may alias the caller's buffer)
    std::vector<TreeLayout> layout;
    std::vector<uint8_t*> nodes;         // device, per round (slices of `arena`)
    std::vector<std::vector<uint8_t>> roots;
    void* arena = nullptr;               // one allocation: folded codewords + all trees
    void* owned_cw0 = nullptr;           // staged copy of a host codeword
    std::vector<uint8_t> last_cw;        // host copy of the last codeword when the commit already brought it back
};

extern "C" {

uint64_t zkb_fri_num_rounds(const zkb_fri_params* p) {
    if (!p) return 0;
    uint64_t n = p->domain_length, r = 0;           // fri.rs:40-50
    while (n > p->expansion_factor && n > 4 * p->num_colinearity_tests) { n /= 2; r++; }
    return r;
}

int zkb_fri_fold(zkb_ctx* c, const void* cw, size_t n, const uint8_t alpha[16], const uint8_t offset[16],
                 const uint8_t omega[16], void* out) {
    if (!c || !cw || !alpha || !offset || !omega || !out) return ZKB_ERR_ARG;
    if (n < 2 || (n & (n - 1))) return set_err(c, ZKB_ERR_NOT_POW2, "fri_fold: codeword length %zu is not a power of two >= 2", n);
    ZKB_CUDA(c, cudaSetDevice(c->device));
    const bool out_dev = is_device_ptr(out);
    DevBuf bin, bout;
    const void* d_in = nullptr;
    ZKB_TRY(stage_in(c, cw, n * sizeof(fe), bin, &d_in));
    fe* d_out = (fe*)out;
    if (!out_dev) { ZKB_TRY(bout.alloc(c, (n / 2) * sizeof(fe))); d_out = (fe*)bout.p; }
    fe w = h_load(omega), off = h_load(offset);
    if (fe_is_zero(off) || fe_is_zero(w)) return set_err(c, ZKB_ERR_DIV_ZERO, "divide by zero");
    fe winv = h_inv(w);
    FoldArgs f;
    f.cw = (const fe*)d_in; f.next = d_out; f.half = n / 2;
    ZKB_TRY(get_pow_table(c, winv, ilog2_u64(n), &f.winv));
    f.exp_mul = 1;
    f.kk_m = fe_to_mont(h_mul(h_load(alpha), h_inv(off)));
    f.wr_inv_m = fe_to_mont(winv);
    ZKB_TRY(launch_fold(c, f));
    if (!out_dev) ZKB_CUDA(c, cudaMemcpyAsync(out, d_out, (n / 2) * sizeof(fe), cudaMemcpyDeviceToHost, c->stream));
    if (!out_dev || bin.p) ZKB_CUDA(c, cudaStreamSynchronize(c->stream));
    return 0;
}

void zkb_fri_layers_free(zkb_fri_layers* l) {
    if (!l) return;
    cudaSetDevice(l->ctx->device);
    dev_free(l->ctx, l->arena);
    dev_free(l->ctx, l->owned_cw0);
    delete l;
}

// Shared body of zkb_fri_commit / zkb_lde_fri_commit.  Exactly one of `codeword` (n values)
// and `coeffs` (n_coeffs <= n coefficients, extended to the coset codeword on the device
// first: fast_coset_evaluate ntt_arithmetics.rs:161-170 with omega/offset of `p`) is set.
// `dev_ps` non-null: the transcript is the library's own proof stream and the whole commit runs WITHOUT host hops -
// the tree kernels do the Fiat-Shamir step on the device (keccak.cuh) and every layer of <= 2^17 values is handled
// by the persistent tail kernel (fri_tail.cu); the roots are pushed to `dev_ps` afterwards (same bytes, same order).
// Otherwise (`fs` callback: any foreign ProofStream) each round hands its root to the host as before.
static int fri_commit_device_fs(zkb_ctx* c, zkb_fri_layers* L, const fe& omega_inv0, const DevPow& winv_tab,
                                const std::vector<fe>& inv_offset, zkb_ps* ps, bool leaf3_done);

static int fri_commit_impl(zkb_ctx* c, const zkb_fri_params* p, const void* codeword, const void* coeffs,
                           size_t n_coeffs, size_t n, zkb_fs_callback fs, void* user, zkb_fri_layers** out, zkb_ps* dev_ps = nullptr) {
    *out = nullptr;
    if (p->domain_length != n) return set_err(c, ZKB_ERR_LENGTH, "Length of the domain doesnt match the length of initial codeword");
    if (n < 2 || (n & (n - 1))) return set_err(c, ZKB_ERR_NOT_POW2, "Leafs len must be power of two (got %zu)", n);
    const uint64_t rounds = zkb_fri_num_rounds(p);
    if (rounds < 1) return set_err(c, ZKB_ERR_ROUNDS, "FRI needs at least one round");
    ZKB_CUDA(c, cudaSetDevice(c->device));
    fe omega = h_load(p->omega), offset = h_load(p->offset);
    // fri.rs:133: omega^(n-1) == omega^-1, i.e. omega^n == 1 (checked once; squaring preserves it)
    if (!fe_eq(h_pow(omega, n), fe_from_u32(1)) || fe_is_zero(omega))
        return set_err(c, ZKB_ERR_ROOT_ORDER, "error in commit: omega does not have the right order!");
    if (fe_is_zero(offset)) return set_err(c, ZKB_ERR_DIV_ZERO, "divide by zero");

    std::unique_ptr<zkb_fri_layers> L(new zkb_fri_layers());
    L->ctx = c;
    L->rounds = rounds;
    size_t arena_bytes = 0;
    std::vector<size_t> cw_off(rounds), node_off(rounds);
    for (uint64_t r = 0; r < rounds; r++) {
        uint64_t len = n >> r;
        L->len.push_back(len);
        TreeLayout tl;
        tl.init(ilog2_u64(len));
        L->layout.push_back(tl);
        if (r > 0) { cw_off[r] = arena_bytes; arena_bytes += len * sizeof(fe); }
        node_off[r] = arena_bytes;
        arena_bytes += tl.total_nodes * 64;
    }
    ZKB_CUDA(c, dev_alloc(c, &L->arena, arena_bytes));
    auto fail = [&](int rc) { dev_free(c, L->arena); dev_free(c, L->owned_cw0); L->arena = nullptr; L->owned_cw0 = nullptr; return rc; };
    DevBuf staged_coeffs;
    bool leaf3_done = false;
    if (codeword && is_device_ptr(codeword)) {
        L->cw.push_back((const fe*)codeword);
    } else {
        cudaError_t e = dev_alloc(c, &L->owned_cw0, n * sizeof(fe));
        if (e != cudaSuccess) return fail(set_err(c, ZKB_ERR_CUDA, "allocating the codeword failed: %s", cudaGetErrorString(e)));
        if (codeword) {
            e = cudaMemcpyAsync(L->owned_cw0, codeword, n * sizeof(fe), cudaMemcpyHostToDevice, c->stream);
            if (e != cudaSuccess) return fail(set_err(c, ZKB_ERR_CUDA, "staging the codeword failed: %s", cudaGetErrorString(e)));
        } else if (n_coeffs == 0) {
            e = cudaMemsetAsync(L->owned_cw0, 0, n * sizeof(fe), c->stream);
            if (e != cudaSuccess) return fail(set_err(c, ZKB_ERR_CUDA, "memset failed: %s", cudaGetErrorString(e)));
        } else {
            const void* d_coeffs = nullptr;
            int rc = stage_in_once(c, coeffs, n_coeffs * sizeof(fe), staged_coeffs, &d_coeffs);
            if (rc) return fail(rc);
            NttOpts o;
            o.has_scale = true;
            o.scale_base = offset;
            // opt-in (measured: no gain, see ntt.cu): the LDE's last pass hashes its own output, layer 0's level-3 nodes come out of the transform
            if (L->layout[0].top >= 3 && ntt_can_fuse_leaves(ilog2_u64(n)) && getenv("ZKB_NTT_LEAF_FUSION") != nullptr) {
                o.leaf3_out = (uint8_t*)L->arena + node_off[0] + L->layout[0].level_off[3] * 64;
                leaf3_done = true;
            }
            rc = ntt_exec(c, omega, (const fe*)d_coeffs, n_coeffs, 0, (fe*)L->owned_cw0, 0, 1, ilog2_u64(n), o);
            if (rc) return fail(rc);
        }
        L->cw.push_back((const fe*)L->owned_cw0);
    }
    for (uint64_t r = 0; r < rounds; r++) {
        L->nodes.push_back((uint8_t*)L->arena + node_off[r]);
        if (r > 0) L->cw.push_back((const fe*)((uint8_t*)L->arena + cw_off[r]));
    }
    DevPow winv_tab;
    fe omega_inv0 = h_inv(omega);
    int rc = get_pow_table(c, omega_inv0, ilog2_u64(n), &winv_tab);
    if (rc) return fail(rc);

    fe alpha = fe_zero();
    fe omega_inv_r = omega_inv0;                         // round-r omega^-1
    // 1/offset_r for every round up front (offset_r = offset^(2^r), fri.rs:161-162): independent
    // of the challenges, so it stays off the per-round critical path
    std::vector<fe> inv_offset(rounds);
    {
        fe io = h_inv(offset);
        for (uint64_t r = 0; r < rounds; r++) { inv_offset[r] = io; io = h_mul(io, io); }
    }
    if (dev_ps && rounds <= ZKB_FS_MAX_ROUNDS && getenv("ZKB_HOST_FS") == nullptr) {
        rc = fri_commit_device_fs(c, L.get(), omega_inv0, winv_tab, inv_offset, dev_ps, leaf3_done);
        if (rc) return fail(rc);
        *out = L.release();
        return 0;
    }
    // roots come back through mapped pinned memory: the top kernel writes root + sequence flag,
    // the host polls (no D2H copy, no stream synchronisation on the critical path)
    RootSignal sig;
    sig.host_root = c->pinned + 128;
    sig.host_flag = reinterpret_cast<volatile uint32_t*>(c->pinned + 64);
    for (uint64_t r = 0; r < rounds; r++) {
        const uint64_t len = L->len[r];
        const TreeLayout& tl = L->layout[r];
        sig.seq = ++c->root_seq;
        const bool polled = len > 1;
        if (r == 0) {
            rc = merkle_build_levels(c, L->cw[0], nullptr, len, tl, L->nodes[0], &sig, nullptr, leaf3_done);
        } else {
            FoldArgs f;
            f.cw = L->cw[r - 1]; f.next = (fe*)L->cw[r]; f.half = len;
            f.winv = winv_tab; f.exp_mul = 1ull << (r - 1);
            f.kk_m = fe_to_mont(h_mul(alpha, inv_offset[r - 1]));
            f.wr_inv_m = fe_to_mont(omega_inv_r);
            rc = merkle_build_levels(c, nullptr, &f, len, tl, L->nodes[r], &sig);   // fold fused with leaf hashing
            omega_inv_r = h_mul(omega_inv_r, omega_inv_r);
        }
        if (rc) return fail(rc);
        if (polled) {
            uint8_t root[64];
            rc = wait_root(c, sig, root);
            if (rc) return fail(rc);
            L->roots.emplace_back(root, root + 64);
        } else {
            cudaError_t e = cudaMemcpyAsync(c->pinned, L->nodes[r] + tl.level_off[tl.log_n] * 64, 64, cudaMemcpyDeviceToHost, c->stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
            if (e != cudaSuccess) return fail(set_err(c, ZKB_ERR_CUDA, "FRI round %llu failed: %s", (unsigned long long)r, cudaGetErrorString(e)));
            L->roots.emplace_back(c->pinned, c->pinned + 64);
        }
        const int want_alpha = r + 1 < rounds;
        uint8_t alpha_le[16] = {0};
        if (fs(user, (uint32_t)r, L->roots.back().data(), want_alpha, alpha_le) != 0)
            return fail(set_err(c, ZKB_ERR_CALLBACK, "Fiat-Shamir callback failed in round %llu", (unsigned long long)r));
        if (want_alpha) {
            alpha = h_load(alpha_le);
            if (fe_ge_p(alpha)) return fail(set_err(c, ZKB_ERR_ARG, "Fiat-Shamir callback returned a non-canonical alpha"));
        }
    }
    *out = L.release();
    return 0;
}


#define ZKB_PINNED_FS_INIT 16384u     // c->pinned: staging of the FsDev head (sponge, kk_m, alpha, inv_off_m2[rounds])
#define ZKB_PINNED_TAIL_OUT 65536u    // c->pinned: roots + last codeword written by the tail kernel

static int fri_commit_device_fs(zkb_ctx* c, zkb_fri_layers* L, const fe& omega_inv0, const DevPow& winv_tab,
                                const std::vector<fe>& inv_offset, zkb_ps* ps, bool leaf3_done) {
    const uint64_t rounds = L->rounds;
    ZKB_TRY(ensure_fs_dev(c, 1));
    FsDev* fs = (FsDev*)c->fs_dev;
    uint32_t* bar = (uint32_t*)((uint8_t*)c->fs_dev + c->fs_dev_count * sizeof(FsDev));
    {   // head of the FsDev: transcript sponge as `ps` stands now, and the per-round constants
        FsDev* h = (FsDev*)(c->pinned + ZKB_PINNED_FS_INIT);
        ps_export_sponge(ps, &h->sp);
        h->kk_m = fe_zero(); h->alpha = fe_zero();
        const fe r2 = ZKB_FE_R2;
        for (uint64_t r = 0; r < rounds; r++) h->inv_off_m2[r] = fe_montmul(fe_to_mont(inv_offset[r]), r2);   // (1/offset_r) * R^2
        const size_t head = offsetof(FsDev, inv_off_m2) + rounds * sizeof(fe);
        ZKB_CUDA(c, cudaMemcpyAsync(fs, h, head, cudaMemcpyHostToDevice, c->stream));
    }
    // rounds handled by the throughput kernels (k_leaf8 / k_node8 / k_tree): every layer above the tail's limit
    uint64_t r0 = 0;
    fe omega_inv_r = omega_inv0;
    // (c->tail_threads == 0: no persistent tail at all - every round runs as its own small launches, still without host hops;
    // for callers that keep several contexts busy on one GPU, where the other lanes hide the latency and a 128-SM cooperative
    // kernel would only get in their way)
    while (r0 < rounds && (c->tail_threads == 0 || !(L->layout[r0].top == 0 && L->layout[r0].log_n <= ZKB_TAIL_MAX_LOG))) {
        const uint64_t r = r0;
        FsHook hook;
        hook.fs = fs; hook.round = (uint32_t)r; hook.want_alpha = r + 1 < rounds;
        if (r == 0) {
            ZKB_TRY(merkle_build_levels(c, L->cw[0], nullptr, L->len[0], L->layout[0], L->nodes[0], nullptr, &hook, leaf3_done));
        } else {
            FoldArgs f;
            f.cw = L->cw[r - 1]; f.next = (fe*)L->cw[r]; f.half = L->len[r];
            f.winv = winv_tab; f.exp_mul = 1ull << (r - 1);
            f.kk_m = fe_zero();
            f.kk_dev = (const uint8_t*)&fs->kk_m; f.kk_stride = 0;
            f.wr_inv_m = fe_to_mont(omega_inv_r);
            ZKB_TRY(merkle_build_levels(c, nullptr, &f, L->len[r], L->layout[r], L->nodes[r], nullptr, &hook));
            omega_inv_r = h_mul(omega_inv_r, omega_inv_r);
        }
        r0++;
    }
    const uint64_t last_len = L->len[rounds - 1];
    const bool mapped_out = r0 < rounds && ZKB_PINNED_TAIL_OUT + ZKB_TAIL_HOST_CW_OFF + last_len * sizeof(fe) <= c->pinned_bytes;
    RootSignal sig;
    sig.host_root = c->pinned + ZKB_PINNED_TAIL_OUT;
    sig.host_flag = reinterpret_cast<volatile uint32_t*>(c->pinned + 64);
    sig.seq = ++c->root_seq;
    if (r0 < rounds) {
        TailArgs a;
        memset(&a, 0, sizeof(a));
        a.first_is_plain = r0 == 0;
        a.cw_in = r0 == 0 ? L->cw[0] : L->cw[r0 - 1];
        a.log_n0 = L->layout[r0].log_n;
        a.n_rounds = (uint32_t)(rounds - r0);
        a.r0 = (uint32_t)r0; a.total_rounds = (uint32_t)rounds;
        for (uint64_t r = r0; r < rounds; r++) { a.cw[r - r0] = (fe*)L->cw[r]; a.nodes[r - r0] = L->nodes[r]; }
        a.winv = winv_tab;
        a.fs = fs; a.bar = bar;
        if (mapped_out) { a.host_out = sig.host_root; a.host_flag = sig.host_flag; a.seq = sig.seq; }
        ZKB_TRY(fri_tail_launch(c, a));
    }
    uint8_t* roots_host = c->pinned + ZKB_PINNED_TAIL_OUT;
    if (mapped_out) {
        uint64_t spins = 0;
        for (;;) {
            const uint32_t f = *sig.host_flag;
            if (f == sig.seq) break;
            if (f == ZKB_TAIL_TIMEOUT_FLAG) return set_err(c, ZKB_ERR_CUDA, "FRI tail kernel: grid barrier timed out");
            if ((++spins & 0xFFFF) == 0) {
                cudaError_t e = cudaStreamQuery(c->stream);
                if (e != cudaSuccess && e != cudaErrorNotReady) return set_err(c, ZKB_ERR_CUDA, "FRI commit failed: %s", cudaGetErrorString(e));
                if (e == cudaSuccess && *sig.host_flag != sig.seq) return set_err(c, ZKB_ERR_CUDA, "FRI tail kernel finished without signalling");
            }
        }
        __sync_synchronize();
        L->last_cw.assign(roots_host + ZKB_TAIL_HOST_CW_OFF, roots_host + ZKB_TAIL_HOST_CW_OFF + last_len * sizeof(fe));
    } else {
        ZKB_CUDA(c, cudaMemcpyAsync(roots_host, fs->roots, rounds * 64, cudaMemcpyDeviceToHost, c->stream));
        ZKB_CUDA(c, cudaStreamSynchronize(c->stream));
    }
    for (uint64_t r = 0; r < rounds; r++) {
        L->roots.emplace_back(roots_host + 64 * r, roots_host + 64 * r + 64);
        zkb_ps_push_root(ps, roots_host + 64 * r, 64);                       // fri.rs:136-137, in round order
    }
    return 0;
}

static int zkb_fri_commit_body(zkb_ctx* c, const zkb_fri_params* p, const void* codeword, size_t n,
                   zkb_fs_callback fs, void* user, zkb_fri_layers** out) {
    if (!c || !p || !codeword || !fs || !out) return ZKB_ERR_ARG;
    return fri_commit_impl(c, p, codeword, nullptr, 0, n, fs, user, out);
}

static int zkb_lde_fri_commit_body(zkb_ctx* c, const zkb_fri_params* p, const void* coeffs, size_t n_coeffs,
                       zkb_fs_callback fs, void* user, zkb_fri_layers** out) {
    if (!c || !p || !fs || !out || (n_coeffs && !coeffs)) return ZKB_ERR_ARG;
    if (n_coeffs > p->domain_length)
        return set_err(c, ZKB_ERR_TOO_LONG, "coset_lde: %zu coefficients exceed root_order %llu", n_coeffs, (unsigned long long)p->domain_length);
    return fri_commit_impl(c, p, nullptr, coeffs, n_coeffs, (size_t)p->domain_length, fs, user, out);
}

uint64_t zkb_fri_layer_count(const zkb_fri_layers* l) { return l ? l->rounds : 0; }
uint64_t zkb_fri_layer_len(const zkb_fri_layers* l, uint64_t r) { return (l && r < l->rounds) ? l->len[r] : 0; }
const void* zkb_fri_layer_device_ptr(const zkb_fri_layers* l, uint64_t r) { return (l && r < l->rounds) ? l->cw[r] : nullptr; }
int zkb_fri_layer_root(const zkb_fri_layers* l, uint64_t r, uint8_t root[64]) {
    if (!l || r >= l->rounds || !root) return ZKB_ERR_ARG;
    memcpy(root, l->roots[r].data(), 64);
    return 0;
}
int zkb_fri_layer_codeword(zkb_fri_layers* l, uint64_t r, void* out) {
    if (!l || r >= l->rounds || !out) return ZKB_ERR_ARG;
    zkb_ctx* c = l->ctx;
    ZKB_CUDA(c, cudaMemcpyAsync(out, l->cw[r], l->len[r] * sizeof(fe), cudaMemcpyDefault, c->stream));
    ZKB_CUDA(c, cudaStreamSynchronize(c->stream));
    return 0;
}

static int zkb_fri_query_body(zkb_fri_layers* l, uint64_t r, const uint64_t* idx_c, size_t ncc, uint8_t* leafs_out, uint8_t* paths_out) {
    if (!l || !idx_c || !leafs_out || !paths_out) return ZKB_ERR_ARG;
    zkb_ctx* c = l->ctx;
    if (r + 1 >= l->rounds) return set_err(c, ZKB_ERR_ARG, "fri_query: no layer after round %llu", (unsigned long long)r);
    if (ncc == 0) return 0;
    ZKB_CUDA(c, cudaSetDevice(c->device));
    const uint64_t len = l->len[r], half = len / 2;
    const uint32_t d_cur = l->layout[r].log_n, d_nxt = l->layout[r + 1].log_n;
    for (size_t s = 0; s < ncc; s++)
        if (idx_c[s] >= half) return set_err(c, ZKB_ERR_INDEX, "fri_query: index %llu out of range", (unsigned long long)idx_c[s]);
    // index lists: a (cur), b (cur), c (next)
    std::vector<uint64_t> ab(2 * ncc), cc(ncc);
    for (size_t s = 0; s < ncc; s++) { ab[2 * s] = idx_c[s]; ab[2 * s + 1] = idx_c[s] + half; cc[s] = idx_c[s]; }
    const size_t pb_cur = (size_t)d_cur * 64, pb_nxt = (size_t)d_nxt * 64;
    DevBuf buf;
    size_t bytes = 3 * ncc * 8 + 3 * ncc * 16 + 2 * ncc * pb_cur + ncc * pb_nxt + 64;
    ZKB_TRY(buf.alloc(c, bytes));
    uint64_t* d_ab = (uint64_t*)buf.p;
    uint64_t* d_c = d_ab + 2 * ncc;
    fe* d_leafs = (fe*)(d_c + ncc + (ncc & 1));
    uint8_t* d_pab = (uint8_t*)(d_leafs + 3 * ncc);
    uint8_t* d_pc = d_pab + 2 * ncc * pb_cur;
    ZKB_CUDA(c, cudaMemcpyAsync(d_ab, ab.data(), 2 * ncc * 8, cudaMemcpyHostToDevice, c->stream));
    ZKB_CUDA(c, cudaMemcpyAsync(d_c, cc.data(), ncc * 8, cudaMemcpyHostToDevice, c->stream));
    { LaunchScope ls(c, K_GATHER); k_gather3<<<(unsigned)((ncc + 127) / 128), 128, 0, c->stream>>>(l->cw[r], l->cw[r + 1], half, d_c, (uint32_t)ncc, d_leafs); }
    ZKB_CUDA(c, cudaGetLastError());
    ZKB_TRY(merkle_open_device(c, l->cw[r], l->layout[r], l->nodes[r], d_ab, 2 * ncc, d_pab));
    if (d_nxt > 0) ZKB_TRY(merkle_open_device(c, l->cw[r + 1], l->layout[r + 1], l->nodes[r + 1], d_c, ncc, d_pc));
    std::vector<uint8_t> hab(2 * ncc * pb_cur), hc(ncc * pb_nxt);
    ZKB_CUDA(c, cudaMemcpyAsync(leafs_out, d_leafs, 3 * ncc * 16, cudaMemcpyDeviceToHost, c->stream));
    ZKB_CUDA(c, cudaMemcpyAsync(hab.data(), d_pab, hab.size(), cudaMemcpyDeviceToHost, c->stream));
    if (!hc.empty()) ZKB_CUDA(c, cudaMemcpyAsync(hc.data(), d_pc, hc.size(), cudaMemcpyDeviceToHost, c->stream));
    ZKB_CUDA(c, cudaStreamSynchronize(c->stream));
    // interleave per s: path(a) || path(b) || path(c)
    uint8_t* o = paths_out;
    for (size_t s = 0; s < ncc; s++) {
        memcpy(o, hab.data() + (2 * s) * pb_cur, pb_cur); o += pb_cur;
        memcpy(o, hab.data() + (2 * s + 1) * pb_cur, pb_cur); o += pb_cur;
        memcpy(o, hc.data() + s * pb_nxt, pb_nxt); o += pb_nxt;
    }
    return 0;
}

// ---- FRI::prove against the built-in proof stream -------------------------------------------
static int ps_callback(void* user, uint32_t, const uint8_t root[64], int want_alpha, uint8_t alpha_out[16]) {
    zkb_ps* ps = (zkb_ps*)user;
    zkb_ps_push_root(ps, root, 64);                         // fri.rs:136-137
    if (want_alpha) {
        uint8_t ch[32];
        zkb_ps_fiat_shamir(ps, 32, ch);                     // fri.rs:145
        zkb_field_sample(ch, 32, alpha_out);                // fri.rs:146
    }
    return 0;
}

static int push_last_codeword(zkb_fri_layers* L, zkb_ps* ps) {
    const uint64_t R = L->rounds;                           // fri.rs:166
    if (L->last_cw.size() == L->len[R - 1] * 16) return zkb_ps_push_codeword(ps, L->last_cw.data(), L->len[R - 1]);
    std::vector<uint8_t> last(L->len[R - 1] * 16);
    ZKB_TRY(zkb_fri_layer_codeword(L, R - 1, last.data()));
    zkb_ps_push_codeword(ps, last.data(), L->len[R - 1]);
    return 0;
}

static int zkb_fri_commit_ps_body(zkb_ctx* c, const zkb_fri_params* p, const void* codeword, size_t n, zkb_ps* ps,
                      zkb_fri_layers** out) {
    if (!c || !p || !codeword || !ps || !out) return ZKB_ERR_ARG;
    ZKB_TRY(fri_commit_impl(c, p, codeword, nullptr, 0, n, ps_callback, ps, out, ps));
    int rc = push_last_codeword(*out, ps);
    if (rc) { zkb_fri_layers_free(*out); *out = nullptr; }
    return rc;
}

static int zkb_lde_fri_commit_ps_body(zkb_ctx* c, const zkb_fri_params* p, const void* coeffs, size_t n_coeffs, zkb_ps* ps,
                          zkb_fri_layers** out) {
    if (!c || !p || !ps || !out) return ZKB_ERR_ARG;
    if (n_coeffs && !coeffs) return ZKB_ERR_ARG;
    if (n_coeffs > p->domain_length)
        return set_err(c, ZKB_ERR_TOO_LONG, "coset_lde: %zu coefficients exceed root_order %llu", n_coeffs, (unsigned long long)p->domain_length);
    ZKB_TRY(fri_commit_impl(c, p, nullptr, coeffs, n_coeffs, (size_t)p->domain_length, ps_callback, ps, out, ps));
    int rc = push_last_codeword(*out, ps);
    if (rc) { zkb_fri_layers_free(*out); *out = nullptr; }
    return rc;
}

static int zkb_fri_prove_body(zkb_ctx* c, const zkb_fri_params* p, const void* codeword, size_t n, zkb_ps* ps,
                  uint64_t* top_indices_out) {
    if (!c || !p || !ps || !top_indices_out) return ZKB_ERR_ARG;
    const uint64_t ncc = p->num_colinearity_tests;
    if (zkb_fri_num_rounds(p) < 2) return set_err(c, ZKB_ERR_ROUNDS, "FRI::prove needs at least two rounds (fri.rs:225)");
    zkb_fri_layers* L = nullptr;
    if (!codeword) return ZKB_ERR_ARG;
    ZKB_TRY(fri_commit_impl(c, p, codeword, nullptr, 0, n, ps_callback, ps, &L, ps));
    struct Guard { zkb_fri_layers* l; ~Guard() { zkb_fri_layers_free(l); } } guard{L};
    const uint64_t R = L->rounds;
    ZKB_TRY(push_last_codeword(L, ps));
    // top-level indices (fri.rs:223-228)
    uint8_t seed[32];
    zkb_ps_fiat_shamir(ps, 32, seed);
    if (ncc > L->len[R - 1]) return set_err(c, ZKB_ERR_ARG, "Cannot sample more indices than available in the last codeword");
    if (zkb_fri_sample_indices(seed, 32, L->len[1], L->len[R - 1], ncc, top_indices_out) != 0)
        return set_err(c, ZKB_ERR_ARG, "sample_indices failed");
    std::vector<uint64_t> idx(top_indices_out, top_indices_out + ncc);
    bool wire = getenv("ZKB_HOST_ASSEMBLY") == nullptr && ncc > 0;
    for (uint64_t r = 0; r < R; r++) if (L->layout[r].log_n < 1 || L->layout[r].log_n > 30) wire = false;
    if (wire) {
        // The query phase as finished objects (fri.rs:174-208): for every layer pair ncc Leafs, then ncc x (Path a, Path b, Path c), written by
        // the opening kernels at their offsets in the proof (merkle.cu k_leafs_wire / k_open_wire); all rounds queued back to back,
        // ONE copy to the host, ONE append to the proof stream.
        std::vector<uint64_t> base(R, 0);
        uint64_t seg = 0;
        for (uint64_t r = 0; r + 1 < R; r++) {
            base[r] = seg;
            seg += ncc * 57 + ncc * (2 * (9 + 72ull * L->layout[r].log_n) + (9 + 72ull * L->layout[r + 1].log_n));
        }
        const size_t per_round = ncc * 3;
        std::vector<uint64_t> hidx((R - 1) * per_round + 1);
        for (uint64_t r = 0; r + 1 < R; r++) {
            const uint64_t half = L->len[r] / 2;
            for (auto& i : idx) i %= half;                   // fri.rs:234-237
            uint64_t* ab = hidx.data() + r * per_round;
            uint64_t* cc = ab + 2 * ncc;
            for (size_t q = 0; q < ncc; q++) { ab[2 * q] = idx[q]; ab[2 * q + 1] = idx[q] + half; cc[q] = idx[q]; }
        }
        hidx.back() = 0;                                     // y_off of the single instance
        DevBuf q;
        const size_t idx_bytes = (hidx.size() * 8 + 255) & ~(size_t)255;
        ZKB_TRY(q.alloc(c, idx_bytes + seg + 16));
        uint64_t* d_idx = (uint64_t*)q.p;
        uint8_t* d_wire = (uint8_t*)q.p + idx_bytes;
        const uint64_t* d_yoff = d_idx + (R - 1) * per_round;
        ZKB_CUDA(c, cudaMemcpyAsync(d_idx, hidx.data(), hidx.size() * 8, cudaMemcpyHostToDevice, c->stream));
        for (uint64_t r = 0; r + 1 < R; r++) {
            const uint64_t half = L->len[r] / 2;
            const uint64_t pc = 9 + 72ull * L->layout[r].log_n, pn = 9 + 72ull * L->layout[r + 1].log_n, trip = 2 * pc + pn;
            const uint64_t* d_ab = d_idx + r * per_round;
            const uint64_t* d_c = d_ab + 2 * ncc;
            ZKB_TRY(fri_leafs_wire_batch(c, L->cw[r], 0, L->cw[r + 1], 0, half, d_c, ncc, 1, d_wire, d_yoff, base[r]));
            ZKB_TRY(merkle_open_wire_batch(c, L->cw[r], L->layout[r], L->nodes[r], d_ab, 2 * ncc, 1, 0, 0, d_wire, d_yoff, 2, base[r] + ncc * 57, trip, pc, false));
            ZKB_TRY(merkle_open_wire_batch(c, L->cw[r + 1], L->layout[r + 1], L->nodes[r + 1], d_c, ncc, 1, 0, 0, d_wire, d_yoff, 1,
                                           base[r] + ncc * 57 + 2 * pc, trip, 0, false));
        }
        uint8_t* dst = ps_body_extend(ps, seg);              // straight into the (pinned) body: no staging copy
        ZKB_CUDA(c, cudaMemcpyAsync(dst, d_wire, seg, cudaMemcpyDeviceToHost, c->stream));
        ZKB_CUDA(c, cudaStreamSynchronize(c->stream));
        ps->has_field = true;                                // Leafs carry field elements
        return 0;
    }
    for (uint64_t r = 0; r + 1 < R; r++) {
        const uint64_t half = L->len[r] / 2;
        for (auto& i : idx) i %= half;                       // fri.rs:234-237
        const size_t d_cur = L->layout[r].log_n, d_nxt = L->layout[r + 1].log_n;
        std::vector<uint8_t> leafs(ncc * 48), paths(ncc * (2 * d_cur + d_nxt) * 64);
        ZKB_TRY(zkb_fri_query(L, r, idx.data(), ncc, leafs.data(), paths.data()));
        for (uint64_t s = 0; s < ncc; s++)                   // fri.rs:189-195
            zkb_ps_push_leafs(ps, leafs.data() + 48 * s, leafs.data() + 48 * s + 16, leafs.data() + 48 * s + 32);
        const uint8_t* q = paths.data();
        for (uint64_t s = 0; s < ncc; s++) {                 // fri.rs:198-202
            zkb_ps_push_path(ps, q, d_cur); q += d_cur * 64;
            zkb_ps_push_path(ps, q, d_cur); q += d_cur * 64;
            zkb_ps_push_path(ps, q, d_nxt); q += d_nxt * 64;
        }
    }
    return 0;
}

int zkb_fri_commit(zkb_ctx* c, const zkb_fri_params* p, const void* codeword, size_t n, zkb_fs_callback fs, void* user, zkb_fri_layers** out) {
    if (!c) return ZKB_ERR_ARG;
    ZKB_ABI_GUARD(c, return zkb_fri_commit_body(c, p, codeword, n, fs, user, out);)
}

int zkb_lde_fri_commit(zkb_ctx* c, const zkb_fri_params* p, const void* coeffs, size_t n_coeffs, zkb_fs_callback fs, void* user, zkb_fri_layers** out) {
    if (!c) return ZKB_ERR_ARG;
    ZKB_ABI_GUARD(c, return zkb_lde_fri_commit_body(c, p, coeffs, n_coeffs, fs, user, out);)
}

int zkb_fri_commit_ps(zkb_ctx* c, const zkb_fri_params* p, const void* codeword, size_t n, zkb_ps* ps, zkb_fri_layers** out) {
    if (!c) return ZKB_ERR_ARG;
    ZKB_ABI_GUARD(c, return zkb_fri_commit_ps_body(c, p, codeword, n, ps, out);)
}

int zkb_lde_fri_commit_ps(zkb_ctx* c, const zkb_fri_params* p, const void* coeffs, size_t n_coeffs, zkb_ps* ps, zkb_fri_layers** out) {
    if (!c) return ZKB_ERR_ARG;
    ZKB_ABI_GUARD(c, return zkb_lde_fri_commit_ps_body(c, p, coeffs, n_coeffs, ps, out);)
}

int zkb_fri_prove(zkb_ctx* c, const zkb_fri_params* p, const void* codeword, size_t n, zkb_ps* ps, uint64_t* top_indices_out) {
    if (!c) return ZKB_ERR_ARG;
    ZKB_ABI_GUARD(c, return zkb_fri_prove_body(c, p, codeword, n, ps, top_indices_out);)
}

int zkb_fri_query(zkb_fri_layers* l, uint64_t r, const uint64_t* idx_c, size_t ncc, uint8_t* leafs_out, uint8_t* paths_out) {
    if (!l) return ZKB_ERR_ARG;
    ZKB_ABI_GUARD(l->ctx, return zkb_fri_query_body(l, r, idx_c, ncc, leafs_out, paths_out);)
}

}  // extern "C"
