// TEST INFRASTRUCTURE - never shipped, never on a product path.
// A CPU stand-in for the GPU entry points of libzkb200.so, built on the C oracle (oracle/libzkoracle.so),
// so that the HOST logic of include/zk_impl.hpp (packing, error mapping, proof-stream assembly in the
// callback path of FRI::prove, FRI::verify) can be exercised by `pytest -m "not gpu"` in a container
// without a GPU.  Linked INTO the test executable, its definitions interpose the library's for calls made
// from the executable; the library's host-only functions (zkb_ps_*, zkb_field_*, hashes, sample_indices,
// Merkle verify) stay the real ones.  The -m gpu run links the same tests against the real library only.
// Each function follows the reference lines named in include/zkb200.h.
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "zkb200.h"

typedef unsigned __int128 u128;
extern "C" {
void zo_mul(const u128*, const u128*, u128*);
void zo_inv(const u128*, u128*);
int zo_ntt(const u128* root, const u128* in, size_t n_in, u128* out);
int zo_intt(const u128* root, const u128* in, size_t n_in, u128* out);
void zo_scale(const u128* factor, const u128* in, size_t n, u128* out);
int zo_coset_lde(const u128* omega, size_t order, const u128* offset, const u128* coeffs, size_t n, u128* out);
int zo_merkle(const u128* vals, size_t n, uint8_t root[64], uint8_t* nodes);
void zo_fri_fold(const u128* cw, size_t n, const u128* alpha, const u128* offset, const u128* omega, u128* out);
}

struct zkb_ctx { std::string err; };
struct zkb_tree { std::vector<u128> vals; std::vector<uint8_t> nodes; uint8_t root[64]; };
struct zkb_fri_layers { std::vector<zkb_tree> layers; };

static u128 ld(const uint8_t p[16]) { u128 v; memcpy(&v, p, 16); return v; }
static size_t next_pow2(size_t n) { size_t p = 1; while (p < n) p <<= 1; return p; }
static int fail(zkb_ctx* c, int code, const char* msg) { if (c) c->err = msg; return code; }
static u128 fmul(u128 a, u128 b) { u128 r; zo_mul(&a, &b, &r); return r; }
static u128 finv(u128 a) { u128 r; zo_inv(&a, &r); return r; }
static u128 fpow(u128 a, u128 e) { u128 r = 1; for (int i = 127; i >= 0; i--) { r = fmul(r, r); if ((e >> i) & 1) r = fmul(r, a); } return r; }
static long degree(const u128* p, size_t n) { long d = -1; for (size_t i = 0; i < n; i++) if (p[i]) d = (long)i; return d; }
static int build_tree(zkb_ctx* c, const u128* vals, size_t n, zkb_tree& t) {
    if (n == 0 || (n & (n - 1))) return fail(c, ZKB_ERR_NOT_POW2, "length must be power of two");
    t.vals.assign(vals, vals + n);
    t.nodes.resize((2 * n - 1) * 64);
    zo_merkle(vals, n, t.root, t.nodes.data());
    return 0;
}

extern "C" {

int zkb_ctx_create(int, void*, zkb_ctx** out) { *out = new zkb_ctx(); return 0; }
void zkb_ctx_destroy(zkb_ctx* c) { delete c; }
const char* zkb_last_error(const zkb_ctx* c) { return c ? c->err.c_str() : "no context"; }

int zkb_ntt(zkb_ctx* c, const uint8_t root[16], const void* in, size_t n_in, void* out) {
    if (n_in == 0) return fail(c, ZKB_ERR_EMPTY, "ntt of an empty vector (ntt.rs:11)");
    u128 r = ld(root);
    zo_ntt(&r, (const u128*)in, n_in, (u128*)out);
    return 0;
}
int zkb_intt(zkb_ctx*, const uint8_t root[16], const void* in, size_t n_in, void* out) {
    u128 r = ld(root);
    zo_intt(&r, (const u128*)in, n_in, (u128*)out);
    return 0;
}
int zkb_poly_scale(zkb_ctx*, const uint8_t factor[16], const void* coeffs, size_t n, void* out) {
    u128 f = ld(factor);
    zo_scale(&f, (const u128*)coeffs, n, (u128*)out);
    return 0;
}
int zkb_coset_lde(zkb_ctx* c, const uint8_t omega[16], uint64_t order, const uint8_t offset[16], const void* coeffs, size_t n, void* out) {
    if (n > order) return fail(c, ZKB_ERR_TOO_LONG, "attempt to subtract with overflow (ntt_arithmetics.rs:168)");
    u128 w = ld(omega), o = ld(offset);
    zo_coset_lde(&w, order, &o, (const u128*)coeffs, n, (u128*)out);
    return 0;
}
static int check_root(zkb_ctx* c, u128 root, uint64_t order) {
    if (fpow(root, order) != 1) return fail(c, ZKB_ERR_ROOT_ORDER, "supplied root does not have supplied root_order");
    if (fpow(root, order / 2) == 1) return fail(c, ZKB_ERR_ROOT_ORDER, "supplied root is not a primitive of root_order");
    return 0;
}
// shared body of fast_multiply (ntt_arithmetics.rs:26-63) and fast_coset_divide (:258-309)
static void transform_pair(u128 root, uint64_t order, u128 deg, const u128* scale_by, const u128* l, size_t nl, const u128* r, size_t nr,
                           bool divide, size_t result_len, u128* out, size_t* n_out) {
    while (deg < order / 2) { root = fmul(root, root); order /= 2; }
    auto inner = [&](const u128* p, size_t n) {
        std::vector<u128> v(p, p + n);
        if (scale_by) zo_scale(scale_by, p, n, v.data());
        if (v.size() < order) v.resize(order, 0);
        std::vector<u128> o(next_pow2(v.size()));
        zo_ntt(&root, v.data(), v.size(), o.data());
        return o;
    };
    std::vector<u128> a = inner(l, nl), b = inner(r, nr), h(order), co(next_pow2(order));
    for (size_t i = 0; i < order; i++) h[i] = divide ? fmul(a[i], finv(b[i])) : fmul(a[i], b[i]);
    zo_intt(&root, h.data(), order, co.data());
    size_t n = result_len < co.size() ? result_len : co.size();
    if (order < 2) n = result_len < 1 ? result_len : 1;
    memcpy(out, co.data(), n * 16);
    *n_out = n;
}
int zkb_poly_mul(zkb_ctx* c, const uint8_t root[16], uint64_t root_order, const void* lhs, size_t nl, const void* rhs, size_t nr, void* out,
                 size_t* n_out) {
    if (int rc = check_root(c, ld(root), root_order)) return rc;
    long dl = degree((const u128*)lhs, nl), dr = degree((const u128*)rhs, nr);
    *n_out = 0;
    if (dl < 0 || dr < 0) return 0;
    transform_pair(ld(root), root_order, (u128)(dl + dr), nullptr, (const u128*)lhs, nl, (const u128*)rhs, nr, false, (size_t)(dl + dr + 1),
                   (u128*)out, n_out);
    return 0;
}
int zkb_coset_div(zkb_ctx* c, const uint8_t root[16], uint64_t root_order, const uint8_t offset[16], const void* lhs, size_t nl, const void* rhs,
                  size_t nr, void* out, size_t* n_out) {
    if (int rc = check_root(c, ld(root), root_order)) return rc;
    long dl = degree((const u128*)lhs, nl), dr = degree((const u128*)rhs, nr);
    *n_out = 0;
    if (dr < 0) return fail(c, ZKB_ERR_DIV_ZERO, "cannot divide by zero polynomial");
    if (dl < 0) return 0;
    if (dl < dr) return fail(c, ZKB_ERR_DEGREE, "cannot divide by polynomial of larger degree");
    u128 off = ld(offset), off_inv = finv(off);
    transform_pair(ld(root), root_order, (u128)dl, &off, (const u128*)lhs, nl, (const u128*)rhs, nr, true, (size_t)(dl - dr + 1), (u128*)out, n_out);
    std::vector<u128> s(*n_out);
    zo_scale(&off_inv, (const u128*)out, *n_out, s.data());
    memcpy(out, s.data(), *n_out * 16);
    return 0;
}

int zkb_merkle_commit(zkb_ctx* c, const void* vals, size_t n, uint8_t root[64]) {
    zkb_tree t;
    if (int rc = build_tree(c, (const u128*)vals, n, t)) return rc;
    memcpy(root, t.root, 64);
    return 0;
}
int zkb_merkle_build(zkb_ctx* c, const void* vals, size_t n, zkb_tree** tree) {
    zkb_tree* t = new zkb_tree();
    if (int rc = build_tree(c, (const u128*)vals, n, *t)) { delete t; return rc; }
    *tree = t;
    return 0;
}
int zkb_merkle_root(const zkb_tree* t, uint8_t root[64]) { memcpy(root, t->root, 64); return 0; }
static void open_one(const zkb_tree& t, uint64_t idx, uint8_t* out) {          // merkle_root.rs:34-53: sibling per level, bottom-up
    size_t n = t.vals.size(), off = 0;
    for (size_t w = n; w > 1; w >>= 1) {
        memcpy(out, &t.nodes[(off + (idx ^ 1)) * 64], 64);
        out += 64;
        off += w;
        idx >>= 1;
    }
}
int zkb_merkle_open(zkb_tree* t, const uint64_t* idx, size_t k, uint8_t* paths_out) {
    size_t depth = 0;
    while (((size_t)1 << depth) < t->vals.size()) depth++;
    for (size_t s = 0; s < k; s++) {
        if (idx[s] >= t->vals.size()) return ZKB_ERR_INDEX;
        open_one(*t, idx[s], paths_out + s * depth * 64);
    }
    return 0;
}
void zkb_merkle_free(zkb_tree* t) { delete t; }

int zkb_fri_commit(zkb_ctx* c, const zkb_fri_params* p, const void* codeword, size_t n, zkb_fs_callback fs, void* user, zkb_fri_layers** out) {
    if (n != p->domain_length) return fail(c, ZKB_ERR_LENGTH, "Length of the domain doesnt match the length of initial codeword");
    const uint64_t R = zkb_fri_num_rounds(p);
    if (R < 1) return fail(c, ZKB_ERR_ROUNDS, "FRI needs at least one round");
    zkb_fri_layers* L = new zkb_fri_layers();
    L->layers.resize(R);
    std::vector<u128> cw((const u128*)codeword, (const u128*)codeword + n);
    u128 omega = ld(p->omega), offset = ld(p->offset);
    for (uint64_t r = 0; r < R; r++) {                                         // fri.rs:128-163
        build_tree(c, cw.data(), cw.size(), L->layers[r]);
        uint8_t alpha[16] = {0};
        if (fs(user, (uint32_t)r, L->layers[r].root, r + 1 < R, alpha) != 0) { delete L; return fail(c, ZKB_ERR_CALLBACK, "Fiat-Shamir callback failed"); }
        if (r + 1 == R) break;
        u128 a = ld(alpha);
        std::vector<u128> next(cw.size() / 2);
        zo_fri_fold(cw.data(), cw.size(), &a, &offset, &omega, next.data());
        cw.swap(next);
        omega = fmul(omega, omega);
        offset = fmul(offset, offset);
    }
    *out = L;
    return 0;
}
uint64_t zkb_fri_layer_count(const zkb_fri_layers* l) { return l->layers.size(); }
uint64_t zkb_fri_layer_len(const zkb_fri_layers* l, uint64_t r) { return l->layers[r].vals.size(); }
int zkb_fri_layer_root(const zkb_fri_layers* l, uint64_t r, uint8_t root[64]) { memcpy(root, l->layers[r].root, 64); return 0; }
int zkb_fri_layer_codeword(zkb_fri_layers* l, uint64_t r, void* out) { memcpy(out, l->layers[r].vals.data(), l->layers[r].vals.size() * 16); return 0; }
int zkb_fri_query(zkb_fri_layers* l, uint64_t r, const uint64_t* idx_c, size_t ncc, uint8_t* leafs_out, uint8_t* paths_out) {   // fri.rs:174-208
    const zkb_tree &cur = l->layers[r], &nxt = l->layers[r + 1];
    size_t half = cur.vals.size() / 2, d_cur = 0;
    while (((size_t)1 << d_cur) < cur.vals.size()) d_cur++;
    for (size_t s = 0; s < ncc; s++) {
        uint64_t a = idx_c[s], b = a + half;
        memcpy(leafs_out + 48 * s, &cur.vals[a], 16);
        memcpy(leafs_out + 48 * s + 16, &cur.vals[b], 16);
        memcpy(leafs_out + 48 * s + 32, &nxt.vals[a], 16);
        uint8_t* o = paths_out + s * (3 * d_cur - 1) * 64;
        open_one(cur, a, o);
        open_one(cur, b, o + d_cur * 64);
        open_one(nxt, a, o + 2 * d_cur * 64);
    }
    return 0;
}
void zkb_fri_layers_free(zkb_fri_layers* l) { delete l; }

// FRI::prove against the library's own stream (fri.rs:210-248), assembled from the pieces above
int zkb_fri_prove(zkb_ctx* c, const zkb_fri_params* p, const void* codeword, size_t n, zkb_ps* ps, uint64_t* top_out) {
    if (zkb_fri_num_rounds(p) < 2) return fail(c, ZKB_ERR_ROUNDS, "FRI::prove needs at least two rounds (fri.rs:225)");
    zkb_fs_callback cb = [](void* user, uint32_t, const uint8_t root[64], int want_alpha, uint8_t alpha_out[16]) -> int {
        zkb_ps* s = static_cast<zkb_ps*>(user);
        zkb_ps_push_root(s, root, 64);
        if (want_alpha) { uint8_t ch[32]; zkb_ps_fiat_shamir(s, 32, ch); zkb_field_sample(ch, 32, alpha_out); }
        return 0;
    };
    zkb_fri_layers* L = nullptr;
    if (int rc = zkb_fri_commit(c, p, codeword, n, cb, ps, &L)) return rc;
    const uint64_t R = L->layers.size(), ncc = p->num_colinearity_tests;
    zkb_ps_push_codeword(ps, L->layers[R - 1].vals.data(), L->layers[R - 1].vals.size());
    uint8_t seed[32];
    zkb_ps_fiat_shamir(ps, 32, seed);
    zkb_fri_sample_indices(seed, 32, L->layers[1].vals.size(), L->layers[R - 1].vals.size(), ncc, top_out);
    std::vector<uint64_t> idx(top_out, top_out + ncc);
    for (uint64_t r = 0; r + 1 < R; r++) {
        size_t len = L->layers[r].vals.size(), d = 0;
        while (((size_t)1 << d) < len) d++;
        for (auto& i : idx) i %= len / 2;
        std::vector<uint8_t> leafs(ncc * 48), paths(ncc * (3 * d - 1) * 64);
        zkb_fri_query(L, r, idx.data(), ncc, leafs.data(), paths.data());
        for (size_t s = 0; s < ncc; s++) zkb_ps_push_leafs(ps, &leafs[48 * s], &leafs[48 * s + 16], &leafs[48 * s + 32]);
        for (size_t s = 0; s < ncc; s++) {
            uint8_t* o = &paths[s * (3 * d - 1) * 64];
            zkb_ps_push_path(ps, o, d);
            zkb_ps_push_path(ps, o + d * 64, d);
            zkb_ps_push_path(ps, o + 2 * d * 64, d - 1);
        }
    }
    delete L;
    return 0;
}

}  // extern "C"
