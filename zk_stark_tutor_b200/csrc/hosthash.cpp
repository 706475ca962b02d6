// hosthash.cpp - host BLAKE2b-512, SHAKE256, proof stream, index sampling, Merkle verify.
//
// Replaces (reference file:line):
//   blake2b512            src/crypto/blake2b512.rs:4-14      (crate blake2 0.10.6, RFC 7693)
//   shake256              src/crypto/shake256.rs:7-19        (crate sha3 0.10.8, FIPS 202)
//   ProofStream           src/proof_stream.rs:14-83, src/stark/proof_stream_enum.rs:67-190,
//                         src/rescue_prime/proof_stream.rs:9-62, src/utils/digest.rs:17-33
//   FRI::sample_index(es) src/fri.rs:60-113
//   MerkleRoot::verify    src/merkle_root.rs:69-95
// These stay on the host by design: they are tiny, serial, and sit between FRI rounds.
#include <stdlib.h>
#include <string.h>
#include <cuda_runtime.h>
#include <algorithm>
#include <mutex>
#include "hosthash.hpp"
#include "fe128.cuh"
#include "keccak.cuh"
#include "../../include/zkb200.h"

namespace zkb {

// ---------------------------------------------------------------- BLAKE2b (any length)
static const uint64_t IV[8] = {0x6a09e667f3bcc908ULL, 0xbb67ae8584caa73bULL, 0x3c6ef372fe94f82bULL,
                               0xa54ff53a5f1d36f1ULL, 0x510e527fade682d1ULL, 0x9b05688c2b3e6c1fULL,
                               0x1f83d9abfb41bd6bULL, 0x5be0cd19137e2179ULL};
static const uint8_t SIGMA[12][16] = {
    {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3},
    {11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4}, {7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8},
    {9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13}, {2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9},
    {12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11}, {13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10},
    {6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5}, {10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0},
    {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3}};
static inline uint64_t rotr(uint64_t x, int n) { return (x >> n) | (x << (64 - n)); }

static void compress(uint64_t h[8], const uint8_t block[128], uint64_t t, bool last) {
    uint64_t m[16], v[16];
    for (int i = 0; i < 16; i++) {
        uint64_t w = 0;
        for (int b = 7; b >= 0; b--) w = (w << 8) | block[8 * i + b];
        m[i] = w;
    }
    for (int i = 0; i < 8; i++) { v[i] = h[i]; v[i + 8] = IV[i]; }
    v[12] ^= t;
    if (last) v[14] = ~v[14];
#define G(a, b, c, d, x, y)                                                      \
    v[a] = v[a] + v[b] + (x); v[d] = rotr(v[d] ^ v[a], 32); v[c] = v[c] + v[d]; \
    v[b] = rotr(v[b] ^ v[c], 24); v[a] = v[a] + v[b] + (y);                      \
    v[d] = rotr(v[d] ^ v[a], 16); v[c] = v[c] + v[d]; v[b] = rotr(v[b] ^ v[c], 63);
    for (int r = 0; r < 12; r++) {
        const uint8_t* s = SIGMA[r];
        G(0, 4, 8, 12, m[s[0]], m[s[1]]) G(1, 5, 9, 13, m[s[2]], m[s[3]])
        G(2, 6, 10, 14, m[s[4]], m[s[5]]) G(3, 7, 11, 15, m[s[6]], m[s[7]])
        G(0, 5, 10, 15, m[s[8]], m[s[9]]) G(1, 6, 11, 12, m[s[10]], m[s[11]])
        G(2, 7, 8, 13, m[s[12]], m[s[13]]) G(3, 4, 9, 14, m[s[14]], m[s[15]])
    }
#undef G
    for (int i = 0; i < 8; i++) h[i] ^= v[i] ^ v[i + 8];
}

void host_blake2b512(const uint8_t* msg, size_t len, uint8_t out[64]) {
    uint64_t h[8];
    for (int i = 0; i < 8; i++) h[i] = IV[i];
    h[0] ^= 0x01010040ULL;
    uint8_t block[128];
    size_t off = 0;
    while (len - off > 128) {               // all but the last block
        compress(h, msg + off, (uint64_t)(off + 128), false);
        off += 128;
    }
    memset(block, 0, 128);
    if (len - off) memcpy(block, msg + off, len - off);
    compress(h, block, (uint64_t)len, true);
    for (int i = 0; i < 8; i++)
        for (int b = 0; b < 8; b++) out[8 * i + b] = (uint8_t)(h[i] >> (8 * b));
}

// ---------------------------------------------------------------- SHAKE256 (Keccak-f[1600])
static const uint64_t KRC[24] = {
    0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808aULL, 0x8000000080008000ULL,
    0x000000000000808bULL, 0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL,
    0x000000000000008aULL, 0x0000000000000088ULL, 0x0000000080008009ULL, 0x000000008000000aULL,
    0x000000008000808bULL, 0x800000000000008bULL, 0x8000000000008089ULL, 0x8000000000008003ULL,
    0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800aULL, 0x800000008000000aULL,
    0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
static const int KROT[24] = {1, 3, 6, 10, 15, 21, 28, 36, 45, 55, 2, 14, 27, 41, 56, 8, 25, 43, 62, 18, 39, 61, 20, 44};
static const int KPIL[24] = {10, 7, 11, 17, 18, 3, 5, 16, 8, 21, 24, 4, 15, 23, 19, 13, 12, 2, 20, 14, 22, 9, 6, 1};
static inline uint64_t rotl(uint64_t x, int n) { return (x << n) | (x >> (64 - n)); }

static void keccakf(uint64_t st[25]) {
    for (int round = 0; round < 24; round++) {
        uint64_t bc[5];
        for (int i = 0; i < 5; i++) bc[i] = st[i] ^ st[i + 5] ^ st[i + 10] ^ st[i + 15] ^ st[i + 20];
        for (int i = 0; i < 5; i++) {
            uint64_t t = bc[(i + 4) % 5] ^ rotl(bc[(i + 1) % 5], 1);
            for (int j = 0; j < 25; j += 5) st[j + i] ^= t;
        }
        uint64_t t = st[1];
        for (int i = 0; i < 24; i++) {
            int j = KPIL[i];
            uint64_t b = st[j];
            st[j] = rotl(t, KROT[i]);
            t = b;
        }
        for (int j = 0; j < 25; j += 5) {
            for (int i = 0; i < 5; i++) bc[i] = st[j + i];
            for (int i = 0; i < 5; i++) st[j + i] ^= (~bc[(i + 1) % 5]) & bc[(i + 2) % 5];
        }
        st[0] ^= KRC[round];
    }
}

void host_shake256(const uint8_t* msg, size_t len, uint8_t* out, size_t out_len) {
    const size_t rate = 136;
    uint64_t st[25];
    memset(st, 0, sizeof(st));
    auto absorb_block = [&](const uint8_t* b) {
        for (size_t i = 0; i < rate / 8; i++) {
            uint64_t w = 0;
            for (int k = 7; k >= 0; k--) w = (w << 8) | b[8 * i + k];
            st[i] ^= w;
        }
        keccakf(st);
    };
    size_t off = 0;
    while (len - off >= rate) { absorb_block(msg + off); off += rate; }
    uint8_t last[136];
    memset(last, 0, rate);
    if (len - off) memcpy(last, msg + off, len - off);
    last[len - off] ^= 0x1F;               // SHAKE domain separation + first pad bit
    last[rate - 1] ^= 0x80;
    absorb_block(last);
    size_t produced = 0;
    while (produced < out_len) {
        size_t take = out_len - produced < rate ? out_len - produced : rate;
        for (size_t i = 0; i < take; i++) out[produced + i] = (uint8_t)(st[i / 8] >> (8 * (i % 8)));
        produced += take;
        if (produced < out_len) keccakf(st);
    }
}

}  // namespace zkb

namespace zkb {
// 64-byte header in front of every body block: magic + kind (1 = cudaHostAlloc, 0 = malloc)
static const uint64_t kBodyMagic = 0x7a6b62505342ull;
static const size_t kPinFrom = 256u << 10;
void* ps_body_alloc(size_t bytes) {
    uint8_t* raw = nullptr;
    uint64_t kind = 0;
    if (bytes >= kPinFrom) {
        void* p = nullptr;
        if (cudaHostAlloc(&p, bytes + 64, cudaHostAllocPortable) == cudaSuccess) { raw = (uint8_t*)p; kind = 1; }
        else cudaGetLastError();                                   // no device / out of pinned memory: pageable memory still works
    }
    if (!raw) {
        raw = (uint8_t*)malloc(bytes + 64);
        if (!raw) throw std::bad_alloc();
    }
    ((uint64_t*)raw)[0] = kBodyMagic; ((uint64_t*)raw)[1] = kind;
    return raw + 64;
}
void ps_body_free(void* p) {
    if (!p) return;
    uint8_t* raw = (uint8_t*)p - 64;
    if (((uint64_t*)raw)[1] == 1) cudaFreeHost(raw); else free(raw);
}
uint8_t* ps_body_extend(zkb_ps* ps, size_t bytes) {
    PsBody& v = ps->body;
    const size_t at = v.size(), need = at + bytes;
    // a body that reaches pinned size is given room for a whole proof at once (~1.2 MB): one cudaHostAlloc, kept by the pool afterwards
    if (v.capacity() < need) v.reserve(std::max(std::max(v.capacity() * 2, need), need >= kPinFrom ? (size_t)(1536u << 10) : need));
    v.resize(need);
    return v.data() + at;
}
}  // namespace zkb

using namespace zkb;

template <typename V>
static void put_be64(V& v, uint64_t x) {
    for (int i = 7; i >= 0; i--) v.push_back((uint8_t)(x >> (8 * i)));
}
static void le16_to_be16(const uint8_t* le, uint8_t* be) {
    for (int i = 0; i < 16; i++) be[i] = le[15 - i];
}

void zkb_ps::push(uint8_t code, const uint8_t* payload, size_t len) {
    body.push_back(code);                     // proof_stream_enum.rs:176-181
    put_be64(body, (uint64_t)len);
    body.insert(body.end(), payload, payload + len);
}
void zkb_ps::header(uint8_t out[16]) const {
    memset(out, 0, 16);                       // proof_stream_enum.rs:186-188: order or 0
    if (has_field) { out[0] = 0xCB; out[1] = 0x80; out[15] = 0x01; }
}

void zkb_ps::transcript_read(size_t off, size_t len, uint8_t* out) const {
    uint8_t hdr[16];
    header(hdr);
    const size_t np = prefix.size();
    for (size_t i = 0; i < len; ) {                       // three segments: prefix, 16-byte header, body
        const size_t o = off + i;
        if (o < np) { size_t k = std::min(len - i, np - o); memcpy(out + i, prefix.data() + o, k); i += k; }
        else if (o < np + 16) { size_t k = std::min(len - i, np + 16 - o); memcpy(out + i, hdr + (o - np), k); i += k; }
        else { size_t k = len - i; memcpy(out + i, body.data() + (o - np - 16), k); i += k; }
    }
}
void zkb_ps::sponge_sync() {
    if (sp_field != has_field) {                          // the header changed (once per stream): start over
        memset(sp_st, 0, sizeof(sp_st));
        sp_absorbed = 0;
        sp_field = has_field;
    }
    const size_t total = transcript_len();
    uint8_t block[136];
    while (total - sp_absorbed >= 136) {
        transcript_read(sp_absorbed, 136, block);
        for (size_t i = 0; i < 17; i++) {
            uint64_t w = 0;
            for (int k = 7; k >= 0; k--) w = (w << 8) | block[8 * i + k];
            sp_st[i] ^= w;
        }
        keccakf(sp_st);
        sp_absorbed += 136;
    }
}
namespace zkb {
void ps_export_sponge(zkb_ps* ps, FsSponge* out) {
    ps->sponge_sync();
    memset(out, 0, sizeof(*out));
    memcpy(out->st, ps->sp_st, sizeof(ps->sp_st));
    out->fill = (uint32_t)(ps->transcript_len() - ps->sp_absorbed);
    ps->transcript_read(ps->sp_absorbed, out->fill, out->buf);
}
}  // namespace zkb

extern "C" {

void zkb_blake2b512(const uint8_t* msg, size_t len, uint8_t out[64]) { host_blake2b512(msg, len, out); }
void zkb_shake256(const uint8_t* msg, size_t len, uint8_t* out, size_t out_len) { host_shake256(msg, len, out, out_len); }

// Transcript buffers are recycled: a proof is ~1.2 MB, which malloc serves with a fresh mmap (page faults on
// every first touch, munmap on free) - measured 2.7 ms of 2.8 ms per assembled proof.  A small free list of
// bodies that keep their capacity removes that (0.3 ms per proof).
static std::mutex g_pool_mu;
static std::vector<zkb::PsBody> g_body_pool;
// <= 1024 bodies and <= 2 GB kept: several batches of ~1.2 MB signature streams in flight, or a few 35 MB proofs of 2^24-value codewords (a pinned
// block of that size costs ~10 ms to allocate - more than the proof's kernels)
static const size_t kPoolMax = 1024, kPoolKeepBytes = 128u << 20, kPoolTotalBytes = (size_t)2 << 30;
static size_t g_pool_bytes = 0;

static int zkb_ps_create_body(const uint8_t* document, size_t document_len, int is_signature, zkb_ps** out) {
    if (!out) return ZKB_ERR_ARG;
    zkb_ps* ps = new zkb_ps();
    {
        std::lock_guard<std::mutex> lk(g_pool_mu);
        if (!g_body_pool.empty()) { ps->body.swap(g_body_pool.back()); g_body_pool.pop_back(); g_pool_bytes -= ps->body.capacity(); }
    }
    ps->body.clear();
    if (is_signature) {
        uint8_t h[64];
        host_blake2b512(document, document_len, h);       // rescue_prime/proof_stream.rs:15-22
        put_be64(ps->prefix, 64);                          // :24-27 digest_prefix
        ps->prefix.insert(ps->prefix.end(), h, h + 64);
    }
    *out = ps;
    return 0;
}
void zkb_ps_free(zkb_ps* ps) {
    if (!ps) return;
    if (ps->body.capacity() >= (64u << 10) && ps->body.capacity() <= kPoolKeepBytes) {
        std::lock_guard<std::mutex> lk(g_pool_mu);
        if (g_body_pool.size() < kPoolMax && g_pool_bytes + ps->body.capacity() <= kPoolTotalBytes) {
            g_pool_bytes += ps->body.capacity();
            g_body_pool.emplace_back(); g_body_pool.back().swap(ps->body);
        }
    }
    delete ps;
}

static int zkb_ps_push_root_body(zkb_ps* ps, const uint8_t* root, size_t len) {
    if (!ps || (!root && len)) return ZKB_ERR_ARG;
    ps->push(0, root, len);
    return 0;
}
static int zkb_ps_push_codeword_body(zkb_ps* ps, const void* vals_host, size_t n) {
    if (!ps || (!vals_host && n)) return ZKB_ERR_ARG;
    std::vector<uint8_t> p(n * 16);
    for (size_t i = 0; i < n; i++) le16_to_be16((const uint8_t*)vals_host + 16 * i, p.data() + 16 * i);
    ps->push(1, p.data(), p.size());
    if (n) ps->has_field = true;
    return 0;
}
static int zkb_ps_push_path_body(zkb_ps* ps, const uint8_t* nodes, size_t count) {
    if (!ps || (!nodes && count)) return ZKB_ERR_ARG;
    // code || len || count x (u64_be(64) || 64 bytes), written in place (a proof holds ~1,350 paths)
    zkb::PsBody& v = ps->body;
    const size_t payload = count * 72, at = v.size();
    if (v.capacity() < at + 9 + payload) v.reserve(std::max(v.capacity() * 2, at + 9 + payload));
    v.resize(at + 9 + payload);
    uint8_t* o = v.data() + at;
    *o++ = 2;
    for (int i = 7; i >= 0; i--) *o++ = (uint8_t)((uint64_t)payload >> (8 * i));
    for (size_t i = 0; i < count; i++) {
        memset(o, 0, 7); o[7] = 64; o += 8;
        memcpy(o, nodes + 64 * i, 64); o += 64;
    }
    return 0;
}
static int zkb_ps_push_leafs_body(zkb_ps* ps, const uint8_t a[16], const uint8_t b[16], const uint8_t c[16]) {
    if (!ps) return ZKB_ERR_ARG;
    uint8_t p[48];
    le16_to_be16(a, p); le16_to_be16(b, p + 16); le16_to_be16(c, p + 32);
    ps->push(3, p, 48);
    ps->has_field = true;
    return 0;
}
static int zkb_ps_push_value_body(zkb_ps* ps, const uint8_t v[16]) {
    if (!ps) return ZKB_ERR_ARG;
    uint8_t p[16];
    le16_to_be16(v, p);
    ps->push(4, p, 16);
    ps->has_field = true;
    return 0;
}
static int zkb_ps_push_object_body(zkb_ps* ps, uint8_t code, const uint8_t* payload, size_t len) {
    if (!ps || (!payload && len) || code > 4) return ZKB_ERR_ARG;
    ps->push(code, payload, len);
    if (code == 3 || code == 4 || (code == 1 && len)) ps->has_field = true;   // proof_stream_enum.rs:76-125
    return 0;
}
int zkb_ps_create(const uint8_t* document, size_t document_len, int is_signature, zkb_ps** out) {
    try { return zkb_ps_create_body(document, document_len, is_signature, out); } catch (...) { return ZKB_ERR_NOMEM; }     // std::bad_alloc must not cross the C ABI; the stream is unusable afterwards
}

int zkb_ps_push_root(zkb_ps* ps, const uint8_t* root, size_t len) {
    try { return zkb_ps_push_root_body(ps, root, len); } catch (...) { return ZKB_ERR_NOMEM; }     // std::bad_alloc must not cross the C ABI; the stream is unusable afterwards
}

int zkb_ps_push_codeword(zkb_ps* ps, const void* vals_host, size_t n) {
    try { return zkb_ps_push_codeword_body(ps, vals_host, n); } catch (...) { return ZKB_ERR_NOMEM; }     // std::bad_alloc must not cross the C ABI; the stream is unusable afterwards
}

int zkb_ps_push_path(zkb_ps* ps, const uint8_t* nodes, size_t count) {
    try { return zkb_ps_push_path_body(ps, nodes, count); } catch (...) { return ZKB_ERR_NOMEM; }     // std::bad_alloc must not cross the C ABI; the stream is unusable afterwards
}

int zkb_ps_push_leafs(zkb_ps* ps, const uint8_t a[16], const uint8_t b[16], const uint8_t c[16]) {
    try { return zkb_ps_push_leafs_body(ps, a, b, c); } catch (...) { return ZKB_ERR_NOMEM; }     // std::bad_alloc must not cross the C ABI; the stream is unusable afterwards
}

int zkb_ps_push_value(zkb_ps* ps, const uint8_t v[16]) {
    try { return zkb_ps_push_value_body(ps, v); } catch (...) { return ZKB_ERR_NOMEM; }     // std::bad_alloc must not cross the C ABI; the stream is unusable afterwards
}

int zkb_ps_push_object(zkb_ps* ps, uint8_t code, const uint8_t* payload, size_t len) {
    try { return zkb_ps_push_object_body(ps, code, payload, len); } catch (...) { return ZKB_ERR_NOMEM; }     // std::bad_alloc must not cross the C ABI; the stream is unusable afterwards
}

size_t zkb_ps_digest(const zkb_ps* ps, uint8_t* out, size_t cap) {
    if (!ps) return 0;
    size_t total = 16 + ps->body.size();
    if (out && cap) {
        uint8_t hdr[16];
        ps->header(hdr);
        size_t n = cap < 16 ? cap : 16;
        memcpy(out, hdr, n);
        if (cap > 16) memcpy(out + 16, ps->body.data(), (cap - 16 < ps->body.size()) ? cap - 16 : ps->body.size());
    }
    return total;
}
int zkb_ps_fiat_shamir(const zkb_ps* ps_c, size_t num_bytes, uint8_t* out) {
    if (!ps_c || !out) return ZKB_ERR_ARG;
    zkb_ps* ps = const_cast<zkb_ps*>(ps_c);                // the sponge is a cache of the transcript's complete blocks
    ps->sponge_sync();
    uint64_t st[25];
    memcpy(st, ps->sp_st, sizeof(st));
    uint8_t last[136];
    const size_t rem = ps->transcript_len() - ps->sp_absorbed;   // < 136
    memset(last, 0, sizeof(last));
    ps->transcript_read(ps->sp_absorbed, rem, last);
    last[rem] ^= 0x1F;
    last[135] ^= 0x80;
    for (size_t i = 0; i < 17; i++) {
        uint64_t w = 0;
        for (int k = 7; k >= 0; k--) w = (w << 8) | last[8 * i + k];
        st[i] ^= w;
    }
    keccakf(st);
    size_t produced = 0;
    while (produced < num_bytes) {
        const size_t take = std::min<size_t>(num_bytes - produced, 136);
        for (size_t i = 0; i < take; i++) out[produced + i] = (uint8_t)(st[i / 8] >> (8 * (i % 8)));
        produced += take;
        if (produced < num_bytes) keccakf(st);
    }
    return 0;
}

int zkb_fri_sample_indices(const uint8_t* seed, size_t seed_len, uint64_t size, uint64_t reduced_size,
                           uint64_t number, uint64_t* out) {
    // fri.rs:85-113 (+ sample_index :60-83)
    if (!seed || !out || size == 0 || reduced_size == 0) return ZKB_ERR_ARG;
    if (number > 2 * reduced_size || number > reduced_size) return ZKB_ERR_ARG;
    uint32_t bit = 0;
    while ((size >> (bit + 1)) != 0) bit++;
    size_t nbytes = bit / 8 + 1;
    std::vector<uint8_t> msg(seed, seed + seed_len);
    std::vector<uint64_t> reduced;
    uint64_t found = 0;
    while (found < number) {
        uint8_t h[64];
        host_blake2b512(msg.data(), msg.size(), h);
        uint64_t acc = 0;
        for (size_t i = 64 - nbytes; i < 64; i++) acc = (acc << 8) ^ h[i];
        uint64_t idx = acc % size, red = idx % reduced_size;
        msg.push_back(0);                        // counter = one more zero byte
        bool seen = false;
        for (uint64_t r : reduced) if (r == red) { seen = true; break; }
        if (!seen) { out[found++] = idx; reduced.push_back(red); }
    }
    return 0;
}

static void leaf_digest(const uint8_t leaf_le[16], uint8_t out[64]) {
    // decimal ASCII of the u128 (field_element.rs:46-50)
    unsigned __int128 v = 0;
    for (int i = 15; i >= 0; i--) v = (v << 8) | leaf_le[i];
    char tmp[40], buf[40];
    int n = 0;
    if (v == 0) tmp[n++] = '0';
    while (v) { tmp[n++] = (char)('0' + (int)(v % 10)); v /= 10; }
    for (int i = 0; i < n; i++) buf[i] = tmp[n - 1 - i];
    host_blake2b512((const uint8_t*)buf, (size_t)n, out);
}

int zkb_merkle_verify(const uint8_t root[64], uint64_t index, const uint8_t* path, size_t path_len,
                      const uint8_t leaf[16]) {
    // merkle_root.rs:69-95
    if (!root || !path || !leaf || path_len == 0 || path_len > 63) return ZKB_ERR_ARG;
    if (index >= (1ull << path_len)) return ZKB_ERR_INDEX;
    uint8_t h[64], cat[128];
    leaf_digest(leaf, h);
    for (size_t i = 0; i < path_len; i++) {
        if ((index & 1) == 0) { memcpy(cat, h, 64); memcpy(cat + 64, path + 64 * i, 64); }
        else { memcpy(cat, path + 64 * i, 64); memcpy(cat + 64, h, 64); }
        host_blake2b512(cat, 128, h);
        index >>= 1;
    }
    return memcmp(h, root, 64) == 0 ? 1 : 0;
}

}  // extern "C"
