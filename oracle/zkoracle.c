/* zkoracle.c - CPU oracle in C for sizes where the Python oracle is too slow.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Never linked into, loaded by or
 * called from the product library; only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs load it (through ctypes).
 *
 * Two families of entry points, both restating SpekalsG3/zk-stark-tutor semantics
 * on canonical little-endian u128 values (the reference cannot be compiled here:
 * no rustc/cargo in the image):
 *
 *   zo_*  "fast":     result-exact restatement with a Montgomery multiplier, a
 *                     table-driven NTT and an iterative Merkle tree.  Used as the
 *                     large-size checker.  Validated against the Python oracle
 *                     (which carries the reference's KATs) in tests/test_oracle_c.py.
 *   zr_*  "faithful": the reference's own ALGORITHMS - bit-serial double-and-add
 *                     mul_mod (src/field/field.rs:117-131), unsigned xgcd inverse
 *                     (src/utils/xgcd.rs:22-48 / field.rs:160-169), MSB-first pow
 *                     (src/field/field_element.rs:108-143), per-coefficient pow in
 *                     scale (src/field/polynomial.rs:109-121), bit-reverse copy +
 *                     powtable NTT (src/fft/ntt.rs:7-49), per-element pow + inverse
 *                     in the FRI fold (src/fri.rs:152-159), recursive allocating
 *                     Merkle with decimal-string leaves (src/merkle_root.rs:7-32).
 *                     Used ONLY as the timed single-core CPU baseline ("port").
 *
 * BLAKE2b-512 below is RFC 7693 (the reference takes it from crate blake2 0.10.6,
 * src/crypto/blake2b512.rs:4-14); pinned by the reference's KATs in the tests.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>

typedef unsigned __int128 u128;
typedef uint64_t u64;

static const u128 P = ((u128)0xCB80000000000000ULL << 64) | 1ULL;   /* field.rs:9-10 */

/* ------------------------------------------------------------------ field (fast) */
static inline u128 f_add(u128 a, u128 b) {
    u128 s = a + b;
    if (s < a || s >= P) s -= P;           /* p > 2^127: the sum may wrap 2^128 */
    return s;
}
static inline u128 f_sub(u128 a, u128 b) { return a >= b ? a - b : a + (P - b); }

static inline void mul128(u128 a, u128 b, u128 *hi, u128 *lo) {
    u64 a0 = (u64)a, a1 = (u64)(a >> 64), b0 = (u64)b, b1 = (u64)(b >> 64);
    u128 p00 = (u128)a0 * b0, p01 = (u128)a0 * b1, p10 = (u128)a1 * b0, p11 = (u128)a1 * b1;
    u128 mid = (p00 >> 64) + (u64)p01 + (u64)p10;
    *lo = (u128)(u64)p00 | (mid << 64);
    *hi = p11 + (p01 >> 64) + (p10 >> 64) + (mid >> 64);
}
/* Montgomery product a*b*2^-128 mod p.  -p^-1 mod 2^128 = p - 2 (p = 1 + 407*2^119,
 * (407*2^119)^2 = 0 mod 2^128). */
static inline u128 mont_mul(u128 a, u128 b) {
    u128 thi, tlo, mhi, mlo;
    mul128(a, b, &thi, &tlo);
    u128 m = tlo * (P - 2);
    mul128(m, P, &mhi, &mlo);
    u128 carry = (tlo != 0);               /* tlo + mlo == 0 mod 2^128, carries iff tlo != 0 */
    u128 s = thi + mhi;
    int ov = s < thi;
    u128 s2 = s + carry;
    ov |= s2 < s;
    if (ov || s2 >= P) s2 -= P;
    return s2;
}
static u128 R2;                             /* 2^256 mod p */
static void init_consts(void) {
    if (R2) return;
    u128 r = 1;                             /* 2^256 mod p by 256 doublings */
    for (int i = 0; i < 256; i++) r = f_add(r, r);
    R2 = r;
}
static inline u128 f_mul(u128 a, u128 b) { return mont_mul(mont_mul(a, b), R2); }
static inline u128 to_mont(u128 a) { return mont_mul(a, R2); }
static u128 f_pow(u128 a, u128 e) {
    u128 acc = 1;
    for (int i = 127; i >= 0; i--) {
        acc = f_mul(acc, acc);
        if ((e >> i) & 1) acc = f_mul(acc, a);
    }
    return acc;
}
static u128 f_inv(u128 a) { return f_pow(a, P - 2); }

/* ------------------------------------------------------------------ BLAKE2b-512 */
static const u64 B2_IV[8] = {
    0x6a09e667f3bcc908ULL, 0xbb67ae8584caa73bULL, 0x3c6ef372fe94f82bULL, 0xa54ff53a5f1d36f1ULL,
    0x510e527fade682d1ULL, 0x9b05688c2b3e6c1fULL, 0x1f83d9abfb41bd6bULL, 0x5be0cd19137e2179ULL};
static const uint8_t B2_SIGMA[12][16] = {
    {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15},
    {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3},
    {11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4},
    {7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8},
    {9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13},
    {2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9},
    {12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11},
    {13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10},
    {6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5},
    {10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0},
    {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15},
    {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3}};
static inline u64 rotr64(u64 x, int n) { return (x >> n) | (x << (64 - n)); }
#define B2_G(a, b, c, d, x, y) do { \
    v[a] = v[a] + v[b] + (x); v[d] = rotr64(v[d] ^ v[a], 32); \
    v[c] = v[c] + v[d];       v[b] = rotr64(v[b] ^ v[c], 24); \
    v[a] = v[a] + v[b] + (y); v[d] = rotr64(v[d] ^ v[a], 16); \
    v[c] = v[c] + v[d];       v[b] = rotr64(v[b] ^ v[c], 63); } while (0)

/* One-block unkeyed BLAKE2b-512 of msg[0..len), len <= 128 (all the path needs:
 * leaves are 1..39 bytes, nodes 128 bytes; len 0 also handled per RFC). */
static void blake2b512_1block(const uint8_t *msg, size_t len, uint8_t out[64]) {
    u64 h[8], v[16], m[16];
    uint8_t blk[128];
    memset(blk, 0, 128);
    memcpy(blk, msg, len);
    memcpy(m, blk, 128);                   /* little-endian host */
    for (int i = 0; i < 8; i++) h[i] = B2_IV[i];
    h[0] ^= 0x01010040ULL;                 /* digest 64, no key, fanout 1, depth 1 */
    for (int i = 0; i < 8; i++) { v[i] = h[i]; v[i + 8] = B2_IV[i]; }
    v[12] ^= (u64)len;
    v[14] = ~v[14];                        /* final block */
    for (int r = 0; r < 12; r++) {
        const uint8_t *s = B2_SIGMA[r];
        B2_G(0, 4, 8, 12, m[s[0]], m[s[1]]);   B2_G(1, 5, 9, 13, m[s[2]], m[s[3]]);
        B2_G(2, 6, 10, 14, m[s[4]], m[s[5]]);  B2_G(3, 7, 11, 15, m[s[6]], m[s[7]]);
        B2_G(0, 5, 10, 15, m[s[8]], m[s[9]]);  B2_G(1, 6, 11, 12, m[s[10]], m[s[11]]);
        B2_G(2, 7, 8, 13, m[s[12]], m[s[13]]); B2_G(3, 4, 9, 14, m[s[14]], m[s[15]]);
    }
    for (int i = 0; i < 8; i++) h[i] ^= v[i] ^ v[i + 8];
    memcpy(out, h, 64);
}

static size_t u128_to_dec(u128 v, char *buf) {     /* ASCII decimal, no padding */
    char tmp[40];
    size_t n = 0;
    if (v == 0) { buf[0] = '0'; return 1; }
    while (v) { tmp[n++] = (char)('0' + (int)(v % 10)); v /= 10; }
    for (size_t i = 0; i < n; i++) buf[i] = tmp[n - 1 - i];
    return n;
}
static void leaf_hash(u128 v, uint8_t out[64]) {   /* merkle_root.rs:25-30 */
    char buf[40];
    size_t n = u128_to_dec(v, buf);
    blake2b512_1block((const uint8_t *)buf, n, out);
}

/* ======================================================================= zo_* */
void zo_blake2b512(const uint8_t *msg, size_t len, uint8_t out[64]) { blake2b512_1block(msg, len, out); }
void zo_mul(const u128 *a, const u128 *b, u128 *out) { init_consts(); *out = f_mul(*a, *b); }
void zo_pow(const u128 *a, const u128 *e, u128 *out) { init_consts(); *out = f_pow(*a, *e); }
void zo_inv(const u128 *a, u128 *out) { init_consts(); *out = f_inv(*a); }

static size_t next_pow2(size_t n) { size_t p = 1; while (p < n) p <<= 1; return p; }

/* ntt.rs:7-49 semantics: out has next_pow2(n_in) elements (n_in < 2: copied). */
int zo_ntt(const u128 *root, const u128 *in, size_t n_in, u128 *out) {
    init_consts();
    if (n_in == 0) return -1;
    if (n_in < 2) { out[0] = in[0]; return 0; }
    size_t n = next_pow2(n_in);
    int bits = 0; while (((size_t)1 << bits) < n) bits++;
    for (size_t k = 0; k < n; k++) {
        size_t r = 0;
        for (int b = 0; b < bits; b++) r |= ((k >> b) & 1) << (bits - 1 - b);
        out[r] = k < n_in ? in[k] : 0;
    }
    u128 *pw = (u128 *)malloc(sizeof(u128) * (n / 2));      /* root^k in Montgomery form */
    u128 rm = to_mont(*root), t = to_mont(1);
    for (size_t k = 0; k < n / 2; k++) { pw[k] = t; t = mont_mul(t, rm); }
    for (size_t size = 2; size <= n; size <<= 1) {
        size_t half = size >> 1, step = n / size;
        #pragma omp parallel for schedule(static) if (n >= 65536)
        for (size_t idx = 0; idx < n / 2; idx++) {
            size_t blk = idx / half, j = idx % half;
            size_t lo = blk * size + j, hi = lo + half;
            u128 e = out[lo], o = mont_mul(out[hi], pw[j * step]);  /* data canonical, twiddle Montgomery */
            out[lo] = f_add(e, o);
            out[hi] = f_sub(e, o);
        }
    }
    free(pw);
    return 0;
}
int zo_intt(const u128 *root, const u128 *in, size_t n_in, u128 *out) {
    init_consts();
    if (n_in < 2) { if (n_in) out[0] = in[0]; return 0; }
    size_t n = next_pow2(n_in);
    u128 rinv = f_inv(*root);
    zo_ntt(&rinv, in, n_in, out);
    u128 ninv = to_mont(f_inv((u128)n));
    #pragma omp parallel for schedule(static) if (n >= 65536)
    for (size_t i = 0; i < n; i++) out[i] = mont_mul(out[i], ninv);
    return 0;
}
/* polynomial.rs:109-121 */
void zo_scale(const u128 *factor, const u128 *in, size_t n, u128 *out) {
    init_consts();
    u128 fm = to_mont(*factor), t = to_mont(1);
    for (size_t i = 0; i < n; i++) { out[i] = mont_mul(in[i], t); t = mont_mul(t, fm); }
}
/* ntt_arithmetics.rs:161-170 */
int zo_coset_lde(const u128 *omega, size_t order, const u128 *offset, const u128 *coeffs, size_t n, u128 *out) {
    if (n > order) return -2;
    u128 *tmp = (u128 *)calloc(order, sizeof(u128));
    zo_scale(offset, coeffs, n, tmp);
    int rc = zo_ntt(omega, tmp, order, out);
    free(tmp);
    return rc;
}
/* merkle_root.rs:7-32.  If nodes != NULL it receives every level bottom-up:
 * n leaf hashes, n/2, ..., 1 (2n-1 digests of 64 bytes). */
int zo_merkle(const u128 *vals, size_t n, uint8_t root[64], uint8_t *nodes) {
    if (n == 0 || (n & (n - 1))) return -3;
    uint8_t *buf = nodes ? nodes : (uint8_t *)malloc((2 * n - 1) * 64);
    #pragma omp parallel for schedule(static) if (n >= 4096)
    for (size_t i = 0; i < n; i++) leaf_hash(vals[i], buf + 64 * i);
    uint8_t *cur = buf;
    for (size_t w = n; w > 1; w >>= 1) {
        uint8_t *nxt = cur + 64 * w;
        #pragma omp parallel for schedule(static) if (w >= 4096)
        for (size_t i = 0; i < w / 2; i++) blake2b512_1block(cur + 128 * i, 128, nxt + 64 * i);
        cur = nxt;
    }
    memcpy(root, cur, 64);
    if (!nodes) free(buf);
    return 0;
}
/* fri.rs:150-159: out[i] = 2^-1((1 + a/x_i) c[i] + (1 - a/x_i) c[i+n/2]), x_i = offset*omega^i */
void zo_fri_fold(const u128 *cw, size_t n, const u128 *alpha, const u128 *offset, const u128 *omega, u128 *out) {
    init_consts();
    size_t half = n / 2;
    u128 winv = to_mont(f_inv(*omega));
    u128 k0 = f_mul(*alpha, f_inv(*offset));            /* alpha / offset */
    u128 two_inv = to_mont(f_inv(2));
    /* walk alpha/x_i = k0 * omega^-i in chunks so the loop parallelises */
    const size_t CH = 4096;
    #pragma omp parallel for schedule(static) if (half >= 65536)
    for (size_t c0 = 0; c0 < half; c0 += CH) {
        u128 ax = f_mul(k0, f_pow(f_inv(*omega), (u128)c0));
        size_t end = c0 + CH < half ? c0 + CH : half;
        for (size_t i = c0; i < end; i++) {
            u128 first = f_mul(f_add(1, ax), cw[i]);
            u128 second = f_mul(f_sub(1, ax), cw[half + i]);
            out[i] = mont_mul(f_add(first, second), two_inv);
            ax = mont_mul(ax, winv);
        }
    }
}

/* ======================================================================= zr_*
 * Faithful-algorithm port: same operation counts as the Rust reference. */
static inline u128 r_sub_mod(u128 a, u128 b) { return a > b ? a - b : (a == b ? 0 : P - b + a); }
static inline u128 r_add_mod(u128 a, u128 b) { return b == 0 ? a : r_sub_mod(a, P - b); }
static u128 r_mul_mod(u128 a, u128 b) {            /* field.rs:117-131 */
    u128 res = 0;
    while (b > 0) {
        if (b & 1) res = r_add_mod(res, a);
        a = r_add_mod(a, a);
        b >>= 1;
    }
    return res;
}
static u128 r_pow(u128 base, u128 e) {             /* field_element.rs:108-143 */
    u128 acc = 1;
    int top = 0;
    for (int i = 127; i >= 0; i--) if ((e >> i) & 1) { top = i; break; }
    for (int i = top; i >= 0; i--) {
        acc = r_mul_mod(acc, acc);
        if ((e >> i) & 1) acc = r_mul_mod(acc, base);
    }
    return acc;
}
static u128 r_inv(u128 a) {
    /* field.rs:160-169 via utils/xgcd.rs:22-48: extended Euclid with u128 divisions
     * (the reference tracks both Bezout coefficients, this port only the one it uses -
     * slightly cheaper, which favours the CPU baseline),
     * tracking the Bezout coefficient of `a` with an explicit sign. */
    u128 r0 = a, r1 = P, s0 = 1, s1 = 0;
    int n0 = 0, n1 = 0;                         /* signs of s0, s1 */
    while (r1 != 0) {
        u128 q = r0 / r1, t = r0 - q * r1;
        r0 = r1; r1 = t;
        /* s = s0 - q*s1 with signs */
        u128 qs = q * s1, ns; int nn;
        if (n0 == n1) { if (s0 >= qs) { ns = s0 - qs; nn = n0; } else { ns = qs - s0; nn = !n0; } }
        else { ns = s0 + qs; nn = n0; }
        s0 = s1; n0 = n1; s1 = ns; n1 = nn;
    }
    if (s0 == 0) return 0;
    return n0 ? r_sub_mod(P, s0) : s0;
}
void zr_mul(const u128 *a, const u128 *b, u128 *out) { *out = r_mul_mod(*a, *b); }
void zr_inv(const u128 *a, u128 *out) { *out = r_inv(*a); }
void zr_pow(const u128 *a, const u128 *e, u128 *out) { *out = r_pow(*a, *e); }

int zr_ntt(const u128 *root, const u128 *in, size_t n_in, u128 *out) {   /* ntt.rs:7-49 */
    if (n_in == 0) return -1;
    if (n_in < 2) { out[0] = in[0]; return 0; }
    size_t n = next_pow2(n_in);
    int bits = 0; while (((size_t)1 << bits) < n) bits++;
    for (size_t k = 0; k < n; k++) {                        /* per-element bit loop */
        size_t r = 0;
        for (int b = 0; b < bits; b++) r |= ((k >> b) & 1) << (bits - 1 - b);
        out[r] = k < n_in ? in[k] : 0;
    }
    u128 *pw = (u128 *)malloc(sizeof(u128) * (n / 2));
    u128 t = 1;
    for (size_t k = 0; k < n / 2; k++) { pw[k] = t; t = r_mul_mod(t, *root); }
    for (size_t size = 2; size <= n; size <<= 1) {
        size_t half = size >> 1, step = n / size;
        for (size_t i = 0; i < n; i += size) {
            size_t k = 0;
            for (size_t j = i; j < i + half; j++) {
                u128 e = out[j], o = r_mul_mod(out[j + half], pw[k]);
                out[j] = r_add_mod(e, o);
                out[j + half] = r_sub_mod(e, o);
                k += step;
            }
        }
    }
    free(pw);
    return 0;
}
int zr_coset_lde(const u128 *omega, size_t order, const u128 *offset, const u128 *coeffs, size_t n, u128 *out) {
    if (n > order) return -2;
    u128 *tmp = (u128 *)calloc(order, sizeof(u128));
    for (size_t i = 0; i < n; i++) tmp[i] = r_mul_mod(r_pow(*offset, (u128)i), coeffs[i]);   /* polynomial.rs:116-117 */
    int rc = zr_ntt(omega, tmp, order, out);
    free(tmp);
    return rc;
}
typedef struct { uint8_t *p; size_t len; } rbytes;           /* utils/bytes.rs: a heap Vec<u8> */
static rbytes rb_hash(const uint8_t *msg, size_t len) {
    rbytes r; r.p = (uint8_t *)malloc(64); r.len = 64;
    blake2b512_1block(msg, len, r.p);
    return r;
}
static rbytes r_commit_(rbytes *leafs, size_t len) {        /* merkle_root.rs:7-19 */
    if (len == 1) { rbytes c; c.p = (uint8_t *)malloc(64); c.len = 64; memcpy(c.p, leafs[0].p, 64); return c; }
    rbytes a = r_commit_(leafs, len / 2), b = r_commit_(leafs + len / 2, len / 2);
    a.p = (uint8_t *)realloc(a.p, 128); memcpy(a.p + 64, b.p, 64); free(b.p);   /* concat = a + b */
    rbytes h = rb_hash(a.p, 128); free(a.p);
    return h;
}
int zr_merkle_commit(const u128 *vals, size_t n, uint8_t root[64]) {     /* merkle_root.rs:21-32 */
    if (n == 0 || (n & (n - 1))) return -3;
    rbytes *leafs = (rbytes *)malloc(sizeof(rbytes) * n);
    for (size_t i = 0; i < n; i++) {
        char *s = (char *)malloc(40);                        /* to_string() + Bytes */
        size_t l = u128_to_dec(vals[i], s);
        leafs[i] = rb_hash((const uint8_t *)s, l);
        free(s);
    }
    rbytes r = r_commit_(leafs, n);
    memcpy(root, r.p, 64); free(r.p);
    for (size_t i = 0; i < n; i++) free(leafs[i].p);
    free(leafs);
    return 0;
}
void zr_fri_fold(const u128 *cw, size_t n, const u128 *alpha, const u128 *offset, const u128 *omega, u128 *out) {
    size_t half = n / 2;                                     /* fri.rs:150-159 */
    u128 two_inv = r_inv(2);
    for (size_t i = 0; i < half; i++) {
        u128 x = r_mul_mod(*offset, r_pow(*omega, (u128)i));
        u128 ax = r_mul_mod(*alpha, r_inv(x));
        u128 first = r_mul_mod(r_add_mod(1, ax), cw[i]);
        u128 second = r_mul_mod(r_sub_mod(1, ax), cw[half + i]);
        out[i] = r_mul_mod(two_inv, r_add_mod(first, second));
    }
}
