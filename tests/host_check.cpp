// Host-side check of the __host__ __device__ arithmetic in csrc/fe128.cuh and
// csrc/blake2b.cuh against the C oracle (oracle/libzkoracle.so).  Test infrastructure:
// built and run by tests/test_host_arith.py on the CPU (no GPU needed).  The device
// PTX paths of the same functions are covered by the -m gpu parity tests.
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <cstdlib>
#include "../zk_stark_tutor_b200/csrc/blake2b.cuh"

typedef unsigned __int128 u128;
extern "C" {
void zo_mul(const u128*, const u128*, u128*);
void zo_blake2b512(const uint8_t*, size_t, uint8_t*);
int zo_merkle(const u128*, size_t, uint8_t*, uint8_t*);
}
using namespace zkb;

static const u128 P = ((u128)0xCB80000000000000ULL << 64) | 1ULL;
static uint64_t rng_state = 0x123456789ULL;
static uint64_t rnd64() { rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17; return rng_state; }
static u128 rndfe(int mode) {
    switch (mode % 8) {
        case 0: return 0;
        case 1: return 1;
        case 2: return P - 1;
        case 3: return (u128)rnd64();
        case 4: return (((u128)rnd64() << 64) | rnd64()) % P;
        case 5: return ((u128)1 << 96) - 1;
        case 6: return ((u128)(rnd64() & 0xFFFFFFFF) << 96) % P;
        default: return (((u128)rnd64() << 64) | rnd64()) % P;
    }
}
static fe tofe(u128 x) { fe r; for (int i = 0; i < 4; i++) r.v[i] = (uint32_t)(x >> (32 * i)); return r; }
static u128 fromfe(fe a) { u128 x = 0; for (int i = 3; i >= 0; i--) x = (x << 32) | a.v[i]; return x; }
static int fails = 0;
#define CHECK(c, ...) do { if (!(c)) { if (fails++ < 10) { printf("FAIL %s:%d ", __FILE__, __LINE__); printf(__VA_ARGS__); printf("\n"); } } } while (0)

int main() {
    const u128 Rm = (((u128)0x347FFFFFFFFFFFFFULL) << 64) | 0xFFFFFFFFFFFFFFFFULL;
    for (int it = 0; it < 200000; it++) {
        u128 a = rndfe(it), b = rndfe(it / 8);
        u128 s = a + b; if (s < a || s >= P) s -= P;
        CHECK(fromfe(fe_add(tofe(a), tofe(b))) == s, "add it=%d", it);
        u128 d = a >= b ? a - b : a + (P - b);
        CHECK(fromfe(fe_sub(tofe(a), tofe(b))) == d, "sub it=%d", it);
        u128 want; zo_mul(&a, &b, &want);
        CHECK(fromfe(fe_mul(tofe(a), tofe(b))) == want, "mul it=%d", it);
        // montmul(a, b*R) == a*b
        u128 bR; zo_mul(&b, &Rm, &bR);
        CHECK(fromfe(fe_montmul(tofe(a), tofe(bR))) == want, "montmul it=%d", it);
        CHECK(fromfe(fe_from_mont(fe_to_mont(tofe(a)))) == a, "mont roundtrip it=%d", it);
        u128 h = fromfe(fe_half(tofe(a)));
        u128 h2 = h + h; if (h2 < h || h2 >= P) h2 -= P;
        CHECK(h2 == a && h < P, "half it=%d", it);
    }
    // montmul with a non-canonical second operand < 2^128 (allowed: a*b < p*2^128)
    for (int it = 0; it < 20000; it++) {
        u128 a = rndfe(it), b = ((u128)rnd64() << 64) | rnd64();
        u128 br = b % P, want, t;
        zo_mul(&a, &br, &t);
        u128 one = 1; zo_mul(&t, &one, &want);
        fe r = fe_montmul(tofe(a), tofe(b));           // = a*b/R
        u128 back; u128 rr = fromfe(r); zo_mul(&rr, &Rm, &back);
        CHECK(back == want && rr < P, "montmul noncanon it=%d", it);
    }
    // pow
    {
        fe g = fe_to_mont(tofe(((u128)0x4040FBED12EE470FULL << 64) | 0xB5038F9C18F6F7D1ULL));
        fe x = g;
        for (int i = 0; i < 119; i++) x = fe_montmul(x, x);
        CHECK(fromfe(fe_from_mont(x)) == 1, "G^(2^119) != 1");
        CHECK(fe_eq(fe_mont_pow(g, 12345), fe_montmul(fe_mont_pow(g, 12344), g)), "pow");
    }
    // leaf encoder, piece by piece: the 4-digit groups (all of them), the split of every 8-digit chunk, the division steps (edges + random)
    for (uint32_t x = 0; x < 10000; x++) {
        uint32_t wv = dec4_ascii(x); char b4[8]; snprintf(b4, 8, "%04u", x);
        CHECK(memcmp(&wv, b4, 4) == 0, "dec4_ascii %u", x);
    }
    for (uint32_t cch = 0; cch < 100000000u; cch++) {
        uint32_t hi = b2_mulhi32(cch, 0xD1B71759u) >> 13;
        if (hi != cch / 10000u) { CHECK(false, "chunk / 10^4 at %u", cch); break; }
    }
    for (int it = 0; it < 20000000; it++) {
        uint32_t r = (uint32_t)(rnd64() % 100000000u), l = (uint32_t)rnd64();
        if (it < 64) { r = (it & 1) ? 99999999u : 0u; l = (it & 2) ? 0xFFFFFFFFu : ((it & 4) ? 0u : l); if (it & 8) r = (uint32_t)(rnd64() % 100000000u); if (it & 16) l = 0xFFFFFFFFu - (uint32_t)(it >> 5); }
        u128 x = ((u128)r << 32) | l;
        uint32_t rem, q = div1e8_step(r, l, rem);
        if (q != (uint32_t)(x / 100000000u) || rem != (uint32_t)(x % 100000000u)) { CHECK(false, "div1e8_step r=%u l=%u", r, l); break; }
    }
    // leaf encoder + hashes
    u128 edge[] = {0, 1, 9, 10, 11, 99, 100, 5462, 999999999, 1000000000, 1000000001,
                   (u128)1000000000 * 1000000000, (u128)1000000000 * 1000000000 - 1,
                   ((u128)1 << 64), ((u128)1 << 64) - 1, P - 1, P - 2, ((u128)1 << 127),
                   (u128)10000000000000000000ULL * 10000000000000000000ULL,            // 10^38
                   (u128)10000000000000000000ULL * 10000000000000000000ULL - 1,        // 10^38-1
                   (u128)10000000000000000000ULL * 1000000000000000000ULL};            // 10^37
    int nedge = sizeof(edge) / sizeof(edge[0]);
    for (int it = 0; it < 100000; it++) {
        u128 a;
        if (it < nedge) a = edge[it];
        else if (it % 3 == 0) { a = rndfe(7); int sh = (int)(rnd64() % 128); a >>= sh; }   // all lengths
        else a = rndfe(it);
        uint32_t w[10];
        uint32_t len = u128_to_dec_words(tofe(a), w);
        char buf[48]; int n = 0; { char tmp[48]; u128 v = a; if (!v) tmp[n++] = '0'; while (v) { tmp[n++] = '0' + (int)(v % 10); v /= 10; } for (int i = 0; i < n; i++) buf[i] = tmp[n - 1 - i]; }
        uint8_t exp[40]; memset(exp, 0, 40); memcpy(exp, buf, n);
        CHECK((int)len == n && memcmp(exp, w, 40) == 0, "dec it=%d len=%u n=%d", it, len, n);
        uint64_t h[8]; uint8_t want[64];
        blake2b_leaf(tofe(a), h);
        zo_blake2b512((const uint8_t*)buf, n, want);
        CHECK(memcmp(h, want, 64) == 0, "leaf hash it=%d", it);
        if (it < 2000) {
            uint64_t l[8], r[8], hn[8]; uint8_t msg[128];
            for (int i = 0; i < 8; i++) { l[i] = rnd64(); r[i] = rnd64(); }
            memcpy(msg, l, 64); memcpy(msg + 64, r, 64);
            blake2b_node(l, r, hn);
            zo_blake2b512(msg, 128, want);
            CHECK(memcmp(hn, want, 64) == 0, "node hash it=%d", it);
        }
    }
    if (fails) { printf("host_check: %d failures\n", fails); return 1; }
    printf("host_check: ok\n");
    return 0;
}
