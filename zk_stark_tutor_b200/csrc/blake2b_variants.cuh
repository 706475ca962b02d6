// blake2b_variants.cuh - EXPERIMENTAL instruction-mix variants of the BLAKE2b compression, used only by the
// integer-pipe probe (probe.cu -> lib/libzkb200_probe.so, tools/probe_run.py): the study behind DESIGN.md 4
// ("25 variants measured, none beats the plain ALU-pipe form").  Not part of libzkb200.so.
#pragma once
#include "blake2b.cuh"

namespace zkb {

#if defined(__CUDACC__)
// ---- device variants of the compression with explicit pipe placement -----------------------
// On B200 the ALU pipe issues LOP3 / plain IADD3 at 32 lanes/clk/SMSP but PRMT, SHF and the
// carry-in IADD3.X at 16 (measured, tools/probe_run.py), while the FMA pipe (IMAD*) is idle in
// a pure-ALU BLAKE2b.  The variants move work across: V bit 0: high-word adds as IMAD.X
// (madc.lo), bit 1: rotr63 as 2 IMAD.WIDE + 2 LOP3 instead of 2 SHF, bit 2: rotr24 likewise
// instead of 2 PRMT, bit 3: rotr16 likewise, bit 4: 3-input add also with explicit madc.
__device__ __forceinline__ uint64_t b2v_pack(uint32_t lo, uint32_t hi) { return ((uint64_t)hi << 32) | lo; }
template <int V> __device__ __forceinline__ uint64_t b2v_add(uint64_t a, uint64_t b) {
    if (V & 1) {
        uint32_t lo, hi;
        asm("add.cc.u32 %0, %2, %4;\n\tmadc.lo.u32 %1, %3, 1, %5;"
            : "=r"(lo), "=r"(hi) : "r"((uint32_t)a), "r"((uint32_t)(a >> 32)), "r"((uint32_t)b), "r"((uint32_t)(b >> 32)));
        return b2v_pack(lo, hi);
    }
    return a + b;
}
template <int V> __device__ __forceinline__ uint64_t b2v_add3(uint64_t a, uint64_t b, uint64_t m) {
    if (V & 16) return b2v_add<V>(b2v_add<V>(a, b), m);
    return a + b + m;
}
// rotate left by K < 32 through the FMA pipe: t = lo*2^K, u = hi*2^K (64-bit products)
template <int K> __device__ __forceinline__ uint64_t b2v_rotl_mad(uint32_t lo, uint32_t hi) {
    uint64_t t, u;
    asm("mad.wide.u32 %0, %1, %2, 0;" : "=l"(t) : "r"(lo), "r"(1u << K));
    asm("mad.wide.u32 %0, %1, %2, 0;" : "=l"(u) : "r"(hi), "r"(1u << K));
    return b2v_pack((uint32_t)t | (uint32_t)(u >> 32), (uint32_t)u | (uint32_t)(t >> 32));
}
template <int V> __device__ __forceinline__ uint64_t b2v_rotr63(uint64_t x) {
    if (V & 2) return b2v_rotl_mad<1>((uint32_t)x, (uint32_t)(x >> 32));
    return b2_rotr63(x);
}
template <int V> __device__ __forceinline__ uint64_t b2v_rotr24(uint64_t x) {
    if (V & 4) return b2v_rotl_mad<8>((uint32_t)(x >> 32), (uint32_t)x);       // rotl40 = word swap + rotl8
    return b2_rotr24(x);
}
template <int V> __device__ __forceinline__ uint64_t b2v_rotr16(uint64_t x) {
    if (V & 8) return b2v_rotl_mad<16>((uint32_t)(x >> 32), (uint32_t)x);      // rotl48 = word swap + rotl16
    return b2_rotr16(x);
}
#define ZKB_B2V_G(a, b, c, d, x, y)                                   \
    do {                                                              \
        a = b2v_add3<V>(a, b, (x)); d = b2_rotr32(d ^ a);             \
        c = b2v_add<V>(c, d);       b = b2v_rotr24<V>(b ^ c);         \
        a = b2v_add3<V>(a, b, (y)); d = b2v_rotr16<V>(d ^ a);         \
        c = b2v_add<V>(c, d);       b = b2v_rotr63<V>(b ^ c);         \
    } while (0)
#define ZKB_B2V_ROUND(s0, s1, s2, s3, s4, s5, s6, s7, s8, s9, s10, s11, s12, s13, s14, s15) \
    do {                                                                                  \
        ZKB_B2V_G(v0, v4, v8, v12, m[s0], m[s1]);   ZKB_B2V_G(v1, v5, v9, v13, m[s2], m[s3]);   \
        ZKB_B2V_G(v2, v6, v10, v14, m[s4], m[s5]);  ZKB_B2V_G(v3, v7, v11, v15, m[s6], m[s7]);  \
        ZKB_B2V_G(v0, v5, v10, v15, m[s8], m[s9]);  ZKB_B2V_G(v1, v6, v11, v12, m[s10], m[s11]); \
        ZKB_B2V_G(v2, v7, v8, v13, m[s12], m[s13]); ZKB_B2V_G(v3, v4, v9, v14, m[s14], m[s15]); \
    } while (0)
template <int V>
__device__ __forceinline__ void blake2b_compress_dev(const uint64_t (&m)[16], uint64_t t, uint64_t (&h)[8]) {
    uint64_t v0 = ZKB_B2_H0, v1 = ZKB_B2_IV1, v2 = ZKB_B2_IV2, v3 = ZKB_B2_IV3;
    uint64_t v4 = ZKB_B2_IV4, v5 = ZKB_B2_IV5, v6 = ZKB_B2_IV6, v7 = ZKB_B2_IV7;
    uint64_t v8 = ZKB_B2_IV0, v9 = ZKB_B2_IV1, v10 = ZKB_B2_IV2, v11 = ZKB_B2_IV3;
    uint64_t v12 = ZKB_B2_IV4 ^ t, v13 = ZKB_B2_IV5, v14 = ~ZKB_B2_IV6, v15 = ZKB_B2_IV7;
    ZKB_B2V_ROUND(0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15);
    ZKB_B2V_ROUND(14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3);
    ZKB_B2V_ROUND(11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4);
    ZKB_B2V_ROUND(7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8);
    ZKB_B2V_ROUND(9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13);
    ZKB_B2V_ROUND(2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9);
    ZKB_B2V_ROUND(12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11);
    ZKB_B2V_ROUND(13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10);
    ZKB_B2V_ROUND(6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5);
    ZKB_B2V_ROUND(10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0);
    ZKB_B2V_ROUND(0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15);
    ZKB_B2V_ROUND(14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3);
    h[0] = ZKB_B2_H0 ^ v0 ^ v8;   h[1] = ZKB_B2_IV1 ^ v1 ^ v9;
    h[2] = ZKB_B2_IV2 ^ v2 ^ v10; h[3] = ZKB_B2_IV3 ^ v3 ^ v11;
    h[4] = ZKB_B2_IV4 ^ v4 ^ v12; h[5] = ZKB_B2_IV5 ^ v5 ^ v13;
    h[6] = ZKB_B2_IV6 ^ v6 ^ v14; h[7] = ZKB_B2_IV7 ^ v7 ^ v15;
}
#endif

#if defined(__CUDACC__)
// ---- compression on separate 32-bit halves ---------------------------------------------------
// 64-bit C variables live in aligned (even, odd) register pairs, so every low-word LOP3 / IADD3
// reads only even registers and every high-word one only odd registers: the register file
// serves one distinct register per bank per cycle (B300_MICROARCH.md "RF banking"), which caps
// the portable code at ~0.57 IPC.  Keeping the halves in independent 32-bit variables lets
// ptxas spread the operands of one instruction over both banks.
struct b2w { uint32_t lo, hi; };
__device__ __forceinline__ void b2h_add3(b2w& a, const b2w& b, uint32_t mlo, uint32_t mhi) {
    asm("add.cc.u32 %0, %0, %2;\n\taddc.u32 %1, %1, %3;\n\tadd.cc.u32 %0, %0, %4;\n\taddc.u32 %1, %1, %5;"
        : "+r"(a.lo), "+r"(a.hi) : "r"(b.lo), "r"(b.hi), "r"(mlo), "r"(mhi));
}
__device__ __forceinline__ void b2h_add(b2w& c, const b2w& d) {
    asm("add.cc.u32 %0, %0, %2;\n\taddc.u32 %1, %1, %3;" : "+r"(c.lo), "+r"(c.hi) : "r"(d.lo), "r"(d.hi));
}
// d = rotr32(d ^ a): swap halves
__device__ __forceinline__ void b2h_xr32(b2w& d, const b2w& a) { uint32_t l = d.hi ^ a.hi, h = d.lo ^ a.lo; d.lo = l; d.hi = h; }
__device__ __forceinline__ void b2h_xr24(b2w& b, const b2w& c) {
    uint32_t l = b.lo ^ c.lo, h = b.hi ^ c.hi;
    b.lo = __byte_perm(l, h, 0x6543); b.hi = __byte_perm(h, l, 0x6543);
}
__device__ __forceinline__ void b2h_xr16(b2w& d, const b2w& a) {
    uint32_t l = d.lo ^ a.lo, h = d.hi ^ a.hi;
    d.lo = __byte_perm(l, h, 0x5432); d.hi = __byte_perm(h, l, 0x5432);
}
__device__ __forceinline__ void b2h_xr63(b2w& b, const b2w& c) {
    uint32_t l = b.lo ^ c.lo, h = b.hi ^ c.hi;
    b.lo = __funnelshift_l(h, l, 1); b.hi = __funnelshift_l(l, h, 1);
}
#define ZKB_B2H_G(a, b, c, d, x, y)                                        \
    do {                                                                   \
        b2h_add3(v[a], v[b], ml[x], mh[x]); b2h_xr32(v[d], v[a]);          \
        b2h_add(v[c], v[d]);                b2h_xr24(v[b], v[c]);          \
        b2h_add3(v[a], v[b], ml[y], mh[y]); b2h_xr16(v[d], v[a]);          \
        b2h_add(v[c], v[d]);                b2h_xr63(v[b], v[c]);          \
    } while (0)
#define ZKB_B2H_ROUND(s0, s1, s2, s3, s4, s5, s6, s7, s8, s9, s10, s11, s12, s13, s14, s15) \
    do {                                                                                  \
        ZKB_B2H_G(0, 4, 8, 12, s0, s1);   ZKB_B2H_G(1, 5, 9, 13, s2, s3);   \
        ZKB_B2H_G(2, 6, 10, 14, s4, s5);  ZKB_B2H_G(3, 7, 11, 15, s6, s7);  \
        ZKB_B2H_G(0, 5, 10, 15, s8, s9);  ZKB_B2H_G(1, 6, 11, 12, s10, s11); \
        ZKB_B2H_G(2, 7, 8, 13, s12, s13); ZKB_B2H_G(3, 4, 9, 14, s14, s15); \
    } while (0)
// ml/mh: message words (low / high halves); t < 2^32; hl/hh: digest halves out
__device__ __forceinline__ void blake2b_compress_h32(const uint32_t (&ml)[16], const uint32_t (&mh)[16], uint32_t t,
                                                     uint32_t (&hl)[8], uint32_t (&hh)[8]) {
    const uint64_t iv[8] = {ZKB_B2_IV0, ZKB_B2_IV1, ZKB_B2_IV2, ZKB_B2_IV3, ZKB_B2_IV4, ZKB_B2_IV5, ZKB_B2_IV6, ZKB_B2_IV7};
    b2w v[16];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        uint64_t h0 = i == 0 ? ZKB_B2_H0 : iv[i];
        v[i].lo = (uint32_t)h0; v[i].hi = (uint32_t)(h0 >> 32);
        uint64_t w = i == 6 ? ~iv[6] : iv[i];
        v[8 + i].lo = (uint32_t)w; v[8 + i].hi = (uint32_t)(w >> 32);
    }
    v[12].lo ^= t;
    ZKB_B2H_ROUND(0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15);
    ZKB_B2H_ROUND(14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3);
    ZKB_B2H_ROUND(11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4);
    ZKB_B2H_ROUND(7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8);
    ZKB_B2H_ROUND(9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13);
    ZKB_B2H_ROUND(2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9);
    ZKB_B2H_ROUND(12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11);
    ZKB_B2H_ROUND(13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10);
    ZKB_B2H_ROUND(6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5);
    ZKB_B2H_ROUND(10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0);
    ZKB_B2H_ROUND(0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15);
    ZKB_B2H_ROUND(14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3);
#pragma unroll
    for (int i = 0; i < 8; i++) {
        uint64_t h0 = i == 0 ? ZKB_B2_H0 : iv[i];
        hl[i] = (uint32_t)h0 ^ v[i].lo ^ v[8 + i].lo;
        hh[i] = (uint32_t)(h0 >> 32) ^ v[i].hi ^ v[8 + i].hi;
    }
}
#endif

#if defined(__CUDACC__)
// ---- pipe-balanced compression (experimental family, selected by CFG bits) --------------------
// B200 issues 1 instruction/clk/SMSP but each of the ALU pipe (LOP3, IADD3, PRMT, SHF) and the
// FMA pipe (IMAD*) accepts one warp instruction every 2 clk, so an all-ALU BLAKE2b runs at half
// the issue rate (ncu: pipe_alu 86 %, pipe_fma 8 %).  These variants re-express adds and rotates
// as IMADs so both pipes work.  K = {1, 2^1, 2^8, 2^16} must come from kernel parameters /
// registers (a literal would be strength-reduced back to shifts).
//   CFG bit 0: 2-input add  = IMAD.WIDE(a.lo, 1, b) ; hi += a.hi
//   CFG bit 1: 3-input add  = two IMAD.WIDE ; hi = t.hi + a.hi + m.hi
//   CFG bit 2: rotr63 via IMAD.HI + IMAD     (4 FMA-pipe instructions, 0 ALU)
//   CFG bit 3: rotr24 likewise,  CFG bit 4: rotr16 likewise
//   CFG bit 5: hi-word adds of bit 0/1 as IMAD (mad.lo x, 1, y) instead of add
__device__ __forceinline__ uint64_t b2x_pack(uint32_t lo, uint32_t hi) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi)); return r; }
__device__ __forceinline__ void b2x_unpack(uint64_t x, uint32_t& lo, uint32_t& hi) { asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(x)); }
template <int CFG> __device__ __forceinline__ uint32_t b2x_hiadd(uint32_t x, uint32_t y, uint32_t k1) {
    if (CFG & 32) { uint32_t r; asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(x), "r"(k1), "r"(y)); return r; }
    return x + y;
}
template <int CFG> __device__ __forceinline__ void b2x_add(b2w& c, const b2w& d, const uint32_t (&K)[4]) {
    if (CFG & 1) {
        uint64_t t;
        asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(t) : "r"(d.lo), "r"(K[0]), "l"(b2x_pack(c.lo, c.hi)));
        uint32_t lo, hi;
        b2x_unpack(t, lo, hi);
        c.lo = lo; c.hi = b2x_hiadd<CFG>(d.hi, hi, K[0]);
    } else {
        b2h_add(c, d);
    }
}
template <int CFG> __device__ __forceinline__ void b2x_add3(b2w& a, const b2w& b, uint32_t mlo, uint32_t mhi, const uint32_t (&K)[4]) {
    if (CFG & 2) {
        uint64_t t;
        asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(t) : "r"(b.lo), "r"(K[0]), "l"(b2x_pack(a.lo, a.hi)));
        asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(t) : "r"(mlo), "r"(K[0]));
        uint32_t lo, hi;
        b2x_unpack(t, lo, hi);
        a.lo = lo;
        if (CFG & 32) a.hi = b2x_hiadd<CFG>(b.hi, b2x_hiadd<CFG>(mhi, hi, K[0]), K[0]);
        else a.hi = hi + b.hi + mhi;
    } else {
        uint64_t r = b2x_pack(a.lo, a.hi) + b2x_pack(b.lo, b.hi) + b2x_pack(mlo, mhi);
        a.lo = (uint32_t)r; a.hi = (uint32_t)(r >> 32);
    }
}
// rotate left by log2(k) < 32 of the 64-bit word {hi, lo} on the FMA pipe
__device__ __forceinline__ void b2x_rotl_fma(uint32_t lo, uint32_t hi, uint32_t k, uint32_t& olo, uint32_t& ohi) {
    uint32_t tl, th;
    asm("mul.hi.u32 %0, %1, %2;" : "=r"(tl) : "r"(hi), "r"(k));      // hi >> (32 - s)
    asm("mul.hi.u32 %0, %1, %2;" : "=r"(th) : "r"(lo), "r"(k));      // lo >> (32 - s)
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(olo) : "r"(lo), "r"(k), "r"(tl));
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(ohi) : "r"(hi), "r"(k), "r"(th));
}
template <int CFG> __device__ __forceinline__ void b2x_xr24(b2w& b, const b2w& c, const uint32_t (&K)[4]) {
    uint32_t l = b.lo ^ c.lo, h = b.hi ^ c.hi;
    if (CFG & 8) b2x_rotl_fma(h, l, K[2], b.lo, b.hi);                // rotl40 = word swap + rotl8
    else { b.lo = __byte_perm(l, h, 0x6543); b.hi = __byte_perm(h, l, 0x6543); }
}
template <int CFG> __device__ __forceinline__ void b2x_xr16(b2w& d, const b2w& a, const uint32_t (&K)[4]) {
    uint32_t l = d.lo ^ a.lo, h = d.hi ^ a.hi;
    if (CFG & 16) b2x_rotl_fma(h, l, K[3], d.lo, d.hi);               // rotl48 = word swap + rotl16
    else { d.lo = __byte_perm(l, h, 0x5432); d.hi = __byte_perm(h, l, 0x5432); }
}
template <int CFG> __device__ __forceinline__ void b2x_xr63(b2w& b, const b2w& c, const uint32_t (&K)[4]) {
    uint32_t l = b.lo ^ c.lo, h = b.hi ^ c.hi;
    if (CFG & 4) b2x_rotl_fma(l, h, K[1], b.lo, b.hi);
    else { b.lo = __funnelshift_l(h, l, 1); b.hi = __funnelshift_l(l, h, 1); }
}
#define ZKB_B2X_G(a, b, c, d, x, y)                                                  \
    do {                                                                             \
        b2x_add3<CFG>(v[a], v[b], ml[x], mh[x], K); b2h_xr32(v[d], v[a]);            \
        b2x_add<CFG>(v[c], v[d], K);                b2x_xr24<CFG>(v[b], v[c], K);    \
        b2x_add3<CFG>(v[a], v[b], ml[y], mh[y], K); b2x_xr16<CFG>(v[d], v[a], K);    \
        b2x_add<CFG>(v[c], v[d], K);                b2x_xr63<CFG>(v[b], v[c], K);    \
    } while (0)
#define ZKB_B2X_ROUND(s0, s1, s2, s3, s4, s5, s6, s7, s8, s9, s10, s11, s12, s13, s14, s15) \
    do {                                                                                  \
        ZKB_B2X_G(0, 4, 8, 12, s0, s1);   ZKB_B2X_G(1, 5, 9, 13, s2, s3);   \
        ZKB_B2X_G(2, 6, 10, 14, s4, s5);  ZKB_B2X_G(3, 7, 11, 15, s6, s7);  \
        ZKB_B2X_G(0, 5, 10, 15, s8, s9);  ZKB_B2X_G(1, 6, 11, 12, s10, s11); \
        ZKB_B2X_G(2, 7, 8, 13, s12, s13); ZKB_B2X_G(3, 4, 9, 14, s14, s15); \
    } while (0)
template <int CFG>
__device__ __forceinline__ void blake2b_compress_x(const uint32_t (&ml)[16], const uint32_t (&mh)[16], uint32_t t,
                                                   const uint32_t (&K)[4], uint32_t (&hl)[8], uint32_t (&hh)[8]) {
    const uint64_t iv[8] = {ZKB_B2_IV0, ZKB_B2_IV1, ZKB_B2_IV2, ZKB_B2_IV3, ZKB_B2_IV4, ZKB_B2_IV5, ZKB_B2_IV6, ZKB_B2_IV7};
    b2w v[16];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        uint64_t h0 = i == 0 ? ZKB_B2_H0 : iv[i];
        v[i].lo = (uint32_t)h0; v[i].hi = (uint32_t)(h0 >> 32);
        uint64_t w = i == 6 ? ~iv[6] : iv[i];
        v[8 + i].lo = (uint32_t)w; v[8 + i].hi = (uint32_t)(w >> 32);
    }
    v[12].lo ^= t;
    ZKB_B2X_ROUND(0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15);
    ZKB_B2X_ROUND(14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3);
    ZKB_B2X_ROUND(11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4);
    ZKB_B2X_ROUND(7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8);
    ZKB_B2X_ROUND(9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13);
    ZKB_B2X_ROUND(2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9);
    ZKB_B2X_ROUND(12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11);
    ZKB_B2X_ROUND(13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10);
    ZKB_B2X_ROUND(6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5);
    ZKB_B2X_ROUND(10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0);
    ZKB_B2X_ROUND(0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15);
    ZKB_B2X_ROUND(14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3);
#pragma unroll
    for (int i = 0; i < 8; i++) {
        uint64_t h0 = i == 0 ? ZKB_B2_H0 : iv[i];
        hl[i] = (uint32_t)h0 ^ v[i].lo ^ v[8 + i].lo;
        hh[i] = (uint32_t)(h0 >> 32) ^ v[i].hi ^ v[8 + i].hi;
    }
}
#endif

#if defined(__CUDACC__)
// ---- round 2: 64-bit adds as ONE accumulating IMAD.WIDE (family "y") ---------------------------
// The x family above wrote c + d as mad.wide.u32 with a 64-bit addend, which ptxas 12.9 SPLITS into IMAD.WIDE + IADD3 + IADD3.X
// on sm_100a (found later with probe kind 40): it added FMA-pipe work without removing any ALU-pipe work.  The pair
// mad.lo.cc / madc.hi (as in fe_montmul) IS fused into one `IMAD.WIDE.U32 Rd, Ra, K1, Rc` with the 64-bit accumulate, so
// c + d = {IMAD.WIDE(d.lo * 1 + c), hi += d.hi}: the low-word IADD3 and the carry-in IADD3.X leave the ALU pipe.
//   CFG = c_frac + 5 * a_frac + 25 * hi_imad + 50 * a_mode
//   c_frac / a_frac in 0..4: how many of every four G functions use the IMAD.WIDE form for their c + d / a + b + m adds
//   hi_imad: the remaining high-word add as IMAD (mad.lo x, 1, y) instead of leaving the choice to ptxas
//   a_mode 0: a + b through IMAD.WIDE, + m as an ordinary 64-bit add; 1: + m through a second IMAD.WIDE, one 3-input high add
template <int HI> __device__ __forceinline__ uint32_t b2y_hiadd(uint32_t x, uint32_t y, uint32_t k1) {
    if (HI) { uint32_t r; asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(x), "r"(k1), "r"(y)); return r; }
    return x + y;
}
// {lo, hi} = x * k1 + {clo, chi}  (one IMAD.WIDE.U32 with accumulate)
__device__ __forceinline__ void b2y_wide(uint32_t x, uint32_t k1, uint32_t clo, uint32_t chi, uint32_t& lo, uint32_t& hi) {
    asm("mad.lo.cc.u32 %0, %2, %3, %4;\n\tmadc.hi.u32 %1, %2, %3, %5;" : "=r"(lo), "=r"(hi) : "r"(x), "r"(k1), "r"(clo), "r"(chi));
}
template <bool W, int HI> __device__ __forceinline__ void b2y_add(b2w& c, const b2w& d, uint32_t k1) {
    if (W) {
        uint32_t lo, hi;
        b2y_wide(d.lo, k1, c.lo, c.hi, lo, hi);
        c.lo = lo; c.hi = b2y_hiadd<HI>(d.hi, hi, k1);
    } else {
        b2h_add(c, d);
    }
}
template <bool W, int HI, int AMODE> __device__ __forceinline__ void b2y_add3(b2w& a, const b2w& b, uint32_t mlo, uint32_t mhi, uint32_t k1) {
    if (W) {
        uint32_t lo, hi;
        b2y_wide(b.lo, k1, a.lo, a.hi, lo, hi);
        if (AMODE == 0) {
            hi = b2y_hiadd<HI>(b.hi, hi, k1);
            asm("add.cc.u32 %0, %0, %2;\n\taddc.u32 %1, %1, %3;" : "+r"(lo), "+r"(hi) : "r"(mlo), "r"(mhi));
        } else {
            uint32_t lo2, hi2;
            b2y_wide(mlo, k1, lo, hi, lo2, hi2);
            lo = lo2; hi = hi2 + b.hi + mhi;
        }
        a.lo = lo; a.hi = hi;
    } else {
        uint64_t r = b2x_pack(a.lo, a.hi) + b2x_pack(b.lo, b.hi) + b2x_pack(mlo, mhi);
        a.lo = (uint32_t)r; a.hi = (uint32_t)(r >> 32);
    }
}
#define ZKB_B2Y_G(gi, a, b, c, d, x, y)                                                                   \
    do {                                                                                                  \
        b2y_add3<((gi) & 3) < AF, HI, AM>(v[a], v[b], ml[x], mh[x], k1); b2h_xr32(v[d], v[a]);            \
        b2y_add<((gi) & 3) < CF, HI>(v[c], v[d], k1);                    b2h_xr24(v[b], v[c]);            \
        b2y_add3<((gi) & 3) < AF, HI, AM>(v[a], v[b], ml[y], mh[y], k1); b2h_xr16(v[d], v[a]);            \
        b2y_add<((gi) & 3) < CF, HI>(v[c], v[d], k1);                    b2h_xr63(v[b], v[c]);            \
    } while (0)
#define ZKB_B2Y_ROUND(s0, s1, s2, s3, s4, s5, s6, s7, s8, s9, s10, s11, s12, s13, s14, s15) \
    do {                                                                                  \
        ZKB_B2Y_G(0, 0, 4, 8, 12, s0, s1);   ZKB_B2Y_G(1, 1, 5, 9, 13, s2, s3);   \
        ZKB_B2Y_G(2, 2, 6, 10, 14, s4, s5);  ZKB_B2Y_G(3, 3, 7, 11, 15, s6, s7);  \
        ZKB_B2Y_G(4, 0, 5, 10, 15, s8, s9);  ZKB_B2Y_G(5, 1, 6, 11, 12, s10, s11); \
        ZKB_B2Y_G(6, 2, 7, 8, 13, s12, s13); ZKB_B2Y_G(7, 3, 4, 9, 14, s14, s15); \
    } while (0)
template <int CFG>
__device__ __forceinline__ void blake2b_compress_y(const uint32_t (&ml)[16], const uint32_t (&mh)[16], uint32_t t,
                                                   uint32_t k1, uint32_t (&hl)[8], uint32_t (&hh)[8]) {
    constexpr int CF = CFG % 5, AF = (CFG / 5) % 5, HI = (CFG / 25) % 2, AM = (CFG / 50) % 2;
    const uint64_t iv[8] = {ZKB_B2_IV0, ZKB_B2_IV1, ZKB_B2_IV2, ZKB_B2_IV3, ZKB_B2_IV4, ZKB_B2_IV5, ZKB_B2_IV6, ZKB_B2_IV7};
    b2w v[16];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        uint64_t h0 = i == 0 ? ZKB_B2_H0 : iv[i];
        v[i].lo = (uint32_t)h0; v[i].hi = (uint32_t)(h0 >> 32);
        uint64_t w = i == 6 ? ~iv[6] : iv[i];
        v[8 + i].lo = (uint32_t)w; v[8 + i].hi = (uint32_t)(w >> 32);
    }
    v[12].lo ^= t;
    ZKB_B2Y_ROUND(0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15);
    ZKB_B2Y_ROUND(14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3);
    ZKB_B2Y_ROUND(11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4);
    ZKB_B2Y_ROUND(7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8);
    ZKB_B2Y_ROUND(9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13);
    ZKB_B2Y_ROUND(2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9);
    ZKB_B2Y_ROUND(12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11);
    ZKB_B2Y_ROUND(13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10);
    ZKB_B2Y_ROUND(6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5);
    ZKB_B2Y_ROUND(10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0);
    ZKB_B2Y_ROUND(0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15);
    ZKB_B2Y_ROUND(14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3);
#pragma unroll
    for (int i = 0; i < 8; i++) {
        uint64_t h0 = i == 0 ? ZKB_B2_H0 : iv[i];
        hl[i] = (uint32_t)h0 ^ v[i].lo ^ v[8 + i].lo;
        hh[i] = (uint32_t)(h0 >> 32) ^ v[i].hi ^ v[8 + i].hi;
    }
}
#endif

}  // namespace zkb
