// blake2b.cuh - single-block BLAKE2b-512 compression and the Merkle leaf encoder.
//
// Replaces src/crypto/blake2b512.rs:4-14 (crate blake2 0.10.6 `Blake2b512`, RFC 7693,
// unkeyed, 64-byte digest) for the two message shapes the Merkle tree needs
// (src/merkle_root.rs:7-32):
//   node  = BLAKE2b-512(left64 || right64)      one full 128-byte block, t = 128
//   leaf  = BLAKE2b-512(ascii_decimal(value))   1..39 bytes, one padded block, t = len
// The leaf preimage is the decimal STRING of the u128 value
// (src/field/field_element.rs:46-50,101-105), so each leaf needs a 128-bit
// binary -> decimal conversion on device (u128_to_dec_block below).
//
// Cost model: 96 G x (4 64-bit adds, 4 64-bit xors, 3 non-trivial rotates)
// = 2112 32-bit ALU-pipe ops + 16 feed-forward LOP3; rotr32 is a register rename.
#pragma once
#include <stdint.h>
#include "fe128.cuh"

namespace zkb {

#define ZKB_B2_IV0 0x6a09e667f3bcc908ULL
#define ZKB_B2_IV1 0xbb67ae8584caa73bULL
#define ZKB_B2_IV2 0x3c6ef372fe94f82bULL
#define ZKB_B2_IV3 0xa54ff53a5f1d36f1ULL
#define ZKB_B2_IV4 0x510e527fade682d1ULL
#define ZKB_B2_IV5 0x9b05688c2b3e6c1fULL
#define ZKB_B2_IV6 0x1f83d9abfb41bd6bULL
#define ZKB_B2_IV7 0x5be0cd19137e2179ULL
#define ZKB_B2_H0 (ZKB_B2_IV0 ^ 0x01010040ULL)   // digest 64, key 0, fanout 1, depth 1

ZKB_HD uint64_t b2_rotr32(uint64_t x) { return (x >> 32) | (x << 32); }
ZKB_HD uint64_t b2_rotr24(uint64_t x) {
#if defined(__CUDA_ARCH__)
    uint32_t lo = (uint32_t)x, hi = (uint32_t)(x >> 32);
    return ((uint64_t)__byte_perm(hi, lo, 0x6543) << 32) | __byte_perm(lo, hi, 0x6543);
#else
    return (x >> 24) | (x << 40);
#endif
}
ZKB_HD uint64_t b2_rotr16(uint64_t x) {
#if defined(__CUDA_ARCH__)
    uint32_t lo = (uint32_t)x, hi = (uint32_t)(x >> 32);
    return ((uint64_t)__byte_perm(hi, lo, 0x5432) << 32) | __byte_perm(lo, hi, 0x5432);
#else
    return (x >> 16) | (x << 48);
#endif
}
ZKB_HD uint64_t b2_rotr63(uint64_t x) {
#if defined(__CUDA_ARCH__)
    uint32_t lo = (uint32_t)x, hi = (uint32_t)(x >> 32);
    return ((uint64_t)__funnelshift_l(lo, hi, 1) << 32) | __funnelshift_l(hi, lo, 1);
#else
    return (x >> 63) | (x << 1);
#endif
}

#define ZKB_B2_G(a, b, c, d, x, y)          \
    do {                                    \
        a = a + b + (x); d = b2_rotr32(d ^ a); \
        c = c + d;       b = b2_rotr24(b ^ c); \
        a = a + b + (y); d = b2_rotr16(d ^ a); \
        c = c + d;       b = b2_rotr63(b ^ c); \
    } while (0)

#define ZKB_B2_ROUND(s0, s1, s2, s3, s4, s5, s6, s7, s8, s9, s10, s11, s12, s13, s14, s15) \
    do {                                                                                  \
        ZKB_B2_G(v0, v4, v8, v12, m[s0], m[s1]);   ZKB_B2_G(v1, v5, v9, v13, m[s2], m[s3]);   \
        ZKB_B2_G(v2, v6, v10, v14, m[s4], m[s5]);  ZKB_B2_G(v3, v7, v11, v15, m[s6], m[s7]);  \
        ZKB_B2_G(v0, v5, v10, v15, m[s8], m[s9]);  ZKB_B2_G(v1, v6, v11, v12, m[s10], m[s11]); \
        ZKB_B2_G(v2, v7, v8, v13, m[s12], m[s13]); ZKB_B2_G(v3, v4, v9, v14, m[s14], m[s15]); \
    } while (0)

// One-block BLAKE2b-512: m[16] message words (little-endian, zero padded), t = byte
// count.  Writes the 8 digest words to h[8].  Fully unrolled so every m[] index is a
// compile-time constant (words known to be zero at compile time cost no registers).
ZKB_HD void blake2b_compress_1block(const uint64_t (&m)[16], uint64_t t, uint64_t (&h)[8]) {
    uint64_t v0 = ZKB_B2_H0, v1 = ZKB_B2_IV1, v2 = ZKB_B2_IV2, v3 = ZKB_B2_IV3;
    uint64_t v4 = ZKB_B2_IV4, v5 = ZKB_B2_IV5, v6 = ZKB_B2_IV6, v7 = ZKB_B2_IV7;
    uint64_t v8 = ZKB_B2_IV0, v9 = ZKB_B2_IV1, v10 = ZKB_B2_IV2, v11 = ZKB_B2_IV3;
    uint64_t v12 = ZKB_B2_IV4 ^ t, v13 = ZKB_B2_IV5, v14 = ~ZKB_B2_IV6, v15 = ZKB_B2_IV7;
    ZKB_B2_ROUND(0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15);
    ZKB_B2_ROUND(14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3);
    ZKB_B2_ROUND(11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4);
    ZKB_B2_ROUND(7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8);
    ZKB_B2_ROUND(9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13);
    ZKB_B2_ROUND(2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9);
    ZKB_B2_ROUND(12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11);
    ZKB_B2_ROUND(13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10);
    ZKB_B2_ROUND(6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5);
    ZKB_B2_ROUND(10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0);
    ZKB_B2_ROUND(0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15);
    ZKB_B2_ROUND(14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3);
    h[0] = ZKB_B2_H0 ^ v0 ^ v8;   h[1] = ZKB_B2_IV1 ^ v1 ^ v9;
    h[2] = ZKB_B2_IV2 ^ v2 ^ v10; h[3] = ZKB_B2_IV3 ^ v3 ^ v11;
    h[4] = ZKB_B2_IV4 ^ v4 ^ v12; h[5] = ZKB_B2_IV5 ^ v5 ^ v13;
    h[6] = ZKB_B2_IV6 ^ v6 ^ v14; h[7] = ZKB_B2_IV7 ^ v7 ^ v15;
}


#if defined(__CUDACC__)
// ---- one compression on FOUR lanes (latency mode) ----------------------------------------------
// The upper levels of a tree are a chain of dependent compressions with almost no parallelism
// across nodes; there the time per level is the latency of ONE compression (~3.4 us for a
// single lane executing all 96 G functions).  Here lane q of a quad owns column q of the 4x4
// BLAKE2b state (a_q, b_q, c_q, d_q): the four column G's run in parallel, the rows are rotated
// across the quad with shuffles for the diagonal G's and rotated back: ~4x fewer instructions
// on the critical path.  The 16 message words are read from shared memory (contiguous, 128 B
// at `msg`); which word a lane needs is a compile-time table indexed by q: the four byte
// offsets of one (round, slot) are packed into one 32-bit constant and picked with ONE PRMT
// (`sel` = 0x4440 | q), so a message word costs PRMT + IADD + LDS.64.
struct B2Sigma { uint8_t s[12][16]; };
__host__ __device__ constexpr B2Sigma b2_sigma() {
    return B2Sigma{{{0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3},
                    {11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4}, {7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8},
                    {9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13}, {2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9},
                    {12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11}, {13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10},
                    {6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5}, {10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0},
                    {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3}}};
}
// message word for lane q: slot 0/1 = column x/y (sigma[2q], sigma[2q+1]), slot 2/3 = diagonal x/y
template <int R, int SLOT>
__device__ __forceinline__ uint64_t b2q_msg(const uint8_t* msg, uint32_t sel) {
    constexpr B2Sigma S = b2_sigma();
    constexpr int o = (SLOT & 1) + 8 * (SLOT >> 1);
    constexpr uint32_t packed = (uint32_t)(S.s[R][o] * 8) | ((uint32_t)(S.s[R][o + 2] * 8) << 8) |
                                ((uint32_t)(S.s[R][o + 4] * 8) << 16) | ((uint32_t)(S.s[R][o + 6] * 8) << 24);
    return *reinterpret_cast<const uint64_t*>(msg + __byte_perm(packed, 0u, sel));
}
template <int R>
__device__ __forceinline__ void b2q_round(uint64_t& a, uint64_t& b, uint64_t& c, uint64_t& d, const uint8_t* msg, uint32_t sel,
                                          uint32_t l1, uint32_t l2, uint32_t l3) {
    uint64_t x = b2q_msg<R, 0>(msg, sel), y = b2q_msg<R, 1>(msg, sel);
    uint64_t x2 = b2q_msg<R, 2>(msg, sel), y2 = b2q_msg<R, 3>(msg, sel);
    ZKB_B2_G(a, b, c, d, x, y);
    b = __shfl_sync(0xFFFFFFFFu, b, l1);
    c = __shfl_sync(0xFFFFFFFFu, c, l2);
    d = __shfl_sync(0xFFFFFFFFu, d, l3);
    ZKB_B2_G(a, b, c, d, x2, y2);
    b = __shfl_sync(0xFFFFFFFFu, b, l3);
    c = __shfl_sync(0xFFFFFFFFu, c, l2);
    d = __shfl_sync(0xFFFFFFFFu, d, l1);
}
// One block of t bytes (final): lane q returns digest words q (h_lo) and q + 4 (h_hi).
// All 32 lanes of the warp must call it (shuffles); quads without work pass any readable msg.
__device__ __forceinline__ void blake2b_quad(const uint8_t* msg, uint64_t t, uint32_t lane, uint64_t& h_lo, uint64_t& h_hi) {
    const uint32_t q = lane & 3u, lane_base = lane & ~3u, sel = 0x4440u | q;
    const uint32_t l1 = lane_base | ((q + 1) & 3), l2 = lane_base | ((q + 2) & 3), l3 = lane_base | ((q + 3) & 3);
    const uint64_t iv_lo = q == 0 ? ZKB_B2_IV0 : q == 1 ? ZKB_B2_IV1 : q == 2 ? ZKB_B2_IV2 : ZKB_B2_IV3;
    const uint64_t iv_hi = q == 0 ? ZKB_B2_IV4 : q == 1 ? ZKB_B2_IV5 : q == 2 ? ZKB_B2_IV6 : ZKB_B2_IV7;
    const uint64_t h0 = q == 0 ? ZKB_B2_H0 : iv_lo;
    uint64_t a = h0, b = iv_hi, c = iv_lo;
    uint64_t d = iv_hi ^ (q == 0 ? t : 0ull) ^ (q == 2 ? ~0ull : 0ull);
    b2q_round<0>(a, b, c, d, msg, sel, l1, l2, l3);  b2q_round<1>(a, b, c, d, msg, sel, l1, l2, l3);
    b2q_round<2>(a, b, c, d, msg, sel, l1, l2, l3);  b2q_round<3>(a, b, c, d, msg, sel, l1, l2, l3);
    b2q_round<4>(a, b, c, d, msg, sel, l1, l2, l3);  b2q_round<5>(a, b, c, d, msg, sel, l1, l2, l3);
    b2q_round<6>(a, b, c, d, msg, sel, l1, l2, l3);  b2q_round<7>(a, b, c, d, msg, sel, l1, l2, l3);
    b2q_round<8>(a, b, c, d, msg, sel, l1, l2, l3);  b2q_round<9>(a, b, c, d, msg, sel, l1, l2, l3);
    b2q_round<10>(a, b, c, d, msg, sel, l1, l2, l3); b2q_round<11>(a, b, c, d, msg, sel, l1, l2, l3);
    h_lo = h0 ^ a ^ c;
    h_hi = iv_hi ^ b ^ d;
}
#endif

// node = H(left || right)
ZKB_HD void blake2b_node(const uint64_t (&l)[8], const uint64_t (&r)[8], uint64_t (&h)[8]) {
    uint64_t m[16];
#pragma unroll
    for (int i = 0; i < 8; i++) { m[i] = l[i]; m[8 + i] = r[i]; }
    blake2b_compress_1block(m, 128, h);
}

// ---- u128 -> ASCII decimal, packed little-endian into ten 32-bit message words ------
// (round 2b: 23 quarter-rate multiplies and ~200 ALU-pipe instructions per value instead of 75 + 286 - an IMAD.WIDE blocks ALU issue,
//  DESIGN.md 4, so the long division's and the digit extraction's wide multiplies were ~10 % of the leaf kernels)
ZKB_HD uint32_t b2_mulhi32(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}
// One step of the long division by D = 10^8: x = r * 2^32 + l with r < D; returns floor(x / D), rem = x mod D.
// M = floor(2^58 / D) = 2882303761 leaves 2^58 - M D = 51711744, so x M / 2^58 = x / D - x * 51711744 / (D 2^58) > x / D - 0.771 for every
// x < D 2^32: floor(x M / 2^58) is the quotient or ONE below it.  The remainder of that estimate is < 2 D < 2^32, i.e. exact in 32-bit
// arithmetic, and one compare finishes the step: one wide multiply, one high multiply and one low multiply instead of the 64-bit division
// by a constant (a 64 x 64 high product + a 64-bit multiply-subtract).
ZKB_HD uint32_t div1e8_step(uint32_t r, uint32_t l, uint32_t& rem) {
    const uint32_t D = 100000000u, M = 2882303761u;
    const uint64_t t = (uint64_t)r * M + b2_mulhi32(l, M);          // floor(x * M / 2^32) < 2^59
    const uint32_t q = (uint32_t)(t >> 26);
    const uint32_t rm = l - q * D;                                   // x - q * D, in [0, 2 D)
    const bool up = rm >= D;
    rem = up ? rm - D : rm;
    return up ? q + 1u : q;
}
// four decimal digits of x < 10^4 as ASCII bytes, most significant digit in the LOWEST byte (string order), without a wide
// multiply: x -> two 2-digit lanes (x / 100 low, x % 100 high), both lanes -> tens with ONE multiply (lane * 103 >> 10; 99 * 103 <
// 2^14, the lanes do not meet), then bytes = tens + (lane - 10 tens) << 8 = (lanes << 8) - 2559 tens, + "0000"
ZKB_HD uint32_t dec4_ascii(uint32_t x) {
    const uint32_t q = (x * 5243u) >> 19;                            // x / 100 for x < 10^4
    const uint32_t v = (x << 16) - q * 6553599u;                     // q | (x - 100 q) << 16
    const uint32_t t = ((v * 103u) >> 10) & 0x000F000Fu;             // tens of both lanes
    return (v << 8) + 0x30303030u - t * 2559u;
}

// Writes the decimal string of `a` (canonical, < 2^128) into w[0..9] (bytes in string
// order, byte 0 = lowest byte of w[0], zero padded) and returns its length (1..39).
ZKB_HD uint32_t u128_to_dec_words(const fe& a, uint32_t (&w)[10]) {
    // long division by 10^8 on 32-bit limbs: value = c4*10^32 + c3*10^24 + c2*10^16 + c1*10^8 + c0, c4 < 10^7 (the value is < 10^39)
    uint32_t c0, c1, c2, c3, c4, r;
    // round 1: 4 limbs; the quotient is < 2^101.5, its top limb <= 42
    const uint32_t q3 = div1e8_step(0u, a.v[3], r);
    const uint32_t q2 = div1e8_step(r, a.v[2], r);
    const uint32_t q1 = div1e8_step(r, a.v[1], r);
    const uint32_t q0 = div1e8_step(r, a.v[0], r);
    c0 = r;
    // round 2: quotient < 2^74.9, top limb < 2^11
    r = q3;
    const uint32_t p2 = div1e8_step(r, q2, r);
    const uint32_t p1 = div1e8_step(r, q1, r);
    const uint32_t p0 = div1e8_step(r, q0, r);
    c1 = r;
    // round 3: quotient < 2^48.4, top limb < 2^17
    r = p2;
    const uint32_t s1 = div1e8_step(r, p1, r);
    const uint32_t s0 = div1e8_step(r, p0, r);
    c2 = r;
    // round 4: quotient < 2^21.8
    c4 = div1e8_step(s1, s0, c3);
    // 40 digit positions (the first is always '0'): chunk -> two groups of four digits -> two message words each
    uint32_t g[10];
    {
        const uint32_t cs[5] = {c4, c3, c2, c1, c0};
#pragma unroll
        for (int i = 0; i < 5; i++) {
            const uint32_t hi = b2_mulhi32(cs[i], 0xD1B71759u) >> 13;    // c / 10^4 (exact for every 32-bit c)
            g[2 * i] = hi;
            g[2 * i + 1] = cs[i] - hi * 10000u;
        }
    }
    uint32_t x[10];
#pragma unroll
    for (int k = 0; k < 10; k++) x[k] = dec4_ascii(g[k]);
    // first non-'0' position z (1..39; the value 0 keeps its last digit): the first word whose group is non-zero, then the lowest
    // non-'0' byte inside it
    // (scan the five chunks, then the two groups of the chunk found: 6 compares instead of 10)
    uint32_t zc = 4u, hs = g[8], ls = g[9], xh = x[8], xl = x[9];
    {
        const uint32_t cs[5] = {c4, c3, c2, c1, c0};
#pragma unroll
        for (int i = 3; i >= 0; i--) {
            const bool nz = cs[i] != 0u;
            zc = nz ? (uint32_t)i : zc;
            hs = nz ? g[2 * i] : hs;  ls = nz ? g[2 * i + 1] : ls;
            xh = nz ? x[2 * i] : xh;  xl = nz ? x[2 * i + 1] : xl;
        }
    }
    const bool in_hi = hs != 0u;
    const uint32_t zw = 2u * zc + (in_hi ? 0u : 1u);
    uint32_t ws = in_hi ? xh : xl;
    ws = (hs | ls) != 0u ? ws : 0x31303030u;                          // the value 0: keep the last '0' (the scan ended on chunk 4, word 9)
    const uint32_t raw = ws ^ 0x30303030u;                            // != 0
#if defined(__CUDA_ARCH__)
    const uint32_t zb = (uint32_t)(__ffs((int)raw) - 1) >> 3;
#else
    const uint32_t zb = (raw & 0xFFu) ? 0u : ((raw & 0xFF00u) ? 1u : ((raw & 0xFF0000u) ? 2u : 3u));
#endif
    const uint32_t z = 4u * zw + zb;
    // shift the 40-byte string left by z bytes (zero fill)
    uint32_t y[11];
#if defined(__CUDA_ARCH__)
    const uint32_t sel = 0x3210u + 0x1111u * zb;
#pragma unroll
    for (int k = 0; k < 10; k++) y[k] = __byte_perm(x[k], k < 9 ? x[k + 1] : 0u, sel);
#else
#pragma unroll
    for (int k = 0; k < 10; k++) {
        uint64_t pair = ((uint64_t)(k < 9 ? x[k + 1] : 0u) << 32) | x[k];
        y[k] = (uint32_t)(pair >> (8 * zb));
    }
#endif
    y[10] = 0;
    // word barrel shifter: zw in 0..9
#pragma unroll
    for (int k = 0; k < 10; k++) y[k] = (zw & 1u) ? y[k + 1] : y[k];
#pragma unroll
    for (int k = 0; k < 10; k++) y[k] = (zw & 2u) ? (k + 2 < 10 ? y[k + 2] : 0u) : y[k];
#pragma unroll
    for (int k = 0; k < 10; k++) y[k] = (zw & 4u) ? (k + 4 < 10 ? y[k + 4] : 0u) : y[k];
#pragma unroll
    for (int k = 0; k < 10; k++) y[k] = (zw & 8u) ? (k + 8 < 10 ? y[k + 8] : 0u) : y[k];
#pragma unroll
    for (int k = 0; k < 10; k++) w[k] = y[k];
    return 40u - z;
}

// leaf = H(decimal string of a)
ZKB_HD void blake2b_leaf(const fe& a, uint64_t (&h)[8]) {
    uint32_t w[10];
    uint32_t len = u128_to_dec_words(a, w);
    uint64_t m[16];
#pragma unroll
    for (int i = 0; i < 5; i++) m[i] = ((uint64_t)w[2 * i + 1] << 32) | w[2 * i];
#pragma unroll
    for (int i = 5; i < 16; i++) m[i] = 0;
    blake2b_compress_1block(m, len, h);
}

}  // namespace zkb
