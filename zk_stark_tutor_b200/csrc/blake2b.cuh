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
// x / 10^9 for x < 2^62 (all the long division below needs)
ZKB_HD uint32_t div1e9(uint64_t x, uint32_t& rem) {
    uint64_t q = x / 1000000000ull;
    rem = (uint32_t)(x - q * 1000000000ull);
    return (uint32_t)q;
}
// 9 decimal digits of c (< 10^9), most significant first, one per byte of d[0..8]
ZKB_HD void digits9(uint32_t c, uint32_t (&d)[9]) {
#pragma unroll
    for (int i = 8; i >= 0; i--) {
        uint32_t q = (uint32_t)(((uint64_t)c * 0xCCCCCCCDull) >> 35);   // c / 10
        d[i] = c - q * 10u;
        c = q;
    }
}

// Writes the decimal string of `a` (canonical, < 2^128) into w[0..9] (bytes in string
// order, byte 0 = lowest byte of w[0], zero padded) and returns its length (1..39).
ZKB_HD uint32_t u128_to_dec_words(const fe& a, uint32_t (&w)[10]) {
    // long division by 10^9 on 32-bit limbs: value = c4*10^36 + c3*10^27 + c2*10^18 + c1*10^9 + c0
    uint32_t l0 = a.v[0], l1 = a.v[1], l2 = a.v[2], l3 = a.v[3];
    uint32_t c0, c1, c2, c3, c4, r;
    // round 1: 4 limbs
    uint32_t q3 = l3 / 1000000000u; r = l3 - q3 * 1000000000u;
    uint32_t q2 = div1e9(((uint64_t)r << 32) | l2, r);
    uint32_t q1 = div1e9(((uint64_t)r << 32) | l1, r);
    uint32_t q0 = div1e9(((uint64_t)r << 32) | l0, r);
    c0 = r;
    // round 2: quotient < 2^98.2, q3 <= 4
    r = q3;
    uint32_t p2 = div1e9(((uint64_t)r << 32) | q2, r);
    uint32_t p1 = div1e9(((uint64_t)r << 32) | q1, r);
    uint32_t p0 = div1e9(((uint64_t)r << 32) | q0, r);
    c1 = r;
    // round 3: quotient < 2^68.3, p2 <= 19
    r = p2;
    uint32_t s1 = div1e9(((uint64_t)r << 32) | p1, r);
    uint32_t s0 = div1e9(((uint64_t)r << 32) | p0, r);
    c2 = r;
    // round 4: quotient < 2^38.4
    uint64_t rest = ((uint64_t)s1 << 32) | s0;
    c4 = div1e9(rest, c3);                                  // c4 <= 340
    uint32_t d[40];
    {
        uint32_t t[9];
        digits9(c4, t); d[0] = t[6]; d[1] = t[7]; d[2] = t[8];
        digits9(c3, t);
#pragma unroll
        for (int i = 0; i < 9; i++) d[3 + i] = t[i];
        digits9(c2, t);
#pragma unroll
        for (int i = 0; i < 9; i++) d[12 + i] = t[i];
        digits9(c1, t);
#pragma unroll
        for (int i = 0; i < 9; i++) d[21 + i] = t[i];
        digits9(c0, t);
#pragma unroll
        for (int i = 0; i < 9; i++) d[30 + i] = t[i];
        d[39] = 0;
    }
    uint32_t x[10];
#pragma unroll
    for (int k = 0; k < 10; k++)
        x[k] = d[4 * k] | (d[4 * k + 1] << 8) | (d[4 * k + 2] << 16) | (d[4 * k + 3] << 24);
    // leading zero digits z (0..38; the value 0 keeps one digit)
    uint32_t z = 38;
#pragma unroll
    for (int k = 9; k >= 0; k--) {
        uint32_t xk = (k == 9) ? (x[9] & 0x00FF0000u ? x[9] : (x[9] | 0x00010000u)) : x[k];
        // index of the lowest non-zero byte of xk
        uint32_t lowbit = xk & (0u - xk);
        uint32_t byte = lowbit > 0x00FFFFFFu ? 3u : (lowbit > 0x0000FFFFu ? 2u : (lowbit > 0xFFu ? 1u : 0u));
        z = xk != 0 ? (uint32_t)(4 * k) + byte : z;
    }
    // ASCII, then shift the 39-byte string left by z bytes (zero fill)
#pragma unroll
    for (int k = 0; k < 9; k++) x[k] += 0x30303030u;
    x[9] += 0x00303030u;
    uint32_t zb = z & 3u, zw = z >> 2;
    uint32_t y[11];
#if defined(__CUDA_ARCH__)
    uint32_t sel = 0x3210u + 0x1111u * zb;
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
    return 39u - z;
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
