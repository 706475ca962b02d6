// fe128.cuh - F_p arithmetic for p = 1 + 407*2^119 = 0xCB800000_00000000_00000000_00000001
// on four 32-bit limbs (little-endian, identical bytes to Rust's in-memory u128).
//
// Replaces (results, not algorithms) src/field/field.rs:101-169 of the reference:
// add_mod :109-115, sub_mod :101-107, neg_mod :133-139, mul_mod :117-131 (bit-serial
// there, Montgomery here), and FieldElement ops src/field/field_element.rs:52-143.
//
// Convention used by every kernel: DATA stays canonical (< p, not in Montgomery form)
// so it can be hashed / shipped to the host at any point; CONSTANTS (twiddles, powers
// of the coset offset, 2^-1, n^-1, alpha/x_i) are kept in Montgomery form c*R mod p,
// R = 2^128, so that montmul(data, c*R) = data*c is again canonical.
//
// p > 2^127, so a+b can exceed 2^128: add/sub carry a real carry-out and do one
// conditional correction; no lazy [0,2p) representation fits in four limbs.
//
// Montgomery reduction exploits p = 1 + c*2^119: p^-1 = 1 - c*2^119 (mod 2^128), so the
// quotient digit m' = T_lo * p^-1 is T_lo with one limb adjusted, and m'*p costs four
// products m'_i * 0xCB800000; 16 + 4 (+1 low) 32x32 products per field multiplication.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define ZKB_HD __host__ __device__ __forceinline__
#define ZKB_D __device__ __forceinline__
#else
#define ZKB_HD inline
#define ZKB_D inline
#endif

namespace zkb {

struct alignas(16) fe {
    uint32_t v[4];
};

static constexpr uint32_t P0 = 1u, P3 = 0xCB800000u;
// R = 2^128 mod p, R2 = 2^256 mod p
#define ZKB_FE_R  {{0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0x347FFFFFu}}
#define ZKB_FE_R2 {{0x0E778236u, 0x5BD53A7Fu, 0x1A6AEDC2u, 0xAAF4AD9Au}}

ZKB_HD fe fe_zero() { fe r; r.v[0] = r.v[1] = r.v[2] = r.v[3] = 0; return r; }
ZKB_HD fe fe_from_u32(uint32_t x) { fe r = fe_zero(); r.v[0] = x; return r; }
ZKB_HD bool fe_is_zero(const fe& a) { return (a.v[0] | a.v[1] | a.v[2] | a.v[3]) == 0; }
ZKB_HD bool fe_eq(const fe& a, const fe& b) {
    return ((a.v[0] ^ b.v[0]) | (a.v[1] ^ b.v[1]) | (a.v[2] ^ b.v[2]) | (a.v[3] ^ b.v[3])) == 0;
}
// a >= p ?   (p = {1,0,0,P3})
ZKB_HD bool fe_ge_p(const fe& a) {
    return a.v[3] > P3 || (a.v[3] == P3 && (a.v[0] | a.v[1] | a.v[2]) != 0);
}

// (a + b) mod p, a, b < p
ZKB_HD fe fe_add(const fe& a, const fe& b) {
    fe r, t;
#if defined(__CUDA_ARCH__)
    uint32_t carry, borrow;
    asm("add.cc.u32 %0, %5, %9;\n\t"
        "addc.cc.u32 %1, %6, %10;\n\t"
        "addc.cc.u32 %2, %7, %11;\n\t"
        "addc.cc.u32 %3, %8, %12;\n\t"
        "addc.u32 %4, 0, 0;"
        : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(carry)
        : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]),
          "r"(b.v[0]), "r"(b.v[1]), "r"(b.v[2]), "r"(b.v[3]));
    asm("sub.cc.u32 %0, %5, 1;\n\t"
        "subc.cc.u32 %1, %6, 0;\n\t"
        "subc.cc.u32 %2, %7, 0;\n\t"
        "subc.cc.u32 %3, %8, 0xCB800000;\n\t"
        "subc.u32 %4, 0, 0;"
        : "=r"(t.v[0]), "=r"(t.v[1]), "=r"(t.v[2]), "=r"(t.v[3]), "=r"(borrow)
        : "r"(r.v[0]), "r"(r.v[1]), "r"(r.v[2]), "r"(r.v[3]));
    bool take = (carry != 0) || (borrow == 0);
#else
    uint64_t c = 0;
    for (int i = 0; i < 4; i++) { c += (uint64_t)a.v[i] + b.v[i]; r.v[i] = (uint32_t)c; c >>= 32; }
    static const uint32_t PL[4] = {P0, 0, 0, P3};
    int64_t bw = 0;
    for (int i = 0; i < 4; i++) { bw += (int64_t)r.v[i] - PL[i]; t.v[i] = (uint32_t)bw; bw >>= 32; }
    bool take = (c != 0) || (bw == 0);
#endif
    for (int i = 0; i < 4; i++) r.v[i] = take ? t.v[i] : r.v[i];
    return r;
}

// (a - b) mod p, a, b < p
ZKB_HD fe fe_sub(const fe& a, const fe& b) {
    fe r;
#if defined(__CUDA_ARCH__)
    uint32_t borrow;
    asm("sub.cc.u32 %0, %5, %9;\n\t"
        "subc.cc.u32 %1, %6, %10;\n\t"
        "subc.cc.u32 %2, %7, %11;\n\t"
        "subc.cc.u32 %3, %8, %12;\n\t"
        "subc.u32 %4, 0, 0;"
        : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(borrow)
        : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]),
          "r"(b.v[0]), "r"(b.v[1]), "r"(b.v[2]), "r"(b.v[3]));
    // borrow is 0 or 0xFFFFFFFF; add p under that mask
    asm("add.cc.u32 %0, %0, %4;\n\t"
        "addc.cc.u32 %1, %1, 0;\n\t"
        "addc.cc.u32 %2, %2, 0;\n\t"
        "addc.u32 %3, %3, %5;"
        : "+r"(r.v[0]), "+r"(r.v[1]), "+r"(r.v[2]), "+r"(r.v[3])
        : "r"(borrow & 1u), "r"(borrow & P3));
#else
    int64_t bw = 0;
    for (int i = 0; i < 4; i++) { bw += (int64_t)a.v[i] - b.v[i]; r.v[i] = (uint32_t)bw; bw >>= 32; }
    if (bw) {
        static const uint32_t PL[4] = {P0, 0, 0, P3};
        uint64_t c = 0;
        for (int i = 0; i < 4; i++) { c += (uint64_t)r.v[i] + PL[i]; r.v[i] = (uint32_t)c; c >>= 32; }
    }
#endif
    return r;
}

ZKB_HD fe fe_neg(const fe& a) { return fe_sub(fe_zero(), a); }

// a/2 mod p: a even -> a>>1, a odd -> (a+p)>>1 (a+p may carry into bit 128)
ZKB_HD fe fe_half(const fe& a) {
    uint32_t odd = a.v[0] & 1u;
    uint32_t m0 = odd, m3 = odd ? P3 : 0u;
    uint64_t c = (uint64_t)a.v[0] + m0;
    uint32_t s0 = (uint32_t)c; c >>= 32;
    c += a.v[1]; uint32_t s1 = (uint32_t)c; c >>= 32;
    c += a.v[2]; uint32_t s2 = (uint32_t)c; c >>= 32;
    c += (uint64_t)a.v[3] + m3; uint32_t s3 = (uint32_t)c; uint32_t s4 = (uint32_t)(c >> 32);
    fe r;
    r.v[0] = (s0 >> 1) | (s1 << 31);
    r.v[1] = (s1 >> 1) | (s2 << 31);
    r.v[2] = (s2 >> 1) | (s3 << 31);
    r.v[3] = (s3 >> 1) | (s4 << 31);
    return r;
}

// Montgomery product a*b*2^-128 mod p.  Needs a*b < p*2^128 (true when either operand
// is canonical); result canonical.
#if defined(__CUDA_ARCH__)
// Device version in PTX.  The 4x4 product is accumulated in two interleaved 8-limb accumulators
// ("even" holds a_i*b_j with i+j even at limb i+j, "odd" those with i+j odd) so that every
// 32x32->64 product lands on a fixed, aligned register pair and maps to one IMAD.WIDE with
// carry-out, rows chained by carry; the two are merged with one 7-limb add.  The reduction
// uses the shape of p = 1 + c*2^119 (P3 = c << 23): p^-1 mod 2^128 = 1 - P3*2^96, so the quotient digit
// m' = T_lo * p^-1 mod 2^128 is T_lo itself with limb 3 replaced by t3 - lo32(t0*P3) - no negation, no carry chain.
// m'*p = m' + (m'*P3 << 96) agrees with T on the low 128 bits, so (T - m'*p) / 2^128 = T_hi - hi128(m'*p) with no borrow from
// below, and hi128(m'*p) = (m'*P3 + m'_3) >> 32 is ONE chain of four wide multiply-adds (the first absorbs m'_3 and with it the
// carry out of limb 3; its low word is t3 again and is dropped).  The difference lies in (-p, p): p is added back under the
// borrow mask.  24 ALU-pipe instructions + 21 quarter-rate multiplies (round 1/2a form with m = -T_lo/p: 34 + 21).
__device__ __forceinline__ fe fe_montmul_ptx(const fe& a, const fe& b) {
    fe r;
    asm("{\n\t"
        ".reg .u32 e0,e1,e2,e3,e4,e5,e6,e7,o0,o1,o2,o3,o4,o5,o6;\n\t"
        ".reg .u32 t1,t2,t3,t4,t5,t6,t7,m3,d0,x0,h0,h1,h2,q1,q2,q3,q4,r0,r1,r2,r3,bw,a0,a3;\n\t"
        ".reg .u64 w0,w1,w2,w3;\n\t"
        // ---- even accumulator: row b0
        "mul.wide.u32 w0, %4, %8;\n\t" "mov.b64 {e0,e1}, w0;\n\t"
        "mul.wide.u32 w1, %6, %8;\n\t" "mov.b64 {e2,e3}, w1;\n\t"
        // odd accumulator: row b0 (a1*b0 at limb 1, a3*b0 at limb 3)
        "mul.wide.u32 w2, %5, %8;\n\t" "mov.b64 {o0,o1}, w2;\n\t"
        "mul.wide.u32 w3, %7, %8;\n\t" "mov.b64 {o2,o3}, w3;\n\t"
        // ---- row b1: odd += a0*b1 (limb 1), a2*b1 (limb 3); even += a1*b1 (limb 2), a3*b1 (limb 4)
        "mad.lo.cc.u32 o0, %4, %9, o0;\n\t"  "madc.hi.cc.u32 o1, %4, %9, o1;\n\t"
        "madc.lo.cc.u32 o2, %6, %9, o2;\n\t" "madc.hi.cc.u32 o3, %6, %9, o3;\n\t"
        "addc.u32 o4, 0, 0;\n\t"
        "mad.lo.cc.u32 e2, %5, %9, e2;\n\t"  "madc.hi.cc.u32 e3, %5, %9, e3;\n\t"
        "madc.lo.cc.u32 e4, %7, %9, 0;\n\t"  "madc.hi.cc.u32 e5, %7, %9, 0;\n\t"
        "addc.u32 e6, 0, 0;\n\t"
        // ---- row b2: even += a0*b2 (limb 2), a2*b2 (limb 4); odd += a1*b2 (limb 3), a3*b2 (limb 5)
        "mad.lo.cc.u32 e2, %4, %10, e2;\n\t" "madc.hi.cc.u32 e3, %4, %10, e3;\n\t"
        "madc.lo.cc.u32 e4, %6, %10, e4;\n\t" "madc.hi.cc.u32 e5, %6, %10, e5;\n\t"
        "addc.u32 e6, e6, 0;\n\t"
        "mad.lo.cc.u32 o2, %5, %10, o2;\n\t" "madc.hi.cc.u32 o3, %5, %10, o3;\n\t"
        "madc.lo.cc.u32 o4, %7, %10, o4;\n\t" "madc.hi.cc.u32 o5, %7, %10, 0;\n\t"
        "addc.u32 o6, 0, 0;\n\t"
        // ---- row b3: odd += a0*b3 (limb 3), a2*b3 (limb 5); even += a1*b3 (limb 4), a3*b3 (limb 6)
        "mad.lo.cc.u32 o2, %4, %11, o2;\n\t" "madc.hi.cc.u32 o3, %4, %11, o3;\n\t"
        "madc.lo.cc.u32 o4, %6, %11, o4;\n\t" "madc.hi.cc.u32 o5, %6, %11, o5;\n\t"
        "addc.u32 o6, o6, 0;\n\t"
        "mad.lo.cc.u32 e4, %5, %11, e4;\n\t" "madc.hi.cc.u32 e5, %5, %11, e5;\n\t"
        "madc.lo.cc.u32 e6, %7, %11, e6;\n\t" "madc.hi.u32 e7, %7, %11, 0;\n\t"
        // ---- merge: T = even + (odd << 32); T0 = e0
        "add.cc.u32 t1, e1, o0;\n\t"  "addc.cc.u32 t2, e2, o1;\n\t" "addc.cc.u32 t3, e3, o2;\n\t"
        "addc.cc.u32 t4, e4, o3;\n\t" "addc.cc.u32 t5, e5, o4;\n\t" "addc.cc.u32 t6, e6, o5;\n\t"
        "addc.u32 t7, e7, o6;\n\t"
        // ---- m' = (t0, t1, t2, m3), m3 = t3 - lo32(t0*P3);  hi128(m'*p) = (m'*P3 + m3) >> 32 = (q1, q2, q3, q4)
        "mul.lo.u32 d0, e0, 0xCB800000;\n\t"
        "sub.u32 m3, t3, d0;\n\t"
        "mad.lo.cc.u32 x0, e0, 0xCB800000, m3;\n\t" "madc.hi.u32 h0, e0, 0xCB800000, 0;\n\t"
        "mad.lo.cc.u32 q1, t1, 0xCB800000, h0;\n\t" "madc.hi.u32 h1, t1, 0xCB800000, 0;\n\t"
        "mad.lo.cc.u32 q2, t2, 0xCB800000, h1;\n\t" "madc.hi.u32 h2, t2, 0xCB800000, 0;\n\t"
        "mad.lo.cc.u32 q3, m3, 0xCB800000, h2;\n\t" "madc.hi.u32 q4, m3, 0xCB800000, 0;\n\t"
        // ---- r = T_hi - (q1..q4) in (-p, p); + p = {1, 0, 0, P3} under the borrow mask
        "sub.cc.u32 r0, t4, q1;\n\t" "subc.cc.u32 r1, t5, q2;\n\t" "subc.cc.u32 r2, t6, q3;\n\t" "subc.cc.u32 r3, t7, q4;\n\t"
        "subc.u32 bw, 0, 0;\n\t"
        "and.b32 a0, bw, 1;\n\t" "and.b32 a3, bw, 0xCB800000;\n\t"
        "add.cc.u32 %0, r0, a0;\n\t" "addc.cc.u32 %1, r1, 0;\n\t" "addc.cc.u32 %2, r2, 0;\n\t" "addc.u32 %3, r3, a3;\n\t"
        "}"
        : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3])
        : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(b.v[0]), "r"(b.v[1]), "r"(b.v[2]), "r"(b.v[3]));
    return r;
}
#endif
ZKB_HD fe fe_montmul(const fe& a, const fe& b) {
#if defined(__CUDA_ARCH__)
    return fe_montmul_ptx(a, b);
#else
    uint32_t t[9];
    // ---- 4x4 schoolbook product, row by row with 64-bit accumulation
    uint64_t acc;
    acc = (uint64_t)a.v[0] * b.v[0];                 t[0] = (uint32_t)acc; acc >>= 32;
    acc += (uint64_t)a.v[1] * b.v[0];                t[1] = (uint32_t)acc; acc >>= 32;
    acc += (uint64_t)a.v[2] * b.v[0];                t[2] = (uint32_t)acc; acc >>= 32;
    acc += (uint64_t)a.v[3] * b.v[0];                t[3] = (uint32_t)acc; t[4] = (uint32_t)(acc >> 32);
#pragma unroll
    for (int i = 1; i < 4; i++) {
        acc = (uint64_t)a.v[0] * b.v[i] + t[i];      t[i] = (uint32_t)acc; acc >>= 32;
        acc += (uint64_t)a.v[1] * b.v[i] + t[i + 1]; t[i + 1] = (uint32_t)acc; acc >>= 32;
        acc += (uint64_t)a.v[2] * b.v[i] + t[i + 2]; t[i + 2] = (uint32_t)acc; acc >>= 32;
        acc += (uint64_t)a.v[3] * b.v[i] + t[i + 3]; t[i + 3] = (uint32_t)acc; t[i + 4] = (uint32_t)(acc >> 32);
    }
    // ---- reduce limbs 0..2 at once: m = -T mod 2^96; T + m*p = T + m + (m*P3) << 96.
    // T_lo96 + m == 2^96 exactly when T_lo96 != 0 (carry 1 into limb 3), else 0.
    uint32_t nz = (t[0] | t[1] | t[2]) != 0 ? 1u : 0u;
    uint32_t m0 = 0u - t[0];
    uint32_t m1 = ~t[1] + (t[0] == 0 ? 1u : 0u);
    uint32_t m2 = ~t[2] + ((t[0] | t[1]) == 0 ? 1u : 0u);
    acc = (uint64_t)m0 * P3 + t[3] + nz;             t[3] = (uint32_t)acc; acc >>= 32;
    acc += (uint64_t)m1 * P3 + t[4];                 t[4] = (uint32_t)acc; acc >>= 32;
    acc += (uint64_t)m2 * P3 + t[5];                 t[5] = (uint32_t)acc; acc >>= 32;
    acc += t[6];                                     t[6] = (uint32_t)acc; acc >>= 32;
    acc += t[7];                                     t[7] = (uint32_t)acc; t[8] = (uint32_t)(acc >> 32);
    // ---- reduce limb 3: m3 = -t3 mod 2^32; t3 + m3 -> 0 carry (t3 != 0); m3*P3 lands on limbs 6,7
    uint32_t m3 = 0u - t[3];
    uint32_t c3 = t[3] != 0 ? 1u : 0u;
    acc = (uint64_t)t[4] + c3;                       t[4] = (uint32_t)acc; acc >>= 32;
    acc += t[5];                                     t[5] = (uint32_t)acc; acc >>= 32;
    acc += (uint64_t)m3 * P3 + t[6];                 t[6] = (uint32_t)acc; acc >>= 32;
    acc += t[7];                                     t[7] = (uint32_t)acc; acc >>= 32;
    uint32_t top = t[8] + (uint32_t)acc;             // value = top*2^128 + t[7..4] < 2p
    fe r; r.v[0] = t[4]; r.v[1] = t[5]; r.v[2] = t[6]; r.v[3] = t[7];
    // conditional subtract p
    uint32_t d0 = r.v[0] - 1u;
    uint32_t b0 = r.v[0] < 1u ? 1u : 0u;
    uint32_t d1 = r.v[1] - b0;
    uint32_t b1 = r.v[1] < b0 ? 1u : 0u;
    uint32_t d2 = r.v[2] - b1;
    uint32_t b2 = r.v[2] < b1 ? 1u : 0u;
    uint64_t d3w = (uint64_t)r.v[3] - P3 - b2;
    uint32_t d3 = (uint32_t)d3w;
    bool borrow = (d3w >> 63) != 0;
    bool take = (top != 0) || !borrow;
    r.v[0] = take ? d0 : r.v[0];
    r.v[1] = take ? d1 : r.v[1];
    r.v[2] = take ? d2 : r.v[2];
    r.v[3] = take ? d3 : r.v[3];
    return r;
#endif
}

ZKB_HD fe fe_mont_one() { fe r = ZKB_FE_R; return r; }
ZKB_HD fe fe_to_mont(const fe& a) { fe r2 = ZKB_FE_R2; return fe_montmul(a, r2); }
ZKB_HD fe fe_from_mont(const fe& a) { return fe_montmul(a, fe_from_u32(1)); }
// plain product of two canonical values (two Montgomery products)
ZKB_HD fe fe_mul(const fe& a, const fe& b) { return fe_montmul(fe_to_mont(a), b); }

// base^e, base in Montgomery form -> result in Montgomery form
ZKB_HD fe fe_mont_pow(fe base, uint64_t e) {
    fe acc = fe_mont_one();
    while (e) {
        if (e & 1) acc = fe_montmul(acc, base);
        base = fe_montmul(base, base);
        e >>= 1;
    }
    return acc;
}

#if defined(__CUDACC__)
ZKB_D fe fe_load(const fe* p) {
    uint4 x = *reinterpret_cast<const uint4*>(p);
    fe r; r.v[0] = x.x; r.v[1] = x.y; r.v[2] = x.z; r.v[3] = x.w; return r;
}
ZKB_D fe fe_ldg(const fe* p) {
    uint4 x = __ldg(reinterpret_cast<const uint4*>(p));
    fe r; r.v[0] = x.x; r.v[1] = x.y; r.v[2] = x.z; r.v[3] = x.w; return r;
}
ZKB_D void fe_store(fe* p, const fe& a) {
    *reinterpret_cast<uint4*>(p) = make_uint4(a.v[0], a.v[1], a.v[2], a.v[3]);
}
#endif

}  // namespace zkb
