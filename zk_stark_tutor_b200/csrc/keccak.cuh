// keccak.cuh - Fiat-Shamir on the device: SHAKE256 (Keccak-f[1600], FIPS 202) over the proof-stream
// transcript, one WARP per sponge.
//
// Replaces, for the rounds of FRI::commit (src/fri.rs:136-146), the host hop
//   push Root -> proof_stream.fiat_shamir_prover(32) -> Field::sample
// i.e. src/proof_stream.rs:36-40 (SHAKE256 of the whole serialised stream, crate sha3 0.10.8 via
// src/crypto/shake256.rs:7-19), src/stark/proof_stream_enum.rs:67-127,161-190 (the Root record:
// code 0 || u64_be(64) || 64 bytes; 16-byte order header) and src/field/field.rs:87-99 (sample).
//
// The reference re-serialises and re-hashes the whole stream per challenge.  Here the sponge is
// INCREMENTAL: the host absorbs what the transcript holds when FRI::commit starts (complete
// 136-byte blocks; SignatureProofStream's prefix and the zero-header quirk are just bytes of that
// transcript) and hands over the 200-byte state plus the partial block; per round the device appends
// the 73-byte Root record, absorbs a block when one completes, and for a challenge finalises a COPY of
// the state (pad 0x1F .. 0x80, one permutation), takes the last 16 of the 32 squeezed bytes as a
// big-endian integer mod p (Field::sample) and multiplies by 1/offset_r for the fold.
//
// Keccak-f on a warp: lane x + 5y holds state lane A[x][y]; theta = 4 + 2 shuffles, rho = a per-lane
// rotate, pi + chi = 3 shuffles (three dependent shuffle stages per round; measured 283 clk per round on B200
// = 3.5 us per permutation; one thread holding all 25 lanes would be bound by ALU issue at ~360 clk per round).  The challenge sits on the critical path between two
// FRI layers.
#pragma once
#include <stdint.h>
#include "fe128.cuh"

namespace zkb {

#define ZKB_FS_RATE 136u            // SHAKE256 rate in bytes
#define ZKB_FS_MAX_ROUNDS 64u

// Sponge state in the middle of a transcript (host and device layout)
struct FsSponge {
    uint64_t st[25];                // after absorbing every complete block
    uint8_t buf[ZKB_FS_RATE + 80];  // the partial block (fill < 136 bytes between calls; room for one 73-byte record)
    uint32_t fill;
    uint32_t pad_;
};

// Device-resident Fiat-Shamir context of one FRI::commit (one per instance of a batch)
struct FsDev {
    FsSponge sp;
    fe kk_m;                                // alpha_r / offset_r in Montgomery form: the fold constant of round r + 1
    fe alpha;                               // the last challenge (canonical)
    fe inv_off_m2[ZKB_FS_MAX_ROUNDS];       // (1 / offset_r) * R^2: montmul(alpha, .) = (alpha / offset_r) * R
    uint8_t roots[ZKB_FS_MAX_ROUNDS][64];   // Merkle root of every round, in order (what the host pushes afterwards)
};

#if defined(__CUDACC__)

__device__ __forceinline__ uint64_t rotl64_var(uint64_t v, uint32_t n) {   // n in 0..63
    uint32_t lo = (uint32_t)v, hi = (uint32_t)(v >> 32);
    if (n & 32u) { uint32_t t = lo; lo = hi; hi = t; }
    const uint32_t s = n & 31u;
    const uint32_t nlo = __funnelshift_l(hi, lo, s), nhi = __funnelshift_l(lo, hi, s);
    return ((uint64_t)nhi << 32) | nlo;
}

// One Keccak-f[1600] permutation; lane l < 25 holds A[l % 5][l / 5] (= state word l).  All 32 lanes call.
// Per round THREE shuffle stages sit on the dependency chain: (1) the four other lanes of the column for theta's
// parity, (2) the parities of columns x-1 / x+1, (3) rho+pi+chi in one go - every lane fetches the three rotated
// values its chi needs straight from their pre-pi source lanes (B[x'][y] = rot(A[(x' + 3y) % 5][x']), x' = x, x+1, x+2).
// Fully unrolled: the round constant is an immediate.
struct KeccakLane {
    uint32_t rho, col1, col2, col3, col4, xm1, xp1, src0, src1, src2;
};
__device__ __forceinline__ KeccakLane keccak_lane(uint32_t lane) {
    // rho offsets r[x][y] at index x + 5y
    const uint32_t RHO = (lane == 0) ? 0 : (lane == 1) ? 1 : (lane == 2) ? 62 : (lane == 3) ? 28 : (lane == 4) ? 27 :
                         (lane == 5) ? 36 : (lane == 6) ? 44 : (lane == 7) ? 6 : (lane == 8) ? 55 : (lane == 9) ? 20 :
                         (lane == 10) ? 3 : (lane == 11) ? 10 : (lane == 12) ? 43 : (lane == 13) ? 25 : (lane == 14) ? 39 :
                         (lane == 15) ? 41 : (lane == 16) ? 45 : (lane == 17) ? 15 : (lane == 18) ? 21 : (lane == 19) ? 8 :
                         (lane == 20) ? 18 : (lane == 21) ? 2 : (lane == 22) ? 61 : (lane == 23) ? 56 : 14;
    const uint32_t l = lane < 25 ? lane : 24;          // idle lanes mirror lane 24 (valid shuffle sources only)
    const uint32_t x = l % 5u, y = l / 5u;
    KeccakLane k;
    k.rho = RHO;
    k.col1 = x + 5u * ((y + 1u) % 5u); k.col2 = x + 5u * ((y + 2u) % 5u);
    k.col3 = x + 5u * ((y + 3u) % 5u); k.col4 = x + 5u * ((y + 4u) % 5u);
    k.xm1 = (x + 4u) % 5u; k.xp1 = (x + 1u) % 5u;
    const uint32_t x1 = (x + 1u) % 5u, x2 = (x + 2u) % 5u;
    k.src0 = ((x + 3u * y) % 5u) + 5u * x;
    k.src1 = ((x1 + 3u * y) % 5u) + 5u * x1;
    k.src2 = ((x2 + 3u * y) % 5u) + 5u * x2;
    return k;
}
// (An exchange through shared memory - one 64-bit STS + k LDS per stage - was measured too: 304 clk per round
// against 283 for the shuffles below; the first version, four stages + round constant from constant memory: 367.)
template <int ROUND>
__device__ __forceinline__ uint64_t keccak_round(uint64_t a, const KeccakLane& k, uint32_t lane) {
    constexpr uint64_t RC[24] = {
        0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808aULL, 0x8000000080008000ULL,
        0x000000000000808bULL, 0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL,
        0x000000000000008aULL, 0x0000000000000088ULL, 0x0000000080008009ULL, 0x000000008000000aULL,
        0x000000008000808bULL, 0x800000000000008bULL, 0x8000000000008089ULL, 0x8000000000008003ULL,
        0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800aULL, 0x800000008000000aULL,
        0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
    const unsigned FULL = 0xFFFFFFFFu;
    // theta
    const uint64_t c = a ^ __shfl_sync(FULL, a, k.col1) ^ __shfl_sync(FULL, a, k.col2) ^ __shfl_sync(FULL, a, k.col3) ^ __shfl_sync(FULL, a, k.col4);
    const uint64_t cp = __shfl_sync(FULL, c, k.xp1);
    a ^= __shfl_sync(FULL, c, k.xm1) ^ ((cp << 1) | (cp >> 63));
    // rho, then pi + chi
    const uint64_t r = rotl64_var(a, k.rho);
    const uint64_t b0 = __shfl_sync(FULL, r, k.src0), b1 = __shfl_sync(FULL, r, k.src1), b2 = __shfl_sync(FULL, r, k.src2);
    a = b0 ^ (~b1 & b2);
    // iota
    if (lane == 0) a ^= RC[ROUND];
    return a;
}
static __device__ __noinline__ uint64_t keccak_f_warp(uint64_t a, uint32_t lane) {
    const KeccakLane k = keccak_lane(lane);
#define ZKB_KR(i) a = keccak_round<i>(a, k, lane)
    ZKB_KR(0);  ZKB_KR(1);  ZKB_KR(2);  ZKB_KR(3);  ZKB_KR(4);  ZKB_KR(5);  ZKB_KR(6);  ZKB_KR(7);
    ZKB_KR(8);  ZKB_KR(9);  ZKB_KR(10); ZKB_KR(11); ZKB_KR(12); ZKB_KR(13); ZKB_KR(14); ZKB_KR(15);
    ZKB_KR(16); ZKB_KR(17); ZKB_KR(18); ZKB_KR(19); ZKB_KR(20); ZKB_KR(21); ZKB_KR(22); ZKB_KR(23);
#undef ZKB_KR
    return a;
}

// SHAKE256(msg)[0..8*nwords) for a message in memory any thread of the warp can read (a test hook and the reference for
// the incremental sponge below).  out[i], i < nwords <= 17: written by lane i.
__device__ __forceinline__ void shake256_warp(const uint8_t* msg, uint32_t len, uint64_t* out, uint32_t nwords, uint32_t lane) {
    uint64_t a = 0;
    uint32_t off = 0;
    for (;;) {
        const uint32_t take = len - off < ZKB_FS_RATE ? len - off : ZKB_FS_RATE;
        uint64_t w = 0;
        if (lane < 17) {
#pragma unroll
            for (uint32_t k = 0; k < 8; k++) {
                const uint32_t i = 8 * lane + k;
                uint32_t byte = i < take ? msg[off + i] : 0u;
                if (take < ZKB_FS_RATE) { if (i == take) byte ^= 0x1Fu; if (i == ZKB_FS_RATE - 1) byte ^= 0x80u; }
                w |= (uint64_t)byte << (8 * k);
            }
        }
        a = keccak_f_warp(a ^ w, lane);
        off += take;
        if (take < ZKB_FS_RATE) break;
    }
    if (lane < nwords) out[lane] = a;
}

// One FRI round of the transcript, executed by one warp on a sponge in SHARED memory:
//   push Root(root)                                    (fri.rs:136-137, proof_stream_enum.rs:76-83)
//   if want_alpha: alpha = sample(fiat_shamir(32))     (fri.rs:145-146); returns alpha on every lane
// `root` = 64 bytes readable by the warp (shared or global).
__device__ __forceinline__ fe fs_round_warp(FsSponge* sp, const uint8_t* root, bool want_alpha, uint32_t lane) {
    uint32_t fill = sp->fill;
    // record: code 0, u64_be(64), 64 root bytes
    for (uint32_t i = lane; i < 73; i += 32) sp->buf[fill + i] = i < 8 ? 0u : i == 8 ? 64u : root[i - 9];
    __syncwarp();
    fill += 73;
    uint64_t a = lane < 25 ? sp->st[lane] : 0ull;
    const uint64_t* bw = reinterpret_cast<const uint64_t*>(sp->buf);
    if (fill >= ZKB_FS_RATE) {
        a = keccak_f_warp(a ^ (lane < 17 ? bw[lane] : 0ull), lane);
        const uint32_t rem = fill - ZKB_FS_RATE;            // <= 72
        uint8_t t0 = 0, t1 = 0, t2 = 0;
        if (lane < rem) t0 = sp->buf[ZKB_FS_RATE + lane];
        if (lane + 32 < rem) t1 = sp->buf[ZKB_FS_RATE + lane + 32];
        if (lane + 64 < rem) t2 = sp->buf[ZKB_FS_RATE + lane + 64];
        __syncwarp();
        if (lane < rem) sp->buf[lane] = t0;
        if (lane + 32 < rem) sp->buf[lane + 32] = t1;
        if (lane + 64 < rem) sp->buf[lane + 64] = t2;
        fill = rem;
        if (lane < 25) sp->st[lane] = a;
    }
    __syncwarp();
    if (lane == 0) sp->fill = fill;
    fe alpha = fe_zero();
    if (want_alpha) {
        uint64_t w = 0;
        if (lane < 17) {
#pragma unroll
            for (uint32_t k = 0; k < 8; k++) {
                const uint32_t i = 8 * lane + k;
                uint32_t byte = i < fill ? sp->buf[i] : 0u;
                if (i == fill) byte ^= 0x1Fu;
                if (i == ZKB_FS_RATE - 1) byte ^= 0x80u;
                w |= (uint64_t)byte << (8 * k);
            }
        }
        const uint64_t t = keccak_f_warp(a ^ w, lane);
        // Field::sample: the last 16 of the 32 squeezed bytes (state words 2, 3) as a big-endian integer, mod p
        const uint64_t w2 = __shfl_sync(0xFFFFFFFFu, t, 2), w3 = __shfl_sync(0xFFFFFFFFu, t, 3);
        const uint32_t w2l = (uint32_t)w2, w2h = (uint32_t)(w2 >> 32), w3l = (uint32_t)w3, w3h = (uint32_t)(w3 >> 32);
        alpha.v[0] = __byte_perm(w3h, 0, 0x0123);           // byte 31 is the least significant
        alpha.v[1] = __byte_perm(w3l, 0, 0x0123);
        alpha.v[2] = __byte_perm(w2h, 0, 0x0123);
        alpha.v[3] = __byte_perm(w2l, 0, 0x0123);
        if (fe_ge_p(alpha)) {                               // value < 2^128 < 2p: one subtraction
            fe p; p.v[0] = P0; p.v[1] = 0; p.v[2] = 0; p.v[3] = P3;
            uint32_t bw_ = 0;
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const uint64_t d = (uint64_t)alpha.v[i] - p.v[i] - bw_;
                alpha.v[i] = (uint32_t)d;
                bw_ = (uint32_t)(d >> 63);
            }
        }
    }
    __syncwarp();
    return alpha;
}

#endif  // __CUDACC__

}  // namespace zkb
