// ntt.cu - forward / inverse NTT, coset LDE and NTT-based polynomial arithmetic.
//
// Replaces the bodies of (reference file:line):
//   ntt                  src/fft/ntt.rs:7-49      (+ bit_reverse_copy, src/utils/bit_reverse_copy.rs)
//   intt                 src/fft/ntt.rs:51-68
//   Polynomial::scale    src/field/polynomial.rs:109-121
//   fast_coset_evaluate  src/fft/ntt_arithmetics.rs:161-170   (the LDE)
//   fast_multiply        src/fft/ntt_arithmetics.rs:5-64
//   fast_coset_divide    src/fft/ntt_arithmetics.rs:239-310
//
// Algorithm (results identical - exact arithmetic - but not the reference's single
// radix-2 loop): N = N1*N2*N3 is transformed in up to three HBM passes (Bailey 4-step
// applied twice).  Every pass is the same kernel: a CTA stages a tile of S x B values
// (S = transform length of the pass, B = independent transforms that are CONTIGUOUS in
// HBM so that every global access is a B*16-byte segment) in shared memory, runs all
// log2(S) radix-2 stages there, multiplies by the inter-pass twiddle w_N^(k*col) taken
// from a two-level power table, and writes the tile back.  Passes 1/2 are DIT with the
// bit reversal folded into the global load address (free); the final pass is DIF with
// the bit reversal folded into the shared-memory read of the (transposing) store.
//   x[n1*M + m] --pass1: N1-point over n1, *w_N^(k1*m)--> A[k1*M + m]      (M = N2*N3)
//   A[k1*M + n2*N3 + n3] --pass2: N2-point over n2, *w_M^(k2*n3)--> in place
//   A[k1*M + k2*N3 + n3] --pass3: N3-point over n3--> X[k1 + N1*k2 + N1*N2*k3]
// Zero padding (ntt.rs pads to the next power of two; the LDE pads N/ef coefficients
// to N) is never materialised: loads beyond n_in read as zero, and the leading DIT
// stages whose odd inputs are all zero are replaced by a broadcast (2 of 24 stages for
// ef = 4).  The coset scaling c_i*offset^i of the LDE is fused into the pass-1 load.
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "ctx.hpp"
#include "ntt.cuh"
#include "merkle_dev.cuh"

namespace zkb {

struct PassParams {
    const fe* in;
    fe* out;
    uint32_t log_s, log_b;
    uint32_t transposed;         // 0: passes 1/2 (b fastest everywhere, DIT), 1: final pass (DIF)
    uint32_t inner_count;
    uint64_t ld_s, ld_b, ld_outer, ld_inner;
    uint64_t st_k, st_b, st_outer, st_inner;
    uint64_t in_batch, out_batch;
    uint64_t n_valid;            // loads at logical index >= n_valid read as zero
    uint32_t has_valid;
    uint32_t skip_log;           // leading DIT stages replaced by a broadcast
    uint64_t tw_mul;             // inter-pass twiddle exponent = k * col * tw_mul
    DevPow tw;
    uint32_t has_tw;
    uint32_t has_scale;          // multiply loaded value by base^(logical index) (LDE)
    DevPow sc;
    uint32_t has_post;           // multiply stored value by `post` (n^-1 for the iNTT)
    fe post;
    const fe* tw_s;              // w_S^j * R, j < S
    const fe* tw_x;              // k_ntt_rr<.., .., false, false, true>: the pass's inter-pass twiddles themselves, tw_x[k * tw_x_cols + col] =
    uint64_t tw_x_cols;          //   w^(k * col * tw_mul) * R (* n^-1 for an inverse transform) - see tw_exact_table/2
    // two-level outer index (four-pass transforms): outer = oa + outer_a_count * ob
    uint32_t outer_a_count;      // 0: single level
    uint64_t ld_outer2, st_outer2;
    // final pass only (k_ntt_rr<.., .., true, true>): the four-step NTT's exchange fused into the store.
    // out[k] *= osc_base^k (running product per thread), then element k goes to peer[k >> log_blk] + peer_row + (k & (blk - 1))
    uint32_t has_oscale;
    DevPow osc;
    uint32_t n_peers, log_blk;
    uint64_t peer_row;
    fe* peer[ZKB_NTT_MAX_PEERS];
};

__device__ __forceinline__ fe pow2lvl(const DevPow& t, uint64_t e) {
    fe lo = fe_ldg(t.lo + (e & ((1ull << t.lo_bits) - 1)));
    fe hi = fe_ldg(t.hi + (e >> t.lo_bits));
    return fe_montmul(hi, lo);
}

__device__ __forceinline__ uint32_t bitrev(uint32_t x, uint32_t bits) { return __brev(x) >> (32u - bits); }

#define ZKB_NTT_THREADS 256
// Which multiplications of the register-radix pass are INLINE (the rest go through the out-of-line mm / mm2): bit 0 the DFT stages,
// bit 1 the step-1 twiddles, bit 2 the step-3 running products.  An out-of-line call is a scheduling barrier and costs ~12 argument
// moves per product; too much inline code stalls on instruction fetch.  Measured on B200, 2^24 forward + inverse:
//   34-ALU-instruction multiplication (round 2a): 0: 2.598 ms (69 KB per kernel), 1: 2.423 (98 KB), 3: 2.380 (111 KB), 7: 2.432 (136 KB)
//   24-ALU-instruction multiplication (fe128.cuh now): 1: 2.245, 3: 2.178 (100 KB), 7: 2.154 (118 KB)
#ifndef ZKB_NTT_INLINE_DFT
#define ZKB_NTT_INLINE_DFT 7
#endif

// block -> (outer offset in, outer offset out, inner)
__device__ __forceinline__ void tile_origin(const PassParams& p, uint64_t& in_base, uint64_t& out_base, uint32_t& inner) {
    const uint32_t outer = blockIdx.x / p.inner_count;
    inner = blockIdx.x % p.inner_count;
    uint32_t oa = outer, ob = 0;
    if (p.outer_a_count) { oa = outer % p.outer_a_count; ob = outer / p.outer_a_count; }
    in_base = (uint64_t)oa * p.ld_outer + (uint64_t)ob * p.ld_outer2 + (uint64_t)inner * p.ld_inner;
    out_base = (uint64_t)oa * p.st_outer + (uint64_t)ob * p.st_outer2 + (uint64_t)inner * p.st_inner;
}

__global__ void __launch_bounds__(ZKB_NTT_THREADS) k_ntt_pass(const PassParams p) {
    extern __shared__ uint4 smem_raw[];
    fe* sm = reinterpret_cast<fe*>(smem_raw);
    const uint32_t S = 1u << p.log_s, B = 1u << p.log_b;
    const uint32_t tid = threadIdx.x;
    uint64_t in_base, out_base;
    uint32_t inner;
    tile_origin(p, in_base, out_base, inner);
    const fe* in = p.in + (uint64_t)blockIdx.y * p.in_batch;
    fe* out = p.out + (uint64_t)blockIdx.y * p.out_batch + out_base;
    fe* tws = sm + (p.transposed ? (size_t)B * (S + 1) : (size_t)S * B);
    for (uint32_t j = tid; j < (S >> 1); j += ZKB_NTT_THREADS) tws[j] = fe_ldg(p.tw_s + j);

    if (!p.transposed) {
        // ---- load: smem row q <- x[rev(q)], only rows q = qq << z have a source ---------
        const uint32_t z = p.skip_log, rows = S >> z;
        for (uint32_t idx = tid; idx < rows * B; idx += ZKB_NTT_THREADS) {
            uint32_t b = idx & (B - 1), qq = idx >> p.log_b;
            uint32_t s = (p.log_s == z) ? 0u : bitrev(qq, p.log_s - z);
            uint64_t lin = in_base + (uint64_t)s * p.ld_s + (uint64_t)b * p.ld_b;
            fe x = fe_zero();
            if (!p.has_valid || lin < p.n_valid) {
                x = fe_ldg(in + lin);
                if (p.has_scale) x = fe_montmul(x, pow2lvl(p.sc, lin));
            }
            uint32_t q0 = qq << z;
            for (uint32_t r = 0; r < (1u << z); r++) sm[(size_t)(q0 + r) * B + b] = x;
        }
        // ---- DIT stages -------------------------------------------------------------------
        for (uint32_t lh = z; lh < p.log_s; lh++) {
            __syncthreads();
            const uint32_t h = 1u << lh;
            for (uint32_t idx = tid; idx < (S >> 1) * B; idx += ZKB_NTT_THREADS) {
                uint32_t b = idx & (B - 1), pi = idx >> p.log_b;
                uint32_t j = pi & (h - 1);
                uint32_t s0 = ((pi >> lh) << (lh + 1)) | j;
                fe w = tws[j << (p.log_s - 1 - lh)];
                fe e = sm[(size_t)s0 * B + b];
                fe o = fe_montmul(sm[(size_t)(s0 + h) * B + b], w);
                sm[(size_t)s0 * B + b] = fe_add(e, o);
                sm[(size_t)(s0 + h) * B + b] = fe_sub(e, o);
            }
        }
        __syncthreads();
        // ---- store (natural order) with the inter-pass twiddle ----------------------------
        for (uint32_t idx = tid; idx < S * B; idx += ZKB_NTT_THREADS) {
            uint32_t b = idx & (B - 1), k = idx >> p.log_b;
            fe x = sm[(size_t)k * B + b];
            if (p.has_tw) {
                uint64_t col = (uint64_t)inner * B + b;
                x = fe_montmul(x, pow2lvl(p.tw, (uint64_t)k * col * p.tw_mul));
            }
            if (p.has_post) x = fe_montmul(x, p.post);
            fe_store(out + (uint64_t)k * p.st_k + (uint64_t)b * p.st_b, x);
        }
    } else {
        const uint32_t pitch = S + 1;      // odd pitch: conflict-free column reads in the store
        for (uint32_t idx = tid; idx < S * B; idx += ZKB_NTT_THREADS) {
            uint32_t s = idx & (S - 1), b = idx >> p.log_s;
            uint64_t lin = in_base + (uint64_t)s * p.ld_s + (uint64_t)b * p.ld_b;
            fe x = fe_zero();
            if (!p.has_valid || lin < p.n_valid) {
                x = fe_ldg(in + lin);
                if (p.has_scale) x = fe_montmul(x, pow2lvl(p.sc, lin));
            }
            sm[(size_t)b * pitch + s] = x;
        }
        // ---- DIF stages: natural in, bit-reversed out ------------------------------------
        for (int lh = (int)p.log_s - 1; lh >= 0; lh--) {
            __syncthreads();
            const uint32_t h = 1u << lh;
            for (uint32_t idx = tid; idx < (S >> 1) * B; idx += ZKB_NTT_THREADS) {
                uint32_t pi = idx & ((S >> 1) - 1), b = idx >> (p.log_s - 1);
                uint32_t j = pi & (h - 1);
                uint32_t s0 = ((pi >> lh) << (lh + 1)) | j;
                fe* row = sm + (size_t)b * pitch;
                fe u = row[s0], v = row[s0 + h];
                row[s0] = fe_add(u, v);
                row[s0 + h] = fe_montmul(fe_sub(u, v), tws[j << (p.log_s - 1 - lh)]);
            }
        }
        __syncthreads();
        for (uint32_t idx = tid; idx < S * B; idx += ZKB_NTT_THREADS) {
            uint32_t b = idx & (B - 1), k = idx >> p.log_b;
            fe x = sm[(size_t)b * pitch + bitrev(k, p.log_s)];
            if (p.has_tw) {
                uint64_t col = (uint64_t)inner * B + b;
                x = fe_montmul(x, pow2lvl(p.tw, (uint64_t)k * col * p.tw_mul));
            }
            if (p.has_post) x = fe_montmul(x, p.post);
            fe_store(out + (uint64_t)k * p.st_k + (uint64_t)b * p.st_b, x);
        }
    }
}


// ---- register-radix pass ------------------------------------------------------------------
// The same tile (S x B values) as k_ntt_pass, but the S = R1*R2-point transform of each column is
// done as two in-register radix-R passes (R <= 16, fully unrolled DIF, bit reversal = register
// renaming) around ONE shared-memory exchange instead of log2(S) shared-memory stages:
//   step 1: thread (s0, b) loads x[R2*s1 + s0], s1 < R1, straight from HBM, R1-point DFT,
//           multiplies by w_S^(s0*ka), writes Y[ka][s0][b] to shared memory
//   step 3: thread (ka, b) reads Y[ka][s0][b], s0 < R2, R2-point DFT -> X[ka + R1*kb], applies the
//           inter-pass twiddle (running product over kb) / n^-1 and stores to HBM.
// Trivial twiddles (w^0) cost nothing: 17 multiplications per 16-point DFT instead of 32.
// One out-of-line copy of the multiplication per kernel: the fully unrolled pass has ~130
// multiplication sites; inlined that is 190 KB of SASS and the pass stalls on instruction
// fetch (ncu: no_instruction is the top stall) - as calls it fits the instruction cache.
__device__ __noinline__ fe mm(fe a, fe b) { return fe_montmul(a, b); }
// Two independent products per call: the two carry chains interleave (ILP 2) and the call overhead halves.
struct fe2 { fe a, b; };
__device__ __noinline__ fe2 mm2(fe a0, fe b0, fe a1, fe b1) {
    fe2 r;
    r.a = fe_montmul(a0, b0);
    r.b = fe_montmul(a1, b1);
    return r;
}

__device__ __forceinline__ fe pow2lvl_c(const DevPow& t, uint64_t e) {
    fe lo = fe_ldg(t.lo + (e & ((1ull << t.lo_bits) - 1)));
    fe hi = fe_ldg(t.hi + (e >> t.lo_bits));
    return mm(hi, lo);
}
// position (in x[]) of the k-th butterfly of stage lh that needs a multiplication (j != 0), and its twiddle index
template <int LOGR, int LH> __host__ __device__ constexpr int mul_bfly(int k) {
    // butterflies i = 0 .. R/2-1 with j = i & (h-1) != 0, in order
    int seen = 0;
    for (int i = 0; i < (1 << (LOGR - 1)); i++) {
        if ((i & ((1 << LH) - 1)) == 0) continue;
        if (seen == k) return i;
        seen++;
    }
    return -1;
}
template <int LOGR, int LH>
__device__ __forceinline__ void dft_stage(fe (&x)[1 << LOGR], const fe* __restrict__ tw, uint32_t tw_stride) {
    constexpr int R = 1 << LOGR, h = 1 << LH;
#pragma unroll
    for (int i = 0; i < R / 2; i++) {
        const int j = i & (h - 1);
        const int s0 = ((i >> LH) << (LH + 1)) | j;
        fe u = x[s0], v = x[s0 + h];
        x[s0] = fe_add(u, v);            // inline: as calls the argument moves (17 per call) cost more than the 27 instructions saved
        x[s0 + h] = fe_sub(u, v);
    }
    constexpr int NM = R / 2 - R / (2 * h);                  // butterflies with j != 0
#if (ZKB_NTT_INLINE_DFT & 1)
    // the products of a DFT stage inline: the scheduler interleaves them with each other and with the neighbouring stage's
    // add / sub (an out-of-line call is a scheduling barrier and costs ~12 argument moves per product)
#pragma unroll
    for (int k = 0; k < NM; k++) {
        const int i0 = mul_bfly<LOGR, LH>(k);
        const int j0 = i0 & (h - 1);
        const int p0 = (((i0 >> LH) << (LH + 1)) | j0) + h;
        x[p0] = fe_montmul(x[p0], tw[(uint32_t)(j0 << (LOGR - 1 - LH)) * tw_stride]);
    }
#else
#pragma unroll
    for (int k = 0; k + 1 < NM; k += 2) {
        const int i0 = mul_bfly<LOGR, LH>(k), i1 = mul_bfly<LOGR, LH>(k + 1);
        const int j0 = i0 & (h - 1), j1 = i1 & (h - 1);
        const int p0 = (((i0 >> LH) << (LH + 1)) | j0) + h, p1 = (((i1 >> LH) << (LH + 1)) | j1) + h;
        fe2 r = mm2(x[p0], tw[(uint32_t)(j0 << (LOGR - 1 - LH)) * tw_stride], x[p1], tw[(uint32_t)(j1 << (LOGR - 1 - LH)) * tw_stride]);
        x[p0] = r.a; x[p1] = r.b;
    }
    if (NM & 1) {
        const int i0 = mul_bfly<LOGR, LH>(NM - 1);
        const int j0 = i0 & (h - 1);
        const int p0 = (((i0 >> LH) << (LH + 1)) | j0) + h;
        x[p0] = mm(x[p0], tw[(uint32_t)(j0 << (LOGR - 1 - LH)) * tw_stride]);
    }
#endif
}
template <int LOGR>
__device__ __forceinline__ void dft_dif(fe (&x)[1 << LOGR], const fe* __restrict__ tw, uint32_t tw_stride) {
    if constexpr (LOGR >= 4) dft_stage<LOGR, 3>(x, tw, tw_stride);
    if constexpr (LOGR >= 3) dft_stage<LOGR, 2>(x, tw, tw_stride);
    if constexpr (LOGR >= 2) dft_stage<LOGR, 1>(x, tw, tw_stride);
    dft_stage<LOGR, 0>(x, tw, tw_stride);
}
template <int LOGR> __host__ __device__ constexpr int brev_c(int i) {
    int r = 0;
    for (int k = 0; k < LOGR; k++) r |= ((i >> k) & 1) << (LOGR - 1 - k);
    return r;
}

template <int LR1, int LR2, bool TRANSPOSED, bool EXCHANGE = false, bool TWX = false>
__global__ void __launch_bounds__(ZKB_NTT_THREADS, 2) k_ntt_rr(const PassParams p) {
    extern __shared__ uint4 smem_raw[];
    constexpr uint32_t R1 = 1u << LR1, R2 = 1u << LR2, S = R1 * R2;
    fe* sm = reinterpret_cast<fe*>(smem_raw);
    const uint32_t B = 1u << p.log_b, pitch = B + 1;
    fe* tws = sm + (size_t)S * pitch;                      // w_S^j * R, j < S
    const uint32_t tid = threadIdx.x;
    uint64_t in_base, out_base;
    uint32_t inner;
    tile_origin(p, in_base, out_base, inner);
    const fe* in = p.in + (uint64_t)blockIdx.y * p.in_batch;
    fe* out = p.out + (uint64_t)blockIdx.y * p.out_batch + out_base;
    for (uint32_t j = tid; j < S; j += ZKB_NTT_THREADS) tws[j] = fe_ldg(p.tw_s + j);
    __syncthreads();
    // ---- step 1
    for (uint32_t d = tid; d < R2 * B; d += ZKB_NTT_THREADS) {
        const uint32_t s0 = TRANSPOSED ? (d & (R2 - 1)) : (d >> p.log_b);
        const uint32_t b = TRANSPOSED ? (d >> LR2) : (d & (B - 1));
        fe x[R1];
#pragma unroll
        for (uint32_t s1 = 0; s1 < R1; s1++) {
            const uint64_t lin = in_base + (uint64_t)(R2 * s1 + s0) * p.ld_s + (uint64_t)b * p.ld_b;
            x[s1] = fe_zero();
            if (!p.has_valid || lin < p.n_valid) {
                x[s1] = fe_ldg(in + lin);
                if (p.has_scale) x[s1] = mm(x[s1], pow2lvl_c(p.sc, lin));
            }
        }
        dft_dif<LR1>(x, tws, R2);
#pragma unroll
        for (uint32_t i = 0; i < R1; i++) {
            const uint32_t ka = brev_c<LR1>(i);
            fe y = x[i];
            const uint32_t e = s0 * ka;                    // < S
#if (ZKB_NTT_INLINE_DFT & 2)
            if (ka != 0 && s0 != 0) y = fe_montmul(y, tws[e]);
#else
            if (ka != 0 && s0 != 0) y = mm(y, tws[e]);
#endif
            sm[(size_t)(ka * R2 + s0) * pitch + b] = y;
        }
    }
    __syncthreads();
    // ---- step 3
    for (uint32_t d = tid; d < R1 * B; d += ZKB_NTT_THREADS) {
        const uint32_t ka = d >> p.log_b, b = d & (B - 1);
        fe x[R2];
#pragma unroll
        for (uint32_t s0 = 0; s0 < R2; s0++) x[s0] = sm[(size_t)(ka * R2 + s0) * pitch + b];
        dft_dif<LR2>(x, tws, R1);
        fe t, step;
        const fe* txp = nullptr;                               // TWX: this thread's column of the exact twiddle table
        if (TWX) txp = p.tw_x + (uint64_t)ka * p.tw_x_cols + ((uint64_t)inner * B + b);
        else if (p.has_tw) {
            const uint64_t col = (uint64_t)inner * B + b;
            t = pow2lvl_c(p.tw, (uint64_t)ka * col * p.tw_mul);
            step = pow2lvl_c(p.tw, (uint64_t)R1 * col * p.tw_mul);
        }
        fe ot, ostep;                                          // EXCHANGE: running osc_base^(output index)
        const uint64_t lin0 = out_base + (uint64_t)ka * p.st_k + (uint64_t)b * p.st_b;
        if (EXCHANGE && p.has_oscale) {
            ot = pow2lvl_c(p.osc, lin0);
            ostep = pow2lvl_c(p.osc, (uint64_t)R1 * p.st_k);
        }
#pragma unroll
        for (uint32_t kb = 0; kb < R2; kb++) {
            fe y = x[brev_c<LR2>(kb)];
            if (TWX) {
                // one look-up and one product per output instead of the running product's two (and the table carries n^-1)
                y = fe_montmul(y, fe_ldg(txp + (uint64_t)(R1 * kb) * p.tw_x_cols));
            } else if (p.has_tw) {
#if (ZKB_NTT_INLINE_DFT & 4)
                y = fe_montmul(y, t);
                if (kb + 1 < R2) t = fe_montmul(t, step);
#else
                if (kb + 1 < R2) { fe2 r = mm2(y, t, t, step); y = r.a; t = r.b; }
                else y = mm(y, t);
#endif
            }
            if (p.has_post) y = mm(y, p.post);
            if (EXCHANGE) {
                if (p.has_oscale) {
                    if (kb + 1 < R2) { fe2 r = mm2(y, ot, ot, ostep); y = r.a; ot = r.b; }
                    else y = mm(y, ot);
                }
                const uint64_t lin = lin0 + (uint64_t)(R1 * kb) * p.st_k;
                fe* dst = p.n_peers ? p.peer[lin >> p.log_blk] + p.peer_row + (lin & ((1ull << p.log_blk) - 1)) : p.out + lin;
                fe_store(dst, y);
            } else {
                fe_store(out + (uint64_t)(ka + R1 * kb) * p.st_k + (uint64_t)b * p.st_b, y);
            }
        }
    }
}

// ---- the LAST pass of an LDE fused with the layer-0 leaf hashing of the Merkle commitment ------------------------------------
// ntt_arithmetics.rs:161-170 (fast_coset_evaluate) -> merkle_root.rs:21-32 (commit), stark.rs:373-381 / fri.rs:136.  The pass is
// bound by the FMA-heavy pipe (IMAD.WIDE), the hashing by the ALU pipe; as two kernels they use one pipe at a time.  Here a CTA
// transforms its tile exactly like k_ntt_rr<.., .., true>, leaves the 4096 outputs in shared memory (and writes them to the codeword
// as before), and then every thread hashes two groups of 8 CONSECUTIVE outputs (a tile's outputs are runs of B >= 16 consecutive
// values) down to their level-3 nodes, as k_leaf8 does.  With two CTAs per SM in different phases, the multiplications of one
// overlap the hashing of the other, and the codeword is not read back from HBM for hashing.
template <int LR1, int LR2>
__global__ void __launch_bounds__(ZKB_NTT_THREADS, 2) k_ntt_rr_leaf(const PassParams p, uint8_t* __restrict__ out3) {
    extern __shared__ uint4 smem_raw[];
    constexpr uint32_t R1 = 1u << LR1, R2 = 1u << LR2, S = R1 * R2;
    constexpr uint32_t ITERS = 16 / R2;                     // (R1 * B) / 256 with S * B = 4096
    fe* sm = reinterpret_cast<fe*>(smem_raw);
    const uint32_t B = 1u << p.log_b, pitch = B + 1;
    fe* tws = sm + (size_t)S * pitch;                      // w_S^j * R, j < S
    const uint32_t tid = threadIdx.x;
    uint64_t in_base, out_base;
    uint32_t inner;
    tile_origin(p, in_base, out_base, inner);
    const fe* in = p.in;
    fe* out = p.out + out_base;
    for (uint32_t j = tid; j < S; j += ZKB_NTT_THREADS) tws[j] = fe_ldg(p.tw_s + j);
    __syncthreads();
    // ---- step 1 (as k_ntt_rr, transposed load)
    for (uint32_t d = tid; d < R2 * B; d += ZKB_NTT_THREADS) {
        const uint32_t s0 = d & (R2 - 1), b = d >> LR2;
        fe x[R1];
#pragma unroll
        for (uint32_t s1 = 0; s1 < R1; s1++) x[s1] = fe_ldg(in + in_base + (uint64_t)(R2 * s1 + s0) * p.ld_s + (uint64_t)b * p.ld_b);
        dft_dif<LR1>(x, tws, R2);
#pragma unroll
        for (uint32_t i = 0; i < R1; i++) {
            const uint32_t ka = brev_c<LR1>(i);
            fe y = x[i];
            if (ka != 0 && s0 != 0) y = mm(y, tws[s0 * ka]);
            sm[(size_t)(ka * R2 + s0) * pitch + b] = y;
        }
    }
    __syncthreads();
    // ---- step 3: all outputs of this thread stay in registers until every thread has read its inputs
    fe y[ITERS][R2];
#pragma unroll
    for (uint32_t it = 0; it < ITERS; it++) {
        const uint32_t d = tid + it * ZKB_NTT_THREADS;
        const uint32_t ka = d >> p.log_b, b = d & (B - 1);
        fe x[R2];
#pragma unroll
        for (uint32_t s0 = 0; s0 < R2; s0++) x[s0] = sm[(size_t)(ka * R2 + s0) * pitch + b];
        dft_dif<LR2>(x, tws, R1);
#pragma unroll
        for (uint32_t kb = 0; kb < R2; kb++) {
            fe v = x[brev_c<LR2>(kb)];
            if (p.has_post) v = mm(v, p.post);
            y[it][kb] = v;
        }
    }
    __syncthreads();
#pragma unroll
    for (uint32_t it = 0; it < ITERS; it++) {
        const uint32_t d = tid + it * ZKB_NTT_THREADS;
        const uint32_t ka = d >> p.log_b, b = d & (B - 1);
#pragma unroll
        for (uint32_t kb = 0; kb < R2; kb++) {
            const uint32_t k = ka + R1 * kb;
            fe_store(out + (uint64_t)k * p.st_k + b, y[it][kb]);     // the codeword (st_b == 1)
            sm[(size_t)k * pitch + b] = y[it][kb];                   // ... and the tile, run k = outputs [k * st_k, k * st_k + B)
        }
    }
    __syncthreads();
    // ---- leaf hashing: 512 groups of 8 consecutive outputs per tile, two per thread -> level-3 nodes
    const uint32_t gshift = p.log_b - 3;                             // groups per run = B / 8
#pragma unroll 1
    for (uint32_t g = tid; g < 512; g += ZKB_NTT_THREADS) {
        const uint32_t k = g >> gshift, h8 = (g & ((1u << gshift) - 1)) << 3;
        const fe* src = sm + (size_t)k * pitch + h8;
        uint64_t h[8];
        reduce8([&](int j, uint64_t* o) { fe v = src[j]; b2_leaf_call(&v, o); }, h);
        g_store_digest(out3, (out_base + (uint64_t)k * p.st_k + h8) >> 3, h);
    }
}

template <int LR1, int LR2>
static int launch_rr_t(zkb_ctx* c, const PassParams& p, uint32_t tiles, uint32_t batch) {
    const size_t S = (size_t)1 << (LR1 + LR2), B = (size_t)1 << p.log_b;
    const size_t smem = (S * (B + 1) + S) * sizeof(fe);
    dim3 grid(tiles, batch);
    LaunchScope ls(c, K_NTT_PASS);
    if (p.transposed && (p.has_oscale || p.n_peers)) k_ntt_rr<LR1, LR2, true, true><<<grid, ZKB_NTT_THREADS, smem, c->stream>>>(p);
    else if (p.transposed) k_ntt_rr<LR1, LR2, true><<<grid, ZKB_NTT_THREADS, smem, c->stream>>>(p);
    else if (p.tw_x) k_ntt_rr<LR1, LR2, false, false, true><<<grid, ZKB_NTT_THREADS, smem, c->stream>>>(p);
    else k_ntt_rr<LR1, LR2, false><<<grid, ZKB_NTT_THREADS, smem, c->stream>>>(p);
    return 0;
}
template <int LR1, int LR2>
static cudaError_t rr_attrs() {
    cudaError_t e = cudaFuncSetAttribute(k_ntt_rr<LR1, LR2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_ntt_rr<LR1, LR2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_ntt_rr<LR1, LR2, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_ntt_rr<LR1, LR2, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    return e;
}
// The dynamic shared-memory opt-in is a per-DEVICE attribute of a kernel: it is set for the context's device
// when the context is created (zkb_ctx_create, after cudaSetDevice), not behind a process-wide flag.
int ntt_device_init(zkb_ctx* c) {
    ZKB_CUDA(c, (rr_attrs<4, 4>()));
    ZKB_CUDA(c, (rr_attrs<4, 3>()));
    ZKB_CUDA(c, (rr_attrs<3, 3>()));
    ZKB_CUDA(c, (rr_attrs<3, 2>()));
    ZKB_CUDA(c, cudaFuncSetAttribute(k_ntt_rr_leaf<4, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    ZKB_CUDA(c, cudaFuncSetAttribute(k_ntt_rr_leaf<4, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    ZKB_CUDA(c, cudaFuncSetAttribute(k_ntt_rr_leaf<3, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    ZKB_CUDA(c, cudaFuncSetAttribute(k_ntt_pass, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    return 0;
}
template <int LR1, int LR2>
static int launch_rr_leaf_t(zkb_ctx* c, const PassParams& p, uint32_t tiles, uint8_t* out3) {
    const size_t S = (size_t)1 << (LR1 + LR2), B = (size_t)1 << p.log_b;
    const size_t smem = (S * (B + 1) + S) * sizeof(fe);
    LaunchScope ls(c, K_NTT_LEAF);
    k_ntt_rr_leaf<LR1, LR2><<<tiles, ZKB_NTT_THREADS, smem, c->stream>>>(p, out3);
    return 0;
}
// the final pass of a single transform + layer-0 leaf hashing (log_s in 6..8, tile = 4096 values, no exchange / batch)
static int launch_pass_rr_leaf(zkb_ctx* c, const PassParams& p, uint32_t tiles, uint8_t* out3) {
    int rc;
    switch (p.log_s) {
        case 8: rc = launch_rr_leaf_t<4, 4>(c, p, tiles, out3); break;
        case 7: rc = launch_rr_leaf_t<4, 3>(c, p, tiles, out3); break;
        case 6: rc = launch_rr_leaf_t<3, 3>(c, p, tiles, out3); break;
        default: return set_err(c, ZKB_ERR_ARG, "internal: fused final pass needs 6 <= log_s <= 8");
    }
    ZKB_TRY(rc);
    ZKB_CUDA(c, cudaGetLastError());
    return 0;
}
// register-radix pass for log_s in 5..8
static int launch_pass_rr(zkb_ctx* c, const PassParams& p, uint32_t tiles, uint32_t batch) {
    int rc;
    switch (p.log_s) {
        case 8: rc = launch_rr_t<4, 4>(c, p, tiles, batch); break;
        case 7: rc = launch_rr_t<4, 3>(c, p, tiles, batch); break;
        case 6: rc = launch_rr_t<3, 3>(c, p, tiles, batch); break;
        case 5: rc = launch_rr_t<3, 2>(c, p, tiles, batch); break;
        default: return set_err(c, ZKB_ERR_ARG, "internal: register-radix pass needs 5 <= log_s <= 8");
    }
    ZKB_TRY(rc);
    ZKB_CUDA(c, cudaGetLastError());
    return 0;
}

static size_t pass_smem_bytes(const PassParams& p) {
    size_t S = (size_t)1 << p.log_s, B = (size_t)1 << p.log_b;
    return ((p.transposed ? B * (S + 1) : S * B) + S / 2) * sizeof(fe);
}

static int launch_pass(zkb_ctx* c, const PassParams& p, uint32_t tiles, uint32_t batch) {
    dim3 grid(tiles, batch);
    { LaunchScope ls(c, K_NTT_PASS); k_ntt_pass<<<grid, ZKB_NTT_THREADS, pass_smem_bytes(p), c->stream>>>(p); }
    ZKB_CUDA(c, cudaGetLastError());
    return 0;
}

static const uint32_t TILE_LOG = 12;     // 4096 elements = 64 KiB per tile

static int tw_s_table(zkb_ctx* c, const fe& root, uint32_t log_n, uint32_t log_s, const fe** out) {
    // w_S = root^(N/S); table of w_S^j * R for j < S/2
    fe ws = root;
    for (uint32_t i = 0; i < log_n - log_s; i++) ws = h_mul(ws, ws);
    DevPow t;
    ZKB_TRY(get_pow_table(c, ws, log_s - 1, &t));
    *out = t.lo;
    return 0;
}

// ---- exact inter-pass twiddle tables ---------------------------------------------------------------------------------------------
// Pass i multiplies output k of column m by w_N^(k m N_0..N_{i-1}): N_i x M_i = N / (N_0..N_{i-1}) distinct factors.  For every pass but
// the first that is a small table (2^16 entries for the middle pass of a 2^24 transform, 1 MB, L2-resident): the pass then does ONE
// look-up and ONE product per output where the running product t *= step does two, and an inverse transform's n^-1 is folded into
// the table (no scaling pass, no scaling products).  Tables are cached per context (least recently used of 12 is evicted).
static const uint64_t TWX_MAX_ENTRIES = 1ull << 20;
struct TwExact { fe root; uint32_t log_n; uint64_t tw_mul, rows, cols; bool has_post; fe* d; uint64_t stamp; };
struct TwExactCache { std::vector<TwExact> v; uint64_t clock = 0; };
static void tw_exact_destroy(void* p) {
    TwExactCache* t = (TwExactCache*)p;
    for (auto& e : t->v) cudaFree(e.d);
    delete t;
}
__global__ void k_tw_exact(fe* __restrict__ out, uint64_t rows, uint64_t cols, uint64_t tw_mul, DevPow tw, uint32_t has_post, fe post) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * cols) return;
    const uint64_t k = i / cols, m = i - k * cols;
    fe t = pow2lvl(tw, k * m * tw_mul);
    if (has_post) t = fe_montmul(t, post);                  // post = n^-1 R: (w R)(n^-1 R) / R = w n^-1 R
    fe_store(out + i, t);
}
static int tw_exact_table(zkb_ctx* c, const fe& root, uint32_t log_n, uint64_t tw_mul, uint64_t rows, uint64_t cols, const DevPow& tw,
                          bool has_post, const fe& post, const fe** out) {
    TwExactCache* cache = nullptr;
    for (auto& at : c->attachments) if (at.destroy == tw_exact_destroy) cache = (TwExactCache*)at.p;
    if (!cache) { cache = new TwExactCache(); c->attachments.push_back({cache, tw_exact_destroy}); }
    cache->clock++;
    for (auto& e : cache->v)
        if (e.log_n == log_n && e.tw_mul == tw_mul && e.rows == rows && e.cols == cols && e.has_post == has_post && fe_eq(e.root, root)) {
            e.stamp = cache->clock; *out = e.d; return 0;
        }
    if (cache->v.size() >= 12) {
        size_t victim = 0;
        for (size_t i = 1; i < cache->v.size(); i++) if (cache->v[i].stamp < cache->v[victim].stamp) victim = i;
        ZKB_CUDA(c, cudaStreamSynchronize(c->stream));
        cudaFree(cache->v[victim].d);
        cache->v.erase(cache->v.begin() + victim);
    }
    TwExact e{root, log_n, tw_mul, rows, cols, has_post, nullptr, cache->clock};
    ZKB_CUDA(c, cudaMalloc(&e.d, sizeof(fe) * rows * cols));
    { LaunchScope ls(c, K_POW_TABLE); k_tw_exact<<<(unsigned)((rows * cols + 255) / 256), 256, 0, c->stream>>>(e.d, rows, cols, tw_mul, tw, has_post ? 1u : 0u, post); }
    cudaError_t le = cudaGetLastError();
    if (le != cudaSuccess) { cudaFree(e.d); return set_err(c, ZKB_ERR_CUDA, "k_tw_exact failed: %s", cudaGetErrorString(le)); }
    cache->v.push_back(e);
    *out = e.d;
    return 0;
}

int ntt_exec(zkb_ctx* c, fe root, const fe* d_in, size_t n_in, size_t in_stride, fe* d_out,
             size_t out_stride, size_t batch, uint32_t log_n, const NttOpts& o) {
    if (batch == 0) return 0;
    const uint64_t N = 1ull << log_n;
    if (log_n == 0) {
        // length-1 transform: identity (the LDE scaling offset^0 = 1 as well)
        for (size_t bi = 0; bi < batch; bi++)
            ZKB_CUDA(c, cudaMemcpyAsync(d_out + bi * out_stride, d_in + bi * in_stride, sizeof(fe),
                                        cudaMemcpyDeviceToDevice, c->stream));
        return 0;
    }
    if (log_n > 31) return set_err(c, ZKB_ERR_ARG, "ntt: length 2^%u not supported on one GPU", log_n);
    if (o.inverse) root = h_inv(root);
    PassParams base;
    memset(&base, 0, sizeof(base));
    if (o.inverse && !o.no_post) {
        base.has_post = 1;
        base.post = fe_to_mont(h_inv(h_from_u64(N)));
    }
    DevPow sc{};
    if (o.has_scale) ZKB_TRY(get_pow_table(c, o.scale_base, log_n, &sc));

    // Pass plan: N = N_0 * ... * N_{P-1}.  2^13..2^32: every pass is 5..8 bits wide and runs in the register-radix
    // kernel (two passes up to 2^16, three up to 2^24, four above: Bailey's split applied P - 1 times).
    uint32_t wdt[4] = {0, 0, 0, 0};
    int passes;
    const bool fast = log_n > TILE_LOG;
    if (!fast) { passes = 1; wdt[0] = log_n; }
    else {
        passes = log_n <= 16 ? 2 : log_n <= 24 ? 3 : 4;
        uint32_t left = log_n;
        for (int i = 0; i < passes; i++) { wdt[i] = (left + (passes - i) - 1) / (passes - i); left -= wdt[i]; }
    }
    if (o.exchange && (!fast || batch != 1)) return set_err(c, ZKB_ERR_ARG, "internal: fused exchange needs one transform of at least 2^%u values", TILE_LOG + 1);
    auto tw_table = [&](uint32_t log_s, const fe** out) -> int {
        fe ws = root;                                        // full table w_S^j * R, j < S
        for (uint32_t i = 0; i < log_n - log_s; i++) ws = h_mul(ws, ws);
        DevPow t;
        ZKB_TRY(get_pow_table(c, ws, log_s, &t));
        *out = t.lo;
        return 0;
    };
    if (passes == 1) {
        // One tile: the literal radix-2 DIT of ntt.rs:26-46 (bit-reversed load, twiddles
        // root^k from one table, explicit e - o*w) - op for op the reference's loop, so the
        // result matches the reference even for a `root` that is not a primitive n-th root
        // (ntt.rs does not check; fast_multiply's order-shrink quirk relies on it).
        PassParams p = base;
        p.in = d_in; p.out = d_out;
        p.log_s = log_n; p.log_b = 0; p.transposed = 0; p.inner_count = 1;
        p.ld_s = 1; p.ld_b = 0; p.st_k = 1; p.st_b = 0;
        p.in_batch = in_stride; p.out_batch = out_stride;
        p.has_valid = n_in < N; p.n_valid = n_in;
        p.skip_log = log_n - ilog2_u64(n_in ? n_in : 1);       // x[i] = 0 for i >= 2^ceil(log2 n_in)
        p.has_scale = o.has_scale; p.sc = sc;
        ZKB_TRY(tw_s_table(c, root, log_n, log_n, &p.tw_s));
        return launch_pass(c, p, 1, (uint32_t)batch);
    }
    DevPow tw;
    ZKB_TRY(get_pow_table(c, root, log_n, &tw));
    fe* A = nullptr;
    ZKB_TRY(scratch_reserve(c, sizeof(fe) * N * batch, (void**)&A));
    // M[i] = N / (N_0 ... N_i): stride of digit n_i in the working layout A[k_0 M_0 + ... + k_i M_i + m]
    uint64_t M[4], Nprod[5];
    Nprod[0] = 1;
    for (int i = 0; i < passes; i++) { Nprod[i + 1] = Nprod[i] << wdt[i]; M[i] = N >> (ilog2_u64(Nprod[i + 1])); }
    // the LAST pass with a small exact twiddle table also applies an inverse transform's n^-1 (folded into the table)
    int post_pass = -1;
    if (base.has_post && !getenv("ZKB_NTT_NO_TWX"))
        for (int i = 0; i + 1 < passes; i++) if ((N >> ilog2_u64(Nprod[i])) <= TWX_MAX_ENTRIES) post_pass = i;
    for (int i = 0; i + 1 < passes; i++) {
        // pass i: N_i-point transforms over digit n_i (stride M_i) for every prefix (k_0 .. k_{i-1}) and column m < M_i,
        // B adjacent columns per tile; then the twiddle w_{M_{i-1}}^(k_i m) = w_N^(k_i m N_0...N_{i-1})
        PassParams p = base;
        p.has_post = 0;
        p.in = i == 0 ? d_in : A; p.out = A;
        p.log_s = wdt[i];
        uint32_t lb = TILE_LOG - wdt[i];
        if ((1ull << lb) > M[i]) lb = ilog2_u64(M[i]);
        p.log_b = lb; p.transposed = 0;
        p.inner_count = (uint32_t)(M[i] >> lb);
        p.ld_s = M[i]; p.ld_b = 1; p.ld_inner = 1ull << lb;
        p.st_k = M[i]; p.st_b = 1; p.st_inner = 1ull << lb;
        p.ld_outer = p.st_outer = i == 0 ? 0 : M[i - 1];
        p.in_batch = i == 0 ? in_stride : N; p.out_batch = N;
        if (i == 0) {
            p.has_valid = n_in < N; p.n_valid = n_in;
            uint64_t rows = (n_in + M[0] - 1) / M[0];      // rows n_0 that hold any data
            if (rows == 0) rows = 1;
            p.skip_log = wdt[0] - ilog2_u64(rows);          // ceil log2
            p.has_scale = o.has_scale; p.sc = sc;
        }
        p.has_tw = 1; p.tw = tw; p.tw_mul = Nprod[i];
        if ((N >> ilog2_u64(Nprod[i])) <= TWX_MAX_ENTRIES && !getenv("ZKB_NTT_NO_TWX")) {
            p.tw_x_cols = M[i];
            ZKB_TRY(tw_exact_table(c, root, log_n, Nprod[i], 1ull << wdt[i], M[i], tw, i == post_pass, base.post, &p.tw_x));
        }
        ZKB_TRY(tw_table(wdt[i], &p.tw_s));
        ZKB_TRY(launch_pass_rr(c, p, (uint32_t)(Nprod[i] * p.inner_count), (uint32_t)batch));
    }
    {   // final pass: N_{P-1}-point transforms along contiguous runs; tile = B consecutive k_0 (stride M_0 in, contiguous out),
        // outer = (k_1 .. k_{P-2}); transposing store X[k_0 + N_0 k_1 + N_0 N_1 k_2 + ...]
        const int f = passes - 1;
        PassParams p = base;
        if (post_pass >= 0) p.has_post = 0;                      // n^-1 already applied by pass `post_pass`
        p.in = A; p.out = d_out;
        p.log_s = wdt[f];
        uint32_t lb = TILE_LOG - wdt[f];
        if (lb > wdt[0]) lb = wdt[0];
        p.log_b = lb; p.transposed = 1;
        p.inner_count = (uint32_t)(Nprod[1] >> lb);
        p.ld_s = 1; p.ld_b = M[0]; p.ld_inner = (1ull << lb) * M[0];
        p.st_k = Nprod[f]; p.st_b = 1; p.st_inner = 1ull << lb;
        uint64_t outer = 1;
        if (passes == 3) { outer = 1ull << wdt[1]; p.ld_outer = M[1]; p.st_outer = Nprod[1]; }
        if (passes == 4) {
            outer = 1ull << (wdt[1] + wdt[2]);
            p.outer_a_count = 1u << wdt[1];
            p.ld_outer = M[1]; p.st_outer = Nprod[1];          // k_1
            p.ld_outer2 = M[2]; p.st_outer2 = Nprod[2];        // k_2
        }
        p.in_batch = N; p.out_batch = out_stride;
        if (o.exchange) {
            const NttExchange& x = *o.exchange;
            p.has_oscale = 1;
            ZKB_TRY(get_pow_table(c, x.oscale_base, log_n, &p.osc));
            p.n_peers = x.n_peers; p.log_blk = x.log_blk; p.peer_row = x.peer_row;
            for (uint32_t q = 0; q < x.n_peers && q < ZKB_NTT_MAX_PEERS; q++) p.peer[q] = x.peer[q];
        }
        ZKB_TRY(tw_table(wdt[f], &p.tw_s));
        if (o.leaf3_out && batch == 1 && !o.exchange && wdt[f] >= 6 && wdt[f] + p.log_b == TILE_LOG && p.log_b >= 4)
            ZKB_TRY(launch_pass_rr_leaf(c, p, (uint32_t)(outer * p.inner_count), o.leaf3_out));
        else if (o.leaf3_out)
            return set_err(c, ZKB_ERR_ARG, "internal: this transform shape has no fused leaf-hashing pass (ntt_can_fuse_leaves)");
        else
            ZKB_TRY(launch_pass_rr(c, p, (uint32_t)(outer * p.inner_count), (uint32_t)batch));
    }
    return 0;
}

// does ntt_exec's final pass for 2^log_n values come in the fused-with-leaf-hashing variant?
bool ntt_can_fuse_leaves(uint32_t log_n) {
    if (log_n <= TILE_LOG + 5) return false;                              // (trees of <= 2^17 leaves are hashed by the latency kernels anyway)
    const int passes = log_n <= 16 ? 2 : log_n <= 24 ? 3 : 4;
    uint32_t left = log_n, w0 = 0, wl = 0;
    for (int i = 0; i < passes; i++) { uint32_t w = (left + (passes - i) - 1) / (passes - i); if (i == 0) w0 = w; wl = w; left -= w; }
    const uint32_t lb = TILE_LOG - wl;
    return wl >= 6 && wl <= 8 && lb <= w0 && lb >= 4;
}

// ---- cross-rank stage of the four-step NTT: `blk` interleaved G-point transforms, element n1 of column i at in[n1 * blk + i],
// output k1 at out[k1 * blk + i].  One thread per column, the whole transform in registers (G <= 16); HBM-bound.
template <int LOGG>
__global__ void __launch_bounds__(256) k_ntt_cross(const fe* __restrict__ in, fe* __restrict__ out, uint64_t blk, const fe* __restrict__ tw_g,
                                                   uint32_t has_post, fe post) {
    constexpr uint32_t G = 1u << LOGG;
    __shared__ fe tws[G];
    if (threadIdx.x < G) tws[threadIdx.x] = fe_ldg(tw_g + threadIdx.x);
    __syncthreads();
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= blk) return;
    fe x[G];
#pragma unroll
    for (uint32_t n1 = 0; n1 < G; n1++) x[n1] = fe_ldg(in + (uint64_t)n1 * blk + i);
    dft_dif<LOGG>(x, tws, 1);
#pragma unroll
    for (uint32_t k1 = 0; k1 < G; k1++) {
        fe y = x[brev_c<LOGG>(k1)];
        if (has_post) y = mm(y, post);
        fe_store(out + (uint64_t)k1 * blk + i, y);
    }
}

// root_g: primitive 2^log_g-th root (canonical); post: nullptr or a Montgomery-form factor applied to every output
int ntt_cross_exec(zkb_ctx* c, const fe& root_g, uint32_t log_g, const fe* d_in, fe* d_out, uint64_t blk, const fe* post) {
    if (log_g < 1 || log_g > 4) return set_err(c, ZKB_ERR_ARG, "four-step NTT: 2, 4, 8 or 16 ranks");
    DevPow t;
    ZKB_TRY(get_pow_table(c, root_g, log_g, &t));
    const unsigned blocks = (unsigned)((blk + 255) / 256);
    const fe pf = post ? *post : fe_zero();
    LaunchScope ls(c, K_NTT_PASS);
    switch (log_g) {
        case 1: k_ntt_cross<1><<<blocks, 256, 0, c->stream>>>(d_in, d_out, blk, t.lo, post != nullptr, pf); break;
        case 2: k_ntt_cross<2><<<blocks, 256, 0, c->stream>>>(d_in, d_out, blk, t.lo, post != nullptr, pf); break;
        case 3: k_ntt_cross<3><<<blocks, 256, 0, c->stream>>>(d_in, d_out, blk, t.lo, post != nullptr, pf); break;
        default: k_ntt_cross<4><<<blocks, 256, 0, c->stream>>>(d_in, d_out, blk, t.lo, post != nullptr, pf); break;
    }
    ZKB_CUDA(c, cudaGetLastError());
    return 0;
}

// out[k] = in[k] * base^k scattered like NttExchange (the unfused fallback for local transforms below 2^13)
__global__ void k_twiddle_scatter(const fe* __restrict__ in, uint64_t n, DevPow sc, PassParams p) {
    const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const fe y = fe_montmul(fe_ldg(in + k), pow2lvl(sc, k));
    fe_store(p.peer[k >> p.log_blk] + p.peer_row + (k & ((1ull << p.log_blk) - 1)), y);
}
int ntt_twiddle_scatter(zkb_ctx* c, const fe* d_in, uint64_t n, const NttExchange& x) {
    PassParams p;
    memset(&p, 0, sizeof(p));
    p.n_peers = x.n_peers; p.log_blk = x.log_blk; p.peer_row = x.peer_row;
    for (uint32_t q = 0; q < x.n_peers && q < ZKB_NTT_MAX_PEERS; q++) p.peer[q] = x.peer[q];
    DevPow sc;
    ZKB_TRY(get_pow_table(c, x.oscale_base, ilog2_u64(n), &sc));
    { LaunchScope ls(c, K_ELEMENTWISE); k_twiddle_scatter<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(d_in, n, sc, p); }
    ZKB_CUDA(c, cudaGetLastError());
    return 0;
}

// ---- small elementwise kernels ------------------------------------------------------------
__global__ void k_scale(const fe* in, fe* out, uint64_t n, DevPow sc) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    fe_store(out + i, fe_montmul(fe_ldg(in + i), pow2lvl(sc, i)));
}
__global__ void k_pointwise_mul(const fe* a, const fe* b, fe* out, uint64_t n) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    fe_store(out + i, fe_montmul(fe_to_mont(fe_ldg(a + i)), fe_ldg(b + i)));
}
// out = a / b pointwise; flags[0] set if any b[i] == 0 (field_element.rs:85 panics there)
__global__ void k_pointwise_div(const fe* a, const fe* b, fe* out, uint64_t n, uint32_t* flag) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    fe d = fe_ldg(b + i);
    if (fe_is_zero(d)) { atomicOr(flag, 1u); fe_store(out + i, fe_zero()); return; }
    // d^(p-2) in Montgomery form, p-2 = 0xCB7FFFFF FFFFFFFF FFFFFFFF FFFFFFFF
    fe base = fe_to_mont(d), acc = fe_mont_one();
    for (int bit = 127; bit >= 0; bit--) {
        acc = fe_montmul(acc, acc);
        uint32_t word = (bit >= 96) ? 0xCB7FFFFFu : 0xFFFFFFFFu;
        if ((word >> (bit & 31)) & 1u) acc = fe_montmul(acc, base);
    }
    fe_store(out + i, fe_montmul(acc, fe_ldg(a + i)));     // (1/d)*R * a / R
}

// ---- the reference's radix-2 loop, literally, for ANY length (ntt.rs:26-46: bit-reversed copy, then one stage per launch with
// explicit e + o*w / e - o*w and twiddles root^(j * n / (2h)) from a power table of `root`).  Only fast_multiply / fast_coset_divide's
// trailing-zero quirk needs it above one tile: there ntt() runs at the operand's padded length with a root of SMALLER order, for
// which the loop is not a DFT and no factorisation applies.  Global-memory passes: slow, correct, rare.
__global__ void k_bitrev_copy(const fe* __restrict__ in, fe* __restrict__ out, uint32_t log_n) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >> log_n) return;
    const uint64_t r = __brevll(i) >> (64 - log_n);
    fe_store(out + r, fe_ldg(in + i));
}
__global__ void k_dit_stage(fe* data, uint32_t log_n, uint32_t lh, DevPow pw) {
    const uint64_t pi = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (pi >> (log_n - 1)) return;
    const uint64_t h = 1ull << lh, j = pi & (h - 1);
    const uint64_t s0 = ((pi >> lh) << (lh + 1)) | j;
    const fe w = pow2lvl(pw, j << (log_n - 1 - lh));                  // root^(j * n / (2h)) * R
    const fe e = fe_load(data + s0), o = fe_montmul(fe_load(data + s0 + h), w);
    fe_store(data + s0, fe_add(e, o));
    fe_store(data + s0 + h, fe_sub(e, o));
}
// powers root^i * R for i < n with root of any order: the two-level table only needs exponent arithmetic, not root^n == 1
static int ntt_literal_big(zkb_ctx* c, const fe& root, const fe* d_in, fe* d_out, uint32_t log_n) {
    DevPow pw;
    ZKB_TRY(get_pow_table(c, root, log_n, &pw));
    const uint64_t n = 1ull << log_n;
    { LaunchScope ls(c, K_ELEMENTWISE); k_bitrev_copy<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(d_in, d_out, log_n); }
    for (uint32_t lh = 0; lh < log_n; lh++) {
        LaunchScope ls(c, K_NTT_PASS);
        k_dit_stage<<<(unsigned)((n / 2 + 255) / 256), 256, 0, c->stream>>>(d_out, log_n, lh, pw);
    }
    ZKB_CUDA(c, cudaGetLastError());
    return 0;
}

static int poly_degree(const fe* p, size_t n) {   // polynomial.rs:46-63; -1 for the zero polynomial
    int d = -1;
    for (size_t i = 0; i < n; i++) if (!fe_is_zero(p[i])) d = (int)i;
    return d;
}

static int check_root(zkb_ctx* c, const fe& root, uint64_t order) {
    // ntt_arithmetics.rs:11-24
    if (order == 0 || (order & (order - 1))) return set_err(c, ZKB_ERR_ROOT_ORDER, "root_order %llu is not a power of two", (unsigned long long)order);
    fe one = fe_from_u32(1);
    if (!fe_eq(h_pow(root, order), one))
        return set_err(c, ZKB_ERR_ROOT_ORDER, "supplied root does not have supplied root_order %llu", (unsigned long long)order);
    if (order > 1 && fe_eq(h_pow(root, order / 2), one))
        return set_err(c, ZKB_ERR_ROOT_ORDER, "supplied root is not a primitive of root_order %llu", (unsigned long long)order);
    return 0;
}

}  // namespace zkb

using namespace zkb;

extern "C" {

int zkb_ntt_batch(zkb_ctx* c, const uint8_t root[16], int inverse, const void* in, size_t n_in,
                  size_t in_stride, void* out, size_t out_stride, size_t batch) {
    if (!c || !root) return ZKB_ERR_ARG;
    if (n_in == 0) return set_err(c, ZKB_ERR_EMPTY, "ntt: empty input (ntt.rs:11 indexes inputs[0])");   // an empty Vec's pointer may be null
    if (!in || !out) return set_err(c, ZKB_ERR_ARG, "ntt: null data pointer");
    if (batch == 0) return 0;
    ZKB_CUDA(c, cudaSetDevice(c->device));
    const uint64_t n = next_pow2_u64(n_in);
    const uint32_t log_n = ilog2_u64(n);
    if (batch > 1 && (in_stride < n_in || out_stride < n)) return set_err(c, ZKB_ERR_ARG, "ntt_batch: strides shorter than the columns");
    const bool out_dev = is_device_ptr(out);
    DevBuf bin, bout;
    const void* d_in = nullptr;
    size_t in_elems = batch > 1 ? in_stride * (batch - 1) + n_in : n_in;
    size_t out_elems = batch > 1 ? out_stride * (batch - 1) + n : n;
    ZKB_TRY(stage_in(c, in, in_elems * sizeof(fe), bin, &d_in));
    fe* d_out = (fe*)out;
    if (!out_dev) { ZKB_TRY(bout.alloc(c, out_elems * sizeof(fe))); d_out = (fe*)bout.p; }
    if (n_in < 2) {
        // ntt.rs: bit_reverse_copy returns a 1-element input unchanged; intt returns it as is
        for (size_t bi = 0; bi < batch; bi++)
            ZKB_CUDA(c, cudaMemcpyAsync(d_out + bi * out_stride, (const fe*)d_in + bi * in_stride, sizeof(fe), cudaMemcpyDeviceToDevice, c->stream));
    } else {
        NttOpts o;
        o.inverse = inverse != 0;
        ZKB_TRY(ntt_exec(c, h_load(root), (const fe*)d_in, n_in, in_stride, d_out, out_stride, batch, log_n, o));
    }
    if (!out_dev) ZKB_CUDA(c, cudaMemcpyAsync(out, d_out, out_elems * sizeof(fe), cudaMemcpyDeviceToHost, c->stream));
    if (!out_dev || bin.p) ZKB_CUDA(c, cudaStreamSynchronize(c->stream));
    return 0;
}

// `count` interleaved transforms of length n (power of two, <= 4096): element j of sequence q is
// in[j * stride + q] (stride >= count).  Same layout out.  This is the cross-GPU stage of the
// four-step NTT (n = number of GPUs) and runs as one strided tile pass.
int zkb_ntt_strided(zkb_ctx* c, const uint8_t root[16], int inverse, const void* in, size_t n, size_t stride,
                    size_t count, void* out) {
    if (!c || !root || !in || !out) return ZKB_ERR_ARG;
    if (n == 0 || (n & (n - 1)) || n > (1u << TILE_LOG)) return set_err(c, ZKB_ERR_ARG, "ntt_strided: length %zu must be a power of two <= 4096", n);
    if (stride < count) return set_err(c, ZKB_ERR_ARG, "ntt_strided: stride shorter than the sequence count");
    if (!is_device_ptr(in) || !is_device_ptr(out)) return set_err(c, ZKB_ERR_ARG, "ntt_strided takes device pointers");
    if (count == 0) return 0;
    ZKB_CUDA(c, cudaSetDevice(c->device));
    if (n == 1) {
        if (in != out) ZKB_CUDA(c, cudaMemcpyAsync(out, in, count * sizeof(fe), cudaMemcpyDeviceToDevice, c->stream));
        return 0;
    }
    const uint32_t log_n = ilog2_u64(n);
    fe r = h_load(root);
    PassParams p;
    memset(&p, 0, sizeof(p));
    if (inverse) {
        r = h_inv(r);
        p.has_post = 1;
        p.post = fe_to_mont(h_inv(h_from_u64(n)));
    }
    p.in = (const fe*)in; p.out = (fe*)out;
    p.log_s = log_n; p.transposed = 0;
    p.ld_s = stride; p.ld_b = 1; p.st_k = stride; p.st_b = 1;
    ZKB_TRY(tw_s_table(c, r, log_n, log_n, &p.tw_s));
    // columns are processed in tiles of B = 2^log_b (a ragged tail gets narrower tiles)
    size_t done = 0;
    while (done < count) {
        uint32_t lb = TILE_LOG - log_n;
        while (lb > 0 && ((size_t)1 << lb) > count - done) lb--;
        const size_t B = (size_t)1 << lb, tiles = (count - done) >> lb;
        PassParams q = p;
        q.in = p.in + done; q.out = p.out + done;
        q.log_b = lb; q.inner_count = (uint32_t)tiles;
        q.ld_inner = B; q.st_inner = B;
        ZKB_TRY(launch_pass(c, q, (uint32_t)tiles, 1));
        done += tiles * B;
    }
    return 0;
}

int zkb_ntt(zkb_ctx* c, const uint8_t root[16], const void* in, size_t n_in, void* out) {
    return zkb_ntt_batch(c, root, 0, in, n_in, 0, out, 0, 1);
}
int zkb_intt(zkb_ctx* c, const uint8_t root[16], const void* in, size_t n_in, void* out) {
    return zkb_ntt_batch(c, root, 1, in, n_in, 0, out, 0, 1);
}

int zkb_poly_scale(zkb_ctx* c, const uint8_t factor[16], const void* coeffs, size_t n, void* out) {
    if (!c || !factor || (n && (!coeffs || !out))) return ZKB_ERR_ARG;
    if (n == 0) return 0;
    ZKB_CUDA(c, cudaSetDevice(c->device));
    const bool out_dev = is_device_ptr(out);
    DevBuf bin, bout;
    const void* d_in = nullptr;
    ZKB_TRY(stage_in(c, coeffs, n * sizeof(fe), bin, &d_in));
    fe* d_out = (fe*)out;
    if (!out_dev) { ZKB_TRY(bout.alloc(c, n * sizeof(fe))); d_out = (fe*)bout.p; }
    DevPow sc;
    ZKB_TRY(get_pow_table(c, h_load(factor), ilog2_u64(next_pow2_u64(n)), &sc));
    { LaunchScope ls(c, K_ELEMENTWISE); k_scale<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>((const fe*)d_in, d_out, n, sc); }
    ZKB_CUDA(c, cudaGetLastError());
    if (!out_dev) ZKB_CUDA(c, cudaMemcpyAsync(out, d_out, n * sizeof(fe), cudaMemcpyDeviceToHost, c->stream));
    if (!out_dev || bin.p) ZKB_CUDA(c, cudaStreamSynchronize(c->stream));
    return 0;
}

int zkb_coset_lde_batch(zkb_ctx* c, const uint8_t omega[16], uint64_t order, const uint8_t offset[16],
                        const void* coeffs, size_t n_coeffs, size_t in_stride, void* out,
                        size_t out_stride, size_t batch) {
    if (!c || !omega || !offset || !out) return ZKB_ERR_ARG;
    if (order == 0 || (order & (order - 1))) return set_err(c, ZKB_ERR_ARG, "coset_lde: order %llu is not a power of two", (unsigned long long)order);
    if (n_coeffs > order) return set_err(c, ZKB_ERR_TOO_LONG, "coset_lde: %zu coefficients exceed root_order %llu", n_coeffs, (unsigned long long)order);
    if (batch == 0) return 0;
    if (batch > 1 && (in_stride < n_coeffs || out_stride < order)) return set_err(c, ZKB_ERR_ARG, "coset_lde_batch: strides shorter than the columns");
    ZKB_CUDA(c, cudaSetDevice(c->device));
    const bool out_dev = is_device_ptr(out);
    size_t out_elems = batch > 1 ? out_stride * (batch - 1) + order : order;
    DevBuf bin, bout;
    fe* d_out = (fe*)out;
    if (!out_dev) { ZKB_TRY(bout.alloc(c, out_elems * sizeof(fe))); d_out = (fe*)bout.p; }
    if (n_coeffs == 0) {
        for (size_t bi = 0; bi < batch; bi++)
            ZKB_CUDA(c, cudaMemsetAsync(d_out + bi * out_stride, 0, order * sizeof(fe), c->stream));
    } else {
        if (!coeffs) return ZKB_ERR_ARG;
        const void* d_in = nullptr;
        size_t in_elems = batch > 1 ? in_stride * (batch - 1) + n_coeffs : n_coeffs;
        ZKB_TRY(stage_in_once(c, coeffs, in_elems * sizeof(fe), bin, &d_in));   // read once, by the first pass
        NttOpts o;
        o.has_scale = true;
        o.scale_base = h_load(offset);
        ZKB_TRY(ntt_exec(c, h_load(omega), (const fe*)d_in, n_coeffs, in_stride, d_out, out_stride, batch, ilog2_u64(order), o));
    }
    if (!out_dev) ZKB_CUDA(c, cudaMemcpyAsync(out, d_out, out_elems * sizeof(fe), cudaMemcpyDeviceToHost, c->stream));
    if (!out_dev || (coeffs && !is_device_ptr(coeffs))) ZKB_CUDA(c, cudaStreamSynchronize(c->stream));
    return 0;
}

int zkb_coset_lde(zkb_ctx* c, const uint8_t omega[16], uint64_t order, const uint8_t offset[16],
                  const void* coeffs, size_t n_coeffs, void* out) {
    return zkb_coset_lde_batch(c, omega, order, offset, coeffs, n_coeffs, 0, out, 0, 1);
}

// fast_multiply / fast_coset_divide: operands are small host polynomials in the reference
// (<= the omicron domain); degrees are taken on the host, the transforms run on the device.
static int poly_binop(zkb_ctx* c, bool divide, const uint8_t root_b[16], uint64_t root_order,
                      const uint8_t* offset_b, const void* lhs, size_t n_lhs, const void* rhs,
                      size_t n_rhs, void* out, size_t* n_out) {
    if (!c || !root_b || !n_out || !out) return ZKB_ERR_ARG;
    ZKB_CUDA(c, cudaSetDevice(c->device));
    // the degrees are taken on the host (polynomial.rs:46-63 scans the coefficients): device operands are brought over first
    std::vector<fe> hl, hr, hout;
    void* out_dev = nullptr;
    if (lhs && n_lhs && is_device_ptr(lhs)) { hl.resize(n_lhs); ZKB_CUDA(c, cudaMemcpyAsync(hl.data(), lhs, n_lhs * sizeof(fe), cudaMemcpyDeviceToHost, c->stream)); lhs = hl.data(); }
    if (rhs && n_rhs && is_device_ptr(rhs)) { hr.resize(n_rhs); ZKB_CUDA(c, cudaMemcpyAsync(hr.data(), rhs, n_rhs * sizeof(fe), cudaMemcpyDeviceToHost, c->stream)); rhs = hr.data(); }
    if (!hl.empty() || !hr.empty()) ZKB_CUDA(c, cudaStreamSynchronize(c->stream));
    if (is_device_ptr(out)) { out_dev = out; hout.resize(n_lhs + n_rhs + 1); out = hout.data(); }
    fe root = h_load(root_b);
    ZKB_TRY(check_root(c, root, root_order));
    const fe* L = (const fe*)lhs; const fe* R = (const fe*)rhs;
    int dl = poly_degree(L, n_lhs), dr = poly_degree(R, n_rhs);
    uint64_t degree; size_t result_len;
    if (!divide) {
        if (dl < 0 || dr < 0) { *n_out = 0; return 0; }               // :26-28
        degree = (uint64_t)dl + (uint64_t)dr; result_len = degree + 1;
    } else {
        if (dr < 0) return set_err(c, ZKB_ERR_DIV_ZERO, "cannot divide by zero polynomial");      // :258
        if (dl < 0) { *n_out = 0; return 0; }                         // :260-262
        if (dl < dr) return set_err(c, ZKB_ERR_DEGREE, "cannot divide by polynomial of larger degree");
        degree = (uint64_t)dl; result_len = (size_t)(dl - dr + 1);
    }
    uint64_t order = root_order;
    while (degree < order / 2) { root = h_mul(root, root); order /= 2; }   // :38-41 / :278-281
    // Operands longer than the shrunk order (trailing zero coefficients): the reference's
    // `extend(len..order)` is a no-op and ntt() pads to ITS next power of two while keeping the
    // order-`order` root (:43-47); only the first `order` outputs are used.  Reproduced
    // literally (single-tile DIT), which needs the padded operand to fit one tile.
    const uint64_t nL = next_pow2_u64(n_lhs > order ? n_lhs : order), nR = next_pow2_u64(n_rhs > order ? n_rhs : order);
    // forward transform of an operand: the factorised engine when its length is the transform order (root primitive for it), else
    // the reference's loop literally - one tile up to 4096 values (ntt_exec's single-pass path), global-memory stages above
    auto forward = [&](const fe* src, fe* dst, uint64_t n, const NttOpts& o) -> int {
        if (n == 1) { ZKB_CUDA(c, cudaMemcpyAsync(dst, src, sizeof(fe), cudaMemcpyDeviceToDevice, c->stream)); return 0; }
        if (n == order || n <= (1ull << TILE_LOG)) return ntt_exec(c, root, src, n, 0, dst, 0, 1, ilog2_u64(n), o);
        fe* tmp = const_cast<fe*>(src);
        if (o.has_scale) {                                               // scale(offset) first (ntt_arithmetics.rs:283-284), in place
            DevPow sc;
            ZKB_TRY(get_pow_table(c, o.scale_base, ilog2_u64(n), &sc));
            { LaunchScope ls(c, K_ELEMENTWISE); k_scale<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(src, tmp, n, sc); }
        }
        return ntt_literal_big(c, root, tmp, dst, ilog2_u64(n));
    };
    const uint32_t log_n = ilog2_u64(order);
    DevBuf buf;
    const size_t total = (size_t)(2 * nL + 2 * nR);
    ZKB_TRY(buf.alloc(c, sizeof(fe) * total + 16));
    fe* dL = (fe*)buf.p; fe* eL = dL + nL; fe* dR = eL + nL; fe* eR = dR + nR;
    uint32_t* flag = (uint32_t*)(eR + nR);
    ZKB_CUDA(c, cudaMemsetAsync(buf.p, 0, sizeof(fe) * total + 16, c->stream));
    ZKB_CUDA(c, cudaMemcpyAsync(dL, L, sizeof(fe) * n_lhs, cudaMemcpyHostToDevice, c->stream));
    ZKB_CUDA(c, cudaMemcpyAsync(dR, R, sizeof(fe) * n_rhs, cudaMemcpyHostToDevice, c->stream));
    NttOpts fwd;
    if (divide) { fwd.has_scale = true; fwd.scale_base = h_load(offset_b); }
    ZKB_TRY(forward(dL, eL, nL, fwd));
    ZKB_TRY(forward(dR, eR, nR, fwd));
    unsigned blocks = (unsigned)((order + 127) / 128);
    {
        LaunchScope ls(c, K_ELEMENTWISE);
        if (divide) k_pointwise_div<<<blocks, 128, 0, c->stream>>>(eL, eR, dL, order, flag);
        else k_pointwise_mul<<<blocks, 128, 0, c->stream>>>(eL, eR, dL, order);
    }
    ZKB_CUDA(c, cudaGetLastError());
    if (order > 1) {
        NttOpts inv; inv.inverse = true;
        ZKB_TRY(ntt_exec(c, root, dL, order, 0, dR, 0, 1, log_n, inv));
    } else {
        ZKB_CUDA(c, cudaMemcpyAsync(dR, dL, sizeof(fe), cudaMemcpyDeviceToDevice, c->stream));
    }
    size_t keep = result_len < order ? result_len : (size_t)order;    // coeffs.drain(result_degree..)
    if (divide) {
        DevPow sc;
        ZKB_TRY(get_pow_table(c, h_inv(h_load(offset_b)), log_n, &sc));
        { LaunchScope ls(c, K_ELEMENTWISE); k_scale<<<(unsigned)((keep + 255) / 256), 256, 0, c->stream>>>(dR, dR, keep, sc); }
    }
    uint32_t hflag = 0;
    ZKB_CUDA(c, cudaMemcpyAsync(out, dR, sizeof(fe) * keep, cudaMemcpyDeviceToHost, c->stream));
    ZKB_CUDA(c, cudaMemcpyAsync(&hflag, flag, 4, cudaMemcpyDeviceToHost, c->stream));
    ZKB_CUDA(c, cudaStreamSynchronize(c->stream));
    if (divide && hflag) return set_err(c, ZKB_ERR_DIV_ZERO, "divide by zero");
    if (out_dev) ZKB_CUDA(c, cudaMemcpy(out_dev, out, sizeof(fe) * keep, cudaMemcpyHostToDevice));
    *n_out = keep;
    return 0;
}

int zkb_poly_mul(zkb_ctx* c, const uint8_t root[16], uint64_t root_order, const void* lhs, size_t n_lhs,
                 const void* rhs, size_t n_rhs, void* out, size_t* n_out) {
    return poly_binop(c, false, root, root_order, nullptr, lhs, n_lhs, rhs, n_rhs, out, n_out);
}
int zkb_coset_div(zkb_ctx* c, const uint8_t root[16], uint64_t root_order, const uint8_t offset[16],
                  const void* lhs, size_t n_lhs, const void* rhs, size_t n_rhs, void* out, size_t* n_out) {
    if (!offset) return ZKB_ERR_ARG;
    return poly_binop(c, true, root, root_order, offset, lhs, n_lhs, rhs, n_rhs, out, n_out);
}

}  // extern "C"
