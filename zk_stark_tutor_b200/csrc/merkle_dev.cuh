// merkle_dev.cuh - device building blocks shared by the Merkle kernels (merkle.cu) and the persistent FRI
// tail kernel (fri_tail.cu): compression wrappers, digest loads / stores, the fused FRI fold, and the
// latency-mode level-by-level reduction in shared memory (one compression per QUAD of lanes).
// Reference: MerkleRoot::commit src/merkle_root.rs:7-32, the fold of FRI::commit src/fri.rs:150-159.
#pragma once
#include "merkle.cuh"
#include "blake2b.cuh"

namespace zkb {

// ---- compression wrappers.  __noinline__ keeps ONE copy of each 2.2k-instruction body per
// kernel instead of one per call site.
static __device__ __noinline__ void b2_leaf_call(const fe* a, uint64_t* h) {
    uint64_t out[8];
    blake2b_leaf(*a, out);
#pragma unroll
    for (int i = 0; i < 8; i++) h[i] = out[i];
}
static __device__ __noinline__ void b2_node_call(const uint64_t* l, const uint64_t* r, uint64_t* h) {
    uint64_t m[16], out[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { m[i] = l[i]; m[8 + i] = r[i]; }
    blake2b_compress_1block(m, 128, out);
#pragma unroll
    for (int i = 0; i < 8; i++) h[i] = out[i];
}

__device__ __forceinline__ void g_store_digest(uint8_t* nodes, uint64_t idx, const uint64_t* h) {
    uint4* dst = reinterpret_cast<uint4*>(nodes + idx * 64);
#pragma unroll
    for (int cidx = 0; cidx < 4; cidx++)
        dst[cidx] = make_uint4((uint32_t)h[2 * cidx], (uint32_t)(h[2 * cidx] >> 32),
                               (uint32_t)h[2 * cidx + 1], (uint32_t)(h[2 * cidx + 1] >> 32));
}
__device__ __forceinline__ void g_load_digest(const uint8_t* nodes, uint64_t idx, uint64_t* h) {
    const uint4* src = reinterpret_cast<const uint4*>(nodes + idx * 64);
#pragma unroll
    for (int cidx = 0; cidx < 4; cidx++) {
        uint4 x = __ldg(src + cidx);
        h[2 * cidx] = ((uint64_t)x.y << 32) | x.x;
        h[2 * cidx + 1] = ((uint64_t)x.w << 32) | x.z;
    }
}

__device__ __forceinline__ fe pow2lvl_m(const DevPow& t, uint64_t e) {
    fe lo = fe_ldg(t.lo + (e & ((1ull << t.lo_bits) - 1)));
    fe hi = fe_ldg(t.hi + (e >> t.lo_bits));
    return fe_montmul(hi, lo);
}

// The FRI split-and-fold (fri.rs:150-159) of element i: k_m = alpha/(offset*omega^i) * R
__device__ __forceinline__ fe fold_one(const FoldArgs& f, uint64_t i, const fe& k_m) {
    fe a = fe_ldg(f.cw + i), b = fe_ldg(f.cw + f.half + i);
    fe s = fe_add(a, b), d = fe_sub(a, b);
    fe v = fe_half(fe_add(s, fe_montmul(k_m, d)));
    fe_store(f.next + i, v);
    return v;
}

// Depth-first reduction of 8 digests produced one at a time by `next(j, out)`:
// 7 node compressions, two pending digests at most per level.
template <typename Next>
__device__ __forceinline__ void reduce8(Next next, uint64_t* h) {
    uint64_t d0[8], d1[8], a[8], b[8];
    next(0, d0); next(1, d1); b2_node_call(d0, d1, a);          // level 1, #0
    next(2, d0); next(3, d1); b2_node_call(d0, d1, d1);         // level 1, #1
    b2_node_call(a, d1, b);                                     // level 2, #0
    next(4, d0); next(5, d1); b2_node_call(d0, d1, a);          // level 1, #2
    next(6, d0); next(7, d1); b2_node_call(d0, d1, d1);         // level 1, #3
    b2_node_call(a, d1, a);                                     // level 2, #1
    b2_node_call(b, a, h);                                      // level 3
}

__device__ __forceinline__ uint32_t dig_word(uint32_t i) { return i * 8 + (i >> 1); }
static inline size_t dig_words_host(size_t i) { return i * 8 + (i >> 1); }

// Reduce `n_in` digests held in `cur` to one, level by level, in shared memory (digest i at 8-byte word
// dig_word(i): the two children of a node are one contiguous 128-byte message).  The k-th reduction
// (k = 0, 1, ...) produces relative level first_level + k; node j of it is written to out(level) + (base' + j) * 64
// when out(level) is non-null, base' = base >> (k + 1), `base` = global index of the first input digest at its
// level.  on_root(q, h_lo, h_hi) runs on the four lanes that produced the LAST digest (lane q holds words q, q + 4).
// Returns the buffer whose word 0..7 hold that digest.  Ends with a __syncthreads().
template <typename Out, typename OnRoot>
__device__ __forceinline__ uint64_t* reduce_in_smem(uint64_t* cur, uint64_t* nxt, uint32_t n_in, Out out, uint32_t first_level,
                                                    uint64_t base, OnRoot on_root) {
    const uint32_t tid = threadIdx.x, lane = tid & 31u, q = lane & 3u;
    const uint32_t quad = tid >> 2, quads = blockDim.x >> 2, warp_quad0 = (tid >> 5) << 3;
    uint32_t level = first_level;
    uint32_t cnt = n_in >> 1;
    // levels wider than one warp's eight quads: all warps, block barrier per level.  (One THREAD per node for the widest,
    // throughput-bound levels was measured and is slower: its register pressure costs every other phase of the kernel more.)
    for (; cnt > 8; cnt >>= 1, level++) {
        base >>= 1;
        uint8_t* const dst = out(level);
        for (uint32_t j0 = 0; j0 < cnt; j0 += quads) {
            if (j0 + warp_quad0 >= cnt) break;                      // no live quad in this warp
            const uint32_t j = j0 + quad;
            const bool live = j < cnt;
            uint64_t h_lo, h_hi;
            blake2b_quad(reinterpret_cast<const uint8_t*>(cur + (live ? 17u * j : 0u)), 128, lane, h_lo, h_hi);
            if (live) {
                uint64_t* s = nxt + dig_word(j);
                s[q] = h_lo; s[4 + q] = h_hi;
                if (dst) {
                    unsigned long long* g = reinterpret_cast<unsigned long long*>(dst + (base + j) * 64);
                    g[q] = h_lo; g[4 + q] = h_hi;
                }
            }
        }
        __syncthreads();
        uint64_t* t = cur; cur = nxt; nxt = t;
    }
    // the last (<= 4) levels fit the eight quads of warp 0: no block barrier on this part of the chain
    const uint32_t tail_levels = cnt >= 1 ? 32u - __clz(cnt) : 0u;
    if (tid < 32) {
        uint64_t* c2 = cur; uint64_t* n2 = nxt;
        uint64_t b2 = base;
        uint32_t lv = level;
        for (uint32_t w = cnt; w >= 1; w >>= 1, lv++) {
            b2 >>= 1;
            uint8_t* const dst = out(lv);
            const bool live = quad < w;
            uint64_t h_lo, h_hi;
            blake2b_quad(reinterpret_cast<const uint8_t*>(c2 + (live ? 17u * quad : 0u)), 128, lane, h_lo, h_hi);
            if (live) {
                uint64_t* s = n2 + dig_word(quad);
                s[q] = h_lo; s[4 + q] = h_hi;
                if (dst) {
                    unsigned long long* g = reinterpret_cast<unsigned long long*>(dst + (b2 + quad) * 64);
                    g[q] = h_lo; g[4 + q] = h_hi;
                }
                if (w == 1) on_root(q, h_lo, h_hi);
            }
            __syncwarp();
            uint64_t* t = c2; c2 = n2; n2 = t;
        }
    }
    if (tail_levels & 1u) { uint64_t* t = cur; cur = nxt; nxt = t; }
    __syncthreads();
    return cur;
}
__device__ __forceinline__ void load_chunk(uint64_t* buf, const uint8_t* src, uint32_t n_nodes, bool coherent) {
    const unsigned long long* p = reinterpret_cast<const unsigned long long*>(src);
    for (uint32_t w = threadIdx.x; w < n_nodes * 8; w += blockDim.x)
        buf[dig_word(w >> 3) + (w & 7)] = coherent ? __ldcg(p + w) : __ldg(p + w);
    __syncthreads();
}

}  // namespace zkb
