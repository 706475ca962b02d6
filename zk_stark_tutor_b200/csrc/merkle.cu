// merkle.cu - BLAKE2b-512 Merkle commitment over field elements.
//
// Replaces the bodies of (reference file:line):
//   MerkleRoot::commit  src/merkle_root.rs:7-32   (leaf = H(decimal string), node = H(l||r))
//   MerkleRoot::open    src/merkle_root.rs:34-66  (the reference re-hashes the whole tree per call)
//   MerkleRoot::verify  src/merkle_root.rs:69-95  (host, hosthash.cpp)
//
// Kernel structure.  BLAKE2b is bound by the ALU pipe (2,014 ALU-pipe instructions per
// compression at one per 2 clk per SM sub-partition), so the goal is that every
// resident warp executes compressions back to back with no barrier and no shrinking tail:
//   k_leaf8   one THREAD per 8 consecutive leaves: 8 leaf hashes + the 7 nodes above them
//             (15 compressions, depth first, nothing shared) -> one level-3 node (8 B/leaf
//             written).  With FoldArgs the 8 values are produced by the FRI split-and-fold
//             of the previous layer and written out as the next codeword ("fold fused with
//             the next round's leaf hashing").
//   k_node8   the same over 8 stored nodes: 7 compressions -> the node three levels up.
//   k_leaf1   one thread per leaf, for layers of <= 2^17 leaves (latency-bound: a warp per CTA when small).
//   k_tree    everything above a level of <= 2^17 nodes, and (after k_leaf1) whole trees of <= 2^17
//             leaves: the latency-mode tree - chunks reduced inside one SM's shared memory with one
//             compression spread over FOUR lanes, the last CTA to arrive finishes (see below).  These
//             phases are bound by the latency of a chain of log2(n) compressions, not by throughput.
//             blockIdx.y = instance of a batch of identical small trees (batch.cu).
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "merkle_dev.cuh"
#include "keccak.cuh"
#include "hosthash.hpp"

namespace zkb {

// thresholds (tunable through ZKB_TREE_LEAF_LOG / ZKB_TREE_NODE_LOG for experiments)
static uint32_t tree_leaf_log() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("ZKB_TREE_LEAF_LOG"); v = e ? atoi(e) : ZKB_TREE_LEAF_LOG; if (v > 22) v = 22; if (v < 1) v = 1; }
    return (uint32_t)v;
}
static uint32_t tree_node_log() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("ZKB_TREE_NODE_LOG"); v = e ? atoi(e) : ZKB_TREE_NODE_LOG; if (v > 22) v = 22; if (v < 1) v = 1; }
    return (uint32_t)v;
}

void TreeLayout::init(uint32_t log_n_) {
    log_n = log_n_;
    uint64_t off = 0;
    for (uint32_t l = 0; l <= 40; l++) { level_off[l] = 0; stored[l] = 0; }
    if (log_n <= tree_leaf_log()) {
        top = 0;                                   // small layer: k_leaf1, then the latency-mode tree
    } else {
        top = 3;
        while (log_n - top > tree_node_log()) top += 3;
        for (uint32_t l = 3; l < top; l += 3) stored[l] = 1;
    }
    for (uint32_t l = top; l <= log_n; l++) stored[l] = 1;
    for (uint32_t l = 0; l <= log_n; l++) {
        if (!stored[l]) continue;
        level_off[l] = off;
        off += 1ull << (log_n - l);
    }
    total_nodes = off;
}

// One thread per 8 leaves -> one level-3 node.  n_groups = n / 8.
template <bool FOLD, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) k_leaf8(const fe* __restrict__ vals, FoldArgs f, uint64_t n_groups, uint8_t* __restrict__ out3) {
    const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_groups) return;
    const uint64_t i0 = g * 8;
    fe k_m;
    if (FOLD) k_m = fe_montmul(f.kk_dev ? fe_ldg(reinterpret_cast<const fe*>(f.kk_dev)) : f.kk_m, pow2lvl_m(f.winv, i0 * f.exp_mul));
    uint64_t h[8];
    reduce8([&](int j, uint64_t* out) {
        fe v;
        if (FOLD) {
            v = fold_one(f, i0 + j, k_m);
            k_m = fe_montmul(k_m, f.wr_inv_m);
        } else {
            v = fe_ldg(vals + i0 + j);
        }
        b2_leaf_call(&v, out);
    }, h);
    g_store_digest(out3, g, h);
}

// One thread per 8 stored nodes -> the node three levels up.
template <int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) k_node8(const uint8_t* __restrict__ in, uint64_t n_groups, uint8_t* __restrict__ out) {
    const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_groups) return;
    uint64_t h[8];
    reduce8([&](int j, uint64_t* o) { g_load_digest(in, g * 8 + j, o); }, h);
    g_store_digest(out, g, h);
}

// ---- latency-mode tree -------------------------------------------------------------------------
// Everything above a level of <= 2^17 nodes (and, after k_leaf1, whole layers of <= 2^17 leaves)
// is a chain of log2(n) dependent compressions with too little parallelism to fill the GPU, so
// it is organised for latency.  The `count` input nodes are cut into <= 128 chunks (one CTA, one
// SM each); a CTA takes its chunk into shared memory and reduces it to one node without leaving
// the SM, each node compressed by a QUAD of lanes (blake2b_quad: ~1.2 us per level instead of
// ~6 us for a thread-per-node level with a grid barrier); warps without a live quad skip the
// level.  The last CTA to arrive (one atomic counter, no spinning, no cooperative launch)
// gathers the chunk roots and finishes the tree the same way, then hands the root to the
// polling host.  Every level is also written to global memory: the tree stores them for
// openings.  Shared layout: digest i at 8-byte word 8i + (i >> 1), i.e. the two children of a
// node are one contiguous 128-byte message and consecutive messages are 136 B apart (banks).
#define ZKB_TREE_MAX_CHUNK 1024
struct TopArgs {
    const uint8_t* nodes_in; // `count` input nodes (power of two, <= 2^20)
    uint32_t count;
    uint32_t chunk_log;      // nodes per CTA = 2^chunk_log (== count: a single CTA does everything)
    uint8_t* level_out[26];  // global destination of relative level r >= 1
    uint64_t batch_stride;   // bytes between the node arenas of the instances of a batch (blockIdx.y); 0 for one tree
    uint32_t* bar;           // arrival counter per instance (zero between launches: the last CTA resets it)
    uint8_t* host_root;      // mapped pinned host memory (or nullptr)
    volatile uint32_t* host_flag;
    uint32_t seq;
    FsDev* fs;               // device Fiat-Shamir (or nullptr): instance blockIdx.y at fs + blockIdx.y
    uint32_t fs_round, fs_want_alpha;
};

// After the root of round `round`: Root(root) into the transcript sponge, the challenge, the next fold constant.
// One warp; `root` = 8 words in shared memory.
__device__ __forceinline__ void fs_after_root(FsDev* fs, FsSponge* sp, const uint64_t* root, uint32_t round, bool want_alpha, uint32_t lane) {
    uint64_t* d = reinterpret_cast<uint64_t*>(sp);
    const uint64_t* s = reinterpret_cast<const uint64_t*>(&fs->sp);
    for (uint32_t i = lane; i < sizeof(FsSponge) / 8; i += 32) d[i] = s[i];
    __syncwarp();
    const fe alpha = fs_round_warp(sp, reinterpret_cast<const uint8_t*>(root), want_alpha, lane);
    uint64_t* g = reinterpret_cast<uint64_t*>(&fs->sp);
    for (uint32_t i = lane; i < sizeof(FsSponge) / 8; i += 32) g[i] = d[i];
    if (lane < 8) reinterpret_cast<uint64_t*>(fs->roots[round])[lane] = root[lane];
    if (lane == 0 && want_alpha) {
        fe_store(&fs->alpha, alpha);
        fe_store(&fs->kk_m, fe_montmul(alpha, fe_load(&fs->inv_off_m2[round])));
    }
}

// One thread per leaf -> level-0 digests (the input of k_tree for small layers).
template <bool FOLD>
__global__ void __launch_bounds__(256, 2) k_leaf1(const fe* __restrict__ vals, FoldArgs f, uint32_t n, uint8_t* __restrict__ out0, BatchArgs b) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t inst = blockIdx.y;                      // instance of a batch (0 for a single tree)
    vals += (uint64_t)inst * b.vals_stride;
    out0 += (uint64_t)inst * b.nodes_stride;
    fe v;
    if (FOLD) {
        f.cw += (uint64_t)inst * b.vals_stride;
        f.next += (uint64_t)inst * b.next_stride;
        const fe kk = f.kk_dev ? fe_ldg(reinterpret_cast<const fe*>(f.kk_dev + (uint64_t)inst * f.kk_stride)) : b.kk_m ? fe_ldg(b.kk_m + inst) : f.kk_m;
        fe k_m = fe_montmul(kk, pow2lvl_m(f.winv, (uint64_t)i * f.exp_mul));
        v = fold_one(f, i, k_m);
    } else {
        v = fe_ldg(vals + i);
    }
    uint64_t h[8];
    b2_leaf_call(&v, h);
    g_store_digest(out0, i, h);
}

__global__ void __launch_bounds__(512, 1) k_tree(TopArgs a) {
    extern __shared__ uint64_t tree_smem[];
    __shared__ uint32_t s_last;
    __shared__ __align__(16) FsSponge s_sponge;
    const uint32_t chunk = 1u << a.chunk_log, chunks = a.count >> a.chunk_log;
    const uint32_t nmax = chunk > chunks ? chunk : chunks;
    const uint64_t boff = (uint64_t)blockIdx.y * a.batch_stride;   // instance blockIdx.y of a batch: every buffer shifts by boff
    uint32_t* const bar = a.bar + blockIdx.y;
    uint64_t* bufA = tree_smem;
    uint64_t* bufB = tree_smem + dig_word(nmax) + 8;
    // stage 0: this CTA's chunk; stage 1 (last CTA to arrive only): the chunk roots
    const uint8_t* src = a.nodes_in + boff + (size_t)blockIdx.x * chunk * 64;
    uint32_t n_in = chunk, first_level = 1;
    uint64_t base = (uint64_t)blockIdx.x * chunk;
    bool last_stage = chunks == 1;
    auto out = [&](uint32_t level) -> uint8_t* { return a.level_out[level] + boff; };
#pragma unroll 1
    for (;;) {
        load_chunk(bufA, src, n_in, first_level != 1);
        const bool signal = last_stage && a.host_root != nullptr;
        const uint64_t* root = reduce_in_smem(bufA, bufB, n_in, out, first_level, base, [&](uint32_t q, uint64_t h_lo, uint64_t h_hi) {
            if (!signal) return;                               // the root: hand it to the polling host
            unsigned long long* hr = reinterpret_cast<unsigned long long*>(a.host_root);
            hr[q] = h_lo; hr[4 + q] = h_hi;
            __threadfence_system();
            __syncwarp(0xFu);
            if (q == 0) { __threadfence_system(); *a.host_flag = a.seq; }   // the flag store is ordered after all four lanes' root stores
        });
        if (last_stage) {
            if (a.fs && threadIdx.x < 32) fs_after_root(a.fs + blockIdx.y, &s_sponge, root, a.fs_round, a.fs_want_alpha != 0, threadIdx.x);
            return;
        }
        if (threadIdx.x == 0) {                // (the __syncthreads closing the reduction ordered the CTA's stores before this)
            __threadfence();
            const uint32_t arrived = atomicAdd(bar, 1u);
            s_last = arrived == gridDim.x - 1;
            if (s_last) { *bar = 0; __threadfence(); }
        }
        __syncthreads();
        if (!s_last) return;
        src = a.level_out[a.chunk_log] + boff;
        n_in = chunks; first_level = a.chunk_log + 1; base = 0; last_stage = true;
    }
}

// Authentication paths.  One warp per opened index.  Stored levels are read; the levels inside
// a group of three whose base is the leaves or a stored level are recomputed from the 8 group
// inputs (8 leaf hashes or 8 loads, then 4 + 2 compressions).
struct OpenArgs {
    const fe* vals;
    const uint8_t* nodes;
    TreeLayout layout;
    const uint64_t* idx;
    uint32_t k;
    uint8_t* out;            // k * log_n * 64 bytes
    uint64_t vals_stride, nodes_stride;   // batch (blockIdx.y): elements / bytes between instances; idx and out are dense (k per instance)
};
__global__ void __launch_bounds__(128) k_open(OpenArgs a) {
    __shared__ uint64_t dig[4][14][8];          // per warp: 8 + 4 + 2 digests
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    a.vals += (uint64_t)blockIdx.y * a.vals_stride;
    a.nodes += (uint64_t)blockIdx.y * a.nodes_stride;
    a.idx += (uint64_t)blockIdx.y * a.k;
    a.out += (uint64_t)blockIdx.y * a.k * a.layout.log_n * 64;
    const uint32_t q = blockIdx.x * 4 + warp;
    if (q >= a.k) return;
    const uint64_t idx = a.idx[q];
    const uint32_t log_n = a.layout.log_n;
    uint4* out = reinterpret_cast<uint4*>(a.out + (uint64_t)q * log_n * 64);
    uint64_t (*D)[8] = dig[warp];
    uint32_t l = 0;
    while (l < log_n) {
        if (a.layout.stored[l] && (l + 1 > log_n || a.layout.stored[l + 1] || l + 1 == log_n)) {
            // sibling available directly, and the next level does not need a recomputation
            const uint4* sib = reinterpret_cast<const uint4*>(a.nodes + (a.layout.level_off[l] + ((idx >> l) ^ 1)) * 64);
            if (lane < 4) out[l * 4 + lane] = __ldg(sib + lane);
            l++;
            continue;
        }
        // group with base level l (leaves when l == 0 and not stored, else a stored level):
        // emits the siblings at levels l, l+1, l+2
        const uint64_t base = (idx >> (l + 3)) << 3;                // first of the 8 group inputs at level l
        if (lane < 8) {
            uint64_t h[8];
            if (a.layout.stored[l]) {
                g_load_digest(a.nodes + a.layout.level_off[l] * 64, base + lane, h);
            } else {
                fe v = fe_ldg(a.vals + base + lane);
                b2_leaf_call(&v, h);
            }
            for (int i = 0; i < 8; i++) D[lane][i] = h[i];
        }
        __syncwarp();
        if (lane < 4) {
            uint64_t h[8];
            b2_node_call(D[2 * lane], D[2 * lane + 1], h);
            for (int i = 0; i < 8; i++) D[8 + lane][i] = h[i];
        }
        __syncwarp();
        if (lane < 2) {
            uint64_t h[8];
            b2_node_call(D[8 + 2 * lane], D[8 + 2 * lane + 1], h);
            for (int i = 0; i < 8; i++) D[12 + lane][i] = h[i];
        }
        __syncwarp();
        const uint32_t p0 = (uint32_t)(idx >> l) & 7u;
        const uint4* s0 = reinterpret_cast<const uint4*>(D[p0 ^ 1]);
        const uint4* s1 = reinterpret_cast<const uint4*>(D[8 + ((p0 >> 1) ^ 1)]);
        const uint4* s2 = reinterpret_cast<const uint4*>(D[12 + ((p0 >> 2) ^ 1)]);
        if (lane < 4) {
            out[l * 4 + lane] = s0[lane];
            out[(l + 1) * 4 + lane] = s1[lane];
            out[(l + 2) * 4 + lane] = s2[lane];
        }
        __syncwarp();
        l += 3;
    }
}

// ---- openings in WIRE FORMAT (src/stark/proof_stream_enum.rs:67-127) ------------------------------------------------------
// The batched provers used to bring raw paths to the host and let host threads wrap every node in its object framing
// (1.16 MB of proof per RPSSS signature, 14,000 nodes): that assembly, not the GPU, bounded the batch workloads.  Here the kernel
// that computes an authentication path writes the finished objects - [Value: 04 | u64_be(16) | value_be16] and
// Path: 02 | u64_be(72 d) | d x (u64_be(64) | node) - at the byte offset the object has inside its proof, so the host appends
// one contiguous segment per proof.  One warp per opening; the record is built in shared memory at the destination's 16-byte
// phase and leaves as aligned 128-bit stores (byte stores only for the ragged head and tail).
struct OpenWireArgs {
    const fe* vals;
    const uint8_t* nodes;
    TreeLayout layout;
    const uint64_t* idx;             // k indices per blockIdx.y
    uint32_t k;
    uint64_t vals_stride, nodes_stride;   // per blockIdx.y
    uint8_t* out;                    // wire buffer (device)
    const uint64_t* y_off;           // byte offset of blockIdx.y's segment inside `out`
    uint32_t qdiv;                   // record q at y_off[y] + base + (q / qdiv) * stride_hi + (q % qdiv) * stride_lo
    uint64_t base, stride_hi, stride_lo;
    uint32_t with_value;
};
#define ZKB_WIRE_MAX_DEPTH 30
#define ZKB_WIRE_REC_BYTES (16 + 25 + 9 + ZKB_WIRE_MAX_DEPTH * 72 + 16)
__device__ __forceinline__ void wire_put_be64(uint8_t* p, uint64_t v, uint32_t lane) {
    if (lane < 8) p[lane] = (uint8_t)(v >> (56 - 8 * lane));
}
__global__ void __launch_bounds__(128) k_open_wire(OpenWireArgs a) {
    __shared__ uint64_t dig[4][14][8];          // per warp: 8 + 4 + 2 digests
    __shared__ __align__(16) uint8_t recs[4][(ZKB_WIRE_REC_BYTES + 15) & ~15];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    a.vals += (uint64_t)blockIdx.y * a.vals_stride;
    a.nodes += (uint64_t)blockIdx.y * a.nodes_stride;
    a.idx += (uint64_t)blockIdx.y * a.k;
    const uint32_t q = blockIdx.x * 4 + warp;
    if (q >= a.k) return;
    const uint64_t idx = a.idx[q];
    const uint32_t log_n = a.layout.log_n;
    uint8_t* const dest = a.out + a.y_off[blockIdx.y] + a.base + (uint64_t)(q / a.qdiv) * a.stride_hi + (uint64_t)(q % a.qdiv) * a.stride_lo;
    const uint32_t m = (uint32_t)(reinterpret_cast<uintptr_t>(dest) & 15u);
    uint8_t* rec = recs[warp] + m;               // record byte p lives at rec[p]: same 16-byte phase as the destination
    uint32_t pos = 0;
    if (a.with_value) {                          // stark.rs:550-553: Value(codeword[i])
        const fe v = fe_ldg(a.vals + idx);
        if (lane == 0) rec[0] = 4;
        wire_put_be64(rec + 1, 16, lane);
        if (lane < 16) rec[9 + lane] = (uint8_t)(v.v[3 - (lane >> 2)] >> (24 - 8 * (lane & 3)));
        pos = 25;
    }
    if (lane == 0) rec[pos] = 2;                 // Path(open(i))
    wire_put_be64(rec + pos + 1, (uint64_t)log_n * 72, lane);
    pos += 9;
    uint64_t (*D)[8] = dig[warp];
    auto put_node = [&](uint32_t level, const uint8_t* src) {          // 64 bytes from shared or global memory
        uint8_t* o = rec + pos + level * 72;
        wire_put_be64(o, 64, lane);
        const uint16_t w = reinterpret_cast<const uint16_t*>(src)[lane];
        o[8 + 2 * lane] = (uint8_t)w;
        o[9 + 2 * lane] = (uint8_t)(w >> 8);
    };
    uint32_t l = 0;
    while (l < log_n) {
        if (a.layout.stored[l] && (l + 1 > log_n || a.layout.stored[l + 1] || l + 1 == log_n)) {
            put_node(l, a.nodes + (a.layout.level_off[l] + ((idx >> l) ^ 1)) * 64);
            l++;
            continue;
        }
        // group with base level l (leaves when l == 0 and not stored, else a stored level): siblings at levels l, l+1, l+2
        const uint64_t base = (idx >> (l + 3)) << 3;
        if (lane < 8) {
            uint64_t h[8];
            if (a.layout.stored[l]) {
                g_load_digest(a.nodes + a.layout.level_off[l] * 64, base + lane, h);
            } else {
                fe v = fe_ldg(a.vals + base + lane);
                b2_leaf_call(&v, h);
            }
            for (int i = 0; i < 8; i++) D[lane][i] = h[i];
        }
        __syncwarp();
        if (lane < 4) {
            uint64_t h[8];
            b2_node_call(D[2 * lane], D[2 * lane + 1], h);
            for (int i = 0; i < 8; i++) D[8 + lane][i] = h[i];
        }
        __syncwarp();
        if (lane < 2) {
            uint64_t h[8];
            b2_node_call(D[8 + 2 * lane], D[8 + 2 * lane + 1], h);
            for (int i = 0; i < 8; i++) D[12 + lane][i] = h[i];
        }
        __syncwarp();
        const uint32_t p0 = (uint32_t)(idx >> l) & 7u;
        put_node(l, reinterpret_cast<const uint8_t*>(D[p0 ^ 1]));
        put_node(l + 1, reinterpret_cast<const uint8_t*>(D[8 + ((p0 >> 1) ^ 1)]));
        put_node(l + 2, reinterpret_cast<const uint8_t*>(D[12 + ((p0 >> 2) ^ 1)]));
        __syncwarp();
        l += 3;
    }
    __syncwarp();
    // ---- the record leaves: bytes [m, m + total) of the 16-byte-aligned buffer recs[warp] -> dest - m + the same offsets
    const uint32_t total = pos + log_n * 72, end = m + total;
    const uint8_t* sb = recs[warp];
    uint8_t* db = dest - m;
    const uint32_t c0 = (m + 15u) >> 4, c1 = end >> 4;                   // full 16-byte chunks [c0, c1)
    if (c0 < c1) {
        for (uint32_t cidx = c0 + lane; cidx < c1; cidx += 32)
            reinterpret_cast<uint4*>(db)[cidx] = reinterpret_cast<const uint4*>(sb)[cidx];
        if (lane < 16) {
            const uint32_t ph = m + lane, pt = (c1 << 4) + lane;
            if (ph < (c0 << 4)) db[ph] = sb[ph];                         // ragged head
            if (pt < end) db[pt] = sb[pt];                               // ragged tail
        }
    } else {
        const uint32_t p = m + lane;                                     // shorter than two chunks: byte by byte
        if (p < end) db[p] = sb[p];
        if (p + 32 < end) db[p + 32] = sb[p + 32];
    }
}

// Leafs objects of one FRI round for every instance (fri.rs:189-195): 03 | u64_be(48) | cur[a], cur[a + half], nxt[a] big-endian
__global__ void k_leafs_wire(const fe* cur, uint64_t cur_stride, const fe* nxt, uint64_t nxt_stride, uint64_t half,
                             const uint64_t* idx, uint32_t k, uint8_t* out, const uint64_t* y_off, uint64_t base) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= k) return;
    const uint32_t b = blockIdx.y;
    cur += (uint64_t)b * cur_stride;
    nxt += (uint64_t)b * nxt_stride;
    const uint64_t i = idx[(uint64_t)b * k + s];
    const fe v[3] = {fe_ldg(cur + i), fe_ldg(cur + i + half), fe_ldg(nxt + i)};
    uint8_t* o = out + y_off[b] + base + (uint64_t)s * 57;
    o[0] = 3;
    for (int j = 0; j < 7; j++) o[1 + j] = 0;
    o[8] = 48;
    for (int e = 0; e < 3; e++)
        for (int j = 0; j < 16; j++) o[9 + 16 * e + j] = (uint8_t)(v[e].v[3 - (j >> 2)] >> (24 - 8 * (j & 3)));
}

// values at k indices -> contiguous (for the Value objects of the Stark query phase)
__global__ void k_gather_vals(const fe* __restrict__ vals, const uint64_t* __restrict__ idx, uint32_t k, fe* __restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < k) fe_store(out + i, fe_ldg(vals + idx[i]));
}

static uint32_t tree_split_log() {   // trees of up to 2^this nodes are reduced by ONE CTA (ZKB_TREE_SPLIT_LOG)
    static int v = -1;
    if (v < 0) { const char* e = getenv("ZKB_TREE_SPLIT_LOG"); v = e ? atoi(e) : 8; if (v > 10) v = 10; if (v < 1) v = 1; }
    return (uint32_t)v;
}
static uint32_t tree_chunks_log() {  // larger ones are cut into 2^this chunks (ZKB_TREE_CHUNKS_LOG; 7: 128 CTAs <= 148 SMs)
    static int v = -1;
    if (v < 0) { const char* e = getenv("ZKB_TREE_CHUNKS_LOG"); v = e ? atoi(e) : 7; if (v > 10) v = 10; if (v < 1) v = 1; }
    return (uint32_t)v;
}
// per-device kernel attributes, set when a context is created on the device (zkb_ctx_create)
int merkle_device_init(zkb_ctx* c) {
    const size_t max_smem = (size_t)(dig_words_host(ZKB_TREE_MAX_CHUNK) + 8 + dig_words_host(ZKB_TREE_MAX_CHUNK / 2) + 8) * sizeof(uint64_t);
    ZKB_CUDA(c, cudaFuncSetAttribute(k_tree, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_smem));
    return 0;
}
static int launch_top(zkb_ctx* c, const TopArgs& a, uint32_t batch = 1) {
    if (a.count > (1u << 20) || a.count < 2) return set_err(c, ZKB_ERR_ARG, "internal: k_tree takes 2..2^20 nodes");
    const uint32_t log_c = ilog2_u64(a.count);
    TopArgs b = a;
    if (log_c <= tree_split_log()) b.chunk_log = log_c;
    else {
        b.chunk_log = log_c > tree_chunks_log() + 1 ? log_c - tree_chunks_log() : 1;
        if (b.chunk_log > 10) b.chunk_log = 10;
    }
    if (batch > 1) {
        // a batch is a throughput problem: the fattest chunks that still give >= 128 CTAs over the whole batch
        // (fewer waves of CTAs, fewer latency-bound finishing stages)
        uint32_t cl = log_c < 10 ? log_c : 10;
        while (cl > b.chunk_log && (uint64_t)batch * (a.count >> cl) < 128) cl--;
        b.chunk_log = cl;
    }
    const uint32_t chunk = 1u << b.chunk_log, chunks = a.count >> b.chunk_log;
    const uint32_t nmax = chunk > chunks ? chunk : chunks;
    uint32_t threads = 2 * nmax;                          // one quad per node of the widest level computed
    if (threads < 128) threads = 128;
    if (threads > 512) threads = 512;
    const size_t smem = (size_t)(dig_words_host(nmax) + 8 + dig_words_host(nmax / 2) + 8) * sizeof(uint64_t);
    if (batch < 1 || batch > ZKB_MAX_BATCH) return set_err(c, ZKB_ERR_ARG, "internal: tree batch %u out of range", batch);
    if (!c->tree_bars) {
        ZKB_CUDA(c, cudaMalloc(&c->tree_bars, ZKB_MAX_BATCH * sizeof(uint32_t)));
        ZKB_CUDA(c, cudaMemsetAsync(c->tree_bars, 0, ZKB_MAX_BATCH * sizeof(uint32_t), c->stream));
    }
    b.bar = c->tree_bars;
    {
        LaunchScope ls(c, K_MERKLE_SMALL);
        k_tree<<<dim3(chunks, batch), threads, smem, c->stream>>>(b);
    }
    ZKB_CUDA(c, cudaGetLastError());
    return 0;
}

int wait_root(zkb_ctx* c, const RootSignal& s, uint8_t root_out[64]) {
    uint64_t spins = 0;
    while (*s.host_flag != s.seq) {
        if ((++spins & 0xFFFF) == 0) {                  // every ~64k polls make sure the kernel is still alive
            cudaError_t e = cudaStreamQuery(c->stream);
            if (e != cudaSuccess && e != cudaErrorNotReady)
                return set_err(c, ZKB_ERR_CUDA, "Merkle kernel failed: %s", cudaGetErrorString(e));
            if (e == cudaSuccess && *s.host_flag != s.seq)
                return set_err(c, ZKB_ERR_CUDA, "Merkle kernel finished without signalling its root");
        }
    }
    __sync_synchronize();
    memcpy(root_out, s.host_root, 64);
    return 0;
}

int merkle_build_levels(zkb_ctx* c, const fe* vals, const FoldArgs* fold, uint64_t n,
                        const TreeLayout& L, uint8_t* nodes, const RootSignal* signal, const FsHook* fs, bool leaf3_done) {
    const uint32_t log_n = L.log_n;
    TopArgs a;
    memset(&a, 0, sizeof(a));
    if (fs && fs->fs) {
        if (n < 2) return set_err(c, ZKB_ERR_ARG, "internal: device Fiat-Shamir needs a tree of at least two leaves");
        a.fs = fs->fs; a.fs_round = fs->round; a.fs_want_alpha = fs->want_alpha;
    }
    if (signal && n > 1) { a.host_root = signal->host_root; a.host_flag = signal->host_flag; a.seq = signal->seq; }
    if (L.top == 0) {                                   // small layer: thread per leaf, then the latency-mode tree
        FoldArgs fa;
        if (fold) fa = *fold; else memset(&fa, 0, sizeof(fa));
        {
            LaunchScope ls(c, K_LEAF1);
            // latency-bound: small layers go out one warp per CTA so every warp gets a scheduler of its own
            const unsigned threads = n <= (1u << 14) ? 32u : n <= (1u << 15) ? 64u : 256u;
            const unsigned blocks = (unsigned)((n + threads - 1) / threads);
            if (fold) k_leaf1<true><<<blocks, threads, 0, c->stream>>>(nullptr, fa, (uint32_t)n, nodes + L.level_off[0] * 64, BatchArgs());
            else k_leaf1<false><<<blocks, threads, 0, c->stream>>>(vals, fa, (uint32_t)n, nodes + L.level_off[0] * 64, BatchArgs());
        }
        ZKB_CUDA(c, cudaGetLastError());
        if (n == 1) return 0;
        a.nodes_in = nodes + L.level_off[0] * 64;
        a.count = (uint32_t)n;
        for (uint32_t l = 1; l <= log_n; l++) a.level_out[l] = nodes + L.level_off[l] * 64;
        return launch_top(c, a);
    }
    const uint64_t groups = n >> 3;
    uint8_t* lvl3 = nodes + L.level_off[3] * 64;
    static int cfg = -1;                      // tuning knob (ZKB_LEAF_CFG): threads x min resident CTAs
    if (cfg < 0) { const char* e = getenv("ZKB_LEAF_CFG"); cfg = e ? atoi(e) : 0; }
    FoldArgs fa;
    if (fold) fa = *fold; else memset(&fa, 0, sizeof(fa));
    if (!leaf3_done) {
        LaunchScope ls(c, fold ? K_FOLD_LEAF_TILE : K_LEAF_TILE);
#define ZKB_LEAF_LAUNCH(T, M)                                                                                   \
        do {                                                                                                    \
            const unsigned blocks = (unsigned)((groups + (T) - 1) / (T));                                       \
            if (fold) k_leaf8<true, T, M><<<blocks, T, 0, c->stream>>>(nullptr, fa, groups, lvl3);               \
            else k_leaf8<false, T, M><<<blocks, T, 0, c->stream>>>(vals, fa, groups, lvl3);                      \
        } while (0)
        if (cfg == 1) ZKB_LEAF_LAUNCH(256, 3);
        else if (cfg == 2) ZKB_LEAF_LAUNCH(128, 4);
        else if (cfg == 3) ZKB_LEAF_LAUNCH(128, 6);
        else if (cfg == 4) ZKB_LEAF_LAUNCH(64, 8);
        else ZKB_LEAF_LAUNCH(256, 2);
    }
    ZKB_CUDA(c, cudaGetLastError());
    uint32_t level = 3;
    while (level < L.top) {
        const uint64_t g = n >> (level + 3);
        {
            LaunchScope ls(c, K_NODE_TILE);
            if (g >= (1u << 17)) k_node8<256, 2><<<(unsigned)((g + 255) / 256), 256, 0, c->stream>>>(nodes + L.level_off[level] * 64, g, nodes + L.level_off[level + 3] * 64);
            else k_node8<64, 8><<<(unsigned)((g + 63) / 64), 64, 0, c->stream>>>(nodes + L.level_off[level] * 64, g, nodes + L.level_off[level + 3] * 64);
        }
        ZKB_CUDA(c, cudaGetLastError());
        level += 3;
    }
    a.nodes_in = nodes + L.level_off[level] * 64;
    a.count = (uint32_t)(n >> level);
    for (uint32_t r = 1; level + r <= log_n; r++) a.level_out[r] = nodes + L.level_off[level + r] * 64;
    if (a.count > 1) ZKB_TRY(launch_top(c, a));
    return 0;
}

int merkle_build_levels_batch(zkb_ctx* c, const fe* vals, const FoldArgs* fold, uint64_t n,
                              const TreeLayout& L, uint8_t* nodes, const BatchArgs& b, const FsHook* fs) {
    if (L.top != 0) return set_err(c, ZKB_ERR_ARG, "internal: batched trees are limited to 2^%u leaves", tree_leaf_log());
    if (b.batch < 1 || b.batch > ZKB_MAX_BATCH) return set_err(c, ZKB_ERR_ARG, "batch of %u trees not supported (max %u)", b.batch, ZKB_MAX_BATCH);
    FoldArgs fa;
    if (fold) fa = *fold; else memset(&fa, 0, sizeof(fa));
    {
        LaunchScope ls(c, K_LEAF1);
        const unsigned threads = n * b.batch <= (1u << 14) ? 32u : n * b.batch <= (1u << 15) ? 64u : 256u;
        const dim3 grid((unsigned)((n + threads - 1) / threads), b.batch);
        if (fold) k_leaf1<true><<<grid, threads, 0, c->stream>>>(nullptr, fa, (uint32_t)n, nodes + L.level_off[0] * 64, b);
        else k_leaf1<false><<<grid, threads, 0, c->stream>>>(vals, fa, (uint32_t)n, nodes + L.level_off[0] * 64, b);
    }
    ZKB_CUDA(c, cudaGetLastError());
    if (n == 1) return 0;
    TopArgs a;
    memset(&a, 0, sizeof(a));
    a.nodes_in = nodes + L.level_off[0] * 64;
    a.count = (uint32_t)n;
    a.batch_stride = b.nodes_stride;
    if (fs && fs->fs) { a.fs = fs->fs; a.fs_round = fs->round; a.fs_want_alpha = fs->want_alpha; }
    for (uint32_t l = 1; l <= L.log_n; l++) a.level_out[l] = nodes + L.level_off[l] * 64;
    return launch_top(c, a, b.batch);
}

int merkle_batch_roots(zkb_ctx* c, const TreeLayout& L, const uint8_t* nodes, const BatchArgs& b, uint8_t* roots_host) {
    ZKB_CUDA(c, cudaMemcpy2DAsync(roots_host, 64, nodes + L.level_off[L.log_n] * 64, b.nodes_stride ? b.nodes_stride : 64, 64, b.batch,
                                  cudaMemcpyDeviceToHost, c->stream));
    ZKB_CUDA(c, ctx_stream_sync(c));
    return 0;
}

int merkle_open_device(zkb_ctx* c, const fe* vals, const TreeLayout& layout, const uint8_t* nodes,
                       const uint64_t* d_idx, size_t k, uint8_t* d_out) {
    if (k == 0) return 0;
    OpenArgs a;
    a.vals = vals; a.nodes = nodes; a.layout = layout; a.idx = d_idx; a.k = (uint32_t)k; a.out = d_out;
    a.vals_stride = 0; a.nodes_stride = 0;
    { LaunchScope ls(c, K_OPEN); k_open<<<(unsigned)((k + 3) / 4), 128, 0, c->stream>>>(a); }
    ZKB_CUDA(c, cudaGetLastError());
    return 0;
}

int merkle_open_device_batch(zkb_ctx* c, const fe* vals, const TreeLayout& layout, const uint8_t* nodes,
                             const uint64_t* d_idx, size_t k, uint8_t* d_out, uint32_t batch,
                             uint64_t vals_stride, uint64_t nodes_stride) {
    if (k == 0 || batch == 0) return 0;
    OpenArgs a;
    a.vals = vals; a.nodes = nodes; a.layout = layout; a.idx = d_idx; a.k = (uint32_t)k; a.out = d_out;
    a.vals_stride = vals_stride; a.nodes_stride = nodes_stride;
    { LaunchScope ls(c, K_OPEN); k_open<<<dim3((unsigned)((k + 3) / 4), batch), 128, 0, c->stream>>>(a); }
    ZKB_CUDA(c, cudaGetLastError());
    return 0;
}

// One launch: k openings per instance (blockIdx.y) written as wire objects; see OpenWireArgs.
int merkle_open_wire_batch(zkb_ctx* c, const fe* vals, const TreeLayout& layout, const uint8_t* nodes, const uint64_t* d_idx, size_t k,
                           uint32_t batch, uint64_t vals_stride, uint64_t nodes_stride, uint8_t* d_out, const uint64_t* d_y_off,
                           uint32_t qdiv, uint64_t base, uint64_t stride_hi, uint64_t stride_lo, bool with_value) {
    if (k == 0 || batch == 0) return 0;
    if (layout.log_n > ZKB_WIRE_MAX_DEPTH || layout.log_n < 1) return set_err(c, ZKB_ERR_ARG, "internal: wire-format openings take trees of 2..2^%d leaves", ZKB_WIRE_MAX_DEPTH);
    OpenWireArgs a;
    a.vals = vals; a.nodes = nodes; a.layout = layout; a.idx = d_idx; a.k = (uint32_t)k;
    a.vals_stride = vals_stride; a.nodes_stride = nodes_stride;
    a.out = d_out; a.y_off = d_y_off; a.qdiv = qdiv ? qdiv : 1; a.base = base; a.stride_hi = stride_hi; a.stride_lo = stride_lo;
    a.with_value = with_value;
    { LaunchScope ls(c, K_OPEN); k_open_wire<<<dim3((unsigned)((k + 3) / 4), batch), 128, 0, c->stream>>>(a); }
    ZKB_CUDA(c, cudaGetLastError());
    return 0;
}
int fri_leafs_wire_batch(zkb_ctx* c, const fe* cur, uint64_t cur_stride, const fe* nxt, uint64_t nxt_stride, uint64_t half,
                         const uint64_t* d_idx, size_t k, uint32_t batch, uint8_t* d_out, const uint64_t* d_y_off, uint64_t base) {
    if (k == 0 || batch == 0) return 0;
    { LaunchScope ls(c, K_GATHER); k_leafs_wire<<<dim3((unsigned)((k + 63) / 64), batch), 64, 0, c->stream>>>(cur, cur_stride, nxt, nxt_stride, half, d_idx, (uint32_t)k, d_out, d_y_off, base); }
    ZKB_CUDA(c, cudaGetLastError());
    return 0;
}

}  // namespace zkb

using namespace zkb;

extern "C" {

int zkb_merkle_build(zkb_ctx* c, const void* vals, size_t n, zkb_tree** out) {
    if (!c || !vals || !out) return ZKB_ERR_ARG;
    *out = nullptr;
    if (n == 0 || (n & (n - 1))) return set_err(c, ZKB_ERR_NOT_POW2, "Leafs len must be power of two (got %zu)", n);
    if (n > (1ull << 36)) return set_err(c, ZKB_ERR_ARG, "tree too large");
    ZKB_CUDA(c, cudaSetDevice(c->device));
    std::unique_ptr<zkb_tree> t(new zkb_tree());
    t->ctx = c;
    t->n = n;
    t->layout.init(ilog2_u64(n));
    if (is_device_ptr(vals)) {
        t->vals = (const fe*)vals;
    } else {
        ZKB_CUDA(c, dev_alloc(c, &t->owned_vals, n * sizeof(fe)));
        cudaError_t e = cudaMemcpyAsync(t->owned_vals, vals, n * sizeof(fe), cudaMemcpyHostToDevice, c->stream);
        if (e != cudaSuccess) { dev_free(c, t->owned_vals); return set_err(c, ZKB_ERR_CUDA, "H2D copy failed: %s", cudaGetErrorString(e)); }
        t->vals = (const fe*)t->owned_vals;
    }
    cudaError_t e = dev_alloc(c, (void**)&t->nodes, t->layout.total_nodes * 64);
    if (e != cudaSuccess) { dev_free(c, t->owned_vals); return set_err(c, ZKB_ERR_CUDA, "cudaMalloc(tree) failed: %s", cudaGetErrorString(e)); }
    int rc = merkle_build_levels(c, t->vals, nullptr, n, t->layout, t->nodes);
    if (rc == 0) {
        e = cudaMemcpyAsync(c->pinned, t->nodes + t->layout.level_off[t->layout.log_n] * 64, 64, cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) rc = set_err(c, ZKB_ERR_CUDA, "merkle build failed: %s", cudaGetErrorString(e));
    }
    if (rc != 0) { dev_free(c, t->nodes); dev_free(c, t->owned_vals); return rc; }
    memcpy(t->root, c->pinned, 64);
    *out = t.release();
    return 0;
}

int zkb_merkle_root(const zkb_tree* t, uint8_t root[64]) {
    if (!t || !root) return ZKB_ERR_ARG;
    memcpy(root, t->root, 64);
    return 0;
}

void zkb_merkle_free(zkb_tree* t) {
    if (!t) return;
    cudaSetDevice(t->ctx->device);
    if (t->owns_nodes) dev_free(t->ctx, t->nodes);
    dev_free(t->ctx, t->owned_vals);
    delete t;
}

int zkb_merkle_commit(zkb_ctx* c, const void* vals, size_t n, uint8_t root[64]) {
    if (!c || !vals || !root) return ZKB_ERR_ARG;
    if (n == 0 || (n & (n - 1))) return set_err(c, ZKB_ERR_NOT_POW2, "Leafs len must be power of two (got %zu)", n);
    ZKB_CUDA(c, cudaSetDevice(c->device));
    TreeLayout L;
    L.init(ilog2_u64(n));
    DevBuf bin;
    const void* d_vals = nullptr;
    ZKB_TRY(stage_in(c, vals, n * sizeof(fe), bin, &d_vals));
    DevBuf nodes;
    ZKB_TRY(nodes.alloc(c, L.total_nodes * 64));
    ZKB_TRY(merkle_build_levels(c, (const fe*)d_vals, nullptr, n, L, (uint8_t*)nodes.p));
    ZKB_CUDA(c, cudaMemcpyAsync(c->pinned, (uint8_t*)nodes.p + L.level_off[L.log_n] * 64, 64, cudaMemcpyDeviceToHost, c->stream));
    ZKB_CUDA(c, cudaStreamSynchronize(c->stream));
    memcpy(root, c->pinned, 64);
    return 0;
}

int zkb_merkle_open(zkb_tree* t, const uint64_t* idx, size_t k, uint8_t* paths_out) {
    if (!t || (k && (!idx || !paths_out))) return ZKB_ERR_ARG;
    zkb_ctx* c = t->ctx;
    if (t->n < 2) return set_err(c, ZKB_ERR_INDEX, "open on a 1-leaf tree (the reference recurses forever)");
    for (size_t i = 0; i < k; i++)
        if (idx[i] >= t->n) return set_err(c, ZKB_ERR_INDEX, "cannot open invalid index %llu", (unsigned long long)idx[i]);
    if (k == 0) return 0;
    ZKB_CUDA(c, cudaSetDevice(c->device));
    const size_t path_bytes = (size_t)t->layout.log_n * 64;
    DevBuf buf;
    const size_t idx_bytes = (k * 8 + 63) & ~(size_t)63;       // keep the 16-byte path stores aligned
    ZKB_TRY(buf.alloc(c, idx_bytes + k * path_bytes));
    uint64_t* d_idx = (uint64_t*)buf.p;
    uint8_t* d_out = (uint8_t*)buf.p + idx_bytes;
    ZKB_CUDA(c, cudaMemcpyAsync(d_idx, idx, k * 8, cudaMemcpyHostToDevice, c->stream));
    ZKB_TRY(merkle_open_device(c, t->vals, t->layout, t->nodes, d_idx, k, d_out));
    ZKB_CUDA(c, cudaMemcpyAsync(paths_out, d_out, k * path_bytes, cudaMemcpyDeviceToHost, c->stream));
    ZKB_CUDA(c, cudaStreamSynchronize(c->stream));
    return 0;
}

// stark.rs:546-560: for every opened index push Value(codeword[i]) then Path(open(i)); one batched
// opening + one D2H instead of k tree rebuilds.
static int zkb_merkle_open_ps_body(zkb_tree* t, const uint64_t* idx, size_t k, zkb_ps* ps) {
    if (!t || !ps || (k && !idx)) return ZKB_ERR_ARG;
    zkb_ctx* c = t->ctx;
    if (t->n < 2) return set_err(c, ZKB_ERR_INDEX, "open on a 1-leaf tree (the reference recurses forever)");
    for (size_t i = 0; i < k; i++)
        if (idx[i] >= t->n) return set_err(c, ZKB_ERR_INDEX, "cannot open invalid index %llu", (unsigned long long)idx[i]);
    if (k == 0) return 0;
    ZKB_CUDA(c, cudaSetDevice(c->device));
    const size_t depth = t->layout.log_n, path_bytes = depth * 64;
    if (getenv("ZKB_HOST_ASSEMBLY") == nullptr && depth >= 1 && depth <= ZKB_WIRE_MAX_DEPTH) {
        // objects framed by the opening kernel (k_open_wire): one copy, one append
        const uint64_t rec = 25 + 9 + 72ull * depth;
        std::vector<uint64_t> hidx(idx, idx + k);
        hidx.push_back(0);                                                   // y_off of the single tree
        DevBuf wb;
        const size_t ib = (hidx.size() * 8 + 255) & ~(size_t)255;
        ZKB_TRY(wb.alloc(c, ib + k * rec + 16));
        uint64_t* d_i = (uint64_t*)wb.p;
        uint8_t* d_wire = (uint8_t*)wb.p + ib;
        ZKB_CUDA(c, cudaMemcpyAsync(d_i, hidx.data(), hidx.size() * 8, cudaMemcpyHostToDevice, c->stream));
        ZKB_TRY(merkle_open_wire_batch(c, t->vals, t->layout, t->nodes, d_i, k, 1, 0, 0, d_wire, d_i + k, 1, 0, rec, 0, true));
        uint8_t* dst = ps_body_extend(ps, k * rec);                          // straight into the (pinned) body: no staging copy
        ZKB_CUDA(c, cudaMemcpyAsync(dst, d_wire, k * rec, cudaMemcpyDeviceToHost, c->stream));
        ZKB_CUDA(c, cudaStreamSynchronize(c->stream));
        ps->has_field = true;
        return 0;
    }
    DevBuf buf;
    const size_t idx_bytes = (k * 8 + 63) & ~(size_t)63, val_bytes = (k * 16 + 63) & ~(size_t)63;
    ZKB_TRY(buf.alloc(c, idx_bytes + val_bytes + k * path_bytes));
    uint64_t* d_idx = (uint64_t*)buf.p;
    fe* d_vals = (fe*)((uint8_t*)buf.p + idx_bytes);
    uint8_t* d_paths = (uint8_t*)buf.p + idx_bytes + val_bytes;
    ZKB_CUDA(c, cudaMemcpyAsync(d_idx, idx, k * 8, cudaMemcpyHostToDevice, c->stream));
    { LaunchScope ls(c, K_GATHER); k_gather_vals<<<(unsigned)((k + 127) / 128), 128, 0, c->stream>>>(t->vals, d_idx, (uint32_t)k, d_vals); }
    ZKB_CUDA(c, cudaGetLastError());
    ZKB_TRY(merkle_open_device(c, t->vals, t->layout, t->nodes, d_idx, k, d_paths));
    std::vector<uint8_t> host(val_bytes + k * path_bytes);
    ZKB_CUDA(c, cudaMemcpyAsync(host.data(), d_vals, host.size(), cudaMemcpyDeviceToHost, c->stream));
    ZKB_CUDA(c, cudaStreamSynchronize(c->stream));
    for (size_t i = 0; i < k; i++) {
        zkb_ps_push_value(ps, host.data() + 16 * i);
        zkb_ps_push_path(ps, host.data() + val_bytes + i * path_bytes, depth);
    }
    return 0;
}

int zkb_merkle_open_ps(zkb_tree* t, const uint64_t* idx, size_t k, zkb_ps* ps) {
    if (!t) return ZKB_ERR_ARG;
    ZKB_ABI_GUARD(t->ctx, return zkb_merkle_open_ps_body(t, idx, k, ps);)
}

}  // extern "C"
