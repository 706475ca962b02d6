// merkle.cu - BLAKE2b-512 Merkle commitment over field elements.
//
// Replaces the bodies of (reference file:line):
//   MerkleRoot::commit  src/merkle_root.rs:7-32   (leaf = H(decimal string), node = H(l||r))
//   MerkleRoot::open    src/merkle_root.rs:34-66  (the reference re-hashes the whole tree per call)
//   MerkleRoot::verify  src/merkle_root.rs:69-95  (host, hosthash.cpp)
//
// Kernel structure.  The work is ALU-pipe bound (2,128 32-bit ops per compression), so
// the design goals are full warps in every hashing step and no HBM round trip between
// the levels of a subtree:
//   k_leaf_tile   1024 leaves per CTA.  Each thread hashes 4 consecutive leaves and the 3
//                 nodes above them in registers (7 compressions, no synchronisation), the
//                 256 level-2 nodes then go through shared memory (128-bit accesses, XOR
//                 swizzled so both the 64-byte stores and the 128-byte loads are
//                 conflict-free) for levels 3, 4, 5 with 4, 2, 1 full warps.  Only the 32
//                 level-5 nodes leave the SM.  With FoldArgs the 4 values are produced by
//                 the FRI split-and-fold of the previous layer instead of being loaded,
//                 and are written out as the next codeword ("fold fused with the next
//                 round's leaf hashing").
//   k_node_tile   the same shape over 1024 stored nodes -> 5 more levels per launch.
//   k_small       one CTA finishes any tree (or tree top) of <= 1024 inputs.
#include <string.h>
#include "merkle.cuh"
#include "blake2b.cuh"

namespace zkb {

void TreeLayout::init(uint32_t log_n_) {
    log_n = log_n_;
    cut = log_n > 10 ? 5 : 0;
    uint64_t off = 0;
    for (uint32_t l = 0; l <= 40; l++) level_off[l] = 0;
    for (uint32_t l = cut; l <= log_n; l++) {
        level_off[l] = off;
        off += 1ull << (log_n - l);
    }
    total_nodes = off;
}

// ---- compression wrappers.  __noinline__ keeps ONE copy of each 2.2k-instruction body per
// kernel (10 call sites in k_leaf_tile would otherwise be 350 KB of code).
__device__ __noinline__ void b2_leaf_call(const fe* a, uint64_t* h) {
    uint64_t out[8];
    blake2b_leaf(*a, out);
#pragma unroll
    for (int i = 0; i < 8; i++) h[i] = out[i];
}
__device__ __noinline__ void b2_block_call(const uint64_t* m_in, uint64_t* h) {
    uint64_t m[16], out[8];
#pragma unroll
    for (int i = 0; i < 16; i++) m[i] = m_in[i];
    blake2b_compress_1block(m, 128, out);
#pragma unroll
    for (int i = 0; i < 8; i++) h[i] = out[i];
}

// ---- swizzled shared-memory digest store: node i = 4 x 16-byte chunks; a 128-byte line
// holds nodes 2j, 2j+1; chunk position p in the line is stored at p ^ (line & 7).
__device__ __forceinline__ void sm_store_digest(uint4* reg, uint32_t i, const uint64_t* h) {
    uint32_t line = i >> 1, half = (i & 1u) << 2;
#pragma unroll
    for (uint32_t cidx = 0; cidx < 4; cidx++) {
        uint32_t pos = (half | cidx) ^ (line & 7u);
        reg[line * 8 + pos] = make_uint4((uint32_t)h[2 * cidx], (uint32_t)(h[2 * cidx] >> 32),
                                         (uint32_t)h[2 * cidx + 1], (uint32_t)(h[2 * cidx + 1] >> 32));
    }
}
__device__ __forceinline__ void sm_load_pair(const uint4* reg, uint32_t line, uint64_t* m) {
#pragma unroll
    for (uint32_t pidx = 0; pidx < 8; pidx++) {
        uint4 x = reg[line * 8 + (pidx ^ (line & 7u))];
        m[2 * pidx] = ((uint64_t)x.y << 32) | x.x;
        m[2 * pidx + 1] = ((uint64_t)x.w << 32) | x.z;
    }
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

// Shared tail of both tile kernels: 256 nodes (one per thread, `h`) at relative level 0 ->
// relative levels 1, 2, 3 (128, 64, 32 nodes).  lvl_ptr[r] = where to store relative level
// r+1 in global memory (nullptr = not stored); tile = CTA index.
__device__ __forceinline__ void tile_tail(uint4* regA, uint4* regB, uint32_t tid, uint64_t tile,
                                          uint64_t* h, uint8_t* out1, uint8_t* out2, uint8_t* out3) {
    uint64_t m[16];
    sm_store_digest(regA, tid, h);
    __syncthreads();
    if (tid < 128) {
        sm_load_pair(regA, tid, m);
        b2_block_call(m, h);
        sm_store_digest(regB, tid, h);
        if (out1) g_store_digest(out1, tile * 128 + tid, h);
    }
    __syncthreads();
    if (tid < 64) {
        sm_load_pair(regB, tid, m);
        b2_block_call(m, h);
        sm_store_digest(regA, tid, h);
        if (out2) g_store_digest(out2, tile * 64 + tid, h);
    }
    __syncthreads();
    if (tid < 32) {
        sm_load_pair(regA, tid, m);
        b2_block_call(m, h);
        g_store_digest(out3, tile * 32 + tid, h);
    }
}

// 1024 leaves per CTA -> 32 level-5 nodes.
template <bool FOLD>
__global__ void __launch_bounds__(256, 2) k_leaf_tile(const fe* __restrict__ vals, FoldArgs f, uint8_t* __restrict__ out5) {
    __shared__ uint4 regA[256 * 4];
    __shared__ uint4 regB[128 * 4];
    const uint32_t tid = threadIdx.x;
    const uint64_t tile = blockIdx.x;
    const uint64_t i0 = tile * 1024 + (uint64_t)tid * 4;
    fe v[4];
    if (FOLD) {
        fe k_m = fe_montmul(f.kk_m, pow2lvl_m(f.winv, i0 * f.exp_mul));   // alpha/(offset*omega^i0) * R
#pragma unroll
        for (int t = 0; t < 4; t++) {
            fe a = fe_ldg(f.cw + i0 + t), b = fe_ldg(f.cw + f.half + i0 + t);
            fe s = fe_add(a, b), d = fe_sub(a, b);
            v[t] = fe_half(fe_add(s, fe_montmul(k_m, d)));
            fe_store(f.next + i0 + t, v[t]);
            k_m = fe_montmul(k_m, f.wr_inv_m);
        }
    } else {
#pragma unroll
        for (int t = 0; t < 4; t++) v[t] = fe_ldg(vals + i0 + t);
    }
    uint64_t m[16], h[8];
    b2_leaf_call(&v[0], m);
    b2_leaf_call(&v[1], m + 8);
    b2_block_call(m, h);                 // level-1 node over leaves 0,1
    b2_leaf_call(&v[2], m);
    b2_leaf_call(&v[3], m + 8);
    b2_block_call(m, m + 8);             // level-1 node over leaves 2,3 (input copied before the write)
#pragma unroll
    for (int i = 0; i < 8; i++) m[i] = h[i];
    b2_block_call(m, h);                 // level-2 node, index tile*256 + tid
    tile_tail(regA, regB, tid, tile, h, nullptr, nullptr, out5);
}

// 1024 stored nodes (relative level 0) per CTA -> relative levels 1..5, all stored.
__global__ void __launch_bounds__(256, 2) k_node_tile(const uint8_t* __restrict__ in, uint8_t* out1, uint8_t* out2,
                                                      uint8_t* out3, uint8_t* out4, uint8_t* out5) {
    __shared__ uint4 regA[256 * 4];
    __shared__ uint4 regB[128 * 4];
    const uint32_t tid = threadIdx.x;
    const uint64_t tile = blockIdx.x;
    const uint64_t i0 = tile * 1024 + (uint64_t)tid * 4;
    uint64_t m[16], h[8], h2[8];
    g_load_digest(in, i0, m); g_load_digest(in, i0 + 1, m + 8);
    b2_block_call(m, h);
    g_store_digest(out1, tile * 512 + tid * 2, h);
    g_load_digest(in, i0 + 2, m); g_load_digest(in, i0 + 3, m + 8);
    b2_block_call(m, h2);
    g_store_digest(out1, tile * 512 + tid * 2 + 1, h2);
#pragma unroll
    for (int i = 0; i < 8; i++) { m[i] = h[i]; m[8 + i] = h2[i]; }
    b2_block_call(m, h);
    g_store_digest(out2, tile * 256 + tid, h);
    tile_tail(regA, regB, tid, tile, h, out3, out4, out5);
}

// One CTA: `count` (power of two, <= 1024) leaves or nodes -> every level up to the root.
// level_out[r] = global destination of relative level r (r = 0 is the leaf-hash level and
// is only written for leaf input).  Dynamic shared memory: 1024 + 512 digests.
struct SmallArgs {
    const fe* vals;          // leaf input (or nullptr)
    const uint8_t* nodes_in; // node input (or nullptr)
    uint32_t count;
    uint8_t* level_out[12];
};
__global__ void __launch_bounds__(256) k_small(SmallArgs a) {
    extern __shared__ uint4 dyn[];
    uint4* cur = dyn;
    uint4* nxt = dyn + 1024 * 4;
    const uint32_t tid = threadIdx.x;
    uint64_t m[16], h[8];
    for (uint32_t i = tid; i < a.count; i += 256) {
        if (a.vals) {
            fe v = fe_ldg(a.vals + i);
            b2_leaf_call(&v, h);
            g_store_digest(a.level_out[0], i, h);
        } else {
            g_load_digest(a.nodes_in, i, h);
        }
        sm_store_digest(cur, i, h);
    }
    uint32_t level = 1;
    for (uint32_t cnt = a.count >> 1; cnt >= 1; cnt >>= 1, level++) {
        __syncthreads();
        for (uint32_t j = tid; j < cnt; j += 256) {
            sm_load_pair(cur, j, m);
            b2_block_call(m, h);
            sm_store_digest(nxt, j, h);
            g_store_digest(a.level_out[level], j, h);
        }
        uint4* t = cur; cur = nxt; nxt = t;
    }
}

// Authentication paths.  One warp per opened index.
struct OpenArgs {
    const fe* vals;
    const uint8_t* nodes;
    TreeLayout layout;
    const uint64_t* idx;
    uint32_t k;
    uint8_t* out;            // k * log_n * 64 bytes
};
__global__ void __launch_bounds__(128) k_open(OpenArgs a) {
    __shared__ uint64_t dig[4][48][8];          // per warp: 32 + 16 digests
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t q = blockIdx.x * 4 + warp;
    if (q >= a.k) return;
    const uint64_t idx = a.idx[q];
    const uint32_t log_n = a.layout.log_n, cut = a.layout.cut;
    uint4* out = reinterpret_cast<uint4*>(a.out + (uint64_t)q * log_n * 64);
    if (cut > 0) {
        // recompute the 32-leaf subtree around idx (cut == 5)
        uint64_t (*A)[8] = dig[warp];
        uint64_t (*Bq)[8] = dig[warp] + 32;
        const uint64_t base = idx & ~31ull;
        {
            fe v = fe_ldg(a.vals + base + lane);
            uint64_t h[8];
            b2_leaf_call(&v, h);
            for (int i = 0; i < 8; i++) A[lane][i] = h[i];
        }
        __syncwarp();
        uint32_t pos = (uint32_t)(idx & 31);
        for (uint32_t t = 0; t < cut; t++) {
            // level t digests are in A (32 >> t of them); emit the sibling, then hash up into Bq
            const uint4* sib = reinterpret_cast<const uint4*>(A[(pos >> t) ^ 1]);
            if (lane < 4) out[t * 4 + lane] = sib[lane];
            uint32_t cnt = 32u >> (t + 1);
            if (lane < cnt && t + 1 < cut) {
                uint64_t m[16], h[8];
                for (int i = 0; i < 8; i++) { m[i] = A[2 * lane][i]; m[8 + i] = A[2 * lane + 1][i]; }
                b2_block_call(m, h);
                for (int i = 0; i < 8; i++) Bq[lane][i] = h[i];
            }
            __syncwarp();
            uint64_t (*tmp)[8] = A; A = Bq; Bq = tmp;
        }
    }
    for (uint32_t l = cut; l < log_n; l++) {
        const uint4* sib = reinterpret_cast<const uint4*>(a.nodes + (a.layout.level_off[l] + ((idx >> l) ^ 1)) * 64);
        if (lane < 4) out[l * 4 + lane] = __ldg(sib + lane);
    }
}

static int launch_small(zkb_ctx* c, const SmallArgs& a) {
    static bool attr_set = false;
    if (!attr_set) {
        ZKB_CUDA(c, cudaFuncSetAttribute(k_small, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
        attr_set = true;
    }
    { LaunchScope ls(c, K_MERKLE_SMALL); k_small<<<1, 256, 96 * 1024, c->stream>>>(a); }
    ZKB_CUDA(c, cudaGetLastError());
    return 0;
}

int merkle_build_levels(zkb_ctx* c, const fe* vals, const FoldArgs* fold, uint64_t n,
                        const TreeLayout& L, uint8_t* nodes) {
    const uint32_t log_n = L.log_n;
    if (L.cut == 0) {
        // small tree: values must already exist (the caller folds separately for small layers)
        if (fold) return set_err(c, ZKB_ERR_ARG, "internal: fused fold needs n > 1024");
        SmallArgs a;
        memset(&a, 0, sizeof(a));
        a.vals = vals; a.count = (uint32_t)n;
        for (uint32_t l = 0; l <= log_n; l++) a.level_out[l] = nodes + L.level_off[l] * 64;
        return launch_small(c, a);
    }
    uint8_t* lvl5 = nodes + L.level_off[5] * 64;
    if (fold) {
        LaunchScope ls(c, K_FOLD_LEAF_TILE);
        k_leaf_tile<true><<<(unsigned)(n >> 10), 256, 0, c->stream>>>(nullptr, *fold, lvl5);
    } else {
        FoldArgs dummy;
        memset(&dummy, 0, sizeof(dummy));
        LaunchScope ls(c, K_LEAF_TILE);
        k_leaf_tile<false><<<(unsigned)(n >> 10), 256, 0, c->stream>>>(vals, dummy, lvl5);
    }
    ZKB_CUDA(c, cudaGetLastError());
    uint32_t level = 5;
    uint64_t m = n >> 5;
    while (m > 1024) {
        const uint8_t* in = nodes + L.level_off[level] * 64;
        {
            LaunchScope ls(c, K_NODE_TILE);
            k_node_tile<<<(unsigned)(m >> 10), 256, 0, c->stream>>>(in,
                nodes + L.level_off[level + 1] * 64, nodes + L.level_off[level + 2] * 64,
                nodes + L.level_off[level + 3] * 64, nodes + L.level_off[level + 4] * 64,
                nodes + L.level_off[level + 5] * 64);
        }
        ZKB_CUDA(c, cudaGetLastError());
        level += 5;
        m >>= 5;
    }
    SmallArgs a;
    memset(&a, 0, sizeof(a));
    a.nodes_in = nodes + L.level_off[level] * 64;
    a.count = (uint32_t)m;
    for (uint32_t r = 1; level + r <= log_n; r++) a.level_out[r] = nodes + L.level_off[level + r] * 64;
    if (m > 1) ZKB_TRY(launch_small(c, a));
    return 0;
}

int merkle_open_device(zkb_ctx* c, const fe* vals, const TreeLayout& layout, const uint8_t* nodes,
                       const uint64_t* d_idx, size_t k, uint8_t* d_out) {
    if (k == 0) return 0;
    OpenArgs a;
    a.vals = vals; a.nodes = nodes; a.layout = layout; a.idx = d_idx; a.k = (uint32_t)k; a.out = d_out;
    { LaunchScope ls(c, K_OPEN); k_open<<<(unsigned)((k + 3) / 4), 128, 0, c->stream>>>(a); }
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

}  // extern "C"
