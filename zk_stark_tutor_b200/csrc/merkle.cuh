// merkle.cuh - internal interface of the Merkle engine (shared by merkle.cu and fri.cu).
#pragma once
#include "ctx.hpp"

namespace zkb {

// Layout of a retained tree.  Levels are counted from the leaves: level 0 = leaf hashes
// (n nodes), level log_n = the root.  Trees of <= 2^17 leaves store every level.  Larger
// trees store level 3, 6, 9, ... (what the per-thread 8-ary subtree kernels emit) down to the
// first level `top` with <= 2^17 nodes, and every level above `top` (<= 9.2n bytes + 4 MiB
// instead of 128n).  The two unstored levels inside each group of three - and the leaf hashes - of an
// authentication path are recomputed at opening time from the 8 group inputs, which the tree
// can always reach (it references the committed values).
#define ZKB_TREE_LEAF_LOG 17     // trees up to 2^17 leaves: k_leaf1 + the latency-mode tree from the leaf level
#define ZKB_TREE_NODE_LOG 17     // larger trees: 8-ary subtree kernels down to <= 2^17 nodes, then the latency-mode tree
struct TreeLayout {
    uint32_t log_n = 0;
    uint32_t top = 0;            // first level handled by the latency-mode tree k_tree (0: it starts from the leaf hashes)
    uint8_t stored[41];
    uint64_t level_off[41];      // node index (64-byte units) of level l inside `nodes` (stored levels only)
    uint64_t total_nodes = 0;
    void init(uint32_t log_n_);
};

// Parameters of the fused FRI fold (fri.rs:150-159) feeding the leaf hasher.
struct FoldArgs {
    const fe* cw;        // current codeword, length 2*half
    fe* next;            // folded codeword out, length half
    uint64_t half;
    DevPow winv;         // powers of omega_0^-1 (top-level domain)
    uint64_t exp_mul;    // 2^round: (omega_r^-1)^i = (omega_0^-1)^(i*exp_mul)
    fe kk_m;             // alpha / offset_r, Montgomery form
    fe wr_inv_m;         // omega_r^-1, Montgomery form
    // device Fiat-Shamir (keccak.cuh): when non-null the fold constant is READ from device memory (instance b of a batch at
    // kk_dev + b * kk_stride bytes) - the previous layer's tree kernel left it there, no host hop in between
    const uint8_t* kk_dev = nullptr;
    uint64_t kk_stride = 0;
};
struct FsDev;
// Device Fiat-Shamir hook of a tree build: after the root, the top kernel appends Root(root) to the transcript sponge in
// fs (instance b of a batch: fs + b), stores the root in fs->roots[round] and - if want_alpha - leaves the next fold constant in fs->kk_m.
struct FsHook {
    FsDev* fs = nullptr;
    uint32_t round = 0;
    uint32_t want_alpha = 0;
};

// Build every stored level of the tree over `n` = 2^log_n values.  If `fold` is non-null
// the values are produced on the fly by folding fold->cw (and written to fold->next; any n),
// otherwise they are read from `vals`.  `nodes` must hold layout.total_nodes * 64 bytes.
// If `signal` is non-null the top kernel also writes the 64-byte root to signal->host_root and
// then stores signal->seq to *signal->host_flag (both in mapped pinned host memory), so the host
// can pick the root up by polling instead of a D2H copy + stream synchronisation.
struct RootSignal {
    uint8_t* host_root;
    volatile uint32_t* host_flag;
    uint32_t seq;
};
// leaf3_done: the level-3 nodes are already in place (the LDE's last pass hashed them, ntt.cu k_ntt_rr_leaf): start above them.
int merkle_build_levels(zkb_ctx* c, const fe* vals, const FoldArgs* fold, uint64_t n,
                        const TreeLayout& layout, uint8_t* nodes, const RootSignal* signal = nullptr, const FsHook* fs = nullptr,
                        bool leaf3_done = false);

// Batched small trees (n <= 2^ZKB_TREE_LEAF_LOG, i.e. layout.top == 0): `batch` independent instances with
// identical layouts, instance b = blockIdx.y working on buffers offset by b * stride.  One leaf launch and one
// tree launch for the whole batch: at RPSSS sizes (4096-point domain) a single instance cannot fill the GPU.
#define ZKB_MAX_BATCH 4096
struct BatchArgs {
    uint32_t batch = 1;
    uint64_t vals_stride = 0;    // elements between the instances' inputs (vals, or fold->cw)
    uint64_t next_stride = 0;    // elements between the instances' folded outputs (fold->next)
    uint64_t nodes_stride = 0;   // bytes between the instances' node arenas
    const fe* kk_m = nullptr;    // per-instance alpha / offset_r in Montgomery form (overrides FoldArgs::kk_m)
};
int merkle_build_levels_batch(zkb_ctx* c, const fe* vals, const FoldArgs* fold, uint64_t n,
                              const TreeLayout& layout, uint8_t* nodes, const BatchArgs& b, const FsHook* fs = nullptr);
// roots of a batch -> host (batch x 64 bytes), one strided copy; synchronises the stream
int merkle_batch_roots(zkb_ctx* c, const TreeLayout& layout, const uint8_t* nodes, const BatchArgs& b, uint8_t* roots_host);
// Spin until *signal->host_flag == signal->seq (falls back to a stream sync on a CUDA error).
int wait_root(zkb_ctx* c, const RootSignal& signal, uint8_t root_out[64]);

// Authentication paths for k indices: out[k][log_n][64] (device), leaf sibling first.
int merkle_open_device(zkb_ctx* c, const fe* vals, const TreeLayout& layout, const uint8_t* nodes,
                       const uint64_t* d_idx, size_t k, uint8_t* d_out);
// The same for `batch` instances: instance b reads vals + b*vals_stride, nodes + b*nodes_stride, indices
// d_idx + b*k and writes d_out + b*k*log_n*64.
int merkle_open_device_batch(zkb_ctx* c, const fe* vals, const TreeLayout& layout, const uint8_t* nodes,
                             const uint64_t* d_idx, size_t k, uint8_t* d_out, uint32_t batch,
                             uint64_t vals_stride, uint64_t nodes_stride);

// Openings written by the device as finished proof-stream objects (proof_stream_enum.rs:67-127): opening q of instance y goes to
// d_out + d_y_off[y] + base + (q / qdiv) * stride_hi + (q % qdiv) * stride_lo as [Value (25 bytes) if with_value] Path (9 + 72 log_n bytes).
int merkle_open_wire_batch(zkb_ctx* c, const fe* vals, const TreeLayout& layout, const uint8_t* nodes, const uint64_t* d_idx, size_t k,
                           uint32_t batch, uint64_t vals_stride, uint64_t nodes_stride, uint8_t* d_out, const uint64_t* d_y_off,
                           uint32_t qdiv, uint64_t base, uint64_t stride_hi, uint64_t stride_lo, bool with_value);
// Leafs objects (57 bytes each) of one FRI round: object s of instance y at d_out + d_y_off[y] + base + 57 s
int fri_leafs_wire_batch(zkb_ctx* c, const fe* cur, uint64_t cur_stride, const fe* nxt, uint64_t nxt_stride, uint64_t half,
                         const uint64_t* d_idx, size_t k, uint32_t batch, uint8_t* d_out, const uint64_t* d_y_off, uint64_t base);

}  // namespace zkb

struct zkb_tree {
    zkb_ctx* ctx = nullptr;
    uint64_t n = 0;
    zkb::TreeLayout layout;
    uint8_t* nodes = nullptr;        // device
    bool owns_nodes = true;
    const zkb::fe* vals = nullptr;   // device
    void* owned_vals = nullptr;      // set when the values were staged from the host
    uint8_t root[64];
};
