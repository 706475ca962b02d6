// merkle.cuh - internal interface of the Merkle engine (shared by merkle.cu and fri.cu).
#pragma once
#include "ctx.hpp"

namespace zkb {

// Layout of a retained tree.  Levels are counted from the leaves: level 0 = leaf hashes
// (n nodes), level log_n = the root.  Only levels >= cut are stored (cut = 5 for trees of
// more than 1024 leaves: 4n bytes instead of 128n); the bottom `cut` levels of an
// authentication path are recomputed at opening time from the 32 leaves around the
// index, which the tree can always reach (it references the committed values).
struct TreeLayout {
    uint32_t log_n = 0;
    uint32_t cut = 0;
    uint64_t level_off[41];      // node index (64-byte units) of level l inside `nodes`
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
};

// Build every stored level of the tree over `n` = 2^log_n values.  If `fold` is non-null
// the values are produced on the fly by folding fold->cw (and written to fold->next),
// otherwise they are read from `vals`.  `nodes` must hold layout.total_nodes * 64 bytes.
int merkle_build_levels(zkb_ctx* c, const fe* vals, const FoldArgs* fold, uint64_t n,
                        const TreeLayout& layout, uint8_t* nodes);

// Authentication paths for k indices: out[k][log_n][64] (device), leaf sibling first.
int merkle_open_device(zkb_ctx* c, const fe* vals, const TreeLayout& layout, const uint8_t* nodes,
                       const uint64_t* d_idx, size_t k, uint8_t* d_out);

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
