// ntt.cuh - internal interface of the NTT engine (ntt.cu), used by fri.cu for the fused
// LDE -> FRI pipeline entry points.
#pragma once
#include "ctx.hpp"

namespace zkb {

#define ZKB_NTT_MAX_PEERS 16
// The four-step NTT's exchange fused into the store of the last local pass: output element k (k < 2^log_n) is
// multiplied by oscale_base^k and written to peer[k >> log_blk] + peer_row + (k & (2^log_blk - 1)).
struct NttExchange {
    fe oscale_base;
    uint32_t n_peers = 0, log_blk = 0;
    uint64_t peer_row = 0;
    fe* peer[ZKB_NTT_MAX_PEERS];
};

struct NttOpts {
    bool has_scale = false;   // x_i *= scale_base^i on load (coset LDE, polynomial.rs:109-121)
    fe scale_base;
    bool inverse = false;     // use root^-1 and multiply by n^-1 (ntt.rs:51-68)
    bool no_post = false;     // inverse without the n^-1 factor (the caller applies it later)
    const NttExchange* exchange = nullptr;   // batch == 1, 2^13 <= n only
    // batch == 1, ntt_can_fuse_leaves(log_n): the last pass also hashes the output as Merkle leaves (8 leaf hashes + 7 nodes per
    // group of 8 consecutive values) and writes the level-3 nodes here (n / 8 x 64 bytes) - what k_leaf8<false> would produce
    uint8_t* leaf3_out = nullptr;
};
bool ntt_can_fuse_leaves(uint32_t log_n);

// d_in / d_out are device pointers; 2^log_n is the transform length; n_in <= 2^log_n values
// are read per column (the rest are zero); `batch` columns at the given element strides.
int ntt_exec(zkb_ctx* c, fe root, const fe* d_in, size_t n_in, size_t in_stride, fe* d_out,
             size_t out_stride, size_t batch, uint32_t log_n, const NttOpts& o);

// cross-rank stage of the four-step NTT (see ntt4.cu) and the unfused twiddle + scatter for small local transforms
int ntt_cross_exec(zkb_ctx* c, const fe& root_g, uint32_t log_g, const fe* d_in, fe* d_out, uint64_t blk, const fe* post);
int ntt_twiddle_scatter(zkb_ctx* c, const fe* d_in, uint64_t n, const NttExchange& x);

}  // namespace zkb
