// ntt.cuh - internal interface of the NTT engine (ntt.cu), used by fri.cu for the fused
// LDE -> FRI pipeline entry points.
#pragma once
#include "ctx.hpp"

namespace zkb {

struct NttOpts {
    bool has_scale = false;   // x_i *= scale_base^i on load (coset LDE, polynomial.rs:109-121)
    fe scale_base;
    bool inverse = false;     // use root^-1 and multiply by n^-1 (ntt.rs:51-68)
};

// d_in / d_out are device pointers; 2^log_n is the transform length; n_in <= 2^log_n values
// are read per column (the rest are zero); `batch` columns at the given element strides.
int ntt_exec(zkb_ctx* c, fe root, const fe* d_in, size_t n_in, size_t in_stride, fe* d_out,
             size_t out_stride, size_t batch, uint32_t log_n, const NttOpts& o);

}  // namespace zkb
