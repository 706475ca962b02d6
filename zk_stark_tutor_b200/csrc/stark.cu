// stark.cu - the front of Stark::prove for batches of trace columns: interpolation of the (randomized) trace and its
// evaluation on the FRI coset, on the device.
//
// Reference: stark.rs:303-326 interpolates every register column over the trace domain omicron^0 .. omicron^(L-1) with
// fast_interpolate_domain (ntt_arithmetics.rs:172-237: divide-and-conquer over arbitrary points, O(L log^2 L) products and
// schoolbook remainders); the polynomials then go through boundary quotients and fast_coset_evaluate (stark.rs:331-386).
// Here: the trace domain is a PREFIX of the order-N subgroup <omicron>, so the unique interpolant of degree < L is
//       p = iNTT_N(values || 0...) mod Z,       Z(x) = prod_{i<L} (x - omicron^i)
// (any degree < N polynomial with the right values on the prefix is congruent to p modulo Z).  The remainder is a fast
// division: with m = N - L quotient coefficients, rev(quo) = rev(p~) * rev(Z)^-1 mod x^m, rem = p~ - Z*quo.  Z, the
// power-series inverse of its reversal and their length-2N transforms depend on (L, N) only and are cached per context,
// so a column costs one iNTT_N, three NTT_2N, two pointwise products and two tiny copy kernels - and every launch carries
// the whole batch of columns.  The coefficients then go straight into the coset LDE; nothing returns to the host.
#include <string.h>
#include <vector>
#include "ctx.hpp"
#include "ntt.cuh"
#include "prefix.cuh"

namespace zkb {

struct PrefixTables {
    uint64_t L, N;
    fe omicron;
    fe* dev = nullptr;          // ginv_hat (2N) | z_hat (2N): length-2N transforms of rev(Z)^-1 mod x^m and of Z
    fe root2n;                  // the primitive 2N-th root they were transformed with
};
static void prefix_tables_destroy(void* p) {
    PrefixTables* t = (PrefixTables*)p;
    if (t->dev) cudaFree(t->dev);
    delete t;
}

// out[b][k] = in[b][top - k], k < count   (coefficient reversal of the first top+1 entries, truncated)
__global__ void k_reverse_take(const fe* __restrict__ in, uint64_t in_stride, uint64_t top, fe* __restrict__ out, uint64_t out_stride, uint64_t count) {
    const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count) return;
    const uint64_t b = blockIdx.y;
    fe_store(out + b * out_stride + k, fe_ldg(in + b * in_stride + top - k));
}
// a[b][k] *= shared[k]  (one operand of the product is the same for every column: the cached transform)
__global__ void k_mul_shared(fe* __restrict__ a, uint64_t stride, const fe* __restrict__ shared, uint64_t count) {
    const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count) return;
    fe* p = a + (uint64_t)blockIdx.y * stride + k;
    fe_store(p, fe_montmul(fe_to_mont(fe_ldg(p)), fe_ldg(shared + k)));
}
// out[b][k] = a[b][k] - s[b][k], k < count
__global__ void k_sub_take(const fe* __restrict__ a, uint64_t a_stride, const fe* __restrict__ s, uint64_t s_stride, fe* __restrict__ out,
                           uint64_t out_stride, uint64_t count) {
    const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count) return;
    const uint64_t b = blockIdx.y;
    fe_store(out + b * out_stride + k, fe_sub(fe_ldg(a + b * a_stride + k), fe_ldg(s + b * s_stride + k)));
}
// degree[b] = max k with coeffs[b][k] != 0 (atomicMax over k + 1; 0 = the zero polynomial)
__global__ void k_degree(const fe* __restrict__ coeffs, uint64_t stride, uint64_t n, unsigned long long* __restrict__ deg_plus_1) {
    const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    if (!fe_is_zero(fe_ldg(coeffs + (uint64_t)blockIdx.y * stride + k))) atomicMax(deg_plus_1 + blockIdx.y, (unsigned long long)(k + 1));
}

static fe primitive_root_h(uint64_t n) {
    uint8_t b[16];
    zkb_primitive_nth_root(n, b);
    return h_load(b);
}

// Host, once per (L, N): Z and rev(Z)^-1 mod x^m (prefix.cuh; L = 284, m = 740 at RPSSS parameters: ~0.3 M field
// multiplications), then both are transformed on the device.
static int get_prefix_tables(zkb_ctx* c, const fe& omicron, uint64_t N, uint64_t L, PrefixTables** out) {
    for (auto& at : c->attachments) {
        if (at.destroy != prefix_tables_destroy) continue;
        PrefixTables* t = (PrefixTables*)at.p;
        if (t->L == L && t->N == N && fe_eq(t->omicron, omicron)) { *out = t; return 0; }
    }
    const uint64_t m = N - L;
    const std::vector<fe> Z = prefix_zerofier(omicron, L);
    const std::vector<fe> g = reversed_series_inverse(Z, m);
    PrefixTables* t = new PrefixTables();
    t->L = L; t->N = N; t->omicron = omicron;
    t->root2n = primitive_root_h(2 * N);
    cudaError_t e = cudaMalloc(&t->dev, 4 * N * sizeof(fe));
    if (e != cudaSuccess) { delete t; return set_err(c, ZKB_ERR_CUDA, "prefix tables: cudaMalloc failed: %s", cudaGetErrorString(e)); }
    DevBuf stage;
    int rc = stage.alloc(c, 4 * N * sizeof(fe));
    if (rc) { prefix_tables_destroy(t); return rc; }
    std::vector<fe> host(4 * N, fe_zero());
    for (uint64_t k = 0; k < m; k++) host[k] = g[k];
    for (uint64_t k = 0; k <= L; k++) host[2 * N + k] = Z[k];
    if (cudaMemcpyAsync(stage.p, host.data(), 4 * N * sizeof(fe), cudaMemcpyHostToDevice, c->stream) != cudaSuccess) {
        prefix_tables_destroy(t);
        return set_err(c, ZKB_ERR_CUDA, "prefix tables: upload failed");
    }
    NttOpts o;
    rc = ntt_exec(c, t->root2n, (const fe*)stage.p, 2 * N, 2 * N, t->dev, 2 * N, 2, ilog2_u64(2 * N), o);
    if (rc == 0 && ctx_stream_sync(c) != cudaSuccess) rc = set_err(c, ZKB_ERR_CUDA, "prefix tables: transform failed");
    if (rc) { prefix_tables_destroy(t); return rc; }
    c->attachments.push_back({t, prefix_tables_destroy});
    *out = t;
    return 0;
}

}  // namespace zkb

using namespace zkb;

extern "C" {

int zkb_trace_lde_batch(zkb_ctx* c, const uint8_t omicron_b[16], uint64_t omicron_order, uint64_t length, const uint8_t omega_b[16],
                        uint64_t order, const uint8_t offset_b[16], const void* values, size_t stride, size_t batch, void* out,
                        size_t out_stride, void* coeffs_out) {
    if (!c || !omicron_b || !omega_b || !offset_b || !values || !out || batch == 0) return ZKB_ERR_ARG;
    const uint64_t N = omicron_order, L = length, n = order;
    if (N < 2 || (N & (N - 1)) || n < N || (n & (n - 1))) return set_err(c, ZKB_ERR_ARG, "trace_lde: domain lengths must be powers of two, order >= omicron_order");
    if (L < 2 || L > N) return set_err(c, ZKB_ERR_TOO_LONG, "trace_lde: trace length %llu outside 2..%llu", (unsigned long long)L, (unsigned long long)N);
    if (stride < L || out_stride < n) return set_err(c, ZKB_ERR_ARG, "trace_lde: strides shorter than the columns");
    if (!is_device_ptr(out) || (coeffs_out && !is_device_ptr(coeffs_out))) return set_err(c, ZKB_ERR_ARG, "trace_lde: outputs are device buffers");
    if (batch > 65535) return set_err(c, ZKB_ERR_ARG, "trace_lde: at most 65535 columns per call");
    ZKB_CUDA(c, cudaSetDevice(c->device));
    const fe omicron = h_load(omicron_b), omega = h_load(omega_b), offset = h_load(offset_b);
    if (!fe_eq(h_pow(omicron, N), fe_from_u32(1)) || fe_eq(h_pow(omicron, N / 2), fe_from_u32(1)))
        return set_err(c, ZKB_ERR_ROOT_ORDER, "trace_lde: omicron is not a primitive root of order %llu", (unsigned long long)N);
    const uint64_t m = N - L;
    PrefixTables* tab = nullptr;
    if (m) ZKB_TRY(get_prefix_tables(c, omicron, N, L, &tab));

    DevBuf in, work;
    const void* d_vals = nullptr;
    const size_t in_elems = stride * (batch - 1) + L;
    ZKB_TRY(stage_in(c, values, in_elems * sizeof(fe), in, &d_vals));
    // work: P (batch x N) | A (batch x 2N) | B (batch x 2N) | C (batch x L, the coefficients; or the caller's coeffs_out)
    const size_t eP = batch * N, eA = batch * 2 * N, eC = coeffs_out ? 0 : batch * L;
    ZKB_TRY(work.alloc(c, (eP + 2 * eA + eC) * sizeof(fe)));
    fe* P = (fe*)work.p; fe* A = P + eP; fe* B = A + eA; fe* C = coeffs_out ? (fe*)coeffs_out : B + eA;
    NttOpts inv; inv.inverse = true;
    NttOpts fwd;
    const uint32_t logN = ilog2_u64(N), log2N = logN + 1;
    const dim3 blk(128);
    auto grid = [&](uint64_t count) { return dim3((unsigned)((count + 127) / 128), (unsigned)batch); };
    ZKB_TRY(ntt_exec(c, omicron, (const fe*)d_vals, L, stride, P, N, batch, logN, inv));          // p~ = iNTT_N(values || 0)
    if (m == 0) {
        ZKB_CUDA(c, cudaMemcpyAsync(C, P, eP * sizeof(fe), cudaMemcpyDeviceToDevice, c->stream));
    } else {
        { LaunchScope ls(c, K_ELEMENTWISE); k_reverse_take<<<grid(m), blk, 0, c->stream>>>(P, N, N - 1, A, 2 * N, m); }       // rev(p~) mod x^m
        ZKB_TRY(ntt_exec(c, tab->root2n, A, m, 2 * N, B, 2 * N, batch, log2N, fwd));
        { LaunchScope ls(c, K_ELEMENTWISE); k_mul_shared<<<grid(2 * N), blk, 0, c->stream>>>(B, 2 * N, tab->dev, 2 * N); }     // * rev(Z)^-1
        ZKB_TRY(ntt_exec(c, tab->root2n, B, 2 * N, 2 * N, A, 2 * N, batch, log2N, inv));            // first m entries: rev(quo)
        { LaunchScope ls(c, K_ELEMENTWISE); k_reverse_take<<<grid(m), blk, 0, c->stream>>>(A, 2 * N, m - 1, B, 2 * N, m); }   // quo
        ZKB_TRY(ntt_exec(c, tab->root2n, B, m, 2 * N, A, 2 * N, batch, log2N, fwd));
        { LaunchScope ls(c, K_ELEMENTWISE); k_mul_shared<<<grid(2 * N), blk, 0, c->stream>>>(A, 2 * N, tab->dev + 2 * N, 2 * N); }   // * Z
        ZKB_TRY(ntt_exec(c, tab->root2n, A, 2 * N, 2 * N, B, 2 * N, batch, log2N, inv));            // Z * quo (degree < N)
        { LaunchScope ls(c, K_ELEMENTWISE); k_sub_take<<<grid(L), blk, 0, c->stream>>>(P, N, B, 2 * N, C, L, L); }            // rem = p~ - Z*quo
    }
    ZKB_CUDA(c, cudaGetLastError());
    NttOpts lde;
    lde.has_scale = true;
    lde.scale_base = offset;
    ZKB_TRY(ntt_exec(c, omega, C, m ? L : N, m ? L : N, (fe*)out, out_stride, batch, ilog2_u64(n), lde));   // stark.rs:373-378 for the trace itself
    if (in.p) ZKB_CUDA(c, ctx_stream_sync(c));     // the caller's host buffer has been consumed
    return 0;
}

int zkb_coset_degree_batch(zkb_ctx* c, const uint8_t omega_b[16], const void* codewords, size_t n, size_t stride, size_t batch, int64_t* degrees_out) {
    if (!c || !omega_b || !codewords || !degrees_out || batch == 0) return ZKB_ERR_ARG;
    if (n < 2 || (n & (n - 1)) || stride < n) return set_err(c, ZKB_ERR_ARG, "coset_degree: bad length / stride");
    if (!is_device_ptr(codewords)) return set_err(c, ZKB_ERR_ARG, "coset_degree takes device codewords");
    if (batch > 65535) return set_err(c, ZKB_ERR_ARG, "coset_degree: at most 65535 codewords per call");
    ZKB_CUDA(c, cudaSetDevice(c->device));
    DevBuf work;
    ZKB_TRY(work.alloc(c, batch * n * sizeof(fe) + batch * 8));
    fe* coeffs = (fe*)work.p;
    unsigned long long* deg = (unsigned long long*)(coeffs + batch * n);
    ZKB_CUDA(c, cudaMemsetAsync(deg, 0, batch * 8, c->stream));
    NttOpts inv; inv.inverse = true;
    // values of p on offset*<omega> -> coefficients of p(offset*x) = c_i * offset^i: same degree as p (offset != 0)
    ZKB_TRY(ntt_exec(c, h_load(omega_b), (const fe*)codewords, n, stride, coeffs, n, batch, ilog2_u64(n), inv));
    { LaunchScope ls(c, K_ELEMENTWISE); k_degree<<<dim3((unsigned)((n + 127) / 128), (unsigned)batch), 128, 0, c->stream>>>(coeffs, n, n, deg); }
    ZKB_CUDA(c, cudaGetLastError());
    std::vector<unsigned long long> host(batch);
    ZKB_CUDA(c, cudaMemcpyAsync(host.data(), deg, batch * 8, cudaMemcpyDeviceToHost, c->stream));
    ZKB_CUDA(c, ctx_stream_sync(c));
    for (size_t b = 0; b < batch; b++) degrees_out[b] = (int64_t)host[b] - 1;     // -1: the zero polynomial (Polynomial::degree() == None)
    return 0;
}

}  // extern "C"
