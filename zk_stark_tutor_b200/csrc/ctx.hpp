// ctx.hpp - library context: one per (process, GPU).  Owns the stream, a scratch
// arena, the cache of power tables (twiddles) and the last error string.
// Single-threaded callers per context, like the reference (no globals,
// no interior mutability in the reference; the C ABI keeps state in explicit handles).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>
#include <memory>
#include <new>
#include "fe128.cuh"
#include "../../include/zkb200.h"

namespace zkb {

// Two-level table of powers of `base` covering exponents [0, 2^log_n):
//   lo[i] = base^i * R        i < 2^lo_bits
//   hi[j] = base^(j<<lo_bits) * R
// so base^e * R = montmul(hi[e >> lo_bits], lo[e & mask]).  128 KiB for log_n = 24.
struct PowTable {
    fe base;            // canonical
    uint32_t log_n;
    uint32_t lo_bits;
    fe* lo = nullptr;   // device
    fe* hi = nullptr;   // device
    uint64_t stamp = 0;
};

struct DevPow {         // what kernels take by value
    const fe* lo;
    const fe* hi;
    uint32_t lo_bits;
};

}  // namespace zkb

struct zkb_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int sm_count = 148;
    std::string err;
    // scratch arena (grown on demand, never shrunk)
    void* scratch = nullptr;
    size_t scratch_bytes = 0;
    // pinned staging for small host<->device traffic (roots, challenges)
    uint8_t* pinned = nullptr;
    size_t pinned_bytes = 0;
    std::vector<std::unique_ptr<zkb::PowTable>> pow_tables;
    uint64_t clock = 0;
    void* host_scratch[3] = {nullptr, nullptr, nullptr};   // grow-only pinned buffers for large D2H results (batch paths)
    size_t host_scratch_bytes[3] = {0, 0, 0};
    size_t assembly_threads = 16;    // host threads per batched call for proof-stream assembly (zkb_ctx_assembly_threads)
    bool zero_copy_inputs = true;    // pinned host LDE inputs are read in place by the first NTT pass (zkb_ctx_zero_copy_inputs)
    uint32_t* tree_bars = nullptr;   // k_tree's arrival counter (device; zero between launches)
    bool blocking_sync = false;      // wait for the stream by polling + sleeping instead of spinning (zkb_ctx_blocking_sync)
    uint32_t tail_threads = 512;     // threads per CTA of the persistent FRI tail kernel (zkb_ctx_tail_threads)
    void* fs_dev = nullptr;          // device Fiat-Shamir contexts (keccak.cuh FsDev x fs_dev_count) + the tail kernel's barrier word
    size_t fs_dev_count = 0;
    uint32_t root_seq = 0;   // sequence number of the last root signalled through pinned memory
    uint64_t launches = 0;   // kernels launched through this context (bench "gpu_launches")
    // optional per-kernel-class device timing (CUDA events on the launching stream)
    bool profiling = false;
    struct ProfRec { int id; cudaEvent_t a, b; };
    std::vector<ProfRec> prof_recs;
    std::vector<cudaEvent_t> prof_pool;
    double prof_ms[16] = {0};
    uint64_t prof_count[16] = {0};
    // caches owned by other translation units (e.g. the prefix-interpolation tables of stark.cu): destroyed with the context
    struct Attachment { void* p; void (*destroy)(void*); };
    std::vector<Attachment> attachments;
};

namespace zkb {

int set_err(zkb_ctx* c, int code, const char* fmt, ...);

#define ZKB_CUDA(ctx, call)                                                              \
    do {                                                                                 \
        cudaError_t e__ = (call);                                                        \
        if (e__ != cudaSuccess)                                                          \
            return zkb::set_err((ctx), ZKB_ERR_CUDA, "%s failed: %s (%s:%d)", #call,     \
                                cudaGetErrorString(e__), __FILE__, __LINE__);            \
    } while (0)

// C++ exceptions must not cross the C ABI (undefined behaviour in ctypes / Rust callers): entry points that allocate on the host
// run their body under this guard and report std::bad_alloc as ZKB_ERR_NOMEM
#define ZKB_ABI_GUARD(ctx, body)                                                                 \
    try { body } catch (const std::bad_alloc&) {                                                 \
        return zkb::set_err((ctx), ZKB_ERR_NOMEM, "out of host memory (std::bad_alloc)");        \
    } catch (...) {                                                                              \
        return zkb::set_err((ctx), ZKB_ERR_NOMEM, "unexpected C++ exception inside the library"); \
    }

#define ZKB_TRY(expr)            \
    do {                         \
        int rc__ = (expr);       \
        if (rc__ != 0) return rc__; \
    } while (0)

// kernel classes for zkb_ctx_profile_* (keep in sync with zkb_kernel_name)
enum KernelId { K_POW_TABLE = 0, K_NTT_PASS = 1, K_ELEMENTWISE = 2, K_LEAF_TILE = 3, K_FOLD_LEAF_TILE = 4,
                K_NODE_TILE = 5, K_MERKLE_SMALL = 6, K_OPEN = 7, K_FOLD = 8, K_GATHER = 9, K_LEAF1 = 10, K_FRI_TAIL = 11, K_NTT_LEAF = 12, K_COUNT = 13 };

// Brackets one launch with events when profiling is on; always counts the launch.
struct LaunchScope {
    zkb_ctx* c; int id; cudaEvent_t a = nullptr, b = nullptr;
    LaunchScope(zkb_ctx* c_, int id_);
    ~LaunchScope();
};
int prof_collect(zkb_ctx* c);

// per-device kernel attributes (dynamic shared-memory opt-in) of ntt.cu / merkle.cu; called by zkb_ctx_create
int ntt_device_init(zkb_ctx* c);
int merkle_device_init(zkb_ctx* c);
int fri_tail_device_init(zkb_ctx* c);
// room for `count` device Fiat-Shamir contexts (keccak.cuh FsDev) in c->fs_dev (+ the tail kernel's barrier word behind them)
int ensure_fs_dev(zkb_ctx* c, size_t count);

// cudaStreamSynchronize(c->stream), or - blocking_sync - cudaStreamQuery + usleep: the calling thread sleeps instead of
// spinning on a core (callers that keep many contexts in flight on their own threads)
cudaError_t ctx_stream_sync(zkb_ctx* c);

int scratch_reserve(zkb_ctx* c, size_t bytes, void** out);
// pinned host buffer `slot` of at least `bytes` (grow-only, owned by the context): D2H copies into pageable memory
// are staged and synchronous (~8 GB/s); into pinned memory they run at PCIe rate
int host_scratch_reserve(zkb_ctx* c, int slot, size_t bytes, uint8_t** out);
int get_pow_table(zkb_ctx* c, const fe& base, uint32_t log_n, DevPow* out);
bool is_device_ptr(const void* p);

// Stream-ordered device memory from the device's default CUDA memory pool (cudaMallocAsync on
// the context's stream; the pool's release threshold is raised at context creation so freed
// blocks are kept and re-used: a 0.8 GB FRI arena costs microseconds per call, not a
// cudaMalloc/cudaFree pair of tens of milliseconds).
cudaError_t dev_alloc(zkb_ctx* c, void** p, size_t bytes);
void dev_free(zkb_ctx* c, void* p);

// RAII device buffer used for transient staging of host inputs
struct DevBuf {
    void* p = nullptr;
    zkb_ctx* ctx = nullptr;
    ~DevBuf() { if (p) dev_free(ctx, p); }
    int alloc(zkb_ctx* c, size_t bytes);
};

// If `p` is a host pointer, stage it into `buf` (async on the ctx stream) and return the
// device pointer; device pointers are passed through.
int stage_in(zkb_ctx* c, const void* p, size_t bytes, DevBuf& buf, const void** dev);
int stage_in_once(zkb_ctx* c, const void* p, size_t bytes, DevBuf& buf, const void** dev);   // pinned host memory is read in place
const void* pinned_device_alias(const void* p);

// ---- host-side field helpers (canonical values) built on the verified host path of
// fe128.cuh
inline fe h_mul(const fe& a, const fe& b) { return fe_mul(a, b); }
inline fe h_pow(const fe& a, uint64_t e) { return fe_from_mont(fe_mont_pow(fe_to_mont(a), e)); }
fe h_inv(const fe& a);                      // a^(p-2); inverse(0) = 0 like field.rs:160-169
inline fe h_from_u64(uint64_t x) { fe r = fe_zero(); r.v[0] = (uint32_t)x; r.v[1] = (uint32_t)(x >> 32); return r; }
inline fe h_load(const uint8_t* le16) { fe r; for (int i = 0; i < 4; i++) r.v[i] = (uint32_t)le16[4 * i] | ((uint32_t)le16[4 * i + 1] << 8) | ((uint32_t)le16[4 * i + 2] << 16) | ((uint32_t)le16[4 * i + 3] << 24); return r; }
inline void h_store(uint8_t* le16, const fe& a) { for (int i = 0; i < 4; i++) { le16[4 * i] = (uint8_t)a.v[i]; le16[4 * i + 1] = (uint8_t)(a.v[i] >> 8); le16[4 * i + 2] = (uint8_t)(a.v[i] >> 16); le16[4 * i + 3] = (uint8_t)(a.v[i] >> 24); } }

inline uint32_t ilog2_u64(uint64_t x) { uint32_t l = 0; while ((1ull << l) < x) l++; return l; }
inline uint64_t next_pow2_u64(uint64_t x) { uint64_t p = 1; while (p < x) p <<= 1; return p; }

}  // namespace zkb
