// core.cu - context lifecycle, scratch arena, power-table cache, host scalar helpers.
#include <stdarg.h>
#include <unistd.h>
#include <stdlib.h>
#include <string.h>
#include "ctx.hpp"
#include "keccak.cuh"

namespace zkb {

int set_err(zkb_ctx* c, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (c) c->err = buf;
    return code;
}

static cudaEvent_t prof_event(zkb_ctx* c) {
    if (!c->prof_pool.empty()) { cudaEvent_t e = c->prof_pool.back(); c->prof_pool.pop_back(); return e; }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}
LaunchScope::LaunchScope(zkb_ctx* c_, int id_) : c(c_), id(id_) {
    c->launches++;
    if (c->profiling) { a = prof_event(c); b = prof_event(c); cudaEventRecord(a, c->stream); }
}
LaunchScope::~LaunchScope() {
    if (a) { cudaEventRecord(b, c->stream); c->prof_recs.push_back({id, a, b}); }
}
int prof_collect(zkb_ctx* c) {
    if (c->prof_recs.empty()) return 0;
    ZKB_CUDA(c, cudaStreamSynchronize(c->stream));
    for (auto& r : c->prof_recs) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) { c->prof_ms[r.id] += ms; c->prof_count[r.id]++; }
        c->prof_pool.push_back(r.a);
        c->prof_pool.push_back(r.b);
    }
    c->prof_recs.clear();
    return 0;
}

cudaError_t dev_alloc(zkb_ctx* c, void** p, size_t bytes) {
    return cudaMallocAsync(p, bytes ? bytes : 16, c->stream);
}
void dev_free(zkb_ctx* c, void* p) {
    if (p) cudaFreeAsync(p, c->stream);
}

int DevBuf::alloc(zkb_ctx* c, size_t bytes) {
    ctx = c;
    ZKB_CUDA(c, dev_alloc(c, &p, bytes));
    return 0;
}

bool is_device_ptr(const void* p) {
    cudaPointerAttributes a;
    cudaError_t e = cudaPointerGetAttributes(&a, p);
    if (e != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

// Pinned (page-locked, UVA-mapped) host memory can be read by a kernel in place over PCIe.
// Returns the device alias of `p`, or nullptr if `p` is not such memory (ZKB_ZERO_COPY=0 disables).
const void* pinned_device_alias(const void* p) {
    static int on = -1;
    if (on < 0) { const char* e = getenv("ZKB_ZERO_COPY"); on = e ? atoi(e) : 1; }
    if (!on) return nullptr;
    cudaPointerAttributes a;
    cudaError_t e = cudaPointerGetAttributes(&a, p);
    if (e != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return a.type == cudaMemoryTypeHost ? a.devicePointer : nullptr;
}

// For inputs the consuming kernel reads exactly ONCE (the coefficients of an LDE): pinned host
// memory is handed to the kernel as is, so the PCIe transfer overlaps the first NTT pass instead
// of preceding it; pageable memory is staged as usual.
int stage_in_once(zkb_ctx* c, const void* p, size_t bytes, DevBuf& buf, const void** dev) {
    if (c->zero_copy_inputs && bytes >= (1u << 20)) {
        const void* alias = pinned_device_alias(p);
        if (alias) { *dev = alias; return 0; }
    }
    return stage_in(c, p, bytes, buf, dev);
}

int stage_in(zkb_ctx* c, const void* p, size_t bytes, DevBuf& buf, const void** dev) {
    if (is_device_ptr(p)) { *dev = p; return 0; }
    ZKB_TRY(buf.alloc(c, bytes));
    ZKB_CUDA(c, cudaMemcpyAsync(buf.p, p, bytes, cudaMemcpyHostToDevice, c->stream));
    *dev = buf.p;
    return 0;
}

int scratch_reserve(zkb_ctx* c, size_t bytes, void** out) {
    if (bytes > c->scratch_bytes) {
        // the arena may still be in use by queued kernels
        ZKB_CUDA(c, cudaStreamSynchronize(c->stream));
        if (c->scratch) cudaFree(c->scratch);
        c->scratch = nullptr;
        c->scratch_bytes = 0;
        size_t want = bytes + (bytes >> 3);
        ZKB_CUDA(c, cudaMalloc(&c->scratch, want));
        c->scratch_bytes = want;
    }
    *out = c->scratch;
    return 0;
}

int host_scratch_reserve(zkb_ctx* c, int slot, size_t bytes, uint8_t** out) {
    if (slot < 0 || slot >= 3) return set_err(c, ZKB_ERR_ARG, "internal: host scratch slot");
    if (bytes > c->host_scratch_bytes[slot]) {
        ZKB_CUDA(c, cudaStreamSynchronize(c->stream));            // a queued copy may still target the old buffer
        if (c->host_scratch[slot]) cudaFreeHost(c->host_scratch[slot]);
        c->host_scratch[slot] = nullptr;
        c->host_scratch_bytes[slot] = 0;
        const size_t want = bytes + (bytes >> 2);
        ZKB_CUDA(c, cudaHostAlloc(&c->host_scratch[slot], want, cudaHostAllocDefault));
        c->host_scratch_bytes[slot] = want;
    }
    *out = (uint8_t*)c->host_scratch[slot];
    return 0;
}

cudaError_t ctx_stream_sync(zkb_ctx* c) {
    if (!c->blocking_sync) return cudaStreamSynchronize(c->stream);
    // poll + sleep: the thread gives its core away while the GPU works (cudaStreamSynchronize spins)
    for (;;) {
        cudaError_t e = cudaStreamQuery(c->stream);
        if (e != cudaErrorNotReady) return e;
        usleep(15);
    }
}

int ensure_fs_dev(zkb_ctx* c, size_t count) {
    if (c->fs_dev_count >= count) return 0;
    ZKB_CUDA(c, cudaStreamSynchronize(c->stream));
    if (c->fs_dev) cudaFree(c->fs_dev);
    c->fs_dev = nullptr; c->fs_dev_count = 0;
    ZKB_CUDA(c, cudaMalloc(&c->fs_dev, count * sizeof(FsDev) + 256));       // + the tail kernel's barrier word
    ZKB_CUDA(c, cudaMemsetAsync(c->fs_dev, 0, count * sizeof(FsDev) + 256, c->stream));
    c->fs_dev_count = count;
    return 0;
}

fe h_inv(const fe& a) {
    if (fe_is_zero(a)) return fe_zero();
    // a^(p-2), p-2 = 0xCB7FFFFF_FFFFFFFF_FFFFFFFF_FFFFFFFF
    const uint32_t e[4] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xCB7FFFFFu};
    fe base = fe_to_mont(a), acc = fe_mont_one();
    for (int i = 127; i >= 0; i--) {
        acc = fe_montmul(acc, acc);
        if ((e[i >> 5] >> (i & 31)) & 1u) acc = fe_montmul(acc, base);
    }
    return fe_from_mont(acc);
}

// out[i] = base^i * R, i < count  (base_m = base*R).  Each thread: square-and-multiply.
__global__ void k_pow_table(fe base_m, fe* out, uint32_t count) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    fe_store(out + i, fe_mont_pow(base_m, i));
}

int get_pow_table(zkb_ctx* c, const fe& base, uint32_t log_n, DevPow* out) {
    c->clock++;
    for (auto& t : c->pow_tables) {
        if (t->log_n == log_n && fe_eq(t->base, base)) {
            t->stamp = c->clock;
            *out = DevPow{t->lo, t->hi, t->lo_bits};
            return 0;
        }
    }
    if (c->pow_tables.size() >= 48) {   // evict least recently used
        size_t victim = 0;
        for (size_t i = 1; i < c->pow_tables.size(); i++)
            if (c->pow_tables[i]->stamp < c->pow_tables[victim]->stamp) victim = i;
        ZKB_CUDA(c, cudaStreamSynchronize(c->stream));
        cudaFree(c->pow_tables[victim]->lo);
        c->pow_tables.erase(c->pow_tables.begin() + victim);
    }
    std::unique_ptr<PowTable> t(new PowTable());
    t->base = base;
    t->log_n = log_n;
    t->lo_bits = log_n < 12 ? log_n : 12;
    uint32_t n_lo = 1u << t->lo_bits, n_hi = 1u << (log_n - t->lo_bits);
    ZKB_CUDA(c, cudaMalloc(&t->lo, sizeof(fe) * (size_t)(n_lo + n_hi)));
    t->hi = t->lo + n_lo;
    fe base_m = fe_to_mont(base);
    fe hi_base_m = base_m;
    for (uint32_t i = 0; i < t->lo_bits; i++) hi_base_m = fe_montmul(hi_base_m, hi_base_m);
    { LaunchScope ls(c, K_POW_TABLE); k_pow_table<<<(n_lo + 255) / 256, 256, 0, c->stream>>>(base_m, t->lo, n_lo); }
    { LaunchScope ls(c, K_POW_TABLE); k_pow_table<<<(n_hi + 255) / 256, 256, 0, c->stream>>>(hi_base_m, t->hi, n_hi); }
    ZKB_CUDA(c, cudaGetLastError());
    t->stamp = c->clock;
    *out = DevPow{t->lo, t->hi, t->lo_bits};
    c->pow_tables.push_back(std::move(t));
    return 0;
}

}  // namespace zkb

using namespace zkb;

extern "C" {

const char* zkb_version(void) { return "zkb200 0.1 (sm_100a)"; }

int zkb_ctx_create(int device, void* stream, zkb_ctx** out) {
    if (!out) return ZKB_ERR_ARG;
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0 || device < 0 || device >= count) {
        cudaGetLastError();
        return ZKB_ERR_CUDA;       // no CPU fallback: no device, no context
    }
    zkb_ctx* c = new zkb_ctx();
    c->device = device;
    if (cudaSetDevice(device) != cudaSuccess) { delete c; return ZKB_ERR_CUDA; }
    if (stream) {
        c->stream = (cudaStream_t)stream;
    } else {
        if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) { delete c; return ZKB_ERR_CUDA; }
        c->own_stream = true;
    }
    cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device);
    {   // keep freed blocks in the default pool instead of returning them to the driver
        cudaMemPool_t pool = nullptr;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            uint64_t keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
    }
    c->pinned_bytes = 1 << 20;
    if (cudaMallocHost(&c->pinned, c->pinned_bytes) != cudaSuccess) { delete c; return ZKB_ERR_CUDA; }
    memset(c->pinned, 0, c->pinned_bytes);          // the root handshake reads a sequence flag from this buffer
    if (ntt_device_init(c) != 0 || merkle_device_init(c) != 0 || fri_tail_device_init(c) != 0) { zkb_ctx_destroy(c); return ZKB_ERR_CUDA; }
    if (const char* e = getenv("ZKB_BLOCKING_SYNC")) c->blocking_sync = atoi(e) != 0;
    if (const char* e = getenv("ZKB_TAIL_THREADS")) { int t = atoi(e); if (t == 0 || t == 128 || t == 256 || t == 512) c->tail_threads = (uint32_t)t; }
    *out = c;
    return 0;
}

void zkb_ctx_destroy(zkb_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    for (auto& t : c->pow_tables) cudaFree(t->lo);
    for (auto& at : c->attachments) at.destroy(at.p);
    for (auto& r : c->prof_recs) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    for (auto e : c->prof_pool) cudaEventDestroy(e);
    if (c->scratch) cudaFree(c->scratch);
    if (c->tree_bars) cudaFree(c->tree_bars);
    if (c->fs_dev) cudaFree(c->fs_dev);
    for (int i = 0; i < 3; i++) if (c->host_scratch[i]) cudaFreeHost(c->host_scratch[i]);
    if (c->pinned) cudaFreeHost(c->pinned);
    if (c->own_stream) cudaStreamDestroy(c->stream);
    delete c;
}

const char* zkb_last_error(const zkb_ctx* c) { return c ? c->err.c_str() : "no context (no CUDA device?)"; }

int zkb_ctx_sync(zkb_ctx* c) {
    if (!c) return ZKB_ERR_ARG;
    ZKB_CUDA(c, cudaStreamSynchronize(c->stream));
    return 0;
}

uint64_t zkb_ctx_launches(const zkb_ctx* c) { return c ? c->launches : 0; }

int zkb_ctx_assembly_threads(zkb_ctx* c, int threads) {
    if (!c || threads < 1) return ZKB_ERR_ARG;
    c->assembly_threads = (size_t)threads;
    return 0;
}

int zkb_ctx_tail_threads(zkb_ctx* c, int threads) {
    if (!c || (threads != 0 && threads != 128 && threads != 256 && threads != 512)) return ZKB_ERR_ARG;
    c->tail_threads = (uint32_t)threads;
    return 0;
}

int zkb_ctx_blocking_sync(zkb_ctx* c, int enable) {
    if (!c) return ZKB_ERR_ARG;
    c->blocking_sync = enable != 0;
    return 0;
}

int zkb_ctx_profile(zkb_ctx* c, int enable) {
    if (!c) return ZKB_ERR_ARG;
    ZKB_TRY(prof_collect(c));
    c->profiling = enable != 0;
    if (enable == 2) for (int i = 0; i < 16; i++) { c->prof_ms[i] = 0; c->prof_count[i] = 0; }
    return 0;
}
int zkb_ctx_profile_read(zkb_ctx* c, int kernel_id, double* total_ms, uint64_t* count) {
    if (!c || kernel_id < 0 || kernel_id >= K_COUNT) return ZKB_ERR_ARG;
    ZKB_TRY(prof_collect(c));
    if (total_ms) *total_ms = c->prof_ms[kernel_id];
    if (count) *count = c->prof_count[kernel_id];
    return 0;
}
int zkb_ctx_zero_copy_inputs(zkb_ctx* c, int enable) {
    if (!c) return ZKB_ERR_ARG;
    c->zero_copy_inputs = enable != 0;
    return 0;
}
const char* zkb_kernel_name(int kernel_id) {
    static const char* names[K_COUNT] = {"k_pow_table", "k_ntt_pass", "k_elementwise", "k_leaf8<false>", "k_leaf8<true>",
                                         "k_node8", "k_tree", "k_open", "k_fold", "k_gather3", "k_leaf1", "k_fri_tail", "k_ntt_rr_leaf"};
    return (kernel_id >= 0 && kernel_id < K_COUNT) ? names[kernel_id] : nullptr;
}

int zkb_dev_alloc(zkb_ctx* c, size_t bytes, void** dptr) {
    if (!c || !dptr) return ZKB_ERR_ARG;
    ZKB_CUDA(c, cudaSetDevice(c->device));
    ZKB_CUDA(c, cudaMalloc(dptr, bytes ? bytes : 16));
    return 0;
}
int zkb_dev_free(zkb_ctx* c, void* dptr) {
    if (!c) return ZKB_ERR_ARG;
    ZKB_CUDA(c, cudaStreamSynchronize(c->stream));
    ZKB_CUDA(c, cudaFree(dptr));
    return 0;
}
int zkb_memcpy(zkb_ctx* c, void* dst, const void* src, size_t bytes) {
    if (!c) return ZKB_ERR_ARG;
    ZKB_CUDA(c, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, c->stream));
    ZKB_CUDA(c, cudaStreamSynchronize(c->stream));
    return 0;
}

// ---- host scalar helpers ------------------------------------------------------------
static const uint8_t GEN_LE[16] = {0xD1, 0xF7, 0xF6, 0x18, 0x9C, 0x8F, 0x03, 0xB5,
                                   0x0F, 0x47, 0xEE, 0x12, 0xED, 0xFB, 0x40, 0x40};   // field.rs:43

void zkb_field_generator(uint8_t out[16]) { memcpy(out, GEN_LE, 16); }

int zkb_primitive_nth_root(uint64_t n, uint8_t out[16]) {
    // field.rs:58-71: square the order-2^119 generator down to order n
    if (n == 0 || (n & (n - 1)) != 0) return ZKB_ERR_ARG;
    fe root = fe_to_mont(h_load(GEN_LE));
    uint32_t log_n = ilog2_u64(n);
    for (uint32_t i = 119; i > log_n; i--) root = fe_montmul(root, root);
    h_store(out, fe_from_mont(root));
    return 0;
}
void zkb_field_mul(const uint8_t a[16], const uint8_t b[16], uint8_t out[16]) { h_store(out, h_mul(h_load(a), h_load(b))); }
void zkb_field_inv(const uint8_t a[16], uint8_t out[16]) { h_store(out, h_inv(h_load(a))); }
void zkb_field_pow(const uint8_t a[16], uint64_t e, uint8_t out[16]) { h_store(out, h_pow(h_load(a), e)); }

void zkb_field_sample(const uint8_t* bytes, size_t len, uint8_t out[16]) {
    // field.rs:87-99: big-endian value of the last 16 bytes, mod p (value < 2^128 < 2p)
    uint8_t le[16] = {0};
    size_t take = len < 16 ? len : 16;
    for (size_t i = 0; i < take; i++) le[i] = bytes[len - 1 - i];
    fe v = h_load(le);
    if (fe_ge_p(v)) {
        fe p; p.v[0] = P0; p.v[1] = 0; p.v[2] = 0; p.v[3] = P3;
        // v - p without the modular wrap
        uint64_t bw = 0; fe r;
        for (int i = 0; i < 4; i++) { uint64_t d = (uint64_t)v.v[i] - p.v[i] - bw; r.v[i] = (uint32_t)d; bw = (d >> 63) & 1; }
        v = r;
    }
    h_store(out, v);
}

}  // extern "C"
