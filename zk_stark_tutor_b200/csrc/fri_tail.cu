// fri_tail.cu - the latency-bound tail of FRI::commit as ONE persistent kernel.
//
// Replaces, for every layer of at most 2^17 values, the per-round sequence of src/fri.rs:135-163
//   MerkleRoot::commit(codeword) -> push Root -> alpha = sample(fiat_shamir(32)) -> split-and-fold
// (merkle_root.rs:7-32, proof_stream.rs:36-40, field.rs:87-99, fri.rs:150-159).
//
// Below 2^17 values a layer cannot fill the GPU: each round is a chain of log2(n) + 1 dependent BLAKE2b
// compressions plus the Fiat-Shamir challenge, and with one launch pair + one host hop per round the
// chain was mostly launch latency and PCIe round trips.  Here 128 co-resident CTAs (cooperative launch)
// walk through ALL remaining rounds without leaving the device:
//   leaf phase   CTA b owns the leaves [b * chunk, (b + 1) * chunk): fold of the previous layer (the
//                folded value is written once as the next codeword), decimal leaf hash, digest to
//                shared memory (and to the retained tree);
//   chunk phase  the chunk is reduced to one node inside shared memory, one compression per QUAD of lanes;
//   ONE grid barrier (arrive-counter in L2);
//   top phase    EVERY CTA reduces the <= 128 chunk roots to the Merkle root by itself (CTA 0 stores the levels),
//                and its warp 0 runs the transcript sponge (keccak.cuh): Root record, SHAKE256 challenge,
//                Field::sample, alpha / offset_r.  All CTAs hold the same alpha: no second barrier, no broadcast.
// At the end CTA 0 hands every root of the commit and the last codeword to the polling host through mapped
// pinned memory.  Tree layout as TreeLayout with top == 0 (every level stored, level l at node offset
// 2^(log_n+1) - 2^(log_n+1-l)), so FRI::query opens these layers like any other tree.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "merkle_dev.cuh"
#include "keccak.cuh"
#include "fri_tail.cuh"

namespace zkb {

__device__ __forceinline__ uint64_t lvl_off(uint32_t log_n, uint32_t l) { return (2ull << log_n) - (2ull << (log_n - l)); }

// Arrive + wait on a monotonically increasing counter in L2.  Release / acquire at gpu scope (both cumulative, and the
// block barriers on either side extend them to the whole CTA) instead of two full membars around a relaxed atomic.
__device__ __forceinline__ void grid_arrive_wait(uint32_t* bar, uint32_t target, volatile uint32_t* timeout_flag, uint32_t* dead) {
    __syncthreads();                                   // the CTA's stores precede the arrival
    if (threadIdx.x == 0) {
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");
        const long long t0 = clock64();
        uint32_t spins = 0, v;
        for (;;) {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
            if (v >= target) break;
            // a co-resident grid cannot deadlock here; the bound only keeps a broken launch from hanging the GPU
            if ((++spins & 0x3FFu) == 0 && clock64() - t0 > (4ll << 30)) { if (timeout_flag) *timeout_flag = ZKB_TAIL_TIMEOUT_FLAG; *dead = 1; break; }
        }
    }
    __syncthreads();
}

__global__ void __launch_bounds__(ZKB_TAIL_THREADS, 1) k_fri_tail(TailArgs a) {
    extern __shared__ uint64_t tree_smem[];
    __shared__ __align__(16) FsSponge s_sp;
    __shared__ __align__(16) fe s_kk;
    __shared__ uint32_t s_dead;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, cta = blockIdx.x;
    uint64_t* bufA = tree_smem;
    uint64_t* bufB = tree_smem + dig_word(ZKB_TAIL_MAX_CHUNK) + 8;
    if (tid < 32) {
        uint64_t* d = reinterpret_cast<uint64_t*>(&s_sp);
        const uint64_t* s = reinterpret_cast<const uint64_t*>(&a.fs->sp);
        for (uint32_t i = lane; i < sizeof(FsSponge) / 8; i += 32) d[i] = __ldcg(s + i);
        if (lane == 0) { s_kk = fe_load(&a.fs->kk_m); s_dead = 0; }
    }
    __syncthreads();
    uint32_t arrivals = 0;
#define ZKB_TAIL_STAMP(slot) do { if (a.dbg && cta == 0 && tid == 0) a.dbg[k * 8 + (slot)] = clock64(); } while (0)
#pragma unroll 1
    for (uint32_t k = 0; k < a.n_rounds; k++) {
        const uint32_t log_n = a.log_n0 - k, n = 1u << log_n, round = a.r0 + k;
        const uint32_t chunk_log = log_n > 9 ? log_n - 7 : (log_n < 2 ? log_n : 2);
        const uint32_t chunk = 1u << chunk_log, chunks = n >> chunk_log;
        uint8_t* const nodes = a.nodes[k];
        ZKB_TAIL_STAMP(0);
        if (cta < chunks) {
            // ---- leaf phase: (fold +) leaf hash of this CTA's chunk
            const bool plain = k == 0 && a.first_is_plain;
            const fe* src = k == 0 ? a.cw_in : a.cw[k - 1];
            fe* dst = a.cw[k];
            const fe kk = s_kk;
            for (uint32_t j = tid; j < chunk; j += blockDim.x) {
                const uint32_t i = cta * chunk + j;
                fe v;
                if (plain) {
                    v = fe_ldg(src + i);
                } else {
                    const fe k_m = fe_montmul(kk, pow2lvl_m(a.winv, (uint64_t)i << (round - 1)));
                    const uint4 xa = __ldcg(reinterpret_cast<const uint4*>(src + i)), xb = __ldcg(reinterpret_cast<const uint4*>(src + n + i));
                    fe fa, fb;
                    fa.v[0] = xa.x; fa.v[1] = xa.y; fa.v[2] = xa.z; fa.v[3] = xa.w;
                    fb.v[0] = xb.x; fb.v[1] = xb.y; fb.v[2] = xb.z; fb.v[3] = xb.w;
                    v = fe_half(fe_add(fe_add(fa, fb), fe_montmul(k_m, fe_sub(fa, fb))));
                    fe_store(dst + i, v);
                }
                uint64_t h[8];
                b2_leaf_call(&v, h);
                uint64_t* s = bufA + dig_word(j);
#pragma unroll
                for (int w = 0; w < 8; w++) s[w] = h[w];
                g_store_digest(nodes, i, h);
            }
            __syncthreads();
            ZKB_TAIL_STAMP(1);
            // ---- chunk phase: levels 1 .. chunk_log of this chunk
            reduce_in_smem(bufA, bufB, chunk, [&](uint32_t level) -> uint8_t* { return nodes + lvl_off(log_n, level) * 64; }, 1,
                           (uint64_t)cta * chunk, [](uint32_t, uint64_t, uint64_t) {});
        }
        ZKB_TAIL_STAMP(2);
        grid_arrive_wait(a.bar, (++arrivals) * gridDim.x, a.host_flag, &s_dead);
        ZKB_TAIL_STAMP(3);
        // ---- top phase (every CTA): chunk roots -> root, then the transcript
        load_chunk(bufA, nodes + lvl_off(log_n, chunk_log) * 64, chunks, true);
        const uint64_t* root = reduce_in_smem(bufA, bufB, chunks,
            [&](uint32_t level) -> uint8_t* { return cta == 0 ? nodes + lvl_off(log_n, level) * 64 : nullptr; },
            chunk_log + 1, 0, [](uint32_t, uint64_t, uint64_t) {});
        ZKB_TAIL_STAMP(4);
        if (tid < 32) {
            const bool want = round + 1 < a.total_rounds;
            const fe alpha = fs_round_warp(&s_sp, reinterpret_cast<const uint8_t*>(root), want, lane);
            if (lane == 0 && want) s_kk = fe_montmul(alpha, fe_ldg(&a.fs->inv_off_m2[round]));
            if (cta == 0 && lane < 8) reinterpret_cast<uint64_t*>(a.fs->roots[round])[lane] = root[lane];
        }
        __syncthreads();
        ZKB_TAIL_STAMP(5);
    }
    // ---- epilogue: results to the host
    if (cta == 0) {
        if (a.host_out) {
            const uint64_t* r = reinterpret_cast<const uint64_t*>(a.fs->roots);
            uint64_t* o = reinterpret_cast<uint64_t*>(a.host_out);
            for (uint32_t i = tid; i < a.total_rounds * 8; i += blockDim.x) o[i] = __ldcg(r + i);
            const uint32_t last_n = 1u << (a.log_n0 - (a.n_rounds - 1));
            const uint4* lc = reinterpret_cast<const uint4*>(a.n_rounds == 1 && a.first_is_plain ? a.cw_in : a.cw[a.n_rounds - 1]);
            uint4* lo = reinterpret_cast<uint4*>(a.host_out + ZKB_TAIL_HOST_CW_OFF);
            for (uint32_t i = tid; i < last_n; i += blockDim.x) lo[i] = __ldcg(lc + i);
            __threadfence_system();
            __syncthreads();
            if (tid == 0 && !s_dead) { __threadfence_system(); *a.host_flag = a.seq; }
        }
        if (tid < 32) {                                   // the sponge continues in device memory (not needed by the host path)
            uint64_t* g = reinterpret_cast<uint64_t*>(&a.fs->sp);
            const uint64_t* d = reinterpret_cast<const uint64_t*>(&s_sp);
            for (uint32_t i = lane; i < sizeof(FsSponge) / 8; i += 32) g[i] = d[i];
        }
    }
}

static size_t tail_smem_bytes() {
    return (dig_words_host(ZKB_TAIL_MAX_CHUNK) + 8 + dig_words_host(ZKB_TAIL_MAX_CHUNK / 2) + 8) * sizeof(uint64_t);
}

int fri_tail_device_init(zkb_ctx* c) {
    ZKB_CUDA(c, cudaFuncSetAttribute(k_fri_tail, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tail_smem_bytes()));
    return 0;
}

int fri_tail_launch(zkb_ctx* c, const TailArgs& a) {
    if (a.n_rounds == 0 || a.n_rounds > ZKB_TAIL_MAX_ROUNDS || a.log_n0 > ZKB_TAIL_MAX_LOG)
        return set_err(c, ZKB_ERR_ARG, "internal: FRI tail takes 1..%u rounds of at most 2^%u values", ZKB_TAIL_MAX_ROUNDS, ZKB_TAIL_MAX_LOG);
    if (c->sm_count < (int)ZKB_TAIL_CTAS) return set_err(c, ZKB_ERR_CUDA, "the persistent FRI tail needs %u SMs (device has %d)", ZKB_TAIL_CTAS, c->sm_count);
    TailArgs args = a;
    static const bool debug = getenv("ZKB_TAIL_DEBUG") != nullptr;
    DevBuf dbg;
    if (debug) {
        ZKB_TRY(dbg.alloc(c, ZKB_TAIL_MAX_ROUNDS * 8 * sizeof(unsigned long long)));
        ZKB_CUDA(c, cudaMemsetAsync(dbg.p, 0, ZKB_TAIL_MAX_ROUNDS * 8 * sizeof(unsigned long long), c->stream));
        args.dbg = (unsigned long long*)dbg.p;
    }
    void* params[] = {&args};
    ZKB_CUDA(c, cudaMemsetAsync(args.bar, 0, sizeof(uint32_t), c->stream));     // arrive counter of the grid barrier
    {
        LaunchScope ls(c, K_FRI_TAIL);
        // 512 threads per CTA = the whole register file of 128 SMs for the duration of the (latency-bound) tail: right for one
        // codeword at a time; a context that shares its GPU with other contexts' kernels (column / proof pipelines) asks for 256
        // (zkb_ctx_tail_threads), which leaves half of every SM to the other lanes' throughput kernels
        ZKB_CUDA(c, cudaLaunchCooperativeKernel((const void*)k_fri_tail, dim3(ZKB_TAIL_CTAS), dim3(c->tail_threads), params, tail_smem_bytes(), c->stream));
    }
    if (debug) {
        unsigned long long h[ZKB_TAIL_MAX_ROUNDS * 8];
        ZKB_CUDA(c, cudaMemcpyAsync(h, dbg.p, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
        ZKB_CUDA(c, cudaStreamSynchronize(c->stream));
        for (uint32_t k = 0; k < a.n_rounds; k++)
            fprintf(stderr, "tail round %2u (2^%2u): leaf %6llu chunk %6llu barrier %6llu load+top %6llu fs %6llu clk; total %6llu\n", a.r0 + k, a.log_n0 - k,
                    h[k * 8 + 1] - h[k * 8], h[k * 8 + 2] - h[k * 8 + 1], h[k * 8 + 3] - h[k * 8 + 2], h[k * 8 + 4] - h[k * 8 + 3],
                    h[k * 8 + 5] - h[k * 8 + 4], h[k * 8 + 5] - h[k * 8]);
    }
    return 0;
}

// ---- SHAKE256 on the device (test hook for the warp sponge: proof_stream.rs:129-145's KAT runs through it) ----
__global__ void k_shake256(const uint8_t* msg, uint32_t len, uint64_t* out) {
    shake256_warp(msg, len, out, 17, threadIdx.x);
}

}  // namespace zkb

using namespace zkb;

extern "C" int zkb_shake256_device(zkb_ctx* c, const uint8_t* msg, size_t len, uint8_t* out, size_t out_len) {
    if (!c || (!msg && len) || !out) return ZKB_ERR_ARG;
    if (out_len > 136 || len > (1u << 30)) return set_err(c, ZKB_ERR_ARG, "shake256_device: at most 136 output bytes");
    ZKB_CUDA(c, cudaSetDevice(c->device));
    DevBuf in, o;
    ZKB_TRY(in.alloc(c, len + 8));
    ZKB_TRY(o.alloc(c, 17 * 8));
    if (len) ZKB_CUDA(c, cudaMemcpyAsync(in.p, msg, len, cudaMemcpyHostToDevice, c->stream));
    { LaunchScope ls(c, K_ELEMENTWISE); k_shake256<<<1, 32, 0, c->stream>>>((const uint8_t*)in.p, (uint32_t)len, (uint64_t*)o.p); }
    ZKB_CUDA(c, cudaGetLastError());
    uint8_t host[136];
    ZKB_CUDA(c, cudaMemcpyAsync(host, o.p, 136, cudaMemcpyDeviceToHost, c->stream));
    ZKB_CUDA(c, cudaStreamSynchronize(c->stream));
    memcpy(out, host, out_len);
    return 0;
}
