// ntt4.cu - ONE NTT spread over the g GPUs of an NVLink / NVSwitch box (BASELINE configs[4]) and the
// multi-GPU column batch (8e.1, configs[3]) behind the C ABI.
//
// Same function as src/fft/ntt.rs:7-49 / :51-68 (X[k] = sum_n x[n] w^(nk), natural order), N = g * L:
// write n = r + g*m (r = rank) and k = k2 + L*k1.  Then
//     X[k2 + L*k1] = sum_r w_g^(r*k1) * [ w^(r*k2) * sum_m x[r + g*m] * w_L^(m*k2) ]
//   1. rank r holds the cyclic slice x[r + g*m] and runs a local L-point NTT (root w^g)
//   2. twiddle by w^(r*k2)
//   3. exchange: the k2 range is cut into g blocks, block q belongs to rank q
//   4. rank q runs L/g interleaved g-point NTTs (root w^L) over the g pieces it received -> X[k2 + L*k1] at [k1][k2 - q*L/g]
// Steps 2 and 3 are FUSED into the last pass of step 1: that pass's stores multiply by the running power of w^r and go
// straight into the receiving GPU's HBM through peer-mapped pointers (NVLink stores; same-process peer access or CUDA IPC
// handles between processes), so the transfer overlaps the pass and there is no pack / all-to-all / unpack.  What remains
// between steps 3 and 4 is a barrier, which the caller provides in stream order (events in one process -
// zkb_ntt_4step does that; any stream-ordered collective, e.g. a one-element NCCL all-reduce, between processes).
// Receive buffers are double-buffered, so one barrier per transform is enough.
#include <string.h>
#include <memory>
#include <thread>
#include <vector>
#include "ctx.hpp"
#include "ntt.cuh"
#include "hosthash.hpp"

using namespace zkb;

struct zkb_ntt4 {
    zkb_ctx* ctx = nullptr;
    uint32_t rank = 0, world = 1, log_world = 0;
    uint64_t n_local = 0;                 // L
    fe* recv[2] = {nullptr, nullptr};     // [n1][k2'] pieces from every rank, double-buffered (plain cudaMalloc: IPC-exportable)
    fe* peer_recv[2][ZKB_NTT_MAX_PEERS];  // every rank's receive buffers as seen from this GPU
    void* ipc_opened[2][ZKB_NTT_MAX_PEERS];
    bool connected = false;
    uint32_t parity = 0;                  // buffer the NEXT scatter writes
    uint32_t pending = 0;                 // buffer the next finish reads
    fe root = fe_zero();                  // of the transform in flight
    bool inverse = false;
};

extern "C" {

int zkb_ntt4_create(zkb_ctx* c, uint32_t rank, uint32_t world, size_t n_local, zkb_ntt4** out) {
    if (!c || !out) return ZKB_ERR_ARG;
    *out = nullptr;
    if (world < 1 || world > ZKB_NTT_MAX_PEERS || (world & (world - 1)) || rank >= world)
        return set_err(c, ZKB_ERR_ARG, "ntt4: world %u must be a power of two <= %u, rank < world", world, ZKB_NTT_MAX_PEERS);
    if (n_local == 0 || (n_local & (n_local - 1)) || n_local < world)
        return set_err(c, ZKB_ERR_ARG, "ntt4: local length %zu must be a power of two >= world", n_local);
    ZKB_CUDA(c, cudaSetDevice(c->device));
    std::unique_ptr<zkb_ntt4> p(new zkb_ntt4());
    p->ctx = c; p->rank = rank; p->world = world; p->log_world = ilog2_u64(world); p->n_local = n_local;
    memset(p->peer_recv, 0, sizeof(p->peer_recv));
    memset(p->ipc_opened, 0, sizeof(p->ipc_opened));
    for (int b = 0; b < 2; b++) {
        cudaError_t e = cudaMalloc((void**)&p->recv[b], n_local * sizeof(fe));
        if (e != cudaSuccess) {
            if (b) cudaFree(p->recv[0]);
            return set_err(c, ZKB_ERR_CUDA, "ntt4: cudaMalloc of the receive buffer failed: %s", cudaGetErrorString(e));
        }
        p->peer_recv[b][rank] = p->recv[b];
    }
    if (world == 1) p->connected = true;
    *out = p.release();
    return 0;
}

void zkb_ntt4_free(zkb_ntt4* p) {
    if (!p) return;
    cudaSetDevice(p->ctx->device);
    cudaStreamSynchronize(p->ctx->stream);
    for (int b = 0; b < 2; b++) {
        for (uint32_t q = 0; q < p->world; q++) if (p->ipc_opened[b][q]) cudaIpcCloseMemHandle(p->ipc_opened[b][q]);
        if (p->recv[b]) cudaFree(p->recv[b]);
    }
    delete p;
}

// 2 x 64 bytes: the CUDA IPC handles of this rank's two receive buffers
int zkb_ntt4_export(zkb_ntt4* p, uint8_t handles[128]) {
    if (!p || !handles) return ZKB_ERR_ARG;
    zkb_ctx* c = p->ctx;
    ZKB_CUDA(c, cudaSetDevice(c->device));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    for (int b = 0; b < 2; b++) {
        cudaIpcMemHandle_t h;
        ZKB_CUDA(c, cudaIpcGetMemHandle(&h, p->recv[b]));
        memcpy(handles + 64 * b, &h, 64);
    }
    return 0;
}

// all_handles: world x 128 bytes, rank-major (what an all-gather of zkb_ntt4_export's output gives)
int zkb_ntt4_connect_ipc(zkb_ntt4* p, const uint8_t* all_handles) {
    if (!p || !all_handles) return ZKB_ERR_ARG;
    zkb_ctx* c = p->ctx;
    ZKB_CUDA(c, cudaSetDevice(c->device));
    for (uint32_t q = 0; q < p->world; q++) {
        if (q == p->rank) continue;
        for (int b = 0; b < 2; b++) {
            cudaIpcMemHandle_t h;
            memcpy(&h, all_handles + 128 * q + 64 * b, 64);
            void* ptr = nullptr;
            ZKB_CUDA(c, cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
            p->ipc_opened[b][q] = ptr;
            p->peer_recv[b][q] = (fe*)ptr;
        }
    }
    p->connected = true;
    return 0;
}

// the ranks of ONE process (one context per GPU, or several ranks emulated on one GPU): peer access + plain pointers
int zkb_ntt4_connect_local(zkb_ntt4* const* plans, size_t world) {
    if (!plans || world == 0) return ZKB_ERR_ARG;
    for (size_t r = 0; r < world; r++)
        if (!plans[r] || plans[r]->world != world || plans[r]->rank != r || plans[r]->n_local != plans[0]->n_local)
            return plans[0] ? set_err(plans[0]->ctx, ZKB_ERR_ARG, "ntt4_connect_local: plans[r] must be rank r of one world") : ZKB_ERR_ARG;
    for (size_t r = 0; r < world; r++) {
        zkb_ctx* c = plans[r]->ctx;
        ZKB_CUDA(c, cudaSetDevice(c->device));
        for (size_t q = 0; q < world; q++) {
            const int peer = plans[q]->ctx->device;
            if (peer != c->device) {
                int can = 0;
                ZKB_CUDA(c, cudaDeviceCanAccessPeer(&can, c->device, peer));
                if (!can) return set_err(c, ZKB_ERR_CUDA, "ntt4: GPU %d cannot access GPU %d (no NVLink / P2P)", c->device, peer);
                cudaError_t e = cudaDeviceEnablePeerAccess(peer, 0);
                if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
                else if (e != cudaSuccess) return set_err(c, ZKB_ERR_CUDA, "cudaDeviceEnablePeerAccess(%d) failed: %s", peer, cudaGetErrorString(e));
            }
            for (int b = 0; b < 2; b++) plans[r]->peer_recv[b][q] = plans[q]->recv[b];
        }
        plans[r]->connected = true;
    }
    return 0;
}

// Steps 1-3 on this rank, asynchronous on the context's stream: x_local = the cyclic slice x[rank + world*m] (device,
// n_local values; host pointers are staged).  `root` = the primitive (world * n_local)-th root of the whole transform.
int zkb_ntt4_scatter(zkb_ntt4* p, const uint8_t root[16], int inverse, const void* x_local) {
    if (!p || !root || !x_local) return ZKB_ERR_ARG;
    zkb_ctx* c = p->ctx;
    if (!p->connected) return set_err(c, ZKB_ERR_ARG, "ntt4: connect the ranks first (zkb_ntt4_connect_ipc / _local)");
    ZKB_CUDA(c, cudaSetDevice(c->device));
    const uint64_t L = p->n_local, g = p->world;
    const uint32_t log_l = ilog2_u64(L);
    fe w = h_load(root);
    if (inverse) w = h_inv(w);
    p->root = w; p->inverse = inverse != 0;
    DevBuf bin;
    const void* d_in = nullptr;
    ZKB_TRY(stage_in(c, x_local, L * sizeof(fe), bin, &d_in));
    const uint32_t b = p->parity;
    p->pending = b; p->parity ^= 1u;
    fe wg = w;                                                  // w^g: root of the local transform
    for (uint32_t i = 0; i < p->log_world; i++) wg = h_mul(wg, wg);
    NttOpts o;                                                  // (w is already inverted: plain forward transform with it)
    if (g == 1) {                                               // one rank: the plain transform (n^-1 included for the inverse)
        o.inverse = inverse != 0;
        if (L == 1) ZKB_CUDA(c, cudaMemcpyAsync(p->recv[b], d_in, sizeof(fe), cudaMemcpyDeviceToDevice, c->stream));
        else ZKB_TRY(ntt_exec(c, h_load(root), (const fe*)d_in, L, 0, p->recv[b], 0, 1, log_l, o));
    } else {
        NttExchange x;
        x.oscale_base = h_pow(w, p->rank);                       // y[k2] *= w^(rank * k2)
        x.n_peers = (uint32_t)g; x.log_blk = log_l - p->log_world; x.peer_row = (uint64_t)p->rank << x.log_blk;
        for (uint32_t q = 0; q < g; q++) x.peer[q] = p->peer_recv[b][q];
        if (log_l > 12) {
            o.exchange = &x;
            ZKB_TRY(ntt_exec(c, wg, (const fe*)d_in, L, 0, nullptr, 0, 1, log_l, o));
        } else {                                                // small local transforms: unfused twiddle + scatter
            DevBuf tmp;
            ZKB_TRY(tmp.alloc(c, L * sizeof(fe)));
            if (L == 1) ZKB_CUDA(c, cudaMemcpyAsync(tmp.p, d_in, sizeof(fe), cudaMemcpyDeviceToDevice, c->stream));
            else ZKB_TRY(ntt_exec(c, wg, (const fe*)d_in, L, 0, (fe*)tmp.p, 0, 1, log_l, o));
            ZKB_TRY(ntt_twiddle_scatter(c, (const fe*)tmp.p, L, x));
        }
    }
    if (bin.p) ZKB_CUDA(c, cudaStreamSynchronize(c->stream));     // the staged input is released with `bin`
    return 0;
}

// Step 4, after EVERY rank's scatter of this transform has completed (the caller orders that in the stream):
// out (device or host, n_local values) = X[k2 + L*k1] at [k1 * (L / world) + (k2 - rank * L / world)].
int zkb_ntt4_finish(zkb_ntt4* p, void* out) {
    if (!p || !out) return ZKB_ERR_ARG;
    zkb_ctx* c = p->ctx;
    ZKB_CUDA(c, cudaSetDevice(c->device));
    const uint64_t L = p->n_local, g = p->world;
    const bool out_dev = is_device_ptr(out);
    DevBuf bout;
    fe* d_out = (fe*)out;
    if (!out_dev) { ZKB_TRY(bout.alloc(c, L * sizeof(fe))); d_out = (fe*)bout.p; }
    fe post;
    const fe* postp = nullptr;
    if (p->inverse) { post = fe_to_mont(h_inv(h_from_u64(L * g))); postp = &post; }
    if (g == 1) {
        ZKB_CUDA(c, cudaMemcpyAsync(d_out, p->recv[p->pending], L * sizeof(fe), cudaMemcpyDeviceToDevice, c->stream));
    } else {
        const fe wl = h_pow(p->root, L);                        // w^L: primitive g-th root
        ZKB_TRY(ntt_cross_exec(c, wl, p->log_world, p->recv[p->pending], d_out, L / g, postp));
    }
    if (!out_dev) {
        ZKB_CUDA(c, cudaMemcpyAsync(out, d_out, L * sizeof(fe), cudaMemcpyDeviceToHost, c->stream));
        ZKB_CUDA(c, cudaStreamSynchronize(c->stream));
    }
    return 0;
}

// One process driving `world` GPUs (or emulating `world` ranks on fewer GPUs): plans[r] on its own context.
// scatter on every rank -> events (every finish waits for every scatter) -> finish on every rank.  Asynchronous when all
// pointers are device pointers: synchronise the contexts before reading `out`.
int zkb_ntt4_run(zkb_ntt4* const* plans, size_t world, const uint8_t root[16], int inverse, const void* const* x_local, void* const* out) {
    if (!plans || !root || !x_local || !out || world == 0) return ZKB_ERR_ARG;
    std::vector<cudaEvent_t> ev(world, nullptr);
    int rc = 0;
    for (size_t r = 0; r < world && !rc; r++) {
        rc = zkb_ntt4_scatter(plans[r], root, inverse, x_local[r]);
        if (rc) break;
        zkb_ctx* c = plans[r]->ctx;
        if (cudaEventCreateWithFlags(&ev[r], cudaEventDisableTiming) != cudaSuccess || cudaEventRecord(ev[r], c->stream) != cudaSuccess)
            rc = set_err(c, ZKB_ERR_CUDA, "ntt4: event record failed");
    }
    for (size_t r = 0; r < world && !rc; r++) {
        zkb_ctx* c = plans[r]->ctx;
        cudaSetDevice(c->device);
        for (size_t q = 0; q < world && !rc; q++)
            if (plans[q]->ctx->stream != c->stream && cudaStreamWaitEvent(c->stream, ev[q], 0) != cudaSuccess)
                rc = set_err(c, ZKB_ERR_CUDA, "ntt4: cudaStreamWaitEvent failed");
        if (!rc) rc = zkb_ntt4_finish(plans[r], out[r]);
    }
    for (auto e : ev) if (e) cudaEventDestroy(e);
    return rc;
}

// Convenience: the whole transform in one call (plans created and released inside; keep plans for repeated transforms).
int zkb_ntt_4step(zkb_ctx* const* ctxs, size_t world, const uint8_t root[16], int inverse, const void* const* x_local, size_t n_local,
                  void* const* out) {
    if (!ctxs || !root || !x_local || !out || world == 0 || world > ZKB_NTT_MAX_PEERS) return ZKB_ERR_ARG;
    std::vector<zkb_ntt4*> plans(world, nullptr);
    int rc = 0;
    for (size_t r = 0; r < world && !rc; r++) rc = zkb_ntt4_create(ctxs[r], (uint32_t)r, (uint32_t)world, n_local, &plans[r]);
    if (!rc && world > 1) rc = zkb_ntt4_connect_local(plans.data(), world);
    if (!rc) rc = zkb_ntt4_run(plans.data(), world, root, inverse, x_local, out);
    for (size_t r = 0; r < world; r++) if (plans[r]) zkb_ntt4_free(plans[r]);      // synchronises each context
    return rc;
}

// ---- configs[3]: independent trace columns over the GPUs of one box, one process (stark.rs:373-381 per register:
// fast_coset_evaluate then the commitment; here LDE + the whole FRI commit per column).  Column i runs on
// ctxs[i % n_ctx] (one host thread per context); no field data crosses NVLink.
int zkb_lde_commit_batch(zkb_ctx* const* ctxs, size_t n_ctx, const zkb_fri_params* p, const void* const* cols, size_t n_coeffs,
                         size_t ncols, uint8_t* roots_out) {
    if (!ctxs || !p || !cols || !roots_out || n_ctx == 0) return ZKB_ERR_ARG;
    const uint64_t rounds = zkb_fri_num_rounds(p);
    std::vector<int> rcs(n_ctx, 0);
    std::vector<std::thread> pool;
    for (size_t k = 0; k < n_ctx; k++) {
        pool.emplace_back([&, k]() {
            zkb_ctx* c = ctxs[k];
            try {
                for (size_t i = k; i < ncols && rcs[k] == 0; i += n_ctx) {
                    zkb_ps* ps = nullptr;
                    zkb_fri_layers* L = nullptr;
                    int rc = zkb_ps_create(nullptr, 0, 0, &ps);
                    if (!rc) rc = zkb_lde_fri_commit_ps(c, p, cols[i], n_coeffs, ps, &L);
                    for (uint64_t r = 0; r < rounds && !rc; r++) rc = zkb_fri_layer_root(L, r, roots_out + (i * rounds + r) * 64);
                    zkb_fri_layers_free(L);
                    zkb_ps_free(ps);
                    rcs[k] = rc;
                }
            } catch (...) { rcs[k] = set_err(c, ZKB_ERR_CUDA, "lde_commit_batch: out of memory on the host"); }
        });
    }
    for (auto& t : pool) t.join();
    for (size_t k = 0; k < n_ctx; k++) if (rcs[k]) return rcs[k];
    return 0;
}

}  // extern "C"
