// probe.cu - integer-pipe issue-rate microbenchmarks (measurement tooling, built into
// lib/libzkb200_probe.so; not part of the product ABI).  bench.py runs them on the box to
// obtain the MEASURED integer roofline denominators the north star asks for:
//   kind 0: ALU pipe  - dependent chains of IADD3 / LOP3 (a = (a + b) ^ c), 2 ops per step
//   kind 1: FMA pipe  - dependent chains of IMAD        (a = a * b + c),   1 op per step
//   kind 2: both      - one ALU pair and two IMADs per step, interleaved
// 8 independent chains per thread hide the 4-cycle pipe latency.  ops/s = lane-operations
// per second over the whole GPU (one SASS instruction = 32 lane-ops).
#include <cuda_runtime.h>
#include <stdint.h>

#define CH 8
#define UNROLL 16

template <int KIND>
__global__ void __launch_bounds__(256) k_probe(uint32_t* out, uint32_t b, uint32_t c, int iters) {
    uint32_t a[CH];
#pragma unroll
    for (int i = 0; i < CH; i++) a[i] = threadIdx.x * 2654435761u + i * 40503u + blockIdx.x;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
#pragma unroll
            for (int i = 0; i < CH; i++) {
                if (KIND == 0) {
                    asm volatile("add.u32 %0, %0, %1;\n\txor.b32 %0, %0, %2;" : "+r"(a[i]) : "r"(b), "r"(c));
                } else if (KIND == 1) {
                    asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));
                } else {
                    if (i & 1) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));
                    else asm volatile("add.u32 %0, %0, %1;\n\txor.b32 %0, %0, %2;" : "+r"(a[i]) : "r"(b), "r"(c));
                }
            }
        }
    }
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < CH; i++) acc ^= a[i];
    if (acc == 0x12345678u) out[0] = acc;      // keeps the chains live
}

extern "C" int zkb_probe_int_pipe(int device, int kind, double* lane_ops_per_s, double* ms_out) {
    if (cudaSetDevice(device) != cudaSuccess) return -1;
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    uint32_t* d = nullptr;
    if (cudaMalloc(&d, 64) != cudaSuccess) return -1;
    const int iters = 2048, blocks = sms * 8, threads = 256;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 5; rep++) {
        cudaEventRecord(e0);
        if (kind == 0) k_probe<0><<<blocks, threads>>>(d, 0x9E3779B9u, 0x7F4A7C15u, iters);
        else if (kind == 1) k_probe<1><<<blocks, threads>>>(d, 0x9E3779B9u, 0x7F4A7C15u, iters);
        else k_probe<2><<<blocks, threads>>>(d, 0x9E3779B9u, 0x7F4A7C15u, iters);
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) { cudaFree(d); return -1; }
        float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    double per_thread = (double)iters * UNROLL * (kind == 0 ? CH * 2 : (kind == 1 ? CH : (CH / 2) * 2 + (CH / 2)));
    *lane_ops_per_s = per_thread * blocks * threads / (best * 1e-3);
    if (ms_out) *ms_out = best;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(d);
    return 0;
}
