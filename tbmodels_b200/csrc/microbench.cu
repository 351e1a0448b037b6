// microbench.cu -- measured FP64 roofline denominators.
//
// MEASURED_PEAKS.json (driver-written) holds HBM GB/s and dense bf16 TFLOP/s only; the H(k) build runs
// on the FP64 pipes, so bench.py measures the two FP64 peaks on the same GPU in the same run:
//   kind 0: register-resident mma.sync.m8n8k4.f64 chains (SASS DMMA.8x8x4), 16 independent accumulators/warp
//   kind 1: register-resident DFMA chains, 16 independent accumulators/thread
// Both run one CTA of 256 threads x 4 per SM for `iters` iterations and are timed with CUDA events.
#include "tbk_kernels.h"

namespace tbk {

namespace {

constexpr int ACC = 16;

__global__ void __launch_bounds__(256) dmma_peak_kernel(double* out, int iters, double seed) {
    double c0[ACC], c1[ACC];
#pragma unroll
    for (int i = 0; i < ACC; ++i) {
        c0[i] = seed * i;
        c1[i] = -seed * i;
    }
    double a = seed + threadIdx.x * 1e-9, b = seed - threadIdx.x * 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ACC; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c0[i]), "+d"(c1[i])
                         : "d"(a), "d"(b));
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < ACC; ++i) s += c0[i] + c1[i];
    if (s == 123.456) out[0] = s;  // keep the chain alive
}

__global__ void __launch_bounds__(256) dfma_peak_kernel(double* out, int iters, double seed) {
    double c[ACC];
#pragma unroll
    for (int i = 0; i < ACC; ++i) c[i] = seed * i;
    const double a = 1.0 + seed * 1e-9, b = seed * 1e-12 + threadIdx.x * 1e-15;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ACC; ++i) c[i] = fma(c[i], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < ACC; ++i) s += c[i];
    if (s == 123.456) out[0] = s;
}

// both at once: do DMMA and DFMA share a datapath?  (reported sum of both flop counts)
__global__ void __launch_bounds__(256) mixed_peak_kernel(double* out, int iters, double seed) {
    double c0[ACC], c1[ACC], f[ACC];
#pragma unroll
    for (int i = 0; i < ACC; ++i) {
        c0[i] = seed * i;
        c1[i] = -seed * i;
        f[i] = seed * i;
    }
    double a = seed + threadIdx.x * 1e-9, b = seed - threadIdx.x * 1e-9;
    const double fa = 1.0 + seed * 1e-9, fb = seed * 1e-12;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ACC; ++i) {
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c0[i]), "+d"(c1[i])
                         : "d"(a), "d"(b));
            f[i] = fma(f[i], fa, fb);
        }
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < ACC; ++i) s += c0[i] + c1[i] + f[i];
    if (s == 123.456) out[0] = s;
}

}  // namespace

double measure_fp64_peak(int kind, int iters) {
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return -1.0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    double* out = nullptr;
    if (cudaMalloc(&out, 8) != cudaSuccess) return -1.0;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int blocks = sms * 4;
    double best = -1.0;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        if (kind == 0) dmma_peak_kernel<<<blocks, 256>>>(out, iters, 1.0);
        else if (kind == 1) dfma_peak_kernel<<<blocks, 256>>>(out, iters, 1.0);
        else mixed_peak_kernel<<<blocks, 256>>>(out, iters, 1.0);
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) { best = -1.0; break; }
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        double flops;
        if (kind == 0) flops = (double)blocks * 8 /*warps*/ * (double)iters * ACC * 512.0;   // 8x8x4 MACs x 2
        else if (kind == 1) flops = (double)blocks * 256 * (double)iters * ACC * 2.0;
        else flops = (double)blocks * 8 * (double)iters * ACC * 512.0 + (double)blocks * 256 * (double)iters * ACC * 2.0;
        const double tf = flops / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;  // first rep is warm-up
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    return best;
}

}  // namespace tbk
