// peak.cu -- FP64 FMA throughput microbenchmark.  MEASURED_PEAKS.json (driver-written) holds
// HBM and bf16 tensor peaks only; the ODE stage is bounded by the FP64 pipe, so bench.py
// measures that denominator the same way (a kernel timed alone, best of several, CUDA events).
#include <cuda_runtime.h>

namespace {
// 8 independent DFMA chains per thread, 1024 iterations, 16 resident warps per SM-quarter.
__global__ void __launch_bounds__(512) fp64_fma_kernel(double* out, double a, double b, int iters) {
    double x0 = threadIdx.x * 1e-9, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
            x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
        }
    }
    double s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
    if (s == 123.456) out[0] = s;  // never true; keeps the chains alive
}
}  // namespace

// returns TFLOP/s (2 flops per FMA), best of `reps`
double cb_fp64_peak_measure(int n_sm, cudaStream_t st, int reps) {
    double* d = nullptr;
    if (cudaMalloc((void**)&d, 8) != cudaSuccess) return 0;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 2048, blocks = n_sm * 4, threads = 512;
    double best = 0;
    for (int r = 0; r < reps + 1; r++) {
        cudaEventRecord(e0, st);
        fp64_fma_kernel<<<blocks, threads, 0, st>>>(d, 0.999999, 1e-7, iters);
        cudaEventRecord(e1, st);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        double flops = 2.0 * 64.0 * iters * (double)blocks * threads;
        if (r > 0 && ms > 0) best = flops / (ms * 1e-3) / 1e12 > best ? flops / (ms * 1e-3) / 1e12 : best;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(d);
    return best;
}
