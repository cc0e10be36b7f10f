// tmem_probe.cu -- self-test of tmem.cuh on the device: every thread of a four-warp CTA fills its private TMEM row
// (256 doubles) with a known pattern through dynamic column addresses, reads it back, and the load / store
// instructions are timed with the SM clock.  cb_tmem_probe is called by the GPU tests before they trust the
// mode-per-lane evolve kernel.
#include <cuda_runtime.h>
#include <stdint.h>

#include "tmem.cuh"

__global__ void __launch_bounds__(128, 1) cb_tmem_probe_kernel(int* mismatches, long long* cycles) {
    __shared__ uint32_t base_smem;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) tmem::alloc(&base_smem, 512);
    tmem::fence_before_sync();
    __syncthreads();
    tmem::fence_after_sync();
    const uint32_t base = base_smem;
    int bad = 0;
    // 32 blocks of 8 doubles per thread
    for (int b = 0; b < 32; b++) {
        double v[8];
#pragma unroll
        for (int i = 0; i < 8; i++) v[i] = 1000.0 * blockIdx.x + threadIdx.x + (b * 8 + i) * 1e-3;
        tmem::st8(tmem::addr(base, warp, 16 * b), v);
    }
    tmem::wait_st();
    for (int b = 31; b >= 0; b--) {
        double v[8];
        tmem::ld8(tmem::addr(base, warp, 16 * b), v);
#pragma unroll
        for (int i = 0; i < 8; i++) bad += v[i] != 1000.0 * blockIdx.x + threadIdx.x + (b * 8 + i) * 1e-3;
    }
    // timing: 256 dependent-free loads, then 256 stores
    double acc = 0;
    long long t0 = clock64();
    for (int rep = 0; rep < 8; rep++)
        for (int b = 0; b < 32; b++) {
            double v[8];
            tmem::ld8(tmem::addr(base, warp, 16 * b), v);
            acc += v[rep & 7];
        }
    long long t1 = clock64();
    for (int rep = 0; rep < 8; rep++)
        for (int b = 0; b < 32; b++) {
            double v[8];
#pragma unroll
            for (int i = 0; i < 8; i++) v[i] = acc + i;
            tmem::st8(tmem::addr(base, warp, 16 * b), v);
        }
    tmem::wait_st();
    long long t2 = clock64();
    if (bad) atomicAdd(mismatches, bad);
    if (threadIdx.x == 0 && blockIdx.x == 0) { cycles[0] = (t1 - t0) / 256; cycles[1] = (t2 - t1) / 256; cycles[2] = (long long)acc; }
    tmem::fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem::dealloc(base, 512);
}

cudaError_t cb_tmem_probe_launch(int n_blocks, int* d_mismatches, long long* d_cycles, cudaStream_t stream) {
    cb_tmem_probe_kernel<<<n_blocks, 128, 0, stream>>>(d_mismatches, d_cycles);
    return cudaGetLastError();
}
