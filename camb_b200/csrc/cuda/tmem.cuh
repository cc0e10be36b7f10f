// tmem.cuh -- Tensor Memory (sm_100a) as thread-private vector storage.
//
// TMEM is 512 columns x 128 lanes x 32 bit per SM (256 KB).  The warp whose index in the CTA is w (mod 4) addresses
// lanes 32 (w mod 4) .. 32 (w mod 4) + 31 with the 32x32b shape: thread t of that warp reads / writes lane 32 (w mod 4) + t,
// i.e. every thread owns a private row of 512 words = 256 doubles.  The mode-per-lane evolve kernel (evolve_mpl.cu) keeps
// the Runge-Kutta stage vectors of its k-mode there: one tcgen05.ld / tcgen05.st moves 8 doubles per thread in a single
// instruction and costs neither registers between uses nor shared-memory bandwidth.  No tensor-core instruction is
// involved; allocation follows the usual protocol (one warp allocates, the address travels through shared memory).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace tmem {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// one fully active warp; ncols a power of two in [32, 512]; writes the base address (lane 0, first column) to *dst (shared)
__device__ __forceinline__ void alloc(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// Address of column `col` in the lane quarter of warp `warp_in_cta`: lane index in bits 31..16, column in bits 15..0.
__device__ __forceinline__ uint32_t addr(uint32_t base, int warp_in_cta, int col) {
    return base + (((uint32_t)(warp_in_cta & 3) * 32u) << 16) + (uint32_t)col;
}

// 8 doubles (16 columns) per thread in one instruction.  The whole warp must execute these together (.sync.aligned).
// tcgen05.ld is asynchronous: the destination registers are defined only after tcgen05.wait::ld.  A load is therefore
// split in two: ld8_issue() starts it into a raw register block, and -- after ONE wait_ld() for all outstanding blocks --
// ld8_take() hands the doubles out.  ld8_take() begins with an empty volatile asm that "rewrites" the sixteen registers,
// so neither the compiler nor ptxas can schedule a consumer of them ahead of the wait (volatile asms keep their order).
struct Raw8 { uint32_t r[16]; };
__device__ __forceinline__ void ld8_issue(uint32_t taddr, Raw8& q) {
    uint32_t* r = q.r;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void ld8_take(Raw8& q, double (&v)[8]) {
    uint32_t* r = q.r;
    asm volatile("" : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]), "+r"(r[9]),
                      "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
                 :
                 : "memory");
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] = __hiloint2double((int)r[2 * i + 1], (int)r[2 * i]);
}
// one block, start to finish
__device__ __forceinline__ void ld8(uint32_t taddr, double (&v)[8]) {
    Raw8 q;
    ld8_issue(taddr, q);
    wait_ld();
    ld8_take(q, v);
}
__device__ __forceinline__ void st8(uint32_t taddr, const double (&v)[8]) {
    uint32_t r[16];
#pragma unroll
    for (int i = 0; i < 8; i++) { r[2 * i] = (uint32_t)__double2loint(v[i]); r[2 * i + 1] = (uint32_t)__double2hiint(v[i]); }
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
                 "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
                 "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
                 : "memory");
}

}  // namespace tmem
