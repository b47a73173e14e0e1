// bulk_copy.cuh -- 1-D bulk asynchronous copies global -> shared memory (the TMA engine's cp.async.bulk, SASS: UBLKCP) with
// mbarrier completion (SYNCS).  One elected thread issues the copy, nobody spends issue slots on LDG / STS pairs, and
// the data arrives while the block does something else; the waiters spin on the barrier's phase bit.
// Requirements of the instruction: source, destination and size are multiples of 16 bytes.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace hj3d {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t arrivals) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(arrivals) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");      // make the initialised barrier visible to the async proxy
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t phase) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "HJ3D_WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra HJ3D_DONE_%=;\n"
      "bra HJ3D_WAIT_%=;\n"
      "HJ3D_DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(phase) : "memory");
}

}  // namespace hj3d
