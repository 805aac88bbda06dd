// Bulk (TMA) global -> shared copies with mbarrier completion: cp.async.bulk (UBLKCP in SASS).  One instruction moves a
// whole contiguous tile (16-byte aligned, size a multiple of 16) without staging through registers and without one
// LDGSTS per 16 bytes; the issuing thread arms the mbarrier with the byte count, everybody waits on its phase parity.
#pragma once
#include <cuda_runtime.h>

namespace tma {

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

// arm the barrier with `bytes` and start the copy (one thread).  The proxy fence orders the block's earlier
// generic-proxy accesses to the destination (made visible to this thread by a preceding __syncthreads) before it.
__device__ __forceinline__ void bulk_load(void *dst, const void *src, unsigned bytes, unsigned long long *bar)
{
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity)
{
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "TMA_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra TMA_DONE;\n"
      "bra TMA_WAIT;\n"
      "TMA_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

// usable when the tile is 16-byte aligned in global memory and its size is a multiple of 16 bytes
__device__ __forceinline__ bool bulk_ok(const void *src, size_t bytes)
{
  return ((reinterpret_cast<unsigned long long>(src) | (unsigned long long)bytes) & 15ull) == 0ull;
}

}  // namespace tma
