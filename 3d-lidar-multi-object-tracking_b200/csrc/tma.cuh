// tma.cuh -- minimal inline-PTX wrappers for the Blackwell bulk async copy engine (TMA, non-tensor form) and the
// shared-memory mbarrier it signals.  `cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes` shows up in
// SASS as UBLKCP, `mbarrier.arrive.expect_tx` as SYNCS.ARRIVE.TRANS64 (B300_MICROARCH.md, "cp.async/LDGSTS").
#pragma once
#include <cstdint>

namespace lmot {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

// producer: announce `bytes` of async traffic on the barrier (and arrive once)
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

// one bulk DMA global -> shared; completes `bytes` on the barrier.  dst/src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void bulk_copy_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// consumer: wait for the phase with the given parity; traps instead of hanging the GPU if the copy never lands
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  for (uint32_t spin = 0; !ok; ++spin) {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(ok)
                 : "r"(smem_u32(bar)), "r"(parity)
                 : "memory");
    if (spin > (1u << 26)) __trap();
  }
}

}  // namespace lmot
