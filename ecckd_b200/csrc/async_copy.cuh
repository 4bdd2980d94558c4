// async_copy.cuh - mbarrier and bulk asynchronous copy (TMA 1-D, SASS UBLKCP) helpers shared by the
// streaming kernels (spectral_reduce.cu, flux_solver.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ecckd {

__device__ __forceinline__ unsigned smem_u32(const void* p) {
  return static_cast<unsigned>(__cvta_generic_to_shared(p));
}
// The mbarrier / bulk-copy helpers take shared-window byte addresses (smem_u32 of the dynamic
// shared-memory base, computed once per thread, plus an offset): converting a generic pointer at
// every call costs an S2R of the CTA-in-cluster id each time.
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned bar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}\n" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
      "@p bra.uni WAIT_DONE;\n\t"
      "bra.uni WAIT_LOOP;\n"
      "WAIT_DONE:\n\t}\n" ::"r"(bar),
      "r"(parity), "r"(0x989680u)  // suspend-time hint: sleep in hardware instead of spinning
      : "memory");
}
// 1-D bulk asynchronous copy global -> shared (SASS: UBLKCP), completion on an mbarrier
__device__ __forceinline__ void bulk_g2s(unsigned smem_dst, const void* gmem_src, unsigned bytes, unsigned bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_dst),
               "l"(gmem_src), "r"(bytes), "r"(bar)
               : "memory");
}
// the same with an L2 cache policy (optical depths are read once: evict-first keeps the
// per-wavenumber data that every group re-reads resident in L2)
__device__ __forceinline__ void bulk_g2s_hint(unsigned smem_dst, const void* gmem_src, unsigned bytes, unsigned bar,
                                              uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;\n" ::
          "r"(smem_dst),
      "l"(gmem_src), "r"(bytes), "r"(bar), "l"(policy)
      : "memory");
}
__device__ __forceinline__ uint64_t policy_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;\n" : "=l"(pol));
  return pol;
}
// per-thread asynchronous copies global -> shared (SASS: LDGSTS), completion by commit / wait groups
__device__ __forceinline__ void cp_async8(unsigned smem_dst, const void* gmem_src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(smem_dst), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async16(unsigned smem_dst, const void* gmem_src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;\n" ::"r"(smem_dst), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }


}  // namespace ecckd
