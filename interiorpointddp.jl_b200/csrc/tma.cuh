// Bulk asynchronous copies global -> shared memory through the TMA engine (cp.async.bulk, 1-D) with mbarrier
// completion, for staging whole per-knot records ahead of the warp that consumes them.
// Requirements of the instruction: 16-byte aligned source / destination, size a multiple of 16 bytes -- which is
// why trajectory and gains records are padded to an even number of doubles (layout.cuh).
// Under the CPU emulator (tests only) the copy happens synchronously at issue time and the wait is a warp barrier,
// so the staging logic (stage indices, phases, drain) is still exercised against the oracle.
#pragma once
#include "simt_compat.cuh"

namespace ipk {

#ifndef IPDDP_SIMT_EMU
IPDDP_D unsigned smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
IPDDP_D void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
IPDDP_D void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
IPDDP_D void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
IPDDP_D void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(dst)),
               "l"(src), "r"(bytes), "r"(smem_addr(bar))
               : "memory");
}
IPDDP_D void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_addr(bar)),
      "r"(parity)
      : "memory");
}
#else
inline void mbar_init(unsigned long long*, unsigned) {}
inline void mbar_init_fence() {}
inline void mbar_expect_tx(unsigned long long*, unsigned) {}
inline void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long*) { memcpy(dst, src, bytes); }
// the wait is a warp-level rendezvous in the emulator: the issuing lane's (synchronous) copy precedes its own wait in
// program order, so no lane reads a stage before it was filled, whatever order the emulator runs the lanes in
inline void mbar_wait(unsigned long long*, unsigned) { emu::yield_barrier(); }
#endif

}  // namespace ipk
