// Per-model host launchers.  Each model translation unit (csrc/models/<name>.cu, or a generated plugin)
// instantiates the kernel templates for its Model_<name> struct and exports one vtable.
#pragma once
#include "kernels_thread.cuh"
#include "kernel_backward.cuh"
#include "kernel_forward.cuh"
#include "vtable.h"

#ifndef IPDDP_BW_EXTRA_SMEM
#define IPDDP_BW_EXTRA_SMEM 0   // A/B only: unused dynamic shared memory per warp of k_backward (lowers the resident warps per SM)
#endif

namespace ipk {

template <class M> struct Launch {
  static void init(const DevView& v, int warm, int b0, int nb, int* list_next, int* counters, cudaStream_t s) {
    if (nb <= 0) return;
    const size_t smem = (size_t)INIT_WARPS * MeritLayout<M>::per_warp_doubles(v.N) * sizeof(double);
    IPDDP_LAUNCH((k_init<M>), (nb + INIT_WARPS - 1) / INIT_WARPS, INIT_WARPS * 32, smem, s, v, warm, b0, nb, list_next,
                 counters);
  }
  static void admit(const DevView& v, const QueueView& q, const int* slots, int n, int inst0, int* list, int* counters,
                    cudaStream_t s) {
    if (n <= 0) return;
    const size_t smem = (size_t)INIT_WARPS * MeritLayout<M>::per_warp_doubles(v.N) * sizeof(double);
    IPDDP_LAUNCH((k_admit<M>), (n + INIT_WARPS - 1) / INIT_WARPS, INIT_WARPS * 32, smem, s, v, q, slots, n, inst0, list,
                 counters);
  }
  static long long smem_merit(int N) {
    const size_t a = FwLayout<M>::bytes(N), b = (size_t)CHK_WARPS * MeritLayout<M>::per_warp_doubles(N) * sizeof(double);
    return (long long)(a > b ? a : b);
  }
  static long long smem_merit_spec(int N) { return (long long)FwLayout<M>::spec_bytes(N); }
  static void derivs(const DevView& v, const ListView& list, cudaStream_t s) {
    const int n = list.total();
    if (n <= 0) return;
    const int th = 128;
    const long long total = (long long)n * v.N;
    IPDDP_LAUNCH((k_derivs<M>), (unsigned)((total + th - 1) / th), th, 0, s, v, list);
  }
  static void backward(const DevView& v, const ListView& list, cudaStream_t s) {
    const int n = list.total();
    if (n <= 0) return;
    if (n <= v.bw_spec_max) {
      const size_t smem = (size_t)BWS_WARPS * BwLayout<M>::BYTES + BWS_WARPS * 8 + BWS_WARPS * 2 * 4;
      IPDDP_LAUNCH((k_backward_spec<M>), n, BWS_WARPS * 32, smem, s, v, list);
      return;
    }
    IPDDP_LAUNCH((k_backward<M>), n, 32, BwLayout<M>::BYTES + IPDDP_BW_EXTRA_SMEM, s, v, list);
  }
  static void check(const DevView& v, const ListView& list, int* list_next, int* list_fwd, int* counters,
                    cudaStream_t s) {
    const int n = list.total();
    if (n <= 0) return;
    const size_t smem = (size_t)CHK_WARPS * MeritLayout<M>::per_warp_doubles(v.N) * sizeof(double);
    IPDDP_LAUNCH((k_check<M>), (n + CHK_WARPS - 1) / CHK_WARPS, CHK_WARPS * 32, smem, s, v, list, list_next, list_fwd,
                 counters);
  }
  static void forward(const DevView& v, const int* list_fwd, int n_upper, int* list_next, int* counters,
                      cudaStream_t s) {
    if (n_upper <= 0) return;
    if (n_upper <= v.fw_spec_max) {
      IPDDP_LAUNCH((k_forward_spec<M>), n_upper, FWS_WARPS * 32, FwLayout<M>::spec_bytes(v.N), s, v, list_fwd, list_next,
                   counters, 0);
      return;
    }
    if (v.list_sort && v.fw_spec_max > 0)   // the heaviest bucket of a bulk round, 8 step sizes at a time (see fwd_heavy_split)
      IPDDP_LAUNCH((k_forward_spec<M>), v.fw_spec_max, FWS_WARPS * 32, FwLayout<M>::spec_bytes(v.N), s, v, list_fwd,
                   list_next, counters, 1);
    IPDDP_LAUNCH((k_forward<M>), (n_upper + FW_WARPS - 1) / FW_WARPS, FW_WARPS * 32, FwLayout<M>::bytes(v.N), s, v,
                 list_fwd, list_next, counters);
  }
  // The speculative tail kernels trade work for latency (4 regularisation values / 8 step sizes at once): worth it exactly
  // while every active instance's CTA is resident at the same time, i.e. up to (CTAs per SM) x (SMs) instances.
  static void spec_caps(int N, int* bw_spec, int* fw_spec) {
    *bw_spec = 592; *fw_spec = 148;
#ifndef IPDDP_SIMT_EMU
    int dev = 0, sms = 0, nb = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return;
    const size_t smem_bw = (size_t)BWS_WARPS * BwLayout<M>::BYTES + BWS_WARPS * 8 + BWS_WARPS * 2 * 4;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_backward_spec<M>, BWS_WARPS * 32, smem_bw) == cudaSuccess && nb > 0)
      *bw_spec = nb * sms;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_forward_spec<M>, FWS_WARPS * 32, FwLayout<M>::spec_bytes(N)) == cudaSuccess && nb > 0)
      *fw_spec = nb * sms;
#else
    (void)N;
#endif
  }
  static int prepare(int max_optin_smem) {
    (void)max_optin_smem;
#ifndef IPDDP_SIMT_EMU
    // long horizons: the per-knot scratch of the merit evaluation may exceed the 48 KB default; ipddp_problem_create
    // checks the actual need of a horizon against the same limit
    const int big = max_optin_smem;
    if (cudaFuncSetAttribute(k_forward<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, big) != cudaSuccess) return -1;
    if (cudaFuncSetAttribute(k_forward_spec<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, big) != cudaSuccess) return -1;
    if (cudaFuncSetAttribute(k_check<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, big) != cudaSuccess) return -1;
    if (cudaFuncSetAttribute(k_init<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, big) != cudaSuccess) return -1;
    if (cudaFuncSetAttribute(k_admit<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, big) != cudaSuccess) return -1;
    // the sweep is occupancy bound through shared memory: ask for the largest shared-memory carveout
    if (cudaFuncSetAttribute(k_backward<M>, cudaFuncAttributePreferredSharedMemoryCarveout, 100) != cudaSuccess) return -1;
    if (cudaFuncSetAttribute(k_backward_spec<M>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             BWS_WARPS * BwLayout<M>::BYTES + 256) != cudaSuccess) return -1;
    if (BwLayout<M>::BYTES > 48 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(k_backward<M>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           BwLayout<M>::BYTES);
      if (e != cudaSuccess) return -1;
    }
#endif
    return 0;
  }
};

template <class M, int I> constexpr int stage_dim(int which) {
  if constexpr (I < M::NSTAGE) {
    typedef typename M::template Stage<I> S;
    return which == 0 ? S::NX : which == 1 ? S::NU : which == 2 ? S::NC : S::NXN;
  } else {
    return 0;
  }
}
#define IPDDP_STAGE_DIMS(w) {stage_dim<M, 0>(w), stage_dim<M, 1>(w), stage_dim<M, 2>(w), stage_dim<M, 3>(w)}

template <class M> const ModelVTable* make_vtable() {
  static const ModelVTable vt = {
      M::NAME, M::NX, M::NU, M::NC, M::NP, M::D_NSLOT, M::DN_NSLOT, M::VF_NSLOT, BwLayout<M>::BYTES,
      M::NSTAGE, M::NXT, Dims<M>::NS, IPDDP_STAGE_DIMS(0), IPDDP_STAGE_DIMS(1), IPDDP_STAGE_DIMS(2), IPDDP_STAGE_DIMS(3),
      &Launch<M>::init, &Launch<M>::derivs, &Launch<M>::backward, &Launch<M>::check, &Launch<M>::forward,
      &Launch<M>::admit, &Launch<M>::smem_merit, &Launch<M>::smem_merit_spec, &Launch<M>::prepare, &Launch<M>::spec_caps};
  return &vt;
}

}  // namespace ipk

#define IPDDP_REGISTER_MODEL(M, sym) \
  extern "C" const ModelVTable* sym() { return ipk::make_vtable<M>(); }
