// Host-side table of one compiled model: dimensions + kernel launchers.
#pragma once
#include "layout.cuh"

struct ModelVTable {
  const char* name;
  int nx, nu, nc, np, d_nslot, dn_nslot, vf_nslot, smem_backward;   // nx/nu/nc: maxima over the stage types of a chain
  // stage chain: number of stage types, their sizes (state, control, constraint, next state), the terminal state size and
  // the stride of state-sized arrays
  int nstage, nxt, ns;
  int snx[MAX_STAGE_TYPES], snu[MAX_STAGE_TYPES], snc[MAX_STAGE_TYPES], snxn[MAX_STAGE_TYPES];
  void (*init)(const DevView&, int warm, int b0, int nb, int* list_next, int* counters, cudaStream_t);
  void (*derivs)(const DevView&, const ListView& list, cudaStream_t);
  void (*backward)(const DevView&, const ListView& list, cudaStream_t);
  void (*check)(const DevView&, const ListView& list, int* list_next, int* list_fwd, int* counters, cudaStream_t);
  void (*forward)(const DevView&, const int* list_fwd, int n_upper, int* list_next, int* counters, cudaStream_t);
  // queue mode: admit `n` queued instances inst0.. into the slots `slots` (NULL: 0..n-1), appending them at list[0..n)
  void (*admit)(const DevView&, const QueueView&, const int* slots, int n, int inst0, int* list, int* counters, cudaStream_t);
  // dynamic shared memory of the merit kernels (k_forward / k_check, and the speculative k_forward_spec) for N knots
  long long (*smem_merit)(int N);
  long long (*smem_merit_spec)(int N);
  int (*prepare)(int max_optin_smem);
  // how many instances the speculative tail kernels can hold in ONE wave on the current device (CTAs per SM x SMs)
  void (*spec_caps)(int N, int* bw_spec, int* fw_spec);
};
