// Host-side table of one compiled model: dimensions + kernel launchers.
#pragma once
#include "layout.cuh"

struct ModelVTable {
  const char* name;
  int nx, nu, nc, np, d_nslot, dn_nslot, vf_nslot, smem_backward;
  void (*init)(const DevView&, int warm, int b0, int nb, int* list_next, int* counters, cudaStream_t);
  void (*derivs)(const DevView&, const int* list, int n, cudaStream_t);
  void (*backward)(const DevView&, const int* list, int n, cudaStream_t);
  void (*check)(const DevView&, const int* list, int n, int* list_next, int* list_fwd, int* counters, cudaStream_t);
  void (*forward)(const DevView&, const int* list_fwd, int n_upper, int* list_next, int* counters, cudaStream_t);
  int (*prepare)();
};
