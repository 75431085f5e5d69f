// TEST INFRASTRUCTURE ONLY -- a minimal SIMT emulator so that the CUDA kernel sources under
// interiorpointddp.jl_b200/csrc compile with g++ and run on the CPU of a container without a GPU.
// It exists to debug kernel *logic* (indexing, operation order) against the oracle before spending
// GPU time; it is built by tests/emu/build_emu.py into tests/emu/libipddp_emu.so and loaded only by
// tests/test_emu_*.py.  The product library (libipddp_b200.so) is never built from or linked with it.
//
// Model: every thread of a block is a fiber (own stack, hand-rolled x86-64 context switch); fibers of
// a block run round-robin on one OS thread, __syncwarp/__syncthreads/shuffles yield to the scheduler,
// which releases the barrier once every live fiber has arrived (warp-level barriers: one step per scheduler
// round; __syncthreads: a fiber stays parked until every live fiber of the block is parked at __syncthreads).  Blocks are distributed over OpenMP
// threads.  Floating point: fma() is a hardware FMA (-mfma), everything else uncontracted
// (-ffp-contract=off), matching nvcc -fmad=false + explicit __fma_rn.
#pragma once
#define IPDDP_SIMT_EMU 1
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <functional>
#include <vector>

struct dim3 { unsigned x, y, z; dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };
struct double4 { double x, y, z, w; };
struct double2 { double x, y; };

namespace emu {
struct Fiber { void* sp; char* stack; bool done; bool at_block_barrier; };
struct BlockState {
  std::vector<Fiber> fibers;
  std::vector<char*> stacks;   // one malloc per fiber (never moved: the frames on them are live between switches)
  std::vector<unsigned char> smem;
  unsigned long long xchg[1024];
  void* sched_sp;
  int current;
  const std::function<void()>* body;
};
extern thread_local BlockState* tls_block;
extern thread_local dim3 tls_threadIdx, tls_blockIdx, tls_blockDim, tls_gridDim;
void yield_barrier();
void yield_block_barrier();
void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body);
inline void* cur_smem() { return tls_block->smem.data(); }
}  // namespace emu

#define threadIdx (emu::tls_threadIdx)
#define blockIdx (emu::tls_blockIdx)
#define blockDim (emu::tls_blockDim)
#define gridDim (emu::tls_gridDim)

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#ifndef __restrict__
#define __restrict__ __restrict
#endif

#define IPDDP_HD inline
#define IPDDP_D inline
#define IPDDP_BOTH
#define IPDDP_TABLE static const
static inline long long emu_d2ll(double x) { long long u; memcpy(&u, &x, 8); return u; }
static inline double emu_ll2d(long long u) { double x; memcpy(&x, &u, 8); return x; }
#define IPDDP_D2LL(x) emu_d2ll(x)
#define IPDDP_LL2D(x) emu_ll2d(x)
#define IPDDP_FMA(a, b, c) fma((a), (b), (c))
#define IPDDP_LDG(p) (*(p))
#define IPDDP_LDCG(p) (*(p))
#define IPDDP_PREFETCH_L2(p) ((void)(p))
#define IPDDP_DYN_SMEM(type, name) type* name = reinterpret_cast<type*>(emu::cur_smem())
#define IPDDP_LAUNCH(kernel, grid, block, smem, stream, ...) \
  emu::launch(dim3(grid), dim3(block), (smem), [&]() { kernel(__VA_ARGS__); })

static inline void __syncwarp(unsigned = 0xffffffffu) { emu::yield_barrier(); }
static inline void __syncthreads() { emu::yield_block_barrier(); }
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int lanemask) {
  static_assert(sizeof(T) <= 8, "shuffle payload");
  emu::BlockState* b = emu::tls_block;
  const int me = emu::tls_threadIdx.x;
  unsigned long long raw = 0;
  memcpy(&raw, &v, sizeof(T));
  b->xchg[me] = raw;
  emu::yield_barrier();
  raw = b->xchg[(me & ~31) | ((me ^ lanemask) & 31)];
  emu::yield_barrier();
  T r;
  memcpy(&r, &raw, sizeof(T));
  return r;
}
template <class T> static inline T __shfl_sync(unsigned, T v, int src) {
  static_assert(sizeof(T) <= 8, "shuffle payload");
  emu::BlockState* b = emu::tls_block;
  const int me = emu::tls_threadIdx.x;
  unsigned long long raw = 0;
  memcpy(&raw, &v, sizeof(T));
  b->xchg[me] = raw;
  emu::yield_barrier();
  raw = b->xchg[(me & ~31) | (src & 31)];
  emu::yield_barrier();
  T r;
  memcpy(&r, &raw, sizeof(T));
  return r;
}
static inline unsigned __ballot_sync(unsigned, int pred) {
  emu::BlockState* b = emu::tls_block;
  const int me = emu::tls_threadIdx.x;
  b->xchg[me] = pred ? 1ull : 0ull;
  emu::yield_barrier();
  unsigned r = 0;
  const int w0 = me & ~31;
  const int nthr = (int)emu::tls_blockDim.x;
  for (int q = 0; q < 32 && w0 + q < nthr; ++q) if (!b->fibers[w0 + q].done && b->xchg[w0 + q]) r |= (1u << q);
  emu::yield_barrier();
  return r;
}
static inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0; }
static inline int __all_sync(unsigned m, int pred) { return __ballot_sync(m, !pred) == 0; }
static inline unsigned __reduce_max_sync(unsigned, unsigned v) {
  emu::BlockState* b = emu::tls_block;
  const int me = emu::tls_threadIdx.x;
  b->xchg[me] = v;
  emu::yield_barrier();
  unsigned r = 0;
  const int w0 = me & ~31;
  for (int q = 0; q < 32; ++q) { unsigned x = (unsigned)b->xchg[w0 + q]; if (x > r) r = x; }
  emu::yield_barrier();
  return r;
}
static inline unsigned __reduce_min_sync(unsigned, unsigned v) {
  emu::BlockState* b = emu::tls_block;
  const int me = emu::tls_threadIdx.x;
  b->xchg[me] = v;
  emu::yield_barrier();
  unsigned r = 0xffffffffu;
  const int w0 = me & ~31;
  for (int q = 0; q < 32; ++q) { unsigned x = (unsigned)b->xchg[w0 + q]; if (x < r) r = x; }
  emu::yield_barrier();
  return r;
}
static inline int __ffs(unsigned x) { return x ? __builtin_ctz(x) + 1 : 0; }
static inline int __ffsll(long long x) { return x ? __builtin_ctzll((unsigned long long)x) + 1 : 0; }
static inline double __hiloint2double(int hi, int lo) { return emu_ll2d(((long long)(unsigned)hi << 32) | (unsigned)lo); }
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((unsigned long long)a * b) >> 32); }
static inline int __double2hiint(double x) { return (int)(emu_d2ll(x) >> 32); }
static inline int __double2loint(double x) { return (int)(emu_d2ll(x) & 0xffffffffll); }
static inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }

// ---- CUDA runtime shims (host memory stands in for device memory)
typedef int cudaError_t;
typedef void* cudaStream_t;
typedef void* cudaEvent_t;
enum { cudaSuccess = 0 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice };
enum { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
enum { cudaDevAttrMaxSharedMemoryPerBlockOptin = 97 };
static inline cudaError_t cudaDeviceGetAttribute(int* v, int, int) { *v = 227 * 1024; return 0; }
struct cudaDeviceProp { int multiProcessorCount; };
static inline const char* cudaGetErrorString(cudaError_t) { return "emulated"; }
static inline cudaError_t cudaGetLastError() { return 0; }
static inline cudaError_t cudaSetDevice(int) { return 0; }
static inline cudaError_t cudaMalloc(void** p, size_t n) { *p = calloc(n ? n : 1, 1); return *p ? 0 : 2; }
static inline cudaError_t cudaFree(void* p) { free(p); return 0; }
static inline cudaError_t cudaMallocHost(void** p, size_t n) { *p = calloc(n ? n : 1, 1); return *p ? 0 : 2; }
static inline cudaError_t cudaFreeHost(void* p) { free(p); return 0; }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memcpy(d, s, n); return 0; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { memcpy(d, s, n); return 0; }
static inline cudaError_t cudaMemset(void* d, int v, size_t n) { memset(d, v, n); return 0; }
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) { memset(d, v, n); return 0; }
static inline cudaError_t cudaStreamCreate(cudaStream_t* s) { *s = (void*)1; return 0; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return 0; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
static inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = (void*)1; return 0; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t) { return 0; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return 0; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return 0; }
enum { cudaErrorNotReady = 600 };
static inline cudaError_t cudaEventQuery(cudaEvent_t) { return 0; }
static inline cudaError_t cudaDeviceSynchronize() { return 0; }
static inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return 0; }
static inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) { p->multiProcessorCount = 1; return 0; }
