// Compilation shim.  Under nvcc this is plain CUDA.  When IPDDP_SIMT_EMU is defined (ONLY by
// tests/emu/build_emu.py, never by the product build) the same kernel sources compile with g++ against
// a fiber-per-lane SIMT emulator so kernel logic can be debugged in a container without a GPU.
#pragma once

#ifdef IPDDP_SIMT_EMU
// the emulator header is force-included (-include) by the test-only build and provides every macro below
#else
#include <cuda_runtime.h>
#include <math.h>
#define IPDDP_HD __device__ __forceinline__
#define IPDDP_TABLE static __device__ const
#define IPDDP_D __device__ __forceinline__
#define IPDDP_BOTH __host__ __device__
#define IPDDP_D2LL(x) __double_as_longlong(x)
#define IPDDP_LL2D(x) __longlong_as_double(x)
#define IPDDP_FMA(a, b, c) __fma_rn((a), (b), (c))
#define IPDDP_LDG(p) __ldg(p)
#define IPDDP_LDCG(p) __ldcg(p)
#define IPDDP_PREFETCH_L2(p) asm volatile("prefetch.global.L2 [%0];" ::"l"(p))
#define IPDDP_DYN_SMEM(type, name) extern __shared__ __align__(16) unsigned char name##_raw[]; type* name = reinterpret_cast<type*>(name##_raw)
#define IPDDP_LAUNCH(kernel, grid, block, smem, stream, ...) kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#endif
