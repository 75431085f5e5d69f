// TEST INFRASTRUCTURE ONLY -- runtime of the SIMT emulator (see cpu_simt.h).
#include "cpu_simt.h"

#include <omp.h>
#if defined(__SANITIZE_ADDRESS__)
#include <sanitizer/asan_interface.h>
#endif
#include <string.h>

#include <utility>

namespace emu {
thread_local BlockState* tls_block = nullptr;
thread_local dim3 tls_threadIdx, tls_blockIdx, tls_blockDim, tls_gridDim;
}

extern "C" void emu_switch(void** save_sp, void* new_sp);
asm(R"(
.text
.globl emu_switch
.type emu_switch,@function
emu_switch:
  pushq %rbp
  pushq %rbx
  pushq %r12
  pushq %r13
  pushq %r14
  pushq %r15
  movq %rsp, (%rdi)
  movq %rsi, %rsp
  popq %r15
  popq %r14
  popq %r13
  popq %r12
  popq %rbx
  popq %rbp
  ret
.size emu_switch,.-emu_switch
)");

namespace emu {

static const size_t kStack = 512 * 1024;

static void trampoline() {
  BlockState* b = tls_block;
  (*b->body)();
  Fiber& f = b->fibers[b->current];
  f.done = true;
  void* dummy;
  emu_switch(&dummy, b->sched_sp);
  abort();
}

void yield_barrier() {
  BlockState* b = tls_block;
  Fiber& f = b->fibers[b->current];
  emu_switch(&f.sp, b->sched_sp);
}

void yield_block_barrier() {
  BlockState* b = tls_block;
  Fiber& f = b->fibers[b->current];
  f.at_block_barrier = true;
  emu_switch(&f.sp, b->sched_sp);
}

static void run_block(BlockState& bs, dim3 grid, dim3 block, unsigned bx, size_t smem, const std::function<void()>& body) {
  const int T = block.x;
  bs.body = &body;
  if ((int)bs.fibers.size() < T) {
    bs.fibers.resize(T);
    while ((int)bs.stacks.size() < T) bs.stacks.push_back((char*)malloc(kStack));
  }
#if defined(__SANITIZE_ADDRESS__)
  __asan_unpoison_memory_region(bs.smem.data(), bs.smem.size());
#endif
  if (bs.smem.size() < smem + 64) bs.smem.resize(smem + 64);
#if defined(__SANITIZE_ADDRESS__)
  // shared memory beyond what the launch asked for is poisoned: an overrun of the dynamic shared-memory layout is reported
  __asan_poison_memory_region(bs.smem.data() + smem, bs.smem.size() - smem);
#endif
  for (int t = 0; t < T; ++t) {
    Fiber& f = bs.fibers[t];
    f.done = false;
    f.at_block_barrier = false;
    f.stack = bs.stacks[t];
#if defined(__SANITIZE_ADDRESS__)
    __asan_unpoison_memory_region(f.stack, kStack);   // redzones of the previous fiber's frames (it never returned)
#endif
    uintptr_t top = ((uintptr_t)(f.stack + kStack)) & ~(uintptr_t)15;
    void** A = (void**)(top - 16);
    *A = (void*)&trampoline;
    void** sp = A - 6;
    for (int q = 0; q < 6; ++q) sp[q] = nullptr;
    f.sp = (void*)sp;
  }
  tls_block = &bs;
  tls_blockIdx = dim3(bx, 0, 0);
  tls_blockDim = block;
  tls_gridDim = grid;
  // Lane order inside a scheduler round (IPDDP_EMU_ORDER = fwd | rev | rand): a fiber runs from one barrier to the next
  // before the following fiber starts, so data exchanged through shared memory WITHOUT a barrier in between is seen
  // "new" by later fibers and "old" by earlier ones.  Race-free kernels give identical results for every order;
  // tests/test_emu_parity.py::test_emulated_lane_order_independence uses this as its race check.
  const char* ord_env = getenv("IPDDP_EMU_ORDER");
  const int ord = !ord_env ? 0 : !strcmp(ord_env, "rev") ? 1 : !strcmp(ord_env, "rand") ? 2 : 0;
  std::vector<int> order(T);
  for (int t = 0; t < T; ++t) order[t] = (ord == 1) ? T - 1 - t : t;
  unsigned long long rng = 0x9E3779B97F4A7C15ull ^ ((unsigned long long)bx * 0xD1B54A32D192ED03ull);
  int alive = T;
  while (alive > 0) {
    alive = 0;
    int parked = 0;
    if (ord == 2)
      for (int t = T - 1; t > 0; --t) {   // Fisher-Yates with xorshift64
        rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17;
        std::swap(order[t], order[(int)(rng % (unsigned long long)(t + 1))]);
      }
    for (int q = 0; q < T; ++q) {
      const int t = order[q];
      Fiber& f = bs.fibers[t];
      if (f.done) continue;
      if (f.at_block_barrier) { alive++; parked++; continue; }
      bs.current = t;
      tls_threadIdx = dim3(t, 0, 0);
      emu_switch(&bs.sched_sp, f.sp);
      if (!f.done) { alive++; if (f.at_block_barrier) parked++; }
    }
    if (alive > 0 && parked == alive)
      for (int t = 0; t < T; ++t) bs.fibers[t].at_block_barrier = false;
  }
}

void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body) {
  const int G = (int)grid.x;
#pragma omp parallel
  {
    static thread_local BlockState bs;
#pragma omp for schedule(dynamic, 1)
    for (int bx = 0; bx < G; ++bx) run_block(bs, grid, block, (unsigned)bx, smem, body);
  }
}

}  // namespace emu
