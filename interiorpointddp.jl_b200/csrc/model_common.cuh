// Common declarations for generated device model headers (csrc/models_gen/*.cuh).
#pragma once
#include "simt_compat.cuh"
#include "detmath.cuh"

// Stage chains (state / control sizes that change along the horizon): a model struct names its stage types through
// `template <int I> using Stage` (a plain model is a chain of one: itself); generated composites use these helpers.
namespace ipk {
template <int I, class... Ts> struct TypeAt;
template <class T0, class... Ts> struct TypeAt<0, T0, Ts...> { typedef T0 type; };
template <int I, class T0, class... Ts> struct TypeAt<I, T0, Ts...> { typedef typename TypeAt<I - 1, Ts...>::type type; };
constexpr int cmax(int a) { return a; }
template <class... R> constexpr int cmax(int a, int b, R... r) { return cmax(a > b ? a : b, r...); }
}  // namespace ipk

// One structurally non-zero entry (i,j) of a derivative matrix.  slot >= 0: index into the compact
// derivative tile that travels through HBM; slot < 0: the entry is the compile-time constant
// CONSTS[-1-slot].
struct MEntry {
  short slot;
  unsigned char i, j;
};

// read-only load of one table entry (4 bytes: slot | i << 16 | j << 24)
IPDDP_D MEntry ld_entry(const MEntry* p) {
  const unsigned w = IPDDP_LDG(reinterpret_cast<const unsigned*>(p));
  MEntry q;
  q.slot = (short)(w & 0xffffu);
  q.i = (unsigned char)((w >> 16) & 0xffu);
  q.j = (unsigned char)(w >> 24);
  return q;
}
