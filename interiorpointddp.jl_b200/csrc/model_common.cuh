// Common declarations for generated device model headers (csrc/models_gen/*.cuh).
#pragma once
#include "simt_compat.cuh"
#include "detmath.cuh"

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
