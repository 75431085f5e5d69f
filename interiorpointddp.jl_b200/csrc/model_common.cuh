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
