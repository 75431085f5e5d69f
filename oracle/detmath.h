// ORACLE (test infrastructure) -- deterministic elementary functions, plain C.
//
// Why: the reference evaluates sin/cos/tan/log/^ through Julia's own libm ports; glibc and CUDA
// libdevice each differ from it (and from each other) by 1-2 ulp, and IPDDP2's filter / inertia
// branch decisions amplify last-bit differences into different iteration counts.
// The oracle therefore fixes ONE evaluation order for every elementary function, using only IEEE
// +,-,*,/ (no FMA contraction: compile with -ffp-contract=off).  The CUDA product carries an
// independent transcription of the same published algorithms (csrc/detmath.cuh); tests/ check the
// two bit-for-bit on the GPU and check this file against libm to <= 2 ulp on the CPU.
//
// Algorithms: classic fdlibm-style argument reduction + minimax polynomials (Sun fdlibm e_log.c,
// e_exp.c, k_sin.c, k_cos.c, e_rem_pio2.c medium-argument path), restated from the published method.
#pragma once
#include <stdint.h>
#include <string.h>
#include <math.h>

static inline uint64_t dm_bits(double x) { uint64_t u; memcpy(&u, &x, 8); return u; }
static inline double dm_from_bits(uint64_t u) { double x; memcpy(&x, &u, 8); return x; }

// ---------------------------------------------------------------- log
static inline double dm_log(double x) {
  const double ln2_hi = 6.93147180369123816490e-01, ln2_lo = 1.90821492927058770002e-10;
  const double Lg1 = 6.666666666666735130e-01, Lg2 = 3.999999999940941908e-01, Lg3 = 2.857142874366239149e-01,
               Lg4 = 2.222219843214978396e-01, Lg5 = 1.818357216161805012e-01, Lg6 = 1.531383769920937332e-01,
               Lg7 = 1.479819860511658591e-01;
  uint64_t ix = dm_bits(x);
  int k = 0;
  if (x != x) return x;                                   // NaN
  if (x < 0.0) return dm_from_bits(0x7ff8000000000000ull); // NaN
  if (x == 0.0) return -INFINITY;
  if (ix == 0x7ff0000000000000ull) return x;              // +Inf
  if (ix < 0x0010000000000000ull) {                       // subnormal: scale up by 2^54
    x = x * 18014398509481984.0;
    ix = dm_bits(x);
    k -= 54;
  }
  uint32_t hx = (uint32_t)(ix >> 32);
  uint32_t lx = (uint32_t)ix;
  k += (int)(hx >> 20) - 1023;
  hx &= 0x000fffffu;
  uint32_t i = (hx + 0x95f64u) & 0x100000u;
  hx = hx | (i ^ 0x3ff00000u);                            // normalise x or x/2
  k += (int)(i >> 20);
  x = dm_from_bits(((uint64_t)hx << 32) | lx);
  double f = x - 1.0;
  double hfsq = 0.5 * f * f;
  double s = f / (2.0 + f);
  double z = s * s;
  double w = z * z;
  double t1 = w * (Lg2 + w * (Lg4 + w * Lg6));
  double t2 = z * (Lg1 + w * (Lg3 + w * (Lg5 + w * Lg7)));
  double R = t2 + t1;
  double dk = (double)k;
  return s * (hfsq + R) + dk * ln2_lo - hfsq + f + dk * ln2_hi;
}

// ---------------------------------------------------------------- exp
static inline double dm_exp(double x) {
  const double ln2HI = 6.93147180369123816490e-01, ln2LO = 1.90821492927058770002e-10,
               invln2 = 1.44269504088896338700e+00;
  const double P1 = 1.66666666666666019037e-01, P2 = -2.77777777770155933842e-03, P3 = 6.61375632143793436117e-05,
               P4 = -1.65339022054652515390e-06, P5 = 4.13813679705723846039e-08;
  if (x != x) return x;
  if (x > 709.782712893383973096) return INFINITY;
  if (x < -745.13321910194110842) return 0.0;
  double fk = rint(x * invln2);
  int k = (int)fk;
  double hi = x - fk * ln2HI;
  double lo = fk * ln2LO;
  double r = hi - lo;
  double t = r * r;
  double c = r - t * (P1 + t * (P2 + t * (P3 + t * (P4 + t * P5))));
  double y = 1.0 - ((lo - (r * c) / (2.0 - c)) - hi);
  // scale by 2^k in two exact steps so that subnormal results round once
  int k1 = k / 2, k2 = k - k1;
  double s1 = dm_from_bits((uint64_t)(1023 + k1) << 52);
  double s2 = dm_from_bits((uint64_t)(1023 + k2) << 52);
  return y * s1 * s2;
}

// ---------------------------------------------------------------- pow (x >= 0)
// exp(y*log(x)); relative error ~ |y log x| * 1e-16, ample for the barrier-parameter schedule and the
// switching condition (reference src/solve.jl:62, src/forward_pass.jl:40-41, src/inertia_correction.jl:263).
static inline double dm_pow(double x, double y) {
  if (y == 0.0) return 1.0;
  if (x != x || y != y) return x + y;
  if (x == 0.0) return y > 0.0 ? 0.0 : INFINITY;
  if (x < 0.0) return dm_from_bits(0x7ff8000000000000ull);
  return dm_exp(y * dm_log(x));
}

// ---------------------------------------------------------------- sin / cos / tan
static inline double dm_ksin(double x, double y) {
  const double S1 = -1.66666666666666324348e-01, S2 = 8.33333333332248946124e-03, S3 = -1.98412698298579493134e-04,
               S4 = 2.75573137070700676789e-06, S5 = -2.50507602534068634195e-08, S6 = 1.58969099521155010221e-10;
  double z = x * x;
  double v = z * x;
  double r = S2 + z * (S3 + z * (S4 + z * (S5 + z * S6)));
  return x - ((z * (0.5 * y - v * r) - y) - v * S1);
}
static inline double dm_kcos(double x, double y) {
  const double C1 = 4.16666666666666019037e-02, C2 = -1.38888888888741095749e-03, C3 = 2.48015872894767294178e-05,
               C4 = -2.75573143513906633035e-07, C5 = 2.08757232129817482790e-09, C6 = -1.13596475577881948265e-11;
  double z = x * x;
  double r = z * (C1 + z * (C2 + z * (C3 + z * (C4 + z * (C5 + z * C6)))));
  double hz = 0.5 * z;
  double w = 1.0 - hz;
  return w + (((1.0 - w) - hz) + (z * r - x * y));
}
// reduce x to y0+y1 in [-pi/4, pi/4], return quadrant.  Exact to ~118 bits for |x| < 2^20*pi/2;
// beyond that the result is still deterministic but loses accuracy (not reached by any workload).
static inline int dm_rem_pio2(double x, double* y0, double* y1) {
  const double invpio2 = 6.36619772367581382433e-01, pio2_1 = 1.57079632673412561417e+00,
               pio2_2 = 6.07710050630396597660e-11, pio2_2t = 2.02226624879595063154e-21;
  double fn = rint(x * invpio2);
  double t = x - fn * pio2_1;
  double w = fn * pio2_2;
  double r = t - w;
  w = fn * pio2_2t - ((t - r) - w);
  *y0 = r - w;
  *y1 = (r - *y0) - w;
  double q = fn - 4.0 * rint(fn * 0.25);   // fn mod 4 in {-2,-1,0,1,2}
  int n = (int)q;
  return n & 3;
}
static inline double dm_sin(double x) {
  if (!(fabs(x) <= 1.7976931348623157e308)) return x - x;  // Inf/NaN -> NaN
  double y0, y1;
  int n = dm_rem_pio2(x, &y0, &y1);
  switch (n) {
    case 0: return dm_ksin(y0, y1);
    case 1: return dm_kcos(y0, y1);
    case 2: return -dm_ksin(y0, y1);
    default: return -dm_kcos(y0, y1);
  }
}
static inline double dm_cos(double x) {
  if (!(fabs(x) <= 1.7976931348623157e308)) return x - x;
  double y0, y1;
  int n = dm_rem_pio2(x, &y0, &y1);
  switch (n) {
    case 0: return dm_kcos(y0, y1);
    case 1: return -dm_ksin(y0, y1);
    case 2: return -dm_kcos(y0, y1);
    default: return dm_ksin(y0, y1);
  }
}
static inline double dm_tan(double x) { return dm_sin(x) / dm_cos(x); }

#define DM_SIN(x) dm_sin(x)
#define DM_COS(x) dm_cos(x)
#define DM_TAN(x) dm_tan(x)
#define DM_LOG(x) dm_log(x)
#define DM_EXP(x) dm_exp(x)
#define DM_POW(x, y) dm_pow(x, y)
