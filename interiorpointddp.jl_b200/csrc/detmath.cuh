// Deterministic FP64 elementary functions for the device path.
//
// CUDA libdevice sin/cos/log differ from any CPU libm by 1-2 ulp; IPDDP2's filter and inertia
// decisions amplify such differences into different iteration counts.  These routines use only
// IEEE add/mul/div (the translation unit is compiled with -fmad=false, so nothing is contracted)
// and therefore return the same bits on every device -- and the same bits as the independent CPU
// transcription the test-suite's oracle carries, which tests/test_gpu_detmath.py checks bit-for-bit.
// Method: classic fdlibm-style reduction + minimax polynomials (log: k*ln2 + log(1+f) via s=f/(2+f);
// exp: k*ln2 split hi/lo + rational P1..P5; sin/cos: two-term Cody-Waite pi/2 reduction + kernel
// polynomials with tail correction).
#pragma once
#include "simt_compat.cuh"

namespace dm {

IPDDP_HD unsigned long long bits(double x) { return (unsigned long long)IPDDP_D2LL(x); }
IPDDP_HD double from_bits(unsigned long long u) { return IPDDP_LL2D((long long)u); }
IPDDP_HD double inf() { return from_bits(0x7ff0000000000000ull); }
IPDDP_HD double nan_() { return from_bits(0x7ff8000000000000ull); }

IPDDP_HD double log(double x) {
  const double ln2_hi = 6.93147180369123816490e-01, ln2_lo = 1.90821492927058770002e-10;
  const double Lg1 = 6.666666666666735130e-01, Lg2 = 3.999999999940941908e-01, Lg3 = 2.857142874366239149e-01,
               Lg4 = 2.222219843214978396e-01, Lg5 = 1.818357216161805012e-01, Lg6 = 1.531383769920937332e-01,
               Lg7 = 1.479819860511658591e-01;
  unsigned long long ix = bits(x);
  int k = 0;
  if (x != x) return x;
  if (x < 0.0) return nan_();
  if (x == 0.0) return -inf();
  if (ix == 0x7ff0000000000000ull) return x;
  if (ix < 0x0010000000000000ull) {
    x = x * 18014398509481984.0;
    ix = bits(x);
    k -= 54;
  }
  unsigned int hx = (unsigned int)(ix >> 32);
  unsigned int lx = (unsigned int)ix;
  k += (int)(hx >> 20) - 1023;
  hx &= 0x000fffffu;
  unsigned int i = (hx + 0x95f64u) & 0x100000u;
  hx = hx | (i ^ 0x3ff00000u);
  k += (int)(i >> 20);
  x = from_bits(((unsigned long long)hx << 32) | lx);
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

IPDDP_HD double exp(double x) {
  const double ln2HI = 6.93147180369123816490e-01, ln2LO = 1.90821492927058770002e-10,
               invln2 = 1.44269504088896338700e+00;
  const double P1 = 1.66666666666666019037e-01, P2 = -2.77777777770155933842e-03, P3 = 6.61375632143793436117e-05,
               P4 = -1.65339022054652515390e-06, P5 = 4.13813679705723846039e-08;
  if (x != x) return x;
  if (x > 709.782712893383973096) return inf();
  if (x < -745.13321910194110842) return 0.0;
  double fk = rint(x * invln2);
  int k = (int)fk;
  double hi = x - fk * ln2HI;
  double lo = fk * ln2LO;
  double r = hi - lo;
  double t = r * r;
  double c = r - t * (P1 + t * (P2 + t * (P3 + t * (P4 + t * P5))));
  double y = 1.0 - ((lo - (r * c) / (2.0 - c)) - hi);
  int k1 = k / 2, k2 = k - k1;
  double s1 = from_bits((unsigned long long)(1023 + k1) << 52);
  double s2 = from_bits((unsigned long long)(1023 + k2) << 52);
  return y * s1 * s2;
}

IPDDP_HD double pow(double x, double y) {
  if (y == 0.0) return 1.0;
  if (x != x || y != y) return x + y;
  if (x == 0.0) return y > 0.0 ? 0.0 : inf();
  if (x < 0.0) return nan_();
  return dm::exp(y * dm::log(x));
}

IPDDP_HD double ksin(double x, double y) {
  const double S1 = -1.66666666666666324348e-01, S2 = 8.33333333332248946124e-03, S3 = -1.98412698298579493134e-04,
               S4 = 2.75573137070700676789e-06, S5 = -2.50507602534068634195e-08, S6 = 1.58969099521155010221e-10;
  double z = x * x;
  double v = z * x;
  double r = S2 + z * (S3 + z * (S4 + z * (S5 + z * S6)));
  return x - ((z * (0.5 * y - v * r) - y) - v * S1);
}
IPDDP_HD double kcos(double x, double y) {
  const double C1 = 4.16666666666666019037e-02, C2 = -1.38888888888741095749e-03, C3 = 2.48015872894767294178e-05,
               C4 = -2.75573143513906633035e-07, C5 = 2.08757232129817482790e-09, C6 = -1.13596475577881948265e-11;
  double z = x * x;
  double r = z * (C1 + z * (C2 + z * (C3 + z * (C4 + z * (C5 + z * C6)))));
  double hz = 0.5 * z;
  double w = 1.0 - hz;
  return w + (((1.0 - w) - hz) + (z * r - x * y));
}
IPDDP_HD int rem_pio2(double x, double* y0, double* y1) {
  const double invpio2 = 6.36619772367581382433e-01, pio2_1 = 1.57079632673412561417e+00,
               pio2_2 = 6.07710050630396597660e-11, pio2_2t = 2.02226624879595063154e-21;
  double fn = rint(x * invpio2);
  double t = x - fn * pio2_1;
  double w = fn * pio2_2;
  double r = t - w;
  w = fn * pio2_2t - ((t - r) - w);
  *y0 = r - w;
  *y1 = (r - *y0) - w;
  double q = fn - 4.0 * rint(fn * 0.25);
  int n = (int)q;
  return n & 3;
}
IPDDP_HD double sin(double x) {
  if (!(fabs(x) <= 1.7976931348623157e308)) return x - x;
  double y0, y1;
  int n = rem_pio2(x, &y0, &y1);
  double s = ksin(y0, y1), c = kcos(y0, y1);
  return n == 0 ? s : (n == 1 ? c : (n == 2 ? -s : -c));
}
IPDDP_HD double cos(double x) {
  if (!(fabs(x) <= 1.7976931348623157e308)) return x - x;
  double y0, y1;
  int n = rem_pio2(x, &y0, &y1);
  double s = ksin(y0, y1), c = kcos(y0, y1);
  return n == 0 ? c : (n == 1 ? -s : (n == 2 ? -c : s));
}
IPDDP_HD double tan(double x) { return dm::sin(x) / dm::cos(x); }

}  // namespace dm

#define DM_SIN(x) dm::sin(x)
#define DM_COS(x) dm::cos(x)
#define DM_TAN(x) dm::tan(x)
#define DM_LOG(x) dm::log(x)
#define DM_EXP(x) dm::exp(x)
#define DM_POW(x, y) dm::pow(x, y)
