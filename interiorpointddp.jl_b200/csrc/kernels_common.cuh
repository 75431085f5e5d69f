// Device helpers shared by all IPDDP2 kernels.
#pragma once
#include "layout.cuh"
#include "model_common.cuh"

#define IPDDP_FULL_MASK 0xffffffffu
#define IPDDP_EPS 2.220446049250313e-16

namespace ipk {

// Julia's max/min propagate NaN (the reference relies on max(), norm(.,Inf); src/solve.jl:107-180)
IPDDP_D double jmax(double a, double b) { return (a != a || b != b) ? dm::nan_() : (a > b ? a : b); }
IPDDP_D double jmin(double a, double b) { return (a != a || b != b) ? dm::nan_() : (a < b ? a : b); }
IPDDP_D bool finite(double x) { return fabs(x) <= 1.7976931348623157e308; }
IPDDP_D bool is_inf(double x) { return fabs(x) > 1.7976931348623157e308 && x == x; }

// The one summation order used for every BLAS-like contraction: 4 interleaved FMA partial sums
// (element i goes to partial i mod 4), combined as (s0+s1)+(s2+s3).
IPDDP_D double dot4(int n, const double* a, int sa, const double* b, int sb) {
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
  int i = 0;
  for (; i + 3 < n; i += 4) {
    s0 = IPDDP_FMA(a[(i + 0) * sa], b[(i + 0) * sb], s0);
    s1 = IPDDP_FMA(a[(i + 1) * sa], b[(i + 1) * sb], s1);
    s2 = IPDDP_FMA(a[(i + 2) * sa], b[(i + 2) * sb], s2);
    s3 = IPDDP_FMA(a[(i + 3) * sa], b[(i + 3) * sb], s3);
  }
  if (i < n) s0 = IPDDP_FMA(a[i * sa], b[i * sb], s0);
  if (i + 1 < n) s1 = IPDDP_FMA(a[(i + 1) * sa], b[(i + 1) * sb], s1);
  if (i + 2 < n) s2 = IPDDP_FMA(a[(i + 2) * sa], b[(i + 2) * sb], s2);
  return (s0 + s1) + (s2 + s3);
}
template <int NN> IPDDP_D double dot4c(const double* a, int sa, const double* b, int sb) {
  double s[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
  for (int i = 0; i < NN; ++i) s[i & 3] = IPDDP_FMA(a[i * sa], b[i * sb], s[i & 3]);
  return (s[0] + s[1]) + (s[2] + s[3]);
}

// Objective, constraints (written to the record), theta = sum_t |c_t|_1 and the barrier Lagrangian of
// one trajectory set, in the reference's evaluation order: objective (src/objectives.jl:37-46),
// constraint! (src/data/methods.jl:20-32), constraint_violation_1norm (:69-76),
// barrier_lagrangian! (:34-67: one running accumulator over (t, finite-lower idx, finite-upper idx),
// times mu, plus J, plus sum_t dot(c_t, phi_t)).
template <class M>
IPDDP_D void eval_metrics(const DevView& v, int set, int b, int Nb, double mu, double* Jout, double* theta_out,
                          double* Lout) {
  typedef Rec<M> R;
  const double* p = v.p + (size_t)b * (M::NP > 0 ? M::NP : 1);
  const double* lo = v.lower + (size_t)b * M::NU;
  const double* up = v.upper + (size_t)b * M::NU;
  double J = 0.0, theta = 0.0, bl = 0.0;
  for (int t = 0; t < Nb; ++t) {
    double* r = v.rec(set, b, t);
    double x[M::NX];
#pragma unroll
    for (int i = 0; i < M::NX; ++i) x[i] = r[R::X + i];
    double Jp;
    if (t < Nb - 1) {
      double u[M::NU > 0 ? M::NU : 1], c[M::NC > 0 ? M::NC : 1];
#pragma unroll
      for (int i = 0; i < M::NU; ++i) u[i] = r[R::U + i];
      M::cost(x, u, p, &Jp);
      if (M::NC > 0) {
        M::con(x, u, p, c);
        if (v.compl_mask) {
#pragma unroll
            for (int i = 0; i < M::NC; ++i) if ((v.compl_mask >> i) & 1ull) c[i] -= mu;
          }
        double n1 = 0.0;
#pragma unroll
        for (int i = 0; i < M::NC; ++i) { r[R::C + i] = c[i]; n1 += fabs(c[i]); }
        theta += n1;
      }
#pragma unroll
      for (int i = 0; i < M::NU; ++i)
        if (!is_inf(lo[i])) bl -= dm::log(r[R::IL + i]);
#pragma unroll
      for (int i = 0; i < M::NU; ++i)
        if (!is_inf(up[i])) bl -= dm::log(r[R::IU + i]);
    } else {
      M::costN(x, p, &Jp);
    }
    J += Jp;
  }
  bl *= mu;
  bl += J;
  for (int t = 0; t < Nb - 1; ++t) {
    const double* r = v.rec(set, b, t);
    bl += dot4c<M::NC>(r + R::C, 1, r + R::PHI, 1);
  }
  bl += 0.0;  // terminal stage: dot of two empty vectors
  *Jout = J;
  *theta_out = theta;
  *Lout = bl;
}

IPDDP_D void reset_filter(const DevView& v, int b) {
  v.filter[(size_t)(0 * IPDDP_FILTER_CAPACITY + 0) * v.B + b] = v.sdv(SD_THETA_MAX, b);
  v.filter[(size_t)(1 * IPDDP_FILTER_CAPACITY + 0) * v.B + b] = -dm::inf();
  v.siv(SI_FILTER_N, b) = 1;
  v.siv(SI_STATUS, b) = 0;
}

}  // namespace ipk
