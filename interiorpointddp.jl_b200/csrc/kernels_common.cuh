// Device helpers shared by all IPDDP2 kernels.
#pragma once
#include "layout.cuh"
#include "model_common.cuh"

#define IPDDP_FULL_MASK 0xffffffffu
#define IPDDP_EPS 2.220446049250313e-16

#ifndef IPDDP_FWD_HEAVY_L
#define IPDDP_FWD_HEAVY_L 16   // trial steps of the last line search that put an instance into the heaviest forward bucket
#endif

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

// Per-warp scratch shared by the warp-per-instance kernels that evaluate merit terms (k_forward, k_check):
// u[NU] | chunk[32] | finite-bound index bytes (2*NU, padded to 8 doubles) | 4 per-knot arrays of N doubles.
template <class M> struct MeritLayout {
  static constexpr int NUP = M::NU > 0 ? M::NU : 1;
  static constexpr int FIXED = NUP + 32 + ((2 * NUP + 7) / 8);
  static IPDDP_BOTH int per_warp_doubles(int N) { return FIXED + 4 * N; }
};

// finite-bound index list in the reference's accumulation order (lower indices, then upper indices;
// src/data/methods.jl:45-53); built by lane 0, counts broadcast
template <class M>
IPDDP_D void warp_bound_list(const double* lo, const double* up, unsigned char* bidx, int lane, int& nlo, int& nbd) {
  int a = 0, q = 0;
  if (lane == 0) {
    for (int i = 0; i < M::NU; ++i) if (!is_inf(lo[i])) bidx[q++] = (unsigned char)i;
    a = q;
    for (int i = 0; i < M::NU; ++i) if (!is_inf(up[i])) bidx[q++] = (unsigned char)i;
  }
  nlo = __shfl_sync(IPDDP_FULL_MASK, a, 0);
  nbd = __shfl_sync(IPDDP_FULL_MASK, q, 0);
  __syncwarp();
}

// Warp-parallel eval_metrics: the per-knot terms (l_t, c_t written to the record, |c_t|_1, c_t'phi_t) are evaluated
// with lane = knot, the barrier logs 32 at a time, and every sum is then accumulated sequentially in exactly the
// order eval_metrics uses -- bit-identical results.  All lanes return the same J, theta, L.
// recs: the instance's knot records of the trajectory set to evaluate (record t at recs + t * TR).
template <class M>
IPDDP_D void warp_eval_metrics(const DevView& v, double* recs, int Nb, double mu, const double* p, int nlo, int nbd,
                               const unsigned char* bidx, double* chunk, double* p_l, double* p_th, double* p_d,
                               int lane, double* Jout, double* theta_out, double* Lout) {
  typedef Rec<M> R;
  constexpr int NX = M::NX, NU = M::NU, NC = M::NC;
  for (int t = lane; t < Nb; t += 32) {
    double* r = recs + (size_t)t * R::STRIDE;
    double x[NX];
#pragma unroll
    for (int i = 0; i < NX; ++i) x[i] = r[R::X + i];
    double Jp;
    if (t < Nb - 1) {
      double u[NU > 0 ? NU : 1], c[NC > 0 ? NC : 1];
#pragma unroll
      for (int i = 0; i < NU; ++i) u[i] = r[R::U + i];
      M::cost(x, u, p, &Jp);
      double n1 = 0.0;
      if (NC > 0) {
        M::con(x, u, p, c);
        if (v.compl_mask) {
#pragma unroll
          for (int i = 0; i < M::NC; ++i) if ((v.compl_mask >> i) & 1ull) c[i] -= mu;
        }
#pragma unroll
        for (int i = 0; i < NC; ++i) { r[R::C + i] = c[i]; n1 += fabs(c[i]); }
      }
      p_th[t] = n1;
      double ph[NC > 0 ? NC : 1];
#pragma unroll
      for (int i = 0; i < NC; ++i) ph[i] = r[R::PHI + i];
      p_d[t] = dot4c<NC>(c, 1, ph, 1);
    } else {
      M::costN(x, p, &Jp);
      p_th[t] = 0.0;
      p_d[t] = 0.0;
    }
    p_l[t] = Jp;
  }
  __syncwarp();
  double Jn = 0.0, theta = 0.0;
  for (int t = 0; t < Nb; ++t) { Jn += p_l[t]; if (t < Nb - 1 && NC > 0) theta += p_th[t]; }
  // barrier term: bl -= log(slack) over (t, lower idx..., upper idx...), one running accumulator
  double bl = 0.0;
  {
    const int total = (Nb - 1) * nbd;
    for (int base = 0; base < total; base += 32) {
      const int q = base + lane;
      double lg = 0.0;
      if (q < total) {
        const int t = q / nbd, s = q - t * nbd;
        const double* r = recs + (size_t)t * R::STRIDE;
        const int i = bidx[s];
        lg = dm::log(s < nlo ? r[R::IL + i] : r[R::IU + i]);
      }
      chunk[lane] = lg;
      __syncwarp();
      const int cnt = (total - base) < 32 ? (total - base) : 32;
      for (int e = 0; e < cnt; ++e) bl -= chunk[e];
      __syncwarp();
    }
  }
  bl *= mu;
  bl += Jn;
  for (int t = 0; t < Nb; ++t) bl += p_d[t];
  *Jout = Jn;
  *theta_out = theta;
  *Lout = bl;
}

// the instance in slot b has terminated: flag it and, in queue mode, hand the slot to the retire / admit step
IPDDP_D void mark_done(const DevView& v, int b, int* counters) {
  v.siv(SI_DONE, b) = 1;
  if (v.done_list) v.done_list[atomicAdd(&counters[CNT_DONE], 1)] = b;
}

// append instance b to the next round's list / to this round's forward list (bucket = expected work, see layout.cuh)
IPDDP_D void append_next(const DevView& v, int* list_next, int* counters, int b) {
  const int sw = v.siv(SI_LASTSW, b);
  const int k = !v.list_sort ? 3 : sw >= 4 ? 0 : sw == 3 ? 1 : sw == 2 ? 2 : 3;
  list_next[(size_t)k * v.B + atomicAdd(&counters[CNT_NEXT + k], 1)] = b;
}
IPDDP_D void append_fwd(const DevView& v, int* list_fwd, int* counters, int b) {
  const int l = v.siv(SI_LASTROLL, b);
  const int k = !v.list_sort ? 3 : l >= IPDDP_FWD_HEAVY_L ? 0 : l >= 6 ? 1 : l >= 3 ? 2 : 3;
  list_fwd[(size_t)k * v.B + atomicAdd(&counters[CNT_FWD + k], 1)] = b;
}
// the forward list's view, built on the device from the counters k_check left
IPDDP_D ListView fwd_view(const DevView& v, const int* list_fwd, const int* counters) {
  ListView l;
  l.base = list_fwd; l.stride = v.B;
  for (int k = 0; k < LIST_BUCKETS; ++k) l.n[k] = counters[CNT_FWD + k];
  return l;
}

// Rounds too big for the speculative line search as a whole still hand their heaviest forward bucket (instances whose
// last line search took >= 16 trial steps: the stragglers that reject dozens of step sizes per iteration) to
// k_forward_spec, 8 step sizes at a time, while k_forward takes the rest; both kernels decide from the same counters.
IPDDP_D bool fwd_heavy_split(const DevView& v, const ListView& lf) {
  return v.list_sort && v.fw_spec_max > 0 && lf.n[0] > 0 && lf.n[0] <= v.fw_spec_max;
}

IPDDP_D void reset_filter(const DevView& v, int b) {
  v.filter[(size_t)(0 * IPDDP_FILTER_CAPACITY + 0) * v.B + b] = v.sdv(SD_THETA_MAX, b);
  v.filter[(size_t)(1 * IPDDP_FILTER_CAPACITY + 0) * v.B + b] = -dm::inf();
  v.siv(SI_FILTER_N, b) = 1;
  v.siv(SI_STATUS, b) = 0;
}

}  // namespace ipk
