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

// Stage dispatch: calls f(StageTag<S>{}) with S = stage type `type` of the (chain) model M.  A plain model is a chain of
// one type: the call is direct, nothing is looked up.
template <class S> struct StageTag { typedef S type; };
template <class M, class F>
IPDDP_D void for_stage(int type, F&& f) {
  static_assert(M::NSTAGE >= 1 && M::NSTAGE <= MAX_STAGE_TYPES, "1..4 stage types");
  if constexpr (M::NSTAGE == 1) {
    (void)type;
    f(StageTag<typename M::template Stage<0>>{});
  } else if constexpr (M::NSTAGE == 2) {
    if (type == 0) f(StageTag<typename M::template Stage<0>>{}); else f(StageTag<typename M::template Stage<1>>{});
  } else if constexpr (M::NSTAGE == 3) {
    if (type == 0) f(StageTag<typename M::template Stage<0>>{});
    else if (type == 1) f(StageTag<typename M::template Stage<1>>{});
    else f(StageTag<typename M::template Stage<2>>{});
  } else {
    if (type == 0) f(StageTag<typename M::template Stage<0>>{});
    else if (type == 1) f(StageTag<typename M::template Stage<1>>{});
    else if (type == 2) f(StageTag<typename M::template Stage<2>>{});
    else f(StageTag<typename M::template Stage<3>>{});
  }
}
#define IPDDP_STAGE(tag) typename decltype(tag)::type
// stride of state-sized arrays (lambda, the x outputs): the largest state a knot of the (chain) model can carry
template <class M> struct Dims { static constexpr int NS = cmax(M::NX, M::NXN, M::NXT); };

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

// Per-warp scratch shared by the warp-per-instance kernels that evaluate merit terms (k_forward, k_check):
// u[NU] | chunk[32] | finite-bound index bytes (2*NU per stage type, padded to 8 doubles) | 4 per-knot arrays of N doubles
// (k_init / k_admit keep the two state buffers of their rollout there: at least 2*NS doubles).
template <class M> struct MeritLayout {
  static constexpr int NUP = M::NU > 0 ? M::NU : 1;
  static constexpr int BIDX = 2 * NUP;                       // bytes of one stage type's index list
  static constexpr int FIXED = NUP + 32 + ((BIDX * M::NSTAGE + 7) / 8);
  static IPDDP_BOTH int per_warp_doubles(int N) { return FIXED + (4 * N > 2 * Dims<M>::NS ? 4 * N : 2 * Dims<M>::NS); }
};

// finite-bound index lists in the reference's accumulation order (lower indices, then upper indices;
// src/data/methods.jl:45-53), one per stage type; built by lane 0, counts broadcast
template <class M> struct BoundLists {
  int nlo[M::NSTAGE], nbd[M::NSTAGE];
  const unsigned char* bidx;
  IPDDP_D const unsigned char* list(int type) const { return bidx + type * MeritLayout<M>::BIDX; }
};
template <class M>
IPDDP_D BoundLists<M> warp_bound_lists(const DevView& v, int b, unsigned char* bidx, int lane) {
  BoundLists<M> bl;
  bl.bidx = bidx;
  for (int type = 0; type < M::NSTAGE; ++type) {
    int a = 0, q = 0;
    if (lane == 0) {
      const double* lo = v.lower_of(b, type);
      const double* up = v.upper_of(b, type);
      unsigned char* dst = bidx + type * MeritLayout<M>::BIDX;
      const int nu = M::NSTAGE == 1 ? M::NU : v.snu[type];
      for (int i = 0; i < nu; ++i) if (!is_inf(lo[i])) dst[q++] = (unsigned char)i;
      a = q;
      for (int i = 0; i < nu; ++i) if (!is_inf(up[i])) dst[q++] = (unsigned char)i;
    }
    bl.nlo[type] = __shfl_sync(IPDDP_FULL_MASK, a, 0);
    bl.nbd[type] = __shfl_sync(IPDDP_FULL_MASK, q, 0);
  }
  __syncwarp();
  return bl;
}

// Objective, constraints (written to the record), theta = sum_t |c_t|_1 and the barrier Lagrangian of one trajectory
// set, in the reference's evaluation order: objective (src/objectives.jl:37-46), constraint! (src/data/methods.jl:20-32),
// constraint_violation_1norm (:69-76), barrier_lagrangian! (:34-67: one running accumulator over (t, finite-lower idx,
// finite-upper idx), times mu, plus J, plus sum_t dot(c_t, phi_t)).  The per-knot terms (l_t, c_t, |c_t|_1, c_t'phi_t)
// are evaluated with lane = knot, the barrier logs 32 at a time, and every sum is then accumulated sequentially in
// exactly that order.  All lanes return the same J, theta, L.
// recs: the instance's knot records of the trajectory set to evaluate (record t at recs + t * TR).
template <class M>
IPDDP_D void warp_eval_metrics(const DevView& v, double* recs, int Nb, double mu, const double* p, const BoundLists<M>& bls,
                               double* chunk, double* p_l, double* p_th, double* p_d,
                               int lane, double* Jout, double* theta_out, double* Lout) {
  constexpr int STRIDE = Rec<M>::STRIDE;
  for (int t = lane; t < Nb; t += 32) {
    double* r = recs + (size_t)t * STRIDE;
    double Jp;
    if (t < Nb - 1) {
      const int type = v.type_of(t);
      for_stage<M>(type, [&](auto tag) {
        typedef IPDDP_STAGE(tag) S;
        typedef Rec<S> R;
        constexpr int NX = S::NX, NU = S::NU, NC = S::NC;
        double x[NX];
#pragma unroll
        for (int i = 0; i < NX; ++i) x[i] = r[R::X + i];
        double u[NU > 0 ? NU : 1], c[NC > 0 ? NC : 1];
#pragma unroll
        for (int i = 0; i < NU; ++i) u[i] = r[R::U + i];
        S::cost(x, u, p, &Jp);
        double n1 = 0.0;
        if (NC > 0) {
          S::con(x, u, p, c);
          if (v.compl_mask[type]) {
#pragma unroll
            for (int i = 0; i < NC; ++i) if ((v.compl_mask[type] >> i) & 1ull) c[i] -= mu;
          }
#pragma unroll
          for (int i = 0; i < NC; ++i) { r[R::C + i] = c[i]; n1 += fabs(c[i]); }
        }
        p_th[t] = n1;
        double ph[NC > 0 ? NC : 1];
#pragma unroll
        for (int i = 0; i < NC; ++i) ph[i] = r[R::PHI + i];
        p_d[t] = dot4c<NC>(c, 1, ph, 1);
      });
    } else {
      typedef typename M::Terminal T;
      double x[T::NXT];
#pragma unroll
      for (int i = 0; i < T::NXT; ++i) x[i] = r[i];
      T::costN(x, p, &Jp);
      p_th[t] = 0.0;
      p_d[t] = 0.0;
    }
    p_l[t] = Jp;
  }
  __syncwarp();
  double Jn = 0.0, theta = 0.0;
  for (int t = 0; t < Nb; ++t) {
    Jn += p_l[t];
    if (t < Nb - 1 && (M::NSTAGE > 1 ? v.snc[v.type_of(t)] > 0 : M::NC > 0)) theta += p_th[t];
  }
  // barrier term: bl -= log(slack) over (t, lower idx..., upper idx...), one running accumulator
  double bl = 0.0;
  if constexpr (M::NSTAGE == 1) {
    typedef Rec<M> R;
    const int nlo = bls.nlo[0], nbd = bls.nbd[0];
    const unsigned char* bidx = bls.list(0);
    const int total = (Nb - 1) * nbd;
    for (int base = 0; base < total; base += 32) {
      const int q = base + lane;
      double lg = 0.0;
      if (q < total) {
        const int t = q / nbd, s = q - t * nbd;
        const double* r = recs + (size_t)t * STRIDE;
        const int i = bidx[s];
        lg = dm::log(s < nlo ? r[R::IL + i] : r[R::IU + i]);
      }
      chunk[lane] = lg;
      __syncwarp();
      const int cnt = (total - base) < 32 ? (total - base) : 32;
      for (int e = 0; e < cnt; ++e) bl -= chunk[e];
      __syncwarp();
    }
  } else {   // stage chain: the number of bounded controls changes with the knot's stage type -- one knot at a time
    for (int t = 0; t < Nb - 1; ++t) {
      const int type = v.type_of(t);
      const int nlo = bls.nlo[type], nbd = bls.nbd[type];
      const unsigned char* bidx = bls.list(type);
      const double* r = recs + (size_t)t * STRIDE;
      const int nx = v.snx[type], nu = v.snu[type], nc = v.snc[type];
      const int oIL = nx + nu + nc, oIU = oIL + nu;
      for (int base = 0; base < nbd; base += 32) {
        const int s = base + lane;
        double lg = 0.0;
        if (s < nbd) { const int i = bidx[s]; lg = dm::log(s < nlo ? r[oIL + i] : r[oIU + i]); }
        chunk[lane] = lg;
        __syncwarp();
        const int cnt = (nbd - base) < 32 ? (nbd - base) : 32;
        for (int e = 0; e < cnt; ++e) bl -= chunk[e];
        __syncwarp();
      }
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
