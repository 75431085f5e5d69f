// k_forward / k_forward_spec: forward_pass! (reference src/forward_pass.jl:1-57).
//
// Per trial step a warp runs
//   (1) rollout! (:98-153), sequential in time: the 3nu+nc affine updates of a knot (u, phi, zl, zu: one
//       4-term dot with delta-x each) are spread over the lanes, gains and nominal values of knot t+1 are
//       prefetched into registers while knot t is computed, the dynamics f(x,u) is evaluated redundantly by
//       every lane (x lives replicated in registers, u goes through shared memory); the fraction-to-boundary
//       test (:59-85) is fused in (warp vote, early exit: same accept/reject decision),
//   (2) the merit evaluation, parallel in time: lane = knot evaluates l_t, c_t (stored), |c_t|_1 and c_t'phi_t;
//       the barrier logs are evaluated 32 at a time in exactly the order the reference accumulates them,
//   (3) ordered (sequential) summations so that J, theta and the barrier Lagrangian carry the reference's
//       summation order (src/objectives.jl:37-46, src/data/methods.jl:34-76),
//   (4) the filter / switching / Armijo / sufficient-decrease logic, uniform over the warp.
// Accepting a step flips the instance's nominal/trial record set (update_nominal_trajectory!,
// src/data/methods.jl:78-91), augments the filter (src/solve.jl:81,95-99) and appends the instance to the
// next round's list.
//
// Two launch shapes with bit-identical results:
//   k_forward       one warp per instance (4 instances per CTA), trial steps one after another: bulk rounds;
//   k_forward_spec  one CTA of FWS_WARPS warps per instance, for rounds with few active instances (the lock-step
//                   tail, where straggler instances backtrack through dozens of step sizes, most of them rejected by
//                   the filter / Armijo tests): warp w evaluates step size gamma / 2^w completely -- (1) into a
//                   private trial record buffer (DevView::spec_traj), then (2)(3) -- and the verdicts
//                   (rollout status, theta, L, J) are consumed by one thread in the order of the reference's
//                   backtracking loop, so counters, statuses and the accepted step are exactly the sequential
//                   ones.  The accepted candidate's records are copied into the instance's trial set.
#pragma once
#include "kernels_common.cuh"
#include "tma.cuh"

#ifndef IPDDP_FW_TMA
#define IPDDP_FW_TMA 1            // 1: the rollout's per-knot gains + nominal records are staged in shared memory by TMA bulk copies
#endif

namespace ipk {

constexpr int FW_WARPS = 4;
constexpr int FWS_WARPS = 8;
constexpr int FW_TMA_STAGES = 3;  // per-warp ring of staged knots (two knots in flight ahead of the one being consumed)
#ifndef IPDDP_FW_MINBLOCKS
#define IPDDP_FW_MINBLOCKS 3      // resident CTAs per SM the register allocation of k_forward is held to (168 registers; measured: 2 -> 65 ms, 3 -> 53 ms, 4 -> 77 ms per 12 bulk rounds)
#endif

template <class M> struct FwLayout : MeritLayout<M> {
  // TMA staging ring of one warp: FW_TMA_STAGES x (gains record | nominal record), then the stages' mbarriers
  static constexpr int K_ = M::NU + M::NC;
  static constexpr int GP = ((K_ + 2 * M::NU) * (M::NX + 1) + 1) & ~1;     // == DevView::G
  static constexpr int STAGE = GP + Rec<M>::STRIDE;
  static constexpr int TMA_DOUBLES = IPDDP_FW_TMA ? FW_TMA_STAGES * STAGE + ((FW_TMA_STAGES + 1) & ~1) : 0;   // even: every warp's ring stays 16-byte aligned
  static size_t bytes(int N) { return (size_t)FW_WARPS * (MeritLayout<M>::per_warp_doubles(N) + TMA_DOUBLES) * sizeof(double); }
  // spec kernel: per-warp merit scratch, per-warp results (theta, L, J), next base step, rollout verdicts, 2 control ints
  static IPDDP_BOTH int spec_doubles(int N) {   // rounded up to even: the TMA staging rings follow at a 16-byte boundary
    return (FWS_WARPS * MeritLayout<M>::per_warp_doubles(N) + 3 * FWS_WARPS + 1 + (FWS_WARPS + 2 + 1) / 2 + 1) & ~1;
  }
  static size_t spec_bytes(int N) { return (size_t)(spec_doubles(N) + FWS_WARPS * TMA_DOUBLES) * sizeof(double); }
};

// per-warp handle on the TMA staging ring (empty when IPDDP_FW_TMA is off)
struct FwStage {
  double* buf;                 // FW_TMA_STAGES x STAGE doubles, 16-byte aligned
  unsigned long long* bars;    // FW_TMA_STAGES mbarriers
  unsigned seq;                // knots staged so far by this warp (stage = seq % STAGES, phase = (seq / STAGES) & 1)
};
template <class M>
IPDDP_D FwStage fw_stage_init(double* base, int lane) {
  FwStage st;
  st.buf = base;
  st.bars = reinterpret_cast<unsigned long long*>(base + FW_TMA_STAGES * FwLayout<M>::STAGE);
  st.seq = 0;
#if IPDDP_FW_TMA
  if (lane == 0) {
    for (int q = 0; q < FW_TMA_STAGES; ++q) mbar_init(&st.bars[q], 1);
    mbar_init_fence();
  }
  __syncwarp();
#endif
  return st;
}

// per-lane output descriptors of the rollout (invariant along a run of one stage type): which gain row, which nominal
// field.  Sized for the largest stage type of a chain; init<S> fills them for stage type S.
template <class M> struct FwDesc {
  static constexpr int NOUT_MAX = M::NU + M::NC + 2 * M::NU, NIT = (NOUT_MAX + 31) / 32;
  int g_off[NIT], g_ld[NIT], n_off[NIT], kind[NIT];   // kind 0 u, 1 phi, 2 zl, 3 zu, -1 none
  double blo[NIT], bup[NIT];
  int type;                                            // stage type the descriptors were filled for (-1: none yet)
  template <class S> IPDDP_D void init(const double* lo, const double* up, int lane, int type_) {
    typedef Rec<S> R;
    constexpr int NU = S::NU, K = S::NU + S::NC, NR = S::NX + 1, NOUT = K + 2 * NU;
    type = type_;
#pragma unroll
    for (int it = 0; it < NIT; ++it) {
      const int o = lane + 32 * it;
      blo[it] = 0.0; bup[it] = 0.0;
      if (o < NU) { kind[it] = 0; g_off[it] = o; g_ld[it] = K; n_off[it] = R::U + o; blo[it] = lo[o]; bup[it] = up[o]; }
      else if (o < K) { kind[it] = 1; g_off[it] = o; g_ld[it] = K; n_off[it] = R::PHI + (o - NU); }
      else if (o < K + NU) { kind[it] = 2; g_off[it] = K * NR + (o - K); g_ld[it] = 2 * NU; n_off[it] = R::ZL + (o - K); }
      else if (o < NOUT) { kind[it] = 3; g_off[it] = K * NR + (o - K); g_ld[it] = 2 * NU; n_off[it] = R::ZU + (o - K - NU); }
      else { kind[it] = -1; g_off[it] = 0; g_ld[it] = 0; n_off[it] = 0; }
    }
  }
};
// descriptors for the stage type of knot t (no-op while the type does not change)
template <class M>
IPDDP_D void fw_desc_for(const DevView& v, FwDesc<M>& d, int b, int t, int lane) {
  const int type = v.type_of(t);
  if (d.type == type) return;
  for_stage<M>(type, [&](auto tag) { d.template init<IPDDP_STAGE(tag)>(v.lower_of(b, type), v.upper_of(b, type), lane, type); });
}

// rollout! for step size gamma by one warp into the trial records at `trial` (record t at trial + t * TR).
// returns 0 pass, 1 non-finite control / state (DomainError analogue, src/forward_pass.jl:18-24),
//         2 fraction-to-boundary violation (src/forward_pass.jl:26-27)
template <class M>
IPDDP_D int fw_rollout(const DevView& v, FwDesc<M>& d, int b, int Nb, int nom, double* trial, double gamma,
                       double one_m_tau, const double* p, double* us, int lane, FwStage& stg) {
  constexpr int STRIDE = Rec<M>::STRIDE;
  constexpr int NXM = Dims<M>::NS;     // a chain's state size changes with the stage type
  constexpr int NIT = FwDesc<M>::NIT;
  int rc = 0;
  double x[NXM], dx[NXM];
  const double* r0 = v.rec(nom, b, 0);
#pragma unroll
  for (int i = 0; i < NXM; ++i) x[i] = (M::NSTAGE == 1 || i < v.snx[v.type_of(0)]) ? r0[i] : 0.0;
  // prefetch registers for knot t
  double ff[NIT], fb[NIT][M::NX], nv[NIT], nil[NIT], niu[NIT], xbar[NXM];
#if IPDDP_FW_TMA
  // Knot t's gains record and nominal record are copied into stage (seq0 + t) % STAGES by one bulk copy each, issued by
  // lane 0 up to STAGES - 1 knots before the warp reads them.  A stage is refilled one knot AFTER its values were loaded
  // into registers, i.e. after those registers have been consumed by the arithmetic of that knot: no read of the stage can
  // still be in flight.  Every issued copy is waited for before the function returns (early exits drain the ring), so the
  // barrier phases stay in step with stg.seq across calls.
  constexpr int SD = FwLayout<M>::STAGE;
  const unsigned seq0 = stg.seq;
  const int nrun = Nb - 1;                 // knots with gains
  int issued = 0, waited = 0;
  auto issue = [&](int t) {
    if (lane == 0) {
      const int sidx = (int)((seq0 + (unsigned)t) % FW_TMA_STAGES);
      double* dst = stg.buf + (size_t)sidx * SD;
      mbar_expect_tx(&stg.bars[sidx], (unsigned)(SD * sizeof(double)));
      bulk_g2s(dst, v.gains + ((size_t)b * (v.N - 1) + t) * v.G, (unsigned)(FwLayout<M>::GP * sizeof(double)), &stg.bars[sidx]);
      bulk_g2s(dst + FwLayout<M>::GP, v.rec(nom, b, t), (unsigned)(STRIDE * sizeof(double)), &stg.bars[sidx]);
    }
    issued = t + 1;
  };
  auto wait_knot = [&](int t) {
    const unsigned q = seq0 + (unsigned)t;
    mbar_wait(&stg.bars[q % FW_TMA_STAGES], (q / FW_TMA_STAGES) & 1u);
    waited = t + 1;
  };
  for (int t = 0; t < FW_TMA_STAGES && t < nrun; ++t) issue(t);
#endif
  // loads knot t's gains rows and nominal values into the prefetch registers (descriptors: the knot's stage type)
  auto prefetch = [&](int t) {
    fw_desc_for<M>(v, d, b, t, lane);
#if IPDDP_FW_TMA
    wait_knot(t);
    const double* g = stg.buf + (size_t)((seq0 + (unsigned)t) % FW_TMA_STAGES) * SD;
    const double* rn = g + FwLayout<M>::GP;
    if (t >= 1 && t - 1 + FW_TMA_STAGES < nrun) issue(t - 1 + FW_TMA_STAGES);   // knot t-1's stage is free (see above)
#else
    const double* g = v.gains + ((size_t)b * (v.N - 1) + t) * v.G;
    const double* rn = v.rec(nom, b, t);
#endif
    for_stage<M>(d.type, [&](auto tag) {
      typedef IPDDP_STAGE(tag) S;
      typedef Rec<S> R;
#pragma unroll
      for (int i = 0; i < S::NX; ++i) xbar[i] = rn[R::X + i];
#pragma unroll
      for (int it = 0; it < NIT; ++it) {
        if (d.kind[it] >= 0) {
          ff[it] = g[d.g_off[it]];
#pragma unroll
          for (int j = 0; j < S::NX; ++j) fb[it][j] = g[d.g_off[it] + (1 + j) * d.g_ld[it]];
          nv[it] = rn[d.n_off[it]];
          if (d.kind[it] == 0) { nil[it] = rn[R::IL + (d.n_off[it] - R::U)]; niu[it] = rn[R::IU + (d.n_off[it] - R::U)]; }
        }
      }
    });
  };
  if (Nb > 1) prefetch(0);
  for (int t = 0; t < Nb; ++t) {
    double* rcur = trial + (size_t)t * STRIDE;
    if (t == Nb - 1) {        // terminal knot: only the state
      if (lane < M::Terminal::NXT) {
        double xv = x[0];
#pragma unroll
        for (int i = 1; i < M::Terminal::NXT; ++i) xv = (lane == i) ? x[i] : xv;
        rcur[lane] = xv;
      }
      break;
    }
    bool viol = false, bad = false;
    const int type_t = d.type;            // the prefetch below may move the descriptors on to the next knot's stage type
    for_stage<M>(type_t, [&](auto tag) {
      typedef IPDDP_STAGE(tag) S;
      typedef Rec<S> R;
      constexpr int NX = S::NX;
      if (lane < NX) {
        double xv = x[0];
#pragma unroll
        for (int i = 1; i < NX; ++i) xv = (lane == i) ? x[i] : xv;
        rcur[R::X + lane] = xv;
      }
#pragma unroll
      for (int i = 0; i < NX; ++i) dx[i] = x[i] - xbar[i];
#pragma unroll
      for (int it = 0; it < NIT; ++it) {
        if (d.kind[it] >= 0) {
          double w = ff[it];
          w *= gamma;
          w += nv[it];
          w = dot4c<NX>(fb[it], 1, dx, 1) + w;
          rcur[d.n_off[it]] = w;
          if (d.kind[it] == 0) {
            const int i = d.n_off[it] - R::U;
            us[i] = w;
            const double il = w - d.blo[it], iu = d.bup[it] - w;
            rcur[R::IL + i] = il;
            rcur[R::IU + i] = iu;
            viol = viol || (nil[it] * one_m_tau > il) || (niu[it] * one_m_tau > iu);
            bad = bad || !finite(w);
          } else if (d.kind[it] >= 2) {
            viol = viol || (nv[it] * one_m_tau > w);
          }
        }
      }
    });
    __syncwarp();
    if (t + 1 < Nb - 1) prefetch(t + 1);
    else {
      const double* rn = v.rec(nom, b, t + 1);
#pragma unroll
      for (int i = 0; i < M::Terminal::NXT; ++i) xbar[i] = rn[i];
    }
    for_stage<M>(type_t, [&](auto tag) {
      typedef IPDDP_STAGE(tag) S;
      double xn[S::NXN];
      S::dyn(x, us, p, xn);
#pragma unroll
      for (int i = 0; i < S::NXN; ++i) { x[i] = xn[i]; bad = bad || !finite(xn[i]); }
    });
    const bool any_bad = __any_sync(IPDDP_FULL_MASK, bad);
    const bool any_viol = __any_sync(IPDDP_FULL_MASK, viol);
    __syncwarp();
    if (any_bad) { rc = 1; break; }
    if (any_viol) { rc = 2; break; }
  }
#if IPDDP_FW_TMA
  while (waited < issued) wait_knot(waited);      // drain the copies an early exit left in flight
  stg.seq = seq0 + (unsigned)issued;
  __syncwarp();
#endif
  return rc;
}

// per-instance invariants of one forward pass + the line-search state
template <class M> struct FwState {
  int b, Nb, nom, cur, fn;
  double mu, one_m_tau, theta_prev, L_prev, theta_min, dL;
  const double* p;
  int l, status, nroll;
  double step, J, L_next, theta_next;
  bool switching, armijo;
};

template <class M>
IPDDP_D void fw_setup(const DevView& v, FwState<M>& s, int b) {
  s.b = b;
  s.Nb = v.horizon[b];
  s.nom = v.nomsel[b];
  s.cur = 1 - s.nom;
  s.p = v.p + (size_t)b * (M::NP > 0 ? M::NP : 1);
  s.mu = v.sdv(SD_MU, b);
  const double tau = jmax(v.opt.tau_min, 1.0 - s.mu);
  s.one_m_tau = 1.0 - tau;
  s.theta_prev = v.sdv(SD_THETA_CURR, b);
  s.L_prev = v.sdv(SD_L_CURR, b);
  s.theta_min = v.sdv(SD_THETA_MIN, b);
  s.fn = v.siv(SI_FILTER_N, b);
  s.l = 0; s.status = 0; s.nroll = 0;
  s.step = 1.0;
  s.switching = false; s.armijo = false;
  s.L_next = 0.0; s.theta_next = 0.0;
  s.J = v.sdv(SD_OBJECTIVE, b);
  s.dL = 0.0;
}

// expected_change_lagrangian (src/forward_pass.jl:87-96): per-knot terms in parallel, ordered sum t descending
template <class M>
IPDDP_D double fw_expected_change(const DevView& v, const FwState<M>& s, double* p_l, double* p_th, int lane) {
  for (int t = lane; t < s.Nb - 1; t += 32) {
    const double* g = v.gains + ((size_t)s.b * (v.N - 1) + t) * v.G;
    const double* q = v.Qu + ((size_t)s.b * (v.N - 1) + t) * M::NU;
    const double* rn = v.rec(s.nom, s.b, t);
    for_stage<M>(v.type_of(t), [&](auto tag) {
      typedef IPDDP_STAGE(tag) S;
      typedef Rec<S> R;
      p_l[t] = dot4c<S::NU>(q, 1, g, 1);
      p_th[t] = dot4c<S::NC>(rn + R::C, 1, g + S::NU, 1);
    });
  }
  __syncwarp();
  double dL = 0.0;
  for (int t = s.Nb - 2; t >= 0; --t) { dL += p_l[t]; dL += p_th[t]; }
  __syncwarp();
  return dL;
}

// (4) for a trial point with merit (theta, L) at step size gamma = s.step: filter / switching / Armijo / sufficient
// decrease.  returns true if the step is accepted; otherwise s.status in {3,4,5} (the caller halves the step, l += 1)
template <class M>
IPDDP_D bool fw_accept(const DevView& v, FwState<M>& s, double Jn, double theta, double L) {
  const double gamma = s.step;
  s.J = Jn;
  bool blocked = false;
  for (int f = 0; f < s.fn; ++f) {
    const double ft = v.filter[(size_t)(0 * IPDDP_FILTER_CAPACITY + f) * v.B + s.b];
    const double fL = v.filter[(size_t)(1 * IPDDP_FILTER_CAPACITY + f) * v.B + s.b];
    if (theta >= ft && L >= fL) { blocked = true; break; }
  }
  s.status = blocked ? 3 : 0;
  if (s.status != 0) return false;
  s.switching = (s.dL < 0.0) && (dm::pow(-gamma * s.dL, v.opt.s_L) * dm::pow(gamma, 1.0 - v.opt.s_L) >
                                 v.opt.delta * dm::pow(s.theta_prev, v.opt.s_theta));
  s.armijo = L - s.L_prev - 10.0 * IPDDP_EPS * fabs(s.L_prev) <= v.opt.eta_L * gamma * s.dL;
  if (theta <= s.theta_min && s.switching) {
    s.status = s.armijo ? 0 : 4;
  } else {
    const bool suff = (theta <= (1.0 - v.opt.gamma_theta) * s.theta_prev) || (L <= s.L_prev - v.opt.gamma_L * s.theta_prev);
    s.status = suff ? 0 : 5;
  }
  if (s.status != 0) return false;
  s.L_next = L;
  s.theta_next = theta;
  return true;
}

// bookkeeping by one thread after the line search ended (accepted, or step < eps => status 7)
template <class M>
IPDDP_D void fw_finish(const DevView& v, const FwState<M>& s, int* list_next, int* counters) {
  const int b = s.b;
  int status = s.status;
  if (s.step < IPDDP_EPS) status = 7;
  v.siv(SI_L, b) = s.l;
  v.siv(SI_NROLL, b) += s.nroll;
  v.siv(SI_LASTROLL, b) = s.nroll;
  v.sdv(SD_STEP, b) = s.step;
  v.sdv(SD_OBJECTIVE, b) = s.J;
  v.siv(SI_SWITCHING, b) = s.switching;
  v.siv(SI_ARMIJO, b) = s.armijo;
  v.siv(SI_STATUS, b) = status;
  if (status != 0) { mark_done(v, b, counters); return; }
  v.nomsel[b] = s.cur;
  const int fn = s.fn;
  if (!s.armijo && !s.switching) {
    if (fn >= IPDDP_FILTER_CAPACITY) { v.siv(SI_STATUS, b) = 9; mark_done(v, b, counters); return; }
    v.filter[(size_t)(0 * IPDDP_FILTER_CAPACITY + fn) * v.B + b] = (1.0 - v.opt.gamma_theta) * s.theta_prev;
    v.filter[(size_t)(1 * IPDDP_FILTER_CAPACITY + fn) * v.B + b] = s.L_prev - v.opt.gamma_L * s.theta_prev;
    v.siv(SI_FILTER_N, b) = fn + 1;
  }
  v.sdv(SD_L_CURR, b) = s.L_next;
  v.sdv(SD_THETA_CURR, b) = s.theta_next;
  v.sdv(SD_L_NEXT, b) = s.L_next;
  v.sdv(SD_THETA_NEXT, b) = s.theta_next;
  const int k = v.siv(SI_K, b) + 1;
  v.siv(SI_K, b) = k;
  if (v.trace_cap > 0) {
    const int row = v.siv(SI_TRACE_N, b);
    if (row < v.trace_cap) {
      double* tr = v.trace + ((size_t)b * v.trace_cap + row) * IPDDP_TRACE_COLS;
      tr[0] = (double)k; tr[1] = (double)v.siv(SI_J, b); tr[2] = s.J; tr[3] = v.sdv(SD_PRIMAL_INF, b);
      tr[4] = v.sdv(SD_DUAL_INF, b); tr[5] = v.sdv(SD_CS_INF, b); tr[6] = s.mu; tr[7] = v.sdv(SD_REG_LAST, b);
      tr[8] = s.step; tr[9] = (double)s.l; tr[10] = s.theta_next; tr[11] = s.L_next;
      v.siv(SI_TRACE_N, b) = row + 1;
    }
  }
  if (k >= v.opt.max_iterations) { v.siv(SI_STATUS, b) = 8; mark_done(v, b, counters); return; }
  append_next(v, list_next, counters, b);
}

template <class M>
__global__ void __launch_bounds__(FW_WARPS * 32, IPDDP_FW_MINBLOCKS) k_forward(DevView v, const int* list_fwd, int* list_next, int* counters) {
  IPDDP_DYN_SMEM(double, sm_all);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int slot = blockIdx.x * FW_WARPS + warp;
  ListView lf = fwd_view(v, list_fwd, counters);
  if (fwd_heavy_split(v, lf)) lf.n[0] = 0;      // the heaviest bucket is k_forward_spec's (heavy_only launch)
  if (slot >= lf.total()) return;
  double* us = sm_all + (size_t)warp * MeritLayout<M>::per_warp_doubles(v.N);
  double* chunk = us + MeritLayout<M>::NUP;
  unsigned char* bidx = reinterpret_cast<unsigned char*>(chunk + 32);
  double* part = us + MeritLayout<M>::FIXED;   // [4][N]: l_t, theta_t, c'phi_t, spare
  double* p_l = part; double* p_th = part + v.N; double* p_d = part + 2 * v.N;

  FwState<M> s;
  fw_setup<M>(v, s, lf.at(slot));
  const BoundLists<M> bls = warp_bound_lists<M>(v, s.b, bidx, lane);
  FwDesc<M> d;
  d.type = -1;
  s.dL = fw_expected_change<M>(v, s, p_l, p_th, lane);

  FwStage stg = fw_stage_init<M>(sm_all + (size_t)FW_WARPS * MeritLayout<M>::per_warp_doubles(v.N) + (size_t)warp * FwLayout<M>::TMA_DOUBLES, lane);
  double* trial = v.rec(s.cur, s.b, 0);
  while (s.step >= IPDDP_EPS) {
    s.nroll++;
    const int rc = fw_rollout<M>(v, d, s.b, s.Nb, s.nom, trial, s.step, s.one_m_tau, s.p, us, lane, stg);
    if (rc == 1) { s.step *= 0.5; continue; }
    if (rc == 2) { s.status = 2; s.step *= 0.5; continue; }
    __syncwarp();
    double Jn, theta, L;
    warp_eval_metrics<M>(v, trial, s.Nb, s.mu, s.p, bls, chunk, p_l, p_th, p_d, lane, &Jn, &theta, &L);
    if (fw_accept<M>(v, s, Jn, theta, L)) break;
    s.step *= 0.5;
    s.l += 1;
  }
  if (lane == 0) fw_finish<M>(v, s, list_next, counters);
}

template <class M>
__global__ void __launch_bounds__(FWS_WARPS * 32) k_forward_spec(DevView v, const int* list_fwd, int* list_next,
                                                                int* counters, int heavy_only) {
  typedef Rec<M> R;
  IPDDP_DYN_SMEM(double, sm_all);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int slot = blockIdx.x;
  ListView lf = fwd_view(v, list_fwd, counters);
  if (heavy_only) {                               // bulk round: only the heaviest bucket, and only if k_forward leaves it to us
    if (!fwd_heavy_split(v, lf)) return;
    lf.n[1] = 0; lf.n[2] = 0; lf.n[3] = 0;
  }
  if (slot >= lf.total()) return;
  constexpr int NUP = MeritLayout<M>::NUP;
  double* us = sm_all + (size_t)warp * MeritLayout<M>::per_warp_doubles(v.N);
  double* chunk = us + NUP;
  unsigned char* bidx = reinterpret_cast<unsigned char*>(chunk + 32);
  double* part = us + MeritLayout<M>::FIXED;
  double* p_l = part; double* p_th = part + v.N; double* p_d = part + 2 * v.N;
  double* res = sm_all + (size_t)FWS_WARPS * MeritLayout<M>::per_warp_doubles(v.N);   // [FWS_WARPS][3]: theta, L, J
  double* next_step = res + 3 * FWS_WARPS;
  int* verdict = reinterpret_cast<int*>(next_step + 1);                               // [FWS_WARPS]
  int* ctl = verdict + FWS_WARPS;                                                     // [0] accepted candidate or -1, [1] go on
  double* tma_base = sm_all + FwLayout<M>::spec_doubles(v.N);

  FwState<M> s;
  fw_setup<M>(v, s, lf.at(slot));
  FwDesc<M> d;
  d.type = -1;
  const BoundLists<M> bls = warp_bound_lists<M>(v, s.b, bidx, lane);
  if (warp == 0) s.dL = fw_expected_change<M>(v, s, p_l, p_th, lane);   // only thread 0 judges
  double* trial = v.spec_traj + ((size_t)slot * FWS_WARPS + warp) * v.N * R::STRIDE;
  FwStage stg = fw_stage_init<M>(tma_base + (size_t)warp * FwLayout<M>::TMA_DOUBLES, lane);

  // thread 0 owns the line-search state; the others only follow s.step
  for (;;) {
    double mine = s.step;
    for (int q = 0; q < warp; ++q) mine *= 0.5;     // the w-th value of the sequential `step *= 0.5` chain
    int rc = 3;                                      // 3: the sequential loop has ended before this step size
    if (mine >= IPDDP_EPS) {
      rc = fw_rollout<M>(v, d, s.b, s.Nb, s.nom, trial, mine, s.one_m_tau, s.p, us, lane, stg);
      if (rc == 0) {
        __syncwarp();
        double Jn, theta, L;
        warp_eval_metrics<M>(v, trial, s.Nb, s.mu, s.p, bls, chunk, p_l, p_th, p_d, lane, &Jn, &theta, &L);
        if (lane == 0) { res[3 * warp + 0] = theta; res[3 * warp + 1] = L; res[3 * warp + 2] = Jn; }
      }
    }
    if (lane == 0) verdict[warp] = rc;
    __syncthreads();
    if (threadIdx.x == 0) {
      int accepted = -1, go_on = 1;
      for (int q = 0; q < FWS_WARPS; ++q) {
        const int r = verdict[q];
        if (r == 3) { go_on = 0; break; }              // s.step < eps: the sequential while condition ends the search
        s.nroll++;
        if (r == 1) { s.step *= 0.5; continue; }
        if (r == 2) { s.status = 2; s.step *= 0.5; continue; }
        if (fw_accept<M>(v, s, res[3 * q + 2], res[3 * q + 0], res[3 * q + 1])) { accepted = q; go_on = 0; break; }
        s.step *= 0.5;
        s.l += 1;
      }
      if (go_on && !(s.step >= IPDDP_EPS)) go_on = 0;
      ctl[0] = accepted; ctl[1] = go_on;
      next_step[0] = s.step;
    }
    __syncthreads();
    const int accepted = ctl[0], go_on = ctl[1];
    if (accepted >= 0) {   // the accepted candidate's records become the instance's trial set (then nomsel flips)
      const double* src = v.spec_traj + ((size_t)slot * FWS_WARPS + accepted) * v.N * R::STRIDE;
      double* dst = v.rec(s.cur, s.b, 0);
      const int n = (s.Nb - 1) * R::STRIDE + M::Terminal::NXT;   // the terminal record only carries x
      for (int e = threadIdx.x; e < n; e += FWS_WARPS * 32) dst[e] = src[e];
    }
    if (!go_on) break;
    s.step = next_step[0];   // all FWS_WARPS candidates were consumed: continue the chain from step / 2^FWS_WARPS
    __syncthreads();
  }
  __syncthreads();          // the copy of the accepted records precedes the nomsel flip
  if (threadIdx.x == 0) fw_finish<M>(v, s, list_next, counters);
}

}  // namespace ipk
