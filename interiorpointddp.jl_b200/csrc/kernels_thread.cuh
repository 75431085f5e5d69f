// The short phases of a round: initialisation / admission and the convergence check with the barrier update (one warp
// per instance), and the thread-per-(instance, knot) derivative kernel.  (The forward pass lives in kernel_forward.cuh.)
// Instance records are contiguous, the active lists are plain index indirection.
#pragma once
#include "kernels_common.cuh"

namespace ipk {

// ---------------------------------------------------------------------------------------------
// k_init: initialize_trajectory! (reference src/solver.jl:54-105) + the prologue of solve!
// (src/solve.jl:14-38): projection of the initial controls strictly inside their bounds, slack
// distances, open-loop rollout, dual reset (src/solve.jl:182-198), J, c, theta, L, theta_max/min, filter.
// warm != 0: keep the stored nominal primal trajectory (solve!(solver), src/solve.jl:6-17).
// ---------------------------------------------------------------------------------------------
// one instance (slot b) by one warp; x1p / ubarp: this instance's initial state and control guess; sm: the warp's
// MeritLayout scratch.  The rollout is sequential in time (lane 0 evaluates the dynamics, state ping-pong in shared
// memory), everything per knot -- projection of the controls, record writes, dual reset -- is spread over the lanes, the
// merit terms come from warp_eval_metrics.  Returns (all lanes) true if the instance takes part in the first round
// (max_iterations > 0).
template <class M>
IPDDP_D bool init_instance(const DevView& v, int warm, int b, const double* x1p, const double* ubarp, double* sm,
                           int lane) {
  const int Nb = v.horizon[b];
  const double* p = v.p + (size_t)b * (M::NP > 0 ? M::NP : 1);
  const double k1 = v.opt.kappa_1, k2 = v.opt.kappa_2;
  int set = 0;
  if (warm) set = v.nomsel[b];
  else if (lane == 0) v.nomsel[b] = 0;

  double* us = sm;
  double* chunk = us + MeritLayout<M>::NUP;
  unsigned char* bidx = reinterpret_cast<unsigned char*>(chunk + 32);
  double* part = us + MeritLayout<M>::FIXED;
  // a chain's state size changes with the stage type: the two state buffers are sized for the largest; they live in
  // the per-knot arrays, which warp_eval_metrics only uses after the rollout
  double* xs = part;
  double* xns = part + Dims<M>::NS;
  if (!warm) {
    for (int i = lane; i < M::NX; i += 32) xs[i] = (M::NSTAGE == 1 || i < v.snx[v.type_of(0)]) ? x1p[i] : 0.0;
    __syncwarp();
  }
  for (int t = 0; t < Nb; ++t) {
    double* r = v.rec(set, b, t);
    if (t == Nb - 1) {
      if (!warm)
        for (int i = lane; i < M::Terminal::NXT; i += 32) r[i] = xs[i];
      break;
    }
    const int type = v.type_of(t);
    for_stage<M>(type, [&](auto tag) {
      typedef IPDDP_STAGE(tag) S;
      typedef Rec<S> R;
      const double* lo = v.lower_of(b, type);
      const double* up = v.upper_of(b, type);
      if (!warm) {
        for (int i = lane; i < S::NX; i += 32) r[R::X + i] = xs[i];
        const double* u0p = ubarp + (size_t)t * M::NU;
        for (int i = lane; i < S::NU; i += 32) {
          const double u0 = u0p[i], l = lo[i], h = up[i];
          double ub;
          if (!is_inf(l) && is_inf(h)) {
            double tmp = jmax(l, 1.0);
            tmp *= k1;
            tmp += l;
            ub = jmax(u0, tmp);
          } else if (!is_inf(h) && is_inf(l)) {
            double tmp = jmax(h, 1.0);
            tmp *= -k1;
            tmp += h;
            ub = jmin(u0, tmp);
          } else if (!is_inf(h) && !is_inf(l)) {
            double t1 = l + jmin(k1 * jmax(1.0, fabs(l)), k2 * (h - l));
            double t2 = h - jmin(k1 * jmax(1.0, fabs(h)), k2 * (h - l));
            ub = jmin(jmax(u0, t1), t2);
          } else {
            ub = u0;
          }
          us[i] = ub;
          r[R::U + i] = ub;
          r[R::IL + i] = ub - l;
          r[R::IU + i] = h - ub;
        }
        __syncwarp();
        if (lane == 0) S::dyn(xs, us, p, xns);
        __syncwarp();
        double* sw = xs; xs = xns; xns = sw;
      }
      // reset_duals!
      for (int i = lane; i < S::NC; i += 32) r[R::PHI + i] = 0.0;
      for (int i = lane; i < S::NU; i += 32) {
        r[R::ZL + i] = is_inf(lo[i]) ? 0.0 : 1.0;
        r[R::ZU + i] = is_inf(up[i]) ? 0.0 : 1.0;
      }
    });
  }
  {
    double* lam = v.lam + (size_t)b * v.N * Dims<M>::NS;
    for (int e = lane; e < Nb * Dims<M>::NS; e += 32) lam[e] = 0.0;
  }
  __syncwarp();

  // reset!(data) + prologue
  const double mu = v.opt.mu_init;
  double J, theta, L;
  const BoundLists<M> bls = warp_bound_lists<M>(v, b, bidx, lane);
  warp_eval_metrics<M>(v, v.rec(set, b, 0), Nb, mu, p, bls, chunk, part, part + v.N, part + 2 * v.N, lane, &J, &theta, &L);
  if (lane == 0) {
    v.sdv(SD_MU, b) = mu;
    v.sdv(SD_REG_LAST, b) = 0.0;
    v.sdv(SD_OBJECTIVE, b) = J;
    v.sdv(SD_PRIMAL_INF, b) = 0.0;
    v.sdv(SD_DUAL_INF, b) = 0.0;
    v.sdv(SD_CS_INF, b) = 0.0;
    v.sdv(SD_L_CURR, b) = L;
    v.sdv(SD_THETA_CURR, b) = theta;
    v.sdv(SD_L_NEXT, b) = 0.0;
    v.sdv(SD_THETA_NEXT, b) = 0.0;
    v.sdv(SD_THETA_MAX, b) = 1e4 * jmax(1.0, theta);
    v.sdv(SD_THETA_MIN, b) = 1e-4 * jmax(1.0, theta);
    v.sdv(SD_STEP, b) = 0.0;
    v.sdv(SD_DUAL_NUM, b) = 0.0;
    for (int f = 0; f < SI_COUNT; ++f) v.siv(f, b) = 0;
    reset_filter(v, b);
    if (v.opt.max_iterations <= 0) v.siv(SI_STATUS, b) = 8;
  }
  return v.opt.max_iterations > 0;
}

constexpr int INIT_WARPS = 4;   // same per-warp scratch as k_check (Launch::smem_merit bounds both)

template <class M>
__global__ void __launch_bounds__(INIT_WARPS * 32) k_init(DevView v, int warm, int b0, int nb, int* list_next,
                                                          int* counters) {
  IPDDP_DYN_SMEM(double, sm_all);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j = blockIdx.x * INIT_WARPS + warp;
  if (j >= nb) return;
  const int b = b0 + j;
  double* sm = sm_all + (size_t)warp * MeritLayout<M>::per_warp_doubles(v.N);
  const bool runs = init_instance<M>(v, warm, b, v.x1 + (size_t)b * M::NX, v.ubar + (size_t)b * (v.N - 1) * M::NU, sm, lane);
  if (lane == 0) {
    if (runs) append_next(v, list_next, counters, b);
    else mark_done(v, b, counters);
  }
}

// Queue mode: admit n queued instances inst0 .. inst0+n-1 into the slots slots[0..n) (NULL: slots 0..n-1): copy the
// instance's parameters, bounds and horizon into the slot, initialise its trajectory from the queue's x1 / ubar
// (k_init's work) and append the slot to the running round's list at list[j] (the host passes the end of the lightest
// bucket: a fresh instance's first backward pass is one sweep).  One warp per admitted instance.
template <class M>
__global__ void __launch_bounds__(INIT_WARPS * 32) k_admit(DevView v, QueueView q, const int* slots, int n, int inst0,
                                                           int* list, int* counters) {
  IPDDP_DYN_SMEM(double, sm_all);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j = blockIdx.x * INIT_WARPS + warp;
  if (j >= n) return;
  const int b = slots ? slots[j] : j;
  const int i = inst0 + j;
  constexpr int NP1 = M::NP > 0 ? M::NP : 1;
  double* ps = const_cast<double*>(v.p) + (size_t)b * NP1;
  constexpr int NB = M::NSTAGE * M::NU;       // bounds per instance: [stage type][control]
  double* lo = const_cast<double*>(v.lower) + (size_t)b * NB;
  double* up = const_cast<double*>(v.upper) + (size_t)b * NB;
  for (int e = lane; e < M::NP; e += 32) ps[e] = q.p[(size_t)i * NP1 + e];
  for (int e = lane; e < NB; e += 32) { lo[e] = q.lower[(size_t)i * NB + e]; up[e] = q.upper[(size_t)i * NB + e]; }
  int hz = q.horizon ? q.horizon[i] : v.N;
  if (hz < 2 || hz > v.N) { if (lane == 0) atomicAdd(&counters[CNT_BAD], 1); hz = hz < 2 ? 2 : v.N; }
  if (lane == 0) {
    const_cast<int*>(v.horizon)[b] = hz;
    v.inst_of[b] = i;
  }
  __syncwarp();
  double* sm = sm_all + (size_t)warp * MeritLayout<M>::per_warp_doubles(v.N);
  const bool runs = init_instance<M>(v, 0, b, q.x1 + (size_t)i * M::NX, q.ubar + (size_t)i * (v.N - 1) * M::NU, sm, lane);
  if (lane == 0) {
    if (runs) list[j] = b;
    else mark_done(v, b, counters);
  }
}

// ---------------------------------------------------------------------------------------------
// k_derivs: evaluate_derivatives!(problem) (reference src/derivatives.jl:1-35) over the
// (active instance x knot) grid.  Consecutive threads take consecutive knots of one instance, so
// every compact-tile slot is written as a run of consecutive doubles.
// ---------------------------------------------------------------------------------------------
struct TileStore {
  double* base;
  int stride;
  IPDDP_D void operator()(int slot, double val) const { base[(size_t)slot * stride] = val; }
};

template <class M>
__global__ void k_derivs(DevView v, ListView list) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int i = (int)(idx / v.N), t = (int)(idx % v.N);
  if (i >= list.total()) return;
  const int b = list.at(i);
  const int Nb = v.horizon[b];
  if (t >= Nb) return;
  const double* p = v.p + (size_t)b * (M::NP > 0 ? M::NP : 1);
  const double* r = v.rec(v.nomsel[b], b, t);
  if (t < Nb - 1) {
    for_stage<M>(v.type_of(t), [&](auto tag) {
      typedef IPDDP_STAGE(tag) S;
      typedef Rec<S> R;
      double x[S::NX], u[S::NU > 0 ? S::NU : 1], phi[S::NC > 0 ? S::NC : 1];
#pragma unroll
      for (int q = 0; q < S::NX; ++q) x[q] = r[R::X + q];
#pragma unroll
      for (int q = 0; q < S::NU; ++q) u[q] = r[R::U + q];
#pragma unroll
      for (int q = 0; q < S::NC; ++q) phi[q] = r[R::PHI + q];
      TileStore st{v.tile + (size_t)b * M::D_NSLOT * v.N + t, v.N};
      S::derivs(x, u, phi, p, st);
    });
  } else {
    typedef typename M::Terminal T;
    double x[T::NXT];
#pragma unroll
    for (int q = 0; q < T::NXT; ++q) x[q] = r[q];
    TileStore st{v.tileN + (size_t)b * (M::DN_NSLOT > 0 ? M::DN_NSLOT : 1), 1};
    T::derivsN(x, p, st);
  }
  if (t == 0) v.siv(SI_NDERIV, b) += 1;
}

// ---------------------------------------------------------------------------------------------
// k_check: optimality errors (reference src/solve.jl:107-180), convergence test and barrier update
// (src/solve.jl:49-73), one warp per instance.  Appends the instance to the forward list, to the next-round list
// (barrier update: `continue` without forward pass, Q6) or marks it done.
// Per-knot terms are evaluated with lane = knot; the max-norms are order independent (NaN-propagating max), the
// sums are then accumulated sequentially in the reference's order (t descending) from shared memory.
// ---------------------------------------------------------------------------------------------
constexpr int CHK_WARPS = 4;

template <class M>
__global__ void __launch_bounds__(CHK_WARPS * 32) k_check(DevView v, ListView list, int* list_next, int* list_fwd,
                                                         int* counters) {
  IPDDP_DYN_SMEM(double, sm_all);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i = blockIdx.x * CHK_WARPS + warp;
  if (i >= list.total()) return;
  const int b = list.at(i);
  if (v.siv(SI_STATUS, b) != 0) {  // backward pass failed (status 1): break
    __syncwarp();
    if (lane == 0) mark_done(v, b, counters);
    return;
  }
  double* us = sm_all + (size_t)warp * MeritLayout<M>::per_warp_doubles(v.N);
  double* chunk = us + MeritLayout<M>::NUP;
  unsigned char* bidx = reinterpret_cast<unsigned char*>(chunk + 32);
  double* part = us + MeritLayout<M>::FIXED;
  double* p_a = part; double* p_b = part + v.N; double* p_c = part + 2 * v.N;
  const int Nb = v.horizon[b];
  const int set = v.nomsel[b];
  const double* p = v.p + (size_t)b * (M::NP > 0 ? M::NP : 1);
  const double mu = v.sdv(SD_MU, b);
  const BoundLists<M> bls = warp_bound_lists<M>(v, b, bidx, lane);

  double primal_inf = 0.0, cs0 = 0.0, csm = 0.0;
  // terminal stage contributes nothing (nu = nc = 0)
  for (int t = lane; t < Nb - 1; t += 32) {
    const double* r = v.rec(set, b, t);
    const int type = v.type_of(t);
    for_stage<M>(type, [&](auto tag) {
      typedef IPDDP_STAGE(tag) S;
      typedef Rec<S> R;
      double m = 0.0;
#pragma unroll
      for (int q = 0; q < S::NC; ++q) m = jmax(m, fabs(r[R::C + q]));
      primal_inf = jmax(primal_inf, m);
      double szl = 0.0, szu = 0.0, sphi = 0.0;
#pragma unroll
      for (int q = 0; q < S::NU; ++q) szl += r[R::ZL + q];
#pragma unroll
      for (int q = 0; q < S::NU; ++q) szu += r[R::ZU + q];
#pragma unroll
      for (int q = 0; q < S::NC; ++q) sphi += fabs(r[R::PHI + q]);
      p_a[t] = szl; p_b[t] = szu; p_c[t] = sphi;
      if (bls.nbd[type] > 0) {
        double a0 = 0.0, am = 0.0, b0 = 0.0, bm = 0.0;
#pragma unroll
        for (int q = 0; q < S::NU; ++q) {
          double w = r[R::IL + q];
          w *= r[R::ZL + q];
          double w0 = w - 0.0, wm = w - mu;
          if (w0 != w0) w0 = 0.0;   // replace!(NaN => 0) after subtracting mu (Q3)
          if (wm != wm) wm = 0.0;
          a0 = jmax(a0, fabs(w0));
          am = jmax(am, fabs(wm));
        }
#pragma unroll
        for (int q = 0; q < S::NU; ++q) {
          double w = r[R::IU + q];
          w *= r[R::ZU + q];
          double w0 = w - 0.0, wm = w - mu;
          if (w0 != w0) w0 = 0.0;
          if (wm != wm) wm = 0.0;
          b0 = jmax(b0, fabs(w0));
          bm = jmax(bm, fabs(wm));
        }
        cs0 = jmax(cs0, a0); cs0 = jmax(cs0, b0);
        csm = jmax(csm, am); csm = jmax(csm, bm);
      }
    });
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    primal_inf = jmax(primal_inf, __shfl_xor_sync(IPDDP_FULL_MASK, primal_inf, o));
    cs0 = jmax(cs0, __shfl_xor_sync(IPDDP_FULL_MASK, cs0, o));
    csm = jmax(csm, __shfl_xor_sync(IPDDP_FULL_MASK, csm, o));
  }
  __syncwarp();
  // sums in the reference's order (t descending); the complementarity error only counts stages that have bounds
  // (src/solve.jl:160-166), the dual error every stage
  double z_norm = 0.0, phi_norm = 0.0, z_norm_cs = 0.0;
  int num_bounds = 0, num_con = 0;
  for (int t = Nb - 2; t >= 0; --t) {
    z_norm += p_a[t];
    z_norm += p_b[t];
    phi_norm += p_c[t];
    if constexpr (M::NSTAGE > 1) {
      const int type = v.type_of(t);
      if (bls.nbd[type] > 0) { z_norm_cs += p_a[t]; z_norm_cs += p_b[t]; }
      num_bounds += bls.nbd[type];
      num_con += v.snc[type];
    }
  }
  __syncwarp();
  if constexpr (M::NSTAGE == 1) {
    num_bounds = bls.nbd[0] * (Nb - 1);
    num_con = M::NC * (Nb - 1);
    z_norm_cs = bls.nbd[0] > 0 ? z_norm : 0.0;      // same additions in the same order
  }
  const double num_ineq = (double)num_bounds;        // a sum of small integers: exact
  const double num_constr = (double)num_con;
  const double s_max = v.opt.s_max;
  const double sd_ = jmax(s_max, (phi_norm + z_norm) / jmax(num_ineq + num_constr, 1.0)) / s_max;
  const double sc_ = jmax(s_max, z_norm_cs / jmax(num_ineq, 1.0)) / s_max;
  const double dual_inf = v.sdv(SD_DUAL_NUM, b) / sd_;
  const double cs_inf = cs0 / sc_;
  const double cs_mu = csm / sc_;
  const double err_mu = jmax(jmax(dual_inf, cs_mu), primal_inf);
  const double err_0 = jmax(jmax(dual_inf, cs_inf), primal_inf);
  const double tol = v.opt.optimality_tolerance;
  if (lane == 0) {
    v.sdv(SD_DUAL_INF, b) = dual_inf;
    v.sdv(SD_PRIMAL_INF, b) = primal_inf;
    v.sdv(SD_CS_INF, b) = cs_inf;
  }
  if (err_0 < tol) {  // converged
    if (lane == 0) mark_done(v, b, counters);
    return;
  }
  if (err_mu <= v.opt.kappa_eps * mu && num_bounds > 0 && mu > tol / 10.0) {
    const double mu_new = jmax(tol / 10.0, jmin(v.opt.kappa_mu * mu, dm::pow(mu, v.opt.theta_mu)));
    double J, theta, L;
    warp_eval_metrics<M>(v, v.rec(set, b, 0), Nb, mu_new, p, bls, chunk, p_a, p_b, p_c, lane, &J, &theta, &L);
    if (lane == 0) {
      v.sdv(SD_MU, b) = mu_new;
      reset_filter(v, b);
      v.sdv(SD_OBJECTIVE, b) = J;
      v.sdv(SD_L_CURR, b) = L;
      v.sdv(SD_THETA_CURR, b) = theta;
      v.siv(SI_J, b) += 1;
      append_next(v, list_next, counters, b);
    }
    return;
  }
  if (lane == 0) append_fwd(v, list_fwd, counters, b);
}

}  // namespace ipk
