// Thread-per-instance kernels: initialisation, convergence check / barrier update, forward pass
// (rollout + filter line search), and the thread-per-(instance, knot) derivative kernel.
//
// These phases are short, sequential-in-time per instance and dominated by streaming each instance's
// own contiguous records, so one thread walks one instance; instance records are contiguous, the
// active lists are plain index indirection.
#pragma once
#include "kernels_common.cuh"

namespace ipk {

// ---------------------------------------------------------------------------------------------
// k_init: initialize_trajectory! (reference src/solver.jl:54-105) + the prologue of solve!
// (src/solve.jl:14-38): projection of the initial controls strictly inside their bounds, slack
// distances, open-loop rollout, dual reset (src/solve.jl:182-198), J, c, theta, L, theta_max/min, filter.
// warm != 0: keep the stored nominal primal trajectory (solve!(solver), src/solve.jl:6-17).
// ---------------------------------------------------------------------------------------------
template <class M>
__global__ void k_init(DevView v, int warm, int* list_next, int* counters) {
  typedef Rec<M> R;
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= v.B) return;
  const int Nb = v.horizon[b];
  const double* p = v.p + (size_t)b * (M::NP > 0 ? M::NP : 1);
  const double* lo = v.lower + (size_t)b * M::NU;
  const double* up = v.upper + (size_t)b * M::NU;
  const double k1 = v.opt.kappa_1, k2 = v.opt.kappa_2;
  int set = 0;
  if (warm) set = v.nomsel[b]; else v.nomsel[b] = 0;

  double x[M::NX], xn[M::NXN], u[M::NU > 0 ? M::NU : 1];
  if (!warm) {
#pragma unroll
    for (int i = 0; i < M::NX; ++i) x[i] = v.x1[(size_t)b * M::NX + i];
  }
  for (int t = 0; t < Nb; ++t) {
    double* r = v.rec(set, b, t);
    if (!warm) {
#pragma unroll
      for (int i = 0; i < M::NX; ++i) r[R::X + i] = x[i];
    }
    if (t < Nb - 1) {
      if (!warm) {
        const double* u0p = v.ubar + ((size_t)b * (v.N - 1) + t) * M::NU;
#pragma unroll
        for (int i = 0; i < M::NU; ++i) {
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
          u[i] = ub;
          r[R::U + i] = ub;
          r[R::IL + i] = ub - l;
          r[R::IU + i] = h - ub;
        }
        M::dyn(x, u, p, xn);
#pragma unroll
        for (int i = 0; i < M::NX; ++i) x[i] = xn[i];
      }
      // reset_duals!
#pragma unroll
      for (int i = 0; i < M::NC; ++i) r[R::PHI + i] = 0.0;
#pragma unroll
      for (int i = 0; i < M::NU; ++i) {
        r[R::ZL + i] = is_inf(lo[i]) ? 0.0 : 1.0;
        r[R::ZU + i] = is_inf(up[i]) ? 0.0 : 1.0;
      }
    }
  }
  for (int t = 0; t < Nb; ++t)
    for (int i = 0; i < M::NX; ++i) v.lam[((size_t)b * v.N + t) * M::NX + i] = 0.0;

  // reset!(data) + prologue
  const double mu = v.opt.mu_init;
  double J, theta, L;
  eval_metrics<M>(v, set, b, Nb, mu, &J, &theta, &L);
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
  if (v.opt.max_iterations > 0) {
    list_next[atomicAdd(&counters[CNT_NEXT], 1)] = b;
  } else {
    v.siv(SI_STATUS, b) = 8;
    v.siv(SI_DONE, b) = 1;
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
__global__ void k_derivs(DevView v, const int* list, int n_list) {
  typedef Rec<M> R;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int i = (int)(idx / v.N), t = (int)(idx % v.N);
  if (i >= n_list) return;
  const int b = list[i];
  const int Nb = v.horizon[b];
  if (t >= Nb) return;
  const double* p = v.p + (size_t)b * (M::NP > 0 ? M::NP : 1);
  const double* r = v.rec(v.nomsel[b], b, t);
  double x[M::NX];
#pragma unroll
  for (int q = 0; q < M::NX; ++q) x[q] = r[R::X + q];
  if (t < Nb - 1) {
    double u[M::NU > 0 ? M::NU : 1], phi[M::NC > 0 ? M::NC : 1];
#pragma unroll
    for (int q = 0; q < M::NU; ++q) u[q] = r[R::U + q];
#pragma unroll
    for (int q = 0; q < M::NC; ++q) phi[q] = r[R::PHI + q];
    TileStore st{v.tile + (size_t)b * M::D_NSLOT * v.N + t, v.N};
    M::derivs(x, u, phi, p, st);
  } else {
    TileStore st{v.tileN + (size_t)b * (M::DN_NSLOT > 0 ? M::DN_NSLOT : 1), 1};
    M::derivsN(x, p, st);
  }
  if (t == 0) v.siv(SI_NDERIV, b) += 1;
}

// ---------------------------------------------------------------------------------------------
// k_check: optimality errors (reference src/solve.jl:107-180), convergence test and barrier update
// (src/solve.jl:49-73).  Appends the instance to the forward list, to the next-round list (barrier
// update: `continue` without forward pass, Q6) or marks it done.
// ---------------------------------------------------------------------------------------------
template <class M>
__global__ void k_check(DevView v, const int* list, int n_list, int* list_next, int* list_fwd, int* counters) {
  typedef Rec<M> R;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_list) return;
  const int b = list[i];
  if (v.siv(SI_STATUS, b) != 0) {  // backward pass failed (status 1): break
    v.siv(SI_DONE, b) = 1;
    return;
  }
  const int Nb = v.horizon[b];
  const int set = v.nomsel[b];
  const double* lo = v.lower + (size_t)b * M::NU;
  const double* up = v.upper + (size_t)b * M::NU;
  const double mu = v.sdv(SD_MU, b);
  int nb_stage = 0;
  for (int q = 0; q < M::NU; ++q) nb_stage += (!is_inf(lo[q])) + (!is_inf(up[q]));

  double num_ineq = 0.0, z_norm = 0.0, phi_norm = 0.0, primal_inf = 0.0;
  double cs0 = 0.0, csm = 0.0, z_norm_cs = 0.0;
  // terminal stage contributes nothing (nu = nc = 0)
  for (int t = Nb - 2; t >= 0; --t) {
    const double* r = v.rec(set, b, t);
    double m = 0.0;
#pragma unroll
    for (int q = 0; q < M::NC; ++q) m = jmax(m, fabs(r[R::C + q]));
    primal_inf = jmax(primal_inf, m);
    double szl = 0.0, szu = 0.0, sphi = 0.0;
#pragma unroll
    for (int q = 0; q < M::NU; ++q) szl += r[R::ZL + q];
#pragma unroll
    for (int q = 0; q < M::NU; ++q) szu += r[R::ZU + q];
#pragma unroll
    for (int q = 0; q < M::NC; ++q) sphi += fabs(r[R::PHI + q]);
    z_norm += szl;
    z_norm += szu;
    phi_norm += sphi;
    num_ineq += (double)nb_stage;
    if (nb_stage > 0) {
      double a0 = 0.0, am = 0.0, b0 = 0.0, bm = 0.0;
#pragma unroll
      for (int q = 0; q < M::NU; ++q) {
        double w = r[R::IL + q];
        w *= r[R::ZL + q];
        double w0 = w - 0.0, wm = w - mu;
        if (w0 != w0) w0 = 0.0;   // replace!(NaN => 0) after subtracting mu (Q3)
        if (wm != wm) wm = 0.0;
        a0 = jmax(a0, fabs(w0));
        am = jmax(am, fabs(wm));
      }
#pragma unroll
      for (int q = 0; q < M::NU; ++q) {
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
      z_norm_cs += szl;
      z_norm_cs += szu;
    }
  }
  const double s_max = v.opt.s_max;
  const double num_constr = (double)(M::NC * (Nb - 1));
  const double sd_ = jmax(s_max, (phi_norm + z_norm) / jmax(num_ineq + num_constr, 1.0)) / s_max;
  const double sc_ = jmax(s_max, z_norm_cs / jmax(num_ineq, 1.0)) / s_max;
  const double dual_inf = v.sdv(SD_DUAL_NUM, b) / sd_;
  const double cs_inf = cs0 / sc_;
  const double cs_mu = csm / sc_;
  v.sdv(SD_DUAL_INF, b) = dual_inf;
  v.sdv(SD_PRIMAL_INF, b) = primal_inf;
  v.sdv(SD_CS_INF, b) = cs_inf;
  const double err_mu = jmax(jmax(dual_inf, cs_mu), primal_inf);
  const double err_0 = jmax(jmax(dual_inf, cs_inf), primal_inf);
  const double tol = v.opt.optimality_tolerance;
  if (err_0 < tol) {  // converged
    v.siv(SI_DONE, b) = 1;
    return;
  }
  const int num_bounds = nb_stage * (Nb - 1);
  if (err_mu <= v.opt.kappa_eps * mu && num_bounds > 0 && mu > tol / 10.0) {
    const double mu_new = jmax(tol / 10.0, jmin(v.opt.kappa_mu * mu, dm::pow(mu, v.opt.theta_mu)));
    v.sdv(SD_MU, b) = mu_new;
    reset_filter(v, b);
    double J, theta, L;
    eval_metrics<M>(v, set, b, Nb, mu_new, &J, &theta, &L);
    v.sdv(SD_OBJECTIVE, b) = J;
    v.sdv(SD_L_CURR, b) = L;
    v.sdv(SD_THETA_CURR, b) = theta;
    v.siv(SI_J, b) += 1;
    list_next[atomicAdd(&counters[CNT_NEXT], 1)] = b;
    return;
  }
  list_fwd[atomicAdd(&counters[CNT_FWD], 1)] = b;
}

// ---------------------------------------------------------------------------------------------
// k_forward: forward_pass! (reference src/forward_pass.jl:1-57) = backtracking line search over
// rollout! (:98-153) with the fraction-to-boundary test (:59-85, fused into the rollout with early
// exit: same accept/reject decision), filter test, switching / Armijo / sufficient-decrease tests;
// then update_nominal_trajectory! (src/data/methods.jl:78-91, here: flip nomsel), filter augmentation
// (src/solve.jl:81,95-99, Q5) and the iteration bookkeeping of src/solve.jl:82-85.
// ---------------------------------------------------------------------------------------------
template <class M>
IPDDP_D int rollout(const DevView& v, int b, int Nb, int nom, int cur, double step, double one_m_tau) {
  typedef Rec<M> R;
  constexpr int K = M::NU + M::NC, NR = M::NX + 1;
  const double* p = v.p + (size_t)b * (M::NP > 0 ? M::NP : 1);
  const double* lo = v.lower + (size_t)b * M::NU;
  const double* up = v.upper + (size_t)b * M::NU;
  double x[M::NX], xn[M::NXN], dx[M::NX], u[M::NU > 0 ? M::NU : 1];
  {
    const double* r0 = v.rec(nom, b, 0);
#pragma unroll
    for (int i = 0; i < M::NX; ++i) x[i] = r0[R::X + i];
  }
  for (int t = 0; t < Nb; ++t) {
    const double* rn = v.rec(nom, b, t);
    double* rc = v.rec(cur, b, t);
#pragma unroll
    for (int i = 0; i < M::NX; ++i) { dx[i] = x[i] - rn[R::X + i]; rc[R::X + i] = x[i]; }
    if (t == Nb - 1) break;
    const double* g = v.gains + ((size_t)b * (v.N - 1) + t) * v.G;
    const double* gi = g + K * NR;
    bool viol = false, bad = false;
#pragma unroll
    for (int i = 0; i < M::NU; ++i) {
      double w = g[i];
      w *= step;
      w += rn[R::U + i];
      w = dot4c<M::NX>(g + i + K, K, dx, 1) + w;
      u[i] = w;
      rc[R::U + i] = w;
      const double il = w - lo[i], iu = up[i] - w;
      rc[R::IL + i] = il;
      rc[R::IU + i] = iu;
      viol = viol || (rn[R::IL + i] * one_m_tau > il) || (rn[R::IU + i] * one_m_tau > iu);
      bad = bad || !finite(w);
    }
#pragma unroll
    for (int q = 0; q < M::NC; ++q) {
      double w = g[M::NU + q];
      w *= step;
      w += rn[R::PHI + q];
      rc[R::PHI + q] = dot4c<M::NX>(g + M::NU + q + K, K, dx, 1) + w;
    }
#pragma unroll
    for (int i = 0; i < M::NU; ++i) {
      double w = gi[i];
      w *= step;
      w += rn[R::ZL + i];
      w = dot4c<M::NX>(gi + i + 2 * M::NU, 2 * M::NU, dx, 1) + w;
      rc[R::ZL + i] = w;
      viol = viol || (rn[R::ZL + i] * one_m_tau > w);
    }
#pragma unroll
    for (int i = 0; i < M::NU; ++i) {
      double w = gi[M::NU + i];
      w *= step;
      w += rn[R::ZU + i];
      w = dot4c<M::NX>(gi + M::NU + i + 2 * M::NU, 2 * M::NU, dx, 1) + w;
      rc[R::ZU + i] = w;
      viol = viol || (rn[R::ZU + i] * one_m_tau > w);
    }
    M::dyn(x, u, p, xn);
#pragma unroll
    for (int i = 0; i < M::NX; ++i) { x[i] = xn[i]; bad = bad || !finite(xn[i]); }
    if (bad) return 1;    // DomainError analogue: reject, halve (src/forward_pass.jl:18-24)
    if (viol) return 2;   // fraction-to-boundary violated somewhere: reject, halve
  }
  return 0;
}

template <class M>
__global__ void k_forward(DevView v, const int* list_fwd, int* list_next, int* counters) {
  constexpr int K = M::NU + M::NC;
  typedef Rec<M> R;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= counters[CNT_FWD]) return;
  const int b = list_fwd[i];
  const int Nb = v.horizon[b];
  const int nom = v.nomsel[b], cur = 1 - nom;
  const double mu = v.sdv(SD_MU, b);
  const double tau = jmax(v.opt.tau_min, 1.0 - mu);
  const double one_m_tau = 1.0 - tau;
  const double theta_prev = v.sdv(SD_THETA_CURR, b), L_prev = v.sdv(SD_L_CURR, b);
  const double theta_min = v.sdv(SD_THETA_MIN, b);
  int l = 0, status = 0, nroll = 0;
  double step = 1.0;
  bool switching = false, armijo = false;
  double L_next = 0.0, theta_next = 0.0, J = v.sdv(SD_OBJECTIVE, b);

  // expected_change_lagrangian (src/forward_pass.jl:87-96), t descending
  double dL = 0.0;
  for (int t = Nb - 2; t >= 0; --t) {
    const double* g = v.gains + ((size_t)b * (v.N - 1) + t) * v.G;
    const double* q = v.Qu + ((size_t)b * (v.N - 1) + t) * M::NU;
    const double* rn = v.rec(nom, b, t);
    dL += dot4c<M::NU>(q, 1, g, 1);
    dL += dot4c<M::NC>(rn + R::C, 1, g + M::NU, 1);
  }
  (void)K;
  const int fn = v.siv(SI_FILTER_N, b);
  while (step >= IPDDP_EPS) {
    const double gamma = step;
    nroll++;
    const int rc = rollout<M>(v, b, Nb, nom, cur, gamma, one_m_tau);
    if (rc == 1) { step *= 0.5; continue; }
    if (rc == 2) { status = 2; step *= 0.5; continue; }
    double theta, L;
    eval_metrics<M>(v, cur, b, Nb, mu, &J, &theta, &L);
    bool blocked = false;
    for (int f = 0; f < fn; ++f) {
      const double ft = v.filter[(size_t)(0 * IPDDP_FILTER_CAPACITY + f) * v.B + b];
      const double fL = v.filter[(size_t)(1 * IPDDP_FILTER_CAPACITY + f) * v.B + b];
      if (theta >= ft && L >= fL) { blocked = true; break; }
    }
    status = blocked ? 3 : 0;
    if (status != 0) { step *= 0.5; l += 1; continue; }
    switching = (dL < 0.0) && (dm::pow(-gamma * dL, v.opt.s_L) * dm::pow(gamma, 1.0 - v.opt.s_L) >
                               v.opt.delta * dm::pow(theta_prev, v.opt.s_theta));
    armijo = L - L_prev - 10.0 * IPDDP_EPS * fabs(L_prev) <= v.opt.eta_L * gamma * dL;
    if (theta <= theta_min && switching) {
      status = armijo ? 0 : 4;
    } else {
      const bool suff = (theta <= (1.0 - v.opt.gamma_theta) * theta_prev) || (L <= L_prev - v.opt.gamma_L * theta_prev);
      status = suff ? 0 : 5;
    }
    if (status != 0) { step *= 0.5; l += 1; continue; }
    L_next = L;
    theta_next = theta;
    break;
  }
  if (step < IPDDP_EPS) status = 7;
  v.siv(SI_L, b) = l;
  v.siv(SI_NROLL, b) += nroll;
  v.sdv(SD_STEP, b) = step;
  v.sdv(SD_OBJECTIVE, b) = J;
  v.siv(SI_SWITCHING, b) = switching;
  v.siv(SI_ARMIJO, b) = armijo;
  v.siv(SI_STATUS, b) = status;
  if (status != 0) {  // line search failed: break
    v.siv(SI_DONE, b) = 1;
    return;
  }
  // accept: update_nominal_trajectory! (pointer flip), filter, bookkeeping
  v.nomsel[b] = cur;
  if (!armijo && !switching) {
    if (fn >= IPDDP_FILTER_CAPACITY) {
      v.siv(SI_STATUS, b) = 9;
      v.siv(SI_DONE, b) = 1;
      return;
    }
    v.filter[(size_t)(0 * IPDDP_FILTER_CAPACITY + fn) * v.B + b] = (1.0 - v.opt.gamma_theta) * theta_prev;
    v.filter[(size_t)(1 * IPDDP_FILTER_CAPACITY + fn) * v.B + b] = L_prev - v.opt.gamma_L * theta_prev;
    v.siv(SI_FILTER_N, b) = fn + 1;
  }
  v.sdv(SD_L_CURR, b) = L_next;
  v.sdv(SD_THETA_CURR, b) = theta_next;
  v.sdv(SD_L_NEXT, b) = L_next;
  v.sdv(SD_THETA_NEXT, b) = theta_next;
  const int k = v.siv(SI_K, b) + 1;
  v.siv(SI_K, b) = k;
  if (v.trace_cap > 0) {
    const int row = v.siv(SI_TRACE_N, b);
    if (row < v.trace_cap) {
      double* tr = v.trace + ((size_t)b * v.trace_cap + row) * IPDDP_TRACE_COLS;
      tr[0] = (double)k; tr[1] = (double)v.siv(SI_J, b); tr[2] = J; tr[3] = v.sdv(SD_PRIMAL_INF, b);
      tr[4] = v.sdv(SD_DUAL_INF, b); tr[5] = v.sdv(SD_CS_INF, b); tr[6] = mu; tr[7] = v.sdv(SD_REG_LAST, b);
      tr[8] = step; tr[9] = (double)l; tr[10] = theta_next; tr[11] = L_next;
      v.siv(SI_TRACE_N, b) = row + 1;
    }
  }
  if (k >= v.opt.max_iterations) {  // src/solve.jl:90
    v.siv(SI_STATUS, b) = 8;
    v.siv(SI_DONE, b) = 1;
    return;
  }
  list_next[atomicAdd(&counters[CNT_NEXT], 1)] = b;
}

}  // namespace ipk
