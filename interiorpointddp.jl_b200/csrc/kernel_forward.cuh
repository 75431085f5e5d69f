// k_forward: forward_pass! (reference src/forward_pass.jl:1-57), one warp per instance.
//
// Per trial step the warp runs
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
#pragma once
#include "kernels_common.cuh"

namespace ipk {

constexpr int FW_WARPS = 4;

template <class M> struct FwLayout {
  static constexpr int NUP = M::NU > 0 ? M::NU : 1;
  // per warp: u[NU] | chunk[32] | idx bytes (2*NU, padded to 8 doubles) | per-knot partials 4 x N (runtime)
  static constexpr int FIXED = NUP + 32 + ((2 * NUP + 7) / 8);
  static IPDDP_BOTH int per_warp_doubles(int N) { return FIXED + 4 * N; }
  static size_t bytes(int N) { return (size_t)FW_WARPS * per_warp_doubles(N) * sizeof(double); }
};

template <class M>
__global__ void __launch_bounds__(FW_WARPS * 32) k_forward(DevView v, const int* list_fwd, int* list_next, int* counters) {
  typedef Rec<M> R;
  constexpr int NX = M::NX, NU = M::NU, NC = M::NC, K = NU + NC, NR = NX + 1;
  constexpr int NOUT = K + 2 * NU, NIT = (NOUT + 31) / 32;
  IPDDP_DYN_SMEM(double, sm_all);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int slot = blockIdx.x * FW_WARPS + warp;
  if (slot >= counters[CNT_FWD]) return;
  const int b = list_fwd[slot];
  const int Nb = v.horizon[b];
  double* us = sm_all + (size_t)warp * FwLayout<M>::per_warp_doubles(v.N);
  double* chunk = us + FwLayout<M>::NUP;
  unsigned char* bidx = reinterpret_cast<unsigned char*>(chunk + 32);
  double* part = chunk + 32 + ((2 * FwLayout<M>::NUP + 7) / 8);   // [4][N]: l_t, theta_t, c'phi_t, scratch
  double* p_l = part; double* p_th = part + v.N; double* p_d = part + 2 * v.N; double* p_s = part + 3 * v.N;

  const int nom = v.nomsel[b], cur = 1 - nom;
  const double* p = v.p + (size_t)b * (M::NP > 0 ? M::NP : 1);
  const double* lo = v.lower + (size_t)b * NU;
  const double* up = v.upper + (size_t)b * NU;
  const double mu = v.sdv(SD_MU, b);
  const double tau = jmax(v.opt.tau_min, 1.0 - mu);
  const double one_m_tau = 1.0 - tau;
  const double theta_prev = v.sdv(SD_THETA_CURR, b), L_prev = v.sdv(SD_L_CURR, b);
  const double theta_min = v.sdv(SD_THETA_MIN, b);
  const int fn = v.siv(SI_FILTER_N, b);

  // finite-bound index list in the reference's accumulation order: lower indices then upper indices
  int nlo = 0, nbd = 0;
  if (lane == 0) {
    int q = 0;
    for (int i = 0; i < NU; ++i) if (!is_inf(lo[i])) bidx[q++] = (unsigned char)i;
    nlo = q;
    for (int i = 0; i < NU; ++i) if (!is_inf(up[i])) bidx[q++] = (unsigned char)i;
    nbd = q;
  }
  nlo = __shfl_sync(IPDDP_FULL_MASK, nlo, 0);
  nbd = __shfl_sync(IPDDP_FULL_MASK, nbd, 0);
  __syncwarp();

  // per-lane output descriptors (loop invariant): which gain row, which nominal field
  int g_off[NIT], g_ld[NIT], n_off[NIT], kind[NIT];   // kind 0 u, 1 phi, 2 zl, 3 zu, -1 none
  double blo[NIT], bup[NIT];
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

  // ---- expected_change_lagrangian (src/forward_pass.jl:87-96): per-knot terms in parallel, ordered sum t descending
  for (int t = lane; t < Nb - 1; t += 32) {
    const double* g = v.gains + ((size_t)b * (v.N - 1) + t) * v.G;
    const double* q = v.Qu + ((size_t)b * (v.N - 1) + t) * NU;
    const double* rn = v.rec(nom, b, t);
    p_l[t] = dot4c<NU>(q, 1, g, 1);
    p_th[t] = dot4c<NC>(rn + R::C, 1, g + NU, 1);
  }
  __syncwarp();
  double dL = 0.0;
  for (int t = Nb - 2; t >= 0; --t) { dL += p_l[t]; dL += p_th[t]; }
  __syncwarp();

  int l = 0, status = 0, nroll = 0;
  double step = 1.0;
  bool switching = false, armijo = false;
  double L_next = 0.0, theta_next = 0.0, J = v.sdv(SD_OBJECTIVE, b);

  while (step >= IPDDP_EPS) {
    const double gamma = step;
    nroll++;
    // ================= (1) rollout =================
    int rc = 0;
    {
      double x[NX], xn[NX], dx[NX];
      const double* r0 = v.rec(nom, b, 0);
#pragma unroll
      for (int i = 0; i < NX; ++i) x[i] = r0[R::X + i];
      // prefetch registers for knot t
      double ff[NIT], fb[NIT][NX], nv[NIT], nil[NIT], niu[NIT], xbar[NX];
      auto prefetch = [&](int t) {
        const double* g = v.gains + ((size_t)b * (v.N - 1) + t) * v.G;
        const double* rn = v.rec(nom, b, t);
#pragma unroll
        for (int i = 0; i < NX; ++i) xbar[i] = rn[R::X + i];
#pragma unroll
        for (int it = 0; it < NIT; ++it) {
          if (kind[it] >= 0) {
            ff[it] = g[g_off[it]];
#pragma unroll
            for (int j = 0; j < NX; ++j) fb[it][j] = g[g_off[it] + (1 + j) * g_ld[it]];
            nv[it] = rn[n_off[it]];
            if (kind[it] == 0) { nil[it] = rn[R::IL + (n_off[it] - R::U)]; niu[it] = rn[R::IU + (n_off[it] - R::U)]; }
          }
        }
      };
      if (Nb > 1) prefetch(0);
      for (int t = 0; t < Nb; ++t) {
        double* rcur = v.rec(cur, b, t);
        if (lane < NX) {
          double xv = x[0];
#pragma unroll
          for (int i = 1; i < NX; ++i) xv = (lane == i) ? x[i] : xv;
          rcur[R::X + lane] = xv;
        }
        if (t == Nb - 1) break;
#pragma unroll
        for (int i = 0; i < NX; ++i) dx[i] = x[i] - xbar[i];
        bool viol = false, bad = false;
#pragma unroll
        for (int it = 0; it < NIT; ++it) {
          if (kind[it] >= 0) {
            double w = ff[it];
            w *= gamma;
            w += nv[it];
            w = dot4c<NX>(fb[it], 1, dx, 1) + w;
            rcur[n_off[it]] = w;
            if (kind[it] == 0) {
              const int i = n_off[it] - R::U;
              us[i] = w;
              const double il = w - blo[it], iu = bup[it] - w;
              rcur[R::IL + i] = il;
              rcur[R::IU + i] = iu;
              viol = viol || (nil[it] * one_m_tau > il) || (niu[it] * one_m_tau > iu);
              bad = bad || !finite(w);
            } else if (kind[it] >= 2) {
              viol = viol || (nv[it] * one_m_tau > w);
            }
          }
        }
        __syncwarp();
        if (t + 1 < Nb - 1) prefetch(t + 1);
        else {
          const double* rn = v.rec(nom, b, t + 1);
#pragma unroll
          for (int i = 0; i < NX; ++i) xbar[i] = rn[R::X + i];
        }
        M::dyn(x, us, p, xn);
#pragma unroll
        for (int i = 0; i < NX; ++i) { x[i] = xn[i]; bad = bad || !finite(xn[i]); }
        const bool any_bad = __any_sync(IPDDP_FULL_MASK, bad);
        const bool any_viol = __any_sync(IPDDP_FULL_MASK, viol);
        __syncwarp();
        if (any_bad) { rc = 1; break; }     // DomainError analogue (src/forward_pass.jl:18-24)
        if (any_viol) { rc = 2; break; }    // fraction-to-boundary (src/forward_pass.jl:26-27)
      }
    }
    if (rc == 1) { step *= 0.5; continue; }
    if (rc == 2) { status = 2; step *= 0.5; continue; }
    __syncwarp();
    // ================= (2) merit terms, lane = knot =================
    for (int t = lane; t < Nb; t += 32) {
      double* r = v.rec(cur, b, t);
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
    // ================= (3) ordered sums =================
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
          const double* r = v.rec(cur, b, t);
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
    J = Jn;
    const double L = bl;
    // ================= (4) acceptance logic =================
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
  (void)p_s;
  if (lane != 0) return;
  // ---- bookkeeping by lane 0
  v.siv(SI_L, b) = l;
  v.siv(SI_NROLL, b) += nroll;
  v.sdv(SD_STEP, b) = step;
  v.sdv(SD_OBJECTIVE, b) = J;
  v.siv(SI_SWITCHING, b) = switching;
  v.siv(SI_ARMIJO, b) = armijo;
  v.siv(SI_STATUS, b) = status;
  if (status != 0) { v.siv(SI_DONE, b) = 1; return; }
  v.nomsel[b] = cur;
  if (!armijo && !switching) {
    if (fn >= IPDDP_FILTER_CAPACITY) { v.siv(SI_STATUS, b) = 9; v.siv(SI_DONE, b) = 1; return; }
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
  if (k >= v.opt.max_iterations) { v.siv(SI_STATUS, b) = 8; v.siv(SI_DONE, b) = 1; return; }
  list_next[atomicAdd(&counters[CNT_NEXT], 1)] = b;
}

}  // namespace ipk
