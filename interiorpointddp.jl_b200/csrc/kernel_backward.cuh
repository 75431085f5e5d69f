// k_backward / k_backward_spec: backward_pass! + inertia_correction! (reference src/backward_pass.jl:1-195,
// src/inertia_correction.jl:257-276).  bw_sweep is one sweep by one warp; k_backward runs the regularisation-restart
// loop with one warp per active instance (bulk rounds), k_backward_spec tries four values of the regularisation schedule
// at once with one CTA per instance (rounds with few active instances), see the comment above that kernel.
//
// The sweep is sequential in time (t = N-1 .. 0) with a restart-from-the-end loop on inertia failure;
// per knot the warp assembles the (nu+nc)^2 KKT matrix in shared memory from the compact derivative
// tile (scatter tables generated with the model), factorises it (ldlt_warp.cuh), solves for the
// nx+1 right-hand sides, forms the gains and updates Vx, Vxx, lambda.  Vx/Vxx/lambda are a
// shared-memory carry between knots and never touch HBM; only gains, Qu, lambda and the per-instance
// scalars are written back.  The numerator of the dual infeasibility (src/solve.jl:118-137) is
// accumulated on the fly because it needs lu, cu, fu (tile) and lambda (sweep) which are in hand here.
//
// Bit-parity notes: every output element is computed by exactly one lane with the operation order the
// CPU oracle uses (dot4 contractions; lhs/rhs terms added in the reference's order:
// lxx -> +fx'Vxx fx -> +vfxx -> +vcxx, ...).  Structural zeros are skipped, which is exact for finite data.
#pragma once
#include "ldlt_warp.cuh"

#ifndef IPDDP_BW_MINBLOCKS
#define IPDDP_BW_MINBLOCKS 20     // resident warps per SM the register allocation of k_backward is held to (96 registers)
#endif

namespace ipk {

// e / D for a compile-time divisor and the small non-negative indices of the gains blocks (e * D < 2^32): one multiply-high
// with ceil(2^32 / D) instead of the ~20 instructions of the signed division.  D == 1 has no 32-bit multiplier
// (ceil(2^32 / 1) wraps to 0): models with a single control (or a 1 x 1 KKT matrix) take the identity.
template <int D> IPDDP_D int div_small(int e) {
  if constexpr (D == 1) return e;
  else return (int)__umulhi((unsigned)e, 0xffffffffu / (unsigned)D + 1u);
}

template <class M> struct BwLayout {
  // Sizes come from M: for a stage chain M carries the maxima over its stage types, so one layout serves every knot; the
  // arithmetic of a knot indexes the buffers with its own stage type's dimensions.  NS = largest state size that can sit in
  // a state-sized buffer (stage states, next states, the terminal state).
  static constexpr int NX = M::NX, NU = M::NU, NC = M::NC, K = NU + NC, NR = NX + 1;
  static constexpr int NS = cmax(M::NX, M::NXN, M::NXT);
  static constexpr int KP = K * (K + 1) / 2;
  static constexpr int pad(int n) { return n > 0 ? n : 1; }
  static constexpr int mx(int a, int b) { return a > b ? a : b; }
  // offsets in doubles
  static constexpr int LHS = 0;
  static constexpr int RHS = LHS + ((pad(KP) + 1) & ~1);   // (16-byte aligned) K x NR: assembled as [Qu B; c cx], negated, solved -> [alpha beta; psi omega]
  static constexpr int FX = RHS + ((pad(RhsL<(K > 0 ? K : 1), NR>::SIZE) + 1) & ~1);   // NXN x NX (row index = next-state component)
  static constexpr int NFU = M::FU_NC;               // controls the dynamics depend on (non-zero columns of fu)
  static constexpr int FU = FX + NS * NS;            // NXN x NFU, compact: column c belongs to control fu_col(c)
  static constexpr int VXX = FU + pad(NS * NFU);     // value Hessian of knot t+1
  static constexpr int VX = VXX + NS * NS;
  static constexpr int LAM = VX + NS;
  static constexpr int CM = LAM + NS;                // C (NX x NX)
  static constexpr int VEC = CM + NS * NS;           // 1/il 1/iu Sigma^L Sigma^U  (4 x NU)
  static constexpr int PHI = VEC + pad(4 * NU);
  static constexpr int LX = PHI + pad(NC);
  static constexpr int NEWV = LX + NS;               // new Vxx (NX*NX), Vx (NX), lam (NX); before the solve: |dual residual| (NU)
  // PRE: buffers that are dead once the factorisation starts; the 4K-double scratch of the 2x2 pivot update
  // (WS) is aliased on top of them
  static constexpr int PRE = NEWV + mx(NS * NS + 2 * NS, NU);
  static constexpr int DSC = NEWV;
  static constexpr int UXT = PRE;                    // fu' Vxx+  (NFU x NXN)
  static constexpr int XXT = UXT + pad(NFU * NS);    // fx' Vxx+  (NX x NXN)
  static constexpr int TILE = XXT + NS * NS;
  static constexpr int VFS = TILE + pad(mx(M::D_NSLOT, M::DN_NSLOT));
  static constexpr int XS = VFS + pad(M::VF_NSLOT);  // x, u copies for the dynamics Hessian contraction
  static constexpr int US = XS + (M::VF_NSLOT > 0 ? NX : 0);
  static constexpr int PRE_END = US + (M::VF_NSLOT > 0 ? NU : 0);
  static constexpr int WS = PRE;
  static constexpr int DBL_END = mx(PRE_END, WS + 4 * K);
  // LDLT scratch (8-byte aligned: it starts with doubles)
  static constexpr int LIST_B = DBL_END * 8;
  static constexpr int BYTES = ((LIST_B + LdltScratch<(K > 0 ? K : 1)>::BYTES + 15) / 16) * 16;
};

// where one sweep writes its outputs (per-instance base pointers: gains + t*G, Qu + t*NU, lam + t*NX)
struct BwOut {
  double* gains;
  double* Qu;
  double* lam;
};

// regularisation schedule on inertia failure (src/inertia_correction.jl:267-274); reg_last is the value the previous
// backward pass ended with
IPDDP_D double bw_next_reg(const DevView& v, double reg, double reg_last) {
  if (reg == 0.0) return (reg_last == 0.0) ? v.opt.reg_1 : jmax(v.opt.reg_min, v.opt.kappa_w_m * reg_last);
  return (reg_last == 0.0) ? v.opt.kappa_bar_w_p * reg : v.opt.kappa_w_p * reg;
}

// (the buffers of the PRE group -- xxt, uxt, tile, vfs, xs, us -- share memory with ws, and dsc with nVxx / nVx / nlam:
// no __restrict__ there).  L = the (chain) model's layout, S = the stage type whose tables / constants are used.
#define IPDDP_BW_POINTERS(S) \
  double* __restrict__ lhs = sm + L::LHS; double* __restrict__ rhs = sm + L::RHS; \
  double* __restrict__ fx = sm + L::FX; double* __restrict__ fu = sm + L::FU; \
  double* __restrict__ Vxx = sm + L::VXX; double* __restrict__ Vx = sm + L::VX; \
  double* __restrict__ lamn = sm + L::LAM; double* __restrict__ Cm = sm + L::CM; \
  double* xxt = sm + L::XXT; double* uxt = sm + L::UXT; \
  double* __restrict__ ra1 = sm + L::VEC; double* __restrict__ ra2 = sm + L::VEC + S::NU; \
  double* __restrict__ t1 = sm + L::VEC + 2 * S::NU; double* __restrict__ t2 = sm + L::VEC + 3 * S::NU; \
  double* __restrict__ phi = sm + L::PHI; double* xs = sm + L::XS; double* us = sm + L::US; \
  double* __restrict__ lx = sm + L::LX; \
  double* tile = sm + L::TILE; double* vfs = sm + L::VFS; double* ws = sm + L::WS; double* dsc = sm + L::DSC; \
  double* nVxx = sm + L::NEWV; double* nVx = sm + L::NEWV + S::NX * S::NX; double* nlam = sm + L::NEWV + S::NX * S::NX + S::NX; \
  unsigned char* smb = reinterpret_cast<unsigned char*>(sm); \
  unsigned char* nzlist = smb + L::LIST_B; \
  const MEntry* tbl = S::tbl(); \
  const double* cst = S::consts(); \
  (void)lhs; (void)rhs; (void)fx; (void)fu; (void)Vxx; (void)Vx; (void)lamn; (void)Cm; (void)xxt; (void)uxt; (void)ra1; (void)ra2; \
  (void)t1; (void)t2; (void)phi; (void)xs; (void)us; (void)lx; (void)tile; (void)vfs; (void)ws; (void)dsc; (void)nVxx; (void)nVx; \
  (void)nlam; (void)nzlist; (void)tbl; (void)cst;

// constant parts of fx / fu of stage type S in a warp's shared memory (once per warp for a plain model, at every knot of a
// stage chain, whose constants change with the stage type)
template <class M, class S>
IPDDP_D void bw_setup(double* sm, int lane) {
  typedef BwLayout<M> L;
  constexpr int NX = S::NX, NN = S::NXN, NFU = S::FU_NC;
  double* fx = sm + L::FX; double* fu = sm + L::FU;
  const MEntry* tbl = S::tbl();
  const double* cst = S::consts();
  for (int e = lane; e < NN * NX; e += 32) fx[e] = 0.0;
  for (int e = lane; e < NN * NFU; e += 32) fu[e] = 0.0;
  __syncwarp();
  for (int e = lane; e < S::D_fx_N; e += 32) {
    const MEntry q = ld_entry(tbl + S::D_fx_OFF + e);
    if (q.slot < 0) fx[q.i + q.j * NN] = IPDDP_LDG(cst - 1 - q.slot);
  }
  for (int e = lane; e < S::D_fu_N; e += 32) {
    const MEntry q = ld_entry(tbl + S::D_fu_OFF + e);
    if (q.slot < 0) fu[q.i + S::fu_idx(q.j) * NN] = IPDDP_LDG(cst - 1 - q.slot);
  }
  __syncwarp();
}

// per-sweep state shared by the knots of one sweep
struct BwSweep {
  double reg, mu, delta_c, dual_num;
  bool second_order;
  unsigned tri_lane;
};

// terminal knot (K = 0): Vxx = lxx_N, Vx = lx_N, lambda = lx_N, on the terminal stage's NXT states
template <class M>
IPDDP_D void bw_terminal(const DevView& v, int b, int t, double* sm, const BwOut& out, int lane, BwSweep& sw) {
  typedef BwLayout<M> L;
  typedef typename M::Terminal T;
  constexpr int NX = T::NXT;
  IPDDP_BW_POINTERS(T)
  auto val = [&](const MEntry& q) -> double { return q.slot >= 0 ? tile[q.slot] : IPDDP_LDG(cst - 1 - q.slot); };
  for (int e = lane; e < T::DN_NSLOT; e += 32) tile[e] = v.tileN[(size_t)b * (M::DN_NSLOT > 0 ? M::DN_NSLOT : 1) + e];
  for (int e = lane; e < NX * NX; e += 32) Cm[e] = 0.0;
  for (int e = lane; e < NX; e += 32) lx[e] = 0.0;
  __syncwarp();
  for (int e = lane; e < T::DN_lxx_N; e += 32) { const MEntry q = ld_entry(tbl + T::DN_lxx_OFF + e); Cm[q.i + q.j * NX] = val(q); }
  for (int e = lane; e < T::DN_lx_N; e += 32) { const MEntry q = ld_entry(tbl + T::DN_lx_OFF + e); lx[q.i] = val(q); }
  __syncwarp();
  // C = lxx (+ vcxx = 0 unless quasi_newton);  Vxx = (0 + 0) + C ;  Vx = lambda = 0 + lx
  // inertia_correction! on the empty KKT matrix resets delta_c (Q4: the value set by a failed knot
  // never reaches a non-empty KKT matrix)
  sw.delta_c = 0.0;
  for (int e = lane; e < NX * NX; e += 32) Vxx[e] = (0.0 + 0.0) + (sw.second_order ? (Cm[e] + 0.0) : Cm[e]);
  for (int e = lane; e < NX; e += 32) {
    const double l0 = 0.0 + lx[e];
    lamn[e] = l0;
    Vx[e] = (l0 + 0.0) + 0.0;
    out.lam[(size_t)t * L::NS + e] = l0;
  }
  __syncwarp();
}

// One running knot t of stage type S (NX states, NN = NXN next states).  Returns 0, or 1 if the inertia test failed.
template <class M, class S>
IPDDP_D int bw_knot(const DevView& v, int b, int t, int set, double* sm, const BwOut& out, int lane, BwSweep& sw,
                    const double* p) {
  typedef BwLayout<M> L;
  typedef Rec<S> R;
  constexpr int NX = S::NX, NN = S::NXN, NU = S::NU, NC = S::NC, K = NU + NC, NR = NX + 1, NFU = S::FU_NC;
  typedef RhsL<K, NR> RL;          // layout of the K x NR right-hand sides in shared memory (ldlt_warp.cuh)
  IPDDP_BW_POINTERS(S)
  const double reg = sw.reg, mu = sw.mu;
  const bool second_order = sw.second_order;
  const unsigned tri_lane = sw.tri_lane;
  auto val = [&](const MEntry& q) -> double { return q.slot >= 0 ? tile[q.slot] : IPDDP_LDG(cst - 1 - q.slot); };
  if constexpr (M::NSTAGE > 1) bw_setup<M, S>(sm, lane);     // the constants of fx / fu belong to the stage type
  {
    const double* r = v.rec(set, b, t);
    if (t > 0) {   // the next knot's record and tile sectors: pull them towards L2 while this knot is factorised
      const double* rn = v.rec(set, b, t - 1);
      if (lane * 16 < Rec<M>::SIZE) IPDDP_PREFETCH_L2(rn + lane * 16);
      for (int e = lane; e < M::D_NSLOT; e += 32) IPDDP_PREFETCH_L2(v.tile + ((size_t)b * M::D_NSLOT + e) * v.N + (t - 1));
    }
    // ---- stage inputs
    for (int e = lane; e < S::D_NSLOT; e += 32) tile[e] = v.tile[((size_t)b * M::D_NSLOT + e) * v.N + t];
    if constexpr (S::VF_NSLOT > 0) {
      for (int e = lane; e < NU; e += 32) us[e] = r[R::U + e];
      for (int e = lane; e < NX; e += 32) xs[e] = r[R::X + e];
    }
    for (int e = lane; e < NC; e += 32) phi[e] = r[R::PHI + e];
    for (int e = lane; e < K * (K + 1) / 2; e += 32) lhs[e] = 0.0;
    for (int e = lane; e < RL::SIZE; e += 32) rhs[e] = 0.0;
    for (int e = lane; e < NX * NX; e += 32) Cm[e] = 0.0;
    for (int e = lane; e < NX; e += 32) lx[e] = 0.0;
    __syncwarp();
    // ---- scatter pass 1: fx, fu (non-constant part), cu -> lhs top-right, cx -> rhs, lu -> rhs col 0,
    //      lux -> rhs B block, lxx -> C, lx, c -> rhs   (rhs holds the un-negated [Qu B; c cx] until the solve)
    for (int e = lane; e < S::D_fx_N; e += 32) { const MEntry q = ld_entry(tbl + S::D_fx_OFF + e); if (q.slot >= 0) fx[q.i + q.j * NN] = tile[q.slot]; }
    for (int e = lane; e < S::D_fu_N; e += 32) { const MEntry q = ld_entry(tbl + S::D_fu_OFF + e); if (q.slot >= 0) fu[q.i + S::fu_idx(q.j) * NN] = tile[q.slot]; }
    for (int e = lane; e < S::D_cu_N; e += 32) { const MEntry q = ld_entry(tbl + S::D_cu_OFF + e); lhs[pk(q.j, NU + q.i)] = val(q); }
    for (int e = lane; e < S::D_cx_N; e += 32) { const MEntry q = ld_entry(tbl + S::D_cx_OFF + e); rhs[RL::at(NU + q.i, 1 + q.j)] = val(q); }
    for (int e = lane; e < S::D_lu_N; e += 32) { const MEntry q = ld_entry(tbl + S::D_lu_OFF + e); rhs[RL::at(q.i, 0)] = val(q); }
    for (int e = lane; e < S::D_lux_N; e += 32) { const MEntry q = ld_entry(tbl + S::D_lux_OFF + e); rhs[RL::at(q.i, 1 + q.j)] = val(q); }
    for (int e = lane; e < S::D_lxx_N; e += 32) { const MEntry q = ld_entry(tbl + S::D_lxx_OFF + e); Cm[q.i + q.j * NX] = val(q); }
    for (int e = lane; e < S::D_lx_N; e += 32) { const MEntry q = ld_entry(tbl + S::D_lx_OFF + e); lx[q.i] = val(q); }
    for (int e = lane; e < NC; e += 32) rhs[RL::at(NU + e, 0)] = r[R::C + e];
    __syncwarp();
    // ---- barrier terms, Qu, dual-infeasibility numerator            (src/backward_pass.jl:62-75)
    for (int i = lane; i < NU; i += 32) {
      // 1/il, 1/iu.  Unbounded controls have il / iu = +Inf (Q1): 1/Inf = 0 exactly, but the FP64 reciprocal would take
      // its slow path for it, so those lanes divide 1 by 1 and select the zero afterwards
      const double il_ = r[R::IL + i], iu_ = r[R::IU + i];
      const bool il_inf = il_ > 1.7976931348623157e308, iu_inf = iu_ > 1.7976931348623157e308;
      double a1 = 1.0 / (il_inf ? 1.0 : il_), a2 = 1.0 / (iu_inf ? 1.0 : iu_);
      a1 = il_inf ? 0.0 : a1;
      a2 = iu_inf ? 0.0 : a2;
      const double zl_i = r[R::ZL + i], zu_i = r[R::ZU + i];
      const double cl = a1 * mu, cu_ = a2 * mu;
      ra1[i] = a1; ra2[i] = a2;
      const double lu_i = rhs[RL::at(i, 0)];
      double dq = 0.0;   // cu' phi
      {
        double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
        int rr = 0;
        for (; rr + 3 < NC; rr += 4) {
          s0 = IPDDP_FMA(lhs[pk(i, NU + rr + 0)], phi[rr + 0], s0);
          s1 = IPDDP_FMA(lhs[pk(i, NU + rr + 1)], phi[rr + 1], s1);
          s2 = IPDDP_FMA(lhs[pk(i, NU + rr + 2)], phi[rr + 2], s2);
          s3 = IPDDP_FMA(lhs[pk(i, NU + rr + 3)], phi[rr + 3], s3);
        }
        if (rr < NC) s0 = IPDDP_FMA(lhs[pk(i, NU + rr)], phi[rr], s0);
        if (rr + 1 < NC) s1 = IPDDP_FMA(lhs[pk(i, NU + rr + 1)], phi[rr + 1], s1);
        if (rr + 2 < NC) s2 = IPDDP_FMA(lhs[pk(i, NU + rr + 2)], phi[rr + 2], s2);
        dq = (s0 + s1) + (s2 + s3);
      }
      double q = dq + lu_i;
      // fu' Vx: controls outside the fu columns contribute (0 + 0) + (0 + 0) = +0
      const int ci = S::fu_idx(i);
      q = (ci != 255 ? dot4c<NN>(fu + ci * NN, 1, Vx, 1) : 0.0) + q;
      q -= cl;
      q += cu_;
      rhs[RL::at(i, 0)] = q;   // Qu
      // dual error numerator: lu + cu'phi - zl + zu + fu'lambda+      (src/solve.jl:127-132)
      double d = dq + lu_i;
      d -= zl_i;
      d += zu_i;
      d = (ci != 255 ? dot4c<NN>(fu + ci * NN, 1, lamn, 1) : 0.0) + d;
      t1[i] = a1 * zl_i;    // Sigma^L
      t2[i] = a2 * zu_i;    // Sigma^U
      dsc[i] = fabs(d);
    }
    // ---- xx_tmp = fx' Vxx+ (NX x NN) ; ux_tmp = fu' Vxx+ (NFU x NN)      (:80, :91)
    for (int e = lane; e < NX * NN; e += 32) {
      const int i = e % NX, j = e / NX;
      xxt[e] = dot4c<NN>(fx + i * NN, 1, Vxx + j * NN, 1);
    }
    for (int e = lane; e < NFU * NN; e += 32) {     // rows of fu' Vxx+ outside the fu columns are exactly zero: not stored
      const int a = e % NFU, j = e / NFU;
      uxt[e] = dot4c<NN>(fu + a * NN, 1, Vxx + j * NN, 1);
    }
    __syncwarp();
    {  // dual_num = max(dual_num, |.|_inf) -- uniform scan, NaN propagating like Julia's max
      double m = 0.0;
      for (int i = 0; i < NU; ++i) m = jmax(m, dsc[i]);
      sw.dual_num = jmax(sw.dual_num, m);
    }
    // ---- C += xx_tmp fx ;  H = Sigma + ux_tmp fu (upper) ; B += ux_tmp fx    (:81, :86-99)
    for (int e = lane; e < NX * NX; e += 32) {
      const int i = e % NX, j = e / NX;
      Cm[e] = dot4c<NN>(xxt + i, NX, fx + j * NN, 1) + Cm[e];
    }
    // H = (fu' Vxx+ fu) + Sigma: outside the fu columns the product is (0 + 0) + (0 + 0) = +0, i.e. H = +0 + Sigma on the
    // diagonal and +0 (the cleared matrix) elsewhere; the fu columns get the dense formula
    for (int i = lane; i < NU; i += 32)
      if (S::fu_idx(i) == 255) lhs[pk(i, i)] = 0.0 + (t1[i] + t2[i]);
    for (int e = lane; e < NFU * (NFU + 1) / 2; e += 32) {
      const unsigned q = tri_decode(e);
      const int a = q & 0xff, b2 = q >> 8;                 // compact columns a <= b2
      const int i = S::fu_col(a), j = S::fu_col(b2);       // controls i <= j
      const double h0 = (a == b2) ? (t1[i] + t2[i]) : 0.0;
      lhs[pk(i, j)] = dot4c<NN>(uxt + a, NFU, fu + b2 * NN, 1) + h0;
    }
    for (int e = lane; e < NFU * NX; e += 32) {             // B += (fu' Vxx+) fx: rows outside the fu columns get +0
      const int a = e % NFU, j = e / NFU;
      const int i = S::fu_col(a);
      rhs[RL::at(i, 1 + j)] = dot4c<NN>(uxt + a, NFU, fx + j * NN, 1) + rhs[RL::at(i, 1 + j)];
    }
    __syncwarp();
    // ---- H += luu
    for (int e = lane; e < S::D_luu_N; e += 32) { const MEntry q = ld_entry(tbl + S::D_luu_OFF + e); lhs[pk(q.i, q.j)] += val(q); }
    __syncwarp();
    if (second_order) {
      if constexpr (S::VF_NSLOT > 0) {   // dynamics Hessian contraction with lambda+ (:102-110), evaluated redundantly per lane
        double vfl[S::VF_NSLOT > 0 ? S::VF_NSLOT : 1];
        auto st = [&](int s_, double x_) { vfl[s_] = x_; };
        S::vf(xs, us, lamn, p, st);
        if (lane == 0)
          for (int s_ = 0; s_ < S::VF_NSLOT; ++s_) vfs[s_] = vfl[s_];
        __syncwarp();
        for (int e = lane; e < S::VF_vfxx_N; e += 32) { const MEntry q = ld_entry(tbl + S::VF_vfxx_OFF + e); Cm[q.i + q.j * NX] += (q.slot >= 0 ? vfs[q.slot] : IPDDP_LDG(cst - 1 - q.slot)); }
        for (int e = lane; e < S::VF_vfux_N; e += 32) { const MEntry q = ld_entry(tbl + S::VF_vfux_OFF + e); rhs[RL::at(q.i, 1 + q.j)] += (q.slot >= 0 ? vfs[q.slot] : IPDDP_LDG(cst - 1 - q.slot)); }
        for (int e = lane; e < S::VF_vfuu_N; e += 32) { const MEntry q = ld_entry(tbl + S::VF_vfuu_OFF + e); lhs[pk(q.i, q.j)] += (q.slot >= 0 ? vfs[q.slot] : IPDDP_LDG(cst - 1 - q.slot)); }
        __syncwarp();
      }
      for (int e = lane; e < S::D_vcuu_N; e += 32) { const MEntry q = ld_entry(tbl + S::D_vcuu_OFF + e); lhs[pk(q.i, q.j)] += val(q); }
      for (int e = lane; e < S::D_vcux_N; e += 32) { const MEntry q = ld_entry(tbl + S::D_vcux_OFF + e); rhs[RL::at(q.i, 1 + q.j)] += val(q); }
      for (int e = lane; e < S::D_vcxx_N; e += 32) { const MEntry q = ld_entry(tbl + S::D_vcxx_OFF + e); Cm[q.i + q.j * NX] += val(q); }
      __syncwarp();
    }
    if (reg > 0.0)
      for (int i = lane; i < NU; i += 32) lhs[pk(i, i)] += reg;
    if (sw.delta_c > 0.0)
      for (int i = lane; i < NC; i += 32) lhs[pk(NU + i, NU + i)] -= sw.delta_c;
    // ---- park the un-negated [Qu B; c cx] in this knot's gains slot (HBM, read back after the solve), write Qu,
    //      and negate in place: rhs = -[Qu B; c cx]                    (:129-136)
    double* g = out.gains + (size_t)t * v.G;
    double* qo = out.Qu + (size_t)t * M::NU;
    for (int e = lane; e < K * NR; e += 32) {     // e = position in the column-major gains block (row e % K, column e / K)
      const int jc = div_small<K>(e), rr = e - jc * K;
      const double w = rhs[RL::at(rr, jc)]; g[e] = w; rhs[RL::at(rr, jc)] = w * -1.0; if (e < NU) qo[e] = w;
    }
    __syncwarp();
    // ---- factorise + inertia                                          (src/inertia_correction.jl:257-276)
    int np = 0;
    const int info = warp_ldlt_factor<K, NR>(lhs, rhs, ws, nzlist, lane, 1e-12, np, tri_lane);
    sw.delta_c = 0.0;
    if (info > 0) sw.delta_c = v.opt.delta_c * dm::pow(mu, v.opt.kappa_c);
    if (np != NU || info != 0) return 1;   // inertia failure: the caller restarts the sweep with a larger reg
    warp_ldlt_solve_forward<K, NR>(lhs, rhs, nzlist, lane);
    // ---- ineq gains to HBM                                            (:159-172)
    double* gi = g + K * NR;
    for (int e = lane; e < NU * NR; e += 32) {
      const int j = div_small<NU>(e);
      const int i = e - j * NU;
      if (j == 0) {
        const double al = rhs[RL::at(i, 0)];
        double cl = ra1[i] * mu;      // chi^L = mu / il, recomputed from the stored reciprocal (same operands, same bits)
        cl -= r[R::ZL + i];
        cl -= t1[i] * al;
        double cu_ = ra2[i] * mu;
        cu_ -= r[R::ZU + i];
        cu_ += t2[i] * al;
        gi[i] = cl;
        gi[NU + i] = cu_;
      } else {
        const double be = rhs[RL::at(i, j)];
        gi[i + j * 2 * NU] = (be * t1[i]) * -1.0;
        gi[NU + i + j * 2 * NU] = be * t2[i];
      }
    }
    // ---- Vxx = beta' B + omega' cx + C ; Vx ; lambda                  (:176-189)
    // 4 lanes per output element: partial sums over i mod 4, butterfly combine = dot4 order
    {
      const int g4 = lane & 3;
      for (int e0 = 0; e0 < NX * NX; e0 += 8) {
        const int e = e0 + (lane >> 2);
        const int i = e % NX, j = e / NX;
        double sa = 0.0, sb = 0.0;
        if (e < NX * NX) {
          // fully unrolled with the loads of the parked copy (L2 hits) issued ahead of the FMA chains
          double pu[(NU + 3) / 4 > 0 ? (NU + 3) / 4 : 1], pc[(NC + 3) / 4 > 0 ? (NC + 3) / 4 : 1];
#pragma unroll
          for (int qq = 0; qq < (NU + 3) / 4; ++qq) { const int q = g4 + 4 * qq; pu[qq] = (q < NU) ? IPDDP_LDCG(g + q + (1 + j) * K) : 0.0; }
#pragma unroll
          for (int qq = 0; qq < (NC + 3) / 4; ++qq) { const int q = g4 + 4 * qq; pc[qq] = (q < NC) ? IPDDP_LDCG(g + NU + q + (1 + j) * K) : 0.0; }
#pragma unroll
          for (int qq = 0; qq < (NU + 3) / 4; ++qq) { const int q = g4 + 4 * qq; if (q < NU) sa = IPDDP_FMA(rhs[RL::at(q, 1 + i)], pu[qq], sa); }
#pragma unroll
          for (int qq = 0; qq < (NC + 3) / 4; ++qq) { const int q = g4 + 4 * qq; if (q < NC) sb = IPDDP_FMA(rhs[RL::at(NU + q, 1 + i)], pc[qq], sb); }
        }
        sa = sa + __shfl_xor_sync(IPDDP_FULL_MASK, sa, 1);
        sa = sa + __shfl_xor_sync(IPDDP_FULL_MASK, sa, 2);
        sb = sb + __shfl_xor_sync(IPDDP_FULL_MASK, sb, 1);
        sb = sb + __shfl_xor_sync(IPDDP_FULL_MASK, sb, 2);
        if (e < NX * NX && g4 == 0) {
          double w = sa;
          w = sb + w;
          w += Cm[e];
          nVxx[e] = w;
        }
      }
      for (int e0 = 0; e0 < NX; e0 += 8) {
        const int i = e0 + (lane >> 2);
        double sa = 0.0, sb = 0.0, sc = 0.0;
        if (i < NX) {
          double px[(NC + 3) / 4 > 0 ? (NC + 3) / 4 : 1], pq[(NU + 3) / 4 > 0 ? (NU + 3) / 4 : 1], pc[(NC + 3) / 4 > 0 ? (NC + 3) / 4 : 1];
#pragma unroll
          for (int qq = 0; qq < (NC + 3) / 4; ++qq) { const int q = g4 + 4 * qq; px[qq] = (q < NC) ? IPDDP_LDCG(g + NU + q + (1 + i) * K) : 0.0; }
#pragma unroll
          for (int qq = 0; qq < (NU + 3) / 4; ++qq) { const int q = g4 + 4 * qq; pq[qq] = (q < NU) ? IPDDP_LDCG(g + q) : 0.0; }
#pragma unroll
          for (int qq = 0; qq < (NC + 3) / 4; ++qq) { const int q = g4 + 4 * qq; pc[qq] = (q < NC) ? IPDDP_LDCG(g + NU + q) : 0.0; }
#pragma unroll
          for (int qq = 0; qq < (NC + 3) / 4; ++qq) { const int q = g4 + 4 * qq; if (q < NC) sc = IPDDP_FMA(px[qq], phi[q], sc); }               // cx' phi
#pragma unroll
          for (int qq = 0; qq < (NU + 3) / 4; ++qq) { const int q = g4 + 4 * qq; if (q < NU) sa = IPDDP_FMA(rhs[RL::at(q, 1 + i)], pq[qq], sa); }   // beta' Qu
#pragma unroll
          for (int qq = 0; qq < (NC + 3) / 4; ++qq) { const int q = g4 + 4 * qq; if (q < NC) sb = IPDDP_FMA(rhs[RL::at(NU + q, 1 + i)], pc[qq], sb); }   // omega' c
        }
        sa = sa + __shfl_xor_sync(IPDDP_FULL_MASK, sa, 1);
        sa = sa + __shfl_xor_sync(IPDDP_FULL_MASK, sa, 2);
        sb = sb + __shfl_xor_sync(IPDDP_FULL_MASK, sb, 1);
        sb = sb + __shfl_xor_sync(IPDDP_FULL_MASK, sb, 2);
        sc = sc + __shfl_xor_sync(IPDDP_FULL_MASK, sc, 1);
        sc = sc + __shfl_xor_sync(IPDDP_FULL_MASK, sc, 2);
        if (i < NX && g4 == 0) {
          double w = lx[i];
          w = sc + w;
          double lv = w;
          w = sa + w;
          w = sb + w;
          w = dot4c<NN>(fx + i * NN, 1, Vx, 1) + w;
          lv = dot4c<NN>(fx + i * NN, 1, lamn, 1) + lv;
          nVx[i] = w;
          nlam[i] = lv;
        }
      }
    }
    __syncwarp();
    for (int e = lane; e < K * NR; e += 32) {     // eq gains replace the parked copy
      const int jc = div_small<K>(e), rr = e - jc * K;
      g[e] = rhs[RL::at(rr, jc)];
    }
    for (int e = lane; e < NX * NX; e += 32) Vxx[e] = nVxx[e];
    for (int e = lane; e < NX; e += 32) {
      Vx[e] = nVx[e];
      lamn[e] = nlam[e];
      out.lam[(size_t)t * L::NS + e] = nlam[e];
    }
    __syncwarp();
  }
  return 0;
}

// One sweep t = Nb-1 .. 0 of backward_pass! with regularisation `reg` by one warp.  Returns 0 on success, 1 if the
// inertia test failed at some knot (the caller restarts with a larger reg).  nkkt += knots visited;
// dual_num = numerator of the dual infeasibility of this sweep.
template <class M>
IPDDP_D int bw_sweep(const DevView& v, int b, int Nb, int set, double reg, double mu, double* sm, const BwOut& out,
                     int lane, int& nkkt, double& dual_num) {
  const double* p = v.p + (size_t)b * (M::NP > 0 ? M::NP : 1);
  BwSweep sw;
  sw.reg = reg; sw.mu = mu; sw.delta_c = 0.0; sw.dual_num = 0.0;
  sw.second_order = (v.opt.quasi_newton == 0);
  sw.tri_lane = ldlt_tri_lane(lane);
  nkkt++;
  bw_terminal<M>(v, b, Nb - 1, sm, out, lane, sw);
  int status = 0;
  for (int t = Nb - 2; t >= 0; --t) {
    nkkt++;
    for_stage<M>(v.type_of(t), [&](auto tag) { status = bw_knot<M, IPDDP_STAGE(tag)>(v, b, t, set, sm, out, lane, sw, p); });
    if (status != 0) break;
  }
  dual_num = sw.dual_num;
  return status;
}

template <class M>
__global__ void __launch_bounds__(32, IPDDP_BW_MINBLOCKS) k_backward(DevView v, ListView list) {
  IPDDP_DYN_SMEM(double, sm);
  const int lane = threadIdx.x;
  const int inst = blockIdx.x;
  if (inst >= list.total()) return;
  const int b = list.at(inst);
  const int Nb = v.horizon[b];
  const int set = v.nomsel[b];
  if constexpr (M::NSTAGE == 1) bw_setup<M, M>(sm, lane);
  const double mu = v.sdv(SD_MU, b);
  const double reg_last = v.sdv(SD_REG_LAST, b);
  const BwOut out = {v.gains + (size_t)b * (v.N - 1) * v.G, v.Qu + (size_t)b * (v.N - 1) * M::NU,
                     v.lam + (size_t)b * v.N * BwLayout<M>::NS};
  double reg = 0.0, dual_num = 0.0;
  int status = 0, nsweep = 0, nkkt = 0;
  while (reg <= v.opt.reg_max) {
    nsweep++;
    status = bw_sweep<M>(v, b, Nb, set, reg, mu, sm, out, lane, nkkt, dual_num);
    if (status == 0) break;
    reg = bw_next_reg(v, reg, reg_last);
  }
  if (lane == 0) {
    v.sdv(SD_REG_LAST, b) = reg;
    v.sdv(SD_DUAL_NUM, b) = dual_num;
    v.siv(SI_STATUS, b) = status;
    v.siv(SI_NBACK, b) += 1;
    v.siv(SI_NSWEEP, b) += nsweep;
    v.siv(SI_NKKT, b) += nkkt;
    v.siv(SI_LASTSW, b) = nsweep;
  }
}

// Speculative restarts for rounds with few active instances (the lock-step tail): one CTA of BWS_WARPS warps per
// instance, warp w runs the sweep with the w-th value of the regularisation schedule (the schedule only depends on
// reg_last, so it is known in advance); the first sweep IN SCHEDULE ORDER that passes every inertia test is the one the
// sequential loop would have ended with, and the counters add up the sweeps the sequential loop would have run.
// Warp 0 writes the instance's gains / Qu / lambda in place, the others into a private pool (DevView::spec_bw) that
// is copied over if one of them wins.
constexpr int BWS_WARPS = 4;

template <class M>
__global__ void __launch_bounds__(BWS_WARPS * 32) k_backward_spec(DevView v, ListView list) {
  IPDDP_DYN_SMEM(double, sm_all);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int inst = blockIdx.x;
  if (inst >= list.total()) return;
  const int b = list.at(inst);
  const int Nb = v.horizon[b];
  const int set = v.nomsel[b];
  constexpr int WD = BwLayout<M>::BYTES / 8;
  double* sm = sm_all + (size_t)warp * WD;
  double* res_dn = sm_all + (size_t)BWS_WARPS * WD;                 // [BWS_WARPS] dual_num
  int* res_i = reinterpret_cast<int*>(res_dn + BWS_WARPS);          // [BWS_WARPS][2]: verdict (0 ok, 1 failed, 3 not run), nkkt
  if constexpr (M::NSTAGE == 1) bw_setup<M, M>(sm, lane);
  const double mu = v.sdv(SD_MU, b);
  const double reg_last = v.sdv(SD_REG_LAST, b);
  const size_t G1 = (size_t)(v.N - 1) * v.G, Q1 = (size_t)(v.N - 1) * M::NU, L1 = (size_t)v.N * BwLayout<M>::NS;
  BwOut out = {v.gains + (size_t)b * G1, v.Qu + (size_t)b * Q1, v.lam + (size_t)b * L1};
  const BwOut main_out = out;
  if (warp > 0) {
    double* pool = v.spec_bw + ((size_t)inst * (BWS_WARPS - 1) + (warp - 1)) * (G1 + Q1 + L1);
    out.gains = pool; out.Qu = pool + G1; out.lam = pool + G1 + Q1;
  }
  double reg = 0.0, dual_num = 0.0;     // every thread tracks the schedule identically
  int status = 0, nsweep = 0, nkkt = 0;
  for (;;) {
    double mine = reg;
    for (int q = 0; q < warp; ++q) mine = bw_next_reg(v, mine, reg_last);
    // candidates after the first one that exceeds reg_max are never reached by the sequential loop
    bool reachable = true;
    {
      double r = reg;
      for (int q = 0; q < warp; ++q) { if (!(r <= v.opt.reg_max)) reachable = false; r = bw_next_reg(v, r, reg_last); }
    }
    int verdict = 3, kk = 0;
    double dn = 0.0;
    if (reachable && mine <= v.opt.reg_max) verdict = bw_sweep<M>(v, b, Nb, set, mine, mu, sm, out, lane, kk, dn);
    if (lane == 0) { res_i[2 * warp] = verdict; res_i[2 * warp + 1] = kk; res_dn[warp] = dn; }
    __syncthreads();
    int winner = -1;
    bool ended = false;
    for (int q = 0; q < BWS_WARPS; ++q) {
      const int r = res_i[2 * q];
      if (r == 3) { ended = true; status = 1; break; }     // reg > reg_max: the while condition of the sequential loop
      nsweep++;
      nkkt += res_i[2 * q + 1];
      dual_num = res_dn[q];
      if (r == 0) { winner = q; status = 0; break; }
      status = 1;
      reg = bw_next_reg(v, reg, reg_last);
    }
    if (winner > 0) {     // copy the winning sweep's outputs over the instance's arrays
      const double* pool = v.spec_bw + ((size_t)inst * (BWS_WARPS - 1) + (winner - 1)) * (G1 + Q1 + L1);
      const int ng = (Nb - 1) * v.G, nq = (Nb - 1) * M::NU, nl = Nb * BwLayout<M>::NS;
      for (int e = threadIdx.x; e < ng; e += BWS_WARPS * 32) main_out.gains[e] = pool[e];
      for (int e = threadIdx.x; e < nq; e += BWS_WARPS * 32) main_out.Qu[e] = pool[G1 + e];
      for (int e = threadIdx.x; e < nl; e += BWS_WARPS * 32) main_out.lam[e] = pool[G1 + Q1 + e];
    }
    __syncthreads();
    if (winner >= 0 || ended) break;
    if (!(reg <= v.opt.reg_max)) { status = 1; break; }
  }
  if (threadIdx.x == 0) {
    v.sdv(SD_REG_LAST, b) = reg;
    v.sdv(SD_DUAL_NUM, b) = dual_num;
    v.siv(SI_STATUS, b) = status;
    v.siv(SI_NBACK, b) += 1;
    v.siv(SI_NSWEEP, b) += nsweep;
    v.siv(SI_NKKT, b) += nkkt;
    v.siv(SI_LASTSW, b) = nsweep;
  }
}

}  // namespace ipk
