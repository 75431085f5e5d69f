// ORACLE (test infrastructure, NOT product code) -- CPU restatement of the IPDDP2 solve of
// mingu6/InteriorPointDDP.jl v0.5.0 for one OCP instance, plus an OpenMP batch driver.
//
// Each function cites the reference file:line it restates.  Quirks are replicated on purpose
// (Q1 Inf slacks, Q3 NaN->0 after subtracting mu, Q4 delta_c flow, Q5 filter
// augmentation rule, Q6 barrier update is not an iteration).  Deviations, all forced:
//   * summation order inside BLAS-like contractions is fixed to dot4 (ldlt.h); the reference's
//     OpenBLAS order is not reproducible,
//   * elementary functions come from detmath.h,
//   * the DomainError catch of src/forward_pass.jl:18-24 becomes "next state / control not finite",
//   * upper-bound-only projection (src/solver.jl:83-84 is broken upstream) implements the intent.
// Parity pinning: tests/test_oracle_golden.py checks this file against the reference's committed
// results tables (experiments/ipddp2/results/*.txt) using the committed params tables.
#include "ipddp_oracle.h"

#include <math.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "ldlt.h"
#include "oracle_model.h"
#include "models_gen/acrobot.h"
#include "models_gen/cartpole.h"
#include "models_gen/concar.h"
#include "models_gen/concar_quad.h"
#include "models_gen/double_integrator.h"
#include "models_gen/pushing.h"
#include "models_gen/ragged.h"

#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

typedef std::vector<double> vec;

const OracleModel* kModels[] = {&gen_cartpole::model,    &gen_acrobot::model, &gen_concar::model,
                                &gen_concar_quad::model, &gen_pushing::model, &gen_double_integrator::model,
                                &gen_ragged_s0::model,   &gen_ragged_s1::model, &gen_ragged_s2::model};
const int kNumModels = sizeof(kModels) / sizeof(kModels[0]);

const OracleModel* find_model(const char* name) {
  for (int i = 0; i < kNumModels; ++i)
    if (strcmp(kModels[i]->name, name) == 0) return kModels[i];
  return nullptr;
}

// Julia's max/min propagate NaN
inline double jmax(double a, double b) { return (a != a || b != b) ? NAN : (a > b ? a : b); }
inline double jmin(double a, double b) { return (a != a || b != b) ? NAN : (a < b ? a : b); }
inline double norm_inf(const vec& v) {
  double m = 0.0;
  for (double x : v) m = jmax(m, fabs(x));
  return m;
}
inline double norm_1(const vec& v) {
  double s = 0.0;
  for (double x : v) s += fabs(x);
  return s;
}
inline double sum_seq(const vec& v) {
  double s = 0.0;
  for (double x : v) s += x;
  return s;
}

static int g_filter_capacity = 0;   // 0 = unbounded like the reference (oracle_set_filter_capacity)

// one set of trajectories (reference src/data/problem.jl:1-24)
struct Traj {
  std::vector<vec> x, u, c, il, iu, phi, zl, zu, lam;
  void alloc(int N, const std::vector<int>& nx, const std::vector<int>& nu, const std::vector<int>& nc) {
    x.resize(N); u.resize(N); c.resize(N); il.resize(N); iu.resize(N);
    phi.resize(N); zl.resize(N); zu.resize(N); lam.resize(N);
    for (int t = 0; t < N; ++t) {
      x[t].assign(nx[t], 0.0); u[t].assign(nu[t], 0.0); c[t].assign(nc[t], 0.0);
      il[t].assign(nu[t], 0.0); iu[t].assign(nu[t], 0.0); phi[t].assign(nc[t], 0.0);
      zl[t].assign(nu[t], 0.0); zu[t].assign(nu[t], 0.0); lam[t].assign(nx[t], 0.0);
    }
  }
};

struct Bound {  // reference src/bounds.jl:1-17
  vec lower, upper;
  std::vector<int> idx_lower, idx_upper;
  void set(const vec& lo, const vec& up) {
    lower = lo; upper = up; idx_lower.clear(); idx_upper.clear();
    for (size_t i = 0; i < lo.size(); ++i) if (!isinf(lo[i])) idx_lower.push_back((int)i);
    for (size_t i = 0; i < up.size(); ++i) if (!isinf(up[i])) idx_upper.push_back((int)i);
  }
  int num_lower() const { return (int)idx_lower.size(); }
  int num_upper() const { return (int)idx_upper.size(); }
};

struct Solver {
  const OracleModel* m = nullptr;              // stage 0's model (dimension queries of uniform problems)
  std::vector<const OracleModel*> mt;          // model of stage t; mt[N-1] carries the terminal cost (costN / derivsN)
  std::vector<std::vector<int>> compl_t;       // indices_compl of stage t
  int N = 0;
  std::vector<int> nx, nu, nc;
  vec p;
  std::vector<Bound> bounds;
  std::vector<int> indices_compl;
  OracleOptions opt;

  Traj nom, cur;
  // derivative caches (reference src/data/model.jl, objectives.jl, constraints.jl)
  std::vector<vec> fx, fu, vfxx, vfux, vfuu, lx, lu, lxx, luu, lux, cx, cu, vcxx, vcux, vcuu;
  // update rule data (reference src/data/update_rule.jl:89-126)
  std::vector<vec> eq, ineq, Qu, C, H, Bm, Vx, Vxx, lhs, x_tmp, u_tmp1, u_tmp2, xx_tmp, ux_tmp;
  std::vector<std::vector<int>> ipiv;

  // SolverData (reference src/data/solver.jl:8-33)
  double max_primal_1 = 0, min_primal_1 = 0, step_size = 0, mu = 0, reg_last = 0, objective = 0, primal_inf = 0,
         dual_inf = 0, cs_inf = 0, L_curr = 0, theta_curr = 0, L_next = 0, theta_next = 0;
  int status = 0, j = 0, k = 0, l = 0;
  bool switching = false, armijo_passed = false;
  std::vector<std::pair<double, double>> filter;

  long long n_backward = 0, n_sweeps = 0, n_kkt = 0, n_rollouts = 0, n_deriv = 0;
  std::vector<double> trace;

  // ---- gain views (reference src/data/update_rule.jl:71-84): eq[t] is K x (nx+1) column-major,
  // rows 0..nu-1 = [alpha | beta], rows nu..K-1 = [psi | omega]; ineq[t] is 2nu x (nx+1).
  int K(int t) const { return nu[t] + nc[t]; }
  double& alpha(int t, int i) { return eq[t][i]; }
  double& beta(int t, int i, int jx) { return eq[t][i + (1 + jx) * K(t)]; }
  double& psi(int t, int r) { return eq[t][nu[t] + r]; }
  double& omega(int t, int r, int jx) { return eq[t][nu[t] + r + (1 + jx) * K(t)]; }
  double& chil(int t, int i) { return ineq[t][i]; }
  double& zetal(int t, int i, int jx) { return ineq[t][i + (1 + jx) * 2 * nu[t]]; }
  double& chiu(int t, int i) { return ineq[t][nu[t] + i]; }
  double& zetau(int t, int i, int jx) { return ineq[t][nu[t] + i + (1 + jx) * 2 * nu[t]]; }

  void setup(const OracleModel* model, int N_, const double* p_, const double* lo, const double* up,
             const int* compl_idx, int n_compl, const OracleOptions* o) {
    m = model; N = N_;
    mt.assign(N, m);
    nx.assign(N, m->nx); nu.assign(N, m->nu); nc.assign(N, m->nc);
    nu[N - 1] = 0; nc[N - 1] = 0;   // terminal stage (every experiment: Objective(term, nx, 0), Constraint(nx, 0))
    p.assign(p_, p_ + m->np);
    if (p.empty()) p.push_back(0.0);
    bounds.resize(N);
    vec l(lo, lo + m->nu), u(up, up + m->nu);
    for (int t = 0; t < N - 1; ++t) bounds[t].set(l, u);
    bounds[N - 1].set(vec(), vec());
    indices_compl.assign(compl_idx, compl_idx + (compl_idx ? n_compl : 0));
    compl_t.assign(N, indices_compl);
    opt = *o;
    allocate();
  }

  // A chain of stage types (reference src/data/problem.jl:44-62: every buffer is sized per timestep): stage t < N-1 uses
  // types[stage_type[t]], the terminal cost is types[ntypes-1]'s costN on a state of its nxt entries.  lo / up: the types'
  // bounds one after another; compl_idx / n_compl: the types' indices_compl one after another with their counts.
  int setup_chain(const OracleModel* const* types, int ntypes, const int* stage_type, int N_, const double* p_,
                  const double* lo, const double* up, const int* compl_idx, const int* n_compl, const OracleOptions* o) {
    N = N_;
    m = types[stage_type[0]];
    const OracleModel* term = types[ntypes - 1];
    mt.assign(N, term);
    nx.assign(N, 0); nu.assign(N, 0); nc.assign(N, 0);
    int np = 0;
    std::vector<int> boff(ntypes, 0), coff(ntypes, 0);
    for (int k = 0, b = 0, c = 0; k < ntypes; ++k) {
      boff[k] = b; coff[k] = c;
      b += types[k]->nu; c += n_compl ? n_compl[k] : 0;
      if (types[k]->np > np) np = types[k]->np;
    }
    bounds.resize(N);
    compl_t.assign(N, std::vector<int>());
    for (int t = 0; t < N - 1; ++t) {
      const int k = stage_type[t];
      if (k < 0 || k >= ntypes) return -1;
      mt[t] = types[k];
      nx[t] = mt[t]->nx; nu[t] = mt[t]->nu; nc[t] = mt[t]->nc;
      bounds[t].set(vec(lo + boff[k], lo + boff[k] + mt[t]->nu), vec(up + boff[k], up + boff[k] + mt[t]->nu));
      if (compl_idx && n_compl) compl_t[t].assign(compl_idx + coff[k], compl_idx + coff[k] + n_compl[k]);
      if (t > 0 && mt[t - 1]->nxn != nx[t]) return -2;        // x_{t+1} = f_t(x_t, u_t) must fit the next stage
    }
    nx[N - 1] = term->nxt;
    if (mt[N - 2]->nxn != nx[N - 1]) return -2;
    bounds[N - 1].set(vec(), vec());
    p.assign(p_, p_ + np);
    if (p.empty()) p.push_back(0.0);
    indices_compl.clear();
    opt = *o;
    allocate();
    return 0;
  }

  void allocate() {
    nom.alloc(N, nx, nu, nc); cur.alloc(N, nx, nu, nc);
    auto A = [&](std::vector<vec>& v, auto f) { v.resize(N); for (int t = 0; t < N; ++t) v[t].assign(f(t), 0.0); };
    auto nxn = [&](int t) { return t < N - 1 ? nx[t + 1] : 0; };
    A(fx, [&](int t) { return nxn(t) * nx[t]; }); A(fu, [&](int t) { return nxn(t) * nu[t]; });
    A(vfxx, [&](int t) { return nx[t] * nx[t]; }); A(vfux, [&](int t) { return nu[t] * nx[t]; });
    A(vfuu, [&](int t) { return nu[t] * nu[t]; });
    A(lx, [&](int t) { return nx[t]; }); A(lu, [&](int t) { return nu[t]; });
    A(lxx, [&](int t) { return nx[t] * nx[t]; }); A(luu, [&](int t) { return nu[t] * nu[t]; });
    A(lux, [&](int t) { return nu[t] * nx[t]; });
    A(cx, [&](int t) { return nc[t] * nx[t]; }); A(cu, [&](int t) { return nc[t] * nu[t]; });
    A(vcxx, [&](int t) { return nx[t] * nx[t]; }); A(vcux, [&](int t) { return nu[t] * nx[t]; });
    A(vcuu, [&](int t) { return nu[t] * nu[t]; });
    A(eq, [&](int t) { return K(t) * (nx[t] + 1); }); A(ineq, [&](int t) { return 2 * nu[t] * (nx[t] + 1); });
    A(Qu, [&](int t) { return nu[t]; }); A(C, [&](int t) { return nx[t] * nx[t]; });
    A(H, [&](int t) { return nu[t] * nu[t]; }); A(Bm, [&](int t) { return nu[t] * nx[t]; });
    A(Vx, [&](int t) { return nx[t]; }); A(Vxx, [&](int t) { return nx[t] * nx[t]; });
    A(lhs, [&](int t) { return K(t) * K(t); });
    A(x_tmp, [&](int t) { return nx[t]; }); A(u_tmp1, [&](int t) { return nu[t]; }); A(u_tmp2, [&](int t) { return nu[t]; });
    A(xx_tmp, [&](int t) { return nx[t] * nxn(t); }); A(ux_tmp, [&](int t) { return nu[t] * nxn(t); });
    ipiv.resize(N);
    for (int t = 0; t < N; ++t) ipiv[t].assign(K(t) + 1, 0);
  }

  Traj& tr(bool nominal) { return nominal ? nom : cur; }

  // ---------------------------------------------------------------- model evaluation helpers
  void dynamics(int t, const vec& x, const vec& u, vec& xn) { mt[t]->dyn(x.data(), u.data(), p.data(), xn.data()); }

  // reference src/objectives.jl:37-46
  double eval_objective(bool nominal) {
    Traj& T = tr(nominal);
    double J = 0.0, Jp = 0.0;
    for (int t = 0; t < N; ++t) {
      if (t < N - 1) mt[t]->cost(T.x[t].data(), T.u[t].data(), p.data(), &Jp);
      else mt[t]->costN(T.x[t].data(), p.data(), &Jp);
      J += Jp;
    }
    objective = J;
    return J;
  }

  // reference src/data/methods.jl:20-32
  void eval_constraint(double mu_, bool nominal) {
    Traj& T = tr(nominal);
    for (int t = 0; t < N; ++t) {
      if (nc[t] > 0) {
        mt[t]->con(T.x[t].data(), T.u[t].data(), p.data(), T.c[t].data());
        for (int i : compl_t[t]) T.c[t][i] -= mu_;
      }
    }
  }

  // reference src/data/methods.jl:69-76
  double constraint_violation_1norm(bool nominal) {
    Traj& T = tr(nominal);
    double v = 0.0;
    for (int t = 0; t < N; ++t) v += norm_1(T.c[t]);
    return v;
  }

  // reference src/data/methods.jl:34-67 (single running accumulator; logs taken of every slack but
  // only finite-bound indices summed, Q2)
  double barrier_lagrangian(bool nominal) {
    Traj& T = tr(nominal);
    double bl = 0.0;
    for (int t = 0; t < N; ++t) {
      for (int i : bounds[t].idx_lower) bl -= dm_log(T.il[t][i]);
      for (int i : bounds[t].idx_upper) bl -= dm_log(T.iu[t][i]);
    }
    bl *= mu;
    eval_objective(nominal);
    bl += objective;
    for (int t = 0; t < N; ++t) bl += dot4(nc[t], T.c[t].data(), 1, T.phi[t].data(), 1);
    return bl;
  }

  // ---------------------------------------------------------------- reference src/solver.jl:54-105
  void initialize_trajectory(const double* x1, const double* ubar) {
    nom.x[0].assign(x1, x1 + nx[0]);
    const double k1 = opt.kappa_1, k2 = opt.kappa_2;
    int off = 0;
    for (int t = 0; t < N; ++t) {
      const Bound& b = bounds[t];
      for (int i = 0; i < nu[t]; ++i) {
        double u0 = ubar[off + i], lo = b.lower[i], up = b.upper[i], ub;
        if (!isinf(lo) && isinf(up)) {
          double tmp = jmax(lo, 1.0);
          tmp *= k1;
          tmp += lo;
          ub = jmax(u0, tmp);
        } else if (!isinf(up) && isinf(lo)) {
          double tmp = jmax(up, 1.0);   // upstream branch is broken (src/solver.jl:83-84); evident intent
          tmp *= -k1;
          tmp += up;
          ub = jmin(u0, tmp);
        } else if (!isinf(up) && !isinf(lo)) {
          double t1 = lo + jmin(k1 * jmax(1.0, fabs(lo)), k2 * (up - lo));
          double t2 = up - jmin(k1 * jmax(1.0, fabs(up)), k2 * (up - lo));
          ub = jmin(jmax(u0, t1), t2);
        } else {
          ub = u0;
        }
        nom.u[t][i] = ub;
      }
      off += nu[t];
      for (int i = 0; i < nu[t]; ++i) {
        nom.il[t][i] = nom.u[t][i] - b.lower[i];
        nom.iu[t][i] = b.upper[i] - nom.u[t][i];
      }
      if (t < N - 1) dynamics(t, nom.x[t], nom.u[t], nom.x[t + 1]);
    }
  }

  // reference src/solve.jl:182-198
  void reset_duals() {
    for (int t = 0; t < N; ++t) {
      for (Traj* T : {&cur, &nom}) {
        std::fill(T->phi[t].begin(), T->phi[t].end(), 0.0);
        std::fill(T->zl[t].begin(), T->zl[t].end(), 0.0);
        std::fill(T->zu[t].begin(), T->zu[t].end(), 0.0);
        for (int i : bounds[t].idx_lower) T->zl[t][i] = 1.0;
        for (int i : bounds[t].idx_upper) T->zu[t][i] = 1.0;
      }
      std::fill(nom.lam[t].begin(), nom.lam[t].end(), 0.0);
    }
  }

  void reset_filter() {  // reference src/solve.jl:101-105
    filter.clear();
    filter.push_back({max_primal_1, -INFINITY});
    status = 0;
  }

  // reference src/solve.jl:14-38 (everything before the while loop)
  void prologue() {
    for (auto* v : {&fx, &fu, &vfxx, &vfux, &vfuu, &lx, &lu, &lxx, &luu, &lux, &cx, &cu, &vcux, &vcuu})
      for (auto& a : *v) std::fill(a.begin(), a.end(), 0.0);
    max_primal_1 = min_primal_1 = step_size = 0.0;
    status = j = k = l = 0;
    mu = reg_last = objective = primal_inf = dual_inf = cs_inf = 0.0;
    L_curr = theta_curr = L_next = theta_next = 0.0;
    switching = armijo_passed = false;
    filter.assign(1, {0.0, 0.0});
    reset_duals();
    n_backward = n_sweeps = n_kkt = n_rollouts = n_deriv = 0;
    trace.clear();

    eval_objective(true);
    mu = opt.mu_init;
    eval_constraint(mu, true);
    theta_curr = constraint_violation_1norm(true);
    L_curr = barrier_lagrangian(true);
    max_primal_1 = 1e4 * jmax(1.0, theta_curr);
    min_primal_1 = 1e-4 * jmax(1.0, theta_curr);
    reset_filter();
  }

  // ---------------------------------------------------------------- reference src/derivatives.jl:1-35
  void evaluate_derivatives() {
    n_deriv++;
    for (int t = 0; t < N; ++t) {
      if (t < N - 1) {
        mt[t]->derivs(nom.x[t].data(), nom.u[t].data(), nom.phi[t].data(), p.data(), fx[t].data(), fu[t].data(),
                  lx[t].data(), lu[t].data(), lxx[t].data(), luu[t].data(), lux[t].data(), cx[t].data(),
                  cu[t].data(), vcxx[t].data(), vcux[t].data(), vcuu[t].data());
      } else {
        mt[t]->derivsN(nom.x[t].data(), p.data(), lx[t].data(), lxx[t].data());
      }
    }
  }

  // ---------------------------------------------------------------- reference src/inertia_correction.jl:257-276
  int inertia_correction(int t, double& reg, double& delta_c) {
    int n = K(t);
    int st = 0;
    delta_c = 0.0;
    int info = sytf2_rook_upper(n, lhs[t].data(), n > 0 ? n : 1, ipiv[t].data());
    if (info > 0) delta_c = opt.delta_c * dm_pow(mu, opt.kappa_c);
    int np = inertia_np_upper(n, lhs[t].data(), n > 0 ? n : 1, ipiv[t].data(), 1e-12);
    if (np != nu[t] || info != 0) {
      if (reg == 0.0) reg = (reg_last == 0.0) ? opt.reg_1 : jmax(opt.reg_min, opt.kappa_w_m * reg_last);
      else reg = (reg_last == 0.0) ? opt.kappa_bar_w_p * reg : opt.kappa_w_p * reg;
      st = 1;
    }
    return st;
  }

  // ---------------------------------------------------------------- reference src/backward_pass.jl:1-195
  int backward_pass() {
    n_backward++;
    double reg = 0.0, delta_c = 0.0;
    Traj& T = nom;
    while (reg <= opt.reg_max) {
      status = 0;
      n_sweeps++;
      for (int t = N - 1; t >= 0; --t) {
        n_kkt++;
        const int n = nx[t], mu_ = nu[t], pc = nc[t], Kt = mu_ + pc;
        const int nn = (t < N - 1) ? nx[t + 1] : 0;
        vec &t1 = u_tmp1[t], &t2 = u_tmp2[t];
        for (int i = 0; i < mu_; ++i) { t1[i] = 1.0 / T.il[t][i]; t2[i] = 1.0 / T.iu[t][i]; }
        for (int i = 0; i < mu_; ++i) { chil(t, i) = t1[i] * mu; chiu(t, i) = t2[i] * mu; }

        // Qu = lu + cu' phi + fu' Vx+ - mu/il + mu/iu            (:71-75)
        for (int i = 0; i < mu_; ++i) {
          double q = lu[t][i];
          q = dot4(pc, &cu[t][i * pc], 1, T.phi[t].data(), 1) + q;
          if (t < N - 1) q = dot4(nn, &fu[t][i * nn], 1, Vx[t + 1].data(), 1) + q;
          q -= chil(t, i);
          q += chiu(t, i);
          Qu[t][i] = q;
        }
        // C = lxx + fx' Vxx+ fx                                  (:78-82)
        C[t] = lxx[t];
        if (t < N - 1) {
          for (int jn = 0; jn < nn; ++jn)
            for (int i = 0; i < n; ++i)
              xx_tmp[t][i + jn * n] = dot4(nn, &fx[t][i * nn], 1, &Vxx[t + 1][jn * nn], 1);
          for (int jx = 0; jx < n; ++jx)
            for (int i = 0; i < n; ++i)
              C[t][i + jx * n] = dot4(nn, &xx_tmp[t][i], n, &fx[t][jx * nn], 1) + C[t][i + jx * n];
        }
        // H = Sigma + fu' Vxx+ fu + luu                          (:85-95)
        for (int i = 0; i < mu_; ++i) { t1[i] *= T.zl[t][i]; t2[i] *= T.zu[t][i]; }
        std::fill(H[t].begin(), H[t].end(), 0.0);
        for (int i = 0; i < mu_; ++i) H[t][i + i * mu_] = t1[i] + t2[i];
        if (t < N - 1) {
          for (int jn = 0; jn < nn; ++jn)
            for (int i = 0; i < mu_; ++i)
              ux_tmp[t][i + jn * mu_] = dot4(nn, &fu[t][i * nn], 1, &Vxx[t + 1][jn * nn], 1);
          for (int ju = 0; ju < mu_; ++ju)
            for (int i = 0; i < mu_; ++i)
              H[t][i + ju * mu_] = dot4(nn, &ux_tmp[t][i], mu_, &fu[t][ju * nn], 1) + H[t][i + ju * mu_];
        }
        for (int e = 0; e < mu_ * mu_; ++e) H[t][e] += luu[t][e];
        // B = lux + fu' Vxx+ fx                                  (:98-99)
        Bm[t] = lux[t];
        if (t < N - 1)
          for (int jx = 0; jx < n; ++jx)
            for (int i = 0; i < mu_; ++i)
              Bm[t][i + jx * mu_] = dot4(nn, &ux_tmp[t][i], mu_, &fx[t][jx * nn], 1) + Bm[t][i + jx * mu_];
        // second-order contraction terms                          (:102-115)
        if (!opt.quasi_newton) {
          if (t < N - 1) {
            mt[t]->vf(T.x[t].data(), T.u[t].data(), T.lam[t + 1].data(), p.data(), vfxx[t].data(), vfux[t].data(),
                  vfuu[t].data());
            for (int e = 0; e < n * n; ++e) C[t][e] += vfxx[t][e];
            for (int e = 0; e < mu_ * n; ++e) Bm[t][e] += vfux[t][e];
            for (int e = 0; e < mu_ * mu_; ++e) H[t][e] += vfuu[t][e];
          }
          for (int e = 0; e < mu_ * mu_; ++e) H[t][e] += vcuu[t][e];
          for (int e = 0; e < mu_ * n; ++e) Bm[t][e] += vcux[t][e];
          for (int e = 0; e < n * n; ++e) C[t][e] += vcxx[t][e];
        }
        if (reg > 0.0)
          for (int i = 0; i < mu_; ++i) H[t][i + i * mu_] += reg;   // (:118-122)

        // KKT system                                              (:125-142)
        vec& M = lhs[t];
        for (int jc = 0; jc < mu_; ++jc)
          for (int i = 0; i < mu_; ++i) M[i + jc * Kt] = H[t][i + jc * mu_];
        for (int r = 0; r < pc; ++r)
          for (int i = 0; i < mu_; ++i) M[i + (mu_ + r) * Kt] = cu[t][r + i * pc];
        for (int r2 = 0; r2 < pc; ++r2)
          for (int r1 = 0; r1 < pc; ++r1) M[mu_ + r1 + (mu_ + r2) * Kt] = 0.0;
        for (int i = 0; i < mu_; ++i) alpha(t, i) = Qu[t][i] * -1.0;
        for (int r = 0; r < pc; ++r) psi(t, r) = T.c[t][r] * -1.0;
        for (int jx = 0; jx < n; ++jx) {
          for (int i = 0; i < mu_; ++i) beta(t, i, jx) = Bm[t][i + jx * mu_] * -1.0;
          for (int r = 0; r < pc; ++r) omega(t, r, jx) = cx[t][r + jx * pc] * -1.0;
        }
        if (delta_c > 0.0)
          for (int r = 0; r < pc; ++r) M[mu_ + r + (mu_ + r) * Kt] -= delta_c;

        status = inertia_correction(t, reg, delta_c);   // (:144)
        if (status != 0) break;

        sytrs_rook_upper(Kt, n + 1, M.data(), Kt > 0 ? Kt : 1, ipiv[t].data(), eq[t].data(), Kt > 0 ? Kt : 1);

        // inequality-dual gains                                   (:159-172)
        for (int jx = 0; jx < n; ++jx)
          for (int i = 0; i < mu_; ++i) { zetal(t, i, jx) = beta(t, i, jx) * t1[i]; zetal(t, i, jx) *= -1.0; }
        for (int i = 0; i < mu_; ++i) { t1[i] *= alpha(t, i); chil(t, i) -= T.zl[t][i]; chil(t, i) -= t1[i]; }
        for (int jx = 0; jx < n; ++jx)
          for (int i = 0; i < mu_; ++i) zetau(t, i, jx) = beta(t, i, jx) * t2[i];
        for (int i = 0; i < mu_; ++i) { t2[i] *= alpha(t, i); chiu(t, i) -= T.zu[t][i]; chiu(t, i) += t2[i]; }

        // Vxx = beta' B + omega' cx + C                           (:176-178)
        for (int jx = 0; jx < n; ++jx)
          for (int i = 0; i < n; ++i) {
            double v = dot4(mu_, &beta(t, 0, i), 1, &Bm[t][jx * mu_], 1);
            v = dot4(pc, &omega(t, 0, i), 1, &cx[t][jx * pc], 1) + v;
            v += C[t][i + jx * n];
            Vxx[t][i + jx * n] = v;
          }
        // Vx, lambda                                              (:181-189)
        for (int i = 0; i < n; ++i) {
          double v = lx[t][i];
          v = dot4(pc, &cx[t][i * pc], 1, T.phi[t].data(), 1) + v;
          double lamv = v;
          v = dot4(mu_, &beta(t, 0, i), 1, Qu[t].data(), 1) + v;
          v = dot4(pc, &omega(t, 0, i), 1, T.c[t].data(), 1) + v;
          if (t < N - 1) v = dot4(nn, &fx[t][i * nn], 1, Vx[t + 1].data(), 1) + v;
          if (t < N - 1) lamv = dot4(nn, &fx[t][i * nn], 1, T.lam[t + 1].data(), 1) + lamv;
          Vx[t][i] = v;
          T.lam[t][i] = lamv;
        }
      }
      if (status == 0) break;
    }
    reg_last = reg;
    return status;
  }

  // ---------------------------------------------------------------- reference src/solve.jl:107-180
  double primal_error() {
    double e = 0.0;
    for (int t = N - 1; t >= 0; --t) e = jmax(e, norm_inf(nom.c[t]));
    return e;
  }
  double dual_error() {
    double num_ineq = 0.0, z_norm = 0.0, phi_norm = 0.0, dinf = 0.0;
    double num_constr = 0.0;
    for (int t = 0; t < N; ++t) num_constr += nc[t];
    for (int t = N - 1; t >= 0; --t) {
      const int mu_ = nu[t], pc = nc[t];
      const int nn = (t < N - 1) ? nx[t + 1] : 0;
      vec& t1 = u_tmp1[t];
      for (int i = 0; i < mu_; ++i) {
        double v = lu[t][i];
        v = dot4(pc, &cu[t][i * pc], 1, nom.phi[t].data(), 1) + v;
        v -= nom.zl[t][i];
        v += nom.zu[t][i];
        if (t < N - 1) v = dot4(nn, &fu[t][i * nn], 1, nom.lam[t + 1].data(), 1) + v;
        t1[i] = v;
      }
      dinf = jmax(dinf, norm_inf(t1));
      z_norm += sum_seq(nom.zl[t]);
      z_norm += sum_seq(nom.zu[t]);
      phi_norm += norm_1(nom.phi[t]);
      num_ineq += bounds[t].num_lower() + bounds[t].num_upper();
    }
    double scaling = jmax(opt.s_max, (phi_norm + z_norm) / jmax(num_ineq + num_constr, 1.0)) / opt.s_max;
    return dinf / scaling;
  }
  double cs_error(double mu_) {
    double num_ineq = 0.0, z_norm = 0.0, cs = 0.0;
    for (int t = N - 1; t >= 0; --t) {
      num_ineq += bounds[t].num_lower() + bounds[t].num_upper();
      if (bounds[t].num_upper() == 0 && bounds[t].num_lower() == 0) continue;
      vec &t1 = u_tmp1[t], &t2 = u_tmp2[t];
      for (int i = 0; i < nu[t]; ++i) {
        double v = nom.il[t][i];
        v *= nom.zl[t][i];
        v -= mu_;
        if (v != v) v = 0.0;   // replace!(NaN => 0) after subtracting mu (Q3)
        t1[i] = v;
      }
      cs = jmax(cs, norm_inf(t1));
      for (int i = 0; i < nu[t]; ++i) {
        double v = nom.iu[t][i];
        v *= nom.zu[t][i];
        v -= mu_;
        if (v != v) v = 0.0;
        t2[i] = v;
      }
      cs = jmax(cs, norm_inf(t2));
      z_norm += sum_seq(nom.zl[t]);
      z_norm += sum_seq(nom.zu[t]);
    }
    double scaling = jmax(opt.s_max, z_norm / jmax(num_ineq, 1.0)) / opt.s_max;
    return cs / scaling;
  }

  // ---------------------------------------------------------------- reference src/forward_pass.jl:98-153
  bool rollout(double step) {
    n_rollouts++;
    cur.x[0] = nom.x[0];
    for (int t = 0; t < N; ++t) {
      const int n = nx[t], mu_ = nu[t], pc = nc[t];
      vec& dx = x_tmp[t];
      for (int i = 0; i < n; ++i) dx[i] = cur.x[t][i] - nom.x[t][i];
      for (int i = 0; i < mu_; ++i) {
        double v = alpha(t, i);
        v *= step;
        v += nom.u[t][i];
        cur.u[t][i] = dot4(n, &beta(t, i, 0), K(t), dx.data(), 1) + v;
      }
      for (int r = 0; r < pc; ++r) {
        double v = psi(t, r);
        v *= step;
        v += nom.phi[t][r];
        cur.phi[t][r] = dot4(n, &omega(t, r, 0), K(t), dx.data(), 1) + v;
      }
      for (int i = 0; i < mu_; ++i) {
        double v = chil(t, i);
        v *= step;
        v += nom.zl[t][i];
        cur.zl[t][i] = dot4(n, &zetal(t, i, 0), 2 * mu_, dx.data(), 1) + v;
      }
      for (int i = 0; i < mu_; ++i) {
        double v = chiu(t, i);
        v *= step;
        v += nom.zu[t][i];
        cur.zu[t][i] = dot4(n, &zetau(t, i, 0), 2 * mu_, dx.data(), 1) + v;
      }
      if (t < N - 1) dynamics(t, cur.x[t], cur.u[t], cur.x[t + 1]);
      for (int i = 0; i < mu_; ++i) {
        cur.il[t][i] = cur.u[t][i] - bounds[t].lower[i];
        cur.iu[t][i] = bounds[t].upper[i] - cur.u[t][i];
      }
      // DomainError analogue (src/forward_pass.jl:18-24): a non-finite control or next state rejects the step
      bool ok = true;
      for (int i = 0; i < mu_; ++i) ok = ok && isfinite(cur.u[t][i]);
      if (t < N - 1) for (int i = 0; i < nx[t + 1]; ++i) ok = ok && isfinite(cur.x[t + 1][i]);
      if (!ok) return false;
    }
    return true;
  }

  // reference src/forward_pass.jl:59-85
  int check_fraction_boundary(double tau) {
    for (int t = 0; t < N; ++t) {
      for (int i = 0; i < nu[t]; ++i) if (nom.il[t][i] * (1.0 - tau) > cur.il[t][i]) return 2;
      for (int i = 0; i < nu[t]; ++i) if (nom.iu[t][i] * (1.0 - tau) > cur.iu[t][i]) return 2;
      for (int i = 0; i < nu[t]; ++i) if (nom.zl[t][i] * (1.0 - tau) > cur.zl[t][i]) return 2;
      for (int i = 0; i < nu[t]; ++i) if (nom.zu[t][i] * (1.0 - tau) > cur.zu[t][i]) return 2;
    }
    return 0;
  }

  // reference src/forward_pass.jl:87-96
  double expected_change_lagrangian() {
    double dL = 0.0;
    for (int t = N - 1; t >= 0; --t) {
      dL += dot4(nu[t], Qu[t].data(), 1, &alpha(t, 0), 1);
      dL += dot4(nc[t], nom.c[t].data(), 1, &psi(t, 0), 1);
    }
    return dL;
  }

  // reference src/forward_pass.jl:1-57
  int forward_pass() {
    const double eps = 2.220446049250313e-16;
    l = 0;
    status = 0;
    step_size = 1.0;
    const double tau = jmax(opt.tau_min, 1.0 - mu);
    const double theta_prev = theta_curr, L_prev = L_curr;
    double theta = theta_prev;
    const double dL = expected_change_lagrangian();
    while (step_size >= eps) {
      const double gamma = step_size;
      if (!rollout(gamma)) { step_size *= 0.5; continue; }
      status = check_fraction_boundary(tau);
      if (status != 0) { step_size *= 0.5; continue; }
      eval_constraint(mu, false);
      theta = constraint_violation_1norm(false);
      double L = barrier_lagrangian(false);
      bool blocked = false;
      for (auto& f : filter) if (theta >= f.first && L >= f.second) { blocked = true; break; }
      status = blocked ? 3 : 0;
      if (status != 0) { step_size *= 0.5; l += 1; continue; }
      switching = (dL < 0.0) &&
                  (dm_pow(-gamma * dL, opt.s_L) * dm_pow(gamma, 1.0 - opt.s_L) > opt.delta * dm_pow(theta_prev, opt.s_theta));
      armijo_passed = L - L_prev - 10.0 * eps * fabs(L_prev) <= opt.eta_L * gamma * dL;
      if (theta <= min_primal_1 && switching) {
        status = armijo_passed ? 0 : 4;
      } else {
        bool suff = (theta <= (1.0 - opt.gamma_theta) * theta_prev) || (L <= L_prev - opt.gamma_L * theta_prev);
        status = suff ? 0 : 5;
      }
      if (status != 0) { step_size *= 0.5; l += 1; continue; }
      L_next = L;
      theta_next = theta;
      break;
    }
    if (step_size < eps) status = 7;
    return status;
  }

  // reference src/data/methods.jl:78-91 + src/solve.jl:80-85,95-99
  void accept_step() {
    nom.x = cur.x; nom.u = cur.u; nom.c = cur.c; nom.il = cur.il; nom.iu = cur.iu;
    nom.phi = cur.phi; nom.zl = cur.zl; nom.zu = cur.zu; nom.lam = cur.lam;
    if (!armijo_passed && !switching) {
      // the device keeps a fixed-capacity filter (IPDDP_FILTER_CAPACITY) and ends the instance with status 9 when an
      // accepted step would overflow it; the reference's filter is an unbounded Vector.  Off (0) unless a test asks.
      if (g_filter_capacity > 0 && (int)filter.size() >= g_filter_capacity) { status = 9; return; }
      filter.push_back({(1.0 - opt.gamma_theta) * theta_curr, L_curr - opt.gamma_L * theta_curr});
    }
    L_curr = L_next;
    theta_curr = theta_next;
    k += 1;
    const double row[ORACLE_TRACE_COLS] = {(double)k, (double)j, objective, primal_inf, dual_inf, cs_inf, mu,
                                           reg_last, step_size, (double)l, theta_curr, L_curr};
    trace.insert(trace.end(), row, row + ORACLE_TRACE_COLS);
  }

  // ---------------------------------------------------------------- reference src/solve.jl:6-93
  int solve() {
    prologue();
    int num_bounds = 0;
    for (int t = 0; t < N; ++t) num_bounds += bounds[t].num_lower() + bounds[t].num_upper();
    while (k < opt.max_iterations) {
      evaluate_derivatives();
      backward_pass();
      if (status != 0) break;
      dual_inf = dual_error();
      primal_inf = primal_error();
      cs_inf = cs_error(0.0);
      double cs_mu = cs_error(mu);
      double err_mu = jmax(jmax(dual_inf, cs_mu), primal_inf);
      double err_0 = jmax(jmax(dual_inf, cs_inf), primal_inf);
      if (err_0 < opt.optimality_tolerance) break;
      if (err_mu <= opt.kappa_eps * mu && num_bounds > 0 && mu > opt.optimality_tolerance / 10.0) {
        mu = jmax(opt.optimality_tolerance / 10.0, jmin(opt.kappa_mu * mu, dm_pow(mu, opt.theta_mu)));
        reset_filter();
        eval_constraint(mu, true);
        L_curr = barrier_lagrangian(true);
        theta_curr = constraint_violation_1norm(true);
        j += 1;
        continue;
      }
      forward_pass();
      if (status != 0) break;
      accept_step();
      if (status != 0) break;
    }
    if (k == opt.max_iterations) status = 8;
    return status;
  }
};

int get_named(Solver& s, const std::string& name, double* out) {
  std::vector<vec>* a = nullptr;
  Traj* T = &s.nom;
  std::string nm = name;
  if (nm.rfind("cur_", 0) == 0) { T = &s.cur; nm = nm.substr(4); }
  if (nm == "x") a = &T->x; else if (nm == "u") a = &T->u; else if (nm == "c") a = &T->c;
  else if (nm == "il") a = &T->il; else if (nm == "iu") a = &T->iu; else if (nm == "phi") a = &T->phi;
  else if (nm == "zl") a = &T->zl; else if (nm == "zu") a = &T->zu; else if (nm == "lam") a = &T->lam;
  else if (nm == "fx") a = &s.fx; else if (nm == "fu") a = &s.fu; else if (nm == "lx") a = &s.lx;
  else if (nm == "lu") a = &s.lu; else if (nm == "lxx") a = &s.lxx; else if (nm == "luu") a = &s.luu;
  else if (nm == "lux") a = &s.lux; else if (nm == "cx") a = &s.cx; else if (nm == "cu") a = &s.cu;
  else if (nm == "vcxx") a = &s.vcxx; else if (nm == "vcux") a = &s.vcux; else if (nm == "vcuu") a = &s.vcuu;
  else if (nm == "vfxx") a = &s.vfxx; else if (nm == "vfux") a = &s.vfux; else if (nm == "vfuu") a = &s.vfuu;
  else if (nm == "eq") a = &s.eq; else if (nm == "ineq") a = &s.ineq; else if (nm == "Qu") a = &s.Qu;
  else if (nm == "Vx") a = &s.Vx; else if (nm == "Vxx") a = &s.Vxx;
  if (!a) return -1;
  int n = 0;
  for (auto& v : *a) {
    if (out) memcpy(out + n, v.data(), v.size() * sizeof(double));
    n += (int)v.size();
  }
  return n;
}

void fill_result(const Solver& s, OracleResult* r) {
  r->status = s.status; r->k = s.k; r->j = s.j; r->l = s.l;
  r->objective = s.objective; r->primal_inf = s.primal_inf; r->dual_inf = s.dual_inf; r->cs_inf = s.cs_inf;
  r->mu = s.mu; r->reg_last = s.reg_last; r->step_size = s.step_size;
  r->barrier_lagrangian = s.L_curr; r->primal_1 = s.theta_curr;
  r->n_backward = s.n_backward; r->n_sweeps = s.n_sweeps; r->n_kkt = s.n_kkt; r->n_rollouts = s.n_rollouts;
  r->n_deriv = s.n_deriv;
}

}  // namespace

extern "C" {

void oracle_set_filter_capacity(int cap) { g_filter_capacity = cap; }

void oracle_default_options(OracleOptions* o) {  // reference src/options.jl:1-38
  o->quasi_newton = 0; o->optimality_tolerance = 1.0e-8; o->max_iterations = 1000; o->reset_cache = 1;
  o->verbose = 0; o->print_frequency = 10; o->mu_init = 1.0; o->ineq_dual_init = 1.0; o->kappa_1 = 0.01;
  o->kappa_2 = 0.01; o->reg_1 = 1e-4; o->reg_min = 1e-20; o->reg_max = 1e40; o->kappa_bar_w_p = 100.0;
  o->kappa_w_p = 8.0; o->kappa_w_m = 1.0 / 3.0; o->kappa_c = 0.25; o->delta_c = 1e-8; o->kappa_eps = 10.0;
  o->kappa_mu = 0.2; o->theta_mu = 1.2; o->tau_min = 0.99; o->s_max = 100.0; o->eta_L = 1e-4; o->s_L = 2.3;
  o->delta = 1.0; o->s_theta = 1.1; o->gamma_alpha = 0.05; o->gamma_theta = 1e-5; o->gamma_L = 1e-5;
  o->kappa_Sigma = 1e10;
}

int oracle_num_models(void) { return kNumModels; }
const char* oracle_model_name(int i) { return (i >= 0 && i < kNumModels) ? kModels[i]->name : nullptr; }
int oracle_model_dims(const char* model, int* nx, int* nu, int* nc, int* np) {
  const OracleModel* m = find_model(model);
  if (!m) return -1;
  *nx = m->nx; *nu = m->nu; *nc = m->nc; *np = m->np;
  return 0;
}

void* oracle_create(const char* model, int N, const double* p, const double* lower, const double* upper,
                    const int* indices_compl, int n_compl, const OracleOptions* opt) {
  const OracleModel* m = find_model(model);
  if (!m || N < 2) return nullptr;
  OracleOptions o;
  if (opt) o = *opt; else oracle_default_options(&o);
  Solver* s = new Solver();
  s->setup(m, N, p, lower, upper, indices_compl, n_compl, &o);
  return s;
}
// Chain of stage types: type_names[ntypes] (registered models), stage_type[N-1] per running stage; bounds and
// indices_compl of the types one after another (see Solver::setup_chain).  p: max over the types' np doubles.
void* oracle_create_chain(const char* const* type_names, int ntypes, const int* stage_type, int N, const double* p,
                          const double* lower, const double* upper, const int* indices_compl, const int* n_compl,
                          const OracleOptions* opt) {
  if (ntypes < 1 || ntypes > 16 || N < 2) return nullptr;
  const OracleModel* types[16];
  for (int k = 0; k < ntypes; ++k) {
    types[k] = find_model(type_names[k]);
    if (!types[k]) return nullptr;
  }
  OracleOptions o;
  if (opt) o = *opt; else oracle_default_options(&o);
  Solver* s = new Solver();
  if (s->setup_chain(types, ntypes, stage_type, N, p, lower, upper, indices_compl, n_compl, &o) != 0) { delete s; return nullptr; }
  return s;
}
void oracle_destroy(void* h) { delete (Solver*)h; }

int oracle_solve(void* h, const double* x1, const double* ubar) {
  Solver* s = (Solver*)h;
  s->initialize_trajectory(x1, ubar);
  return s->solve();
}
int oracle_resolve(void* h) { return ((Solver*)h)->solve(); }
void oracle_get_result(void* h, OracleResult* r) { fill_result(*(Solver*)h, r); }
int oracle_trace_rows(void* h) { return (int)(((Solver*)h)->trace.size() / ORACLE_TRACE_COLS); }
void oracle_get_trace(void* h, double* out) {
  Solver* s = (Solver*)h;
  memcpy(out, s->trace.data(), s->trace.size() * sizeof(double));
}

void oracle_initialize(void* h, const double* x1, const double* ubar) {
  Solver* s = (Solver*)h;
  s->initialize_trajectory(x1, ubar);
  s->prologue();
}
void oracle_eval_derivatives(void* h) { ((Solver*)h)->evaluate_derivatives(); }
int oracle_backward_pass(void* h) { return ((Solver*)h)->backward_pass(); }
void oracle_errors(void* h, double* d, double* pr, double* cs0, double* csmu) {
  Solver* s = (Solver*)h;
  *d = s->dual_error(); *pr = s->primal_error(); *cs0 = s->cs_error(0.0); *csmu = s->cs_error(s->mu);
  s->dual_inf = *d; s->primal_inf = *pr; s->cs_inf = *cs0;
}
int oracle_forward_pass(void* h) { return ((Solver*)h)->forward_pass(); }
void oracle_accept_step(void* h) { ((Solver*)h)->accept_step(); }
void oracle_set_mu(void* h, double mu) { ((Solver*)h)->mu = mu; }
int oracle_get_array(void* h, const char* name, double* out) { return get_named(*(Solver*)h, name, out); }

int oracle_solve_batch(const char* model, int B, int N, const int* horizons, const double* p, const double* lower,
                       const double* upper, const double* x1, const double* ubar, const OracleOptions* opt,
                       int nthreads, OracleResult* results, double* xout, double* uout) {
  const OracleModel* m = find_model(model);
  if (!m) return -1;
  OracleOptions o;
  if (opt) o = *opt; else oracle_default_options(&o);
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
  const int np = m->np > 0 ? m->np : 1;
#pragma omp parallel for schedule(dynamic, 1)
  for (int b = 0; b < B; ++b) {
    const int Nb = horizons ? horizons[b] : N;
    Solver s;
    s.setup(m, Nb, p + (size_t)b * np, lower + (size_t)b * m->nu, upper + (size_t)b * m->nu, nullptr, 0, &o);
    s.initialize_trajectory(x1 + (size_t)b * m->nx, ubar + (size_t)b * (N - 1) * m->nu);
    s.solve();
    fill_result(s, &results[b]);
    if (xout)
      for (int t = 0; t < Nb; ++t) memcpy(xout + ((size_t)b * N + t) * m->nx, s.nom.x[t].data(), m->nx * sizeof(double));
    if (uout)
      for (int t = 0; t < Nb - 1; ++t)
        memcpy(uout + ((size_t)b * (N - 1) + t) * m->nu, s.nom.u[t].data(), m->nu * sizeof(double));
  }
  return 0;
}

void oracle_detmath(int fn, int n, const double* x, const double* y, double* out) {
  for (int i = 0; i < n; ++i) {
    switch (fn) {
      case 0: out[i] = dm_sin(x[i]); break;
      case 1: out[i] = dm_cos(x[i]); break;
      case 2: out[i] = dm_tan(x[i]); break;
      case 3: out[i] = dm_log(x[i]); break;
      case 4: out[i] = dm_exp(x[i]); break;
      default: out[i] = dm_pow(x[i], y[i]); break;
    }
  }
}
int oracle_sytf2_rook(int n, double* A, int lda, int* ipiv) { return sytf2_rook_upper(n, A, lda, ipiv); }
void oracle_sytrs_rook(int n, int nrhs, const double* A, int lda, const int* ipiv, double* B, int ldb) {
  sytrs_rook_upper(n, nrhs, A, lda, ipiv, B, ldb);
}
int oracle_inertia_np(int n, const double* A, int lda, const int* ipiv, double tol) {
  return inertia_np_upper(n, A, lda, ipiv, tol);
}
}
