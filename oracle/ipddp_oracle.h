// ORACLE (test infrastructure) -- C API of the CPU restatement of InteriorPointDDP.jl's IPDDP2 solve.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
// this library.  The product (libipddp_b200.so) never links, includes or calls it.
#pragma once
#ifdef __cplusplus
extern "C" {
#endif

// Mirror of reference src/options.jl:1-38 (31 fields, same order, same defaults via oracle_default_options).
typedef struct OracleOptions {
  int quasi_newton;
  double optimality_tolerance;
  int max_iterations;
  int reset_cache;
  int verbose;
  int print_frequency;
  double mu_init;
  double ineq_dual_init;
  double kappa_1;
  double kappa_2;
  double reg_1;
  double reg_min;
  double reg_max;
  double kappa_bar_w_p;
  double kappa_w_p;
  double kappa_w_m;
  double kappa_c;
  double delta_c;
  double kappa_eps;
  double kappa_mu;
  double theta_mu;
  double tau_min;
  double s_max;
  double eta_L;
  double s_L;
  double delta;
  double s_theta;
  double gamma_alpha;
  double gamma_theta;
  double gamma_L;
  double kappa_Sigma;
} OracleOptions;

// SolverData scalars (reference src/data/solver.jl:8-33) + work counters used by the benchmark.
typedef struct OracleResult {
  int status, k, j, l;
  double objective, primal_inf, dual_inf, cs_inf, mu, reg_last, step_size;
  double barrier_lagrangian, primal_1;
  long long n_backward, n_sweeps, n_kkt, n_rollouts, n_deriv;
} OracleResult;

// one trace row per accepted iteration, recorded right after the forward pass accepted a step
// (the numeric columns of reference src/print.jl:13-29 plus theta and L)
#define ORACLE_TRACE_COLS 12
// cols: k, j, objective, primal_inf, dual_inf, cs_inf, mu, reg_last, step_size, l, theta(primal_1), barrier_lagrangian

void oracle_default_options(OracleOptions* o);
int oracle_num_models(void);
const char* oracle_model_name(int i);
int oracle_model_dims(const char* model, int* nx, int* nu, int* nc, int* np);

// N = number of knots (running stages 0..N-2 use the model, stage N-1 is the terminal stage with nu=nc=0).
// lower/upper: nu doubles (+-inf allowed), applied to every running stage.  indices_compl: 0-based, may be NULL.
void* oracle_create(const char* model, int N, const double* p, const double* lower, const double* upper,
                    const int* indices_compl, int n_compl, const OracleOptions* opt);
// chain of stage types (state / control sizes may change along the horizon; reference src/data/problem.jl:44-62):
// type_names[ntypes] registered models, stage_type[N-1]; lower/upper and indices_compl/n_compl: per type, concatenated.
void* oracle_create_chain(const char* const* type_names, int ntypes, const int* stage_type, int N, const double* p,
                          const double* lower, const double* upper, const int* indices_compl, const int* n_compl,
                          const OracleOptions* opt);
void oracle_set_filter_capacity(int cap);   // 0 = unbounded (reference); n = the device's fixed filter capacity
void oracle_destroy(void* h);

// reference solve!(solver, x1, controls) (src/solve.jl:1-4); ubar is (N-1)*nu doubles
int oracle_solve(void* h, const double* x1, const double* ubar);
// reference solve!(solver) (src/solve.jl:6-93): warm start from the stored nominal trajectory
int oracle_resolve(void* h);
void oracle_get_result(void* h, OracleResult* r);
int oracle_trace_rows(void* h);
void oracle_get_trace(void* h, double* out);  // rows x ORACLE_TRACE_COLS

// phase-level entry points for kernel parity tests
void oracle_initialize(void* h, const double* x1, const double* ubar);   // initialize_trajectory! + solve! prologue
void oracle_eval_derivatives(void* h);
int oracle_backward_pass(void* h);
void oracle_errors(void* h, double* dual_inf, double* primal_inf, double* cs_inf0, double* cs_inf_mu);
int oracle_forward_pass(void* h);
void oracle_accept_step(void* h);   // the post-forward bookkeeping of src/solve.jl:80-85
void oracle_set_mu(void* h, double mu);

// named array access: nominal "x","u","c","il","iu","phi","zl","zu","lam"; current "cur_x",...;
// derivative caches "fx","fu","lx","lu","lxx","luu","lux","cx","cu","vcxx","vcux","vcuu";
// gains "eq","ineq","Qu"; value "Vx","Vxx".  Arrays are concatenated over stages 0..N-1.
// Returns the number of doubles (call with out=NULL to size).
int oracle_get_array(void* h, const char* name, double* out);

// batch driver (OpenMP over instances) used as the CPU baseline.  Arrays are instance-major.
// horizons may be NULL (all = N).  results: B entries.  xout/uout optional (B*N*nx, B*(N-1)*nu).
int oracle_solve_batch(const char* model, int B, int N, const int* horizons, const double* p, const double* lower,
                       const double* upper, const double* x1, const double* ubar, const OracleOptions* opt,
                       int nthreads, OracleResult* results, double* xout, double* uout);

// elementwise evaluation of the deterministic math layer: fn 0 sin,1 cos,2 tan,3 log,4 exp,5 pow(x, y[i])
void oracle_detmath(int fn, int n, const double* x, const double* y, double* out);
// LAPACK-restatement test hooks
int oracle_sytf2_rook(int n, double* A, int lda, int* ipiv);
void oracle_sytrs_rook(int n, int nrhs, const double* A, int lda, const int* ipiv, double* B, int ldb);
int oracle_inertia_np(int n, const double* A, int lda, const int* ipiv, double tol);

#ifdef __cplusplus
}
#endif
