/* ipddp_b200.h -- C ABI of libipddp_b200.so: batched IPDDP2 on NVIDIA B200 (sm_100a), FP64.
 *
 * Drop-in boundary for the hot path of mingu6/InteriorPointDDP.jl v0.5.0.  The reference has no FFI
 * layer of its own; this ABI sits directly under its exported Julia API (reference
 * src/InteriorPointDDP.jl:29-45) and each entry point names the reference function(s) it replaces.
 * A Julia wrapper binds these with `ccall` (see INTEGRATION.md and julia/InteriorPointDDPB200.jl);
 * tests and bench.py bind them with Python ctypes.
 *
 * Conventions
 *   - plain pointers and sizes only; caller owns every host buffer; the handle owns all device memory.
 *   - return value: 0 = ok, negative = API / CUDA error (text via ipddp_last_error()).  The algorithmic
 *     outcome is per instance in `status`, with the reference's codes (reference src/data/solver.jl:5-7):
 *     0 solved, 1 backward pass failed (reg > reg_max), 7 line search exhausted, 8 max iterations,
 *     plus 9 = filter capacity exceeded (the reference's filter is an unbounded Vector).
 *   - all matrices column-major, all reals FP64.
 *   - a "knot" is a timestep; N knots = N-1 running stages (nx,nu,nc of the model) + 1 terminal stage
 *     (nx, 0, 0), as in every reference experiment.  Per-instance horizons may be shorter than N.
 *   - one handle = one device + one stream (its own, or the caller's: ipddp_set_stream); calls on a handle must be
 *     serialised by the caller.
 */
#ifndef IPDDP_B200_H
#define IPDDP_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define IPDDP_ABI_VERSION 2
#define IPDDP_TRACE_COLS 12
#define IPDDP_FILTER_CAPACITY 64

/* Mirror of reference `Options{T}` (src/options.jl:1-38): 31 fields, same order.  Unused upstream
 * fields (reset_cache, ineq_dual_init, gamma_alpha, kappa_Sigma) are kept for layout compatibility. */
typedef struct ipddp_options {
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
} ipddp_options;

/* Per-solve device timing and work counters (the reference keeps wall/solver/fn_eval seconds in
 * SolverData, src/data/solver.jl:16-18; here the split is per kernel, measured with CUDA events). */
typedef struct ipddp_stats {
  int iterations;              /* lock-step outer rounds executed (max over instances of k + j + 1) */
  long long launches;          /* kernels launched by the last ipddp_solve */
  double ms_total;             /* device time of the whole solve */
  double ms_init, ms_derivs, ms_backward, ms_check, ms_forward;
  long long sum_backward, sum_sweeps, sum_kkt, sum_rollouts, sum_deriv_stages; /* summed over instances */
  long long n_converged;       /* instances with status 0 */
  long long n_active_rounds;   /* sum over rounds of active instances (occupancy of the lock-step loop) */
  double sum_active_sq;        /* sum over rounds of (active instances)^2: sum_active_sq / n_active_rounds = the round size
                                  the average instance-iteration ran in (work-weighted occupancy) */
} ipddp_stats;

typedef struct ipddp_problem ipddp_problem;

int ipddp_abi_version(void);
const char* ipddp_last_error(void);

/* reference Options{T}() defaults (src/options.jl:1-38) */
void ipddp_default_options(ipddp_options* opt);

/* Model registry.  A model is the compiled counterpart of the reference's generated closures
 * (Dynamics/Objective/Constraint, src/dynamics.jl:15-47, src/objectives.jl:12-33,
 * src/constraints.jl:16-50).  Built-in: cartpole, acrobot, concar, concar_quad, pushing,
 * double_integrator, and the stage chain `ragged` (a test problem whose state grows from 2 to 3 and whose controls shrink
 * from 3 to 2 along the horizon).  ipddp_model_load registers a plugin .so produced by the code generator. */
int ipddp_num_models(void);
const char* ipddp_model_name(int index);
int ipddp_model_dims(const char* model, int* nx, int* nu, int* nc, int* np, int* tile_slots);
int ipddp_model_load(const char* plugin_path);

/* Stage chains -- horizons whose state and control sizes change from stage to stage (reference README.md:18,
 * src/data/problem.jl:44-62: every buffer is sized per timestep, src/solver.jl:11-26 takes one Dynamics / Objective /
 * Constraint / Bound per stage).  A chain model is compiled from up to 4 stage TYPES (each with its own nx, nu, nc and next
 * state size nxn; the last type carries the terminal cost on nxt states); a problem assigns a type to each running stage:
 *   ipddp_model_stages     the types of a model: arrays of nstage entries (a plain model has one type);
 *   ipddp_set_stage_types  stage_types[N-1] of a problem (checked: stage t must map onto the state size of knot t+1), to
 *                          be called before ipddp_set_inputs / ipddp_solve_queue when nstage > 1;
 *   ipddp_set_stage_compl  indices_compl of one stage type (ipddp_problem_create's list belongs to type 0);
 *   ipddp_stage_layout     per-knot sizes nx[N], nu[N], nc[N] (the terminal knot: nxt, 0, 0) next to ipddp_layout's offsets.
 * Array shapes of a chain: every per-stage array is strided by the MAXIMUM over the types (ipddp_model_dims returns the
 * maxima) and zero-padded: ubar [B][N-1][nu], lower/upper [B][nstage][nu] (one bound vector per stage type), x1 [B][nx],
 * states [B][N][ns] with ns = max(nx, nxn, nxt) (= nx for a plain model), controls [B][N-1][nu].  Per-instance horizons
 * are not available for chains. */
int ipddp_model_stages(const char* model, int* nstage, int* nx, int* nu, int* nc, int* nxn, int* nxt);
int ipddp_set_stage_types(ipddp_problem* h, const int* stage_types);
int ipddp_set_stage_compl(ipddp_problem* h, int stage_type, const int* indices_compl, int n_compl);
int ipddp_stage_layout(ipddp_problem* h, int* nx, int* nu, int* nc);

/* Replaces Solver(T, dynamics, objectives, constraints, bounds; options) (src/solver.jl:11-26) and the
 * workspace constructors behind it (src/data/*.jl) for a batch of B instances with up to N knots.
 * indices_compl: 0-based constraint indices that get `- mu` (src/data/methods.jl:27-29), may be NULL.
 * trace_capacity: rows of per-iteration trace kept per instance (0 = none).  device: CUDA ordinal.
 * Limits (an error is returned, nothing is truncated): B >= 1, N >= 2, nu + nc <= 64 per stage type, N bounded by the
 * shared memory of the merit kernels (the message states the bound); any number of states. */
int ipddp_problem_create(const char* model, int B, int N, const int* indices_compl, int n_compl,
                         const ipddp_options* opt, int device, int trace_capacity, ipddp_problem** out);
int ipddp_problem_destroy(ipddp_problem* h);
int ipddp_set_options(ipddp_problem* h, const ipddp_options* opt);
/* Execution tuning that never changes results (no reference counterpart).  h == NULL sets the default for problems
 * created afterwards.  Keys: "fw_spec_max" -- rounds with at most this many active instances run the forward pass with
 * one CTA per instance that tries 8 step sizes of the backtracking sequence at once (0 = never);
 * "bw_spec_max" -- rounds with at most this many active instances run the backward pass with one CTA per instance
 * that tries 4 values of the regularisation schedule at once (0 = never).  The default (-1 as the global default) is the
 * number of instances whose speculative CTAs are resident at the same time on the device (CTAs per SM x SMs);
 * "list_sort" -- 1 (default): the active lists are bucketed by expected work, heaviest instances first, so that a launch
 * does not end with its longest-running warps; 0: arrival order;
 * "bulk_slots" (global) -- ipddp_solve_many admits at most this many batches into their bulk rounds at the same time
 * (default 3), so that the low-occupancy tail of one batch overlaps the bulk rounds of the next. */
int ipddp_set_tuning(ipddp_problem* h, const char* key, int value);

/* Per-timestep offset tables of the instance records (doubles from the start of one instance's
 * block): traj_off[t], gain_off[t] for t = 0..N-1, and the strides.  Any pointer may be NULL. */
int ipddp_layout(ipddp_problem* h, long long* traj_off, long long* gain_off, long long* traj_stride,
                 long long* gain_stride, long long* tile_stride);

/* Inputs of solve!(solver, x1, controls) (src/solve.jl:1-4), batched; HOST pointers, copied H2D.
 *   x1 [B*nx], ubar [B*(N-1)*nu], params [B*np] (NULL if np==0), lower/upper [B*nu] (+-inf allowed;
 *   the reference's Bound, src/bounds.jl:1-26), horizons [B] knots per instance (NULL = all N). */
int ipddp_set_inputs(ipddp_problem* h, const double* x1, const double* ubar, const double* params,
                     const double* lower, const double* upper, const int* horizons);
/* Same, but the buffers are already resident in device memory (device pointers, D2D copy). */
int ipddp_set_inputs_device(ipddp_problem* h, const double* x1, const double* ubar, const double* params,
                            const double* lower, const double* upper, const int* horizons);

/* Replaces solve!(solver, x1, controls): initialize_trajectory! (src/solver.jl:54-105) + solve!(solver)
 * (src/solve.jl:6-93) for every instance.  Asynchronous phases run on the handle's stream; returns after
 * all instances terminated.  warm_start != 0 skips initialize_trajectory! (src/solve.jl:6 semantics). */
int ipddp_solve(ipddp_problem* h, int warm_start);

/* Streaming solve of a QUEUE of instances through the handle's B resident slots (B from ipddp_problem_create): every
 * lock-step round works on the instances resident at that moment; when an instance terminates, its results are written to
 * the output arrays at the instance's queue index and its slot is re-initialised with the next queued instance before the
 * following round.  The rounds stay full until the queue runs dry and a straggler only holds its own slot, so one handle
 * with memory for B instances sustains the throughput that otherwise needs several batches in flight (ipddp_solve_many).
 * Per instance this is solve!(solver, x1, controls) (reference src/solve.jl:1-93) exactly as in ipddp_solve -- the slot
 * an instance lands in never enters its computations.  Q may be smaller or (much) larger than B.
 *   inputs  [Q] instances laid out like ipddp_set_inputs' arrays; HOST pointers, or device pointers if inputs_on_device.
 *   outputs any pointer may be NULL; HOST arrays of Q (x [Q*N*nx], u [Q*(N-1)*nu], zeros beyond an instance's horizon),
 *           or device arrays if outputs_on_device.  Host copies (H2D of the inputs, D2H of the results) are issued on the
 *           handle's stream and are part of ipddp_stats.ms_total.
 * Not kept in this mode: per-iteration traces and duals (they live in the slots; use ipddp_solve). */
typedef struct ipddp_queue {
  int Q;
  const double* x1;        /* [Q*nx] */
  const double* ubar;      /* [Q*(N-1)*nu] */
  const double* params;    /* [Q*np], NULL if np == 0 */
  const double* lower;     /* [Q*nu] */
  const double* upper;     /* [Q*nu] */
  const int* horizons;     /* [Q] knots per instance, NULL = all N */
  int inputs_on_device;
  int *status, *k, *j, *l;                                       /* SolverData, reference src/data/solver.jl:8-33 */
  double *objective, *primal_inf, *dual_inf, *cs_inf, *mu, *reg_last, *step_size;
  int *n_backward, *n_sweeps, *n_kkt, *n_rollouts;               /* work counters (ipddp_get_counters) */
  double* x;               /* get_trajectory(solver): nominal states */
  double* u;               /* nominal controls */
  int outputs_on_device;
} ipddp_queue;
int ipddp_solve_queue(ipddp_problem* h, const ipddp_queue* queue);

/* Launch on the caller's CUDA stream (cudaStream_t as void*) instead of the handle's own; NULL restores the handle's
 * stream.  Synchronises the previous stream first. */
int ipddp_set_stream(ipddp_problem* h, void* cuda_stream);

/* Several handles at once: n independent problems (same device, e.g. different models) progress concurrently, each on
 * its own stream; one host thread polls completion events and immediately enqueues the next round of whichever problem
 * is ready, so the lock-step tail of one batch overlaps the bulk rounds of the others.  Handles that finish start another solve of
 * their inputs until total_solves (>= n) solves are complete.  elapsed_ms: device time from the first enqueue to the
 * last completion (CUDA events); agg: counters summed over all solves (per-kernel times are not split here). */
int ipddp_solve_many(ipddp_problem** problems, int n, int total_solves, int warm_start, double* elapsed_ms,
                     ipddp_stats* agg);

/* Phase-level entry points (for kernel parity tests and profiling), all instances:
 *   initialize       initialize_trajectory! + the prologue of solve! (src/solve.jl:14-38)
 *   eval_derivatives evaluate_derivatives!(problem) (src/derivatives.jl:31-35)
 *   backward_pass    backward_pass! + inertia_correction! (src/backward_pass.jl, src/inertia_correction.jl:257-276)
 *   check            dual/primal/cs errors, convergence test, barrier update (src/solve.jl:49-73,107-180)
 *   forward_pass     forward_pass! + update_nominal_trajectory! + filter update (src/forward_pass.jl, src/solve.jl:80-85)
 * n_forward (may be NULL) receives how many instances went through the forward pass in `check`'s verdict. */
int ipddp_initialize(ipddp_problem* h);
int ipddp_eval_derivatives(ipddp_problem* h);
int ipddp_backward_pass(ipddp_problem* h);
int ipddp_check(ipddp_problem* h, int* n_forward);
int ipddp_forward_pass(ipddp_problem* h);

/* SolverData fields per instance (src/data/solver.jl:8-33); any pointer may be NULL; HOST arrays of B. */
int ipddp_get_results(ipddp_problem* h, int* status, int* k, int* j, int* l, double* objective,
                      double* primal_inf, double* dual_inf, double* cs_inf, double* mu, double* reg_last,
                      double* step_size);
/* get_trajectory(solver) (src/solver.jl:46-48): nominal states x [B*N*nx], controls u [B*(N-1)*nu]. */
int ipddp_get_trajectory(ipddp_problem* h, double* x, double* u);
/* nominal duals: phi [B*(N-1)*nc], zl, zu [B*(N-1)*nu], lam [B*N*nx] */
int ipddp_get_duals(ipddp_problem* h, double* phi, double* zl, double* zu, double* lam);
/* per-instance work counters [B]: backward passes, sweeps, KKT steps, rollouts */
int ipddp_get_counters(ipddp_problem* h, int* n_backward, int* n_sweeps, int* n_kkt, int* n_rollouts);
/* named raw arrays for tests: "x","u","c","il","iu","phi","zl","zu" (nominal) and "cur_*" [B][N-1 or N][dim];
 * "lam" [B][N][nx]; "gains" [B][N-1][(K+2nu)(nx+1)]; "Qu" [B][N-1][nu]; "tile" [B][slots][N]; "tileN" [B][slotsN].
 * Returns the number of doubles (out == NULL: size query). */
long long ipddp_get_array(ipddp_problem* h, const char* name, double* out);
/* trace rows of instance b: cols k, j, objective, primal_inf, dual_inf, cs_inf, mu, reg_last, step_size, l,
 * theta, barrier_lagrangian -- one row per accepted iteration. */
int ipddp_get_trace(ipddp_problem* h, int b, double* rows, int* nrows);
int ipddp_get_stats(ipddp_problem* h, ipddp_stats* st);
/* the CUDA stream the handle launches on (cudaStream_t as void*), for external event timing */
void* ipddp_stream(ipddp_problem* h);

/* FP64 FMA throughput microbenchmark (TFLOP/s) used as the measured roofline denominator for the
 * backward pass; and an HBM copy probe (GB/s). */
double ipddp_measure_fp64_tflops(int device);
double ipddp_measure_hbm_gbs(int device);

/* Kernel-level test hooks (used by tests/ only).
 *   ipddp_test_detmath: elementwise device evaluation of the deterministic math layer (fn 0..5) and of the
 *       reciprocal-based exact division x / y used inside the LDLT (fn 6);
 *       fn 0 sin, 1 cos, 2 tan, 3 log, 4 exp, 5 pow(x[i], y[i]).
 *   ipddp_test_ldlt: one warp per matrix runs the device dsytf2_rook('U') / inertia / dsytrs_rook path on
 *       nmat dense column-major n x n matrices (upper triangle read) with 5 right-hand sides each.
 *       Outputs: Aout [nmat][n*n] (upper triangle = factors), ipiv [nmat][n] (LAPACK 1-based convention),
 *       info [nmat], npos [nmat] (positive eigenvalues of D, tol 1e-12), X [nmat][n*5] (solution). */
int ipddp_test_detmath(int fn, int n, const double* x, const double* y, double* out, int device);
int ipddp_test_ldlt(int n, int nmat, const double* A, const double* Bm, double* Aout, int* ipiv, int* info,
                    int* npos, double* X, int device);

#ifdef __cplusplus
}
#endif
#endif
