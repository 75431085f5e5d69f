// libipddp_b200.so -- C ABI (include/ipddp_b200.h): handle management, device memory, the lock-step
// batched solve loop (reference src/solve.jl:40-88 for every instance at once) and result access.
//
// Solve loop per round, over the list of still-active instances (active-set compaction = list of
// instance ids rebuilt by the kernels with atomics; finished instances cost nothing):
//   k_derivs   (instance x knot grid)        evaluate_derivatives!
//   k_backward (one warp per instance)       backward_pass! + inertia_correction!
//   k_check    (one warp per instance)       errors, convergence, barrier update -> forward list / next list
//   k_forward  (one warp per instance)       forward_pass! + accept -> next list
// then one 32-byte D2H of the list counters decides the next round's grid sizes.  ipddp_solve runs a fixed batch to
// completion; ipddp_solve_queue streams a queue of instances through the handle's slots (k_admit / k_retire between
// rounds) so that the rounds stay full; ipddp_solve_many overlaps several handles.
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <thread>
#include <vector>

#include "kernels_common.cuh"
#include "ldlt_warp.cuh"
#include "kernel_backward.cuh"
#include "kernel_forward.cuh"
#include "vtable.h"

extern "C" {
const ModelVTable* ipddp_vtable_cartpole();
const ModelVTable* ipddp_vtable_acrobot();
const ModelVTable* ipddp_vtable_concar();
const ModelVTable* ipddp_vtable_concar_quad();
const ModelVTable* ipddp_vtable_pushing();
const ModelVTable* ipddp_vtable_double_integrator();
const ModelVTable* ipddp_vtable_ragged();
}

namespace {

thread_local std::string g_err;
// defaults for new problems: -1 = as many instances as the speculative kernels hold in one wave on this device
// (ModelVTable::spec_caps: e.g. cartpole 148 / 740, acrobot 296 / 2 664), capped by the memory of their private pools
int g_fw_spec_max = -1;
int g_bw_spec_max = -1;
int g_bulk_slots = 3;      // ipddp_solve_many: batches admitted into their bulk rounds at the same time
int g_list_sort = 1;       // active lists bucketed by expected work, heaviest first (0 = arrival order, for A/B runs)
int fail(const std::string& m) { g_err = m; return -1; }
#define CK(call)                                                                                    \
  do {                                                                                              \
    cudaError_t e_ = (call);                                                                        \
    if (e_ != cudaSuccess) return fail(std::string(#call) + ": " + cudaGetErrorString(e_));         \
  } while (0)

std::vector<const ModelVTable*>& registry() {
  static std::vector<const ModelVTable*> r = {ipddp_vtable_cartpole(),    ipddp_vtable_acrobot(),
                                              ipddp_vtable_concar(),      ipddp_vtable_concar_quad(),
                                              ipddp_vtable_pushing(),     ipddp_vtable_double_integrator(),
                                              ipddp_vtable_ragged()};
  return r;
}
const ModelVTable* find(const char* name) {
  for (auto* m : registry())
    if (strcmp(m->name, name) == 0) return m;
  return nullptr;
}

// Field `f` (0 x, 1 u, 2 c, 3 il, 4 iu, 5 phi, 6 zl, 7 zu) of knot t's record of an instance whose knot has nx / nu / nc
// entries: offset inside the record and number of entries.
__device__ __forceinline__ void field_of(int f, int nx, int nu, int nc, int* off, int* dim) {
  const int oU = nx, oC = oU + nu, oIL = oC + nc, oIU = oIL + nu, oPHI = oIU + nu, oZL = oPHI + nc, oZU = oZL + nu;
  switch (f) {
    case 0: *off = 0; *dim = nx; break;
    case 1: *off = oU; *dim = nu; break;
    case 2: *off = oC; *dim = nc; break;
    case 3: *off = oIL; *dim = nu; break;
    case 4: *off = oIU; *dim = nu; break;
    case 5: *off = oPHI; *dim = nc; break;
    case 6: *off = oZL; *dim = nu; break;
    default: *off = oZU; *dim = nu; break;
  }
}
// sizes of knot t of instance b: its stage type's (nx, nu, nc); the terminal knot carries nxt states only
__device__ __forceinline__ void knot_dims(const DevView& v, int t, int Nb, int* nx, int* nu, int* nc) {
  if (t >= Nb - 1) { *nx = v.nxt; *nu = 0; *nc = 0; return; }
  const int k = v.type_of(t);
  *nx = v.snx[k]; *nu = v.snu[k]; *nc = v.snc[k];
}

// gather one field of the trajectory records into a dense [B][nst][dim] array (dim = the largest size of the field over
// the stage types; entries a knot does not have, and knots beyond an instance's horizon, are zero)
__global__ void k_gather(DevView v, int use_cur, int field, int dim, int nst, double* out) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)v.B * nst * dim;
  if (idx >= total) return;
  const int i = (int)(idx % dim);
  const int t = (int)((idx / dim) % nst);
  const int b = (int)(idx / ((long long)dim * nst));
  const int set = use_cur ? 1 - v.nomsel[b] : v.nomsel[b];
  const int Nb = v.horizon[b];
  double val = 0.0;
  if (t < Nb) {
    int nx, nu, nc, off, n;
    knot_dims(v, t, Nb, &nx, &nu, &nc);
    field_of(field, nx, nu, nc, &off, &n);
    if (i < n) val = v.rec(set, b, t)[off + i];
  }
  out[idx] = val;
}

__global__ void k_fp64_peak(double* out, int iters) {
  double a0 = threadIdx.x * 1e-9 + 1.0, a1 = a0 + 0.1, a2 = a0 + 0.2, a3 = a0 + 0.3, a4 = a0 + 0.4, a5 = a0 + 0.5,
         a6 = a0 + 0.6, a7 = a0 + 0.7;
  const double m = 1.0000001, c = 1e-9;
  for (int i = 0; i < iters; ++i) {
    a0 = IPDDP_FMA(a0, m, c); a1 = IPDDP_FMA(a1, m, c); a2 = IPDDP_FMA(a2, m, c); a3 = IPDDP_FMA(a3, m, c);
    a4 = IPDDP_FMA(a4, m, c); a5 = IPDDP_FMA(a5, m, c); a6 = IPDDP_FMA(a6, m, c); a7 = IPDDP_FMA(a7, m, c);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
}
__global__ void k_copy(const double4* __restrict__ a, double4* __restrict__ b, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    b[i] = a[i];
}

__global__ void k_test_detmath(int fn, int n, const double* x, const double* y, double* out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double r;
  switch (fn) {
    case 0: r = dm::sin(x[i]); break;
    case 1: r = dm::cos(x[i]); break;
    case 2: r = dm::tan(x[i]); break;
    case 3: r = dm::log(x[i]); break;
    case 4: r = dm::exp(x[i]); break;
    case 6: r = ipk::DivBy(y[i])(x[i]); break;     // x / y through the reciprocal-based exact division of the LDLT
    default: r = dm::pow(x[i], y[i]); break;
  }
  out[i] = r;
}

// one warp per matrix: dense -> packed smem, factor (+ inertia), solve 5 rhs, packed -> dense
template <int N>
__global__ void k_test_ldlt(int nmat, const double* A, const double* Bm, double* Aout, int* ipiv_out,
                            int* info_out, int* np_out, double* X) {
  IPDDP_DYN_SMEM(double, sm);
  const int lane = threadIdx.x, m = blockIdx.x;
  if (m >= nmat) return;
  constexpr int n = N, kp = N * (N + 1) / 2;
  double* lhs = sm;
  double* rhs = lhs + ((kp + 1) & ~1);
  double* ws = rhs + ((n * 6 + 1) & ~1);
  unsigned char* scratch = reinterpret_cast<unsigned char*>(ws + 4 * n);
  for (int j = 0; j < n; ++j)
    for (int i = lane; i <= j; i += 32) lhs[ipk::pk(i, j)] = A[(size_t)m * n * n + i + (size_t)j * n];
  typedef ipk::RhsL<N, 5> RL;
  for (int e = lane; e < RL::SIZE; e += 32) rhs[e] = 0.0;
  __syncwarp();
  for (int e = lane; e < n * 5; e += 32) rhs[RL::at(e % n, e / n)] = Bm[(size_t)m * n * 5 + e];
  __syncwarp();
  int np = 0;
  const int info = ipk::warp_ldlt_factor<N, 5>(lhs, rhs, ws, scratch, lane, 1e-12, np, ipk::ldlt_tri_lane(lane));
  if (info == 0) ipk::warp_ldlt_solve_forward<N, 5>(lhs, rhs, scratch, lane);
  __syncwarp();
  for (int j = 0; j < n; ++j)
    for (int i = lane; i <= j; i += 32) Aout[(size_t)m * n * n + i + (size_t)j * n] = lhs[ipk::pk(i, j)];
  for (int e = lane; e < n * 5; e += 32) X[(size_t)m * n * 5 + e] = rhs[RL::at(e % n, e / n)];
  for (int e = lane; e < n; e += 32) ipiv_out[(size_t)m * n + e] = ipk::LdltScratch<N>::ipiv_of(ipk::LdltScratch<N>::info(scratch)[e]);
  if (lane == 0) { info_out[m] = info; np_out[m] = np; }
}

// Queue mode: write the results of the instances whose slots are listed in done[0..n) to the queue's output arrays at
// the instance's queue index: SolverData scalars, work counters and the nominal trajectory (get_trajectory,
// reference src/solver.jl:46-48).  One CTA per slot.  Knots beyond the instance's horizon are written as zeros.
__global__ void k_retire(DevView v, QueueView q, const int* done, int n) {
  if ((int)blockIdx.x >= n) return;
  const int b = done[blockIdx.x];
  const size_t i = (size_t)v.inst_of[b];
  const size_t Q = (size_t)q.Q;
  if (threadIdx.x == 0) {
    const int sf[QSI_COUNT] = {SI_STATUS, SI_K, SI_J, SI_L, SI_NBACK, SI_NSWEEP, SI_NKKT, SI_NROLL};
    const int df[QSD_COUNT] = {SD_OBJECTIVE, SD_PRIMAL_INF, SD_DUAL_INF, SD_CS_INF, SD_MU, SD_REG_LAST, SD_STEP};
    for (int f = 0; f < QSI_COUNT; ++f) q.si[(size_t)f * Q + i] = v.siv(sf[f], b);
    for (int f = 0; f < QSD_COUNT; ++f) q.sd[(size_t)f * Q + i] = v.sdv(df[f], b);
  }
  const int Nb = v.horizon[b];
  const double* r0 = v.rec(v.nomsel[b], b, 0);
  if (q.x) {
    double* xo = q.x + i * (size_t)v.N * v.ns;
    for (int e = threadIdx.x; e < v.N * v.ns; e += blockDim.x) {
      const int t = e / v.ns, c = e - t * v.ns;
      int nx = 0, nu = 0, nc = 0;
      if (t < Nb) knot_dims(v, t, Nb, &nx, &nu, &nc);
      xo[e] = c < nx ? r0[(size_t)t * v.TR + c] : 0.0;
    }
  }
  if (q.u) {
    double* uo = q.u + i * (size_t)(v.N - 1) * v.nu;
    for (int e = threadIdx.x; e < (v.N - 1) * v.nu; e += blockDim.x) {
      const int t = e / v.nu, c = e - t * v.nu;
      int nx = 0, nu = 0, nc = 0;
      if (t < Nb - 1) knot_dims(v, t, Nb, &nx, &nu, &nc);
      uo[e] = c < nu ? r0[(size_t)t * v.TR + nx + c] : 0.0;
    }
  }
}

// horizons handed over in device memory cannot be validated on the host: count the out-of-range entries
__global__ void k_count_bad_horizons(const int* hz, int n, int N, int* bad) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && (hz[i] < 2 || hz[i] > N)) atomicAdd(bad, 1);
}

}  // namespace

struct ipddp_problem {
  const ModelVTable* vt = nullptr;
  DevView v;
  int device = 0;
  cudaStream_t stream = nullptr;      // the stream every launch of this handle goes to
  cudaStream_t own_stream = nullptr;  // created with the handle; `stream` points here unless ipddp_set_stream replaced it
  std::vector<void*> allocs;
  int* d_compl = nullptr;
  double *d_p = nullptr, *d_lower = nullptr, *d_upper = nullptr, *d_x1 = nullptr, *d_ubar = nullptr;
  int* d_horizon = nullptr;
  int* d_list[2] = {nullptr, nullptr};
  int* d_list_fwd = nullptr;
  int* d_done[2] = {nullptr, nullptr};   // queue mode: slots freed in the running / the previous round
  int* d_counters = nullptr;
  int* h_counters = nullptr;             // pinned
  int* h_si = nullptr;                   // pinned snapshot of the per-instance int scalars
  double* h_sd = nullptr;                // pinned snapshot of the per-instance double scalars
  // grow-only scratch: device staging for gathers and for the queue's inputs / outputs, pinned host staging
  void* d_stage = nullptr; size_t d_stage_cap = 0;
  void* d_qin = nullptr;   size_t d_qin_cap = 0;
  void* d_qout = nullptr;  size_t d_qout_cap = 0;
  void* h_stage = nullptr; size_t h_stage_cap = 0;
  int cur = 0, n_active = 0, hstate = 0;
  int nb[LIST_BUCKETS] = {0, 0, 0, 0};    // bucket sizes of d_list[cur] (their sum is n_active)
  ListView view(int which) const {
    ListView l;
    l.base = d_list[which]; l.stride = v.B;
    for (int k = 0; k < LIST_BUCKETS; ++k) l.n[k] = nb[k];
    return l;
  }
  // the bucket sizes of the next round's list as the kernels left them in the (host copy of the) counters
  void take_next_counts() {
    n_active = 0;
    for (int k = 0; k < LIST_BUCKETS; ++k) { nb[k] = h_counters[CNT_NEXT + k]; n_active += nb[k]; }
  }
  bool inputs_set = false;
  int spec_cap = 0;              // instances the speculative-forward record pool (DevView::spec_traj) was sized for
  int bw_spec_cap = 0;           // instances the speculative-backward output pool (DevView::spec_bw) was sized for
  size_t bw_pool_doubles() const { return (size_t)(v.N - 1) * (v.G + v.nu) + (size_t)v.N * v.ns; }
  unsigned char* d_stage_type = nullptr;   // stage chains: [N-1] stage type per running stage
  bool stages_set = false;
  cudaEvent_t ev[9] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  ipddp_stats st;

  template <class T> int alloc(T** p, size_t n) {
    void* q = nullptr;
    cudaError_t e = cudaMalloc(&q, (n > 0 ? n : 1) * sizeof(T));
    if (e != cudaSuccess) return fail(std::string("cudaMalloc: ") + cudaGetErrorString(e));
    allocs.push_back(q);
    *p = (T*)q;
    return 0;
  }
};

namespace {

#define NEED_HANDLE(h) do { if (!(h)) return fail("null problem handle"); } while (0)

int grow_device(void** p, size_t* cap, size_t need) {
  if (need <= *cap) return 0;
  if (*p) { cudaFree(*p); *p = nullptr; *cap = 0; }
  CK(cudaMalloc(p, need));
  *cap = need;
  return 0;
}
int grow_pinned(void** p, size_t* cap, size_t need) {
  if (need <= *cap) return 0;
  if (*p) { cudaFreeHost(*p); *p = nullptr; *cap = 0; }
  CK(cudaMallocHost(p, need));
  *cap = need;
  return 0;
}

int run_init(ipddp_problem* h, int warm) {
  CK(cudaMemsetAsync(h->d_counters, 0, CNT_COUNT * sizeof(int), h->stream));
  h->cur = 0;
  h->vt->init(h->v, warm, 0, h->v.B, h->d_list[0], h->d_counters, h->stream);
  h->st.launches += 1;
  CK(cudaMemcpyAsync(h->h_counters, h->d_counters, CNT_COUNT * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  CK(cudaGetLastError());
  h->take_next_counts();
  return 0;
}

// summed per-instance work counters of a snapshot of the SI block
void add_counters(ipddp_stats& st, const int* si, int B) {
  for (int b = 0; b < B; ++b) {
    st.sum_backward += si[(size_t)SI_NBACK * B + b];
    st.sum_sweeps += si[(size_t)SI_NSWEEP * B + b];
    st.sum_kkt += si[(size_t)SI_NKKT * B + b];
    st.sum_rollouts += si[(size_t)SI_NROLL * B + b];
    st.sum_deriv_stages += (long long)si[(size_t)SI_NDERIV * B + b];
    st.n_converged += (si[(size_t)SI_STATUS * B + b] == 0);
  }
}

}  // namespace

extern "C" {

int ipddp_abi_version(void) { return IPDDP_ABI_VERSION; }
const char* ipddp_last_error(void) { return g_err.c_str(); }

void ipddp_default_options(ipddp_options* o) {  // reference src/options.jl:1-38
  if (!o) return;
  o->quasi_newton = 0; o->optimality_tolerance = 1.0e-8; o->max_iterations = 1000; o->reset_cache = 1;
  o->verbose = 0; o->print_frequency = 10; o->mu_init = 1.0; o->ineq_dual_init = 1.0; o->kappa_1 = 0.01;
  o->kappa_2 = 0.01; o->reg_1 = 1e-4; o->reg_min = 1e-20; o->reg_max = 1e40; o->kappa_bar_w_p = 100.0;
  o->kappa_w_p = 8.0; o->kappa_w_m = 1.0 / 3.0; o->kappa_c = 0.25; o->delta_c = 1e-8; o->kappa_eps = 10.0;
  o->kappa_mu = 0.2; o->theta_mu = 1.2; o->tau_min = 0.99; o->s_max = 100.0; o->eta_L = 1e-4; o->s_L = 2.3;
  o->delta = 1.0; o->s_theta = 1.1; o->gamma_alpha = 0.05; o->gamma_theta = 1e-5; o->gamma_L = 1e-5;
  o->kappa_Sigma = 1e10;
}

int ipddp_num_models(void) { return (int)registry().size(); }
const char* ipddp_model_name(int i) { return (i >= 0 && i < (int)registry().size()) ? registry()[i]->name : nullptr; }
int ipddp_model_dims(const char* model, int* nx, int* nu, int* nc, int* np, int* tile_slots) {
  const ModelVTable* m = model ? find(model) : nullptr;
  if (!m) return fail(std::string("unknown model ") + (model ? model : "(null)"));
  if (nx) *nx = m->nx; if (nu) *nu = m->nu; if (nc) *nc = m->nc; if (np) *np = m->np;
  if (tile_slots) *tile_slots = m->d_nslot;
  return 0;
}
int ipddp_model_load(const char* path) {
  if (!path) return fail("null plugin path");
  void* so = dlopen(path, RTLD_NOW | RTLD_LOCAL);
  if (!so) return fail(std::string("dlopen: ") + dlerror());
  typedef const ModelVTable* (*fn_t)();
  fn_t f = (fn_t)dlsym(so, "ipddp_plugin_vtable");
  if (!f) return fail("plugin does not export ipddp_plugin_vtable");
  const ModelVTable* vt = f();
  if (!vt || !vt->name || !vt->init || !vt->derivs || !vt->backward || !vt->check || !vt->forward || !vt->admit ||
      !vt->prepare || !vt->spec_caps)
    return fail("plugin returned an incomplete model table");
  for (auto*& m : registry())
    if (strcmp(m->name, vt->name) == 0) { m = vt; return 0; }
  registry().push_back(vt);
  return 0;
}

int ipddp_problem_create(const char* model, int B, int N, const int* indices_compl, int n_compl,
                         const ipddp_options* opt, int device, int trace_capacity, ipddp_problem** out) {
  if (!out) return fail("null output pointer");
  const ModelVTable* vt = model ? find(model) : nullptr;
  if (!vt) return fail(std::string("unknown model ") + (model ? model : "(null)"));
  if (B < 1 || N < 2) return fail("need B >= 1 and N >= 2");
  if (vt->nu + vt->nc > 64) return fail("KKT dimension nu+nc > 64 not supported");
  CK(cudaSetDevice(device));
  int optin = 0;
  CK(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
  // the merit kernels keep four per-knot arrays per warp in shared memory: the horizon is bounded by the opt-in limit
  if ((long long)vt->smem_merit(N) > (long long)optin)
    return fail("N too large: the per-knot scratch of k_forward / k_check (" + std::to_string(vt->smem_merit(N)) +
                " B) exceeds the device's shared memory per block (" + std::to_string(optin) + " B)");
  if (vt->prepare(optin) != 0) return fail("cudaFuncSetAttribute failed");
  ipddp_problem* h = new ipddp_problem();
  memset(&h->st, 0, sizeof(h->st));
  h->vt = vt;
  h->device = device;
  DevView& v = h->v;
  memset(&v, 0, sizeof(v));
  v.B = B; v.N = N; v.nx = vt->nx; v.nu = vt->nu; v.nc = vt->nc; v.np = vt->np;
  v.nstage = vt->nstage; v.nxt = vt->nxt; v.ns = vt->ns;
  for (int k = 0; k < MAX_STAGE_TYPES; ++k) { v.snx[k] = vt->snx[k]; v.snu[k] = vt->snu[k]; v.snc[k] = vt->snc[k]; }
  v.TR = (vt->nx + 5 * vt->nu + 2 * vt->nc + 1) & ~1;                  // strides padded to an even number of doubles
  v.G = ((vt->nu + vt->nc + 2 * vt->nu) * (vt->nx + 1) + 1) & ~1;
  if (opt) v.opt = *opt; else ipddp_default_options(&v.opt);
  v.trace_cap = trace_capacity > 0 ? trace_capacity : 0;
  {
    int cap_bw = 592, cap_fw = 148;
    vt->spec_caps(N, &cap_bw, &cap_fw);
    // private pools: at most ~3 GB each
    const double bw_one = 3.0 * ((double)(N - 1) * (v.G + v.nu) + (double)N * v.ns) * 8.0, fw_one = 8.0 * N * v.TR * 8.0;
    const int lim_bw = (int)(3.0e9 / bw_one), lim_fw = (int)(3.0e9 / fw_one);
    if (cap_bw > lim_bw) cap_bw = lim_bw;
    if (cap_fw > lim_fw) cap_fw = lim_fw;
    v.fw_spec_max = g_fw_spec_max >= 0 ? g_fw_spec_max : cap_fw;
    v.bw_spec_max = g_bw_spec_max >= 0 ? g_bw_spec_max : cap_bw;
  }
  v.list_sort = g_list_sort;
  if ((long long)vt->smem_merit_spec(N) > (long long)optin) v.fw_spec_max = 0;   // long horizons: bulk line search only
  v.n_compl = (indices_compl && n_compl > 0) ? n_compl : 0;
  for (int q = 0; q < v.n_compl; ++q) {
    if (indices_compl[q] < 0 || indices_compl[q] >= vt->snc[0]) { delete h; return fail("indices_compl out of range"); }
    v.compl_mask[0] |= 1ull << indices_compl[q];     // stage type 0; the other types of a chain: ipddp_set_stage_compl
  }
  const size_t np1 = vt->np > 0 ? vt->np : 1;
  int rc = 0;
  rc |= h->alloc(&h->d_compl, (size_t)v.n_compl);
  rc |= h->alloc(&h->d_p, (size_t)B * np1);
  rc |= h->alloc(&h->d_lower, (size_t)B * vt->nstage * vt->nu);
  rc |= h->alloc(&h->d_upper, (size_t)B * vt->nstage * vt->nu);
  if (vt->nstage > 1) rc |= h->alloc(&h->d_stage_type, (size_t)(N - 1));
  rc |= h->alloc(&h->d_x1, (size_t)B * vt->nx);
  rc |= h->alloc(&h->d_ubar, (size_t)B * (N - 1) * vt->nu);
  rc |= h->alloc(&h->d_horizon, (size_t)B);
  rc |= h->alloc(&v.traj, (size_t)2 * B * N * v.TR);
  rc |= h->alloc(&v.nomsel, (size_t)B);
  rc |= h->alloc(&v.inst_of, (size_t)B);
  rc |= h->alloc(&v.lam, (size_t)B * N * vt->ns);
  rc |= h->alloc(&v.tile, (size_t)B * (vt->d_nslot > 0 ? vt->d_nslot : 1) * N);
  rc |= h->alloc(&v.tileN, (size_t)B * (vt->dn_nslot > 0 ? vt->dn_nslot : 1));
  rc |= h->alloc(&v.gains, (size_t)B * (N - 1) * v.G);
  rc |= h->alloc(&v.Qu, (size_t)B * (N - 1) * vt->nu);
  rc |= h->alloc(&v.sd, (size_t)SD_COUNT * B);
  rc |= h->alloc(&v.si, (size_t)SI_COUNT * B);
  rc |= h->alloc(&v.filter, (size_t)2 * IPDDP_FILTER_CAPACITY * B);
  rc |= h->alloc(&v.trace, (size_t)B * v.trace_cap * IPDDP_TRACE_COLS);
  if (v.fw_spec_max > B) v.fw_spec_max = B;
  rc |= h->alloc(&v.spec_traj, (size_t)(v.fw_spec_max > 0 ? v.fw_spec_max : 0) * ipk::FWS_WARPS * N * v.TR);
  h->spec_cap = v.fw_spec_max;
  if (v.bw_spec_max > B) v.bw_spec_max = B;
  rc |= h->alloc(&v.spec_bw, (size_t)(v.bw_spec_max > 0 ? v.bw_spec_max : 0) * (ipk::BWS_WARPS - 1) * h->bw_pool_doubles());
  h->bw_spec_cap = v.bw_spec_max;
  rc |= h->alloc(&h->d_list[0], (size_t)LIST_BUCKETS * B);
  rc |= h->alloc(&h->d_list[1], (size_t)LIST_BUCKETS * B);
  rc |= h->alloc(&h->d_list_fwd, (size_t)LIST_BUCKETS * B);
  rc |= h->alloc(&h->d_done[0], (size_t)B);
  rc |= h->alloc(&h->d_done[1], (size_t)B);
  rc |= h->alloc(&h->d_counters, (size_t)CNT_COUNT);
  if (rc != 0) { ipddp_problem_destroy(h); return -1; }
  v.compl_idx = h->d_compl; v.p = h->d_p; v.lower = h->d_lower; v.upper = h->d_upper; v.x1 = h->d_x1;
  v.ubar = h->d_ubar; v.horizon = h->d_horizon;
  v.stage_type = h->d_stage_type;      // NULL for a plain model
  h->stages_set = vt->nstage == 1;
  auto finish = [&]() -> int {   // anything failing from here on releases the half-built handle
    if (v.n_compl) CK(cudaMemcpy(h->d_compl, indices_compl, v.n_compl * sizeof(int), cudaMemcpyHostToDevice));
    CK(cudaMemset(v.traj, 0, (size_t)2 * B * N * v.TR * sizeof(double)));
    CK(cudaMemset(v.si, 0, (size_t)SI_COUNT * B * sizeof(int)));
    CK(cudaMemset(v.sd, 0, (size_t)SD_COUNT * B * sizeof(double)));
    CK(cudaMemset(v.nomsel, 0, (size_t)B * sizeof(int)));
    CK(cudaMemset(v.inst_of, 0, (size_t)B * sizeof(int)));
    CK(cudaStreamCreate(&h->own_stream));
    h->stream = h->own_stream;
    CK(cudaMallocHost((void**)&h->h_counters, CNT_COUNT * sizeof(int)));
    CK(cudaMallocHost((void**)&h->h_si, (size_t)SI_COUNT * B * sizeof(int)));
    CK(cudaMallocHost((void**)&h->h_sd, (size_t)SD_COUNT * B * sizeof(double)));
    for (int i = 0; i < 9; ++i) CK(cudaEventCreate(&h->ev[i]));
    return 0;
  };
  if (finish() != 0) { ipddp_problem_destroy(h); return -1; }
  *out = h;
  return 0;
}

int ipddp_problem_destroy(ipddp_problem* h) {
  if (!h) return 0;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  for (void* p : h->allocs) cudaFree(p);
  if (h->d_stage) cudaFree(h->d_stage);
  if (h->d_qin) cudaFree(h->d_qin);
  if (h->d_qout) cudaFree(h->d_qout);
  if (h->h_stage) cudaFreeHost(h->h_stage);
  if (h->h_counters) cudaFreeHost(h->h_counters);
  if (h->h_si) cudaFreeHost(h->h_si);
  if (h->h_sd) cudaFreeHost(h->h_sd);
  for (int i = 0; i < 9; ++i)
    if (h->ev[i]) cudaEventDestroy(h->ev[i]);
  if (h->own_stream) cudaStreamDestroy(h->own_stream);
  delete h;
  return 0;
}

int ipddp_set_options(ipddp_problem* h, const ipddp_options* opt) {
  NEED_HANDLE(h);
  if (!opt) return fail("null options");
  h->v.opt = *opt;
  return 0;
}

int ipddp_set_stream(ipddp_problem* h, void* cuda_stream) {
  NEED_HANDLE(h);
  CK(cudaSetDevice(h->device));
  CK(cudaStreamSynchronize(h->stream));
  h->stream = cuda_stream ? (cudaStream_t)cuda_stream : h->own_stream;
  return 0;
}

int ipddp_model_stages(const char* model, int* nstage, int* nx, int* nu, int* nc, int* nxn, int* nxt) {
  const ModelVTable* m = model ? find(model) : nullptr;
  if (!m) return fail(std::string("unknown model ") + (model ? model : "(null)"));
  if (nstage) *nstage = m->nstage;
  if (nxt) *nxt = m->nxt;
  for (int k = 0; k < m->nstage; ++k) {
    if (nx) nx[k] = m->snx[k];
    if (nu) nu[k] = m->snu[k];
    if (nc) nc[k] = m->snc[k];
    if (nxn) nxn[k] = m->snxn[k];
  }
  return 0;
}

int ipddp_set_stage_types(ipddp_problem* h, const int* stage_types) {
  NEED_HANDLE(h);
  if (!stage_types) return fail("null stage type table");
  const DevView& v = h->v;
  const ModelVTable* vt = h->vt;
  for (int t = 0; t < v.N - 1; ++t)
    if (stage_types[t] < 0 || stage_types[t] >= vt->nstage) return fail("stage type out of range");
  for (int t = 0; t < v.N - 1; ++t) {
    const int k = stage_types[t];
    const int next_nx = t + 1 < v.N - 1 ? vt->snx[stage_types[t + 1]] : vt->nxt;
    if (vt->snxn[k] != next_nx)
      return fail("stage " + std::to_string(t) + " maps to " + std::to_string(vt->snxn[k]) + " states but the next knot has " +
                  std::to_string(next_nx));
  }
  if (vt->nstage == 1) return 0;
  CK(cudaSetDevice(h->device));
  std::vector<unsigned char> tmp(v.N - 1);
  for (int t = 0; t < v.N - 1; ++t) tmp[t] = (unsigned char)stage_types[t];
  CK(cudaMemcpy(h->d_stage_type, tmp.data(), tmp.size(), cudaMemcpyHostToDevice));
  h->stages_set = true;
  return 0;
}

int ipddp_set_stage_compl(ipddp_problem* h, int stage_type, const int* indices_compl, int n_compl) {
  NEED_HANDLE(h);
  if (stage_type < 0 || stage_type >= h->vt->nstage) return fail("stage type out of range");
  unsigned long long m = 0ull;
  for (int q = 0; q < n_compl; ++q) {
    if (!indices_compl || indices_compl[q] < 0 || indices_compl[q] >= h->vt->snc[stage_type]) return fail("indices_compl out of range");
    m |= 1ull << indices_compl[q];
  }
  h->v.compl_mask[stage_type] = m;
  return 0;
}

int ipddp_stage_layout(ipddp_problem* h, int* nx, int* nu, int* nc) {
  NEED_HANDLE(h);
  const DevView& v = h->v;
  std::vector<unsigned char> types(v.N > 1 ? v.N - 1 : 1, 0);
  if (v.nstage > 1) {
    if (!h->stages_set) return fail("stage chain: ipddp_set_stage_types not called");
    CK(cudaSetDevice(h->device));
    CK(cudaMemcpy(types.data(), h->d_stage_type, (size_t)(v.N - 1), cudaMemcpyDeviceToHost));
  }
  for (int t = 0; t < v.N; ++t) {
    const bool term = t == v.N - 1;
    if (nx) nx[t] = term ? v.nxt : v.snx[types[t]];
    if (nu) nu[t] = term ? 0 : v.snu[types[t]];
    if (nc) nc[t] = term ? 0 : v.snc[types[t]];
  }
  return 0;
}

int ipddp_set_tuning(ipddp_problem* h, const char* key, int value) {
  const std::string k = key ? key : "";
  if (k == "fw_spec_max") {   // h == NULL: default for problems created afterwards
    if (!h) { g_fw_spec_max = value < 0 ? -1 : value; return 0; }     // -1: automatic (one wave of speculative CTAs)
    if (value < 0) value = 0;
    if (value > h->v.B) value = h->v.B;
    int optin = 0;
    CK(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, h->device));
    if ((long long)h->vt->smem_merit_spec(h->v.N) > (long long)optin) value = 0;
    if (value > h->spec_cap) {   // grow the private trial-record pool
      CK(cudaSetDevice(h->device));
      double* q = nullptr;
      CK(cudaMalloc((void**)&q, (size_t)value * ipk::FWS_WARPS * h->v.N * h->v.TR * sizeof(double)));
      h->allocs.push_back(q);    // the old pool is released with the handle
      h->v.spec_traj = q;
      h->spec_cap = value;
    }
    h->v.fw_spec_max = value;
    return 0;
  }
  if (k == "list_sort") {
    if (h) h->v.list_sort = value != 0; else g_list_sort = value != 0;
    return 0;
  }
  if (k == "bulk_slots") {
    g_bulk_slots = value < 1 ? 1 : value;
    return 0;
  }
  if (k == "bw_spec_max") {
    if (!h) { g_bw_spec_max = value < 0 ? -1 : value; return 0; }
    if (value < 0) value = 0;
    if (value > h->v.B) value = h->v.B;
    if (value > h->bw_spec_cap) {
      CK(cudaSetDevice(h->device));
      double* q = nullptr;
      CK(cudaMalloc((void**)&q, (size_t)value * (ipk::BWS_WARPS - 1) * h->bw_pool_doubles() * sizeof(double)));
      h->allocs.push_back(q);
      h->v.spec_bw = q;
      h->bw_spec_cap = value;
    }
    h->v.bw_spec_max = value;
    return 0;
  }
  return fail("unknown tuning key " + k);
}

int ipddp_layout(ipddp_problem* h, long long* traj_off, long long* gain_off, long long* traj_stride,
                 long long* gain_stride, long long* tile_stride) {
  NEED_HANDLE(h);
  const DevView& v = h->v;
  for (int t = 0; t < v.N; ++t) {
    if (traj_off) traj_off[t] = (long long)t * v.TR;
    if (gain_off) gain_off[t] = (long long)t * v.G;
  }
  if (traj_stride) *traj_stride = (long long)v.N * v.TR;
  if (gain_stride) *gain_stride = (long long)(v.N - 1) * v.G;
  if (tile_stride) *tile_stride = (long long)h->vt->d_nslot * v.N;
  return 0;
}

static int set_inputs_impl(ipddp_problem* h, const double* x1, const double* ubar, const double* params,
                           const double* lower, const double* upper, const int* horizons, cudaMemcpyKind kind) {
  NEED_HANDLE(h);
  const DevView& v = h->v;
  CK(cudaSetDevice(h->device));
  if (!x1 || !ubar || !lower || !upper) return fail("x1, ubar, lower, upper are required");
  if (v.np > 0 && !params) return fail("params required (np > 0)");
  if (!h->stages_set) return fail("stage chain: ipddp_set_stage_types not called");
  if (v.nstage > 1 && horizons) return fail("stage chains solve the full horizon: per-instance horizons are not supported");
  cudaStream_t s = h->stream;
  if (horizons && kind == cudaMemcpyHostToDevice)
    for (int b = 0; b < v.B; ++b)
      if (horizons[b] < 2 || horizons[b] > v.N) return fail("horizon out of range [2, N]");
  CK(cudaMemcpyAsync(h->d_x1, x1, (size_t)v.B * v.nx * sizeof(double), kind, s));
  CK(cudaMemcpyAsync(h->d_ubar, ubar, (size_t)v.B * (v.N - 1) * v.nu * sizeof(double), kind, s));
  if (v.np > 0) CK(cudaMemcpyAsync(h->d_p, params, (size_t)v.B * v.np * sizeof(double), kind, s));
  else CK(cudaMemsetAsync(h->d_p, 0, (size_t)v.B * sizeof(double), s));
  CK(cudaMemcpyAsync(h->d_lower, lower, (size_t)v.B * v.nstage * v.nu * sizeof(double), kind, s));
  CK(cudaMemcpyAsync(h->d_upper, upper, (size_t)v.B * v.nstage * v.nu * sizeof(double), kind, s));
  bool check_device_horizons = false;
  if (horizons) {
    CK(cudaMemcpyAsync(h->d_horizon, horizons, (size_t)v.B * sizeof(int), kind, s));
    check_device_horizons = (kind != cudaMemcpyHostToDevice);
  } else {
    if (grow_pinned(&h->h_stage, &h->h_stage_cap, (size_t)v.B * sizeof(int)) != 0) return -1;
    int* hz = (int*)h->h_stage;
    for (int b = 0; b < v.B; ++b) hz[b] = v.N;
    CK(cudaMemcpyAsync(h->d_horizon, hz, (size_t)v.B * sizeof(int), cudaMemcpyHostToDevice, s));
  }
  if (check_device_horizons) {   // a horizon outside [2, N] would index out of bounds in every kernel
    CK(cudaMemsetAsync(h->d_counters, 0, CNT_COUNT * sizeof(int), s));
    IPDDP_LAUNCH(k_count_bad_horizons, (v.B + 255) / 256, 256, 0, s, h->d_horizon, v.B, v.N, h->d_counters + CNT_BAD);
    CK(cudaMemcpyAsync(h->h_counters, h->d_counters, CNT_COUNT * sizeof(int), cudaMemcpyDeviceToHost, s));
  }
  CK(cudaStreamSynchronize(s));
  if (check_device_horizons && h->h_counters[CNT_BAD] != 0) { h->inputs_set = false; return fail("horizon out of range [2, N]"); }
  h->inputs_set = true;
  return 0;
}

int ipddp_set_inputs(ipddp_problem* h, const double* x1, const double* ubar, const double* params,
                     const double* lower, const double* upper, const int* horizons) {
  return set_inputs_impl(h, x1, ubar, params, lower, upper, horizons, cudaMemcpyHostToDevice);
}
int ipddp_set_inputs_device(ipddp_problem* h, const double* x1, const double* ubar, const double* params,
                            const double* lower, const double* upper, const int* horizons) {
  return set_inputs_impl(h, x1, ubar, params, lower, upper, horizons, cudaMemcpyDeviceToDevice);
}

int ipddp_initialize(ipddp_problem* h) {
  NEED_HANDLE(h);
  if (!h->inputs_set) return fail("ipddp_set_inputs not called");
  CK(cudaSetDevice(h->device));
  return run_init(h, 0);
}
int ipddp_eval_derivatives(ipddp_problem* h) {
  NEED_HANDLE(h);
  CK(cudaSetDevice(h->device));
  h->vt->derivs(h->v, h->view(h->cur), h->stream);
  CK(cudaStreamSynchronize(h->stream));
  CK(cudaGetLastError());
  return 0;
}
int ipddp_backward_pass(ipddp_problem* h) {
  NEED_HANDLE(h);
  CK(cudaSetDevice(h->device));
  h->vt->backward(h->v, h->view(h->cur), h->stream);
  CK(cudaStreamSynchronize(h->stream));
  CK(cudaGetLastError());
  return 0;
}
int ipddp_check(ipddp_problem* h, int* n_forward) {
  NEED_HANDLE(h);
  CK(cudaSetDevice(h->device));
  CK(cudaMemsetAsync(h->d_counters, 0, CNT_COUNT * sizeof(int), h->stream));
  h->vt->check(h->v, h->view(h->cur), h->d_list[1 - h->cur], h->d_list_fwd, h->d_counters, h->stream);
  CK(cudaMemcpyAsync(h->h_counters, h->d_counters, CNT_COUNT * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  CK(cudaGetLastError());
  if (n_forward) {
    *n_forward = 0;
    for (int k = 0; k < LIST_BUCKETS; ++k) *n_forward += h->h_counters[CNT_FWD + k];
  }
  return 0;
}
int ipddp_forward_pass(ipddp_problem* h) {
  NEED_HANDLE(h);
  CK(cudaSetDevice(h->device));
  h->vt->forward(h->v, h->d_list_fwd, h->n_active, h->d_list[1 - h->cur], h->d_counters, h->stream);
  CK(cudaMemcpyAsync(h->h_counters, h->d_counters, CNT_COUNT * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  CK(cudaGetLastError());
  h->take_next_counts();
  h->cur = 1 - h->cur;
  return 0;
}

// One lock-step round on the handle's stream over the n instances of list[cur]: the four phase kernels, timed with CUDA
// events (ev[0..4]) on the launching stream; ends with the asynchronous D2H copy of the list counters.
static int enqueue_round(ipddp_problem* h, const DevView& v, const ListView& list, int cur, bool clear_counters) {
  cudaStream_t s = h->stream;
  cudaEvent_t* ev = h->ev;
  const int n = list.total();
  CK(cudaEventRecord(ev[0], s));
  h->vt->derivs(v, list, s);
  CK(cudaEventRecord(ev[1], s));
  h->vt->backward(v, list, s);
  CK(cudaEventRecord(ev[2], s));
  if (clear_counters) CK(cudaMemsetAsync(h->d_counters, 0, CNT_COUNT * sizeof(int), s));
  h->vt->check(v, list, h->d_list[1 - cur], h->d_list_fwd, h->d_counters, s);
  CK(cudaEventRecord(ev[3], s));
  h->vt->forward(v, h->d_list_fwd, n, h->d_list[1 - cur], h->d_counters, s);
  CK(cudaEventRecord(ev[4], s));
  CK(cudaMemcpyAsync(h->h_counters, h->d_counters, CNT_COUNT * sizeof(int), cudaMemcpyDeviceToHost, s));
  h->st.launches += 4;
  h->st.iterations += 1;
  h->st.n_active_rounds += n;
  h->st.sum_active_sq += (double)n * (double)n;
  return 0;
}
static int add_round_times(ipddp_problem* h) {
  float ms = 0.f;
  cudaEvent_t* ev = h->ev;
  CK(cudaEventElapsedTime(&ms, ev[0], ev[1])); h->st.ms_derivs += ms;
  CK(cudaEventElapsedTime(&ms, ev[1], ev[2])); h->st.ms_backward += ms;
  CK(cudaEventElapsedTime(&ms, ev[2], ev[3])); h->st.ms_check += ms;
  CK(cudaEventElapsedTime(&ms, ev[3], ev[4])); h->st.ms_forward += ms;
  return 0;
}

int ipddp_solve(ipddp_problem* h, int warm_start) {
  NEED_HANDLE(h);
  if (!h->inputs_set) return fail("ipddp_set_inputs not called");
  CK(cudaSetDevice(h->device));
  const DevView& v = h->v;
  ipddp_stats& st = h->st;
  memset(&st, 0, sizeof(st));
  cudaStream_t s = h->stream;
  cudaEvent_t* ev = h->ev;
  CK(cudaEventRecord(ev[6], s));
  CK(cudaEventRecord(ev[0], s));
  if (run_init(h, warm_start) != 0) return -1;
  CK(cudaEventRecord(ev[1], s));
  CK(cudaEventSynchronize(ev[1]));
  float ms = 0.f;
  CK(cudaEventElapsedTime(&ms, ev[0], ev[1]));
  st.ms_init = ms;
  while (h->n_active > 0) {
    if (enqueue_round(h, v, h->view(h->cur), h->cur, true) != 0) return -1;
    CK(cudaStreamSynchronize(s));
    CK(cudaGetLastError());
    if (add_round_times(h) != 0) return -1;
    h->take_next_counts();
    h->cur = 1 - h->cur;
  }
  CK(cudaMemcpyAsync(h->h_si, v.si, (size_t)SI_COUNT * v.B * sizeof(int), cudaMemcpyDeviceToHost, s));
  CK(cudaEventRecord(ev[7], s));
  CK(cudaEventSynchronize(ev[7]));
  CK(cudaEventElapsedTime(&ms, ev[6], ev[7]));
  st.ms_total = ms;
  add_counters(st, h->h_si, v.B);
  return 0;
}

// Streaming solve: the Q queued instances flow through the handle's B resident slots.  Every lock-step round works
// on the instances resident at that moment; a slot whose instance terminated in round r is written out (k_retire) and
// re-initialised with the next queued instance (k_admit) before round r+1, so the rounds stay full until the queue is
// empty and a straggler instance only ever holds its own slot.  Per instance this is exactly ipddp_solve's sequence
// (reference src/solve.jl:40-90): which slot an instance lands in does not enter any of its computations.
int ipddp_solve_queue(ipddp_problem* h, const ipddp_queue* io) {
  NEED_HANDLE(h);
  if (!io) return fail("null queue description");
  const DevView& hv = h->v;
  const int Q = io->Q, B = hv.B, N = hv.N;
  if (Q < 1) return fail("need Q >= 1 queued instances");
  if (!io->x1 || !io->ubar || !io->lower || !io->upper) return fail("x1, ubar, lower, upper are required");
  if (hv.np > 0 && !io->params) return fail("params required (np > 0)");
  if (hv.trace_cap > 0) return fail("per-iteration traces are kept per slot: use ipddp_solve for traced solves");
  CK(cudaSetDevice(h->device));
  cudaStream_t s = h->stream;
  ipddp_stats& st = h->st;
  memset(&st, 0, sizeof(st));
  const size_t np1 = hv.np > 0 ? hv.np : 1;
  const size_t n_x1 = (size_t)Q * hv.nx, n_ub = (size_t)Q * (N - 1) * hv.nu, n_p = (size_t)Q * np1,
               n_b = (size_t)Q * hv.nstage * hv.nu;
  const size_t n_xo = (size_t)Q * N * hv.ns, n_uo = (size_t)Q * (N - 1) * hv.nu;
  if (!h->stages_set) return fail("stage chain: ipddp_set_stage_types not called");
  if (hv.nstage > 1 && io->horizons) return fail("stage chains solve the full horizon: per-instance horizons are not supported");
  CK(cudaEventRecord(h->ev[6], s));
  QueueView q;
  memset(&q, 0, sizeof(q));
  q.Q = Q;
  // ---- inputs: device pointers are used in place, host pointers are staged (one H2D copy each, on the stream)
  if (io->inputs_on_device) {
    q.x1 = io->x1; q.ubar = io->ubar; q.p = io->params; q.lower = io->lower; q.upper = io->upper; q.horizon = io->horizons;
  } else {
    if (io->horizons)
      for (int i = 0; i < Q; ++i)
        if (io->horizons[i] < 2 || io->horizons[i] > N) return fail("horizon out of range [2, N]");
    const size_t bytes = (n_x1 + n_ub + n_p + 2 * n_b) * sizeof(double) + (size_t)Q * sizeof(int);
    if (grow_device(&h->d_qin, &h->d_qin_cap, bytes) != 0) return -1;
    double* d = (double*)h->d_qin;
    double* dx1 = d; double* dub = dx1 + n_x1; double* dp = dub + n_ub; double* dlo = dp + n_p; double* dup = dlo + n_b;
    int* dhz = (int*)(dup + n_b);
    CK(cudaMemcpyAsync(dx1, io->x1, n_x1 * sizeof(double), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(dub, io->ubar, n_ub * sizeof(double), cudaMemcpyHostToDevice, s));
    if (hv.np > 0) CK(cudaMemcpyAsync(dp, io->params, n_p * sizeof(double), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(dlo, io->lower, n_b * sizeof(double), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(dup, io->upper, n_b * sizeof(double), cudaMemcpyHostToDevice, s));
    if (io->horizons) CK(cudaMemcpyAsync(dhz, io->horizons, (size_t)Q * sizeof(int), cudaMemcpyHostToDevice, s));
    q.x1 = dx1; q.ubar = dub; q.p = dp; q.lower = dlo; q.upper = dup; q.horizon = io->horizons ? dhz : nullptr;
  }
  // ---- outputs: scalars always go through the handle's buffers (the statistics need them on the host); trajectories
  //      are written in place when the caller's arrays live on the device
  const bool want_x = io->x != nullptr, want_u = io->u != nullptr;
  const bool dev_out = io->outputs_on_device != 0;
  {
    size_t bytes = (size_t)QSD_COUNT * Q * sizeof(double) + (size_t)QSI_COUNT * Q * sizeof(int);
    if (!dev_out) bytes += ((want_x ? n_xo : 0) + (want_u ? n_uo : 0)) * sizeof(double);
    if (grow_device(&h->d_qout, &h->d_qout_cap, bytes) != 0) return -1;
    double* d = (double*)h->d_qout;
    q.sd = d; d += (size_t)QSD_COUNT * Q;
    if (!dev_out) {
      if (want_x) { q.x = d; d += n_xo; }
      if (want_u) { q.u = d; d += n_uo; }
    } else {
      q.x = io->x; q.u = io->u;
    }
    q.si = (int*)d;
  }
  // ---- the rounds
  int next_inst = 0, retired = 0, n_active = 0, cur = 0, n_free = B, r = 0;
  struct LogFile {   // per-round series for profiling; closed on every return path
    FILE* f;
    ~LogFile() { if (f) fclose(f); }
  } qlog_guard = {getenv("IPDDP_QUEUE_LOG") ? fopen(getenv("IPDDP_QUEUE_LOG"), "a") : nullptr};
  FILE* qlog = qlog_guard.f;
  for (int k = 0; k < LIST_BUCKETS; ++k) h->nb[k] = 0;
  const int light = LIST_BUCKETS - 1;   // fresh instances join the bucket of the one-sweep instances
  const bool runs = hv.opt.max_iterations > 0;   // otherwise every instance terminates inside k_admit (status 8)
  while (retired < Q) {
    DevView v = hv;
    v.done_list = h->d_done[r & 1];
    CK(cudaMemsetAsync(h->d_counters, 0, CNT_COUNT * sizeof(int), s));
    const int n_admit = n_free < Q - next_inst ? n_free : Q - next_inst;
    CK(cudaEventRecord(h->ev[5], s));     // ev[5] .. ev[0]: admission (k_admit) of this round, retirement of the previous one
    if (n_admit > 0) {
      h->vt->admit(v, q, r == 0 ? nullptr : h->d_done[(r - 1) & 1], n_admit, next_inst,
                   h->d_list[cur] + (size_t)light * B + h->nb[light], h->d_counters, s);
      st.launches += 1;
      next_inst += n_admit;
      if (runs) { n_active += n_admit; h->nb[light] += n_admit; }
    }
    if (n_active > 0) {
      if (enqueue_round(h, v, h->view(cur), cur, false) != 0) { cudaStreamSynchronize(s); return -1; }
    } else {
      CK(cudaMemcpyAsync(h->h_counters, h->d_counters, CNT_COUNT * sizeof(int), cudaMemcpyDeviceToHost, s));
    }
    CK(cudaStreamSynchronize(s));
    CK(cudaGetLastError());
    if (h->h_counters[CNT_BAD] != 0) return fail("horizon out of range [2, N] in the queue");
    if (n_active > 0) {
      const double b0 = st.ms_backward, f0 = st.ms_forward;
      if (add_round_times(h) != 0) return -1;
      float ms_a = 0.f;
      CK(cudaEventElapsedTime(&ms_a, h->ev[5], h->ev[0]));
      st.ms_init += ms_a;
      if (qlog) fprintf(qlog, "{\"round\": %d, \"active\": %d, \"buckets\": [%d, %d, %d, %d], \"admitted\": %d, \"admit_ms\": %.3f, "
                        "\"bw_ms\": %.3f, \"fw_ms\": %.3f, \"done\": %d}\n", r, n_active, h->nb[0], h->nb[1], h->nb[2], h->nb[3],
                        n_admit, ms_a, st.ms_backward - b0, st.ms_forward - f0, h->h_counters[CNT_DONE]);
    }
    const int n_done = h->h_counters[CNT_DONE];
    if (n_done > 0) {
      IPDDP_LAUNCH(k_retire, n_done, 128, 0, s, v, q, h->d_done[r & 1], n_done);
      st.launches += 1;
    }
    if (n_done == 0 && n_admit == 0 && n_active == 0) return fail("queue stalled");   // cannot happen
    retired += n_done;
    n_free = n_done;
    if (n_active > 0) { h->take_next_counts(); n_active = h->n_active; }
    cur = 1 - cur;
    r += 1;
  }
  if (qlog) fflush(qlog);
  // ---- results to the caller
  {
    const size_t sd_b = (size_t)QSD_COUNT * Q * sizeof(double), si_b = (size_t)QSI_COUNT * Q * sizeof(int);
    if (grow_pinned(&h->h_stage, &h->h_stage_cap, sd_b + si_b) != 0) return -1;
    double* hsd = (double*)h->h_stage;
    int* hsi = (int*)((char*)h->h_stage + sd_b);
    CK(cudaMemcpyAsync(hsd, q.sd, sd_b, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(hsi, q.si, si_b, cudaMemcpyDeviceToHost, s));
    int* io_i[QSI_COUNT] = {io->status, io->k, io->j, io->l, io->n_backward, io->n_sweeps, io->n_kkt, io->n_rollouts};
    double* io_d[QSD_COUNT] = {io->objective, io->primal_inf, io->dual_inf, io->cs_inf, io->mu, io->reg_last, io->step_size};
    if (dev_out) {
      for (int f = 0; f < QSI_COUNT; ++f)
        if (io_i[f]) CK(cudaMemcpyAsync(io_i[f], q.si + (size_t)f * Q, (size_t)Q * sizeof(int), cudaMemcpyDeviceToDevice, s));
      for (int f = 0; f < QSD_COUNT; ++f)
        if (io_d[f]) CK(cudaMemcpyAsync(io_d[f], q.sd + (size_t)f * Q, (size_t)Q * sizeof(double), cudaMemcpyDeviceToDevice, s));
    } else {
      if (want_x) CK(cudaMemcpyAsync(io->x, q.x, n_xo * sizeof(double), cudaMemcpyDeviceToHost, s));
      if (want_u) CK(cudaMemcpyAsync(io->u, q.u, n_uo * sizeof(double), cudaMemcpyDeviceToHost, s));
    }
    CK(cudaEventRecord(h->ev[7], s));
    CK(cudaEventSynchronize(h->ev[7]));
    CK(cudaGetLastError());
    if (!dev_out) {
      for (int f = 0; f < QSI_COUNT; ++f)
        if (io_i[f]) memcpy(io_i[f], hsi + (size_t)f * Q, (size_t)Q * sizeof(int));
      for (int f = 0; f < QSD_COUNT; ++f)
        if (io_d[f]) memcpy(io_d[f], hsd + (size_t)f * Q, (size_t)Q * sizeof(double));
    }
    for (int i = 0; i < Q; ++i) {
      st.sum_backward += hsi[(size_t)QSI_NBACK * Q + i];
      st.sum_sweeps += hsi[(size_t)QSI_NSWEEP * Q + i];
      st.sum_kkt += hsi[(size_t)QSI_NKKT * Q + i];
      st.sum_rollouts += hsi[(size_t)QSI_NROLL * Q + i];
      st.n_converged += (hsi[(size_t)QSI_STATUS * Q + i] == 0);
    }
    st.sum_deriv_stages = st.sum_backward;   // one derivative evaluation per backward pass
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, h->ev[6], h->ev[7]));
    st.ms_total = ms;
  }
  h->inputs_set = false;   // the slots hold whatever instances came last: ipddp_solve needs fresh inputs
  return 0;
}

// Several independent problems progress concurrently, each on its own stream, driven by one host thread that polls
// completion events and immediately enqueues the next round of whichever problem became ready.  This hides the
// lock-step tail of one batch (a few straggler instances at <1 % occupancy) behind the bulk rounds of the others.
// total_solves >= n: problems that finish early start another solve of their inputs until total_solves are done.
// (One problem with more queued instances than slots: ipddp_solve_queue does the same with one handle's memory.)
int ipddp_solve_many(ipddp_problem** hs, int n, int total_solves, int warm_start, double* elapsed_ms,
                     ipddp_stats* agg) {
  if (!hs || n < 1 || total_solves < n) return fail("need n >= 1 handles and total_solves >= n");
  for (int i = 0; i < n; ++i) {
    if (!hs[i]) return fail("null problem handle");
    if (!hs[i]->inputs_set) return fail("ipddp_set_inputs not called on every handle");
    if (hs[i]->device != hs[0]->device) return fail("all handles must live on one device");
  }
  CK(cudaSetDevice(hs[0]->device));
  ipddp_stats tot;
  memset(&tot, 0, sizeof(tot));
  enum { H_WAITING = 0, H_INIT = 1, H_ROUND = 2, H_STATS = 3, H_FINISHED = 4 };
  int started = 0, finished = 0;
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(hs[0]->ev[6], hs[0]->stream));
  // any error below leaves other handles' streams busy: drain the device before handing the error to the caller
  auto bail = [&]() -> int { cudaDeviceSynchronize(); return -1; };
  auto enqueue_init = [&](ipddp_problem* h) -> int {
    memset(&h->st, 0, sizeof(h->st));
    CK(cudaMemsetAsync(h->d_counters, 0, CNT_COUNT * sizeof(int), h->stream));
    h->cur = 0;
    h->vt->init(h->v, warm_start, 0, h->v.B, h->d_list[0], h->d_counters, h->stream);
    h->st.launches += 1;
    CK(cudaMemcpyAsync(h->h_counters, h->d_counters, CNT_COUNT * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaEventRecord(h->ev[8], h->stream));
    h->hstate = H_INIT;
    return 0;
  };
  // Speculative restarts spend 3 of 4 warps on sweeps that are usually discarded: worth it when the GPU is otherwise
  // idle (the tail of a lone batch), not while another problem still runs bulk rounds that could use those SM slots.
  auto bulk_elsewhere = [&](const ipddp_problem* self) -> bool {
    for (int i = 0; i < n; ++i)
      if (hs[i] != self && (hs[i]->hstate == H_INIT || (hs[i]->hstate == H_ROUND && hs[i]->n_active > 4096))) return true;
    return false;
  };
  // Admission: batches are started staggered, at most g_bulk_slots of them in their bulk rounds at a time.  Batches
  // that start together run their low-occupancy middle and tail rounds together, which is exactly what overlapping
  // batches is meant to avoid; a batch is admitted when an earlier one has dropped below B/8 active instances.
  for (int i = 0; i < n; ++i) hs[i]->hstate = H_WAITING;
  std::vector<int> runs(n, 0);
  auto in_bulk = [&](ipddp_problem* h) -> bool {
    if (h->hstate == H_INIT) return true;
    return h->hstate == H_ROUND && (long long)h->n_active * 8 > h->v.B;
  };
  auto admit = [&]() -> int {
    int bulk = 0;
    for (int i = 0; i < n; ++i) bulk += in_bulk(hs[i]) ? 1 : 0;
    while (bulk < g_bulk_slots && started < total_solves) {
      int pick = -1;     // the waiting handle that has run least often in this call (every handle runs at least once)
      for (int i = 0; i < n; ++i)
        if (hs[i]->hstate == H_WAITING && (pick < 0 || runs[i] < runs[pick])) pick = i;
      if (pick < 0) break;
      if (enqueue_init(hs[pick]) != 0) return -1;
      runs[pick]++;
      started++;
      bulk++;
    }
    return 0;
  };
  while (finished < total_solves) {
    bool progressed = false;
    if (admit() != 0) return bail();
    for (int i = 0; i < n; ++i) {
      ipddp_problem* h = hs[i];
      if (h->hstate == H_FINISHED || h->hstate == H_WAITING) continue;
      cudaError_t qe = cudaEventQuery(h->ev[8]);
      if (qe == cudaErrorNotReady) continue;
      if (qe != cudaSuccess) { fail(std::string("cudaEventQuery: ") + cudaGetErrorString(qe)); return bail(); }
      progressed = true;
      if (h->hstate == H_STATS) {
        tot.iterations += h->st.iterations;
        tot.launches += h->st.launches;
        tot.n_active_rounds += h->st.n_active_rounds;
        tot.sum_active_sq += h->st.sum_active_sq;
        add_counters(tot, h->h_si, h->v.B);
        finished++;
        if (cudaEventRecord(h->ev[7], h->stream) != cudaSuccess) { fail("cudaEventRecord"); return bail(); }
        h->hstate = (started < total_solves) ? H_WAITING : H_FINISHED;
        continue;
      }
      if (h->hstate == H_ROUND) h->cur = 1 - h->cur;
      h->take_next_counts();
      if (h->n_active > 0) {
        DevView v = h->v;
        if (h->n_active > 32 && h->n_active <= v.bw_spec_max && bulk_elsewhere(h)) v.bw_spec_max = 32;
        if (enqueue_round(h, v, h->view(h->cur), h->cur, true) != 0) return bail();
        if (cudaEventRecord(h->ev[8], h->stream) != cudaSuccess) { fail("cudaEventRecord"); return bail(); }
        h->hstate = H_ROUND;
      } else {
        // the stream is idle (its last event was observed): snapshot the per-instance scalars
        if (cudaMemcpyAsync(h->h_si, h->v.si, (size_t)SI_COUNT * h->v.B * sizeof(int), cudaMemcpyDeviceToHost, h->stream) != cudaSuccess ||
            cudaEventRecord(h->ev[8], h->stream) != cudaSuccess) { fail("snapshot of the instance scalars failed"); return bail(); }
        h->hstate = H_STATS;
      }
    }
    if (!progressed) std::this_thread::yield();
  }
  float best = 0.f;
  for (int i = 0; i < n; ++i) {
    CK(cudaEventSynchronize(hs[i]->ev[7]));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, hs[0]->ev[6], hs[i]->ev[7]));
    if (ms > best) best = ms;
  }
  CK(cudaGetLastError());
  tot.ms_total = best;
  if (elapsed_ms) *elapsed_ms = best;
  if (agg) *agg = tot;
  return 0;
}

// SolverData scalars: one D2H copy of each scalar block into pinned memory, then plain host copies
int ipddp_get_results(ipddp_problem* h, int* status, int* k, int* j, int* l, double* objective,
                      double* primal_inf, double* dual_inf, double* cs_inf, double* mu, double* reg_last,
                      double* step_size) {
  NEED_HANDLE(h);
  const DevView& v = h->v;
  const size_t B = (size_t)v.B;
  CK(cudaSetDevice(h->device));
  CK(cudaMemcpyAsync(h->h_si, v.si, (size_t)SI_COUNT * B * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaMemcpyAsync(h->h_sd, v.sd, (size_t)SD_COUNT * B * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  auto gi = [&](int f, int* dst) { if (dst) memcpy(dst, h->h_si + (size_t)f * B, B * sizeof(int)); };
  auto gd = [&](int f, double* dst) { if (dst) memcpy(dst, h->h_sd + (size_t)f * B, B * sizeof(double)); };
  gi(SI_STATUS, status); gi(SI_K, k); gi(SI_J, j); gi(SI_L, l);
  gd(SD_OBJECTIVE, objective); gd(SD_PRIMAL_INF, primal_inf); gd(SD_DUAL_INF, dual_inf);
  gd(SD_CS_INF, cs_inf); gd(SD_MU, mu); gd(SD_REG_LAST, reg_last); gd(SD_STEP, step_size);
  return 0;
}

static int gather(ipddp_problem* h, int use_cur, int field, int dim, int nst, double* out_host) {
  const DevView& v = h->v;
  const long long total = (long long)v.B * nst * dim;
  if (total == 0) return 0;
  if (grow_device(&h->d_stage, &h->d_stage_cap, (size_t)total * sizeof(double)) != 0) return -1;
  double* d = (double*)h->d_stage;
  const int th = 256;
  IPDDP_LAUNCH(k_gather, (unsigned)((total + th - 1) / th), th, 0, h->stream, v, use_cur, field, dim, nst, d);
  CK(cudaMemcpyAsync(out_host, d, total * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

int ipddp_get_trajectory(ipddp_problem* h, double* x, double* u) {
  NEED_HANDLE(h);
  const DevView& v = h->v;
  CK(cudaSetDevice(h->device));
  if (x && gather(h, 0, 0, v.ns, v.N, x) != 0) return -1;
  if (u && gather(h, 0, 1, v.nu, v.N - 1, u) != 0) return -1;
  return 0;
}

int ipddp_get_duals(ipddp_problem* h, double* phi, double* zl, double* zu, double* lam) {
  NEED_HANDLE(h);
  const DevView& v = h->v;
  CK(cudaSetDevice(h->device));
  if (phi && gather(h, 0, 5, v.nc, v.N - 1, phi) != 0) return -1;
  if (zl && gather(h, 0, 6, v.nu, v.N - 1, zl) != 0) return -1;
  if (zu && gather(h, 0, 7, v.nu, v.N - 1, zu) != 0) return -1;
  if (lam) {
    CK(cudaMemcpy(lam, v.lam, (size_t)v.B * v.N * v.ns * sizeof(double), cudaMemcpyDeviceToHost));
    // The costate lives in its own buffer here and keeps the values of the last sweep.  In the reference it is part of
    // the nominal dual set: update_nominal_trajectory! overwrites it with the (never written) current costate, i.e. zeros,
    // after every accepted step (src/data/methods.jl:89), and the next sweep rewrites it.  A solve that ends with an
    // accepted step -- max_iterations reached, status 8 -- therefore reports zeros.
    std::vector<int> status(v.B);
    CK(cudaMemcpy(status.data(), v.si + (size_t)SI_STATUS * v.B, (size_t)v.B * sizeof(int), cudaMemcpyDeviceToHost));
    for (int b = 0; b < v.B; ++b)
      if (status[b] == 8) memset(lam + (size_t)b * v.N * v.ns, 0, (size_t)v.N * v.ns * sizeof(double));
  }
  return 0;
}

int ipddp_get_counters(ipddp_problem* h, int* n_backward, int* n_sweeps, int* n_kkt, int* n_rollouts) {
  NEED_HANDLE(h);
  const DevView& v = h->v;
  const size_t B = (size_t)v.B;
  CK(cudaSetDevice(h->device));
  CK(cudaMemcpyAsync(h->h_si, v.si, (size_t)SI_COUNT * B * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  int* dst[4] = {n_backward, n_sweeps, n_kkt, n_rollouts};
  const int f[4] = {SI_NBACK, SI_NSWEEP, SI_NKKT, SI_NROLL};
  for (int q = 0; q < 4; ++q)
    if (dst[q]) memcpy(dst[q], h->h_si + (size_t)f[q] * B, B * sizeof(int));
  return 0;
}

long long ipddp_get_array(ipddp_problem* h, const char* name, double* out) {
  NEED_HANDLE(h);
  if (!name) return fail("null array name");
  const DevView& v = h->v;
  if (cudaSetDevice(h->device) != cudaSuccess) return fail("cudaSetDevice");
  std::string nm = name;
  int use_cur = 0;
  if (nm.rfind("cur_", 0) == 0) { use_cur = 1; nm = nm.substr(4); }
  struct F { const char* n; int field, dim, nst; } fields[] = {
      {"x", 0, v.ns, v.N}, {"u", 1, v.nu, v.N - 1}, {"c", 2, v.nc, v.N - 1}, {"il", 3, v.nu, v.N - 1},
      {"iu", 4, v.nu, v.N - 1}, {"phi", 5, v.nc, v.N - 1}, {"zl", 6, v.nu, v.N - 1}, {"zu", 7, v.nu, v.N - 1}};
  for (auto& f : fields)
    if (nm == f.n) {
      const long long n = (long long)v.B * f.nst * f.dim;
      if (out && gather(h, use_cur, f.field, f.dim, f.nst, out) != 0) return -1;
      return n;
    }
  const double* src = nullptr;
  long long n = 0;
  if (nm == "lam") { src = v.lam; n = (long long)v.B * v.N * v.ns; }
  else if (nm == "gains") {      // without the padding double of the device records
    const long long G = (long long)(v.nu + v.nc + 2 * v.nu) * (v.nx + 1), rows = (long long)v.B * (v.N - 1);
    if (out && rows > 0) {
      std::vector<double> tmp((size_t)rows * v.G);
      cudaError_t e = cudaMemcpy(tmp.data(), v.gains, tmp.size() * sizeof(double), cudaMemcpyDeviceToHost);
      if (e != cudaSuccess) return fail(cudaGetErrorString(e));
      for (long long r = 0; r < rows; ++r) memcpy(out + r * G, tmp.data() + r * v.G, (size_t)G * sizeof(double));
    }
    return rows * G;
  }
  else if (nm == "Qu") { src = v.Qu; n = (long long)v.B * (v.N - 1) * v.nu; }
  else if (nm == "tile") { src = v.tile; n = (long long)v.B * h->vt->d_nslot * v.N; }
  else if (nm == "tileN") { src = v.tileN; n = (long long)v.B * h->vt->dn_nslot; }
  else if (nm == "sd") { src = v.sd; n = (long long)SD_COUNT * v.B; }
  else return fail("unknown array " + nm);
  if (out && n > 0) {
    cudaError_t e = cudaMemcpy(out, src, n * sizeof(double), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) return fail(cudaGetErrorString(e));
  }
  return n;
}

int ipddp_get_trace(ipddp_problem* h, int b, double* rows, int* nrows) {
  NEED_HANDLE(h);
  const DevView& v = h->v;
  CK(cudaSetDevice(h->device));
  if (b < 0 || b >= v.B) return fail("instance out of range");
  int n = 0;
  CK(cudaMemcpy(&n, v.si + (size_t)SI_TRACE_N * v.B + b, sizeof(int), cudaMemcpyDeviceToHost));
  if (nrows) *nrows = n;
  if (rows && n > 0)
    CK(cudaMemcpy(rows, v.trace + (size_t)b * v.trace_cap * IPDDP_TRACE_COLS,
                  (size_t)n * IPDDP_TRACE_COLS * sizeof(double), cudaMemcpyDeviceToHost));
  return 0;
}

int ipddp_get_stats(ipddp_problem* h, ipddp_stats* st) {
  NEED_HANDLE(h);
  if (!st) return fail("null stats pointer");
  *st = h->st;
  return 0;
}
void* ipddp_stream(ipddp_problem* h) { return h ? (void*)h->stream : nullptr; }

double ipddp_measure_fp64_tflops(int device) {
  if (cudaSetDevice(device) != cudaSuccess) return -1.0;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return -1.0;
  const int blocks = prop.multiProcessorCount * 8, th = 256, iters = 20000;
  double* d = nullptr;
  if (cudaMalloc((void**)&d, (size_t)blocks * th * sizeof(double)) != cudaSuccess) return -1.0;
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  IPDDP_LAUNCH(k_fp64_peak, blocks, th, 0, 0, d, 1000);
  double best = 0.0;
  for (int r = 0; r < 5; ++r) {
    cudaEventRecord(a, 0);
    IPDDP_LAUNCH(k_fp64_peak, blocks, th, 0, 0, d, iters);
    cudaEventRecord(b, 0);
    cudaEventSynchronize(b);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, a, b);
    const double tf = 2.0 * 8.0 * (double)iters * blocks * th / (ms * 1e-3) / 1e12;
    if (tf > best) best = tf;
  }
  cudaEventDestroy(a); cudaEventDestroy(b);
  cudaFree(d);
  return best;
}

double ipddp_measure_hbm_gbs(int device) {
  if (cudaSetDevice(device) != cudaSuccess) return -1.0;
  const long long n4 = (1ll << 30) / 32;   // 1 GiB per buffer
  double4 *x = nullptr, *y = nullptr;
  if (cudaMalloc((void**)&x, n4 * 32) != cudaSuccess) return -1.0;
  if (cudaMalloc((void**)&y, n4 * 32) != cudaSuccess) { cudaFree(x); return -1.0; }
  cudaMemset(x, 0, n4 * 32);
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  double best = 0.0;
  for (int r = 0; r < 6; ++r) {
    cudaEventRecord(a, 0);
    IPDDP_LAUNCH(k_copy, 148 * 16, 512, 0, 0, x, y, n4);
    cudaEventRecord(b, 0);
    cudaEventSynchronize(b);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, a, b);
    const double gbs = 2.0 * n4 * 32 / (ms * 1e-3) / 1e9;
    if (r > 0 && gbs > best) best = gbs;
  }
  cudaEventDestroy(a); cudaEventDestroy(b);
  cudaFree(x); cudaFree(y);
  return best;
}

int ipddp_test_detmath(int fn, int n, const double* x, const double* y, double* out, int device) {
  CK(cudaSetDevice(device));
  double *dx = nullptr, *dy = nullptr, *dout = nullptr;
  CK(cudaMalloc((void**)&dx, (size_t)n * 8)); CK(cudaMalloc((void**)&dy, (size_t)n * 8)); CK(cudaMalloc((void**)&dout, (size_t)n * 8));
  CK(cudaMemcpy(dx, x, (size_t)n * 8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dy, y, (size_t)n * 8, cudaMemcpyHostToDevice));
  IPDDP_LAUNCH(k_test_detmath, (n + 255) / 256, 256, 0, 0, fn, n, dx, dy, dout);
  CK(cudaMemcpy(out, dout, (size_t)n * 8, cudaMemcpyDeviceToHost));
  CK(cudaGetLastError());
  cudaFree(dx); cudaFree(dy); cudaFree(dout);
  return 0;
}

int ipddp_test_ldlt(int n, int nmat, const double* A, const double* Bm, double* Aout, int* ipiv, int* info,
                    int* npos, double* X, int device) {
  if (n < 1 || n > 64) return fail("n out of range");
  CK(cudaSetDevice(device));
  const size_t sa = (size_t)nmat * n * n * 8, sb = (size_t)nmat * n * 5 * 8;
  double *dA = nullptr, *dB = nullptr, *dAo = nullptr, *dX = nullptr;
  int *dip = nullptr, *dinfo = nullptr, *dnp = nullptr;
  CK(cudaMalloc((void**)&dA, sa)); CK(cudaMalloc((void**)&dAo, sa)); CK(cudaMalloc((void**)&dB, sb)); CK(cudaMalloc((void**)&dX, sb));
  CK(cudaMalloc((void**)&dip, (size_t)nmat * n * 4)); CK(cudaMalloc((void**)&dinfo, (size_t)nmat * 4)); CK(cudaMalloc((void**)&dnp, (size_t)nmat * 4));
  CK(cudaMemcpy(dA, A, sa, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, Bm, sb, cudaMemcpyHostToDevice));
  CK(cudaMemset(dAo, 0, sa));
  const int kp = n * (n + 1) / 2;
  const size_t smem = ((size_t)(kp + 2 + n * 6 + 2 + 4 * n) * 8 + ((size_t)n * 21 + 15) / 16 * 16 + 31) / 16 * 16;
#define IPDDP_LDLT_CASE(NN) case NN: IPDDP_LAUNCH((k_test_ldlt<NN>), nmat, 32, smem, 0, nmat, dA, dB, dAo, dip, dinfo, dnp, dX); break;
  switch (n) {
    IPDDP_LDLT_CASE(1) IPDDP_LDLT_CASE(2) IPDDP_LDLT_CASE(3) IPDDP_LDLT_CASE(4) IPDDP_LDLT_CASE(5) IPDDP_LDLT_CASE(8)
    IPDDP_LDLT_CASE(14) IPDDP_LDLT_CASE(15) IPDDP_LDLT_CASE(16) IPDDP_LDLT_CASE(17) IPDDP_LDLT_CASE(31) IPDDP_LDLT_CASE(32)
    IPDDP_LDLT_CASE(33) IPDDP_LDLT_CASE(35) IPDDP_LDLT_CASE(48) IPDDP_LDLT_CASE(64)
    default: return fail("ipddp_test_ldlt: n must be one of 1,2,3,4,5,8,14,15,16,17,31,32,33,35,48,64");
  }
#undef IPDDP_LDLT_CASE
  CK(cudaMemcpy(Aout, dAo, sa, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(X, dX, sb, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(ipiv, dip, (size_t)nmat * n * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(info, dinfo, (size_t)nmat * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(npos, dnp, (size_t)nmat * 4, cudaMemcpyDeviceToHost));
  CK(cudaGetLastError());
  cudaFree(dA); cudaFree(dB); cudaFree(dAo); cudaFree(dX); cudaFree(dip); cudaFree(dinfo); cudaFree(dnp);
  return 0;
}

}  // extern "C"
