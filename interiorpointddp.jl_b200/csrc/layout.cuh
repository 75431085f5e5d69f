// Device data layout of one batched problem (all FP64 unless noted).
//
// Instance-major records ("array of instance records"): every kernel addresses instance b through
// b * stride, so active-set compaction is a pure index indirection (list of instance ids) and never
// moves data.  Inside an instance:
//   traj   [2 sets][B][N][TR]   one record per knot: x | u | c | il | iu | phi | zl | zu  (TR = nx+5nu+2nc)
//                               set `nomsel[b]` is the nominal trajectory, the other one the trial
//                               ("current") trajectory; accepting a step flips nomsel (no copy;
//                               reference copies 9 arrays, src/data/methods.jl:78-91)
//   lam    [B][N][nx]           costate of the last sweep
//   tile   [B][D_NSLOT][N]      compact derivative tile, knot index fastest: the derivative kernel
//                               (one thread per (instance, knot)) writes 32 consecutive knots per warp
//                               store = coalesced; the backward kernel reads a knot's slots with one
//                               32-byte sector per slot covering 4 consecutive knots
//   tileN  [B][DN_NSLOT]        terminal-stage derivative slots
//   gains  [B][N-1][G]          eq block K x (nx+1) col-major ([alpha beta; psi omega]) followed by the
//                               ineq block 2nu x (nx+1) ([chil zetal; chiu zetau]) -- the reference's
//                               `eq`/`ineq` matrices (src/data/update_rule.jl:71-84); G = (K+2nu)(nx+1)
//   Qu     [B][N-1][nu]
//   sd/si  [field][B]           SolverData scalars, structure of arrays
//   filter [2][FCAP][B]
#pragma once
#include "../../include/ipddp_b200.h"
#include "simt_compat.cuh"

enum SD {  // double scalars
  SD_MU = 0, SD_REG_LAST, SD_OBJECTIVE, SD_PRIMAL_INF, SD_DUAL_INF, SD_CS_INF, SD_L_CURR, SD_THETA_CURR,
  SD_L_NEXT, SD_THETA_NEXT, SD_THETA_MAX, SD_THETA_MIN, SD_STEP, SD_DUAL_NUM, SD_COUNT
};
enum SI {  // int scalars
  SI_STATUS = 0, SI_K, SI_J, SI_L, SI_FILTER_N, SI_SWITCHING, SI_ARMIJO, SI_DONE, SI_NBACK, SI_NSWEEP, SI_NKKT,
  SI_NROLL, SI_NDERIV, SI_TRACE_N, SI_LASTSW, SI_LASTROLL, SI_COUNT
};
// Active lists are kept in LIST_BUCKETS buckets, heaviest expected work first (bucket 0), so that a launch hands its
// longest-running instances to the first wave of warps instead of the last (a launch lasts as long as its slowest warp):
// the next-round list by the sweeps of the instance's last backward pass, the forward list by the trial steps (rollouts) of
// its last line search.  list[k * B + i] = i-th instance of bucket k; counters[CNT_NEXT + k] / [CNT_FWD + k] = its size.
constexpr int LIST_BUCKETS = 4;
enum CNT { CNT_NEXT = 0, CNT_FWD = 4, CNT_DONE = 8, CNT_BAD = 9, CNT_COUNT = 16 };
struct ListView {      // kernel argument: a bucketed list with its (host-known) bucket sizes
  const int* base;
  int stride;          // B
  int n[LIST_BUCKETS];
  IPDDP_BOTH int total() const { return n[0] + n[1] + n[2] + n[3]; }
  IPDDP_D int at(int i) const {
    if (i < n[0]) return base[i];
    i -= n[0];
    if (i < n[1]) return base[stride + i];
    i -= n[1];
    if (i < n[2]) return base[2 * stride + i];
    return base[3 * stride + (i - n[2])];
  }
};

// Queue mode (ipddp_solve_queue): Q queued instances flow through the B resident slots of a handle.  Inputs are read
// from the queue arrays when an instance is admitted into a slot, results are written to the queue's output arrays
// when it retires (k_admit / k_retire).
enum QSI { QSI_STATUS = 0, QSI_K, QSI_J, QSI_L, QSI_NBACK, QSI_NSWEEP, QSI_NKKT, QSI_NROLL, QSI_COUNT };
enum QSD { QSD_OBJECTIVE = 0, QSD_PRIMAL_INF, QSD_DUAL_INF, QSD_CS_INF, QSD_MU, QSD_REG_LAST, QSD_STEP, QSD_COUNT };
struct QueueView {
  int Q;
  const double *x1, *ubar, *p, *lower, *upper;   // [Q][...] like the batch inputs
  const int* horizon;                            // [Q] or NULL (= N)
  int* si;                                       // [QSI_COUNT][Q]
  double* sd;                                    // [QSD_COUNT][Q]
  double* x;                                     // [Q][N][nx] or NULL
  double* u;                                     // [Q][N-1][nu] or NULL
};

constexpr int MAX_STAGE_TYPES = 4;

struct DevView {
  int B, N;
  int nx, nu, nc, np;        // of a plain model; of a stage chain: the maxima over its stage types (they size every stride)
  // Stage chains (reference src/data/problem.jl:44-62: state / control sizes may change along the horizon): running stage t
  // is of type stage_type[t] (NULL: one type), with its own sizes snx/snu/snc; the terminal knot carries nxt states.
  // Storage stays uniform -- every record, gains block, Qu, lambda and tile column is strided by the maxima -- while the
  // arithmetic of a knot is instantiated for its stage type (ipk::for_stage).
  int nstage, nxt, ns;       // ns = stride of state-sized arrays (lambda, x outputs) = the largest state of any knot
  int snx[MAX_STAGE_TYPES], snu[MAX_STAGE_TYPES], snc[MAX_STAGE_TYPES];
  const unsigned char* stage_type;   // [N-1] or NULL
  int TR, G;                 // record strides (doubles), padded to even: x|u|c|il|iu|phi|zl|zu [+pad], eq|ineq gains [+pad]
  int n_compl;
  const int* compl_idx;
  unsigned long long compl_mask[MAX_STAGE_TYPES];   // per stage type: bit i set <=> constraint i is in indices_compl
  // inputs
  const double* p;           // [B][np]
  const double* lower;       // [B][nstage][nu]: bounds per instance and stage type
  const double* upper;       // [B][nstage][nu]
  const double* x1;          // [B][nx]
  const double* ubar;        // [B][(N-1) nu]
  const int* horizon;        // [B]
  // state
  double* traj;              // [2][B][N][TR]
  int* nomsel;               // [B]
  double* lam;               // [B][N][nx]
  double* tile;              // [B][D_NSLOT][N]
  double* tileN;             // [B][DN_NSLOT]
  double* gains;             // [B][N-1][G]
  double* Qu;                // [B][N-1][nu]
  double* sd;                // [SD_COUNT][B]
  int* si;                   // [SI_COUNT][B]
  double* filter;            // [2][FCAP][B]
  double* trace;             // [B][trace_cap][IPDDP_TRACE_COLS]
  int trace_cap;
  int fw_spec_max;           // rounds with at most this many active instances use k_forward_spec
  int bw_spec_max;           // rounds with at most this many active instances use k_backward_spec
  double* spec_bw;           // [bw_spec_max][BWS_WARPS-1][(N-1)(G+nu) + N nx] private gains / Qu / lambda of speculative sweeps
  double* spec_traj;         // [fw_spec_max][FWS_WARPS][N][TR] private trial records of the speculative line search
  int* done_list;            // queue mode (ipddp_solve_queue): slots whose instance terminated in this round, else NULL
  int* inst_of;              // queue mode: [B] queue index of the instance resident in slot b
  int list_sort;             // 1: active lists bucketed by expected work (heaviest first), 0: everything in one bucket (A/B)
  ipddp_options opt;

  IPDDP_D int type_of(int t) const { return stage_type ? (int)stage_type[t] : 0; }
  IPDDP_D const double* lower_of(int b, int type) const { return lower + ((size_t)b * nstage + type) * nu; }
  IPDDP_D const double* upper_of(int b, int type) const { return upper + ((size_t)b * nstage + type) * nu; }
  IPDDP_D double* rec(int set, int b, int t) const { return traj + (((size_t)set * B + b) * N + t) * TR; }
  IPDDP_D double& sdv(int f, int b) const { return sd[(size_t)f * B + b]; }
  IPDDP_D int& siv(int f, int b) const { return si[(size_t)f * B + b]; }
};

// offsets inside a trajectory record
template <class M> struct Rec {
  static constexpr int X = 0, U = M::NX, C = U + M::NU, IL = C + M::NC, IU = IL + M::NU, PHI = IU + M::NU,
                       ZL = PHI + M::NC, ZU = ZL + M::NU, SIZE = ZU + M::NU,
                       STRIDE = (SIZE + 1) & ~1;   // records are padded to an even number of doubles (16-byte multiples: bulk copies)
};
