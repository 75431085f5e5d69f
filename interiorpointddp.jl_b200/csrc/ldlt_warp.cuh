// Warp-cooperative symmetric-indefinite factorisation / solve / inertia for one small KKT matrix
// held in shared memory (packed upper triangle, column-major: A(i,j), i<=j, at i + j(j+1)/2).
//
// Replaces LAPACK dsytrf_rook('U') (reached by the reference at src/inertia_correction.jl:261 through
// FastLapackInterface), dsytrs_rook (`ldiv!(bk, eq[t])`, src/backward_pass.jl:148) and `inertia!`
// (src/inertia_correction.jl:54-205) for n <= 64, i.e. the unblocked dsytf2_rook path.  One warp works
// on one matrix; every matrix / right-hand-side element sees exactly the operation sequence of the
// unblocked LAPACK algorithms (rank-1 update as fma(x_i, -d*x_j, a_ij) like OpenBLAS' dsyr; everything
// else plain IEEE ops), so factors, pivot sequence, `info` and the solution do not depend on the lane mapping.
//
// The kernel that calls this is bound by instruction issue and dependent-instruction latency, not by
// FLOPs or bytes, so the code is organised around the pivot steps that IPDDP2's KKT matrices
// actually take:
//   * ldlt_step_fast handles a pivot column k < 32 that ends in a 1x1 pivot either without interchange or with
//     the single interchange dsytf2_rook finds in its first rook iteration (> 90 % of the steps).  It keeps the
//     pivot column in registers from the search to the elimination, decides "no interchange" with one warp vote
//     (|a_kk| >= alpha*|a_ik| for all i  <=>  |a_kk| >= alpha*colmax, since rounding is monotone) and the rook
//     acceptance |a_pp| >= alpha*rowmax the same way, finds imax with one REDUX.MAX on the high words (a second
//     one only on ties), and performs the symmetric interchange as stores of register values.  Anything else
//     (2x2 pivots, longer rook searches, NaNs, tiny pivots, singular columns, k >= 32) falls through -- before
//     any memory is modified -- to ldlt_step, the general implementation;
//   * trailing updates are SPARSE: pivot columns of IPDDP2 KKT matrices are mostly zero (measured: 23 %
//     non-zero entries, 7 % of the rank-1 element updates do arithmetic), so the non-zero row indices are
//     compacted with a ballot and only the nnz(nnz+1)/2 affected elements are touched.  Skipped
//     elements would receive fma(0, t, a) = a, i.e. the skip is exact for finite data (only the sign of
//     an exact zero can differ).  The non-zero row MASK of every pivot column is kept and reused by the second
//     triangular solve (its skipped terms are fma(0, b, s) = s as well);
//   * dsytrs_rook's first loop (interchange, rank-1 downdate of B, scaling) runs in the same k-descending
//     pivot order as the factorisation, so it is fused into it: the multipliers are still in registers.
//     The scaling B(k,:) *= 1/a_kk of a 1x1 pivot row is deferred (row k is not touched again by the first loop)
//     and applied to all rows at once before the second loop; only the U' solve is a separate pass;
//   * the inertia of D is counted while the pivots are produced (a D block is final when it is chosen);
//   * rows >= 32 (second slot per lane) are only touched while the pivot index is >= 32.
#pragma once
#include "kernels_common.cuh"

// test-only instrumentation (emulator builds with -DIPDDP_LDLT_STATS): how many pivot steps take which path
#if defined(IPDDP_SIMT_EMU) && defined(IPDDP_LDLT_STATS)
extern "C" long long g_ipddp_ldlt_steps[3];
#define IPDDP_LDLT_COUNT(w) do { if (lane == 0) __atomic_fetch_add(&g_ipddp_ldlt_steps[w], 1, __ATOMIC_RELAXED); } while (0)
#else
#define IPDDP_LDLT_COUNT(w) do { } while (0)
#endif


namespace ipk {

IPDDP_D int coff(int j) { return (j * (j + 1)) >> 1; }
IPDDP_D int pk(int i, int j) { return i + coff(j); }   // requires i <= j

// (a, b) with a <= b for the p-th element of a packed upper triangle, as a | b << 8
IPDDP_D unsigned tri_decode(int p) {
  int b = (int)((sqrtf(8.0f * (float)p + 1.0f) - 1.0f) * 0.5f);
  while (coff(b) > p) --b;
  while (coff(b + 1) <= p) ++b;
  return (unsigned)(p - coff(b)) | ((unsigned)b << 8);
}

// x / d for many numerators x and one divisor d.  With rd = RN(1/d) (one real division), q = RN(x*rd),
// r = x - d*q (exact in one FMA) and RN(q + r*rd) is the correctly rounded quotient (Markstein's theorem) as long as
// nothing under- or overflows; operands outside [2^-400, 2^400] (zeros, infinities, NaNs, denormals) take the plain
// division, except +-0 / d which is x*rd exactly.  3 instructions instead of ~25 -- and the plain FP64 division falls
// into a ~60-instruction slow path for zero numerators, which KKT matrices with a zero block produce all the time.
// The fallback division of DivBy sits out of line: the ~25 inlined instructions of an FP64 division per call site (18 sites
// in the 2x2 pivot path) are cold code in the middle of the hot pivot step.  Measured: -2.8 % sweep time.
#if !defined(IPDDP_SIMT_EMU)
static __device__ __noinline__ double div_rare(double x, double d) { return x / d; }
#else
IPDDP_D double div_rare(double x, double d) { return x / d; }
#endif
struct DivBy {
  double d, rd;
  bool ok;
  IPDDP_D explicit DivBy(double d_) : d(d_), rd(1.0 / d_) {
    const double a = fabs(d_);
    ok = a >= 0x1p-400 && a <= 0x1p400;
  }
  IPDDP_D double operator()(double x) const {
    const double ax = fabs(x);
    if (ok && ax >= 0x1p-400 && ax <= 0x1p400) {
      const double q = x * rd;
      const double r = IPDDP_FMA(-d, q, x);
      return IPDDP_FMA(r, rd, q);
    }
    if (ok && x == 0.0) return x * rd;
    return div_rare(x, d);
  }
};

// Maximum of the non-negative candidates (v0 at index lane, v1 at index lane+32; vld* = candidate present)
// and the 64-bit mask of the indices that attain it.
template <bool TWO>
IPDDP_D double warp_max_ties(double v0, bool vld0, double v1, bool vld1, unsigned long long& ties, bool two = TWO) {
  const unsigned h0 = vld0 ? (unsigned)__double2hiint(v0) : 0u, l0 = vld0 ? (unsigned)__double2loint(v0) : 0u;
  const unsigned h1 = (two && vld1) ? (unsigned)__double2hiint(v1) : 0u, l1 = (two && vld1) ? (unsigned)__double2loint(v1) : 0u;
  const unsigned mh = __reduce_max_sync(IPDDP_FULL_MASK, two ? (h0 > h1 ? h0 : h1) : h0);
  unsigned lc = (h0 == mh) ? l0 : 0u;
  if (two) { const unsigned lc1 = (h1 == mh) ? l1 : 0u; lc = lc > lc1 ? lc : lc1; }
  const unsigned ml = __reduce_max_sync(IPDDP_FULL_MASK, lc);
  const unsigned m0 = __ballot_sync(IPDDP_FULL_MASK, vld0 && h0 == mh && l0 == ml);
  unsigned long long t = m0;
  if (two) t |= (unsigned long long)__ballot_sync(IPDDP_FULL_MASK, vld1 && h1 == mh && l1 == ml) << 32;
  ties = t;
  return __hiloint2double((int)mh, (int)ml);
}

// symmetric interchange of rows/columns a < b inside the leading (b+1)x(b+1) block (dsytf2_rook style:
// trailing columns are NOT touched)
template <bool TWO>
IPDDP_D void warp_sym_swap(double* A, int a, int b, int lane, bool two = TWO) {
  const int ca = coff(a), cb = coff(b);
#pragma unroll
  for (int s = 0; s < (TWO ? 2 : 1); ++s) {
    if (s == 1 && !two) continue;
    const int i = lane + 32 * s;
    if (i < b && i != a) {
      const int pa = (i < a) ? ca + i : coff(i) + a, pb = cb + i;
      const double t = A[pa]; A[pa] = A[pb]; A[pb] = t;
    }
  }
  if (lane == 0) {
    const double t = A[ca + a]; A[ca + a] = A[cb + b]; A[cb + b] = t;
  }
}

// compacts the indices i (< 64) flagged by (f0 at lane, f1 at lane+32) into list[] ascending; returns count,
// m0 / m1 = the flag masks
template <bool TWO>
IPDDP_D int warp_compact(bool f0, bool f1, unsigned char* list, int lane, unsigned& m0, unsigned& m1, bool two = TWO) {
  m0 = __ballot_sync(IPDDP_FULL_MASK, f0);
  m1 = 0u;
  const unsigned lt = (1u << lane) - 1u;
  if (f0) list[__popc(m0 & lt)] = (unsigned char)lane;
  int n = __popc(m0);
  if (two) {
    m1 = __ballot_sync(IPDDP_FULL_MASK, f1);
    if (f1) list[n + __popc(m1 & lt)] = (unsigned char)(lane + 32);
    n += __popc(m1);
  }
  return n;
}

// Access to the K x NR right-hand sides in shared memory (column-major: element (r, j) at r + j K).  The first solve loop
// touches whole rows -- row k broadcast, row i updated, row interchanges -- NR = nx + 1 doubles each; all loads of a row
// update are issued before its FMAs.  (A row-major layout with rows padded to 16 bytes and 128-bit accesses was measured
// and lost 0.7 %: profiles/r2_ab/summary.md, series 4.)
template <int K, int NR> struct RhsL {
  static constexpr int LD = NR;
  static constexpr int SIZE = K * NR;
  static IPDDP_D int at(int r, int j) { return r + j * K; }
  static IPDDP_D void load_row(const double* B, int r, double (&v)[LD]) {
#pragma unroll
    for (int j = 0; j < NR; ++j) v[j] = B[r + j * K];
  }
  static IPDDP_D void store_row(double* B, int r, const double (&v)[LD]) {
#pragma unroll
    for (int j = 0; j < NR; ++j) B[r + j * K] = v[j];
  }
  // B(i,:) = fma(x, -B(k,:), B(i,:)) for one row i by one lane (the dger of dsytrs_rook's first loop)
  static IPDDP_D void downdate_row(double* B, int i, int k, double x) {
    double bk[LD], bi[LD];
    load_row(B, k, bk);
    load_row(B, i, bi);
#pragma unroll
    for (int j = 0; j < NR; ++j) bi[j] = IPDDP_FMA(x, -bk[j], bi[j]);
    store_row(B, i, bi);
  }
  // rows a <-> b
  static IPDDP_D void swap_rows(double* B, int a, int b, int lane) {
    if (lane < NR) { const double t = B[a + lane * K]; B[a + lane * K] = B[b + lane * K]; B[b + lane * K] = t; }
  }
};

// Scratch used by the factorisation and kept for the second solve (base must be 8-byte aligned):
// dinv: K doubles (deferred 1/a_kk row scalings, 1.0 where the step scaled in place); info: K 64-bit words, low half =
// non-zero multiplier rows 0..31 of the pivot column, high half = LAPACK's ipiv entry (1-based, negative for 2x2
// blocks) -- one shared-memory access per column in the second solve; nzhi: K uint32 (rows 32..63, K > 32 only);
// list: K bytes (current compaction).  The 2x2 pivot update additionally needs w: 4K doubles, passed separately.
template <int K> struct LdltScratch {
  static constexpr int DINV = 0;
  static constexpr int INFO = DINV + K * 8;
  static constexpr int NZHI = INFO + K * 8;
  static constexpr int LIST = NZHI + K * 4;
  static constexpr int BYTES = ((LIST + K + 15) / 16) * 16;
  static IPDDP_D double* dinv(unsigned char* s) { return reinterpret_cast<double*>(s + DINV); }
  static IPDDP_D unsigned long long* info(unsigned char* s) { return reinterpret_cast<unsigned long long*>(s + INFO); }
  static IPDDP_D unsigned* nzhi(unsigned char* s) { return reinterpret_cast<unsigned*>(s + NZHI); }
  static IPDDP_D unsigned char* list(unsigned char* s) { return s + LIST; }
  static IPDDP_D const double* dinv(const unsigned char* s) { return reinterpret_cast<const double*>(s + DINV); }
  static IPDDP_D const unsigned long long* info(const unsigned char* s) { return reinterpret_cast<const unsigned long long*>(s + INFO); }
  static IPDDP_D const unsigned* nzhi(const unsigned char* s) { return reinterpret_cast<const unsigned*>(s + NZHI); }
  static IPDDP_D unsigned long long pack(unsigned nzlo, int ipiv) { return (unsigned long long)nzlo | ((unsigned long long)(unsigned)ipiv << 32); }
  static IPDDP_D int ipiv_of(unsigned long long w) { return (int)(unsigned)(w >> 32); }
};

// General pivot step (column k) of dsytf2_rook('U') fused with dsytrs_rook's first loop on Bm.
// Returns kstep (1 or 2).  TWO = rows >= 32 may be involved (k >= 32).
template <int K, int NR, bool TWO>
IPDDP_D int ldlt_step(int k, double* __restrict__ A, double* __restrict__ Bm, double* __restrict__ w,
                      unsigned char* __restrict__ scratch, int lane,
                      unsigned tri_lane, double tol, int& info, int& np) {
  typedef LdltScratch<K> S;
  double* dinv = S::dinv(scratch);
  unsigned long long* cinfo = S::info(scratch);
  unsigned* nzhi = S::nzhi(scratch);
  unsigned char* list = S::list(scratch);
  const double alpha = 0.6403882032022076;   // (1 + sqrt(17)) / 8
  const double sfmin = 2.2250738585072014e-308;
  const int i0 = lane, i1 = lane + 32;
  const bool two = TWO && k >= 32;     // rows >= 32 exist in the leading block: uniform, so one instance serves every k
  int kstep = 1, p = k, kp = k;
  const int ck = coff(k);
  const double absakk = fabs(A[ck + k]);
  double colmax = 0.0;
  int imax = 0;
  if (k > 0) {
    unsigned long long ties;
    const bool v0 = i0 < k, v1 = two && i1 < k;
    colmax = warp_max_ties<TWO>(v0 ? fabs(A[ck + i0]) : 0.0, v0, v1 ? fabs(A[ck + i1]) : 0.0, v1, ties, two);
    imax = __ffsll((long long)ties) - 1;
  }
  if (fmax(absakk, colmax) == 0.0) {
    // exactly singular column: no elimination (LAPACK sets info and moves on); dsytrs would divide by zero
    if (info == 0) info = k + 1;
    if (lane == 0) { cinfo[k] = S::pack(0u, k + 1); nzhi[k] = 0u; dinv[k] = 1.0; }
    return 1;
  }
  if (absakk < alpha * colmax) {
    for (;;) {
      // largest off-diagonal magnitude in row/column imax of the leading block; ties: the row segment
      // (c > imax) is searched first and wins, lowest index inside a segment
      const int ci = coff(imax);
      const bool v0 = i0 <= k && i0 != imax, v1 = two && i1 <= k && i1 != imax;
      const double a0 = v0 ? fabs(i0 < imax ? A[ci + i0] : A[coff(i0) + imax]) : 0.0;
      const double a1 = v1 ? fabs(i1 < imax ? A[ci + i1] : A[coff(i1) + imax]) : 0.0;
      unsigned long long ties;
      const double rowmax = warp_max_ties<TWO>(a0, v0, a1, v1, ties, two);
      const unsigned long long rowpart = ties & ~((2ull << imax) - 1ull);
      const int jmax = __ffsll((long long)(rowpart ? rowpart : ties)) - 1;
      if (!(fabs(A[ci + imax]) < alpha * rowmax)) { kp = imax; break; }
      else if (p == jmax || rowmax <= colmax) { kp = imax; kstep = 2; break; }
      else { p = imax; colmax = rowmax; imax = jmax; }
    }
  }
  __syncwarp();
  if (kstep == 2 && p != k) {   // first interchange: k <-> p  (matrix and right-hand sides)
    warp_sym_swap<TWO>(A, p, k, lane, two);
    RhsL<K, NR>::swap_rows(Bm, k, p, lane);
    __syncwarp();
  }
  const int kk = k - kstep + 1;
  if (kp != kk) {               // second interchange: kk <-> kp
    warp_sym_swap<TWO>(A, kp, kk, lane, two);
    if (kstep == 2 && lane == 0) {
      const int pa = pk(k - 1, k), pb = pk(kp, k);
      const double t = A[pa]; A[pa] = A[pb]; A[pb] = t;
    }
    RhsL<K, NR>::swap_rows(Bm, kk, kp, lane);
    __syncwarp();
  }
  if (kstep == 1) {
    const double akk = A[ck + k];
    if (akk > tol) np += 1;
    int nnz = 0;
    unsigned m0 = 0u, m1 = 0u;
    double x0 = 0.0, x1 = 0.0;
    const bool big = fabs(akk) >= sfmin;
    const double rinv = 1.0 / akk;            // used by the column scaling (big) and by dsytrs' B(k,:) scaling
    const double d11 = big ? rinv : akk;
    if (k > 0) {
      double* x = A + ck;
      x0 = (i0 < k) ? x[i0] : 0.0;
      x1 = (two && i1 < k) ? x[i1] : 0.0;
      nnz = warp_compact<TWO>(x0 != 0.0, x1 != 0.0, list, lane, m0, m1, two);
      if (nnz > 0) {
        __syncwarp();
        if (!big) {   // tiny pivot: LAPACK divides the column first, then updates with -akk
          if (x0 != 0.0) { x0 = x0 / akk; x[i0] = x0; }
          if (two && x1 != 0.0) { x1 = x1 / akk; x[i1] = x1; }
          __syncwarp();
        }
        const int P = (nnz * (nnz + 1)) >> 1;
        for (int pp = lane; pp < P; pp += 32) {
          const unsigned q = (pp < 32) ? (tri_lane & 0xffffu) : (pp < 64) ? (tri_lane >> 16) : tri_decode(pp);
          const int i = list[q & 0xff], j = list[q >> 8];
          const int e = coff(j) + i;
          A[e] = IPDDP_FMA(x[i], -d11 * x[j], A[e]);
        }
        if (big) {    // scale the multipliers (registers keep the scaled values for the B downdate)
          __syncwarp();
          if (x0 != 0.0) { x0 = x0 * d11; x[i0] = x0; }
          if (two && x1 != 0.0) { x1 = x1 * d11; x[i1] = x1; }
        }
      }
    }
    if (lane == 0) { cinfo[k] = S::pack(m0, kp + 1); nzhi[k] = m1; dinv[k] = 1.0; }
    // dsytrs first loop for this pivot: B(0:k-1,:) -= x * B(k,:), then B(k,:) *= 1/akk
    if (x0 != 0.0) RhsL<K, NR>::downdate_row(Bm, i0, k, x0);
    if (two && x1 != 0.0) RhsL<K, NR>::downdate_row(Bm, i1, k, x1);
    __syncwarp();
    if (lane < NR) Bm[RhsL<K, NR>::at(k, lane)] = Bm[RhsL<K, NR>::at(k, lane)] * rinv;
    __syncwarp();
  } else {
    double* xk = A + ck;
    double* xkm1 = A + coff(k - 1);
    const double d12 = xk[k - 1];
    const DivBy by12(d12);
    // D^-1 scaled by d12 (dsytf2_rook / dsytrs_rook): d22 = a(k-1,k-1)/d12, d11 = a(k,k)/d12, denom = d11*d22 - 1
    const double d22 = by12(xkm1[k - 1]);
    const double d11 = by12(xk[k]);
    const DivBy bydn(d11 * d22 - 1.0);
    {   // inertia of the 2x2 block (reference inertia!, atol = tol)
      const double e11 = xkm1[k - 1], e22 = xk[k];
      if (d12 != 0.0) {
        const double a11 = fabs(e11), a22 = fabs(e22);
        const double s1 = 2.0 * fmax(fmax(a11, fabs(d12)), a22);
        const DivBy bys(s1);
        // (e11/s1)*e22 - (d12/s1)*d12 if |e11| >= |e22|, else e11*(e22/s1) - ...: one expression, the product commutes
        const bool first = a11 >= a22;
        const double smin = fabs(bys(first ? e11 : e22) * (first ? e22 : e11) - bys(d12) * d12);
        const double trace = e11 + e22;
        if (0.5 * s1 <= tol) {
        } else if (smin > tol || trace == 0.0) {
          np += 1;
        } else if (trace >= 0.0) {
          np += 1;
        }
      } else {   // reference: zero super-diagonal => two 1x1 blocks
        if (e11 > tol) np += 1;
        if (e22 > tol) np += 1;
      }
    }
    int nnz = 0;
    unsigned m0 = 0u, m1 = 0u;
    unsigned fm = 0u;                      // bit s: this lane's row lane + 32 s takes part in the update
    const int ns = (TWO && two) ? 2 : 1;
    if (k > 1) {
      const int m = k - 1;   // rows/columns 0..m-1 get updated
      const double t = bydn.rd;   // 1.0 / (d11 * d22 - 1.0)
      double* wk = w; double* wkm1 = w + K; double* rk = w + 2 * K; double* rkm1 = w + 3 * K;
#pragma unroll 1
      for (int s = 0; s < ns; ++s) {
        const int j = lane + 32 * s;
        if (j < m) {
          const double ak = xk[j], akm1 = xkm1[j];
          if ((ak != 0.0) || (akm1 != 0.0)) {
            fm |= 1u << s;
            wkm1[j] = t * (d11 * akm1 - ak);
            wk[j] = t * (d22 * ak - akm1);
            rk[j] = by12(ak);
            rkm1[j] = by12(akm1);
          }
        }
      }
      nnz = warp_compact<TWO>((fm & 1u) != 0u, (fm & 2u) != 0u, list, lane, m0, m1, two);
      __syncwarp();
      const int P = (nnz * (nnz + 1)) >> 1;
      for (int pp = lane; pp < P; pp += 32) {
        const unsigned q = (pp < 32) ? (tri_lane & 0xffffu) : (pp < 64) ? (tri_lane >> 16) : tri_decode(pp);
        const int i = list[q & 0xff], j = list[q >> 8];
        const int e = coff(j) + i;
        A[e] = A[e] - rk[i] * wk[j] - rkm1[i] * wkm1[j];
      }
#pragma unroll 1
      for (int s = 0; s < ns; ++s) {
        const int j = lane + 32 * s;
        if ((fm >> s) & 1u) {
          xk[j] = by12(wk[j]);
          xkm1[j] = by12(wkm1[j]);
        }
      }
      __syncwarp();
    }
    if (lane == 0) {
      cinfo[k] = S::pack(m0, -(p + 1)); cinfo[k - 1] = S::pack(m0, -(kp + 1));
      nzhi[k] = m1; nzhi[k - 1] = m1; dinv[k] = 1.0; dinv[k - 1] = 1.0;
    }
    // dsytrs first loop for the 2x2 block: two rank-1 downdates of B, then the 2x2 solve
#pragma unroll 1
    for (int s = 0; s < ns; ++s) {
      const int i = lane + 32 * s;
      if ((fm >> s) & 1u) {
        const double xa = xk[i], xb = xkm1[i];
        typedef RhsL<K, NR> RL;
        double bk[RL::LD], bkm[RL::LD], bi[RL::LD];
        RL::load_row(Bm, k, bk); RL::load_row(Bm, k - 1, bkm); RL::load_row(Bm, i, bi);
#pragma unroll
        for (int j = 0; j < NR; ++j) {
          const double bv = IPDDP_FMA(xa, -bk[j], bi[j]);
          bi[j] = IPDDP_FMA(xb, -bkm[j], bv);
        }
        RL::store_row(Bm, i, bi);
      }
    }
    __syncwarp();
    if (lane < NR) {
      // akm1 = a(k-1,k-1)/d12 = d22, ak = a(k,k)/d12 = d11, denom = akm1*ak - 1 = d11*d22 - 1 (same product)
      const double bkm1 = by12(Bm[RhsL<K, NR>::at(k - 1, lane)]);
      const double bk = by12(Bm[RhsL<K, NR>::at(k, lane)]);
      Bm[RhsL<K, NR>::at(k - 1, lane)] = bydn(d11 * bkm1 - bk);
      Bm[RhsL<K, NR>::at(k, lane)] = bydn(d22 * bk - bkm1);
    }
    __syncwarp();
  }
  return kstep;
}

// Fast pivot step for 0 <= k < 32 (see the header).  Returns false -- with nothing modified -- if the step is not a
// 1x1 pivot reached with at most one interchange, or if NaNs / tiny pivots / singular columns are involved.
template <int K, int NR>
IPDDP_D bool ldlt_step_fast(int k, double* __restrict__ A, double* __restrict__ Bm,
                            unsigned char* __restrict__ scratch, int lane, unsigned tri_lane,
                            double tol, int& np) {
  typedef LdltScratch<K> S;
  const double alpha = 0.6403882032022076;   // (1 + sqrt(17)) / 8
  const double sfmin = 2.2250738585072014e-308;
  const int ck = coff(k);
  const bool in = lane < k;
  double x = in ? A[ck + lane] : 0.0;        // pivot column, rows 0..k-1
  double piv = A[ck + k];
  int kp = k;
  const bool keep = __all_sync(IPDDP_FULL_MASK, fabs(piv) >= alpha * fabs(x));   // false on any NaN
  if (!keep) {
    // imax: first row attaining max |x|
    const double ax = fabs(x);
    const unsigned hi = in ? (unsigned)__double2hiint(ax) : 0u;
    const unsigned mh = __reduce_max_sync(IPDDP_FULL_MASK, hi);
    if (mh >= 0x7ff00000u || piv != piv) return false;   // NaN (and Inf) rows show up in the maximum: no extra vote
    unsigned cand = __ballot_sync(IPDDP_FULL_MASK, in && hi == mh);
    if (cand & (cand - 1u)) {   // several rows share the high word: compare the low words among them
      const unsigned lo = (in && hi == mh) ? (unsigned)__double2loint(ax) : 0u;
      const unsigned ml = __reduce_max_sync(IPDDP_FULL_MASK, lo);
      cand = __ballot_sync(IPDDP_FULL_MASK, in && hi == mh && lo == ml);
    }
    const int imax = __ffs(cand) - 1;
    const int ci = coff(imax);
    // row / column imax of the leading (k+1) x (k+1) block, signed: it becomes the pivot column after the interchange
    const bool vr = lane <= k && lane != imax;
    const int pa = (lane < imax) ? ci + lane : coff(lane) + imax;
    const double a = vr ? A[pa] : 0.0;
    const double aii = A[ci + imax];
    // rook acceptance of the first candidate: |a_ii| >= alpha * rowmax  (false on any NaN)
    if (!__all_sync(IPDDP_FULL_MASK, fabs(aii) >= alpha * fabs(a))) return false;
    if (!(fabs(aii) >= sfmin)) return false;   // tiny pivot: general path
    // ---- committed: symmetric interchange k <-> imax as stores of register values (row / column imax receives the
    //      old column k) and the row interchange of the right-hand sides
    if (in && lane != imax) A[pa] = x;
    if (lane == 0) { A[ci + imax] = piv; A[ck + k] = aii; }
    RhsL<K, NR>::swap_rows(Bm, k, imax, lane);
    x = (lane == imax) ? x : (in ? a : 0.0);
    piv = aii;
    kp = imax;
  } else if (!(fabs(piv) >= sfmin)) {
    return false;                              // singular column (piv == 0 => column all zero), tiny pivot
  }
  // ---- committed: 1x1 pivot piv at (k,k)
  const unsigned nzm = __ballot_sync(IPDDP_FULL_MASK, x != 0.0);
  const double rinv = 1.0 / piv;
  if (piv > tol) np += 1;
  if (lane == 0) {
    S::info(scratch)[k] = S::pack(nzm, kp + 1);
    S::dinv(scratch)[k] = rinv;
  }                                              // (nzhi[k] is only read for k > 32: nothing to clear here)
  if (nzm == 0u) return true;                // nothing to eliminate; B(k,:) scaling is deferred
  unsigned char* list = S::list(scratch);
  if (x != 0.0) list[__popc(nzm & ((1u << lane) - 1u))] = (unsigned char)lane;
  __syncwarp();
  const int nnz = __popc(nzm);
  const int P = (nnz * (nnz + 1)) >> 1;
  for (int p0 = 0; p0 < P; p0 += 32) {
    const int pp = p0 + lane;
    const bool act = pp < P;
    const unsigned q = (p0 == 0) ? (tri_lane & 0xffffu) : (p0 == 32) ? (tri_lane >> 16) : tri_decode(act ? pp : 0);
    const int i = act ? list[q & 0xff] : 0, j = act ? list[q >> 8] : 0;
    const double xi = __shfl_sync(IPDDP_FULL_MASK, x, i), xj = __shfl_sync(IPDDP_FULL_MASK, x, j);
    if (act) {
      const int e = coff(j) + i;
      A[e] = IPDDP_FMA(xi, -rinv * xj, A[e]);
    }
  }
  if (x != 0.0) {
    const double xs = x * rinv;                // multiplier
    A[ck + lane] = xs;
    // dsytrs first loop for this pivot: B(0:k-1,:) -= xs * B(k,:)   (B(k,:) *= 1/piv is deferred: dinv[k])
    RhsL<K, NR>::downdate_row(Bm, lane, k, xs);
  } else if (kp != k && in) {
    A[ck + lane] = 0.0;                        // the interchanged column's zeros
  }
  __syncwarp();
  return true;
}

// The same fast step for a pivot column 32 <= k < 64: rows 0..31 sit in lane slot 0, rows 32..k-1 in slot 1 (the first
// k - 32 lanes).  The multipliers of the trailing update are read from the column in shared memory (not shuffled, two
// slots), which costs one more __syncwarp than the single-slot step; everything else is the same sequence.
template <int K, int NR>
IPDDP_D bool ldlt_step_fast2(int k, double* __restrict__ A, double* __restrict__ Bm,
                             unsigned char* __restrict__ scratch, int lane, unsigned tri_lane,
                             double tol, int& np) {
  typedef LdltScratch<K> S;
  const double alpha = 0.6403882032022076;   // (1 + sqrt(17)) / 8
  const double sfmin = 2.2250738585072014e-308;
  const int ck = coff(k);
  const int i1 = lane + 32;
  const bool in1 = i1 < k;
  double* xc = A + ck;
  double x0 = xc[lane];                       // rows 0..31 (k >= 32: all present)
  double x1 = in1 ? xc[i1] : 0.0;             // rows 32..k-1
  double piv = xc[k];
  int kp = k;
  const bool keep = __all_sync(IPDDP_FULL_MASK, fabs(piv) >= alpha * fabs(x0) && fabs(piv) >= alpha * fabs(x1));
  if (!keep) {
    if (__any_sync(IPDDP_FULL_MASK, x0 != x0 || x1 != x1) || piv != piv) return false;
    // imax: first row attaining max |x| over both slots
    const double ax0 = fabs(x0), ax1 = fabs(x1);
    const unsigned h0 = (unsigned)__double2hiint(ax0), h1 = in1 ? (unsigned)__double2hiint(ax1) : 0u;
    const unsigned mh = __reduce_max_sync(IPDDP_FULL_MASK, h0 > h1 ? h0 : h1);
    unsigned c0 = __ballot_sync(IPDDP_FULL_MASK, h0 == mh), c1 = __ballot_sync(IPDDP_FULL_MASK, in1 && h1 == mh);
    if (__popc(c0) + __popc(c1) > 1) {        // several rows share the high word: compare the low words among them
      const unsigned l0 = (h0 == mh) ? (unsigned)__double2loint(ax0) : 0u;
      const unsigned l1 = (in1 && h1 == mh) ? (unsigned)__double2loint(ax1) : 0u;
      const unsigned ml = __reduce_max_sync(IPDDP_FULL_MASK, l0 > l1 ? l0 : l1);
      c0 = __ballot_sync(IPDDP_FULL_MASK, h0 == mh && l0 == ml);
      c1 = __ballot_sync(IPDDP_FULL_MASK, in1 && h1 == mh && l1 == ml);
    }
    const int imax = c0 ? __ffs(c0) - 1 : 32 + __ffs(c1) - 1;
    const int ci = coff(imax);
    // row / column imax of the leading (k+1) x (k+1) block (signed): rows lane (slot 0) and lane + 32 <= k (slot 1)
    const bool v0 = lane != imax, v1 = i1 <= k && i1 != imax;
    const int pa0 = (lane < imax) ? ci + lane : coff(lane) + imax;
    const int pa1 = (i1 < imax) ? ci + i1 : coff(i1) + imax;
    const double a0 = v0 ? A[pa0] : 0.0;
    const double a1 = v1 ? A[pa1] : 0.0;
    const double aii = A[ci + imax];
    if (!__all_sync(IPDDP_FULL_MASK, fabs(aii) >= alpha * fabs(a0) && fabs(aii) >= alpha * fabs(a1))) return false;
    if (!(fabs(aii) >= sfmin)) return false;
    // ---- committed: symmetric interchange k <-> imax (row / column imax receives the old column k), right-hand sides,
    //      and the new (unscaled) pivot column written back for the trailing update
    if (lane != imax) { A[pa0] = x0; xc[lane] = a0; }
    if (in1 && i1 != imax) { A[pa1] = x1; xc[i1] = a1; }
    if (lane == 0) { A[ci + imax] = piv; xc[k] = aii; }
    RhsL<K, NR>::swap_rows(Bm, k, imax, lane);
    x0 = (lane == imax) ? x0 : a0;
    x1 = in1 ? ((i1 == imax) ? x1 : a1) : 0.0;
    piv = aii;
    kp = imax;
  } else if (!(fabs(piv) >= sfmin)) {
    return false;
  }
  const unsigned nz0 = __ballot_sync(IPDDP_FULL_MASK, x0 != 0.0), nz1 = __ballot_sync(IPDDP_FULL_MASK, x1 != 0.0);
  const double rinv = 1.0 / piv;
  if (piv > tol) np += 1;
  if (lane == 0) {
    S::info(scratch)[k] = S::pack(nz0, kp + 1);
    S::nzhi(scratch)[k] = nz1;
    S::dinv(scratch)[k] = rinv;
  }
  if ((nz0 | nz1) == 0u) return true;
  unsigned char* list = S::list(scratch);
  const unsigned lt = (1u << lane) - 1u;
  const int n0 = __popc(nz0);
  if (x0 != 0.0) list[__popc(nz0 & lt)] = (unsigned char)lane;
  if (x1 != 0.0) list[n0 + __popc(nz1 & lt)] = (unsigned char)i1;
  __syncwarp();
  const int nnz = n0 + __popc(nz1);
  const int P = (nnz * (nnz + 1)) >> 1;
  for (int pp = lane; pp < P; pp += 32) {
    const unsigned q = (pp < 32) ? (tri_lane & 0xffffu) : (pp < 64) ? (tri_lane >> 16) : tri_decode(pp);
    const int i = list[q & 0xff], j = list[q >> 8];
    const int e = coff(j) + i;
    A[e] = IPDDP_FMA(xc[i], -rinv * xc[j], A[e]);
  }
  __syncwarp();                                // the unscaled column was read by other lanes: now it can be scaled
  if (x0 != 0.0) {
    const double xs = x0 * rinv;
    xc[lane] = xs;
    RhsL<K, NR>::downdate_row(Bm, lane, k, xs);
  }
  if (x1 != 0.0) {
    const double xs = x1 * rinv;
    xc[i1] = xs;
    RhsL<K, NR>::downdate_row(Bm, i1, k, xs);
  }
  __syncwarp();
  return true;
}

// dsytf2_rook('U') on the packed matrix A of order K, fused with the first (U D) loop of dsytrs_rook on
// the NR right-hand sides in Bm (column-major, leading dimension K).  Returns info; np_out = number of
// positive eigenvalues of D.  If info != 0 the contents of Bm are meaningless (the caller restarts).
// scratch: LdltScratch<K>::BYTES, 8-byte aligned.
// (row, column) pairs of the packed triangle elements `lane` and `lane + 32`: one register serves trailing updates of
// up to 10 non-zero rows (55 pairs) without decoding.  Loop invariant: compute once per kernel.
IPDDP_D unsigned ldlt_tri_lane(int lane) { return tri_decode(lane) | (tri_decode(lane + 32) << 16); }

template <int K, int NR>
IPDDP_D int warp_ldlt_factor(double* __restrict__ A, double* __restrict__ Bm, double* __restrict__ w,
                             unsigned char* __restrict__ scratch, int lane, double tol,
                             int& np_out, unsigned tri_lane) {
  int info = 0, np = 0;
  int k = K - 1;
  if (K > 32) {
    while (k >= 32) {
      if (ldlt_step_fast2<K, NR>(k, A, Bm, scratch, lane, tri_lane, tol, np)) { k -= 1; continue; }
      k -= ldlt_step<K, NR, true>(k, A, Bm, w, scratch, lane, tri_lane, tol, info, np);
    }
  }
  while (k >= 0) {
    while (k >= 0 && ldlt_step_fast<K, NR>(k, A, Bm, scratch, lane, tri_lane, tol, np)) {   // fast steps in their own inner loop
      IPDDP_LDLT_COUNT(0);
      k -= 1;
    }
    if (k < 0) break;
    // one instance of the general step for every k (rows >= 32 are gated by a uniform run-time flag): half the code
    const int ks = ldlt_step<K, NR, (K > 32)>(k, A, Bm, w, scratch, lane, tri_lane, tol, info, np);
    IPDDP_LDLT_COUNT(ks);
    k -= ks;
  }
  __syncwarp();
  np_out = np;
  return info;
}

// second loop of dsytrs_rook('U'): U' X = B, k ascending, preceded by the deferred row scalings of the first loop.
// 4 lanes per right-hand side accumulate the dgemv('T') dot product in the dot4 order (partial sums by i mod 4,
// ascending i, 2-step butterfly); zero multipliers are skipped through the per-column row masks.
template <int K, int NR>
IPDDP_D void warp_ldlt_solve_forward_wide(const double* __restrict__ A, double* __restrict__ Bm,
                                          const unsigned char* __restrict__ scratch, int lane);

template <int K, int NR>
IPDDP_D void warp_ldlt_solve_forward(const double* __restrict__ A, double* __restrict__ Bm,
                                     const unsigned char* __restrict__ scratch, int lane) {
  if constexpr (NR > 8) {     // more than 7 states: the right-hand sides are taken in groups of 8 columns
    warp_ldlt_solve_forward_wide<K, NR>(A, Bm, scratch, lane);
    return;
  }
  typedef LdltScratch<K> S;
  const double* dinv = S::dinv(scratch);
  const unsigned long long* cinfo = S::info(scratch);
  const unsigned* nzhi = S::nzhi(scratch);
#pragma unroll
  for (int s = 0; s < (K > 32 ? 2 : 1); ++s) {
    const int r = lane + 32 * s;
    if (r < K) {
      const double d = dinv[r];
      typedef RhsL<K, NR> RL;
      double br[RL::LD];
      RL::load_row(Bm, r, br);
#pragma unroll
      for (int j = 0; j < NR; ++j) br[j] = br[j] * d;
      RL::store_row(Bm, r, br);
    }
  }
  __syncwarp();
  const int g = lane & 3;
  const int j = lane >> 2;
  const bool act = j < NR;
  const unsigned gm = 0x11111111u << g;
  typedef RhsL<K, NR> RL;
  const int jc = act ? j : 0;
  auto bj = [&](int i) -> double& { return Bm[RL::at(i, jc)]; };
  int k = 0;
  while (k < K) {
    const unsigned long long cw = cinfo[k];      // column 0 carries an empty mask
    const int pv = S::ipiv_of(cw);
    const bool one = pv > 0;
    const unsigned mlo = (unsigned)cw;
    unsigned mhi = 0u;
    if (K > 32 && k > 32) mhi = nzhi[k];        // rows >= 32 can only be non-zero in columns k > 32
    const bool any = (mlo | mhi) != 0u;
    double sa = 0.0, sb = 0.0;
    if (any) {
      const double* xa = A + coff(k);
      const double* xb = A + coff(k + (one ? 0 : 1));
      if (act) {
        unsigned m = mlo & gm;
        while (m) {
          const int i = __ffs(m) - 1;
          m &= m - 1u;
          const double bv = bj(i);
          sa = IPDDP_FMA(xa[i], bv, sa);
          if (!one) sb = IPDDP_FMA(xb[i], bv, sb);
        }
        if (K > 32 && mhi != 0u) {
          m = mhi & gm;
          while (m) {
            const int i = 32 + __ffs(m) - 1;
            m &= m - 1u;
            const double bv = bj(i);
            sa = IPDDP_FMA(xa[i], bv, sa);
            if (!one) sb = IPDDP_FMA(xb[i], bv, sb);
          }
        }
      }
      __syncwarp();
      sa = sa + __shfl_xor_sync(IPDDP_FULL_MASK, sa, 1);
      sa = sa + __shfl_xor_sync(IPDDP_FULL_MASK, sa, 2);
      if (!one) {
        sb = sb + __shfl_xor_sync(IPDDP_FULL_MASK, sb, 1);
        sb = sb + __shfl_xor_sync(IPDDP_FULL_MASK, sb, 2);
      }
    }
    if (one) {
      // B(k,:) -= sa, then the interchange k <-> kp, done by the lane that owns column j in one read-modify-write
      const int kp = pv - 1;
      if (any || kp != k) {
        if (act && g == 0) {
          double v = bj(k);
          if (any) v = v - sa;
          if (kp != k) { const double t = bj(kp); bj(kp) = v; v = t; }
          bj(k) = v;
        }
        __syncwarp();
      }
      k += 1;
    } else {
      if (any) {
        if (act && g == 0) { bj(k) = bj(k) - sa; bj(k + 1) = bj(k + 1) - sb; }
        __syncwarp();
      }
      int kp = -pv - 1;
      if (kp != k) { RL::swap_rows(Bm, k, kp, lane); __syncwarp(); }
      kp = -S::ipiv_of(cinfo[k + 1]) - 1;
      if (kp != k + 1) { RL::swap_rows(Bm, k + 1, kp, lane); __syncwarp(); }
      k += 2;
    }
  }
}

// The same loop for NR > 8 right-hand sides (models with more than 7 states; the reference's experiments have at most 4):
// lane (g, j) serves the columns j, j + 8, j + 16, ... one after the other inside every pivot step.  The columns of B are
// independent of each other, so every column sees exactly the operations, in exactly the order, of the loop above.
template <int K, int NR>
IPDDP_D void warp_ldlt_solve_forward_wide(const double* __restrict__ A, double* __restrict__ Bm,
                                          const unsigned char* __restrict__ scratch, int lane) {
  typedef LdltScratch<K> S;
  typedef RhsL<K, NR> RL;
  const double* dinv = S::dinv(scratch);
  const unsigned long long* cinfo = S::info(scratch);
  const unsigned* nzhi = S::nzhi(scratch);
  constexpr int NG = (NR + 7) / 8;
#pragma unroll
  for (int s = 0; s < (K > 32 ? 2 : 1); ++s) {
    const int r = lane + 32 * s;
    if (r < K) {
      const double d = dinv[r];
      double br[RL::LD];
      RL::load_row(Bm, r, br);
#pragma unroll
      for (int j = 0; j < NR; ++j) br[j] = br[j] * d;
      RL::store_row(Bm, r, br);
    }
  }
  __syncwarp();
  const int g = lane & 3;
  const int j0 = lane >> 2;
  const unsigned gm = 0x11111111u << g;
  int k = 0;
  while (k < K) {
    const unsigned long long cw = cinfo[k];      // column 0 carries an empty mask
    const int pv = S::ipiv_of(cw);
    const bool one = pv > 0;
    const unsigned mlo = (unsigned)cw;
    unsigned mhi = 0u;
    if (K > 32 && k > 32) mhi = nzhi[k];        // rows >= 32 can only be non-zero in columns k > 32
    const bool any = (mlo | mhi) != 0u;
    const double* xa = A + coff(k);
    const double* xb = A + coff(k + (one ? 0 : 1));
    const int kp1 = pv - 1;                      // 1 x 1 pivot: the row interchanged with row k
    for (int q = 0; q < NG; ++q) {
      const int j = j0 + 8 * q;
      const bool act = j < NR;
      const int jc = act ? j : 0;
      auto bj = [&](int i) -> double& { return Bm[RL::at(i, jc)]; };
      double sa = 0.0, sb = 0.0;
      if (any) {
        if (act) {
          unsigned m = mlo & gm;
          while (m) {
            const int i = __ffs(m) - 1;
            m &= m - 1u;
            const double bv = bj(i);
            sa = IPDDP_FMA(xa[i], bv, sa);
            if (!one) sb = IPDDP_FMA(xb[i], bv, sb);
          }
          if (K > 32 && mhi != 0u) {
            m = mhi & gm;
            while (m) {
              const int i = 32 + __ffs(m) - 1;
              m &= m - 1u;
              const double bv = bj(i);
              sa = IPDDP_FMA(xa[i], bv, sa);
              if (!one) sb = IPDDP_FMA(xb[i], bv, sb);
            }
          }
        }
        __syncwarp();
        sa = sa + __shfl_xor_sync(IPDDP_FULL_MASK, sa, 1);
        sa = sa + __shfl_xor_sync(IPDDP_FULL_MASK, sa, 2);
        if (!one) {
          sb = sb + __shfl_xor_sync(IPDDP_FULL_MASK, sb, 1);
          sb = sb + __shfl_xor_sync(IPDDP_FULL_MASK, sb, 2);
        }
      }
      if (act && g == 0) {
        if (one) {
          if (any || kp1 != k) {     // B(k, j) -= sa, then the interchange k <-> kp in one read-modify-write
            double v = bj(k);
            if (any) v = v - sa;
            if (kp1 != k) { const double t = bj(kp1); bj(kp1) = v; v = t; }
            bj(k) = v;
          }
        } else if (any) {
          bj(k) = bj(k) - sa;
          bj(k + 1) = bj(k + 1) - sb;
        }
      }
    }
    __syncwarp();
    if (one) {
      k += 1;
    } else {
      int kp = -pv - 1;
      if (kp != k) { RL::swap_rows(Bm, k, kp, lane); __syncwarp(); }
      kp = -S::ipiv_of(cinfo[k + 1]) - 1;
      if (kp != k + 1) { RL::swap_rows(Bm, k + 1, kp, lane); __syncwarp(); }
      k += 2;
    }
  }
}

}  // namespace ipk
