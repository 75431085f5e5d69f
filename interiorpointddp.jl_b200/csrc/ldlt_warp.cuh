// Warp-cooperative symmetric-indefinite factorisation / solve / inertia for one small KKT matrix
// held in shared memory (packed upper triangle, column-major: A(i,j), i<=j, at i + j(j+1)/2).
//
// Replaces LAPACK dsytrf_rook('U') (reached by the reference at src/inertia_correction.jl:261 through
// FastLapackInterface), dsytrs_rook (`ldiv!(bk, eq[t])`, src/backward_pass.jl:148) and `inertia!`
// (src/inertia_correction.jl:54-205) for n <= 64, i.e. the unblocked dsytf2_rook path.  One warp works
// on one matrix: pivot searches are warp arg-max reductions that keep IDAMAX's first-maximum
// tie-break, the symmetric interchanges and the rank-1 / rank-2 trailing updates are spread over the
// lanes element-wise.  Every matrix element sees exactly the operation sequence of the unblocked
// LAPACK algorithm (rank-1 update as fma(x_i, -d*x_j, a_ij) like OpenBLAS' dsyr; everything else
// plain IEEE ops), so the factors, the pivot sequence and `info` do not depend on the lane mapping.
#pragma once
#include "kernels_common.cuh"

namespace ipk {

IPDDP_D int pk(int i, int j) { return i + ((j * (j + 1)) >> 1); }   // requires i <= j

// first-maximum arg-max over the warp: larger value wins, ties go to the smaller index
IPDDP_D void warp_argmax(double& val, int& idx) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    const double ov = __shfl_xor_sync(IPDDP_FULL_MASK, val, off);
    const int oi = __shfl_xor_sync(IPDDP_FULL_MASK, idx, off);
    if (ov > val || (ov == val && oi < idx)) { val = ov; idx = oi; }
  }
}

// symmetric interchange of rows/columns a < b inside the leading (b+1)x(b+1) block (dsytf2_rook style:
// trailing columns are NOT touched)
IPDDP_D void warp_sym_swap(double* A, int a, int b, int lane) {
  for (int i = lane; i < b; i += 32) {
    if (i < a) {
      const int pa = pk(i, a), pb = pk(i, b);
      const double t = A[pa]; A[pa] = A[pb]; A[pb] = t;
    } else if (i > a) {
      const int pa = pk(a, i), pb = pk(i, b);
      const double t = A[pa]; A[pa] = A[pb]; A[pb] = t;
    }
  }
  if (lane == 0) {
    const int pa = pk(a, a), pb = pk(b, b);
    const double t = A[pa]; A[pa] = A[pb]; A[pb] = t;
  }
}

// dsytf2_rook('U').  A: packed upper (n(n+1)/2), ipiv: n ints (LAPACK 1-based convention),
// ij: table of (i | j<<8) for every packed index, w: scratch of 4n doubles.  Returns info.
IPDDP_D int warp_sytf2_rook(int n, double* A, int* ipiv, const unsigned short* ij, double* w, int lane) {
  const double alpha = 0.6403882032022076;   // (1 + sqrt(17)) / 8
  const double sfmin = 2.2250738585072014e-308;
  int info = 0;
  int k = n - 1;   // 0-based pivot column
  while (k >= 0) {
    int kstep = 1, p = k, kp = k;
    const double absakk = fabs(A[pk(k, k)]);
    double colmax = 0.0;
    int imax = 0;
    if (k > 0) {
      double v = -1.0; int vi = 0x7fffffff;
      for (int i = lane; i < k; i += 32) {
        const double a = fabs(A[pk(i, k)]);
        if (a > v) { v = a; vi = i; }
      }
      warp_argmax(v, vi);
      colmax = v; imax = vi;
    }
    bool singular = false;
    if (fmax(absakk, colmax) == 0.0) {
      if (info == 0) info = k + 1;
      kp = k;
      singular = true;
    } else {
      if (!(absakk < alpha * colmax)) {
        kp = k;
      } else {
        for (;;) {
          // rowmax: largest off-diagonal magnitude in row/column imax of the leading block;
          // row segment (imax, j), j = imax+1..k, is searched first, the column segment replaces it
          // only if strictly larger
          double rv = -1.0; int rj = 0x7fffffff;
          for (int j = imax + 1 + lane; j <= k; j += 32) {
            const double a = fabs(A[pk(imax, j)]);
            if (a > rv) { rv = a; rj = j; }
          }
          warp_argmax(rv, rj);
          double rowmax = 0.0; int jmax = 0;
          if (imax != k) { rowmax = rv; jmax = rj; }
          if (imax > 0) {
            double cv = -1.0; int ci = 0x7fffffff;
            for (int i = lane; i < imax; i += 32) {
              const double a = fabs(A[pk(i, imax)]);
              if (a > cv) { cv = a; ci = i; }
            }
            warp_argmax(cv, ci);
            if (cv > rowmax) { rowmax = cv; jmax = ci; }
          }
          if (!(fabs(A[pk(imax, imax)]) < alpha * rowmax)) {
            kp = imax;
            break;
          } else if (p == jmax || rowmax <= colmax) {
            kp = imax;
            kstep = 2;
            break;
          } else {
            p = imax;
            colmax = rowmax;
            imax = jmax;
          }
        }
      }
      __syncwarp();
      if (kstep == 2 && p != k) {   // first interchange: k <-> p
        warp_sym_swap(A, p, k, lane);
        __syncwarp();
      }
      const int kk = k - kstep + 1;
      if (kp != kk) {               // second interchange: kk <-> kp
        warp_sym_swap(A, kp, kk, lane);
        if (kstep == 2 && lane == 0) {
          const int pa = pk(k - 1, k), pb = pk(kp, k);
          const double t = A[pa]; A[pa] = A[pb]; A[pb] = t;
        }
        __syncwarp();
      }
      if (kstep == 1) {
        if (k > 0) {
          double* x = A + pk(0, k);
          const double akk = A[pk(k, k)];
          const int ne = (k * (k + 1)) >> 1;
          if (fabs(akk) >= sfmin) {
            const double d11 = 1.0 / akk;
            for (int e = lane; e < ne; e += 32) {
              const unsigned short q = ij[e];
              const double xj = x[q >> 8];
              if (xj != 0.0) A[e] = IPDDP_FMA(x[q & 0xff], -d11 * xj, A[e]);
            }
            __syncwarp();
            for (int i = lane; i < k; i += 32) x[i] = x[i] * d11;
          } else {
            for (int i = lane; i < k; i += 32) x[i] = x[i] / akk;
            __syncwarp();
            for (int e = lane; e < ne; e += 32) {
              const unsigned short q = ij[e];
              const double xj = x[q >> 8];
              if (xj != 0.0) A[e] = IPDDP_FMA(x[q & 0xff], -akk * xj, A[e]);
            }
          }
          __syncwarp();
        }
      } else {
        if (k > 1) {
          const int m = k - 1;   // rows/columns 0..m-1 get updated
          double* xk = A + pk(0, k);
          double* xkm1 = A + pk(0, k - 1);
          const double d12 = A[pk(k - 1, k)];
          const double d22 = A[pk(k - 1, k - 1)] / d12;
          const double d11 = A[pk(k, k)] / d12;
          const double t = 1.0 / (d11 * d22 - 1.0);
          double* wk = w; double* wkm1 = w + n; double* rk = w + 2 * n; double* rkm1 = w + 3 * n;
          for (int j = lane; j < m; j += 32) {
            const double ak = xk[j], akm1 = xkm1[j];
            wkm1[j] = t * (d11 * akm1 - ak);
            wk[j] = t * (d22 * ak - akm1);
            rk[j] = ak / d12;
            rkm1[j] = akm1 / d12;
          }
          __syncwarp();
          const int ne = (m * (m + 1)) >> 1;
          for (int e = lane; e < ne; e += 32) {
            const unsigned short q = ij[e];
            const int i = q & 0xff, j = q >> 8;
            A[e] = A[e] - rk[i] * wk[j] - rkm1[i] * wkm1[j];
          }
          for (int j = lane; j < m; j += 32) {
            xk[j] = wk[j] / d12;
            xkm1[j] = wkm1[j] / d12;
          }
          __syncwarp();
        }
      }
    }
    (void)singular;
    if (lane == 0) {
      if (kstep == 1) {
        ipiv[k] = kp + 1;
      } else {
        ipiv[k] = -(p + 1);
        ipiv[k - 1] = -(kp + 1);
      }
    }
    k -= kstep;
  }
  __syncwarp();
  return info;
}

// number of positive eigenvalues of D with absolute tolerance tol (reference `inertia!`, atol = 1e-12);
// executed redundantly by every lane (uniform, broadcast reads)
IPDDP_D int warp_inertia_np(int n, const double* A, const int* ipiv, double tol) {
  int np = 0;
  int i = 0;
  // blocks are parsed from the bottom as get_D! does; 2x2 blocks are (i, i+1) with both ipiv < 0.
  // Walk upwards from the top using the pairing implied by the bottom-up parse: count the negative
  // entries below to know the parity.
  // A bottom-up parse pairs negatives from the end; runs of negatives always have even length
  // (LAPACK marks both rows of a 2x2 block), so a top-down pairing is identical.
  while (i < n) {
    bool two = false;
    double d12 = 0.0;
    if (i + 1 < n && ipiv[i] < 0 && ipiv[i + 1] < 0) {
      d12 = A[pk(i, i + 1)];
      two = (d12 != 0.0);
    }
    if (two) {
      const double d11 = A[pk(i, i)], d22 = A[pk(i + 1, i + 1)];
      const double a11 = fabs(d11), a22 = fabs(d22);
      const double s1 = 2.0 * fmax(fmax(a11, fabs(d12)), a22);
      double smin;
      if (a11 >= a22) smin = fabs((d11 / s1) * d22 - (d12 / s1) * d12);
      else            smin = fabs(d11 * (d22 / s1) - (d12 / s1) * d12);
      const double trace = d11 + d22;
      if (0.5 * s1 <= tol) {
      } else if (smin > tol || trace == 0.0) {
        np += 1;
      } else if (trace >= 0.0) {
        np += 1;
      }
      i += 2;
    } else {
      if (i + 1 < n && ipiv[i] < 0 && ipiv[i + 1] < 0) {
        // 2x2 block whose off-diagonal is exactly zero: reference treats both rows as 1x1 blocks
        if (A[pk(i, i)] > tol) np += 1;
        if (A[pk(i + 1, i + 1)] > tol) np += 1;
        i += 2;
      } else {
        if (A[pk(i, i)] > tol) np += 1;
        i += 1;
      }
    }
  }
  return np;
}

// dsytrs_rook('U') on NR right-hand sides held column-major in Bm (leading dimension n).
// Backward substitution spreads each column update over the lanes; the forward substitution's
// dgemv('T') dot products use 4 lanes per right-hand side (partial sums i mod 4, then a 2-step
// butterfly) -- the dot4 order.
template <int NR>
IPDDP_D void warp_sytrs_rook(int n, const double* A, const int* ipiv, double* Bm, int lane) {
  int k = n - 1;
  while (k >= 0) {
    if (ipiv[k] > 0) {
      const int kp = ipiv[k] - 1;
      if (kp != k && lane < NR) {
        const double t = Bm[k + lane * n]; Bm[k + lane * n] = Bm[kp + lane * n]; Bm[kp + lane * n] = t;
      }
      __syncwarp();
      const double* x = A + pk(0, k);
#pragma unroll
      for (int j = 0; j < NR; ++j) {
        const double t = -Bm[k + j * n];
        for (int i = lane; i < k; i += 32) Bm[i + j * n] = IPDDP_FMA(x[i], t, Bm[i + j * n]);
      }
      __syncwarp();
      if (lane < NR) Bm[k + lane * n] = Bm[k + lane * n] * (1.0 / A[pk(k, k)]);
      __syncwarp();
      k -= 1;
    } else {
      int kp = -ipiv[k] - 1;
      if (kp != k && lane < NR) {
        const double t = Bm[k + lane * n]; Bm[k + lane * n] = Bm[kp + lane * n]; Bm[kp + lane * n] = t;
      }
      __syncwarp();
      kp = -ipiv[k - 1] - 1;
      if (kp != k - 1 && lane < NR) {
        const double t = Bm[k - 1 + lane * n]; Bm[k - 1 + lane * n] = Bm[kp + lane * n]; Bm[kp + lane * n] = t;
      }
      __syncwarp();
      if (k > 1) {
        const double* xk = A + pk(0, k);
        const double* xkm1 = A + pk(0, k - 1);
#pragma unroll
        for (int j = 0; j < NR; ++j) {
          const double tk = -Bm[k + j * n];
          const double tkm1 = -Bm[k - 1 + j * n];
          for (int i = lane; i < k - 1; i += 32) {
            double bv = IPDDP_FMA(xk[i], tk, Bm[i + j * n]);
            Bm[i + j * n] = IPDDP_FMA(xkm1[i], tkm1, bv);
          }
        }
      }
      __syncwarp();
      if (lane < NR) {
        const double akm1k = A[pk(k - 1, k)];
        const double akm1 = A[pk(k - 1, k - 1)] / akm1k;
        const double ak = A[pk(k, k)] / akm1k;
        const double denom = akm1 * ak - 1.0;
        const double bkm1 = Bm[k - 1 + lane * n] / akm1k;
        const double bk = Bm[k + lane * n] / akm1k;
        Bm[k - 1 + lane * n] = (ak * bkm1 - bk) / denom;
        Bm[k + lane * n] = (akm1 * bk - bkm1) / denom;
      }
      __syncwarp();
      k -= 2;
    }
  }
  // forward: U' X = B
  const int g = lane & 3;
  k = 0;
  while (k < n) {
    const bool one = ipiv[k] > 0;
    const int ncol = one ? 1 : 2;   // pivot columns handled in this step
    for (int c = 0; c < ncol; ++c) {
      const int kc = k + c;
      if (k > 0) {
        const double* x = A + pk(0, kc);
        for (int j0 = 0; j0 < NR; j0 += 8) {
          const int j = j0 + (lane >> 2);
          double s = 0.0;
          if (j < NR)
            for (int i = g; i < k; i += 4) s = IPDDP_FMA(x[i], Bm[i + j * n], s);
          s = s + __shfl_xor_sync(IPDDP_FULL_MASK, s, 1);
          s = s + __shfl_xor_sync(IPDDP_FULL_MASK, s, 2);
          if (j < NR && g == 0) Bm[kc + j * n] = Bm[kc + j * n] - s;
        }
      }
    }
    __syncwarp();
    if (one) {
      const int kp = ipiv[k] - 1;
      if (kp != k && lane < NR) {
        const double t = Bm[k + lane * n]; Bm[k + lane * n] = Bm[kp + lane * n]; Bm[kp + lane * n] = t;
      }
      k += 1;
    } else {
      int kp = -ipiv[k] - 1;
      if (kp != k && lane < NR) {
        const double t = Bm[k + lane * n]; Bm[k + lane * n] = Bm[kp + lane * n]; Bm[kp + lane * n] = t;
      }
      __syncwarp();
      kp = -ipiv[k + 1] - 1;
      if (kp != k + 1 && lane < NR) {
        const double t = Bm[k + 1 + lane * n]; Bm[k + 1 + lane * n] = Bm[kp + lane * n]; Bm[kp + lane * n] = t;
      }
      k += 2;
    }
    __syncwarp();
  }
}

}  // namespace ipk
