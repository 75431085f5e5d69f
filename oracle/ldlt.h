// ORACLE (test infrastructure) -- symmetric-indefinite factorisation with rook pivoting.
//
// The reference factorises every stage KKT matrix with LAPACK dsytrf_rook('U') through
// FastLapackInterface (reference src/inertia_correction.jl:261), solves with dsytrs_rook through
// `ldiv!(bk, eq[t])` (reference src/backward_pass.jl:148) and counts the inertia of D with
// `inertia!` (reference src/inertia_correction.jl:54-205, get_D! :207-255).  LAPACK/OpenBLAS are a
// third-party dependency that is NOT in the reference tree (Julia 1.10.4 OpenBLAS_jll);
// this file restates the published unblocked algorithm dsytf2_rook / dsytrs_rook (n <= 64 never
// reaches the blocked code).  tests/test_ldlt_vs_lapack.py pins it bit-for-bit (factors, ipiv, info)
// against the OpenBLAS binary bundled with SciPy.
//
// Floating-point conventions (fixed here, reproduced by the CUDA kernels):
//   * rank-1 update of a 1x1 pivot:  a_ij = fma(x_i, t, a_ij),  t = -d11 * x_j   (OpenBLAS dsyr = axpy with FMA)
//   * 2x2 update, scalings, pivot tests: plain IEEE ops in the order written in LAPACK
//   * triangular solves: dger as fma(a_ik, -b_kj, b_ij); dgemv('T') as b_kj - dot4(a_:k, b_:j)
#pragma once
#include <math.h>
#include <float.h>

// 4-way interleaved dot product with FMA: the ONE summation order used for every BLAS-like
// contraction in the oracle (OpenBLAS kernels are SIMD-interleaved too; the exact order of the
// reference's BLAS build is not reproducible).
static inline double dot4(int n, const double* a, int sa, const double* b, int sb) {
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
  int i = 0;
  for (; i + 3 < n; i += 4) {
    s0 = fma(a[(i + 0) * sa], b[(i + 0) * sb], s0);
    s1 = fma(a[(i + 1) * sa], b[(i + 1) * sb], s1);
    s2 = fma(a[(i + 2) * sa], b[(i + 2) * sb], s2);
    s3 = fma(a[(i + 3) * sa], b[(i + 3) * sb], s3);
  }
  if (i < n) s0 = fma(a[i * sa], b[i * sb], s0);
  if (i + 1 < n) s1 = fma(a[(i + 1) * sa], b[(i + 1) * sb], s1);
  if (i + 2 < n) s2 = fma(a[(i + 2) * sa], b[(i + 2) * sb], s2);
  return (s0 + s1) + (s2 + s3);
}

// index of first element of maximum |x| (BLAS idamax, 0-based; n >= 1)
static inline int idamax0(int n, const double* x, int inc) {
  int im = 0;
  double vm = fabs(x[0]);
  for (int i = 1; i < n; ++i) {
    double v = fabs(x[i * inc]);
    if (v > vm) { vm = v; im = i; }
  }
  return im;
}

static inline void swap_strided(int n, double* x, int ix, double* y, int iy) {
  for (int i = 0; i < n; ++i) { double t = x[i * ix]; x[i * ix] = y[i * iy]; y[i * iy] = t; }
}

// dsytf2_rook, UPLO='U'.  A is n x n column-major with leading dimension lda; only the upper
// triangle is referenced.  ipiv uses LAPACK's 1-based convention (negative pairs for 2x2 blocks).
// Returns info (0, or k>0 for the first exactly-zero pivot column).
static inline int sytf2_rook_upper(int n, double* A, int lda, int* ipiv) {
#define A_(i, j) A[((i) - 1) + ((j) - 1) * lda]
  const double alpha = (1.0 + sqrt(17.0)) / 8.0;
  const double sfmin = DBL_MIN;  // dlamch('S') for IEEE double
  int info = 0;
  int k = n;
  while (k >= 1) {
    int kstep = 1, p = k, kp = k;
    double absakk = fabs(A_(k, k));
    int imax = 0;
    double colmax = 0.0;
    if (k > 1) {
      imax = 1 + idamax0(k - 1, &A_(1, k), 1);
      colmax = fabs(A_(imax, k));
    }
    if (fmax(absakk, colmax) == 0.0) {
      if (info == 0) info = k;
      kp = k;
    } else {
      if (!(absakk < alpha * colmax)) {
        kp = k;
      } else {
        for (;;) {
          int jmax = 0;
          double rowmax = 0.0;
          if (imax != k) {
            jmax = imax + 1 + idamax0(k - imax, &A_(imax, imax + 1), lda);
            rowmax = fabs(A_(imax, jmax));
          }
          if (imax > 1) {
            int itemp = 1 + idamax0(imax - 1, &A_(1, imax), 1);
            double dtemp = fabs(A_(itemp, imax));
            if (dtemp > rowmax) { rowmax = dtemp; jmax = itemp; }
          }
          if (!(fabs(A_(imax, imax)) < alpha * rowmax)) {
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
      // first swap (2x2 only): rows/columns k and p
      if (kstep == 2 && p != k) {
        if (p > 1) swap_strided(p - 1, &A_(1, k), 1, &A_(1, p), 1);
        if (p < k - 1) swap_strided(k - p - 1, &A_(p + 1, k), 1, &A_(p, p + 1), lda);
        double t = A_(k, k); A_(k, k) = A_(p, p); A_(p, p) = t;
      }
      // second swap: rows/columns kk and kp
      int kk = k - kstep + 1;
      if (kp != kk) {
        if (kp > 1) swap_strided(kp - 1, &A_(1, kk), 1, &A_(1, kp), 1);
        if (kk > 1 && kp < kk - 1) swap_strided(kk - kp - 1, &A_(kp + 1, kk), 1, &A_(kp, kp + 1), lda);
        double t = A_(kk, kk); A_(kk, kk) = A_(kp, kp); A_(kp, kp) = t;
        if (kstep == 2) { t = A_(k - 1, k); A_(k - 1, k) = A_(kp, k); A_(kp, k) = t; }
      }
      if (kstep == 1) {
        if (k > 1) {
          if (fabs(A_(k, k)) >= sfmin) {
            double d11 = 1.0 / A_(k, k);
            for (int j = 1; j <= k - 1; ++j) {
              if (A_(j, k) == 0.0) continue;   // dsyr skips zero x(j) (keeps signed zeros untouched)
              double t = -d11 * A_(j, k);
              for (int i = 1; i <= j; ++i) A_(i, j) = fma(A_(i, k), t, A_(i, j));
            }
            for (int i = 1; i <= k - 1; ++i) A_(i, k) = A_(i, k) * d11;
          } else {
            double d11 = A_(k, k);
            for (int i = 1; i <= k - 1; ++i) A_(i, k) = A_(i, k) / d11;
            for (int j = 1; j <= k - 1; ++j) {
              if (A_(j, k) == 0.0) continue;
              double t = -d11 * A_(j, k);
              for (int i = 1; i <= j; ++i) A_(i, j) = fma(A_(i, k), t, A_(i, j));
            }
          }
        }
      } else {
        if (k > 2) {
          double d12 = A_(k - 1, k);
          double d22 = A_(k - 1, k - 1) / d12;
          double d11 = A_(k, k) / d12;
          double t = 1.0 / (d11 * d22 - 1.0);
          for (int j = k - 2; j >= 1; --j) {
            double wkm1 = t * (d11 * A_(j, k - 1) - A_(j, k));
            double wk = t * (d22 * A_(j, k) - A_(j, k - 1));
            for (int i = j; i >= 1; --i)
              A_(i, j) = A_(i, j) - (A_(i, k) / d12) * wk - (A_(i, k - 1) / d12) * wkm1;
            A_(j, k) = wk / d12;
            A_(j, k - 1) = wkm1 / d12;
          }
        }
      }
    }
    if (kstep == 1) {
      ipiv[k - 1] = kp;
    } else {
      ipiv[k - 1] = -p;
      ipiv[k - 2] = -kp;
    }
    k -= kstep;
  }
  return info;
#undef A_
}

// dsytrs_rook, UPLO='U': solves A X = B in place (B is n x nrhs column-major, ldb).
static inline void sytrs_rook_upper(int n, int nrhs, const double* A, int lda, const int* ipiv, double* B, int ldb) {
#define A_(i, j) A[((i) - 1) + ((j) - 1) * lda]
#define B_(i, j) B[((i) - 1) + ((j) - 1) * ldb]
  int k = n;
  while (k >= 1) {
    if (ipiv[k - 1] > 0) {
      int kp = ipiv[k - 1];
      if (kp != k) swap_strided(nrhs, &B_(k, 1), ldb, &B_(kp, 1), ldb);
      for (int j = 1; j <= nrhs; ++j) {
        double t = -B_(k, j);
        for (int i = 1; i <= k - 1; ++i) B_(i, j) = fma(A_(i, k), t, B_(i, j));
      }
      double r = 1.0 / A_(k, k);
      for (int j = 1; j <= nrhs; ++j) B_(k, j) = B_(k, j) * r;
      k -= 1;
    } else {
      int kp = -ipiv[k - 1];
      if (kp != k) swap_strided(nrhs, &B_(k, 1), ldb, &B_(kp, 1), ldb);
      kp = -ipiv[k - 2];
      if (kp != k - 1) swap_strided(nrhs, &B_(k - 1, 1), ldb, &B_(kp, 1), ldb);
      if (k > 2) {
        for (int j = 1; j <= nrhs; ++j) {
          double t = -B_(k, j);
          for (int i = 1; i <= k - 2; ++i) B_(i, j) = fma(A_(i, k), t, B_(i, j));
        }
        for (int j = 1; j <= nrhs; ++j) {
          double t = -B_(k - 1, j);
          for (int i = 1; i <= k - 2; ++i) B_(i, j) = fma(A_(i, k - 1), t, B_(i, j));
        }
      }
      double akm1k = A_(k - 1, k);
      double akm1 = A_(k - 1, k - 1) / akm1k;
      double ak = A_(k, k) / akm1k;
      double denom = akm1 * ak - 1.0;
      for (int j = 1; j <= nrhs; ++j) {
        double bkm1 = B_(k - 1, j) / akm1k;
        double bk = B_(k, j) / akm1k;
        B_(k - 1, j) = (ak * bkm1 - bk) / denom;
        B_(k, j) = (akm1 * bk - bkm1) / denom;
      }
      k -= 2;
    }
  }
  k = 1;
  while (k <= n) {
    if (ipiv[k - 1] > 0) {
      if (k > 1)
        for (int j = 1; j <= nrhs; ++j) B_(k, j) = B_(k, j) - dot4(k - 1, &A_(1, k), 1, &B_(1, j), 1);
      int kp = ipiv[k - 1];
      if (kp != k) swap_strided(nrhs, &B_(k, 1), ldb, &B_(kp, 1), ldb);
      k += 1;
    } else {
      if (k > 1) {
        for (int j = 1; j <= nrhs; ++j) B_(k, j) = B_(k, j) - dot4(k - 1, &A_(1, k), 1, &B_(1, j), 1);
        for (int j = 1; j <= nrhs; ++j) B_(k + 1, j) = B_(k + 1, j) - dot4(k - 1, &A_(1, k + 1), 1, &B_(1, j), 1);
      }
      int kp = -ipiv[k - 1];
      if (kp != k) swap_strided(nrhs, &B_(k, 1), ldb, &B_(kp, 1), ldb);
      kp = -ipiv[k];
      if (kp != k + 1) swap_strided(nrhs, &B_(k + 1, 1), ldb, &B_(kp, 1), ldb);
      k += 2;
    }
  }
#undef A_
#undef B_
}

// Number of positive eigenvalues of the block-diagonal factor D, as `inertia!` with atol=1e-12,
// rtol=0 computes it (reference src/inertia_correction.jl:54-205; D rebuilt as in get_D! :207-255:
// diag(LD) plus, where ipiv<0, the super-diagonal entry of the 2x2 block).  Only `np` is used by
// the solver (reference src/inertia_correction.jl:265-266).
static inline int inertia_np_upper(int n, const double* A, int lda, const int* ipiv, double tol) {
#define A_(i, j) A[((i) - 1) + ((j) - 1) * lda]
  int np = 0;
  // get_D!: e[i] (1-based) holds the super-diagonal entry A(i-1,i) of a 2x2 block whose lower row is i
  double e[65];
  if (n == 0) return 0;
  for (int q = 0; q <= n; ++q) e[q] = 0.0;
  {
    int i = n;
    while (i > 1) {
      if (ipiv[i - 1] < 0) { e[i] = A_(i - 1, i); e[i - 1] = 0.0; i -= 1; }
      else e[i] = 0.0;
      i -= 1;
    }
  }
  int i = 1;
  while (i <= n) {
    int two = (i < n) && (e[i + 1] != 0.0);   // D[i,i+1] != 0
    if (two) {
      double d11 = A_(i, i), d12 = e[i + 1], d22 = A_(i + 1, i + 1);
      double a11 = fabs(d11), a22 = fabs(d22);
      double s1 = 2.0 * fmax(fmax(a11, fabs(d12)), a22);
      double smin;
      if (a11 >= a22) smin = fabs((d11 / s1) * d22 - (d12 / s1) * d12);
      else            smin = fabs(d11 * (d22 / s1) - (d12 / s1) * d12);
      double trace = d11 + d22;
      if (0.5 * s1 <= tol) {
        // both eigenvalues numerically zero
      } else if (smin > tol || trace == 0.0) {
        np += 1;            // one positive, one negative
      } else if (trace >= 0.0) {
        np += 1;            // one zero, one with the sign of the trace
      }
      i += 2;
    } else {
      if (A_(i, i) > tol) np += 1;
      i += 1;
    }
  }
  return np;
#undef A_
}
