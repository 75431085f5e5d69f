"""Pins the oracle's restatement of dsytf2_rook / dsytrs_rook / inertia against a real LAPACK binary:
the OpenBLAS bundled with SciPy (scipy_dsytrf_rook_, scipy_dsytrs_rook_).  The reference reaches the
same routines through Julia's OpenBLAS_jll (reference src/inertia_correction.jl:261, src/backward_pass.jl:148).
Factors, ipiv and info must be bit-identical; the solve agrees to rounding (OpenBLAS's dger/dgemv
kernels have their own summation order)."""
import ctypes as C
import glob
import os

import numpy as np
import pytest


def _openblas():
    import scipy
    base = os.path.join(os.path.dirname(os.path.dirname(scipy.__file__)), "scipy.libs")
    libs = glob.glob(os.path.join(base, "libscipy_openblas*.so*"))
    if not libs:
        pytest.skip("no bundled OpenBLAS")
    return C.CDLL(libs[0])


def _lapack_sytrf_rook(L, A):
    n = A.shape[0]
    a = np.asfortranarray(A.copy())
    ipiv = np.zeros(max(n, 1), dtype=np.int32)
    work = np.zeros(max(1, 64 * n))
    info = C.c_int(0)
    uplo = C.c_char(b"U")
    nn, lda, lwork = C.c_int(n), C.c_int(max(n, 1)), C.c_int(work.size)
    L.scipy_dsytrf_rook_(C.byref(uplo), C.byref(nn), a.ctypes.data_as(C.c_void_p), C.byref(lda),
                         ipiv.ctypes.data_as(C.c_void_p), work.ctypes.data_as(C.c_void_p), C.byref(lwork),
                         C.byref(info), C.c_size_t(1))
    return a, ipiv, info.value


def _lapack_sytrs_rook(L, a, ipiv, Bm):
    n, nrhs = Bm.shape
    b = np.asfortranarray(Bm.copy())
    info = C.c_int(0)
    uplo = C.c_char(b"U")
    nn, nr, lda, ldb = C.c_int(n), C.c_int(nrhs), C.c_int(n), C.c_int(n)
    L.scipy_dsytrs_rook_(C.byref(uplo), C.byref(nn), C.byref(nr), a.ctypes.data_as(C.c_void_p), C.byref(lda),
                         ipiv.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p), C.byref(ldb),
                         C.byref(info), C.c_size_t(1))
    return b


def _oracle_factor(oracle, A):
    n = A.shape[0]
    a = np.asfortranarray(A.copy())
    ipiv = np.zeros(n + 1, dtype=np.int32)
    info = oracle.lib().oracle_sytf2_rook(n, a.ctypes.data_as(C.POINTER(C.c_double)), n,
                                          ipiv.ctypes.data_as(C.POINTER(C.c_int)))
    return a, ipiv[:n], info


def _matrices(rng, count):
    out = []
    for i in range(count):
        kind = i % 4
        if kind == 0:      # dense symmetric indefinite
            n = int(rng.integers(1, 36))
            M = rng.standard_normal((n, n)); M = M + M.T
        elif kind == 1:    # KKT [H A'; A 0] with sparse / rank-deficient A
            m = int(rng.integers(2, 22)); p = int(rng.integers(1, 15)); n = m + p
            H = rng.standard_normal((m, m)); H = H @ H.T + np.diag(rng.uniform(0, 5, m))
            A = rng.standard_normal((p, m)) * (rng.uniform(size=(p, m)) < 0.3)
            if i % 8 == 1 and p > 1:
                A[-1] = A[0]
            M = np.zeros((n, n)); M[:m, :m] = H; M[:m, m:] = A.T; M[m:, :m] = A
        elif kind == 2:    # sparse with zero diagonal entries
            n = int(rng.integers(2, 30))
            M = rng.standard_normal((n, n)) * (rng.uniform(size=(n, n)) < 0.25); M = M + M.T
            np.fill_diagonal(M, 0.0)
        else:              # badly scaled
            n = int(rng.integers(2, 36))
            M = rng.standard_normal((n, n)); M = M + M.T
            s = 10.0 ** rng.uniform(-6, 6, n); M = M * s[:, None] * s[None, :]
        out.append(M)
    return out


def test_factor_bit_identical_to_openblas(oracle_mod):
    L = _openblas()
    rng = np.random.default_rng(7)
    n2x2 = 0
    for M in _matrices(rng, 600):
        a_ref, ip_ref, info_ref = _lapack_sytrf_rook(L, M)
        a_or, ip_or, info_or = _oracle_factor(oracle_mod, M)
        n = M.shape[0]
        assert info_or == info_ref
        assert np.array_equal(ip_or, ip_ref[:n])
        iu = np.triu_indices(n)
        assert np.array_equal(a_or[iu].view(np.int64), a_ref[iu].view(np.int64)), "factor bits differ"
        n2x2 += int((ip_ref[:n] < 0).any())
    assert n2x2 > 100   # the 2x2-pivot path is exercised


def test_solve_matches_openblas(oracle_mod):
    L = _openblas()
    rng = np.random.default_rng(11)
    for i in range(200):
        m = int(rng.integers(2, 22)); p = int(rng.integers(1, min(m, 14) + 1)); n = m + p
        H = rng.standard_normal((m, m)); H = H @ H.T + np.eye(m)
        A = rng.standard_normal((p, m))
        M = np.zeros((n, n)); M[:m, :m] = H; M[:m, m:] = A.T; M[m:, :m] = A
        Bm = rng.standard_normal((n, 5))
        a_ref, ip_ref, info = _lapack_sytrf_rook(L, M)
        assert info == 0
        x_ref = _lapack_sytrs_rook(L, a_ref, ip_ref, Bm)
        a_or, ip_or, _ = _oracle_factor(oracle_mod, M)
        b = np.asfortranarray(Bm.copy())
        ipc = np.ascontiguousarray(ip_or, dtype=np.int32)
        oracle_mod.lib().oracle_sytrs_rook(n, 5, a_or.ctypes.data_as(C.POINTER(C.c_double)), n,
                                           ipc.ctypes.data_as(C.POINTER(C.c_int)),
                                           b.ctypes.data_as(C.POINTER(C.c_double)), n)
        assert np.allclose(b, x_ref, rtol=1e-9, atol=1e-11)
        assert np.allclose(M @ b, Bm, rtol=1e-8, atol=1e-8)


def test_inertia_counts_match_eigenvalues(oracle_mod):
    rng = np.random.default_rng(5)
    for i in range(300):
        m = int(rng.integers(1, 22)); p = int(rng.integers(0, 15)); n = m + p
        H = rng.standard_normal((m, m)); H = H + H.T + (2.0 * (i % 3)) * np.eye(m)
        A = rng.standard_normal((p, m))
        M = np.zeros((n, n)); M[:m, :m] = H; M[:m, m:] = A.T; M[m:, :m] = A
        a_or, ip_or, info = _oracle_factor(oracle_mod, M)
        if info != 0:
            continue
        ipc = np.ascontiguousarray(ip_or, dtype=np.int32)
        npos = oracle_mod.lib().oracle_inertia_np(n, a_or.ctypes.data_as(C.POINTER(C.c_double)), n,
                                                  ipc.ctypes.data_as(C.POINTER(C.c_int)), 1e-12)
        ev = np.linalg.eigvalsh(M)
        if np.min(np.abs(ev)) < 1e-8:
            continue
        assert npos == int((ev > 0).sum())
