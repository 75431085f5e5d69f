"""ORACLE (test infrastructure) -- ctypes binding of oracle/libipddp_oracle.so.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libipddp_oracle.so")
TRACE_COLS = 12
TRACE_NAMES = ["k", "j", "objective", "primal_inf", "dual_inf", "cs_inf", "mu", "reg_last", "step_size", "l",
               "theta", "barrier_lagrangian"]


class OracleOptions(C.Structure):
    _fields_ = [
        ("quasi_newton", C.c_int), ("optimality_tolerance", C.c_double), ("max_iterations", C.c_int),
        ("reset_cache", C.c_int), ("verbose", C.c_int), ("print_frequency", C.c_int),
        ("mu_init", C.c_double), ("ineq_dual_init", C.c_double), ("kappa_1", C.c_double), ("kappa_2", C.c_double),
        ("reg_1", C.c_double), ("reg_min", C.c_double), ("reg_max", C.c_double), ("kappa_bar_w_p", C.c_double),
        ("kappa_w_p", C.c_double), ("kappa_w_m", C.c_double), ("kappa_c", C.c_double), ("delta_c", C.c_double),
        ("kappa_eps", C.c_double), ("kappa_mu", C.c_double), ("theta_mu", C.c_double), ("tau_min", C.c_double),
        ("s_max", C.c_double), ("eta_L", C.c_double), ("s_L", C.c_double), ("delta", C.c_double),
        ("s_theta", C.c_double), ("gamma_alpha", C.c_double), ("gamma_theta", C.c_double), ("gamma_L", C.c_double),
        ("kappa_Sigma", C.c_double),
    ]


class OracleResult(C.Structure):
    _fields_ = [
        ("status", C.c_int), ("k", C.c_int), ("j", C.c_int), ("l", C.c_int),
        ("objective", C.c_double), ("primal_inf", C.c_double), ("dual_inf", C.c_double), ("cs_inf", C.c_double),
        ("mu", C.c_double), ("reg_last", C.c_double), ("step_size", C.c_double),
        ("barrier_lagrangian", C.c_double), ("primal_1", C.c_double),
        ("n_backward", C.c_longlong), ("n_sweeps", C.c_longlong), ("n_kkt", C.c_longlong),
        ("n_rollouts", C.c_longlong), ("n_deriv", C.c_longlong),
    ]


_lib = None


def build(force: bool = False) -> str:
    """Compile the oracle with the committed Makefile (idempotent)."""
    import glob
    import hashlib
    h = hashlib.sha256()
    for f in sorted(glob.glob(os.path.join(HERE, "*.h")) + glob.glob(os.path.join(HERE, "*.cpp")) +
                    glob.glob(os.path.join(HERE, "models_gen", "*.h")) + [os.path.join(HERE, "Makefile")]):
        with open(f, "rb") as fh:
            h.update(fh.read())
    stamp, stamp_file = h.hexdigest(), LIB_PATH + ".hash"
    stale = not os.path.exists(LIB_PATH) or not os.path.exists(stamp_file) or open(stamp_file).read() != stamp
    if force or stale:   # content hash, not mtime: checkouts reset mtimes
        subprocess.check_call(["make", "-C", HERE, "-s", "-B"])
        with open(stamp_file, "w") as fh:
            fh.write(stamp)
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
        L.oracle_default_options.argtypes = [C.POINTER(OracleOptions)]
        L.oracle_model_dims.argtypes = [C.c_char_p, ip, ip, ip, ip]
        L.oracle_create.restype = C.c_void_p
        L.oracle_create.argtypes = [C.c_char_p, C.c_int, dp, dp, dp, ip, C.c_int, C.POINTER(OracleOptions)]
        L.oracle_create_chain.restype = C.c_void_p
        L.oracle_create_chain.argtypes = [C.POINTER(C.c_char_p), C.c_int, ip, C.c_int, dp, dp, dp, ip, ip,
                                          C.POINTER(OracleOptions)]
        L.oracle_destroy.argtypes = [C.c_void_p]
        L.oracle_solve.argtypes = [C.c_void_p, dp, dp]
        L.oracle_resolve.argtypes = [C.c_void_p]
        L.oracle_get_result.argtypes = [C.c_void_p, C.POINTER(OracleResult)]
        L.oracle_trace_rows.argtypes = [C.c_void_p]
        L.oracle_get_trace.argtypes = [C.c_void_p, dp]
        L.oracle_initialize.argtypes = [C.c_void_p, dp, dp]
        L.oracle_eval_derivatives.argtypes = [C.c_void_p]
        L.oracle_backward_pass.argtypes = [C.c_void_p]
        L.oracle_errors.argtypes = [C.c_void_p, dp, dp, dp, dp]
        L.oracle_forward_pass.argtypes = [C.c_void_p]
        L.oracle_accept_step.argtypes = [C.c_void_p]
        L.oracle_set_mu.argtypes = [C.c_void_p, C.c_double]
        L.oracle_get_array.argtypes = [C.c_void_p, C.c_char_p, dp]
        L.oracle_solve_batch.argtypes = [C.c_char_p, C.c_int, C.c_int, ip, dp, dp, dp, dp, dp,
                                         C.POINTER(OracleOptions), C.c_int, C.POINTER(OracleResult), dp, dp]
        L.oracle_detmath.argtypes = [C.c_int, C.c_int, dp, dp, dp]
        L.oracle_sytf2_rook.argtypes = [C.c_int, dp, C.c_int, ip]
        L.oracle_sytrs_rook.argtypes = [C.c_int, C.c_int, dp, C.c_int, ip, dp, C.c_int]
        L.oracle_inertia_np.argtypes = [C.c_int, dp, C.c_int, ip, C.c_double]
        L.oracle_model_name.restype = C.c_char_p
        L.oracle_model_name.argtypes = [C.c_int]
        _lib = L
    return _lib


def set_filter_capacity(cap: int):
    """0 (default) = unbounded filter like the reference; n > 0 = the device's fixed capacity (status 9 on overflow)."""
    lib().oracle_set_filter_capacity(int(cap))


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def default_options(**kw) -> OracleOptions:
    o = OracleOptions()
    lib().oracle_default_options(C.byref(o))
    for k, v in kw.items():
        setattr(o, k, v)
    return o


def model_dims(model: str):
    nx, nu, nc, npar = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    if lib().oracle_model_dims(model.encode(), C.byref(nx), C.byref(nu), C.byref(nc), C.byref(npar)) != 0:
        raise KeyError(model)
    return nx.value, nu.value, nc.value, npar.value


class OracleSolver:
    """One OCP instance on the CPU oracle (mirrors reference Solver + solve!)."""

    def __init__(self, model, N, p, lower, upper, options=None, indices_compl=None):
        self.model, self.N = model, N
        self.nx, self.nu, self.nc, self.np = model_dims(model)
        p = np.ascontiguousarray(np.atleast_1d(np.asarray(p, dtype=np.float64)))
        if p.size == 0:
            p = np.zeros(1)
        self._p = p
        self._lo = np.ascontiguousarray(lower, dtype=np.float64)
        self._up = np.ascontiguousarray(upper, dtype=np.float64)
        self.options = options or default_options()
        ic = np.ascontiguousarray(indices_compl if indices_compl is not None else [], dtype=np.int32)
        self.h = lib().oracle_create(model.encode(), N, _dp(p), _dp(self._lo), _dp(self._up),
                                     _ip(ic) if ic.size else None, int(ic.size), C.byref(self.options))
        if not self.h:
            raise RuntimeError("oracle_create failed")

    def __del__(self):
        if getattr(self, "h", None):
            lib().oracle_destroy(self.h)
            self.h = None

    def solve(self, x1, ubar):
        x1 = np.ascontiguousarray(x1, dtype=np.float64)
        ubar = np.ascontiguousarray(ubar, dtype=np.float64).reshape(-1)
        assert ubar.size == (self.N - 1) * self.nu
        lib().oracle_solve(self.h, _dp(x1), _dp(ubar))
        return self.result()

    def resolve(self):
        lib().oracle_resolve(self.h)
        return self.result()

    def result(self) -> OracleResult:
        r = OracleResult()
        lib().oracle_get_result(self.h, C.byref(r))
        return r

    def trace(self) -> np.ndarray:
        n = lib().oracle_trace_rows(self.h)
        out = np.zeros((n, TRACE_COLS))
        if n:
            lib().oracle_get_trace(self.h, _dp(out))
        return out

    def initialize(self, x1, ubar):
        x1 = np.ascontiguousarray(x1, dtype=np.float64)
        ubar = np.ascontiguousarray(ubar, dtype=np.float64).reshape(-1)
        lib().oracle_initialize(self.h, _dp(x1), _dp(ubar))

    def eval_derivatives(self):
        lib().oracle_eval_derivatives(self.h)

    def backward_pass(self) -> int:
        return lib().oracle_backward_pass(self.h)

    def errors(self):
        v = [C.c_double() for _ in range(4)]
        lib().oracle_errors(self.h, *[C.byref(x) for x in v])
        return tuple(x.value for x in v)

    def forward_pass(self) -> int:
        return lib().oracle_forward_pass(self.h)

    def accept_step(self):
        lib().oracle_accept_step(self.h)

    def set_mu(self, mu):
        lib().oracle_set_mu(self.h, float(mu))

    def array(self, name: str) -> np.ndarray:
        n = lib().oracle_get_array(self.h, name.encode(), None)
        if n < 0:
            raise KeyError(name)
        out = np.zeros(n)
        if n:
            lib().oracle_get_array(self.h, name.encode(), _dp(out))
        return out


class OracleChainSolver(OracleSolver):
    """One OCP instance whose stage types change along the horizon (oracle_create_chain): `types` = registered model names,
    `stage_type[t]` for the running stages t = 0..N-2, `lowers` / `uppers` / `indices_compl` one entry per type.  `solve`
    takes the initial controls as the concatenation of the stages' (differently sized) vectors."""

    def __init__(self, types, stage_type, N, p, lowers, uppers, options=None, indices_compl=None):
        self.types, self.N = list(types), N
        self.stage_type = np.ascontiguousarray(stage_type, dtype=np.int32)
        assert self.stage_type.size == N - 1
        dims = [model_dims(t) for t in self.types]
        self.stage_nu = [dims[k][1] for k in self.stage_type]
        self.stage_nx = [dims[k][0] for k in self.stage_type]
        p = np.ascontiguousarray(np.atleast_1d(np.asarray(p, dtype=np.float64)))
        if p.size == 0:
            p = np.zeros(1)
        self._p = p
        self._lo = np.ascontiguousarray(np.concatenate([np.asarray(l, dtype=np.float64).reshape(-1) for l in lowers]))
        self._up = np.ascontiguousarray(np.concatenate([np.asarray(u, dtype=np.float64).reshape(-1) for u in uppers]))
        self.options = options or default_options()
        ic = indices_compl if indices_compl is not None else [[] for _ in self.types]
        self._ncompl = np.ascontiguousarray([len(c) for c in ic], dtype=np.int32)
        flat = np.ascontiguousarray([i for c in ic for i in c], dtype=np.int32)
        self._ic = flat if flat.size else np.zeros(1, dtype=np.int32)
        names = (C.c_char_p * len(self.types))(*[t.encode() for t in self.types])
        self.h = lib().oracle_create_chain(names, len(self.types), _ip(self.stage_type), N, _dp(p), _dp(self._lo), _dp(self._up),
                                           _ip(self._ic), _ip(self._ncompl), C.byref(self.options))
        if not self.h:
            raise RuntimeError("oracle_create_chain failed (unknown type, or stage dimensions that do not chain)")

    def solve(self, x1, ubar):
        x1 = np.ascontiguousarray(x1, dtype=np.float64)
        ubar = np.ascontiguousarray(ubar, dtype=np.float64).reshape(-1)
        assert ubar.size == sum(self.stage_nu)
        lib().oracle_solve(self.h, _dp(x1), _dp(ubar))
        return self.result()


def solve_batch(model, N, p, lower, upper, x1, ubar, options=None, horizons=None, nthreads=0, want_traj=False):
    """OpenMP batch solve.  p [B,np], lower/upper [B,nu], x1 [B,nx], ubar [B,(N-1)*nu]."""
    nx, nu, nc, npar = model_dims(model)
    B = x1.shape[0]
    p = np.ascontiguousarray(p, dtype=np.float64).reshape(B, -1)
    if p.shape[1] == 0:
        p = np.zeros((B, 1))
    lower = np.ascontiguousarray(lower, dtype=np.float64)
    upper = np.ascontiguousarray(upper, dtype=np.float64)
    x1 = np.ascontiguousarray(x1, dtype=np.float64)
    ubar = np.ascontiguousarray(ubar, dtype=np.float64).reshape(B, -1)
    options = options or default_options()
    res = (OracleResult * B)()
    xo = np.zeros((B, N, nx)) if want_traj else None
    uo = np.zeros((B, N - 1, nu)) if want_traj else None
    hz = np.ascontiguousarray(horizons, dtype=np.int32) if horizons is not None else None
    rc = lib().oracle_solve_batch(model.encode(), B, N, _ip(hz) if hz is not None else None, _dp(p), _dp(lower),
                                  _dp(upper), _dp(x1), _dp(ubar), C.byref(options), nthreads, res,
                                  _dp(xo) if want_traj else None, _dp(uo) if want_traj else None)
    if rc != 0:
        raise RuntimeError("oracle_solve_batch failed")
    return res, xo, uo


def detmath(fn: int, x, y=None):
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.ascontiguousarray(y if y is not None else np.zeros_like(x), dtype=np.float64)
    out = np.zeros_like(x)
    lib().oracle_detmath(fn, x.size, _dp(x), _dp(y), _dp(out))
    return out
