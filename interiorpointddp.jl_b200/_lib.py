"""ctypes binding of the C ABI declared in include/ipddp_b200.h.

`load()` returns the CUDA product library (libipddp_b200.so, built in-tree by build.py).  There is no CPU
fallback: if the library is missing or no CUDA device is usable, calls fail loudly.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libipddp_b200.so")
TRACE_COLS = 12
TRACE_NAMES = ["k", "j", "objective", "primal_inf", "dual_inf", "cs_inf", "mu", "reg_last", "step_size", "l",
               "theta", "barrier_lagrangian"]

EXPORTS = [
    "ipddp_abi_version", "ipddp_last_error", "ipddp_default_options", "ipddp_num_models", "ipddp_model_name",
    "ipddp_model_dims", "ipddp_model_load", "ipddp_problem_create", "ipddp_problem_destroy", "ipddp_set_options",
    "ipddp_set_tuning", "ipddp_layout", "ipddp_set_inputs", "ipddp_set_inputs_device", "ipddp_solve", "ipddp_solve_many", "ipddp_solve_queue", "ipddp_set_stream", "ipddp_model_stages", "ipddp_set_stage_types",
    "ipddp_set_stage_compl", "ipddp_stage_layout",
    "ipddp_initialize",
    "ipddp_eval_derivatives", "ipddp_backward_pass", "ipddp_check", "ipddp_forward_pass", "ipddp_get_results",
    "ipddp_get_trajectory", "ipddp_get_duals", "ipddp_get_counters", "ipddp_get_array", "ipddp_get_trace",
    "ipddp_get_stats", "ipddp_stream", "ipddp_measure_fp64_tflops", "ipddp_measure_hbm_gbs",
    "ipddp_test_detmath", "ipddp_test_ldlt",
]


class Options(C.Structure):
    """Mirror of reference Options{T} (src/options.jl:1-38), C layout of `ipddp_options`."""
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


class Stats(C.Structure):
    _fields_ = [
        ("iterations", C.c_int), ("launches", C.c_longlong), ("ms_total", C.c_double), ("ms_init", C.c_double),
        ("ms_derivs", C.c_double), ("ms_backward", C.c_double), ("ms_check", C.c_double), ("ms_forward", C.c_double),
        ("sum_backward", C.c_longlong), ("sum_sweeps", C.c_longlong), ("sum_kkt", C.c_longlong),
        ("sum_rollouts", C.c_longlong), ("sum_deriv_stages", C.c_longlong), ("n_converged", C.c_longlong),
        ("n_active_rounds", C.c_longlong), ("sum_active_sq", C.c_double),
    ]


class Queue(C.Structure):
    """C layout of `ipddp_queue` (ipddp_solve_queue): Q queued instances, input and output arrays (host or device)."""
    _fields_ = [
        ("Q", C.c_int), ("x1", C.c_void_p), ("ubar", C.c_void_p), ("params", C.c_void_p), ("lower", C.c_void_p),
        ("upper", C.c_void_p), ("horizons", C.c_void_p), ("inputs_on_device", C.c_int),
        ("status", C.c_void_p), ("k", C.c_void_p), ("j", C.c_void_p), ("l", C.c_void_p),
        ("objective", C.c_void_p), ("primal_inf", C.c_void_p), ("dual_inf", C.c_void_p), ("cs_inf", C.c_void_p),
        ("mu", C.c_void_p), ("reg_last", C.c_void_p), ("step_size", C.c_void_p),
        ("n_backward", C.c_void_p), ("n_sweeps", C.c_void_p), ("n_kkt", C.c_void_p), ("n_rollouts", C.c_void_p),
        ("x", C.c_void_p), ("u", C.c_void_p), ("outputs_on_device", C.c_int),
    ]


class Lib:
    """Typed handle on one shared library exporting the ipddp_* C ABI."""

    def __init__(self, path: str):
        if not os.path.exists(path):
            raise FileNotFoundError(
                f"{path} not found: build it with `python interiorpointddp.jl_b200/build.py` "
                "(there is no CPU fallback)")
        self.path = path
        L = C.CDLL(path)
        self.L = L
        dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
        vp = C.c_void_p
        L.ipddp_last_error.restype = C.c_char_p
        L.ipddp_default_options.argtypes = [C.POINTER(Options)]
        L.ipddp_model_name.restype = C.c_char_p
        L.ipddp_model_name.argtypes = [C.c_int]
        L.ipddp_model_dims.argtypes = [C.c_char_p, ip, ip, ip, ip, ip]
        L.ipddp_model_load.argtypes = [C.c_char_p]
        L.ipddp_problem_create.argtypes = [C.c_char_p, C.c_int, C.c_int, ip, C.c_int, C.POINTER(Options), C.c_int,
                                           C.c_int, C.POINTER(vp)]
        L.ipddp_problem_destroy.argtypes = [vp]
        L.ipddp_set_options.argtypes = [vp, C.POINTER(Options)]
        L.ipddp_set_tuning.argtypes = [vp, C.c_char_p, C.c_int]
        L.ipddp_layout.argtypes = [vp, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong), C.POINTER(C.c_longlong),
                                   C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)]
        L.ipddp_set_inputs.argtypes = [vp, dp, dp, dp, dp, dp, ip]
        L.ipddp_set_inputs_device.argtypes = [vp, vp, vp, vp, vp, vp, vp]
        L.ipddp_solve.argtypes = [vp, C.c_int]
        L.ipddp_solve_queue.argtypes = [vp, C.POINTER(Queue)]
        L.ipddp_set_stream.argtypes = [vp, vp]
        L.ipddp_model_stages.argtypes = [C.c_char_p, ip, ip, ip, ip, ip, ip]
        L.ipddp_set_stage_types.argtypes = [vp, ip]
        L.ipddp_set_stage_compl.argtypes = [vp, C.c_int, ip, C.c_int]
        L.ipddp_stage_layout.argtypes = [vp, ip, ip, ip]
        L.ipddp_solve_many.argtypes = [C.POINTER(vp), C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(Stats)]
        for f in ("ipddp_initialize", "ipddp_eval_derivatives", "ipddp_backward_pass", "ipddp_forward_pass"):
            getattr(L, f).argtypes = [vp]
        L.ipddp_check.argtypes = [vp, ip]
        L.ipddp_get_results.argtypes = [vp, ip, ip, ip, ip, dp, dp, dp, dp, dp, dp, dp]
        L.ipddp_get_trajectory.argtypes = [vp, dp, dp]
        L.ipddp_get_duals.argtypes = [vp, dp, dp, dp, dp]
        L.ipddp_get_counters.argtypes = [vp, ip, ip, ip, ip]
        L.ipddp_get_array.restype = C.c_longlong
        L.ipddp_get_array.argtypes = [vp, C.c_char_p, dp]
        L.ipddp_get_trace.argtypes = [vp, C.c_int, dp, ip]
        L.ipddp_get_stats.argtypes = [vp, C.POINTER(Stats)]
        L.ipddp_stream.restype = vp
        L.ipddp_stream.argtypes = [vp]
        L.ipddp_measure_fp64_tflops.restype = C.c_double
        L.ipddp_measure_fp64_tflops.argtypes = [C.c_int]
        L.ipddp_measure_hbm_gbs.restype = C.c_double
        L.ipddp_measure_hbm_gbs.argtypes = [C.c_int]
        L.ipddp_test_detmath.argtypes = [C.c_int, C.c_int, dp, dp, dp, C.c_int]
        L.ipddp_test_ldlt.argtypes = [C.c_int, C.c_int, dp, dp, dp, ip, ip, ip, dp, C.c_int]

    def check(self, rc, what=""):
        if rc != 0:
            raise RuntimeError(f"{what} failed: {self.L.ipddp_last_error().decode()}")

    def default_options(self, **kw) -> Options:
        o = Options()
        self.L.ipddp_default_options(C.byref(o))
        for k, v in kw.items():
            if not hasattr(o, k):
                raise AttributeError(f"Options has no field {k}")
            setattr(o, k, v)
        return o

    def models(self):
        return [self.L.ipddp_model_name(i).decode() for i in range(self.L.ipddp_num_models())]

    def model_stages(self, model: str):
        """(nstage, [(nx, nu, nc, nxn) per stage type], nxt) of a model; a plain model has one stage type."""
        n, nxt = C.c_int(), C.c_int()
        a = [(C.c_int * 4)() for _ in range(4)]
        self.check(self.L.ipddp_model_stages(model.encode(), C.byref(n), a[0], a[1], a[2], a[3], C.byref(nxt)), "ipddp_model_stages")
        return n.value, [tuple(int(x[k]) for x in a) for k in range(n.value)], nxt.value

    def model_dims(self, model: str):
        v = [C.c_int() for _ in range(5)]
        self.check(self.L.ipddp_model_dims(model.encode(), *[C.byref(x) for x in v]), "ipddp_model_dims")
        return tuple(x.value for x in v)


_product = None


def load() -> Lib:
    """The CUDA product library.  Raises if it has not been built."""
    global _product
    if _product is None:
        _product = Lib(LIB_PATH)
    return _product


def dptr(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def iptr(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_int))
