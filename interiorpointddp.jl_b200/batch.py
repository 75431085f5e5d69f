"""BatchSolver: the batched counterpart of the reference's `Solver` + `solve!` over the C ABI.

One BatchSolver owns one `ipddp_problem` handle (B instances of one model, up to N knots, one GPU).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional

import numpy as np

from . import _lib
from ._lib import Options, Stats, dptr, iptr


@dataclass
class BatchResult:
    """Per-instance SolverData fields (reference src/data/solver.jl:8-33)."""
    status: np.ndarray
    k: np.ndarray
    j: np.ndarray
    l: np.ndarray
    objective: np.ndarray
    primal_inf: np.ndarray
    dual_inf: np.ndarray
    cs_inf: np.ndarray
    mu: np.ndarray
    reg_last: np.ndarray
    step_size: np.ndarray

    @property
    def converged(self) -> np.ndarray:
        return self.status == 0


class BatchSolver:
    def __init__(self, model: str, B: int, N: int, options: Optional[Options] = None, device: int = 0,
                 trace_capacity: int = 0, indices_compl=None, lib: Optional[_lib.Lib] = None):
        self.lib = lib or _lib.load()
        self.model, self.B, self.N = model, int(B), int(N)
        self.nx, self.nu, self.nc, self.np, self.tile_slots = self.lib.model_dims(model)
        # stage chains: nx / nu / nc are the maxima over the stage types (array strides); ns = stride of the state outputs
        self.nstage, self.stages, self.nxt = self.lib.model_stages(model)
        self.ns = max([self.nxt] + [max(st[0], st[3]) for st in self.stages])
        self.options = options or self.lib.default_options()
        ic = np.ascontiguousarray(indices_compl if indices_compl is not None else [], dtype=np.int32)
        h = C.c_void_p()
        self.lib.check(self.lib.L.ipddp_problem_create(model.encode(), self.B, self.N, iptr(ic) if ic.size else None,
                                                       int(ic.size), C.byref(self.options), int(device),
                                                       int(trace_capacity), C.byref(h)), "ipddp_problem_create")
        self.h = h
        self._keep = None

    def close(self):
        if getattr(self, "h", None):
            self.lib.L.ipddp_problem_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ inputs
    def set_inputs(self, x1, ubar, params=None, lower=None, upper=None, horizons=None):
        """Host buffers (numpy), copied H2D.  Shapes: x1 [B,nx], ubar [B,(N-1)*nu], params [B,np],
        lower/upper [B,nu] (default: unbounded), horizons [B] int."""
        B, N = self.B, self.N
        x1 = np.ascontiguousarray(x1, dtype=np.float64).reshape(B, self.nx)
        ubar = np.ascontiguousarray(ubar, dtype=np.float64).reshape(B, (N - 1) * self.nu)
        nb = self.nstage * self.nu          # bounds per instance: one vector per stage type
        if lower is None:
            lower = np.full((B, nb), -np.inf)
        if upper is None:
            upper = np.full((B, nb), np.inf)
        lower = np.ascontiguousarray(np.broadcast_to(np.asarray(lower, dtype=np.float64).reshape(-1, nb), (B, nb)))
        upper = np.ascontiguousarray(np.broadcast_to(np.asarray(upper, dtype=np.float64).reshape(-1, nb), (B, nb)))
        p = None
        if self.np > 0:
            p = np.ascontiguousarray(params, dtype=np.float64).reshape(B, self.np)
        hz = None
        if horizons is not None:
            hz = np.ascontiguousarray(horizons, dtype=np.int32).reshape(B)
        self._keep = (x1, ubar, p, lower, upper, hz)
        self.lib.check(self.lib.L.ipddp_set_inputs(self.h, dptr(x1), dptr(ubar), dptr(p) if p is not None else None,
                                                   dptr(lower), dptr(upper), iptr(hz) if hz is not None else None),
                       "ipddp_set_inputs")

    def set_inputs_device(self, x1_ptr, ubar_ptr, params_ptr, lower_ptr, upper_ptr, horizons_ptr=None):
        """Raw device pointers (ints), e.g. torch.Tensor.data_ptr(); D2D copy."""
        self.lib.check(self.lib.L.ipddp_set_inputs_device(self.h, x1_ptr, ubar_ptr, params_ptr, lower_ptr, upper_ptr,
                                                          horizons_ptr), "ipddp_set_inputs_device")

    def set_batch(self, batch):
        self.set_inputs(batch.x1, batch.ubar, batch.p if self.np > 0 else None, batch.lower, batch.upper,
                        batch.horizons)

    def set_stage_types(self, stage_types, indices_compl=None):
        """Stage chains: the stage type of every running stage t = 0..N-2 (ipddp_set_stage_types) and, optionally, the
        indices_compl of each stage type (list of lists, ipddp_set_stage_compl)."""
        st = np.ascontiguousarray(stage_types, dtype=np.int32).reshape(self.N - 1)
        self.lib.check(self.lib.L.ipddp_set_stage_types(self.h, iptr(st)), "ipddp_set_stage_types")
        for k, ic in enumerate(indices_compl or []):
            a = np.ascontiguousarray(ic, dtype=np.int32)
            self.lib.check(self.lib.L.ipddp_set_stage_compl(self.h, k, iptr(a) if a.size else None, int(a.size)),
                           "ipddp_set_stage_compl")

    def stage_layout(self):
        """per-knot sizes (nx[N], nu[N], nc[N]) of the problem (ipddp_stage_layout)"""
        a = [np.zeros(self.N, dtype=np.int32) for _ in range(3)]
        self.lib.check(self.lib.L.ipddp_stage_layout(self.h, *[iptr(x) for x in a]), "ipddp_stage_layout")
        return tuple(a)

    def set_tuning(self, key: str, value: int):
        """Execution tuning that never changes results (see ipddp_set_tuning)."""
        self.lib.check(self.lib.L.ipddp_set_tuning(self.h, key.encode(), int(value)), "ipddp_set_tuning")

    def set_stream(self, cuda_stream: int):
        """Launch on the caller's CUDA stream (raw cudaStream_t, e.g. torch.cuda.Stream().cuda_stream); 0 = own stream."""
        self.lib.check(self.lib.L.ipddp_set_stream(self.h, C.c_void_p(int(cuda_stream) or None)), "ipddp_set_stream")

    # ------------------------------------------------------------------ solve and phases
    def solve(self, warm_start: bool = False) -> BatchResult:
        self.lib.check(self.lib.L.ipddp_solve(self.h, int(warm_start)), "ipddp_solve")
        return self.results()

    def solve_queue(self, x1, ubar, params=None, lower=None, upper=None, horizons=None, want_traj=True):
        """Streaming solve of Q >= 1 queued instances through this solver's B slots (ipddp_solve_queue); host arrays.
        Returns (BatchResult over the Q instances, counters dict, x [Q,N,nx] or None, u [Q,N-1,nu] or None)."""
        N = self.N
        x1 = np.ascontiguousarray(x1, dtype=np.float64).reshape(-1, self.nx)
        Q = x1.shape[0]
        ubar = np.ascontiguousarray(ubar, dtype=np.float64).reshape(Q, (N - 1) * self.nu)
        nb = self.nstage * self.nu
        lower = np.full((Q, nb), -np.inf) if lower is None else lower
        upper = np.full((Q, nb), np.inf) if upper is None else upper
        lower = np.ascontiguousarray(np.broadcast_to(np.asarray(lower, dtype=np.float64).reshape(-1, nb), (Q, nb)))
        upper = np.ascontiguousarray(np.broadcast_to(np.asarray(upper, dtype=np.float64).reshape(-1, nb), (Q, nb)))
        p = np.ascontiguousarray(params, dtype=np.float64).reshape(Q, self.np) if self.np > 0 else None
        hz = np.ascontiguousarray(horizons, dtype=np.int32).reshape(Q) if horizons is not None else None
        ints = [np.zeros(Q, dtype=np.int32) for _ in range(4)]
        dbl = [np.zeros(Q) for _ in range(7)]
        cnt = [np.zeros(Q, dtype=np.int32) for _ in range(4)]
        x = np.zeros((Q, N, self.ns)) if want_traj else None
        u = np.zeros((Q, N - 1, self.nu)) if want_traj else None
        q = make_queue(Q, x1, ubar, p, lower, upper, hz, ints + dbl + cnt, x, u)
        self.lib.check(self.lib.L.ipddp_solve_queue(self.h, C.byref(q)), "ipddp_solve_queue")
        return (BatchResult(*ints, *dbl), dict(n_backward=cnt[0], n_sweeps=cnt[1], n_kkt=cnt[2], n_rollouts=cnt[3]), x, u)

    def initialize(self):
        self.lib.check(self.lib.L.ipddp_initialize(self.h), "ipddp_initialize")

    def eval_derivatives(self):
        self.lib.check(self.lib.L.ipddp_eval_derivatives(self.h), "ipddp_eval_derivatives")

    def backward_pass(self):
        self.lib.check(self.lib.L.ipddp_backward_pass(self.h), "ipddp_backward_pass")

    def check(self) -> int:
        n = C.c_int()
        self.lib.check(self.lib.L.ipddp_check(self.h, C.byref(n)), "ipddp_check")
        return n.value

    def forward_pass(self):
        self.lib.check(self.lib.L.ipddp_forward_pass(self.h), "ipddp_forward_pass")

    # ------------------------------------------------------------------ outputs
    def results(self) -> BatchResult:
        B = self.B
        ints = [np.zeros(B, dtype=np.int32) for _ in range(4)]
        dbl = [np.zeros(B) for _ in range(7)]
        self.lib.check(self.lib.L.ipddp_get_results(self.h, *[iptr(a) for a in ints], *[dptr(a) for a in dbl]),
                       "ipddp_get_results")
        return BatchResult(*ints, *dbl)

    def trajectory(self):
        """get_trajectory(solver): nominal states [B,N,nx] and controls [B,N-1,nu]."""
        x = np.zeros((self.B, self.N, self.ns))
        u = np.zeros((self.B, self.N - 1, self.nu))
        self.lib.check(self.lib.L.ipddp_get_trajectory(self.h, dptr(x), dptr(u)), "ipddp_get_trajectory")
        return x, u

    def duals(self):
        phi = np.zeros((self.B, self.N - 1, self.nc))
        zl = np.zeros((self.B, self.N - 1, self.nu))
        zu = np.zeros((self.B, self.N - 1, self.nu))
        lam = np.zeros((self.B, self.N, self.ns))
        self.lib.check(self.lib.L.ipddp_get_duals(self.h, dptr(phi), dptr(zl), dptr(zu), dptr(lam)), "ipddp_get_duals")
        return phi, zl, zu, lam

    def counters(self):
        a = [np.zeros(self.B, dtype=np.int32) for _ in range(4)]
        self.lib.check(self.lib.L.ipddp_get_counters(self.h, *[iptr(x) for x in a]), "ipddp_get_counters")
        return dict(n_backward=a[0], n_sweeps=a[1], n_kkt=a[2], n_rollouts=a[3])

    def array(self, name: str) -> np.ndarray:
        n = self.lib.L.ipddp_get_array(self.h, name.encode(), None)
        if n < 0:
            raise KeyError(f"{name}: {self.lib.L.ipddp_last_error().decode()}")
        out = np.zeros(int(n))
        if n and self.lib.L.ipddp_get_array(self.h, name.encode(), dptr(out)) < 0:
            raise RuntimeError(f"ipddp_get_array({name}) failed: {self.lib.L.ipddp_last_error().decode()}")
        return out

    def trace(self, b: int) -> np.ndarray:
        n = C.c_int()
        self.lib.check(self.lib.L.ipddp_get_trace(self.h, b, None, C.byref(n)), "ipddp_get_trace")
        out = np.zeros((n.value, _lib.TRACE_COLS))
        if n.value:
            self.lib.check(self.lib.L.ipddp_get_trace(self.h, b, dptr(out), C.byref(n)), "ipddp_get_trace")
        return out

    def stats(self) -> Stats:
        st = Stats()
        self.lib.check(self.lib.L.ipddp_get_stats(self.h, C.byref(st)), "ipddp_get_stats")
        return st


def _addr(a):
    if a is None:
        return None
    return int(a) if isinstance(a, (int, np.integer)) else a.ctypes.data


def make_queue(Q, x1, ubar, params, lower, upper, horizons, scalars, x, u, inputs_on_device=False,
               outputs_on_device=False) -> _lib.Queue:
    """Fill an `ipddp_queue`: arrays are numpy arrays (host) or raw device addresses (ints); `scalars` is the list of the
    15 per-instance output arrays in the struct's order (status k j l | objective primal_inf dual_inf cs_inf mu reg_last
    step_size | n_backward n_sweeps n_kkt n_rollouts), entries may be None."""
    q = _lib.Queue()
    q.Q = int(Q)
    q.x1, q.ubar, q.params, q.lower, q.upper, q.horizons = (_addr(a) for a in (x1, ubar, params, lower, upper, horizons))
    q.inputs_on_device = int(inputs_on_device)
    names = ["status", "k", "j", "l", "objective", "primal_inf", "dual_inf", "cs_inf", "mu", "reg_last", "step_size",
             "n_backward", "n_sweeps", "n_kkt", "n_rollouts"]
    for nm, a in zip(names, scalars):
        setattr(q, nm, _addr(a))
    q.x, q.u = _addr(x), _addr(u)
    q.outputs_on_device = int(outputs_on_device)
    return q


def solve_many(solvers, total_solves=None, warm_start=False):
    """Concurrent solve of several BatchSolvers (one device, one stream each), see ipddp_solve_many.
    Returns (elapsed_ms, Stats summed over all solves)."""
    lib = solvers[0].lib
    n = len(solvers)
    arr = (C.c_void_p * n)(*[s.h for s in solvers])
    ms = C.c_double()
    st = Stats()
    lib.check(lib.L.ipddp_solve_many(arr, n, int(total_solves or n), int(warm_start), C.byref(ms), C.byref(st)),
              "ipddp_solve_many")
    return ms.value, st
