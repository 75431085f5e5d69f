"""Host-side mirror of the reference's exported Julia API (reference src/InteriorPointDDP.jl:29-45):

    Dynamics, Objective, Constraint, Bound, Options, Solver, solve (= solve!), get_trajectory

Same names, argument meaning and result fields as the reference, extended additively for batching
(per-instance x1, controls, parameter vector p, horizon).  Where the reference traces user closures with
Symbolics.jl and `eval`s generated Julia (src/dynamics.jl:15-47, src/objectives.jl:12-33,
src/constraints.jl:16-50), this traces them with SymPy, emits CUDA device functions (codegen/generate.py),
compiles them with nvcc into a model plugin and registers it through `ipddp_model_load`.  The solve itself
always runs on the GPU through the C ABI (no CPU fallback).

Example (reference experiments/ipddp2/double_integrator.jl:27-63):

    f = lambda x, u: [x[0] + dt * x[1], x[1] + dt * u[0]]
    dynamics = [Dynamics(f, 2, 3) for k in range(N - 1)]
    objective = [Objective(lambda x, u: dt * (u[1] + u[2]), 2, 3)] * (N - 1) + [Objective(term, 2, 0)]
    constraints = [Constraint(lambda x, u: [u[1] - u[2] - u[0] * x[1]], 2, 3)] * (N - 1) + [Constraint(2, 0)]
    bounds = [Bound([-10, 0, 0], [10, inf, inf])] * (N - 1) + [Bound(float, 0)]
    solver = Solver(float, dynamics, objective, constraints, bounds, options=Options(optimality_tolerance=1e-7))
    solve(solver, x1, ubar)
    x_sol, u_sol = get_trajectory(solver)
    solver.data.k, solver.data.status, solver.data.objective
"""
from __future__ import annotations

import hashlib
import inspect
import os
import subprocess
from dataclasses import dataclass, field
from typing import Callable, List, Optional, Sequence

import numpy as np

from . import _lib
from .batch import BatchSolver
from .codegen import generate, workloads

INF = float("inf")
HERE = os.path.dirname(os.path.abspath(__file__))
PLUGIN_DIR = os.path.join(HERE, "_plugins")


def _arity(fn) -> int:
    return len([p for p in inspect.signature(fn).parameters.values()
                if p.default is inspect.Parameter.empty and p.kind in (p.POSITIONAL_ONLY, p.POSITIONAL_OR_KEYWORD)])


def _with_p(fn, nargs_without_p):
    """Accept closures written without the runtime parameter vector (the reference's signature)."""
    if fn is None:
        return None
    return fn if _arity(fn) > nargs_without_p else (lambda *a: fn(*a[:nargs_without_p]))


def _inplace(fn, nargs_without_p):
    """Adapter for a closure in the reference's in-place form `fn!(out, args...[, p])` (src/dynamics.jl:49-61,
    src/constraints.jl:60-64): the tracer hands it a symbolic buffer of the right shape as `out`."""
    if fn is None:
        return None
    g = _with_p(fn, nargs_without_p + 1)
    h = lambda *a: g(*a)
    h._ipddp_inplace = True
    return h


def _filled(fn, n):
    """value-returning view (x, u, p) -> list of an in-place vector closure fn!(out[n], x, u, p)"""
    import sympy as sp

    def g(x, u, p):
        out = sp.zeros(n, 1)
        fn(out, x, u, p)
        return list(out)
    return g


# ------------------------------------------------------------------------------------------------
# constructors (reference src/dynamics.jl:15, src/objectives.jl:12, src/constraints.jl:16,52, src/bounds.jl:12-26)
# ------------------------------------------------------------------------------------------------
class Dynamics:
    """Dynamics(f, num_state, num_control; quasi_newton=false): x+ = f(x, u[, p])   (reference src/dynamics.jl:15)

    User-provided derivatives (reference src/dynamics.jl:58-61):
    Dynamics(f, fx, fu, num_next_state, num_state, num_control; vfxx=None, vfux=None, vfuu=None) -- fx, fu are closures
    (x, u[, p]) -> matrix; the contractions (x, u, v[, p]) -> matrix are optional and stay zero when omitted, as in the
    reference.  The closures are traced symbolically into CUDA device functions, so by default they return their value;
    `inplace=True` takes the reference's own in-place form instead: `f(y, x, u)`, `fx(J, x, u)`, `vfxx(H, x, u, v)` fill
    a (symbolic) output buffer."""

    def __init__(self, f: Callable, *args, quasi_newton: bool = False, vfxx=None, vfux=None, vfuu=None,
                 inplace: bool = False):
        self.f = _with_p(f, 2)
        self._src = f
        self.user_derivs = {}
        if len(args) == 5 and callable(args[0]) and callable(args[1]):
            fx, fu, num_next_state, num_state, num_control = args
            wrap = _inplace if inplace else _with_p
            if inplace:
                self.f = _filled(_with_p(f, 3), int(num_next_state))
            self.user_derivs = {"fx": wrap(fx, 2), "fu": wrap(fu, 2)}
            for k, fn in (("vfxx", vfxx), ("vfux", vfux), ("vfuu", vfuu)):
                if fn is not None:
                    self.user_derivs[k] = wrap(fn, 3)
            self._extra_src = [fx, fu, vfxx, vfux, vfuu]
        else:
            num_state, num_control = args
            self._extra_src = []
        self.num_state, self.num_control, self.quasi_newton = int(num_state), int(num_control), bool(quasi_newton)


class Objective:
    """Objective(f, num_state, num_control): stage cost l(x, u[, p]); num_control = 0 for the terminal stage."""

    def __init__(self, f: Callable, num_state: int, num_control: int):
        self.f = _with_p(f, 2)
        self.num_state, self.num_control = int(num_state), int(num_control)
        self._src = f


class Constraint:
    """Constraint(c, num_state, num_control; quasi_newton=false, indices_compl=nothing) or the empty
    Constraint(num_state, num_control)."""

    def __init__(self, *args, quasi_newton: bool = False, indices_compl: Optional[Sequence[int]] = None,
                 vcxx=None, vcux=None, vcuu=None, inplace: bool = False):
        self.user_derivs = {}
        self._extra_src = []
        if len(args) == 6 and callable(args[0]) and callable(args[1]):
            # user-provided derivatives (reference src/constraints.jl:60-64):
            # Constraint(c, cx, cu, num_constraint, num_state, num_control; indices_compl, vcxx, vcux, vcuu)
            # inplace=True: the reference's in-place closures c(out, x, u), cx(J, x, u), vcxx(H, x, u, v)
            c, cx, cu, num_constraint, nx, nu = args
            wrap = _inplace if inplace else _with_p
            self.c = _filled(_with_p(c, 3), int(num_constraint)) if inplace else _with_p(c, 2)
            self._src = c
            self.user_derivs = {"cx": wrap(cx, 2), "cu": wrap(cu, 2)}
            for k, fn in (("vcxx", vcxx), ("vcux", vcux), ("vcuu", vcuu)):
                if fn is not None:
                    self.user_derivs[k] = wrap(fn, 3)
            self._extra_src = [cx, cu, vcxx, vcux, vcuu]
        elif callable(args[0]):
            c, nx, nu = args
            self.c = _with_p(c, 2)
            self._src = c
        else:
            nx, nu = args
            self.c = None
            self._src = None
        self.num_state, self.num_control = int(nx), int(nu)
        self.quasi_newton = bool(quasi_newton)
        # reference indices are 1-based (Julia); here 0-based
        self.indices_compl = list(indices_compl) if indices_compl is not None else []


class Bound:
    """Bound(lower, upper) | Bound(T, num_control) (unbounded) | Bound(num_control, lower, upper)."""

    def __init__(self, *args):
        if len(args) == 2 and isinstance(args[0], type):
            n = int(args[1])
            self.lower, self.upper = np.full(n, -INF), np.full(n, INF)
        elif len(args) == 3:
            n, lo, up = args
            self.lower, self.upper = np.full(int(n), float(lo)), np.full(int(n), float(up))
        else:
            lo, up = args
            self.lower = np.asarray(lo, dtype=np.float64).reshape(-1)
            self.upper = np.asarray(up, dtype=np.float64).reshape(-1)
        assert self.lower.shape == self.upper.shape
        self.indices_lower = [i for i, v in enumerate(self.lower) if not np.isinf(v)]
        self.indices_upper = [i for i, v in enumerate(self.upper) if not np.isinf(v)]
        self.num_lower, self.num_upper = len(self.indices_lower), len(self.indices_upper)


def Options(**kw) -> _lib.Options:
    """Options{T}(; kwargs...) (reference src/options.jl:1-38).  Greek field names map to ASCII:
    mu_init, kappa_1, kappa_2, kappa_bar_w_p, kappa_w_p, kappa_w_m, kappa_c, delta_c, kappa_eps, kappa_mu,
    theta_mu, tau_min, eta_L, s_L, delta, s_theta, gamma_alpha, gamma_theta, gamma_L, kappa_Sigma."""
    return _lib.load().default_options(**kw)


# ------------------------------------------------------------------------------------------------
# model plugin cache: closures -> CUDA -> nvcc -> .so -> ipddp_model_load
# ------------------------------------------------------------------------------------------------
_loaded_models = {}   # (library id, model name) -> digest of the source the registered model was built from


def _copy_options(o: _lib.Options) -> _lib.Options:
    c = _lib.Options()
    for fname, _ in _lib.Options._fields_:
        setattr(c, fname, getattr(o, fname))
    return c


def model_digest(md: workloads.ModelDef, bundles=None) -> str:
    """sha256 over the emitted device source with the model name neutralised, the dimensions and indices_compl."""
    saved = md.name
    md.name = "M"
    try:
        src = generate.emit_device(md, bundles if bundles is not None else generate.trace(md))
    finally:
        md.name = saved
    key = f"{src}|{md.nx}|{md.nu}|{md.np_}|{sorted(md.indices_compl)}"
    return hashlib.sha256(key.encode()).hexdigest()[:12]


def _compile_plugin(name: str, src: str, force: bool = False) -> str:
    """Compiles an emitted model header for sm_100a into `<name>_<hash>.so` exporting `ipddp_plugin_vtable` (three lines
    of CUDA around the header).  The plugin embeds the kernel templates and the DevView layout, so the
    cache key covers them and the compiler flags too."""
    os.makedirs(PLUGIN_DIR, exist_ok=True)
    src = src.replace('#include "../model_common.cuh"', f'#include "{os.path.join(HERE, "csrc", "model_common.cuh")}"')
    from . import build as _b
    tag = hashlib.sha256((src + _b.content_hash(_b._headers(), " ".join(_b.FLAGS))).encode()).hexdigest()[:12]
    so = os.path.join(PLUGIN_DIR, f"{name}_{tag}.so")
    if os.path.exists(so) and not force:
        return so
    cuh = os.path.join(PLUGIN_DIR, f"{name}_{tag}.cuh")
    cu = os.path.join(PLUGIN_DIR, f"{name}_{tag}.cu")
    with open(cuh, "w") as fh:
        fh.write(src)
    with open(cu, "w") as fh:
        fh.write(f'#include "{cuh}"\n#include "{os.path.join(HERE, "csrc", "model_register.cuh")}"\n'
                 f'IPDDP_REGISTER_MODEL(Model_{name}, ipddp_plugin_vtable)\n')
    r = subprocess.run([_b.NVCC] + _b.FLAGS + ["-shared", cu, "-o", so], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for model plugin:\n" + r.stdout + r.stderr)
    return so


def build_model_plugin(md: workloads.ModelDef, force: bool = False, bundles=None) -> str:
    """Generates the device code of `md` and compiles it into a plugin."""
    bundles = bundles if bundles is not None else generate.trace(md)
    return _compile_plugin(md.name, generate.emit_device(md, bundles), force)


def build_chain_plugin(chain, all_bundles, force: bool = False) -> str:
    """Plugin of a stage chain (several stage types, sizes that change along the horizon): one header with the stage
    structs and the composite model (generate.emit_device_chain)."""
    src = generate.emit_device_chain(chain, [generate.emit_device(md, b) for md, b in zip(chain.stages, all_bundles)])
    return _compile_plugin(chain.name, src, force)


@dataclass
class SolverData:
    """Mirror of reference SolverData (src/data/solver.jl:8-33) after a solve; arrays of length B
    (scalars when the solver was built for a single instance)."""
    status: object = 0
    k: object = 0
    j: object = 0
    l: object = 0
    objective: object = 0.0
    primal_inf: object = 0.0
    dual_inf: object = 0.0
    cs_inf: object = 0.0
    mu: object = 0.0
    reg_last: object = 0.0
    step_size: object = 0.0
    wall_time: float = 0.0       # seconds (device time of the whole batch solve)
    solver_time: float = 0.0     # wall_time minus derivative-evaluation time
    fn_eval_time: float = 0.0


class Solver:
    """Solver(T, dynamics, objectives, constraints, bounds=nothing; options=nothing)  (reference src/solver.jl:11-26)

    One Dynamics / Objective / Constraint / Bound per stage, as in the reference; the state and control sizes may change
    along the horizon (reference README.md:18, src/data/problem.jl:44-62).  Stages built from the same objects form one
    stage TYPE (one set of compiled device functions); a horizon may use up to 4 types.  The terminal stage has
    num_control = 0 and no constraints.  Batched extension: `batch` instances share the model; `num_parameter` runtime
    parameters per instance are the trailing argument `p` of the closures; per-instance horizons <= N are available when
    all running stages are of one type."""

    def __init__(self, T, dynamics: List[Dynamics], objectives: List[Objective], constraints: List[Constraint],
                 bounds: Optional[List[Bound]] = None, options: Optional[_lib.Options] = None, batch: int = 1,
                 num_parameter: int = 0, device: int = 0, name: Optional[str] = None, trace_capacity: int = 0):
        assert T in (float, np.float64), "FP64 only (the reference experiments are all Float64)"
        N = len(objectives)
        assert len(dynamics) + 1 == N == len(constraints), "need N-1 dynamics, N objectives, N constraints"
        oN, cN = objectives[-1], constraints[-1]
        assert oN.num_control == 0 and cN.c is None, "terminal stage must have num_control = 0 and no constraints"
        if bounds is None:
            bounds = [Bound(float, d.num_control) for d in dynamics] + [Bound(float, 0)]
        assert len(bounds) == N
        # ---- stage types: running stages built from the same closures (and equal bounds) share one type
        keys, types, stage_type = [], [], []
        for t in range(N - 1):
            d, o, c, bd = dynamics[t], objectives[t], constraints[t], bounds[t]
            key = (id(d._src), id(o._src), id(c._src), d.quasi_newton, c.quasi_newton, tuple(c.indices_compl),
                   bd.lower.tobytes(), bd.upper.tobytes())
            if key not in keys:
                keys.append(key)
                types.append((d, o, c, bd))
            stage_type.append(keys.index(key))
        if len(types) > 4:
            raise NotImplementedError("at most 4 distinct stage types per horizon")
        last = stage_type[-1]           # the type of the last running stage carries the terminal cost: it goes last
        order = [k for k in range(len(types)) if k != last] + [last]
        types = [types[k] for k in order]
        stage_type = [order.index(k) for k in stage_type]
        self.stage_type = stage_type
        self.N, self.batch, self.num_parameter = N, int(batch), int(num_parameter)
        term_f = _with_p(oN._src, 2)
        mds = []
        for k, (d, o, c, bd) in enumerate(types):
            cfun = c.c if c.c is not None else (lambda x, u, p: [])
            is_last = k == len(types) - 1
            mds.append(workloads.ModelDef(
                name=f"pending_s{k}", nx=d.num_state, nu=d.num_control, np_=self.num_parameter, f=d.f, stage_cost=o.f,
                term_cost=(lambda x, p: term_f(x, [], p)) if is_last else (lambda x, p: 0.0 * x[0]), c=cfun,
                lower=(lambda p, bd=bd: list(bd.lower)), upper=(lambda p, bd=bd: list(bd.upper)),
                u_init=[0.0] * d.num_control, dt=0.0, indices_compl=list(c.indices_compl),
                user_derivs={**d.user_derivs, **c.user_derivs}, user_dynamics=bool(d.user_derivs),
                user_constraint=bool(c.user_derivs), qn_dynamics=d.quasi_newton, qn_constraint=c.quasi_newton,
                nx_term=oN.num_state if is_last else None))
        self.stage_defs = mds
        self.bounds_by_type = [bd for (_, _, _, bd) in types]
        self.bound = self.bounds_by_type[0]
        self.lib = _lib.load()
        # The model's identity is the device code that was traced from the closures (plus its dimensions): globals and
        # closure cells the closures read are folded into that code at trace time, so two Solvers share a compiled model
        # exactly when their emitted sources agree.  A user-chosen `name` is bound to the source it was first built from
        # in this process; the same name with different code is rebuilt and re-registered (ipddp_model_load replaces it).
        all_bundles = [generate.trace(md) for md in mds]
        for md, bd in zip(mds, all_bundles):      # the kernels' limits, before any compiler is started
            n_c = bd["con"].outputs[0][1]
            if md.nu < 1:
                raise ValueError("a running stage needs at least one control")
            if md.nu + n_c > 64:
                raise ValueError(f"num_control + num_constraint = {md.nu + n_c} > 64: the warp-level KKT factorisation holds at "
                                 "most 64 rows (two lane slots per column)")
        chain = len(mds) > 1 or all_bundles[0]["dyn"].outputs[0][1] != mds[0].nx or (oN.num_state != mds[0].nx)
        digest = hashlib.sha256("|".join(model_digest(md, bd) for md, bd in zip(mds, all_bundles)).encode()).hexdigest()[:12]
        tag = name or ("user_" + digest)
        if chain:
            for k, md in enumerate(mds):
                md.name = f"{tag}_s{k}"
            self.model_def = workloads.ChainDef(tag, mds)
        else:
            mds[0].name = tag
            self.model_def = mds[0]
        if _loaded_models.get((id(self.lib), tag)) != digest:
            so = build_chain_plugin(self.model_def, all_bundles) if chain else build_model_plugin(mds[0], bundles=all_bundles[0])
            self.lib.check(self.lib.L.ipddp_model_load(so.encode()), "ipddp_model_load")
            _loaded_models[(id(self.lib), tag)] = digest
        # per-object quasi_newton only removes that object's contractions (ModelDef.qn_*); Options.quasi_newton is the
        # user's and is neither forced nor mutated here
        self.options = _copy_options(options) if options is not None else self.lib.default_options()
        self.quasi_newton = bool(self.options.quasi_newton)
        self._bs = BatchSolver(tag, self.batch, N, options=self.options, device=device, trace_capacity=trace_capacity,
                               indices_compl=mds[0].indices_compl, lib=self.lib)
        self._bs.set_stage_types(stage_type, [md.indices_compl for md in mds])
        self.nx, self.nu, self.nc = self._bs.nx, self._bs.nu, self._bs.nc      # maxima over the stage types
        self.data = SolverData()

    @classmethod
    def from_workload(cls, workload: str, batch: int, N: int = 101, options=None, device: int = 0, trace_capacity: int = 0):
        """One of the reference's six benchmark classes (built into the library)."""
        self = cls.__new__(cls)
        self.lib = _lib.load()
        md = workloads.get(workload)
        self.model_def = md
        self.N, self.nx, self.nu, self.batch, self.num_parameter = N, md.nx, md.nu, int(batch), md.np_
        self.options = options if options is not None else self.lib.default_options()
        self._bs = BatchSolver(workload, self.batch, N, options=self.options, device=device, trace_capacity=trace_capacity,
                               lib=self.lib)
        self.nc = self._bs.nc
        self.bound = None
        self.bounds_by_type = None
        self.stage_type = [0] * (N - 1)
        self.quasi_newton = False
        self.data = SolverData()
        return self

    # --------------------------------------------------------------------------------------------
    def _fill_data(self, r):
        st = self._bs.stats()
        one = self.batch == 1
        g = (lambda a: a[0].item()) if one else (lambda a: a)
        self.data = SolverData(status=g(r.status), k=g(r.k), j=g(r.j), l=g(r.l), objective=g(r.objective),
                               primal_inf=g(r.primal_inf), dual_inf=g(r.dual_inf), cs_inf=g(r.cs_inf), mu=g(r.mu),
                               reg_last=g(r.reg_last), step_size=g(r.step_size), wall_time=st.ms_total * 1e-3,
                               solver_time=(st.ms_total - st.ms_derivs) * 1e-3, fn_eval_time=st.ms_derivs * 1e-3)
        return self.data


def solve(solver: Solver, x1=None, controls=None, params=None, lower=None, upper=None, horizons=None):
    """solve!(solver, x1, controls) (reference src/solve.jl:1-4); without x1/controls: solve!(solver)
    (warm start from the stored nominal trajectory, src/solve.jl:6-17).

    x1: [nx] or [B, nx]; controls: list of N per-stage vectors (last one empty), or [N-1, nu], or [B, N-1, nu];
    params: [np] or [B, np]; lower/upper: per-instance bounds [B, nu] (default: the Solver's Bound)."""
    B, N, nx, nu = solver.batch, solver.N, solver.nx, solver.nu
    if x1 is None and controls is None:
        r = solver._bs.solve(warm_start=True)
        return solver._fill_data(r)
    x1 = np.asarray(x1, dtype=np.float64)
    if x1.ndim == 1 and x1.size < nx:       # stage chain: the first stage may have fewer states than the largest one
        x1 = np.concatenate([x1, np.zeros(nx - x1.size)])
    x1 = np.broadcast_to(x1.reshape(-1, nx), (B, nx))
    if isinstance(controls, (list, tuple)) and len(controls) == N and np.size(controls[-1]) == 0:
        # the reference's list of per-stage vectors (possibly of different lengths): padded to the largest control size
        padded = np.zeros((N - 1, nu))
        for t in range(N - 1):
            c = np.asarray(controls[t], dtype=np.float64).reshape(-1)
            padded[t, :c.size] = c
        controls = padded
    controls = np.asarray(controls, dtype=np.float64)
    controls = np.broadcast_to(controls.reshape(-1, (N - 1) * nu), (B, (N - 1) * nu))
    p = None
    if solver.num_parameter > 0:
        p = np.broadcast_to(np.asarray(params, dtype=np.float64).reshape(-1, solver.num_parameter), (B, solver.num_parameter))
    if solver.bounds_by_type is not None:   # one bound vector per stage type, padded with +-inf
        def pad(vecs, fill):
            out = np.full((len(vecs), nu), fill)
            for k, v in enumerate(vecs):
                out[k, :v.size] = v
            return out.reshape(-1)
        if lower is None:
            lower = pad([bd.lower for bd in solver.bounds_by_type], -np.inf)
        if upper is None:
            upper = pad([bd.upper for bd in solver.bounds_by_type], np.inf)
    solver._bs.set_inputs(x1, controls, p, lower, upper, horizons)
    r = solver._bs.solve()
    return solver._fill_data(r)


def get_trajectory(solver: Solver):
    """get_trajectory(solver) -> (nominal_states, nominal_controls) (reference src/solver.jl:46-48).
    Single instance: lists of per-stage vectors like the reference; batched: arrays [B,N,nx], [B,N-1,nu]."""
    x, u = solver._bs.trajectory()
    if solver.batch == 1:
        nxs, nus, _ = solver._bs.stage_layout()
        return ([x[0, t, :nxs[t]] for t in range(solver.N)],
                [u[0, t, :nus[t]] for t in range(solver.N - 1)] + [np.zeros(0)])
    return x, u
