"""The reference-facing host API (Dynamics / Objective / Constraint / Bound / Options / Solver / solve /
get_trajectory, mirroring reference src/InteriorPointDDP.jl:29-45) on the GPU: user closures are traced,
emitted as CUDA, compiled with nvcc into a model plugin and solved through the C ABI.  Checked against the
reference's golden row and against the oracle's built-in restatement of the same closures."""
import math

import numpy as np
import pytest

import helpers

pytestmark = pytest.mark.gpu


def test_double_integrator_via_reference_api(oracle_mod):
    """reference experiments/ipddp2/double_integrator.jl:27-63 -> results/double_integrator.txt: 31 iterations,
    objective 1.26574863e+00, primal infeasibility 2.97128544e-09."""
    from ipddp_b200 import Dynamics, Objective, Constraint, Bound, Options, Solver, solve, get_trajectory
    dt, N = 0.01, 101
    f = lambda x, u: [x[0] + dt * x[1], x[1] + dt * u[0]]
    stage_obj = lambda x, u: dt * (u[1] + u[2])
    term_obj = lambda x, u: 500.0 * ((x[0] - 1.0) * (x[0] - 1.0) + (x[1] - 0.0) * (x[1] - 0.0))
    dyn = Dynamics(f, 2, 3)
    stage = Objective(stage_obj, 2, 3)
    path = Constraint(lambda x, u: [u[1] - u[2] - u[0] * x[1]], 2, 3)
    bound = Bound([-10.0, 0.0, 0.0], [10.0, math.inf, math.inf])
    solver = Solver(float, [dyn] * (N - 1), [stage] * (N - 1) + [Objective(term_obj, 2, 0)],
                    [path] * (N - 1) + [Constraint(2, 0)], [bound] * (N - 1) + [Bound(float, 0)],
                    options=Options(optimality_tolerance=1e-7))
    ubar = [np.array([0.01, 0.01, 0.01]) for _ in range(N - 1)] + [np.zeros(0)]
    solve(solver, np.zeros(2), ubar)
    d = solver.data
    assert d.status == 0 and d.k == 31
    assert abs(d.objective - 1.26574863e+00) < 5e-9
    assert abs(d.primal_inf - 2.97128544e-09) < 1e-12
    x_sol, u_sol = get_trajectory(solver)
    assert len(x_sol) == N and len(u_sol) == N and u_sol[-1].size == 0
    assert abs(x_sol[-1][0] - 1.0) < 0.1
    # same closures as the oracle's built-in 'double_integrator': identical iterates expected
    o = oracle_mod.OracleSolver("double_integrator", N, [], bound.lower, bound.upper,
                                options=oracle_mod.default_options(optimality_tolerance=1e-7))
    ro = o.solve(np.zeros(2), np.tile([0.01, 0.01, 0.01], N - 1))
    assert ro.k == d.k
    helpers.assert_same_bits(d.objective, ro.objective, "objective")
    helpers.assert_same_bits(np.concatenate(x_sol), o.array("x"), "states")
    # warm start: solve!(solver) (reference src/solve.jl:6-17)
    solve(solver)
    rw = o.resolve()
    assert solver.data.k == rw.k and solver.data.status == rw.status
    helpers.assert_same_bits(solver.data.objective, rw.objective, "warm-start objective")


def test_user_model_with_parameters_batched(oracle_mod):
    """A parametrised user model (concar: reference experiments/ipddp2/concar.jl:31-131) through the mirror API,
    batched over instances with per-instance parameters, bounds and initial states."""
    import sympy as sp
    from ipddp_b200 import Dynamics, Objective, Constraint, Bound, Options, Solver, solve, get_trajectory
    from ipddp_b200 import instances
    dt, r_car, N, B = 0.05, 0.02, 41, 6

    def g(x, u):
        return [x[3] * sp.cos(x[2]), x[3] * sp.sin(x[2]), u[1], u[0]]

    def f(x, u, p):
        k1 = g(x, u)
        xm = [x[i] + dt * 0.5 * k1[i] for i in range(4)]
        k2 = g(xm, u)
        return [x[i] + dt * k2[i] for i in range(4)]

    def stage(x, u, p):
        return dt * (u[0] * 5.0 * u[0] + u[1] * 1.0 * u[1]) + 50.0 * sum(u[2:6])

    def term(x, u, p):
        xN = [1.0, 1.0, math.pi / 4, 0.0]
        return 200.0 * sum((x[i] - xN[i]) * (x[i] - xN[i]) for i in range(4))

    def con(x, u, p):
        out = []
        for i in range(4):
            ox, oy, orad = p[2 + 3 * i], p[3 + 3 * i], p[4 + 3 * i]
            d = [x[0] - ox, x[1] - oy]
            out.append((orad + r_car) * (orad + r_car) - (d[0] * d[0] + d[1] * d[1]) - u[2 + i] + u[6 + i])
        return out

    b = instances.make_batch("concar", B, N)
    dyn, st, pc = Dynamics(f, 4, 10), Objective(stage, 4, 10), Constraint(con, 4, 10)
    solver = Solver(float, [dyn] * (N - 1), [st] * (N - 1) + [Objective(term, 4, 0)], [pc] * (N - 1) + [Constraint(4, 0)],
                    None, options=Options(optimality_tolerance=1e-7), batch=B, num_parameter=14)
    solve(solver, b.x1, b.ubar, params=b.p, lower=b.lower, upper=b.upper)
    res, xo, uo = oracle_mod.solve_batch("concar", N, b.p, b.lower, b.upper, b.x1, b.ubar,
                                         options=oracle_mod.default_options(optimality_tolerance=1e-7), want_traj=True)
    assert [int(k) for k in solver.data.k] == [r.k for r in res]
    assert [int(s) for s in solver.data.status] == [r.status for r in res]
    x, u = get_trajectory(solver)
    helpers.assert_same_bits(x, xo, "states")
    helpers.assert_same_bits(u, uo, "controls")


def test_indices_compl(oracle_mod):
    """constraint entries listed in indices_compl get `- mu` (reference src/data/methods.jl:27-29)."""
    from ipddp_b200 import _lib, instances
    from ipddp_b200.batch import BatchSolver
    lib = _lib.load()
    N, B = 31, 3
    b = instances.make_batch("acrobot", B, N)
    s = BatchSolver("acrobot", B, N, options=lib.default_options(optimality_tolerance=1e-7, max_iterations=60),
                    indices_compl=[4, 5], lib=lib, trace_capacity=60)
    s.set_batch(b)
    r = s.solve()
    for i in range(B):
        o = oracle_mod.OracleSolver("acrobot", N, b.p[i], b.lower[i], b.upper[i], indices_compl=[4, 5],
                                    options=oracle_mod.default_options(optimality_tolerance=1e-7, max_iterations=60))
        ro = o.solve(b.x1[i], b.ubar[i])
        assert (int(r.status[i]), int(r.k[i]), int(r.j[i])) == (ro.status, ro.k, ro.j)
        helpers.assert_same_bits(s.trace(i), o.trace(), f"trace inst {i}")
    s.close()


def test_quasi_newton_option(oracle_mod):
    """Options.quasi_newton drops the second-order contraction terms (reference src/backward_pass.jl:102)."""
    helpers_opts = dict(optimality_tolerance=1e-7, max_iterations=40, quasi_newton=1)
    from ipddp_b200 import _lib, instances
    from ipddp_b200.batch import BatchSolver
    lib = _lib.load()
    N, B = 41, 3
    b = instances.make_batch("concar", B, N)
    s = BatchSolver("concar", B, N, options=lib.default_options(**helpers_opts), lib=lib)
    s.set_batch(b)
    r = s.solve()
    res, _, _ = oracle_mod.solve_batch("concar", N, b.p, b.lower, b.upper, b.x1, b.ubar,
                                       options=oracle_mod.default_options(**helpers_opts))
    for i in range(B):
        assert (int(r.status[i]), int(r.k[i])) == (res[i].status, res[i].k)
        helpers.assert_same_bits(r.objective[i], res[i].objective, f"objective inst {i}")
    s.close()


def test_user_provided_derivative_constructors(oracle_mod):
    """Dynamics(f, fx, fu, nx', nx, nu; vfxx, vfux, vfuu) and Constraint(c, cx, cu, nc, nx, nu; vcxx, vcux, vcuu)
    (reference src/dynamics.jl:58-61, src/constraints.jl:60-64) with hand-written derivatives of the double integrator:
    same iterates as the symbolically differentiated model (= the oracle's built-in restatement)."""
    from ipddp_b200 import Dynamics, Objective, Constraint, Bound, Options, Solver, solve, get_trajectory
    dt, N = 0.01, 101
    f = lambda x, u: [x[0] + dt * x[1], x[1] + dt * u[0]]
    fx = lambda x, u: [[1.0, dt], [0.0, 1.0]]
    fu = lambda x, u: [[0.0, 0.0, 0.0], [dt, 0.0, 0.0]]
    zero = lambda r, c: [[0.0] * c for _ in range(r)]
    dyn = Dynamics(f, fx, fu, 2, 2, 3, vfxx=lambda x, u, v: zero(2, 2), vfux=lambda x, u, v: zero(3, 2),
                   vfuu=lambda x, u, v: zero(3, 3))
    c = lambda x, u: [u[1] - u[2] - u[0] * x[1]]
    cx = lambda x, u: [[0.0, -u[0]]]
    cu = lambda x, u: [[-x[1], 1.0, -1.0]]
    path = Constraint(c, cx, cu, 1, 2, 3, vcxx=lambda x, u, v: zero(2, 2),
                      vcux=lambda x, u, v: [[0.0, -v[0]], [0.0, 0.0], [0.0, 0.0]], vcuu=lambda x, u, v: zero(3, 3))
    stage = Objective(lambda x, u: dt * (u[1] + u[2]), 2, 3)
    term = Objective(lambda x, u: 500.0 * ((x[0] - 1.0) * (x[0] - 1.0) + (x[1] - 0.0) * (x[1] - 0.0)), 2, 0)
    bound = Bound([-10.0, 0.0, 0.0], [10.0, math.inf, math.inf])
    solver = Solver(float, [dyn] * (N - 1), [stage] * (N - 1) + [term], [path] * (N - 1) + [Constraint(2, 0)],
                    [bound] * (N - 1) + [Bound(float, 0)], options=Options(optimality_tolerance=1e-7))
    ubar = [np.array([0.01, 0.01, 0.01]) for _ in range(N - 1)] + [np.zeros(0)]
    solve(solver, np.zeros(2), ubar)
    d = solver.data
    assert d.status == 0 and d.k == 31 and abs(d.objective - 1.26574863e+00) < 5e-9
    o = oracle_mod.OracleSolver("double_integrator", N, [], bound.lower, bound.upper,
                                options=oracle_mod.default_options(optimality_tolerance=1e-7))
    ro = o.solve(np.zeros(2), np.tile([0.01, 0.01, 0.01], N - 1))
    assert ro.k == d.k
    helpers.assert_same_bits(d.objective, ro.objective, "objective")
    x_sol, _ = get_trajectory(solver)
    helpers.assert_same_bits(np.concatenate(x_sol), o.array("x"), "states")


def test_solver_with_stage_sizes_that_change_along_the_horizon(oracle_mod):
    """The reference's constructor call with one Dynamics / Objective / Constraint / Bound per stage whose state and
    control sizes change mid-way (README.md:18): Solver groups the stages into stage types, compiles a chain model and
    reproduces the oracle's per-stage solve bit for bit."""
    from ipddp_b200 import Dynamics, Objective, Constraint, Bound, Solver, solve, get_trajectory
    from ipddp_b200.codegen import workloads
    ch = workloads.get_chain("ragged")
    N = 41
    st = ch.stage_types(N)
    dyn = [Dynamics(lambda x, u, md=md: md.f(x, u, []), md.nx, md.nu) for md in ch.stages]
    obj = [Objective(lambda x, u, md=md: md.stage_cost(x, u, []), md.nx, md.nu) for md in ch.stages]
    con = [Constraint(lambda x, u, md=md: md.c(x, u, []), md.nx, md.nu) if md.nc > 0 else Constraint(md.nx, md.nu) for md in ch.stages]
    bnd = [Bound(np.array(md.lower([])), np.array(md.upper([]))) for md in ch.stages]
    termo = Objective(lambda x, u: ch.stages[-1].term_cost(x, []), 3, 0)
    solver = Solver(float, [dyn[k] for k in st], [obj[k] for k in st] + [termo], [con[k] for k in st] + [Constraint(3, 0)],
                    [bnd[k] for k in st] + [Bound(float, 0)], options=None)
    assert solver._bs.nstage == 3 and solver.stage_type == st
    ubar = [np.asarray(ch.stages[k].u_init) for k in st] + [np.zeros(0)]
    data = solve(solver, np.zeros(2), ubar)
    xs, us = get_trajectory(solver)
    o = oracle_mod.OracleChainSolver([md.name for md in ch.stages], st, N, [], [md.lower([]) for md in ch.stages],
                                     [md.upper([]) for md in ch.stages])
    res = o.solve(np.zeros(2), np.concatenate(ubar))
    assert (int(data.status), int(data.k), int(data.j)) == (res.status, res.k, res.j)
    helpers.assert_same_bits(data.objective, res.objective, "objective")
    assert [len(x) for x in xs] == [ch.stages[k].nx for k in st] + [3] and [len(u) for u in us[:-1]] == [ch.stages[k].nu for k in st]
    helpers.assert_same_bits(np.concatenate(xs), o.array("x"), "ragged states")
    helpers.assert_same_bits(np.concatenate(us), o.array("u"), "ragged controls")
