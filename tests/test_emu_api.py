"""The reference-facing host API (Dynamics / Objective / Constraint / Bound / Options / Solver / solve / get_trajectory,
api.py) end to end WITHOUT a GPU: the test bodies of tests/test_gpu_api.py run against the SIMT emulator -- the same
tracing, the same emitted model headers, the same C ABI calls; only the compiler of the plugin (g++ against
tests/emu/cpu_simt.h instead of nvcc) and the library behind `_lib.load()` differ (tests/emu/emu_plugins.py, test
infrastructure).  The GPU suite runs the same bodies on the B200."""
import os
import sys

import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu"))

import test_gpu_api as cases  # noqa: E402  (its pytestmark only applies to tests collected from that module)


@pytest.fixture()
def emu_api(monkeypatch):
    import emu_plugins
    return emu_plugins.install(monkeypatch)


def test_emulated_double_integrator_via_reference_api(emu_api, oracle_mod):
    cases.test_double_integrator_via_reference_api(oracle_mod)


def test_emulated_user_provided_derivative_constructors(emu_api, oracle_mod):
    cases.test_user_provided_derivative_constructors(oracle_mod)


def test_emulated_solver_with_stage_sizes_that_change_along_the_horizon(emu_api, oracle_mod):
    cases.test_solver_with_stage_sizes_that_change_along_the_horizon(oracle_mod)


def test_emulated_user_model_with_parameters_batched(emu_api, oracle_mod):
    cases.test_user_model_with_parameters_batched(oracle_mod)


def test_emulated_indices_compl(emu_api, oracle_mod):
    cases.test_indices_compl(oracle_mod)


def test_emulated_quasi_newton_option(emu_api, oracle_mod):
    cases.test_quasi_newton_option(oracle_mod)


def test_emulated_inplace_user_closures_solve_like_value_closures(emu_api, oracle_mod):
    """The reference's in-place user-derivative closures (`f!(y, x, u)`, `fx!(J, x, u)`, ...: src/dynamics.jl:49-61,
    src/constraints.jl:60-64) behind `inplace=True`: same model digest, same solve (the golden double-integrator row)
    as the value-returning closures of test_user_provided_derivative_constructors."""
    import math
    import numpy as np
    from ipddp_b200 import Dynamics, Objective, Constraint, Bound, Options, Solver, solve, get_trajectory
    import helpers
    dt, N = 0.01, 101

    def f(y, x, u):
        y[0] = x[0] + dt * x[1]
        y[1] = x[1] + dt * u[0]

    def fx(J, x, u):
        J[0, 0] = 1.0; J[0, 1] = dt; J[1, 1] = 1.0

    def fu(J, x, u):
        J[1, 0] = dt

    def c(out, x, u):
        out[0] = u[1] - u[2] - u[0] * x[1]

    def cx(J, x, u):
        J[0, 1] = -u[0]

    def cu(J, x, u):
        J[0, 0] = -x[1]; J[0, 1] = 1.0; J[0, 2] = -1.0

    def vcux(H, x, u, v):
        H[0, 1] = -v[0]

    dyn = Dynamics(f, fx, fu, 2, 2, 3, inplace=True)
    path = Constraint(c, cx, cu, 1, 2, 3, vcux=vcux, inplace=True)
    stage = Objective(lambda x, u: dt * (u[1] + u[2]), 2, 3)
    term = Objective(lambda x, u: 500.0 * ((x[0] - 1.0) * (x[0] - 1.0) + (x[1] - 0.0) * (x[1] - 0.0)), 2, 0)
    bound = Bound([-10.0, 0.0, 0.0], [10.0, math.inf, math.inf])
    solver = Solver(float, [dyn] * (N - 1), [stage] * (N - 1) + [term], [path] * (N - 1) + [Constraint(2, 0)],
                    [bound] * (N - 1) + [Bound(float, 0)], options=Options(optimality_tolerance=1e-7))
    solve(solver, np.zeros(2), [np.array([0.01, 0.01, 0.01]) for _ in range(N - 1)] + [np.zeros(0)])
    d = solver.data
    assert d.status == 0 and d.k == 31 and abs(d.objective - 1.26574863e+00) < 5e-9
    o = oracle_mod.OracleSolver("double_integrator", N, [], bound.lower, bound.upper,
                                options=oracle_mod.default_options(optimality_tolerance=1e-7))
    ro = o.solve(np.zeros(2), np.tile([0.01, 0.01, 0.01], N - 1))
    assert ro.k == d.k
    helpers.assert_same_bits(d.objective, ro.objective, "objective")
    x_sol, _ = get_trajectory(solver)
    helpers.assert_same_bits(np.concatenate(x_sol), o.array("x"), "states")


def test_emulated_experiment_harness_reproduces_the_reference_table(emu_api, tmp_path):
    """tools/run_experiments.py (one batched launch per class, tables in the reference's file format) on the
    emulator: the double-integrator table comes out with the reference's own text in the seed / iterations / status /
    objective columns and its primal infeasibility (experiments/ipddp2/results/double_integrator.txt; the timing columns
    are this machine's), and parses with the reference's regexes.  (The 100-row classes take the emulator minutes per row;
    the B200 regenerates all six tables: profiles/r2_experiments.)"""
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.join(os.path.dirname(here), "tools"))
    import run_experiments
    from ipddp_b200 import results_io
    fname = "double_integrator"
    summ = run_experiments.run_class(emu_api, "double_integrator", fname, str(tmp_path))
    assert summ["same_both"] == summ["instances"] == summ["converged_gpu"] == 1
    got = open(tmp_path / (fname + ".txt")).read().splitlines()
    want = open(os.path.join(here, "golden", "results", fname + ".txt")).read().splitlines()
    assert got[0] == want[0] and len(got) == len(want) == 2
    ca, cb = got[1].split(), want[1].split()
    assert ca[:5] == cb[:5], (got[1], want[1])
    assert len(results_io.read_results(str(tmp_path / (fname + ".txt")))) == 1


def test_emulated_fresh_objects_per_stage_are_one_stage_type(emu_api):
    """`[Dynamics(f, nx, nu) for k = 1:N-1]` (reference experiments/ipddp2/pushing_1_obs.jl:98,104: a new object per stage
    around the same closure) must give the same single stage type as one shared object (cartpole_friction.jl:53)."""
    import math
    import numpy as np
    from ipddp_b200 import Dynamics, Objective, Constraint, Bound, Options, Solver
    dt, N = 0.01, 101
    f = lambda x, u: [x[0] + dt * x[1], x[1] + dt * u[0]]
    stage_obj = lambda x, u: dt * (u[1] + u[2])
    term_obj = lambda x, u: 500.0 * ((x[0] - 1.0) * (x[0] - 1.0) + (x[1] - 0.0) * (x[1] - 0.0))
    con = lambda x, u: [u[1] - u[2] - u[0] * x[1]]
    solver = Solver(float, [Dynamics(f, 2, 3) for _ in range(N - 1)],
                    [Objective(stage_obj, 2, 3) for _ in range(N - 1)] + [Objective(term_obj, 2, 0)],
                    [Constraint(con, 2, 3) for _ in range(N - 1)] + [Constraint(2, 0)],
                    [Bound([-10.0, 0.0, 0.0], [10.0, math.inf, math.inf]) for _ in range(N - 1)] + [Bound(float, 0)],
                    options=Options(optimality_tolerance=1e-7))
    assert solver._bs.nstage == 1 and set(solver.stage_type) == {0}


def test_solver_refuses_models_beyond_the_kernels_limits(emu_api):
    """num_control + num_constraint > 64 is refused with a plain message before any compiler runs (the C ABI would refuse it
    too: ipddp_problem_create)."""
    import math
    from ipddp_b200 import Dynamics, Objective, Constraint, Bound, Solver
    nu, N = 65, 3
    dyn = Dynamics(lambda x, u: [x[0] + 0.1 * u[0]], 1, nu)
    with pytest.raises(ValueError, match="> 64"):
        Solver(float, [dyn] * (N - 1), [Objective(lambda x, u: sum(ui * ui for ui in u), 1, nu)] * (N - 1)
               + [Objective(lambda x, u: x[0] * x[0], 1, 0)], [Constraint(1, nu)] * (N - 1) + [Constraint(1, 0)],
               [Bound(float, nu)] * (N - 1) + [Bound(float, 0)])
