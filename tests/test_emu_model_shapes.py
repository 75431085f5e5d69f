"""Model shapes beyond the reference's experiments (all of which have 2 or 4 states, 3 to 21 controls, 1 to 14 constraints),
through the whole solve on the emulator against a scratch build of the oracle that carries the same traced closures
(tests/emu/user_model_harness.py) -- bit for bit, speculative and bulk kernels:

  wide64           K = nu + nc = 64, the C ABI's limit: KKT assembly beyond one lane slot per column, the two-slot pivot
                   steps of the warp LDL^T on real KKT matrices, the gains / record strides
  single_control   nu = 1 (a pendulum has one torque): the gains' index arithmetic divides by nu and by K
  one_by_one       nx = 1, nu = 1, no constraints: 1 x 1 KKT matrices, 6-double records
  no_constraints   nc = 0 with one- and two-sided bounds
  state16          more than 7 states: more right-hand sides per KKT system than the warp's 8 column groups
  k36_state9       both at once (K = 36, 9 states): the grouped second solve with rows beyond 32
Every model also at the shortest horizon (2 knots: one running stage and the terminal stage).
"""
import os
import sys

import numpy as np
import pytest

import helpers

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu"))
INF = float("inf")
DT = 0.05


def wide_model(name, nf, extra=0):
    """nx = 4, nu = 4 nf + extra, nc = 2 nf (K = 6 nf + extra): forces u[0:nf] = s+ - s- (L1 split: c[0:nf], the first
    `extra` of them with one more non-negative slack each), bilinear couplings x[i % 4] * f_i + 0.05 = t_i with slack t
    (c[nf:2nf])."""
    import sympy as sp
    from ipddp_b200.codegen import workloads
    nu, dt = 4 * nf + extra, DT

    def f(x, u, p):
        fs = u[0:nf]
        a = sum(fs[0::2]) - 0.5 * sum(fs[1::2])
        b = 0.3 * sum(((-1) ** i) * fs[i] for i in range(nf))
        return [x[0] + dt * x[2], x[1] + dt * x[3], x[2] + dt * (a - 0.2 * x[0] * x[0]), x[3] + dt * (b - sp.sin(x[1]))]

    def stage(x, u, p):
        return dt * (sum(0.05 * u[i] * u[i] for i in range(nf)) + sum(u[nf:3 * nf]) + 2.0 * sum(u[3 * nf:4 * nf])
                     + 3.0 * sum(u[4 * nf:nu]))

    def term(x, p):
        tgt = [0.5, -0.25, 0.0, 0.0]
        return 50.0 * sum((x[i] - tgt[i]) * (x[i] - tgt[i]) for i in range(4))

    def c(x, u, p):
        out = [u[i] - u[nf + i] + u[2 * nf + i] + (u[4 * nf + i] if i < extra else 0) for i in range(nf)]
        out += [x[i % 4] * u[i] + 0.05 - u[3 * nf + i] for i in range(nf)]
        return out
    return workloads.ModelDef(name=name, nx=4, nu=nu, np_=0, f=f, stage_cost=stage, term_cost=term, c=c,
                              lower=lambda p: [-5.0] * nf + [0.0] * (nu - nf), upper=lambda p: [5.0] * nf + [INF] * (nu - nf),
                              u_init=[0.0] * nf + [0.01] * (nu - nf), dt=dt)


def small_models():
    import sympy as sp
    from ipddp_b200.codegen import workloads
    dt = DT
    M = workloads.ModelDef
    single = M(name="single_control", nx=2, nu=1, np_=0,
               f=lambda x, u, p: [x[0] + dt * x[1], x[1] + dt * (u[0] - x[0] * x[0] * x[0])],
               stage_cost=lambda x, u, p: dt * (u[0] * u[0]), term_cost=lambda x, p: 10.0 * ((x[0] - 0.7) ** 2 + x[1] ** 2),
               c=lambda x, u, p: [0.01 * u[0] * x[1]], lower=lambda p: [-3.0], upper=lambda p: [3.0], u_init=[0.1], dt=dt)
    one = M(name="one_by_one", nx=1, nu=1, np_=0, f=lambda x, u, p: [x[0] + dt * (u[0] - x[0] * x[0] * x[0])],
            stage_cost=lambda x, u, p: dt * (u[0] * u[0]), term_cost=lambda x, p: 10.0 * (x[0] - 0.7) ** 2,
            c=lambda x, u, p: [], lower=lambda p: [-INF], upper=lambda p: [1.5], u_init=[0.1], dt=dt)
    nocon = M(name="no_constraints", nx=3, nu=2, np_=0,
              f=lambda x, u, p: [x[0] + dt * x[1], x[1] + dt * (u[0] - sp.sin(x[0])), x[2] + dt * (u[1] * x[0])],
              stage_cost=lambda x, u, p: dt * (u[0] * u[0] + 0.5 * u[1] * u[1] + 0.1 * x[2] * x[2]),
              term_cost=lambda x, p: 20.0 * ((x[0] - 1.0) ** 2 + x[1] ** 2 + x[2] ** 2), c=lambda x, u, p: [],
              lower=lambda p: [-2.0, -INF], upper=lambda p: [2.0, 3.0], u_init=[0.0, 0.0], dt=dt)

    def f12(x, u, p):
        out = [x[i] + dt * x[6 + i] for i in range(6)]
        out += [x[6 + i] + dt * (p[0] * u[i] - p[1] * sp.sin(x[i]) - 0.1 * x[6 + i] + 0.05 * x[(i + 1) % 6] * u[(i + 2) % 6])
                for i in range(6)]
        return out
    s12 = M(name="state12", nx=12, nu=9, np_=2, f=f12,
            stage_cost=lambda x, u, p: dt * (sum(0.1 * u[i] * u[i] for i in range(6)) + sum(u[6:9])
                                             + 0.01 * sum(x[i] * x[i] for i in range(12))),
            term_cost=lambda x, p: 30.0 * sum((x[i] - 0.3 * ((-1) ** i)) ** 2 for i in range(6)) + sum(x[6 + i] ** 2 for i in range(6)),
            c=lambda x, u, p: [x[i] * u[i] + u[i + 3] - u[6 + i] + 0.02 for i in range(3)],
            lower=lambda p: [-3.0] * 6 + [0.0] * 3, upper=lambda p: [3.0] * 6 + [INF] * 3, u_init=[0.0] * 6 + [0.01] * 3, dt=dt)

    def f16(x, u, p):
        return [x[i] + dt * (0.5 * x[(i + 3) % 16] - 0.4 * x[i] + (u[i % 4] if i % 4 == 0 else 0.0) + 0.1 * sp.sin(x[(i + 1) % 16]))
                for i in range(16)]
    s16 = M(name="state16", nx=16, nu=6, np_=0, f=f16,
            stage_cost=lambda x, u, p: dt * (sum(0.1 * u[i] * u[i] for i in range(4)) + 2.0 * (u[4] + u[5])),
            term_cost=lambda x, p: 5.0 * sum((x[i] - 0.05 * i) ** 2 for i in range(16)),
            c=lambda x, u, p: [x[0] * u[0] - u[4] + 0.02, x[5] * u[1] - u[5] + 0.02],
            lower=lambda p: [-2.0] * 4 + [0.0] * 2, upper=lambda p: [2.0] * 4 + [INF] * 2, u_init=[0.0] * 4 + [0.01] * 2, dt=dt)

    def f9(x, u, p):
        return [x[i] + dt * (0.4 * x[(i + 1) % 9] - 0.3 * x[i] + (u[i % 8] if i < 16 else 0.0)
                             + 0.05 * sp.sin(x[(i + 2) % 9]) * u[(i + 1) % 8]) for i in range(9)]
    k36 = M(name="k36_state9", nx=9, nu=24, np_=0, f=f9,
            stage_cost=lambda x, u, p: dt * (sum(0.05 * u[i] * u[i] for i in range(8)) + sum(u[8:24])),
            term_cost=lambda x, p: 10.0 * sum((x[i] - 0.05 * i) ** 2 for i in range(9)),
            c=lambda x, u, p: [u[i] - u[8 + i] + u[16 + i] for i in range(8)]
            + [x[i % 9] * u[i] + 0.03 - u[16 + (i + 4) % 8] * 0.5 for i in range(4)],
            lower=lambda p: [-4.0] * 8 + [0.0] * 16, upper=lambda p: [4.0] * 8 + [INF] * 16, u_init=[0.0] * 8 + [0.01] * 16, dt=dt)
    return [single, one, nocon, s16, k36]      # (state12: tests/tools/model_shape_fuzz.py territory; kept as a definition)


PARAMS = {"state12": [1.0, 2.0]}
NAMES = ["wide64", "single_control", "one_by_one", "no_constraints", "state16", "k36_state9"]


@pytest.fixture(scope="module")
def world(tmp_path_factory):
    """All models traced once; one scratch oracle and one emulator library that know them all."""
    import user_model_harness as H
    from ipddp_b200.codegen import generate
    mds = [wide_model("wide64", 10, 4)] + small_models()
    assert [md.name for md in mds] == NAMES
    models = [(md, generate.trace(md)) for md in mds]
    return dict(models={md.name: md for md, _ in models}, emu=H.emulator_with_models(models),
                orc=H.scratch_oracle(tmp_path_factory.mktemp("scratch"), models))


@pytest.mark.parametrize("name", NAMES)
def test_emulated_model_shape(world, name):
    from ipddp_b200.batch import BatchSolver
    md, emu, orc = world["models"][name], world["emu"], world["orc"]
    nx, nu, nc, np_, _slots = emu.model_dims(name)
    assert (nx, nu) == (md.nx, md.nu)
    if name.startswith("wide"):
        assert nu + nc == int(name[4:])
    for N in (9, 2):
        _solve_and_compare(md, emu, orc, name, nx, nu, np_, N)


def _solve_and_compare(md, emu, orc, name, nx, nu, np_, N):
    from ipddp_b200.batch import BatchSolver
    B, maxit = 2, 25
    rng = np.random.default_rng(5)
    x1 = 0.2 * rng.standard_normal((B, nx))
    ubar = np.tile(np.asarray(md.u_init, dtype=np.float64), (B, N - 1))
    P = np.zeros((B, 0)) if np_ == 0 else np.asarray(PARAMS[name]) * (1.0 + 0.05 * rng.standard_normal((B, np_)))
    lower = np.array([md.lower(list(P[i])) for i in range(B)], dtype=np.float64).reshape(B, nu)
    upper = np.array([md.upper(list(P[i])) for i in range(B)], dtype=np.float64).reshape(B, nu)
    oopt = orc.default_options(optimality_tolerance=1e-7, max_iterations=maxit)
    res, xo, uo = orc.solve_batch(name, N, P, lower, upper, x1, ubar, options=oopt, want_traj=True)
    assert N == 2 or max(r.k for r in res) >= 4, [(r.status, r.k) for r in res]          # a real solve, not an immediate exit
    for spec in (-1, 0):
        emu.L.ipddp_set_tuning(None, b"fw_spec_max", spec)
        emu.L.ipddp_set_tuning(None, b"bw_spec_max", spec)
        try:
            s = BatchSolver(name, B, N, options=emu.default_options(optimality_tolerance=1e-7, max_iterations=maxit), lib=emu)
            s.set_inputs(x1, ubar, P if np_ > 0 else None, lower, upper)
            r = s.solve()
            x, u = s.trajectory()
            cnt = s.counters()
            s.close()
        finally:
            emu.L.ipddp_set_tuning(None, b"fw_spec_max", -1)
            emu.L.ipddp_set_tuning(None, b"bw_spec_max", -1)
        for i in range(B):
            o = res[i]
            got = (int(r.status[i]), int(r.k[i]), int(r.j[i]), int(r.l[i]))
            assert got == (o.status, o.k, o.j, o.l), (name, spec, i, got, (o.status, o.k, o.j, o.l))
            for field in ("objective", "primal_inf", "dual_inf", "cs_inf", "mu", "reg_last", "step_size"):
                helpers.assert_same_bits(getattr(r, field)[i], getattr(o, field), f"{name} inst {i} {field}")
            assert (cnt["n_sweeps"][i], cnt["n_kkt"][i], cnt["n_rollouts"][i]) == (o.n_sweeps, o.n_kkt, o.n_rollouts)
        helpers.assert_same_bits(x, xo, f"{name} states")
        helpers.assert_same_bits(u, uo, f"{name} controls")


def test_emulated_chain_with_a_single_control_stage_and_a_state_that_grows_past_seven(tmp_path, monkeypatch):
    """A stage chain (reference README.md:18) built from shapes no built-in model has: a single-control stage, a stage that
    maps 2 states onto 9, and a 9-state stage with the terminal cost -- per-stage arithmetic of both fixes above inside the
    chain dispatch, resident batch (speculative and bulk kernels) and queue mode."""
    import sympy as sp
    import build_emu
    import emu_plugins
    import user_model_harness as H
    from ipddp_b200 import _lib
    from ipddp_b200.codegen import generate, workloads
    dt, M = DT, workloads.ModelDef
    zero = lambda x, p: 0.0 * x[0]
    s0 = M(name="odd_s0", nx=2, nu=1, np_=0, f=lambda x, u, p: [x[0] + dt * x[1], x[1] + dt * (u[0] - sp.sin(x[0]))],
           stage_cost=lambda x, u, p: dt * u[0] * u[0], term_cost=zero, c=lambda x, u, p: [], lower=lambda p: [-2.0],
           upper=lambda p: [INF], u_init=[0.05], dt=dt)
    s1 = M(name="odd_s1", nx=2, nu=3, np_=0,
           f=lambda x, u, p: [x[0] + dt * x[1], x[1] + dt * u[0]] + [dt * (x[i % 2] * 0.5 + 0.1 * i) + 0.0 * u[1] for i in range(7)],
           stage_cost=lambda x, u, p: dt * (u[0] * u[0] + u[1] + u[2]), term_cost=zero, c=lambda x, u, p: [u[1] - u[2] - u[0] * x[1]],
           lower=lambda p: [-3.0, 0.0, 0.0], upper=lambda p: [3.0, INF, INF], u_init=[0.01] * 3, dt=dt)
    s2 = M(name="odd_s2", nx=9, nu=2, np_=0,
           f=lambda x, u, p: [x[i] + dt * (0.3 * x[(i + 1) % 9] - 0.2 * x[i] + (u[i % 2] if i < 4 else 0.0)) for i in range(9)],
           stage_cost=lambda x, u, p: dt * (0.5 * u[0] * u[0] + u[1]) + 0.1 * dt * x[8] * x[8],
           term_cost=lambda x, p: 20.0 * sum((x[i] - 0.1 * i) ** 2 for i in range(9)), c=lambda x, u, p: [],
           lower=lambda p: [-3.0, 0.0], upper=lambda p: [3.0, INF], u_init=[0.01] * 2, dt=dt, nx_term=9)
    chain = workloads.ChainDef("odd", [s0, s1, s2])
    monkeypatch.setitem(workloads.CHAINS, "odd", lambda: chain)
    bundles = [generate.trace(md) for md in chain.stages]
    orc = H.scratch_oracle(tmp_path, list(zip(chain.stages, bundles)))
    emu = _lib.Lib(build_emu.build())
    src = generate.emit_device_chain(chain, [generate.emit_device(md, b) for md, b in zip(chain.stages, bundles)])
    emu.check(emu.L.ipddp_model_load(emu_plugins.compile_plugin("odd", src).encode()), "ipddp_model_load")
    try:
        for spec in (0,):                    # bulk kernels; the queue run below takes the speculative ones
            emu.L.ipddp_set_tuning(None, b"fw_spec_max", spec)
            emu.L.ipddp_set_tuning(None, b"bw_spec_max", spec)
            helpers.chain_parity(emu, orc, "odd", 3, 9, maxit=40)
    finally:
        emu.L.ipddp_set_tuning(None, b"fw_spec_max", -1)
        emu.L.ipddp_set_tuning(None, b"bw_spec_max", -1)
    helpers.chain_parity(emu, orc, "odd", 5, 9, maxit=40, queue_slots=2)
