"""Models wider than the reference's experiments.  The C ABI accepts nu + nc <= 64, the largest built-in workload (cartpole)
has 35: a synthetic model with K = 48 and one with K = 64 (L1-split forces, bilinear state-force couplings with slacks,
bounds on every control) through the whole solve on the emulator -- KKT assembly beyond one lane slot per column, the
two-slot pivot steps of the warp LDL^T on real KKT matrices, the gains / record strides -- against the oracle built with
the same traced closures (tests/emu/user_model_harness.py).  Bit for bit, speculative and bulk kernels."""
import os
import sys

import numpy as np
import pytest

import helpers

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu"))


def wide_model(name, nf, extra=0, dt=0.05):
    """nx = 4, nu = 4 nf + extra, nc = 2 nf (K = 6 nf + extra): forces u[0:nf] = s+ - s- (L1 split: c[0:nf], the first
    `extra` of them with one more non-negative slack each), bilinear couplings x[i % 4] * f_i + 0.05 = t_i with slack t
    (c[nf:2nf])."""
    import sympy as sp
    from ipddp_b200.codegen import workloads
    nu, inf = 4 * nf + extra, float("inf")

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
                              lower=lambda p: [-5.0] * nf + [0.0] * (nu - nf), upper=lambda p: [5.0] * nf + [inf] * (nu - nf),
                              u_init=[0.0] * nf + [0.01] * (nu - nf), dt=dt)


@pytest.mark.parametrize("nf,extra", [pytest.param(8, 0, id="K48"), pytest.param(10, 4, id="K64")])
def test_emulated_wide_model(nf, extra, oracle_mod, tmp_path):
    import user_model_harness as H
    from ipddp_b200.batch import BatchSolver
    from ipddp_b200.codegen import generate
    K = 6 * nf + extra
    md = wide_model(f"wide{K}", nf, extra)
    bundles = generate.trace(md)
    emu = H.emulator_with_model(md, bundles)
    orc = H.scratch_oracle(tmp_path, md, bundles)
    nx, nu, nc, _np, _slots = emu.model_dims(md.name)
    assert (nx, nu, nc) == (4, 4 * nf + extra, 2 * nf) and nu + nc == K
    B, N, maxit = 2, 9, 25
    rng = np.random.default_rng(5)
    x1 = 0.2 * rng.standard_normal((B, 4))
    ubar = np.tile(np.asarray(md.u_init), (B, N - 1))
    lower = np.tile(np.asarray(md.lower([])), (B, 1))
    upper = np.tile(np.asarray(md.upper([])), (B, 1))
    oopt = orc.default_options(optimality_tolerance=1e-7, max_iterations=maxit)
    res, xo, uo = orc.solve_batch(md.name, N, np.zeros((B, 0)), lower, upper, x1, ubar, options=oopt, want_traj=True)
    assert max(r.k for r in res) >= 8, [(r.status, r.k) for r in res]          # a real solve, not an immediate exit
    for spec in (-1, 0):
        emu.L.ipddp_set_tuning(None, b"fw_spec_max", spec)
        emu.L.ipddp_set_tuning(None, b"bw_spec_max", spec)
        try:
            s = BatchSolver(md.name, B, N, options=emu.default_options(optimality_tolerance=1e-7, max_iterations=maxit), lib=emu)
            s.set_inputs(x1, ubar, None, lower, upper)
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
            assert got == (o.status, o.k, o.j, o.l), (spec, i, got, (o.status, o.k, o.j, o.l))
            for name in ("objective", "primal_inf", "dual_inf", "cs_inf", "mu", "reg_last", "step_size"):
                helpers.assert_same_bits(getattr(r, name)[i], getattr(o, name), f"K={K} inst {i} {name}")
            assert (cnt["n_sweeps"][i], cnt["n_kkt"][i], cnt["n_rollouts"][i]) == (o.n_sweeps, o.n_kkt, o.n_rollouts)
        helpers.assert_same_bits(x, xo, "states")
        helpers.assert_same_bits(u, uo, "controls")
