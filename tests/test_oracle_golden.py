"""Pins the CPU oracle to the reference's committed known-answer tables
(reference experiments/ipddp2/results/*.txt, instances rebuilt from experiments/ipddp2/params/*.txt).

Acceptance: identical iteration count AND 9-digit objective on >= 90 % cartpole,
>= 95 % concar_quad, >= 50 % acrobot / concar, 1/1 double_integrator; objective within 1e-8 relative on
every iteration-matched row of the well-conditioned classes (acrobot/concar have a few rows that reach
the same count through a different local optimum).  (Iteration-count equality is only a statistical invariant on long runs:
last-bit differences of summation order get amplified by filter/inertia branch decisions, SURVEY App. D.)"""
import numpy as np
import pytest

import ipddp_b200  # noqa: F401
from ipddp_b200 import instances

CASES = [("double_integrator", 1, 1.0), ("cartpole", 100, 0.90), ("concar_quad", 100, 0.95),
         ("acrobot", 100, 0.50), ("concar", 100, 0.50)]


@pytest.mark.parametrize("wl,n,frac", CASES)
def test_golden_table(oracle_mod, wl, n, frac):
    g = instances.load_golden_results(wl)
    b = instances.make_batch(wl, n, 101)
    opt = oracle_mod.default_options(optimality_tolerance=1e-7)
    res, _, _ = oracle_mod.solve_batch(wl, 101, b.p, b.lower, b.upper, b.x1, b.ubar, options=opt)
    both = 0
    for i in range(n):
        r = res[i]
        it_ok = r.k == g["iterations"][i]
        rel = abs(r.objective - g["objective"][i]) / max(1.0, abs(g["objective"][i]))
        if it_ok and rel <= 1e-8:
            assert (r.status == 0) == bool(g["converged"][i])
            both += 1
        elif it_ok and wl in ("double_integrator", "cartpole", "concar_quad"):
            # on the well-conditioned classes an iteration-matched row never lands in another optimum
            raise AssertionError(f"{wl} seed {i+1}: iteration-matched row with objective off by {rel:.2e}")
    assert both >= frac * n, f"{wl}: only {both}/{n} rows reproduce the reference table"


def test_cartpole_seed1_known_answer(oracle_mod):
    """Config 1 of BASELINE.json: cartpole N=101 seed 1 -> 60 iterations, objective 9.29397628e-01
    (reference experiments/ipddp2/results/cartpole_friction.txt:2)."""
    b = instances.make_batch("cartpole", 1, 101)
    s = oracle_mod.OracleSolver("cartpole", 101, b.p[0], b.lower[0], b.upper[0],
                                options=oracle_mod.default_options(optimality_tolerance=1e-7))
    r = s.solve(b.x1[0], b.ubar[0])
    assert r.status == 0 and r.k == 60
    assert abs(r.objective - 9.29397628e-01) < 5e-10
    assert abs(r.primal_inf - 4.56853303e-14) < 1e-13
    tr = s.trace()
    assert tr.shape == (60, oracle_mod.TRACE_COLS)
    assert tr[-1, 0] == 60


def test_pushing_statistical(oracle_mod):
    """pushing_1_obs is chaotic: only distribution-level agreement is asserted."""
    g = instances.load_golden_results("pushing")
    n = 16
    b = instances.make_batch("pushing", n, 101)
    opt = oracle_mod.default_options(optimality_tolerance=1e-7)
    res, _, _ = oracle_mod.solve_batch("pushing", 101, b.p, b.lower, b.upper, b.x1, b.ubar, options=opt)
    ks = np.array([r.k for r in res]); objs = np.array([r.objective for r in res])
    assert 0.4 * np.median(g["iterations"]) < np.median(ks) < 2.0 * np.median(g["iterations"])
    assert 0.5 * np.median(g["objective"]) < np.median(objs) < 2.0 * np.median(g["objective"])


def test_package_params_tables_are_the_reference_fixtures():
    """interiorpointddp.jl_b200/data/params (read by the instance generator at run time) == tests/golden/params
    (byte-identical copies of the reference's experiments/ipddp2/params/*.txt)."""
    import filecmp
    import os
    here = os.path.dirname(os.path.abspath(__file__))
    gold = os.path.join(here, "golden", "params")
    names = sorted(os.listdir(gold))
    assert names == sorted(os.listdir(instances.PARAMS_DIR)) and len(names) == 4
    for n in names:
        assert filecmp.cmp(os.path.join(gold, n), os.path.join(instances.PARAMS_DIR, n), shallow=False), n


def test_table_residual_is_the_one_ulp_noise_floor(oracle_mod):
    """Why only ~54 % of the acrobot rows reproduce the reference's iteration counts: the count is a chaotic function of
    last-bit rounding.  Moving every model parameter to the next representable double makes the oracle disagree WITH
    ITSELF on as many rows as it disagrees with the reference's table (tests/tools/perturbation_study.py, all classes:
    profiles/r2_perturbation_stability.json) -- no implementation that is not bit-identical to the reference's toolchain
    (OpenBLAS kernel choice, Symbolics' expression order, Julia's libm) can match more rows than that."""
    wl, n = "acrobot", 100
    g = instances.load_golden_results(wl)
    b = instances.make_batch(wl, n, 101)
    opt = oracle_mod.default_options(optimality_tolerance=1e-7)
    base, _, _ = oracle_mod.solve_batch(wl, 101, b.p, b.lower, b.upper, b.x1, b.ubar, options=opt)
    pert, _, _ = oracle_mod.solve_batch(wl, 101, np.nextafter(b.p, np.inf), b.lower, b.upper, b.x1, b.ubar, options=opt)
    vs_table = sum(1 for i in range(n) if base[i].k == g["iterations"][i]
                   and abs(base[i].objective - g["objective"][i]) <= 1e-8 * max(1.0, abs(g["objective"][i])))
    vs_self = sum(1 for i in range(n) if base[i].k == pert[i].k and f"{base[i].objective:.8e}" == f"{pert[i].objective:.8e}")
    assert vs_table >= 50 and vs_self >= 40
    assert vs_self <= vs_table + 15, (vs_self, vs_table)      # the table is reproduced about as well as the oracle reproduces itself
    # the perturbation is harmless where it should be: every instance still converges to (nearly) the same optimum or another
    # local one of similar cost
    assert sum(1 for r in pert if r.status == 0) >= 95
