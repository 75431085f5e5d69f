"""Wide randomized GPU-vs-oracle parity sweep (run on a B200; a few minutes): every workload on instances far from the
ones the test-suite uses, resident batches and queues of awkward sizes.  Prints one JSON line per case.
    python tests/tools/parity_sweep.py [instances_per_workload]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402
import oracle  # noqa: E402
import ipddp_b200  # noqa: E402,F401
from ipddp_b200 import _lib, instances  # noqa: E402
from ipddp_b200.batch import BatchSolver  # noqa: E402


def same_bits(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return bool(((a.view(np.int64) == b.view(np.int64)) | ((a == 0) & (b == 0)) | (np.isnan(a) & np.isnan(b))).all())


def case(lib, wl, n, N, first, slots, vary):
    b = instances.make_batch(wl, n, N, vary_horizon=vary, first=first)
    opt = lib.default_options(optimality_tolerance=1e-7)
    s = BatchSolver(wl, slots or n, N, options=opt, lib=lib)
    t0 = time.perf_counter()
    if slots:
        r, cnt, x, u = s.solve_queue(b.x1, b.ubar, b.p if s.np > 0 else None, b.lower, b.upper, b.horizons)
    else:
        s.set_batch(b)
        r = s.solve(); x, u = s.trajectory(); cnt = s.counters()
    tg = time.perf_counter() - t0
    s.close()
    t0 = time.perf_counter()
    res, xo, uo = oracle.solve_batch(wl, N, b.p, b.lower, b.upper, b.x1, b.ubar, options=oracle.default_options(optimality_tolerance=1e-7),
                                     horizons=b.horizons, want_traj=True)
    to = time.perf_counter() - t0
    bad = 0
    for i, o in enumerate(res):
        ok = ((int(r.status[i]), int(r.k[i]), int(r.j[i]), int(r.l[i])) == (o.status, o.k, o.j, o.l)
              and (cnt["n_sweeps"][i], cnt["n_kkt"][i], cnt["n_rollouts"][i]) == (o.n_sweeps, o.n_kkt, o.n_rollouts)
              and all(same_bits(getattr(r, nm)[i], getattr(o, nm)) for nm in ("objective", "primal_inf", "dual_inf", "cs_inf", "mu", "reg_last", "step_size"))
              and same_bits(x[i], xo[i]) and same_bits(u[i], uo[i]))
        bad += 0 if ok else 1
    st = np.bincount(r.status, minlength=10).tolist()
    print(json.dumps(dict(workload=wl, instances=n, knots=N, first=first, slots=slots or n, vary_horizon=vary, mismatches=bad,
                          status_histogram={k: v for k, v in enumerate(st) if v}, gpu_s=round(tg, 2), oracle_s=round(to, 2))), flush=True)
    return bad


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 384
    lib = _lib.load()
    bad = 0
    for wl, N, vary in (("cartpole", 101, False), ("acrobot", 101, False), ("concar", 101, True), ("concar_quad", 101, False),
                        ("pushing", 101, True), ("double_integrator", 101, False)):
        bad += case(lib, wl, n, N, 20000, 0, vary)
    for wl, nq, slots in (("double_integrator", 50, 1), ("concar_quad", 1000, 333), ("concar", 97, 7), ("cartpole", 1, 5), ("acrobot", 640, 640)):
        bad += case(lib, wl, nq, 61, 31000, slots, False)
    print(json.dumps(dict(total_mismatches=bad)))
    sys.exit(1 if bad else 0)
