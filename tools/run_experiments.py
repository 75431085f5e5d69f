"""Regenerates the reference's results tables (experiments/ipddp2/results/*.txt, written by e.g.
experiments/ipddp2/cartpole_friction.jl:151-161) on the GPU: the 100 seeded instances of each class are solved as ONE
batch per class, and the table is written in the reference's column format.  With --compare the tables are diffed
against the committed golden copies (tests/golden/results): iteration-count and 9-digit-objective agreement per class.

    python tools/run_experiments.py [--out gpurun_out/results] [--compare]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import ipddp_b200  # noqa: F401,E402
from ipddp_b200 import _lib, instances, results_io  # noqa: E402
from ipddp_b200.batch import BatchSolver  # noqa: E402

CLASSES = [("cartpole", "cartpole_friction"), ("acrobot", "acrobot_contact"), ("concar", "concar"),
           ("concar_quad", "concar_quad"), ("pushing", "pushing_1_obs"), ("double_integrator", "double_integrator")]


def run_class(lib, wl, fname, out, rows=None):
    """Solves the first `rows` (default: all) seeded instances of one class as ONE batch and writes the class' results table
    (and params table) in the reference's formats; returns the agreement with the committed golden table."""
    g = instances.load_golden_results(wl)
    n = len(g["seed"]) if rows is None else min(rows, len(g["seed"]))
    g = {k: v[:n] for k, v in g.items()}
    b = instances.make_batch(wl, n, 101)
    s = BatchSolver(wl, n, 101, options=lib.default_options(optimality_tolerance=1e-7), lib=lib)
    s.set_batch(b)
    r = s.solve()
    st = s.stats()
    s.close()
    wall_ms = st.ms_total / n      # batch device time amortised per instance
    solver_ms = (st.ms_total - st.ms_derivs) / n
    # the reference's writer format (experiments/ipddp2/cartpole_friction.jl:151-168), checked byte for byte against the
    # reference's own tables by tests/test_results_io.py; the file parses with experiments/utils.jl's read_results
    results_io.write_results(os.path.join(out, fname + ".txt"), np.arange(1, n + 1), r.k, r.status == 0, r.objective,
                             r.primal_inf, np.full(n, wall_ms), np.full(n, solver_ms))
    back = results_io.read_results(os.path.join(out, fname + ".txt"))
    assert back.iters == [int(x) for x in r.k] and back.status == [bool(x) for x in (r.status == 0)]
    if b.p.shape[1] > 0:
        os.makedirs(os.path.join(out, "params"), exist_ok=True)
        results_io.write_params(os.path.join(out, "params", fname + ".txt"), b.p.tolist())
    close = np.abs(r.objective - g["objective"]) <= 1e-8 * np.maximum(1.0, np.abs(g["objective"]))
    return dict(instances=n, converged_gpu=int((r.status == 0).sum()), converged_reference=int(g["converged"].sum()),
                same_iterations=int((r.k == g["iterations"]).sum()), same_objective_1e8=int(close.sum()),
                same_both=int(((r.k == g["iterations"]) & close).sum()),
                mean_iterations_gpu=float(r.k.mean()), mean_iterations_reference=float(g["iterations"].mean()),
                batch_ms=round(st.ms_total, 1), ms_per_instance=round(wall_ms, 3))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "results"))
    ap.add_argument("--compare", action="store_true")
    args = ap.parse_args()
    os.makedirs(args.out, exist_ok=True)
    lib = _lib.load()
    summary = {}
    for wl, fname in CLASSES:
        summary[fname] = run_class(lib, wl, fname, args.out)
        print(fname, json.dumps(summary[fname]), flush=True)
    json.dump(summary, open(os.path.join(args.out, "summary.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
