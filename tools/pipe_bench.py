"""Throughput of ipddp_solve_many for different admission settings (A/B tool).
    python tools/pipe_bench.py <K> <F> <slots,slots,...> [B]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ipddp_b200  # noqa: E402,F401
from ipddp_b200 import _lib, instances  # noqa: E402
from ipddp_b200.batch import BatchSolver, solve_many  # noqa: E402

K, F = int(sys.argv[1]), int(sys.argv[2])
slots = [int(x) for x in sys.argv[3].split(",")]
B = int(sys.argv[4]) if len(sys.argv) > 4 else 16384
lib = _lib.load()
b = instances.make_batch("cartpole", B, 101)
opt = lib.default_options(optimality_tolerance=1e-7)
solvers = [BatchSolver("cartpole", B, 101, options=opt, lib=lib) for _ in range(F)]
for s in solvers:
    s.set_batch(b)
solve_many(solvers[:1], total_solves=1)   # warm-up
for sl in slots:
    lib.L.ipddp_set_tuning(None, b"bulk_slots", sl)
    ms, st = solve_many(solvers, total_solves=K)
    print(json.dumps(dict(K=K, F=F, bulk_slots=sl, ms=round(ms, 1), solves_per_s=round(st.n_converged / (ms * 1e-3), 1),
                          kkt_per_s=round(st.sum_kkt / (ms * 1e-3), 0), launches=st.launches)), flush=True)
for s in solvers:
    s.close()
