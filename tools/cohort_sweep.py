"""Single-solve time of one 16384-instance cartpole batch as a function of the number of cohorts, and the
pipelined throughput of 3 batches in flight with cohorts."""
import sys, os, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import ipddp_b200
from ipddp_b200 import _lib, instances
from ipddp_b200.batch import BatchSolver, solve_many
lib = _lib.load()
wl = sys.argv[1] if len(sys.argv) > 1 else "cartpole"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
Ss = [int(x) for x in (sys.argv[3] if len(sys.argv) > 3 else "1,4,8,16,32").split(",")]
b = instances.make_batch(wl, B, 101)
opt = lib.default_options(optimality_tolerance=1e-7)
s = BatchSolver(wl, B, 101, options=opt, lib=lib)
s.set_batch(b)
ref = None
for S in Ss:
    s.set_cohorts(S)
    t0 = time.time(); r = s.solve(); dt = time.time() - t0
    st = s.stats()
    if ref is None: ref = r
    same = bool(np.array_equal(ref.k, r.k) and np.array_equal(ref.objective.view(np.int64), r.objective.view(np.int64)))
    print(json.dumps(dict(cohorts=S, wall_s=round(dt, 2), ms_total=round(st.ms_total, 1), solves_per_s=round(st.n_converged / dt, 1), same_as_S1=same, launches=st.launches)), flush=True)
F = 3
Sbest = int(sys.argv[4]) if len(sys.argv) > 4 else 8
solvers = [s] + [BatchSolver(wl, B, 101, options=opt, lib=lib) for _ in range(F - 1)]
for q in solvers:
    q.set_batch(b); q.set_cohorts(Sbest)
ms, st = solve_many(solvers, total_solves=F)
print(json.dumps(dict(pipelined_handles=F, cohorts=Sbest, ms=round(ms, 1), solves_per_s=round(st.n_converged / (ms * 1e-3), 1))), flush=True)
