"""Per-kernel device-time split of one batched solve (CUDA events inside libipddp_b200.so)."""
import sys, os, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import ipddp_b200
from ipddp_b200 import _lib, instances
from ipddp_b200.batch import BatchSolver

def run(wl, B, N=int(os.environ.get('IPDDP_KNOTS', '101')), reps=1):
    lib = _lib.load()
    b = instances.make_batch(wl, B, N, vary_horizon=os.environ.get('IPDDP_VARY_HORIZON', '') != '')
    s = BatchSolver(wl, B, N, options=lib.default_options(optimality_tolerance=1e-7), lib=lib)
    s.set_batch(b)
    out = None
    for _ in range(reps):
        t0 = time.time(); r = s.solve(); wall = time.time() - t0
        st = s.stats()
        out = dict(workload=wl, B=B, wall_s=round(wall, 3), ms_total=round(st.ms_total, 1), rounds=st.iterations, launches=st.launches,
                   ms_init=round(st.ms_init, 2), ms_derivs=round(st.ms_derivs, 1), ms_backward=round(st.ms_backward, 1),
                   ms_check=round(st.ms_check, 1), ms_forward=round(st.ms_forward, 1), kkt=st.sum_kkt, rollouts=st.sum_rollouts,
                   conv=int(st.n_converged), solves_per_s=round(st.n_converged / (st.ms_total * 1e-3), 1),
                   kkt_per_s=round(st.sum_kkt / (st.ms_backward * 1e-3), 0), active_frac=round(st.n_active_rounds / max(1, st.iterations) / B, 3))
    s.close()
    print(json.dumps(out), flush=True)

if __name__ == "__main__":
    wl = sys.argv[1]
    for B in [int(x) for x in sys.argv[2].split(",")]:
        run(wl, B, reps=int(sys.argv[3]) if len(sys.argv) > 3 else 1)
