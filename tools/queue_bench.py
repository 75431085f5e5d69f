"""Streaming throughput of ipddp_solve_queue: Q queued instances of a workload through B resident slots (device-resident
inputs, device-timed).    python tools/queue_bench.py <workload> <Q> <B,B,...> [knots] [lib.so]"""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import ipddp_b200  # noqa: E402,F401
from ipddp_b200 import _lib, instances  # noqa: E402
from ipddp_b200.batch import BatchSolver, make_queue  # noqa: E402


def run(lib, wl, Q, B, N, vary):
    b = instances.make_batch(wl, Q, N, vary_horizon=vary)
    dev = torch.device("cuda", 0)
    s = BatchSolver(wl, B, N, options=lib.default_options(optimality_tolerance=1e-7), lib=lib)
    t = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in
         dict(x1=b.x1, ubar=b.ubar, p=b.p if s.np > 0 else np.zeros((Q, 1)), lower=b.lower, upper=b.upper).items()}
    hz = torch.from_numpy(b.horizons.astype(np.int32)).to(dev)
    st_i = torch.zeros(Q, dtype=torch.int32, device=dev)
    k_i = torch.zeros(Q, dtype=torch.int32, device=dev)
    sc = [st_i.data_ptr(), k_i.data_ptr()] + [None] * 13
    q = make_queue(Q, t["x1"].data_ptr(), t["ubar"].data_ptr(), t["p"].data_ptr() if s.np > 0 else None, t["lower"].data_ptr(),
                   t["upper"].data_ptr(), hz.data_ptr(), sc, None, None, inputs_on_device=True, outputs_on_device=True)
    for rep in range(int(os.environ.get("IPDDP_REPS", "1"))):
        lib.check(lib.L.ipddp_solve_queue(s.h, C.byref(q)), "ipddp_solve_queue")
        st = s.stats()
        print(json.dumps(dict(workload=wl, Q=Q, B=B, N=N, rep=rep, ms=round(st.ms_total, 1), converged=int(st.n_converged),
                              solves_per_s=round(st.n_converged / (st.ms_total * 1e-3), 1), rounds=st.iterations,
                              mean_active=round(st.n_active_rounds / max(1, st.iterations), 1),
                              kkt_per_s_bw=round(st.sum_kkt / max(1e-9, st.ms_backward * 1e-3), 0),
                              ms_derivs=round(st.ms_derivs, 1), ms_backward=round(st.ms_backward, 1), ms_check=round(st.ms_check, 1),
                              ms_forward=round(st.ms_forward, 1), mean_k=float(k_i.double().mean()))), flush=True)
    s.close()


def apply_tuning(lib):
    """IPDDP_TUNE="key=value,key=value": global ipddp_set_tuning defaults for the problems created afterwards."""
    for kv in filter(None, os.environ.get("IPDDP_TUNE", "").split(",")):
        k, v = kv.split("=")
        lib.check(lib.L.ipddp_set_tuning(None, k.encode(), int(v)), "ipddp_set_tuning")


if __name__ == "__main__":
    wl = sys.argv[1]
    Q = int(sys.argv[2])
    Bs = [int(x) for x in sys.argv[3].split(",")]
    N = int(sys.argv[4]) if len(sys.argv) > 4 else 101
    lib = _lib.Lib(sys.argv[5]) if len(sys.argv) > 5 else _lib.load()
    apply_tuning(lib)
    for B in Bs:
        run(lib, wl, Q, B, N, vary=(wl == "pushing"))
