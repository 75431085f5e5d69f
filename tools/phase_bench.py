"""Per-phase timing of single rounds through the phase-level C ABI (ipddp_eval_derivatives / backward_pass / check /
forward_pass), for A/B comparisons of kernel variants.

    python tools/phase_bench.py <workload> <B,B,...> [rounds] [lib.so]

For each B: initialise, run `rounds` lock-step rounds, print the wall time of every phase call of the LAST rounds
(the phase calls synchronise their stream, so wall time = kernel time + ~20 us launch/sync overhead).  B = 8 gives
the lone-warp latency (8 warps on 148 SMs), B = 16384 the bulk throughput."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import ipddp_b200  # noqa: E402,F401
from ipddp_b200 import _lib, instances  # noqa: E402
from ipddp_b200.batch import BatchSolver  # noqa: E402


def run(lib, wl, B, N, rounds):
    b = instances.make_batch(wl, B, N)
    s = BatchSolver(wl, B, N, options=lib.default_options(optimality_tolerance=1e-7), lib=lib)
    s.set_batch(b)
    s.initialize()
    rows = []
    series = os.environ.get("IPDDP_SERIES", "") != ""
    prev = None
    for r in range(rounds):
        t0 = time.perf_counter(); s.eval_derivatives()
        t1 = time.perf_counter(); s.backward_pass()
        t2 = time.perf_counter(); nf = s.check()
        t3 = time.perf_counter(); s.forward_pass()
        t4 = time.perf_counter()
        rows.append(((t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, (t4 - t3) * 1e3, nf))
        if series:
            c = s.counters()
            if prev is not None or True:
                p0 = prev or {k: np.zeros_like(v) for k, v in c.items()}
                dk = int((c["n_kkt"] - p0["n_kkt"]).sum()); ds = int((c["n_sweeps"] - p0["n_sweeps"]).sum())
                db = int((c["n_backward"] - p0["n_backward"]).sum()); dr = int((c["n_rollouts"] - p0["n_rollouts"]).sum())
                mx = int((c["n_sweeps"] - p0["n_sweeps"]).max())
                print(json.dumps(dict(round=r, active=db, sweeps=ds, max_sweeps=mx, kkt=dk, bw_ms=round(rows[-1][1], 2),
                                      Mkkt_per_s=round(dk / rows[-1][1] / 1e3, 2), fw_ms=round(rows[-1][3], 2), rollouts=dr)), flush=True)
            prev = {k: v.copy() for k, v in c.items()}
            if db == 0:
                break
    cnt = s.counters()
    import hashlib
    x, u = s.trajectory()
    res = s.results()
    digest = hashlib.sha256(np.ascontiguousarray(x).tobytes() + np.ascontiguousarray(u).tobytes()
                            + np.ascontiguousarray(res.objective).tobytes() + np.ascontiguousarray(res.k).tobytes()
                            + np.ascontiguousarray(cnt["n_kkt"]).tobytes()).hexdigest()[:16]
    s.close()
    a = np.array([r[:4] for r in rows])
    tail = a[max(1, rounds // 2):]
    out = dict(workload=wl, B=B, rounds=rounds,
               first_ms=dict(derivs=round(a[0, 0], 3), backward=round(a[0, 1], 3), check=round(a[0, 2], 3), forward=round(a[0, 3], 3)),
               mean_ms_late=dict(derivs=round(tail[:, 0].mean(), 3), backward=round(tail[:, 1].mean(), 3),
                                 check=round(tail[:, 2].mean(), 3), forward=round(tail[:, 3].mean(), 3)),
               sum_ms=dict(derivs=round(a[:, 0].sum(), 2), backward=round(a[:, 1].sum(), 2), check=round(a[:, 2].sum(), 2),
                           forward=round(a[:, 3].sum(), 2)),
               digest=digest, kkt=int(cnt["n_kkt"].sum()), rollouts=int(cnt["n_rollouts"].sum()),
               kkt_per_s=round(cnt["n_kkt"].sum() / (a[:, 1].sum() * 1e-3), 0),
               us_per_kkt_per_warp=round(a[:, 1].sum() * 1e3 / max(1, cnt["n_kkt"].max()), 2))
    print(json.dumps(out), flush=True)


def apply_tuning(lib):
    """IPDDP_TUNE="key=value,key=value": global ipddp_set_tuning defaults for the problems created afterwards."""
    for kv in filter(None, os.environ.get("IPDDP_TUNE", "").split(",")):
        k, v = kv.split("=")
        lib.check(lib.L.ipddp_set_tuning(None, k.encode(), int(v)), "ipddp_set_tuning")


if __name__ == "__main__":
    wl = sys.argv[1]
    Bs = [int(x) for x in sys.argv[2].split(",")]
    rounds = int(sys.argv[3]) if len(sys.argv) > 3 else 12
    lib = _lib.Lib(sys.argv[4]) if len(sys.argv) > 4 else _lib.load()
    N = int(os.environ.get("IPDDP_KNOTS", "101"))
    apply_tuning(lib)
    for B in Bs:
        run(lib, wl, B, N, rounds)
