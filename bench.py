#!/usr/bin/env python
"""bench.py -- headline benchmark of the batched IPDDP2 hot path (BASELINE.json).

metric   : converged OCP solves/sec (batched)            unit: solves/s
workload : BASELINE configs[1] = cartpole swing-up, batch of 16384 random initial states, N = 101 knots,
           optimality_tolerance 1e-7, on ONE B200 (per GPU; weak scaling over GPUs: every rank solves its
           own 16384 instances, no data-path collective, one NCCL all-reduce of statistics at the end).
step     : one batch of 16384 synthetic instances solved to termination (initialise + derivative / backward / check /
           forward rounds).  The K steps of a timed region are queued behind each other on ONE problem handle whose
           resident slots are refilled instance by instance (ipddp_solve_queue), so the rounds stay full.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B] [--workload W]

`value`  : device-timed (CUDA events on the library's stream), inputs and outputs resident in HBM.
`e2e`    : the same metric through the C ABI with HOST buffers: pinned H2D of the inputs + solve + D2H of
           the SolverData scalars, work counters and the trajectories, inside the timed region.
`configs`: short runs of BASELINE configs 3-5 (acrobot 8192 x 201 knots sharded over the GPUs, concar / concar_quad
           4096, pushing 2048 with horizons 61..141), each with its KKT rate, roofline fractions and the oracle's rate.
--impl reference : the CPU restatement of the reference (oracle/, -O3, OpenMP over instances on all host cores)
           on a bounded sample of the same workload.  (The reference itself is Julia; no Julia here.)
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "oracle")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402

METRIC = "converged OCP solves/sec (batched)"
UNIT = "solves/s"


def kkt_flops(n, m, p):
    """Algorithmic FLOPs per backward-pass timestep-KKT; n' = n."""
    K = m + p
    n2 = n
    return ((2 * p * m + 2 * n2 * m + 3 * m) + (2 * n * n2 * n2 + 2 * n * n * n2)
            + (2 * m * n2 * n2 + 2 * m * m * n2 + m * m + 3 * m) + (2 * m * n * n2) + 2 * (n * n + m * n + m * m)
            + K ** 3 / 3.0 + 2 * K * K * (n + 1) + (2 * m * n + 6 * m) + (2 * n * n * m + 2 * n * n * p + n * n)
            + (4 * p * n + 2 * m * n + 4 * n * n2))


def kkt_bytes_dense(n, m, p):
    """Algorithmic bytes per timestep-KKT with dense derivative tiles through HBM."""
    K = m + p
    tile = n * n + n * m + n + m + n * n + m * m + m * n + p * n + p * m + n * n + m * n + m * m
    return 8 * (tile + (n + 5 * m + 2 * p) + (K + 2 * m) * (n + 1) + m + n)


def kkt_bytes_compact(n, m, p, slots):
    """Bytes this implementation's layout moves per timestep-KKT (compact tile of `slots` doubles)."""
    K = m + p
    return 8 * (slots + (n + 5 * m + 2 * p) + (K + 2 * m) * (n + 1) + m + n)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.index = index
        self.rows = []
        self._stop = threading.Event()
        self._t = None

    def _run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for nm, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(max(mx)) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}



# BASELINE.json configs 3-5 (config 2 is the headline workload): short runs reported under "configs".
# batch = instances of the whole job (sharded over the ranks: strong scaling), knots = max horizon.
OTHER_CONFIGS = [
    dict(name="config3", workload="acrobot", batch=8192, knots=201, vary=False, span=0,
         what="acrobot batch 8192, N=200 stages (201 knots), sharded over the GPUs"),
    dict(name="config4a", workload="concar", batch=4096, knots=101, vary=False, span=0,
         what="concar (state+control constraints and bounds) batch 4096"),
    dict(name="config4b", workload="concar_quad", batch=4096, knots=101, vary=False, span=0, what="concar_quad batch 4096"),
    dict(name="config5", workload="pushing", batch=2048, knots=141, vary=True, span=80,
         what="planar pushing (contact complementarity) batch 2048, horizons 61..141 knots"),
]


def reference_sample(args, cores, per_core):
    """Bounded sample of the headline workload for the CPU arm: instances 1000.. of the canonical stream (past the 100
    reference rows), not the first ones."""
    n = args.cpu_sample if args.cpu_sample > 0 else cores * per_core
    return max(cores, min(args.batch, n)), 1000


def run_reference(args, rank, world):
    """CPU arm: the oracle (literal restatement of the reference, -O3, OpenMP over instances) on all host cores."""
    if rank != 0:
        return
    import oracle
    import ipddp_b200  # noqa: F401
    from ipddp_b200 import instances
    cores = os.cpu_count() or 1
    sample, first = reference_sample(args, cores, 48)
    b = instances.make_batch(args.workload, sample, args.knots, first=first)
    opt = oracle.default_options(optimality_tolerance=args.tol)
    conv = 0
    tot = 0.0
    kkt = 0
    for step in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        res, _, _ = oracle.solve_batch(args.workload, args.knots, b.p, b.lower, b.upper, b.x1, b.ubar, options=opt,
                                       horizons=b.horizons, nthreads=cores)
        dt = time.perf_counter() - t0
        if step >= args.warmup:
            tot += dt
            conv += sum(1 for r in res if r.status == 0)
            kkt += sum(r.n_kkt for r in res)
    val = conv / tot
    sample_txt = (f"instances {first}..{first + sample - 1} of the workload's stream per step ({sample} of the {args.batch} "
                  f"a GPU step solves; per-instance rate), {cores} OpenMP threads")
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * tot / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{args.workload} batch, N={args.knots} knots, tol {args.tol:g} (bounded sample of {sample})",
                   "batch_per_gpu": args.batch, "knots": args.knots},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample_txt,
                         "kkt_steps_per_s": kkt / tot},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=16384, help="instances per GPU and step")
    ap.add_argument("--slots", type=int, default=0,
                    help="resident instance slots of the handle (0 = 8 x --batch, capped by the queue length and halved until "
                         "the handle fits the free device memory: fewer, fuller rounds)")
    ap.add_argument("--workload", default="cartpole")
    ap.add_argument("--knots", type=int, default=101)
    ap.add_argument("--tol", type=float, default=1e-7)
    ap.add_argument("--cpu-sample", type=int, default=0,
                    help="instances of the CPU baseline sample (0 = 128 x cores for cpu_baseline, 48 x cores per step for --impl reference)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the short runs of BASELINE configs 3-5")
    ap.add_argument("--no-single", action="store_true", help="skip the lone-batch (one step, nothing queued behind it) run")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import ctypes as C
    import torch
    import torch.distributed as dist
    import ipddp_b200  # noqa: F401
    from ipddp_b200 import _lib, instances
    from ipddp_b200.batch import BatchSolver, make_queue

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    nx, nu, nc, npar, slots = lib.model_dims(args.workload)
    B, N, K = args.batch, args.knots, args.steps
    S = args.slots if args.slots > 0 else min(8 * B, max(B, K * B))
    opt = lib.default_options(optimality_tolerance=args.tol)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    hbm_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # every rank takes its own contiguous range of the workload's canonical instance stream; a step is one batch of B
    # instances, the K steps of a timed region are queued behind each other on ONE handle with S resident slots
    batch = instances.make_batch(args.workload, B, N, first=rank * B)
    names = ["x1", "ubar", "p", "lower", "upper"]
    host1 = dict(x1=batch.x1, ubar=batch.ubar, p=batch.p if npar > 0 else np.zeros((B, 1)), lower=batch.lower, upper=batch.upper)
    # measured (131 072 vs 65 536 slots on a 327 680-instance cartpole queue): 5 149 vs 5 031 solves/s for 83 vs 42 GB
    solver = None
    io_bytes = K * B * 8 * (nx + (N - 1) * nu + max(npar, 1) + 2 * nu + N * nx + (N - 1) * nu + 16)   # queue inputs + outputs
    while solver is None:
        try:
            solver = BatchSolver(args.workload, S, N, options=opt, device=local_rank, lib=lib)
            if args.slots <= 0 and S > B and torch.cuda.mem_get_info(dev)[0] < 1.5 * io_bytes:
                solver.close()
                solver = None
                raise RuntimeError("no room left for the queue's inputs and outputs")
        except RuntimeError:
            if args.slots > 0 or S <= B:
                raise
            S = max(B, S // 2)

    class DevQueue:
        """`steps` copies of the batch as one queue, inputs and outputs resident in HBM."""

        def __init__(self, steps):
            self.Q = Q = steps * B
            self.t = {k: torch.from_numpy(np.ascontiguousarray(host1[k])).to(dev).repeat(steps, 1) for k in names}
            self.hz = torch.from_numpy(batch.horizons.astype(np.int32)).to(dev).repeat(steps)
            self.oi = [torch.zeros(Q, dtype=torch.int32, device=dev) for _ in range(4)]
            self.od = [torch.zeros(Q, dtype=torch.float64, device=dev) for _ in range(7)]
            self.oc = [torch.zeros(Q, dtype=torch.int32, device=dev) for _ in range(4)]
            self.x = torch.zeros((Q, N, nx), dtype=torch.float64, device=dev)
            self.u = torch.zeros((Q, N - 1, nu), dtype=torch.float64, device=dev)
            t = self.t
            self.q = make_queue(Q, t["x1"].data_ptr(), t["ubar"].data_ptr(), t["p"].data_ptr() if npar > 0 else None,
                                t["lower"].data_ptr(), t["upper"].data_ptr(), self.hz.data_ptr(),
                                [a.data_ptr() for a in self.oi + self.od + self.oc], self.x.data_ptr(), self.u.data_ptr(),
                                inputs_on_device=True, outputs_on_device=True)

        def solve(self):
            lib.check(lib.L.ipddp_solve_queue(solver.h, C.byref(self.q)), "ipddp_solve_queue")
            return solver.stats()

    # ---------------------------------------------------- warm-up: W untimed steps through the same path
    if args.warmup > 0:
        DevQueue(args.warmup).solve()
        torch.cuda.empty_cache()

    # ---------------------------------------------------- timed region A (`value`): K steps, inputs resident in HBM.
    # Device time by CUDA events on the library's stream from the first enqueue to the last completion
    # (ipddp_stats.ms_total), per-kernel CUDA events of the same region for the rooflines.
    dq = DevQueue(K)
    barrier()
    with ClockSampler(local_rank) as clk:
        st = dq.solve()
        barrier()
    clocks = clk.summary()
    _free, _total = torch.cuda.mem_get_info(dev)
    mem_used_gb = round((_total - _free) / 1e9, 1)    # slots + this region's queue (inputs and outputs) on the device
    agg = dict(ms_total=st.ms_total, ms_derivs=st.ms_derivs, ms_backward=st.ms_backward, ms_check=st.ms_check,
               ms_forward=st.ms_forward, kkt=st.sum_kkt, sweeps=st.sum_sweeps, rollouts=st.sum_rollouts,
               backward_calls=st.sum_backward, conv=st.n_converged, launches=st.launches, rounds=st.iterations,
               active_rounds=st.n_active_rounds, active_sq=st.sum_active_sq)
    status_d, k_d, prim_d = dq.oi[0], dq.oi[1], dq.od[1]
    conv_mask = status_d == 0
    ksum = float(k_d.double().sum().item())
    pr_max = float(prim_d[conv_mask].max().item()) if bool(conv_mask.any()) else 0.0
    del dq
    torch.cuda.empty_cache()

    # ---------------------------------------------------- lone batch: one step with nothing queued behind it (the
    # lock-step tail of its slowest instances is exposed)
    single = None
    if not args.no_single:
        d1 = DevQueue(1)
        barrier()
        s1 = d1.solve()
        barrier()
        single = dict(ms=s1.ms_total, conv=s1.n_converged, rounds=s1.iterations, active_rounds=s1.n_active_rounds)
        del d1
        torch.cuda.empty_cache()

    # ---------------------------------------------------- timed region B (`e2e`): the same K steps through HOST buffers:
    # pinned inputs -> H2D, solve, D2H of the SolverData scalars, counters and trajectories, all inside the timed region
    Q = K * B

    def pin(a, reps):
        return torch.from_numpy(np.ascontiguousarray(a)).repeat(reps, *([1] * (a.ndim - 1))).pin_memory()
    hin = {k: pin(host1[k], K) for k in names}
    hhz = pin(batch.horizons.astype(np.int32), K)
    hoi = [torch.zeros(Q, dtype=torch.int32).pin_memory() for _ in range(4)]
    hod = [torch.zeros(Q, dtype=torch.float64).pin_memory() for _ in range(7)]
    hoc = [torch.zeros(Q, dtype=torch.int32).pin_memory() for _ in range(4)]
    hx = torch.zeros((Q, N, nx), dtype=torch.float64).pin_memory()
    hu = torch.zeros((Q, N - 1, nu), dtype=torch.float64).pin_memory()
    hq = make_queue(Q, hin["x1"].data_ptr(), hin["ubar"].data_ptr(), hin["p"].data_ptr() if npar > 0 else None,
                    hin["lower"].data_ptr(), hin["upper"].data_ptr(), hhz.data_ptr(),
                    [a.data_ptr() for a in hoi + hod + hoc], hx.data_ptr(), hu.data_ptr())
    barrier()
    t0 = time.perf_counter()
    lib.check(lib.L.ipddp_solve_queue(solver.h, C.byref(hq)), "ipddp_solve_queue")
    conv_e2e = int((hoi[0].numpy() == 0).sum())     # the device->host read of the step results
    barrier()
    wall_e2e = time.perf_counter() - t0
    st_e2e = solver.stats()
    h2d = sum(int(t.numel() * t.element_size()) for k_, t in hin.items() if not (npar == 0 and k_ == "p")) // K + B * 4
    d2h = B * (8 * 4 + 7 * 8) + (int(hx.numel()) + int(hu.numel())) * 8 // K

    # ------------------------------------------------------------------ reduce over ranks
    vals = torch.tensor([agg["ms_total"], wall_e2e, single["ms"] if single else 0.0], dtype=torch.float64, device=dev)
    sums = torch.tensor([agg["conv"], conv_e2e, agg["kkt"], agg["launches"] + st_e2e.launches, ksum,
                         single["conv"] if single else 0.0, pr_max], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(vals, op=dist.ReduceOp.MAX)      # time = max over ranks
        mx = sums[-1:].clone()
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)      # the single NCCL reduction of convergence statistics
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sums[-1] = mx[0]
    ms_total, wall_e2e_max, ms_single = [float(x) for x in vals.tolist()]
    conv_all, conv_e2e_all, kkt_all, launches_all, ksum_all, conv_single, pr_max = [float(x) for x in sums.tolist()]

    # ------------------------------------------------------------------ BASELINE configs 3-5, short runs (every rank its shard)
    cfg_lines = []
    if not args.no_configs:
        fp64_peak_cfg = float(lib.L.ipddp_measure_fp64_tflops(local_rank))
        for cfg in OTHER_CONFIGS:
            wl, Bc, Nc = cfg["workload"], cfg["batch"], cfg["knots"]
            lo, hi = (Bc * rank) // world, (Bc * (rank + 1)) // world
            cb = instances.make_batch(wl, hi - lo, Nc, vary_horizon=cfg["vary"], first=lo, horizon_span=cfg["span"] or 40)
            cs = BatchSolver(wl, hi - lo, Nc, options=opt, device=local_rank, lib=lib)
            cnx, cnu, cnc, cnp, cslots = lib.model_dims(wl)
            barrier()
            for rep in range(2):   # first pass = warm-up
                r_, cnt_, _, _ = cs.solve_queue(cb.x1, cb.ubar, cb.p if cnp > 0 else None, cb.lower, cb.upper, cb.horizons,
                                                want_traj=False)
            cst = cs.stats()
            cs.close()
            barrier()
            v = torch.tensor([cst.ms_total, cst.ms_backward], dtype=torch.float64, device=dev)
            sm_ = torch.tensor([float(cst.n_converged), float(cst.sum_kkt), float(cst.iterations), float(cst.n_active_rounds)],
                               dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(v, op=dist.ReduceOp.MAX)
                dist.all_reduce(sm_, op=dist.ReduceOp.SUM)
            ms_c, ms_bw = [float(x) for x in v.tolist()]
            conv_c, kkt_c, rounds_c, act_c = [float(x) for x in sm_.tolist()]
            line_c = {"name": cfg["name"], "workload": cfg["what"], "batch": Bc, "knots": Nc, "n_gpus": world,
                      "scaling": "strong" if world > 1 else None, "instances_per_gpu": hi - lo,
                      "solves_per_s": conv_c / (ms_c * 1e-3), "ms": ms_c, "converged_fraction": conv_c / Bc,
                      "rounds": rounds_c / world, "mean_active_fraction": act_c / max(1.0, rounds_c) / (hi - lo),
                      "backward_kkt_steps_per_s": kkt_c / (ms_bw * 1e-3) if ms_bw > 0 else None,
                      "roofline_frac_hbm": (kkt_bytes_dense(cnx, cnu, cnc) * kkt_c / world / (ms_bw * 1e-3) / 1e9 / hbm_peak) if ms_bw > 0 else None,
                      "roofline_frac_fp64": (kkt_flops(cnx, cnu, cnc) * kkt_c / world / (ms_bw * 1e-3) / 1e12 / fp64_peak_cfg)
                      if ms_bw > 0 and fp64_peak_cfg > 0 else None}
            if rank == 0 and world == 1 and not args.no_cpu_baseline:
                import oracle
                cores = os.cpu_count() or 1
                ns = min(Bc, 4 * cores)
                ob = instances.make_batch(wl, ns, Nc, vary_horizon=cfg["vary"], first=1000, horizon_span=cfg["span"] or 40)
                t1 = time.perf_counter()
                ores, _, _ = oracle.solve_batch(wl, Nc, ob.p, ob.lower, ob.upper, ob.x1, ob.ubar,
                                                options=oracle.default_options(optimality_tolerance=args.tol),
                                                horizons=ob.horizons, nthreads=cores)
                dt = time.perf_counter() - t1
                line_c["cpu"] = {"solves_per_s": sum(1 for o in ores if o.status == 0) / dt, "cores": cores, "kind": "port",
                                 "sample": f"{ns} instances (1000.. of the stream), one pass, {dt:.1f} s"}
            cfg_lines.append(line_c)

    if rank == 0:
        value = conv_all / (ms_total * 1e-3)
        e2e_val = conv_e2e_all / wall_e2e_max
        # rooflines of the dominant kernel (backward sweep) from the per-kernel CUDA events of timed region A, rank 0
        fp64_peak = float(lib.L.ipddp_measure_fp64_tflops(local_rank))
        flops_kkt = kkt_flops(nx, nu, nc)
        tb = agg["ms_backward"] * 1e-3
        kkt_rank = agg["kkt"]
        gb_dense = kkt_bytes_dense(nx, nu, nc) * kkt_rank / tb / 1e9
        gb_compact = kkt_bytes_compact(nx, nu, nc, slots) * kkt_rank / tb / 1e9
        tf = flops_kkt * kkt_rank / tb / 1e12
        # dram__bytes_read+write of k_backward from the committed ncu --set full capture (profiles/ncu_traffic.json):
        # bytes per timestep-KKT x the timestep-KKTs of an average launch of this run
        ncu_bytes_per_kkt, ncu_src = None, None
        try:
            nt = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))[args.workload]
            ncu_bytes_per_kkt, ncu_src = float(nt["dram_bytes_per_kkt_step"]), nt.get("source")
        except Exception:
            pass
        launches_bw = max(1, agg["rounds"])
        traffic = ncu_bytes_per_kkt * kkt_rank / launches_bw if ncu_bytes_per_kkt else None
        roof_hbm = {"bound": "hbm", "kernel": "k_backward", "achieved": gb_dense, "peak": hbm_peak, "unit": "GB/s",
                    "frac": gb_dense / hbm_peak, "traffic": traffic, "peak_source": hbm_src,
                    "traffic_source": f"ncu capture ({ncu_src}): bytes per timestep-KKT x timestep-KKTs per launch of this run" if traffic else None,
                    "algorithmic_bytes_per_launch": kkt_bytes_dense(nx, nu, nc) * kkt_rank / launches_bw,
                    "launches": launches_bw, "avg_launch_ms": agg["ms_backward"] / launches_bw,
                    "region": "the timed region `value` is measured on (per-kernel CUDA events on the launching stream)",
                    "note": "achieved = SURVEY 8(d) dense-tile bytes per timestep-KKT x KKT steps / backward-kernel time; "
                            f"this layout moves {kkt_bytes_compact(nx, nu, nc, slots)} B per KKT step (compact tile), i.e. "
                            f"{gb_compact:.1f} GB/s actual"}
        roof_fp64 = {"bound": "fp64", "kernel": "k_backward", "achieved": tf, "peak": fp64_peak, "unit": "TFLOP/s",
                     "frac": tf / fp64_peak if fp64_peak > 0 else None, "peak_source": "DFMA microbenchmark, measured live",
                     "flops_per_kkt_step": flops_kkt}
        nst = agg["backward_calls"] * (N - 1)
        td = agg["ms_derivs"] * 1e-3
        tfw = agg["ms_forward"] * 1e-3
        d_bytes = 8 * (slots + nx + nu + nc)
        f_bytes = 8 * ((nu + nc + 2 * nu) * (nx + 1) + (nx + 3 * nu + nc) + (nx + 5 * nu + 2 * nc))
        ktot = agg["ms_derivs"] + agg["ms_backward"] + agg["ms_check"] + agg["ms_forward"]
        kernels = {
            "k_derivs": {"ms": agg["ms_derivs"], "share": agg["ms_derivs"] / ktot,
                         "GBps_compact_layout": d_bytes * nst / td / 1e9 if td > 0 else None,
                         "frac_hbm": d_bytes * nst / td / 1e9 / hbm_peak if td > 0 else None},
            "k_backward": {"ms": agg["ms_backward"], "share": agg["ms_backward"] / ktot, "kkt_steps_per_s": kkt_rank / tb},
            "k_check": {"ms": agg["ms_check"], "share": agg["ms_check"] / ktot},
            "k_forward": {"ms": agg["ms_forward"], "share": agg["ms_forward"] / ktot, "rollouts": agg["rollouts"],
                          "GBps": f_bytes * (N - 1) * agg["rollouts"] / tfw / 1e9 if tfw > 0 else None,
                          "frac_hbm": f_bytes * (N - 1) * agg["rollouts"] / tfw / 1e9 / hbm_peak if tfw > 0 else None},
            "host_gaps_ms": agg["ms_total"] - ktot,
        }
        cpu = None
        parity = {"parity_checked": 0, "parity_mismatches": None}
        if world == 1 and not args.no_cpu_baseline:
            import oracle
            cores = os.cpu_count() or 1
            sample = max(cores, min(B, args.cpu_sample if args.cpu_sample > 0 else cores * 128))
            sb = batch.slice(0, sample)
            t0 = time.perf_counter()
            ores, oxs, ous = oracle.solve_batch(args.workload, N, sb.p, sb.lower, sb.upper, sb.x1, sb.ubar,
                                                options=oracle.default_options(optimality_tolerance=args.tol),
                                                horizons=sb.horizons, nthreads=cores, want_traj=True)
            dt = time.perf_counter() - t0
            oc = sum(1 for r in ores if r.status == 0)
            cpu = {"value": oc / dt, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": f"first {sample} instances of the step's batch, one pass, {cores} OpenMP threads, {dt:.1f} s "
                             "(trajectories returned: the same results are the checker of parity_checked)",
                   "kkt_steps_per_s": sum(r.n_kkt for r in ores) / dt}
            # the GPU results of the e2e region (host buffers), first step, same instances: bitwise against the oracle
            gs, gk, gobj = hoi[0].numpy()[:sample], hoi[1].numpy()[:sample], hod[0].numpy()[:sample]
            gkkt = hoc[2].numpy()[:sample]
            gx, gu = hx.numpy()[:sample], hu.numpy()[:sample]
            bad = 0
            for i, o in enumerate(ores):
                same = (int(gs[i]) == o.status and int(gk[i]) == o.k and int(gkkt[i]) == o.n_kkt
                        and np.float64(gobj[i]).view(np.int64) == np.float64(o.objective).view(np.int64))
                if same:
                    xa, xb = gx[i], np.asarray(oxs[i]).reshape(gx[i].shape)
                    ua, ub = gu[i], np.asarray(ous[i]).reshape(gu[i].shape)
                    same = bool((((xa.view(np.int64) == xb.view(np.int64)) | ((xa == 0) & (xb == 0))).all())
                                and (((ua.view(np.int64) == ub.view(np.int64)) | ((ua == 0) & (ub == 0))).all()))
                bad += 0 if same else 1
            parity = {"parity_checked": sample, "parity_mismatches": bad,
                      "parity_what": "status, k, timestep-KKT count, objective bits, state and control trajectory bits of the GPU "
                                     "e2e results vs the oracle, instance by instance"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": args.warmup,
            "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{args.workload} swing-up batch of {B} random initial states per GPU, N={N} knots, tol {args.tol:g}",
                       "batch_per_gpu": B, "knots": N, "resident_slots": S, "handles": 1,
                       "device_memory_gb_per_gpu": mem_used_gb,
                       "l2": "working set (trajectories+gains > 8 GB) far exceeds the 126 MB L2",
                       "parallelism": f"batch sharded over {world} GPU(s), no data-path collective",
                       "timing": "value/e2e: the K steps (K x batch instances) are queued on ONE handle with resident_slots instance "
                                 "slots (ipddp_solve_queue: a slot whose instance terminated is refilled before the next round); "
                                 "value = device time by CUDA events on the library's stream from first enqueue to last completion, "
                                 "inputs and outputs in HBM; e2e = wall clock around the same call with pinned HOST buffers; "
                                 "roofline/kernels/lockstep come from the per-kernel CUDA events of the `value` region itself"},
            "sequential": ({"value": conv_single / (ms_single * 1e-3), "ms_per_step": ms_single, "steps": 1,
                            "rounds": single["rounds"], "mean_active_fraction": single["active_rounds"] / max(1, single["rounds"]) / min(S, B),
                            "what": "ONE batch with nothing queued behind it: the lock-step tail of its slowest instances is exposed"}
                           if single else None),
            "converged_fraction": conv_all / (K * B * world), "mean_iterations": ksum_all / (K * B * world), "max_primal_inf": pr_max,
            "backward_kkt_steps_per_s": kkt_all / (ms_total * 1e-3),
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": int(launches_all),
            "clocks": clocks,
            "roofline": roof_hbm, "roofline_fp64": roof_fp64, "kernels": kernels,
            "lockstep": {"rounds": agg["rounds"], "rounds_per_step": agg["rounds"] / K,
                         "mean_active_fraction": agg["active_rounds"] / max(1, agg["rounds"]) / S,
                         "work_weighted_active_fraction": agg["active_sq"] / max(1, agg["active_rounds"]) / S,
                         "note": "mean_active_fraction averages over rounds (the ~1000 nearly empty rounds of the final tail count "
                                 "like full ones); work_weighted = the occupancy of the round the average instance-iteration ran in"},
            "cpu_baseline": cpu,
            "configs": cfg_lines,
        }
        line.update(parity)
        print(json.dumps(line))
    solver.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
