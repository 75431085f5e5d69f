#!/usr/bin/env python
"""bench.py -- headline benchmark of the batched IPDDP2 hot path (BASELINE.json).

metric   : converged OCP solves/sec (batched)            unit: solves/s
workload : BASELINE configs[1] = cartpole swing-up, batch of 16384 random initial states, N = 101 knots,
           optimality_tolerance 1e-7, on ONE B200 (per GPU; weak scaling over GPUs: every rank solves its
           own 16384 instances, no data-path collective, one NCCL all-reduce of statistics at the end).
step     : one complete batched solve (initialise + derivative / backward / check / forward rounds until
           every instance terminated) of one batch of synthetic instances.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B] [--workload W]

`value`  : device-timed (CUDA events on the library's stream), inputs already resident in HBM.
`e2e`    : the same metric through the C ABI with HOST buffers: pinned H2D of the inputs + solve + D2H of
           the SolverData scalars and the trajectories, inside the timed region.
--impl reference : the CPU restatement of the reference (oracle/, OpenMP over instances on all host cores)
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


def run_reference(args, rank, world):
    """CPU arm: the oracle (literal restatement of the reference, -O2, OpenMP over instances) on all host cores."""
    if rank != 0:
        return
    import oracle
    import ipddp_b200  # noqa: F401
    from ipddp_b200 import instances
    cores = os.cpu_count() or 1
    sample = max(cores, min(args.batch, args.cpu_sample if args.cpu_sample > 0 else cores * 48))
    b = instances.make_batch(args.workload, sample, args.knots)
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
    sample_txt = f"first {sample} instances of the workload per step, {cores} OpenMP threads"
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
    ap.add_argument("--batch", type=int, default=16384, help="instances per GPU")
    ap.add_argument("--workload", default="cartpole")
    ap.add_argument("--knots", type=int, default=101)
    ap.add_argument("--tol", type=float, default=1e-7)
    ap.add_argument("--cpu-sample", type=int, default=0, help="instances of the CPU baseline sample (0 = 128 x cores for cpu_baseline, 48 x cores per step for --impl reference)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--inflight", type=int, default=8, help="independent batches in flight during the timed steps")
    ap.add_argument("--seq-steps", type=int, default=2,
                    help="steps of the sequential region (one batch at a time, per-kernel CUDA events: rooflines)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import ipddp_b200  # noqa: F401
    from ipddp_b200 import _lib, instances
    from ipddp_b200.batch import BatchSolver, solve_many

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    nx, nu, nc, npar, slots = lib.model_dims(args.workload)
    B, N = args.batch, args.knots

    # every rank takes its own contiguous range of the workload's canonical instance stream
    batch = instances.make_batch(args.workload, B, N, first=rank * B)
    opt = lib.default_options(optimality_tolerance=args.tol)
    F = max(1, min(args.inflight, args.steps))
    solvers = [BatchSolver(args.workload, B, N, options=opt, device=local_rank, lib=lib) for _ in range(F)]
    solver = solvers[0]

    # pinned host copies (e2e path) and device-resident copies (value path)
    def pin(a):
        return torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    h = {k: pin(v) for k, v in dict(x1=batch.x1, ubar=batch.ubar, p=batch.p if npar > 0 else np.zeros((B, 1)),
                                     lower=batch.lower, upper=batch.upper).items()}
    h_hz = torch.from_numpy(batch.horizons.astype(np.int32)).pin_memory()
    d = {k: t.to(dev, non_blocking=True) for k, t in h.items()}
    d_hz = h_hz.to(dev)
    torch.cuda.synchronize()

    def set_device_inputs(sv):
        sv.set_inputs_device(d["x1"].data_ptr(), d["ubar"].data_ptr(), d["p"].data_ptr() if npar > 0 else None,
                             d["lower"].data_ptr(), d["upper"].data_ptr(), d_hz.data_ptr())

    def set_host_inputs(sv):
        sv.lib.check(sv.lib.L.ipddp_set_inputs(
            sv.h, _lib.dptr(h["x1"].numpy()), _lib.dptr(h["ubar"].numpy()),
            _lib.dptr(h["p"].numpy()) if npar > 0 else None, _lib.dptr(h["lower"].numpy()), _lib.dptr(h["upper"].numpy()),
            _lib.iptr(h_hz.numpy())), "ipddp_set_inputs")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for sv in solvers:
        set_device_inputs(sv)
    # warm-up: W untimed steps through the same pipelined path as the timed region, plus one sequential solve
    if args.warmup > 0:
        solve_many(solvers[:min(F, args.warmup)], total_solves=args.warmup)
        solver.solve()

    # ---------------------------------------------------- timed region A: K steps one after another (kernels timed alone)
    agg = dict(ms_total=0.0, ms_derivs=0.0, ms_backward=0.0, ms_check=0.0, ms_forward=0.0, ms_init=0.0, kkt=0, sweeps=0,
               rollouts=0, deriv=0, conv=0, launches=0, rounds=0, active_rounds=0, backward_calls=0)
    seq_steps = max(1, min(args.steps, args.seq_steps))
    barrier()
    for _ in range(seq_steps):
        set_device_inputs(solver)
        solver.solve()
        st = solver.stats()
        agg["ms_total"] += st.ms_total; agg["ms_derivs"] += st.ms_derivs; agg["ms_backward"] += st.ms_backward
        agg["ms_check"] += st.ms_check; agg["ms_forward"] += st.ms_forward; agg["ms_init"] += st.ms_init
        agg["kkt"] += st.sum_kkt; agg["sweeps"] += st.sum_sweeps; agg["rollouts"] += st.sum_rollouts
        agg["deriv"] += st.sum_deriv_stages; agg["conv"] += st.n_converged; agg["launches"] += st.launches
        agg["rounds"] += st.iterations; agg["active_rounds"] += st.n_active_rounds; agg["backward_calls"] += st.sum_backward
    barrier()
    res = solver.results()

    # ---------------------------------------------------- timed region B: the same K steps, up to F batches in flight
    # (independent problem handles on their own streams, ipddp_solve_many): whole-job throughput, inputs resident in HBM
    barrier()
    with ClockSampler(local_rank) as clk:
        ms_pipe, st_pipe = solve_many(solvers, total_solves=args.steps)
        barrier()
    clocks = clk.summary()

    # ---------------------------------------------------- timed region C: end to end through host buffers, F in flight
    hxs = [torch.from_numpy(np.zeros((B, N, nx))).pin_memory() for _ in range(F)]
    hus = [torch.from_numpy(np.zeros((B, N - 1, nu))).pin_memory() for _ in range(F)]
    barrier()
    t0 = time.perf_counter()
    conv_e2e = 0
    done = 0
    while done < args.steps:
        nb = min(F, args.steps - done)
        for sv in solvers[:nb]:
            set_host_inputs(sv)
        solve_many(solvers[:nb], total_solves=nb)
        for q, sv in enumerate(solvers[:nb]):
            r = sv.results()
            sv.lib.check(sv.lib.L.ipddp_get_trajectory(sv.h, _lib.dptr(hxs[q].numpy()), _lib.dptr(hus[q].numpy())), "get_trajectory")
            conv_e2e += int((r.status == 0).sum())
        done += nb
    barrier()
    wall_e2e = time.perf_counter() - t0
    h2d = sum(int(t.numel() * t.element_size()) for k_, t in h.items() if not (npar == 0 and k_ == "p")) + int(h_hz.numel() * 4)
    d2h = B * (4 * 4 + 7 * 8) + int(hxs[0].numel() * 8) + int(hus[0].numel() * 8)

    # ------------------------------------------------------------------ reduce over ranks
    vals = torch.tensor([agg["ms_total"], wall_e2e, ms_pipe], dtype=torch.float64, device=dev)
    sums = torch.tensor([agg["conv"], conv_e2e, st_pipe.sum_kkt, float(st_pipe.n_converged), agg["rollouts"], st_pipe.launches,
                         int((res.status == 0).sum()), float(res.k.sum()), float(res.primal_inf[res.status == 0].max(initial=0.0))],
                        dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(vals, op=dist.ReduceOp.MAX)      # time = max over ranks
        mx = sums[-1:].clone()
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)      # the single NCCL reduction of convergence statistics
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sums[-1] = mx[0]
    ms_total, wall_e2e_max, ms_pipe_max = [float(x) for x in vals.tolist()]
    conv_all, conv_e2e_all, kkt_pipe_all, conv_pipe_all, roll_all, launches_all, conv_last, ksum, pr_max = [float(x) for x in sums.tolist()]

    if rank == 0:
        value = conv_pipe_all / (ms_pipe_max * 1e-3)
        value_sequential = conv_all / (ms_total * 1e-3)   # over the seq_steps sequential steps
        e2e_val = conv_e2e_all / wall_e2e_max
        # roofline of the dominant kernel (backward sweep) on rank 0
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        hbm_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
        fp64_peak = float(lib.L.ipddp_measure_fp64_tflops(local_rank))
        flops_kkt = kkt_flops(nx, nu, nc)
        tb = agg["ms_backward"] * 1e-3
        kkt_rank = agg["kkt"]
        gb_dense = kkt_bytes_dense(nx, nu, nc) * kkt_rank / tb / 1e9
        gb_compact = kkt_bytes_compact(nx, nu, nc, slots) * kkt_rank / tb / 1e9
        tf = flops_kkt * kkt_rank / tb / 1e12
        # dram__bytes_read+write of one k_backward launch from the committed ncu --set full capture
        # (profiles/r1_backward_summary.md): bytes per KKT step x the KKT steps of an average launch here
        ncu_bytes_per_kkt = None
        try:
            ncu_bytes_per_kkt = float(json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))[args.workload]["dram_bytes_per_kkt_step"])
        except Exception:
            pass
        launches_bw = max(1, agg["rounds"])
        traffic = ncu_bytes_per_kkt * kkt_rank / launches_bw if ncu_bytes_per_kkt else None
        roof_hbm = {"bound": "hbm", "kernel": "k_backward", "achieved": gb_dense, "peak": hbm_peak, "unit": "GB/s",
                    "frac": gb_dense / hbm_peak, "traffic": traffic, "peak_source": hbm_src,
                    "algorithmic_bytes_per_launch": kkt_bytes_dense(nx, nu, nc) * kkt_rank / launches_bw,
                    "note": "achieved = SURVEY 8(d) dense-tile bytes per timestep-KKT x KKT steps / backward-kernel time; "
                            f"this layout moves {kkt_bytes_compact(nx, nu, nc, slots)} B per KKT step (compact tile), i.e. "
                            f"{gb_compact:.1f} GB/s actual"}
        roof_fp64 = {"bound": "fp64", "kernel": "k_backward", "achieved": tf, "peak": fp64_peak, "unit": "TFLOP/s",
                     "frac": tf / fp64_peak if fp64_peak > 0 else None, "peak_source": "DFMA microbenchmark, measured live",
                     "flops_per_kkt_step": flops_kkt}
        nst = agg["deriv"] * (N - 1) if agg["deriv"] else 0
        td = agg["ms_derivs"] * 1e-3
        tfw = agg["ms_forward"] * 1e-3
        d_bytes = 8 * (slots + nx + nu + nc)
        f_bytes = 8 * ((nu + nc + 2 * nu) * (nx + 1) + (nx + 3 * nu + nc) + (nx + 5 * nu + 2 * nc))
        kernels = {
            "k_derivs": {"ms": agg["ms_derivs"], "GBps_compact_layout": d_bytes * nst / td / 1e9 if td > 0 else None,
                         "frac_hbm": d_bytes * nst / td / 1e9 / hbm_peak if td > 0 else None},
            "k_backward": {"ms": agg["ms_backward"], "kkt_steps_per_s": kkt_rank / tb},
            "k_check": {"ms": agg["ms_check"]},
            "k_forward": {"ms": agg["ms_forward"], "rollouts": agg["rollouts"],
                          "GBps": f_bytes * (N - 1) * agg["rollouts"] / tfw / 1e9 if tfw > 0 else None,
                          "frac_hbm": f_bytes * (N - 1) * agg["rollouts"] / tfw / 1e9 / hbm_peak if tfw > 0 else None},
            "k_init": {"ms": agg["ms_init"]},
        }
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            import oracle
            cores = os.cpu_count() or 1
            sample = max(cores, min(B, args.cpu_sample if args.cpu_sample > 0 else cores * 128))
            sb = instances.make_batch(args.workload, sample, N)
            t0 = time.perf_counter()
            ores, _, _ = oracle.solve_batch(args.workload, N, sb.p, sb.lower, sb.upper, sb.x1, sb.ubar,
                                            options=oracle.default_options(optimality_tolerance=args.tol),
                                            horizons=sb.horizons, nthreads=cores)
            dt = time.perf_counter() - t0
            oc = sum(1 for r in ores if r.status == 0)
            cpu = {"value": oc / dt, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": f"first {sample} instances of the workload, one pass, {cores} OpenMP threads, {dt:.1f} s",
                   "kkt_steps_per_s": sum(r.n_kkt for r in ores) / dt}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_pipe_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{args.workload} swing-up batch of {B} random initial states per GPU, N={N} knots, tol {args.tol:g}",
                       "batch_per_gpu": B, "knots": N, "l2": "working set (trajectories+gains > 8 GB) far exceeds the 126 MB L2",
                       "parallelism": f"batch sharded over {world} GPU(s), no data-path collective",
                       "steps_in_flight": F,
                       "timing": "value/e2e: the K steps run with up to steps_in_flight independent batches in flight on their own "
                                 "streams (ipddp_solve_many), device time by CUDA events from first enqueue to last completion; "
                                 "roofline/kernels/sequential: sequential.steps of the same steps one after another, per-kernel "
                                 "CUDA events on the launching stream"},
            "sequential": {"value": value_sequential, "ms_per_step": ms_total / seq_steps, "steps": seq_steps},
            "converged_fraction": conv_last / (B * world), "mean_iterations": ksum / (B * world), "max_primal_inf": pr_max,
            "backward_kkt_steps_per_s": kkt_pipe_all / (ms_pipe_max * 1e-3),
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": int(launches_all),
            "clocks": clocks,
            "roofline": roof_hbm, "roofline_fp64": roof_fp64, "kernels": kernels,
            "lockstep": {"rounds_per_step": agg["rounds"] / seq_steps,
                         "mean_active_fraction": agg["active_rounds"] / max(1, agg["rounds"]) / B},
            "cpu_baseline": cpu,
        }
        print(json.dumps(line))
    for sv in solvers:
        sv.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
