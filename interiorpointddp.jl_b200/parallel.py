"""Multi-GPU sharding of a batch.

Instances are independent, so the batch is split into contiguous shards, one per rank (one process per
GPU), with NO collective on the data path.  The only exchange is one reduction of a small vector of
convergence statistics at the end (NCCL all-reduce over NVLink when the tensors live on the GPU, gloo on
the CPU in the tests).
"""
from __future__ import annotations

from typing import Dict, Tuple

import numpy as np

STAT_SUM = ["instances", "converged", "iterations", "backward_passes", "sweeps", "kkt_steps", "rollouts",
            "status_0", "status_1", "status_7", "status_8", "status_9"]
STAT_MAX = ["max_primal_inf", "max_iterations", "device_ms"]


def shard_bounds(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous balanced split: the first (total % world) ranks get one extra instance."""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def local_stats(status, k, primal_inf, counters: Dict[str, np.ndarray], device_ms: float) -> Dict[str, float]:
    status = np.asarray(status)
    conv = status == 0
    s = {"instances": float(status.size), "converged": float(conv.sum()), "iterations": float(np.sum(k)),
         "backward_passes": float(np.sum(counters["n_backward"])), "sweeps": float(np.sum(counters["n_sweeps"])),
         "kkt_steps": float(np.sum(counters["n_kkt"])), "rollouts": float(np.sum(counters["n_rollouts"]))}
    for code in (0, 1, 7, 8, 9):
        s[f"status_{code}"] = float((status == code).sum())
    s["max_primal_inf"] = float(np.max(np.asarray(primal_inf)[conv])) if conv.any() else 0.0
    s["max_iterations"] = float(np.max(k)) if status.size else 0.0
    s["device_ms"] = float(device_ms)
    return s


def reduce_stats(stats: Dict[str, float], device=None, group=None) -> Dict[str, float]:
    """All-reduce (SUM for counts, MAX for extrema / times) over the default process group.
    Without an initialised process group this is the identity (single GPU)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return dict(stats)
    dev = device if device is not None else torch.device("cpu")
    vs = torch.tensor([stats[n] for n in STAT_SUM], dtype=torch.float64, device=dev)
    vm = torch.tensor([stats[n] for n in STAT_MAX], dtype=torch.float64, device=dev)
    dist.all_reduce(vs, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(vm, op=dist.ReduceOp.MAX, group=group)
    out = {n: float(v) for n, v in zip(STAT_SUM, vs.tolist())}
    out.update({n: float(v) for n, v in zip(STAT_MAX, vm.tolist())})
    return out


def solve_sharded(workload: str, total: int, N: int, options=None, device: int = 0, rank: int = 0, world: int = 1,
                  seed: int = 0, vary_horizon: bool = False, lib=None, reduce_device=None):
    """Each rank solves its shard of the workload's canonical instance stream on its own GPU and returns
    (local BatchResult, reduced statistics).  `lib` / `reduce_device`: the loaded C-ABI library and the device the
    statistics vector is reduced on (defaults: the CUDA library, this rank's GPU -> NCCL); the world-size-2 CPU test passes
    its emulator build and the CPU (gloo)."""
    import torch
    from . import instances
    from .batch import BatchSolver
    lo, hi = shard_bounds(total, rank, world)
    batch = instances.make_batch(workload, hi - lo, N, seed=seed, first=lo, vary_horizon=vary_horizon)
    s = BatchSolver(workload, hi - lo, N, options=options, device=device, lib=lib)
    s.set_batch(batch)
    r = s.solve()
    st = local_stats(r.status, r.k, r.primal_inf, s.counters(), s.stats().ms_total)
    s.close()
    red = reduce_stats(st, device=reduce_device if reduce_device is not None else torch.device("cuda", device))
    return r, red
