"""Synthetic instance batches for the six reference workloads.

ONE generator shared by the GPU harness, the tests and the CPU oracle baseline, so both sides solve
bit-identical inputs.  Instance i < 100 reproduces the reference's seed i+1 from the committed params
tables (tests/golden/params/*.txt; Julia's Xoshiro draws cannot be regenerated without Julia); instance
i >= 100 draws from the same ranges (reference experiments/ipddp2/*.jl) with a counter-based SplitMix64
stream keyed by (seed, instance, draw).  The params tables ship with the package (data/params, byte-identical to
tests/golden/params, which tests/test_oracle_golden.py checks); the results tables are test fixtures and stay under tests/.
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass
from typing import Optional

import numpy as np

from .codegen import workloads

_HERE = os.path.dirname(os.path.abspath(__file__))
_REPO = os.path.dirname(_HERE)
PARAMS_DIR = os.path.join(_HERE, "data", "params")
RESULTS_DIR = os.path.join(_REPO, "tests", "golden", "results")   # known answers: only tests and tools read them

_PARAM_FILE = {"cartpole": "cartpole_friction", "acrobot": "acrobot_contact", "concar": "concar",
               "concar_quad": "concar", "pushing": "pushing_1_obs"}
_RESULT_FILE = {"cartpole": "cartpole_friction", "acrobot": "acrobot_contact", "concar": "concar",
                "concar_quad": "concar_quad", "pushing": "pushing_1_obs", "double_integrator": "double_integrator"}


def _splitmix64(x: np.ndarray) -> np.ndarray:
    x = (x + np.uint64(0x9E3779B97F4A7C15)).astype(np.uint64)
    z = x
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def uniform(seed: int, inst: np.ndarray, draw: int) -> np.ndarray:
    """U[0,1) doubles, one per instance, for draw index `draw` (counter based, order independent)."""
    with np.errstate(over="ignore"):
        key = (np.uint64(seed) * np.uint64(0xD1342543DE82EF95)
               + inst.astype(np.uint64) * np.uint64(0x2545F4914F6CDD1D)
               + np.uint64(draw) * np.uint64(0x9E3779B97F4A7C15))
        z = _splitmix64(_splitmix64(key))
    return (z >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def load_params_table(workload: str) -> Optional[np.ndarray]:
    f = _PARAM_FILE.get(workload)
    if f is None:
        return None
    path = os.path.join(PARAMS_DIR, f + ".txt")
    if not os.path.exists(path):
        return None
    return np.loadtxt(path, ndmin=2)


def load_golden_results(workload: str):
    """Returns dict(seed, iterations, converged, objective, primal) arrays from the reference table."""
    path = os.path.join(RESULTS_DIR, _RESULT_FILE[workload] + ".txt")
    seeds, its, conv, obj, pr = [], [], [], [], []
    with open(path) as fh:
        next(fh)
        for line in fh:
            t = line.split()
            if len(t) < 5:
                continue
            seeds.append(int(t[0])); its.append(int(t[1])); conv.append(t[2] == "true")
            obj.append(float(t[3])); pr.append(float(t[4]))
    return dict(seed=np.array(seeds), iterations=np.array(its), converged=np.array(conv),
                objective=np.array(obj), primal=np.array(pr))


@dataclass
class Batch:
    workload: str
    N: int                      # knots (max horizon)
    p: np.ndarray               # [B, np]
    lower: np.ndarray           # [B, nu]
    upper: np.ndarray           # [B, nu]
    x1: np.ndarray              # [B, nx]
    ubar: np.ndarray            # [B, (N-1)*nu]
    horizons: np.ndarray        # [B] int32 knots per instance (<= N)

    @property
    def B(self) -> int:
        return self.x1.shape[0]

    def slice(self, lo: int, hi: int) -> "Batch":
        return Batch(self.workload, self.N, self.p[lo:hi].copy(), self.lower[lo:hi].copy(),
                     self.upper[lo:hi].copy(), self.x1[lo:hi].copy(), self.ubar[lo:hi].copy(),
                     self.horizons[lo:hi].copy())


_PUSH_BLOCKS = np.array([
    [0.07, 0.12, 0.03711], [0.06, 0.12, 0.0355938], [0.08, 0.12, 0.0387237],
    [0.07, 0.13, 0.0393039], [0.06, 0.13, 0.0378424], [0.08, 0.13, 0.0366212],
    [0.07, 0.11, 0.0349493], [0.06, 0.11, 0.0333738], [0.08, 0.11, 0.0408633]])


def make_batch(workload: str, B: int, N: int = 101, seed: int = 0, use_reference_rows: bool = True,
               random_x1: bool = True, vary_horizon: bool = False, first: int = 0, horizon_span: int = 40) -> Batch:
    """Instances first..first+B-1 of the workload's canonical stream.  vary_horizon: per-instance horizons drawn uniformly
    from N - horizon_span .. N knots (reference rows keep N)."""
    md = workloads.get(workload)
    nx, nu = md.nx, md.nu
    inst = np.arange(first, first + B, dtype=np.int64)
    U = lambda d: uniform(seed, inst, d)
    x1 = np.zeros((B, nx))
    if workload == "cartpole":
        p = np.stack([0.9 + 0.2 * U(0), 0.15 + 0.1 * U(1), 0.45 + 0.1 * U(2), 0.05 + 0.1 * U(3), 0.05 + 0.1 * U(4)], 1)
        if random_x1:
            q = np.stack([-0.25 + 0.5 * U(5), -0.5 + 1.0 * U(6)], 1)
            x1 = np.concatenate([q, q], 1)
    elif workload == "acrobot":
        p = np.stack([0.9 + 0.2 * U(0), np.full(B, 0.333), 0.9 + 0.2 * U(1), np.full(B, 0.5),
                      0.9 + 0.2 * U(2), np.full(B, 0.333), 0.9 + 0.2 * U(3), np.full(B, 0.5)], 1)
    elif workload in ("concar", "concar_quad"):
        cols = [1.5 + U(0), 3.0 + 2.0 * U(1)]
        bases = [(0.25, 0.25), (0.75, 0.75), (0.25, 0.75), (0.75, 0.25)]
        d = 2
        for (bx, by) in bases:
            cols += [bx + (U(d) - 0.5) * 0.2, by + (U(d + 1) - 0.5) * 0.2, 0.05 + U(d + 2) * 0.15]
            d += 3
        p = np.stack(cols, 1)
        x1 = np.zeros((B, nx))
        x1[:, 2] = math.pi / 8 + U(d + 2) * (math.pi / 4)
    elif workload == "pushing":
        blk = _PUSH_BLOCKS[np.minimum((U(0) * 9).astype(np.int64), 8)]
        obs = np.stack([0.2 + 0.3 * (U(1) - 0.5), 0.2 + 0.1 * (U(2) - 0.5), 0.05 + 0.02 * (U(3) - 0.5)], 1)
        mu_f = 0.2 + 0.1 * (U(4) - 0.5)
        r_total = np.maximum(blk[:, 0], blk[:, 1]) + 0.01
        p = np.concatenate([blk, mu_f[:, None], obs, r_total[:, None]], 1)
    elif workload == "double_integrator":
        p = np.zeros((B, 0))
    else:
        raise KeyError(workload)

    tab = load_params_table(workload) if use_reference_rows else None
    if tab is not None:
        for i in range(B):
            g = first + i
            if g >= tab.shape[0]:
                break
            row = tab[g]
            if workload in ("cartpole", "acrobot"):
                p[i] = row
                x1[i] = 0.0
            elif workload in ("concar", "concar_quad"):
                p[i] = row[:14]
                x1[i] = row[14:18]
            elif workload == "pushing":
                rt = max(row[0], row[1]) + 0.01
                p[i] = np.concatenate([row[:7], [rt]])
                x1[i] = 0.0

    lower = np.array([md.lower(list(p[i])) for i in range(B)], dtype=np.float64).reshape(B, nu)
    upper = np.array([md.upper(list(p[i])) for i in range(B)], dtype=np.float64).reshape(B, nu)
    ubar = np.tile(np.asarray(md.u_init, dtype=np.float64), (B, N - 1))
    horizons = np.full(B, N, dtype=np.int32)
    if vary_horizon:
        lo_h = max(2, N - horizon_span)
        hz = lo_h + np.minimum((uniform(seed, inst, 99) * (N - lo_h + 1)).astype(np.int64), N - lo_h)
        horizons = hz.astype(np.int32)
        if tab is not None:
            for i in range(B):
                if first + i < tab.shape[0]:
                    horizons[i] = N
    return Batch(workload, N, np.ascontiguousarray(p), lower, upper, np.ascontiguousarray(x1), ubar, horizons)
