"""World-size-2 check of the sharding host logic on CPU (gloo): the shards are disjoint, cover the batch, rebuild
exactly the instances of the unsharded stream, and the single statistics all-reduce reproduces the single-process
numbers.  The per-shard solves are done by the oracle here (no GPU in this container); on the GPU the same
`parallel.py` functions wrap BatchSolver (tests/test_gpu_multi.py)."""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    import socket
    with socket.socket() as so:
        so.bind(("127.0.0.1", 0))
        return str(so.getsockname()[1])

WORKER = r'''
import os, sys, json
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "oracle"))
import numpy as np, torch, torch.distributed as dist
import ipddp_b200
from ipddp_b200 import instances, parallel
import oracle
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
total, N, wl = 7, 21, "concar"
lo, hi = parallel.shard_bounds(total, rank, world)
b = instances.make_batch(wl, hi - lo, N, first=lo)
res, _, _ = oracle.solve_batch(wl, N, b.p, b.lower, b.upper, b.x1, b.ubar, options=oracle.default_options(optimality_tolerance=1e-7))
cnt = dict(n_backward=np.array([r.n_backward for r in res]), n_sweeps=np.array([r.n_sweeps for r in res]),
           n_kkt=np.array([r.n_kkt for r in res]), n_rollouts=np.array([r.n_rollouts for r in res]))
st = parallel.local_stats([r.status for r in res], [r.k for r in res], [r.primal_inf for r in res], cnt, 1.0 + rank)
red = parallel.reduce_stats(st)
x1_all = [None] * world
dist.all_gather_object(x1_all, (lo, hi, b.x1.tolist(), b.p.tolist()))
if rank == 0:
    print("RESULT " + json.dumps(dict(red=red, shards=x1_all)))
dist.destroy_process_group()
'''


def test_two_rank_sharding_and_reduction(tmp_path):
    import ipddp_b200  # noqa: F401
    from ipddp_b200 import instances, parallel
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=_free_port(), WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r)), stdout=subprocess.PIPE,
                              stderr=subprocess.PIPE, text=True) for r in range(2)]
    outs = [p.communicate(timeout=600) for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    import json
    line = [l for l in outs[0][0].splitlines() if l.startswith("RESULT ")][0]
    got = json.loads(line[len("RESULT "):])
    # shards: disjoint, covering, identical to the unsharded stream
    total, N, wl = 7, 21, "concar"
    full = instances.make_batch(wl, total, N)
    spans = sorted((s[0], s[1]) for s in got["shards"])
    assert spans == [(0, 4), (4, 7)]
    for lo, hi, x1, p in got["shards"]:
        assert np.array_equal(np.array(x1), full.x1[lo:hi]) and np.array_equal(np.array(p), full.p[lo:hi])
    # reduction == single-process statistics
    res, _, _ = oracle.solve_batch(wl, N, full.p, full.lower, full.upper, full.x1, full.ubar,
                                   options=oracle.default_options(optimality_tolerance=1e-7))
    red = got["red"]
    assert red["instances"] == total
    assert red["converged"] == sum(1 for r in res if r.status == 0)
    assert red["iterations"] == sum(r.k for r in res)
    assert red["kkt_steps"] == sum(r.n_kkt for r in res)
    assert red["rollouts"] == sum(r.n_rollouts for r in res)
    assert red["max_iterations"] == max(r.k for r in res)
    assert red["device_ms"] == 2.0      # MAX over ranks of the per-rank time


def test_shard_bounds_properties():
    from ipddp_b200 import parallel
    for total in (0, 1, 7, 16384, 8192, 100003):
        for world in (1, 2, 3, 4, 8):
            spans = [parallel.shard_bounds(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


WORKER_SHARDED = r'''
import os, sys, json
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests", "emu"))
import numpy as np, torch, torch.distributed as dist
import ipddp_b200
from ipddp_b200 import _lib, parallel
import build_emu
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
lib = _lib.Lib(build_emu.LIB)
opt = lib.default_options(optimality_tolerance=1e-7, max_iterations=60)
r, red = parallel.solve_sharded("concar", 5, 11, options=opt, rank=rank, world=world, lib=lib,
                                reduce_device=torch.device("cpu"))
out = [None] * world
dist.all_gather_object(out, dict(status=r.status.tolist(), k=r.k.tolist(), objective=[float(x).hex() for x in r.objective]))
if rank == 0:
    print("RESULT " + json.dumps(dict(red=red, shards=out)))
dist.destroy_process_group()
'''


def test_two_rank_solve_sharded_on_the_emulator(tmp_path, oracle_mod):
    """`parallel.solve_sharded` -- the function every rank runs on its GPU (tests/test_gpu_multi.py, bench.py's configs) --
    with world size 2 over gloo: each rank solves its shard through the C ABI of the emulator build of the kernel sources,
    the statistics are reduced once; the shards together equal the oracle's solve of the whole batch bit for bit."""
    sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
    import build_emu
    build_emu.build()                      # once, before the ranks start
    import ipddp_b200  # noqa: F401
    from ipddp_b200 import instances
    script = tmp_path / "worker_sharded.py"
    script.write_text(WORKER_SHARDED.format(root=ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=_free_port(), WORLD_SIZE="2")
    env.pop("IPDDP_EMU_ORDER", None)
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r)), stdout=subprocess.PIPE,
                              stderr=subprocess.PIPE, text=True) for r in range(2)]
    outs = [p.communicate(timeout=900) for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    import json
    got = json.loads([l for l in outs[0][0].splitlines() if l.startswith("RESULT ")][0][len("RESULT "):])
    total, N, wl = 5, 11, "concar"
    full = instances.make_batch(wl, total, N)
    res, _, _ = oracle_mod.solve_batch(wl, N, full.p, full.lower, full.upper, full.x1, full.ubar,
                                       options=oracle_mod.default_options(optimality_tolerance=1e-7, max_iterations=60))
    status = sum((s["status"] for s in got["shards"]), [])
    ks = sum((s["k"] for s in got["shards"]), [])
    objs = sum((s["objective"] for s in got["shards"]), [])
    assert status == [r.status for r in res] and ks == [r.k for r in res]
    assert objs == [float(r.objective).hex() for r in res]
    red = got["red"]
    assert red["instances"] == total and red["converged"] == sum(1 for r in res if r.status == 0)
    assert red["iterations"] == sum(r.k for r in res) and red["kkt_steps"] == sum(r.n_kkt for r in res)
    assert red["rollouts"] == sum(r.n_rollouts for r in res) and red["max_iterations"] == max(r.k for r in res)
    assert sum(red[f"status_{c}"] for c in (0, 1, 7, 8, 9)) == total
