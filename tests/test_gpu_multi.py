"""Multi-GPU shard equivalence: N ranks (one process per GPU, NCCL) solving contiguous shards give the same
per-instance answers and the same reduced statistics as one GPU solving the whole batch.  Skipped with < 2 GPUs."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, json
sys.path.insert(0, {root!r})
import numpy as np, torch, torch.distributed as dist
import ipddp_b200
from ipddp_b200 import _lib, parallel
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
lib = _lib.load()
r, red = parallel.solve_sharded("acrobot", 96, 101, options=lib.default_options(optimality_tolerance=1e-7), device=lr, rank=rank, world=world)
out = [None] * world
dist.all_gather_object(out, (rank, r.k.tolist(), r.status.tolist(), r.objective.tolist()))
if rank == 0:
    print("RESULT " + json.dumps(dict(red=red, parts=out)))
dist.destroy_process_group()
'''


def test_two_gpu_shards_match_single_gpu(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from ipddp_b200 import _lib, parallel
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29547", str(script)]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stdout + p.stderr
    line = [l for l in p.stdout.splitlines() if l.startswith("RESULT ")][0]
    got = json.loads(line[len("RESULT "):])
    lib = _lib.load()
    r, red1 = parallel.solve_sharded("acrobot", 96, 101, options=lib.default_options(optimality_tolerance=1e-7))
    ks = sum((part[1] for part in sorted(got["parts"])), [])
    objs = sum((part[3] for part in sorted(got["parts"])), [])
    assert ks == r.k.tolist()
    assert np.array_equal(np.array(objs).view(np.int64), r.objective.view(np.int64))
    for name in parallel.STAT_SUM:
        assert got["red"][name] == red1[name], name
