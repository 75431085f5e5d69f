"""Tiny solves through every kernel of the path -- bulk and tail (speculative) variants, a K > 32 model (cartpole: two
lane slots), a model with second-order dynamics terms (concar), ragged horizons (pushing) -- printing a digest of the
results per (workload, kernel variant).  No torch import.

Used as the race / memory sanity check of the round: the digests of a run on the B200 must equal those of the SIMT
emulator (tests/emu) in all three lane orders,

    python tools/sanitize.py                                                           # on the GPU box
    IPDDP_EMU_ORDER=fwd|rev|rand IPDDP_LIB=tests/emu/libipddp_emu.so python tools/sanitize.py   # here

(profiles/r1_sanity/ holds both sides).  compute-sanitizer is closed on this GPU pool, so memcheck / racecheck are
replaced by this comparison plus tests/test_emu_parity.py::test_emulated_lane_order_independence; on a box where the
sanitizer is allowed the same command runs under `compute-sanitizer --tool racecheck`.

Each argument is workload:B:N:max_iterations."""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import ipddp_b200  # noqa: E402,F401
from ipddp_b200 import _lib, instances  # noqa: E402
from ipddp_b200.batch import BatchSolver  # noqa: E402


def main(argv):
    lib = _lib.Lib(os.environ["IPDDP_LIB"]) if os.environ.get("IPDDP_LIB") else _lib.load()   # (emulator dry run)
    specs = argv or ["cartpole:3:9:25", "concar:3:11:30", "pushing:3:9:20"]
    for spec in specs:
        wl, B, N, iters = spec.split(":")
        B, N, iters = int(B), int(N), int(iters)
        for bw, fw, label in ((592, 148, "tail"), (0, 0, "bulk")):
            lib.L.ipddp_set_tuning(None, b"bw_spec_max", bw)
            lib.L.ipddp_set_tuning(None, b"fw_spec_max", fw)
            b = instances.make_batch(wl, B, N)
            if wl == "pushing":
                b.horizons = np.array([N - (i % 3) for i in range(B)], dtype=np.int32)
            s = BatchSolver(wl, B, N, options=lib.default_options(optimality_tolerance=1e-7, max_iterations=iters), lib=lib)
            s.set_batch(b)
            r = s.solve()
            x, u = s.trajectory()
            h = hashlib.sha256(np.ascontiguousarray(x).tobytes() + np.ascontiguousarray(u).tobytes()
                               + np.ascontiguousarray(r.k).tobytes()).hexdigest()[:16]
            print(json.dumps(dict(workload=wl, B=B, N=N, kernels=label, status=[int(v) for v in r.status],
                                  k=[int(v) for v in r.k], digest=h)), flush=True)
            s.close()


if __name__ == "__main__":
    main(sys.argv[1:])
