"""Small solves that reach every kernel of the library: bulk and speculative kernels, the heavy forward split, queue mode
(k_admit / k_retire), warm start, varying horizons and a stage chain.  Iterations are capped: coverage, not convergence.
Run against the AddressSanitizer build of the CPU emulator by tests/test_emu_parity.py::test_emulated_memcheck_asan
(compute-sanitizer is closed on the GPU pool), and plain on the GPU as a launch-coverage check.
    IPDDP_LIB=tests/emu/libipddp_emu_asan.so LD_PRELOAD=$(gcc -print-file-name=libasan.so) python tools/sanitize_small.py [max_iterations]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import ipddp_b200  # noqa: E402,F401
from ipddp_b200 import _lib, instances  # noqa: E402
from ipddp_b200.batch import BatchSolver  # noqa: E402
from ipddp_b200.codegen import workloads  # noqa: E402
from helpers import chain_inputs  # noqa: E402

MAXIT = int(sys.argv[1]) if len(sys.argv) > 1 else 12
lib = _lib.Lib(os.environ["IPDDP_LIB"]) if os.environ.get("IPDDP_LIB") else _lib.load()   # IPDDP_LIB: the CPU emulator, for a dry run
launches = 0


def run(model, B, N, tuning=(), vary=False, queue=0, warm=False):
    global launches
    opt = lib.default_options(optimality_tolerance=1e-7, max_iterations=MAXIT)
    s = BatchSolver(model, B, N, options=opt, lib=lib)
    for k, v in tuning:
        s.set_tuning(k, v)
    n = queue or B
    b = instances.make_batch(model, n, N, vary_horizon=vary, horizon_span=min(40, N // 2))
    if queue:
        s.solve_queue(b.x1, b.ubar, b.p, b.lower, b.upper, b.horizons)
    else:
        s.set_batch(b)
        s.solve()
        if warm:
            s.solve(warm_start=True)
        s.trajectory()
    st = s.stats()
    launches += st.launches
    print(f"{model:14s} B={B} N={N} tuning={dict(tuning)} queue={queue} warm={warm}: rounds={st.iterations} launches={st.launches}",
          flush=True)
    s.close()


bulk = (("fw_spec_max", 0), ("bw_spec_max", 0))
run("cartpole", 6, 31)                                             # speculative tail kernels
run("cartpole", 6, 31, tuning=bulk, warm=True)                     # bulk kernels (TMA-staged rollout), warm start
run("cartpole", 6, 31, tuning=(("fw_spec_max", 2), ("bw_spec_max", 0)))   # heavy forward split
run("cartpole", 4, 31, queue=10)                                   # k_admit / k_retire
run("cartpole", 4, 31, tuning=bulk, queue=10)
run("pushing", 5, 41, vary=True)                                   # varying horizons, complementarity constraints
run("pushing", 5, 41, tuning=bulk, vary=True)
run("acrobot", 3, 41, tuning=bulk)
run("concar", 3, 31)

for tuning in ((), bulk):
    chain = workloads.get_chain("ragged")
    B, N = 3, 13
    st_types, x1, ubar, lower, upper = chain_inputs(chain, B, N)
    s = BatchSolver("ragged", B, N, options=lib.default_options(optimality_tolerance=1e-7, max_iterations=MAXIT), lib=lib)
    for k, v in tuning:
        s.set_tuning(k, v)
    s.set_stage_types(st_types)
    s.set_inputs(x1, ubar, None, lower, upper)
    s.solve()
    launches += s.stats().launches
    print(f"ragged chain    B={B} N={N} tuning={dict(tuning)}: rounds={s.stats().iterations}", flush=True)
    s.solve_queue(np.tile(x1, (2, 1)), np.tile(ubar, (2, 1)), None, np.tile(lower, (2, 1)), np.tile(upper, (2, 1)))
    launches += s.stats().launches
    s.close()
print(f"sanitize_small done: {launches} kernel launches")
