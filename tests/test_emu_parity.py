"""CPU-side logic checks of the CUDA kernel SOURCES: the product's .cu/.cuh files are compiled with g++ against
the fiber-based SIMT emulator in tests/emu (test infrastructure, never loaded by the package) and driven
through the same C ABI.  Small sizes only; the real parity tests run on the GPU (test_gpu_*.py)."""
import os
import sys

import numpy as np
import pytest

import helpers

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu"))


@pytest.fixture(scope="module")
def emu():
    import build_emu
    from ipddp_b200 import _lib
    return _lib.Lib(build_emu.build())


@pytest.mark.parametrize("wl,B,N,maxit", [("double_integrator", 2, 101, 1000), ("cartpole", 2, 9, 25),
                                          ("acrobot", 2, 9, 30), ("concar", 3, 11, 60), ("concar_quad", 2, 11, 60),
                                          ("pushing", 2, 9, 30)])
def test_emulated_full_solve(emu, oracle_mod, wl, B, N, maxit):
    helpers.full_solve_parity(emu, oracle_mod, wl, B, N, maxit=maxit, n_trace=B)


@pytest.mark.parametrize("wl,B,N,maxit", [("cartpole", 2, 9, 25), ("concar", 3, 11, 60), ("pushing", 2, 9, 30)])
def test_emulated_bulk_kernels(emu, oracle_mod, wl, B, N, maxit):
    """The small batches above all take the speculative kernels (k_forward_spec, k_backward_spec: few active
    instances); this forces the bulk kernels (k_forward, k_backward) on the same problems."""
    emu.L.ipddp_set_tuning(None, b"fw_spec_max", 0)
    emu.L.ipddp_set_tuning(None, b"bw_spec_max", 0)
    try:
        helpers.full_solve_parity(emu, oracle_mod, wl, B, N, maxit=maxit, n_trace=B)
    finally:
        emu.L.ipddp_set_tuning(None, b"fw_spec_max", -1)
        emu.L.ipddp_set_tuning(None, b"bw_spec_max", -1)


@pytest.mark.parametrize("order", ["rev", "rand"])
def test_emulated_lane_order_independence(emu, oracle_mod, order, monkeypatch):
    """Race check of the warp-synchronous shared-memory code.  The emulator runs each lane from one barrier to the next
    before the following lane starts, so a value handed from lane to lane through shared memory without a barrier in
    between is seen "new" by later lanes and "old" by earlier ones: a missing __syncwarp changes the results for some
    lane order (checked by hand: dropping the barrier after the row-list compaction in ldlt_step_fast changes the
    digests).  Default order = ascending (every other test); here descending and shuffled, bulk and tail kernels, against
    the oracle.  (compute-sanitizer's racecheck is not available on the GPU pool.)"""
    monkeypatch.setenv("IPDDP_EMU_ORDER", order)
    for spec in (148, 0):
        emu.L.ipddp_set_tuning(None, b"fw_spec_max", spec)
        emu.L.ipddp_set_tuning(None, b"bw_spec_max", 592 if spec else 0)
        try:
            helpers.full_solve_parity(emu, oracle_mod, "cartpole", 2, 9, maxit=25, n_trace=2)
            helpers.full_solve_parity(emu, oracle_mod, "pushing", 2, 9, maxit=20, vary_horizon=True, n_trace=2)
        finally:
            emu.L.ipddp_set_tuning(None, b"fw_spec_max", -1)
            emu.L.ipddp_set_tuning(None, b"bw_spec_max", -1)
    helpers.phase_parity(emu, oracle_mod, "concar", B=2, N=7, rounds=2)


def test_emulated_memcheck_asan():
    """memcheck substitute (compute-sanitizer is closed on the GPU pool): the kernel sources compiled with
    -fsanitize=address against the emulator, where "device" memory is heap memory with redzones and the shared memory
    beyond the launch's request is poisoned, run tiny solves through every kernel (tools/sanitize.py) in a subprocess.
    No report, and the digests equal the ones the B200 produced (profiles/r1_sanity/gpu_digests.jsonl).  Negative
    control done by hand: reading v.horizon[v.B] in k_init is reported with file and line."""
    import subprocess
    here = os.path.dirname(os.path.abspath(__file__))
    root = os.path.dirname(here)
    asan = subprocess.run(["gcc", "-print-file-name=libasan.so"], capture_output=True, text=True).stdout.strip()
    if not os.path.isabs(asan) or not os.path.exists(asan):
        pytest.skip("libasan not available")
    env = dict(os.environ, IPDDP_EMU_DEFS="-fsanitize=address -fno-omit-frame-pointer -g",
               IPDDP_EMU_LDFLAGS="-fsanitize=address", IPDDP_EMU_LIB="libipddp_emu_asan.so")
    subprocess.run([sys.executable, os.path.join(here, "emu", "build_emu.py")], env=env, check=True, capture_output=True)
    env = dict(os.environ, LD_PRELOAD=asan, ASAN_OPTIONS="detect_leaks=0:detect_stack_use_after_return=0",
               IPDDP_LIB=os.path.join(here, "emu", "libipddp_emu_asan.so"))
    env.pop("IPDDP_EMU_ORDER", None)
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "sanitize.py")], env=env, capture_output=True,
                       text=True, timeout=900)
    assert r.returncode == 0 and "AddressSanitizer" not in r.stderr, r.stderr[-3000:]
    want = open(os.path.join(root, "profiles", "r1_sanity", "gpu_digests.jsonl")).read().strip().splitlines()
    assert r.stdout.strip().splitlines() == want
    # round-2 paths: queue mode (k_admit / k_retire), warm start, varying horizons, heavy forward split, a stage chain
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "sanitize_small.py"), "3"], env=env,
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "AddressSanitizer" not in r.stderr and "sanitize_small done" in r.stdout, r.stderr[-3000:]


def test_emulated_forward_heavy_split(oracle_mod):
    """Bulk rounds hand their heaviest forward bucket to k_forward_spec while k_forward takes the rest.  An emulator
    build with the bucket threshold lowered to two trial steps (-DIPDDP_FWD_HEAVY_L=2) and a speculative capacity
    of 1-2 instances makes small batches take exactly that path; results must still equal the oracle bit for bit."""
    import subprocess
    from ipddp_b200 import _lib
    here = os.path.dirname(os.path.abspath(__file__))
    env = dict(os.environ, IPDDP_EMU_DEFS="-DIPDDP_FWD_HEAVY_L=2", IPDDP_EMU_LIB="libipddp_emu_heavy2.so")
    subprocess.run([sys.executable, os.path.join(here, "emu", "build_emu.py")], env=env, check=True, capture_output=True)
    lib = _lib.Lib(os.path.join(here, "emu", "libipddp_emu_heavy2.so"))
    for cap, wl, B, N, maxit in ((1, "cartpole", 3, 9, 25), (2, "pushing", 4, 9, 30), (1, "concar", 3, 11, 60)):
        lib.L.ipddp_set_tuning(None, b"fw_spec_max", cap)
        lib.L.ipddp_set_tuning(None, b"bw_spec_max", 0)
        helpers.full_solve_parity(lib, oracle_mod, wl, B, N, maxit=maxit, n_trace=B)


def test_emulated_tma_staged_rollout(oracle_mod):
    """-DIPDDP_FW_TMA=1: the rollout reads its per-knot gains / nominal records from a shared-memory ring filled by TMA
    bulk copies.  Under the emulator the copies are synchronous, which still checks the ring logic (stage indices across
    trial steps and early exits, record padding, drain) against the oracle, bulk and speculative kernels."""
    import subprocess
    from ipddp_b200 import _lib
    here = os.path.dirname(os.path.abspath(__file__))
    env = dict(os.environ, IPDDP_EMU_DEFS="-DIPDDP_FW_TMA=1", IPDDP_EMU_LIB="libipddp_emu_tma.so")
    subprocess.run([sys.executable, os.path.join(here, "emu", "build_emu.py")], env=env, check=True, capture_output=True)
    lib = _lib.Lib(os.path.join(here, "emu", "libipddp_emu_tma.so"))
    for spec in (148, 0):
        lib.L.ipddp_set_tuning(None, b"fw_spec_max", spec)
        helpers.full_solve_parity(lib, oracle_mod, "cartpole", 2, 9, maxit=25, n_trace=2)
        helpers.full_solve_parity(lib, oracle_mod, "pushing", 3, 9, maxit=30, vary_horizon=True, n_trace=2)
        helpers.full_solve_parity(lib, oracle_mod, "double_integrator", 2, 2)
    helpers.phase_parity(lib, oracle_mod, "concar", B=2, N=7, rounds=2)


@pytest.mark.parametrize("spec", [True, False])
def test_emulated_stage_chain(emu, oracle_mod, spec):
    """State and control sizes that change along the horizon (stage chain `ragged`: 2 -> 3 states, 3 -> 2 controls, a
    constraint that disappears), bulk and speculative kernels, against the oracle's per-stage models."""
    emu.L.ipddp_set_tuning(None, b"fw_spec_max", 148 if spec else 0)
    emu.L.ipddp_set_tuning(None, b"bw_spec_max", 592 if spec else 0)
    try:
        helpers.chain_parity(emu, oracle_mod, "ragged", 3, 13, maxit=60)
    finally:
        emu.L.ipddp_set_tuning(None, b"fw_spec_max", -1)
        emu.L.ipddp_set_tuning(None, b"bw_spec_max", -1)


def test_emulated_stage_chain_queue(emu, oracle_mod):
    helpers.chain_parity(emu, oracle_mod, "ragged", 5, 9, maxit=40, queue_slots=2)


def test_emulated_stage_chain_rejects_mismatched_stages(emu):
    from ipddp_b200.batch import BatchSolver
    s = BatchSolver("ragged", 2, 11, lib=emu)
    with pytest.raises(RuntimeError, match="maps to"):
        s.set_stage_types([0] * 10)            # type 0 keeps 2 states, the terminal cost wants 3
    with pytest.raises(RuntimeError, match="ipddp_set_stage_types not called"):
        s.set_inputs(np.zeros((2, 3)), np.zeros((2, 30)))
    s.close()


def test_emulated_varying_horizon(emu, oracle_mod):
    helpers.full_solve_parity(emu, oracle_mod, "concar", 4, 13, maxit=80, vary_horizon=True, first=100, n_trace=4)


@pytest.mark.parametrize("wl", ["double_integrator", "cartpole", "concar", "pushing"])
def test_emulated_phases(emu, oracle_mod, wl):
    helpers.phase_parity(emu, oracle_mod, wl, B=2, N=7, rounds=2)


def test_emulated_ldlt(emu, oracle_mod):
    helpers.ldlt_parity(emu, oracle_mod, np.random.default_rng(3), nmat=24, nmax=64)     # the C ABI accepts nu + nc <= 64


def test_emulated_exact_division(emu):
    helpers.division_parity(emu, np.random.default_rng(6), n=100000)


def test_emulated_detmath(emu, oracle_mod):
    helpers.detmath_parity(emu, oracle_mod, np.random.default_rng(4), n=5000)


def test_emulated_solve_many(emu, oracle_mod):
    """ipddp_solve_many (several problems in flight, event-polled) gives the same per-instance answers as
    sequential solves, and its aggregated counters add up."""
    from ipddp_b200 import instances
    from ipddp_b200.batch import BatchSolver, solve_many
    opt = emu.default_options(optimality_tolerance=1e-7)
    solvers, batches = [], []
    for q, (B, N) in enumerate([(2, 21), (3, 11), (1, 31)]):
        b = instances.make_batch("double_integrator" if q != 1 else "concar", B, N, first=10 * q)
        s = BatchSolver(b.workload, B, N, options=opt, lib=emu)
        s.set_batch(b)
        solvers.append(s); batches.append(b)
    ms, st = solve_many(solvers, total_solves=5)
    got = [s.results() for s in solvers]
    cnt = [s.counters() for s in solvers]
    conv = 0
    for s, r in zip(solvers, got):
        r2 = s.solve()      # sequential re-solve of the same inputs
        assert np.array_equal(r.k, r2.k) and np.array_equal(r.status, r2.status)
        helpers.assert_same_bits(r.objective, r2.objective, "objective")
    # 5 solves: handles 0,1,2 once, then the first two that finished once more
    assert st.n_converged >= sum(int((r.status == 0).sum()) for r in got)
    assert st.sum_kkt >= sum(int(c["n_kkt"].sum()) for c in cnt)
    assert st.launches > 0 and st.iterations > 0
    for s in solvers:
        s.close()


@pytest.mark.parametrize("wl,Q,B,N,maxit,vary", [("concar", 7, 3, 11, 60, True), ("double_integrator", 5, 8, 21, 1000, False),
                                                  ("cartpole", 4, 2, 9, 12, False), ("concar", 3, 2, 11, 0, False)])
def test_emulated_queue(emu, oracle_mod, wl, Q, B, N, maxit, vary):
    """ipddp_solve_queue: Q queued instances streamed through B slots (more than, fewer than the slots; max_iterations
    = 0 terminates inside the admission kernel) == the oracle instance by instance, bit for bit."""
    helpers.queue_parity(emu, oracle_mod, wl, Q, B, N, maxit=maxit, vary_horizon=vary)


@pytest.mark.parametrize("case,wl,B,N", [("status1", "cartpole", 2, 9), ("status7", "concar", 2, 11), ("status9", "concar", 2, 11)])
def test_emulated_forced_statuses(emu, oracle_mod, case, wl, B, N):
    helpers.forced_status_parity(emu, oracle_mod, case, (wl, B), N)


def test_emulated_queue_bad_horizon(emu):
    from ipddp_b200 import instances
    from ipddp_b200.batch import BatchSolver
    b = instances.make_batch("double_integrator", 3, 11)
    s = BatchSolver("double_integrator", 2, 11, lib=emu)
    hz = b.horizons.copy(); hz[2] = 12
    with pytest.raises(RuntimeError, match="horizon out of range"):
        s.solve_queue(b.x1, b.ubar, None, b.lower, b.upper, hz)
    s.close()


def test_emulated_one_sided_bounds(emu, oracle_mod):
    """Every branch of the control projection (reference src/solver.jl:70-95): upper-only, lower-only, two-sided with the
    guess outside, unbounded -- per instance, speculative and bulk kernels."""
    for spec in (-1, 0):
        emu.L.ipddp_set_tuning(None, b"fw_spec_max", spec)
        emu.L.ipddp_set_tuning(None, b"bw_spec_max", spec)
        try:
            helpers.full_solve_parity(emu, oracle_mod, "concar", 4, 11, maxit=40, n_trace=4, mutate=helpers.one_sided_bounds)
        finally:
            emu.L.ipddp_set_tuning(None, b"fw_spec_max", -1)
            emu.L.ipddp_set_tuning(None, b"bw_spec_max", -1)


def test_emulated_options_reach_the_kernels(emu, oracle_mod):
    """The 31 Options fields travel through the C ABI as one struct (reference src/options.jl:1-38).  Three option sets that
    together move every barrier / regularisation / line-search parameter away from its default, plus the projection
    parameters on instances whose guesses lie outside their bounds: each set changes the iterates, and the kernels follow
    the oracle bit for bit -- a kernel that kept a default where it should read the option would not.  (One option at a
    time: the GPU suite, test_gpu_parity.py::test_every_option_reaches_the_kernels.)"""
    changed = helpers.options_parity(emu, oracle_mod, helpers.OPTION_GROUPS, wl="acrobot", B=2, N=11, maxit=40, first=0)
    assert all(changed), changed
    changed = helpers.options_parity(emu, oracle_mod, [dict(kappa_1=1e-3, kappa_2=0.2), dict(kappa_2=1e-4)], wl="concar", B=4,
                                     N=11, maxit=60, first=0, mutate=helpers.one_sided_bounds)
    assert all(changed), changed


def test_emulated_duals_after_converged_and_max_iteration_exits(emu, oracle_mod):
    helpers.duals_parity(emu, oracle_mod)
