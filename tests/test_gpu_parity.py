"""GPU parity tests (run with -m gpu on a B200): the CUDA path, called through the C ABI of
libipddp_b200.so, against the CPU oracle on the same seeded inputs.  Bar: bit equality of status, iteration
counts, accepted-step trace, objective, infeasibilities and trajectories (DESIGN.md "Parity").  Golden
known-answer rows of the reference are re-checked on the GPU output directly."""
import numpy as np
import pytest

import helpers

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu():
    import torch
    assert torch.cuda.is_available(), "no CUDA device"
    from ipddp_b200 import _lib
    return _lib.load()


def test_detmath_bit_exact(gpu, oracle_mod):
    helpers.detmath_parity(gpu, oracle_mod, np.random.default_rng(4), n=200000)


def test_exact_division_bit_exact(gpu):
    helpers.division_parity(gpu, np.random.default_rng(5))


def test_ldlt_bit_exact(gpu, oracle_mod):
    helpers.ldlt_parity(gpu, oracle_mod, np.random.default_rng(3), nmat=400, nmax=35)


@pytest.mark.parametrize("wl", ["double_integrator", "cartpole", "acrobot", "concar", "concar_quad", "pushing"])
def test_phase_parity(gpu, oracle_mod, wl):
    helpers.phase_parity(gpu, oracle_mod, wl, B=4, N=21, rounds=3)


@pytest.mark.parametrize("wl,B", [("double_integrator", 1), ("cartpole", 32), ("acrobot", 16), ("concar", 16),
                                  ("concar_quad", 16), ("pushing", 8)])
def test_full_solve_parity(gpu, oracle_mod, wl, B):
    helpers.full_solve_parity(gpu, oracle_mod, wl, B, 101, n_trace=3)


@pytest.mark.parametrize("wl,B", [("cartpole", 24), ("acrobot", 12), ("pushing", 8)])
def test_bulk_kernels_parity(gpu, oracle_mod, wl, B):
    """Batches this small normally take the speculative tail kernels (k_forward_spec: 8 step sizes at once,
    k_backward_spec: 4 regularisation values at once); forcing the thresholds to 0 runs the bulk kernels
    (k_forward, k_backward) on the same instances.  Both shapes must reproduce the oracle bit for bit."""
    gpu.L.ipddp_set_tuning(None, b"fw_spec_max", 0)
    gpu.L.ipddp_set_tuning(None, b"bw_spec_max", 0)
    try:
        helpers.full_solve_parity(gpu, oracle_mod, wl, B, 101, n_trace=2)
    finally:
        gpu.L.ipddp_set_tuning(None, b"fw_spec_max", -1)
        gpu.L.ipddp_set_tuning(None, b"bw_spec_max", -1)


def test_tail_kernels_equal_bulk_kernels_mid_size(gpu):
    """Size-independent property at a few hundred instances: the speculative tail kernels and the bulk kernels give
    identical per-instance answers and identical work counters (sweeps, KKT steps, rollouts)."""
    from ipddp_b200 import instances
    from ipddp_b200.batch import BatchSolver
    B = 320
    opt = gpu.default_options(optimality_tolerance=1e-7)
    b = instances.make_batch("cartpole", B, 101, first=500)
    out = []
    for fw, bw in ((0, 0), (B, B)):
        s = BatchSolver("cartpole", B, 101, options=opt, lib=gpu)
        s.set_tuning("fw_spec_max", fw)
        s.set_tuning("bw_spec_max", bw)
        s.set_batch(b)
        r = s.solve()
        x, u = s.trajectory()
        c = s.counters()
        out.append((r, x, u, c))
        s.close()
    (r0, x0, u0, c0), (r1, x1, u1, c1) = out
    assert np.array_equal(r0.k, r1.k) and np.array_equal(r0.status, r1.status) and np.array_equal(r0.l, r1.l)
    helpers.assert_same_bits(r0.objective, r1.objective, "objective")
    helpers.assert_same_bits(r0.reg_last, r1.reg_last, "reg_last")
    helpers.assert_same_bits(r0.step_size, r1.step_size, "step size")
    helpers.assert_same_bits(x0, x1, "states")
    helpers.assert_same_bits(u0, u1, "controls")
    for key in ("n_backward", "n_sweeps", "n_kkt", "n_rollouts"):
        assert np.array_equal(c0[key], c1[key]), key


def test_varying_horizon_parity(gpu, oracle_mod):
    """config 5: per-instance horizons (offset tables / horizon vector), synthetic instances."""
    helpers.full_solve_parity(gpu, oracle_mod, "pushing", 8, 101, vary_horizon=True, first=100, n_trace=2)
    helpers.full_solve_parity(gpu, oracle_mod, "concar", 12, 61, vary_horizon=True, first=200, n_trace=2)


def test_max_iterations_and_short_horizon(gpu, oracle_mod):
    helpers.full_solve_parity(gpu, oracle_mod, "cartpole", 4, 101, maxit=7)          # status 8
    helpers.full_solve_parity(gpu, oracle_mod, "double_integrator", 1, 2)           # single running stage
    helpers.full_solve_parity(gpu, oracle_mod, "acrobot", 3, 3, maxit=50)


@pytest.mark.parametrize("wl,B,N,vary", [("acrobot", 2048, 201, False), ("concar", 1024, 101, False),
                                         ("concar_quad", 1024, 101, False), ("pushing", 512, 141, True)])
def test_baseline_configs_sampled_vs_oracle(gpu, oracle_mod, wl, B, N, vary):
    """BASELINE configs 3-5 at a scale where the bulk kernels run (thousands of instances, N = 201 for acrobot,
    per-instance horizons for pushing): a random sample of instances is re-solved on the oracle and must agree bit
    for bit, including the work counters."""
    helpers.sampled_parity(gpu, oracle_mod, wl, B, N, sample=256 if wl != "pushing" else 96, vary_horizon=vary, first=100, seed=7)


def test_golden_table_on_gpu(gpu):
    """The reference's own known answers (experiments/ipddp2/results/cartpole_friction.txt), straight from the GPU."""
    from ipddp_b200 import instances
    from ipddp_b200.batch import BatchSolver
    for wl, frac in (("cartpole", 0.90), ("concar_quad", 0.95), ("double_integrator", 1.0)):
        g = instances.load_golden_results(wl)
        n = len(g["seed"])
        b = instances.make_batch(wl, n, 101)
        s = BatchSolver(wl, n, 101, options=gpu.default_options(optimality_tolerance=1e-7), lib=gpu)
        s.set_batch(b)
        r = s.solve()
        ok = (r.k == g["iterations"]) & (np.abs(r.objective - g["objective"]) <= 1e-8 * np.maximum(1.0, np.abs(g["objective"])))
        assert ok.sum() >= frac * n, f"{wl}: {ok.sum()}/{n}"
        assert np.array_equal(r.status[ok] == 0, g["converged"][ok])
        s.close()


def test_batch_consistency_and_permutation(gpu):
    """instance i inside a batch == instance i alone; results do not depend on the position in the batch
    (size-independent property, run at a few thousand instances)."""
    from ipddp_b200 import instances
    from ipddp_b200.batch import BatchSolver
    B = 2048
    opt = gpu.default_options(optimality_tolerance=1e-7)
    b = instances.make_batch("cartpole", B, 101)
    s = BatchSolver("cartpole", B, 101, options=opt, lib=gpu)
    s.set_batch(b)
    r = s.solve()
    x, u = s.trajectory()
    s.close()
    assert (r.status == 0).mean() > 0.97
    perm = np.random.default_rng(0).permutation(B)
    s2 = BatchSolver("cartpole", B, 101, options=opt, lib=gpu)
    s2.set_inputs(b.x1[perm], b.ubar[perm], b.p[perm], b.lower[perm], b.upper[perm], b.horizons[perm])
    r2 = s2.solve()
    x2, u2 = s2.trajectory()
    s2.close()
    assert np.array_equal(r2.k, r.k[perm]) and np.array_equal(r2.status, r.status[perm])
    helpers.assert_same_bits(r2.objective, r.objective[perm], "objective under permutation")
    helpers.assert_same_bits(x2, x[perm], "states under permutation")
    for i in (0, 777, 2047):
        s1 = BatchSolver("cartpole", 1, 101, options=opt, lib=gpu)
        s1.set_inputs(b.x1[i:i + 1], b.ubar[i:i + 1], b.p[i:i + 1], b.lower[i:i + 1], b.upper[i:i + 1])
        r1 = s1.solve()
        x1_, u1_ = s1.trajectory()
        s1.close()
        assert int(r1.k[0]) == int(r.k[i]) and int(r1.status[0]) == int(r.status[i])
        helpers.assert_same_bits(u1_[0], u[i], f"controls of instance {i} alone vs in batch")


def test_full_size_batch_properties(gpu, oracle_mod):
    """BASELINE config 2 at full size (16384 cartpole instances): every instance terminates with a valid
    status, >= 99 % converge with primal infeasibility below tolerance, the first 100 reproduce the small-batch
    answers bit-for-bit, and a random sample of 256 instances matches the oracle exactly."""
    from ipddp_b200 import instances
    from ipddp_b200.batch import BatchSolver
    B = 16384
    opt = gpu.default_options(optimality_tolerance=1e-7)
    b = instances.make_batch("cartpole", B, 101)
    s = BatchSolver("cartpole", B, 101, options=opt, lib=gpu)
    s.set_batch(b)
    r = s.solve()
    st = s.stats()
    s.close()
    assert np.isin(r.status, [0, 1, 7, 8]).all()
    assert (r.status == 0).mean() >= 0.99
    conv = r.status == 0
    assert (r.primal_inf[conv] < 1e-7).all() and np.isfinite(r.objective[conv]).all()
    assert st.sum_kkt > 0 and st.n_converged == conv.sum()
    s100 = BatchSolver("cartpole", 100, 101, options=opt, lib=gpu)
    s100.set_batch(b.slice(0, 100))
    r100 = s100.solve()
    s100.close()
    assert np.array_equal(r100.k, r.k[:100])
    helpers.assert_same_bits(r100.objective, r.objective[:100], "first 100 of the full batch")
    idx = np.random.default_rng(1).choice(B, 256, replace=False)
    oopt = oracle_mod.default_options(optimality_tolerance=1e-7)
    res, _, _ = oracle_mod.solve_batch("cartpole", 101, b.p[idx], b.lower[idx], b.upper[idx], b.x1[idx], b.ubar[idx], options=oopt)
    for q, i in enumerate(idx):
        assert (int(r.status[i]), int(r.k[i])) == (res[q].status, res[q].k)
        helpers.assert_same_bits(r.objective[i], res[q].objective, f"instance {i} objective vs oracle")


@pytest.mark.parametrize("wl,Q,B,N,vary", [("cartpole", 3000, 1024, 101, False), ("pushing", 700, 256, 141, True),
                                             ("acrobot", 300, 512, 201, False)])
def test_queue_streaming_parity(gpu, oracle_mod, wl, Q, B, N, vary):
    """ipddp_solve_queue: Q instances streamed through B slots (admission / retirement between rounds) give, per queue
    index, exactly what one resident batch of the same instances gives (every scalar, counter and trajectory), and a
    sample agrees with the oracle bit for bit."""
    from ipddp_b200 import instances
    from ipddp_b200.batch import BatchSolver
    b = instances.make_batch(wl, Q, N, vary_horizon=vary, first=50)
    opt = gpu.default_options(optimality_tolerance=1e-7)
    s = BatchSolver(wl, B, N, options=opt, lib=gpu)
    r, cnt, x, u = s.solve_queue(b.x1, b.ubar, b.p if s.np > 0 else None, b.lower, b.upper, b.horizons)
    stq = s.stats()
    s.close()
    s2 = BatchSolver(wl, Q, N, options=opt, lib=gpu)
    s2.set_batch(b)
    r2 = s2.solve()
    x2, u2 = s2.trajectory()
    c2 = s2.counters()
    s2.close()
    for name in ("status", "k", "j", "l"):
        assert np.array_equal(getattr(r, name), getattr(r2, name)), name
    for name in ("objective", "primal_inf", "dual_inf", "cs_inf", "mu", "reg_last", "step_size"):
        helpers.assert_same_bits(getattr(r, name), getattr(r2, name), name)
    for key in ("n_backward", "n_sweeps", "n_kkt", "n_rollouts"):
        assert np.array_equal(cnt[key], c2[key]), key
    helpers.assert_same_bits(x, x2, "states")
    helpers.assert_same_bits(u, u2, "controls")
    assert stq.n_converged == int((r2.status == 0).sum()) and stq.sum_kkt == int(c2["n_kkt"].sum())
    if Q > B:   # the rounds stay (nearly) full until the queue runs dry
        assert stq.n_active_rounds / stq.iterations > 0.3 * B
    idx = np.sort(np.random.default_rng(5).choice(Q, 24, replace=False))
    oopt = oracle_mod.default_options(optimality_tolerance=1e-7)
    res, xo, uo = oracle_mod.solve_batch(wl, N, b.p[idx], b.lower[idx], b.upper[idx], b.x1[idx], b.ubar[idx], options=oopt,
                                         horizons=b.horizons[idx], want_traj=True)
    for q, i in enumerate(idx):
        assert (int(r.status[i]), int(r.k[i]), int(cnt["n_kkt"][i])) == (res[q].status, res[q].k, res[q].n_kkt)
        helpers.assert_same_bits(r.objective[i], res[q].objective, f"queue index {i} objective vs oracle")
        helpers.assert_same_bits(x[i], xo[q], f"queue index {i} states vs oracle")


def test_queue_device_buffers(gpu):
    """inputs and outputs of ipddp_solve_queue resident in device memory (torch tensors) == host buffers."""
    import torch
    from ipddp_b200 import instances
    from ipddp_b200.batch import BatchSolver, make_queue
    Q, B, N = 600, 256, 61
    b = instances.make_batch("concar_quad", Q, N)
    opt = gpu.default_options(optimality_tolerance=1e-7)
    s = BatchSolver("concar_quad", B, N, options=opt, lib=gpu)
    r, cnt, x, u = s.solve_queue(b.x1, b.ubar, b.p, b.lower, b.upper, b.horizons)
    dev = torch.device("cuda", 0)
    t = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in dict(x1=b.x1, ubar=b.ubar, p=b.p, lower=b.lower, upper=b.upper).items()}
    hz = torch.from_numpy(b.horizons.astype(np.int32)).to(dev)
    oi = [torch.zeros(Q, dtype=torch.int32, device=dev) for _ in range(4)]
    od = [torch.zeros(Q, dtype=torch.float64, device=dev) for _ in range(7)]
    oc = [torch.zeros(Q, dtype=torch.int32, device=dev) for _ in range(4)]
    xd = torch.zeros((Q, N, s.nx), dtype=torch.float64, device=dev)
    ud = torch.zeros((Q, N - 1, s.nu), dtype=torch.float64, device=dev)
    q = make_queue(Q, t["x1"].data_ptr(), t["ubar"].data_ptr(), t["p"].data_ptr(), t["lower"].data_ptr(), t["upper"].data_ptr(),
                   hz.data_ptr(), [a.data_ptr() for a in oi + od + oc], xd.data_ptr(), ud.data_ptr(),
                   inputs_on_device=True, outputs_on_device=True)
    import ctypes as C
    gpu.check(gpu.L.ipddp_solve_queue(s.h, C.byref(q)), "ipddp_solve_queue")
    torch.cuda.synchronize()
    s.close()
    assert np.array_equal(oi[0].cpu().numpy(), r.status) and np.array_equal(oi[1].cpu().numpy(), r.k)
    helpers.assert_same_bits(od[0].cpu().numpy(), r.objective, "objective")
    assert np.array_equal(oc[2].cpu().numpy(), cnt["n_kkt"])
    helpers.assert_same_bits(xd.cpu().numpy(), x, "states")
    helpers.assert_same_bits(ud.cpu().numpy(), u, "controls")


@pytest.mark.parametrize("case", ["status1", "status7", "status9"])
def test_forced_failure_statuses(gpu, oracle_mod, case):
    helpers.forced_status_parity(gpu, oracle_mod, case, {"status1": ("cartpole", 48), "status7": ("concar", 16),
                                                         "status9": ("concar", 16)}[case], 101)


def test_stage_chain_parity(gpu, oracle_mod):
    """SURVEY 8 a13 / reference README.md:18: state and control sizes that change along the horizon.  The chain model
    `ragged` (2 -> 3 states, 3 -> 2 controls, a constraint that disappears mid-way; three stage types) on the device
    against the oracle's per-stage models, bit for bit: resident batch (speculative kernels), bulk kernels, and streamed
    through fewer slots than instances."""
    helpers.chain_parity(gpu, oracle_mod, "ragged", 48, 101)
    gpu.L.ipddp_set_tuning(None, b"fw_spec_max", 0)
    gpu.L.ipddp_set_tuning(None, b"bw_spec_max", 0)
    try:
        helpers.chain_parity(gpu, oracle_mod, "ragged", 24, 61)
    finally:
        gpu.L.ipddp_set_tuning(None, b"fw_spec_max", -1)
        gpu.L.ipddp_set_tuning(None, b"bw_spec_max", -1)
    helpers.chain_parity(gpu, oracle_mod, "ragged", 40, 41, queue_slots=16)


def test_stage_chain_rejects_mismatched_stages(gpu):
    from ipddp_b200.batch import BatchSolver
    s = BatchSolver("ragged", 2, 11, lib=gpu)
    with pytest.raises(RuntimeError, match="maps to"):
        s.set_stage_types([0] * 10)            # type 0 keeps 2 states, the terminal cost wants 3
    with pytest.raises(RuntimeError, match="ipddp_set_stage_types not called"):
        s.set_inputs(np.zeros((2, 3)), np.zeros((2, 30)))
    s.close()


def test_error_convention_gpu(gpu):
    """API misuse -> return code + ipddp_last_error, algorithmic outcome -> per-instance status."""
    helpers.api_error_convention(gpu)


def test_one_sided_bounds(gpu, oracle_mod):
    """Upper-only, lower-only, two-sided (guess outside) and unbounded controls per instance: every branch of the control
    projection of reference src/solver.jl:70-95, incl. the upper-only one no experiment uses."""
    helpers.full_solve_parity(gpu, oracle_mod, "concar", 16, 41, maxit=300, n_trace=4, mutate=helpers.one_sided_bounds)


def test_every_option_reaches_the_kernels(gpu, oracle_mod):
    """One Options field at a time (reference src/options.jl:1-38) moved away from its default: each changes the iterates of
    small pushing solves, and the GPU follows the oracle bit for bit every time."""
    variants = [{k: v} for k, v in helpers.OPTION_VARIANTS.items()] + helpers.OPTION_GROUPS
    changed = helpers.options_parity(gpu, oracle_mod, variants)
    assert all(changed), [v for v, c in zip(variants, changed) if not c]
    changed = helpers.options_parity(gpu, oracle_mod, [dict(kappa_1=1e-3, kappa_2=0.2), dict(kappa_2=1e-4)], wl="concar", B=4,
                                     N=11, maxit=60, first=0, mutate=helpers.one_sided_bounds)
    assert all(changed), changed


def test_ldlt_bit_exact_up_to_the_largest_kkt(gpu, oracle_mod):
    """ipddp_problem_create accepts models with nu + nc <= 64: the warp LDL^T at n = 48 and 64 (two lane slots per column
    beyond 32) against the oracle's dsytf2_rook / dsytrs_rook, as test_ldlt_bit_exact does up to the largest built-in model."""
    helpers.ldlt_parity(gpu, oracle_mod, np.random.default_rng(13), nmat=200, nmax=64)


def test_duals_after_converged_and_max_iteration_exits(gpu, oracle_mod):
    helpers.duals_parity(gpu, oracle_mod)
