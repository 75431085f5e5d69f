"""Parity checks shared by the GPU tests (-m gpu, through libipddp_b200.so on a B200) and the CPU-side
emulator tests (same C ABI, kernels compiled against tests/emu/cpu_simt.h).  The checker is always the
oracle (oracle/); the bar is BIT equality for every compared quantity (see DESIGN.md "Parity")."""
import ctypes as C

import numpy as np

import ipddp_b200  # noqa: F401
from ipddp_b200 import instances
from ipddp_b200.batch import BatchSolver
from ipddp_b200.codegen import generate, workloads


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.int64)


def assert_same_bits(a, b, what):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    # +0.0 and -0.0 are both accepted as zero (structural zeros are skipped on the device)
    same = (bits(a) == bits(b)) | ((a == 0.0) & (b == 0.0)) | (np.isnan(a) & np.isnan(b))
    if not same.all():
        idx = np.argwhere(~same)[0]
        raise AssertionError(f"{what}: {int((~same).sum())} of {a.size} entries differ, first at {tuple(idx)}: "
                             f"{a[tuple(idx)]!r} vs {b[tuple(idx)]!r}")


def full_solve_parity(lib, oracle, wl, B, N=101, maxit=1000, tol=1e-7, vary_horizon=False, first=0, n_trace=4,
                      check_traj=True, mutate=None):
    """Whole-solve parity: status, k, j, l, objective, errors, step sequence (trace) and trajectories.  mutate(batch)
    may edit the instances (bounds, initial guesses) before both sides solve them."""
    b = instances.make_batch(wl, B, N, vary_horizon=vary_horizon, first=first)
    if mutate is not None:
        mutate(b)
    opt = lib.default_options(optimality_tolerance=tol, max_iterations=maxit)
    s = BatchSolver(wl, B, N, options=opt, trace_capacity=maxit, lib=lib)
    s.set_batch(b)
    r = s.solve()
    oopt = oracle.default_options(optimality_tolerance=tol, max_iterations=maxit)
    res, xo, uo = oracle.solve_batch(wl, N, b.p, b.lower, b.upper, b.x1, b.ubar, options=oopt,
                                     horizons=b.horizons, want_traj=True)
    for i in range(B):
        o = res[i]
        got = (int(r.status[i]), int(r.k[i]), int(r.j[i]))
        assert got == (o.status, o.k, o.j), f"{wl} inst {first+i}: (status,k,j) gpu {got} vs oracle {(o.status, o.k, o.j)}"
        for name in ("objective", "primal_inf", "dual_inf", "cs_inf", "mu", "reg_last", "step_size"):
            assert_same_bits(getattr(r, name)[i], getattr(o, name), f"{wl} inst {first+i} {name}")
        assert int(r.l[i]) == o.l
    cnt = s.counters()
    for i in range(B):
        assert (cnt["n_backward"][i], cnt["n_sweeps"][i], cnt["n_kkt"][i], cnt["n_rollouts"][i]) == \
            (res[i].n_backward, res[i].n_sweeps, res[i].n_kkt, res[i].n_rollouts), f"{wl} inst {i}: work counters"
    if check_traj:
        x, u = s.trajectory()
        assert_same_bits(x, xo, f"{wl} states")
        assert_same_bits(u, uo, f"{wl} controls")
    for i in range(min(n_trace, B)):
        o = oracle.OracleSolver(wl, int(b.horizons[i]), b.p[i], b.lower[i], b.upper[i], options=oopt)
        o.solve(b.x1[i], b.ubar[i][:(int(b.horizons[i]) - 1) * s.nu])
        assert_same_bits(s.trace(i), o.trace(), f"{wl} inst {first+i} accepted-step trace")
    st = s.stats()
    s.close()
    return r, st


def one_sided_bounds(b):
    """Edits a concar batch so that every branch of the control projection and of the slack / dual bookkeeping occurs
    (reference src/solver.jl:70-95, src/bounds.jl:12-26): upper-only (the branch that is broken upstream, restated by its
    evident intent on both sides), lower-only with a guess below the bound, two-sided with guesses outside on either
    side, and no bound at all."""
    inf = np.inf
    nu = b.lower.shape[1]
    ub = b.ubar.reshape(b.B, -1, nu)
    for i in range(b.B):
        k = i % 4
        if k == 0:      # control 0 upper-only, guess above the bound -> projected down
            b.lower[i, 0] = -inf; ub[i, :, 0] = b.upper[i, 0] + 0.5
        elif k == 1:    # control 1 lower-only, guess below the bound -> projected up; control 0 upper-only, guess inside
            b.upper[i, 1] = inf; ub[i, :, 1] = b.lower[i, 1] - 1.0
            b.lower[i, 0] = -inf
        elif k == 2:    # two-sided, guesses outside on both sides
            ub[i, :, 0] = b.upper[i, 0] + 1.0; ub[i, :, 1] = b.lower[i, 1] - 1.0
        else:           # no bounds on the first two controls; an upper bound on a slack-like control
            b.lower[i, 0] = -inf; b.upper[i, 0] = inf; b.lower[i, 1] = -inf; b.upper[i, 1] = inf
            b.upper[i, 2] = 0.75
    b.ubar = np.ascontiguousarray(ub.reshape(b.B, -1))


def queue_parity(lib, oracle, wl, Q, B, N=101, maxit=1000, tol=1e-7, vary_horizon=False, first=0):
    """ipddp_solve_queue (Q queued instances through B slots) vs the oracle: every SolverData scalar, the work counters
    and the trajectories, bit for bit, per queue index."""
    b = instances.make_batch(wl, Q, N, vary_horizon=vary_horizon, first=first)
    opt = lib.default_options(optimality_tolerance=tol, max_iterations=maxit)
    s = BatchSolver(wl, B, N, options=opt, lib=lib)
    r, cnt, x, u = s.solve_queue(b.x1, b.ubar, b.p if s.np > 0 else None, b.lower, b.upper, b.horizons)
    st = s.stats()
    s.close()
    oopt = oracle.default_options(optimality_tolerance=tol, max_iterations=maxit)
    res, xo, uo = oracle.solve_batch(wl, N, b.p, b.lower, b.upper, b.x1, b.ubar, options=oopt, horizons=b.horizons,
                                     want_traj=True)
    for i in range(Q):
        o = res[i]
        got = (int(r.status[i]), int(r.k[i]), int(r.j[i]), int(r.l[i]))
        assert got == (o.status, o.k, o.j, o.l), f"{wl} queue index {i}: (status,k,j,l) {got} vs oracle {(o.status, o.k, o.j, o.l)}"
        for name in ("objective", "primal_inf", "dual_inf", "cs_inf", "mu", "reg_last", "step_size"):
            assert_same_bits(getattr(r, name)[i], getattr(o, name), f"{wl} queue index {i} {name}")
        assert (cnt["n_backward"][i], cnt["n_sweeps"][i], cnt["n_kkt"][i], cnt["n_rollouts"][i]) == \
            (o.n_backward, o.n_sweeps, o.n_kkt, o.n_rollouts), f"{wl} queue index {i}: work counters"
    assert_same_bits(x, xo, f"{wl} states")
    assert_same_bits(u, uo, f"{wl} controls")
    assert st.n_converged == sum(1 for o in res if o.status == 0) and st.sum_kkt == sum(o.n_kkt for o in res)
    return r, st


def chain_inputs(chain, B, N, seed=0):
    """Padded batch inputs of a stage chain: (stage_types, x1 [B,nx], ubar [B,(N-1)*nu], lower/upper [B,nstage*nu]) plus
    the per-instance ragged pieces the oracle takes.  Instances differ in their initial state."""
    st = chain.stage_types(N)
    nu_max = max(md.nu for md in chain.stages)
    nx0 = chain.stages[st[0]].nx
    nx_max = max(md.nx for md in chain.stages)
    rng = np.random.default_rng(seed)
    x1 = np.zeros((B, nx_max))
    x1[:, :nx0] = 0.2 * rng.standard_normal((B, nx0))
    x1[0, :] = 0.0
    ubar = np.zeros((B, N - 1, nu_max))
    for t, k in enumerate(st):
        ubar[:, t, :chain.stages[k].nu] = np.asarray(chain.stages[k].u_init)
    lower = np.full((B, len(chain.stages), nu_max), -np.inf)
    upper = np.full((B, len(chain.stages), nu_max), np.inf)
    for k, md in enumerate(chain.stages):
        lower[:, k, :md.nu] = md.lower([])
        upper[:, k, :md.nu] = md.upper([])
    return st, x1, ubar.reshape(B, -1), lower.reshape(B, -1), upper.reshape(B, -1)


def chain_parity(lib, oracle, name, B, N, maxit=1000, tol=1e-7, queue_slots=0):
    """A horizon whose state / control sizes change from stage to stage (reference README.md:18, src/data/problem.jl:44-62):
    the device's chain model against the oracle's per-stage models -- status, iteration counts, objective and error bits,
    work counters, accepted-step trace and the (ragged) trajectories."""
    chain = workloads.get_chain(name)
    st, x1, ubar, lower, upper = chain_inputs(chain, B, N)
    opt = lib.default_options(optimality_tolerance=tol, max_iterations=maxit)
    s = BatchSolver(name, queue_slots or B, N, options=opt, trace_capacity=0 if queue_slots else maxit, lib=lib)
    assert s.nstage == len(chain.stages)
    s.set_stage_types(st)
    nxs, nus, ncs = s.stage_layout()
    assert list(nus[:-1]) == [chain.stages[k].nu for k in st] and nus[-1] == 0 and nxs[-1] == s.nxt
    if queue_slots:
        r, cnt, x, u = s.solve_queue(x1, ubar, None, lower, upper)
    else:
        s.set_inputs(x1, ubar, None, lower, upper)
        r = s.solve()
        x, u = s.trajectory()
        cnt = s.counters()
    nu_max = s.nu
    oopt = oracle.default_options(optimality_tolerance=tol, max_iterations=maxit)
    for i in range(B):
        o = oracle.OracleChainSolver([md.name for md in chain.stages], st, N, [], [md.lower([]) for md in chain.stages],
                                     [md.upper([]) for md in chain.stages], options=oopt)
        ub = ubar[i].reshape(N - 1, nu_max)
        res = o.solve(x1[i, :chain.stages[st[0]].nx], np.concatenate([ub[t, :chain.stages[k].nu] for t, k in enumerate(st)]))
        got = (int(r.status[i]), int(r.k[i]), int(r.j[i]), int(r.l[i]))
        assert got == (res.status, res.k, res.j, res.l), f"{name} inst {i}: (status,k,j,l) {got} vs oracle {(res.status, res.k, res.j, res.l)}"
        for nm in ("objective", "primal_inf", "dual_inf", "cs_inf", "mu", "reg_last", "step_size"):
            assert_same_bits(getattr(r, nm)[i], getattr(res, nm), f"{name} inst {i} {nm}")
        assert (cnt["n_backward"][i], cnt["n_sweeps"][i], cnt["n_kkt"][i], cnt["n_rollouts"][i]) == \
            (res.n_backward, res.n_sweeps, res.n_kkt, res.n_rollouts), f"{name} inst {i}: work counters"
        xo, uo = o.array("x"), o.array("u")
        px = pu = 0
        for t in range(N):
            assert_same_bits(x[i, t, :nxs[t]], xo[px:px + nxs[t]], f"{name} inst {i} state of knot {t}")
            assert not x[i, t, nxs[t]:].any()
            px += nxs[t]
            if t < N - 1:
                assert_same_bits(u[i, t, :nus[t]], uo[pu:pu + nus[t]], f"{name} inst {i} control of knot {t}")
                pu += nus[t]
        if not queue_slots and i < 2:
            assert_same_bits(s.trace(i), o.trace(), f"{name} inst {i} accepted-step trace")
    s.close()
    return r


def _tile_map(wl):
    md = workloads.get(wl)
    bd = generate.trace(md)
    ents, consts, nslot = generate._device_entries(bd["derivs"])
    entsN, constsN, nslotN = generate._device_entries(bd["derivsN"])
    rows = {m: r for (m, r, _) in bd["derivs"].outputs}
    rowsN = {m: r for (m, r, _) in bd["derivsN"].outputs}
    return ents, nslot, rows, entsN, nslotN, rowsN


def phase_parity(lib, oracle, wl, B=3, N=9, rounds=3):
    """Kernel-level parity after each phase of the first few iterations:
    initialise -> derivatives (compact tile vs the oracle's dense matrices) -> backward pass (gains, Qu,
    costate, reg) -> errors/check -> forward pass (trial trajectory, accepted step)."""
    b = instances.make_batch(wl, B, N)
    opt = lib.default_options(optimality_tolerance=1e-7)
    s = BatchSolver(wl, B, N, options=opt, lib=lib)
    s.set_batch(b)
    nx, nu, nc = s.nx, s.nu, s.nc
    K = nu + nc
    G = (K + 2 * nu) * (nx + 1)
    oopt = oracle.default_options(optimality_tolerance=1e-7)
    orc = [oracle.OracleSolver(wl, N, b.p[i], b.lower[i], b.upper[i], options=oopt) for i in range(B)]
    for i, o in enumerate(orc):
        o.initialize(b.x1[i], b.ubar[i])
    s.initialize()
    ents, nslot, rows, entsN, nslotN, rowsN = _tile_map(wl)

    def cmp_traj(prefix=""):
        for name, dim, nst in (("x", nx, N), ("u", nu, N - 1), ("c", nc, N - 1), ("il", nu, N - 1), ("iu", nu, N - 1),
                               ("phi", nc, N - 1), ("zl", nu, N - 1), ("zu", nu, N - 1)):
            g = s.array(prefix + name).reshape(B, nst, dim)
            for i, o in enumerate(orc):
                ref = o.array(prefix + name)
                assert_same_bits(g[i].reshape(-1), ref[:nst * dim], f"{wl} {prefix}{name} inst {i}")

    cmp_traj()
    for rnd in range(rounds):
        s.eval_derivatives()
        tile = s.array("tile").reshape(B, max(nslot, 0), N) if nslot else None
        tileN = s.array("tileN").reshape(B, nslotN) if nslotN else None
        for i, o in enumerate(orc):
            o.eval_derivatives()
            dense = {m: o.array(m) for m in rows}
            sizes = {m: dense[m].size // N if m not in ("fx", "fu") else dense[m].size // (N - 1) for m in rows}
            for en, slot in ents:
                if slot < 0:
                    continue
                per = {"fx": nx * nx, "fu": nx * nu, "lx": nx, "lu": nu, "lxx": nx * nx, "luu": nu * nu, "lux": nu * nx,
                       "cx": nc * nx, "cu": nc * nu, "vcxx": nx * nx, "vcux": nu * nx, "vcuu": nu * nu}[en.mat]
                for t in range(N - 1):
                    off = t * per
                    if en.mat in ("lx", "lxx", "vcxx"):
                        off = t * per      # these exist for all N stages with the same size
                    ref = dense[en.mat][off + en.i + en.j * rows[en.mat]]
                    assert_same_bits(tile[i, slot, t], ref, f"{wl} tile {en.mat}[{en.i},{en.j}] t={t} inst {i}")
            for en, slot in entsN:
                if slot < 0:
                    continue
                per = {"lx": nx, "lxx": nx * nx}[en.mat]
                ref = o.array(en.mat)[(N - 1) * per + en.i + en.j * rowsN[en.mat]]
                assert_same_bits(tileN[i, slot], ref, f"{wl} tileN {en.mat}[{en.i},{en.j}] inst {i}")
        s.backward_pass()
        gains = s.array("gains").reshape(B, N - 1, G)
        Qu = s.array("Qu").reshape(B, N - 1, nu)
        lam = s.array("lam").reshape(B, N, nx)
        sd = s.array("sd").reshape(-1, B)
        for i, o in enumerate(orc):
            st = o.backward_pass()
            assert st == 0
            eq = o.array("eq")[:(N - 1) * K * (nx + 1)].reshape(N - 1, K * (nx + 1))
            iq = o.array("ineq")[:(N - 1) * 2 * nu * (nx + 1)].reshape(N - 1, 2 * nu * (nx + 1))
            assert_same_bits(gains[i, :, :K * (nx + 1)], eq, f"{wl} eq gains inst {i} round {rnd}")
            assert_same_bits(gains[i, :, K * (nx + 1):], iq, f"{wl} ineq gains inst {i} round {rnd}")
            assert_same_bits(Qu[i].reshape(-1), o.array("Qu")[:(N - 1) * nu], f"{wl} Qu inst {i}")
            assert_same_bits(lam[i].reshape(-1), o.array("lam"), f"{wl} costate inst {i}")
            assert_same_bits(sd[1, i], o.result().reg_last, f"{wl} reg_last inst {i}")
        nf = s.check()
        r = s.results()
        need_fwd = 0
        for i, o in enumerate(orc):
            d, pr, cs0, csm = o.errors()
            assert_same_bits(r.dual_inf[i], d, f"{wl} dual_inf inst {i}")
            assert_same_bits(r.primal_inf[i], pr, f"{wl} primal_inf inst {i}")
            assert_same_bits(r.cs_inf[i], cs0, f"{wl} cs_inf inst {i}")
            need_fwd += 1
        # the first rounds never converge nor update the barrier parameter on these workloads with mu = 1
        if nf != need_fwd:
            break
        s.forward_pass()
        for i, o in enumerate(orc):
            assert o.forward_pass() == 0
            o.accept_step()
        cmp_traj()
        r = s.results()
        for i, o in enumerate(orc):
            ro = o.result()
            assert (int(r.k[i]), int(r.l[i])) == (ro.k, ro.l)
            assert_same_bits(r.step_size[i], ro.step_size, f"{wl} step inst {i}")
            assert_same_bits(r.objective[i], ro.objective, f"{wl} objective inst {i}")
    s.close()


def ldlt_parity(lib, oracle, rng, nmat=200, nmax=35, device=0):
    """Device dsytf2_rook / inertia / dsytrs_rook vs the oracle's restatement: bit-identical factors, pivots,
    info, inertia and solutions."""
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
    for n in sorted(set([1, 2, 3, 4, 8, 14, 15, 17, 32, 33, 35, 48, 64])):
        if n > nmax:
            continue
        mats = []
        for i in range(nmat):
            kind = i % 6
            if kind == 4:     # small integers: many equal magnitudes (first-index tie rules), exact zeros, singular columns
                M = rng.integers(-2, 3, size=(n, n)).astype(np.float64) * (rng.uniform(size=(n, n)) < 0.5)
                M = np.triu(M) + np.triu(M, 1).T
            elif kind == 5 and n >= 4:   # IPDDP-like KKT: diagonal H (wide dynamic range) + sparse constraint rows, zero block
                m = max(1, (3 * n) // 5); p = n - m
                H = np.diag(10.0 ** rng.uniform(-8, 8, m))
                H[0, 1] = H[1, 0] = rng.standard_normal()
                A = rng.integers(-1, 2, size=(p, m)).astype(np.float64) * (rng.uniform(size=(p, m)) < 0.2) * rng.uniform(0.5, 2.0, (p, m))
                M = np.zeros((n, n)); M[:m, :m] = H; M[:m, m:] = A.T; M[m:, :m] = A
            elif kind == 0 or n < 4:
                M = rng.standard_normal((n, n)); M = M + M.T
            elif kind == 1:
                m = max(1, (2 * n) // 3); p = n - m
                H = rng.standard_normal((m, m)); H = H @ H.T + np.diag(rng.uniform(0, 5, m))
                A = rng.standard_normal((p, m)) * (rng.uniform(size=(p, m)) < 0.4)
                M = np.zeros((n, n)); M[:m, :m] = H; M[:m, m:] = A.T; M[m:, :m] = A
            elif kind == 2:
                M = rng.standard_normal((n, n)) * (rng.uniform(size=(n, n)) < 0.3); M = M + M.T
                np.fill_diagonal(M, 0.0)
            else:
                M = rng.standard_normal((n, n)); M = M + M.T
                sc = 10.0 ** rng.uniform(-5, 5, n); M = M * sc[:, None] * sc[None, :]
            mats.append(np.asfortranarray(M))
        A = np.stack([m.reshape(-1, order="F") for m in mats])
        Bm = rng.standard_normal((nmat, n * 5))
        Aout = np.zeros_like(A); X = np.zeros_like(Bm)
        ipiv = np.zeros((nmat, n), dtype=np.int32); info = np.zeros(nmat, dtype=np.int32); npos = np.zeros(nmat, dtype=np.int32)
        rc = lib.L.ipddp_test_ldlt(n, nmat, A.ctypes.data_as(dp), Bm.ctypes.data_as(dp), Aout.ctypes.data_as(dp),
                                   ipiv.ctypes.data_as(ip), info.ctypes.data_as(ip), npos.ctypes.data_as(ip),
                                   X.ctypes.data_as(dp), device)
        lib.check(rc, "ipddp_test_ldlt")
        n2 = 0
        for i in range(nmat):
            a = mats[i].copy(order="F")
            piv = np.zeros(n + 1, dtype=np.int32)
            inf_o = oracle.lib().oracle_sytf2_rook(n, a.ctypes.data_as(dp), n, piv.ctypes.data_as(ip))
            assert inf_o == info[i], f"n={n} mat {i}: info {info[i]} vs {inf_o}"
            assert np.array_equal(piv[:n], ipiv[i]), f"n={n} mat {i}: pivots differ"
            iu = np.triu_indices(n)
            got = Aout[i].reshape(n, n, order="F")
            assert_same_bits(got[iu], a[iu], f"n={n} mat {i} factors")
            np_o = oracle.lib().oracle_inertia_np(n, a.ctypes.data_as(dp), n, piv.ctypes.data_as(ip), 1e-12)
            assert np_o == npos[i]
            n2 += int((piv[:n] < 0).any())
            if inf_o == 0:
                bo = np.asfortranarray(Bm[i].reshape(n, 5, order="F").copy())
                oracle.lib().oracle_sytrs_rook(n, 5, a.ctypes.data_as(dp), n, piv.ctypes.data_as(ip),
                                               bo.ctypes.data_as(dp), n)
                assert_same_bits(X[i].reshape(n, 5, order="F"), bo, f"n={n} mat {i} solution")
        if n >= 4:
            assert n2 > nmat // 10


def detmath_parity(lib, oracle, rng, n=200000, device=0):
    dp = C.POINTER(C.c_double)
    cases = {
        0: np.concatenate([rng.uniform(-10, 10, n), rng.uniform(-1e4, 1e4, n), rng.uniform(-1e-3, 1e-3, 1000), [0.0, np.inf, np.nan]]),
        3: np.concatenate([rng.uniform(1e-300, 10, n), np.exp(rng.uniform(-700, 700, n)), [0.0, -1.0, np.inf, 5e-324, 1.0]]),
        4: np.concatenate([rng.uniform(-745, 710, n), rng.uniform(-1, 1, n), [800.0, -800.0, 0.0]]),
    }
    cases[1] = cases[0]; cases[2] = cases[0]
    for fn, x in cases.items():
        x = np.ascontiguousarray(x); y = np.zeros_like(x); out = np.zeros_like(x)
        lib.check(lib.L.ipddp_test_detmath(fn, x.size, x.ctypes.data_as(dp), y.ctypes.data_as(dp), out.ctypes.data_as(dp), device), "detmath")
        ref = oracle.detmath(fn, x)
        assert np.array_equal(bits(out), bits(ref)) or np.all((bits(out) == bits(ref)) | (np.isnan(out) & np.isnan(ref))), f"detmath fn {fn}"
    x = np.ascontiguousarray(rng.uniform(1e-9, 3, n)); y = np.ascontiguousarray(rng.uniform(-3, 4, n)); out = np.zeros_like(x)
    lib.check(lib.L.ipddp_test_detmath(5, n, x.ctypes.data_as(dp), y.ctypes.data_as(dp), out.ctypes.data_as(dp), device), "detmath")
    assert np.array_equal(bits(out), bits(oracle.detmath(5, x, y)))


def division_parity(lib, rng, n=400000, device=0):
    """The reciprocal-based division used in the 2x2 pivot path (csrc/ldlt_warp.cuh: DivBy, Markstein's sequence with a
    plain-division fallback for out-of-range operands) returns the correctly rounded quotient, i.e. the bits of x / y."""
    dp = C.POINTER(C.c_double)

    def rnd(lo, hi, m):   # random sign, random mantissa, exponent uniform in [lo, hi]
        return rng.choice([-1.0, 1.0], m) * rng.uniform(1.0, 2.0, m) * np.exp2(rng.integers(lo, hi + 1, m).astype(np.float64))
    xs = [rnd(-60, 60, n), rnd(-390, 390, n), rnd(-1020, 1020, n // 4), rnd(-10, 10, n // 4)]
    ys = [rnd(-60, 60, n), rnd(-390, 390, n), rnd(-1020, 1020, n // 4), rnd(-10, 10, n // 4)]
    # divisors / numerators with extreme mantissas (all ones, all zeros), zeros, infinities, NaN, denormals
    edge = np.array([0.0, -0.0, 1.0, -1.0, np.inf, -np.inf, np.nan, 5e-324, 2.2250738585072014e-308, 1.7976931348623157e308,
                     np.nextafter(2.0, 1.0), np.nextafter(1.0, 2.0), 3.0, 1.0 / 3.0, 2.0 ** -400, 2.0 ** 400, 2.0 ** -401, 2.0 ** 401])
    ex, ey = np.meshgrid(edge, edge)
    xs.append(ex.ravel()); ys.append(ey.ravel())
    ones = np.nextafter(np.exp2(rng.integers(-50, 50, n // 4).astype(np.float64)) * 2.0, 0.0)   # mantissa all ones
    xs.append(rnd(-50, 50, n // 4)); ys.append(ones)
    xs.append(ones); ys.append(rnd(-50, 50, n // 4))
    x = np.ascontiguousarray(np.concatenate(xs)); y = np.ascontiguousarray(np.concatenate(ys)); out = np.zeros_like(x)
    lib.check(lib.L.ipddp_test_detmath(6, x.size, x.ctypes.data_as(dp), y.ctypes.data_as(dp), out.ctypes.data_as(dp), device), "division")
    with np.errstate(all="ignore"):
        ref = x / y
    same = (bits(out) == bits(ref)) | (np.isnan(out) & np.isnan(ref))
    assert same.all(), f"{int((~same).sum())} quotients differ, first: {x[~same][0]!r} / {y[~same][0]!r} -> {out[~same][0]!r} vs {ref[~same][0]!r}"


def sampled_parity(lib, oracle, wl, B, N, sample=6, vary_horizon=False, first=0, seed=0, tol=1e-7):
    """A batch at BASELINE-config scale on the device, a random sample of its instances on the oracle: same status,
    iteration counts, work counters, objective bits and trajectories for the sampled instances; every instance ends
    with a valid status."""
    b = instances.make_batch(wl, B, N, vary_horizon=vary_horizon, first=first)
    s = BatchSolver(wl, B, N, options=lib.default_options(optimality_tolerance=tol), lib=lib)
    s.set_batch(b)
    r = s.solve()
    x, u = s.trajectory()
    cnt = s.counters()
    s.close()
    assert np.isin(r.status, [0, 1, 7, 8]).all()
    idx = np.sort(np.random.default_rng(seed).choice(B, sample, replace=False))
    oopt = oracle.default_options(optimality_tolerance=tol)
    res, xo, uo = oracle.solve_batch(wl, N, b.p[idx], b.lower[idx], b.upper[idx], b.x1[idx], b.ubar[idx], options=oopt,
                                     horizons=b.horizons[idx], want_traj=True)
    for q, i in enumerate(idx):
        o = res[q]
        assert (int(r.status[i]), int(r.k[i]), int(r.j[i]), int(r.l[i])) == (o.status, o.k, o.j, o.l), f"{wl} inst {i}"
        assert (cnt["n_backward"][i], cnt["n_sweeps"][i], cnt["n_kkt"][i], cnt["n_rollouts"][i]) == \
            (o.n_backward, o.n_sweeps, o.n_kkt, o.n_rollouts), f"{wl} inst {i}: work counters"
        for name in ("objective", "primal_inf", "dual_inf", "cs_inf", "mu", "reg_last", "step_size"):
            assert_same_bits(getattr(r, name)[i], getattr(o, name), f"{wl} inst {i} {name}")
        assert_same_bits(x[i], xo[q], f"{wl} inst {i} states")
        assert_same_bits(u[i], uo[q], f"{wl} inst {i} controls")
    return r


FORCED_STATUS_OPTIONS = {
    "status1": dict(reg_max=1e-6),                                   # the first inertia failure asks for reg_1 = 1e-4 > reg_max
    "status7": dict(gamma_theta=1.0, gamma_L=1e30, delta=1e300),      # no step can pass the sufficient-decrease tests
    # mu never moves, every accepted step augments the filter, the fraction-to-boundary rule keeps the steps short:
    "status9": dict(kappa_eps=0.0, delta=1e300, eta_L=1e300, tau_min=1e-3),
}


def forced_status_parity(lib, oracle, case, wl_B, N):
    """The algorithm's failure exits, forced through the options and compared with the oracle bit for bit:
    status 1 -- regularisation exceeds reg_max in the backward pass (reference src/backward_pass.jl:55-58,
    src/inertia_correction.jl:264-275); status 7 -- the line search runs out of step sizes (src/forward_pass.jl:55);
    status 9 -- the device's fixed-capacity filter overflows (the oracle emulates the capacity on request; the
    reference's filter is an unbounded Vector)."""
    wl, B = wl_B
    kw = FORCED_STATUS_OPTIONS[case]
    if case == "status9":
        oracle.set_filter_capacity(64)
    try:
        b = instances.make_batch(wl, B, N, first=7)
        s = BatchSolver(wl, B, N, options=lib.default_options(optimality_tolerance=1e-7, **kw), lib=lib)
        s.set_batch(b)
        r = s.solve()
        x, u = s.trajectory()
        cnt = s.counters()
        s.close()
        res, xo, uo = oracle.solve_batch(wl, N, b.p, b.lower, b.upper, b.x1, b.ubar,
                                         options=oracle.default_options(optimality_tolerance=1e-7, **kw), want_traj=True)
    finally:
        oracle.set_filter_capacity(0)
    want = {"status1": 1, "status7": 7, "status9": 9}[case]
    assert (r.status == want).sum() >= (B if case != "status1" else 1), (case, r.status)
    for i in range(B):
        o = res[i]
        assert (int(r.status[i]), int(r.k[i]), int(r.j[i]), int(r.l[i])) == (o.status, o.k, o.j, o.l), (case, i)
        assert (cnt["n_sweeps"][i], cnt["n_kkt"][i], cnt["n_rollouts"][i]) == (o.n_sweeps, o.n_kkt, o.n_rollouts), (case, i)
        for name in ("objective", "primal_inf", "dual_inf", "cs_inf", "mu", "reg_last", "step_size"):
            assert_same_bits(getattr(r, name)[i], getattr(o, name), f"{case} inst {i} {name}")
    assert_same_bits(x, xo, f"{case} states")
    assert_same_bits(u, uo, f"{case} controls")


def api_error_convention(lib):
    """The C ABI's error convention (include/ipddp_b200.h): API misuse returns a non-zero code and leaves a message in
    ipddp_last_error(); nothing throws, nothing is left half-built; algorithmic outcomes are per-instance status codes."""
    import ctypes as C
    import pytest
    from ipddp_b200 import instances
    from ipddp_b200.batch import BatchSolver
    L = lib.L
    err = lambda: L.ipddp_last_error().decode()
    opt = lib.default_options()
    h = C.c_void_p()
    # unknown model, bad sizes
    assert L.ipddp_problem_create(b"no_such_model", 2, 11, None, 0, C.byref(opt), 0, 0, C.byref(h)) != 0
    assert "unknown model" in err() and not h.value
    assert L.ipddp_problem_create(b"concar", 0, 11, None, 0, C.byref(opt), 0, 0, C.byref(h)) != 0
    assert "B >= 1" in err()
    assert L.ipddp_problem_create(b"concar", 2, 1, None, 0, C.byref(opt), 0, 0, C.byref(h)) != 0
    nx, nu, nc, np_, _ = lib.model_dims("concar")
    bad_ic = np.array([nc], dtype=np.int32)      # complementarity index out of range
    assert L.ipddp_problem_create(b"concar", 2, 11, bad_ic.ctypes.data_as(C.POINTER(C.c_int)), 1, C.byref(opt), 0, 0,
                                  C.byref(h)) != 0
    assert "indices_compl" in err()
    with pytest.raises(RuntimeError, match="unknown model"):
        lib.model_dims("no_such_model")
    assert L.ipddp_model_load(b"/nonexistent/plugin.so") != 0 and "dlopen" in err()
    # a valid problem: calls in the wrong order / with bad inputs
    B, N = 2, 11
    s = BatchSolver("concar", B, N, options=opt, lib=lib)
    try:
        assert L.ipddp_solve(s.h, 0) != 0 and "ipddp_set_inputs not called" in err()
        assert L.ipddp_initialize(s.h) != 0
        hs = (C.c_void_p * 1)(s.h)
        ms, st = C.c_double(), type(s.stats())()
        assert L.ipddp_solve_many(hs, 1, 1, 0, C.byref(ms), C.byref(st)) != 0
        b = instances.make_batch("concar", B, N)
        with pytest.raises(RuntimeError, match="required"):      # missing parameter vector (np > 0) / missing arrays
            s.lib.check(L.ipddp_set_inputs(s.h, None, None, None, None, None, None), "ipddp_set_inputs")
        hz = np.array([N, N + 1], dtype=np.int32)
        with pytest.raises(RuntimeError, match="horizon out of range"):
            s.set_inputs(b.x1, b.ubar, b.p, b.lower, b.upper, hz)
        hz[1] = 1
        with pytest.raises(RuntimeError, match="horizon out of range"):
            s.set_inputs(b.x1, b.ubar, b.p, b.lower, b.upper, hz)
        with pytest.raises(RuntimeError, match="unknown tuning key"):
            s.set_tuning("no_such_key", 1)
        s.set_batch(b)
        assert L.ipddp_solve_many(hs, 1, 0, 0, C.byref(ms), C.byref(st)) != 0 and "total_solves" in err()
        # after all that misuse the handle still solves, and the outcome is in the status codes (0 here)
        r = s.solve()
        assert r.status.tolist() == [0, 0]
        # algorithmic failure is NOT an API error: max_iterations = 3 ends with status 8 and return code 0
        o3 = lib.default_options(max_iterations=3)
        s3 = BatchSolver("concar", B, N, options=o3, lib=lib)
        s3.set_batch(b)
        r3 = s3.solve()
        assert r3.status.tolist() == [8, 8] and r3.k.tolist() == [3, 3]
        s3.close()
        buf = np.zeros(4)
        assert L.ipddp_get_array(s.h, b"no_such_array", buf.ctypes.data_as(C.POINTER(C.c_double))) < 0
        assert "unknown array" in err()
    finally:
        s.close()


# Every Options field the algorithm reads (reference src/options.jl:1-38), moved away from its default by a value that
# changes the iterates of small pushing / concar solves (checked on the oracle: tests/test_emu_parity.py asserts it).
OPTION_VARIANTS = dict(mu_init=0.3, kappa_1=0.05, reg_1=1e-3, kappa_bar_w_p=50.0, kappa_w_p=5.0, kappa_w_m=0.5, kappa_eps=5.0,
                       kappa_mu=0.3, theta_mu=1.4, tau_min=0.95, s_max=10.0, s_L=2.0, delta=1e-4, s_theta=1.2, eta_L=0.4,
                       gamma_theta=0.3)
OPTION_GROUPS = [dict(mu_init=0.3, kappa_eps=5.0, kappa_mu=0.3, theta_mu=1.4, s_max=10.0, tau_min=0.95),
                 dict(reg_1=1e-3, kappa_bar_w_p=50.0, kappa_w_p=5.0, kappa_w_m=0.5, reg_min=1e-3),
                 dict(s_L=2.0, delta=1e-4, s_theta=1.2, eta_L=0.4, gamma_theta=0.3, gamma_L=0.3)]


def options_parity(lib, oracle, variants, wl="pushing", B=4, N=21, maxit=120, first=3, mutate=None):
    """Solves the same instances once per option set in `variants` (list of dicts) on the library and on the oracle:
    status, k, j, l, objective and trajectories bit for bit.  Returns, per variant, whether it changed the iterates with
    respect to the default options (so that callers can assert the sweep is discriminating)."""
    b = instances.make_batch(wl, B, N, first=first)
    if mutate is not None:
        mutate(b)

    def run(kw):
        base = dict(optimality_tolerance=1e-7, max_iterations=maxit)
        base.update(kw)
        s = BatchSolver(wl, B, N, options=lib.default_options(**base), lib=lib)
        s.set_batch(b)
        r = s.solve()
        x, u = s.trajectory()
        s.close()
        res, xo, uo = oracle.solve_batch(wl, N, b.p, b.lower, b.upper, b.x1, b.ubar, options=oracle.default_options(**base),
                                         want_traj=True)
        for i in range(B):
            o = res[i]
            got = (int(r.status[i]), int(r.k[i]), int(r.j[i]), int(r.l[i]))
            assert got == (o.status, o.k, o.j, o.l), f"{wl} options {kw} inst {i}: {got} vs oracle {(o.status, o.k, o.j, o.l)}"
            assert_same_bits(r.objective[i], o.objective, f"{wl} options {kw} inst {i} objective")
        assert_same_bits(x, xo, f"{wl} options {kw} states")
        assert_same_bits(u, uo, f"{wl} options {kw} controls")
        return [(o.status, o.k, o.objective) for o in res]
    base = run({})
    return [run(kw) != base for kw in variants]


def duals_parity(lib, oracle, wl="concar", B=3, N=11):
    """ipddp_get_duals (reference problem.nominal_*_duals) after a solve that converges and after one that stops at
    max_iterations: phi, zl, zu and the costate against the oracle.  The costate of a status-8 exit is zero in the reference
    (update_nominal_trajectory! clears it after the last accepted step, src/data/methods.jl:89)."""
    b = instances.make_batch(wl, B, N)
    for maxit, want in ((5, 8), (80, 0)):
        s = BatchSolver(wl, B, N, options=lib.default_options(optimality_tolerance=1e-7, max_iterations=maxit), lib=lib)
        s.set_batch(b)
        r = s.solve()
        assert (r.status == want).all(), (maxit, r.status)
        phi, zl, zu, lam = s.duals()
        s.close()
        for i in range(B):
            o = oracle.OracleSolver(wl, N, b.p[i], b.lower[i], b.upper[i],
                                    options=oracle.default_options(optimality_tolerance=1e-7, max_iterations=maxit))
            ro = o.solve(b.x1[i], b.ubar[i])
            assert ro.status == want
            for name, arr in (("phi", phi), ("zl", zl), ("zu", zu), ("lam", lam)):
                assert_same_bits(np.asarray(arr[i]).reshape(-1), o.array(name), f"{wl} {name} inst {i} (status {want})")
            if want == 8:
                assert not lam[i].any()
