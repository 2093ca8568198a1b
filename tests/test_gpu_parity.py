"""GPU parity: the CUDA path, called through the C-ABI (bluerov2_b200.solver -> include/bluerov2_b200.h), against the
CPU oracle and the committed golden fixtures.  Run on the B200 box: ``pytest -m gpu``.

Tolerances (fp64 everywhere; the north star asks for |u_gpu - u_ref|_inf < 1e-4):
  * linearisation (A, B, b):   1e-11 absolute   (same arithmetic, different summation order / libm vs CUDA sincos)
  * RTI step (X, U, u0):       1e-6  absolute   (two IPMs stopped at 1e-11 / 1e-12 on a QP with cond ~1e5)
  * EKF (x, P):                1e-6  relative   (forward differences with d = 1e-6 amplify 1-ulp sincos differences by 1e6)
"""
import numpy as np
import pytest

from bluerov2_b200 import traj, workloads as wl

pytestmark = pytest.mark.gpu

TOL_LIN = 1e-11
TOL_U = 1e-6
NORTH_STAR_TOL = 1e-4


@pytest.fixture(scope="module")
def solver_mod():
    from bluerov2_b200 import solver
    if solver.device_count() < 1:
        pytest.fail("no CUDA device visible: the gpu-marked tests must run on the B200 box")
    return solver


def _replay(solver_mod, g, tag, use_torch=False):
    N = int(g[tag + "_N"])
    B = g[tag + "_X0"].shape[0]
    tr = traj.circle() if str(g[tag + "_ref"]) == "circle" else traj.lemniscate()
    s = solver_mod.BatchSolver(B, N)
    s.set_iterate(g[tag + "_X0"], g[tag + "_U0"])
    worst = 0.0
    for t in range(g[tag + "_x0"].shape[0]):
        yref = traj.window_batch(tr, g[tag + "_lines"][t], N)
        if use_torch:
            import torch
            dev = torch.device("cuda", 0)
            u0, th, st = s.solve(torch.from_numpy(g[tag + "_x0"][t]).to(dev), torch.from_numpy(yref).to(dev),
                                 torch.from_numpy(g[tag + "_p"]).to(dev))
            torch.cuda.synchronize()
            u0, th, st = u0.cpu().numpy(), th.cpu().numpy(), st.cpu().numpy()
        else:
            u0, th, st = s.solve(g[tag + "_x0"][t], yref, g[tag + "_p"])
        assert (st == 0).all(), (tag, t, st)
        X, U = s.get_iterate()
        eu = np.abs(U - g[tag + "_Uout"][t]).max()
        ex = np.abs(X - g[tag + "_Xout"][t]).max()
        worst = max(worst, eu)
        assert eu < TOL_U and ex < TOL_U, (tag, t, eu, ex)
        assert np.array_equal(u0, U[:, 0, :])
        # thrust allocation, bluerov2_dob.cpp:390-395
        rc = 0.026546960744430276
        want = np.stack([-u0[:, 0] + u0[:, 1] + u0[:, 3], -u0[:, 0] - u0[:, 1] - u0[:, 3], u0[:, 0] + u0[:, 1] - u0[:, 3],
                         u0[:, 0] - u0[:, 1] + u0[:, 3], -u0[:, 2], -u0[:, 2]], axis=1) / rc
        assert np.allclose(th, want, rtol=1e-14, atol=1e-12)
        # the golden iterate is carried (both sides then linearise at the same point)
        s.set_iterate(g[tag + "_Xout"][t], g[tag + "_Uout"][t])
    s.close()
    return worst


@pytest.mark.parametrize("tag", ["nom40", "act40", "lem40", "act20", "act80"])
def test_golden_rti_cases(solver_mod, golden, tag):
    worst = _replay(solver_mod, golden["rti_cases"], tag)
    assert worst < NORTH_STAR_TOL


def test_golden_rti_cases_device_pointers(solver_mod, golden):
    """same through br2_batch_solve_device with torch CUDA tensors on torch's current stream"""
    _replay(solver_mod, golden["rti_cases"], "act40", use_torch=True)


@pytest.mark.parametrize("N", [80, 40, 20, 10])
def test_cold_start_known_answer(solver_mod, golden, N):
    """main_bluerov2.c-style cold solve from the generated initial guess (acados_solver_bluerov2.c:681-708)."""
    from oracle import NOMINAL_P
    g = golden["rti_cases"]
    s = solver_mod.BatchSolver(3, N)       # three identical instances: also checks instance independence
    x0 = np.tile(g["cold_x0"], (3, 1))
    yref = np.tile(traj.circle()[:N + 1][None], (3, 1, 1))
    u0, th, st = s.solve(x0, yref, np.tile(NOMINAL_P, (3, 1)))
    assert (st == 0).all()
    X, U = s.get_iterate()
    assert np.abs(U - g[f"cold_N{N}_U"][None]).max() < TOL_U
    assert np.abs(X - g[f"cold_N{N}_X"][None]).max() < TOL_U
    assert np.array_equal(U[0], U[1]) and np.array_equal(U[0], U[2])
    s.close()


def test_linearization_matches_oracle(solver_mod, oracle):
    N, B = 40, 64
    w = wl.tracking_batch(B, N, seed=11, pos_spread=1.0)
    rng = np.random.default_rng(5)
    X = w["X"] + rng.uniform(-0.3, 0.3, w["X"].shape)
    U = rng.uniform(-50, 50, w["U"].shape)
    X[:4, :, 6:9] = 0.0       # sign(0) branch of d|v|v/dv
    p = w["p"].copy()
    p[:, :4] = rng.uniform(-10, 10, (B, 4))
    s = solver_mod.BatchSolver(B, N)
    s.set_iterate(X, U)
    s.set_option("qp_iter_max", 1)
    s.solve(w["x0"], w["yref"], p)
    A, Bm, b = s.linearization()
    Ts = wl.time_steps(N)
    for i in range(B):
        Ao, Bo, bo = oracle.linearize(Ts, p[i], X[i], U[i])
        assert np.abs(A[i] - Ao).max() < TOL_LIN
        assert np.abs(Bm[i] - Bo).max() < TOL_LIN
        assert np.abs(b[i] - bo).max() < TOL_LIN
    s.close()


def test_per_stage_parameters(solver_mod, oracle):
    N, B = 20, 4
    w = wl.tracking_batch(B, N, seed=9)
    rng = np.random.default_rng(1)
    pp = np.tile(w["p"][:, None, :], (1, N + 1, 1))
    pp[:, :, :4] = rng.uniform(-8, 8, (B, N + 1, 4))
    s = solver_mod.BatchSolver(B, N)
    s.set_iterate(w["X"], w["U"])
    u0, th, st = s.solve(w["x0"], w["yref"], pp)
    assert (st == 0).all()
    Ts = wl.time_steps(N)
    for i in range(B):
        X, U = w["X"][i].copy(), w["U"][i].copy()
        sto, _ = oracle.rti_step(Ts, w["x0"][i], w["yref"][i], pp[i], X, U)
        assert sto == 0 and np.abs(U[0] - u0[i]).max() < TOL_U
    s.close()


def test_weights_bounds_and_ragged_batch(solver_mod, oracle):
    """non-default W / bounds (runtime-settable, SURVEY finding 5) and a batch that is not a multiple of anything"""
    N, B = 10, 37
    w = wl.tracking_batch(B, N, seed=21, pos_spread=2.0)
    W = np.array([300, 480, 200, 10, 10, 200, 10, 10, 10, 10, 10, 10, 1, 1, 0.1, 0.05])   # generate_c_code.py:34
    We = W[:12] * 2.0
    lbu, ubu = np.array([-3.0, -4, -5, -0.5]), np.array([2.0, 4, 5, 0.5])
    s = solver_mod.BatchSolver(B, N)
    s.set_weights(W, We)
    s.set_bounds(lbu, ubu)
    s.set_iterate(w["X"], w["U"])
    u0, th, st = s.solve(w["x0"], w["yref"], w["p"])
    assert (st == 0).all()
    Ts = wl.time_steps(N)
    X, U = w["X"].copy(), w["U"].copy()
    sto, _, _ = oracle.rti_step_batch(Ts, w["x0"], w["yref"], w["p"], X, U, W=W, We=We, lbu=lbu, ubu=ubu)
    assert (sto == 0).all()
    assert np.abs(U[:, 0] - u0).max() < TOL_U
    Xg, Ug = s.get_iterate()
    assert np.abs(Ug - U).max() < TOL_U
    assert (Ug <= ubu + 1e-9).all() and (Ug >= lbu - 1e-9).all()
    assert (np.abs(Ug - ubu) < 1e-6).any() or (np.abs(Ug - lbu) < 1e-6).any(), "no bound active in the active-bound case"
    s.close()


@pytest.mark.parametrize("B,N,spread", [(4096, 40, 0.5), (4096, 40, 3.0),          # BASELINE config 2 (both input sets)
                                         (8192, 40, 0.5),                            # config 4: the per-GPU shard of 65536 / 8
                                         (8192, 10, 0.5), (8192, 20, 3.0), (8192, 80, 0.5)])   # config 5: horizon sweep
def test_full_size_batch_properties(solver_mod, oracle, B, N, spread):
    """BASELINE configurations at full size: a sample of instances against the oracle, the rest through size-independent
    properties: bounds respected, states = exact roll-out of the inputs (dynamics residual of the QP), instance
    independence (a permuted batch gives the permuted answer bit for bit)."""
    w = wl.tracking_batch(B, N, seed=0, pos_spread=spread)
    s = solver_mod.BatchSolver(B, N)
    s.set_iterate(w["X"], w["U"])
    u0, th, st = s.solve(w["x0"], w["yref"], w["p"])
    assert (st == 0).all(), np.unique(st, return_counts=True)
    it, info = s.stats()
    assert it.max() < 50 and it.min() >= 1
    X, U = s.get_iterate()
    assert (np.abs(U) <= 50 + 1e-9).all()
    assert np.abs(X[:, 0] - w["x0"]).max() < 1e-12                 # stage-0 state pinned to x0 (lbx = ubx)
    A, Bm, b = s.linearization()
    dX, dU = X - w["X"], U - w["U"]
    roll = np.einsum("bkij,bkj->bki", A, dX[:, :-1]) + np.einsum("bkij,bkj->bki", Bm, dU) + b
    assert np.abs(roll - dX[:, 1:]).max() < 1e-9
    Ts = wl.time_steps(N)
    idx = np.arange(0, B, B // 64)
    Xo, Uo = w["X"][idx].copy(), w["U"][idx].copy()
    sto, _, _ = oracle.rti_step_batch(Ts, w["x0"][idx], w["yref"][idx], w["p"][idx], Xo, Uo)
    assert (sto == 0).all()
    assert np.abs(Uo - U[idx]).max() < TOL_U
    if spread >= 3.0:
        assert (np.abs(U) > 50 - 1e-6).any()
    # permutation invariance
    perm = np.random.default_rng(3).permutation(B)
    s.set_iterate(w["X"][perm], w["U"][perm])
    u0p, _, _ = s.solve(w["x0"][perm], w["yref"][perm], w["p"][perm])
    assert np.array_equal(u0p, u0[perm])
    s.close()


def test_closed_loop_ticks_match_oracle(solver_mod, oracle):
    """20 closed-loop ticks (plant = nominal ERK4 at 0.05 s), iterate carried on the device, parity on every tick."""
    N, B, T = 40, 32, 20
    w = wl.tracking_batch(B, N, seed=4, pos_spread=1.5)
    Ts = wl.time_steps(N)
    s = solver_mod.BatchSolver(B, N)
    s.set_iterate(w["X"], w["U"])
    Xo, Uo = w["X"].copy(), w["U"].copy()
    x0, lines = w["x0"].copy(), w["lines"].copy()
    for t in range(T):
        yref = traj.window_batch(w["traj"], lines, N)
        u0, th, st = s.solve(x0, yref, w["p"])
        sto, _, _ = oracle.rti_step_batch(Ts, x0, yref, w["p"], Xo, Uo)
        assert (st == 0).all() and (sto == 0).all()
        assert np.abs(u0 - Uo[:, 0]).max() < 1e-5, (t, np.abs(u0 - Uo[:, 0]).max())
        for i in range(B):
            x0[i] = oracle.erk4(x0[i], Uo[i, 0], w["p"][i], 0.05)
        lines = lines + 1
        # keep both sides on the same linearisation point so that differences do not compound through the loop
        s.set_iterate(Xo, Uo)
    s.close()


def test_ekf_golden(solver_mod, golden):
    g = golden["ekf_cases"]
    T, B = g["thr"].shape[:2]
    s = solver_mod.BatchSolver(B, 10)
    s.ekf_reset()
    for t in range(T):
        wf, p = s.ekf(g["thr"][t], g["meas"][t], g["acc"][t], compensate=True)
        x, P = s.ekf_state()
        sx = np.maximum(1.0, np.abs(g["ex"][t]))
        assert (np.abs(x - g["ex"][t]) / sx).max() < 1e-6, (t, (np.abs(x - g["ex"][t]) / sx).max())
        assert np.abs(P - g["eP"][t]).max() < 1e-6 * max(1.0, np.abs(g["eP"][t]).max()), t
        assert np.abs(wf - g["wf"][t]).max() < 1e-6 * max(1.0, np.abs(g["wf"][t]).max())
        assert np.allclose(p, wl.dob_params(x, True), rtol=1e-12, atol=1e-12)
        # carry the golden state: differences must not compound through the filter
        s.set_ekf_state(g["ex"][t], g["eP"][t])
    s.close()


def test_errors_are_loud(solver_mod):
    with pytest.raises(solver_mod.SolverError):
        solver_mod.BatchSolver(0, 40)
    with pytest.raises(solver_mod.SolverError):
        solver_mod.BatchSolver(4, 100000)
    s = solver_mod.BatchSolver(2, 10)
    with pytest.raises(solver_mod.SolverError):
        s.set_bounds(np.ones(4), -np.ones(4))
    with pytest.raises(ValueError):
        s.solve(np.zeros((3, 12)), np.zeros((2, 11, 16)), np.zeros((2, 16)))
    s.close()


def test_interior_fast_path_equals_ipm(solver_mod, oracle):
    """The interior-solution fast path (unconstrained Riccati solve accepted when it lies inside the box) and the
    interior-point iteration return the same QP minimiser; instances with active bounds fall through to the IPM."""
    N, B = 40, 512
    Ts = wl.time_steps(N)
    for spread in (0.5, 3.0):
        w = wl.tracking_batch(B, N, seed=17, pos_spread=spread)
        res = {}
        for fp in (0, 1):
            s = solver_mod.BatchSolver(B, N)
            s.set_option("fast_path", fp)
            s.set_iterate(w["X"], w["U"])
            u0, th, st = s.solve(w["x0"], w["yref"], w["p"])
            assert (st == 0).all()
            it, info = s.stats()
            X, U = s.get_iterate()
            res[fp] = (U.copy(), it.copy())
            s.close()
        assert np.abs(res[0][0] - res[1][0]).max() < TOL_U
        assert (res[0][1] >= 2).all()                      # the IPM never stops after one iteration from a cold start
        inside = np.abs(res[0][0]).max(axis=(1, 2)) < 50 - 1e-3
        assert ((res[1][1] == 1) == inside).mean() > 0.99   # fast path taken exactly where no bound is active
        if spread == 0.5:
            assert (res[1][1] == 1).mean() > 0.9
        else:
            assert (res[1][1] > 1).any()
        idx = np.arange(0, B, 16)
        Xo, Uo = w["X"][idx].copy(), w["U"][idx].copy()
        sto, _, _ = oracle.rti_step_batch(Ts, w["x0"][idx], w["yref"][idx], w["p"][idx], Xo, Uo)
        assert (sto == 0).all() and np.abs(Uo - res[1][0][idx]).max() < TOL_U


def test_device_windowing_equals_explicit_yref(solver_mod):
    """ref_cb on the device (bluerov2_dob.cpp:218-265): rows line..line+N clamped to the last row == the explicit window"""
    N, B = 20, 96
    tr = traj.lemniscate()
    rng = np.random.default_rng(8)
    lines = rng.integers(0, tr.shape[0] - N - 1, size=B).astype(np.int32)
    lines[:6] = [tr.shape[0] - 1, tr.shape[0] - 2, tr.shape[0] - N, tr.shape[0] - N - 1, 0, tr.shape[0] + 5]   # clamped tails
    yref = traj.window_batch(tr, lines.astype(np.int64), N)
    x0 = yref[:, 0, :12] + rng.uniform(-0.4, 0.4, (B, 12))
    p = np.tile(wl.NOMINAL_P, (B, 1))
    X = np.repeat(x0[:, None, :], N + 1, axis=1).copy(); U = np.zeros((B, N, 4))
    a = solver_mod.BatchSolver(B, N); a.set_iterate(X, U)
    ua, ta, sa = a.solve(x0, yref, p)
    b = solver_mod.BatchSolver(B, N); b.set_iterate(X, U); b.set_trajectory(tr)
    ub, tb, sb = b.solve_windowed(x0, lines, p)
    assert (sa == 0).all() and (sb == 0).all()
    assert np.array_equal(ua, ub) and np.array_equal(ta, tb)
    Xa, Ua = a.get_iterate(); Xb, Ub = b.get_iterate()
    assert np.array_equal(Xa, Xb) and np.array_equal(Ua, Ub)
    c = solver_mod.BatchSolver(B, N)
    with pytest.raises(solver_mod.SolverError):
        c.solve_windowed(x0, lines, p)            # no trajectory uploaded
    a.close(); b.close(); c.close()


def test_dob_mpc_closed_loop(solver_mod, oracle):
    """BASELINE config 3 in small: DOB-MPC ticks (EKF -> parameters -> RTI solve) with sampled wave disturbances on the
    lemniscate reference; GPU EKF + solve against oracle EKF + solve, device-resident hand-over of p.  The loop is
    settled first (MPC only) and the plant is driven by the uncompensated command: see bench.record_closed_loop_dob."""
    import torch
    N, B, T, SETTLE = 40, 48, 8, 40
    dev = torch.device("cuda", 0)
    w = wl.tracking_batch(B, N, seed=6, reference="lemniscate", pos_spread=0.2, level=True)
    amp, tau0 = wl.wave_disturbance(B, seed=1)
    Ts = wl.time_steps(N)
    s = solver_mod.BatchSolver(B, N)
    s.set_trajectory(w["traj"])
    s.set_iterate(w["X"], w["U"])
    x, lines = w["x0"].copy(), w["lines"].copy()
    vel_prev = x[:, 6:12].copy()
    for t in range(SETTLE):
        vel_prev = x[:, 6:12].copy()
        u0, thr, st = s.solve_windowed(x, lines.astype(np.int32), w["p"])
        assert (st == 0).all()
        x = wl.plant_step(x, u0, w["p"], 0.05, dist=wl.wave_at(amp, tau0, t))
        lines = lines + 1
    assert np.abs(u0).max() < 2.0          # the gentle regime the reference's filter lives in
    Xn, Un = s.get_iterate()               # iterate of the plant-driving (uncompensated) controller
    Xo, Uo = Xn.copy(), Un.copy()          # iterate of the compensated controller (oracle side; GPU side is reset to it)
    ox = np.zeros((B, 18)); oP = np.zeros((B, 18, 18))
    for i in range(B):
        ox[i], oP[i] = oracle.ekf_init()
    ox[:, :12] = x
    s.set_ekf_state(ox, oP)
    for t in range(SETTLE, SETTLE + T):
        dist = wl.wave_at(amp, tau0, t)
        acc = (x[:, 6:12] - vel_prev) / 0.05
        vel_prev = x[:, 6:12].copy()
        yref = traj.window_batch(w["traj"], lines, N)
        # --- GPU: EKF writes p on the device, the solve reads it there ---
        s.set_iterate(Xo, Uo)
        wf, p_dev = s.ekf(torch.from_numpy(thr).to(dev), torch.from_numpy(x).to(dev), torch.from_numpy(acc).to(dev))
        u0, th, st = s.solve_windowed(torch.from_numpy(x).to(dev), torch.from_numpy(lines.astype(np.int32)).to(dev), p_dev)
        torch.cuda.synchronize()
        u0, st, p_gpu = u0.cpu().numpy(), st.cpu().numpy(), p_dev.cpu().numpy()
        # --- oracle ---
        oracle.ekf_step_batch(ox, oP, thr, x, acc)
        p_or = wl.dob_params(ox, True)
        assert np.isfinite(ox).all()
        # EKF parity: the filter differentiates by forward differences with d = 1e-6 (bluerov2_dob.cpp:726,742), which
        # turns 1-ulp sincos differences into ~1e-10 Jacobian noise that the 18x18 inverse amplifies; parameters are
        # compared at 1e-5 relative ...
        assert np.abs(p_gpu - p_or).max() < 1e-5 * max(1.0, np.abs(p_or).max()), (t, np.abs(p_gpu - p_or).max())
        assert np.abs(p_or[:, :3]).max() > 10.0         # the compensation is really engaged
        # ... and the solve is compared on identical inputs (the parameters the GPU filter produced)
        sto, _, _ = oracle.rti_step_batch(Ts, x, yref, p_gpu, Xo, Uo)
        assert (st == 0).all() and (sto == 0).all(), (t, st, sto)
        assert np.abs(u0 - Uo[:, 0]).max() < 1e-5, (t, np.abs(u0 - Uo[:, 0]).max())
        s.set_ekf_state(ox, oP)             # carry the oracle's filter state on both sides
        # plant: driven by the uncompensated command
        stn, _, _ = oracle.rti_step_batch(Ts, x, yref, w["p"], Xn, Un)
        assert (stn == 0).all()
        thr = np.stack([oracle.thrust_alloc(Un[i, 0]) for i in range(B)])
        x = wl.plant_step(x, Un[:, 0].copy(), w["p"], 0.05, dist=dist)
        lines = lines + 1
    s.close()


@pytest.mark.parametrize("N", [1, 2, 3, 5, 64, 256])
def test_horizon_edge_cases(solver_mod, oracle, N):
    """horizons shorter than the prefetch ring, odd, and the engine maximum (NMAX = 256); both solution paths"""
    B = 8
    Ts = wl.time_steps(N)
    for spread, fp in ((0.3, 1), (3.0, 0)):
        w = wl.tracking_batch(B, N, seed=31 + N, pos_spread=spread)
        s = solver_mod.BatchSolver(B, N)
        s.set_option("fast_path", fp)
        s.set_iterate(w["X"], w["U"])
        u0, th, st = s.solve(w["x0"], w["yref"], w["p"])
        X, U = w["X"].copy(), w["U"].copy()
        sto, _, _ = oracle.rti_step_batch(Ts, w["x0"], w["yref"], w["p"], X, U)
        assert (st == 0).all() and (sto == 0).all(), (N, st, sto)
        Xg, Ug = s.get_iterate()
        # N = 256 (Ts = 3.9 ms: input curvature Ts*R = 2e-4, 1024 inequality rows): the stopping tolerance of either IPM maps
        # to ~1e-5 in u; outside the BASELINE horizons (10..80), checked at the north-star tolerance
        tol = TOL_U if N <= 64 else NORTH_STAR_TOL
        assert np.abs(Ug - U).max() < tol and np.abs(Xg - X).max() < tol, (N, np.abs(Ug - U).max())
        s.close()


def test_status_codes_and_nan_guard(solver_mod):
    """acados status conventions: 2 = QP iteration limit, 1 = NaN detected (iterate left untouched), and a healthy
    neighbour instance is not affected"""
    N, B = 20, 4
    w = wl.tracking_batch(B, N, seed=3, pos_spread=3.0)
    s = solver_mod.BatchSolver(B, N)
    s.set_option("fast_path", 0)
    s.set_option("qp_iter_max", 1)
    s.set_iterate(w["X"], w["U"])
    _, _, st = s.solve(w["x0"], w["yref"], w["p"])
    assert (st == 2).all()
    s.set_option("qp_iter_max", 50)
    s.set_option("fast_path", 1)
    x0 = w["x0"].copy()
    x0[1, 4] = np.nan
    s.set_iterate(w["X"], w["U"])
    u0, th, st = s.solve(x0, w["yref"], w["p"])
    assert st[1] == 1 and (np.delete(st, 1) == 0).all(), st
    X, U = s.get_iterate()
    assert np.array_equal(X[1], w["X"][1]) and np.array_equal(U[1], w["U"][1])     # NaN never reaches the iterate
    assert np.isfinite(X[[0, 2, 3]]).all() and np.isfinite(u0[[0, 2, 3]]).all()
    s.close()


def test_single_instance_batch(solver_mod, oracle):
    """B = 1 (what the acados ABI drives) over several closed-loop ticks at N = 20, Ts = 0.05 (BASELINE config 1)"""
    N = 20
    Ts = wl.time_steps(N)
    w = wl.tracking_batch(1, N, seed=12, pos_spread=0.5)
    s = solver_mod.BatchSolver(1, N)
    s.set_iterate(w["X"], w["U"])
    X, U = w["X"].copy(), w["U"].copy()
    x0, line = w["x0"].copy(), w["lines"].copy()
    for t in range(10):
        yref = traj.window_batch(w["traj"], line, N)
        u0, _, st = s.solve(x0, yref, w["p"])
        sto, _, _ = oracle.rti_step_batch(Ts, x0, yref, w["p"], X, U)
        assert st[0] == 0 and sto[0] == 0
        assert np.abs(u0 - U[:, 0]).max() < 1e-5
        x0 = wl.plant_step(x0, U[:, 0].copy(), w["p"], 0.05)
        line = line + 1
        s.set_iterate(X, U)
    s.close()


def test_device_closed_loop_matches_host_loop(solver_mod, oracle):
    """Device-resident closed loop: solve -> plant step -> row counter, nothing leaves the GPU between ticks;
    compared with the same loop stepped on the host (oracle plant), including the wave disturbance and body acceleration."""
    import torch
    N, B, T = 20, 64, 12
    dev = torch.device("cuda", 0)
    w = wl.tracking_batch(B, N, seed=13, reference="lemniscate", pos_spread=0.4)
    amp, tau0 = wl.wave_disturbance(B, seed=2)
    s = solver_mod.BatchSolver(B, N)
    s.set_trajectory(w["traj"])
    s.set_iterate(w["X"], w["U"])
    x = torch.from_numpy(w["x0"].copy()).to(dev)
    lines = torch.from_numpy(w["lines"].astype(np.int32)).to(dev)
    p = torch.from_numpy(w["p"]).to(dev)
    d_amp, d_tau0 = torch.from_numpy(amp).to(dev), torch.from_numpy(tau0).to(dev)
    acc = torch.zeros((B, 6), dtype=torch.float64, device=dev)
    out = None
    xh, lh = w["x0"].copy(), w["lines"].copy()
    s2 = solver_mod.BatchSolver(B, N); s2.set_trajectory(w["traj"]); s2.set_iterate(w["X"], w["U"])
    for t in range(T):
        out = s.solve_windowed(x, lines, p, out=out)
        solver_mod.plant_step(x, out[0], p, 0.05, wave=(d_amp, d_tau0), tick=t, body_acc=acc, lines=lines)
        # host-stepped twin
        u0h, _, sth = s2.solve_windowed(xh, lh.astype(np.int32), w["p"])
        assert (sth == 0).all()
        xn = wl.plant_step(xh, u0h, w["p"], 0.05, dist=wl.wave_at(amp, tau0, t))
        acch = (xn[:, 6:12] - xh[:, 6:12]) / 0.05
        xh, lh = xn, lh + 1
    torch.cuda.synchronize()
    assert (out[2].cpu().numpy() == 0).all()
    assert np.array_equal(lines.cpu().numpy(), lh.astype(np.int32))
    assert np.abs(x.cpu().numpy() - xh).max() < 1e-9
    assert np.abs(acc.cpu().numpy() - acch).max() < 1e-7
    # one plant step against the oracle's RK4
    x1 = torch.from_numpy(w["x0"].copy()).to(dev)
    u1 = torch.from_numpy(np.random.default_rng(0).uniform(-20, 20, (B, 4))).to(dev)
    solver_mod.plant_step(x1, u1, p, 0.05)
    want = np.stack([oracle.erk4(w["x0"][i], u1.cpu().numpy()[i], w["p"][i], 0.05) for i in range(B)])
    assert np.abs(x1.cpu().numpy() - want).max() < 1e-12
    s.close(); s2.close()


def test_hint_carried_across_ticks(solver_mod):
    """closed loop from a large offset: bounds are active for the first ticks, then release.  The hint that steers fast
    path vs interior-point iteration is carried on the device across ticks; both settings must produce the same
    controls (each method returns the unique QP minimiser), with the fast path taking over once the bounds release."""
    N, B, T = 40, 256, 40
    w = wl.tracking_batch(B, N, seed=23, pos_spread=3.0)
    sols = []
    for fp in (0, 1):
        s = solver_mod.BatchSolver(B, N)
        s.set_option("fast_path", fp)
        s.set_trajectory(w["traj"])
        s.set_iterate(w["X"], w["U"])
        sols.append(s)
    x = [w["x0"].copy(), w["x0"].copy()]
    lines = w["lines"].astype(np.int32)
    frac_fast = []
    for t in range(T):
        us = []
        for k, s in enumerate(sols):
            u0, _, st = s.solve_windowed(x[k], lines + t, w["p"])
            assert (st == 0).all(), (t, k, np.unique(st, return_counts=True))
            us.append(u0.copy())
        assert np.abs(us[0] - us[1]).max() < 1e-5, (t, np.abs(us[0] - us[1]).max())
        it, _ = sols[1].stats()
        frac_fast.append(float((it == 1).mean()))
        for k in range(2):
            x[k] = wl.plant_step(x[k], us[k], w["p"], 0.05)
    assert np.abs(x[0] - x[1]).max() < 1e-4
    assert frac_fast[0] < 0.9 and frac_fast[-1] > 0.9, frac_fast      # saturated at first, interior at the end
    for s in sols:
        s.close()


def test_ampc_filter_and_rls_match_oracle(solver_mod, oracle):
    """AMPC estimator chain (bluerov2_ampc_node.cpp:26-29: EKF -> RLSFF): the EKF with the AMPC model (ekf_model 1) and
    the RLS-VFF kernel against the oracle over 70 ticks (both error windows wrap), oracle filter state carried on both
    sides so that differences do not compound through the forward-difference Jacobians."""
    B, T = 192, 70
    rng = np.random.default_rng(11)
    s = solver_mod.BatchSolver(B, 10)
    s.set_option("ekf_model", 1)
    s.ekf_reset(); s.rls_reset()
    ox = np.zeros((B, 18)); oP = np.zeros((B, 18, 18))
    for i in range(B):
        ox[i], oP[i] = oracle.ekf_init()
    ost = oracle.rls_init(B)
    assert np.array_equal(s.rls_state(), ost)
    oracle.ekf_set_model(1)
    try:
        for t in range(T):
            thr = rng.uniform(-10, 10, (B, 6)); acc = rng.uniform(-1, 1, (B, 6)) * (0.02 if t > 40 else 1.0)
            meas = np.tile(np.array([0, 0, -20, 0, 0, 0, 0, 0, 0, 0, 0, 0.0]), (B, 1)) + rng.uniform(-0.3, 0.3, (B, 12))
            s.ekf(thr, meas, acc)
            oracle.ekf_step_batch(ox, oP, thr, meas, acc)
            ex, eP = s.ekf_state()
            assert (np.abs(ex - ox) / np.maximum(1.0, np.abs(ox))).max() < 1e-6, t
            s.set_ekf_state(ox, oP)
            p = s.rls(meas, acc, compensate=True)
            po = oracle.rls_step_batch(ost, ox, acc, meas, compensate=True)
            st = s.rls_state()
            # same arithmetic in the same order on both sides (FMA contraction aside); the lambda update is a threshold
            # decision, so it must agree exactly
            assert np.array_equal(st[:, :, 20], ost[:, :, 20]), t
            assert np.array_equal(st[:, :, 22:24], ost[:, :, 22:24])
            assert np.allclose(st[:, :, :20], ost[:, :, :20], rtol=1e-9, atol=1e-10), t
            assert np.allclose(st[:, :, 24:79], ost[:, :, 24:79], rtol=1e-9, atol=1e-10), t
            f_ok = np.isclose(st[:, :, 21], ost[:, :, 21], rtol=1e-8) | (np.isnan(st[:, :, 21]) & np.isnan(ost[:, :, 21]))
            assert f_ok.all(), t
            assert np.allclose(p, po, rtol=1e-9, atol=1e-9), t
            s.set_rls_state(ost)
    finally:
        oracle.ekf_set_model(0)
    lam = ost[:, :, 20]
    assert lam.min() < 0.9 < lam.max()          # the forgetting factor moved both ways over the run
    # compensate off: only p[0..3] written (reference brace placement, bluerov2_ampc.cpp:345-379)
    p_in = np.full((B, 16), 3.0)
    p = s.rls(meas, acc, compensate=False, out=p_in.copy())
    assert np.array_equal(p[:, :4], np.zeros((B, 4))) and np.array_equal(p[:, 4:], p_in[:, 4:])
    s.close()


def test_ampc_tick_on_device_equals_host_chain(solver_mod, oracle):
    """One AMPC tick as the node runs it (EKF -> RLSFF -> solve) with device pointers on one stream equals the same tick
    through the host entry points, and the solve agrees with the oracle fed the oracle's parameters."""
    import torch
    dev = torch.device("cuda", 0)
    B, N = 64, 20
    w = wl.tracking_batch(B, N, seed=3, pos_spread=0.3)
    rng = np.random.default_rng(4)
    thr = rng.uniform(-5, 5, (B, 6)); acc = rng.uniform(-0.5, 0.5, (B, 6))
    meas = w["x0"].copy()
    outs = []
    for use_dev in (False, True):
        s = solver_mod.BatchSolver(B, N)
        s.set_option("ekf_model", 1)
        s.set_iterate(w["X"], w["U"])
        if use_dev:
            d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)   # noqa: E731
            s.ekf(d(thr), d(meas), d(acc))
            p = s.rls(d(meas), d(acc), compensate=True)
            u0, th, st = s.solve(d(w["x0"]), d(w["yref"]), p)
            torch.cuda.synchronize()
            outs.append((u0.cpu().numpy(), th.cpu().numpy(), st.cpu().numpy(), p.cpu().numpy()))
        else:
            s.ekf(thr, meas, acc)
            p = s.rls(meas, acc, compensate=True)
            u0, th, st = s.solve(w["x0"], w["yref"], p)
            outs.append((u0, th, st, p))
        s.close()
    for a, b in zip(outs[0], outs[1]):
        assert np.array_equal(a, b)
    u0, th, st, p = outs[0]
    assert (st == 0).all()
    X, U = w["X"].copy(), w["U"].copy()
    so, _, _ = oracle.rti_step_batch(wl.time_steps(N), w["x0"], w["yref"], p, X, U)
    assert (so == 0).all() and np.abs(u0 - U[:, 0]).max() < TOL_U


def test_yaw_unwrap_matches_oracle_bit_for_bit(solver_mod, oracle):
    """node glue ahead of the solve (bluerov2_dob.cpp:272-304): float accumulators, only adds / compares / conversions ->
    the kernel must agree with the C restatement exactly, through the host and the device entry point"""
    import torch
    B, T = 300, 120
    rng = np.random.default_rng(9)
    true = rng.uniform(-0.5, 0.5, B); rate = rng.uniform(-0.7, 0.7, B)
    s = solver_mod.BatchSolver(B, 10)
    st = np.zeros((B, 2), dtype=np.float32)
    assert np.array_equal(s.yaw_state(), st)
    dev = torch.device("cuda", 0)
    for t in range(T):
        true = true + rate + rng.normal(0, 0.05, B)
        psi = (true + np.pi) % (2 * np.pi) - np.pi
        x0 = rng.uniform(-1, 1, (B, 12)); x0[:, 5] = psi
        want = oracle.yaw_unwrap_batch(st, psi)
        if t % 2 == 0:
            got = s.unwrap_yaw(x0.copy())
        else:
            d = torch.from_numpy(x0).to(dev)
            s.unwrap_yaw(d); torch.cuda.synchronize()
            got = d.cpu().numpy()
        assert np.array_equal(got[:, 5], want), t
        assert np.array_equal(np.delete(got, 5, axis=1), np.delete(x0, 5, axis=1))      # other columns untouched
        assert np.array_equal(s.yaw_state(), st), t
    assert np.abs(true).max() > 4 * np.pi
    s.yaw_reset()
    assert np.array_equal(s.yaw_state(), np.zeros((B, 2), dtype=np.float32))
    s.close()


def test_active_set_fast_path_equals_ipm(solver_mod, oracle):
    """Option active_set_path: instances whose previous solution had active bounds are solved by the pinned LQR + KKT check
    (exact whenever it accepts) instead of interior-point iterations.  Closed loop from a 3 m start (saturated thrusters for
    the first ticks): every tick against the interior-point-only solver on the same inputs, and against the oracle."""
    N, B, T = 40, 512, 10
    w = wl.tracking_batch(B, N, seed=21, pos_spread=3.0)
    Ts = wl.time_steps(N)
    s_as = solver_mod.BatchSolver(B, N); s_as.set_option("active_set_path", 1)
    s_ip = solver_mod.BatchSolver(B, N); s_ip.set_option("fast_path", 0); s_ip.set_option("active_set_path", 0)
    X, U = w["X"].copy(), w["U"].copy()
    x0, lines = w["x0"].copy(), w["lines"].copy()
    s_as.set_iterate(X, U); s_ip.set_iterate(X, U)
    took_as = took_sat = 0
    for t in range(T):
        yref = traj.window_batch(w["traj"], lines, N)
        u_a, _, st_a = s_as.solve(x0, yref, w["p"])
        u_i, _, st_i = s_ip.solve(x0, yref, w["p"])
        assert (st_a == 0).all() and (st_i == 0).all(), (t, np.unique(st_a), np.unique(st_i))
        Xa, Ua = s_as.get_iterate(); Xi, Ui = s_ip.get_iterate()
        assert np.abs(Ua - Ui).max() < 1e-5 and np.abs(Xa - Xi).max() < 1e-5, (t, np.abs(Ua - Ui).max())   # iterates carried separately
        assert (np.abs(Ua) <= 50 + 1e-9).all()
        it_a, _ = s_as.stats(); it_i, _ = s_ip.stats()
        sat = (np.abs(np.abs(Ui) - 50.0) < 1e-6).any(axis=(1, 2))
        if t >= 1:      # from the second tick on the previous solution's active set is the guess
            took_sat += int(sat.sum()); took_as += int((sat & (it_a <= 3) & (it_i > 3)).sum())
        if t == 2:
            Xo, Uo = X[:64].copy(), U[:64].copy()       # the interior-point solver's iterate before this tick
            so, _, _ = oracle.rti_step_batch(Ts, x0[:64], yref[:64], w["p"][:64], Xo, Uo)
            assert (so == 0).all() and np.abs(Uo - Ui[:64]).max() < TOL_U and np.abs(Uo - Ua[:64]).max() < 1e-5
        for i in range(B):
            x0[i] = oracle.erk4(x0[i], Ui[i, 0], w["p"][i], 0.05)
        lines = lines + 1
        X, U = Xi, Ui                # (each solver carries its own iterate: set_iterate would clear the active-set history)
    s_as.close(); s_ip.close()
    assert took_sat > 0 and took_as > 0.5 * took_sat, (took_as, took_sat)     # the guess was accepted on most saturated instances


def test_plant_wrench_replay_matches_host(solver_mod):
    """applyBodyWrench mode 2 (bluerov2_dob.cpp:818-874): the device plant replaying the reference's wrench series (fixture
    tests/golden/wrench_table.npz) row by row, per-instance start rows, clamped at the end, against the numpy plant"""
    import torch
    dev = torch.device("cuda", 0)
    B = 300
    table = wl.wrench_table()
    assert table.shape == (496, 4)
    rng = np.random.default_rng(3)
    phase = rng.integers(0, 520, B).astype(np.int32)          # some instances run off the end of the series
    w = wl.tracking_batch(B, 10, seed=2)
    x = w["x0"].copy()
    u = rng.uniform(-20, 20, (B, 4))
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)   # noqa: E731
    dx, du, dp, dt_, dph = d(x), d(u), d(w["p"]), d(table), d(phase)
    acc = torch.zeros((B, 6), dtype=torch.float64, device=dev)
    for tick in (0, 1, 7, 495, 600):
        solver_mod.plant_step_replay(dx, du, dp, dt_, phase=dph, tick=tick, h=0.05, body_acc=acc)
        dist = wl.wrench_at(table, tick, phase)
        assert np.array_equal(dist[phase + tick >= 495], np.tile(table[-1], (int((phase + tick >= 495).sum()), 1)))
        xn = wl.plant_step(x, u, w["p"], 0.05, dist=dist)
        assert np.abs(dx.cpu().numpy() - xn).max() < 1e-10, tick
        assert np.abs(acc.cpu().numpy() - (xn[:, 6:] - x[:, 6:]) / 0.05).max() < 1e-7
        x = xn
        dx.copy_(d(x))


def test_consolidated_node_fill_matches_oracle(solver_mod, oracle):
    """BLUEROV2_CTRL (src/ctrller/mpc.cpp:139-187, 199-262): DOMPC parameters (/disturbance divided by the hard-coded 0.0325...,
    p3 = 0) and a per-tick reference preview whose input columns stay 0, through the explicit-yref entry point, against the oracle
    fed the same fill; closed loop over a few ticks (the reference arrives per tick, not from a resident trajectory)"""
    N, B, T = 20, 96, 5
    Ts = wl.time_steps(N)
    w = wl.tracking_batch(B, N, seed=21, pos_spread=1.5)
    rng = np.random.default_rng(5)
    s = solver_mod.BatchSolver(B, N); s.set_iterate(w["X"], w["U"])
    X, U = w["X"].copy(), w["U"].copy()
    x0, lines = w["x0"].copy(), w["lines"].copy()
    for t in range(T):
        disturb = rng.uniform(-0.3, 0.3, (B, 3))                       # the /disturbance message of this tick
        p = wl.ctrl_params(disturb, dompc=True)
        assert np.array_equal(p[:, 3], np.zeros(B)) and np.allclose(p[:, 0], disturb[:, 0] / 0.032546960744430276)
        preview = traj.window_batch(w["traj"], lines, N)[:, :, :12]    # the /ref_traj preview: 12 state references per stage
        yref = wl.ctrl_yref(preview)
        assert np.array_equal(yref[:, :, 12:], np.zeros((B, N + 1, 4)))
        u0, th, st = s.solve(x0, yref, p)
        so, _, _ = oracle.rti_step_batch(Ts, x0, yref, p, X, U)
        assert (st == 0).all() and (so == 0).all()
        assert np.abs(u0 - U[:, 0]).max() < TOL_U, t
        x0 = wl.plant_step(x0, U[:, 0].copy(), w["p"], 0.05)
        lines = lines + 1
    Xg, Ug = s.get_iterate()
    assert np.abs(Ug - U).max() < TOL_U and np.abs(Xg - X).max() < TOL_U
    # controller type MPC: p[0..3] = 0, nominal hydrodynamics
    assert np.array_equal(wl.ctrl_params(disturb, dompc=False), np.tile(wl.NOMINAL_P, (B, 1)))
    s.close()


def test_eskf_matches_oracle(solver_mod):
    """batched IMU error-state Kalman filter (csrc/eskf.cu <- bluerov2_states/src/Eskf.cpp:97-331) against the numpy restatement
    (oracle/eskf.py): IMU-rate predictions with a GPS-rate update every other step, attitudes up to large angles, default and
    custom noise parameters; state and covariance after every call"""
    from scipy.spatial.transform import Rotation
    from oracle import eskf as E
    B = 70
    rng = np.random.default_rng(4)
    for prm in ({}, dict(q_xi=0.01, r_th=0.05, b_a=(0.01, -0.02, 0.03), b_g=(1e-3, 2e-3, -1e-3))):
        f = solver_mod.BatchEskf(B, **prm)
        R0 = Rotation.from_euler("ZYX", rng.uniform([-3, -1.2, -3], [3, 1.2, 3], (B, 3))).as_matrix()
        p0, v0 = rng.uniform(-5, 5, (B, 3)), rng.uniform(-1, 1, (B, 3))
        st = np.concatenate([p0, v0, R0.reshape(B, 9), np.zeros((B, 3))], axis=1)
        f.set_state(st, np.zeros((B, 21, 21)))
        refs = [E.Eskf(p0[i], v0[i], R0[i], **prm) for i in range(B)]
        for step in range(8):
            imu = np.concatenate([rng.normal(size=(B, 3)) * 0.5 + [0, 0, 9.81], rng.normal(size=(B, 3)) * 0.3], axis=1)
            f.predict(imu)
            for i in range(B):
                refs[i].predict(imu[i])
            if step % 2 == 1:
                pm = np.stack([r.p for r in refs]) + rng.normal(size=(B, 3)) * 0.05
                vm = np.stack([r.v for r in refs]) + rng.normal(size=(B, 3)) * 0.05
                Rg = np.stack([r.R @ E.so3_exp(rng.normal(size=3) * 0.02) for r in refs])
                th = rng.uniform(-20, 20, (B, 6))
                xi, y = f.update(pm, vm, Rg, th, imu, Rg)
                xo = np.stack([refs[i].update(pm[i], vm[i], Rg[i], th[i], imu[i], Rg[i]) for i in range(B)])
                assert np.abs(xi - xo).max() < 1e-9 * max(1.0, np.abs(xo).max()), step
            s_gpu, P_gpu = f.get_state()
            s_ref = np.stack([np.concatenate([r.p, r.v, r.R.ravel(), r.xi]) for r in refs])
            P_ref = np.stack([r.P for r in refs])
            assert np.abs(s_gpu - s_ref).max() < 1e-9 * max(1.0, np.abs(s_ref).max()), step
            assert np.abs(P_gpu - P_ref).max() < 1e-10 * max(1.0, np.abs(P_ref).max()), step
        R = s_gpu[:, 6:15].reshape(B, 3, 3)
        assert np.abs(R @ R.transpose(0, 2, 1) - np.eye(3)).max() < 1e-12          # attitudes stay rotations
        f.close()


def test_ekf_full_size_batch(solver_mod, oracle):
    """the EKF at config 3's batch (4096 filters): one step from random states and a random SPD covariance, three draws; every instance
    against the oracle at 1e-6 relative, the covariance symmetric to rounding and positive on its diagonal -- the tensor-core products,
    the zero-padded 18 = 8 + 8 + 2 tiling and the structural-zero tile skipping see every lane pattern here.  (Each draw starts afresh:
    iterating a filter from RANDOM disturbance states and thrusts leaves the model's sane range within two steps, on any implementation.)"""
    B = 4096
    rng = np.random.default_rng(21)
    s = solver_mod.BatchSolver(B, 10)
    for draw in range(3):
        ex = np.zeros((B, 18)); eP = np.zeros((B, 18, 18))
        ex[:, :3] = rng.uniform(-2, 2, (B, 3)); ex[:, 3:5] = rng.uniform(-0.4, 0.4, (B, 2)); ex[:, 5] = rng.uniform(-3, 3, B)
        ex[:, 6:12] = rng.uniform(-0.6, 0.6, (B, 6)); ex[:, 12:] = rng.uniform(-5, 5, (B, 6))
        M = rng.uniform(-1, 1, (B, 18, 18))
        eP[:] = 0.05 * (M @ M.transpose(0, 2, 1)) / 18 + 1e-3 * np.eye(18)
        s.set_ekf_state(ex, eP)
        ox, oP = ex.copy(), eP.copy()
        thr = rng.uniform(-10, 10, (B, 6))
        meas = ox[:, :12] + rng.normal(0, 0.01, (B, 12))
        acc = rng.uniform(-0.3, 0.3, (B, 6))
        wf, p = s.ekf(thr, meas, acc)
        owf = oracle.ekf_step_batch(ox, oP, thr, meas, acc)
        x, P = s.ekf_state()
        assert np.isfinite(x).all() and np.isfinite(P).all()
        assert (np.abs(x - ox) / np.maximum(1.0, np.abs(ox))).max() < 1e-6, draw
        assert np.abs(P - oP).max() < 1e-6 * max(1.0, np.abs(oP).max()), draw
        assert np.abs(wf - owf).max() < 1e-6 * max(1.0, np.abs(owf).max()), draw
        assert np.abs(P - P.transpose(0, 2, 1)).max() < 1e-9 * max(1.0, np.abs(P).max())
        assert (np.einsum("bii->bi", P) > 0).all()
    s.close()
