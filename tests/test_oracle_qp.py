"""The NLP/QP layer of the oracle: solver-independent certificates (acados/HPIPM cannot run here)."""
import numpy as np
import pytest

from bluerov2_b200 import traj, workloads as wl
from tests.qp_dense import build_qp, condense, solve_box_qp, kkt_residuals


def _one_step(oracle, N, x0, yref, p, X, U):
    from oracle import W_DEFAULT, WE_DEFAULT, LBU, UBU
    Ts = wl.time_steps(N)
    A, B, b = oracle.linearize(Ts, p, X, U)
    qp = build_qp(A, B, b, Ts, W_DEFAULT, WE_DEFAULT, X, U, yref, x0, LBU, UBU)
    Xn, Un = X.copy(), U.copy()
    st, info = oracle.rti_step(Ts, x0, yref, p, Xn, Un)
    return qp, st, info, Xn, Un


@pytest.mark.parametrize("N,spread,seed", [(40, 0.5, 0), (40, 3.0, 0), (20, 3.0, 1), (10, 3.0, 2), (80, 2.0, 3)])
def test_rti_step_solves_the_condensed_qp(oracle, N, spread, seed):
    """du from the Riccati IPM == du from a dense active-set solve of the condensed QP (unique minimiser),
    and it satisfies the dense KKT conditions to 1e-8 (scaled)."""
    w = wl.tracking_batch(6, N, seed=seed, pos_spread=spread)
    n_active = 0
    for i in range(6):
        X, U = w["X"][i].copy(), w["U"][i].copy()
        x0 = w["x0"][i].copy()
        line = int(w["lines"][i])
        for tick in range(3):
            yref = traj.window(w["traj"], line + tick, N)
            qp, st, info, Xn, Un = _one_step(oracle, N, x0, yref, w["p"][i], X, U)
            assert st == 0
            H, g, G, c = condense(qp)
            lb, ub = qp["lb"].ravel(), qp["ub"].ravel()
            v_ref, _, _ = solve_box_qp(H, g, lb, ub)
            du = (Un - U).ravel()
            scale = max(1.0, np.abs(g).max())
            stat, feas = kkt_residuals(H, g, lb, ub, du)
            assert feas < 1e-9 and stat < 1e-8 * scale, (stat, feas, scale)
            assert np.abs(du - v_ref).max() < 1e-7, np.abs(du - v_ref).max()
            # states are the exact roll-out of du
            assert np.abs((Xn - X).ravel() - (c + G @ du)).max() < 1e-9
            n_active += int((np.abs(Un) > 50 - 1e-7).sum())
            x0 = oracle.erk4(x0, Un[0], w["p"][i], 0.05)
            X, U = Xn, Un
    if spread >= 3.0 and N == 40:
        assert n_active > 0, "active-bound set did not activate any bound"


def test_golden_rti_cases_reproduce(oracle, golden):
    """rti_cases.npz was generated with the ERK routed through the reference's CasADi VDE; the hand restatement
    must land on the same iterates (dynamics agree to 1e-15, QP solved to 1e-12)."""
    g = golden["rti_cases"]
    for tag in ("nom40", "act40", "lem40", "act20", "act80"):
        N = int(g[tag + "_N"])
        Ts = wl.time_steps(N)
        tr = traj.circle() if str(g[tag + "_ref"]) == "circle" else traj.lemniscate()
        X, U = g[tag + "_X0"].copy(), g[tag + "_U0"].copy()
        for t in range(g[tag + "_x0"].shape[0]):
            yref = traj.window_batch(tr, g[tag + "_lines"][t], N)
            st, info, _ = oracle.rti_step_batch(Ts, g[tag + "_x0"][t], yref, g[tag + "_p"], X, U)
            assert (st == 0).all()
            assert np.abs(U - g[tag + "_Uout"][t]).max() < 1e-7, (tag, t)
            assert np.abs(X - g[tag + "_Xout"][t]).max() < 1e-7, (tag, t)
            X, U = g[tag + "_Xout"][t].copy(), g[tag + "_Uout"][t].copy()


def test_cold_start_known_answers(oracle, golden):
    """SURVEY 8c known-answer u0 for the main_bluerov2.c-style cold solve on circle.txt rows 0..N
    (independently derived by the survey's throw-away numpy prototype)."""
    from oracle import NOMINAL_P, X_INIT
    want = {80: (-16.2663, -16.2821, 0.015788, 8.8851), 40: (-4.57709, -8.86062, 0.0141913, 2.31820),
            20: (-0.893938, -3.11109, 0.0121816, 0.633124)}
    x0 = golden["rti_cases"]["cold_x0"]
    circ = traj.circle()
    for N, u0 in want.items():
        X = np.tile(X_INIT, (N + 1, 1)).copy(); U = np.zeros((N, 4))
        st, info = oracle.rti_step(wl.time_steps(N), x0, circ[:N + 1], NOMINAL_P, X, U)
        assert st == 0
        assert np.allclose(U[0], u0, rtol=2e-5, atol=2e-6), (N, U[0])
        assert np.abs(U - golden["rti_cases"][f"cold_N{N}_U"]).max() < 1e-8


def test_per_stage_parameters(oracle):
    """bluerov2_acados_update_params is per stage (acados_solver_bluerov2.c:835-883): identical rows == shared p."""
    from oracle import NOMINAL_P
    N = 20
    w = wl.tracking_batch(1, N, seed=9)
    Xa, Ua = w["X"][0].copy(), w["U"][0].copy()
    Xb, Ub = Xa.copy(), Ua.copy()
    p = NOMINAL_P.copy(); p[:4] = [3, -2, 1, 0.5]
    oracle.rti_step(wl.time_steps(N), w["x0"][0], w["yref"][0], p, Xa, Ua)
    oracle.rti_step(wl.time_steps(N), w["x0"][0], w["yref"][0], np.tile(p, (N + 1, 1)), Xb, Ub)
    assert np.array_equal(Ua, Ub)
    pp = np.tile(p, (N + 1, 1)); pp[5:, 0] = -7.0
    Xc, Uc = w["X"][0].copy(), w["U"][0].copy()
    oracle.rti_step(wl.time_steps(N), w["x0"][0], w["yref"][0], pp, Xc, Uc)
    assert np.abs(Uc - Ua).max() > 1e-3


def test_thrust_allocation(oracle):
    u = np.array([1.0, -2.0, 3.0, 0.5]); rc = 0.026546960744430276
    want = np.array([(-1 - 2 + 0.5), (-1 + 2 - 0.5), (1 - 2 - 0.5), (1 + 2 + 0.5), -3, -3]) / rc
    assert np.allclose(oracle.thrust_alloc(u), want, rtol=1e-15)


@pytest.mark.parametrize("N,spread,seed", [(40, 0.5, 0), (40, 3.0, 0), (20, 3.0, 1), (10, 3.0, 2)])
def test_rti_step_matches_scipy_bvls(oracle, N, spread, seed):
    """Third-party check of the QP layer.  The Gauss-Newton QP of an RTI step is a bounded linear least-squares problem:
    with H = L L' (Cholesky of the condensed Hessian), 1/2 du'H du + g'du = 1/2 |L'du + L^-1 g|^2 + const.  scipy's
    lsq_linear(method="bvls") (Stark & Parker's bounded-variable least squares, an active-set method that shares no code and no
    algorithm with the oracle's Riccati interior-point iteration) must land on the same du to 1e-7."""
    from scipy.linalg import cholesky, solve_triangular
    from scipy.optimize import lsq_linear
    w = wl.tracking_batch(6, N, seed=seed, pos_spread=spread)
    n_active = 0
    for i in range(6):
        X, U = w["X"][i].copy(), w["U"][i].copy()
        x0 = w["x0"][i].copy()
        line = int(w["lines"][i])
        for tick in range(3):
            yref = traj.window(w["traj"], line + tick, N)
            qp, st, info, Xn, Un = _one_step(oracle, N, x0, yref, w["p"][i], X, U)
            assert st == 0
            H, g, G, c = condense(qp)
            L = cholesky(H, lower=True)
            res = lsq_linear(L.T, -solve_triangular(L, g, lower=True), bounds=(qp["lb"].ravel(), qp["ub"].ravel()),
                             method="bvls", tol=1e-15, max_iter=2000)
            assert res.status > 0, res.message
            du = (Un - U).ravel()
            assert np.abs(du - res.x).max() < 1e-7, (i, tick, np.abs(du - res.x).max())
            n_active += int(res.active_mask.astype(bool).sum())
            x0 = oracle.erk4(x0, Un[0], w["p"][i], 0.05)
            X, U = Xn, Un
    if spread >= 3.0 and N == 40:
        assert n_active > 0, "active-bound set did not activate any bound"


@pytest.mark.parametrize("N,spread,seed", [(40, 0.5, 0), (40, 3.0, 1), (10, 3.0, 2), (80, 1.0, 3)])
def test_dense_condensed_variant_matches_riccati(oracle, N, spread, seed):
    """orc_set_qp_mode(1): the same RTI step with the QP solved by full condensing + a dense interior-point iteration (the cost profile of
    the reference's FULL_CONDENSING_HPIPM; bench.py's cpu_baseline.dense_condensed) lands on the iterate of the Riccati variant"""
    B = 6
    w = wl.tracking_batch(B, N, seed=seed, pos_spread=spread)
    Ts = wl.time_steps(N)
    yref = traj.window_batch(w["traj"], w["lines"], N)
    Xa, Ua = w["X"].copy(), w["U"].copy()
    Xb, Ub = w["X"].copy(), w["U"].copy()
    try:
        oracle.set_qp_mode(0)
        sa, ia, _ = oracle.rti_step_batch(Ts, w["x0"], yref, w["p"], Xa, Ua, nthreads=1)
        oracle.set_qp_mode(1)
        sb, ib, _ = oracle.rti_step_batch(Ts, w["x0"], yref, w["p"], Xb, Ub, nthreads=1)
    finally:
        oracle.set_qp_mode(0)
    assert (sa == 0).all() and (sb == 0).all()
    assert np.abs(Ua - Ub).max() < 1e-9 and np.abs(Xa - Xb).max() < 1e-9
    assert np.array_equal(ia[:, 0], ib[:, 0])            # the same iteration, solve for solve
