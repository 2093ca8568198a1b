"""Oracle vs the reference's own CasADi-generated C (dynamics / forward VDE): the pinned layer."""
import numpy as np


def test_restatement_matches_casadi_golden(oracle, golden):
    """tests/golden/casadi_vde.npz holds outputs of bluerov2_expl_vde_forw / _ode_fun (reference code)."""
    g = golden["casadi_vde"]
    n = g["x"].shape[0]
    worst = 0.0
    for i in range(n):
        f, dSx, dSu = oracle.vde_forw_cm(g["x"][i], g["Sx_cm"][i], g["Su_cm"][i], g["u"][i], g["p"][i])
        sc = max(1.0, np.abs(g["dSx_cm"][i]).max(), np.abs(g["dSu_cm"][i]).max(), np.abs(g["f"][i]).max())
        worst = max(worst, np.abs(f - g["f"][i]).max() / sc, np.abs(dSx - g["dSx_cm"][i]).max() / sc,
                    np.abs(dSu - g["dSu_cm"][i]).max() / sc)
        assert np.abs(oracle.ode(g["x"][i], g["u"][i], g["p"][i]) - g["f_ode"][i]).max() <= 1e-12 * sc
    assert worst < 1e-13, worst


def test_known_answer_ode(oracle):
    """SURVEY 8c known-answer vector (reference CasADi output)."""
    from oracle import NOMINAL_P
    x = np.array([0.1, -0.2, -20, 0.05, -0.03, 0.7, 0.5, -0.4, 0.2, 0.01, -0.02, 0.3])
    u = np.array([3, -2, 1.5, 0.4])
    p = NOMINAL_P.copy(); p[:4] = [1, -0.5, 0.25, 0.1]
    want = np.array([0.64192955796202422, 0.0052905517699709578, 0.1946752492656664, 0.0013951966728900488,
                     -0.0049812544266958289, 0.29875992661720896, -25.348612237931547, -17.944745367846586,
                     -7.1707058817155147, -0.36888299222655957, 0.10651812538627911, 10.169530481708806])
    assert np.abs(oracle.ode(x, u, p) - want).max() < 1e-13


def test_adjoint_golden_consistent_with_jacobian(oracle, golden):
    """bluerov2_expl_vde_adj returns [Jx' lam; Ju' lam]: cross-checks the analytic Jacobian."""
    g = golden["casadi_vde"]
    for i in range(0, g["x"].shape[0], 7):
        Jx, Ju = oracle.jac(g["x"][i], g["u"][i], g["p"][i])
        adj = np.concatenate([Jx.T @ g["lam"][i], Ju.T @ g["lam"][i]])
        assert np.abs(adj - g["adj"][i]).max() < 1e-11 * max(1, np.abs(g["adj"][i]).max())


def test_jacobian_sparsity(oracle, golden):
    """48 structural non-zeros, columns 0..2 identically zero; 5 constant non-zeros in Ju."""
    g = golden["casadi_vde"]
    pat = {0: {3, 4, 5, 6, 7, 8}, 1: {3, 4, 5, 6, 7, 8}, 2: {3, 4, 6, 7, 8}, 3: {3, 4, 5, 9, 10, 11}, 4: {3, 10, 11},
           5: {3, 4, 10, 11}, 6: {4, 6}, 7: {3, 4, 7}, 8: {3, 4, 8}, 9: {3, 4, 10, 11}, 10: {4, 9, 11}, 11: {9, 10, 11}}
    nz = np.zeros((12, 12), bool)
    for i in range(20, 60):
        Jx, Ju = oracle.jac(g["x"][i], g["u"][i], g["p"][i])
        nz |= Jx != 0
        assert (Ju != 0).sum() == 5
    for r in range(12):
        assert set(np.nonzero(nz[r])[0]) == pat[r], r
    from oracle import NOMINAL_P
    _, Ju = oracle.jac(g["x"][0], g["u"][0], NOMINAL_P)
    assert np.allclose([Ju[6, 0], Ju[7, 1], Ju[8, 2], Ju[11, 1], Ju[11, 3]], [-8.208, 9.461, -4.504, -0.6146, 26.275], rtol=2e-4)


def test_live_casadi_when_reference_present(oracle, casadi_ref):
    from oracle import NOMINAL_P
    rng = np.random.default_rng(5)
    for _ in range(200):
        x = rng.uniform(-1, 1, 12) * np.array([5, 5, 5, 1, 1, 3, 2, 2, 2, 1, 1, 1]); x[2] -= 20
        u = rng.uniform(-50, 50, 4)
        p = NOMINAL_P.copy(); p[:4] = rng.uniform(-20, 20, 4)
        Sx, Su = rng.standard_normal(144), rng.standard_normal(48)
        a = casadi_ref.vde_forw(x, Sx, Su, u, p)
        b = oracle.vde_forw_cm(x, Sx, Su, u, p)
        sc = max(1.0, np.abs(a[1]).max(), np.abs(a[2]).max())
        for q, r in zip(a, b):
            assert np.abs(q - r).max() < 1e-13 * sc
    # cost functions: y = [x; u], y_e = x (bluerov2.py:144,153-154)
    assert np.array_equal(casadi_ref.cost_y(x, u, p), np.concatenate([x, u]))
    assert np.array_equal(casadi_ref.cost_y_e(x, p), x)


def test_erk4_sens_restatement_vs_casadi_route(oracle, casadi_ref):
    from oracle import NOMINAL_P
    rng = np.random.default_rng(6)
    for _ in range(50):
        x = rng.uniform(-1, 1, 12) * np.array([5, 5, 5, 0.8, 0.8, 3, 2, 2, 2, 1, 1, 1]); x[2] -= 20
        u = rng.uniform(-50, 50, 4)
        oracle.use_casadi(None)
        a = oracle.erk4_sens(x, u, NOMINAL_P, 0.025)
        oracle.use_casadi(casadi_ref)
        b = oracle.erk4_sens(x, u, NOMINAL_P, 0.025)
        oracle.use_casadi(None)
        for q, r in zip(a, b):
            assert np.abs(q - r).max() < 1e-13 * max(1, np.abs(r).max())
        # positions do not enter f: A[:, 0:3] = [I; 0] exactly
        assert np.array_equal(a[1][:, :3], np.eye(12)[:, :3])
        # finite-difference check of the sensitivities
        eps = 1e-6
        for j in (3, 5, 7, 11):
            xp = x.copy(); xp[j] += eps
            xm = x.copy(); xm[j] -= eps
            fd = (oracle.erk4(xp, u, NOMINAL_P, 0.025) - oracle.erk4(xm, u, NOMINAL_P, 0.025)) / (2 * eps)
            assert np.abs(fd - a[1][:, j]).max() < 1e-7
        assert np.abs(oracle.erk4(x, u, NOMINAL_P, 0.025) - a[0]).max() < 1e-14
