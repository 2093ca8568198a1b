"""EKF restatement (bluerov2_dob.cpp:495-752): structural checks + golden reproduction."""
import numpy as np


def test_init(oracle):
    x, P = oracle.ekf_init()
    assert np.array_equal(x, [0, 0, -20, 0, 0, 0, 0, 0, 0, 0, 0, 0, 6, 6, 6, 0, 0, 0])   # bluerov2_dob.cpp:64
    assert np.array_equal(P, np.eye(18))


def test_process_model_against_numpy(oracle):
    """f(): independent numpy transcription of the kinetics incl. Coriolis and invM(i,i) of the coupled M."""
    rng = np.random.default_rng(0)
    m, Ix, Iy, Iz, ZG, g, buoy = 11.26, 0.3, 0.63, 0.58, 0.02, 9.81, 0.661618
    am = [1.7182, 0, 5.468, 0, 1.2481, 0.4006]
    M = np.diag([m + am[0], m + am[1], m + am[2], Ix + am[3], Iy + am[4], Iz + am[5]])
    M[0, 4] = m * ZG; M[1, 3] = -m * ZG; M[3, 1] = -m * ZG; M[4, 0] = m * ZG
    iM = np.diag(np.linalg.inv(M))
    Dl = [-11.7391, -20, -31.8678, -25, -44.9085, -5]; Dnl = [-18.18, -21.66, -36.99, -1.55, -1.55, -1.55]
    K = np.array([[0.7071067811847433, 0.7071067811847433, -0.7071067811919605, -0.7071067811919605, 0.0, 0.0],
                  [0.7071067811883519, -0.7071067811883519, 0.7071067811811348, -0.7071067811811348, 0.0, 0.0],
                  [0, 0, 0, 0, 1, 1],
                  [0.051265241636155506, -0.05126524163615552, 0.05126524163563227, -0.05126524163563227, -0.11050000000000001, 0.11050000000000003],
                  [-0.05126524163589389, -0.051265241635893896, 0.05126524163641713, 0.05126524163641713, -0.002499999999974481, -0.002499999999974481],
                  [0.16652364696949604, -0.16652364696949604, -0.17500892834341342, 0.17500892834341342, 0.0, 0.0]])
    for _ in range(20):
        x = rng.uniform(-1, 1, 18); u = rng.uniform(-10, 10, 6)
        f = oracle.ekf_f(x, u)
        tau = K @ u
        s3, c3, s4, c4 = np.sin(x[3]), np.cos(x[3]), np.sin(x[4]), np.cos(x[4])
        want6 = iM[0] * (tau[0] + m * x[11] * x[7] - m * x[10] * x[8] - buoy * s4 + x[12] + Dl[0] * x[6] + Dnl[0] * abs(x[6]) * x[6])
        want11 = iM[5] * (tau[5] - (Iy - Ix) * x[9] * x[10] + x[17] + Dl[5] * x[11] + Dnl[5] * abs(x[11]) * x[11])
        want3 = x[9] + np.sin(x[5]) * s4 / c4 * x[10] + c3 * s4 / c4 * x[11]          # sin(psi) quirk (:646)
        assert abs(f[6] - want6) < 1e-12 and abs(f[11] - want11) < 1e-12 and abs(f[3] - want3) < 1e-13
        assert np.all(f[12:] == 0)
        y = oracle.ekf_h(x, u)      # any 6-vector as body_acc
        assert np.array_equal(y[:12], x[:12])
        want12 = M[0, 0] * u[0] - m * x[11] * x[7] + m * x[10] * x[8] + buoy * s4 - x[12] - Dl[0] * x[6] - Dnl[0] * abs(x[6]) * x[6]
        assert abs(y[12] - want12) < 1e-12


def test_golden_reproduces_and_covariance_stays_symmetric(oracle, golden):
    g = golden["ekf_cases"]
    T, B = g["thr"].shape[:2]
    ex = np.zeros((B, 18)); eP = np.zeros((B, 18, 18))
    for i in range(B):
        ex[i], eP[i] = oracle.ekf_init()
    for t in range(T):
        wf = oracle.ekf_step_batch(ex, eP, g["thr"][t], g["meas"][t], g["acc"][t])
        assert np.abs(ex - g["ex"][t]).max() < 1e-9
        assert np.abs(eP - g["eP"][t]).max() < 1e-9
        assert np.abs(wf - g["wf"][t]).max() < 1e-9
        assert np.abs(eP - np.swapaxes(eP, 1, 2)).max() < 1e-9       # Joseph form
        assert (np.linalg.eigvalsh(0.5 * (eP + np.swapaxes(eP, 1, 2))) > -1e-12).all()


def test_batch_equals_single(oracle, golden):
    g = golden["ekf_cases"]
    x, P = oracle.ekf_init()
    wf = oracle.ekf_step(x, P, g["thr"][0, 3], g["meas"][0, 3], g["acc"][0, 3])
    assert np.array_equal(x, g["ex"][0, 3]) and np.array_equal(wf, g["wf"][0, 3])


def test_full_step_against_independent_numpy_restatement(oracle, golden):
    """The whole predict / gain / Joseph step of orc_ekf_step against tests/ekf_numpy.py, a restatement written from
    bluerov2_dob.cpp alone with LAPACK's LU inverse standing in for Eigen's.  Several consecutive steps on the golden inputs:
    the finite-difference Jacobians (d = 1e-6) amplify the last-bit differences between the two summation orders by 1e6, and
    the gain multiplies them by |S^-1| ~ 1e5; agreement is asserted at 1e-6 relative to the state / covariance scale."""
    from tests.ekf_numpy import ekf_step
    g = golden["ekf_cases"]
    T, B = g["thr"].shape[:2]
    worst = 0.0
    for i in range(0, B, 3):
        x, P = oracle.ekf_init()
        xn, Pn = x.copy(), P.copy()
        for t in range(T):
            wf = oracle.ekf_step(x, P, g["thr"][t, i], g["meas"][t, i], g["acc"][t, i])
            # the numpy side restarts from the oracle's state every tick: one-step agreement, no drift compounding
            x2, P2, wf2 = ekf_step(xn, Pn, g["thr"][t, i], g["meas"][t, i], g["acc"][t, i])
            ex = np.abs(x2 - x).max() / max(1.0, np.abs(x).max())
            eP = np.abs(P2 - P).max() / max(1.0, np.abs(P).max())
            ew = np.abs(wf2 - wf).max() / max(1.0, np.abs(wf).max())
            worst = max(worst, ex, eP, ew)
            assert ex < 1e-6 and eP < 1e-6 and ew < 1e-6, (i, t, ex, eP, ew)
            xn, Pn = x.copy(), P.copy()
    print("worst relative one-step difference", worst)
