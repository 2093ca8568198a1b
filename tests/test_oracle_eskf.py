"""The ESKF restatement (oracle/eskf.py <- bluerov2_states/src/Eskf.cpp:97-331): what can be pinned without Sophus / Eigen / ROS.
SO(3) exp / log against scipy's Rotation; the filter algebra against an independent dense formulation; structure of F and H."""
import numpy as np
from scipy.spatial.transform import Rotation

from oracle import eskf as E


def test_so3_exp_log_against_scipy():
    rng = np.random.default_rng(0)
    for scale in (1e-9, 1e-5, 1e-3, 0.3, 2.0, 3.1):
        for _ in range(25):
            w = rng.normal(size=3); w *= scale / np.linalg.norm(w) * rng.uniform(0.2, 1.0)
            R = E.so3_exp(w)
            assert np.abs(R - Rotation.from_rotvec(w).as_matrix()).max() < 1e-14
            assert np.abs(R @ R.T - np.eye(3)).max() < 1e-14
            assert np.abs(E.so3_log(R) - w).max() < 1e-12 * max(1.0, scale)
            assert np.abs(E.so3_log(R) - Rotation.from_matrix(R).as_rotvec()).max() < 1e-9
    # all four branches of the quaternion extraction (rotations by ~pi about each axis)
    for ax in np.eye(3):
        w = ax * 3.1 + 0.01
        assert np.abs(E.so3_log(E.so3_exp(w)) - w).max() < 1e-10


def test_rpy_matches_scipy_zyx():
    rng = np.random.default_rng(1)
    for _ in range(50):
        R = Rotation.from_euler("ZYX", rng.uniform([-3, -1.4, -3], [3, 1.4, 3])).as_matrix()
        yaw, pitch, roll = Rotation.from_matrix(R).as_euler("ZYX")
        assert np.allclose(E.rpy_of(R), [roll, pitch, yaw], atol=1e-12)


def test_transition_and_measurement_structure():
    """set_F / set_H (Eskf.cpp:143-158, 303-313): which blocks are filled, and with what"""
    f = E.Eskf([0, 0, -20], [0.3, 0, 0], Rotation.from_euler("ZYX", [0.4, 0.1, -0.2]).as_matrix())
    imu = np.array([0.1, -0.2, 9.7, 0.02, -0.01, 0.3])
    F = f.F(imu, E.DT)
    mask = np.eye(21, dtype=bool)
    for r, c in ((0, 3), (3, 6), (3, 12), (3, 15), (6, 6), (6, 9)):
        mask[r:r + 3, c:c + 3] = True
    assert np.all(F[~mask] == 0.0)
    assert np.array_equal(F[0:3, 3:6], np.eye(3) * E.DT) and np.array_equal(F[3:6, 15:18], np.eye(3) * E.DT)
    assert np.array_equal(F[6:9, 9:12], -np.eye(3) * E.DT) and np.allclose(F[3:6, 12:15], -f.R * E.DT)
    assert np.allclose(F[6:9, 6:9], Rotation.from_rotvec(-(imu[3:] - f.b_g) * E.DT).as_matrix(), atol=1e-15)
    assert np.allclose(F[3:6, 6:9], -f.R @ E.hat(imu[:3] - f.b_a) * E.DT)
    assert np.array_equal(F[9:, 9:], np.eye(12))


def test_update_is_the_kalman_update_with_the_reference_injection():
    """one predict + update against a dense re-derivation: K = P H'(H P H' + R)^-1 through a linear solve instead of an
    inverse, P+ = (I - K H) P, and the injection quirks (velocity twice, biases / g untouched)"""
    rng = np.random.default_rng(2)
    R0 = Rotation.from_euler("ZYX", [0.7, -0.05, 0.1]).as_matrix()
    f = E.Eskf([1, 2, -20], [0.2, -0.1, 0.05], R0)
    for _ in range(5):
        f.predict(np.array([0.0, 0.0, 9.81, 0, 0, 0]) + rng.normal(size=6) * 0.05)
    P, p, v, R, xi = f.P.copy(), f.p.copy(), f.v.copy(), f.R.copy(), f.xi.copy()
    assert np.abs(P - P.T).max() < 1e-15 and np.linalg.eigvalsh(P).min() > 0
    meas = (p + rng.normal(size=3) * 0.05, v + rng.normal(size=3) * 0.05, R @ E.so3_exp(rng.normal(size=3) * 0.01),
            rng.uniform(-10, 10, 6), np.array([0.0, 0.0, 9.81, 0, 0, 0]) + rng.normal(size=6) * 0.05, R0)
    y = f.innovation(*meas)
    H = np.zeros((12, 21)); H[0:9, 0:9] = np.eye(9); H[9:12, 18:21] = -np.eye(3)
    S = H @ P @ H.T + f.Rm
    Kg = np.linalg.solve(S.T, (P @ H.T).T).T
    dx = Kg @ y
    xw = f.update(*meas)
    assert np.allclose(f.P, (np.eye(21) - Kg @ H) @ P, atol=1e-14)
    assert np.allclose(f.p, p + dx[:3]) and np.allclose(f.v, v + 2 * dx[3:6]) and np.allclose(f.xi, xi + dx[18:])
    assert np.allclose(f.R, R @ Rotation.from_rotvec(dx[6:9]).as_matrix(), atol=1e-14) and np.allclose(xw, f.R @ f.xi)
    assert np.array_equal(f.b_a, np.array(E.DEFAULTS["b_a"])) and np.array_equal(f.g, [0, 0, -E.G])


def test_filter_finds_a_constant_disturbance():
    """closed form sanity: a vehicle at rest, level, held by its thrusters against a constant body force; the disturbance state
    converges to that force (sign convention of the thrust residual, Eskf.cpp:229-272)"""
    f = E.Eskf([0, 0, -20], [0, 0, 0], np.eye(3))
    d = np.array([4.0, -3.0, 2.0])                       # true disturbance force, body frame
    wb = E.MASS * E.G - E.BUOY
    # at rest with R = I: specific force (0, 0, g) after bias removal; M_rb a - xi + M_a (a + g_B) + D v + g(rpy) = tau
    acc = np.array([0.0, 0.0, E.G])
    imu = np.concatenate([acc + f.b_a, f.b_g])
    tau = np.array([E.MASS * 0, 0, E.MASS * E.G]) - d + E.ADDED[:3] * (acc + np.array([0, 0, -E.G])) + np.array([0, 0, -wb])
    th = np.linalg.lstsq(E.K_ALLOC, tau, rcond=None)[0]
    for _ in range(400):
        f.predict(imu)
        f.update(np.array([0, 0, -20.0]), np.zeros(3), np.eye(3), th, imu, np.eye(3))
    assert np.abs(f.xi - d).max() < 0.05, f.xi
