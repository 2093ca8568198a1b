"""CPU restatement (numpy) of the reference's IMU error-state Kalman filter, the alternative disturbance observer of the
consolidated node (bluerov2_states/src/Eskf.cpp:97-331, Dynamics.cpp:57-165, Config.cpp:82-161, launch/config/imudo.yaml).
TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.

PARITY UNPINNED against the reference itself: the filter is a ROS nodelet on Eigen + Sophus + tf (none of them vendored or
installed here), and the reference holds no fixture for it.  What pins this file instead: the SO(3) exp / log written out below
are checked against scipy.spatial.transform.Rotation (tests/test_oracle_eskf.py), the covariance algebra against numpy's own
linalg (np.linalg.inv is LAPACK's partial-pivot LU, the algorithm class of Eigen's .inverse() for a 12 x 12 matrix).

State.  Nominal: p (3), v (3), R (3 x 3 rotation, body -> inertial), xi (3, disturbance force, body frame); constants b_a, b_g
(accelerometer / gyro bias, launch/config/imudo.yaml), g = (0, 0, -9.81).  Error state, 21: [dp, dv, dtheta, db_g, db_a, dg, dxi]
-- the column order set_F / set_H use (Eskf.cpp:143-158, 303-313).  Quirks kept: the velocity correction is injected twice
(Eskf.cpp:320), the biases and g are never injected, the attitude measurement and the gravity direction inside dynamics_Ma come
from ground truth (Eskf.cpp:170-176, Dynamics.cpp:180-181) -- here they are inputs (R_meas, R_gt).
"""
from __future__ import annotations

import numpy as np

DT = 1.0 / 50.0                       # Eskf.cpp:103
MASS, IX, IY, IZ, ZG, G, BUOY = 11.26, 0.3, 0.63, 0.58, 0.02, 9.81, 0.661618          # ImuDo.h:101-110
ADDED = np.array([1.7182, 0, 5.468, 0, 1.2481, 0.4006])
DL = np.array([-11.7391, -20, -31.8678, -25, -44.9085, -5])
DNL = np.array([-18.18, -21.66, -36.99, -1.55, -1.55, -1.55])
K_ALLOC = np.array([[0.7071067811847433, 0.7071067811847433, -0.7071067811919605, -0.7071067811919605, 0.0, 0.0],
                    [0.7071067811883519, -0.7071067811883519, 0.7071067811811348, -0.7071067811811348, 0.0, 0.0],
                    [0, 0, 0, 0, 1, 1]])                                                  # rows 0..2 of K (Config.cpp:167-172)
# launch/config/imudo.yaml
DEFAULTS = dict(q_p=0.001, q_v=0.001, q_r=0.001, q_q=0.001, q_xi=0.001, r_p=0.01, r_v=0.02, r_r=0.0006, r_th=0.012,
                b_a=(-4.342596682195816e-07, -3.581072716118436e-18, -0.009999999990570729),
                b_g=(-2.66013609366142e-20, -1.933924486945935e-19, -3.870624673211354e-16))


def hat(w):
    return np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0.0]])


def so3_exp(w):
    """Rodrigues: exp(hat(w)); series below 1e-4 rad (the CUDA kernel uses the same switch)"""
    w = np.asarray(w, float)
    th2 = float(w @ w)
    K = hat(w)
    if th2 < 1e-8:
        a, b = 1.0 - th2 / 6.0, 0.5 - th2 / 24.0
    else:
        th = np.sqrt(th2)
        a, b = np.sin(th) / th, (1.0 - np.cos(th)) / th2
    return np.eye(3) + a * K + b * (K @ K)


def so3_log(R):
    """rotation vector of R through its unit quaternion (w >= 0), theta = 2 atan2(|q_v|, w): well conditioned near 0 and pi"""
    t = R[0, 0] + R[1, 1] + R[2, 2]
    if t > 0:
        s = np.sqrt(t + 1.0) * 2
        q = np.array([0.25 * s, (R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s])
    elif R[0, 0] > R[1, 1] and R[0, 0] > R[2, 2]:
        s = np.sqrt(1.0 + R[0, 0] - R[1, 1] - R[2, 2]) * 2
        q = np.array([(R[2, 1] - R[1, 2]) / s, 0.25 * s, (R[0, 1] + R[1, 0]) / s, (R[0, 2] + R[2, 0]) / s])
    elif R[1, 1] > R[2, 2]:
        s = np.sqrt(1.0 + R[1, 1] - R[0, 0] - R[2, 2]) * 2
        q = np.array([(R[0, 2] - R[2, 0]) / s, (R[0, 1] + R[1, 0]) / s, 0.25 * s, (R[1, 2] + R[2, 1]) / s])
    else:
        s = np.sqrt(1.0 + R[2, 2] - R[0, 0] - R[1, 1]) * 2
        q = np.array([(R[1, 0] - R[0, 1]) / s, (R[0, 2] + R[2, 0]) / s, (R[1, 2] + R[2, 1]) / s, 0.25 * s])
    if q[0] < 0:
        q = -q
    n = np.sqrt(q[1] ** 2 + q[2] ** 2 + q[3] ** 2)
    k = 2.0 / q[0] - 2.0 * n * n / (3.0 * q[0] ** 3) if n < 1e-6 else 2.0 * np.arctan2(n, q[0]) / n
    return k * q[1:]


def rpy_of(R):
    """tf::Matrix3x3::getEulerYPR as ros_utilities::q2rpy uses it (ros_utilities.cpp:32-45): (roll, pitch, yaw)"""
    if abs(R[2, 0]) >= 1.0:
        if R[2, 0] < 0:
            return np.array([np.arctan2(R[0, 1], R[0, 2]), np.pi / 2, 0.0])
        return np.array([np.arctan2(-R[0, 1], -R[0, 2]), -np.pi / 2, 0.0])
    pitch = -np.arcsin(R[2, 0])
    c = np.cos(pitch)
    return np.array([np.arctan2(R[2, 1] / c, R[2, 2] / c), pitch, np.arctan2(R[1, 0] / c, R[0, 0] / c)])


class Eskf:
    """one filter instance; ``predict`` / ``update`` are Eskf.cpp:97-141 / 197-300"""

    def __init__(self, p, v, R, **prm):
        c = dict(DEFAULTS); c.update(prm)
        self.p, self.v, self.R = np.array(p, float), np.array(v, float), np.array(R, float).reshape(3, 3)
        self.xi = np.zeros(3)                                                   # init_disturb, Config.cpp:113-116
        self.b_a, self.b_g = np.array(c["b_a"], float), np.array(c["b_g"], float)
        self.g = np.array([0.0, 0.0, -G])                                       # Config.cpp:90
        self.P = np.zeros((21, 21))                                             # ImuDo.h:155
        self.Q = np.diag([c["q_p"]] * 3 + [c["q_v"]] * 3 + [c["q_r"]] * 3 + [c["q_q"]] * 9 + [c["q_xi"]] * 3)   # Config.cpp:128-139
        self.Rm = np.diag([c["r_p"]] * 3 + [c["r_v"]] * 3 + [c["r_r"]] * 3 + [c["r_th"]] * 3)                  # :147-155

    def F(self, imu, dt):
        """set_F (Eskf.cpp:143-158), evaluated with the ALREADY propagated attitude (predict calls it last)"""
        a, w = imu[:3] - self.b_a, imu[3:] - self.b_g
        F = np.eye(21)
        F[0:3, 3:6] = np.eye(3) * dt
        F[3:6, 6:9] = -self.R @ hat(a) * dt
        F[3:6, 12:15] = -self.R * dt
        F[3:6, 15:18] = np.eye(3) * dt
        F[6:9, 6:9] = so3_exp(-w * dt)
        F[6:9, 9:12] = -np.eye(3) * dt
        return F

    def predict(self, imu):
        imu = np.asarray(imu, float)
        dt = DT
        a, w = imu[:3] - self.b_a, imu[3:] - self.b_g
        Ra = self.R @ a
        self.p = self.p + self.v * dt + 0.5 * Ra * dt * dt + 0.5 * self.g * dt * dt
        self.v = self.v + Ra * dt + self.g * dt
        self.R = self.R @ so3_exp(w * dt)
        F = self.F(imu, dt)
        self.P = F @ self.P @ F.T + self.Q

    def innovation(self, p_meas, v_meas, R_meas, thrusts, imu_raw, R_gt):
        y = np.zeros(12)
        y[0:3] = p_meas - self.p
        y[3:6] = v_meas - self.v
        y[6:9] = so3_log(self.R.T @ R_meas)
        v_B = self.R.T @ self.v
        ic = imu_raw - np.concatenate([self.b_a, self.b_g])
        mrb = np.array([MASS * ic[0] + MASS * ZG * ic[4], MASS * ic[1] - MASS * ZG * ic[3], MASS * ic[2]])     # M_rb rows 0..2
        g_B = R_gt.T @ np.array([0.0, 0.0, -G])                                  # Dynamics.cpp:180-181
        ma = ADDED[:3] * (ic[:3] + g_B)                                          # dynamics_Ma
        d = (-DL[:3] - DNL[:3] * np.abs(v_B)) * v_B                              # dynamics_D head(3)
        r, pch, _ = rpy_of(self.R)
        wb = MASS * G - BUOY
        gg = np.array([wb * np.sin(pch), -wb * np.cos(pch) * np.sin(r), -wb * np.cos(pch) * np.cos(r)])        # dynamics_g
        tau = K_ALLOC @ thrusts
        y[9:12] = tau - (mrb - self.xi + ma + d + gg)
        return y

    def update(self, p_meas, v_meas, R_meas, thrusts, imu_raw, R_gt):
        args = [np.asarray(a, float) for a in (p_meas, v_meas, R_meas, thrusts, imu_raw, R_gt)]
        args[2], args[5] = args[2].reshape(3, 3), args[5].reshape(3, 3)
        y = self.innovation(*args)
        H = np.zeros((12, 21))
        H[0:3, 0:3] = np.eye(3); H[3:6, 3:6] = np.eye(3); H[6:9, 6:9] = np.eye(3); H[9:12, 18:21] = -np.eye(3)
        S = H @ self.P @ H.T + self.Rm
        Kg = self.P @ H.T @ np.linalg.inv(S)
        dx = Kg @ y
        self.P = (np.eye(21) - Kg @ H) @ self.P
        # inject (Eskf.cpp:315-331): velocity twice, biases / g not at all
        self.p = self.p + dx[0:3]
        self.v = self.v + dx[3:6] + dx[3:6]
        self.R = self.R @ so3_exp(dx[6:9])
        self.xi = self.xi + dx[18:21]
        return self.R @ self.xi                                                  # what goes out on /xi (Eskf.cpp:64-77)
