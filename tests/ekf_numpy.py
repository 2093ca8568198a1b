"""Independent numpy restatement of one BLUEROV2_DOB::EKF step, written from the reference's C++ alone
(bluerov2_dobmpc/src/bluerov2_dob.cpp:41-65 constants, :495-545 EKF, :621-633 RK4, :637-695 f, :699-719 h, :722-752 Jacobians;
constants bluerov2_dob.h:171-208) -- NOT from oracle/bluerov2_oracle.c.  np.linalg.inv is LAPACK's partial-pivot LU, the same
algorithm class as Eigen's general .inverse() for an 18 x 18 matrix.  Used only by tests, as the second opinion on orc_ekf_step."""
import numpy as np

DT, MASS, IX, IY, IZ, ZG, G, BUOY = 0.05, 11.26, 0.3, 0.63, 0.58, 0.02, 9.81, 0.661618
ADDED = np.array([1.7182, 0, 5.468, 0, 1.2481, 0.4006])
DL = np.array([-11.7391, -20, -31.8678, -25, -44.9085, -5])
DNL = np.array([-18.18, -21.66, -36.99, -1.55, -1.55, -1.55])
M = np.diag([MASS + ADDED[0], MASS + ADDED[1], MASS + ADDED[2], IX + ADDED[3], IY + ADDED[4], IZ + ADDED[5]])
M[0, 4] = MASS * ZG; M[1, 3] = -MASS * ZG; M[3, 1] = -MASS * ZG; M[4, 0] = MASS * ZG
INVM = np.linalg.inv(M)
K = np.array([[0.7071067811847433, 0.7071067811847433, -0.7071067811919605, -0.7071067811919605, 0.0, 0.0],
              [0.7071067811883519, -0.7071067811883519, 0.7071067811811348, -0.7071067811811348, 0.0, 0.0],
              [0, 0, 0, 0, 1, 1],
              [0.051265241636155506, -0.05126524163615552, 0.05126524163563227, -0.05126524163563227, -0.11050000000000001, 0.11050000000000003],
              [-0.05126524163589389, -0.051265241635893896, 0.05126524163641713, 0.05126524163641713, -0.002499999999974481, -0.002499999999974481],
              [0.16652364696949604, -0.16652364696949604, -0.17500892834341342, 0.17500892834341342, 0.0, 0.0]])
NOISE_Q = np.diag([DT ** 4 / 4] * 6 + [DT ** 2] * 12)
NOISE_R = np.eye(18) * (DT ** 4 / 4)


def f(x, u):
    KAu = K @ u
    s3, c3, s4, c4, s5, c5 = np.sin(x[3]), np.cos(x[3]), np.sin(x[4]), np.cos(x[4]), np.sin(x[5]), np.cos(x[5])
    xd = np.zeros(18)
    xd[0] = (c5 * c4) * x[6] + (-s5 * c3 + c5 * s4 * s3) * x[7] + (s5 * s3 + c5 * c3 * s4) * x[8]
    xd[1] = (s5 * c4) * x[6] + (c5 * c3 + s3 * s4 * s5) * x[7] + (-c5 * s3 + s4 * s5 * c3) * x[8]
    xd[2] = (-s4) * x[6] + (c4 * s3) * x[7] + (c4 * c3) * x[8]
    xd[3] = x[9] + (s5 * s4 / c4) * x[10] + c3 * s4 / c4 * x[11]          # sin(psi), as in the reference (:646)
    xd[4] = c3 * x[10] + s3 * x[11]
    xd[5] = (s3 / c4) * x[10] + (c3 / c4) * x[11]
    xd[6] = INVM[0, 0] * (KAu[0] + MASS * x[11] * x[7] - MASS * x[10] * x[8] - BUOY * s4 + x[12] + DL[0] * x[6] + DNL[0] * abs(x[6]) * x[6])
    xd[7] = INVM[1, 1] * (KAu[1] - MASS * x[11] * x[6] + MASS * x[9] * x[8] + BUOY * c4 * s3 + x[13] + DL[1] * x[7] + DNL[1] * abs(x[7]) * x[7])
    xd[8] = INVM[2, 2] * (KAu[2] + MASS * x[10] * x[6] - MASS * x[9] * x[7] + BUOY * c4 * c3 + x[14] + DL[2] * x[8] + DNL[2] * abs(x[8]) * x[8])
    xd[9] = INVM[3, 3] * (KAu[3] + (IY - IZ) * x[10] * x[11] - MASS * ZG * G * c4 * s3 + x[15] + DL[3] * x[9] + DNL[3] * abs(x[9]) * x[9])
    xd[10] = INVM[4, 4] * (KAu[4] + (IZ - IX) * x[9] * x[11] - MASS * ZG * G * s4 + x[16] + DL[4] * x[10] + DNL[4] * abs(x[10]) * x[10])
    xd[11] = INVM[5, 5] * (KAu[5] - (IY - IX) * x[9] * x[10] + x[17] + DL[5] * x[11] + DNL[5] * abs(x[11]) * x[11])
    return xd


def rk4(x, u):
    k1 = f(x, u) * DT
    k2 = f(x + k1 / 2, u) * DT
    k3 = f(x + k2 / 3, u) * DT          # /3, as in the reference (:630)
    k4 = f(x + k3, u) * DT
    return x + (k1 + 2 * k2 + 2 * k3 + k4) / 6


def h(x, acc):
    s3, c3, s4, c4 = np.sin(x[3]), np.cos(x[3]), np.sin(x[4]), np.cos(x[4])
    y = np.zeros(18)
    y[:12] = x[:12]
    y[12] = M[0, 0] * acc[0] - MASS * x[11] * x[7] + MASS * x[10] * x[8] + BUOY * s4 - x[12] - DL[0] * x[6] - DNL[0] * abs(x[6]) * x[6]
    y[13] = M[1, 1] * acc[1] + MASS * x[11] * x[6] - MASS * x[9] * x[8] - BUOY * c4 * s3 - x[13] - DL[1] * x[7] - DNL[1] * abs(x[7]) * x[7]
    y[14] = M[2, 2] * acc[2] - MASS * x[10] * x[6] + MASS * x[9] * x[7] - BUOY * c4 * c3 - x[14] - DL[2] * x[8] - DNL[2] * abs(x[8]) * x[8]
    y[15] = M[3, 3] * acc[3] - (IY - IZ) * x[10] * x[11] + MASS * ZG * G * c4 * s3 - x[15] - DL[3] * x[9] - DNL[3] * abs(x[9]) * x[9]
    y[16] = M[4, 4] * acc[4] - (IZ - IX) * x[9] * x[11] + MASS * ZG * G * s4 - x[16] - DL[4] * x[10] - DNL[4] * abs(x[10]) * x[10]
    y[17] = M[5, 5] * acc[5] + (IY - IX) * x[9] * x[10] - x[17] - DL[5] * x[11] - DNL[5] * abs(x[11]) * x[11]
    return y


def jac(fun, x, d=1e-6):
    f0 = fun(x)
    J = np.zeros((18, 18))
    for i in range(18):
        x1 = x.copy(); x1[i] += d
        J[:, i] = (fun(x1) - f0) / d
    return J


def ekf_step(esti_x, esti_P, thrusts, meas12, body_acc):
    """returns (esti_x, esti_P, wf_disturbance) after one EKF() call"""
    meas_u = np.asarray(thrusts, float)
    tau = K @ meas_u
    meas_y = np.concatenate([meas12, tau])
    F = jac(lambda x: rk4(x, meas_u), esti_x)
    x_pred = rk4(esti_x, meas_u)
    P_pred = F @ esti_P @ F.T + NOISE_Q
    H = jac(lambda x: h(x, body_acc), x_pred)
    y_err = meas_y - h(x_pred, body_acc)
    Kal = P_pred @ H.T @ np.linalg.inv(H @ P_pred @ H.T + NOISE_R)
    x_new = x_pred + Kal @ y_err
    IKH = np.eye(18) - Kal @ H
    P_new = IKH @ P_pred @ IKH.T + Kal @ NOISE_R @ Kal.T
    y, e = meas_y, x_new
    s3, c3, s4, c4, s5, c5 = np.sin(y[3]), np.cos(y[3]), np.sin(y[4]), np.cos(y[4]), np.sin(y[5]), np.cos(y[5])
    wf = np.array([
        (c5 * c4) * e[12] + (-s5 * c3 + c5 * s4 * s3) * e[13] + (s5 * s3 + c5 * c3 * s4) * e[14],
        (s5 * c4) * e[12] + (c5 * c3 + s3 * s4 * s5) * e[13] + (-c5 * s3 + s4 * s5 * c3) * e[14],
        (-s4) * e[12] + (c4 * s3) * e[13] + (c4 * c3) * e[14],
        e[15] + (s5 * s4 / c4) * e[16] + c3 * s4 / c4 * e[17],
        c3 * e[16] + s3 * e[17],
        (s3 / c4) * e[16] + (c3 / c4) * e[17]])
    return x_new, P_new, wf
