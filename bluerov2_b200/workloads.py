"""Synthetic batched inputs for the BASELINE.json configurations (host side, numpy).

Everything here is input *generation*; nothing is solved.  Seeds and distributions are part of the benchmark
definition, so tests, bench.py and the CPU baseline all draw the same tensors.
"""
from __future__ import annotations

import numpy as np

from . import traj as _traj

NX, NU, NP, NY = 12, 4, 16, 16

# bluerov2_dob.cpp:340-353: added mass, linear damping, quadratic damping handed to the OCP every tick
NOMINAL_P = np.array([0, 0, 0, 0, 1.7182, 0, 5.468, 0.4006, -11.7391, -20, -31.8678, -5,
                      -18.18, -21.66, -36.99, -1.55], dtype=np.float64)
ROTOR_CONSTANT = 0.026546960744430276      # bluerov2_dob.h:180
COMPENSATE_COEF = 0.032546960744430276     # bluerov2_dob.h:179


def time_steps(N: int, Tf: float = 1.0) -> np.ndarray:
    """generate_c_code.py:17,24,82: Tf = 1 s split into N equal shooting intervals."""
    return np.full(N, Tf / N, dtype=np.float64)


def tracking_batch(B: int, N: int, seed: int = 0, reference: str = "circle", pos_spread: float = 0.5,
                   level: bool = False):
    """Config 2/4/5 inputs: random phase on the reference, x0 = first reference row + uniform offsets.

    ``pos_spread`` = 0.5 m is the nominal set; 3.0 m forces active input bounds (SURVEY 8d "second set");
    ``level`` zeroes the initial roll / pitch offsets (config 3).
    Returns dict(lines, x0, yref, p, X, U, traj): X_k = x0 for all k and U = 0 is the tick-0 warm start.
    """
    rng = np.random.default_rng(seed)
    tr = _traj.circle() if reference == "circle" else _traj.lemniscate()
    lines = rng.integers(0, tr.shape[0] - (N + 1), size=B)
    yref = _traj.window_batch(tr, lines, N)
    d = np.concatenate([
        rng.uniform(-pos_spread, pos_spread, (B, 3)),
        rng.uniform(-0.1, 0.1, (B, 2)),
        rng.uniform(-0.3, 0.3, (B, 1)),
        rng.uniform(-0.5, 0.5, (B, 3)),
        rng.uniform(-0.1, 0.1, (B, 3)),
    ], axis=1)
    if level:
        # no initial roll / pitch (angles and rates): the OCP model has neither roll/pitch damping nor thrust coupling,
        # so they then stay zero -- the regime in which the reference's EKF (explicit RK4 at 50 ms on roll dynamics
        # with |lambda dt| > 2.8) remains stable.  Same random stream, so the other offsets are unchanged.
        d[:, 3:5] = 0.0
        d[:, 9:11] = 0.0
    x0 = yref[:, 0, :12] + d
    p = np.tile(NOMINAL_P, (B, 1))
    X = np.repeat(x0[:, None, :], N + 1, axis=1).copy()
    U = np.zeros((B, N, NU))
    return dict(lines=lines, x0=np.ascontiguousarray(x0), yref=yref, p=p, X=X, U=U, traj=tr)


def wave_disturbance(B: int, seed: int = 1):
    """Config 3 'sampled wave disturbances': mode 0 of applyBodyWrench (bluerov2_dob.cpp:774-797).

    F = sin(tau) * A with A_x,y,z ~ U(0.5, 1) * 6 N, T_z = sin(tau) * A_y / 3, tau advancing 0.05 * 2.5 per tick.
    Returns (amp[B,4] = (A_x, A_y, A_z, A_y/3), tau0[B]).
    """
    rng = np.random.default_rng(seed)
    A = rng.uniform(0.5, 1.0, (B, 3)) * 6.0
    amp = np.concatenate([A, A[:, 1:2] / 3.0], axis=1)
    tau0 = rng.uniform(0, 2 * np.pi, B)
    return amp, tau0


def wave_at(amp: np.ndarray, tau0: np.ndarray, tick: int) -> np.ndarray:
    tau = tau0 + 0.125 * tick
    return np.sin(tau)[:, None] * amp


def wrench_table() -> np.ndarray:
    """The wrench series of applyBodyWrench mode 2 (bluerov2_dob.cpp:818-874): [496, 4] = (fx, fy, fz, tz), the reference's
    config/force{x,y,z}.txt and torquez.txt column-stacked (fixture tests/golden/wrench_table.npz, made by make_wrench_table.py)."""
    import os
    return np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "wrench_table.npz"))["table"]


def wrench_at(table: np.ndarray, tick: int, phase: np.ndarray | None = None, B: int | None = None) -> np.ndarray:
    """mode 2: row min(tick + phase[b], rows - 1) of the series for every instance -> dist [B, 4] (all four columns at the same
    counter, fx_counter++ per tick)"""
    ph = np.zeros(B, dtype=np.int64) if phase is None else np.asarray(phase, dtype=np.int64)
    return table[np.clip(tick + ph, 0, table.shape[0] - 1)]


def dob_params(esti_x: np.ndarray, compensate: bool = True) -> np.ndarray:
    """OCP parameter fill of BLUEROV2_DOB::solve (bluerov2_dob.cpp:324-355) from EKF states [B,18]."""
    B = esti_x.shape[0]
    p = np.tile(NOMINAL_P, (B, 1))
    if compensate:
        p[:, 0] = esti_x[:, 12] / COMPENSATE_COEF
        p[:, 1] = esti_x[:, 13] / COMPENSATE_COEF
        p[:, 2] = esti_x[:, 14] / ROTOR_CONSTANT
        p[:, 3] = esti_x[:, 17] / ROTOR_CONSTANT
    return p


def ctrl_params(disturb: np.ndarray | None = None, dompc: bool = False) -> np.ndarray:
    """OCP parameter fill of the consolidated node, BLUEROV2_CTRL::set_mpc_constraints (src/ctrller/mpc.cpp:139-187):
    ctrller_type MPC -> p[0..3] = 0; DOMPC -> the ``/disturbance`` estimate (x, y, z) [B,3] divided by the hard-coded
    0.032546960744430276 (x, y) / rotor_constant (z) and p[3] = 0 (the yaw moment is never compensated there, :165);
    p[4..15] nominal in both modes."""
    if not dompc:
        return NOMINAL_P[None, :].copy() if disturb is None else np.tile(NOMINAL_P, (np.asarray(disturb).shape[0], 1))
    d = np.ascontiguousarray(disturb, dtype=np.float64)
    p = np.tile(NOMINAL_P, (d.shape[0], 1))
    p[:, 0] = d[:, 0] / COMPENSATE_COEF
    p[:, 1] = d[:, 1] / COMPENSATE_COEF
    p[:, 2] = d[:, 2] / ROTOR_CONSTANT
    p[:, 3] = 0.0
    return p


def ctrl_yref(ref12: np.ndarray) -> np.ndarray:
    """Horizon reference of the consolidated node, BLUEROV2_CTRL::set_ref (src/ctrller/mpc.cpp:199-262): only the 12 state
    columns of each yref row are written from the ``/ref_traj`` preview, the input reference stays 0.
    ref12 [..., N+1, 12] -> yref [..., N+1, 16]."""
    r = np.asarray(ref12, dtype=np.float64)
    y = np.zeros(r.shape[:-1] + (NY,))
    y[..., :12] = r
    return y


# ---------------------------------------------------------------------------------------------------------
# Nominal plant for closed-loop input generation (numpy, vectorised over the batch).  Same equations as the
# OCP model (bluerov2.py:103-137) -- this is workload synthesis on the host, not a solver path.
# ---------------------------------------------------------------------------------------------------------
_M, _IX, _IY, _IZ, _ZG, _G, _BUOY = 11.26, 0.3, 0.63, 0.58, 0.02, 9.81, 0.66


def plant_ode(x: np.ndarray, u: np.ndarray, p: np.ndarray, dist: np.ndarray | None = None) -> np.ndarray:
    """x[B,12], u[B,4], p[B,16] -> xdot[B,12]; ``dist`` [B,4] adds a true disturbance (X, Y, Z, N) on top of p[:, :4]."""
    rc = ROTOR_CONSTANT
    phi, th, psi = x[:, 3], x[:, 4], x[:, 5]
    uu, v, w, pp, q, r = (x[:, i] for i in range(6, 12))
    sphi, cphi, sth, cth, spsi, cpsi = np.sin(phi), np.cos(phi), np.sin(th), np.cos(th), np.sin(psi), np.cos(psi)
    kt0 = -4 * 0.707 * u[:, 0] / rc
    kt1 = 4 * 0.707 * u[:, 1] / rc
    kt2 = -2 * u[:, 2] / rc
    kt5 = ((2 * 0.167 - 2 * 0.175) * u[:, 1] + (2 * 0.167 + 2 * 0.175) * u[:, 3]) / rc
    d = p[:, :4] if dist is None else p[:, :4] + dist
    f = np.empty_like(x)
    f[:, 0] = cpsi * cth * uu + (-spsi * cphi + cpsi * sth * sphi) * v + (spsi * sphi + cpsi * cphi * sth) * w
    f[:, 1] = spsi * cth * uu + (cpsi * cphi + sphi * sth * spsi) * v + (-cpsi * sphi + sth * spsi * cphi) * w
    f[:, 2] = -sth * uu + cth * sphi * v + cth * cphi * w
    f[:, 3] = pp + spsi * sth / cth * q + cphi * sth / cth * r
    f[:, 4] = cphi * q + sphi * r
    f[:, 5] = sphi / cth * q + cphi / cth * r
    f[:, 6] = (kt0 - _BUOY * sth + d[:, 0] + p[:, 8] * uu + p[:, 12] * np.abs(uu) * uu) / (_M + p[:, 4])
    f[:, 7] = (kt1 + _BUOY * cth * sphi + d[:, 1] + p[:, 9] * v + p[:, 13] * np.abs(v) * v) / (_M + p[:, 5])
    f[:, 8] = (kt2 + _BUOY * cth * cphi + d[:, 2] + p[:, 10] * w + p[:, 14] * np.abs(w) * w) / (_M + p[:, 6])
    f[:, 9] = ((_IY - _IZ) * q * r - _M * _ZG * _G * cth * sphi) / _IX
    f[:, 10] = ((_IZ - _IX) * pp * r - _M * _ZG * _G * sth) / _IY
    f[:, 11] = (kt5 - (_IY - _IX) * pp * q + d[:, 3] + p[:, 11] * r + p[:, 15] * np.abs(r) * r) / (_IZ + p[:, 7])
    return f


def plant_step(x: np.ndarray, u: np.ndarray, p: np.ndarray, h: float = 0.05, dist: np.ndarray | None = None) -> np.ndarray:
    """one classical RK4 step of the nominal plant (20 Hz node rate, bluerov2_dob_node.cpp:7)"""
    k1 = plant_ode(x, u, p, dist)
    k2 = plant_ode(x + 0.5 * h * k1, u, p, dist)
    k3 = plant_ode(x + 0.5 * h * k2, u, p, dist)
    k4 = plant_ode(x + h * k3, u, p, dist)
    return x + h * (k1 / 6 + k2 / 3 + k3 / 3 + k4 / 6)
