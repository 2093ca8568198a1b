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


def tracking_batch(B: int, N: int, seed: int = 0, reference: str = "circle", pos_spread: float = 0.5):
    """Config 2/4/5 inputs: random phase on the reference, x0 = first reference row + uniform offsets.

    ``pos_spread`` = 0.5 m is the nominal set; 3.0 m forces active input bounds (SURVEY 8d "second set").
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
