"""Reference trajectories and horizon windowing (host side, numpy).

Re-creates the benchmark inputs of the reference without shipping its data files:

* ``circle()``      -- bluerov2_path/config/traj/circle.py:10-56 (r = 2 m, v = 1.5 m/s, 0.05 s, 240 s), including its
  two quirks: column 6/7 are the *scalars* ``velocity_body.flatten()[0:2]`` = (1.5, 1.5*cos(0.0375)) (circle.py:45-46)
  and the heave reference u3 = 57.5 sits outside the +-50 input box (circle.py:55).
* ``lemniscate()``  -- bluerov2_path/config/traj/lemniscate.py:8-37 (amp 2, frq 0.5, 60 s).

Both are rounded through ``"%f"`` exactly like ``np.savetxt(..., fmt='%f')`` (circle.py:72), so the arrays equal what
``readDataFromFile`` (bluerov2_dob.cpp:182-216) parses from ``circle.txt`` / ``lemniscate.txt``
(checked against the reference files in tests/test_traj.py when /root/reference is present).

``window()`` is the reference's horizon fill rule ``BLUEROV2_DOB::ref_cb`` (bluerov2_dob.cpp:218-265) ==
``BLUEROV2_PATH::read_N_pub`` (bluerov2_path/src/bluerov2_path.cpp:79-118): rows line..line+N, clamped to the last row.
The reference's inner loop runs ``j <= NY`` (one past the row, UB); columns 0..15 are what is defined.
"""
from __future__ import annotations

import numpy as np

SAMPLE_TIME = 0.05


def _through_percent_f(a: np.ndarray) -> np.ndarray:
    flat = np.array([float("%f" % v) for v in a.ravel()], dtype=np.float64)
    return flat.reshape(a.shape)


def circle(duration: float = 240.0, r: float = 2.0, v: float = 1.5, z0: float = -20.0) -> np.ndarray:
    n = int(duration / SAMPLE_TIME + 1)
    t = np.append(np.arange(0, duration, SAMPLE_TIME), duration)
    traj = np.zeros((n, 16))
    traj[:, 0] = -r * np.cos(t * v / r)
    traj[:, 1] = -r * np.sin(t * v / r)
    traj[:, 2] = z0
    traj[:, 5] = t * v / r - 0.5 * np.pi
    psi = traj[:, 5]
    # velocity_body.flatten()[0] and [1] of the (n,2,n) product in circle.py:39-46
    traj[:, 6] = np.cos(psi[0]) * (v * np.cos(psi[0])) + np.sin(psi[0]) * (v * np.sin(psi[0]))
    traj[:, 7] = np.cos(psi[0]) * (v * np.cos(psi[1])) + np.sin(psi[0]) * (v * np.sin(psi[1]))
    traj[:, 14] = 57.5
    return _through_percent_f(traj)


def lemniscate(duration: float = 60.0, amp: float = 2.0, frq: float = 0.5, z0: float = -20.0) -> np.ndarray:
    n = int(duration / SAMPLE_TIME + 1)
    t = np.append(np.arange(0, duration, SAMPLE_TIME), duration)
    traj = np.zeros((n, 16))
    traj[:, 0] = amp * np.cos(t * frq)
    traj[:, 1] = amp * np.sin(t * frq) * np.cos(t * frq)
    traj[:, 2] = z0
    traj[:, 6] = -amp * frq * np.sin(t * frq)
    traj[:, 7] = amp * frq * np.cos(t * 2 * frq)
    return _through_percent_f(traj)


def window(traj: np.ndarray, line: int, N: int) -> np.ndarray:
    """yref[(N+1) x 16] for the tick that starts at trajectory row ``line`` (ref_cb, bluerov2_dob.cpp:218-265)."""
    idx = np.minimum(np.arange(line, line + N + 1), traj.shape[0] - 1)
    return np.ascontiguousarray(traj[idx, :16])


def window_batch(traj: np.ndarray, lines: np.ndarray, N: int) -> np.ndarray:
    idx = np.minimum(lines[:, None] + np.arange(N + 1)[None, :], traj.shape[0] - 1)
    return np.ascontiguousarray(traj[idx, :16])
