#!/usr/bin/env python
"""One adaptive-MPC tick per call for B vehicles through the host entry points: what bluerov2_ampc_node does every 50 ms
(bluerov2_dobmpc/src/bluerov2_ampc_node.cpp:26-29: EKF(); RLSFF(); solve();).

    python examples/ampc_tick.py [B] [ticks]          (needs a CUDA device)
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bluerov2_b200 import solver as S, traj, workloads as wl   # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
T = int(sys.argv[2]) if len(sys.argv) > 2 else 60
N = 40
w = wl.tracking_batch(B, N, seed=0, pos_spread=0.3, level=True)
sol = S.BatchSolver(B, N)
sol.set_option("ekf_model", 1)                                    # the AMPC filter: no damping in f / h (bluerov2_ampc.cpp:658-696)
sol.set_iterate(w["X"], w["U"])
x, lines = w["x0"].copy(), w["lines"].copy()
thr, acc = np.zeros((B, 6)), np.zeros((B, 6))
p = np.tile(wl.NOMINAL_P, (B, 1))
wrap = lambda a: (a + np.pi) % (2 * np.pi) - np.pi                # noqa: E731
# the vehicles join the reference mid-trajectory, so the node's accumulators (pre_yaw, yaw_sum) start at the current heading
sol.set_yaw_state(np.stack([wrap(x[:, 5]), x[:, 5]], axis=1).astype(np.float32))
true_dist = np.tile(np.array([4.0, -3.0, 2.0, 0.2]), (B, 1))      # a constant wrench the estimators have to find
for t in range(T):
    x0 = x.copy()
    x0[:, 5] = wrap(x0[:, 5])                                     # what tf getRPY hands the node ...
    sol.unwrap_yaw(x0)                                            # ... and what solve() makes of it (bluerov2_dob.cpp:272-304)
    sol.ekf(thr, x, acc)                                          # BLUEROV2_AMPC::EKF
    sol.rls(x, acc, compensate=True, out=p)                       # BLUEROV2_AMPC::RLSFF -> p[0..3] = theta(2) / coefficient
    yref = traj.window_batch(w["traj"], lines, N)
    u0, thr, status = sol.solve(x0, yref, p)                      # BLUEROV2_AMPC::solve
    xn = wl.plant_step(x, u0, w["p"], 0.05, dist=true_dist)
    acc = (xn[:, 6:] - x[:, 6:]) / 0.05
    x, lines = xn, lines + 1
st = sol.rls_state()
print(f"B = {B}, {T} ticks: forgetting factors lambda in [{st[:, :, 20].min():.2f}, {st[:, :, 20].max():.2f}], "
      f"estimated external force X mean {st[:, 0, 2].mean():+.2f}, statuses non-zero: {int((status != 0).sum())}")
sol.close()
