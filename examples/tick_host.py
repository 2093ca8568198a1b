#!/usr/bin/env python
"""One control tick = one call (br2_batch_tick_host): B vehicles tracking the circle, host buffers, the whole tick one CUDA graph.
  - windowed reference: the trajectory is uploaded once, a tick names one row per vehicle (ref_cb, bluerov2_dob.cpp:218-265);
  - the OCP parameters are supplied with the first call and stay in the solver (bluerov2_acados_update_params semantics);
  - then the same loop with an explicit (N+1) x 16 window per vehicle, the next tick's window announced one tick ahead.
Needs a B200 (there is no CPU path).  python examples/tick_host.py [B]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bluerov2_b200 import solver as S, traj, workloads as wl   # noqa: E402

B, N, T = (int(sys.argv[1]) if len(sys.argv) > 1 else 1024), 40, 40
w = wl.tracking_batch(B, N, seed=0, reference="circle", pos_spread=0.5)
sol = S.BatchSolver(B, N)
sol.set_trajectory(w["traj"])
sol.set_iterate(w["X"], w["U"])

pinned = lambda shape, dtype=np.float64: S.pinned_empty(shape, dtype)      # noqa: E731  (page-locked: the tick becomes a cached graph)
x0, lines, p = pinned((B, 12)), pinned((B,), np.int32), pinned((B, 16))
out = (pinned((B, 4)), pinned((B, 6)), pinned((B,), np.int32))            # u0, thruster commands, acados status per vehicle
x0[:], lines[:], p[:] = w["x0"], w["lines"], w["p"]

t0 = time.perf_counter()
for t in range(T):
    u0, thrust, status = sol.tick(x0, p=p if t == 0 else None, lines=lines, out=out)     # parameters: once
    assert not status.any()
    x0[:] = wl.plant_step(x0, u0, w["p"], 0.05)            # the vehicles move (host-side nominal plant, stands for the sensors)
    lines += 1
dt = time.perf_counter() - t0
print(f"windowed: {T} ticks of {B} vehicles, {1e3 * dt / T:.3f} ms per tick including the host-side plant step, |u0|max {np.abs(out[0]).max():.3f}")

# explicit reference windows, announced one tick ahead
win = [S.pinned_empty((B, N + 1, 16), write_combined=True) for _ in range(2)]
win[0][...] = traj.window_batch(w["traj"], lines.astype(np.int64), N)
for t in range(T):
    nxt = win[(t + 1) & 1]
    nxt[...] = traj.window_batch(w["traj"], lines.astype(np.int64) + 1, N)       # the planner's window for the NEXT tick
    sol.set_next_yref(nxt)                                                        # ... uploaded while this tick computes
    u0, thrust, status = sol.tick(x0, yref=win[t & 1], out=out)
    assert not status.any()
    x0[:] = wl.plant_step(x0, u0, w["p"], 0.05)
    lines += 1
print(f"explicit windows: {T} more ticks, position error to the reference {np.abs(x0[:, :3] - win[T & 1][:, 0, :3]).max():.3f} m")
sol.close()
