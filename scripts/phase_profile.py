#!/usr/bin/env python
"""scripts/phase_profile.py -- where the IPM kernel's time goes: SM cycles per phase summed over warps (clock64 around each sweep, library
built with -DBR2_PROFILE: BR2_VARIANT=prof BR2_NVCC_DEFS=-DBR2_PROFILE python -m bluerov2_b200.build), for the headline fast path, the
forced interior-point iteration and the saturated start.  Run with BR2_VARIANT=prof.  Cycles include the time a warp waits for its
turn on the SM, so the shares are shares of the kernel's warp-time."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bluerov2_b200 import solver as S, traj, workloads as wl

B, N = 4096, 40
out = {}
for name, spread, opts, ticks in (("fast_path", 0.5, {}, (5, 10)), ("forced_ipm", 0.5, {"fast_path": 0}, (5, 10)),
                                  ("saturated_tick0", 3.0, {}, (0, 1)), ("saturated_tick1", 3.0, {}, (1, 2))):
    w = wl.tracking_batch(B, N, seed=0, pos_spread=spread)
    s = S.BatchSolver(B, N)
    for k, v in opts.items():
        s.set_option(k, v)
    s.set_iterate(w["X"], w["U"])
    x0, lines = w["x0"].copy(), w["lines"].copy()
    tot, t_ipm, its = None, [], []
    for t in range(ticks[1]):
        yref = traj.window_batch(w["traj"], lines, N)
        if t == ticks[0]:
            s.phase_cycles(reset=True)
        u0, th, st = s.solve(x0, yref, w["p"])
        if t >= ticks[0]:
            t_ipm.append(s.last_kernel_times()[1] * 1e3); its.append(float(s.stats()[0].mean()))
        x0 = wl.plant_step(x0, u0, w["p"], 0.05)
        lines = lines + 1
    pc = s.phase_cycles()
    total = sum(pc.values())
    out[name] = {"ipm_ms_mean": float(np.mean(t_ipm)), "mean_iterations": float(np.mean(its)),
                 "share": {k: round(v / max(total, 1), 4) for k, v in pc.items() if v},
                 "cycles_per_instance_and_tick": {k: round(v / B / (ticks[1] - ticks[0])) for k, v in pc.items() if v}}
    s.close()
print(json.dumps(out, indent=1))
