#!/usr/bin/env python
"""scripts/active_set_probe.py -- cost of the saturated regime: closed loop from a 3 m start (thrusters saturated during the first
ticks), IPM kernel time per tick with (a) the default paths (interior fast path + primal-dual active-set iteration), (b) the
active-set iteration switched off (round-1 default: saturated instances go to the interior-point iteration), (c) interior-point only."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bluerov2_b200 import solver as S, traj, workloads as wl

B, N, T = 4096, 40, 12
w = wl.tracking_batch(B, N, seed=0, pos_spread=3.0)
out = {}
for name, opts in (("default", {}), ("no_active_set", {"active_set_path": 0}), ("ipm_only", {"fast_path": 0})):
    s = S.BatchSolver(B, N)
    for k, v in opts.items():
        s.set_option(k, v)
    s.set_iterate(w["X"], w["U"])
    x0, lines = w["x0"].copy(), w["lines"].copy()
    rows = []
    for t in range(T):
        yref = traj.window_batch(w["traj"], lines, N)
        u0, th, st = s.solve(x0, yref, w["p"])
        it, _ = s.stats()
        tl, ti = s.last_kernel_times()
        X, U = s.get_iterate()
        sat = int((np.abs(np.abs(U) - 50.0) < 1e-6).any(axis=(1, 2)).sum())
        rows.append({"tick": t, "ipm_ms": round(ti * 1e3, 4), "mean_iterations": round(float(it.mean()), 3), "max_iterations": int(it.max()), "lin_ms": round(tl * 1e3, 4), "saturated_instances": sat,
                     "nonzero_status": int((st != 0).sum())})
        x0 = wl.plant_step(x0, u0, w["p"], 0.05)
        lines = lines + 1
    s.close()
    out[name] = rows
print(json.dumps(out))
for name, rows in out.items():
    print(name, "ipm ms per tick:", [r["ipm_ms"] for r in rows], "saturated:", [r["saturated_instances"] for r in rows], file=sys.stderr)
