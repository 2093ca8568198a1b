#!/usr/bin/env python
"""Cycles of the segments of one factor-sweep stage (clock64; library built with -DBR2_PROFILE -DBR2_PROFILE_STAGE into lib_sprof:
BR2_VARIANT=sprof BR2_NVCC_DEFS="-DBR2_PROFILE -DBR2_PROFILE_STAGE" python -m bluerov2_b200.build), for a lone warp per SM (B = 148),
one warp per scheduler (592) and the full resident set (2368).  Run with BR2_VARIANT=sprof."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bluerov2_b200 import solver as S, traj, workloads as wl

SEG = ("wait+refill", "loads+shuffles", "W'=Z'P (12 DMMA)", "H=W'Z (12 DMMA)", "s+,exchange,barrier", "cofactor,det,1/det", "K (2 DMMA),F store,shuffles", "update (4 DMMA),p,vin shuffles")
N = 40
for B in (148, 592, 2368):
    w = wl.tracking_batch(B, N, seed=0, pos_spread=0.3)
    s = S.BatchSolver(B, N)
    s.set_iterate(w["X"], w["U"])
    x0, lines = w["x0"].copy(), w["lines"].copy()
    T0, T1 = 3, 8
    for t in range(T1):
        yref = traj.window_batch(w["traj"], lines, N)
        if t == T0:
            s.phase_cycles(reset=True)
        u0, th, st = s.solve(x0, yref, w["p"])
        x0 = wl.plant_step(x0, u0, w["p"], 0.05); lines = lines + 1
    pc = list(s.phase_cycles().values())
    per = [v / B / (T1 - T0) / N for v in pc]
    print(json.dumps({"B": B, "factor_sweep_cycles_per_stage": round(per[0], 1), "segments": {SEG[i]: round(per[5 + i], 1) for i in range(8)},
                      "rollout_cycles_per_stage": round(per[2], 1), "epilogue_per_instance": round(per[12] * N)}))
    s.close()
