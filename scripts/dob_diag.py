"""diagnostic: status histogram / filter health of the config-3 workload over ticks (GPU)"""
import sys, numpy as np
sys.path.insert(0, ".")
import bench
from bluerov2_b200 import solver as S, workloads as wl
B, N = 4096, 40
w = wl.tracking_batch(B, N, seed=0, reference="lemniscate", pos_spread=float(sys.argv[2]) if len(sys.argv) > 2 else 0.2, level=True)
sol = S.BatchSolver(B, N)
rec, start = bench.record_closed_loop_dob(sol, w, 60, N, seed=0, settle=int(sys.argv[1]) if len(sys.argv) > 1 else 60)
ex, eP = start[2].copy(), start[3].copy()
if len(sys.argv) > 3:
    eP *= float(sys.argv[3]); ex[:, 12:15] = float(sys.argv[4])
sol.set_iterate(start[0], start[1]); sol.set_ekf_state(ex, eP)
for t in range(60):
    wf, p = sol.ekf(rec["thr"][t], rec["meas"][t], rec["acc"][t], compensate=True)
    u0, th, st = sol.solve_windowed(rec["meas"][t], rec["lines"][t], p)
    ex, eP = sol.ekf_state()
    bad = ~np.isfinite(ex).all(axis=1)
    if t % 6 == 0 or (st != 0).any():
        it, info = sol.stats()
        print(t, "status", dict(zip(*np.unique(st, return_counts=True))), "ekf nonfinite", int(bad.sum()), "|p|max", float(np.nanmax(np.abs(p[:, :4]))),
              "|thr|max", float(np.abs(rec["thr"][t]).max()), "|u0|max", float(np.nanmax(np.abs(u0))), "iters", float(it.mean()))
bad = np.where(~np.isfinite(ex).all(axis=1))[0]
np.set_printoptions(linewidth=220, precision=3, suppress=True)
print("bad instances", bad[:12], "lines at start", rec["lines"][0][bad[:12]])
for i in bad[:3]:
    for t in range(0, 8):
        print(i, t, "meas", rec["meas"][t][i, 3:], "acc", rec["acc"][t][i], "thr", rec["thr"][t][i])
