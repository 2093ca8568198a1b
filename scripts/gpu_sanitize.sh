#!/bin/bash
# compute-sanitizer over the kernels: memcheck (out-of-bounds / misaligned) and racecheck (shared-memory hazards between the lanes of a warp
# that hand data over through shared memory) on a small EKF + tick + solve workload
O=gpurun_out; mkdir -p $O
cat > /tmp/san_work.py <<'PY'
import numpy as np, sys, os
sys.path.insert(0, os.getcwd())
from bluerov2_b200 import solver as S, traj, workloads as wl
B, N = 64, 12
w = wl.tracking_batch(B, N, seed=0, pos_spread=3.0)
s = S.BatchSolver(B, N)
s.set_iterate(w["X"], w["U"]); s.set_trajectory(w["traj"])
x0, lines = w["x0"].copy(), w["lines"].copy()
rng = np.random.default_rng(0)
s.ekf_reset()
for t in range(3):
    yref = traj.window_batch(w["traj"], lines, N)
    u0, th, st = s.solve(x0, yref, w["p"])
    wf, p = s.ekf(th, x0, rng.uniform(-0.1, 0.1, (B, 6)))
    x0 = wl.plant_step(x0, u0, w["p"], 0.05); lines = lines + 1
s.set_option("fast_path", 0)
u0, th, st = s.solve(x0, traj.window_batch(w["traj"], lines, N), w["p"])
print("status", int((st != 0).sum()), "ok")
# one-call ticks: device graph with the plant step, host tick with resident parameters and an announced reference window
import torch
s.set_option("fast_path", 1)
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
out = (pin(np.zeros((B, 4))), pin(np.zeros((B, 6))), pin(np.zeros(B, dtype=np.int32)))
refs = [pin(traj.window_batch(w["traj"], lines + t, N)) for t in range(4)]
hx, hp = pin(x0), pin(w["p"])
for t in range(3):
    s.set_next_yref(refs[t + 1])
    s.tick(hx, p=hp if t == 0 else None, yref=refs[t], out=out)
d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
dx, dl, dp, da = d(x0), d(lines.astype(np.int32)), d(w["p"]), torch.zeros((B, 6), dtype=torch.float64, device="cuda")
for t in range(3):
    s.tick(dx, p=dp, lines=dl, body_acc=da, plant_h=0.05)
torch.cuda.synchronize()
f = S.BatchEskf(B) if hasattr(S, "BatchEskf") else None
print("ticks ok")
s.close()
PY
for tool in memcheck racecheck; do
  timeout 1500 compute-sanitizer --tool $tool --print-limit 20 python /tmp/san_work.py > $O/sanitize_$tool.log 2>&1; echo "$tool rc=$?"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|Invalid|status" $O/sanitize_$tool.log | head -12
done
