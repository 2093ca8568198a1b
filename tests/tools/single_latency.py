#!/usr/bin/env python
"""BASELINE config 1: ONE instance, circle reference, N = 20 (Ts = 0.05 s), closed loop for 200 ticks, driven through the
acados-generated C-ABI from C (tests/abi/_bin/dob_tick) -- per-tick latency as the node would see it (`time_tot`), next to
the CPU oracle's single-instance latency on this host.  Prints one JSON line."""
import json, os, struct, subprocess, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from bluerov2_b200 import traj, workloads as wl
from oracle import Oracle, NOMINAL_P

N, T = 20, 200
o = Oracle()
w = wl.tracking_batch(1, N, seed=0, pos_spread=0.5)
Ts = wl.time_steps(N)
X, U = w["X"][0].copy(), w["U"][0].copy()
x0, line = w["x0"][0].copy(), int(w["lines"][0])
ticks, t_cpu, u_cpu = [], [], []
for t in range(T):
    yref = traj.window(w["traj"], line + t, N)
    ticks.append((x0.copy(), yref))
    t0 = time.perf_counter()
    st, _ = o.rti_step(Ts, x0, yref, NOMINAL_P, X, U)
    t_cpu.append(time.perf_counter() - t0)
    u_cpu.append(U[0].copy())
    x0 = o.erk4(x0, U[0], NOMINAL_P, 0.05)
fin, fout = "/tmp/single_in.bin", "/tmp/single_out.bin"
with open(fin, "wb") as f:
    f.write(struct.pack("ii", N, T))
    for x0t, yref in ticks:
        f.write(x0t.tobytes()); f.write(NOMINAL_P.tobytes()); f.write(np.ascontiguousarray(yref).tobytes())
exe = os.path.join(ROOT, "tests", "abi", "_bin", "dob_tick")
r = subprocess.run([exe, fin, fout], capture_output=True, text=True)
assert r.returncode == 0, r.stderr
raw = open(fout, "rb").read()
rec = 8 + 8 * 6
t_gpu = np.array([np.frombuffer(raw, dtype=np.float64, count=6, offset=t * rec + 8)[4] for t in range(T)])
u_gpu = np.array([np.frombuffer(raw, dtype=np.float64, count=6, offset=t * rec + 8)[:4] for t in range(T)])
st = np.array([struct.unpack_from("i", raw, t * rec)[0] for t in range(T)])
print(json.dumps({"config": "1 instance, circle reference, N=20, Ts=0.05 s, 200 closed-loop ticks through the acados C-ABI",
                  "gpu_ms_per_tick_median": float(np.median(t_gpu[5:]) * 1e3), "gpu_ms_per_tick_p99": float(np.quantile(t_gpu[5:], 0.99) * 1e3),
                  "cpu_oracle_ms_per_tick_median": float(np.median(t_cpu[5:]) * 1e3), "status_nonzero": int((st != 0).sum()),
                  "max_abs_u_diff": float(np.abs(u_gpu - np.array(u_cpu)).max()), "deadline_ms": 50.0}))
