#!/usr/bin/env python
"""scripts/ekf_phase_profile.py -- SM cycles per phase of the EKF kernel (clock64, library built with -DBR2_PROFILE; run with
BR2_VARIANT=prof) and its CUDA-event time on a config-3 style batch."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bluerov2_b200 import solver as S

B = int(os.environ.get("EKF_B", "4096"))
rng = np.random.default_rng(0)
dev = torch.device("cuda:0")
s = S.BatchSolver(B, 40, device=0)
d = lambda a: torch.from_numpy(a).to(dev)   # noqa: E731
thr = d(rng.uniform(-10, 10, (B, 6))); meas = d(np.concatenate([rng.uniform(-1, 1, (B, 3)), rng.uniform(-0.3, 0.3, (B, 3)), rng.uniform(-0.5, 0.5, (B, 6))], 1))
acc = d(rng.uniform(-0.2, 0.2, (B, 6)))
for _ in range(5):
    s.ekf(thr, meas, acc)
torch.cuda.synchronize()
s.ekf_phase_cycles(reset=True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
K = 20
e0.record()
for _ in range(K):
    s.ekf(thr, meas, acc)
e1.record(); torch.cuda.synchronize()
pc = s.ekf_phase_cycles()
tot = sum(pc.values())
print(json.dumps({"B": B, "ekf_ms": e0.elapsed_time(e1) / K, "share": {k: round(v / max(tot, 1), 4) for k, v in pc.items()},
                  "cycles_per_instance": {k: round(v / B / K) for k, v in pc.items()}}, indent=1))
