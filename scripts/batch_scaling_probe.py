#!/usr/bin/env python
"""kernel times against the batch size (resident QP warps: 16 per SM = 2368): how much of the QP kernel is one instance's latency,
how much is contention between resident warps.  Prints one JSON line per batch size."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from bluerov2_b200 import solver as S, workloads as wl

dev = torch.device("cuda:0")
for B in [int(v) for v in (sys.argv[1:] or "148 592 1184 1776 2368 2960 3552 4096 4736 7104 8192".split())]:
    w = wl.tracking_batch(B, 40, seed=0, reference="circle", pos_spread=0.3)
    sol = S.BatchSolver(B, 40, device=0)
    loop = bench.DeviceLoop(S, sol, w, dev)
    sol.set_option("kernel_timing", 0)
    dt, itm, bad, _ = bench.timed_device_loop(torch, None, loop, 10, 50, dev, False)
    tl, tq = bench.kernel_times(torch, loop, 5, 30, dev)
    print(json.dumps({"B": B, "warps_per_scheduler": B / 592, "tick_ms": 1e3 * dt / 50, "lin_ms": 1e3 * tl, "qp_ms": 1e3 * tq,
                      "qp_us_per_1k_instances": 1e6 * tq / B * 1000 / 1000, "its": itm, "bad": bad}), flush=True)
    sol.close()
