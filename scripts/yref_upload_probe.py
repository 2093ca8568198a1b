#!/usr/bin/env python
"""scripts/yref_upload_probe.py -- the explicit-yref host tick: raw H2D rate of the reference windows, and the tick with the upload in
1 / 2 / 4 pipelined instance ranges (BR2_YREF_CHUNKS, read at first use: one process per setting), same buffers (graph) vs distinct."""
import json, os, subprocess, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

if len(sys.argv) > 1:
    import torch
    from bluerov2_b200 import solver as S, traj, workloads as wl
    B, N, K = 4096, 40, 40
    w = wl.tracking_batch(B, N, seed=0)
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
    s = S.BatchSolver(B, N); s.set_option("kernel_timing", 0)
    out = (pin(np.empty((B, 4))), pin(np.empty((B, 6))), pin(np.empty((B,), dtype=np.int32)))
    hp = pin(w["p"])
    ys = [pin(w["yref"]) for _ in range(K + 5)]
    xs = [pin(w["x0"]) for _ in range(K + 5)]
    res = {"chunks": os.environ.get("BR2_YREF_CHUNKS", "4")}
    # raw H2D
    d = torch.empty((B, N + 1, 16), dtype=torch.float64, device="cuda")
    t_ = torch.from_numpy(ys[0])
    for _ in range(3): d.copy_(t_, non_blocking=True)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(20): d.copy_(t_, non_blocking=True)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 20
    res["h2d_ms"] = 1e3 * dt; res["h2d_gbs"] = t_.numel() * 8 / dt / 1e9
    for mode in ("same_buffers", "distinct_buffers"):
        s.set_iterate(w["X"], w["U"])
        for t in range(5):
            i = 0 if mode == "same_buffers" else t
            s.tick(xs[i], p=hp, yref=ys[i], out=out)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for t in range(5, 5 + K):
            i = 0 if mode == "same_buffers" else t
            s.tick(xs[i], p=hp, yref=ys[i], out=out)
        torch.cuda.synchronize()
        res[mode + "_ms"] = 1e3 * (time.perf_counter() - t0) / K
    res["graphs"] = s.graphs_built()
    print(json.dumps(res))
else:
    for c in ("1", "2", "4"):
        e = dict(os.environ); e["BR2_YREF_CHUNKS"] = c
        r = subprocess.run([sys.executable, __file__, "run"], env=e, capture_output=True, text=True)
        print(r.stdout.strip() or r.stderr[-500:])
