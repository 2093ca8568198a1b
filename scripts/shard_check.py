#!/usr/bin/env python
"""scripts/shard_check.py -- run under torchrun on N GPUs: the peer-to-peer thrust exchange (PeerThrustExchange) against an NCCL
all-gather of the same blocks, over a few closed-loop ticks; prints one JSON line from rank 0."""
import json, os, sys
import numpy as np
import torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bluerov2_b200 import solver as S, workloads as wl
from bluerov2_b200.sharding import PeerThrustExchange

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
B, N, T = 512, 20, 6
w = wl.tracking_batch(B, N, seed=100 * rank, pos_spread=2.0)
sol = S.BatchSolver(B, N, device=local)
sol.set_trajectory(w["traj"]); sol.set_iterate(w["X"], w["U"])
ex = PeerThrustExchange(sol, B)
d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
x, lines, p = d(w["x0"]), d(w["lines"].astype(np.int32)), d(w["p"])
acc = torch.zeros((B, 6), dtype=torch.float64, device=dev)
out = (torch.empty((B, 4), dtype=torch.float64, device=dev), torch.empty((B, 6), dtype=torch.float64, device=dev),
       torch.empty((B,), dtype=torch.int32, device=dev))
worst, ok = 0.0, True
for t in range(T):
    sol.tick(x, p=p, lines=lines, body_acc=acc, out=out, plant_h=0.05)
    ex.wait()
    full = ex.result(sol.tick_count(), dev).clone()
    ref = torch.empty((world * B, 6), dtype=torch.float64, device=dev)
    dist.all_gather_into_tensor(ref, out[1])
    torch.cuda.synchronize()
    ok = ok and bool(torch.equal(full, ref)) and bool((out[2] == 0).all())
    dist.barrier()
res = torch.tensor([1.0 if ok else 0.0], device=dev)
dist.all_reduce(res, op=dist.ReduceOp.MIN)
if rank == 0:
    print(json.dumps({"world": world, "ticks": T, "peer_exchange_equals_nccl_all_gather": bool(res.item() == 1.0), "graphs": sol.graphs_built()}))
sol.close()
dist.destroy_process_group()
