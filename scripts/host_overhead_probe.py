"""Where the per-tick overhead of the host path goes (one B200): wall-clock per tick of
  A  tick_device, no synchronisation (the device-resident number)
  B  tick_device + stream synchronize every tick                      -> launch + completion latency
  C  tick_host, the same pinned buffers every tick                     -> + uploads, mapped outputs
  D  tick_host, a distinct input buffer per tick (what bench.py's e2e does) -> + graph upload-node updates
  E  as D, raw ctypes call on prebuilt br2_tick_io structs             -> minus the Python wrapper
  F  as D with option tick_graph = 0 (direct stream issue)
Usage: python scripts/host_overhead_probe.py [--batch 4096] [--steps 200] > gpurun_out/host_overhead.json"""
import argparse, ctypes as C, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--N", type=int, default=40)
    a = ap.parse_args()
    import torch
    import bench
    from bluerov2_b200 import solver as S, workloads as wl
    dev = torch.device("cuda:0")
    B, N, K, W = a.batch, a.N, a.steps, 10
    w = wl.tracking_batch(B, N, seed=0, reference="circle", pos_spread=0.3)
    out = {}

    def wall(fn, n):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for t in range(n):
            fn(t)
        torch.cuda.synchronize()
        return 1e6 * (time.perf_counter() - t0) / n

    sol = S.BatchSolver(B, N, device=0)
    sol.set_option("kernel_timing", 0)
    loop = bench.DeviceLoop(S, sol, w, dev)
    loop.restart()
    for t in range(W):
        loop.tick(t)
    out["A_device_nosync_us"] = wall(loop.tick, K)
    st = torch.cuda.current_stream(dev)
    out["B_device_sync_us"] = wall(lambda t: (loop.tick(t), st.synchronize()), K)
    # recorded closed-loop states so that the host ticks see the same sequence
    xs, ls = bench.record_states(torch, loop, W + K, dev) if hasattr(bench, "record_states") else (None, None)
    pin = lambda x: torch.from_numpy(np.ascontiguousarray(x)).pin_memory().numpy()   # noqa: E731
    h_x0 = [pin(x) for x in xs[:W + K]]
    h_l = [pin(l.astype(np.int32)) for l in ls[:W + K]]
    h_p = pin(w["p"])
    o = (pin(np.empty((B, 4))), pin(np.empty((B, 6))), pin(np.empty((B,), dtype=np.int32)))
    sol.set_iterate(w["X"], w["U"])
    sol.tick(h_x0[0], p=h_p, lines=h_l[0], out=o)                        # parameters supplied once: resident from here on
    same = lambda t: sol.tick(h_x0[0], lines=h_l[0], out=o)             # noqa: E731
    for t in range(W):
        same(t)
    out["C_host_same_buffers_us"] = wall(same, K)
    dist_ = lambda t: sol.tick(h_x0[t % (W + K)], lines=h_l[t % (W + K)], out=o)   # noqa: E731
    for t in range(W):
        dist_(t)
    sol.set_iterate(w["X"], w["U"])
    for t in range(W):
        dist_(t)
    out["D_host_distinct_us"] = wall(dist_, K)
    sol.set_iterate(w["X"], w["U"])
    for t in range(W):
        dist_(t)
    out["D_with_status_check_us"] = wall(lambda t: (dist_(t), bool((o[2] == 0).all())), K)
    # raw ctypes on prebuilt structs
    ios = []
    for t in range(W + K):
        io = S._TickIO(h_x0[t].ctypes.data, None, None, None, h_l[t].ctypes.data, None, o[0].ctypes.data, o[1].ctypes.data,
                       None, o[2].ctypes.data, None, None, 0.0, 0, 0, 1)
        ios.append((io, C.byref(io)))
    L, h = sol._L, sol._h
    raw = lambda t: L.br2_batch_tick_host(h, ios[t % (W + K)][1])      # noqa: E731
    sol.set_iterate(w["X"], w["U"])
    for t in range(W):
        raw(t)
    out["E_host_distinct_raw_ctypes_us"] = wall(raw, K)
    sol.set_iterate(w["X"], w["U"])
    # python wrapper alone (no GPU work): cost of the cache lookup etc.
    t0 = time.perf_counter()
    for t in range(K):
        args = (h_x0[t], h_p, None, h_l[t], None, None, o[0], o[1], o[2], None, None, None)
        key = tuple(id(x) for x in args) + (0, True, 0.0)
        sol._tick_cache.get(key)
    out["python_key_lookup_us"] = 1e6 * (time.perf_counter() - t0) / K
    dtb, okb, _, _ = bench.host_loop(torch, sol, w, xs, ls, 5, K, False, a.N)
    out["G_bench_host_loop_us"] = 1e6 * dtb / K
    dtb, okb, _, _ = bench.host_loop(torch, sol, w, xs, ls, 5, K, False, a.N)
    out["G_bench_host_loop_again_us"] = 1e6 * dtb / K
    sol.set_iterate(w["X"], w["U"])
    for t in range(W):
        dist_(t)
    out["D_again_us"] = wall(lambda t: (dist_(t), bool((o[2] == 0).all())), K)
    sol.set_option("tick_graph", 0)
    for t in range(W):
        dist_(t)
    out["F_host_distinct_no_graph_us"] = wall(dist_, K)
    sol.set_option("tick_graph", 1)
    out["graphs_built"] = sol.graphs_built()
    out["graph_updates"] = sol.graph_updates()
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
