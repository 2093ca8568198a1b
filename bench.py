#!/usr/bin/env python
"""bench.py -- SQP-RTI steps/sec of the batched BlueROV2 OCP (BASELINE.json metric) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B] [--horizon N]

A "step" is one control tick of the hot path over one batch: for every instance the call sequence of
BLUEROV2_DOB::solve (bluerov2_dob.cpp:307-395) -- x0, parameters, yref window, one SQP-RTI iteration (linearisation + QP), u0 and
the 4->6 thrust allocation.  Like the node, which keeps the trajectory file in memory and advances a line counter
(bluerov2_dob.cpp:367), the solver holds the trajectory on the device and a tick names one row per instance (ref_cb).

Headline workload = BASELINE config 2 (batch 4096 per GPU -- 8192 per GPU at 8 GPUs = config 4 --, random x0 around the circle
reference, N = 40, fp64), closed loop with the nominal ERK4 plant at 0.05 s, iterate carried between ticks.

  value : whole-job steps/s, device-resident closed loop: one br2_batch_tick_device per tick (lineariser -> QP kernels -> plant
          step on the device state, replayed as one CUDA graph), CUDA events on the launching stream, max over ranks.
  e2e   : the same closed-loop ticks through the C-ABI entry point br2_batch_tick_host(solver, &io) with pinned HOST buffers (what a
          C / C++ caller such as the reference's nodes does; br2_tick_io structs filled before the loop): H2D of x0 / row indices and
          D2H of u0 / thrust / status inside the timed region, a distinct input buffer per tick, the status words read on the host
          every tick; the OCP parameters of config 2 do not change and are supplied once (update_params semantics).  The same ticks
          through the Python wrapper (BatchSolver.tick) are reported beside it (e2e.through_python_wrapper).
  roofline / cpu_baseline / sub-records (forced interior-point iteration, saturated start, config 3, config 5, explicit-yref
  host path): see DESIGN.md "Measurement".

--impl reference times the reference's CPU algorithm for the same metric: acados/HPIPM cannot be built here, so it is the oracle
port (oracle/bluerov2_oracle.c, ERK routed through the reference's own CasADi-generated VDE from oracle/_ref when that was
built), OpenMP over instances on all host cores.  That leg and the cpu_baseline leg are the only places this file touches oracle/.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from bluerov2_b200 import traj, workloads as wl  # noqa: E402

METRIC = "SQP-RTI steps/sec (batched 6-DOF OCP, N=40)"
UNIT = "steps/s"
BYTES_PER_STAGE_ITER = 4384          # 548 doubles per instance, stage and QP iteration (factorisation + solve)


def config_of(B: int, world: int, N: int, pos_spread: float) -> dict:
    """the workload description BOTH arms print (identical keys and strings)"""
    return {"workload": f"config {4 if world * B == 65536 else 2}: batch {B} per GPU, random x0 around the circle reference (pos spread "
                        f"{pos_spread} m), N={N}, Ts={1.0 / N:g} s, fp64, closed loop (nominal ERK4 plant at 0.05 s), iterate carried "
                        f"between ticks, seed 0",
            "batch_per_gpu": B, "global_batch": world * B, "horizon": N, "pos_spread": pos_spread}


def default_batch(world: int) -> int:
    return 8192 if world >= 8 else 4096      # config 4 = 65536 = 8 x 8192; configs 2 / 3 = 4096 on one GPU


def measured_traffic(kind: str, B: int, N: int):
    """DRAM bytes per launch from the committed ncu capture of the same configuration (profiles/ipm_traffic.json), else None"""
    try:
        with open(os.path.join(ROOT, "profiles", "ipm_traffic.json")) as f:
            for e in json.load(f)["entries"]:
                if e.get("kernel", "ipm_kernel") == kind and e["batch"] == B and e["horizon"] == N:
                    return float(e["dram_bytes_per_launch"]), e["source"]
    except Exception:
        pass
    return None, None


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """SM clock and throttle reasons during the timed region (NVML, 20 ms period)."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception:
            pass

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                 "hw_power_brake_slowdown": 0x80, "sync_boost": 0x10, "applications_clocks_setting": 0x2}
        while not self._stop_evt.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                r = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for n, bit in names.items():
                    if r & bit:
                        self.reasons.add(n)
            except Exception:
                pass
            self._stop_evt.wait(0.02)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        med = int(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ---------------------------------------------------------------------------------------------------------
def cpu_leg(N: int, budget_s: float, ticks_wanted: int, threads: int = 0, seed: int = 0, pos_spread: float = 0.5, dense: bool = False):
    """the oracle port timed on the host cores over a bounded closed-loop sample of the same workload.
    dense: the QP by full condensing + a dense interior-point iteration instead of the Riccati recursion -- the cost profile of the
    reference's FULL_CONDENSING_HPIPM (plain C loops, so reported beside the Riccati arm, never instead of it)"""
    from oracle import Oracle, CasadiRef
    from oracle.oracle import REF_SO
    o = Oracle()
    o.set_qp_mode(1 if dense else 0)
    kind_note = ("oracle port (oracle/bluerov2_oracle.c, full condensing + dense IPM)" if dense
                 else "oracle port (oracle/bluerov2_oracle.c, Riccati IPM)")
    if os.path.exists(REF_SO):
        o.use_casadi(CasadiRef())
        kind_note += " with the ERK driven by the reference's CasADi-generated bluerov2_expl_vde_forw (oracle/_ref)"
    Ts = wl.time_steps(N)
    # all host cores this process may run on (torchrun exports OMP_NUM_THREADS=1: ask for the threads explicitly)
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    if threads <= 0:
        threads = cores
    # calibrate on 2 x cores instances, then size the sample to the budget
    wcal = wl.tracking_batch(2 * cores, N, seed=seed, pos_spread=pos_spread)
    X, U = wcal["X"].copy(), wcal["U"].copy()
    o.rti_step_batch(Ts, wcal["x0"], wcal["yref"], wcal["p"], X.copy(), U.copy(), nthreads=threads)   # thread-pool warm-up
    t0 = time.perf_counter()
    _, _, used = o.rti_step_batch(Ts, wcal["x0"], wcal["yref"], wcal["p"], X, U, nthreads=threads)
    per_inst = (time.perf_counter() - t0) / (2 * cores)
    Bs = int(max(used, min(4096, budget_s / max(per_inst, 1e-6) / max(ticks_wanted, 1))))
    Bs = max(used, (Bs // used) * used)
    w = wl.tracking_batch(Bs, N, seed=seed, pos_spread=pos_spread)
    X, U = w["X"].copy(), w["U"].copy()
    x0, lines = w["x0"].copy(), w["lines"].copy()
    times, iters = [], []
    for _ in range(ticks_wanted):
        yref = traj.window_batch(w["traj"], lines, N)
        t0 = time.perf_counter()
        st, info, used = o.rti_step_batch(Ts, x0, yref, w["p"], X, U, nthreads=threads)
        times.append(time.perf_counter() - t0)
        iters.append(float(info[:, 0].mean()))
        x0 = wl.plant_step(x0, U[:, 0].copy(), w["p"], 0.05)
        lines = lines + 1
    o.set_qp_mode(0)
    return dict(batch=Bs, times=times, iters=iters, cores=int(used), note=kind_note)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    K, W, N = args.steps, args.warmup, args.horizon
    B = args.batch or default_batch(world)
    r = cpu_leg(N, budget_s=args.cpu_budget, ticks_wanted=K + W, seed=0, pos_spread=args.pos_spread)
    t = float(np.sum(r["times"][W:]))
    value = r["batch"] * K / t
    sample = (f"{r['batch']} instances x {K} closed-loop ticks (after {W} warm-up ticks): the first {r['batch']} instances of the "
              f"workload (same generator and seed as the GPU arm); {r['note']}")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": K, "warmup": W,
        "ms_per_step": 1e3 * t / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "gpu_launches": 0,
        "config": config_of(B, world, N, args.pos_spread),
        "mean_ipm_iterations": float(np.mean(r["iters"][W:])),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": sample,
                         "algorithm": "interior-point iteration on every instance (no interior / active-set shortcut): compare with "
                                      "the GPU arm's sub-record forced_ipm for like-for-like, with its headline for the product path"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "acados/HPIPM/BLASFEO are un-vendored and not installable here; this is the acados-equivalent restatement "
                "(same RTI step, Riccati IPM -- ~10x fewer flops than the reference's full condensing + dense HPIPM).",
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------
class DeviceLoop:
    """device-resident closed loop: state, row counters and parameters stay in fixed device buffers, one tick() per control tick
    (solve + plant step in place), thrusts written straight into the rank's slot of the tick's gather buffer"""

    def __init__(self, S, sol, w, dev, world=1):
        import torch
        from bluerov2_b200.sharding import PipelinedThrustGather
        self.torch, self.sol, self.w, self.dev = torch, sol, w, dev
        B = w["x0"].shape[0]
        self.B = B
        d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)   # noqa: E731
        self.x_init, self.lines_init = d(w["x0"]), d(w["lines"].astype(np.int32))
        self.x, self.lines, self.p = self.x_init.clone(), self.lines_init.clone(), d(w["p"])
        self.acc = torch.zeros((B, 6), dtype=torch.float64, device=dev)
        self.u0 = torch.empty((B, 4), dtype=torch.float64, device=dev)
        self.st = torch.empty((B,), dtype=torch.int32, device=dev)
        # Every rank ends a tick with all ranks' thrust vectors.  Default: peer-to-peer, fused into the solve (the QP epilogue stores
        # into every rank's gather buffer over NVLink, PeerThrustExchange) -- no collective kernel on the tick's path.
        # BR2_GATHER=nccl: one NCCL all-gather per tick instead (double-buffered: it overlaps the next tick's lineariser).
        self.distributed = world > 1
        self.mode = os.environ.get("BR2_GATHER", "peer") if self.distributed else "none"
        self.gather = PipelinedThrustGather(world * B, dev, depth=int(os.environ.get("BR2_GATHER_DEPTH", "2")))
        self.thr = torch.empty((B, 6), dtype=torch.float64, device=dev)
        self.peer, self.gather_note = None, None
        if self.mode == "peer":
            from bluerov2_b200.sharding import PeerThrustExchange, PeerExchangeUnavailable
            try:
                self.peer = PeerThrustExchange(sol, B)
            except PeerExchangeUnavailable as e:        # raised on every rank together: all of them run the NCCL all-gather instead
                self.mode, self.gather_note = "nccl", str(e)
        sol.set_trajectory(w["traj"])

    def restart(self):
        import torch.distributed as dist
        self.sol.set_iterate(self.w["X"], self.w["U"])
        self.x.copy_(self.x_init); self.lines.copy_(self.lines_init)
        if self.peer is not None:
            self.torch.cuda.synchronize(self.dev); dist.barrier()      # every rank's stores have landed before the flags restart
        self.sol.set_tick_index(0)
        if self.peer is not None:
            dist.barrier()

    def tick(self, t):
        if self.mode == "nccl":
            self.sol.tick(self.x, p=self.p, lines=self.lines, body_acc=self.acc, out=(self.u0, self.gather.slot(t), self.st), plant_h=0.05)
            self.gather.all_gather_async(t)     # ONE all-gather per tick, enqueued behind the solve
        else:
            self.sol.tick(self.x, p=self.p, lines=self.lines, body_acc=self.acc, out=(self.u0, self.thr, self.st), plant_h=0.05)

    def finish(self):
        if self.mode == "nccl":
            self.gather.wait_all()              # the last ticks' collectives belong to the timed region
        elif self.peer is not None:
            self.peer.wait()                    # ... and so does the arrival of every rank's last tick


def timed_device_loop(torch, dist, loop, W, K, dev, distributed, sampler=None, per_tick=False):
    """W warm-up ticks, then K ticks bracketed by barrier + synchronize, CUDA events on the launching stream.
    Returns (seconds, mean QP iterations per instance and tick, non-zero statuses, [per-tick ms])"""
    sol = loop.sol
    stream = torch.cuda.current_stream(dev)

    def barrier():
        loop.finish()
        if distributed:
            dist.barrier()
        torch.cuda.synchronize(dev)

    loop.restart()
    for t in range(W):
        loop.tick(t)
    barrier()
    sol.ipm_iterations_total(reset=True); sol.nonzero_status_total(reset=True)
    n_ev = K + 1 if per_tick else 2
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(n_ev)]
    if sampler is not None:
        sampler.start()
    ev[0].record(stream)
    for t in range(W, W + K):
        loop.tick(t)
        if per_tick:
            ev[t - W + 1].record(stream)
    loop.finish()
    if not per_tick:
        ev[1].record(stream)
    barrier()
    dt = ev[0].elapsed_time(ev[-1]) * 1e-3
    ticks_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(K)] if per_tick else None
    return dt, sol.ipm_iterations_total(reset=True) / (loop.B * K), sol.nonzero_status_total(reset=True), ticks_ms


def kernel_times(torch, loop, W, K, dev):
    """per-kernel device times (CUDA events recorded by the library on the launching stream between its kernels): the timed
    ticks replayed once more with option kernel_timing on, reading the events after each tick (outside the steps/s measurement)"""
    sol = loop.sol
    sol.set_option("kernel_timing", 1)
    loop.restart()
    t_lin, t_qp = [], []
    for t in range(W + K):
        loop.tick(t)
        if t >= W:
            torch.cuda.synchronize(dev)
            a, b = sol.last_kernel_times()
            t_lin.append(a); t_qp.append(b)
    loop.finish()
    sol.set_option("kernel_timing", 0)
    return float(np.mean(t_lin)), float(np.mean(t_qp))


def record_states(torch, loop, T, dev):
    """untimed: the closed loop's per-tick inputs (x0, row index) as host arrays, for the host-API loops"""
    loop.restart()
    xs, ls = [], []
    for t in range(T):
        xs.append(loop.x.cpu().numpy().copy()); ls.append(loop.lines.cpu().numpy().copy())
        loop.tick(t)
    loop.finish()
    torch.cuda.synchronize(dev)
    return xs, ls


def roofline_of(B, N, iters_mean, t_qp, kernel):
    peak, peak_src = peaks()
    alg = B * BYTES_PER_STAGE_ITER * N * iters_mean
    achieved = alg / t_qp / 1e9
    traffic, src = measured_traffic(kernel, B, N)
    return {"kernel": kernel, "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
            "traffic_source": src, "algorithmic_bytes": alg, "peak_source": peak_src,
            "formula": f"B*{BYTES_PER_STAGE_ITER}*N*mean_qp_iterations / t_qp"}


def host_loop(torch, sol, w, xs, ls, W, K, explicit_yref, N, announce_next=False, write_combined=False, c_abi=False):
    """closed-loop ticks through the host API with pinned host buffers, a distinct input buffer per tick; wall clock around K ticks.
    explicit_yref: upload the (N+1) x 16 reference window per instance like ocp_nlp_cost_model_set("yref") x (N+1) does
    (bluerov2_dob.cpp:370-372) instead of naming a trajectory row.
    c_abi: the ticks go through the C-ABI entry point itself -- br2_batch_tick_host(solver, &io) on br2_tick_io structs filled before
    the loop, what a C / C++ caller such as the reference's nodes does -- instead of through the Python wrapper around it"""
    from bluerov2_b200 import solver as S_mod
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()   # noqa: E731
    B = w["x0"].shape[0]
    h_x0 = [pin(a) for a in xs[:W + K]]
    h_p = pin(w["p"])
    out = (pin(np.empty((B, 4))), pin(np.empty((B, 6))), pin(np.empty((B,), dtype=np.int32)))
    if explicit_yref:
        def window(l):
            # write-combined pinned memory when asked for: the host only writes these 21.5 MB, the copy engine reads them
            a = traj.window_batch(w["traj"], l.astype(np.int64), N)
            if not write_combined:
                return pin(a)
            b = S_mod.pinned_empty(a.shape, np.float64, write_combined=True)
            b[...] = a
            return b
        h_ref = [window(l) for l in ls[:W + K]]
        if announce_next:
            # the reference window of tick t + 1 is known while tick t runs (the measurement is not): registered before the call, it is
            # uploaded beside tick t's kernels (br2_batch_set_next_yref_host)
            def call(t):
                if t + 1 < W + K:
                    sol.set_next_yref(h_ref[t + 1])
                return sol.tick(h_x0[t], p=h_p, yref=h_ref[t], out=out)
        else:
            call = lambda t: sol.tick(h_x0[t], p=h_p, yref=h_ref[t], out=out)   # noqa: E731
        h2d = (B * 12 + B * 16 + B * (N + 1) * 16) * 8
    else:
        h_ref = [pin(l.astype(np.int32)) for l in ls[:W + K]]
        # config 2 is plain MPC: the OCP parameters do not change from tick to tick.  They are supplied with the first call and persist in the
        # solver (as in the reference, whose nodes call bluerov2_acados_update_params only when a parameter changes); every tick uploads the
        # measured state and the trajectory row index
        call = lambda t: sol.tick(h_x0[t], p=h_p if t == 0 else None, lines=h_ref[t], out=out)  # noqa: E731
        h2d = B * 12 * 8 + B * 4                 # x0 (fp64) and one trajectory row index per instance (int32)
        if c_abi:
            import ctypes as C
            ios = [S_mod._TickIO(h_x0[t].ctypes.data, None, h_p.ctypes.data if t == 0 else None, None, h_ref[t].ctypes.data, None,
                                 out[0].ctypes.data, out[1].ctypes.data, None, out[2].ctypes.data, None, None, 0.0, 0, 0, 1) for t in range(W + K)]
            refs = [C.byref(io) for io in ios]
            tick_host, handle = sol._L.br2_batch_tick_host, sol._h

            def call(t):                        # noqa: F811
                rc = tick_host(handle, refs[t])
                if rc:
                    sol._check(rc)
    sol.set_iterate(w["X"], w["U"])
    ok = True
    for t in range(W):
        call(t)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for t in range(W, W + K):
        call(t)
        ok = ok and not out[2].any()            # the status words of this tick, read on the host
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    ok = ok and bool(np.isfinite(out[0]).all())
    return dt, ok, h2d, (B * 4 + B * 6) * 8 + B * 4


# ---------------------------------------------------------------------------------------------------------
def sub_records(torch, S, dev, args, N, cpu):
    """the regimes the headline does not time (rank 0, one GPU): nested records of the one JSON line"""
    out = {}
    W, K = 5, 30

    def solver_for(B, Nh, spread, **opts):
        w = wl.tracking_batch(B, Nh, seed=0, reference="circle", pos_spread=spread)
        sol = S.BatchSolver(B, Nh, device=dev.index)
        sol.set_option("kernel_timing", 0)
        for k, v in opts.items():
            sol.set_option(k, v)
        return sol, w, DeviceLoop(S, sol, w, dev)

    def _forced_ipm():
        # ---- forced interior-point iteration (the reference's algorithm on every instance): like for like with the CPU arm ----
        sol, w, loop = solver_for(args.batch_sub, N, args.pos_spread, fast_path=0)
        dt, itm, bad, _ = timed_device_loop(torch, None, loop, W, K, dev, False)
        tl, tq = kernel_times(torch, loop, W, K, dev)
        out["forced_ipm"] = {"value": args.batch_sub * K / dt, "unit": UNIT, "ms_per_step": 1e3 * dt / K, "mean_ipm_iterations": itm,
                             "nonzero_status": bad, "kernels": {"linearize_ms": 1e3 * tl, "qp_ms": 1e3 * tq},
                             "roofline": roofline_of(args.batch_sub, N, itm, tq, "ipm_kernel"),
                             "what": "option fast_path = 0: Mehrotra predictor-corrector Riccati IPM on every instance, cold-started every tick"}
        sol.close()

    def _saturated_start():
        # ---- saturated start: 3 m position spread, thrusters saturated during the first ticks; timed FROM TICK 0 ----
        sol, w, loop = solver_for(args.batch_sub, N, 3.0)
        timed_device_loop(torch, None, loop, 0, 8, dev, False)              # throw-away pass: graphs built, code paths warm
        dt, itm, bad, ticks = timed_device_loop(torch, None, loop, 0, 12, dev, False, per_tick=True)
        tq_ticks = []
        sol.set_option("kernel_timing", 1)
        loop.restart()
        for t in range(12):
            loop.tick(t); torch.cuda.synchronize(dev)
            tq_ticks.append(1e3 * sol.last_kernel_times()[1])
        it0, _ = sol.stats()
        out["saturated_start"] = {"value": args.batch_sub * 12 / dt, "unit": UNIT, "tick_ms": [round(x, 4) for x in ticks],
                                  "qp_ms_per_tick": [round(x, 4) for x in tq_ticks], "mean_qp_iterations": itm, "nonzero_status": bad,
                                  "what": "pos spread 3.0 m, ticks 0..11 timed from tick 0 (about a third of the instances start with "
                                          "saturated thrusters): interior fast path + primal-dual active-set iteration, IPM fallback"}
        sol.close()

    def _config5_horizon_sweep():
        # ---- config 5: horizon sweep at batch 8192 ----
        sweep = []
        for Nh in (10, 20, 40, 80):
            sol, w, loop = solver_for(8192, Nh, args.pos_spread)
            dt, itm, bad, _ = timed_device_loop(torch, None, loop, W, K, dev, False)
            tl, tq = kernel_times(torch, loop, W, K, dev)
            rf = roofline_of(8192, Nh, itm, tq, "pdas_kernel")
            sweep.append({"horizon": Nh, "value": 8192 * K / dt, "ms_per_step": 1e3 * dt / K, "linearize_ms": 1e3 * tl, "qp_ms": 1e3 * tq,
                          "roofline_frac": rf["frac"], "achieved_gbs": rf["achieved"], "nonzero_status": bad})
            sol.close()
        out["config5_horizon_sweep"] = {"batch": 8192, "unit": UNIT, "points": sweep}

    def _config3_dob():
        # ---- config 3: DOB-MPC (EKF -> parameters -> solve as ONE tick), sampled wave disturbances, lemniscate reference ----
        out["config3_dob"] = dob_record(torch, S, dev, args, N)

    # each record on its own: a failure is reported in its place, the others (and the headline line) stand
    for name, fn in (("forced_ipm", _forced_ipm), ("saturated_start", _saturated_start), ("config5_horizon_sweep", _config5_horizon_sweep),
                     ("config3_dob", _config3_dob)):
        try:
            fn()
        except Exception as e:       # noqa: BLE001
            out[name] = {"error": repr(e)[:300]}
    return out


def dob_record(torch, S, dev, args, N):
    """BASELINE config 3.  Inputs recorded from the settled closed loop driven by the UNCOMPENSATED command (the node's compensation
    gain 1/0.0325, bluerov2_dob.cpp:326-338, over-compensates ~30x on any plant whose thrust scale is the OCP model's, so the
    compensated command is computed -- the timed path -- but not fed back); the timed loop replays them through fixed device
    buffers (one device-to-device copy of the packed record per tick, inside the timed region) into br2_batch_tick_device(ekf = 1)."""
    B, W, K, settle = args.batch_sub, 5, 30, 60
    w = wl.tracking_batch(B, N, seed=0, reference="lemniscate", pos_spread=min(args.pos_spread, 0.2), level=True)
    sol = S.BatchSolver(B, N, device=dev.index)
    sol.set_option("kernel_timing", 0)
    sol.set_trajectory(w["traj"]); sol.set_iterate(w["X"], w["U"])
    amp, tau0 = wl.wave_disturbance(B, seed=1)
    x, lines = w["x0"].copy(), w["lines"].copy()
    vel_prev, thr = x[:, 6:12].copy(), np.zeros((B, 6))
    rec = {"thr": [], "meas": [], "acc": [], "lines": []}
    start = None
    for t in range(settle + W + K):
        if t == settle:
            X0, U0 = sol.get_iterate()
            sol.ekf_reset()
            ex, eP = sol.ekf_state()
            ex[:, :12] = x                  # filter starts at the true pose
            start = (X0, U0, ex, eP)
        if t >= settle:
            rec["thr"].append(thr.copy()); rec["meas"].append(x.copy())
            rec["acc"].append((x[:, 6:12] - vel_prev) / 0.05); rec["lines"].append(lines.astype(np.int32))
        vel_prev = x[:, 6:12].copy()
        u0, th, st = sol.solve_windowed(x, lines.astype(np.int32), w["p"])
        if (st != 0).any():
            return {"error": f"solver status != 0 while recording the DOB workload at tick {t}"}
        thr = th.copy()                                   # thrust feedback = previous command
        x = wl.plant_step(x, u0, w["p"], 0.05, dist=wl.wave_at(amp, tau0, t))
        lines = lines + 1
    # one packed record per tick [meas B x 12 | thr B x 6 | acc B x 6 | row index B (int32, in B/2 doubles)] so that "new sensor data
    # arrives" is ONE device-to-device copy per tick into the fixed buffers the tick graph reads
    off = np.cumsum([0, B * 12, B * 6, B * 6, (B + 1) // 2])

    def packed(t):
        a = np.zeros(off[-1])
        a[off[0]:off[1]], a[off[1]:off[2]], a[off[2]:off[3]] = rec["meas"][t].ravel(), rec["thr"][t].ravel(), rec["acc"][t].ravel()
        a[off[3]:].view(np.int32)[:B] = rec["lines"][t]
        return torch.from_numpy(a).to(dev)
    src = [packed(t) for t in range(W + K)]
    buf = torch.empty_like(src[0])
    bx, bt, ba = buf[off[0]:off[1]].view(B, 12), buf[off[1]:off[2]].view(B, 6), buf[off[2]:off[3]].view(B, 6)
    bl = buf[off[3]:].view(torch.int32)[:B]
    out = (torch.empty((B, 4), dtype=torch.float64, device=dev), torch.empty((B, 6), dtype=torch.float64, device=dev),
           torch.empty((B,), dtype=torch.int32, device=dev))
    wf = torch.empty((B, 6), dtype=torch.float64, device=dev)

    def tick(t):
        buf.copy_(src[t])
        sol.tick(bx, lines=bl, thrusts=bt, body_acc=ba, ekf=1, compensate=True, out=out, wf_dist=wf)

    def restart():
        sol.set_iterate(start[0], start[1]); sol.set_ekf_state(start[2], start[3])

    stream = torch.cuda.current_stream(dev)
    restart()
    for t in range(W):
        tick(t)
    torch.cuda.synchronize(dev)
    sol.ipm_iterations_total(reset=True); sol.nonzero_status_total(reset=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for t in range(W, W + K):
        tick(t)
    e1.record(stream)
    torch.cuda.synchronize(dev)
    dt = e0.elapsed_time(e1) * 1e-3
    itm, bad = sol.ipm_iterations_total(reset=True) / (B * K), sol.nonzero_status_total(reset=True)
    # host path: the tick through br2_batch_tick_host (EKF inputs up, u0 / thrust / status / disturbance down)
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()   # noqa: E731
    hsrc = [tuple(pin(rec[k][t]) for k in ("meas", "thr", "acc", "lines")) for t in range(W + K)]
    hout = (pin(np.empty((B, 4))), pin(np.empty((B, 6))), pin(np.empty((B,), dtype=np.int32)))
    hwf = pin(np.empty((B, 6)))
    restart()
    for t in range(W):
        sol.tick(hsrc[t][0], lines=hsrc[t][3], thrusts=hsrc[t][1], body_acc=hsrc[t][2], ekf=1, out=hout, wf_dist=hwf)
    t0 = time.perf_counter()
    for t in range(W, W + K):
        sol.tick(hsrc[t][0], lines=hsrc[t][3], thrusts=hsrc[t][1], body_acc=hsrc[t][2], ekf=1, out=hout, wf_dist=hwf)
    dth = time.perf_counter() - t0
    sol.close()
    return {"value": B * K / dt, "unit": UNIT, "ms_per_step": 1e3 * dt / K, "mean_qp_iterations": itm, "nonzero_status": bad,
            "gpu_launches_per_tick": 4, "kernels": ["ekf_kernel", "linearize_kernel", "pdas_kernel", "ipm_kernel"],
            "e2e": {"value": B * K / dth, "unit": UNIT, "ms_per_step": 1e3 * dth / K,
                    "h2d_bytes_per_step": (B * 12 + B * 6 + B * 6) * 8 + B * 4, "d2h_bytes_per_step": (B * 4 + B * 6 + B * 6) * 8 + B * 4},
            "what": f"batch {B} DOB-MPC: 18-state EKF -> OCP parameters -> RTI solve as one br2_batch_tick_device (one CUDA graph), sampled wave "
                    f"disturbances, lemniscate reference, N={N}; inputs recorded from the settled closed loop (see dob_record)"}


# ---------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from bluerov2_b200 import solver as S

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the solver has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    distributed = world > 1
    if distributed:
        dist.init_process_group("nccl", device_id=dev)
    K, W, N = args.steps, args.warmup, args.horizon
    B = args.batch or default_batch(world)
    args.batch_sub = args.batch or 4096
    if W < 3:
        raise SystemExit("--warmup must be >= 3")

    sampler = ClockSampler(local)
    # every rank draws the same generator with its own seed offset: shards of one job
    w = wl.tracking_batch(B, N, seed=1000 * rank, reference="circle", pos_spread=args.pos_spread)
    sol = S.BatchSolver(B, N, device=local)
    sol.set_option("kernel_timing", 0)
    if args.no_fast_path:
        sol.set_option("fast_path", 0)
    if args.active_set_path is not None:
        sol.set_option("active_set_path", args.active_set_path)
    loop = DeviceLoop(S, sol, w, dev, world)

    dt, iters_mean, n_bad, _ = timed_device_loop(torch, dist, loop, W, K, dev, distributed, sampler)
    clocks = sampler.stop()
    t_lin, t_qp = kernel_times(torch, loop, W, K, dev)
    graphs = sol.graphs_built()

    # ---- e2e: the same closed-loop ticks through the host API ----
    # the host loops warm up for at least 20 ticks (W is the contract's minimum): a synchronous host tick is sensitive to cold code paths
    # and page-locked buffers in a way the enqueue-only device loop is not, and K can be as small as 20
    Wh = max(W, 20)
    xs, ls = record_states(torch, loop, Wh + K, dev)
    # e2e = the C-ABI call with host buffers (what the reference's C++ nodes would make); the same ticks through the Python wrapper beside it
    # throw-away pass of the same length: host graphs built, code paths warm, and the pinned input buffers of the timed passes come out of
    # torch's pinned-memory cache instead of fresh cudaHostAlloc calls (first-use page-locking shows up as ~2 % on the tick)
    host_loop(torch, sol, w, xs, ls, Wh, K, False, N, c_abi=True)
    dt_e2e, e2e_ok, h2d, d2h = host_loop(torch, sol, w, xs, ls, Wh, K, False, N, c_abi=True)
    dt_py, py_ok, _, _ = host_loop(torch, sol, w, xs, ls, Wh, K, False, N)
    dt_exp = dt_ann = None
    extra_errors = {}
    if world == 1 and not args.quick:
        # the secondary host-path legs must not take the headline line down with them (pinned-memory limits of the box, ...)
        try:
            dt_exp, exp_ok, h2d_exp, _ = host_loop(torch, sol, w, xs, ls, Wh, min(K, 50), True, N)
        except Exception as e:       # noqa: BLE001
            extra_errors["e2e_explicit_yref"] = repr(e)[:300]
        try:
            dt_ann, ann_ok, _, _ = host_loop(torch, sol, w, xs, ls, Wh, min(K, 50), True, N, announce_next=True, write_combined=True)
        except Exception as e:       # noqa: BLE001
            extra_errors["announced_one_tick_ahead"] = repr(e)[:300]

    by_rank = None
    if distributed:
        # every rank's own device time beside the max: names what the slowest rank is (chip-to-chip variation or the exchange)
        mine = torch.tensor([1e3 * dt / K, 1e3 * t_lin, 1e3 * t_qp], dtype=torch.float64, device=dev)
        allr = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        by_rank = {k: [round(float(a[i]), 5) for a in allr] for i, k in enumerate(("ms_per_step", "lineariser_ms", "qp_ms"))}
        tt = torch.tensor([dt, dt_e2e, dt_py], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt, dt_e2e, dt_py = float(tt[0]), float(tt[1]), float(tt[2])
        cnt = torch.tensor([float(iters_mean), float(n_bad)], dtype=torch.float64, device=dev)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
        iters_mean, n_bad = float(cnt[0]) / world, int(cnt[1])

    if rank == 0:
        line = {
            "metric": METRIC, "value": world * B * K / dt, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": 1e3 * dt / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": config_of(B, world, N, args.pos_spread),
            "mean_qp_iterations": iters_mean, "nonzero_status": n_bad, "fast_path": not args.no_fast_path,
            "l2": "per-tick working set (stage records + iterates) "
                  f"{B * (N + 1) * 352 * 8 / 1e6:.0f} MB > 126 MB L2",
            "parallelism": f"{world} x independent shards" + ("" if not distributed else
                                                              ", thrust vectors exchanged peer-to-peer from the QP epilogue (stores into every rank's gather "
                                                              "buffer over NVLink, one flag per rank and tick; no collective kernel)" if loop.mode == "peer" else
                                                              ", NO exchange of the thrust vectors (BR2_GATHER=none: diagnostic)" if loop.mode == "none" else
                                                              ", one NCCL all-gather of the thrust vectors per tick (double-buffered: it overlaps the next "
                                                              "tick's lineariser)") +
                           (f" [{loop.gather_note}]" if loop.gather_note else ""),
            "e2e": {"value": world * B * K / dt_e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ok": e2e_ok,
                    "ms_per_step": 1e3 * dt_e2e / K, "api": "br2_batch_tick_host(solver, &io) -- the C-ABI entry point on br2_tick_io structs filled before "
                                                                 "the loop, as a C / C++ caller holds them (windowed reference: one trajectory row index per "
                                                                 "instance; parameters supplied once, resident in the solver); status words read on the host "
                                                                 "every tick",
                    "overhead_over_device_time": dt_e2e / dt - 1.0,
                    "through_python_wrapper": {"value": world * B * K / dt_py, "unit": UNIT, "ms_per_step": 1e3 * dt_py / K, "ok": py_ok,
                                               "overhead_over_device_time": dt_py / dt - 1.0,
                                               "api": "BatchSolver.tick(...) with a fresh pinned input buffer per tick"}},
            "gpu_launches": 4 * K,
            "gpu_launches_per_tick": {"count": 4, "kernels": ["linearize_kernel", "pdas_kernel", "ipm_kernel (fallback list, normally empty)",
                                                              "plant_kernel"], "graphs_instantiated": graphs,
                                      "how": "one cudaGraphLaunch per tick (br2_batch_tick_device)"},
            "kernels": {"linearize_ms": 1e3 * t_lin, "qp_ms": 1e3 * t_qp, "qp_share_of_step": t_qp / (dt / K)},
            "by_rank": by_rank,
            "roofline": roofline_of(B, N, iters_mean, t_qp, "ipm_kernel" if args.no_fast_path else "pdas_kernel"),
            "clocks": clocks,
        }
        if dt_exp is not None:
            Ke = min(K, 50)
            line["e2e_explicit_yref"] = {"value": B * Ke / dt_exp, "unit": UNIT, "ms_per_step": 1e3 * dt_exp / Ke, "h2d_bytes_per_step": h2d_exp,
                                         "d2h_bytes_per_step": d2h, "ok": exp_ok, "fraction_of_windowed_e2e": (B * Ke / dt_exp) / (B * K / dt_py),
                                         "api": "br2_batch_tick_host with the explicit (N+1) x 16 reference window per instance "
                                                "(ocp_nlp_cost_model_set \"yref\" x (N+1), bluerov2_dob.cpp:370-372)",
                                         }
            if dt_ann is not None:
                line["e2e_explicit_yref"]["announced_one_tick_ahead"] = {
                    "value": B * Ke / dt_ann, "unit": UNIT, "ms_per_step": 1e3 * dt_ann / Ke, "ok": ann_ok,
                    "h2d_bytes_per_step": h2d_exp, "fraction_of_windowed_e2e": (B * Ke / dt_ann) / (B * K / dt_py),
                    "api": "the same call, the NEXT tick's window registered before it (br2_batch_set_next_yref_host): "
                           "uploaded beside the running tick's kernels; x0 and p still go up when the tick is called; "
                           "windows in write-combined pinned memory (br2_host_alloc)"}
        if extra_errors:
            line["secondary_leg_errors"] = extra_errors
        if world == 1 and not args.no_cpu:
            r = cpu_leg(N, budget_s=args.cpu_budget, ticks_wanted=3 + 2, seed=0, pos_spread=args.pos_spread)
            tcpu = float(np.sum(r["times"][2:]))
            line["cpu_baseline"] = {"value": r["batch"] * 3 / tcpu, "unit": UNIT, "cores": r["cores"], "kind": "port",
                                    "sample": f"{r['batch']} instances x 3 closed-loop ticks (after 2 warm-up ticks) of the same workload; {r['note']}",
                                    "mean_ipm_iterations": float(np.mean(r["iters"][2:]))}
            try:
                rd = cpu_leg(N, budget_s=args.cpu_budget / 3, ticks_wanted=2 + 1, seed=0, pos_spread=args.pos_spread, dense=True)
                td = float(np.sum(rd["times"][1:]))
                line["cpu_baseline"]["dense_condensed"] = {
                    "value": rd["batch"] * 2 / td, "unit": UNIT, "cores": rd["cores"],
                    "sample": f"{rd['batch']} instances x 2 closed-loop ticks (after 1 warm-up tick); {rd['note']}",
                    "what": "the same RTI step with the QP solved the way the reference's configuration does (FULL_CONDENSING_HPIPM: states "
                            "eliminated, dense interior-point iteration on the N*nu inputs), in plain C without BLASFEO-class kernels"}
            except Exception as e:       # noqa: BLE001
                line["cpu_baseline"]["dense_condensed"] = {"error": repr(e)[:200]}
        if world == 1 and not args.quick:
            sol.close()
            try:
                sub = sub_records(torch, S, dev, args, N, line.get("cpu_baseline"))
            except Exception as e:       # noqa: BLE001  (the headline line is printed regardless)
                sub = {"error": repr(e)[:300]}
            line["sub_records"] = sub
            if "cpu_baseline" in line and "value" in sub.get("forced_ipm", {}):
                cb = line["cpu_baseline"]
                cb["like_for_like"] = {
                    "gpu_forced_ipm_over_cpu": sub["forced_ipm"]["value"] / cb["value"],
                    "gpu_default_over_cpu": line["value"] / cb["value"],
                    "note": "the CPU arm runs the interior-point iteration on every instance; forced_ipm is the GPU doing the same, the "
                            "headline additionally uses the exact interior / active-set shortcuts (same minimiser, fewer factorisations)"}
        print(json.dumps(line), flush=True)
    try:
        sol.close()
    except Exception:
        pass
    if distributed:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=None, help="instances per GPU (default 4096; 8192 at 8 GPUs = BASELINE config 4)")
    ap.add_argument("--horizon", type=int, default=40)
    ap.add_argument("--pos-spread", type=float, default=0.5)
    ap.add_argument("--cpu-budget", type=float, default=20.0, help="seconds of CPU work for the cpu_baseline / reference arm")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--quick", action="store_true", help="headline + e2e only (no sub-records)")
    ap.add_argument("--no-fast-path", action="store_true", help="always run the interior-point iteration (diagnostic)")
    ap.add_argument("--active-set-path", type=int, default=None, choices=[0, 1],
                    help="override the library default of the primal-dual active-set iteration (diagnostic)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
