#!/usr/bin/env python
"""bench.py -- SQP-RTI steps/sec of the batched BlueROV2 OCP (BASELINE.json metric) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B] [--horizon N]

A "step" is one control tick of the hot path over one batch: for every instance the call sequence of
BLUEROV2_DOB::solve (bluerov2_dob.cpp:307-395) -- x0, parameters, yref window, one SQP-RTI iteration
(linearisation + Riccati IPM), u0 and the 4->6 thrust allocation.  Like the node, which keeps the trajectory file in
memory and advances a line counter (bluerov2_dob.cpp:367), the solver holds the trajectory on the device and a tick
passes one row index per instance (reference windowing on the device, ref_cb).  Workload = BASELINE config 2 (batch 4096 per GPU,
random x0 around the circle reference, N = 40, fp64): closed loop, plant = nominal ERK4 at 0.05 s, iterate carried
between ticks.  The closed-loop input sequence (x0_t, yref_t) is generated ONCE before the timed region (untimed
pass through the same CUDA solver + a numpy plant) and then replayed from the same initial iterate, so the timed
ticks see exactly the warm-started problems of ticks W..W+K-1.

  value : whole-job steps/s, inputs resident in HBM, device time (CUDA events), max over ranks.
  e2e   : the same ticks through the public host API (BatchSolver.solve_windowed -> br2_batch_solve_windowed_host) with
          pinned HOST buffers: H2D of x0 / row indices / p and D2H of u0/thrust/status inside the timed region.
  roofline / cpu_baseline: see DESIGN.md "Measurement".

--impl reference times the reference's CPU algorithm for the same metric: acados/HPIPM cannot be built here, so it
is the oracle port (oracle/bluerov2_oracle.c, ERK routed through the reference's own CasADi-generated VDE from
oracle/_ref when that was built), OpenMP over instances on all host cores.  That leg and the cpu_baseline leg are
the only places this file touches oracle/.
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
BYTES_PER_STAGE_ITER = 4384          # 548 doubles per instance, stage and IPM iteration


def measured_traffic(B, N, fast_path):
    """DRAM bytes per ipm_kernel launch from the committed ncu capture of the same configuration, else None"""
    try:
        with open(os.path.join(ROOT, "profiles", "ipm_traffic.json")) as f:
            for e in json.load(f)["entries"]:
                if e["batch"] == B and e["horizon"] == N and bool(e["fast_path"]) == bool(fast_path):
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
def make_workload(B: int, N: int, seed: int, pos_spread: float):
    w = wl.tracking_batch(B, N, seed=seed, reference="circle", pos_spread=pos_spread)
    return w


def record_closed_loop(solver, w, ticks: int, N: int):
    """untimed: run `ticks` closed-loop ticks through the CUDA solver, return the per-tick inputs (host arrays)"""
    x0, lines = w["x0"].copy(), w["lines"].copy()
    solver.set_iterate(w["X"], w["U"])
    x0s, lns = [], []
    solver.set_trajectory(w["traj"])
    for _ in range(ticks):
        x0s.append(x0.copy()); lns.append(lines.astype(np.int32))
        u0, _, st = solver.solve_windowed(x0, lns[-1], w["p"])
        if (st != 0).any():
            raise RuntimeError(f"solver status != 0 while recording the workload: {np.unique(st, return_counts=True)}")
        x0 = wl.plant_step(x0, u0, w["p"], 0.05)
        lines = lines + 1
    return x0s, lns


def record_closed_loop_dob(solver, w, ticks: int, N: int, seed: int, settle: int = 60):
    """untimed, BASELINE config 3: inputs of DOB-MPC ticks (thruster forces, pose/velocity measurement, finite-differenced
    body acceleration, trajectory row) from a plant driven by sampled wave disturbances (mode 0 of applyBodyWrench,
    bluerov2_dob.cpp:774-797).  The loop is first settled for `settle` ticks (the filter of the reference is explicit
    RK4 at 50 ms on roll/pitch dynamics with |lambda dt| > 2.8: it only lives in the gentle regime |u| < 1), and the
    plant is driven by the UNCOMPENSATED command: the node's compensation gain 1/0.0325 (bluerov2_dob.cpp:326-338) over-
    compensates ~30x on any plant whose thrust scale is the OCP model's, so the compensated command is computed (timed
    path) but not fed back.  Returns the per-tick inputs and the state to restart the replay from."""
    B = w["x0"].shape[0]
    amp, tau0 = wl.wave_disturbance(B, seed=seed + 1)
    x, lines = w["x0"].copy(), w["lines"].copy()
    solver.set_trajectory(w["traj"])
    solver.set_iterate(w["X"], w["U"])
    vel_prev, thr = x[:, 6:12].copy(), np.zeros((B, 6))
    rec = {"thr": [], "meas": [], "acc": [], "lines": []}
    start = None
    for t in range(settle + ticks):
        if t == settle:
            X0, U0 = solver.get_iterate()
            solver.ekf_reset()
            ex, eP = solver.ekf_state()
            ex[:, :12] = x                  # filter starts at the true pose
            start = (X0, U0, ex, eP)
        if t >= settle:
            rec["thr"].append(thr.copy()); rec["meas"].append(x.copy())
            rec["acc"].append((x[:, 6:12] - vel_prev) / 0.05); rec["lines"].append(lines.astype(np.int32))
        vel_prev = x[:, 6:12].copy()
        u0, th, st = solver.solve_windowed(x, lines.astype(np.int32), w["p"])
        if (st != 0).any():
            raise RuntimeError(f"solver status != 0 while recording the DOB workload: {np.unique(st, return_counts=True)}")
        thr = th.copy()                                   # thrust feedback = previous command
        x = wl.plant_step(x, u0, w["p"], 0.05, dist=wl.wave_at(amp, tau0, t))
        lines = lines + 1
    return rec, start


def cpu_leg(N: int, budget_s: float, ticks_wanted: int, threads: int = 0, seed: int = 0, pos_spread: float = 0.5):
    """the oracle port timed on the host cores over a bounded closed-loop sample of the same workload"""
    from oracle import Oracle, CasadiRef
    from oracle.oracle import REF_SO
    o = Oracle()
    kind_note = "oracle port (oracle/bluerov2_oracle.c, Riccati IPM)"
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
    return dict(batch=Bs, times=times, iters=iters, cores=int(used), note=kind_note)


# ---------------------------------------------------------------------------------------------------------
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    K, W, N = args.steps, args.warmup, args.horizon
    r = cpu_leg(N, budget_s=args.cpu_budget, ticks_wanted=K + W, seed=0, pos_spread=args.pos_spread)
    t = float(np.sum(r["times"][W:]))
    value = r["batch"] * K / t
    sample = (f"{r['batch']} instances x {K} closed-loop ticks (after {W} warm-up ticks) of config 2 "
              f"(same generator and seed as the GPU arm, first {r['batch']} instances); {r['note']}")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": K, "warmup": W,
        "ms_per_step": 1e3 * t / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "gpu_launches": 0,
        "config": {"workload": f"config 2: random x0 around the circle reference, N={N}, Ts={1.0 / N:g} s, fp64; CPU arm on a "
                               f"bounded sample (batch {r['batch']})", "batch_per_step": r["batch"], "horizon": N,
                   "mean_ipm_iterations": float(np.mean(r["iters"][W:]))},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "acados/HPIPM/BLASFEO are un-vendored and not installable here; this is the acados-equivalent restatement "
                "(same RTI step, Riccati IPM -- ~10x fewer flops than the reference's full condensing + dense HPIPM).",
    }
    print(json.dumps(line), flush=True)


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
    K, W, N, B = args.steps, args.warmup, args.horizon, args.batch
    if W < 3:
        raise SystemExit("--warmup must be >= 3")

    sampler = ClockSampler(local)
    dob = args.workload == "dob"
    w = wl.tracking_batch(B, N, seed=1000 * rank, reference="lemniscate" if dob else "circle",
                          pos_spread=min(args.pos_spread, 0.2) if dob else args.pos_spread, level=dob)
    sol = S.BatchSolver(B, N, device=local)
    if args.no_fast_path:
        sol.set_option("fast_path", 0)
    if args.active_set_path is not None:
        sol.set_option("active_set_path", args.active_set_path)
    if dob:
        rec, dob0 = record_closed_loop_dob(sol, w, W + K, N, seed=1000 * rank)
        x0s, lns = rec["meas"], rec["lines"]
        d_thr = [torch.from_numpy(a).to(dev) for a in rec["thr"]]
        d_acc = [torch.from_numpy(a).to(dev) for a in rec["acc"]]
        ekf_out = (torch.empty((B, 6), dtype=torch.float64, device=dev), torch.empty((B, 16), dtype=torch.float64, device=dev))
    else:
        x0s, lns = record_closed_loop(sol, w, W + K, N)

    # ---- device-resident inputs, one distinct buffer per tick ----
    d_x0 = [torch.from_numpy(a).to(dev) for a in x0s]
    d_lines = [torch.from_numpy(a).to(dev) for a in lns]
    d_p = torch.from_numpy(w["p"]).to(dev)

    def restart():
        if dob:
            sol.set_iterate(dob0[0], dob0[1])
            sol.set_ekf_state(dob0[2], dob0[3])
        else:
            sol.set_iterate(w["X"], w["U"])
    from bluerov2_b200.sharding import PipelinedThrustGather
    # all ranks' thrust vectors, double-buffered: the all-gather of tick t overlaps the lineariser of tick t + 1
    gather = PipelinedThrustGather(world * B, dev, depth=int(os.environ.get("BR2_GATHER_DEPTH", "2")))
    out_u0, out_st = torch.empty((B, 4), dtype=torch.float64, device=dev), torch.empty((B,), dtype=torch.int32, device=dev)
    out = (out_u0, gather.slot(0), out_st)
    stream = torch.cuda.current_stream(dev)

    def tick(t):
        o = (out_u0, gather.slot(t), out_st)    # thrusts land directly in this rank's slot of the tick's gather buffer
        if dob:     # EKF writes the OCP parameters on the device; the solve reads them there (same stream, no host hop)
            sol.ekf(d_thr[t], d_x0[t], d_acc[t], compensate=True, out=ekf_out)
            sol.solve_windowed(d_x0[t], d_lines[t], ekf_out[1], out=o)
        else:
            sol.solve_windowed(d_x0[t], d_lines[t], d_p, out=o)
        if distributed:
            gather.all_gather_async(t)          # ONE all-gather per tick, enqueued behind the solve

    def barrier():
        gather.wait_all()                       # the last ticks' collectives belong to the timed region
        if distributed:
            dist.barrier()
        torch.cuda.synchronize(dev)

    restart()
    for t in range(W):
        tick(t)
    barrier()
    sol.ipm_iterations_total(reset=True)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    t_lin, t_ipm = [], []
    sampler.start()
    ev[0].record(stream)
    for t in range(W, W + K):
        tick(t)
    gather.wait_all()                           # the stream waits for the collectives still in flight before the end event
    ev[1].record(stream)
    barrier()
    clocks = sampler.stop()
    dt = ev[0].elapsed_time(ev[1]) * 1e-3
    iters_total = sol.ipm_iterations_total(reset=True)
    st = out[2].cpu().numpy()
    n_bad = int((st != 0).sum())

    # per-kernel device times (CUDA events recorded by the library on the launching stream around each kernel):
    # replay the timed ticks once more, reading the events after each tick (outside the steps/s measurement)
    restart()
    for t in range(W + K):
        tick(t)
        if t >= W:
            torch.cuda.synchronize(dev)
            a, b = sol.last_kernel_times()
            t_lin.append(a); t_ipm.append(b)

    # ---- e2e: host buffers through the public host API ----
    pin = lambda a: torch.from_numpy(a).pin_memory().numpy()  # noqa: E731
    h_x0 = [pin(a) for a in x0s]
    h_lines = [torch.from_numpy(a).pin_memory().numpy() for a in lns]
    h_p = pin(w["p"])
    h_out = (pin(np.empty((B, 4))), pin(np.empty((B, 6))), torch.empty((B,), dtype=torch.int32).pin_memory().numpy())
    if dob:
        h_thr = [pin(a) for a in rec["thr"]]
        h_acc = [pin(a) for a in rec["acc"]]
        h_ekf = (pin(np.empty((B, 6))), pin(np.empty((B, 16))))

    def host_tick(t):
        if dob:     # the two public host calls of a DOB-MPC tick; p makes the round trip through the host like in the node
            sol.ekf(h_thr[t], h_x0[t], h_acc[t], compensate=True, out=h_ekf)
            sol.solve_windowed(h_x0[t], h_lines[t], h_ekf[1], out=h_out)
        else:
            sol.solve_windowed(h_x0[t], h_lines[t], h_p, out=h_out)

    restart()
    for t in range(W):
        host_tick(t)
    barrier()
    t0 = time.perf_counter()
    for t in range(W, W + K):
        host_tick(t)
    barrier()
    dt_e2e = time.perf_counter() - t0
    e2e_ok = bool((h_out[2] == 0).all()) and bool(np.isfinite(h_out[0]).all())
    if dob:     # thrusts, measurement (= x0), body acceleration, row index up; disturbance, p, u0, thrust, status down; p up again
        h2d = (B * 6 + B * 12 + B * 6 + B * 16) * 8 + B * 4
        d2h = (B * 6 + B * 16 + B * 4 + B * 6) * 8 + B * 4
    else:
        h2d = (B * 12 + B * 16) * 8 + B * 4      # x0, p (fp64) and one trajectory row index per instance (int32)
        d2h = (B * 4 + B * 6) * 8 + B * 4

    if distributed:
        tt = torch.tensor([dt, dt_e2e], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt, dt_e2e = float(tt[0]), float(tt[1])
        cnt = torch.tensor([float(iters_total), float(n_bad)], dtype=torch.float64, device=dev)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
        iters_total, n_bad = float(cnt[0]), int(cnt[1])
        iters_mean = iters_total / (world * B * K)
    else:
        iters_mean = iters_total / (B * K)

    if rank == 0:
        peak, peak_src = peaks()
        t_ipm_avg = float(np.mean(t_ipm))
        alg_bytes = B * BYTES_PER_STAGE_ITER * N * iters_mean
        achieved = alg_bytes / t_ipm_avg / 1e9
        traffic, traffic_src = measured_traffic(B, N, not args.no_fast_path)
        line = {
            "metric": METRIC, "value": world * B * K / dt, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": 1e3 * dt / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": (f"config 3: batch {B} per GPU DOB-MPC (18-state EKF -> parameters -> RTI solve, back to back on "
                                    f"one stream), sampled wave disturbances, lemniscate reference, N={N}, fp64; inputs recorded from the settled "
                                    f"closed loop driven by the uncompensated command (see record_closed_loop_dob)"
                                    if dob else
                                    f"config 2: batch {B} per GPU, random x0 around the circle reference (pos spread "
                                    f"{args.pos_spread} m), N={N}, Ts={1.0 / N:g} s, fp64, closed loop (nominal ERK4 plant at 0.05 s), "
                                    f"iterate carried between ticks"), "batch_per_gpu": B, "global_batch": world * B, "horizon": N,
                       "mean_ipm_iterations": iters_mean, "nonzero_status": n_bad, "fast_path": not args.no_fast_path,
                       "l2": "per-tick working set (stage records + factors + iterates) "
                             f"{B * N * (208 + 64 + 64) * 8 / 1e6:.0f} MB > 126 MB L2; distinct input buffers per step",
                       "parallelism": f"{world} x independent shards" + (", one NCCL all-gather of the thrust vectors per tick (double-buffered: it overlaps the next tick's lineariser)" if distributed else "")},
            "e2e": {"value": world * B * K / dt_e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ok": e2e_ok,
                    "ms_per_step": 1e3 * dt_e2e / K},
            "gpu_launches": (3 if dob else 2) * K,
            "kernels": {"linearize_ms": 1e3 * float(np.mean(t_lin)), "ipm_ms": 1e3 * t_ipm_avg,
                        "ipm_share_of_step": t_ipm_avg / (dt / K)},
            "roofline": {"kernel": "ipm_kernel (Riccati sweeps)", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "algorithmic_bytes": alg_bytes,
                         "peak_source": peak_src,
                         "formula": f"B*{BYTES_PER_STAGE_ITER}*N*mean_ipm_iterations / t_ipm"},
            "clocks": clocks,
        }
        # ---- CPU baseline on this box's host cores (bounded sample) ----
        if world == 1 and not args.no_cpu and not dob:
            r = cpu_leg(N, budget_s=args.cpu_budget, ticks_wanted=3 + 2, seed=0, pos_spread=args.pos_spread)
            tcpu = float(np.sum(r["times"][2:]))
            line["cpu_baseline"] = {"value": r["batch"] * 3 / tcpu, "unit": UNIT, "cores": r["cores"], "kind": "port",
                                    "sample": f"{r['batch']} instances x 3 closed-loop ticks (after 2 warm-up ticks) of the same workload; {r['note']}",
                                    "mean_ipm_iterations": float(np.mean(r["iters"][2:]))}
        print(json.dumps(line), flush=True)
    sol.close()
    if distributed:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=4096, help="instances per GPU")
    ap.add_argument("--horizon", type=int, default=40)
    ap.add_argument("--pos-spread", type=float, default=0.5)
    ap.add_argument("--workload", default="tracking", choices=["tracking", "dob"],
                    help="tracking = BASELINE config 2 (default, the headline); dob = config 3 (EKF + solve per tick)")
    ap.add_argument("--cpu-budget", type=float, default=20.0, help="seconds of CPU work for the cpu_baseline / reference arm")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-fast-path", action="store_true", help="always run the interior-point iteration (diagnostic)")
    ap.add_argument("--active-set-path", type=int, default=None, choices=[0, 1],
                    help="override the library default of the active-set fast path (diagnostic)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
