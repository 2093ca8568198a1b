"""One control tick as one call / one CUDA graph (br2_batch_tick_device / _host == `EKF(); solve();`, the body of the node's loop,
bluerov2_dobmpc/src/bluerov2_dob_node.cpp:13-31): the graph replays must be indistinguishable from the kernel-by-kernel entry
points they bundle, and a steady closed loop must build exactly one graph."""
import numpy as np
import pytest

from bluerov2_b200 import traj, workloads as wl

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def solver_mod():
    from bluerov2_b200 import solver
    if solver.device_count() < 1:
        pytest.fail("no CUDA device visible: the gpu-marked tests must run on the B200 box")
    return solver


def _pin(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()


def test_device_closed_loop_tick_graph_equals_separate_calls(solver_mod):
    """solve -> plant (wave wrench, body acceleration, row counter) as ONE graph launch per tick against the same loop issued
    call by call: bit-identical states, one graph built; from a saturated start so that all three QP paths run"""
    import torch
    dev = torch.device("cuda", 0)
    N, B, T = 20, 256, 14
    w = wl.tracking_batch(B, N, seed=5, reference="lemniscate", pos_spread=2.5)
    amp, tau0 = wl.wave_disturbance(B, seed=3)
    d = lambda a, dt=None: torch.from_numpy(np.ascontiguousarray(a if dt is None else a.astype(dt))).to(dev)   # noqa: E731
    res = []
    for graph in (1, 0):
        s = solver_mod.BatchSolver(B, N)
        s.set_trajectory(w["traj"]); s.set_iterate(w["X"], w["U"])
        x, lines, p = d(w["x0"]), d(w["lines"], np.int32), d(w["p"])
        wave = (d(amp), d(tau0))
        acc = torch.zeros((B, 6), dtype=torch.float64, device=dev)
        out = None
        sts = []
        for t in range(T):
            if graph:
                out = s.tick(x, p=p, lines=lines, body_acc=acc, out=out, plant_h=0.05, wave=wave)
            else:
                out = s.solve_windowed(x, lines, p, out=out)
                solver_mod.plant_step(x, out[0], p, 0.05, wave=wave, tick=t, body_acc=acc, lines=lines)
            sts.append(out[2].cpu().numpy().copy())
        torch.cuda.synchronize()
        it, _ = s.stats()
        res.append((x.cpu().numpy(), lines.cpu().numpy(), acc.cpu().numpy(), out[0].cpu().numpy(), np.array(sts), it, s.graphs_built()))
        s.close()
    for a, b in zip(res[0][:5], res[1][:5]):
        assert np.array_equal(a, b)
    assert (res[0][4] == 0).all()
    assert res[0][6] == 1 and res[1][6] == 0          # one graph for the whole loop; none without tick()


def test_dob_tick_equals_ekf_then_solve(solver_mod):
    """ekf = 1: EKF -> parameters -> solve in one graph == br2_batch_ekf_device + br2_batch_solve_windowed_device"""
    import torch
    dev = torch.device("cuda", 0)
    N, B, T = 20, 128, 6
    w = wl.tracking_batch(B, N, seed=6, reference="lemniscate", pos_spread=0.2, level=True)
    rng = np.random.default_rng(1)
    thr = rng.uniform(-5, 5, (T, B, 6)); acc = rng.uniform(-0.3, 0.3, (T, B, 6))
    meas = w["x0"][None] + rng.uniform(-0.02, 0.02, (T, B, 12))
    d = lambda a, dt=None: torch.from_numpy(np.ascontiguousarray(a if dt is None else a.astype(dt))).to(dev)   # noqa: E731
    outs = []
    for fused in (1, 0):
        s = solver_mod.BatchSolver(B, N)
        s.set_trajectory(w["traj"]); s.set_iterate(w["X"], w["U"])
        lines = d(w["lines"], np.int32)
        x, th_in, ac = d(meas[0]), d(thr[0]), d(acc[0])          # fixed device buffers, refilled every tick
        wf = torch.empty((B, 6), dtype=torch.float64, device=dev)
        out = None
        us = []
        for t in range(T):
            x.copy_(d(meas[t])); th_in.copy_(d(thr[t])); ac.copy_(d(acc[t]))
            if fused:
                out = s.tick(x, lines=lines, thrusts=th_in, body_acc=ac, ekf=1, compensate=True, out=out, wf_dist=wf)
            else:
                wf2, pp = s.ekf(th_in, x, ac, compensate=True, out=(wf, torch.empty((B, 16), dtype=torch.float64, device=dev)))
                out = s.solve_windowed(x, lines, pp, out=out)
            us.append(out[0].cpu().numpy().copy())
        torch.cuda.synchronize()
        ex, eP = s.ekf_state()
        outs.append((np.array(us), wf.cpu().numpy(), ex, eP, out[2].cpu().numpy(), s.graphs_built()))
        s.close()
    for a, b in zip(outs[0][:5], outs[1][:5]):
        assert np.array_equal(a, b)
    assert (outs[0][4] == 0).all() and outs[0][5] == 1


def test_host_tick_equals_host_solve(solver_mod):
    """pinned host buffers: copies + kernels of a tick as one graph == br2_batch_solve_host / _windowed_host (+ ekf_host);
    pageable buffers take the stream path and give the same numbers"""
    N, B = 20, 96
    w = wl.tracking_batch(B, N, seed=8, pos_spread=2.0)
    ref = solver_mod.BatchSolver(B, N); ref.set_trajectory(w["traj"]); ref.set_iterate(w["X"], w["U"])
    u_ref, th_ref, st_ref = ref.solve(w["x0"], w["yref"], w["p"])
    for pinned in (True, False):
        f = _pin if pinned else (lambda a: np.ascontiguousarray(a).copy())
        s = solver_mod.BatchSolver(B, N); s.set_trajectory(w["traj"]); s.set_iterate(w["X"], w["U"])
        x0, yref, p = f(w["x0"]), f(w["yref"]), f(w["p"])
        out = (f(np.zeros((B, 4))), f(np.zeros((B, 6))), f(np.zeros(B, dtype=np.int32)))
        u0, th, st = s.tick(x0, p=p, yref=yref, out=out)
        assert np.array_equal(u0, u_ref) and np.array_equal(th, th_ref) and np.array_equal(st, st_ref)
        assert s.graphs_built() == 0                       # first sighting of these buffers: stream path
        # the second consecutive tick with the same buffers instantiates the graph, the third replays it
        s.tick(x0, p=p, yref=yref, out=out)
        assert s.graphs_built() == (1 if pinned else 0)
        s.tick(x0, p=p, yref=yref, out=out)
        assert s.graphs_built() == (1 if pinned else 0)
        lines = f(w["lines"].astype(np.int32))
        s.set_iterate(w["X"], w["U"])
        u1, _, _ = s.tick(x0, p=p, lines=lines, out=out)
        assert np.array_equal(u1, u_ref)                   # window of the same rows == the explicit yref
        s.set_iterate(w["X"], w["U"])
        u2, _, _ = s.tick(x0, p=p, lines=lines, out=out)   # ... and through its graph
        assert np.array_equal(u2, u_ref)
        assert s.graphs_built() == (2 if pinned else 0)
        s.close()
    ref.close()


def test_tick_argument_errors(solver_mod):
    s = solver_mod.BatchSolver(4, 10)
    x0 = np.zeros((4, 12)); p = np.zeros((4, 16)); yref = np.zeros((4, 11, 16))
    with pytest.raises(solver_mod.SolverError):
        s.tick(x0, p=p)                                     # neither yref nor lines
    with pytest.raises(solver_mod.SolverError):
        s.tick(x0, p=p, lines=np.zeros(4, dtype=np.int32))  # no trajectory set
    with pytest.raises(solver_mod.SolverError):
        s.tick(x0, yref=yref)                               # no p and no filter
    with pytest.raises(solver_mod.SolverError):
        s.tick(x0, yref=yref, ekf=1)                        # filter without its inputs
    with pytest.raises(solver_mod.SolverError):
        s.tick(x0, p=p, yref=yref, plant_h=0.05)            # plant is a device-path option
    s.close()


def test_pipelined_explicit_yref_upload_equals_plain_host_solve(solver_mod):
    """B >= 1024 with explicit reference windows: the upload is cut into four instance ranges, range c linearised and solved
    while the later ranges are still on the wire; instances are independent, so the result is bit-identical to the one-piece
    host solve -- over several closed-loop ticks from a saturated start (hints, visiting order and fallback list cross the ranges)"""
    N, B, T = 10, 1024 + 7, 5
    w = wl.tracking_batch(B, N, seed=9, pos_spread=2.5)
    a = solver_mod.BatchSolver(B, N); a.set_iterate(w["X"], w["U"])
    b = solver_mod.BatchSolver(B, N); b.set_iterate(w["X"], w["U"])
    x0, lines = w["x0"].copy(), w["lines"].copy()
    pin = lambda v: _pin(v)   # noqa: E731
    hx, hp = pin(x0), pin(w["p"])
    hy = pin(np.zeros((B, N + 1, 16)))
    out = (pin(np.zeros((B, 4))), pin(np.zeros((B, 6))), pin(np.zeros(B, dtype=np.int32)))
    for t in range(T):
        yref = traj.window_batch(w["traj"], lines, N)
        ua, tha, sta = a.solve(x0, yref, w["p"])                      # br2_batch_solve_host: one piece
        hx[:] = x0; hy[:] = yref
        ub, thb, stb = b.tick(hx, p=hp, yref=hy, out=out)             # br2_batch_tick_host: pipelined ranges (a graph from tick 1 on)
        assert np.array_equal(ua, ub) and np.array_equal(tha, thb) and np.array_equal(sta, stb), t
        assert (sta == 0).all()
        x0 = wl.plant_step(x0, ua, w["p"], 0.05)
        lines = lines + 1
    Xa, Ua = a.get_iterate(); Xb, Ub = b.get_iterate()
    assert np.array_equal(Xa, Xb) and np.array_equal(Ua, Ub)
    assert b.graphs_built() == 1
    a.close(); b.close()


def test_host_graph_replayed_on_new_input_buffers(solver_mod):
    """a caller that hands in a fresh pinned input buffer every tick (messages arriving) still replays ONE graph: its upload nodes
    are re-pointed; results equal the stream path tick by tick"""
    N, B, T = 10, 1024, 6
    w = wl.tracking_batch(B, N, seed=10, pos_spread=1.0)
    a = solver_mod.BatchSolver(B, N); a.set_iterate(w["X"], w["U"]); a.set_option("tick_graph", 0)
    b = solver_mod.BatchSolver(B, N); b.set_iterate(w["X"], w["U"])
    hp = _pin(w["p"])
    outa = (_pin(np.zeros((B, 4))), _pin(np.zeros((B, 6))), _pin(np.zeros(B, dtype=np.int32)))
    outb = (_pin(np.zeros((B, 4))), _pin(np.zeros((B, 6))), _pin(np.zeros(B, dtype=np.int32)))
    x0, lines = w["x0"].copy(), w["lines"].copy()
    for t in range(T):
        hx, hy = _pin(x0), _pin(traj.window_batch(w["traj"], lines, N))       # new buffers every tick
        ua, _, sta = a.tick(hx, p=hp, yref=hy, out=outa)
        ub, _, stb = b.tick(hx, p=hp, yref=hy, out=outb)
        assert np.array_equal(ua, ub) and np.array_equal(sta, stb) and (sta == 0).all(), t
        x0 = wl.plant_step(x0, ua.copy(), w["p"], 0.05)
        lines = lines + 1
    assert a.graphs_built() == 0 and b.graphs_built() == 1 and b.graph_updates() == T - 2
    a.close(); b.close()


def test_reference_announced_one_tick_ahead_equals_plain_host_ticks(solver_mod):
    """br2_batch_set_next_yref_host: the next tick's explicit reference window, registered before the call, is uploaded beside the running
    tick's kernels into the buffer that tick does not read; the tick that names it uploads only x0 / p.  Same results as plain host
    ticks, bit for bit, over a closed loop from a saturated start -- also when an announced window ends up NOT being the one passed
    (ordinary path) and when nothing is announced in between"""
    N, B, T = 10, 1024 + 5, 8
    w = wl.tracking_batch(B, N, seed=12, pos_spread=1.0)
    a = solver_mod.BatchSolver(B, N); a.set_iterate(w["X"], w["U"])
    b = solver_mod.BatchSolver(B, N); b.set_iterate(w["X"], w["U"])
    hp = _pin(w["p"])
    outa = (_pin(np.zeros((B, 4))), _pin(np.zeros((B, 6))), _pin(np.zeros(B, dtype=np.int32)))
    outb = (_pin(np.zeros((B, 4))), _pin(np.zeros((B, 6))), _pin(np.zeros(B, dtype=np.int32)))
    lines = w["lines"].copy()
    def wc(a):                                          # write-combined pinned memory from the library's own allocator
        b = solver_mod.pinned_empty(a.shape, np.float64, write_combined=True)
        b[...] = a
        return b
    refs = [wc(traj.window_batch(w["traj"], lines + t, N)) for t in range(T + 1)]       # the reference is a trajectory: known ahead
    decoy = _pin(np.zeros((B, N + 1, 16)))
    x0 = w["x0"].copy()
    for t in range(T):
        hx = _pin(x0)
        ua, tha, sta = a.tick(hx, p=hp, yref=refs[t], out=outa)
        if t == 4:
            b.set_next_yref(decoy)                     # announced, but the next tick passes another window: ordinary upload
        elif t != 5:
            b.set_next_yref(refs[t + 1])               # (t == 5: nothing announced; tick 6 takes the ordinary path)
        ub, thb, stb = b.tick(hx, p=hp, yref=refs[t], out=outb)
        assert np.array_equal(ua, ub) and np.array_equal(tha, thb) and np.array_equal(sta, stb), t
        assert (sta == 0).all()
        x0 = wl.plant_step(x0, ua.copy(), w["p"], 0.05)
    Xa, Ua = a.get_iterate(); Xb, Ub = b.get_iterate()
    assert np.array_equal(Xa, Xb) and np.array_equal(Ua, Ub)
    with pytest.raises(solver_mod.SolverError):
        b.set_next_yref(np.zeros((B, N + 1, 16)))      # pageable memory is refused
    a.close(); b.close()


def test_parameters_persist_between_host_ticks(solver_mod):
    """p == NULL on the host path: the parameters of the last call that supplied them stay in the solver (bluerov2_acados_update_params
    semantics); ticks without p equal ticks that pass the same p every time; a solver that never got parameters refuses"""
    N, B, T = 10, 300, 5
    w = wl.tracking_batch(B, N, seed=14, pos_spread=0.5)
    a = solver_mod.BatchSolver(B, N); a.set_iterate(w["X"], w["U"]); a.set_trajectory(w["traj"])
    b = solver_mod.BatchSolver(B, N); b.set_iterate(w["X"], w["U"]); b.set_trajectory(w["traj"])
    p2 = w["p"].copy(); p2[:, 0] += 0.3                       # a second parameter set (disturbance estimate changed)
    hp, hp2 = _pin(w["p"]), _pin(p2)
    outa = (_pin(np.zeros((B, 4))), _pin(np.zeros((B, 6))), _pin(np.zeros(B, dtype=np.int32)))
    outb = (_pin(np.zeros((B, 4))), _pin(np.zeros((B, 6))), _pin(np.zeros(B, dtype=np.int32)))
    x0, lines = w["x0"].copy(), w["lines"].astype(np.int32)
    with pytest.raises(solver_mod.SolverError):
        b.tick(_pin(x0), lines=_pin(lines), out=outb)         # nothing resident yet
    for t in range(T):
        hx, hl = _pin(x0), _pin(lines + t)
        pa = hp if t < 3 else hp2
        ua, _, sta = a.tick(hx, p=pa, lines=hl, out=outa)
        ub, _, stb = b.tick(hx, p=(hp if t == 0 else hp2 if t == 3 else None), lines=hl, out=outb)
        assert np.array_equal(ua, ub) and np.array_equal(sta, stb) and (sta == 0).all(), t
        x0 = wl.plant_step(x0, ua.copy(), pa, 0.05)
    a.close(); b.close()
