"""Host-side mirror of the batched C-ABI (include/bluerov2_b200.h), via ctypes.

``BatchSolver`` is to B instances what the capsule of the acados-generated solver is to one
(bluerov2_dobmpc/scripts/c_generated_code/acados_solver_bluerov2.h:79-167): create once, then per control tick
set (x0, yref, p) and solve, exactly the call sequence of ``BLUEROV2_DOB::solve`` (bluerov2_dob.cpp:307-395).
The arithmetic is entirely in the CUDA library; this file only marshals pointers.  There is NO CPU path: the
constructor raises if the library is missing or no CUDA device is visible.

Inputs may be

* numpy arrays (host): ``solve`` goes through ``br2_batch_solve_host`` (H2D, kernels, D2H, synchronise);
* torch CUDA tensors (float64, contiguous, on the solver's device): ``solve`` only enqueues on torch's current
  stream through ``br2_batch_solve_device`` and returns torch tensors.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build as _build

NX, NU, NP, NY, NTHRUST, NEKF = 12, 4, 16, 16, 6, 18

_D = C.POINTER(C.c_double)
_I = C.POINTER(C.c_int)
_lib = None


class SolverError(RuntimeError):
    pass


def load_library(path: str | None = None):
    """dlopen the product library (built in-tree by bluerov2_b200.build).  Raises if it does not exist."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    path = path or _build.LIB
    if not os.path.exists(path):
        raise SolverError(f"{path} not built: run `python -m bluerov2_b200.build` (needs nvcc); there is no CPU fallback")
    L = C.CDLL(path, mode=C.RTLD_GLOBAL)
    V = C.c_void_p
    sig = {
        "br2_last_error": (C.c_char_p, []),
        "br2_version": (C.c_char_p, []),
        "br2_device_count": (C.c_int, []),
        "br2_batch_create": (C.c_int, [C.POINTER(V), C.c_int, C.c_int, V, C.c_int]),
        "br2_batch_free": (C.c_int, [V]),
        "br2_batch_size": (C.c_int, [V]),
        "br2_batch_horizon": (C.c_int, [V]),
        "br2_batch_set_weights": (C.c_int, [V, V, V]),
        "br2_batch_set_bounds": (C.c_int, [V, V, V]),
        "br2_batch_set_time_steps": (C.c_int, [V, V]),
        "br2_batch_set_option_int": (C.c_int, [V, C.c_char_p, C.c_int]),
        "br2_batch_set_option_double": (C.c_int, [V, C.c_char_p, C.c_double]),
        "br2_batch_reset": (C.c_int, [V, C.c_int]),
        "br2_batch_set_iterate_host": (C.c_int, [V, V, V]),
        "br2_batch_get_iterate_host": (C.c_int, [V, V, V]),
        "br2_batch_iterate_device": (C.c_int, [V, C.POINTER(V), C.POINTER(V)]),
        "br2_batch_solve_device": (C.c_int, [V, V, V, V, C.c_int, V, V, V, V]),
        "br2_batch_solve_host": (C.c_int, [V, V, V, V, C.c_int, V, V, V]),
        "br2_batch_set_trajectory": (C.c_int, [V, V, C.c_int]),
        "br2_batch_solve_windowed_device": (C.c_int, [V, V, V, V, C.c_int, V, V, V, V]),
        "br2_batch_solve_windowed_host": (C.c_int, [V, V, V, V, C.c_int, V, V, V]),
        "br2_batch_get_stats_host": (C.c_int, [V, V, V]),
        "br2_batch_get_linearization_host": (C.c_int, [V, V, V]),
        "br2_batch_last_solve_time": (C.c_double, [V]),
        "br2_batch_last_kernel_times": (C.c_int, [V, _D, _D]),
        "br2_batch_ipm_iterations_total": (C.c_longlong, [V, C.c_int]),
        "br2_batch_set_next_yref_host": (C.c_int, [V, V]),
        "br2_host_alloc": (C.c_int, [V, C.c_size_t, C.c_int]),
        "br2_host_free": (C.c_int, [V]),
        "br2_batch_phase_cycles": (C.c_int, [V, V, C.c_int]),
        "br2_batch_ekf_phase_cycles": (C.c_int, [V, V, C.c_int]),
        "br2_batch_nonzero_status_total": (C.c_longlong, [V, C.c_int]),
        "br2_batch_tick_device": (C.c_int, [V, V, V]),
        "br2_batch_tick_host": (C.c_int, [V, V]),
        "br2_batch_graphs_built": (C.c_int, [V]),
        "br2_batch_graph_updates": (C.c_int, [V]),
        "br2_batch_shard_init": (C.c_int, [V, C.c_int, C.c_int]),
        "br2_batch_shard_handle": (C.c_int, [V, V]),
        "br2_batch_shard_connect": (C.c_int, [V, C.c_int, V]),
        "br2_batch_shard_wait": (C.c_int, [V, V]),
        "br2_batch_shard_gathered": (C.c_int, [V, C.c_int, C.POINTER(V)]),
        "br2_batch_tick_count": (C.c_int, [V]),
        "br2_batch_set_tick_index": (C.c_int, [V, C.c_int]),
        "br2_plant_step_device": (C.c_int, [C.c_int, V, V, V, V, V, V, C.c_int, C.c_double, V, V, V]),
        "br2_batch_ekf_reset": (C.c_int, [V]),
        "br2_batch_ekf_device": (C.c_int, [V, V, V, V, V, V, C.c_int, V]),
        "br2_batch_ekf_host": (C.c_int, [V, V, V, V, V, V, C.c_int]),
        "br2_batch_ekf_get_state_host": (C.c_int, [V, V, V]),
        "br2_batch_ekf_set_state_host": (C.c_int, [V, V, V]),
        "br2_batch_yaw_reset": (C.c_int, [V]),
        "br2_batch_yaw_unwrap_device": (C.c_int, [V, V, V]),
        "br2_batch_yaw_unwrap_host": (C.c_int, [V, V]),
        "br2_batch_yaw_get_state_host": (C.c_int, [V, V]),
        "br2_batch_yaw_set_state_host": (C.c_int, [V, V]),
        "br2_batch_rls_reset": (C.c_int, [V]),
        "br2_batch_rls_device": (C.c_int, [V, V, V, V, C.c_int, V]),
        "br2_batch_rls_host": (C.c_int, [V, V, V, V, C.c_int]),
        "br2_batch_rls_get_state_host": (C.c_int, [V, V]),
        "br2_batch_rls_set_state_host": (C.c_int, [V, V]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def _is_torch(a) -> bool:
    return type(a).__module__.startswith("torch")


def _np(a, shape, dtype=np.float64):
    a = np.ascontiguousarray(a, dtype=dtype)
    if a.shape != tuple(shape):
        raise ValueError(f"expected shape {tuple(shape)}, got {a.shape}")
    return a


def _ptr(a):
    if a is None:
        return None
    if _is_torch(a):
        return C.c_void_p(a.data_ptr())
    return C.c_void_p(a.ctypes.data)


class _TickIO(C.Structure):
    """struct br2_tick_io (include/bluerov2_b200.h)"""
    _fields_ = [("x0", C.c_void_p), ("yref", C.c_void_p), ("p", C.c_void_p), ("thrusts", C.c_void_p), ("lines", C.c_void_p),
                ("body_acc", C.c_void_p), ("u0", C.c_void_p), ("thrust", C.c_void_p), ("wf_dist", C.c_void_p), ("status", C.c_void_p),
                ("wave_amp", C.c_void_p), ("wave_tau0", C.c_void_p), ("plant_h", C.c_double),
                ("p_per_stage", C.c_int), ("ekf", C.c_int), ("compensate", C.c_int)]


class _HostArgs:
    """Host-array marshalling for the hot calls.  Callers pass the same numpy buffers tick after tick; checking layout and
    building a ctypes pointer costs ~2.5 us per array per call, 15 us per solve -- 4 % of a 0.4 ms tick.  An array object
    that already has the right dtype / shape / contiguity is validated once and remembered (by identity; the entry keeps
    the array alive, so its id and data pointer cannot be recycled)."""

    def __init__(self, cap: int = 64):
        self._d, self._cap = {}, cap

    def get(self, a, shape, dtype=np.float64, out: bool = False):
        """-> (array kept alive by the caller for the duration of the C call, c_void_p)"""
        e = self._d.get(id(a))
        if e is not None and e[0] is a and e[2] == shape and e[3] is dtype:
            return a, e[1]
        ok = (isinstance(a, np.ndarray) and a.dtype == dtype and a.flags.c_contiguous and a.shape == tuple(shape)
              and (a.flags.writeable or not out))
        if not ok:
            if out:
                raise ValueError(f"output buffer must be a writable C-contiguous {np.dtype(dtype).name} array of shape {tuple(shape)}")
            b = _np(a, shape, dtype)              # converted copy: valid for this call only
            return b, C.c_void_p(b.ctypes.data)
        if len(self._d) >= self._cap:
            self._d.clear()
        ptr = C.c_void_p(a.ctypes.data)
        self._d[id(a)] = (a, ptr, shape, dtype)
        return a, ptr


class BatchSolver:
    """B independent BlueROV2 OCP instances advanced by one SQP-RTI step per ``solve`` call."""

    def __init__(self, batch: int, N: int = 40, time_steps=None, device: int = 0, lib_path: str | None = None):
        self._L = load_library(lib_path)
        self._h = C.c_void_p()
        self._hostargs = _HostArgs()
        self._tick_cache = {}
        self._arr_cache = {}
        self._tick_tpl = {}
        ts = None if time_steps is None else _np(time_steps, (N,))
        self._check(self._L.br2_batch_create(C.byref(self._h), int(batch), int(N), _ptr(ts), int(device)))
        self.B, self.N, self.device = int(batch), int(N), int(device)

    # -- plumbing --------------------------------------------------------------------------------------
    def _check(self, rc: int):
        if rc != 0:
            raise SolverError(f"bluerov2_b200 error {rc}: {self._L.br2_last_error().decode()}")

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._L.br2_batch_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _dev_check(self, t, shape, dtype=None):
        import torch
        dtype = dtype or torch.float64
        if not t.is_cuda or t.device.index != self.device:
            raise ValueError(f"tensor must live on cuda:{self.device}")
        if t.dtype != dtype or not t.is_contiguous() or tuple(t.shape) != tuple(shape):
            raise ValueError(f"expected contiguous {dtype} tensor of shape {tuple(shape)}, got {t.dtype} {tuple(t.shape)}")
        return t

    # -- configuration (== the options baked by generate_c_code.py / settable through ocp_nlp_*_set) ----
    def set_weights(self, W=None, We=None):
        W = None if W is None else _np(W, (16,))
        We = None if We is None else _np(We, (12,))
        self._check(self._L.br2_batch_set_weights(self._h, _ptr(W), _ptr(We)))

    def set_bounds(self, lbu, ubu):
        lbu, ubu = _np(lbu, (4,)), _np(ubu, (4,))
        self._check(self._L.br2_batch_set_bounds(self._h, _ptr(lbu), _ptr(ubu)))

    def set_time_steps(self, ts):
        ts = _np(ts, (self.N,))
        self._check(self._L.br2_batch_set_time_steps(self._h, _ptr(ts)))

    def set_option(self, name: str, value):
        if isinstance(value, (int, np.integer)) and not isinstance(value, bool):
            self._check(self._L.br2_batch_set_option_int(self._h, name.encode(), int(value)))
        else:
            self._check(self._L.br2_batch_set_option_double(self._h, name.encode(), float(value)))

    # -- iterate (nlp_out) -----------------------------------------------------------------------------
    def reset(self, mode: int = 0):
        self._check(self._L.br2_batch_reset(self._h, int(mode)))

    def set_iterate(self, X=None, U=None):
        X = None if X is None else _np(X, (self.B, self.N + 1, NX))
        U = None if U is None else _np(U, (self.B, self.N, NU))
        self._check(self._L.br2_batch_set_iterate_host(self._h, _ptr(X), _ptr(U)))

    def get_iterate(self):
        X = np.empty((self.B, self.N + 1, NX))
        U = np.empty((self.B, self.N, NU))
        self._check(self._L.br2_batch_get_iterate_host(self._h, _ptr(X), _ptr(U)))
        return X, U

    # -- the RTI step ----------------------------------------------------------------------------------
    def solve(self, x0, yref, p, out=None):
        """One SQP-RTI step for all instances.  Returns (u0[B,4], thrust[B,6], status[B]).

        ``p`` is [B,16] (one parameter vector per instance, what the nodes do: bluerov2_dob.cpp:324-355) or
        [B,N+1,16] (per stage, acados_solver_bluerov2.c:835-883).  ``out`` = (u0, thrust, status) buffers to reuse.
        """
        per_stage = int(len(p.shape) == 3)
        pshape = (self.B, self.N + 1, NP) if per_stage else (self.B, NP)
        if _is_torch(x0):
            import torch
            self._dev_check(x0, (self.B, NX)); self._dev_check(yref, (self.B, self.N + 1, NY)); self._dev_check(p, pshape)
            if out is None:
                dev = x0.device
                out = (torch.empty((self.B, NU), dtype=torch.float64, device=dev),
                       torch.empty((self.B, NTHRUST), dtype=torch.float64, device=dev),
                       torch.empty((self.B,), dtype=torch.int32, device=dev))
            u0, th, st = out
            self._dev_check(u0, (self.B, NU)); self._dev_check(th, (self.B, NTHRUST)); self._dev_check(st, (self.B,), torch.int32)
            stream = C.c_void_p(torch.cuda.current_stream(x0.device).cuda_stream)
            self._check(self._L.br2_batch_solve_device(self._h, _ptr(x0), _ptr(yref), _ptr(p), per_stage,
                                                       _ptr(u0), _ptr(th), _ptr(st), stream))
            return u0, th, st
        H = self._hostargs.get
        (k0, a0), (k1, a1), (k2, a2) = H(x0, (self.B, NX)), H(yref, (self.B, self.N + 1, NY)), H(p, pshape)
        if out is None:
            out = (np.empty((self.B, NU)), np.empty((self.B, NTHRUST)), np.empty((self.B,), dtype=np.int32))
        u0, th, st = out
        (_, o0), (_, o1), (_, o2) = (H(u0, (self.B, NU), out=True), H(th, (self.B, NTHRUST), out=True),
                                     H(st, (self.B,), np.int32, out=True))
        self._check(self._L.br2_batch_solve_host(self._h, a0, a1, a2, per_stage, o0, o1, o2))
        return u0, th, st

    def set_trajectory(self, traj):
        """Upload the reference trajectory (rows x 16) once; ``solve_windowed`` then takes one row index per instance."""
        traj = np.ascontiguousarray(traj, dtype=np.float64)
        if traj.ndim != 2 or traj.shape[1] != NY:
            raise ValueError(f"trajectory must be [rows, {NY}], got {traj.shape}")
        self._check(self._L.br2_batch_set_trajectory(self._h, _ptr(traj), int(traj.shape[0])))
        self.traj_rows = int(traj.shape[0])

    def solve_windowed(self, x0, lines, p, out=None):
        """Like ``solve`` but the horizon reference is windowed on the device from the uploaded trajectory:
        ``lines[b]`` is instance b's first row (``line_number`` in BLUEROV2_DOB::solve, bluerov2_dob.cpp:367)."""
        per_stage = int(len(p.shape) == 3)
        pshape = (self.B, self.N + 1, NP) if per_stage else (self.B, NP)
        if _is_torch(x0):
            import torch
            self._dev_check(x0, (self.B, NX)); self._dev_check(lines, (self.B,), torch.int32); self._dev_check(p, pshape)
            if out is None:
                dev = x0.device
                out = (torch.empty((self.B, NU), dtype=torch.float64, device=dev),
                       torch.empty((self.B, NTHRUST), dtype=torch.float64, device=dev),
                       torch.empty((self.B,), dtype=torch.int32, device=dev))
            u0, th, st = out
            self._dev_check(u0, (self.B, NU)); self._dev_check(th, (self.B, NTHRUST)); self._dev_check(st, (self.B,), torch.int32)
            stream = C.c_void_p(torch.cuda.current_stream(x0.device).cuda_stream)
            self._check(self._L.br2_batch_solve_windowed_device(self._h, _ptr(x0), _ptr(lines), _ptr(p), per_stage,
                                                                _ptr(u0), _ptr(th), _ptr(st), stream))
            return u0, th, st
        H = self._hostargs.get
        (k0, a0), (k1, a1), (k2, a2) = H(x0, (self.B, NX)), H(lines, (self.B,), np.int32), H(p, pshape)
        if out is None:
            out = (np.empty((self.B, NU)), np.empty((self.B, NTHRUST)), np.empty((self.B,), dtype=np.int32))
        u0, th, st = out
        (_, o0), (_, o1), (_, o2) = (H(u0, (self.B, NU), out=True), H(th, (self.B, NTHRUST), out=True),
                                     H(st, (self.B,), np.int32, out=True))
        self._check(self._L.br2_batch_solve_windowed_host(self._h, a0, a1, a2, per_stage, o0, o1, o2))
        return u0, th, st

    def tick(self, x0, p=None, yref=None, lines=None, thrusts=None, body_acc=None, ekf: int = 0, compensate: bool = True,
             out=None, wf_dist=None, plant_h: float = 0.0, wave=None):
        """One control tick as ONE call (br2_batch_tick_device / _host): [EKF (-> RLS) -> p ->] RTI step [-> plant step on x0 in
        place].  All arrays torch CUDA tensors (enqueue on torch's current stream) or all numpy arrays (synchronous).  A caller
        that passes the same buffers every tick replays one cached CUDA graph.  Returns (u0, thrust, status)."""
        if out is not None:
            # hot path: a call seen before (same objects; the cache entry keeps them alive, so an id cannot have been recycled) is one
            # dictionary lookup and one foreign call -- the wrapper is on the critical path of a synchronous host tick
            w0, w1 = wave if wave is not None else (None, None)
            ent = self._tick_cache.get((id(x0), id(p), id(yref), id(lines), id(thrusts), id(body_acc), id(out[0]), id(out[1]), id(out[2]),
                                        id(wf_dist), id(w0), id(w1), ekf, compensate, plant_h))
            if ent is not None:
                if ent[3]:
                    import torch
                    rc = self._L.br2_batch_tick_device(self._h, ent[2], C.c_void_p(torch.cuda.current_stream(x0.device).cuda_stream))
                else:
                    rc = self._L.br2_batch_tick_host(self._h, ent[2])
                if rc:
                    self._check(rc)
                return out
            # a new buffer in a call shape seen before (a caller that hands in a fresh message buffer every tick): only x0 and the
            # reference are new -- validate those two, copy the validated struct of the sibling call and re-point it
            tpl = self._tick_tpl.get((id(p), id(thrusts), id(body_acc), id(out[0]), id(out[1]), id(out[2]), id(wf_dist), id(w0), id(w1),
                                      ekf, compensate, plant_h, yref is None, lines is None))
            if tpl is not None and not tpl[3]:
                B, N = self.B, self.N
                ref, rshape, rtype = (lines, (B,), np.int32) if yref is None else (yref, (B, N + 1, NY), np.float64)
                if (type(x0) is np.ndarray and x0.dtype == np.float64 and x0.flags.c_contiguous and x0.shape == (B, NX)
                        and type(ref) is np.ndarray and ref.dtype == rtype and ref.flags.c_contiguous and ref.shape == rshape):
                    try:
                        px, pr = C.addressof(C.c_char.from_buffer(x0)), C.addressof(C.c_char.from_buffer(ref))
                    except (TypeError, ValueError):
                        px, pr = x0.ctypes.data, ref.ctypes.data
                    io = _TickIO.from_buffer_copy(tpl[0])
                    io.x0 = px
                    if yref is None:
                        io.lines = pr
                    else:
                        io.yref = pr
                    if len(self._tick_cache) >= 1024:
                        for k in list(self._tick_cache)[:512]:
                            del self._tick_cache[k]
                    ref_io = C.byref(io)
                    self._tick_cache[(id(x0), id(p), id(yref), id(lines), id(thrusts), id(body_acc), id(out[0]), id(out[1]), id(out[2]),
                                      id(wf_dist), id(w0), id(w1), ekf, compensate, plant_h)] = (io, (x0, ref), ref_io, False)
                    rc = self._L.br2_batch_tick_host(self._h, ref_io)
                    if rc:
                        self._check(rc)
                    return out
        dev = _is_torch(x0)
        if out is None:
            if dev:
                import torch
                out = (torch.empty((self.B, NU), dtype=torch.float64, device=x0.device),
                       torch.empty((self.B, NTHRUST), dtype=torch.float64, device=x0.device),
                       torch.empty((self.B,), dtype=torch.int32, device=x0.device))
            else:
                out = (np.empty((self.B, NU)), np.empty((self.B, NTHRUST)), np.empty((self.B,), dtype=np.int32))
        args = (x0, p, yref, lines, thrusts, body_acc, out[0], out[1], out[2], wf_dist) + (tuple(wave) if wave is not None else (None, None))
        key = tuple(id(a) for a in args) + (ekf, compensate, plant_h)
        per_stage = int(p is not None and len(p.shape) == 3)
        shapes = ((self.B, NX), (self.B, self.N + 1, NP) if per_stage else (self.B, NP), (self.B, self.N + 1, NY), (self.B,),
                  (self.B, 6), (self.B, 6), (self.B, NU), (self.B, NTHRUST), (self.B,), (self.B, 6), (self.B, 4), (self.B,))
        ints = (3, 8)       # lines, status
        ptrs = []
        for i, (a, shp) in enumerate(zip(args, shapes)):
            if a is None:
                ptrs.append(None)
                continue
            if dev:
                import torch
                self._dev_check(a, shp, torch.int32 if i in ints else torch.float64)
                ptrs.append(a.data_ptr())
            else:
                # (validated once per array object and role; the address without numpy's ctypes helper object: this path runs on
                # every tick of a caller that hands in a fresh message buffer each time)
                hit = self._arr_cache.get((id(a), i))
                if hit is not None:
                    ptrs.append(hit[1])
                    continue
                want = np.int32 if i in ints else np.float64
                if not (type(a) is np.ndarray and a.dtype == want and a.flags.c_contiguous and a.shape == tuple(shp)):
                    raise ValueError(f"tick(): argument {i} must be a contiguous {want.__name__} array of shape {tuple(shp)}")
                try:
                    ptr = C.addressof(C.c_char.from_buffer(a))
                except (TypeError, ValueError):          # read-only or empty buffer
                    ptr = a.ctypes.data
                if len(self._arr_cache) >= 4096:
                    self._arr_cache.clear()
                self._arr_cache[(id(a), i)] = (a, ptr)
                ptrs.append(ptr)
        io = _TickIO(ptrs[0], ptrs[2], ptrs[1], ptrs[4], ptrs[3], ptrs[5], ptrs[6], ptrs[7], ptrs[9], ptrs[8], ptrs[10], ptrs[11],
                     float(plant_h), per_stage, int(ekf), int(bool(compensate)))
        if len(self._tick_cache) >= 1024:          # (a caller cycling through many buffers: forget the older half)
            for k in list(self._tick_cache)[:512]:
                del self._tick_cache[k]
        ent = self._tick_cache[key] = (io, args, C.byref(io), dev)
        w0, w1 = wave if wave is not None else (None, None)
        if len(self._tick_tpl) >= 64:
            self._tick_tpl.clear()
        self._tick_tpl[(id(p), id(thrusts), id(body_acc), id(out[0]), id(out[1]), id(out[2]), id(wf_dist), id(w0), id(w1),
                        ekf, compensate, plant_h, yref is None, lines is None)] = (io, args, None, dev)
        if dev:
            import torch
            stream = C.c_void_p(torch.cuda.current_stream(x0.device).cuda_stream)
            self._check(self._L.br2_batch_tick_device(self._h, ent[2], stream))
        else:
            self._check(self._L.br2_batch_tick_host(self._h, ent[2]))
        return out

    def set_next_yref(self, yref_next):
        """host path, explicit reference: register the NEXT tick's (N+1) x 16 window (a pinned numpy array that stays untouched until that
        tick has returned) before calling tick(); it is uploaded while the current tick computes (br2_batch_set_next_yref_host)"""
        if yref_next is None:
            self._check(self._L.br2_batch_set_next_yref_host(self._h, None))
            return
        if not (isinstance(yref_next, np.ndarray) and yref_next.dtype == np.float64 and yref_next.flags.c_contiguous
                and yref_next.shape == (self.B, self.N + 1, NY)):
            raise ValueError(f"set_next_yref(): contiguous float64 array of shape {(self.B, self.N + 1, NY)} expected")
        self._next_yref_keepalive = yref_next
        self._check(self._L.br2_batch_set_next_yref_host(self._h, C.c_void_p(yref_next.ctypes.data)))

    def graphs_built(self) -> int:
        return int(self._L.br2_batch_graphs_built(self._h))

    # -- sharding: peer-to-peer exchange of the thrust vectors, see bluerov2_b200.sharding.PeerThrustExchange --------
    def shard_init(self, rank: int, world: int):
        self._check(self._L.br2_batch_shard_init(self._h, int(rank), int(world)))

    def shard_handle(self) -> bytes:
        buf = C.create_string_buffer(64)
        self._check(self._L.br2_batch_shard_handle(self._h, buf))
        return buf.raw

    def shard_connect(self, peer: int, handle: bytes):
        if len(handle) != 64:
            raise ValueError("a CUDA IPC handle has 64 bytes")
        self._check(self._L.br2_batch_shard_connect(self._h, int(peer), C.create_string_buffer(handle, 64)))

    def shard_wait(self, stream=None):
        """enqueue a wait until every rank has published the ticks enqueued so far"""
        if stream is None:
            import torch
            stream = torch.cuda.current_stream(self.device).cuda_stream
        self._check(self._L.br2_batch_shard_wait(self._h, C.c_void_p(stream)))

    def shard_gathered_ptr(self, parity: int) -> int:
        p = C.c_void_p()
        self._check(self._L.br2_batch_shard_gathered(self._h, int(parity), C.byref(p)))
        return int(p.value)

    def tick_count(self) -> int:
        return int(self._L.br2_batch_tick_count(self._h))

    def graph_updates(self) -> int:
        return int(self._L.br2_batch_graph_updates(self._h))

    def set_tick_index(self, next_tick: int = 0):
        """index the next tick's plant step uses for the wave phase (tau = tau0 + 0.125 * index); counts up by itself"""
        self._check(self._L.br2_batch_set_tick_index(self._h, int(next_tick)))

    def stats(self):
        """(ipm_iterations[B], info[B,4] = (mu, stationarity residual, max |dynamics gap|, stationarity scale))."""
        it = np.empty((self.B,), dtype=np.int32)
        info = np.empty((self.B, 4))
        self._check(self._L.br2_batch_get_stats_host(self._h, _ptr(it), _ptr(info)))
        return it, info

    def linearization(self):
        """(A[B,N,12,12], B[B,N,12,4], b[B,N,12]) of the last solve."""
        AB = np.empty((self.B, self.N, 12, 16))
        b = np.empty((self.B, self.N, 12))
        self._check(self._L.br2_batch_get_linearization_host(self._h, _ptr(AB), _ptr(b)))
        return np.ascontiguousarray(AB[..., :12]), np.ascontiguousarray(AB[..., 12:]), b

    def last_solve_time(self) -> float:
        return float(self._L.br2_batch_last_solve_time(self._h))

    def last_kernel_times(self):
        """(linearisation seconds, Riccati-IPM seconds) of the last solve, CUDA events on the launching stream."""
        a, b = C.c_double(), C.c_double()
        self._check(self._L.br2_batch_last_kernel_times(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def ipm_iterations_total(self, reset: bool = False) -> int:
        return int(self._L.br2_batch_ipm_iterations_total(self._h, int(reset)))

    def nonzero_status_total(self, reset: bool = False) -> int:
        return int(self._L.br2_batch_nonzero_status_total(self._h, int(reset)))

    PHASES = ("factor_abs", "factor_as", "fwd_closed_loop", "primal_check", "costate_check", "ipm_start", "factor_ipm", "fwd_affine",
              "e1_centring", "bwd_corrector", "fwd_corrector", "e2_update", "epilogue", "lin_state_trajectory", "lin_jacobians", "lin_sensitivities")

    def phase_cycles(self, reset: bool = False) -> dict:
        """SM cycles per phase of the IPM kernel summed over warps (library built with -DBR2_PROFILE; zeros otherwise)"""
        out = np.zeros(16, dtype=np.uint64)
        self._check(self._L.br2_batch_phase_cycles(self._h, out.ctypes.data_as(C.c_void_p), int(reset)))
        return {n: int(out[i]) for i, n in enumerate(self.PHASES)}

    EKF_PHASES = ("load", "rk4_F", "P_pred", "h_H", "S", "inverse", "gain", "state_joseph", "store")

    def ekf_phase_cycles(self, reset: bool = False) -> dict:
        """SM cycles per phase of the EKF kernel summed over warps (library built with -DBR2_PROFILE; zeros otherwise)"""
        out = np.zeros(12, dtype=np.uint64)
        self._check(self._L.br2_batch_ekf_phase_cycles(self._h, out.ctypes.data_as(C.c_void_p), int(reset)))
        return {n: int(out[i]) for i, n in enumerate(self.EKF_PHASES)}

    # -- EKF (BLUEROV2_DOB::EKF) -----------------------------------------------------------------------
    def ekf_reset(self):
        self._check(self._L.br2_batch_ekf_reset(self._h))

    def ekf(self, thrusts, meas, body_acc, compensate: bool = True, out=None):
        """One EKF step for all instances.  Returns (wf_dist[B,6], p[B,16])."""
        if _is_torch(thrusts):
            import torch
            self._dev_check(thrusts, (self.B, 6)); self._dev_check(meas, (self.B, 12)); self._dev_check(body_acc, (self.B, 6))
            if out is None:
                out = (torch.empty((self.B, 6), dtype=torch.float64, device=thrusts.device),
                       torch.empty((self.B, NP), dtype=torch.float64, device=thrusts.device))
            wf, pp = out
            stream = C.c_void_p(torch.cuda.current_stream(thrusts.device).cuda_stream)
            self._check(self._L.br2_batch_ekf_device(self._h, _ptr(thrusts), _ptr(meas), _ptr(body_acc), _ptr(wf), _ptr(pp),
                                                     int(bool(compensate)), stream))
            return wf, pp
        H = self._hostargs.get
        (k0, a0), (k1, a1), (k2, a2) = H(thrusts, (self.B, 6)), H(meas, (self.B, 12)), H(body_acc, (self.B, 6))
        if out is None:
            out = (np.empty((self.B, 6)), np.empty((self.B, NP)))
        wf, pp = out
        (_, o0), (_, o1) = H(wf, (self.B, 6), out=True), H(pp, (self.B, NP), out=True)
        self._check(self._L.br2_batch_ekf_host(self._h, a0, a1, a2, o0, o1, int(bool(compensate))))
        return wf, pp

    def ekf_state(self):
        x = np.empty((self.B, NEKF)); P = np.empty((self.B, NEKF, NEKF))
        self._check(self._L.br2_batch_ekf_get_state_host(self._h, _ptr(x), _ptr(P)))
        return x, P

    def set_ekf_state(self, x=None, P=None):
        x = None if x is None else _np(x, (self.B, NEKF))
        P = None if P is None else _np(P, (self.B, NEKF, NEKF))
        self._check(self._L.br2_batch_ekf_set_state_host(self._h, _ptr(x), _ptr(P)))

    # -- continuous yaw (top of BLUEROV2_DOB::solve, bluerov2_dob.cpp:272-304) ---------------------------------
    def yaw_reset(self):
        self._check(self._L.br2_batch_yaw_reset(self._h))

    def unwrap_yaw(self, x0):
        """In place on ``x0[:, 5]``: measured yaw in (-pi, pi] -> the node's continuous ``yaw_sum`` (float accumulators, as in
        the reference).  numpy [B,12] (round trip through the device) or a CUDA tensor (enqueued on torch's current stream)."""
        if _is_torch(x0):
            import torch
            self._dev_check(x0, (self.B, NX))
            stream = C.c_void_p(torch.cuda.current_stream(x0.device).cuda_stream)
            self._check(self._L.br2_batch_yaw_unwrap_device(self._h, _ptr(x0), stream))
            return x0
        _, a0 = self._hostargs.get(x0, (self.B, NX), out=True)
        self._check(self._L.br2_batch_yaw_unwrap_host(self._h, a0))
        return x0

    def yaw_state(self):
        st = np.empty((self.B, 2), dtype=np.float32)
        self._check(self._L.br2_batch_yaw_get_state_host(self._h, _ptr(st)))
        return st

    def set_yaw_state(self, st):
        self._check(self._L.br2_batch_yaw_set_state_host(self._h, _ptr(_np(st, (self.B, 2), dtype=np.float32))))

    # -- RLS with variable forgetting factor (BLUEROV2_AMPC::RLSFF, bluerov2_ampc.cpp:731-1004) ---------------
    RLS_STRIDE = 80

    def rls_reset(self):
        self._check(self._L.br2_batch_rls_reset(self._h))

    def rls(self, meas, body_acc, compensate: bool = True, out=None):
        """One RLSFF() call for all instances, fed by the EKF state in the solver (call ``ekf`` first; the AMPC node runs
        EKF -> RLSFF -> solve, bluerov2_ampc_node.cpp:26-29).  Returns p[B,16]; ``out`` is in/out: with ``compensate`` False
        only p[:, 0:4] = 0 is written (the reference never fills p[4..15] in that case)."""
        if _is_torch(meas):
            import torch
            self._dev_check(meas, (self.B, 12)); self._dev_check(body_acc, (self.B, 6))
            if out is None:
                out = torch.zeros((self.B, NP), dtype=torch.float64, device=meas.device)
            stream = C.c_void_p(torch.cuda.current_stream(meas.device).cuda_stream)
            self._check(self._L.br2_batch_rls_device(self._h, _ptr(meas), _ptr(body_acc), _ptr(out), int(bool(compensate)), stream))
            return out
        H = self._hostargs.get
        (k0, a0), (k1, a1) = H(meas, (self.B, 12)), H(body_acc, (self.B, 6))
        if out is None:
            out = np.zeros((self.B, NP))
        _, o0 = H(out, (self.B, NP), out=True)
        self._check(self._L.br2_batch_rls_host(self._h, a0, a1, o0, int(bool(compensate))))
        return out

    def rls_state(self):
        """[B, 4 axes (X, Y, Z, N), 80]: theta[0:4], P[4:20], lambda[20], F[21], window counts[22:24], windows[24:29], [29:79]."""
        st = np.empty((self.B, 4, self.RLS_STRIDE))
        self._check(self._L.br2_batch_rls_get_state_host(self._h, _ptr(st)))
        return st

    def set_rls_state(self, st):
        self._check(self._L.br2_batch_rls_set_state_host(self._h, _ptr(_np(st, (self.B, 4, self.RLS_STRIDE)))))


def plant_step_replay(x, u, p, table, phase=None, tick: int = 0, h: float = 0.05, body_acc=None, lines=None):
    """Nominal plant on the device with the wrench series of applyBodyWrench mode 2 (bluerov2_dob.cpp:818-874): ``table`` [rows,4]
    CUDA tensor (fx, fy, fz, tz), row min(tick + phase[b], rows - 1) per instance.  In place on ``x``; torch's current stream."""
    import torch
    L = load_library()
    B = x.shape[0]
    for t, shp, dt in ((x, (B, NX), torch.float64), (u, (B, NU), torch.float64), (p, (B, NP), torch.float64), (table, (table.shape[0], 4), torch.float64)):
        if not t.is_cuda or t.dtype != dt or not t.is_contiguous() or tuple(t.shape) != shp:
            raise ValueError(f"expected contiguous CUDA {dt} tensor of shape {shp}")
    L.br2_plant_step_replay_device.restype = C.c_int
    L.br2_plant_step_replay_device.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_double,
                                               C.c_void_p, C.c_void_p, C.c_void_p]
    stream = C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)
    rc = L.br2_plant_step_replay_device(int(B), _ptr(x), _ptr(u), _ptr(p), _ptr(table), int(table.shape[0]), _ptr(phase), int(tick), float(h),
                                        _ptr(body_acc), _ptr(lines), stream)
    if rc != 0:
        raise SolverError(f"bluerov2_b200 error {rc}: {L.br2_last_error().decode()}")
    return x


def plant_step(x, u, p, h: float = 0.05, dist=None, wave=None, tick: int = 0, body_acc=None, lines=None):
    """Nominal plant on the device (torch CUDA tensors, in place on ``x``): one RK4 step of the OCP model per instance.
    ``wave`` = (amp[B,4], tau0[B]) adds the sampled wave wrench at ``tick``; ``body_acc`` [B,6] receives the finite-
    differenced body velocities; ``lines`` [B] int32 is incremented.  Enqueues on torch's current stream."""
    import torch
    L = load_library()
    B = x.shape[0]
    for t, shp, dt in ((x, (B, NX), torch.float64), (u, (B, NU), torch.float64), (p, (B, NP), torch.float64)):
        if not t.is_cuda or t.dtype != dt or not t.is_contiguous() or tuple(t.shape) != shp:
            raise ValueError(f"expected contiguous CUDA {dt} tensor of shape {shp}")
    amp, tau0 = wave if wave is not None else (None, None)
    stream = C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)
    rc = L.br2_plant_step_device(int(B), _ptr(x), _ptr(u), _ptr(p), _ptr(dist), _ptr(amp), _ptr(tau0), int(tick), float(h),
                                 _ptr(body_acc), _ptr(lines), stream)
    if rc != 0:
        raise SolverError(f"bluerov2_b200 error {rc}: {L.br2_last_error().decode()}")
    return x


# bluerov2_states/launch/config/imudo.yaml (the library bakes the same values when no parameters are passed)
ESKF_DEFAULTS = dict(q_p=0.001, q_v=0.001, q_r=0.001, q_q=0.001, q_xi=0.001, r_p=0.01, r_v=0.02, r_r=0.0006, r_th=0.012,
                     b_a=(-4.342596682195816e-07, -3.581072716118436e-18, -0.009999999990570729),
                     b_g=(-2.66013609366142e-20, -1.933924486945935e-19, -3.870624673211354e-16))


class _EskfParams(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("q_p", "q_v", "q_r", "q_q", "q_xi", "r_p", "r_v", "r_r", "r_th")] + [("b_a", C.c_double * 3), ("b_g", C.c_double * 3)]


class BatchEskf:
    """B independent IMU error-state Kalman filters (include/bluerov2_b200.h br2_eskf_*; reference: bluerov2_states/src/Eskf.cpp).
    numpy arrays in, numpy arrays out (host entry points); state [B,18] = p, v, R (row-major), xi; covariance [B,21,21]."""

    def __init__(self, batch: int, device: int = 0, **prm):
        self._L = L = load_library()
        V = C.c_void_p
        for name, args in (("br2_eskf_create", [C.POINTER(V), C.c_int, V, C.c_int]), ("br2_eskf_free", [V]),
                           ("br2_eskf_set_state_host", [V, V, V]), ("br2_eskf_get_state_host", [V, V, V]), ("br2_eskf_predict_host", [V, V]),
                           ("br2_eskf_update_host", [V] * 9)):
            getattr(L, name).restype = C.c_int
            getattr(L, name).argtypes = args
        self._h = V()
        par = None
        if prm:
            d = dict(ESKF_DEFAULTS); d.update(prm)
            par = _EskfParams(d["q_p"], d["q_v"], d["q_r"], d["q_q"], d["q_xi"], d["r_p"], d["r_v"], d["r_r"], d["r_th"],
                              (C.c_double * 3)(*d["b_a"]), (C.c_double * 3)(*d["b_g"]))
        rc = L.br2_eskf_create(C.byref(self._h), int(batch), C.byref(par) if par is not None else None, int(device))
        if rc != 0:
            raise SolverError(f"bluerov2_b200 error {rc}: {L.br2_last_error().decode()}")
        self.B = int(batch)

    def _check(self, rc):
        if rc != 0:
            raise SolverError(f"bluerov2_b200 error {rc}: {self._L.br2_last_error().decode()}")

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._L.br2_eskf_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_state(self, state=None, P=None):
        state = None if state is None else _np(state, (self.B, 18))
        P = None if P is None else _np(P, (self.B, 21, 21))
        self._check(self._L.br2_eskf_set_state_host(self._h, _ptr(state), _ptr(P)))

    def get_state(self):
        state, P = np.empty((self.B, 18)), np.empty((self.B, 21, 21))
        self._check(self._L.br2_eskf_get_state_host(self._h, _ptr(state), _ptr(P)))
        return state, P

    def predict(self, imu):
        self._check(self._L.br2_eskf_predict_host(self._h, _ptr(_np(imu, (self.B, 6)))))

    def update(self, gps_p, gps_v, R_meas, thrusts, imu_raw, R_gt):
        """-> (xi_world [B,3], innovation [B,12])"""
        xi, y = np.empty((self.B, 3)), np.empty((self.B, 12))
        a = [_np(gps_p, (self.B, 3)), _np(gps_v, (self.B, 3)), _np(np.reshape(R_meas, (self.B, 9)), (self.B, 9)), _np(thrusts, (self.B, 6)),
             _np(imu_raw, (self.B, 6)), _np(np.reshape(R_gt, (self.B, 9)), (self.B, 9))]
        self._check(self._L.br2_eskf_update_host(self._h, *[_ptr(v) for v in a], _ptr(xi), _ptr(y)))
        return xi, y


def device_count() -> int:
    return int(load_library().br2_device_count())


class _HostBlock:
    """owner of one br2_host_alloc block (freed when the last array viewing it goes away)"""

    def __init__(self, ptr, lib):
        self.ptr, self._lib = ptr, lib

    def __del__(self):
        try:
            self._lib.br2_host_free(C.c_void_p(self.ptr))
        except Exception:
            pass


def pinned_empty(shape, dtype=np.float64, write_combined: bool = False) -> np.ndarray:
    """numpy array in pinned host memory (br2_host_alloc) for the buffers of the host entry points.  write_combined: for large inputs the
    CPU only writes (reference windows): the copy engine reads them without snooping the CPU caches; reading such an array back on the
    CPU is slow."""
    L = load_library()
    shape = tuple(int(v) for v in (shape if isinstance(shape, (tuple, list)) else (shape,)))
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    p = C.c_void_p()
    rc = L.br2_host_alloc(C.byref(p), C.c_size_t(max(n, 1)), int(bool(write_combined)))
    if rc != 0:
        raise SolverError(f"bluerov2_b200 error {rc}: {L.br2_last_error().decode()}")
    buf = (C.c_char * max(n, 1)).from_address(p.value)
    buf._owner = _HostBlock(p.value, L)                      # the ctypes buffer (kept alive by the array) keeps the block alive
    return np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
