"""ctypes front-end of the CPU oracle.  TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(HERE, "libbluerov2_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libbluerov2_casadi_ref.so")

NX, NU, NP, NY = 12, 4, 16, 16

# bluerov2_dob.cpp:340-353 (p4..p15); p0..p3 = disturbance, zero when not compensating
NOMINAL_P = np.array([0, 0, 0, 0, 1.7182, 0, 5.468, 0.4006, -11.7391, -20, -31.8678, -5,
                      -18.18, -21.66, -36.99, -1.55], dtype=np.float64)
# acados_solver_bluerov2.c:424-459 (W), :468-479 (W_e), :547-571 (bounds), :681-708 (initial guess)
W_DEFAULT = np.array([300, 480, 200, 10, 10, 200, 40, 40, 10, 10, 10, 10, 1, 1, 0.1, 0.05], dtype=np.float64)
WE_DEFAULT = W_DEFAULT[:12].copy()
LBU = np.full(4, -50.0)
UBU = np.full(4, 50.0)
X_INIT = np.array([0, 0, -20, 0, 0, 0, 0, 0, 0, 0, 0, 0], dtype=np.float64)


def build(verbose: bool = False) -> None:
    """Compile the C restatement and, when /root/reference is present, oracle/_ref (make all)."""
    out = subprocess.run(["make", "-C", HERE, "all"], capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + out.stdout + out.stderr)
    if verbose:
        print(out.stdout)


def _p(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _c(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None:
        assert a.shape == tuple(shape), (a.shape, shape)
    return a


class CasadiRef:
    """The reference's CasADi-generated functions (bluerov2_model/*.c, bluerov2_cost/*.c), compiled in place."""

    def __init__(self, path: str = REF_SO):
        if not os.path.exists(path):
            raise FileNotFoundError(path + " (run `make -C oracle ref` where /root/reference exists)")
        self.lib = C.CDLL(path)
        self._sig = [C.POINTER(C.POINTER(C.c_double)), C.POINTER(C.POINTER(C.c_double)), C.c_void_p, C.c_void_p, C.c_int]

    def _call(self, name, ins, out_shapes):
        fn = getattr(self.lib, name)
        fn.argtypes = self._sig
        fn.restype = C.c_int
        ins = [np.ascontiguousarray(a, dtype=np.float64) for a in ins]
        outs = [np.zeros(s, dtype=np.float64) for s in out_shapes]
        arg = (C.POINTER(C.c_double) * len(ins))(*[_p(a) for a in ins])
        res = (C.POINTER(C.c_double) * len(outs))(*[_p(a) for a in outs])
        rc = fn(arg, res, None, None, 0)
        assert rc == 0
        return outs

    def ode(self, x, u, p):
        return self._call("bluerov2_expl_ode_fun", [x, u, p], [(12,)])[0]

    def vde_forw(self, x, Sx_cm, Su_cm, u, p):
        """Sx_cm/Su_cm are flat column-major; returns (f, dSx_cm, dSu_cm)."""
        return self._call("bluerov2_expl_vde_forw", [x, Sx_cm, Su_cm, u, p], [(12,), (144,), (48,)])

    def vde_adj(self, x, lam, u, p):
        return self._call("bluerov2_expl_vde_adj", [x, lam, u, p], [(16,)])[0]

    def cost_y(self, x, u, p):
        return self._call("bluerov2_cost_y_fun", [x, u, np.zeros(0), p], [(16,)])[0]

    def cost_y_e(self, x, p):
        return self._call("bluerov2_cost_y_e_fun", [x, np.zeros(0), np.zeros(0), p], [(12,)])[0]

    def fn_ptr(self, name="bluerov2_expl_vde_forw"):
        return C.cast(getattr(self.lib, name), C.c_void_p)


class Oracle:
    def __init__(self, path: str = PORT_SO):
        if not os.path.exists(path):
            build()
        self.lib = L = C.CDLL(path)
        D = C.POINTER(C.c_double)
        I = C.POINTER(C.c_int)
        L.orc_ode.argtypes = [D, D, D, D]
        L.orc_jac.argtypes = [D, D, D, D, D]
        L.orc_vde_forw_cm.argtypes = [D] * 8
        L.orc_set_casadi_vde.argtypes = [C.c_void_p]
        L.orc_erk4_sens.argtypes = [D, D, D, C.c_double, D, D, D]
        L.orc_erk4.argtypes = [D, D, D, C.c_double, D]
        L.orc_linearize.argtypes = [C.c_int, D, D, C.c_int, D, D, D, D, D]
        L.orc_rti_step.argtypes = [C.c_int, D, D, D, D, D, D, D, D, C.c_int, D, D, C.c_int, C.c_double, D]
        L.orc_rti_step.restype = C.c_int
        L.orc_rti_step_batch.argtypes = [C.c_int, C.c_int, D, D, D, D, D, D, D, D, D, D, C.c_int, C.c_double, I, D, C.c_int]
        L.orc_rti_step_batch.restype = C.c_int
        L.orc_thrust_alloc.argtypes = [D, D]
        L.orc_ekf_init.argtypes = [D, D]
        L.orc_ekf_f.argtypes = [D, D, D]
        L.orc_ekf_h.argtypes = [D, D, D]
        L.orc_ekf_step.argtypes = [D, D, D, D, D, D]
        L.orc_ekf_step_batch.argtypes = [C.c_int, D, D, D, D, D, D, C.c_int]
        L.orc_ekf_set_model.argtypes = [C.c_int]
        L.orc_set_qp_mode.argtypes = [C.c_int]
        L.orc_get_qp_mode.restype = C.c_int
        L.orc_yaw_unwrap.argtypes = [C.POINTER(C.c_float), C.c_double]
        L.orc_yaw_unwrap.restype = C.c_double
        L.orc_rls_init.argtypes = [D]
        L.orc_rls_step.argtypes = [D, D, D, D, C.c_int, D]
        L.orc_rls_step_batch.argtypes = [C.c_int, D, D, D, D, C.c_int, D]
        self._ref = None

    # -- routing of the ERK through the reference's CasADi VDE ------------------------------------------
    def use_casadi(self, ref: CasadiRef | None):
        self._ref = ref  # keep alive
        self.lib.orc_set_casadi_vde(ref.fn_ptr() if ref is not None else None)

    # -- model -----------------------------------------------------------------------------------------
    def ode(self, x, u, p):
        x, u, p = _c(x, (12,)), _c(u, (4,)), _c(p, (16,))
        f = np.zeros(12)
        self.lib.orc_ode(_p(x), _p(u), _p(p), _p(f))
        return f

    def jac(self, x, u, p):
        x, u, p = _c(x, (12,)), _c(u, (4,)), _c(p, (16,))
        Jx, Ju = np.zeros((12, 12)), np.zeros((12, 4))
        self.lib.orc_jac(_p(x), _p(u), _p(p), _p(Jx), _p(Ju))
        return Jx, Ju

    def vde_forw_cm(self, x, Sx_cm, Su_cm, u, p):
        x, u, p = _c(x, (12,)), _c(u, (4,)), _c(p, (16,))
        Sx_cm, Su_cm = _c(Sx_cm).reshape(144), _c(Su_cm).reshape(48)
        f, dSx, dSu = np.zeros(12), np.zeros(144), np.zeros(48)
        self.lib.orc_vde_forw_cm(_p(x), _p(Sx_cm), _p(Su_cm), _p(u), _p(p), _p(f), _p(dSx), _p(dSu))
        return f, dSx, dSu

    def erk4_sens(self, x, u, p, h):
        x, u, p = _c(x, (12,)), _c(u, (4,)), _c(p, (16,))
        xn, A, B = np.zeros(12), np.zeros((12, 12)), np.zeros((12, 4))
        self.lib.orc_erk4_sens(_p(x), _p(u), _p(p), float(h), _p(xn), _p(A), _p(B))
        return xn, A, B

    def erk4(self, x, u, p, h):
        x, u, p = _c(x, (12,)), _c(u, (4,)), _c(p, (16,))
        xn = np.zeros(12)
        self.lib.orc_erk4(_p(x), _p(u), _p(p), float(h), _p(xn))
        return xn

    def linearize(self, Ts, p, X, U):
        N = len(Ts)
        Ts, X, U = _c(Ts, (N,)), _c(X, (N + 1, 12)), _c(U, (N, 4))
        p = _c(p)
        stride = 0 if p.ndim == 1 else 16
        A, B, b = np.zeros((N, 12, 12)), np.zeros((N, 12, 4)), np.zeros((N, 12))
        self.lib.orc_linearize(N, _p(Ts), _p(p), stride, _p(X), _p(U), _p(A), _p(B), _p(b))
        return A, B, b

    # -- RTI step --------------------------------------------------------------------------------------
    def rti_step(self, Ts, x0, yref, p, X, U, W=W_DEFAULT, We=WE_DEFAULT, lbu=LBU, ubu=UBU,
                 max_iter=50, tol=1e-12):
        """One SQP-RTI step.  X ((N+1)x12) and U (Nx4) are updated IN PLACE.  Returns (status, info[8])."""
        N = len(Ts)
        Ts, x0, yref = _c(Ts, (N,)), _c(x0, (12,)), _c(yref, (N + 1, 16))
        p = _c(p)
        stride = 0 if p.ndim == 1 else 16
        assert X.dtype == np.float64 and X.flags.c_contiguous and X.shape == (N + 1, 12)
        assert U.dtype == np.float64 and U.flags.c_contiguous and U.shape == (N, 4)
        W, We, lbu, ubu = _c(W, (16,)), _c(We, (12,)), _c(lbu, (4,)), _c(ubu, (4,))
        info = np.zeros(8)
        st = self.lib.orc_rti_step(N, _p(Ts), _p(W), _p(We), _p(lbu), _p(ubu), _p(x0), _p(yref), _p(p), stride,
                                   _p(X), _p(U), int(max_iter), float(tol), _p(info))
        return st, info

    def rti_step_batch(self, Ts, x0, yref, p, X, U, W=W_DEFAULT, We=WE_DEFAULT, lbu=LBU, ubu=UBU,
                       max_iter=50, tol=1e-12, nthreads=0):
        N = len(Ts)
        nb = x0.shape[0]
        Ts, x0, yref, p = _c(Ts, (N,)), _c(x0, (nb, 12)), _c(yref, (nb, N + 1, 16)), _c(p, (nb, 16))
        assert X.dtype == np.float64 and X.flags.c_contiguous and X.shape == (nb, N + 1, 12)
        assert U.dtype == np.float64 and U.flags.c_contiguous and U.shape == (nb, N, 4)
        W, We, lbu, ubu = _c(W, (16,)), _c(We, (12,)), _c(lbu, (4,)), _c(ubu, (4,))
        status = np.zeros(nb, dtype=np.int32)
        info = np.zeros((nb, 8))
        used = self.lib.orc_rti_step_batch(nb, N, _p(Ts), _p(W), _p(We), _p(lbu), _p(ubu), _p(x0), _p(yref), _p(p),
                                           _p(X), _p(U), int(max_iter), float(tol),
                                           status.ctypes.data_as(C.POINTER(C.c_int)), _p(info), int(nthreads))
        return status, info, used

    def thrust_alloc(self, u0):
        u0 = _c(u0, (4,))
        t = np.zeros(6)
        self.lib.orc_thrust_alloc(_p(u0), _p(t))
        return t

    # -- EKF -------------------------------------------------------------------------------------------
    def ekf_init(self):
        x, P = np.zeros(18), np.zeros((18, 18))
        self.lib.orc_ekf_init(_p(x), _p(P))
        return x, P

    def ekf_f(self, x, u):
        x, u = _c(x, (18,)), _c(u, (6,))
        o = np.zeros(18)
        self.lib.orc_ekf_f(_p(x), _p(u), _p(o))
        return o

    def ekf_h(self, x, acc):
        x, acc = _c(x, (18,)), _c(acc, (6,))
        o = np.zeros(18)
        self.lib.orc_ekf_h(_p(x), _p(acc), _p(o))
        return o

    def ekf_step(self, esti_x, esti_P, thrusts, meas12, body_acc):
        """esti_x (18,), esti_P (18,18) updated IN PLACE; returns world-frame disturbance (6,)."""
        assert esti_x.dtype == np.float64 and esti_x.shape == (18,) and esti_x.flags.c_contiguous
        assert esti_P.dtype == np.float64 and esti_P.shape == (18, 18) and esti_P.flags.c_contiguous
        thrusts, meas12, body_acc = _c(thrusts, (6,)), _c(meas12, (12,)), _c(body_acc, (6,))
        wf = np.zeros(6)
        self.lib.orc_ekf_step(_p(esti_x), _p(esti_P), _p(thrusts), _p(meas12), _p(body_acc), _p(wf))
        return wf

    def ekf_step_batch(self, esti_x, esti_P, thrusts, meas12, body_acc, nthreads=0):
        nb = esti_x.shape[0]
        assert esti_x.dtype == np.float64 and esti_x.shape == (nb, 18) and esti_x.flags.c_contiguous
        assert esti_P.dtype == np.float64 and esti_P.shape == (nb, 18, 18) and esti_P.flags.c_contiguous
        thrusts, meas12, body_acc = _c(thrusts, (nb, 6)), _c(meas12, (nb, 12)), _c(body_acc, (nb, 6))
        wf = np.zeros((nb, 6))
        self.lib.orc_ekf_step_batch(nb, _p(esti_x), _p(esti_P), _p(thrusts), _p(meas12), _p(body_acc), _p(wf), int(nthreads))
        return wf

    def set_qp_mode(self, mode: int):
        """0 = Riccati interior-point iteration (default), 1 = full condensing + dense interior-point iteration (the cost profile of the
        reference's FULL_CONDENSING_HPIPM); process-wide, read by rti_step / rti_step_batch"""
        self.lib.orc_set_qp_mode(int(mode))

    def ekf_set_model(self, model: int):
        """0 = BLUEROV2_DOB filter (default), 1 = BLUEROV2_AMPC filter (Dl = 0, no quadratic damping). Process-wide."""
        self.lib.orc_ekf_set_model(int(model))

    # -- RLS with variable forgetting factor (BLUEROV2_AMPC::RLSFF) -----------------------------------------
    RLS_STRIDE = 80
    RLS_THETA, RLS_P, RLS_LAMBDA, RLS_F, RLS_NN, RLS_ND, RLS_EN, RLS_ED = 0, 4, 20, 21, 22, 23, 24, 29

    def rls_init(self, nb=None):
        """state (4, 80) -- or (nb, 4, 80) -- as the AMPC constructor leaves it (theta 0, P = I, lambda 0.9)."""
        st = np.zeros((4, self.RLS_STRIDE))
        self.lib.orc_rls_init(_p(st))
        return st if nb is None else np.ascontiguousarray(np.broadcast_to(st, (nb, 4, self.RLS_STRIDE)))

    def rls_step_batch(self, state, esti_x, body_acc, meas12, compensate=True, p_out=None):
        """state (nb,4,80) updated IN PLACE; returns p_out (nb,16) (pass one in to see which entries are left untouched)."""
        nb = state.shape[0]
        assert state.dtype == np.float64 and state.shape == (nb, 4, self.RLS_STRIDE) and state.flags.c_contiguous
        esti_x, body_acc, meas12 = _c(esti_x, (nb, 18)), _c(body_acc, (nb, 6)), _c(meas12, (nb, 12))
        if p_out is None:
            p_out = np.zeros((nb, 16))
        assert p_out.dtype == np.float64 and p_out.shape == (nb, 16) and p_out.flags.c_contiguous
        self.lib.orc_rls_step_batch(nb, _p(state), _p(esti_x), _p(body_acc), _p(meas12), int(bool(compensate)), _p(p_out))
        return p_out

    # -- continuous yaw (BLUEROV2_DOB::solve, bluerov2_dob.cpp:272-304) ------------------------------------
    def yaw_unwrap_batch(self, state, psi):
        """state (nb,2) float32 = (pre_yaw, yaw_sum) updated IN PLACE; psi (nb,) measured yaw in (-pi, pi]; returns x0[psi] (nb,)"""
        nb = state.shape[0]
        assert state.dtype == np.float32 and state.shape == (nb, 2) and state.flags.c_contiguous
        psi = _c(psi, (nb,))
        out = np.empty(nb)
        for i in range(nb):
            out[i] = self.lib.orc_yaw_unwrap(state[i].ctypes.data_as(C.POINTER(C.c_float)), float(psi[i]))
        return out

