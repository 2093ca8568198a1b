#!/usr/bin/env python
"""Who is right where the CUDA path and the oracle disagree on the ill-posed soak case (N = 10, tilted start, 1.5 m)?  For the worst
(tick, instance) pairs with status 0 on both sides: the condensed QP of that step, both answers' objective and KKT residuals, scipy's
bounded-variable least squares as the referee, the condition number of the condensed Hessian."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from bluerov2_b200 import solver as S, traj, workloads as wl
from oracle import Oracle, W_DEFAULT, WE_DEFAULT, LBU, UBU
from qp_dense import build_qp, condense, kkt_residuals
from scipy.linalg import cholesky, solve_triangular
from scipy.optimize import lsq_linear

o = Oracle()
B, N, T, spread = 1024, 10, 20, 1.5
w = wl.tracking_batch(B, N, seed=7, pos_spread=spread)
Ts = wl.time_steps(N)
s = S.BatchSolver(B, N); s.set_iterate(w["X"], w["U"])
Xo, Uo = w["X"].copy(), w["U"].copy()
x0, lines = w["x0"].copy(), w["lines"].copy()
out = []
for t in range(T):
    yref = traj.window_batch(w["traj"], lines, N)
    Xb, Ub = Xo.copy(), Uo.copy()                        # linearisation point of this tick (both sides)
    u0, th, st = s.solve(x0, yref, w["p"])
    it, info = s.stats()
    Xg, Ug = s.get_iterate()
    sto, _, used = o.rti_step_batch(Ts, x0, yref, w["p"], Xo, Uo)
    okk = (st == 0) & (sto == 0)
    err = np.where(okk, np.abs(u0 - Uo[:, 0]).max(axis=1), 0.0)
    for i in np.argsort(-err)[:2]:
        if err[i] < 1e-6:
            continue
        A, Bm, b = o.linearize(Ts, w["p"][i], Xb[i], Ub[i])
        qp = build_qp(A, Bm, b, Ts, W_DEFAULT, WE_DEFAULT, Xb[i], Ub[i], yref[i], x0[i], LBU, UBU)
        H, g, G, c = condense(qp)
        lb, ub = qp["lb"].ravel(), qp["ub"].ravel()
        obj = lambda v: 0.5 * v @ H @ v + g @ v       # noqa: E731
        dg, do = (Ug[i] - Ub[i]).ravel(), (Uo[i] - Ub[i]).ravel()
        rec = {"tick": t, "inst": int(i), "u0_err": float(err[i]), "gpu_iters": int(it[i]), "cond_H": float(np.linalg.cond(H)),
               "finite_linearisation": bool(np.isfinite(H).all()), "max_abs_iterate": float(np.abs(Xb[i]).max()),
               "obj_gpu": float(obj(dg)), "obj_oracle": float(obj(do)),
               "box_violation_gpu": float(np.maximum(lb - dg, dg - ub).max()), "box_violation_oracle": float(np.maximum(lb - do, do - ub).max())}
        try:
            L = cholesky(H, lower=True)
            res = lsq_linear(L.T, -solve_triangular(L, g, lower=True), bounds=(lb, ub), method="bvls", tol=1e-15, max_iter=5000)
            rec.update({"obj_bvls": float(obj(res.x)), "gpu_minus_bvls": float(np.abs(dg - res.x).max()), "oracle_minus_bvls": float(np.abs(do - res.x).max())})
        except Exception as e:
            rec["bvls"] = repr(e)[:80]
        out.append(rec)
    dead = ~np.isfinite(Xo).all(axis=(1, 2)) | ~np.isfinite(Uo).all(axis=(1, 2)) | (np.abs(Xo).max(axis=(1, 2)) > 1e6)
    if dead.any():
        Xo[dead] = np.where(np.isfinite(x0[dead]), x0[dead], 0.0)[:, None, :]; Uo[dead] = 0.0
    for i in range(B):
        xn = o.erk4(x0[i], Uo[i, 0], w["p"][i], 0.05)
        x0[i] = xn if np.isfinite(xn).all() and np.abs(xn).max() < 1e6 else w["x0"][i]
    lines = lines + 1
    s.set_iterate(Xo, Uo)
for r in sorted(out, key=lambda r: -r["u0_err"])[:12]:
    print(json.dumps(r))
