"""Regenerates tests/golden/*.npz.  Run HERE (needs /root/reference for the CasADi part):

    python tests/golden/make_golden.py

* casadi_vde.npz   -- inputs/outputs of the reference's OWN CasADi-generated functions
                      (oracle/_ref/libbluerov2_casadi_ref.so compiled from
                      /root/reference/bluerov2_dobmpc/scripts/c_generated_code/bluerov2_model/*.c).
                      These are reference outputs: they pin the dynamics/Jacobian layer.
* rti_cases.npz    -- RTI-step cases (inputs + oracle outputs at tol 1e-12, ERK routed through the CasADi VDE).
                      The NLP/QP layer of the reference (acados+HPIPM) cannot run here, so these pin the oracle
                      against itself over time and give the GPU tests fixed inputs; the solver-independent
                      certificate is the dense KKT check in tests/test_oracle_qp.py.
* ekf_cases.npz    -- EKF restatement outputs (Eigen absent: oracle-generated).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import Oracle, CasadiRef, NOMINAL_P, X_INIT  # noqa: E402
from bluerov2_b200 import workloads as wl  # noqa: E402
from bluerov2_b200 import traj  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def casadi_vectors():
    ref = CasadiRef()
    rng = np.random.default_rng(20261017)
    n = 96
    X = rng.uniform(-1, 1, (n, 12)) * np.array([5, 5, 5, 1.2, 1.2, 3.2, 2, 2, 2, 1, 1, 1])
    X[:, 2] -= 20
    X[:8, 6:9] = 0.0           # sign(0) = 0 branch of d|v|v/dv
    U = rng.uniform(-50, 50, (n, 4))
    P = np.tile(NOMINAL_P, (n, 1))
    P[:, :4] = rng.uniform(-20, 20, (n, 4))
    P[n // 2:, 4:] *= rng.uniform(0.7, 1.3, (n - n // 2, 12))
    Sx = rng.standard_normal((n, 144))
    Su = rng.standard_normal((n, 48))
    Sx[0] = np.eye(12).ravel()
    Su[0] = 0
    f = np.zeros((n, 12)); dSx = np.zeros((n, 144)); dSu = np.zeros((n, 48)); f_ode = np.zeros((n, 12))
    adj = np.zeros((n, 16)); lam = rng.standard_normal((n, 12))
    for i in range(n):
        f[i], dSx[i], dSu[i] = ref.vde_forw(X[i], Sx[i], Su[i], U[i], P[i])
        f_ode[i] = ref.ode(X[i], U[i], P[i])
        adj[i] = ref.vde_adj(X[i], lam[i], U[i], P[i])
    np.savez_compressed(os.path.join(HERE, "casadi_vde.npz"), x=X, u=U, p=P, Sx_cm=Sx, Su_cm=Su, f=f, dSx_cm=dSx,
                        dSu_cm=dSu, f_ode=f_ode, lam=lam, adj=adj)


def rti_cases():
    o = Oracle()
    o.use_casadi(CasadiRef())
    out = {}
    circ = traj.circle()
    # (a) the cold-start case of main_bluerov2.c / generate_c_code.py:89-93 on the circle reference
    x0 = np.array([-2, 0, -20, 0, 0, -1.570796, 0, 0, 0, 0, 0, 0.0])
    for N in (80, 40, 20, 10):
        Ts = wl.time_steps(N)
        X = np.tile(X_INIT, (N + 1, 1)).copy(); U = np.zeros((N, 4))
        st, info = o.rti_step(Ts, x0, circ[:N + 1], NOMINAL_P, X, U)
        assert st == 0
        out[f"cold_N{N}_X"] = X; out[f"cold_N{N}_U"] = U; out[f"cold_N{N}_info"] = info
    out["cold_x0"] = x0
    # (b) closed-loop batches: nominal and active-bound sets, 6 ticks each, N = 40 and N = 20
    for tag, N, B, spread, refname, seed in (("nom40", 40, 16, 0.5, "circle", 0), ("act40", 40, 16, 3.0, "circle", 0),
                                             ("lem40", 40, 8, 1.0, "lemniscate", 2), ("act20", 20, 8, 3.0, "circle", 5),
                                             ("act80", 80, 4, 2.0, "circle", 7)):
        Ts = wl.time_steps(N)
        w = wl.tracking_batch(B, N, seed=seed, reference=refname, pos_spread=spread)
        rng = np.random.default_rng(seed + 100)
        p = w["p"].copy()
        p[:, :4] = rng.uniform(-8, 8, (B, 4)) * (tag != "nom40")
        X = w["X"].copy(); U = w["U"].copy(); x0b = w["x0"].copy(); lines = w["lines"].copy()
        T = 5
        rec = {k: [] for k in ("x0", "lines", "Xout", "Uout", "info", "status")}
        out[f"{tag}_X0"] = X.copy(); out[f"{tag}_U0"] = U.copy()
        for t in range(T):
            yref = traj.window_batch(w["traj"], lines, N)
            rec["x0"].append(x0b.copy()); rec["lines"].append(lines.copy())
            st, info, _ = o.rti_step_batch(Ts, x0b, yref, p, X, U)
            assert (st == 0).all(), (tag, t, st)
            rec["Xout"].append(X.copy()); rec["Uout"].append(U.copy()); rec["info"].append(info.copy()); rec["status"].append(st.copy())
            for i in range(B):
                x0b[i] = o.erk4(x0b[i], U[i, 0], p[i], 0.05)
            lines = lines + 1
        for k, v in rec.items():
            out[f"{tag}_{k}"] = np.stack(v)
        out[f"{tag}_p"] = p
        out[f"{tag}_N"] = np.array(N)
        out[f"{tag}_ref"] = np.array(refname)
    np.savez_compressed(os.path.join(HERE, "rti_cases.npz"), **out)


def ekf_cases():
    o = Oracle()
    rng = np.random.default_rng(77)
    B, T = 12, 8
    ex = np.zeros((B, 18)); eP = np.zeros((B, 18, 18))
    for i in range(B):
        ex[i], eP[i] = o.ekf_init()
    rec = {k: [] for k in ("thr", "meas", "acc", "ex", "eP", "wf")}
    pose = np.tile(np.array([0, 0, -20, 0, 0, 0, 0, 0, 0, 0, 0, 0.0]), (B, 1)) + rng.uniform(-0.2, 0.2, (B, 12))
    for t in range(T):
        thr = rng.uniform(-15, 15, (B, 6))
        acc = rng.uniform(-1, 1, (B, 6))
        pose = pose + rng.uniform(-0.05, 0.05, (B, 12))
        wf = o.ekf_step_batch(ex, eP, thr, pose, acc)
        rec["thr"].append(thr); rec["meas"].append(pose.copy()); rec["acc"].append(acc)
        rec["ex"].append(ex.copy()); rec["eP"].append(eP.copy()); rec["wf"].append(wf)
    np.savez_compressed(os.path.join(HERE, "ekf_cases.npz"), **{k: np.stack(v) for k, v in rec.items()})


if __name__ == "__main__":
    casadi_vectors()
    rti_cases()
    ekf_cases()
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))
