"""Independent numpy check of one RTI step's QP: condense to a dense box-QP in du and solve it WITHOUT any
Riccati machinery (projected Newton / active-set on the 4N-variable problem).  Used only by tests."""
import numpy as np

NX, NU = 12, 4


def build_qp(A, B, b, Ts, W, We, X, U, yref, x0, lbu, ubu):
    N = len(Ts)
    Qd = np.vstack([Ts[:, None] * W[None, :12], We[None, :]])
    Rd = Ts[:, None] * W[None, 12:]
    q = np.vstack([Qd[:N] * (X[:N] - yref[:N, :12]), (We * (X[N] - yref[N, :12]))[None, :]])
    r = Rd * (U - yref[:N, 12:])
    return dict(N=N, A=A, B=B, b=b, Qd=Qd, Rd=Rd, q=q, r=r, lb=lbu[None, :] - U, ub=ubu[None, :] - U, dx0=x0 - X[0])


def condense(qp):
    """dx = c + G du (stacked over stages 0..N), so that the QP becomes min 1/2 du'H du + g'du."""
    N = qp["N"]
    G = np.zeros(((N + 1) * NX, N * NU))
    c = np.zeros((N + 1) * NX)
    c[:NX] = qp["dx0"]
    for k in range(N):
        r0, r1 = k * NX, (k + 1) * NX
        G[r1:r1 + NX, :] = qp["A"][k] @ G[r0:r1, :]
        G[r1:r1 + NX, k * NU:(k + 1) * NU] += qp["B"][k]
        c[r1:r1 + NX] = qp["A"][k] @ c[r0:r1] + qp["b"][k]
    Qbar = qp["Qd"].ravel()
    H = G.T @ (Qbar[:, None] * G) + np.diag(qp["Rd"].ravel())
    g = G.T @ (Qbar * c + qp["q"].ravel()) + qp["r"].ravel()
    return H, g, G, c


def solve_box_qp(H, g, lb, ub, iters=200):
    """Primal active-set with exact Newton on the free set (finite termination for strictly convex box QPs)."""
    n = len(g)
    v = np.clip(np.zeros(n), lb, ub)
    for _ in range(iters):
        grad = H @ v + g
        at_l = (v <= lb) & (grad > 0)
        at_u = (v >= ub) & (grad < 0)
        free = ~(at_l | at_u)
        if np.abs(grad[free]).max(initial=0.0) < 1e-11 * max(1.0, np.abs(g).max()):
            break
        d = np.zeros(n)
        d[free] = -np.linalg.solve(H[np.ix_(free, free)], grad[free])
        # longest feasible step, then clip (projected Newton)
        with np.errstate(divide="ignore", invalid="ignore"):
            amax = np.where(d > 0, (ub - v) / d, np.where(d < 0, (lb - v) / d, np.inf))
        a = min(1.0, amax[free].min(initial=np.inf))
        v = np.clip(v + a * d, lb, ub)
    grad = H @ v + g
    lam_l = np.where(v <= lb, np.maximum(grad, 0), 0.0)
    lam_u = np.where(v >= ub, np.maximum(-grad, 0), 0.0)
    return v, lam_l, lam_u


def kkt_residuals(H, g, lb, ub, v):
    grad = H @ v + g
    lam_l = np.where(v - lb < 1e-9, np.maximum(grad, 0), 0.0)
    lam_u = np.where(ub - v < 1e-9, np.maximum(-grad, 0), 0.0)
    stat = np.abs(grad - lam_l + lam_u).max()
    feas = max(np.maximum(lb - v, 0).max(), np.maximum(v - ub, 0).max())
    return stat, feas
