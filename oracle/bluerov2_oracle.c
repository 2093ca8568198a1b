/* oracle/bluerov2_oracle.c -- plain-C restatement of the reference's SQP-RTI hot path and EKF.
 *
 * TEST INFRASTRUCTURE ONLY (see bluerov2_oracle.h for the rules and the parity-pin status).
 * Written for clarity: dense 12x12 loops, everything stored, nothing shared with the CUDA product.
 */
#include "bluerov2_oracle.h"
#include <math.h>
#include <string.h>
#include <stdlib.h>
#include <stdio.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define NX ORC_NX
#define NU ORC_NU
#define NP ORC_NP

/* ------------------------------------------------------------------------------------------------
 * OCP model -- restates bluerov2_dobmpc/scripts/bluerov2.py:77-137 term by term
 * (same expression tree as the generated bluerov2_expl_ode_fun.c:66-321).
 * Quirks kept on purpose: dphi uses sin(psi) (bluerov2.py:133), dtheta = cos(phi) q + sin(phi) r (:134),
 * no Coriolis in du/dv/dw (:119-121 commented out), K rounded to 0.707/0.167/0.175 with zero roll and
 * pitch rows (:95-100), dp/dq carry neither damping nor thrust (:126-127).
 * ---------------------------------------------------------------------------------------------- */
static const double M_MASS = 11.26, IX = 0.3, IY = 0.63, IZ = 0.58, ZG = 0.02, GRAV = 9.81;
static const double BUOY = 0.66;
static const double RC = 0.026546960744430276;
static const double KM[6][6] = {
    {0.707, 0.707, -0.707, -0.707, 0, 0}, {0.707, -0.707, 0.707, -0.707, 0, 0}, {0, 0, 0, 0, 1, 1},
    {0, 0, 0, 0, 0, 0},                   {0, 0, 0, 0, 0, 0},                   {0.167, -0.167, -0.175, 0.175, 0, 0}};

void orc_ode(const double *x, const double *u, const double *p, double *f)
{
    const double phi = x[3], th = x[4], psi = x[5];
    const double uu = x[6], v = x[7], w = x[8], pp = x[9], q = x[10], r = x[11];
    double t[6], Kt[6];
    t[0] = (-u[0] + u[1] + u[3]) / RC;
    t[1] = (-u[0] - u[1] - u[3]) / RC;
    t[2] = (u[0] + u[1] - u[3]) / RC;
    t[3] = (u[0] - u[1] + u[3]) / RC;
    t[4] = -u[2] / RC;
    t[5] = -u[2] / RC;
    for (int i = 0; i < 6; i++) {
        double s = 0;
        for (int j = 0; j < 6; j++) s += KM[i][j] * t[j];
        Kt[i] = s;
    }
    const double sphi = sin(phi), cphi = cos(phi), sth = sin(th), cth = cos(th), spsi = sin(psi), cpsi = cos(psi);
    /* kinematics, bluerov2.py:130-135 */
    f[0] = (cpsi * cth) * uu + (-spsi * cphi + cpsi * sth * sphi) * v + (spsi * sphi + cpsi * cphi * sth) * w;
    f[1] = (spsi * cth) * uu + (cpsi * cphi + sphi * sth * spsi) * v + (-cpsi * sphi + sth * spsi * cphi) * w;
    f[2] = (-sth) * uu + (cth * sphi) * v + (cth * cphi) * w;
    f[3] = pp + (spsi * sth / cth) * q + cphi * sth / cth * r;
    f[4] = cphi * q + sphi * r;
    f[5] = (sphi / cth) * q + (cphi / cth) * r;
    /* kinetics, bluerov2.py:123-128 */
    f[6] = 1 / (M_MASS + p[4]) * (Kt[0] - BUOY * sth + p[0] + p[8] * uu + p[12] * fabs(uu) * uu);
    f[7] = 1 / (M_MASS + p[5]) * (Kt[1] + BUOY * cth * sphi + p[1] + p[9] * v + p[13] * fabs(v) * v);
    f[8] = 1 / (M_MASS + p[6]) * (Kt[2] + BUOY * cth * cphi + p[2] + p[10] * w + p[14] * fabs(w) * w);
    f[9] = 1 / IX * (Kt[3] + (IY - IZ) * q * r - M_MASS * ZG * GRAV * cth * sphi);
    f[10] = 1 / IY * (Kt[4] + (IZ - IX) * pp * r - M_MASS * ZG * GRAV * sth);
    f[11] = 1 / (IZ + p[7]) * (Kt[5] - (IY - IX) * pp * q + p[3] + p[11] * r + p[15] * fabs(r) * r);
}

/* Analytic Jacobian of the above.  d(|v| v)/dv = 2|v|, which is what CasADi's sign(v)*v + |v| evaluates to
 * (bluerov2_expl_vde_forw.c:65, sign(0) = 0).  Sparsity: 48 non-zeros in Jx (columns 0..2 identically 0),
 * 5 constant non-zeros in Ju. */
void orc_jac(const double *x, const double *u, const double *p, double *Jx, double *Ju)
{
    (void)u;
    const double phi = x[3], th = x[4], psi = x[5];
    const double uu = x[6], v = x[7], w = x[8], pp = x[9], q = x[10], r = x[11];
    const double sphi = sin(phi), cphi = cos(phi), sth = sin(th), cth = cos(th), spsi = sin(psi), cpsi = cos(psi);
    const double tth = sth / cth, sec2 = 1.0 / (cth * cth);
    memset(Jx, 0, sizeof(double) * NX * NX);
    memset(Ju, 0, sizeof(double) * NX * NU);
#define J(i, j) Jx[(i) * NX + (j)]
    /* row 0: dx */
    J(0, 3) = (spsi * sphi + cpsi * sth * cphi) * v + (spsi * cphi - cpsi * sphi * sth) * w;
    J(0, 4) = -cpsi * sth * uu + cpsi * cth * sphi * v + cpsi * cphi * cth * w;
    J(0, 5) = -spsi * cth * uu + (-cpsi * cphi - spsi * sth * sphi) * v + (cpsi * sphi - spsi * cphi * sth) * w;
    J(0, 6) = cpsi * cth;
    J(0, 7) = -spsi * cphi + cpsi * sth * sphi;
    J(0, 8) = spsi * sphi + cpsi * cphi * sth;
    /* row 1: dy */
    J(1, 3) = (-cpsi * sphi + cphi * sth * spsi) * v + (-cpsi * cphi - sth * spsi * sphi) * w;
    J(1, 4) = -spsi * sth * uu + sphi * cth * spsi * v + cth * spsi * cphi * w;
    J(1, 5) = cpsi * cth * uu + (-spsi * cphi + sphi * sth * cpsi) * v + (spsi * sphi + sth * cpsi * cphi) * w;
    J(1, 6) = spsi * cth;
    J(1, 7) = cpsi * cphi + sphi * sth * spsi;
    J(1, 8) = -cpsi * sphi + sth * spsi * cphi;
    /* row 2: dz */
    J(2, 3) = cth * cphi * v - cth * sphi * w;
    J(2, 4) = -cth * uu - sth * sphi * v - sth * cphi * w;
    J(2, 6) = -sth;
    J(2, 7) = cth * sphi;
    J(2, 8) = cth * cphi;
    /* row 3: dphi = p + sin(psi) tan(th) q + cos(phi) tan(th) r */
    J(3, 3) = -sphi * tth * r;
    J(3, 4) = (spsi * q + cphi * r) * sec2;
    J(3, 5) = cpsi * tth * q;
    J(3, 9) = 1.0;
    J(3, 10) = spsi * tth;
    J(3, 11) = cphi * tth;
    /* row 4: dtheta = cos(phi) q + sin(phi) r */
    J(4, 3) = -sphi * q + cphi * r;
    J(4, 10) = cphi;
    J(4, 11) = sphi;
    /* row 5: dpsi = (sin(phi) q + cos(phi) r)/cos(th) */
    J(5, 3) = (cphi * q - sphi * r) / cth;
    J(5, 4) = (sphi * q + cphi * r) * sth * sec2;
    J(5, 10) = sphi / cth;
    J(5, 11) = cphi / cth;
    /* rows 6..8: du, dv, dw */
    const double imx = 1 / (M_MASS + p[4]), imy = 1 / (M_MASS + p[5]), imz = 1 / (M_MASS + p[6]);
    J(6, 4) = -BUOY * cth * imx;
    J(6, 6) = (p[8] + 2 * p[12] * fabs(uu)) * imx;
    J(7, 3) = BUOY * cth * cphi * imy;
    J(7, 4) = -BUOY * sth * sphi * imy;
    J(7, 7) = (p[9] + 2 * p[13] * fabs(v)) * imy;
    J(8, 3) = -BUOY * cth * sphi * imz;
    J(8, 4) = -BUOY * sth * cphi * imz;
    J(8, 8) = (p[10] + 2 * p[14] * fabs(w)) * imz;
    /* rows 9..11: dp, dq, dr */
    const double mzg = M_MASS * ZG * GRAV, imn = 1 / (IZ + p[7]);
    J(9, 3) = -mzg * cth * cphi / IX;
    J(9, 4) = mzg * sth * sphi / IX;
    J(9, 10) = (IY - IZ) * r / IX;
    J(9, 11) = (IY - IZ) * q / IX;
    J(10, 4) = -mzg * cth / IY;
    J(10, 9) = (IZ - IX) * r / IY;
    J(10, 11) = (IZ - IX) * pp / IY;
    J(11, 9) = -(IY - IX) * q * imn;
    J(11, 10) = -(IY - IX) * pp * imn;
    J(11, 11) = (p[11] + 2 * p[15] * fabs(r)) * imn;
#undef J
    /* d f / d u : thrusts t0..t5 are linear in u (bluerov2.py:103-115) */
    Ju[6 * NU + 0] = (0.707 * (-1 - 1 - 1 - 1) / RC) * imx;
    Ju[7 * NU + 1] = (0.707 * (1 + 1 + 1 + 1) / RC) * imy;
    Ju[8 * NU + 2] = (-2.0 / RC) * imz;
    Ju[11 * NU + 1] = ((0.167 + 0.167 - 0.175 - 0.175) / RC) * imn;
    Ju[11 * NU + 3] = ((0.167 + 0.167 + 0.175 + 0.175) / RC) * imn;
}

void orc_vde_forw_cm(const double *x, const double *Sx, const double *Su, const double *u, const double *p,
                     double *f, double *dSx, double *dSu)
{
    double Jx[NX * NX], Ju[NX * NU];
    orc_ode(x, u, p, f);
    orc_jac(x, u, p, Jx, Ju);
    for (int j = 0; j < NX; j++)
        for (int i = 0; i < NX; i++) {
            double s = 0;
            for (int k = 0; k < NX; k++) s += Jx[i * NX + k] * Sx[k + NX * j];
            dSx[i + NX * j] = s;
        }
    for (int j = 0; j < NU; j++)
        for (int i = 0; i < NX; i++) {
            double s = Ju[i * NU + j];
            for (int k = 0; k < NX; k++) s += Jx[i * NX + k] * Su[k + NX * j];
            dSu[i + NX * j] = s;
        }
}

typedef int (*casadi_fn_t)(const double **arg, double **res, int *iw, double *w, int mem);
static casadi_fn_t g_casadi_vde = 0;
void orc_set_casadi_vde(void *fn) { g_casadi_vde = (casadi_fn_t)fn; }

static void vde_eval(const double *x, const double *Sx, const double *Su, const double *u, const double *p,
                     double *f, double *dSx, double *dSu)
{
    if (g_casadi_vde) {
        /* argument order of bluerov2_expl_vde_forw: (x, Sx, Su, u, p) -> (f, dSx, dSu) */
        const double *arg[5] = {x, Sx, Su, u, p};
        double *res[3] = {f, dSx, dSu};
        g_casadi_vde(arg, res, 0, 0, 0);
    } else {
        orc_vde_forw_cm(x, Sx, Su, u, p, f, dSx, dSu);
    }
}

/* acados ERK with forward sensitivities [upstream]: classical tableau c=(0,1/2,1/2,1), b=(1/6,1/3,1/3,1/6),
 * one step of length h on z=[x, Sx, Su], z0=[x, I, 0]; 4 stages / 1 step per acados_solver_bluerov2.c:633,639. */
void orc_erk4_sens(const double *x, const double *u, const double *p, double h, double *xn, double *A, double *B)
{
    enum { NZ = NX + NX * NX + NX * NU };
    double z0[NZ], zs[NZ], k[4][NZ];
    memset(z0, 0, sizeof z0);
    memcpy(z0, x, sizeof(double) * NX);
    for (int i = 0; i < NX; i++) z0[NX + i + NX * i] = 1.0;
    static const double c[4] = {0.0, 0.5, 0.5, 1.0};
    static const double bw[4] = {1.0 / 6, 1.0 / 3, 1.0 / 3, 1.0 / 6};
    for (int s = 0; s < 4; s++) {
        for (int i = 0; i < NZ; i++) zs[i] = z0[i] + (s ? c[s] * h * k[s - 1][i] : 0.0);
        vde_eval(zs, zs + NX, zs + NX + NX * NX, u, p, k[s], k[s] + NX, k[s] + NX + NX * NX);
    }
    for (int i = 0; i < NZ; i++) {
        double acc = 0;
        for (int s = 0; s < 4; s++) acc += bw[s] * k[s][i];
        zs[i] = z0[i] + h * acc;
    }
    memcpy(xn, zs, sizeof(double) * NX);
    for (int i = 0; i < NX; i++)
        for (int j = 0; j < NX; j++) A[i * NX + j] = zs[NX + i + NX * j];
    for (int i = 0; i < NX; i++)
        for (int j = 0; j < NU; j++) B[i * NU + j] = zs[NX + NX * NX + i + NX * j];
}

void orc_erk4(const double *x, const double *u, const double *p, double h, double *xn)
{
    double k1[NX], k2[NX], k3[NX], k4[NX], xs[NX];
    orc_ode(x, u, p, k1);
    for (int i = 0; i < NX; i++) xs[i] = x[i] + 0.5 * h * k1[i];
    orc_ode(xs, u, p, k2);
    for (int i = 0; i < NX; i++) xs[i] = x[i] + 0.5 * h * k2[i];
    orc_ode(xs, u, p, k3);
    for (int i = 0; i < NX; i++) xs[i] = x[i] + h * k3[i];
    orc_ode(xs, u, p, k4);
    for (int i = 0; i < NX; i++) xn[i] = x[i] + h * (k1[i] / 6 + k2[i] / 3 + k3[i] / 3 + k4[i] / 6);
}

void orc_linearize(int N, const double *Ts, const double *p, int p_stride, const double *X, const double *U,
                   double *A, double *B, double *b)
{
    for (int k = 0; k < N; k++) {
        double xn[NX];
        orc_erk4_sens(X + k * NX, U + k * NU, p + (size_t)k * p_stride, Ts[k], xn, A + k * NX * NX, B + k * NX * NU);
        for (int i = 0; i < NX; i++) b[k * NX + i] = xn[i] - X[(k + 1) * NX + i];
    }
}

/* ------------------------------------------------------------------------------------------------
 * OCP-QP:   min  sum_k 1/2 dx'Q dx + q'dx + 1/2 du'R du + r'du   (+ terminal)
 *           s.t. dx_{k+1} = A_k dx_k + B_k du_k + b_k,  dx_0 given,  lb_k <= du_k <= ub_k
 * Strictly convex (R > 0), so the solution is unique: any exact solver, HPIPM on the condensed form
 * included, returns it.  Solved here by a Mehrotra predictor-corrector IPM in "absolute" form: each
 * Newton system is the LQR with R~ = R + lam_l/t_l + lam_u/t_u and a shifted input gradient, so x and
 * the costates never have to be carried as iterates -- only (du, t, lam).
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    int N;
    double *P;     /* (N+1) x 144 cost-to-go Hessians */
    double *Kfb;   /* N x 48   feedback gains  K = Lam^-1 B'PA */
    double *S;     /* N x 48   B'PA */
    double *Lc;    /* N x 16   Cholesky factor of Lam = R~ + B'PB (lower, row-major) */
    double *Pb;    /* N x 12   P_{k+1} b_k */
    double *pv;    /* (N+1) x 12 cost-to-go gradients */
    double *kff;   /* N x 4 */
} ricc_ws;

static int chol4(const double *M, double *L)
{
    memset(L, 0, sizeof(double) * 16);
    for (int i = 0; i < 4; i++)
        for (int j = 0; j <= i; j++) {
            double s = M[i * 4 + j];
            for (int k = 0; k < j; k++) s -= L[i * 4 + k] * L[j * 4 + k];
            if (i == j) {
                if (!(s > 0)) return 1;
                L[i * 4 + i] = sqrt(s);
            } else
                L[i * 4 + j] = s / L[j * 4 + j];
        }
    return 0;
}
static void chol4_solve(const double *L, double *v) /* in place: v <- (L L')^-1 v */
{
    for (int i = 0; i < 4; i++) {
        double s = v[i];
        for (int k = 0; k < i; k++) s -= L[i * 4 + k] * v[k];
        v[i] = s / L[i * 4 + i];
    }
    for (int i = 3; i >= 0; i--) {
        double s = v[i];
        for (int k = i + 1; k < 4; k++) s -= L[k * 4 + i] * v[k];
        v[i] = s / L[i * 4 + i];
    }
}

/* backward factorisation with the current barrier diagonal Rt (N x 4) */
static int ricc_factor(const orc_qp *qp, const double *Rt, ricc_ws *ws)
{
    const int N = qp->N;
    double *P = ws->P + (size_t)N * 144;
    memset(P, 0, sizeof(double) * 144);
    for (int i = 0; i < NX; i++) P[i * NX + i] = qp->Qd[N * NX + i];
    for (int k = N - 1; k >= 0; k--) {
        const double *A = qp->A + (size_t)k * 144, *B = qp->B + (size_t)k * 48, *b = qp->b + (size_t)k * 12;
        const double *Pn = ws->P + (size_t)(k + 1) * 144;
        double *Pk = ws->P + (size_t)k * 144, *S = ws->S + (size_t)k * 48, *K = ws->Kfb + (size_t)k * 48;
        double PA[144], PB[48], Lam[16];
        for (int i = 0; i < NX; i++) {
            for (int j = 0; j < NX; j++) {
                double s = 0;
                for (int l = 0; l < NX; l++) s += Pn[i * NX + l] * A[l * NX + j];
                PA[i * NX + j] = s;
            }
            for (int j = 0; j < NU; j++) {
                double s = 0;
                for (int l = 0; l < NX; l++) s += Pn[i * NX + l] * B[l * NU + j];
                PB[i * NU + j] = s;
            }
            double s = 0;
            for (int l = 0; l < NX; l++) s += Pn[i * NX + l] * b[l];
            ws->Pb[k * NX + i] = s;
        }
        for (int a = 0; a < NU; a++) {
            for (int c = 0; c < NU; c++) {
                double s = (a == c) ? Rt[k * NU + a] : 0.0;
                for (int l = 0; l < NX; l++) s += B[l * NU + a] * PB[l * NU + c];
                Lam[a * 4 + c] = s;
            }
            for (int j = 0; j < NX; j++) {
                double s = 0;
                for (int l = 0; l < NX; l++) s += B[l * NU + a] * PA[l * NX + j];
                S[a * NX + j] = s;
            }
        }
        for (int a = 0; a < NU; a++)
            for (int c = 0; c < a; c++) Lam[a * 4 + c] = Lam[c * 4 + a] = 0.5 * (Lam[a * 4 + c] + Lam[c * 4 + a]);
        if (chol4(Lam, ws->Lc + (size_t)k * 16)) return 1;
        for (int j = 0; j < NX; j++) {
            double col[4] = {S[0 * NX + j], S[1 * NX + j], S[2 * NX + j], S[3 * NX + j]};
            chol4_solve(ws->Lc + (size_t)k * 16, col);
            for (int a = 0; a < NU; a++) K[a * NX + j] = col[a];
        }
        for (int i = 0; i < NX; i++)
            for (int j = 0; j < NX; j++) {
                double s = (i == j) ? qp->Qd[k * NX + i] : 0.0;
                for (int l = 0; l < NX; l++) s += A[l * NX + i] * PA[l * NX + j];
                for (int a = 0; a < NU; a++) s -= S[a * NX + i] * K[a * NX + j];
                Pk[i * NX + j] = s;
            }
        for (int i = 0; i < NX; i++)
            for (int j = 0; j < i; j++) Pk[i * NX + j] = Pk[j * NX + i] = 0.5 * (Pk[i * NX + j] + Pk[j * NX + i]);
    }
    return 0;
}

/* Newton step in residual (delta) form.  The iterate keeps dx = roll-out(du) and pi = adjoint(dx), so the only
 * non-zero KKT residual is the input-gradient one and the step solves the homogeneous LQR
 *     min sum 1/2 ddx'Q ddx + 1/2 ddu'R~ ddu + gh'ddu,   ddx_{k+1} = A ddx_k + B ddu_k,  ddx_0 = 0
 * whose vector recursion needs no P (b = 0, q = 0):  kff = Lam^-1 (gh + B'p+),  p = A'p+ - S'kff.
 * Working with steps instead of new points keeps tiny steps on active bounds free of cancellation. */
static void ricc_solve(const orc_qp *qp, const double *gh, ricc_ws *ws, double *ddx, double *ddu)
{
    const int N = qp->N;
    memset(ws->pv + (size_t)N * NX, 0, sizeof(double) * NX);
    for (int k = N - 1; k >= 0; k--) {
        const double *A = qp->A + (size_t)k * 144, *B = qp->B + (size_t)k * 48;
        const double *S = ws->S + (size_t)k * 48;
        const double *w = ws->pv + (size_t)(k + 1) * NX;
        double g[NU];
        for (int a = 0; a < NU; a++) {
            double s = gh[k * NU + a];
            for (int l = 0; l < NX; l++) s += B[l * NU + a] * w[l];
            g[a] = s;
        }
        chol4_solve(ws->Lc + (size_t)k * 16, g);
        memcpy(ws->kff + k * NU, g, sizeof g);
        for (int i = 0; i < NX; i++) {
            double s = 0;
            for (int l = 0; l < NX; l++) s += A[l * NX + i] * w[l];
            for (int a = 0; a < NU; a++) s -= S[a * NX + i] * g[a];
            ws->pv[k * NX + i] = s;
        }
    }
    memset(ddx, 0, sizeof(double) * NX);
    for (int k = 0; k < N; k++) {
        const double *A = qp->A + (size_t)k * 144, *B = qp->B + (size_t)k * 48;
        const double *K = ws->Kfb + (size_t)k * 48;
        const double *xk = ddx + k * NX;
        double *uk = ddu + k * NU, *xn = ddx + (k + 1) * NX;
        for (int a = 0; a < NU; a++) {
            double s = -ws->kff[k * NU + a];
            for (int j = 0; j < NX; j++) s -= K[a * NX + j] * xk[j];
            uk[a] = s;
        }
        for (int i = 0; i < NX; i++) {
            double s = 0;
            for (int j = 0; j < NX; j++) s += A[i * NX + j] * xk[j];
            for (int a = 0; a < NU; a++) s += B[i * NU + a] * uk[a];
            xn[i] = s;
        }
    }
}

static double max_step(int n, const double *v, const double *dv)
{
    double a = 1.0;
    for (int i = 0; i < n; i++)
        if (dv[i] < 0) {
            double c = -v[i] / dv[i];
            if (c < a) a = c;
        }
    return a;
}

void orc_qp_kkt(const orc_qp *qp, const double *dx, const double *du, const double *pi, const double *lam_l,
                const double *lam_u, double *res)
{
    const int N = qp->N;
    double rs = 0, re = 0, ri = 0, rc = 0, lmin = 1e300;
    for (int i = 0; i < NX; i++) re = fmax(re, fabs(dx[i] - qp->dx0[i]));
    for (int k = 0; k < N; k++) {
        const double *A = qp->A + (size_t)k * 144, *B = qp->B + (size_t)k * 48, *b = qp->b + (size_t)k * 12;
        /* Lagrangian: ... + pi_{k+1}'(A dx_k + B du_k + b - dx_{k+1}) ; pi_0 multiplies the initial condition.
         * d/d dx_k (k>=1): Q dx_k + q_k + A'pi_{k+1} - pi_k = 0 ;  d/d du_k: R du + r + B'pi_{k+1} - lam_l + lam_u = 0 */
        for (int i = 0; i < NX; i++) {
            double s = qp->Qd[k * NX + i] * dx[k * NX + i] + qp->q[k * NX + i] - pi[k * NX + i];
            for (int l = 0; l < NX; l++) s += A[l * NX + i] * pi[(k + 1) * NX + l];
            if (k > 0) rs = fmax(rs, fabs(s)); /* stage 0 state is fixed: pi_0 is free */
        }
        for (int a = 0; a < NU; a++) {
            double s = qp->Rd[k * NU + a] * du[k * NU + a] + qp->r[k * NU + a] - lam_l[k * NU + a] + lam_u[k * NU + a];
            for (int l = 0; l < NX; l++) s += B[l * NU + a] * pi[(k + 1) * NX + l];
            rs = fmax(rs, fabs(s));
            double tl = du[k * NU + a] - qp->lb[k * NU + a], tu = qp->ub[k * NU + a] - du[k * NU + a];
            ri = fmax(ri, fmax(-tl, -tu));
            rc = fmax(rc, fmax(fabs(tl * lam_l[k * NU + a]), fabs(tu * lam_u[k * NU + a])));
            lmin = fmin(lmin, fmin(lam_l[k * NU + a], lam_u[k * NU + a]));
        }
        for (int i = 0; i < NX; i++) {
            double s = b[i] - dx[(k + 1) * NX + i];
            for (int j = 0; j < NX; j++) s += A[i * NX + j] * dx[k * NX + j];
            for (int a = 0; a < NU; a++) s += B[i * NU + a] * du[k * NU + a];
            re = fmax(re, fabs(s));
        }
    }
    for (int i = 0; i < NX; i++) {
        double s = qp->Qd[N * NX + i] * dx[N * NX + i] + qp->q[N * NX + i] - pi[N * NX + i];
        rs = fmax(rs, fabs(s));
    }
    res[0] = rs; res[1] = re; res[2] = ri; res[3] = rc; res[4] = lmin;
}

int orc_qp_solve(const orc_qp *qp, int max_iter, double tol, double *dx, double *du, double *pi, double *lam_l,
                 double *lam_u, orc_qp_stats *st)
{
    const int N = qp->N, nb = N * NU;
    ricc_ws ws;
    ws.N = N;
    double *mem = (double *)malloc(sizeof(double) * ((size_t)(N + 1) * 144 + (size_t)N * (48 + 48 + 16 + 12 + 4) + (size_t)(N + 1) * 12 + 17 * (size_t)nb + (size_t)(N + 1) * 12));
    double *m = mem;
    ws.P = m;   m += (size_t)(N + 1) * 144;
    ws.Kfb = m; m += (size_t)N * 48;
    ws.S = m;   m += (size_t)N * 48;
    ws.Lc = m;  m += (size_t)N * 16;
    ws.Pb = m;  m += (size_t)N * 12;
    ws.kff = m; m += (size_t)N * 4;
    ws.pv = m;  m += (size_t)(N + 1) * 12;
    double *tl = m, *tu = tl + nb, *ll = tu + nb, *lu = ll + nb, *Rt = lu + nb, *rh = Rt + nb;
    double *dtl = rh + nb, *dtu = dtl + nb, *dll = dtu + nb, *dlu = dll + nb, *v = dlu + nb, *vt = v + nb;
    double *cl = vt + nb, *cu = cl + nb, *rl = cu + nb, *ru = rl + nb, *gu = ru + nb;
    double *xt = gu + nb; /* (N+1) x 12 trial states */
    int status = 2, it = 0;
    double mu = 0, res_stat = 0, res_ineq = 0, stat_scale = 1.0;

    /* cold start (qp_solver_warm_start 0): du = 0 pushed strictly inside the box, slacks EXACTLY consistent
     * (t_l = v - lb, t_u = ub - v), lam = mu0/t.  The iteration is then a feasible-start method: every step keeps
     * dt_l = dv, dt_u = -dv, so the bound residuals stay zero in exact arithmetic and are dropped from the
     * formulas -- recomputing v - lb - t_l by subtraction of O(50) numbers costs 1e-14 absolute, which the
     * 1/t factors amplify without bound once t ~ 1e-13 (active bounds). */
    const double thr = 1e-1, mu0 = 1.0;
    for (int i = 0; i < nb; i++) {
        v[i] = fmin(fmax(0.0, qp->lb[i] + thr), qp->ub[i] - thr);
        if (qp->ub[i] - qp->lb[i] < 2 * thr) v[i] = 0.5 * (qp->lb[i] + qp->ub[i]);
        tl[i] = v[i] - qp->lb[i];
        tu[i] = qp->ub[i] - v[i];
        ll[i] = mu0 / tl[i];
        lu[i] = mu0 / tu[i];
        rl[i] = ru[i] = 0.0;
    }
    for (it = 0; it < max_iter; it++) {
        mu = 0;
        res_ineq = 0;
        for (int i = 0; i < nb; i++) mu += ll[i] * tl[i] + lu[i] * tu[i];
        mu /= (2.0 * nb);
        /* iterate-consistent states and costates: dx = roll-out(v), pi = adjoint(dx).  The reduced input gradient
         * gu = R v + r + B'pi+ is then the only residual; stationarity residual = gu - lam_l + lam_u. */
        {
            memcpy(xt, qp->dx0, sizeof(double) * NX);
            for (int k = 0; k < N; k++) {
                const double *A = qp->A + (size_t)k * 144, *B = qp->B + (size_t)k * 48, *b = qp->b + (size_t)k * 12;
                for (int i = 0; i < NX; i++) {
                    double s = b[i];
                    for (int j = 0; j < NX; j++) s += A[i * NX + j] * xt[k * NX + j];
                    for (int a = 0; a < NU; a++) s += B[i * NU + a] * v[k * NU + a];
                    xt[(k + 1) * NX + i] = s;
                }
            }
            for (int i = 0; i < NX; i++) pi[N * NX + i] = qp->Qd[N * NX + i] * xt[N * NX + i] + qp->q[N * NX + i];
            res_stat = 0;
            for (int k = N - 1; k >= 0; k--) {
                const double *A = qp->A + (size_t)k * 144, *B = qp->B + (size_t)k * 48;
                for (int a = 0; a < NU; a++) {
                    double s = qp->Rd[k * NU + a] * v[k * NU + a] + qp->r[k * NU + a];
                    for (int l = 0; l < NX; l++) s += B[l * NU + a] * pi[(k + 1) * NX + l];
                    gu[k * NU + a] = s;
                    res_stat = fmax(res_stat, fabs(s - ll[k * NU + a] + lu[k * NU + a]));
                }
                for (int i = 0; i < NX; i++) {
                    double s = qp->Qd[k * NX + i] * xt[k * NX + i] + qp->q[k * NX + i];
                    for (int l = 0; l < NX; l++) s += A[l * NX + i] * pi[(k + 1) * NX + l];
                    pi[k * NX + i] = s;
                }
            }
        }
        if (getenv("ORC_DEBUG")) fprintf(stderr, "it %d mu %.3e stat %.3e\n", it, mu, res_stat);
        if (it == 0) stat_scale = fmax(1.0, res_stat);
        if (mu < tol && res_stat < tol * stat_scale) { status = 0; break; }

        /* predictor (sigma = 0): R~ ddu + B'dpi+ = -gu + c_l/t_l - c_u/t_u with c = sigma*mu - corr = 0 */
        for (int i = 0; i < nb; i++) {
            Rt[i] = qp->Rd[i] + ll[i] / tl[i] + lu[i] / tu[i];
            rh[i] = gu[i];
        }
        if (ricc_factor(qp, Rt, &ws)) { status = 4; break; }
        ricc_solve(qp, rh, &ws, xt, vt);
        for (int i = 0; i < nb; i++) {
            dtl[i] = vt[i];
            dtu[i] = -vt[i];
            dll[i] = -ll[i] - ll[i] * dtl[i] / tl[i];
            dlu[i] = -lu[i] - lu[i] * dtu[i] / tu[i];
        }
        double a_aff = fmin(fmin(max_step(nb, tl, dtl), max_step(nb, tu, dtu)), fmin(max_step(nb, ll, dll), max_step(nb, lu, dlu)));
        double mu_aff = 0;
        for (int i = 0; i < nb; i++)
            mu_aff += (ll[i] + a_aff * dll[i]) * (tl[i] + a_aff * dtl[i]) + (lu[i] + a_aff * dlu[i]) * (tu[i] + a_aff * dtu[i]);
        mu_aff /= (2.0 * nb);
        double sigma = mu_aff / mu;
        sigma = sigma * sigma * sigma;

        /* corrector: centring + Mehrotra second-order term */
        for (int i = 0; i < nb; i++) {
            cl[i] = sigma * mu - dtl[i] * dll[i];
            cu[i] = sigma * mu - dtu[i] * dlu[i];
            rh[i] = gu[i] - cl[i] / tl[i] + cu[i] / tu[i];
        }
        ricc_solve(qp, rh, &ws, xt, vt);
        for (int i = 0; i < nb; i++) {
            dtl[i] = vt[i];
            dtu[i] = -vt[i];
            dll[i] = cl[i] / tl[i] - ll[i] - ll[i] * dtl[i] / tl[i];
            dlu[i] = cu[i] / tu[i] - lu[i] - lu[i] * dtu[i] / tu[i];
        }
        double ap = fmin(max_step(nb, tl, dtl), max_step(nb, tu, dtu));
        double ad = fmin(max_step(nb, ll, dll), max_step(nb, lu, dlu));
        /* separate primal/dual step lengths (legal here: x and pi are recomputed from v every iteration, so no
         * residual bookkeeping depends on a common step); fraction-to-boundary tau = max(0.995, 1 - mu) gives
         * superlinear terminal convergence, capped below 1 so no slack ever reaches zero exactly. */
        const double tau = fmin(fmax(0.995, 1.0 - mu), 1.0 - 1e-8);
        ap = fmin(1.0, tau * ap);
        ad = fmin(1.0, tau * ad);
        if (getenv("ORC_DEBUG")) fprintf(stderr, "   a_aff %.4f sigma %.3e ap %.4f ad %.4f\n", a_aff, sigma, ap, ad);
        for (int i = 0; i < nb; i++) {
            v[i] += ap * dtl[i];
            tl[i] += ap * dtl[i];
            tu[i] += ap * dtu[i];
            ll[i] += ad * dll[i];
            lu[i] += ad * dlu[i];
        }
    }
    /* final primal: exact roll-out of the converged inputs; costates from the adjoint recursion */
    memcpy(du, v, sizeof(double) * nb);
    memcpy(dx, qp->dx0, sizeof(double) * NX);
    for (int k = 0; k < N; k++) {
        const double *A = qp->A + (size_t)k * 144, *B = qp->B + (size_t)k * 48, *b = qp->b + (size_t)k * 12;
        for (int i = 0; i < NX; i++) {
            double s = b[i];
            for (int j = 0; j < NX; j++) s += A[i * NX + j] * dx[k * NX + j];
            for (int a = 0; a < NU; a++) s += B[i * NU + a] * du[k * NU + a];
            dx[(k + 1) * NX + i] = s;
        }
    }
    for (int i = 0; i < NX; i++) pi[N * NX + i] = qp->Qd[N * NX + i] * dx[N * NX + i] + qp->q[N * NX + i];
    for (int k = N - 1; k >= 0; k--) {
        const double *A = qp->A + (size_t)k * 144;
        for (int i = 0; i < NX; i++) {
            double s = qp->Qd[k * NX + i] * dx[k * NX + i] + qp->q[k * NX + i];
            for (int l = 0; l < NX; l++) s += A[l * NX + i] * pi[(k + 1) * NX + l];
            pi[k * NX + i] = s;
        }
    }
    memcpy(lam_l, ll, sizeof(double) * nb);
    memcpy(lam_u, lu, sizeof(double) * nb);
    if (st) {
        st->iters = it;
        st->status = status;
        st->mu = mu;
        st->res_stat = res_stat;
        st->res_ineq = res_ineq;
        st->res_comp = mu;
    }
    free(mem);
    return status;
}

/* ------------------------------------------------------------------------------------------------
 * The same QP by FULL CONDENSING + a dense interior-point iteration -- the cost profile of the reference's own QP path
 * (qp_solver FULL_CONDENSING_HPIPM, acados_solver_bluerov2.c:146,664-668: acados eliminates the states, HPIPM runs its dense
 * Mehrotra IPM on the N*nu = 160 remaining variables).  States eliminated through dx_k = c_k + sum_j Gam[k][j] du_j
 * (Gam[j+1][j] = B_j, Gam[k+1][j] = A_k Gam[k][j]; c_0 = dx0, c_{k+1} = A_k c_k + b_k), condensed Hessian
 * H = sum_k Gam_k' Q_k Gam_k + R, gradient g = sum_k Gam_k' (Q_k c_k + q_k) + r; the iteration is the one of orc_qp_solve
 * (same cold start, same predictor-corrector, same step rule) with every LQR solve replaced by a dense Cholesky solve of
 * H + diag(lam_l / t_l + lam_u / t_u).  Plain C loops: a BLASFEO-class implementation would run the same flops several times
 * faster, so this arm is reported beside the Riccati arm, never instead of it.  Selected by orc_set_qp_mode(1).
 * ---------------------------------------------------------------------------------------------- */
static int g_qp_mode = 0;
void orc_set_qp_mode(int mode) { g_qp_mode = mode; }
int orc_get_qp_mode(void) { return g_qp_mode; }

static int chol_dense(int n, double *M)            /* in place, lower triangle; returns non-zero on a non-positive pivot */
{
    for (int j = 0; j < n; j++) {
        double d = M[(size_t)j * n + j];
        for (int k = 0; k < j; k++) d -= M[(size_t)j * n + k] * M[(size_t)j * n + k];
        if (!(d > 0.0)) return 1;
        d = sqrt(d);
        M[(size_t)j * n + j] = d;
        for (int i = j + 1; i < n; i++) {
            double sacc = M[(size_t)i * n + j];
            for (int k = 0; k < j; k++) sacc -= M[(size_t)i * n + k] * M[(size_t)j * n + k];
            M[(size_t)i * n + j] = sacc / d;
        }
    }
    return 0;
}
static void chol_dense_solve(int n, const double *L, double *v)   /* v <- (L L')^-1 v */
{
    for (int i = 0; i < n; i++) {
        double sacc = v[i];
        for (int k = 0; k < i; k++) sacc -= L[(size_t)i * n + k] * v[k];
        v[i] = sacc / L[(size_t)i * n + i];
    }
    for (int i = n - 1; i >= 0; i--) {
        double sacc = v[i];
        for (int k = i + 1; k < n; k++) sacc -= L[(size_t)k * n + i] * v[k];
        v[i] = sacc / L[(size_t)i * n + i];
    }
}

int orc_qp_solve_dense(const orc_qp *qp, int max_iter, double tol, double *dx, double *du, double *pi, double *lam_l,
                       double *lam_u, orc_qp_stats *st)
{
    const int N = qp->N, n = N * NU;
    const size_t nn = (size_t)n * n;
    /* Gam: for every stage k = 1..N the 12 x (4k) block row, stored with row length n */
    double *mem = (double *)malloc(sizeof(double) * ((size_t)(N + 1) * NX * n + (size_t)(N + 1) * NX + 2 * nn + 19 * (size_t)n));
    if (!mem) return 4;
    double *Gam = mem, *c = Gam + (size_t)(N + 1) * NX * n, *H = c + (size_t)(N + 1) * NX, *Mw = H + nn, *g = Mw + nn;
    double *tl = g + n, *tu = tl + n, *ll = tu + n, *lu = ll + n, *rh = lu + n, *dtl = rh + n, *dtu = dtl + n, *dll = dtu + n;
    double *dlu = dll + n, *v = dlu + n, *vt = v + n, *cl = vt + n, *cu = cl + n, *gu = cu + n, *w1 = gu + n, *w2 = w1 + n;
    (void)w1; (void)w2;
    memset(Gam, 0, sizeof(double) * (size_t)(N + 1) * NX * n);
    memcpy(c, qp->dx0, sizeof(double) * NX);
    for (int k = 0; k < N; k++) {
        const double *A = qp->A + (size_t)k * 144, *B = qp->B + (size_t)k * 48, *b = qp->b + (size_t)k * 12;
        double *Gk = Gam + (size_t)k * NX * n, *Gn = Gam + (size_t)(k + 1) * NX * n;
        for (int i = 0; i < NX; i++) {
            double sacc = b[i];
            for (int j = 0; j < NX; j++) sacc += A[i * NX + j] * c[k * NX + j];
            c[(k + 1) * NX + i] = sacc;
            for (int col = 0; col < k * NU; col++) {
                double t = 0.0;
                for (int j = 0; j < NX; j++) t += A[i * NX + j] * Gk[(size_t)j * n + col];
                Gn[(size_t)i * n + col] = t;
            }
            for (int a = 0; a < NU; a++) Gn[(size_t)i * n + k * NU + a] = B[i * NU + a];
        }
    }
    memset(H, 0, sizeof(double) * nn);
    memset(g, 0, sizeof(double) * n);
    for (int k = 1; k <= N; k++) {
        const double *Gk = Gam + (size_t)k * NX * n;
        const int nc = k * NU;
        for (int i = 0; i < NX; i++) {
            const double qd = qp->Qd[k * NX + i], ge = qd * c[k * NX + i] + qp->q[k * NX + i];
            const double *row = Gk + (size_t)i * n;
            for (int a = 0; a < nc; a++) {
                const double ra = row[a];
                if (ra == 0.0) continue;
                g[a] += ra * ge;
                const double qa = qd * ra;
                for (int bb = 0; bb <= a; bb++) H[(size_t)a * n + bb] += qa * row[bb];
            }
        }
    }
    for (int a = 0; a < n; a++) {
        H[(size_t)a * n + a] += qp->Rd[a];
        g[a] += qp->r[a];
        for (int bb = 0; bb < a; bb++) H[(size_t)bb * n + a] = H[(size_t)a * n + bb];
    }
    int status = 2, it = 0;
    double mu = 0, res_stat = 0, stat_scale = 1.0;
    const double thr = 1e-1, mu0 = 1.0;
    for (int i = 0; i < n; i++) {
        v[i] = fmin(fmax(0.0, qp->lb[i] + thr), qp->ub[i] - thr);
        if (qp->ub[i] - qp->lb[i] < 2 * thr) v[i] = 0.5 * (qp->lb[i] + qp->ub[i]);
        tl[i] = v[i] - qp->lb[i];
        tu[i] = qp->ub[i] - v[i];
        ll[i] = mu0 / tl[i];
        lu[i] = mu0 / tu[i];
    }
    for (it = 0; it < max_iter; it++) {
        mu = 0;
        for (int i = 0; i < n; i++) mu += ll[i] * tl[i] + lu[i] * tu[i];
        mu /= (2.0 * n);
        res_stat = 0;
        for (int a = 0; a < n; a++) {
            double sacc = g[a];
            for (int bb = 0; bb < n; bb++) sacc += H[(size_t)a * n + bb] * v[bb];
            gu[a] = sacc;
            res_stat = fmax(res_stat, fabs(sacc - ll[a] + lu[a]));
        }
        if (it == 0) stat_scale = fmax(1.0, res_stat);
        if (mu < tol && res_stat < tol * stat_scale) { status = 0; break; }
        memcpy(Mw, H, sizeof(double) * nn);
        for (int i = 0; i < n; i++) Mw[(size_t)i * n + i] += ll[i] / tl[i] + lu[i] / tu[i];
        if (chol_dense(n, Mw)) { status = 4; break; }
        for (int i = 0; i < n; i++) vt[i] = -gu[i];
        chol_dense_solve(n, Mw, vt);
        for (int i = 0; i < n; i++) {
            dtl[i] = vt[i];
            dtu[i] = -vt[i];
            dll[i] = -ll[i] - ll[i] * dtl[i] / tl[i];
            dlu[i] = -lu[i] - lu[i] * dtu[i] / tu[i];
        }
        double a_aff = fmin(fmin(max_step(n, tl, dtl), max_step(n, tu, dtu)), fmin(max_step(n, ll, dll), max_step(n, lu, dlu)));
        double mu_aff = 0;
        for (int i = 0; i < n; i++)
            mu_aff += (ll[i] + a_aff * dll[i]) * (tl[i] + a_aff * dtl[i]) + (lu[i] + a_aff * dlu[i]) * (tu[i] + a_aff * dtu[i]);
        mu_aff /= (2.0 * n);
        double sigma = mu_aff / mu;
        sigma = sigma * sigma * sigma;
        for (int i = 0; i < n; i++) {
            cl[i] = sigma * mu - dtl[i] * dll[i];
            cu[i] = sigma * mu - dtu[i] * dlu[i];
            rh[i] = gu[i] - cl[i] / tl[i] + cu[i] / tu[i];
            vt[i] = -rh[i];
        }
        chol_dense_solve(n, Mw, vt);
        for (int i = 0; i < n; i++) {
            dtl[i] = vt[i];
            dtu[i] = -vt[i];
            dll[i] = cl[i] / tl[i] - ll[i] - ll[i] * dtl[i] / tl[i];
            dlu[i] = cu[i] / tu[i] - lu[i] - lu[i] * dtu[i] / tu[i];
        }
        double ap = fmin(max_step(n, tl, dtl), max_step(n, tu, dtu));
        double ad = fmin(max_step(n, ll, dll), max_step(n, lu, dlu));
        const double tau = fmin(fmax(0.995, 1.0 - mu), 1.0 - 1e-8);
        ap = fmin(1.0, tau * ap);
        ad = fmin(1.0, tau * ad);
        for (int i = 0; i < n; i++) {
            v[i] += ap * dtl[i];
            tl[i] += ap * dtl[i];
            tu[i] += ap * dtu[i];
            ll[i] += ad * dll[i];
            lu[i] += ad * dlu[i];
        }
    }
    memcpy(du, v, sizeof(double) * n);
    /* states by the exact roll-out of the inputs, costates by the adjoint recursion (as orc_qp_solve) */
    memcpy(dx, qp->dx0, sizeof(double) * NX);
    for (int k = 0; k < N; k++) {
        const double *A = qp->A + (size_t)k * 144, *B = qp->B + (size_t)k * 48, *b = qp->b + (size_t)k * 12;
        for (int i = 0; i < NX; i++) {
            double sacc = b[i];
            for (int j = 0; j < NX; j++) sacc += A[i * NX + j] * dx[k * NX + j];
            for (int a = 0; a < NU; a++) sacc += B[i * NU + a] * du[k * NU + a];
            dx[(k + 1) * NX + i] = sacc;
        }
    }
    for (int i = 0; i < NX; i++) pi[N * NX + i] = qp->Qd[N * NX + i] * dx[N * NX + i] + qp->q[N * NX + i];
    for (int k = N - 1; k >= 0; k--) {
        const double *A = qp->A + (size_t)k * 144;
        for (int i = 0; i < NX; i++) {
            double sacc = qp->Qd[k * NX + i] * dx[k * NX + i] + qp->q[k * NX + i];
            for (int l = 0; l < NX; l++) sacc += A[l * NX + i] * pi[(k + 1) * NX + l];
            pi[k * NX + i] = sacc;
        }
    }
    memcpy(lam_l, ll, sizeof(double) * n);
    memcpy(lam_u, lu, sizeof(double) * n);
    if (st) { st->iters = it; st->status = status; st->mu = mu; st->res_stat = res_stat; st->res_ineq = 0; st->res_comp = mu; }
    free(mem);
    return status;
}

/* ------------------------------------------------------------------------------------------------
 * One SQP-RTI step (preparation + feedback, rti_phase 0).  Follows acados_solver_bluerov2.c:
 *   cost NONLINEAR_LS with y=[x;u], scaling = Ts on stages 0..N-1, terminal unscaled (:389-394,:424-479);
 *   Gauss-Newton Hessian (cost Hessian structurally empty, bluerov2_cost_y_hess.c:60);
 *   bounds lbu <= u <= ubu on stages 0..N-1 (:547-571); stage-0 state fixed to x0 (lbx=ubx, idxbxe :501-543);
 *   full step (step_length 1, :653); no shifting of (X,U) between calls.
 * ---------------------------------------------------------------------------------------------- */
int orc_rti_step(int N, const double *Ts, const double *W, const double *We, const double *lbu, const double *ubu,
                 const double *x0, const double *yref, const double *p, int p_stride, double *X, double *U,
                 int max_iter, double tol, double *info)
{
    size_t nd = (size_t)N * (144 + 48 + 12) + (size_t)(N + 1) * 12 * 4 + (size_t)N * 4 * 7;
    double *mem = (double *)malloc(sizeof(double) * nd), *m = mem;
    double *A = m;  m += (size_t)N * 144;
    double *B = m;  m += (size_t)N * 48;
    double *b = m;  m += (size_t)N * 12;
    double *Qd = m; m += (size_t)(N + 1) * 12;
    double *q = m;  m += (size_t)(N + 1) * 12;
    double *dx = m; m += (size_t)(N + 1) * 12;
    double *pi = m; m += (size_t)(N + 1) * 12;
    double *Rd = m; m += (size_t)N * 4;
    double *r = m;  m += (size_t)N * 4;
    double *lb = m; m += (size_t)N * 4;
    double *ub = m; m += (size_t)N * 4;
    double *du = m; m += (size_t)N * 4;
    double *ll = m; m += (size_t)N * 4;
    double *lu = m; m += (size_t)N * 4;
    double dx0[NX];

    orc_linearize(N, Ts, p, p_stride, X, U, A, B, b);
    double bmax = 0;
    for (int i = 0; i < N * NX; i++) bmax = fmax(bmax, fabs(b[i]));
    for (int k = 0; k < N; k++) {
        for (int i = 0; i < NX; i++) {
            Qd[k * NX + i] = Ts[k] * W[i];
            q[k * NX + i] = Ts[k] * W[i] * (X[k * NX + i] - yref[k * ORC_NY + i]);
        }
        for (int a = 0; a < NU; a++) {
            Rd[k * NU + a] = Ts[k] * W[NX + a];
            r[k * NU + a] = Ts[k] * W[NX + a] * (U[k * NU + a] - yref[k * ORC_NY + NX + a]);
            lb[k * NU + a] = lbu[a] - U[k * NU + a];
            ub[k * NU + a] = ubu[a] - U[k * NU + a];
        }
    }
    for (int i = 0; i < NX; i++) {
        Qd[N * NX + i] = We[i];
        q[N * NX + i] = We[i] * (X[N * NX + i] - yref[N * ORC_NY + i]);
        dx0[i] = x0[i] - X[i];
    }
    orc_qp qp = {N, A, B, b, Qd, Rd, q, r, lb, ub, dx0};
    orc_qp_stats st;
    int status = g_qp_mode == 1 ? orc_qp_solve_dense(&qp, max_iter, tol, dx, du, pi, ll, lu, &st)
                                : orc_qp_solve(&qp, max_iter, tol, dx, du, pi, ll, lu, &st);
    int nan = 0;
    for (int i = 0; i < N * NU; i++) nan |= !(du[i] == du[i]);
    if (!nan) {
        for (int i = 0; i < (N + 1) * NX; i++) X[i] += dx[i];
        for (int i = 0; i < N * NU; i++) U[i] += du[i];
    } else
        status = 1;
    if (info) {
        info[0] = st.iters; info[1] = st.status; info[2] = st.mu; info[3] = st.res_stat;
        info[4] = st.res_ineq; info[5] = st.res_comp; info[6] = bmax; info[7] = 0;
    }
    free(mem);
    return status;
}

int orc_rti_step_batch(int nb, int N, const double *Ts, const double *W, const double *We, const double *lbu,
                       const double *ubu, const double *x0, const double *yref, const double *p, double *X, double *U,
                       int max_iter, double tol, int *status, double *info, int nthreads)
{
    int used = 1;
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
    used = nthreads;
#pragma omp parallel for schedule(dynamic, 4) num_threads(nthreads)
#endif
    for (int i = 0; i < nb; i++) {
        int s = orc_rti_step(N, Ts, W, We, lbu, ubu, x0 + (size_t)i * NX, yref + (size_t)i * (N + 1) * ORC_NY,
                             p + (size_t)i * NP, 0, X + (size_t)i * (N + 1) * NX, U + (size_t)i * N * NU, max_iter, tol,
                             info ? info + (size_t)i * 8 : 0);
        if (status) status[i] = s;
    }
    return used;
}

void orc_thrust_alloc(const double *u, double *t)
{
    /* bluerov2_dob.cpp:390-395 */
    t[0] = (-u[0] + u[1] + u[3]) / RC;
    t[1] = (-u[0] - u[1] - u[3]) / RC;
    t[2] = (u[0] + u[1] - u[3]) / RC;
    t[3] = (u[0] - u[1] + u[3]) / RC;
    t[4] = (-u[2]) / RC;
    t[5] = (-u[2]) / RC;
}

/* ------------------------------------------------------------------------------------------------
 * EKF disturbance observer -- restates BLUEROV2_DOB::EKF / RK4 / f / h / compute_jacobian_F/H
 * (bluerov2_dob.cpp:495-545, 621-752) with the constants of bluerov2_dob.h:170-208 and the constructor
 * (bluerov2_dob.cpp:41-65).  Quirks kept: k3 = f(x + k2/3) (:630); invM(i,i) = diagonal of the inverse of the
 * COUPLED 6x6 mass matrix (:41-47); forward differences with d = 1e-6 (:726,:742); sin(psi) in the roll
 * kinematics (:646); h uses M(i,i)*body_acc (:702-707).
 * ---------------------------------------------------------------------------------------------- */
#define EN 18
static const double E_DT = 0.05, E_MASS = 11.26, E_IX = 0.3, E_IY = 0.63, E_IZ = 0.58, E_ZG = 0.02, E_G = 9.81;
static const double E_BUOY = 0.661618;
static const double E_AM[6] = {1.7182, 0, 5.468, 0, 1.2481, 0.4006};
/* damping of the filter's process/measurement model: model 0 = BLUEROV2_DOB (bluerov2_dob.h:182-183), model 1 =
 * BLUEROV2_AMPC, whose f/h (bluerov2_ampc.cpp:658-696) are the same expressions with Dl = 0 (:41-42) and no
 * quadratic damping.  Selected process-wide by orc_ekf_set_model (test infrastructure: not thread-safe). */
static const double E_DL_DOB[6] = {-11.7391, -20, -31.8678, -25, -44.9085, -5};
static const double E_DNL_DOB[6] = {-18.18, -21.66, -36.99, -1.55, -1.55, -1.55};
static double E_DL[6] = {-11.7391, -20, -31.8678, -25, -44.9085, -5};
static double E_DNL[6] = {-18.18, -21.66, -36.99, -1.55, -1.55, -1.55};
void orc_ekf_set_model(int model)
{
    for (int i = 0; i < 6; i++) {
        E_DL[i] = model == 1 ? 0.0 : E_DL_DOB[i];
        E_DNL[i] = model == 1 ? 0.0 : E_DNL_DOB[i];
    }
}
static const double E_K[6][6] = {
    {0.7071067811847433, 0.7071067811847433, -0.7071067811919605, -0.7071067811919605, 0.0, 0.0},
    {0.7071067811883519, -0.7071067811883519, 0.7071067811811348, -0.7071067811811348, 0.0, 0.0},
    {0, 0, 0, 0, 1, 1},
    {0.051265241636155506, -0.05126524163615552, 0.05126524163563227, -0.05126524163563227, -0.11050000000000001, 0.11050000000000003},
    {-0.05126524163589389, -0.051265241635893896, 0.05126524163641713, 0.05126524163641713, -0.002499999999974481, -0.002499999999974481},
    {0.16652364696949604, -0.16652364696949604, -0.17500892834341342, 0.17500892834341342, 0.0, 0.0}};

/* general inverse, Gaussian elimination with partial pivoting (stands in for Eigen's PartialPivLU inverse) */
static int mat_inverse(int n, const double *Ain, double *Ainv)
{
    double *a = (double *)malloc(sizeof(double) * n * n);
    memcpy(a, Ain, sizeof(double) * n * n);
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) Ainv[i * n + j] = (i == j);
    for (int c = 0; c < n; c++) {
        int piv = c;
        for (int i = c + 1; i < n; i++)
            if (fabs(a[i * n + c]) > fabs(a[piv * n + c])) piv = i;
        if (a[piv * n + c] == 0) { free(a); return 1; }
        if (piv != c)
            for (int j = 0; j < n; j++) {
                double t = a[c * n + j]; a[c * n + j] = a[piv * n + j]; a[piv * n + j] = t;
                t = Ainv[c * n + j]; Ainv[c * n + j] = Ainv[piv * n + j]; Ainv[piv * n + j] = t;
            }
        double d = 1.0 / a[c * n + c];
        for (int j = 0; j < n; j++) { a[c * n + j] *= d; Ainv[c * n + j] *= d; }
        for (int i = 0; i < n; i++)
            if (i != c) {
                double f = a[i * n + c];
                if (f != 0)
                    for (int j = 0; j < n; j++) { a[i * n + j] -= f * a[c * n + j]; Ainv[i * n + j] -= f * Ainv[c * n + j]; }
            }
    }
    free(a);
    return 0;
}

static void ekf_mass(double *Mdiag, double *invMdiag)
{
    double M[36], iM[36];
    memset(M, 0, sizeof M);
    M[0] = E_MASS + E_AM[0]; M[7] = E_MASS + E_AM[1]; M[14] = E_MASS + E_AM[2];
    M[21] = E_IX + E_AM[3];  M[28] = E_IY + E_AM[4];  M[35] = E_IZ + E_AM[5];
    M[0 * 6 + 4] = E_MASS * E_ZG;
    M[1 * 6 + 3] = -E_MASS * E_ZG;
    M[3 * 6 + 1] = -E_MASS * E_ZG;
    M[4 * 6 + 0] = E_MASS * E_ZG;
    mat_inverse(6, M, iM);
    for (int i = 0; i < 6; i++) { Mdiag[i] = M[i * 6 + i]; invMdiag[i] = iM[i * 6 + i]; }
}

void orc_ekf_init(double *esti_x, double *esti_P)
{
    static const double x0[EN] = {0, 0, -20, 0, 0, 0, 0, 0, 0, 0, 0, 0, 6, 6, 6, 0, 0, 0};
    memcpy(esti_x, x0, sizeof x0);
    for (int i = 0; i < EN; i++)
        for (int j = 0; j < EN; j++) esti_P[i * EN + j] = (i == j);
}

void orc_ekf_f(const double *x, const double *u, double *xdot)
{
    double Md[6], iM[6], KAu[6];
    ekf_mass(Md, iM);
    for (int i = 0; i < 6; i++) {
        double s = 0;
        for (int j = 0; j < 6; j++) s += E_K[i][j] * u[j];
        KAu[i] = s;
    }
    const double s3 = sin(x[3]), c3 = cos(x[3]), s4 = sin(x[4]), c4 = cos(x[4]), s5 = sin(x[5]), c5 = cos(x[5]);
    xdot[0] = (c5 * c4) * x[6] + (-s5 * c3 + c5 * s4 * s3) * x[7] + (s5 * s3 + c5 * c3 * s4) * x[8];
    xdot[1] = (s5 * c4) * x[6] + (c5 * c3 + s3 * s4 * s5) * x[7] + (-c5 * s3 + s4 * s5 * c3) * x[8];
    xdot[2] = (-s4) * x[6] + (c4 * s3) * x[7] + (c4 * c3) * x[8];
    xdot[3] = x[9] + (s5 * s4 / c4) * x[10] + c3 * s4 / c4 * x[11];
    xdot[4] = (c3) * x[10] + (s3) * x[11];
    xdot[5] = (s3 / c4) * x[10] + (c3 / c4) * x[11];
    xdot[6] = iM[0] * (KAu[0] + E_MASS * x[11] * x[7] - E_MASS * x[10] * x[8] - E_BUOY * s4 + x[12] + E_DL[0] * x[6] + E_DNL[0] * fabs(x[6]) * x[6]);
    xdot[7] = iM[1] * (KAu[1] - E_MASS * x[11] * x[6] + E_MASS * x[9] * x[8] + E_BUOY * c4 * s3 + x[13] + E_DL[1] * x[7] + E_DNL[1] * fabs(x[7]) * x[7]);
    xdot[8] = iM[2] * (KAu[2] + E_MASS * x[10] * x[6] - E_MASS * x[9] * x[7] + E_BUOY * c4 * c3 + x[14] + E_DL[2] * x[8] + E_DNL[2] * fabs(x[8]) * x[8]);
    xdot[9] = iM[3] * (KAu[3] + (E_IY - E_IZ) * x[10] * x[11] - E_MASS * E_ZG * E_G * c4 * s3 + x[15] + E_DL[3] * x[9] + E_DNL[3] * fabs(x[9]) * x[9]);
    xdot[10] = iM[4] * (KAu[4] + (E_IZ - E_IX) * x[9] * x[11] - E_MASS * E_ZG * E_G * s4 + x[16] + E_DL[4] * x[10] + E_DNL[4] * fabs(x[10]) * x[10]);
    xdot[11] = iM[5] * (KAu[5] - (E_IY - E_IX) * x[9] * x[10] + x[17] + E_DL[5] * x[11] + E_DNL[5] * fabs(x[11]) * x[11]);
    for (int i = 12; i < EN; i++) xdot[i] = 0;
}

void orc_ekf_h(const double *x, const double *acc, double *y)
{
    double Md[6], iM[6];
    ekf_mass(Md, iM);
    const double s3 = sin(x[3]), c3 = cos(x[3]), s4 = sin(x[4]), c4 = cos(x[4]);
    for (int i = 0; i < 12; i++) y[i] = x[i];
    y[12] = Md[0] * acc[0] - E_MASS * x[11] * x[7] + E_MASS * x[10] * x[8] + E_BUOY * s4 - x[12] - E_DL[0] * x[6] - E_DNL[0] * fabs(x[6]) * x[6];
    y[13] = Md[1] * acc[1] + E_MASS * x[11] * x[6] - E_MASS * x[9] * x[8] - E_BUOY * c4 * s3 - x[13] - E_DL[1] * x[7] - E_DNL[1] * fabs(x[7]) * x[7];
    y[14] = Md[2] * acc[2] - E_MASS * x[10] * x[6] + E_MASS * x[9] * x[7] - E_BUOY * c4 * c3 - x[14] - E_DL[2] * x[8] - E_DNL[2] * fabs(x[8]) * x[8];
    y[15] = Md[3] * acc[3] - (E_IY - E_IZ) * x[10] * x[11] + E_MASS * E_ZG * E_G * c4 * s3 - x[15] - E_DL[3] * x[9] - E_DNL[3] * fabs(x[9]) * x[9];
    y[16] = Md[4] * acc[4] - (E_IZ - E_IX) * x[9] * x[11] + E_MASS * E_ZG * E_G * s4 - x[16] - E_DL[4] * x[10] - E_DNL[4] * fabs(x[10]) * x[10];
    y[17] = Md[5] * acc[5] + (E_IY - E_IX) * x[9] * x[10] - x[17] - E_DL[5] * x[11] - E_DNL[5] * fabs(x[11]) * x[11];
}

static void ekf_rk4(const double *x, const double *u, double *xn)
{
    double k1[EN], k2[EN], k3[EN], k4[EN], xs[EN];
    orc_ekf_f(x, u, k1);
    for (int i = 0; i < EN; i++) { k1[i] *= E_DT; xs[i] = x[i] + k1[i] / 2; }
    orc_ekf_f(xs, u, k2);
    for (int i = 0; i < EN; i++) { k2[i] *= E_DT; xs[i] = x[i] + k2[i] / 3; } /* sic: /3, bluerov2_dob.cpp:630 */
    orc_ekf_f(xs, u, k3);
    for (int i = 0; i < EN; i++) { k3[i] *= E_DT; xs[i] = x[i] + k3[i]; }
    orc_ekf_f(xs, u, k4);
    for (int i = 0; i < EN; i++) { k4[i] *= E_DT; xn[i] = x[i] + (k1[i] + 2 * k2[i] + 2 * k3[i] + k4[i]) / 6; }
}

static void matmul(int n, const double *A, const double *B, double *C, int transB)
{
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) {
            double s = 0;
            for (int k = 0; k < n; k++) s += A[i * n + k] * (transB ? B[j * n + k] : B[k * n + j]);
            C[i * n + j] = s;
        }
}

void orc_ekf_step(double *ex, double *eP, const double *thr, const double *meas12, const double *acc, double *wf)
{
    const double d = 1e-6, dt = E_DT;
    double y[EN], tau[6], F[EN * EN], H[EN * EN], f0[EN], f1[EN], x1[EN], xp[EN], Pp[EN * EN], T1[EN * EN], T2[EN * EN];
    double Sm[EN * EN], Si[EN * EN], Kal[EN * EN], yp[EN], IKH[EN * EN];
    for (int i = 0; i < 6; i++) {
        double s = 0;
        for (int j = 0; j < 6; j++) s += E_K[i][j] * thr[j];
        tau[i] = s;
    }
    for (int i = 0; i < 12; i++) y[i] = meas12[i];
    for (int i = 0; i < 6; i++) y[12 + i] = tau[i];
    /* F = d RK4 / d x by forward differences (:722-735) */
    ekf_rk4(ex, thr, f0);
    for (int c = 0; c < EN; c++) {
        memcpy(x1, ex, sizeof x1);
        x1[c] += d;
        ekf_rk4(x1, thr, f1);
        for (int i = 0; i < EN; i++) F[i * EN + c] = (f1[i] - f0[i]) / d;
    }
    ekf_rk4(ex, thr, xp); /* x_pred (:527) */
    matmul(EN, F, eP, T1, 0);
    matmul(EN, T1, F, Pp, 1);
    for (int i = 0; i < EN; i++) Pp[i * EN + i] += (i < 6) ? pow(dt, 4) / 4 : pow(dt, 2); /* noise_Q (:59-62) */
    /* H by forward differences at x_pred (:738-752) */
    orc_ekf_h(xp, acc, f0);
    for (int c = 0; c < EN; c++) {
        memcpy(x1, xp, sizeof x1);
        x1[c] += d;
        orc_ekf_h(x1, acc, f1);
        for (int i = 0; i < EN; i++) H[i * EN + c] = (f1[i] - f0[i]) / d;
    }
    orc_ekf_h(xp, acc, yp);
    /* Kal = Pp H' (H Pp H' + R)^-1 (:535), R = I dt^4/4 (bluerov2_dob.h:208) */
    matmul(EN, H, Pp, T1, 0);
    matmul(EN, T1, H, Sm, 1);
    for (int i = 0; i < EN; i++) Sm[i * EN + i] += pow(dt, 4) / 4;
    mat_inverse(EN, Sm, Si);
    matmul(EN, Pp, H, T1, 1);
    matmul(EN, T1, Si, Kal, 0);
    for (int i = 0; i < EN; i++) {
        double s = xp[i];
        for (int j = 0; j < EN; j++) s += Kal[i * EN + j] * (y[j] - yp[j]);
        ex[i] = s;
    }
    /* Joseph form (:537) */
    matmul(EN, Kal, H, T1, 0);
    for (int i = 0; i < EN; i++)
        for (int j = 0; j < EN; j++) IKH[i * EN + j] = (i == j) - T1[i * EN + j];
    matmul(EN, IKH, Pp, T1, 0);
    matmul(EN, T1, IKH, T2, 1);
    matmul(EN, Kal, Kal, T1, 1);
    for (int i = 0; i < EN * EN; i++) eP[i] = T2[i] + T1[i] * (pow(dt, 4) / 4);
    /* world-frame disturbance (:540-545) */
    if (wf) {
        const double s3 = sin(y[3]), c3 = cos(y[3]), s4 = sin(y[4]), c4 = cos(y[4]), s5 = sin(y[5]), c5 = cos(y[5]);
        wf[0] = (c5 * c4) * ex[12] + (-s5 * c3 + c5 * s4 * s3) * ex[13] + (s5 * s3 + c5 * c3 * s4) * ex[14];
        wf[1] = (s5 * c4) * ex[12] + (c5 * c3 + s3 * s4 * s5) * ex[13] + (-c5 * s3 + s4 * s5 * c3) * ex[14];
        wf[2] = (-s4) * ex[12] + (c4 * s3) * ex[13] + (c4 * c3) * ex[14];
        wf[3] = ex[15] + (s5 * s4 / c4) * ex[16] + c3 * s4 / c4 * ex[17];
        wf[4] = (c3) * ex[16] + (s3) * ex[17];
        wf[5] = (s3 / c4) * ex[16] + (c3 / c4) * ex[17];
    }
}

void orc_ekf_step_batch(int nb, double *ex, double *eP, const double *thr, const double *meas12, const double *acc,
                        double *wf, int nthreads)
{
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#pragma omp parallel for schedule(static) num_threads(nthreads)
#endif
    for (int i = 0; i < nb; i++)
        orc_ekf_step(ex + (size_t)i * EN, eP + (size_t)i * EN * EN, thr + (size_t)i * 6, meas12 + (size_t)i * 12,
                     acc + (size_t)i * 6, wf ? wf + (size_t)i * 6 : 0);
}

/* ------------------------------------------------------------------------------------------------
 * Recursive least squares with variable forgetting factor -- restates BLUEROV2_AMPC::RLSFF
 * (bluerov2_ampc.cpp:731-1004) with the initialisation of the constructor (:62-79) and the constants of
 * bluerov2_ampc.h:231,239-240 (numParams 4, FF_n 5, FF_d 50).  Four independent estimators (axes X, Y, Z, N):
 *   regressor x = [body_acc, vel, 1, vel*|vel|], target y = esti_x(12 | 13 | 14 | 17) (the EKF's disturbance estimate),
 *   e = y - x.theta; e is appended to a short (5) and a long (50) window (oldest dropped); F = var_short / var_long
 *   (population variances; 0/0 = NaN on the first tick compares false, as in the reference);
 *   F > 0.8 ? lambda = max(lambda - 0.01, 0.5) : lambda = min(lambda + 0.01, 1);
 *   K = P x / (lambda + x'Px); theta += K e; P = (P - (K x') P) / lambda.
 * State of one axis, ORC_RLS_STRIDE doubles: theta[4] | P[16] row-major | lambda | F | n_short | n_long |
 * short window [5] oldest first | long window [50] oldest first | pad.
 * ---------------------------------------------------------------------------------------------- */
#define RLS_NP 4
#define RLS_FFN 5
#define RLS_FFD 50
enum { RLS_THETA = 0, RLS_P = 4, RLS_LAMBDA = 20, RLS_F = 21, RLS_NN = 22, RLS_ND = 23, RLS_EN = 24, RLS_ED = 29, RLS_STRIDE = 80 };

void orc_rls_init(double *st)
{
    for (int a = 0; a < 4; a++) {
        double *s = st + a * RLS_STRIDE;
        memset(s, 0, sizeof(double) * RLS_STRIDE);
        for (int i = 0; i < RLS_NP; i++) s[RLS_P + i * RLS_NP + i] = 1.0;   /* P = I (:63-68) */
        s[RLS_LAMBDA] = 0.9;                                                 /* :70-73 */
    }
}

static double window_push_var(double *w, double *count, int cap, double e)
{
    int n = (int)*count;
    if (n < cap) w[n++] = e;
    else { for (int i = 1; i < cap; i++) w[i - 1] = w[i]; w[cap - 1] = e; }   /* push_back + erase(begin) */
    *count = n;
    double sum = 0.0;
    for (int i = 0; i < n; i++) sum += w[i];
    const double mean = sum / n;
    double var = 0.0;
    for (int i = 0; i < n; i++) var += (w[i] - mean) * (w[i] - mean);         /* std::pow(v - mean, 2) */
    return var / n;
}

static void rls_axis(double *s, double y, double acc, double vel)
{
    const double threshold = 0.8;
    const double x[RLS_NP] = {acc, vel, 1.0, vel * fabs(vel)};
    double *th = s + RLS_THETA, *P = s + RLS_P;
    double pred = 0.0;
    for (int i = 0; i < RLS_NP; i++) pred += x[i] * th[i];
    const double e = y - pred;
    const double vn = window_push_var(s + RLS_EN, s + RLS_NN, RLS_FFN, e);
    const double vd = window_push_var(s + RLS_ED, s + RLS_ND, RLS_FFD, e);
    const double F = vn / vd;
    s[RLS_F] = F;
    double lam = s[RLS_LAMBDA];
    if (F > threshold) lam = (lam - 0.01 >= 0.5) ? lam - 0.01 : 0.5;
    else lam = (lam + 0.01 <= 1) ? lam + 0.01 : 1;
    s[RLS_LAMBDA] = lam;
    double Px[RLS_NP], K[RLS_NP], xPx = 0.0;
    for (int i = 0; i < RLS_NP; i++) {
        double t = 0.0;
        for (int j = 0; j < RLS_NP; j++) t += P[i * RLS_NP + j] * x[j];
        Px[i] = t;
    }
    for (int i = 0; i < RLS_NP; i++) xPx += x[i] * Px[i];
    for (int i = 0; i < RLS_NP; i++) K[i] = Px[i] / (lam + xPx);
    for (int i = 0; i < RLS_NP; i++) th[i] += K[i] * e;
    double Pn[RLS_NP * RLS_NP];
    for (int i = 0; i < RLS_NP; i++)
        for (int j = 0; j < RLS_NP; j++) {
            double t = 0.0;
            for (int k = 0; k < RLS_NP; k++) t += (K[i] * x[k]) * P[k * RLS_NP + j];   /* (K x') P */
            Pn[i * RLS_NP + j] = (P[i * RLS_NP + j] - t) / lam;
        }
    memcpy(P, Pn, sizeof Pn);
}

/* one RLSFF() call: state[4][ORC_RLS_STRIDE]; esti_x[18] from the EKF; body_acc[6] = (x y z phi theta psi);
 * meas12[6..11] = body velocities (v_linear_body, v_angular_body).  p_out (may be null): the OCP parameter vector
 * BLUEROV2_AMPC::solve builds from it (:340-382): p[0..3] = theta(2) / (compensate_coef | rotor_constant) and the
 * nominal p[4..15] when compensate, else p[0..3] = 0 and p[4..15] LEFT UNTOUCHED (brace placement :345-379). */
void orc_rls_step(double *st, const double *esti_x, const double *body_acc, const double *meas12, int compensate,
                  double *p_out)
{
    static const int yi[4] = {12, 13, 14, 17}, ai[4] = {0, 1, 2, 5}, vi[4] = {6, 7, 8, 11};
    for (int a = 0; a < 4; a++) rls_axis(st + a * RLS_STRIDE, esti_x[yi[a]], body_acc[ai[a]], meas12[vi[a]]);
    if (p_out) {
        const double comp = 0.032546960744430276;
        if (!compensate) { p_out[0] = p_out[1] = p_out[2] = p_out[3] = 0.0; return; }
        static const double nominal[12] = {1.7182, 0, 5.468, 0.4006, -11.7391, -20, -31.8678, -5, -18.18, -21.66, -36.99, -1.55};
        p_out[0] = st[0 * RLS_STRIDE + 2] / comp;
        p_out[1] = st[1 * RLS_STRIDE + 2] / comp;
        p_out[2] = st[2 * RLS_STRIDE + 2] / RC;
        p_out[3] = st[3 * RLS_STRIDE + 2] / RC;
        for (int i = 0; i < 12; i++) p_out[4 + i] = nominal[i];
    }
}

void orc_rls_step_batch(int nb, double *st, const double *esti_x, const double *body_acc, const double *meas12,
                        int compensate, double *p_out)
{
    for (int i = 0; i < nb; i++)
        orc_rls_step(st + (size_t)i * 4 * RLS_STRIDE, esti_x + (size_t)i * EN, body_acc + (size_t)i * 6,
                     meas12 + (size_t)i * 12, compensate, p_out ? p_out + (size_t)i * 16 : 0);
}

/* ------------------------------------------------------------------------------------------------
 * Continuous-yaw accumulator at the top of BLUEROV2_DOB::solve (bluerov2_dob.cpp:272-304).  pre_yaw, yaw_sum and
 * yaw_diff are floats in the reference (bluerov2_dob.h:234-236); psi is a double (struct Euler, :79-83).  Every
 * conversion below is the one the C++ expression performs.  state[2] = (pre_yaw, yaw_sum); returns x0[psi].
 * ---------------------------------------------------------------------------------------------- */
double orc_yaw_unwrap(float *state, double psi)
{
    const double TWO_PI = 2 * 3.14159265358979323846;
    const float pre_yaw = state[0];
    float yaw_diff;
    if (pre_yaw >= 0 && psi >= 0) {
        yaw_diff = (float)(psi - pre_yaw);
    } else if (pre_yaw >= 0 && psi < 0) {
        if (TWO_PI + psi - pre_yaw >= pre_yaw + fabs(psi)) yaw_diff = (float)(-(pre_yaw + fabs(psi)));
        else yaw_diff = (float)(TWO_PI + psi - pre_yaw);
    } else if (pre_yaw < 0 && psi >= 0) {
        if (TWO_PI - psi + pre_yaw >= fabsf(pre_yaw) + psi) yaw_diff = (float)(fabsf(pre_yaw) + psi);
        else yaw_diff = (float)(-(TWO_PI - psi + pre_yaw));
    } else {
        yaw_diff = (float)(psi - pre_yaw);
    }
    state[1] = state[1] + yaw_diff;
    state[0] = (float)psi;
    return (double)state[1];
}

