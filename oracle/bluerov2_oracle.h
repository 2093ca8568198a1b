/* oracle/bluerov2_oracle.h -- CPU restatement of the BlueROV2 SQP-RTI hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under bluerov2_b200/ (the product) may include, link or load this.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it, and
 * only as the checker / the CPU arm -- never as the thing shipped.
 *
 * Parity pin status (see DESIGN.md "Oracle"):
 *   - dynamics / forward VDE: PINNED against the reference's own CasADi-generated C
 *     (oracle/_ref/libbluerov2_casadi_ref.so, built from /root/reference by oracle/Makefile) and against
 *     golden vectors generated from it (tests/golden/casadi_vde.npz).
 *   - NLP/QP layer (acados SQP_RTI + HPIPM): "parity unpinned" -- acados/HPIPM/BLASFEO are un-vendored
 *     external dependencies (version not pinned by the reference, README.md:43-48) and cannot be built
 *     here.  The restatement follows the reference's call sites and options
 *     (acados_solver_bluerov2.c:137-164,389-394,424-479,501-571,611-675,681-708) and solves the
 *     strictly convex QP of each RTI step to 1e-12, so its solution is the unique QP solution any
 *     correct solver (HPIPM included) converges to.
 *   - EKF: restates bluerov2_dob.cpp:41-65,495-545,621-752 line by line (Eigen absent: general LU
 *     inverse with partial pivoting stands in for MatrixXd::inverse()).
 *
 * Conventions: all matrices row-major unless the name says "cm" (column-major, CasADi's layout).
 */
#ifndef BLUEROV2_ORACLE_H_
#define BLUEROV2_ORACLE_H_

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_NX 12
#define ORC_NU 4
#define ORC_NP 16
#define ORC_NY 16
#define ORC_NMAX 256

/* ---- model (bluerov2.py:77-137  ==  bluerov2_expl_ode_fun.c:66-321) ---- */
void orc_ode(const double *x, const double *u, const double *p, double *f);
/* analytic Jacobians: Jx[i*12+j] = d f_i / d x_j (row-major), Ju[i*4+j] */
void orc_jac(const double *x, const double *u, const double *p, double *Jx, double *Ju);
/* same contract as bluerov2_expl_vde_forw (bluerov2_expl_vde_forw.c:73): column-major Sx[12x12], Su[12x4] */
void orc_vde_forw_cm(const double *x, const double *Sx, const double *Su, const double *u, const double *p,
                     double *f, double *dSx, double *dSu);

/* Route the ERK through the reference's CasADi forward VDE (function pointer dlsym'd by the caller from
 * oracle/_ref) instead of the hand restatement; pass NULL to go back. */
void orc_set_casadi_vde(void *casadi_fn);

/* one classical RK4 step of length h on [x | Sx=I | Su=0] (acados ERK, 4 stages, 1 step:
 * acados_solver_bluerov2.c:633,639).  A (12x12) and B (12x4) row-major. */
void orc_erk4_sens(const double *x, const double *u, const double *p, double h,
                   double *xn, double *A, double *B);
/* plain RK4 of the OCP model (plant for closed-loop studies) */
void orc_erk4(const double *x, const double *u, const double *p, double h, double *xn);

/* ---- OCP-QP (the QP of one RTI step) ---- */
typedef struct {
    int N;
    const double *A;    /* N x 144 */
    const double *B;    /* N x 48  */
    const double *b;    /* N x 12  */
    const double *Qd;   /* (N+1) x 12 diagonal state Hessians (stage k<N: Ts_k*W, terminal: We) */
    const double *Rd;   /* N x 4 */
    const double *q;    /* (N+1) x 12 */
    const double *r;    /* N x 4 */
    const double *lb;   /* N x 4  bounds on du */
    const double *ub;   /* N x 4 */
    const double *dx0;  /* 12 */
} orc_qp;

typedef struct {
    int iters;
    int status;         /* 0 ok, 2 max-iter */
    double mu, res_stat, res_ineq, res_comp;
} orc_qp_stats;

/* Mehrotra predictor-corrector primal-dual IPM; every Newton system is an unconstrained LQR solved by a
 * Riccati recursion.  Outputs dx ((N+1)x12), du (Nx4), pi ((N+1)x12 costates), lam_l/lam_u (Nx4). */
int orc_qp_solve(const orc_qp *qp, int max_iter, double tol,
                 double *dx, double *du, double *pi, double *lam_l, double *lam_u, orc_qp_stats *st);
/* the same QP by full condensing + a dense interior-point iteration (the cost profile of FULL_CONDENSING_HPIPM); orc_rti_step uses
 * it when orc_set_qp_mode(1) */
int orc_qp_solve_dense(const orc_qp *qp, int max_iter, double tol,
                       double *dx, double *du, double *pi, double *lam_l, double *lam_u, orc_qp_stats *st);
void orc_set_qp_mode(int mode);    /* 0 = Riccati IPM (default), 1 = full condensing + dense IPM; process-wide */
int orc_get_qp_mode(void);
/* KKT residuals of a candidate (dx,du,pi,lam) -- the solver-independent certificate used by the tests.
 * res[0]=stationarity, res[1]=dynamics, res[2]=bound violation, res[3]=complementarity, res[4]=min(lam) */
void orc_qp_kkt(const orc_qp *qp, const double *dx, const double *du, const double *pi,
                const double *lam_l, const double *lam_u, double *res);

/* ---- one SQP-RTI step (acados ocp_nlp_solve with nlp_solver=SQP_RTI, rti_phase 0) ----
 * Ts[N]; W[16] diag; We[12] diag; lbu/ubu[4]; x0[12]; yref[(N+1)*16] (terminal row: first 12 used);
 * p: parameters, p + k*p_stride is stage k's 16-vector (p_stride = 0: one vector for all stages);
 * X[(N+1)*12], U[N*4] in/out (linearisation point -> full-step result).
 * info[8]: ipm iters, qp status, mu, res_stat, res_ineq, res_comp, |b|inf (dynamics gap), reserved.
 * returns acados-style status (0 success, 2 max iter, 1 NaN). */
int orc_rti_step(int N, const double *Ts, const double *W, const double *We,
                 const double *lbu, const double *ubu,
                 const double *x0, const double *yref, const double *p, int p_stride,
                 double *X, double *U, int max_iter, double tol, double *info);

/* linearisation only: fills A (N*144), B (N*48), b (N*12) at (X,U) */
void orc_linearize(int N, const double *Ts, const double *p, int p_stride,
                   const double *X, const double *U, double *A, double *B, double *b);

/* batched RTI step, OpenMP over instances (CPU baseline).  Arrays are [B][...] contiguous per instance,
 * p is [B][16].  status[B], info[B][8] (may be NULL).  Returns number of threads used. */
int orc_rti_step_batch(int nb, int N, const double *Ts, const double *W, const double *We,
                       const double *lbu, const double *ubu,
                       const double *x0, const double *yref, const double *p,
                       double *X, double *U, int max_iter, double tol,
                       int *status, double *info, int nthreads);

/* 4 -> 6 thrust allocation (bluerov2_dob.cpp:390-395) */
void orc_thrust_alloc(const double *u0, double *thrust6);

/* ---- 18-state disturbance-observer EKF (bluerov2_dob.cpp:495-545 + helpers) ----
 * esti_x[18], esti_P[18*18] in/out; thrusts6 = measured thruster forces (meas_u); meas12 = pose (6) +
 * body velocities (6); body_acc[6]; wf_dist[6] out (world-frame disturbance, :540-545). */
void orc_ekf_init(double *esti_x, double *esti_P);
void orc_ekf_f(const double *x18, const double *u6, double *xdot18);
void orc_ekf_h(const double *x18, const double *body_acc6, double *y18);
void orc_ekf_step(double *esti_x, double *esti_P, const double *thrusts6, const double *meas12,
                  const double *body_acc6, double *wf_dist);
void orc_ekf_step_batch(int nb, double *esti_x, double *esti_P, const double *thrusts6,
                        const double *meas12, const double *body_acc6, double *wf_dist, int nthreads);
/* model 0 (default): BLUEROV2_DOB damping; model 1: BLUEROV2_AMPC's filter (bluerov2_ampc.cpp:41-42,658-696: Dl = 0,
 * no quadratic damping).  Process-wide switch. */
void orc_ekf_set_model(int model);

/* ---- RLS with variable forgetting factor, BLUEROV2_AMPC::RLSFF (bluerov2_ampc.cpp:731-1004, init :62-79) ----
 * state[4][ORC_RLS_STRIDE] (axes X, Y, Z, N): theta[4] | P[16] | lambda | F | n_short | n_long | short window[5] |
 * long window[50] | pad.  p_out[16] (may be null): parameters as BLUEROV2_AMPC::solve fills them (:340-382). */
#define ORC_RLS_STRIDE 80
void orc_rls_init(double *state);
void orc_rls_step(double *state, const double *esti_x18, const double *body_acc6, const double *meas12,
                  int compensate, double *p_out);
void orc_rls_step_batch(int nb, double *state, const double *esti_x18, const double *body_acc6,
                        const double *meas12, int compensate, double *p_out);

/* ---- continuous yaw of BLUEROV2_DOB::solve (bluerov2_dob.cpp:272-304): state[2] = (pre_yaw, yaw_sum) as FLOATS
 * (bluerov2_dob.h:234-236); returns the value the node writes into x0[psi] ---- */
double orc_yaw_unwrap(float *state, double psi);

#ifdef __cplusplus
}
#endif
#endif
