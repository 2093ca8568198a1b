/* acados_c/ocp_nlp_interface.h -- the slice of acados' C interface that the BlueROV2 solver ABI and its callers
 * use, served by the B200 engine instead of libacados.
 *
 * Reference call sites (bluerov2_dobmpc/...):
 *   src/bluerov2_dob.cpp:320-321   ocp_nlp_constraints_model_set(cfg, dims, in, 0, "lbx"/"ubx", x0)
 *   src/bluerov2_dob.cpp:371       ocp_nlp_cost_model_set(cfg, dims, in, i, "yref", yref[i])
 *   src/bluerov2_dob.cpp:384       capsule->nlp_out->inf_norm_res            (struct member read)
 *   src/bluerov2_dob.cpp:386       ocp_nlp_get(cfg, solver, "time_tot", &t)
 *   src/bluerov2_dob.cpp:388       ocp_nlp_out_get(cfg, dims, out, 0, "u", u0)
 *   src/ctrller/mpc.cpp:50-57,75,77,121-137   same calls
 *   scripts/c_generated_code/main_bluerov2.c:144-146,213-219,224-227,247-248
 *        "idxbx", ocp_nlp_out_set "x"/"u", ocp_nlp_solver_opts_set "rti_phase", nlp_dims->N,
 *        "kkt_norm_inf", "sqp_iter"
 *   scripts/c_generated_code/acados_solver_bluerov2.c:1001-1028 (print_stats) "stat_n", "stat_m", "statistics"
 * Semantics follow acados: values are copied on set; getters write into caller buffers; an unknown field prints
 * a message and exit(1)s.  The objects are opaque except for the members callers read (dims->N,
 * out->inf_norm_res).
 */
#ifndef BR2_ACADOS_C_OCP_NLP_INTERFACE_H_
#define BR2_ACADOS_C_OCP_NLP_INTERFACE_H_

#include "acados/utils/types.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef enum { SQP, SQP_RTI, INVALID_NLP_SOLVER } ocp_nlp_solver_t;
typedef enum { LINEAR_LS, NONLINEAR_LS, CONVEX_OVER_NONLINEAR, EXTERNAL, INVALID_COST } ocp_nlp_cost_t;
typedef enum { CONTINUOUS_MODEL, DISCRETE_MODEL, INVALID_DYNAMICS } ocp_nlp_dynamics_t;
typedef enum { BGH, BGP, INVALID_CONSTRAINT } ocp_nlp_constraints_t;
typedef enum { NO_REGULARIZE, MIRROR, PROJECT, PROJECT_REDUC_HESS, CONVEXIFY, INVALID_REGULARIZE } ocp_nlp_reg_t;
typedef enum { ERK, IRK, GNSF, LIFTED_IRK, INVALID_SIM_SOLVER } sim_solver_t;
typedef enum { GAUSS_LEGENDRE, GAUSS_RADAU_IIA } sim_collocation_type;
typedef enum
{
    PARTIAL_CONDENSING_HPIPM,
    FULL_CONDENSING_HPIPM,
    FULL_CONDENSING_QPOASES,
    /* what this library actually runs: Riccati-recursion IPM on the un-condensed OCP-QP, sm_100a */
    RICCATI_IPM_B200,
    INVALID_QP_SOLVER
} ocp_qp_solver_t;

typedef struct { sim_solver_t sim_solver; } sim_solver_plan_t;
typedef struct { ocp_qp_solver_t qp_solver; } ocp_qp_solver_plan_t;

typedef struct ocp_nlp_plan_t
{
    ocp_qp_solver_plan_t ocp_qp_solver_plan;
    sim_solver_plan_t *sim_solver_plan;
    ocp_nlp_solver_t nlp_solver;
    ocp_nlp_reg_t regularization;
    ocp_nlp_cost_t *nlp_cost;
    ocp_nlp_dynamics_t *nlp_dynamics;
    ocp_nlp_constraints_t *nlp_constraints;
    int N;
} ocp_nlp_plan_t;

typedef struct ocp_nlp_config
{
    int N;
    void *ctx;      /* engine context (private) */
} ocp_nlp_config;

typedef struct ocp_nlp_dims
{
    int *nv, *nx, *nu, *ni, *nz, *ns;   /* per stage, N+1 entries */
    int N;
    void *ctx;
} ocp_nlp_dims;

typedef struct ocp_nlp_in
{
    double *Ts;     /* N time steps */
    void *ctx;
} ocp_nlp_in;

typedef struct ocp_nlp_out
{
    double *x;      /* (N+1) x nx host mirror of the iterate */
    double *u;      /* N x nu */
    int sqp_iter;
    int qp_iter;
    double inf_norm_res;
    double total_time;
    void *ctx;
} ocp_nlp_out;

typedef struct ocp_nlp_solver
{
    ocp_nlp_config *config;
    ocp_nlp_dims *dims;
    void *opts;
    void *ctx;
} ocp_nlp_solver;

/* setters: value is copied.  Fields: constraints "lbx" "ubx" "idxbx" "idxbxe" (stage 0), "lbu" "ubu" "idxbu";
 * cost "yref" "W" "scaling"; in "Ts"; out "x" "u" (+ "sl" "su" "lam" "t" "z" "pi" accepted and ignored). */
ACADOS_SYMBOL_EXPORT int ocp_nlp_constraints_model_set(ocp_nlp_config *config, ocp_nlp_dims *dims, ocp_nlp_in *in,
                                                       int stage, const char *field, void *value);
ACADOS_SYMBOL_EXPORT int ocp_nlp_cost_model_set(ocp_nlp_config *config, ocp_nlp_dims *dims, ocp_nlp_in *in,
                                                int stage, const char *field, void *value);
ACADOS_SYMBOL_EXPORT int ocp_nlp_in_set(ocp_nlp_config *config, ocp_nlp_dims *dims, ocp_nlp_in *in, int stage,
                                        const char *field, void *value);
ACADOS_SYMBOL_EXPORT void ocp_nlp_out_set(ocp_nlp_config *config, ocp_nlp_dims *dims, ocp_nlp_out *out, int stage,
                                          const char *field, void *value);
/* getters: "x" "u" "kkt_norm_inf" */
ACADOS_SYMBOL_EXPORT void ocp_nlp_out_get(ocp_nlp_config *config, ocp_nlp_dims *dims, ocp_nlp_out *out, int stage,
                                          const char *field, void *value);
/* "time_tot" "time_lin" "time_qp" (double), "sqp_iter" "qp_iter" "stat_n" "stat_m" "status" (int),
 * "statistics" (double[(stat_n+1) * min(sqp_iter+1, stat_m)], column-major) */
ACADOS_SYMBOL_EXPORT void ocp_nlp_get(ocp_nlp_config *config, ocp_nlp_solver *solver, const char *field,
                                      void *return_value_);
/* "rti_phase" (int: 0 preparation+feedback, 1 preparation, 2 feedback), "qp_iter_max"/"qp_solver_iter_max" (int),
 * "qp_tol" (double), "print_level" (int); the generated-code option strings ("globalization", "qp_hpipm_mode", ...)
 * are accepted and ignored. */
ACADOS_SYMBOL_EXPORT void ocp_nlp_solver_opts_set(ocp_nlp_config *config, void *opts_, const char *field, void *value);
ACADOS_SYMBOL_EXPORT int ocp_nlp_solve(ocp_nlp_solver *solver, ocp_nlp_in *nlp_in, ocp_nlp_out *nlp_out);
ACADOS_SYMBOL_EXPORT int ocp_nlp_precompute(ocp_nlp_solver *solver, ocp_nlp_in *nlp_in, ocp_nlp_out *nlp_out);

#ifdef __cplusplus
}
#endif
#endif
