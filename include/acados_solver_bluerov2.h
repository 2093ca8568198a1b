/* acados_solver_bluerov2.h -- drop-in for the acados-generated header of the same name
 * (reference: bluerov2_dobmpc/scripts/c_generated_code/acados_solver_bluerov2.h:42-167).
 *
 * Same dimension macros, same public capsule struct, same 24 exported bluerov2_acados_* entry points, so
 * bluerov2_dob.cpp / bluerov2_dob_patty.cpp / ctrller/mpc.cpp / bluerov2_ampc.cpp and main_bluerov2.c compile
 * and link unchanged (CMakeLists.txt:41,72,94-97).  Behind it sits the batched B200 engine with batch = 1
 * (include/bluerov2_b200.h is the batched view).  Each prototype cites the reference definition it replaces
 * (file acados_solver_bluerov2.c).
 */
#ifndef ACADOS_SOLVER_bluerov2_H_
#define ACADOS_SOLVER_bluerov2_H_

#include "acados/utils/types.h"
#include "acados_c/ocp_nlp_interface.h"
#include "acados_c/external_function_interface.h"

/* ---- problem dimensions (reference .h:42-71; callers size their arrays with N, NY, NP: bluerov2_dob.h:68-77,166) ---- */
#define BLUEROV2_NX 12      /* states */
#define BLUEROV2_NZ 0       /* algebraic variables */
#define BLUEROV2_NU 4       /* controls */
#define BLUEROV2_NP 16      /* parameters per stage */
#define BLUEROV2_NBX 0      /* state bounds, stages 1..N-1 */
#define BLUEROV2_NBX0 12    /* state bounds at stage 0 (all equalities: x0) */
#define BLUEROV2_NBU 4      /* input bounds */
#define BLUEROV2_NSBX 0
#define BLUEROV2_NSBU 0
#define BLUEROV2_NSH 0
#define BLUEROV2_NSG 0
#define BLUEROV2_NSPHI 0
#define BLUEROV2_NSHN 0
#define BLUEROV2_NSGN 0
#define BLUEROV2_NSPHIN 0
#define BLUEROV2_NSBXN 0
#define BLUEROV2_NS 0
#define BLUEROV2_NSN 0      /* no soft constraints anywhere */
#define BLUEROV2_NG 0       /* general linear constraints */
#define BLUEROV2_NBXN 0     /* terminal state bounds */
#define BLUEROV2_NGN 0
#define BLUEROV2_NY0 16     /* residual size, stage 0: y = [x; u] */
#define BLUEROV2_NY 16      /* residual size, stages 1..N-1 */
#define BLUEROV2_NYN 12     /* terminal residual: y = x */
#ifndef BLUEROV2_N              /* overridable like the generated header's */
#define BLUEROV2_N 80           /* shooting intervals (Ts = 1/80 s) */
#endif
#define BLUEROV2_NH 0       /* nonlinear constraints */
#define BLUEROV2_NPHI 0
#define BLUEROV2_NHN 0
#define BLUEROV2_NPHIN 0
#define BLUEROV2_NR 0

#ifdef __cplusplus
extern "C" {
#endif

/* ---- the capsule (reference .h:79-127).  Callers reach into it (nlp_config, nlp_dims, nlp_in, nlp_out, nlp_solver:
 * bluerov2_dob.cpp:320,371,384,386,388), so member names, types and ORDER are the generated ones. ---- */
typedef struct bluerov2_solver_capsule {
    /* acados objects */
    ocp_nlp_in *nlp_in;
    ocp_nlp_out *nlp_out, *sens_out;
    ocp_nlp_solver *nlp_solver;
    void *nlp_opts;
    ocp_nlp_plan_t *nlp_solver_plan;
    ocp_nlp_config *nlp_config;
    ocp_nlp_dims *nlp_dims;
    unsigned int nlp_np;                                                            /* number of parameters */
    /* per-stage external functions: dynamics (N of each), path cost (N - 1 of each) */
    external_function_param_casadi *forw_vde_casadi, *expl_ode_fun;
    external_function_param_casadi *cost_y_fun, *cost_y_fun_jac_ut_xt, *cost_y_hess;
    /* stage-0 and terminal cost, by value */
    external_function_param_casadi cost_y_0_fun, cost_y_0_fun_jac_ut_xt, cost_y_0_hess;
    external_function_param_casadi cost_y_e_fun, cost_y_e_fun_jac_ut_xt, cost_y_e_hess;
} bluerov2_solver_capsule;

#define BR2_GEN ACADOS_SYMBOL_EXPORT
typedef bluerov2_solver_capsule br2_capsule_t_;     /* shorthand for the prototypes below only */

/* ---- life cycle ---- */
BR2_GEN bluerov2_solver_capsule *bluerov2_acados_create_capsule(void);             /* .c:87-93   plain malloc; needs no CUDA context
                                                                                      (called from a member initialiser, bluerov2_dob.h:168) */
BR2_GEN int bluerov2_acados_free_capsule(br2_capsule_t_ *c);                        /* .c:96-100 */
BR2_GEN int bluerov2_acados_create(br2_capsule_t_ *c);                              /* .c:103-108 N = BLUEROV2_N, Ts = 0.0125 s */
BR2_GEN int bluerov2_acados_create_with_discretization(br2_capsule_t_ *c, int n_steps, double *steps);
                                                                                    /* .c:734-783 returns 1 if n_steps != BLUEROV2_N and steps == NULL */
BR2_GEN int bluerov2_acados_reset(br2_capsule_t_ *c, int reset_qp_solver_mem);      /* .c:797-830 zero x, u (and the multipliers this engine does not carry) */
BR2_GEN int bluerov2_acados_free(br2_capsule_t_ *c);                                /* .c:954-998 */
/* ---- problem data ---- */
BR2_GEN int bluerov2_acados_update_time_steps(br2_capsule_t_ *c, int N, double *steps);   /* .c:111-132 returns 1 on N mismatch; sets Ts and the
                                                                                             cost scaling of every stage */
BR2_GEN int bluerov2_acados_update_qp_solver_cond_N(br2_capsule_t_ *c, int cond_N); /* .c:788-794 always prints and exit(1)s, as the reference */
BR2_GEN int bluerov2_acados_update_params(br2_capsule_t_ *c, int stage, double *p, int np);      /* .c:835-883 np != 16 prints and exit(1)s */
BR2_GEN int bluerov2_acados_update_params_sparse(br2_capsule_t_ *c, int stage, int *idx, double *p, int n_update);   /* .c:886-943 */
/* ---- the hot call and its companions ---- */
BR2_GEN int bluerov2_acados_solve(br2_capsule_t_ *c);                               /* .c:945-951 one SQP-RTI step (ocp_nlp_solve); acados status */
BR2_GEN void bluerov2_acados_print_stats(br2_capsule_t_ *c);                        /* .c:1001-1028 */
BR2_GEN int bluerov2_acados_custom_update(br2_capsule_t_ *c, double *data, int data_len);        /* .c:1030-1036 prints two lines, returns 1 */
/* ---- getters (.c:1040-1047): return the capsule member of the same name ---- */
#define BR2_GETTER(type, what) BR2_GEN type bluerov2_acados_get_##what(br2_capsule_t_ *c);
BR2_GETTER(ocp_nlp_in *, nlp_in)
BR2_GETTER(ocp_nlp_out *, nlp_out)
BR2_GETTER(ocp_nlp_out *, sens_out)
BR2_GETTER(ocp_nlp_solver *, nlp_solver)
BR2_GETTER(ocp_nlp_config *, nlp_config)
BR2_GETTER(void *, nlp_opts)
BR2_GETTER(ocp_nlp_dims *, nlp_dims)
BR2_GETTER(ocp_nlp_plan_t *, nlp_plan)
#undef BR2_GETTER
#undef BR2_GEN

#ifdef __cplusplus
}
#endif
#endif
