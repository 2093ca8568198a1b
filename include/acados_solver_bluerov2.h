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

/* reference .h:42-71 */
#define BLUEROV2_NX     12
#define BLUEROV2_NZ     0
#define BLUEROV2_NU     4
#define BLUEROV2_NP     16
#define BLUEROV2_NBX    0
#define BLUEROV2_NBX0   12
#define BLUEROV2_NBU    4
#define BLUEROV2_NSBX   0
#define BLUEROV2_NSBU   0
#define BLUEROV2_NSH    0
#define BLUEROV2_NSG    0
#define BLUEROV2_NSPHI  0
#define BLUEROV2_NSHN   0
#define BLUEROV2_NSGN   0
#define BLUEROV2_NSPHIN 0
#define BLUEROV2_NSBXN  0
#define BLUEROV2_NS     0
#define BLUEROV2_NSN    0
#define BLUEROV2_NG     0
#define BLUEROV2_NBXN   0
#define BLUEROV2_NGN    0
#define BLUEROV2_NY0    16
#define BLUEROV2_NY     16
#define BLUEROV2_NYN    12
#ifndef BLUEROV2_N
#define BLUEROV2_N      80
#endif
#define BLUEROV2_NH     0
#define BLUEROV2_NPHI   0
#define BLUEROV2_NHN    0
#define BLUEROV2_NPHIN  0
#define BLUEROV2_NR     0

#ifdef __cplusplus
extern "C" {
#endif

/* reference .h:79-127 -- member names and order preserved (callers dereference nlp_config, nlp_dims, nlp_in,
 * nlp_out, nlp_solver: bluerov2_dob.cpp:320,371,384,386,388) */
typedef struct bluerov2_solver_capsule
{
    ocp_nlp_in *nlp_in;
    ocp_nlp_out *nlp_out;
    ocp_nlp_out *sens_out;
    ocp_nlp_solver *nlp_solver;
    void *nlp_opts;
    ocp_nlp_plan_t *nlp_solver_plan;
    ocp_nlp_config *nlp_config;
    ocp_nlp_dims *nlp_dims;

    unsigned int nlp_np;

    external_function_param_casadi *forw_vde_casadi;
    external_function_param_casadi *expl_ode_fun;

    external_function_param_casadi *cost_y_fun;
    external_function_param_casadi *cost_y_fun_jac_ut_xt;
    external_function_param_casadi *cost_y_hess;

    external_function_param_casadi cost_y_0_fun;
    external_function_param_casadi cost_y_0_fun_jac_ut_xt;
    external_function_param_casadi cost_y_0_hess;

    external_function_param_casadi cost_y_e_fun;
    external_function_param_casadi cost_y_e_fun_jac_ut_xt;
    external_function_param_casadi cost_y_e_hess;
} bluerov2_solver_capsule;

/* .c:87-93: plain malloc; needs no CUDA context (the nodes call it in a member initialiser, bluerov2_dob.h:168) */
ACADOS_SYMBOL_EXPORT bluerov2_solver_capsule * bluerov2_acados_create_capsule(void);
/* .c:96-100 */
ACADOS_SYMBOL_EXPORT int bluerov2_acados_free_capsule(bluerov2_solver_capsule *capsule);
/* .c:103-108: N = BLUEROV2_N, generated time steps (Ts = 0.0125 s) */
ACADOS_SYMBOL_EXPORT int bluerov2_acados_create(bluerov2_solver_capsule * capsule);
/* .c:797-830: zero x, u (and the multipliers this engine does not carry) */
ACADOS_SYMBOL_EXPORT int bluerov2_acados_reset(bluerov2_solver_capsule* capsule, int reset_qp_solver_mem);
/* .c:734-783: returns 1 when n_time_steps != BLUEROV2_N and new_time_steps == NULL */
ACADOS_SYMBOL_EXPORT int bluerov2_acados_create_with_discretization(bluerov2_solver_capsule * capsule, int n_time_steps, double* new_time_steps);
/* .c:111-132: returns 1 on N mismatch; sets Ts and the cost scaling of every stage */
ACADOS_SYMBOL_EXPORT int bluerov2_acados_update_time_steps(bluerov2_solver_capsule * capsule, int N, double* new_time_steps);
/* .c:788-794: always prints and exit(1)s (no partial condensing in the reference either) */
ACADOS_SYMBOL_EXPORT int bluerov2_acados_update_qp_solver_cond_N(bluerov2_solver_capsule * capsule, int qp_solver_cond_N);
/* .c:835-883: np != 16 prints and exit(1)s */
ACADOS_SYMBOL_EXPORT int bluerov2_acados_update_params(bluerov2_solver_capsule * capsule, int stage, double *value, int np);
/* .c:886-943 */
ACADOS_SYMBOL_EXPORT int bluerov2_acados_update_params_sparse(bluerov2_solver_capsule * capsule, int stage, int *idx, double *p, int n_update);

/* .c:945-951: one SQP-RTI step (ocp_nlp_solve); returns the acados status */
ACADOS_SYMBOL_EXPORT int bluerov2_acados_solve(bluerov2_solver_capsule * capsule);
/* .c:954-998 */
ACADOS_SYMBOL_EXPORT int bluerov2_acados_free(bluerov2_solver_capsule * capsule);
/* .c:1001-1028 */
ACADOS_SYMBOL_EXPORT void bluerov2_acados_print_stats(bluerov2_solver_capsule * capsule);
/* .c:1030-1036: prints two lines, returns 1 */
ACADOS_SYMBOL_EXPORT int bluerov2_acados_custom_update(bluerov2_solver_capsule* capsule, double* data, int data_len);

/* .c:1040-1047 */
ACADOS_SYMBOL_EXPORT ocp_nlp_in *bluerov2_acados_get_nlp_in(bluerov2_solver_capsule * capsule);
ACADOS_SYMBOL_EXPORT ocp_nlp_out *bluerov2_acados_get_nlp_out(bluerov2_solver_capsule * capsule);
ACADOS_SYMBOL_EXPORT ocp_nlp_out *bluerov2_acados_get_sens_out(bluerov2_solver_capsule * capsule);
ACADOS_SYMBOL_EXPORT ocp_nlp_solver *bluerov2_acados_get_nlp_solver(bluerov2_solver_capsule * capsule);
ACADOS_SYMBOL_EXPORT ocp_nlp_config *bluerov2_acados_get_nlp_config(bluerov2_solver_capsule * capsule);
ACADOS_SYMBOL_EXPORT void *bluerov2_acados_get_nlp_opts(bluerov2_solver_capsule * capsule);
ACADOS_SYMBOL_EXPORT ocp_nlp_dims *bluerov2_acados_get_nlp_dims(bluerov2_solver_capsule * capsule);
ACADOS_SYMBOL_EXPORT ocp_nlp_plan_t *bluerov2_acados_get_nlp_plan(bluerov2_solver_capsule * capsule);

#ifdef __cplusplus
} /* extern "C" */
#endif

#endif  // ACADOS_SOLVER_bluerov2_H_
