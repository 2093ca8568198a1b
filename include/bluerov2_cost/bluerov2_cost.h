/* bluerov2_cost/bluerov2_cost.h -- drop-in for c_generated_code/bluerov2_cost/bluerov2_cost.h:45-116.
 * The NLS residual is y = [x; u] (terminal y = x), bluerov2.py:144,153-154; its Jacobian is a permuted identity and its Hessian
 * structurally empty.  Same nine functions with CasADi's six entry points each as the generated header; the bodies are hand-written
 * (bluerov2_b200/csrc/acados_abi.cu, host code).  The engine itself never calls them: its Gauss-Newton Hessian Ts*W is baked
 * into the kernels. */
#ifndef BR2_DROPIN_BLUEROV2_COST_H
#define BR2_DROPIN_BLUEROV2_COST_H
#include "acados/utils/types.h"
#ifdef __cplusplus
extern "C" {
#endif

#ifndef BR2_CASADI_FUNCTION
#define BR2_CASADI_FUNCTION(f)                                                                         \
    int f(const real_t **arg, real_t **res, int *iw, real_t *w, void *mem);                            \
    int f##_work(int *sz_arg, int *sz_res, int *sz_iw, int *sz_w);                                     \
    const int *f##_sparsity_in(int i);                                                                 \
    const int *f##_sparsity_out(int i);                                                                \
    int f##_n_in(void);                                                                                \
    int f##_n_out(void);
#endif

/* initial node (ref :45-66), path nodes (:70-91): (x[12], u[4], z[0], p[16]) */
BR2_CASADI_FUNCTION(bluerov2_cost_y_0_fun)              /* -> y[16]                                              */
BR2_CASADI_FUNCTION(bluerov2_cost_y_0_fun_jac_ut_xt)    /* -> y[16], dy/d[u;x]' (16 x 16, 16 unit entries), 16x0 */
BR2_CASADI_FUNCTION(bluerov2_cost_y_0_hess)             /* (x, u, z, lam_y[16], p) -> 16 x 16, no entries        */
BR2_CASADI_FUNCTION(bluerov2_cost_y_fun)
BR2_CASADI_FUNCTION(bluerov2_cost_y_fun_jac_ut_xt)
BR2_CASADI_FUNCTION(bluerov2_cost_y_hess)
/* terminal node (ref :95-116): (x[12], u[0], z[0], p[16]) */
BR2_CASADI_FUNCTION(bluerov2_cost_y_e_fun)              /* -> y[12]                                              */
BR2_CASADI_FUNCTION(bluerov2_cost_y_e_fun_jac_ut_xt)    /* -> y[12], dy/dx' (12 x 12 identity), 12x0             */
BR2_CASADI_FUNCTION(bluerov2_cost_y_e_hess)             /* (x, u, z, lam_y[12], p) -> 12 x 12, no entries        */

#ifdef __cplusplus
}
#endif
#endif
