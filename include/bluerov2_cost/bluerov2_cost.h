/* bluerov2_cost/bluerov2_cost.h -- drop-in for c_generated_code/bluerov2_cost/bluerov2_cost.h.
 * The NLS residual is y = [x; u] (terminal y = x), bluerov2.py:144,153-154; its Jacobian is a permuted identity
 * and its Hessian structurally empty, so the engine never calls these: they exist for link compatibility. */
#ifndef BR2_DROPIN_BLUEROV2_COST_H
#define BR2_DROPIN_BLUEROV2_COST_H
#include "acados/utils/types.h"
#ifdef __cplusplus
extern "C" {
#endif
/* CasADi evaluation signature: inputs in[], outputs out[], unused work vectors */
#define BR2_COST_EVAL(f) int f(const real_t **in, real_t **out, int *iwork, real_t *rwork, void *mem);
BR2_COST_EVAL(bluerov2_cost_y_0_fun)   /* (x[12], u[4], z[0], p[16]) -> y[16]   ref bluerov2_cost_y_0_fun.c */
BR2_COST_EVAL(bluerov2_cost_y_fun)     /* same residual on the path stages      ref bluerov2_cost_y_fun.c:60 */
BR2_COST_EVAL(bluerov2_cost_y_e_fun)   /* (x[12], u[0], z[0], p[16]) -> y[12]   ref bluerov2_cost_y_e_fun.c:58 */
#undef BR2_COST_EVAL
#ifdef __cplusplus
}
#endif
#endif
