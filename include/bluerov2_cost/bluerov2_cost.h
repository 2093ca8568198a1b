/* bluerov2_cost/bluerov2_cost.h -- drop-in for c_generated_code/bluerov2_cost/bluerov2_cost.h.
 * The NLS residual is y = [x; u] (terminal y = x), bluerov2.py:144,153-154; its Jacobian is a permuted identity
 * and its Hessian structurally empty, so the engine never calls these: they exist for link compatibility. */
#ifndef bluerov2_COST
#define bluerov2_COST
#include "acados/utils/types.h"
#ifdef __cplusplus
extern "C" {
#endif
/* (x[12], u[4], z[0], p[16]) -> y[16]                     reference: bluerov2_cost_y_fun.c:60 */
int bluerov2_cost_y_0_fun(const real_t** arg, real_t** res, int* iw, real_t* w, void *mem);
int bluerov2_cost_y_fun(const real_t** arg, real_t** res, int* iw, real_t* w, void *mem);
/* (x[12], u[0], z[0], p[16]) -> y[12]                     reference: bluerov2_cost_y_e_fun.c:58 */
int bluerov2_cost_y_e_fun(const real_t** arg, real_t** res, int* iw, real_t* w, void *mem);
#ifdef __cplusplus
}
#endif
#endif
