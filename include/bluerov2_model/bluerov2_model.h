/* bluerov2_model/bluerov2_model.h -- drop-in for c_generated_code/bluerov2_model/bluerov2_model.h:46-67.
 * Same CasADi calling convention; the bodies are hand-written (bluerov2_b200/csrc/model.cuh evaluated on the
 * host) instead of CasADi-generated.  Sparsities are dense column-major like the generated code. */
#ifndef BR2_DROPIN_BLUEROV2_MODEL_H
#define BR2_DROPIN_BLUEROV2_MODEL_H
#include "acados/utils/types.h"
#ifdef __cplusplus
extern "C" {
#endif

/* The six entry points CasADi's C code generator emits per function f: evaluation (arg / res pointer arrays, integer and
 * real work vectors, memory slot), work sizes, CCS sparsity of input / output i, number of inputs / outputs. */
#ifndef BR2_CASADI_FUNCTION
#define BR2_CASADI_FUNCTION(f)                                                                         \
    int f(const real_t **arg, real_t **res, int *iw, real_t *w, void *mem);                            \
    int f##_work(int *sz_arg, int *sz_res, int *sz_iw, int *sz_w);                                     \
    const int *f##_sparsity_in(int i);                                                                 \
    const int *f##_sparsity_out(int i);                                                                \
    int f##_n_in(void);                                                                                \
    int f##_n_out(void);
#endif

BR2_CASADI_FUNCTION(bluerov2_expl_ode_fun)   /* (x[12], u[4], p[16]) -> f[12]                               ref bluerov2_expl_ode_fun.c:66  */
BR2_CASADI_FUNCTION(bluerov2_expl_vde_forw)  /* (x, Sx[12x12], Su[12x4], u, p) -> (f, Jx Sx, Jx Su + Ju)    ref bluerov2_expl_vde_forw.c:73 */
BR2_CASADI_FUNCTION(bluerov2_expl_vde_adj)   /* (x, lam[12], u, p) -> [Jx' lam; Ju' lam] (16)               ref bluerov2_expl_vde_adj.c:69  */

#ifdef __cplusplus
}
#endif
#endif
