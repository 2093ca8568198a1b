/* bluerov2_model/bluerov2_model.h -- drop-in for c_generated_code/bluerov2_model/bluerov2_model.h:46-67.
 * Same CasADi calling convention; the bodies are hand-written (bluerov2_b200/csrc/model.cuh evaluated on the
 * host) instead of CasADi-generated.  Sparsities are dense column-major like the generated code. */
#ifndef bluerov2_MODEL
#define bluerov2_MODEL
#include "acados/utils/types.h"
#ifdef __cplusplus
extern "C" {
#endif

/* (x[12], u[4], p[16]) -> f[12]                          reference: bluerov2_expl_ode_fun.c:66 */
int bluerov2_expl_ode_fun(const real_t** arg, real_t** res, int* iw, real_t* w, void *mem);
int bluerov2_expl_ode_fun_work(int *, int *, int *, int *);
const int *bluerov2_expl_ode_fun_sparsity_in(int);
const int *bluerov2_expl_ode_fun_sparsity_out(int);
int bluerov2_expl_ode_fun_n_in(void);
int bluerov2_expl_ode_fun_n_out(void);

/* (x, Sx[12x12], Su[12x4], u, p) -> (f, Jx Sx, Jx Su + Ju)   reference: bluerov2_expl_vde_forw.c:73 */
int bluerov2_expl_vde_forw(const real_t** arg, real_t** res, int* iw, real_t* w, void *mem);
int bluerov2_expl_vde_forw_work(int *, int *, int *, int *);
const int *bluerov2_expl_vde_forw_sparsity_in(int);
const int *bluerov2_expl_vde_forw_sparsity_out(int);
int bluerov2_expl_vde_forw_n_in(void);
int bluerov2_expl_vde_forw_n_out(void);

/* (x, lam[12], u, p) -> [Jx' lam; Ju' lam] (16)            reference: bluerov2_expl_vde_adj.c:69 */
int bluerov2_expl_vde_adj(const real_t** arg, real_t** res, int* iw, real_t* w, void *mem);
int bluerov2_expl_vde_adj_work(int *, int *, int *, int *);
const int *bluerov2_expl_vde_adj_sparsity_in(int);
const int *bluerov2_expl_vde_adj_sparsity_out(int);
int bluerov2_expl_vde_adj_n_in(void);
int bluerov2_expl_vde_adj_n_out(void);

#ifdef __cplusplus
} /* extern "C" */
#endif
#endif  // bluerov2_MODEL
