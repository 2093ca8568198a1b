/* acados_c/external_function_interface.h -- the capsule embeds external_function_param_casadi objects
 * (acados_solver_bluerov2.h:93-117).  In acados they wrap CasADi-generated C; here the dynamics and cost are
 * hand-written CUDA, so the struct only keeps the fields the generated glue touches (function pointers, np, p)
 * so that callers which inspect or size the capsule keep compiling.  Layout is private to this library. */
#ifndef BR2_ACADOS_C_EXTERNAL_FUNCTION_INTERFACE_H_
#define BR2_ACADOS_C_EXTERNAL_FUNCTION_INTERFACE_H_
#include "acados/utils/types.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct external_function_param_casadi
{
    /* generic-function part (acados external_function_generic) */
    void (*evaluate)(void *, void *, void *, void *, void *);
    void (*get_nparam)(void *, int *);
    void (*set_param)(void *, double *);
    void (*set_param_sparse)(void *, int n_update, int *idx, double *p);
    /* CasADi part */
    void *ptr_ext_mem;
    int (*casadi_fun)(const double **, double **, int *, double *, void *);
    int (*casadi_work)(int *, int *, int *, int *);
    const int *(*casadi_sparsity_in)(int);
    const int *(*casadi_sparsity_out)(int);
    int (*casadi_n_in)(void);
    int (*casadi_n_out)(void);
    double **args;
    double **res;
    double *w;
    int *iw;
    int *args_size;
    int *res_size;
    int *args_num;
    int *args_size_tot;
    int *res_num;
    int *res_size_tot;
    int in_num;
    int out_num;
    int iw_size;
    int w_size;
    int np;
    double *p;
} external_function_param_casadi;

ACADOS_SYMBOL_EXPORT void external_function_param_casadi_create(external_function_param_casadi *fun, int np);
ACADOS_SYMBOL_EXPORT void external_function_param_casadi_free(external_function_param_casadi *fun);

#ifdef __cplusplus
}
#endif
#endif
