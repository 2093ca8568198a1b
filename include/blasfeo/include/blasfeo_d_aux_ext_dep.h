/* blasfeo/include/blasfeo_d_aux_ext_dep.h -- main_bluerov2.c:45 uses d_print_exp_tran_mat (:230,:232). */
#ifndef BR2_BLASFEO_D_AUX_EXT_DEP_H_
#define BR2_BLASFEO_D_AUX_EXT_DEP_H_
#ifdef __cplusplus
extern "C" {
#endif
/* print the transposed of a column-major row x col matrix A (leading dimension lda), exponential format */
__attribute__((visibility("default"))) void d_print_exp_tran_mat(int row, int col, double *A, int lda);
__attribute__((visibility("default"))) void d_print_exp_mat(int row, int col, double *A, int lda);
__attribute__((visibility("default"))) void d_print_mat(int row, int col, double *A, int lda);
#ifdef __cplusplus
}
#endif
#endif
