/* acados/utils/print.h -- included by the reference's nodes (bluerov2_dob.h:25) and main_bluerov2.c:39. */
#ifndef BR2_ACADOS_UTILS_PRINT_H_
#define BR2_ACADOS_UTILS_PRINT_H_
#include "acados/utils/types.h"
#include "acados_c/ocp_nlp_interface.h"
#ifdef __cplusplus
extern "C" {
#endif
/* prints x and u of every stage */
ACADOS_SYMBOL_EXPORT void ocp_nlp_out_print(ocp_nlp_dims *dims, ocp_nlp_out *nlp_out);
#ifdef __cplusplus
}
#endif
#endif
