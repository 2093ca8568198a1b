/* acados/ocp_nlp/ocp_nlp_cost_ls.h -- included by bluerov2_dob.h:29; the nodes use no symbol of it. */
#ifndef BR2_ACADOS_OCP_NLP_COST_LS_H_
#define BR2_ACADOS_OCP_NLP_COST_LS_H_
#include "acados/utils/types.h"
#endif
