/* acados/ocp_nlp/ocp_nlp_constraints_bgh.h -- included by bluerov2_dob.h:28; the nodes use no symbol of it.
 * The BGH constraint module of this OCP reduces to: stage-0 state fixed (lbx = ubx = x0, idxbxe all 12) and
 * box bounds on the 4 inputs; both are set through ocp_nlp_constraints_model_set (acados_c/ocp_nlp_interface.h). */
#ifndef BR2_ACADOS_OCP_NLP_CONSTRAINTS_BGH_H_
#define BR2_ACADOS_OCP_NLP_CONSTRAINTS_BGH_H_
#include "acados/utils/types.h"
#endif
