/* bluerov2_constraints/bluerov2_constraints.h -- the reference's generated header declares nothing (only box
 * bounds, no nonlinear constraint functions): c_generated_code/bluerov2_constraints/bluerov2_constraints.h. */
#ifndef BR2_DROPIN_BLUEROV2_CONSTRAINTS_H
#define BR2_DROPIN_BLUEROV2_CONSTRAINTS_H
#endif
