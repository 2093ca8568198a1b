/* blasfeo/include/blasfeo_d_aux.h -- included by bluerov2_dob.h:31; the nodes use no symbol of it. */
#ifndef BR2_BLASFEO_D_AUX_H_
#define BR2_BLASFEO_D_AUX_H_
#endif
