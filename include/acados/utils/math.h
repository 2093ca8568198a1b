/* acados/utils/math.h -- main_bluerov2.c:40 includes it for MIN(). */
#ifndef BR2_ACADOS_UTILS_MATH_H_
#define BR2_ACADOS_UTILS_MATH_H_
#include "acados/utils/types.h"
#ifndef MIN
#define MIN(a, b) (((a) < (b)) ? (a) : (b))
#endif
#ifndef MAX
#define MAX(a, b) (((a) > (b)) ? (a) : (b))
#endif
#endif
