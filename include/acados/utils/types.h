/* acados/utils/types.h -- the subset of acados' types the BlueROV2 solver ABI and its callers use.
 * Stands in for the header of the same path in an acados install (the reference includes it from
 * c_generated_code/acados_solver_bluerov2.h:37).  Values of the return codes follow acados. */
#ifndef BR2_ACADOS_UTILS_TYPES_H_
#define BR2_ACADOS_UTILS_TYPES_H_

#include <stdbool.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32) || defined(__CYGWIN__)
#define ACADOS_SYMBOL_EXPORT __declspec(dllexport)
#else
#define ACADOS_SYMBOL_EXPORT __attribute__((visibility("default")))
#endif

typedef double real_t;
typedef int int_t;

#define MAX_STR_LEN 256
#define ACADOS_EPS 1e-12
#define ACADOS_NEG_INFTY (-1.0e9)
#define ACADOS_POS_INFTY (+1.0e9)

enum return_values
{
    ACADOS_SUCCESS = 0,
    ACADOS_NAN_DETECTED = 1,
    ACADOS_MAXITER = 2,
    ACADOS_MINSTEP = 3,
    ACADOS_QP_FAILURE = 4,
    ACADOS_READY = 5,
};

#ifdef __cplusplus
}
#endif
#endif
