/* tests/abi/a10_behaviour.c -- behaviour of the generated-ABI functions around the solve, through the same
 * headers a node includes.  Reference semantics: acados_solver_bluerov2.c:788-794 (_update_qp_solver_cond_N: message + exit(1)),
 * :797-830 (_reset: x, u, ... of every stage to zero), :835-844 (_update_params: np != 16 -> message + exit(1)), :886-943
 * (_update_params_sparse), :1030-1036 (_custom_update: message, returns 1).
 *
 *   a10_behaviour cond_N | params_np | sparse_np | custom_update      no CUDA device needed (the misuse paths come first)
 *   a10_behaviour gpu                                                 prints "key value" lines checked by tests/test_abi.py
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "acados_c/ocp_nlp_interface.h"
#include "acados_c/external_function_interface.h"
#include "bluerov2_model/bluerov2_model.h"
#include "acados_solver_bluerov2.h"

#define N 20

static void problem(double *x0, double *p, double *yref)
{
    static const double nominal[12] = {1.7182, 0, 5.468, 0.4006, -11.7391, -20, -31.8678, -5, -18.18, -21.66, -36.99, -1.55};
    for (int i = 0; i < BLUEROV2_NX; i++) x0[i] = 0.0;
    x0[0] = 0.3; x0[1] = -0.2; x0[2] = -19.6; x0[5] = 0.25; x0[6] = 0.1;
    p[0] = 1.5; p[1] = -0.7; p[2] = 0.4; p[3] = 0.05;
    memcpy(p + 4, nominal, sizeof nominal);
    for (int k = 0; k <= N; k++) {
        double *y = yref + (size_t)k * BLUEROV2_NY;
        for (int i = 0; i < BLUEROV2_NY; i++) y[i] = 0.0;
        y[0] = 0.05 * k; y[2] = -20.0; y[6] = 1.0;
    }
}

static bluerov2_solver_capsule *make(void)
{
    double ts[N];
    for (int i = 0; i < N; i++) ts[i] = 1.0 / N;
    bluerov2_solver_capsule *c = bluerov2_acados_create_capsule();
    if (bluerov2_acados_create_with_discretization(c, N, ts)) { fprintf(stderr, "create failed\n"); exit(3); }
    return c;
}

static int tick(bluerov2_solver_capsule *c, const double *x0, const double *yref, double *u0)
{
    ocp_nlp_constraints_model_set(c->nlp_config, c->nlp_dims, c->nlp_in, 0, "lbx", (void *)x0);
    ocp_nlp_constraints_model_set(c->nlp_config, c->nlp_dims, c->nlp_in, 0, "ubx", (void *)x0);
    for (int i = 0; i <= N; i++)
        ocp_nlp_cost_model_set(c->nlp_config, c->nlp_dims, c->nlp_in, i, "yref", (void *)(yref + (size_t)i * BLUEROV2_NY));
    int st = bluerov2_acados_solve(c);
    ocp_nlp_out_get(c->nlp_config, c->nlp_dims, c->nlp_out, 0, "u", u0);
    return st;
}

static double maxabs_iterate(bluerov2_solver_capsule *c)
{
    double m = 0.0, v[BLUEROV2_NX];
    for (int i = 0; i <= N; i++) {
        ocp_nlp_out_get(c->nlp_config, c->nlp_dims, c->nlp_out, i, "x", v);
        for (int j = 0; j < BLUEROV2_NX; j++) m = fmax(m, fabs(v[j]));
        if (i < N) {
            ocp_nlp_out_get(c->nlp_config, c->nlp_dims, c->nlp_out, i, "u", v);
            for (int j = 0; j < BLUEROV2_NU; j++) m = fmax(m, fabs(v[j]));
        }
    }
    return m;
}

int main(int argc, char **argv)
{
    const char *mode = argc > 1 ? argv[1] : "";
    double x0[BLUEROV2_NX], p[BLUEROV2_NP], yref[(N + 1) * BLUEROV2_NY], u[4][BLUEROV2_NU];
    problem(x0, p, yref);
    if (!strcmp(mode, "cond_N")) {
        bluerov2_solver_capsule *c = bluerov2_acados_create_capsule();
        bluerov2_acados_update_qp_solver_cond_N(c, 5);
        printf("returned\n");                               /* must not be reached */
        return 0;
    }
    if (!strcmp(mode, "params_np")) {
        bluerov2_solver_capsule *c = bluerov2_acados_create_capsule();
        bluerov2_acados_update_params(c, 0, p, BLUEROV2_NP - 1);
        printf("returned\n");
        return 0;
    }
    if (!strcmp(mode, "sparse_np")) {
        bluerov2_solver_capsule *c = bluerov2_acados_create_capsule();
        int idx[BLUEROV2_NP + 1] = {0};
        double val[BLUEROV2_NP + 1] = {0};
        bluerov2_acados_update_params_sparse(c, 0, idx, val, BLUEROV2_NP + 1);
        printf("returned\n");
        return 0;
    }
    if (!strcmp(mode, "custom_update")) {
        int rc = bluerov2_acados_custom_update(NULL, NULL, 0);
        printf("custom_update_rc %d\n", rc);
        return 0;
    }
    if (!strcmp(mode, "lbx_ne_ubx") || !strcmp(mode, "scaling_ne_ts")) {
        /* settings acados accepts but this solver cannot represent: refused with a message and exit(1) when the solve starts */
        bluerov2_solver_capsule *c = make();
        double ub[BLUEROV2_NX];
        memcpy(ub, x0, sizeof ub);
        ocp_nlp_constraints_model_set(c->nlp_config, c->nlp_dims, c->nlp_in, 0, "lbx", x0);
        if (!strcmp(mode, "lbx_ne_ubx")) ub[3] += 0.5;
        ocp_nlp_constraints_model_set(c->nlp_config, c->nlp_dims, c->nlp_in, 0, "ubx", ub);
        if (!strcmp(mode, "scaling_ne_ts")) { double sc = 0.123; ocp_nlp_cost_model_set(c->nlp_config, c->nlp_dims, c->nlp_in, 2, "scaling", &sc); }
        bluerov2_acados_solve(c);
        printf("returned\n");
        return 0;
    }
    if (strcmp(mode, "gpu")) { fprintf(stderr, "usage: %s cond_N|params_np|sparse_np|custom_update|gpu\n", argv[0]); return 2; }

    /* ---- A: dense parameters, two ticks; then _reset and a solve from the zeroed iterate ---- */
    bluerov2_solver_capsule *A = make();
    for (int i = 0; i <= N; i++) bluerov2_acados_update_params(A, i, p, BLUEROV2_NP);
    int st = tick(A, x0, yref, u[0]);
    st |= tick(A, x0, yref, u[1]);
    printf("status_A %d\n", st);
    printf("iterate_before_reset %.17g\n", maxabs_iterate(A));
    int rc = bluerov2_acados_reset(A, 1);
    printf("reset_rc %d\n", rc);
    printf("iterate_after_reset %.17g\n", maxabs_iterate(A));
    st = tick(A, x0, yref, u[2]);
    printf("status_after_reset %d\n", st);

    /* ---- B: fresh capsule, iterate zeroed by hand (ocp_nlp_out_set), parameters through _update_params_sparse ---- */
    bluerov2_solver_capsule *B = make();
    double zx[BLUEROV2_NX] = {0}, zu[BLUEROV2_NU] = {0};
    for (int i = 0; i <= N; i++) ocp_nlp_out_set(B->nlp_config, B->nlp_dims, B->nlp_out, i, "x", zx);
    for (int i = 0; i < N; i++) ocp_nlp_out_set(B->nlp_config, B->nlp_dims, B->nlp_out, i, "u", zu);
    int idx_a[10] = {15, 0, 1, 2, 3, 4, 5, 6, 7, 8}, idx_b[6] = {9, 10, 11, 12, 13, 14};
    double val_a[10], val_b[6];
    for (int i = 0; i < 10; i++) val_a[i] = p[idx_a[i]];
    for (int i = 0; i < 6; i++) val_b[i] = p[idx_b[i]];
    for (int i = 0; i <= N; i++) {
        rc |= bluerov2_acados_update_params_sparse(B, i, idx_a, val_a, 10);
        rc |= bluerov2_acados_update_params_sparse(B, i, idx_b, val_b, 6);
    }
    printf("sparse_rc %d\n", rc);
    st = tick(B, x0, yref, u[3]);
    printf("status_B %d\n", st);
    double d_reset = 0.0, d_moved = 0.0;
    for (int j = 0; j < BLUEROV2_NU; j++) {
        d_reset = fmax(d_reset, fabs(u[2][j] - u[3][j]));   /* solve after _reset == cold solve from zeros with sparse-set p */
        d_moved = fmax(d_moved, fabs(u[2][j] - u[1][j]));   /* ... and it is NOT the warm-started answer */
    }
    printf("reset_vs_cold %.17g\n", d_reset);
    printf("reset_vs_warm %.17g\n", d_moved);
    printf("u_after_reset %.17g %.17g %.17g %.17g\n", u[2][0], u[2][1], u[2][2], u[2][3]);
    printf("custom_update_rc %d\n", bluerov2_acados_custom_update(A, NULL, 0));
    rc = bluerov2_acados_free(A) | bluerov2_acados_free_capsule(A) | bluerov2_acados_free(B) | bluerov2_acados_free_capsule(B);
    printf("free_rc %d\n", rc);
    return 0;
}
