/* tests/abi/dob_tick.c -- drives the solver through the acados-generated C-ABI exactly like the reference's node:
 * the per-tick call sequence of BLUEROV2_DOB::solve (bluerov2_dobmpc/src/bluerov2_dob.cpp:307-388) and the one-off
 * construction of bluerov2_dob.h:168 + bluerov2_dob.cpp:28-38, written against the same headers the node includes
 * (bluerov2_dob.h:25-35).  Test harness: reads ticks from a binary file, writes u0 / status / iterate to another.
 *
 * file in : int32 N, int32 T, then T x { double x0[12], double p[16], double yref[(N+1)*16] }
 * file out: T x { int32 status, int32 pad, double u0[4], double time_tot, double inf_norm_res },
 *           then double X[(N+1)*12], double U[N*4]
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "acados/utils/print.h"
#include "acados_c/ocp_nlp_interface.h"
#include "acados_c/external_function_interface.h"
#include "acados/ocp_nlp/ocp_nlp_constraints_bgh.h"
#include "acados/ocp_nlp/ocp_nlp_cost_ls.h"
#include "blasfeo/include/blasfeo_d_aux.h"
#include "blasfeo/include/blasfeo_d_aux_ext_dep.h"
#include "bluerov2_model/bluerov2_model.h"
#include "acados_solver_bluerov2.h"

int main(int argc, char **argv)
{
    if (argc < 3) { fprintf(stderr, "usage: %s in.bin out.bin\n", argv[0]); return 2; }
    FILE *fi = fopen(argv[1], "rb");
    if (!fi) { perror(argv[1]); return 2; }
    int hdr[2];
    if (fread(hdr, sizeof(int), 2, fi) != 2) return 2;
    const int N = hdr[0], T = hdr[1];
    double *yref = (double *)malloc(sizeof(double) * (N + 1) * BLUEROV2_NY);
    double *ts = (double *)malloc(sizeof(double) * N);
    for (int i = 0; i < N; i++) ts[i] = 1.0 / N;

    bluerov2_solver_capsule *capsule = bluerov2_acados_create_capsule();
    int status = (N == BLUEROV2_N) ? bluerov2_acados_create(capsule)
                                   : bluerov2_acados_create_with_discretization(capsule, N, ts);
    if (N == BLUEROV2_N) status |= bluerov2_acados_update_time_steps(capsule, N, ts);
    if (status) { fprintf(stderr, "create failed: %d\n", status); return 1; }
    ocp_nlp_config *cfg = capsule->nlp_config;          /* the nodes reach into the capsule: bluerov2_dob.cpp:320 */
    ocp_nlp_dims *dims = capsule->nlp_dims;
    if (dims->N != N) { fprintf(stderr, "dims->N = %d\n", dims->N); return 1; }

    FILE *fo = fopen(argv[2], "wb");
    if (!fo) { perror(argv[2]); return 2; }
    double x0[BLUEROV2_NX], p[BLUEROV2_NP], u0[BLUEROV2_NU];
    for (int t = 0; t < T; t++) {
        if (fread(x0, sizeof(double), BLUEROV2_NX, fi) != BLUEROV2_NX) return 2;
        if (fread(p, sizeof(double), BLUEROV2_NP, fi) != BLUEROV2_NP) return 2;
        if (fread(yref, sizeof(double), (size_t)(N + 1) * BLUEROV2_NY, fi) != (size_t)(N + 1) * BLUEROV2_NY) return 2;
        if (t == 0) {
            /* warm start X_k = x0, U = 0 (the benchmark's tick-0 iterate) through ocp_nlp_out_set (main_bluerov2.c:213) */
            double zu[BLUEROV2_NU] = {0, 0, 0, 0};
            for (int i = 0; i <= N; i++) ocp_nlp_out_set(cfg, dims, capsule->nlp_out, i, "x", x0);
            for (int i = 0; i < N; i++) ocp_nlp_out_set(cfg, dims, capsule->nlp_out, i, "u", zu);
        }
        ocp_nlp_constraints_model_set(cfg, dims, capsule->nlp_in, 0, "lbx", x0);                      /* :320 */
        ocp_nlp_constraints_model_set(cfg, dims, capsule->nlp_in, 0, "ubx", x0);                      /* :321 */
        for (int i = 0; i < N + 1; i++) bluerov2_acados_update_params(capsule, i, p, BLUEROV2_NP);    /* :354 */
        for (int i = 0; i <= N; i++)
            ocp_nlp_cost_model_set(cfg, dims, capsule->nlp_in, i, "yref", yref + (size_t)i * BLUEROV2_NY);   /* :371 */
        int st = bluerov2_acados_solve(capsule);                                                       /* :375 */
        double kkt = capsule->nlp_out->inf_norm_res, cpu_time = 0;                                     /* :384 */
        ocp_nlp_get(cfg, capsule->nlp_solver, "time_tot", &cpu_time);                                  /* :386 */
        ocp_nlp_out_get(cfg, dims, capsule->nlp_out, 0, "u", (void *)u0);                              /* :388 */
        int sti[2] = {st, 0};
        fwrite(sti, sizeof(int), 2, fo);
        fwrite(u0, sizeof(double), BLUEROV2_NU, fo);
        fwrite(&cpu_time, sizeof(double), 1, fo);
        fwrite(&kkt, sizeof(double), 1, fo);
    }
    double xs[BLUEROV2_NX], us[BLUEROV2_NU];
    for (int i = 0; i <= N; i++) { ocp_nlp_out_get(cfg, dims, capsule->nlp_out, i, "x", xs); fwrite(xs, sizeof(double), BLUEROV2_NX, fo); }
    for (int i = 0; i < N; i++) { ocp_nlp_out_get(cfg, dims, capsule->nlp_out, i, "u", us); fwrite(us, sizeof(double), BLUEROV2_NU, fo); }
    fclose(fo);
    fclose(fi);
    bluerov2_acados_print_stats(capsule);
    status = bluerov2_acados_free(capsule);
    status |= bluerov2_acados_free_capsule(capsule);
    free(yref);
    free(ts);
    return status;
}
