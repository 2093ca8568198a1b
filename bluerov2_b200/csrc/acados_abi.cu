// acados_abi.cu -- the reference-facing C-ABI: the acados-generated solver interface of the BlueROV2 OCP
// (include/acados_solver_bluerov2.h == c_generated_code/acados_solver_bluerov2.h:129-167) and the slice of the
// acados C interface the reference's callers use (include/acados_c/ocp_nlp_interface.h), served by the batched
// B200 engine with batch = 1.  Host code only; every solve goes to the GPU through br2_batch_solve_host.
//
// What each function mirrors is cited from bluerov2_dobmpc/scripts/c_generated_code/acados_solver_bluerov2.c
// ("gen.c" below) and from the callers (src/bluerov2_dob.cpp, src/ctrller/mpc.cpp, main_bluerov2.c).
// Semantics kept: values copied on set; unknown fields print and exit(1) like acados; update_params with np != 16
// and update_qp_solver_cond_N exit(1) (gen.c:839-844, :788-794); status codes are acados' (0 success, 1 NaN,
// 2 max iter, 4 QP failure).
#include "../../include/acados_solver_bluerov2.h"
#include "../../include/acados/utils/print.h"
#include "../../include/blasfeo/include/blasfeo_d_aux_ext_dep.h"
#include "../../include/bluerov2_b200.h"
#include "../../include/bluerov2_model/bluerov2_model.h"
#include "../../include/bluerov2_cost/bluerov2_cost.h"
#include "model.cuh"

#include <chrono>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

namespace {

constexpr int NX = 12, NU = 4, NP = 16, NY = 16;

// One context shared by the config / dims / in / out / solver objects of a capsule.
struct Ctx {
    br2_batch_solver* eng;
    int N;
    double* Ts;        // [N]
    double* scaling;   // [N] cost scaling per stage as set through ocp_nlp_cost_model_set("scaling"); the engine ties it to Ts
    double* yref;      // [N+1][16]
    double* p;         // [N+1][16]
    double* X;         // [N+1][12] host mirror of the iterate
    double* U;         // [N][4]
    double lbx0[NX], ubx0[NX];
    double W[16], We[12], lbu[NU], ubu[NU];
    bool iterate_dirty;    // host mirror newer than the device copy
    bool weights_dirty, bounds_dirty, ts_dirty;
    int rti_phase, print_level;
    // results of the last solve
    int status, qp_iter, qp_status;
    double time_tot, time_lin, time_qp, res_stat, res_eq, mu;
    double u0[NU], thrust[6];
};

struct Opts { Ctx* ctx; };

[[noreturn]] void die(const char* fn, const char* field)
{
    // acados: "error: <fn>: field <x> not available" then exit(1)
    fprintf(stderr, "\nerror: %s: field '%s' not available in the bluerov2 B200 solver\n", fn, field);
    exit(1);
}

Ctx* ctx_of(void* c, const char* fn)
{
    if (!c) {
        fprintf(stderr, "\nerror: %s: object was not created by bluerov2_acados_create\n", fn);
        exit(1);
    }
    return (Ctx*)c;
}

void check_stage(const char* fn, int stage, int lo, int hi)
{
    if (stage < lo || stage > hi) {
        fprintf(stderr, "\nerror: %s: stage %d outside [%d, %d]\n", fn, stage, lo, hi);
        exit(1);
    }
}

// ---- external_function_param_casadi glue (gen.c:287-349): only the parameter plumbing is live ----
void ef_set_param(void* self, double* p)
{
    external_function_param_casadi* f = (external_function_param_casadi*)self;
    if (f->p && p) memcpy(f->p, p, sizeof(double) * f->np);
}
void ef_set_param_sparse(void* self, int n_update, int* idx, double* p)
{
    external_function_param_casadi* f = (external_function_param_casadi*)self;
    for (int i = 0; i < n_update; i++)
        if (idx[i] >= 0 && idx[i] < f->np) f->p[idx[i]] = p[i];
}
void ef_get_nparam(void* self, int* np) { *np = ((external_function_param_casadi*)self)->np; }

int push_config(Ctx* c)
{
    int rc = BR2_OK;
    if (c->weights_dirty) { rc = br2_batch_set_weights(c->eng, c->W, c->We); c->weights_dirty = false; }
    if (rc == BR2_OK && c->bounds_dirty) { rc = br2_batch_set_bounds(c->eng, c->lbu, c->ubu); c->bounds_dirty = false; }
    if (rc == BR2_OK && c->ts_dirty) { rc = br2_batch_set_time_steps(c->eng, c->Ts); c->ts_dirty = false; }
    return rc;
}

}  // namespace

// =====================================================================================================
// acados C interface slice
// =====================================================================================================
extern "C" {

void external_function_param_casadi_create(external_function_param_casadi* fun, int np)
{
    fun->np = np;
    fun->p = (double*)calloc(np > 0 ? np : 1, sizeof(double));
    fun->set_param = ef_set_param;
    fun->set_param_sparse = ef_set_param_sparse;
    fun->get_nparam = ef_get_nparam;
}
void external_function_param_casadi_free(external_function_param_casadi* fun)
{
    free(fun->p);
    fun->p = nullptr;
}

int ocp_nlp_constraints_model_set(ocp_nlp_config* config, ocp_nlp_dims*, ocp_nlp_in*, int stage, const char* field, void* value)
{
    Ctx* c = ctx_of(config ? config->ctx : nullptr, "ocp_nlp_constraints_model_set");
    check_stage("ocp_nlp_constraints_model_set", stage, 0, c->N);
    if (!strcmp(field, "lbx") || !strcmp(field, "ubx")) {
        // only the stage-0 state is constrained (nbx = 0 on the path, nbx0 = 12 all equalities: gen.c:228,501-543)
        if (stage != 0) die("ocp_nlp_constraints_model_set (stage > 0 has nbx = 0)", field);
        memcpy(field[0] == 'l' ? c->lbx0 : c->ubx0, value, sizeof(double) * NX);
    } else if (!strcmp(field, "idxbx") || !strcmp(field, "idxbxe") || !strcmp(field, "idxbu")) {
        // index sets are fixed by the problem structure: accept the identity (main_bluerov2.c:144), reject anything else
        const int n = (field[4] == 'u') ? NU : NX;
        const int* idx = (const int*)value;
        for (int i = 0; i < n; i++)
            if (idx[i] != i) {
                fprintf(stderr, "\nerror: ocp_nlp_constraints_model_set: %s must be the identity for this OCP\n", field);
                exit(1);
            }
    } else if (!strcmp(field, "lbu") || !strcmp(field, "ubu")) {
        if (stage >= c->N) die("ocp_nlp_constraints_model_set (terminal stage has no inputs)", field);
        // one box for all stages (gen.c:547-571 sets the same lbu/ubu on every stage)
        memcpy(field[0] == 'l' ? c->lbu : c->ubu, value, sizeof(double) * NU);
        c->bounds_dirty = true;
    } else {
        die("ocp_nlp_constraints_model_set", field);
    }
    return ACADOS_SUCCESS;
}

int ocp_nlp_cost_model_set(ocp_nlp_config* config, ocp_nlp_dims*, ocp_nlp_in*, int stage, const char* field, void* value)
{
    Ctx* c = ctx_of(config ? config->ctx : nullptr, "ocp_nlp_cost_model_set");
    check_stage("ocp_nlp_cost_model_set", stage, 0, c->N);
    if (!strcmp(field, "yref") || !strcmp(field, "y_ref")) {
        // stage < N: ny = 16; terminal: ny_e = 12 (bluerov2_dob.cpp:370-372 passes a 16-row for every stage)
        memcpy(c->yref + (size_t)stage * NY, value, sizeof(double) * (stage < c->N ? NY : NX));
    } else if (!strcmp(field, "W")) {
        // column-major ny x ny; this OCP's W is diagonal (gen.c:424-479) and shared by stages 0..N-1
        const double* Wm = (const double*)value;
        const int n = stage < c->N ? NY : NX;
        for (int i = 0; i < n; i++)
            for (int j = 0; j < n; j++)
                if (i != j && Wm[i + n * j] != 0.0) {
                    fprintf(stderr, "\nerror: ocp_nlp_cost_model_set: W must be diagonal for the bluerov2 B200 solver\n");
                    exit(1);
                }
        for (int i = 0; i < n; i++) (stage < c->N ? c->W : c->We)[i] = Wm[i + n * i];
        c->weights_dirty = true;
    } else if (!strcmp(field, "scaling")) {
        // cost scaling == time step on stages 0..N-1 (gen.c:389-394, :126-129); the engine ties the two together
        // -- kept separately and checked against Ts when the solve starts: a caller that sets the two apart gets a diagnostic
        if (stage < c->N) c->scaling[stage] = *(const double*)value;
    } else {
        die("ocp_nlp_cost_model_set", field);
    }
    return ACADOS_SUCCESS;
}

int ocp_nlp_in_set(ocp_nlp_config* config, ocp_nlp_dims*, ocp_nlp_in*, int stage, const char* field, void* value)
{
    Ctx* c = ctx_of(config ? config->ctx : nullptr, "ocp_nlp_in_set");
    if (!strcmp(field, "Ts")) {
        check_stage("ocp_nlp_in_set", stage, 0, c->N - 1);
        c->Ts[stage] = *(const double*)value;
        c->ts_dirty = true;
    } else {
        die("ocp_nlp_in_set", field);
    }
    return ACADOS_SUCCESS;
}

void ocp_nlp_out_set(ocp_nlp_config* config, ocp_nlp_dims*, ocp_nlp_out*, int stage, const char* field, void* value)
{
    Ctx* c = ctx_of(config ? config->ctx : nullptr, "ocp_nlp_out_set");
    check_stage("ocp_nlp_out_set", stage, 0, c->N);
    if (!strcmp(field, "x")) {
        memcpy(c->X + (size_t)stage * NX, value, sizeof(double) * NX);
        c->iterate_dirty = true;
    } else if (!strcmp(field, "u")) {
        if (stage < c->N) {   // gen.c:812-816 also "sets" u at stage N (nu_N = 0: nothing is copied)
            memcpy(c->U + (size_t)stage * NU, value, sizeof(double) * NU);
            c->iterate_dirty = true;
        }
    } else if (!strcmp(field, "sl") || !strcmp(field, "su") || !strcmp(field, "lam") || !strcmp(field, "t") ||
               !strcmp(field, "z") || !strcmp(field, "pi")) {
        // multipliers / slacks are not carried between RTI steps (qp_solver_warm_start 0): nothing to store
    } else {
        die("ocp_nlp_out_set", field);
    }
}

void ocp_nlp_out_get(ocp_nlp_config* config, ocp_nlp_dims*, ocp_nlp_out* out, int stage, const char* field, void* value)
{
    Ctx* c = ctx_of(config ? config->ctx : nullptr, "ocp_nlp_out_get");
    check_stage("ocp_nlp_out_get", stage, 0, c->N);
    if (!strcmp(field, "x")) {
        memcpy(value, c->X + (size_t)stage * NX, sizeof(double) * NX);
    } else if (!strcmp(field, "u")) {
        if (stage < c->N) memcpy(value, c->U + (size_t)stage * NU, sizeof(double) * NU);
    } else if (!strcmp(field, "kkt_norm_inf")) {
        *(double*)value = out ? out->inf_norm_res : fmax(c->res_stat, c->res_eq);
    } else {
        die("ocp_nlp_out_get", field);
    }
}

void ocp_nlp_get(ocp_nlp_config* config, ocp_nlp_solver*, const char* field, void* ret)
{
    Ctx* c = ctx_of(config ? config->ctx : nullptr, "ocp_nlp_get");
    if (!strcmp(field, "time_tot")) *(double*)ret = c->time_tot;
    else if (!strcmp(field, "time_lin") || !strcmp(field, "time_sim")) *(double*)ret = c->time_lin;
    else if (!strcmp(field, "time_qp") || !strcmp(field, "time_qp_sol")) *(double*)ret = c->time_qp;
    else if (!strcmp(field, "sqp_iter")) *(int*)ret = 1;          // SQP_RTI: one iteration per call
    else if (!strcmp(field, "qp_iter")) *(int*)ret = c->qp_iter;
    else if (!strcmp(field, "status")) *(int*)ret = c->status;
    else if (!strcmp(field, "stat_n")) *(int*)ret = 2;            // columns qp_stat, qp_iter (gen.c:1019)
    else if (!strcmp(field, "stat_m")) *(int*)ret = 2;
    else if (!strcmp(field, "statistics")) {
        // (stat_n + 1) columns x nrow = min(sqp_iter + 1, stat_m) = 2 rows, column-major (gen.c:1011-1027)
        double* s = (double*)ret;
        const int nrow = 2;
        s[0 + 0 * nrow] = 0; s[1 + 0 * nrow] = 1;
        s[0 + 1 * nrow] = 0; s[1 + 1 * nrow] = c->qp_status;
        s[0 + 2 * nrow] = 0; s[1 + 2 * nrow] = c->qp_iter;
    } else {
        die("ocp_nlp_get", field);
    }
}

void ocp_nlp_solver_opts_set(ocp_nlp_config* config, void* opts_, const char* field, void* value)
{
    Ctx* c = ctx_of(config ? config->ctx : (opts_ ? ((Opts*)opts_)->ctx : nullptr), "ocp_nlp_solver_opts_set");
    if (!strcmp(field, "rti_phase")) {
        const int ph = *(const int*)value;
        if (ph < 0 || ph > 2) {
            fprintf(stderr, "\nerror: ocp_nlp_solver_opts_set: invalid value for rti_phase field, must be in [0, 2], got %d\n", ph);
            exit(1);
        }
        c->rti_phase = ph;
    } else if (!strcmp(field, "qp_iter_max") || !strcmp(field, "qp_solver_iter_max")) {
        if (br2_batch_set_option_int(c->eng, "qp_iter_max", *(const int*)value) != BR2_OK) {
            fprintf(stderr, "\nerror: ocp_nlp_solver_opts_set: %s\n", br2_last_error());
            exit(1);
        }
    } else if (!strcmp(field, "qp_tol") || !strcmp(field, "qp_tol_stat") || !strcmp(field, "qp_tol_comp")) {
        if (br2_batch_set_option_double(c->eng, "qp_tol", *(const double*)value) != BR2_OK) {
            fprintf(stderr, "\nerror: ocp_nlp_solver_opts_set: %s\n", br2_last_error());
            exit(1);
        }
    } else if (!strcmp(field, "print_level")) {
        c->print_level = *(const int*)value;
    } else if (!strcmp(field, "globalization") || !strcmp(field, "full_step_dual") || !strcmp(field, "step_length") ||
               !strcmp(field, "levenberg_marquardt") || !strcmp(field, "qp_hpipm_mode") || !strcmp(field, "qp_warm_start") ||
               !strcmp(field, "qp_tol_eq") || !strcmp(field, "qp_tol_ineq") || !strcmp(field, "ext_cost_num_hess") ||
               !strcmp(field, "exact_hess") || !strcmp(field, "qp_cond_N")) {
        // baked: fixed full step, Gauss-Newton, cold-started IPM (gen.c:611-675)
    } else {
        die("ocp_nlp_solver_opts_set", field);
    }
}

int ocp_nlp_precompute(ocp_nlp_solver* solver, ocp_nlp_in*, ocp_nlp_out*)
{
    Ctx* c = ctx_of(solver ? solver->ctx : nullptr, "ocp_nlp_precompute");
    return push_config(c) == BR2_OK ? ACADOS_SUCCESS : ACADOS_QP_FAILURE;
}

int ocp_nlp_solve(ocp_nlp_solver* solver, ocp_nlp_in*, ocp_nlp_out* out)
{
    Ctx* c = ctx_of(solver ? solver->ctx : nullptr, "ocp_nlp_solve");
    if (c->rti_phase == 1) {
        // PREPARATION only: linearisation needs nothing that arrives later, but the engine fuses it in front of
        // the feedback phase; report READY and do the work when the feedback phase is requested.
        c->status = ACADOS_READY;
        return c->status;
    }
    // what this solver cannot represent is refused loudly, like every other unsupported setting of this ABI:
    // the stage-0 state bounds must be equalities (lbx_0 == ubx_0 = x0: all 12 are flagged idxbxe, gen.c:501-543) ...
    for (int i = 0; i < NX; i++)
        if (!(c->lbx0[i] == c->ubx0[i])) {
            fprintf(stderr, "\nerror: bluerov2_acados_solve: lbx_0[%d] = %g differs from ubx_0[%d] = %g; the bluerov2 B200 solver fixes the "
                            "initial state (all stage-0 state bounds are equalities in the generated problem)\n", i, c->lbx0[i], i, c->ubx0[i]);
            exit(1);
        }
    // ... and the cost scaling of a stage is its time step (acados_solver_bluerov2.c:126-129, 389-394)
    for (int k = 0; k < c->N; k++)
        if (!(c->scaling[k] == c->Ts[k])) {
            fprintf(stderr, "\nerror: bluerov2_acados_solve: cost scaling %g of stage %d differs from its time step %g; the bluerov2 B200 "
                            "solver ties the two together\n", c->scaling[k], k, c->Ts[k]);
            exit(1);
        }
    const auto t0 = std::chrono::steady_clock::now();
    int rc = push_config(c);
    if (rc == BR2_OK && c->iterate_dirty) {
        rc = br2_batch_set_iterate_host(c->eng, c->X, c->U);
        c->iterate_dirty = false;
    }
    int status = 0;
    // x0 = lbx = ubx (all 12 stage-0 bounds are equalities, idxbxe: gen.c:501-543)
    if (rc == BR2_OK) rc = br2_batch_solve_host(c->eng, c->lbx0, c->yref, c->p, 1, c->u0, c->thrust, &status);
    if (rc == BR2_OK) rc = br2_batch_get_iterate_host(c->eng, c->X, c->U);
    double info[4] = {0, 0, 0, 0};
    int iters = 0;
    if (rc == BR2_OK) rc = br2_batch_get_stats_host(c->eng, &iters, info);
    if (rc != BR2_OK) {
        fprintf(stderr, "bluerov2_acados_solve: %s\n", br2_last_error());
        c->status = ACADOS_QP_FAILURE;
        return c->status;
    }
    br2_batch_last_kernel_times(c->eng, &c->time_lin, &c->time_qp);
    c->status = status;
    c->qp_iter = iters;
    c->qp_status = (status == ACADOS_MAXITER) ? 1 : (status == 0 ? 0 : 3);
    c->mu = info[0];
    c->res_stat = info[1];
    c->res_eq = info[2];
    c->time_tot = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (out) {
        out->sqp_iter = 1;
        out->qp_iter = iters;
        // residual of the NLP at the linearisation point: dynamics gap, and the QP's stationarity/complementarity
        out->inf_norm_res = fmax(fmax(c->res_stat, c->res_eq), c->mu);
        out->total_time = c->time_tot;
    }
    if (c->print_level > 0)
        printf("bluerov2 B200 RTI: status %d, ipm iterations %d, mu %.2e, res_stat %.2e, |b|inf %.2e, %.3f ms\n", status,
               iters, c->mu, c->res_stat, c->res_eq, c->time_tot * 1e3);
    return status;
}

void ocp_nlp_out_print(ocp_nlp_dims* dims, ocp_nlp_out* out)
{
    Ctx* c = ctx_of(out ? out->ctx : (dims ? dims->ctx : nullptr), "ocp_nlp_out_print");
    for (int k = 0; k <= c->N; k++) {
        printf("stage %d\nx =", k);
        for (int i = 0; i < NX; i++) printf(" %e", c->X[k * NX + i]);
        if (k < c->N) {
            printf("\nu =");
            for (int i = 0; i < NU; i++) printf(" %e", c->U[k * NU + i]);
        }
        printf("\n");
    }
}

// blasfeo auxiliary printers (main_bluerov2.c:230,232)
void d_print_exp_tran_mat(int row, int col, double* A, int lda)
{
    for (int j = 0; j < col; j++) {
        for (int i = 0; i < row; i++) printf("%e\t", A[i + lda * j]);
        printf("\n");
    }
    printf("\n");
}
void d_print_exp_mat(int row, int col, double* A, int lda)
{
    for (int i = 0; i < row; i++) {
        for (int j = 0; j < col; j++) printf("%e\t", A[i + lda * j]);
        printf("\n");
    }
    printf("\n");
}
void d_print_mat(int row, int col, double* A, int lda)
{
    for (int i = 0; i < row; i++) {
        for (int j = 0; j < col; j++) printf("%9.5f ", A[i + lda * j]);
        printf("\n");
    }
    printf("\n");
}

// =====================================================================================================
// generated-solver ABI
// =====================================================================================================
bluerov2_solver_capsule* bluerov2_acados_create_capsule(void)
{
    // gen.c:87-93: plain malloc, no CUDA context needed (called from a member initialiser, bluerov2_dob.h:168)
    bluerov2_solver_capsule* capsule = (bluerov2_solver_capsule*)calloc(1, sizeof(bluerov2_solver_capsule));
    return capsule;
}

int bluerov2_acados_free_capsule(bluerov2_solver_capsule* capsule)
{
    free(capsule);
    return 0;
}

int bluerov2_acados_create(bluerov2_solver_capsule* capsule)
{
    return bluerov2_acados_create_with_discretization(capsule, BLUEROV2_N, nullptr);
}

int bluerov2_acados_update_time_steps(bluerov2_solver_capsule* capsule, int N, double* new_time_steps)
{
    if (N != capsule->nlp_solver_plan->N) {
        fprintf(stderr,
                "bluerov2_acados_update_time_steps: given number of time steps (= %d) differs from the currently allocated "
                "number of time steps (= %d)!\nPlease recreate with new discretization and provide a new vector of time_stamps!\n",
                N, capsule->nlp_solver_plan->N);
        return 1;
    }
    for (int i = 0; i < N; i++) {
        ocp_nlp_in_set(capsule->nlp_config, capsule->nlp_dims, capsule->nlp_in, i, "Ts", &new_time_steps[i]);
        ocp_nlp_cost_model_set(capsule->nlp_config, capsule->nlp_dims, capsule->nlp_in, i, "scaling", &new_time_steps[i]);
    }
    return 0;
}

int bluerov2_acados_create_with_discretization(bluerov2_solver_capsule* capsule, int N, double* new_time_steps)
{
    if (N != BLUEROV2_N && !new_time_steps) {
        fprintf(stderr,
                "bluerov2_acados_create_with_discretization: new_time_steps is NULL but the number of shooting intervals (= %d) "
                "differs from the number of shooting intervals (= %d) during code generation! Please provide a new vector of "
                "time_stamps!\n", N, BLUEROV2_N);
        return 1;
    }
    capsule->nlp_np = NP;
    int device = 0;
    if (const char* e = getenv("BLUEROV2_CUDA_DEVICE")) device = atoi(e);

    Ctx* c = (Ctx*)calloc(1, sizeof(Ctx));
    c->N = N;
    c->Ts = (double*)calloc(N, sizeof(double));
    c->scaling = (double*)calloc(N, sizeof(double));
    c->yref = (double*)calloc((size_t)(N + 1) * NY, sizeof(double));      // yref defaults to zero (gen.c:403-421)
    c->p = (double*)calloc((size_t)(N + 1) * NP, sizeof(double));         // parameters default to zero (gen.c:355-364)
    c->X = (double*)calloc((size_t)(N + 1) * NX, sizeof(double));
    c->U = (double*)calloc((size_t)N * NU, sizeof(double));
    br2_ocp_defaults dflt;                                  // the one table of baked problem data (include/bluerov2_b200.h)
    br2_get_ocp_defaults(&dflt);
    for (int k = 0; k < N; k++) c->scaling[k] = c->Ts[k] = new_time_steps ? new_time_steps[k] : dflt.Tf / dflt.N;   // gen.c:389-394 (0.0125)
    memcpy(c->W, dflt.W, sizeof c->W);                                                     // gen.c:424-459
    memcpy(c->We, dflt.We, sizeof c->We);                                                  // gen.c:468-479
    memcpy(c->lbu, dflt.lbu, sizeof c->lbu); memcpy(c->ubu, dflt.ubu, sizeof c->ubu);      // gen.c:547-571
    memcpy(c->lbx0, dflt.x_init, sizeof c->lbx0); memcpy(c->ubx0, dflt.x_init, sizeof c->ubx0);   // gen.c:522-541
    for (int k = 0; k <= N; k++) memcpy(c->X + k * NX, dflt.x_init, sizeof dflt.x_init);  // gen.c:681-708
    c->iterate_dirty = true;

    // the engine needs a CUDA device; like a failed ocp_nlp_precompute (gen.c:725-728) a failure here is fatal
    if (br2_batch_create(&c->eng, 1, N, c->Ts, device) != BR2_OK) {
        fprintf(stderr, "\nbluerov2_acados_create: %s\n\n", br2_last_error());
        exit(1);
    }

    // 1) plan
    ocp_nlp_plan_t* plan = (ocp_nlp_plan_t*)calloc(1, sizeof(ocp_nlp_plan_t));
    plan->N = N;
    plan->nlp_solver = SQP_RTI;
    plan->regularization = NO_REGULARIZE;
    plan->ocp_qp_solver_plan.qp_solver = RICCATI_IPM_B200;
    plan->nlp_cost = (ocp_nlp_cost_t*)calloc(N + 1, sizeof(ocp_nlp_cost_t));
    plan->nlp_dynamics = (ocp_nlp_dynamics_t*)calloc(N, sizeof(ocp_nlp_dynamics_t));
    plan->nlp_constraints = (ocp_nlp_constraints_t*)calloc(N + 1, sizeof(ocp_nlp_constraints_t));
    plan->sim_solver_plan = (sim_solver_plan_t*)calloc(N, sizeof(sim_solver_plan_t));
    for (int k = 0; k <= N; k++) { plan->nlp_cost[k] = NONLINEAR_LS; plan->nlp_constraints[k] = BGH; }
    for (int k = 0; k < N; k++) { plan->nlp_dynamics[k] = CONTINUOUS_MODEL; plan->sim_solver_plan[k].sim_solver = ERK; }
    capsule->nlp_solver_plan = plan;
    // 2) config, dims
    ocp_nlp_config* cfg = (ocp_nlp_config*)calloc(1, sizeof(ocp_nlp_config));
    cfg->N = N; cfg->ctx = c;
    capsule->nlp_config = cfg;
    ocp_nlp_dims* dims = (ocp_nlp_dims*)calloc(1, sizeof(ocp_nlp_dims));
    dims->N = N; dims->ctx = c;
    int* dmem = (int*)calloc((size_t)6 * (N + 1), sizeof(int));
    dims->nv = dmem; dims->nx = dmem + (N + 1); dims->nu = dmem + 2 * (N + 1); dims->ni = dmem + 3 * (N + 1);
    dims->nz = dmem + 4 * (N + 1); dims->ns = dmem + 5 * (N + 1);
    for (int k = 0; k <= N; k++) {
        dims->nx[k] = NX; dims->nu[k] = k < N ? NU : 0; dims->nv[k] = dims->nx[k] + dims->nu[k];
        dims->ni[k] = k == 0 ? NX + NU : (k < N ? NU : 0);
    }
    capsule->nlp_dims = dims;
    // 3) external functions: parameter plumbing only (the model is hand-written CUDA, csrc/model.cuh)
    capsule->forw_vde_casadi = (external_function_param_casadi*)calloc(N, sizeof(external_function_param_casadi));
    capsule->expl_ode_fun = (external_function_param_casadi*)calloc(N, sizeof(external_function_param_casadi));
    for (int k = 0; k < N; k++) {
        external_function_param_casadi_create(&capsule->forw_vde_casadi[k], NP);
        external_function_param_casadi_create(&capsule->expl_ode_fun[k], NP);
    }
    const int nm = N > 1 ? N - 1 : 1;
    capsule->cost_y_fun = (external_function_param_casadi*)calloc(nm, sizeof(external_function_param_casadi));
    capsule->cost_y_fun_jac_ut_xt = (external_function_param_casadi*)calloc(nm, sizeof(external_function_param_casadi));
    capsule->cost_y_hess = (external_function_param_casadi*)calloc(nm, sizeof(external_function_param_casadi));
    for (int k = 0; k < N - 1; k++) {
        external_function_param_casadi_create(&capsule->cost_y_fun[k], NP);
        external_function_param_casadi_create(&capsule->cost_y_fun_jac_ut_xt[k], NP);
        external_function_param_casadi_create(&capsule->cost_y_hess[k], NP);
    }
    external_function_param_casadi* single[] = {&capsule->cost_y_0_fun, &capsule->cost_y_0_fun_jac_ut_xt, &capsule->cost_y_0_hess,
                                                &capsule->cost_y_e_fun, &capsule->cost_y_e_fun_jac_ut_xt, &capsule->cost_y_e_hess};
    for (external_function_param_casadi* f : single) external_function_param_casadi_create(f, NP);
    // 5) nlp_in, 6) opts, 7) nlp_out / sens_out, 8) solver
    ocp_nlp_in* in = (ocp_nlp_in*)calloc(1, sizeof(ocp_nlp_in));
    in->Ts = c->Ts; in->ctx = c;
    capsule->nlp_in = in;
    Opts* opts = (Opts*)calloc(1, sizeof(Opts));
    opts->ctx = c;
    capsule->nlp_opts = opts;
    for (ocp_nlp_out** o : {&capsule->nlp_out, &capsule->sens_out}) {
        *o = (ocp_nlp_out*)calloc(1, sizeof(ocp_nlp_out));
        (*o)->x = c->X; (*o)->u = c->U; (*o)->ctx = c;
    }
    ocp_nlp_solver* solver = (ocp_nlp_solver*)calloc(1, sizeof(ocp_nlp_solver));
    solver->config = cfg; solver->dims = dims; solver->opts = opts; solver->ctx = c;
    capsule->nlp_solver = solver;
    // 9) precompute
    int status = ocp_nlp_precompute(solver, in, capsule->nlp_out);
    if (status != ACADOS_SUCCESS) {
        printf("\nocp_nlp_precompute failed!\n\n");
        exit(1);
    }
    return status;
}

int bluerov2_acados_update_qp_solver_cond_N(bluerov2_solver_capsule*, int)
{
    printf("\nacados_update_qp_solver_cond_N() failed, since no partial condensing solver is used!\n\n");
    exit(1);
    return -1;
}

int bluerov2_acados_reset(bluerov2_solver_capsule* capsule, int)
{
    // gen.c:797-830: x, u (and sl, su, lam, t, z, pi) of every stage to zero
    Ctx* c = ctx_of(capsule->nlp_config->ctx, "bluerov2_acados_reset");
    memset(c->X, 0, sizeof(double) * (size_t)(c->N + 1) * NX);
    memset(c->U, 0, sizeof(double) * (size_t)c->N * NU);
    c->iterate_dirty = true;
    return 0;
}

int bluerov2_acados_update_params(bluerov2_solver_capsule* capsule, int stage, double* p, int np)
{
    const int casadi_np = 16;
    if (casadi_np != np) {
        printf("acados_update_params: trying to set %i parameters for external functions."
               " External function has %i parameters. Exiting.\n", np, casadi_np);
        exit(1);
    }
    Ctx* c = ctx_of(capsule->nlp_config->ctx, "bluerov2_acados_update_params");
    const int N = c->N;
    if (stage < N && stage >= 0) {
        capsule->forw_vde_casadi[stage].set_param(capsule->forw_vde_casadi + stage, p);
        capsule->expl_ode_fun[stage].set_param(capsule->expl_ode_fun + stage, p);
        if (stage == 0) {
            capsule->cost_y_0_fun.set_param(&capsule->cost_y_0_fun, p);
            capsule->cost_y_0_fun_jac_ut_xt.set_param(&capsule->cost_y_0_fun_jac_ut_xt, p);
            capsule->cost_y_0_hess.set_param(&capsule->cost_y_0_hess, p);
        } else {
            capsule->cost_y_fun[stage - 1].set_param(capsule->cost_y_fun + stage - 1, p);
            capsule->cost_y_fun_jac_ut_xt[stage - 1].set_param(capsule->cost_y_fun_jac_ut_xt + stage - 1, p);
            capsule->cost_y_hess[stage - 1].set_param(capsule->cost_y_hess + stage - 1, p);
        }
        memcpy(c->p + (size_t)stage * NP, p, sizeof(double) * NP);
    } else {
        // gen.c:869-878: every other stage index lands on the terminal node (cost only; the cost ignores p)
        capsule->cost_y_e_fun.set_param(&capsule->cost_y_e_fun, p);
        capsule->cost_y_e_fun_jac_ut_xt.set_param(&capsule->cost_y_e_fun_jac_ut_xt, p);
        capsule->cost_y_e_hess.set_param(&capsule->cost_y_e_hess, p);
        memcpy(c->p + (size_t)N * NP, p, sizeof(double) * NP);
    }
    return 0;
}

int bluerov2_acados_update_params_sparse(bluerov2_solver_capsule* capsule, int stage, int* idx, double* p, int n_update)
{
    const int casadi_np = 16;
    if (casadi_np < n_update) {
        printf("bluerov2_acados_update_params_sparse: trying to set %d parameters for external functions."
               " External function has %d parameters. Exiting.\n", n_update, casadi_np);
        exit(1);
    }
    Ctx* c = ctx_of(capsule->nlp_config->ctx, "bluerov2_acados_update_params_sparse");
    const int N = c->N;
    const int row = (stage < N && stage >= 0) ? stage : N;
    for (int i = 0; i < n_update; i++)
        if (idx[i] >= 0 && idx[i] < NP) c->p[(size_t)row * NP + idx[i]] = p[i];
    if (row < N) {
        capsule->forw_vde_casadi[row].set_param_sparse(capsule->forw_vde_casadi + row, n_update, idx, p);
        capsule->expl_ode_fun[row].set_param_sparse(capsule->expl_ode_fun + row, n_update, idx, p);
    }
    return 0;
}

int bluerov2_acados_solve(bluerov2_solver_capsule* capsule)
{
    return ocp_nlp_solve(capsule->nlp_solver, capsule->nlp_in, capsule->nlp_out);
}

int bluerov2_acados_free(bluerov2_solver_capsule* capsule)
{
    if (!capsule || !capsule->nlp_config) return 0;
    Ctx* c = (Ctx*)capsule->nlp_config->ctx;
    const int N = c->N;
    for (int k = 0; k < N; k++) {
        external_function_param_casadi_free(&capsule->forw_vde_casadi[k]);
        external_function_param_casadi_free(&capsule->expl_ode_fun[k]);
    }
    free(capsule->forw_vde_casadi);
    free(capsule->expl_ode_fun);
    for (int k = 0; k < N - 1; k++) {
        external_function_param_casadi_free(&capsule->cost_y_fun[k]);
        external_function_param_casadi_free(&capsule->cost_y_fun_jac_ut_xt[k]);
        external_function_param_casadi_free(&capsule->cost_y_hess[k]);
    }
    free(capsule->cost_y_fun);
    free(capsule->cost_y_fun_jac_ut_xt);
    free(capsule->cost_y_hess);
    external_function_param_casadi* single[] = {&capsule->cost_y_0_fun, &capsule->cost_y_0_fun_jac_ut_xt, &capsule->cost_y_0_hess,
                                                &capsule->cost_y_e_fun, &capsule->cost_y_e_fun_jac_ut_xt, &capsule->cost_y_e_hess};
    for (external_function_param_casadi* f : single) external_function_param_casadi_free(f);
    free(capsule->nlp_opts);
    free(capsule->nlp_in);
    free(capsule->nlp_out);
    free(capsule->sens_out);
    free(capsule->nlp_solver);
    free(capsule->nlp_dims->nv);
    free(capsule->nlp_dims);
    free(capsule->nlp_config);
    ocp_nlp_plan_t* plan = capsule->nlp_solver_plan;
    free(plan->nlp_cost); free(plan->nlp_dynamics); free(plan->nlp_constraints); free(plan->sim_solver_plan);
    free(plan);
    br2_batch_free(c->eng);
    free(c->Ts); free(c->scaling); free(c->yref); free(c->p); free(c->X); free(c->U);
    free(c);
    memset(capsule, 0, sizeof *capsule);
    return 0;
}

void bluerov2_acados_print_stats(bluerov2_solver_capsule* capsule)
{
    int sqp_iter, stat_m, stat_n, tmp_int;
    ocp_nlp_get(capsule->nlp_config, capsule->nlp_solver, "sqp_iter", &sqp_iter);
    ocp_nlp_get(capsule->nlp_config, capsule->nlp_solver, "stat_n", &stat_n);
    ocp_nlp_get(capsule->nlp_config, capsule->nlp_solver, "stat_m", &stat_m);
    double stat[1200];
    ocp_nlp_get(capsule->nlp_config, capsule->nlp_solver, "statistics", stat);
    int nrow = sqp_iter + 1 < stat_m ? sqp_iter + 1 : stat_m;
    printf("iter\tres_stat\tres_eq\t\tres_ineq\tres_comp\tqp_stat\tqp_iter\talpha");
    if (stat_n > 8) printf("\t\tqp_res_stat\tqp_res_eq\tqp_res_ineq\tqp_res_comp");
    printf("\n");
    printf("iter\tqp_stat\tqp_iter\n");
    for (int i = 0; i < nrow; i++) {
        for (int j = 0; j < stat_n + 1; j++) {
            tmp_int = (int)stat[i + j * nrow];
            printf("%d\t", tmp_int);
        }
        printf("\n");
    }
}

int bluerov2_acados_custom_update(bluerov2_solver_capsule*, double*, int)
{
    printf("\ndummy function that can be called in between solver calls to update parameters or numerical data efficiently in C.\n");
    printf("nothing set yet..\n");
    return 1;
}

ocp_nlp_in* bluerov2_acados_get_nlp_in(bluerov2_solver_capsule* capsule) { return capsule->nlp_in; }
ocp_nlp_out* bluerov2_acados_get_nlp_out(bluerov2_solver_capsule* capsule) { return capsule->nlp_out; }
ocp_nlp_out* bluerov2_acados_get_sens_out(bluerov2_solver_capsule* capsule) { return capsule->sens_out; }
ocp_nlp_solver* bluerov2_acados_get_nlp_solver(bluerov2_solver_capsule* capsule) { return capsule->nlp_solver; }
ocp_nlp_config* bluerov2_acados_get_nlp_config(bluerov2_solver_capsule* capsule) { return capsule->nlp_config; }
void* bluerov2_acados_get_nlp_opts(bluerov2_solver_capsule* capsule) { return capsule->nlp_opts; }
ocp_nlp_dims* bluerov2_acados_get_nlp_dims(bluerov2_solver_capsule* capsule) { return capsule->nlp_dims; }
ocp_nlp_plan_t* bluerov2_acados_get_nlp_plan(bluerov2_solver_capsule* capsule) { return capsule->nlp_solver_plan; }

// =====================================================================================================
// model / cost functions with the CasADi calling convention (include/bluerov2_model/bluerov2_model.h).
// Exported because the reference's libacados_ocp_solver_bluerov2.so exports them (c_generated_code/Makefile
// links the model objects into it); the engine never calls them -- its dynamics are the same model.cuh code
// compiled for the device.  They are NOT a CPU solve path: there is none.
// =====================================================================================================
#define BR2_MODEL_API __attribute__((visibility("default")))

static const int kSpX[] = {12, 1, 0, 12, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11};
static const int kSpU[] = {4, 1, 0, 4, 0, 1, 2, 3};
static const int kSpP[] = {16, 1, 0, 16, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15};
static int kSpSx[3 + 12 + 144], kSpSu[3 + 4 + 48];
static const int* dense_sp(int* buf, int rows, int cols)
{
    if (buf[0] == 0) {
        buf[1] = cols;
        for (int j = 0; j <= cols; j++) buf[2 + j] = j * rows;
        for (int j = 0; j < cols; j++)
            for (int i = 0; i < rows; i++) buf[3 + cols + j * rows + i] = i;
        buf[0] = rows;
    }
    return buf;
}

BR2_MODEL_API int bluerov2_expl_ode_fun(const real_t** arg, real_t** res, int*, real_t*, void*)
{
    br2::ModelConst mc; mc.set(arg[2]);
    br2::Trig t; br2::trig_of(arg[0], t);
    if (res[0]) br2::ode(arg[0], arg[1], mc, t, res[0]);
    return 0;
}
BR2_MODEL_API int bluerov2_expl_ode_fun_work(int* a, int* b, int* c, int* d) { if (a) *a = 3; if (b) *b = 1; if (c) *c = 0; if (d) *d = 0; return 0; }
BR2_MODEL_API const int* bluerov2_expl_ode_fun_sparsity_in(int i) { return i == 0 ? kSpX : i == 1 ? kSpU : i == 2 ? kSpP : nullptr; }
BR2_MODEL_API const int* bluerov2_expl_ode_fun_sparsity_out(int i) { return i == 0 ? kSpX : nullptr; }
BR2_MODEL_API int bluerov2_expl_ode_fun_n_in(void) { return 3; }
BR2_MODEL_API int bluerov2_expl_ode_fun_n_out(void) { return 1; }

BR2_MODEL_API int bluerov2_expl_vde_forw(const real_t** arg, real_t** res, int*, real_t*, void*)
{
    const double *x = arg[0], *Sx = arg[1], *Su = arg[2], *u = arg[3], *p = arg[4];
    br2::ModelConst mc; mc.set(p);
    br2::Trig t; br2::trig_of(x, t);
    br2::Jac J; br2::jac_of(x, mc, t, J);
    if (res[0]) br2::ode(x, u, mc, t, res[0]);
    if (res[1]) for (int j = 0; j < 12; j++) br2::jac_mul(J, Sx + 12 * j, res[1] + 12 * j);
    if (res[2]) for (int j = 0; j < 4; j++) { br2::jac_mul(J, Su + 12 * j, res[2] + 12 * j); br2::ju_add(mc, j, res[2] + 12 * j); }
    return 0;
}
BR2_MODEL_API int bluerov2_expl_vde_forw_work(int* a, int* b, int* c, int* d) { if (a) *a = 5; if (b) *b = 3; if (c) *c = 0; if (d) *d = 0; return 0; }
BR2_MODEL_API const int* bluerov2_expl_vde_forw_sparsity_in(int i)
{
    return i == 0 ? kSpX : i == 1 ? dense_sp(kSpSx, 12, 12) : i == 2 ? dense_sp(kSpSu, 12, 4) : i == 3 ? kSpU : i == 4 ? kSpP : nullptr;
}
BR2_MODEL_API const int* bluerov2_expl_vde_forw_sparsity_out(int i)
{
    return i == 0 ? kSpX : i == 1 ? dense_sp(kSpSx, 12, 12) : i == 2 ? dense_sp(kSpSu, 12, 4) : nullptr;
}
BR2_MODEL_API int bluerov2_expl_vde_forw_n_in(void) { return 5; }
BR2_MODEL_API int bluerov2_expl_vde_forw_n_out(void) { return 3; }

BR2_MODEL_API int bluerov2_expl_vde_adj(const real_t** arg, real_t** res, int*, real_t*, void*)
{
    // (x, lam, u, p) -> [Jx' lam; Ju' lam]: apply J to the unit vectors (12 sparse products) -- not on any hot path
    const double *x = arg[0], *lam = arg[1], *p = arg[3];
    br2::ModelConst mc; mc.set(p);
    br2::Trig t; br2::trig_of(x, t);
    br2::Jac J; br2::jac_of(x, mc, t, J);
    if (!res[0]) return 0;
    for (int j = 0; j < 12; j++) {
        double e[12] = {0}, c[12];
        e[j] = 1.0;
        br2::jac_mul(J, e, c);
        double s = 0;
        for (int i = 0; i < 12; i++) s += c[i] * lam[i];
        res[0][j] = s;
    }
    for (int a = 0; a < 4; a++) {
        double c[12] = {0};
        br2::ju_add(mc, a, c);
        double s = 0;
        for (int i = 0; i < 12; i++) s += c[i] * lam[i];
        res[0][12 + a] = s;
    }
    return 0;
}
BR2_MODEL_API int bluerov2_expl_vde_adj_work(int* a, int* b, int* c, int* d) { if (a) *a = 4; if (b) *b = 1; if (c) *c = 0; if (d) *d = 0; return 0; }
BR2_MODEL_API const int* bluerov2_expl_vde_adj_sparsity_in(int i) { return i == 0 ? kSpX : i == 1 ? kSpX : i == 2 ? kSpU : i == 3 ? kSpP : nullptr; }
BR2_MODEL_API const int* bluerov2_expl_vde_adj_sparsity_out(int i) { return i == 0 ? kSpP : nullptr; }
BR2_MODEL_API int bluerov2_expl_vde_adj_n_in(void) { return 4; }
BR2_MODEL_API int bluerov2_expl_vde_adj_n_out(void) { return 1; }

// ---- cost functions: NLS residual y = [x; u] (terminal y = x), bluerov2.py:144,153-154 -------------------------------------
// The nine functions of c_generated_code/bluerov2_cost/bluerov2_cost.h:45-114 with CasADi's six entry points each.  A null
// input reads as zeros and a null output is skipped, like the generated code.  The engine never calls them (its Gauss-Newton
// Hessian Ts*W is baked into the kernels); they complete the drop-in's symbol table and are checked against the reference's own
// generated C in tests/test_abi.py.
//   _fun            (x, u, z, p)        -> y
//   _fun_jac_ut_xt  (x, u, z, p)        -> y, d y / d [u; x] transposed ((nu+nx) x ny, one unit entry per column), (ny x 0)
//   _hess           (x, u, z, lam_y, p) -> (nu+nx) x (nu+nx), structurally empty: the residual is linear
static const int kSpEmpty[] = {0, 0, 0};
static const int kSpJacUtXt[] = {16, 16, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16,
                                 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 0, 1, 2, 3};        // column j = y_j: row of x_j is 4 + j, of u_a is a
static const int kSpNy0[] = {16, 0, 0};
static const int kSpHess[] = {16, 16, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
static const int kSpJacE[] = {12, 12, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11};
static const int kSpNyE0[] = {12, 0, 0};
static const int kSpHessE[] = {12, 12, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};

static void cost_y(const real_t** arg, real_t* y, bool terminal)
{
    for (int i = 0; i < 12; i++) y[i] = arg[0] ? arg[0][i] : 0.0;
    if (!terminal)
        for (int i = 0; i < 4; i++) y[12 + i] = arg[1] ? arg[1][i] : 0.0;
}
#define BR2_COST_HELPERS(f, n_in, n_out, SP_IN, SP_OUT)                                                                       \
    BR2_MODEL_API int f##_work(int* a, int* b, int* c, int* d) { if (a) *a = n_in; if (b) *b = n_out; if (c) *c = 0; if (d) *d = 0; return 0; } \
    BR2_MODEL_API const int* f##_sparsity_in(int i) { static const int* const t[] = SP_IN; return (i >= 0 && i < n_in) ? t[i] : nullptr; }     \
    BR2_MODEL_API const int* f##_sparsity_out(int i) { static const int* const t[] = SP_OUT; return (i >= 0 && i < n_out) ? t[i] : nullptr; }  \
    BR2_MODEL_API int f##_n_in(void) { return n_in; }                                                                          \
    BR2_MODEL_API int f##_n_out(void) { return n_out; }
#define BR2_LIST(...) {__VA_ARGS__}
#define BR2_COST_STAGE(pre)                                                                                                    \
    BR2_MODEL_API int pre##_fun(const real_t** arg, real_t** res, int*, real_t*, void*) { if (res[0]) cost_y(arg, res[0], false); return 0; } \
    BR2_COST_HELPERS(pre##_fun, 4, 1, BR2_LIST(kSpX, kSpU, kSpEmpty, kSpP), BR2_LIST(kSpP))                                   \
    BR2_MODEL_API int pre##_fun_jac_ut_xt(const real_t** arg, real_t** res, int*, real_t*, void*)                              \
    {                                                                                                                          \
        if (res[0]) cost_y(arg, res[0], false);                                                                                \
        if (res[1]) for (int i = 0; i < 16; i++) res[1][i] = 1.0;                                                              \
        return 0;                                                                                                              \
    }                                                                                                                          \
    BR2_COST_HELPERS(pre##_fun_jac_ut_xt, 4, 3, BR2_LIST(kSpX, kSpU, kSpEmpty, kSpP), BR2_LIST(kSpP, kSpJacUtXt, kSpNy0))     \
    BR2_MODEL_API int pre##_hess(const real_t**, real_t**, int*, real_t*, void*) { return 0; }                                \
    BR2_COST_HELPERS(pre##_hess, 5, 1, BR2_LIST(kSpX, kSpU, kSpEmpty, kSpP, kSpP), BR2_LIST(kSpHess))
BR2_COST_STAGE(bluerov2_cost_y_0)      // bluerov2_cost.h:45-66
BR2_COST_STAGE(bluerov2_cost_y)        // :70-91
// terminal node: y = x, no input (bluerov2_cost.h:95-116)
BR2_MODEL_API int bluerov2_cost_y_e_fun(const real_t** arg, real_t** res, int*, real_t*, void*) { if (res[0]) cost_y(arg, res[0], true); return 0; }
BR2_COST_HELPERS(bluerov2_cost_y_e_fun, 4, 1, BR2_LIST(kSpX, kSpEmpty, kSpEmpty, kSpP), BR2_LIST(kSpX))
BR2_MODEL_API int bluerov2_cost_y_e_fun_jac_ut_xt(const real_t** arg, real_t** res, int*, real_t*, void*)
{
    if (res[0]) cost_y(arg, res[0], true);
    if (res[1]) for (int i = 0; i < 12; i++) res[1][i] = 1.0;
    return 0;
}
BR2_COST_HELPERS(bluerov2_cost_y_e_fun_jac_ut_xt, 4, 3, BR2_LIST(kSpX, kSpEmpty, kSpEmpty, kSpP), BR2_LIST(kSpX, kSpJacE, kSpNyE0))
BR2_MODEL_API int bluerov2_cost_y_e_hess(const real_t**, real_t**, int*, real_t*, void*) { return 0; }
BR2_COST_HELPERS(bluerov2_cost_y_e_hess, 5, 1, BR2_LIST(kSpX, kSpEmpty, kSpEmpty, kSpX, kSpP), BR2_LIST(kSpHessE))
#undef BR2_COST_STAGE
#undef BR2_COST_HELPERS
#undef BR2_LIST

}  // extern "C"
