// engine.cu -- host side of the batched SQP-RTI engine and its C-ABI (include/bluerov2_b200.h).
//
// Owns the HBM-resident state of B OCP instances (iterate, stage records, IPM workspaces, EKF state) and
// enqueues the kernels of kernels.cu / ekf.cu.  No CPU fallback exists: creation fails without a CUDA device.
#include "engine.h"
#include "../../include/bluerov2_b200.h"

#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

using namespace br2;

static thread_local char g_err[512] = "";
static int fail(int code, const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}
#define CK(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess) return fail(BR2_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

struct br2_batch_solver {
    int B, N, device, sm_count;
    double Ts[NMAX];
    double W[16], We[12], lbu[4], ubu[4];
    int max_iter, fast_path, ekf_model, active_set;
    double tol;
    // device state
    double *d_Ts, *d_X, *d_U, *d_S, *d_u0, *d_thrust, *d_info;
    int *d_status, *d_iters, *d_counter, *d_hint, *d_aset, *d_order, *d_fb;
    double *d_x0, *d_yref, *d_p;              // staging for the host API
    double* d_traj;                           // reference trajectory for device-side windowing [traj_rows][16]
    int* d_lines;                             // staging: first trajectory row per instance
    int traj_rows;
    double *d_ex, *d_eP, *d_thr, *d_meas, *d_acc, *d_wf, *d_pout;   // EKF state + staging
    double* d_rls;                            // RLS-VFF state [B][4][RLS_STRIDE] (AMPC)
    float* d_yaw;                             // continuous-yaw accumulators [B][2] = (pre_yaw, yaw_sum), floats as in the node
    cudaStream_t stream_xchg; cudaEvent_t ev_xfork, ev_xdone;   // side branch that ships the previous tick's thrusts to the peers
    cudaStream_t stream_c[4];          // compute streams of the ranges of a pipelined tick
    cudaEvent_t ev_done[4];
    cudaStream_t stream, stream_x0;   // host API: main stream; second stream carrying the x0 upload past the lineariser
    cudaEvent_t ev0, ev1, ev_mid;   // solve start / end / between linearisation and IPM
    cudaEvent_t ev_x0;              // x0 upload complete (the IPM kernel is its first reader)
    unsigned long long* d_iter_total;   // IPM iterations executed, summed over instances and solves
    bool timed;
    // tick graphs: the kernels (and, for host buffers, the copies) of one control tick instantiated as a CUDA graph, keyed on the
    // caller's buffers and on a generation counter that every change of options / weights / bounds / trajectory bumps
    struct H2DNode { cudaGraphNode_t node; int which; size_t off, bytes; void* dst; };   // upload node: input index, offset inside it
    struct TickGraph { br2_tick_io io; int host; int xchg; unsigned gen; unsigned long long stamp; cudaGraphExec_t exec; cudaGraph_t graph;
                       H2DNode up[12]; int nup; } tg[4];
    // sharding: local gather buffer + flags (one allocation, exported through CUDA IPC), mapped peers
    ShardView shard;
    double* d_shard;          // [2][world * B][6] doubles, then [world] int flags, then [1] int time-out marker
    void* peer_base[MAX_SHARDS];   // cudaIpcOpenMemHandle mappings (to close)
    long long ticks_host;     // ticks enqueued so far (host-side mirror of ctr[CTR_TICK])
    long long xchg_host;      // exchanges enqueued so far (mirror of ctr[CTR_XTICK]): tick t is shipped as a side branch of tick t + 1
    int graph_updates;      // replays of a host graph on NEW input buffers (upload nodes re-pointed, nothing re-instantiated)
    unsigned gen;
    unsigned long long tick_stamp;
    int graphs_built, kernel_timing, tick_graph;
    cudaEvent_t ev_fork, ev_chunk[4];
    // host path, explicit reference known one tick ahead (br2_batch_set_next_yref_host): the next window is uploaded on a stream of
    // its own behind this tick's small uploads, into the one of two buffers the running tick does not read
    double* d_yref_pf[2];
    cudaStream_t stream_pf;
    cudaEvent_t ev_pf, ev_up_p, ev_up_x0;   // prefetch complete; this tick's p / x0 uploads complete (external events inside the tick graph)
    int p_resident;                        // parameters of the last host call that supplied them sit in d_p: 1 = [B][16], 2 = per stage
    const double* next_yref;               // registered for the tick after the coming one
    const double* pf_host;                 // host buffer whose window sits (or is arriving) in d_yref_pf[pf_slot]
    int pf_slot;
    struct Miss { br2_tick_io io; int host; int xchg; unsigned gen; bool valid; } miss[4];   // the last keys that missed the graph cache
    unsigned miss_next;
};

// generated-C defaults: acados_solver_bluerov2.c:424-459 (W), :468-479 (W_e), :547-571 (bounds), :522-541 / :681-708 (x init);
// pinned on the reference's scripts/acados_ocp.json by tests/test_ocp_json_pins.py
static const double kW[16] = {300, 480, 200, 10, 10, 200, 40, 40, 10, 10, 10, 10, 1, 1, 0.1, 0.05};

extern "C" void br2_get_ocp_defaults(br2_ocp_defaults* d)
{
    if (!d) return;
    memset(d, 0, sizeof *d);
    d->N = 80; d->nx = NX; d->nu = NU; d->np = NP; d->ny = NY; d->ny_e = NX;
    d->nbu = NU; d->nbx0 = NX; d->nbxe0 = NX;
    d->qp_iter_max = 50; d->qp_warm_start = 0; d->erk_stages = 4; d->erk_steps = 1;
    d->Tf = 1.0;
    memcpy(d->W, kW, sizeof kW);
    memcpy(d->We, kW, sizeof(double) * 12);
    for (int i = 0; i < 4; i++) { d->lbu[i] = -50.0; d->ubu[i] = 50.0; }
    d->x_init[2] = -20.0;
}

extern "C" const char* br2_last_error(void) { return g_err; }
extern "C" const char* br2_version(void) { return "bluerov2_b200 0.1 (sm_100a)"; }
extern "C" int br2_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

template <class T>
static cudaError_t dalloc(T** p, size_t n) { return cudaMalloc((void**)p, n * sizeof(T)); }

// Every entry point works on the solver's device and leaves the caller's current device as it found it.
struct DeviceGuard {
    int prev = -1;
    cudaError_t err;
    explicit DeviceGuard(int dev)
    {
        if (cudaGetDevice(&prev) != cudaSuccess) { cudaGetLastError(); prev = -1; }
        err = (prev == dev) ? cudaSuccess : cudaSetDevice(dev);
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};
#define ON_DEVICE(s) DeviceGuard guard_((s)->device); CK(guard_.err)

extern "C" int br2_batch_free(br2_batch_solver* s)
{
    if (!s) return BR2_OK;
    DeviceGuard guard_(s->device);
    void* ptrs[] = {s->d_Ts, s->d_X, s->d_U, s->d_S, s->d_u0, s->d_thrust, s->d_info, s->d_status,
                    s->d_iters, s->d_counter, s->d_x0, s->d_yref, s->d_p, s->d_ex, s->d_eP, s->d_thr, s->d_meas,
                    s->d_acc, s->d_wf, s->d_pout, s->d_iter_total, s->d_hint, s->d_traj, s->d_lines, s->d_rls, s->d_yaw, s->d_aset, s->d_order, s->d_fb};
    for (void* p : ptrs)
        if (p) cudaFree(p);
    if (s->ev0) cudaEventDestroy(s->ev0);
    if (s->ev1) cudaEventDestroy(s->ev1);
    if (s->ev_mid) cudaEventDestroy(s->ev_mid);
    if (s->ev_pf) cudaEventDestroy(s->ev_pf);
    if (s->ev_up_p) cudaEventDestroy(s->ev_up_p);
    if (s->ev_up_x0) cudaEventDestroy(s->ev_up_x0);
    if (s->stream_pf) cudaStreamDestroy(s->stream_pf);
    for (int i = 0; i < 2; i++)
        if (s->d_yref_pf[i]) cudaFree(s->d_yref_pf[i]);
    if (s->ev_x0) cudaEventDestroy(s->ev_x0);
    if (s->ev_fork) cudaEventDestroy(s->ev_fork);
    for (auto e : s->ev_chunk)
        if (e) cudaEventDestroy(e);
    for (auto e : s->ev_done)
        if (e) cudaEventDestroy(e);
    for (auto c : s->stream_c)
        if (c) cudaStreamDestroy(c);
    if (s->stream_xchg) cudaStreamDestroy(s->stream_xchg);
    if (s->ev_xfork) cudaEventDestroy(s->ev_xfork);
    if (s->ev_xdone) cudaEventDestroy(s->ev_xdone);
    for (int r = 0; r < MAX_SHARDS; r++)
        if (s->peer_base[r]) cudaIpcCloseMemHandle(s->peer_base[r]);
    if (s->d_shard) cudaFree(s->d_shard);
    for (auto& g : s->tg) {
        if (g.exec) cudaGraphExecDestroy(g.exec);
        if (g.graph) cudaGraphDestroy(g.graph);
    }
    if (s->stream) cudaStreamDestroy(s->stream);
    if (s->stream_x0) cudaStreamDestroy(s->stream_x0);
    free(s);
    return BR2_OK;
}

extern "C" int br2_batch_create(br2_batch_solver** out, int batch, int N, const double* time_steps, int device)
{
    if (!out) return fail(BR2_EINVAL, "br2_batch_create: out is NULL");
    *out = nullptr;
    if (batch < 1) return fail(BR2_EINVAL, "br2_batch_create: batch = %d", batch);
    if (N < 1 || N > NMAX) return fail(BR2_EINVAL, "br2_batch_create: N = %d outside [1, %d]", N, NMAX);
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(BR2_ECUDA, "br2_batch_create: no CUDA device (%s); this library has no CPU path",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
    }
    if (device < 0 || device >= ndev) return fail(BR2_EINVAL, "br2_batch_create: device %d of %d", device, ndev);
    DeviceGuard guard_(device);
    CK(guard_.err);
    br2_batch_solver* s = (br2_batch_solver*)calloc(1, sizeof(br2_batch_solver));
    if (!s) return fail(BR2_ENOMEM, "br2_batch_create: out of host memory");
    s->B = batch; s->N = N; s->device = device;
#define CKF(call)                                                                                        \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess) {                                                                         \
            br2_batch_free(s);                                                                           \
            return fail(BR2_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
        }                                                                                                \
    } while (0)
    cudaDeviceProp prop;
    CKF(cudaGetDeviceProperties(&prop, device));
    s->sm_count = prop.multiProcessorCount;
    br2_ocp_defaults dflt;
    br2_get_ocp_defaults(&dflt);
    for (int k = 0; k < N; k++) s->Ts[k] = time_steps ? time_steps[k] : dflt.Tf / N;
    memcpy(s->W, dflt.W, sizeof s->W);
    memcpy(s->We, dflt.We, sizeof s->We);
    memcpy(s->lbu, dflt.lbu, sizeof s->lbu);
    memcpy(s->ubu, dflt.ubu, sizeof s->ubu);
    s->max_iter = dflt.qp_iter_max;
    s->fast_path = 1;
    s->active_set = 1;      // primal-dual active-set iteration on by default (option "active_set_path")
    s->tol = 1e-12;
    s->kernel_timing = 1;
    s->tick_graph = 1;
    const size_t B = batch;
#define DA(p, n)                                                                                          \
    do {                                                                                                  \
        cudaError_t e_ = dalloc(&s->p, (n));                                                              \
        if (e_ != cudaSuccess) {                                                                          \
            br2_batch_free(s);                                                                            \
            return fail(BR2_ENOMEM, "cudaMalloc(%s, %zu elements) failed: %s", #p, (size_t)(n), cudaGetErrorString(e_)); \
        }                                                                                                 \
    } while (0)
    DA(d_Ts, N); DA(d_X, B * (N + 1) * NX); DA(d_U, B * N * NU);
    DA(d_S, B * (N + 1) * SREC);
    DA(d_u0, B * 4); DA(d_thrust, B * 6); DA(d_info, B * 4); DA(d_status, B); DA(d_iters, B); DA(d_counter, CTR_COUNT); DA(d_order, 2 * B); DA(d_fb, B);
    DA(d_x0, B * NX); DA(d_yref, B * (N + 1) * NY); DA(d_p, B * (N + 1) * NP);
    DA(d_ex, B * 18); DA(d_eP, B * 324); DA(d_thr, B * 6); DA(d_meas, B * 12); DA(d_acc, B * 6); DA(d_wf, B * 6);
    DA(d_pout, B * NP);
    DA(d_rls, B * 4 * RLS_STRIDE);
    DA(d_yaw, B * 2);
    DA(d_iter_total, 1 + 16 + 1); // [0] iteration counter, [1..16] phase cycle counters of the instrumentation build, [17] non-zero statuses
    DA(d_hint, B);
    DA(d_aset, B * N);
    DA(d_lines, B);
#undef DA
    CKF(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
    CKF(cudaStreamCreateWithFlags(&s->stream_x0, cudaStreamNonBlocking));
    CKF(cudaEventCreateWithFlags(&s->ev_x0, cudaEventDisableTiming));
    CKF(cudaEventCreateWithFlags(&s->ev_fork, cudaEventDisableTiming));
    CKF(cudaStreamCreateWithFlags(&s->stream_pf, cudaStreamNonBlocking));
    CKF(cudaEventCreateWithFlags(&s->ev_pf, cudaEventDisableTiming));
    CKF(cudaEventCreateWithFlags(&s->ev_up_p, cudaEventDisableTiming));
    CKF(cudaEventCreateWithFlags(&s->ev_up_x0, cudaEventDisableTiming));
    for (auto& e : s->ev_chunk) CKF(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    for (auto& e : s->ev_done) CKF(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    for (auto& c : s->stream_c) CKF(cudaStreamCreateWithFlags(&c, cudaStreamNonBlocking));
    CKF(cudaStreamCreateWithFlags(&s->stream_xchg, cudaStreamNonBlocking));
    CKF(cudaEventCreateWithFlags(&s->ev_xfork, cudaEventDisableTiming));
    CKF(cudaEventCreateWithFlags(&s->ev_xdone, cudaEventDisableTiming));
    CKF(cudaEventCreate(&s->ev0));
    CKF(cudaEventCreate(&s->ev1));
    CKF(cudaEventCreate(&s->ev_mid));
    CKF(cudaMemset(s->d_iter_total, 0, sizeof(unsigned long long) * 18));
    CKF(cudaMemcpy(s->d_Ts, s->Ts, sizeof(double) * N, cudaMemcpyHostToDevice));
    CKF(cudaMemset(s->d_status, 0, sizeof(int) * B));
    CKF(cudaMemset(s->d_iters, 0, sizeof(int) * B));
    CKF(cudaMemset(s->d_hint, 0, sizeof(int) * B));
    CKF(cudaMemset(s->d_aset, 0, sizeof(int) * B * N));
    CKF(cudaMemset(s->d_counter, 0, sizeof(int) * CTR_COUNT));
    {   // visiting order of the IPM kernel: identity in both halves until a solve has ranked the instances
        int* h = (int*)malloc(sizeof(int) * 2 * B);
        if (!h) { br2_batch_free(s); return fail(BR2_ENOMEM, "br2_batch_create: out of host memory"); }
        for (size_t i = 0; i < 2 * B; i++) h[i] = (int)(i % B);
        cudaError_t e3 = cudaMemcpy(s->d_order, h, sizeof(int) * 2 * B, cudaMemcpyHostToDevice);
        free(h);
        if (e3 != cudaSuccess) { br2_batch_free(s); return fail(BR2_ECUDA, "cudaMemcpy(order) failed: %s", cudaGetErrorString(e3)); }
    }
    configure_kernels();
    configure_ekf();
    CKF(cudaMemset(s->d_info, 0, sizeof(double) * B * 4));
    CKF(cudaMemset(s->d_S, 0, sizeof(double) * B * (N + 1) * SREC));
    *out = s;
    int rc = br2_batch_reset(s, 0);
    if (rc == BR2_OK) rc = br2_batch_ekf_reset(s);
    if (rc == BR2_OK) rc = br2_batch_rls_reset(s);
    if (rc == BR2_OK) rc = br2_batch_yaw_reset(s);
    if (rc == BR2_OK) { cudaError_t e2 = cudaMemset(s->d_pout, 0, sizeof(double) * B * NP); if (e2 != cudaSuccess) rc = fail(BR2_ECUDA, "cudaMemset failed"); }
    if (rc != BR2_OK) { br2_batch_free(s); *out = nullptr; }
    return rc;
#undef CKF
}

extern "C" int br2_batch_size(const br2_batch_solver* s) { return s ? s->B : 0; }
extern "C" int br2_batch_horizon(const br2_batch_solver* s) { return s ? s->N : 0; }

extern "C" int br2_batch_set_weights(br2_batch_solver* s, const double* W16, const double* We12)
{
    if (!s) return fail(BR2_EINVAL, "null solver");
    if (W16) {
        for (int i = 0; i < 16; i++)
            if (!(W16[i] >= 0) || (i >= 12 && !(W16[i] > 0))) return fail(BR2_EINVAL, "W[%d] = %g: need W >= 0 and R > 0", i, W16[i]);
        memcpy(s->W, W16, sizeof(double) * 16);
    }
    if (We12) {
        for (int i = 0; i < 12; i++)
            if (!(We12[i] >= 0)) return fail(BR2_EINVAL, "We[%d] = %g: need >= 0", i, We12[i]);
        memcpy(s->We, We12, sizeof(double) * 12);
    }
    s->gen++;
    return BR2_OK;
}

extern "C" int br2_batch_set_bounds(br2_batch_solver* s, const double* lbu4, const double* ubu4)
{
    if (!s || !lbu4 || !ubu4) return fail(BR2_EINVAL, "null argument");
    for (int i = 0; i < 4; i++)
        if (!(lbu4[i] < ubu4[i])) return fail(BR2_EINVAL, "bounds %d: lbu %g >= ubu %g", i, lbu4[i], ubu4[i]);
    memcpy(s->lbu, lbu4, sizeof(double) * 4);
    memcpy(s->ubu, ubu4, sizeof(double) * 4);
    s->gen++;
    return BR2_OK;
}

extern "C" int br2_batch_set_time_steps(br2_batch_solver* s, const double* ts)
{
    if (!s || !ts) return fail(BR2_EINVAL, "null argument");
    for (int k = 0; k < s->N; k++)
        if (!(ts[k] > 0)) return fail(BR2_EINVAL, "time_steps[%d] = %g", k, ts[k]);
    memcpy(s->Ts, ts, sizeof(double) * s->N);
    ON_DEVICE(s);
    CK(cudaMemcpy(s->d_Ts, s->Ts, sizeof(double) * s->N, cudaMemcpyHostToDevice));
    return BR2_OK;
}

extern "C" int br2_batch_set_option_int(br2_batch_solver* s, const char* name, int v)
{
    if (!s || !name) return fail(BR2_EINVAL, "null argument");
    s->gen++;
    if (!strcmp(name, "kernel_timing")) { s->kernel_timing = v != 0; return BR2_OK; }
    if (!strcmp(name, "tick_graph")) { s->tick_graph = v != 0; return BR2_OK; }
    if (!strcmp(name, "qp_iter_max")) {
        if (v < 1) return fail(BR2_EINVAL, "qp_iter_max = %d", v);
        s->max_iter = v;
        return BR2_OK;
    }
    if (!strcmp(name, "fast_path")) {
        s->fast_path = v != 0;
        return BR2_OK;
    }
    if (!strcmp(name, "active_set_path")) {   // active-set fast path for instances whose previous solution had active bounds
        s->active_set = v != 0;
        return BR2_OK;
    }
    if (!strcmp(name, "ekf_model")) {      // 0 = BLUEROV2_DOB filter, 1 = BLUEROV2_AMPC filter (bluerov2_ampc.cpp:658-696)
        if (v != 0 && v != 1) return fail(BR2_EINVAL, "ekf_model = %d (0 = dob, 1 = ampc)", v);
        s->ekf_model = v;
        return BR2_OK;
    }
    return fail(BR2_EINVAL, "unknown int option '%s'", name);
}
extern "C" int br2_batch_set_option_double(br2_batch_solver* s, const char* name, double v)
{
    if (!s || !name) return fail(BR2_EINVAL, "null argument");
    s->gen++;
    if (!strcmp(name, "qp_tol")) {
        if (!(v > 0)) return fail(BR2_EINVAL, "qp_tol = %g", v);
        s->tol = v;
        return BR2_OK;
    }
    return fail(BR2_EINVAL, "unknown double option '%s'", name);
}

extern "C" int br2_batch_reset(br2_batch_solver* s, int mode)
{
    if (!s) return fail(BR2_EINVAL, "null solver");
    ON_DEVICE(s);
    const size_t nX = (size_t)s->B * (s->N + 1) * NX, nU = (size_t)s->B * s->N * NU;
    CK(cudaMemset(s->d_U, 0, sizeof(double) * nU));
    CK(cudaMemset(s->d_X, 0, sizeof(double) * nX));
    CK(cudaMemset(s->d_hint, 0, sizeof(int) * s->B));
    if (mode == 0) {
        // x_k = (0, 0, -20, 0, ...) for every stage: acados_solver_bluerov2.c:681-708
        double* h = (double*)calloc(nX, sizeof(double));
        if (!h) return fail(BR2_ENOMEM, "out of host memory");
        br2_ocp_defaults dflt;
        br2_get_ocp_defaults(&dflt);
        for (size_t i = 0; i < nX / NX; i++) memcpy(h + i * NX, dflt.x_init, sizeof dflt.x_init);
        cudaError_t e = cudaMemcpy(s->d_X, h, sizeof(double) * nX, cudaMemcpyHostToDevice);
        free(h);
        CK(e);
    }
    return BR2_OK;
}

extern "C" int br2_batch_set_iterate_host(br2_batch_solver* s, const double* X, const double* U)
{
    if (!s) return fail(BR2_EINVAL, "null solver");
    ON_DEVICE(s);
    CK(cudaDeviceSynchronize());
    if (X) CK(cudaMemcpy(s->d_X, X, sizeof(double) * s->B * (s->N + 1) * NX, cudaMemcpyHostToDevice));
    if (U) CK(cudaMemcpy(s->d_U, U, sizeof(double) * s->B * s->N * NU, cudaMemcpyHostToDevice));
    CK(cudaMemset(s->d_hint, 0, sizeof(int) * s->B));   // a new iterate invalidates the active-set history
    return BR2_OK;
}
extern "C" int br2_batch_get_iterate_host(br2_batch_solver* s, double* X, double* U)
{
    if (!s) return fail(BR2_EINVAL, "null solver");
    ON_DEVICE(s);
    CK(cudaDeviceSynchronize());
    if (X) CK(cudaMemcpy(X, s->d_X, sizeof(double) * s->B * (s->N + 1) * NX, cudaMemcpyDeviceToHost));
    if (U) CK(cudaMemcpy(U, s->d_U, sizeof(double) * s->B * s->N * NU, cudaMemcpyDeviceToHost));
    return BR2_OK;
}
extern "C" int br2_batch_iterate_device(br2_batch_solver* s, double** X, double** U)
{
    if (!s) return fail(BR2_EINVAL, "null solver");
    if (X) *X = s->d_X;
    if (U) *U = s->d_U;
    return BR2_OK;
}

static void fill_args(br2_batch_solver* s, SolveArgs& a, const double* d_x0, const double* d_yref, const int* d_lines,
                      const double* d_p, int p_per_stage, double* d_u0, double* d_thrust, int* d_status)
{
    a.B = s->B; a.N = s->N; a.lo = 0; a.hi = s->B; a.housekeeping = 2; a.qidx = 0;
    a.x0 = d_x0; a.yref = d_yref; a.p = d_p;
    a.traj = s->d_traj; a.lines = d_lines; a.traj_rows = s->traj_rows;
    a.p_inst_stride = p_per_stage ? (s->N + 1) * NP : NP;
    a.p_stage_stride = p_per_stage ? NP : 0;
    a.Ts = s->d_Ts;
    memcpy(a.W, s->W, sizeof a.W); memcpy(a.We, s->We, sizeof a.We);
    memcpy(a.lbu, s->lbu, sizeof a.lbu); memcpy(a.ubu, s->ubu, sizeof a.ubu);
    a.X = s->d_X; a.U = s->d_U; a.S = s->d_S;
    a.u0 = d_u0 ? d_u0 : s->d_u0;
    a.thrust = d_thrust ? d_thrust : s->d_thrust;
    a.status = d_status ? d_status : s->d_status;
    a.iters = s->d_iters; a.info = s->d_info; a.ctr = s->d_counter; a.order = s->d_order; a.fb = s->d_fb; a.shard = s->shard; a.iter_total = s->d_iter_total; a.prof = s->d_iter_total + 1; a.bad_total = s->d_iter_total + 17; a.hint = s->d_hint; a.fast_path = s->fast_path; a.aset = s->d_aset; a.active_set = s->active_set;
    a.max_iter = s->max_iter; a.tol = s->tol;
}

static int solve_enqueue(br2_batch_solver* s, const double* d_x0, const double* d_yref, const int* d_lines, const double* d_p,
                         int p_per_stage, double* d_u0, double* d_thrust, int* d_status, cudaStream_t st,
                         cudaEvent_t before_ipm = nullptr);

// Host output buffer that the device can write directly (pinned + mapped under unified addressing: cudaHostAlloc /
// cudaHostRegister / torch pin_memory): returns its device alias, else null (-> staged copy).  Writing u0 / thrust /
// status from the kernel epilogue straight into such a buffer removes three device-to-host copies from every host call.
template <class T>
static T* mapped_alias(T* host)
{
    if (!host) return nullptr;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, host) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    if (at.type != cudaMemoryTypeHost || !at.devicePointer) return nullptr;
    return (T*)at.devicePointer;
}

// common tail of the two host entry points: x0 rides a second stream past the lineariser (the IPM kernel is its first
// reader); outputs go straight to mapped host buffers when possible
static int solve_host_common(br2_batch_solver* s, const double* x0, const double* d_yref, const int* d_lines, int p_per_stage,
                             double* u0, double* thrust, int* status)
{
    const size_t B = s->B;
    cudaStream_t st = s->stream;
    // the lineariser goes out first (its inputs -- p, reference -- are already enqueued): everything below is host work and
    // an upload that the GPU does not have to wait for before it starts
    ON_DEVICE(s);
    SolveArgs a;
    fill_args(s, a, s->d_x0, d_yref, d_lines, s->d_p, p_per_stage, nullptr, nullptr, nullptr);
    CK(cudaEventRecord(s->ev0, st));
    launch_linearize(a, st);
    if (s->kernel_timing) CK(cudaEventRecord(s->ev_mid, st));
    CK(cudaMemcpyAsync(s->d_x0, x0, sizeof(double) * B * NX, cudaMemcpyHostToDevice, s->stream_x0));
    CK(cudaEventRecord(s->ev_x0, s->stream_x0));
    double* a_u0 = mapped_alias(u0);
    double* a_th = mapped_alias(thrust);
    int* a_st = mapped_alias(status);
    if (a_u0) a.u0 = a_u0;
    if (a_th) a.thrust = a_th;
    if (a_st) a.status = a_st;
    CK(cudaStreamWaitEvent(st, s->ev_x0, 0));
    launch_ipm(a, s->sm_count, st);
    CK(cudaEventRecord(s->ev1, st));
    s->timed = true;
    s->ticks_host++;
    CK(cudaGetLastError());
    if (u0 && !a_u0) CK(cudaMemcpyAsync(u0, s->d_u0, sizeof(double) * B * 4, cudaMemcpyDeviceToHost, st));
    if (thrust && !a_th) CK(cudaMemcpyAsync(thrust, s->d_thrust, sizeof(double) * B * 6, cudaMemcpyDeviceToHost, st));
    if (status && !a_st) CK(cudaMemcpyAsync(status, s->d_status, sizeof(int) * B, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return BR2_OK;
}

extern "C" int br2_batch_solve_device(br2_batch_solver* s, const double* d_x0, const double* d_yref, const double* d_p,
                                      int p_per_stage, double* d_u0, double* d_thrust, int* d_status, void* stream)
{
    if (!s || !d_x0 || !d_yref || !d_p) return fail(BR2_EINVAL, "br2_batch_solve_device: null argument");
    return solve_enqueue(s, d_x0, d_yref, nullptr, d_p, p_per_stage, d_u0, d_thrust, d_status, (cudaStream_t)stream);
}

extern "C" int br2_batch_set_trajectory(br2_batch_solver* s, const double* traj, int rows)
{
    if (!s || !traj || rows < 1) return fail(BR2_EINVAL, "br2_batch_set_trajectory: bad argument");
    ON_DEVICE(s);
    CK(cudaDeviceSynchronize());
    if (s->d_traj) { cudaFree(s->d_traj); s->d_traj = nullptr; }
    cudaError_t e = dalloc(&s->d_traj, (size_t)rows * NY);
    if (e != cudaSuccess) return fail(BR2_ENOMEM, "cudaMalloc(trajectory, %d rows) failed: %s", rows, cudaGetErrorString(e));
    CK(cudaMemcpy(s->d_traj, traj, sizeof(double) * (size_t)rows * NY, cudaMemcpyHostToDevice));
    s->traj_rows = rows;
    s->gen++;
    return BR2_OK;
}

extern "C" int br2_batch_solve_windowed_device(br2_batch_solver* s, const double* d_x0, const int* d_lines, const double* d_p,
                                               int p_per_stage, double* d_u0, double* d_thrust, int* d_status, void* stream)
{
    if (!s || !d_x0 || !d_lines || !d_p) return fail(BR2_EINVAL, "br2_batch_solve_windowed_device: null argument");
    if (!s->d_traj) return fail(BR2_EINVAL, "br2_batch_solve_windowed_device: no trajectory set (br2_batch_set_trajectory)");
    return solve_enqueue(s, d_x0, nullptr, d_lines, d_p, p_per_stage, d_u0, d_thrust, d_status, (cudaStream_t)stream);
}

extern "C" int br2_batch_solve_windowed_host(br2_batch_solver* s, const double* x0, const int* lines, const double* p,
                                             int p_per_stage, double* u0, double* thrust, int* status)
{
    if (!s || !x0 || !lines || !p) return fail(BR2_EINVAL, "br2_batch_solve_windowed_host: null argument");
    if (!s->d_traj) return fail(BR2_EINVAL, "br2_batch_solve_windowed_host: no trajectory set (br2_batch_set_trajectory)");
    ON_DEVICE(s);
    const size_t B = s->B, N = s->N;
    cudaStream_t st = s->stream;
    CK(cudaMemcpyAsync(s->d_p, p, sizeof(double) * B * (p_per_stage ? (N + 1) * NP : NP), cudaMemcpyHostToDevice, st));
    s->p_resident = p_per_stage ? 2 : 1;
    CK(cudaMemcpyAsync(s->d_lines, lines, sizeof(int) * B, cudaMemcpyHostToDevice, st));
    return solve_host_common(s, x0, nullptr, s->d_lines, p_per_stage, u0, thrust, status);
}

static int solve_enqueue(br2_batch_solver* s, const double* d_x0, const double* d_yref, const int* d_lines, const double* d_p,
                         int p_per_stage, double* d_u0, double* d_thrust, int* d_status, cudaStream_t st, cudaEvent_t before_ipm)
{
    ON_DEVICE(s);
    SolveArgs a;
    fill_args(s, a, d_x0, d_yref, d_lines, d_p, p_per_stage, d_u0, d_thrust, d_status);
    CK(cudaEventRecord(s->ev0, st));
    launch_linearize(a, st);
    if (s->kernel_timing) CK(cudaEventRecord(s->ev_mid, st));
    if (before_ipm) CK(cudaStreamWaitEvent(st, before_ipm, 0));
    launch_ipm(a, s->sm_count, st);
    CK(cudaEventRecord(s->ev1, st));
    s->timed = true;
    s->ticks_host++;
    CK(cudaGetLastError());
    return BR2_OK;
}

extern "C" int br2_batch_solve_host(br2_batch_solver* s, const double* x0, const double* yref, const double* p,
                                    int p_per_stage, double* u0, double* thrust, int* status)
{
    if (!s || !x0 || !yref || !p) return fail(BR2_EINVAL, "br2_batch_solve_host: null argument");
    ON_DEVICE(s);
    const size_t B = s->B, N = s->N;
    cudaStream_t st = s->stream;
    CK(cudaMemcpyAsync(s->d_p, p, sizeof(double) * B * (p_per_stage ? (N + 1) * NP : NP), cudaMemcpyHostToDevice, st));
    s->p_resident = p_per_stage ? 2 : 1;
    CK(cudaMemcpyAsync(s->d_yref, yref, sizeof(double) * B * (N + 1) * NY, cudaMemcpyHostToDevice, st));
    return solve_host_common(s, x0, s->d_yref, nullptr, p_per_stage, u0, thrust, status);
}

extern "C" int br2_batch_get_stats_host(br2_batch_solver* s, int* iters, double* info)
{
    if (!s) return fail(BR2_EINVAL, "null solver");
    ON_DEVICE(s);
    CK(cudaDeviceSynchronize());
    if (iters) CK(cudaMemcpy(iters, s->d_iters, sizeof(int) * s->B, cudaMemcpyDeviceToHost));
    if (info) CK(cudaMemcpy(info, s->d_info, sizeof(double) * s->B * 4, cudaMemcpyDeviceToHost));
    return BR2_OK;
}

extern "C" int br2_batch_get_linearization_host(br2_batch_solver* s, double* AB, double* b)
{
    if (!s) return fail(BR2_EINVAL, "null solver");
    ON_DEVICE(s);
    CK(cudaDeviceSynchronize());
    const size_t N = s->N, nrec = (size_t)s->B * (N + 1);
    double* h = (double*)malloc(sizeof(double) * nrec * SREC);
    if (!h) return fail(BR2_ENOMEM, "out of host memory");
    cudaError_t e = cudaMemcpy(h, s->d_S, sizeof(double) * nrec * SREC, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess)
        for (size_t b_ = 0; b_ < (size_t)s->B; b_++)
            for (size_t k = 0; k < N; k++) {
                const double* g = h + (b_ * (N + 1) + k) * SREC + S_G;
                const size_t i = b_ * N + k;
                if (AB)   // un-permute the fragment order (layout.h) into row-major 12 x 16
                    for (int r = 0; r < 12; r++)
                        for (int c = 0; c < 16; c++) AB[i * 192 + r * 16 + c] = g[g_off(r, c)];
                if (b) memcpy(b + i * 12, g + G_B_OFF, sizeof(double) * 12);
            }
    free(h);
    CK(e);
    return BR2_OK;
}

extern "C" double br2_batch_last_solve_time(br2_batch_solver* s)
{
    if (!s || !s->timed) return 0.0;
    DeviceGuard guard_(s->device);
    if (cudaEventSynchronize(s->ev1) != cudaSuccess) return 0.0;
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, s->ev0, s->ev1) != cudaSuccess) return 0.0;
    return ms * 1e-3;
}

extern "C" int br2_batch_last_kernel_times(br2_batch_solver* s, double* t_linearize, double* t_ipm)
{
    if (!s || !s->timed) return fail(BR2_EINVAL, "no solve recorded");
    if (!s->kernel_timing) return fail(BR2_EINVAL, "option kernel_timing is 0: no event between the kernels");
    ON_DEVICE(s);
    CK(cudaEventSynchronize(s->ev1));
    float a = 0.f, b = 0.f;
    CK(cudaEventElapsedTime(&a, s->ev0, s->ev_mid));
    CK(cudaEventElapsedTime(&b, s->ev_mid, s->ev1));
    if (t_linearize) *t_linearize = a * 1e-3;
    if (t_ipm) *t_ipm = b * 1e-3;
    return BR2_OK;
}

extern "C" long long br2_batch_ipm_iterations_total(br2_batch_solver* s, int reset)
{
    if (!s) return -1;
    DeviceGuard guard_(s->device);
    if (cudaDeviceSynchronize() != cudaSuccess) return -1;
    unsigned long long v = 0;
    if (cudaMemcpy(&v, s->d_iter_total, sizeof v, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    if (reset) cudaMemset(s->d_iter_total, 0, sizeof v);
    return (long long)v;
}

// ---- one control tick as one call / one CUDA graph -----------------------------------------------------------------
static bool pinned_host(const void* p)
{
    if (!p) return true;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeHost;
}

// sharded and fully connected, with a finished tick whose thrust block has not been shipped yet?
static int shard_pending(const br2_batch_solver* s)
{
    if (s->shard.world <= 1 || s->ticks_host <= s->xchg_host) return 0;
    for (int r = 0; r < s->shard.world; r++)
        if (!s->shard.buf[r]) return 0;
    return 1;
}

// Everything of a tick, issued on `st` (plus stream_x0 for the forked x0 upload of the host path).  Runs either inside a
// stream capture (graph construction) or directly.
static int tick_issue(br2_batch_solver* s, const br2_tick_io& io, int host, cudaStream_t st, int xchg)
{
    const size_t B = s->B, N = s->N;
    // the timing events must stay usable from outside a graph: inside a capture they are recorded as EXTERNAL event nodes
    cudaStreamCaptureStatus capst = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &capst) != cudaSuccess) { cudaGetLastError(); capst = cudaStreamCaptureStatusNone; }
    const unsigned evflag = capst == cudaStreamCaptureStatusNone ? cudaEventRecordDefault : cudaEventRecordExternal;
    if (xchg) {
        // sharded: the thrust block of the PREVIOUS tick leaves for the peers on a side branch, concurrently with this tick
        CK(cudaEventRecord(s->ev_xfork, st));
        CK(cudaStreamWaitEvent(s->stream_xchg, s->ev_xfork, 0));
        launch_exchange(s->shard, s->B, s->d_counter, s->stream_xchg);
        CK(cudaEventRecord(s->ev_xdone, s->stream_xchg));
    }
    const int p_per_stage = io.ekf ? 0 : (io.p ? io.p_per_stage : s->p_resident == 2);
    const size_t psz = sizeof(double) * B * (p_per_stage ? (N + 1) * NP : NP);
    const double *d_x0 = io.x0, *d_yref = io.yref, *d_p = io.p ? io.p : s->d_p, *d_thr = io.thrusts, *d_acc = io.body_acc;
    const int* d_lines = io.lines;
    bool forked = false, chunked = false;
    int nchunk = 4;
    const int pf = host >= 2 ? host - 2 : -1;        // host path variant: the reference window is resident in d_yref_pf[pf]
    if (host) {
        if (io.ekf) {
            CK(cudaMemcpyAsync(s->d_x0, io.x0, sizeof(double) * B * NX, cudaMemcpyHostToDevice, st));
            CK(cudaMemcpyAsync(s->d_thr, io.thrusts, sizeof(double) * B * 6, cudaMemcpyHostToDevice, st));
            CK(cudaMemcpyAsync(s->d_acc, io.body_acc, sizeof(double) * B * 6, cudaMemcpyHostToDevice, st));
            d_thr = s->d_thr; d_acc = s->d_acc;
        } else {
            // (p == NULL: the parameters of the last call that supplied them are still in d_p -- they persist like the capsule's after
            // bluerov2_acados_update_params)
            if (io.p) CK(cudaMemcpyAsync(s->d_p, io.p, psz, cudaMemcpyHostToDevice, st));
            d_p = s->d_p;
            if (io.yref && io.p) CK(cudaEventRecordWithFlags(s->ev_up_p, st, evflag));      // (the prefetch of the next reference queues behind this)
        }
        // explicit reference windows are the big upload (B x (N+1) x 16 doubles: 21.5 MB at B = 4096, N = 40 -- longer on the wire
        // than the kernels of the tick): for large batches without a filter it is cut into four instance ranges and
        // pipelined, range c being linearised and solved while ranges c+1.. are still in flight (below)
        static const int kChunks = [] { const char* e = getenv("BR2_YREF_CHUNKS"); int v = e ? atoi(e) : 4; return v == 1 || v == 2 || v == 4 ? v : 4; }();
        nchunk = kChunks;
        chunked = !io.lines && !io.ekf && B >= 1024 && nchunk > 1 && pf < 0;
        if (io.lines) { CK(cudaMemcpyAsync(s->d_lines, io.lines, sizeof(int) * B, cudaMemcpyHostToDevice, st)); d_lines = s->d_lines; }
        else if (!chunked && pf < 0) { CK(cudaMemcpyAsync(s->d_yref, io.yref, sizeof(double) * B * (N + 1) * NY, cudaMemcpyHostToDevice, st)); }
        if (!io.lines) d_yref = pf >= 0 ? s->d_yref_pf[pf] : s->d_yref;
        d_x0 = s->d_x0;
        if (!io.ekf) {
            // x0 rides a second stream past the lineariser: the QP kernel is its first reader
            if (chunked) launch_tick_begin(s->d_counter, st);    // the ranges' kernels do no housekeeping of their own
            CK(cudaEventRecord(s->ev_fork, st));
            CK(cudaStreamWaitEvent(s->stream_x0, s->ev_fork, 0));
            CK(cudaMemcpyAsync(s->d_x0, io.x0, sizeof(double) * B * NX, cudaMemcpyHostToDevice, s->stream_x0));
            if (io.yref) CK(cudaEventRecordWithFlags(s->ev_up_x0, s->stream_x0, evflag));     // (before ev_x0: that one joins the capture)
            CK(cudaEventRecord(s->ev_x0, s->stream_x0));
            forked = true;
            if (chunked) {
                const size_t row = (N + 1) * NY;
                for (int c = 0; c < nchunk; c++) {
                    const size_t lo = B * c / nchunk, hi = B * (c + 1) / nchunk;
                    CK(cudaMemcpyAsync(s->d_yref + lo * row, io.yref + lo * row, sizeof(double) * (hi - lo) * row, cudaMemcpyHostToDevice, s->stream_x0));
                    CK(cudaEventRecord(s->ev_chunk[c], s->stream_x0));
                }
            }
        }
    }
    // outputs: device buffers, or -- host path -- the caller's buffers themselves when the device can write them (mapped)
    double* o_u0 = host ? nullptr : io.u0;
    double* o_th = host ? nullptr : io.thrust;
    int* o_st = host ? nullptr : io.status;
    if (host) { o_u0 = mapped_alias(io.u0); o_th = mapped_alias(io.thrust); o_st = mapped_alias(io.status); }
    if (io.ekf) {
        EkfArgs e;
        e.B = s->B; e.esti_x = s->d_ex; e.esti_P = s->d_eP; e.thrusts = d_thr; e.meas = d_x0; e.body_acc = d_acc;
        e.wf_dist = host ? s->d_wf : io.wf_dist; e.p_out = s->d_pout; e.compensate = io.compensate; e.model = s->ekf_model;
        launch_ekf(e, st);
        if (io.ekf == 2) {
            RlsArgs r;
            r.B = s->B; r.state = s->d_rls; r.esti_x = s->d_ex; r.body_acc = d_acc; r.meas = d_x0; r.p_out = s->d_pout;
            r.compensate = io.compensate;
            launch_rls(r, st);
        }
        d_p = s->d_pout;
    }
    SolveArgs a;
    fill_args(s, a, d_x0, d_yref, d_lines, d_p, p_per_stage, o_u0, o_th, o_st);
    CK(cudaEventRecordWithFlags(s->ev0, st, evflag));
    if (chunked) {
        // range c on its own stream: waits for its part of the upload (and for x0), then lineariser -> pdas_kernel; the ranges
        // overlap each other and the rest of the upload (a range alone leaves most of the chip idle: its time is the latency of
        // one instance, not its share of the batch)
        for (int c = 0; c < nchunk; c++) {
            SolveArgs ac = a;
            ac.lo = (int)(B * c / nchunk); ac.hi = (int)(B * (c + 1) / nchunk); ac.housekeeping = 0; ac.qidx = c;
            cudaStream_t sc = s->stream_c[c];
            CK(cudaStreamWaitEvent(sc, s->ev_chunk[c], 0));      // (recorded after ev_x0 and after the fork: joins the capture)
            launch_linearize(ac, sc);
            launch_pdas(ac, s->sm_count, sc);
            CK(cudaEventRecord(s->ev_done[c], sc));
            CK(cudaStreamWaitEvent(st, s->ev_done[c], 0));
        }
        if (s->kernel_timing) CK(cudaEventRecordWithFlags(s->ev_mid, st, evflag));     // (no lineariser / QP split in this mode)
        launch_ipm_fallback(a, s->sm_count, st);
    } else {
        launch_linearize(a, st);
        if (s->kernel_timing) CK(cudaEventRecordWithFlags(s->ev_mid, st, evflag));
        if (forked) CK(cudaStreamWaitEvent(st, s->ev_x0, 0));
        launch_ipm(a, s->sm_count, st);
    }
    CK(cudaEventRecordWithFlags(s->ev1, st, evflag));
    if (!host && io.plant_h > 0) {
        PlantArgs pl;
        pl.B = s->B; pl.x = const_cast<double*>(io.x0); pl.u = a.u0; pl.p = d_p; pl.dist = nullptr;
        pl.wave_amp = io.wave_amp; pl.wave_tau0 = io.wave_tau0; pl.body_acc = io.body_acc; pl.lines = const_cast<int*>(io.lines);
        pl.h = io.plant_h; pl.tick = -1; pl.tick_ctr = s->d_counter + CTR_TICK; pl.table = nullptr; pl.table_rows = 0; pl.table_phase = nullptr;
        launch_plant(pl, st);
    }
    if (host) {
        if (io.u0 && !o_u0) CK(cudaMemcpyAsync(io.u0, s->d_u0, sizeof(double) * B * 4, cudaMemcpyDeviceToHost, st));
        if (io.thrust && !o_th) CK(cudaMemcpyAsync(io.thrust, s->d_thrust, sizeof(double) * B * 6, cudaMemcpyDeviceToHost, st));
        if (io.status && !o_st) CK(cudaMemcpyAsync(io.status, s->d_status, sizeof(int) * B, cudaMemcpyDeviceToHost, st));
        if (io.ekf && io.wf_dist) CK(cudaMemcpyAsync(io.wf_dist, s->d_wf, sizeof(double) * B * 6, cudaMemcpyDeviceToHost, st));
    }
    if (xchg) CK(cudaStreamWaitEvent(st, s->ev_xdone, 0));
    s->timed = true;
    CK(cudaGetLastError());
    return BR2_OK;
}

static int tick_validate(br2_batch_solver* s, const br2_tick_io* io, int host, const char* who)
{
    if (!s || !io) return fail(BR2_EINVAL, "%s: null argument", who);
    if (!io->x0) return fail(BR2_EINVAL, "%s: x0 is NULL", who);
    if (!io->yref == !io->lines) return fail(BR2_EINVAL, "%s: exactly one of yref / lines must be given", who);
    if (io->lines && !s->d_traj) return fail(BR2_EINVAL, "%s: lines given but no trajectory set (br2_batch_set_trajectory)", who);
    if (io->ekf < 0 || io->ekf > 2) return fail(BR2_EINVAL, "%s: ekf = %d", who, io->ekf);
    if (io->ekf && (!io->thrusts || !io->body_acc)) return fail(BR2_EINVAL, "%s: ekf needs thrusts and body_acc", who);
    if (!io->ekf && !io->p && !s->p_resident)
        return fail(BR2_EINVAL, "%s: p is NULL, no filter produces it and no earlier host call has supplied the parameters", who);
    if (host && io->plant_h > 0) return fail(BR2_EINVAL, "%s: the plant step is a device-path option", who);
    if ((io->wave_amp == nullptr) != (io->wave_tau0 == nullptr)) return fail(BR2_EINVAL, "%s: wave_amp and wave_tau0 go together", who);
    return BR2_OK;
}

// the host INPUT buffers of a tick (index = H2DNode::which) and their sizes: a host graph can be replayed on other input buffers
// by re-pointing its upload nodes (cudaGraphExecMemcpyNodeSetParams1D), everything else of the key must match
static int tick_inputs(const br2_batch_solver* s, const br2_tick_io& io, const void* ptr[5], size_t bytes[5])
{
    const size_t B = s->B, N = s->N;
    ptr[0] = io.x0; bytes[0] = sizeof(double) * B * NX;
    ptr[1] = io.yref; bytes[1] = sizeof(double) * B * (N + 1) * NY;
    ptr[2] = io.lines; bytes[2] = sizeof(int) * B;
    ptr[3] = io.ekf ? (const void*)io.thrusts : (const void*)io.p; bytes[3] = io.ekf ? sizeof(double) * B * 6 : sizeof(double) * B * (io.p_per_stage ? (N + 1) * NP : NP);
    ptr[4] = io.ekf ? io.body_acc : nullptr; bytes[4] = sizeof(double) * B * 6;
    return 5;
}
static bool same_but_inputs(const br2_tick_io& a, const br2_tick_io& b)
{
    if ((a.yref == nullptr) != (b.yref == nullptr) || (a.lines == nullptr) != (b.lines == nullptr) || (a.p == nullptr) != (b.p == nullptr)) return false;
    br2_tick_io x = a, y = b;
    x.x0 = y.x0 = nullptr; x.yref = y.yref = nullptr; x.lines = y.lines = nullptr;
    if (a.ekf) { x.thrusts = y.thrusts = nullptr; x.body_acc = y.body_acc = nullptr; } else { x.p = y.p = nullptr; }
    return !memcmp(&x, &y, sizeof x);
}

// cached graph of this (io, host) or a newly captured one; nullptr (and rc == BR2_OK) when graphs are not to be used
static int tick_graph_for(br2_batch_solver* s, const br2_tick_io& io, int host, int xchg, cudaGraphExec_t* out)
{
    *out = nullptr;
    if (!s->tick_graph) return BR2_OK;
    br2_batch_solver::TickGraph* slot = &s->tg[0];
    for (auto& g : s->tg) {
        if (g.exec && g.host == host && g.xchg == xchg && g.gen == s->gen && !memcmp(&g.io, &io, sizeof io)) {
            g.stamp = ++s->tick_stamp;
            *out = g.exec;
            return BR2_OK;
        }
        if (!g.exec) { if (slot->exec) slot = &g; }
        else if (slot->exec && g.stamp < slot->stamp) slot = &g;      // least recently used
    }
    // a graph is worth building only for a caller that comes back with the same buffers: the first sighting of a key goes the
    // stream path and is remembered (the last four: double-buffered callers alternate), the second one instantiates -- so a
    // caller that cycles through many distinct buffers never thrashes the cache
    if (host) {
        // same tick on other input buffers: re-point the upload nodes of a cached graph
        for (auto& g : s->tg) {
            if (!(g.exec && g.host == 1 && g.xchg == xchg && g.gen == s->gen && g.nup > 0 && same_but_inputs(g.io, io))) continue;
            const void* np_[5]; size_t nb_[5];
            tick_inputs(s, io, np_, nb_);
            bool okk = true;
            for (int i = 0; i < g.nup && okk; i++) {
                const auto& u = g.up[i];
                okk = cudaGraphExecMemcpyNodeSetParams1D(g.exec, u.node, u.dst, (const char*)np_[u.which] + u.off, u.bytes, cudaMemcpyHostToDevice) == cudaSuccess;
            }
            if (!okk) { cudaGetLastError(); break; }
            g.io = io; g.stamp = ++s->tick_stamp;
            s->graph_updates++;
            *out = g.exec;
            return BR2_OK;
        }
    }
    bool seen = false;
    for (auto& m : s->miss)
        if (m.valid && m.host == host && m.xchg == xchg && m.gen == s->gen && (host ? same_but_inputs(m.io, io) : !memcmp(&m.io, &io, sizeof io))) { seen = true; m.valid = false; }
    if (!seen) {
        auto& m = s->miss[s->miss_next++ % 4];
        m.io = io; m.host = host; m.xchg = xchg; m.gen = s->gen; m.valid = true;
        return BR2_OK;
    }
    cudaGraph_t graph = nullptr;
    CK(cudaStreamBeginCapture(s->stream, cudaStreamCaptureModeThreadLocal));
    int rc = tick_issue(s, io, host, s->stream, xchg);
    cudaError_t e = cudaStreamEndCapture(s->stream, &graph);
    if (rc != BR2_OK || e != cudaSuccess || !graph) {
        if (graph) cudaGraphDestroy(graph);
        cudaGetLastError();
        return rc != BR2_OK ? rc : fail(BR2_ECUDA, "tick graph capture failed: %s", cudaGetErrorString(e));
    }
    cudaGraphExec_t exec = nullptr;
    e = cudaGraphInstantiate(&exec, graph, 0);
    if (e != cudaSuccess) { cudaGraphDestroy(graph); cudaGetLastError(); return fail(BR2_ECUDA, "cudaGraphInstantiate failed: %s", cudaGetErrorString(e)); }
    if (slot->exec) cudaGraphExecDestroy(slot->exec);
    if (slot->graph) cudaGraphDestroy(slot->graph);
    slot->io = io; slot->host = host; slot->xchg = xchg; slot->gen = s->gen; slot->stamp = ++s->tick_stamp; slot->exec = exec; slot->graph = graph; slot->nup = 0;
    if (host) {
        // remember the upload nodes: which input each reads and where inside it
        const void* ip[5]; size_t ib[5];
        tick_inputs(s, io, ip, ib);
        cudaGraphNode_t nodes[64];
        size_t nn = 64;
        bool complete = cudaGraphGetNodes(graph, nodes, &nn) == cudaSuccess && nn <= 64;
        for (size_t i = 0; complete && i < nn; i++) {
            cudaGraphNodeType ty;
            if (cudaGraphNodeGetType(nodes[i], &ty) != cudaSuccess) { complete = false; break; }
            if (ty != cudaGraphNodeTypeMemcpy) continue;
            cudaMemcpy3DParms mp;
            if (cudaGraphMemcpyNodeGetParams(nodes[i], &mp) != cudaSuccess) { complete = false; break; }
            if (mp.kind != cudaMemcpyHostToDevice) continue;
            const char* src = (const char*)mp.srcPtr.ptr;
            const size_t bytes = mp.extent.width * mp.extent.height * mp.extent.depth;
            int which = -1;
            for (int k = 0; k < 5; k++)
                if (ip[k] && src >= (const char*)ip[k] && src + bytes <= (const char*)ip[k] + ib[k]) which = k;
            if (which < 0 || slot->nup >= 12) { complete = false; break; }
            slot->up[slot->nup++] = {nodes[i], which, (size_t)(src - (const char*)ip[which]), bytes, mp.dstPtr.ptr};
        }
        if (!complete) { cudaGetLastError(); slot->nup = 0; }      // this graph is replayed on its own buffers only
    }
    s->graphs_built++;
    *out = exec;
    return BR2_OK;
}

// ---- sharding: peer-to-peer exchange of the thrust vectors (kernel side: finish_instance in kernels.cu) -----------------------
extern "C" int br2_batch_shard_init(br2_batch_solver* s, int rank, int world)
{
    if (!s) return fail(BR2_EINVAL, "null solver");
    if (world < 1 || world > MAX_SHARDS || rank < 0 || rank >= world) return fail(BR2_EINVAL, "br2_batch_shard_init: rank %d of %d (max %d)", rank, world, MAX_SHARDS);
    if (s->d_shard) return fail(BR2_EINVAL, "br2_batch_shard_init: already initialised");
    ON_DEVICE(s);
    const size_t nd = (size_t)2 * world * s->B * 6;
    const size_t bytes = nd * sizeof(double) + (MAX_SHARDS + 8) * sizeof(int);
    CK(cudaMalloc((void**)&s->d_shard, bytes));
    CK(cudaMemset(s->d_shard, 0, bytes));
    memset(&s->shard, 0, sizeof s->shard);
    s->shard.world = world; s->shard.rank = rank;
    s->shard.buf[rank] = s->d_shard;
    s->shard.flag[rank] = (int*)(s->d_shard + nd);
    s->gen++;
    return BR2_OK;
}

extern "C" int br2_batch_shard_handle(br2_batch_solver* s, void* handle64)
{
    if (!s || !handle64 || !s->d_shard) return fail(BR2_EINVAL, "br2_batch_shard_handle: not initialised");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    ON_DEVICE(s);
    CK(cudaIpcGetMemHandle((cudaIpcMemHandle_t*)handle64, s->d_shard));
    return BR2_OK;
}

extern "C" int br2_batch_shard_connect(br2_batch_solver* s, int peer, const void* handle64)
{
    if (!s || !handle64 || !s->d_shard) return fail(BR2_EINVAL, "br2_batch_shard_connect: not initialised");
    if (peer < 0 || peer >= s->shard.world || peer == s->shard.rank) return fail(BR2_EINVAL, "br2_batch_shard_connect: peer %d", peer);
    if (s->peer_base[peer]) return fail(BR2_EINVAL, "br2_batch_shard_connect: peer %d already connected", peer);
    ON_DEVICE(s);
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, sizeof h);
    void* base = nullptr;
    CK(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
    s->peer_base[peer] = base;
    const size_t nd = (size_t)2 * s->shard.world * s->B * 6;
    s->shard.buf[peer] = (double*)base;
    s->shard.flag[peer] = (int*)((double*)base + nd);
    s->gen++;
    return BR2_OK;
}

extern "C" int br2_batch_shard_gathered(br2_batch_solver* s, int parity, double** d_buf)
{
    if (!s || !d_buf || !s->d_shard) return fail(BR2_EINVAL, "br2_batch_shard_gathered: not initialised");
    *d_buf = s->d_shard + (size_t)(parity & 1) * s->shard.world * s->B * 6;
    return BR2_OK;
}

extern "C" int br2_batch_tick_count(br2_batch_solver* s) { return s ? (int)s->ticks_host : 0; }

extern "C" int br2_batch_shard_wait(br2_batch_solver* s, void* stream)
{
    if (!s || !s->d_shard) return fail(BR2_EINVAL, "br2_batch_shard_wait: not initialised");
    for (int r = 0; r < s->shard.world; r++)
        if (!s->shard.buf[r]) return fail(BR2_EINVAL, "br2_batch_shard_wait: rank %d is not connected", r);
    ON_DEVICE(s);
    const size_t nd = (size_t)2 * s->shard.world * s->B * 6;
    int* flags = (int*)(s->d_shard + nd);
    while (s->xchg_host < s->ticks_host) {       // (normally one: the last tick's block)
        launch_exchange(s->shard, s->B, s->d_counter, (cudaStream_t)stream);
        s->xchg_host++;
    }
    launch_shard_wait(flags, s->shard.world, (int)s->ticks_host, flags + MAX_SHARDS, (cudaStream_t)stream);
    CK(cudaGetLastError());
    return BR2_OK;
}

extern "C" int br2_batch_graphs_built(const br2_batch_solver* s) { return s ? s->graphs_built : 0; }
extern "C" int br2_batch_graph_updates(const br2_batch_solver* s) { return s ? s->graph_updates : 0; }

extern "C" int br2_batch_set_tick_index(br2_batch_solver* s, int next_tick)
{
    if (!s) return fail(BR2_EINVAL, "null solver");
    ON_DEVICE(s);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(s->d_counter + CTR_TICK, &next_tick, sizeof(int), cudaMemcpyHostToDevice));
    s->ticks_host = next_tick;
    s->xchg_host = next_tick;
    CK(cudaMemcpy(s->d_counter + CTR_XTICK, &next_tick, sizeof(int), cudaMemcpyHostToDevice));
    // sharded: the local flag array restarts with the tick index (every rank does this between two barriers)
    if (s->d_shard) CK(cudaMemset(s->d_shard + (size_t)2 * s->shard.world * s->B * 6, 0, (MAX_SHARDS + 8) * sizeof(int)));
    return BR2_OK;
}

extern "C" int br2_batch_tick_device(br2_batch_solver* s, const br2_tick_io* io_in, void* stream)
{
    int rc = tick_validate(s, io_in, 0, "br2_batch_tick_device");
    if (rc != BR2_OK) return rc;
    ON_DEVICE(s);
    br2_tick_io io;
    memset(&io, 0, sizeof io);                          // (padding bytes take part in the cache comparison)
    io.x0 = io_in->x0; io.yref = io_in->yref; io.p = io_in->p; io.thrusts = io_in->thrusts; io.lines = io_in->lines;
    io.body_acc = io_in->body_acc; io.u0 = io_in->u0; io.thrust = io_in->thrust; io.wf_dist = io_in->wf_dist; io.status = io_in->status;
    io.wave_amp = io_in->wave_amp; io.wave_tau0 = io_in->wave_tau0; io.plant_h = io_in->plant_h;
    io.p_per_stage = io_in->p_per_stage; io.ekf = io_in->ekf; io.compensate = io_in->compensate;
    cudaStream_t st = (cudaStream_t)stream;
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cap) != cudaSuccess) { cudaGetLastError(); cap = cudaStreamCaptureStatusNone; }
    cudaGraphExec_t exec = nullptr;
    const int xchg = shard_pending(s);
    if (cap == cudaStreamCaptureStatusNone) {
        rc = tick_graph_for(s, io, 0, xchg, &exec);
        if (rc != BR2_OK) return rc;
    }
    s->ticks_host++;
    s->xchg_host += xchg;
    if (exec) { CK(cudaGraphLaunch(exec, st)); s->timed = true; return BR2_OK; }
    return tick_issue(s, io, 0, st, xchg);                    // the caller is capturing its own graph, or graphs are switched off
}

// Pinned host memory for the host entry points.  write_combined != 0: write-combined pages -- the CPU writes them (streaming), only the
// copy engine reads them, without snooping the CPU caches on the way: the choice for large per-tick inputs such as reference windows.
extern "C" int br2_host_alloc(void** out, size_t bytes, int write_combined)
{
    if (!out || !bytes) return fail(BR2_EINVAL, "br2_host_alloc: bad argument");
    CK(cudaHostAlloc(out, bytes, cudaHostAllocPortable | cudaHostAllocMapped | (write_combined ? cudaHostAllocWriteCombined : 0)));
    return BR2_OK;
}
extern "C" int br2_host_free(void* p)
{
    if (p) CK(cudaFreeHost(p));
    return BR2_OK;
}

extern "C" int br2_batch_set_next_yref_host(br2_batch_solver* s, const double* yref_next)
{
    if (!s) return fail(BR2_EINVAL, "null solver");
    if (yref_next && !pinned_host(yref_next)) return fail(BR2_EINVAL, "br2_batch_set_next_yref_host: the buffer must be pinned host memory");
    s->next_yref = yref_next;
    return BR2_OK;
}

extern "C" int br2_batch_tick_host(br2_batch_solver* s, const br2_tick_io* io_in)
{
    int rc = tick_validate(s, io_in, 1, "br2_batch_tick_host");
    if (rc != BR2_OK) return rc;
    ON_DEVICE(s);
    br2_tick_io io;
    memset(&io, 0, sizeof io);
    io.x0 = io_in->x0; io.yref = io_in->yref; io.p = io_in->p; io.thrusts = io_in->thrusts; io.lines = io_in->lines;
    io.body_acc = io_in->body_acc; io.u0 = io_in->u0; io.thrust = io_in->thrust; io.wf_dist = io_in->wf_dist; io.status = io_in->status;
    io.p_per_stage = io_in->p_per_stage; io.ekf = io_in->ekf; io.compensate = io_in->compensate;
    cudaGraphExec_t exec = nullptr;
    // copies from / to pageable memory are staged by the driver at enqueue time: only pinned buffers go into a graph
    const bool pinned = pinned_host(io.x0) && pinned_host(io.yref) && pinned_host(io.p) && pinned_host(io.thrusts) && pinned_host(io.lines) &&
                        pinned_host(io.body_acc) && pinned_host(io.u0) && pinned_host(io.thrust) && pinned_host(io.wf_dist) && pinned_host(io.status);
    if (io.p && !io.ekf) {
        const int mode = io.p_per_stage ? 2 : 1;
        if (s->p_resident && s->p_resident != mode) s->gen++;      // (graphs that rely on the resident layout are stale)
        s->p_resident = mode;
    }
    const int xchg = shard_pending(s);
    // explicit reference registered one tick ahead (br2_batch_set_next_yref_host): it is resident (or arriving) in d_yref_pf[pf_slot]
    const int pf = (io.yref && !io.ekf && io.yref == s->pf_host) ? s->pf_slot : -1;
    const int variant = pf >= 0 ? 2 + pf : 1;
    if (pf >= 0) CK(cudaStreamWaitEvent(s->stream, s->ev_pf, 0));
    if (pinned) {
        rc = tick_graph_for(s, io, variant, xchg, &exec);
        if (rc != BR2_OK) return rc;
    }
    s->ticks_host++;
    s->xchg_host += xchg;
    if (exec) { CK(cudaGraphLaunch(exec, s->stream)); s->timed = true; }
    else {
        rc = tick_issue(s, io, variant, s->stream, xchg);
        if (rc != BR2_OK) return rc;
    }
    if (s->next_yref) {
        // upload the NEXT tick's reference window now: behind this tick's own uploads on the wire (its p and x0 copies carry external
        // events), beside its kernels, into the buffer this tick does not read
        const int slot = pf >= 0 ? 1 - pf : 0;
        const size_t bytes = sizeof(double) * (size_t)s->B * (s->N + 1) * NY;
        if (!s->d_yref_pf[slot]) CK(cudaMalloc((void**)&s->d_yref_pf[slot], bytes));
        if (!io.ekf) {
            CK(cudaStreamWaitEvent(s->stream_pf, s->ev_up_p, 0));
            CK(cudaStreamWaitEvent(s->stream_pf, s->ev_up_x0, 0));
        }
        CK(cudaMemcpyAsync(s->d_yref_pf[slot], s->next_yref, bytes, cudaMemcpyHostToDevice, s->stream_pf));
        CK(cudaEventRecord(s->ev_pf, s->stream_pf));
        s->pf_host = s->next_yref; s->pf_slot = slot; s->next_yref = nullptr;
    } else if (pf >= 0) {
        s->pf_host = nullptr;                      // consumed
    }
    CK(cudaStreamSynchronize(s->stream));
    return BR2_OK;
}

extern "C" long long br2_batch_nonzero_status_total(br2_batch_solver* s, int reset)
{
    if (!s) return -1;
    DeviceGuard guard_(s->device);
    if (cudaDeviceSynchronize() != cudaSuccess) return -1;
    unsigned long long v = 0;
    if (cudaMemcpy(&v, s->d_iter_total + 17, sizeof v, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    if (reset) cudaMemset(s->d_iter_total + 17, 0, sizeof v);
    return (long long)v;
}

// instrumentation build (-DBR2_PROFILE): cycles per kernel phase summed over warps since the last reset; zeros otherwise
extern "C" int br2_batch_phase_cycles(br2_batch_solver* s, unsigned long long* out16, int reset)
{
    if (!s || !out16) return fail(BR2_EINVAL, "null argument");
    ON_DEVICE(s);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(out16, s->d_iter_total + 1, sizeof(unsigned long long) * 16, cudaMemcpyDeviceToHost));
    if (reset) CK(cudaMemset(s->d_iter_total + 1, 0, sizeof(unsigned long long) * 16));
    return BR2_OK;
}

extern "C" int br2_batch_ekf_phase_cycles(br2_batch_solver* s, unsigned long long* out12, int reset)
{
    if (!s || !out12) return fail(BR2_EINVAL, "null argument");
    ON_DEVICE(s);
    CK(cudaDeviceSynchronize());
    ekf_phase_cycles(out12, reset);
    CK(cudaGetLastError());
    return BR2_OK;
}

// ---- nominal plant (closed-loop studies on the device) ----
extern "C" int br2_plant_step_device(int batch, double* d_x, const double* d_u, const double* d_p, const double* d_dist,
                                     const double* d_wave_amp, const double* d_wave_tau0, int tick, double h,
                                     double* d_body_acc, int* d_lines, void* stream)
{
    if (batch < 1 || !d_x || !d_u || !d_p || !(h > 0)) return fail(BR2_EINVAL, "br2_plant_step_device: bad argument");
    if ((d_wave_amp == nullptr) != (d_wave_tau0 == nullptr)) return fail(BR2_EINVAL, "br2_plant_step_device: wave_amp and wave_tau0 go together");
    // the launch goes to the device that owns the state vector, whatever the caller's current device is
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, d_x) != cudaSuccess || (at.type != cudaMemoryTypeDevice && at.type != cudaMemoryTypeManaged)) {
        cudaGetLastError();
        return fail(BR2_EINVAL, "br2_plant_step_device: d_x is not a device pointer");
    }
    DeviceGuard guard_(at.device);
    CK(guard_.err);
    PlantArgs a;
    a.B = batch; a.x = d_x; a.u = d_u; a.p = d_p; a.dist = d_dist; a.wave_amp = d_wave_amp; a.wave_tau0 = d_wave_tau0;
    a.body_acc = d_body_acc; a.lines = d_lines; a.h = h; a.tick = tick; a.tick_ctr = nullptr; a.table = nullptr; a.table_rows = 0; a.table_phase = nullptr;
    launch_plant(a, (cudaStream_t)stream);
    CK(cudaGetLastError());
    return BR2_OK;
}

extern "C" int br2_plant_step_replay_device(int batch, double* d_x, const double* d_u, const double* d_p, const double* d_table, int rows,
                                            const int* d_phase, int tick, double h, double* d_body_acc, int* d_lines, void* stream)
{
    if (batch < 1 || !d_x || !d_u || !d_p || !d_table || rows < 1 || tick < 0 || !(h > 0)) return fail(BR2_EINVAL, "br2_plant_step_replay_device: bad argument");
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, d_x) != cudaSuccess || (at.type != cudaMemoryTypeDevice && at.type != cudaMemoryTypeManaged)) {
        cudaGetLastError();
        return fail(BR2_EINVAL, "br2_plant_step_replay_device: d_x is not a device pointer");
    }
    DeviceGuard guard_(at.device);
    CK(guard_.err);
    PlantArgs a;
    a.B = batch; a.x = d_x; a.u = d_u; a.p = d_p; a.dist = nullptr; a.wave_amp = nullptr; a.wave_tau0 = nullptr;
    a.table = d_table; a.table_rows = rows; a.table_phase = d_phase;
    a.body_acc = d_body_acc; a.lines = d_lines; a.h = h; a.tick = tick; a.tick_ctr = nullptr;
    launch_plant(a, (cudaStream_t)stream);
    CK(cudaGetLastError());
    return BR2_OK;
}

// ---- EKF ------------------------------------------------------------------------------------------------
extern "C" int br2_batch_ekf_reset(br2_batch_solver* s)
{
    if (!s) return fail(BR2_EINVAL, "null solver");
    ON_DEVICE(s);
    const size_t B = s->B;
    double* hx = (double*)calloc(B * 18, sizeof(double));
    double* hP = (double*)calloc(B * 324, sizeof(double));
    if (!hx || !hP) { free(hx); free(hP); return fail(BR2_ENOMEM, "out of host memory"); }
    for (size_t i = 0; i < B; i++) {
        hx[i * 18 + 2] = -20.0;                                  // esti_x, bluerov2_dob.cpp:64
        hx[i * 18 + 12] = hx[i * 18 + 13] = hx[i * 18 + 14] = 6.0;
        for (int j = 0; j < 18; j++) hP[i * 324 + j * 19] = 1.0;  // P0 = I, bluerov2_dob.h:203
    }
    cudaError_t e1 = cudaMemcpy(s->d_ex, hx, sizeof(double) * B * 18, cudaMemcpyHostToDevice);
    cudaError_t e2 = cudaMemcpy(s->d_eP, hP, sizeof(double) * B * 324, cudaMemcpyHostToDevice);
    free(hx); free(hP);
    CK(e1); CK(e2);
    return BR2_OK;
}

extern "C" int br2_batch_ekf_device(br2_batch_solver* s, const double* d_thrusts, const double* d_meas,
                                    const double* d_body_acc, double* d_wf_dist, double* d_p_out, int compensate, void* stream)
{
    if (!s || !d_thrusts || !d_meas || !d_body_acc) return fail(BR2_EINVAL, "br2_batch_ekf_device: null argument");
    ON_DEVICE(s);
    EkfArgs a;
    a.B = s->B; a.esti_x = s->d_ex; a.esti_P = s->d_eP;
    a.thrusts = d_thrusts; a.meas = d_meas; a.body_acc = d_body_acc;
    a.wf_dist = d_wf_dist; a.p_out = d_p_out; a.compensate = compensate; a.model = s->ekf_model;
    launch_ekf(a, (cudaStream_t)stream);
    CK(cudaGetLastError());
    return BR2_OK;
}

extern "C" int br2_batch_ekf_host(br2_batch_solver* s, const double* thrusts, const double* meas, const double* body_acc,
                                  double* wf_dist, double* p_out, int compensate)
{
    if (!s || !thrusts || !meas || !body_acc) return fail(BR2_EINVAL, "br2_batch_ekf_host: null argument");
    ON_DEVICE(s);
    const size_t B = s->B;
    cudaStream_t st = s->stream;
    CK(cudaMemcpyAsync(s->d_thr, thrusts, sizeof(double) * B * 6, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(s->d_meas, meas, sizeof(double) * B * 12, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(s->d_acc, body_acc, sizeof(double) * B * 6, cudaMemcpyHostToDevice, st));
    int rc = br2_batch_ekf_device(s, s->d_thr, s->d_meas, s->d_acc, s->d_wf, s->d_pout, compensate, st);
    if (rc != BR2_OK) return rc;
    if (wf_dist) CK(cudaMemcpyAsync(wf_dist, s->d_wf, sizeof(double) * B * 6, cudaMemcpyDeviceToHost, st));
    if (p_out) CK(cudaMemcpyAsync(p_out, s->d_pout, sizeof(double) * B * NP, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return BR2_OK;
}

extern "C" int br2_batch_ekf_get_state_host(br2_batch_solver* s, double* esti_x, double* esti_P)
{
    if (!s) return fail(BR2_EINVAL, "null solver");
    ON_DEVICE(s);
    CK(cudaDeviceSynchronize());
    if (esti_x) CK(cudaMemcpy(esti_x, s->d_ex, sizeof(double) * s->B * 18, cudaMemcpyDeviceToHost));
    if (esti_P) CK(cudaMemcpy(esti_P, s->d_eP, sizeof(double) * s->B * 324, cudaMemcpyDeviceToHost));
    return BR2_OK;
}
extern "C" int br2_batch_ekf_set_state_host(br2_batch_solver* s, const double* esti_x, const double* esti_P)
{
    if (!s) return fail(BR2_EINVAL, "null solver");
    ON_DEVICE(s);
    CK(cudaDeviceSynchronize());
    if (esti_x) CK(cudaMemcpy(s->d_ex, esti_x, sizeof(double) * s->B * 18, cudaMemcpyHostToDevice));
    if (esti_P) CK(cudaMemcpy(s->d_eP, esti_P, sizeof(double) * s->B * 324, cudaMemcpyHostToDevice));
    return BR2_OK;
}

// ---- RLS with variable forgetting factor (AMPC) ---------------------------------------------------------------
extern "C" int br2_batch_rls_reset(br2_batch_solver* s)
{
    if (!s) return fail(BR2_EINVAL, "null solver");
    ON_DEVICE(s);
    const size_t n = (size_t)s->B * 4;
    double* h = (double*)calloc(n * RLS_STRIDE, sizeof(double));
    if (!h) return fail(BR2_ENOMEM, "out of host memory");
    for (size_t i = 0; i < n; i++) {
        for (int j = 0; j < 4; j++) h[i * RLS_STRIDE + RLS_P + j * 5] = 1.0;   // P = I, bluerov2_ampc.cpp:63-68
        h[i * RLS_STRIDE + RLS_LAMBDA] = 0.9;                                    // :70-73
    }
    cudaError_t e = cudaMemcpy(s->d_rls, h, sizeof(double) * n * RLS_STRIDE, cudaMemcpyHostToDevice);
    free(h);
    CK(e);
    return BR2_OK;
}

extern "C" int br2_batch_rls_device(br2_batch_solver* s, const double* d_meas, const double* d_body_acc, double* d_p_out,
                                    int compensate, void* stream)
{
    if (!s || !d_meas || !d_body_acc) return fail(BR2_EINVAL, "br2_batch_rls_device: null argument");
    ON_DEVICE(s);
    RlsArgs a;
    a.B = s->B; a.state = s->d_rls; a.esti_x = s->d_ex; a.body_acc = d_body_acc; a.meas = d_meas; a.p_out = d_p_out;
    a.compensate = compensate;
    launch_rls(a, (cudaStream_t)stream);
    CK(cudaGetLastError());
    return BR2_OK;
}

extern "C" int br2_batch_rls_host(br2_batch_solver* s, const double* meas, const double* body_acc, double* p_out, int compensate)
{
    if (!s || !meas || !body_acc) return fail(BR2_EINVAL, "br2_batch_rls_host: null argument");
    ON_DEVICE(s);
    const size_t B = s->B;
    cudaStream_t st = s->stream;
    CK(cudaMemcpyAsync(s->d_meas, meas, sizeof(double) * B * 12, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(s->d_acc, body_acc, sizeof(double) * B * 6, cudaMemcpyHostToDevice, st));
    // p_out is in/out on the host: with compensate == 0 the reference leaves p[4..15] as they were
    if (p_out) CK(cudaMemcpyAsync(s->d_pout, p_out, sizeof(double) * B * NP, cudaMemcpyHostToDevice, st));
    int rc = br2_batch_rls_device(s, s->d_meas, s->d_acc, p_out ? s->d_pout : nullptr, compensate, st);
    if (rc != BR2_OK) return rc;
    if (p_out) CK(cudaMemcpyAsync(p_out, s->d_pout, sizeof(double) * B * NP, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return BR2_OK;
}

extern "C" int br2_batch_rls_get_state_host(br2_batch_solver* s, double* state)
{
    if (!s || !state) return fail(BR2_EINVAL, "null argument");
    ON_DEVICE(s);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(state, s->d_rls, sizeof(double) * s->B * 4 * RLS_STRIDE, cudaMemcpyDeviceToHost));
    return BR2_OK;
}
extern "C" int br2_batch_rls_set_state_host(br2_batch_solver* s, const double* state)
{
    if (!s || !state) return fail(BR2_EINVAL, "null argument");
    ON_DEVICE(s);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(s->d_rls, state, sizeof(double) * s->B * 4 * RLS_STRIDE, cudaMemcpyHostToDevice));
    return BR2_OK;
}

// ---- continuous yaw (node glue, bluerov2_dob.cpp:272-304) ------------------------------------------------------
extern "C" int br2_batch_yaw_reset(br2_batch_solver* s)
{
    if (!s) return fail(BR2_EINVAL, "null solver");
    ON_DEVICE(s);
    CK(cudaMemset(s->d_yaw, 0, sizeof(float) * 2 * s->B));      // yaw_sum = pre_yaw = 0 (bluerov2_dob.h:234-235)
    return BR2_OK;
}
extern "C" int br2_batch_yaw_unwrap_device(br2_batch_solver* s, double* d_x0, void* stream)
{
    if (!s || !d_x0) return fail(BR2_EINVAL, "br2_batch_yaw_unwrap_device: null argument");
    ON_DEVICE(s);
    launch_yaw_unwrap(s->B, s->d_yaw, d_x0, (cudaStream_t)stream);
    CK(cudaGetLastError());
    return BR2_OK;
}
extern "C" int br2_batch_yaw_unwrap_host(br2_batch_solver* s, double* x0)
{
    if (!s || !x0) return fail(BR2_EINVAL, "br2_batch_yaw_unwrap_host: null argument");
    ON_DEVICE(s);
    const size_t n = sizeof(double) * s->B * NX;
    CK(cudaMemcpyAsync(s->d_x0, x0, n, cudaMemcpyHostToDevice, s->stream));
    launch_yaw_unwrap(s->B, s->d_yaw, s->d_x0, s->stream);
    CK(cudaMemcpyAsync(x0, s->d_x0, n, cudaMemcpyDeviceToHost, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    return BR2_OK;
}
extern "C" int br2_batch_yaw_get_state_host(br2_batch_solver* s, float* state)
{
    if (!s || !state) return fail(BR2_EINVAL, "null argument");
    ON_DEVICE(s);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(state, s->d_yaw, sizeof(float) * 2 * s->B, cudaMemcpyDeviceToHost));
    return BR2_OK;
}
extern "C" int br2_batch_yaw_set_state_host(br2_batch_solver* s, const float* state)
{
    if (!s || !state) return fail(BR2_EINVAL, "null argument");
    ON_DEVICE(s);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(s->d_yaw, state, sizeof(float) * 2 * s->B, cudaMemcpyHostToDevice));
    return BR2_OK;
}


// ---- IMU error-state Kalman filter (eskf.cu) --------------------------------------------------------------------------------
struct br2_eskf {
    int B, device;
    double *d_state, *d_P, *d_in, *d_out;      // d_in: staging for the host entry points (39 doubles / instance), d_out: 15
    EskfArgs base;
    cudaStream_t stream;
};

extern "C" int br2_eskf_free(br2_eskf* f)
{
    if (!f) return BR2_OK;
    DeviceGuard guard_(f->device);
    for (void* p : {(void*)f->d_state, (void*)f->d_P, (void*)f->d_in, (void*)f->d_out})
        if (p) cudaFree(p);
    if (f->stream) cudaStreamDestroy(f->stream);
    free(f);
    return BR2_OK;
}

extern "C" int br2_eskf_create(br2_eskf** out, int batch, const br2_eskf_params* prm, int device)
{
    if (!out) return fail(BR2_EINVAL, "br2_eskf_create: out is NULL");
    *out = nullptr;
    if (batch < 1) return fail(BR2_EINVAL, "br2_eskf_create: batch = %d", batch);
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(BR2_ECUDA, "br2_eskf_create: no CUDA device (%s); this library has no CPU path", e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
    }
    if (device < 0 || device >= ndev) return fail(BR2_EINVAL, "br2_eskf_create: device %d of %d", device, ndev);
    DeviceGuard guard_(device);
    CK(guard_.err);
    br2_eskf* f = (br2_eskf*)calloc(1, sizeof(br2_eskf));
    if (!f) return fail(BR2_ENOMEM, "out of host memory");
    f->B = batch; f->device = device;
    // launch/config/imudo.yaml
    br2_eskf_params d = {0.001, 0.001, 0.001, 0.001, 0.001, 0.01, 0.02, 0.0006, 0.012,
                         {-4.342596682195816e-07, -3.581072716118436e-18, -0.009999999990570729},
                         {-2.66013609366142e-20, -1.933924486945935e-19, -3.870624673211354e-16}};
    if (prm) d = *prm;
    EskfArgs& a = f->base;
    memset(&a, 0, sizeof a);
    a.B = batch;
    for (int i = 0; i < 21; i++) a.Qd[i] = i < 3 ? d.q_p : i < 6 ? d.q_v : i < 9 ? d.q_r : i < 18 ? d.q_q : d.q_xi;      // Config.cpp:128-139
    for (int i = 0; i < 12; i++) a.Rd[i] = i < 3 ? d.r_p : i < 6 ? d.r_v : i < 9 ? d.r_r : d.r_th;                          // :147-155
    memcpy(a.b_a, d.b_a, sizeof a.b_a); memcpy(a.b_g, d.b_g, sizeof a.b_g);
    const size_t B = batch;
    bool ok = cudaMalloc((void**)&f->d_state, sizeof(double) * B * 18) == cudaSuccess && cudaMalloc((void**)&f->d_P, sizeof(double) * B * 441) == cudaSuccess &&
              cudaMalloc((void**)&f->d_in, sizeof(double) * B * 39) == cudaSuccess && cudaMalloc((void**)&f->d_out, sizeof(double) * B * 15) == cudaSuccess &&
              cudaStreamCreateWithFlags(&f->stream, cudaStreamNonBlocking) == cudaSuccess;
    if (ok) {
        double* h = (double*)calloc(B * 18, sizeof(double));
        ok = h != nullptr;
        if (ok) {
            for (size_t i = 0; i < B; i++) h[i * 18 + 6] = h[i * 18 + 10] = h[i * 18 + 14] = 1.0;      // R = I
            ok = cudaMemcpy(f->d_state, h, sizeof(double) * B * 18, cudaMemcpyHostToDevice) == cudaSuccess &&
                 cudaMemset(f->d_P, 0, sizeof(double) * B * 441) == cudaSuccess;
            free(h);
        }
    }
    if (!ok) { cudaError_t le = cudaGetLastError(); br2_eskf_free(f); return fail(BR2_ENOMEM, "br2_eskf_create: allocation failed (%s)", cudaGetErrorString(le)); }
    a.state = f->d_state; a.P = f->d_P;
    configure_eskf();
    *out = f;
    return BR2_OK;
}

extern "C" int br2_eskf_set_state_host(br2_eskf* f, const double* state, const double* P)
{
    if (!f) return fail(BR2_EINVAL, "null filter");
    DeviceGuard guard_(f->device); CK(guard_.err);
    CK(cudaDeviceSynchronize());
    if (state) CK(cudaMemcpy(f->d_state, state, sizeof(double) * f->B * 18, cudaMemcpyHostToDevice));
    if (P) CK(cudaMemcpy(f->d_P, P, sizeof(double) * f->B * 441, cudaMemcpyHostToDevice));
    return BR2_OK;
}
extern "C" int br2_eskf_get_state_host(br2_eskf* f, double* state, double* P)
{
    if (!f) return fail(BR2_EINVAL, "null filter");
    DeviceGuard guard_(f->device); CK(guard_.err);
    CK(cudaDeviceSynchronize());
    if (state) CK(cudaMemcpy(state, f->d_state, sizeof(double) * f->B * 18, cudaMemcpyDeviceToHost));
    if (P) CK(cudaMemcpy(P, f->d_P, sizeof(double) * f->B * 441, cudaMemcpyDeviceToHost));
    return BR2_OK;
}

extern "C" int br2_eskf_predict_device(br2_eskf* f, const double* d_imu, void* stream)
{
    if (!f || !d_imu) return fail(BR2_EINVAL, "br2_eskf_predict_device: null argument");
    DeviceGuard guard_(f->device); CK(guard_.err);
    EskfArgs a = f->base;
    a.imu = d_imu;
    launch_eskf_predict(a, (cudaStream_t)stream);
    CK(cudaGetLastError());
    return BR2_OK;
}
extern "C" int br2_eskf_update_device(br2_eskf* f, const double* d_gps_p, const double* d_gps_v, const double* d_R_meas, const double* d_thrusts,
                                      const double* d_imu_raw, const double* d_R_gt, double* d_xi_world, double* d_innov, void* stream)
{
    if (!f || !d_gps_p || !d_gps_v || !d_R_meas || !d_thrusts || !d_imu_raw || !d_R_gt) return fail(BR2_EINVAL, "br2_eskf_update_device: null argument");
    DeviceGuard guard_(f->device); CK(guard_.err);
    EskfArgs a = f->base;
    a.gps_p = d_gps_p; a.gps_v = d_gps_v; a.R_meas = d_R_meas; a.thrusts = d_thrusts; a.imu = d_imu_raw; a.R_gt = d_R_gt;
    a.xi_world = d_xi_world; a.innov = d_innov;
    launch_eskf_update(a, (cudaStream_t)stream);
    CK(cudaGetLastError());
    return BR2_OK;
}
extern "C" int br2_eskf_predict_host(br2_eskf* f, const double* imu)
{
    if (!f || !imu) return fail(BR2_EINVAL, "br2_eskf_predict_host: null argument");
    DeviceGuard guard_(f->device); CK(guard_.err);
    CK(cudaMemcpyAsync(f->d_in, imu, sizeof(double) * f->B * 6, cudaMemcpyHostToDevice, f->stream));
    int rc = br2_eskf_predict_device(f, f->d_in, f->stream);
    if (rc != BR2_OK) return rc;
    CK(cudaStreamSynchronize(f->stream));
    return BR2_OK;
}
extern "C" int br2_eskf_update_host(br2_eskf* f, const double* gps_p, const double* gps_v, const double* R_meas, const double* thrusts,
                                    const double* imu_raw, const double* R_gt, double* xi_world, double* innov)
{
    if (!f || !gps_p || !gps_v || !R_meas || !thrusts || !imu_raw || !R_gt) return fail(BR2_EINVAL, "br2_eskf_update_host: null argument");
    DeviceGuard guard_(f->device); CK(guard_.err);
    const size_t B = f->B;
    double* d = f->d_in;        // [gps_p 3 | gps_v 3 | R_meas 9 | thrusts 6 | imu 6 | R_gt 9] x B, section by section
    double *dp = d, *dv = d + 3 * B, *dRm = d + 6 * B, *dth = d + 15 * B, *dimu = d + 21 * B, *dRg = d + 27 * B;
    CK(cudaMemcpyAsync(dp, gps_p, sizeof(double) * B * 3, cudaMemcpyHostToDevice, f->stream));
    CK(cudaMemcpyAsync(dv, gps_v, sizeof(double) * B * 3, cudaMemcpyHostToDevice, f->stream));
    CK(cudaMemcpyAsync(dRm, R_meas, sizeof(double) * B * 9, cudaMemcpyHostToDevice, f->stream));
    CK(cudaMemcpyAsync(dth, thrusts, sizeof(double) * B * 6, cudaMemcpyHostToDevice, f->stream));
    CK(cudaMemcpyAsync(dimu, imu_raw, sizeof(double) * B * 6, cudaMemcpyHostToDevice, f->stream));
    CK(cudaMemcpyAsync(dRg, R_gt, sizeof(double) * B * 9, cudaMemcpyHostToDevice, f->stream));
    int rc = br2_eskf_update_device(f, dp, dv, dRm, dth, dimu, dRg, f->d_out, f->d_out + 3 * B, f->stream);
    if (rc != BR2_OK) return rc;
    if (xi_world) CK(cudaMemcpyAsync(xi_world, f->d_out, sizeof(double) * B * 3, cudaMemcpyDeviceToHost, f->stream));
    if (innov) CK(cudaMemcpyAsync(innov, f->d_out + 3 * B, sizeof(double) * B * 12, cudaMemcpyDeviceToHost, f->stream));
    CK(cudaStreamSynchronize(f->stream));
    return BR2_OK;
}
