/* bluerov2_b200.h -- batched C-ABI of the B200 SQP-RTI engine (extension; not in the reference).
 *
 * The reference solves ONE OCP instance per process through the acados-generated ABI
 * (include/acados_solver_bluerov2.h, mirroring bluerov2_dobmpc/scripts/c_generated_code/acados_solver_bluerov2.h).
 * This header is the batched view of the same engine: B independent instances of the BlueROV2 OCP
 * (nx = 12, nu = 4, np = 16, horizon N) advanced by one SQP-RTI step per call.  Each entry point names the
 * reference call sequence it batches.  Plain C types only: pointers and sizes, no torch / C++ types.
 *
 * Conventions
 *   - all arrays are dense, row-major, double, instance-major: x0[B][12], yref[B][N+1][16] (terminal row: first
 *     12 entries used, like double yref[N+1][NY] in bluerov2_dob.h:68-71), p[B][16] (or [B][N+1][16] when
 *     p_per_stage != 0), u0[B][4], thrust[B][6], status[B] (acados codes: 0 success, 1 NaN, 2 max iter, 4 QP failure).
 *   - "_device" entry points take device pointers valid on the solver's GPU and a cudaStream_t passed as void*;
 *     they only enqueue work.  "_host" entry points copy in, solve, copy out and synchronise.
 *   - every function returns 0 on success, a negative BR2_E* code on misuse / CUDA failure (message via
 *     br2_last_error()).  There is no CPU fallback: without a CUDA device br2_batch_create fails.
 */
#ifndef BLUEROV2_B200_H_
#define BLUEROV2_B200_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32) || defined(__CYGWIN__)
#define BR2_API __declspec(dllexport)
#else
#define BR2_API __attribute__((visibility("default")))
#endif

#define BR2_NX 12
#define BR2_NU 4
#define BR2_NP 16
#define BR2_NY 16
#define BR2_NTHRUST 6
#define BR2_NEKF 18

#define BR2_OK 0
#define BR2_EINVAL (-1)
#define BR2_ECUDA (-2)
#define BR2_ENOMEM (-3)

typedef struct br2_batch_solver br2_batch_solver;

BR2_API const char *br2_last_error(void);
BR2_API const char *br2_version(void);
/* number of CUDA devices visible (0 when none / no driver) */
BR2_API int br2_device_count(void);

/* The problem data the generated solver bakes in (acados_solver_bluerov2.c / scripts/acados_ocp.json), as this library
 * bakes them: ONE table used by br2_batch_create and by the acados-ABI shim.  Host-only, needs no CUDA device.
 *   N, Tf            BLUEROV2_N = 80 shooting intervals over tf = 1.0 s -> time step Tf / N = 0.0125 (:137, :389)
 *   W, We            diagonals of the stage / terminal least-squares weights (:424-459, :468-479); cost scaling = time step on
 *                    stages 0..N-1, terminal unscaled (:389-394)
 *   lbu, ubu         input box on stages 0..N-1 (:547-571)
 *   x_init           initial guess of every state node and the default stage-0 bounds lbx_0 = ubx_0 (:522-541, :681-708)
 *   nbxe0            number of stage-0 state bounds flagged as equalities (:228) -- all 12
 *   qp_iter_max      :668;   erk_stages / erk_steps   integrator: ERK, 4 stages, 1 step per interval (:633, :639)
 *   qp_warm_start    0 (json solver_options.qp_solver_warm_start): every QP is solved from a cold start */
typedef struct br2_ocp_defaults {
    int N, nx, nu, np, ny, ny_e, nbu, nbx0, nbxe0, qp_iter_max, qp_warm_start, erk_stages, erk_steps;
    double Tf;
    double W[16], We[12], lbu[4], ubu[4], x_init[12];
} br2_ocp_defaults;
BR2_API void br2_get_ocp_defaults(br2_ocp_defaults *out);

/* == bluerov2_acados_create_with_discretization (acados_solver_bluerov2.c:734-783) for `batch` instances on CUDA
 * device `device`.  time_steps[N] may be NULL: N equal steps over Tf = 1 s (generate_c_code.py:17,24).
 * Defaults baked exactly as the generated C: W, W_e (:424-479), |u| <= 50 (:547-571), qp_iter_max 50 (:668),
 * initial guess x_k = (0,0,-20,0..), u_k = 0 (:681-708). */
BR2_API int br2_batch_create(br2_batch_solver **out, int batch, int N, const double *time_steps, int device);
BR2_API int br2_batch_free(br2_batch_solver *s);

BR2_API int br2_batch_size(const br2_batch_solver *s);
BR2_API int br2_batch_horizon(const br2_batch_solver *s);

/* == ocp_nlp_cost_model_set(..., "W", ...) for all stages: diagonal stage weights W[16] and terminal We[12] */
BR2_API int br2_batch_set_weights(br2_batch_solver *s, const double *W16, const double *We12);
/* == ocp_nlp_constraints_model_set(..., "lbu"/"ubu", ...) for all stages */
BR2_API int br2_batch_set_bounds(br2_batch_solver *s, const double *lbu4, const double *ubu4);
/* == bluerov2_acados_update_time_steps (acados_solver_bluerov2.c:111-132): Ts and cost scaling per stage */
BR2_API int br2_batch_set_time_steps(br2_batch_solver *s, const double *time_steps);
/* options: "qp_iter_max" (int, default 50), "qp_tol" (double, default 1e-12), "fast_path" (int, default 1: try the
 * unconstrained Riccati solution first and accept it when it lies inside the input box -- it is then the exact QP
 * minimiser; 0 = always run the interior-point iteration), "active_set_path" (int, default 1: when the unconstrained solution leaves the
 * box, the primal-dual active-set iteration -- inputs outside the box pinned at the violated bound, the LQR re-solved, a costate
 * sweep checking the multiplier signs of the pinned inputs, up to six solves -- ; accepted solutions satisfy the KKT conditions of
 * the strictly convex QP, i.e. are its minimiser; everything else falls through to the interior-point iteration; 0 = off),
 * "ekf_model" (int, 0 = DOB filter, 1 = AMPC filter), "kernel_timing" (int, default 1: an event between the lineariser and the QP
 * kernels feeds br2_batch_last_kernel_times; 0 removes it from the tick), "tick_graph" (int, default 1: br2_batch_tick_* replay
 * cached CUDA graphs; 0 = always enqueue kernel by kernel) */
BR2_API int br2_batch_set_option_int(br2_batch_solver *s, const char *name, int v);
BR2_API int br2_batch_set_option_double(br2_batch_solver *s, const char *name, double v);

/* iterate (X[B][N+1][12], U[B][N][4]) -- the linearisation point carried from tick to tick (nlp_out).
 * mode 0: the create-time initial guess (:681-708); mode 1: all zeros like bluerov2_acados_reset (:797-830). */
BR2_API int br2_batch_reset(br2_batch_solver *s, int mode);
BR2_API int br2_batch_set_iterate_host(br2_batch_solver *s, const double *X, const double *U);
BR2_API int br2_batch_get_iterate_host(br2_batch_solver *s, double *X, double *U);
/* device pointers of the iterate (owned by the solver) */
BR2_API int br2_batch_iterate_device(br2_batch_solver *s, double **X, double **U);

/* One SQP-RTI step for every instance == per instance the sequence of BLUEROV2_DOB::solve
 * (bluerov2_dob.cpp:320-321 lbx/ubx = x0, :324-355 update_params, :370-372 yref, :375 bluerov2_acados_solve,
 * :388 ocp_nlp_out_get "u", :390-395 thrust allocation).  Device-resident inputs/outputs; enqueues on `stream`.
 * Any of d_u0 / d_thrust / d_status may be NULL (internal buffers are used). */
BR2_API int br2_batch_solve_device(br2_batch_solver *s, const double *d_x0, const double *d_yref, const double *d_p,
                           int p_per_stage, double *d_u0, double *d_thrust, int *d_status, void *stream);
/* Same with HOST buffers: H2D of x0/yref/p, solve, D2H of u0/thrust/status, synchronise.  Pinned host memory
 * makes the copies asynchronous; pageable memory works too. */
BR2_API int br2_batch_solve_host(br2_batch_solver *s, const double *x0, const double *yref, const double *p,
                         int p_per_stage, double *u0, double *thrust, int *status);

/* One control tick as ONE call == the body of the node's loop, `EKF(); solve();` (bluerov2_dobmpc/src/bluerov2_dob_node.cpp:13-31):
 *   [EKF (-> RLS, AMPC) -> parameters] -> SQP-RTI step -> u0, thrusts [-> nominal plant step, device closed loops].
 * The kernels of a tick are instantiated as one CUDA graph per distinct set of buffers (a small cache keyed on this struct):
 * a caller that reuses its buffers from tick to tick -- every closed loop does -- pays one graph launch per tick instead of
 * four to six kernel launches, events and copies.  Pointers are device pointers for _device and host pointers for _host
 * (pinned host buffers are captured into the graph, copies included; pageable ones take the stream path).
 *   x0        [B][12]  measured state; with ekf != 0 also the filter's measurement (meas_y, bluerov2_dob.cpp:501-503)
 *   yref      [B][N+1][16] explicit reference, or NULL with lines != NULL: windowed from br2_batch_set_trajectory
 *   lines     [B] first trajectory row per instance (int)
 *   p         [B][16] ([B][N+1][16] when p_per_stage) OCP parameters; ignored when ekf != 0 (the filter writes them).  NULL (ekf == 0):
 *             the parameters of the last _host call that supplied them -- they persist in the solver like the capsule's after
 *             bluerov2_acados_update_params, which the reference's nodes call only when a parameter changes
 *   ekf       0: no filter.  1: BLUEROV2_DOB::EKF -> p (bluerov2_dob.cpp:324-355).  2: EKF -> RLSFF -> p (BLUEROV2_AMPC::solve,
 *             bluerov2_ampc.cpp:340-382).  thrusts[B][6], body_acc[B][6] are the filter's inputs; compensate = COMPENSATE_D
 *   u0 [B][4], thrust [B][6], status [B] outputs (any may be NULL: internal buffers); wf_dist [B][6] world-frame disturbance (or NULL)
 *   plant_h   > 0 (device only): after the solve, one nominal plant step of length plant_h on x0 IN PLACE with the new u0
 *             (br2_plant_step_device: optional wave_amp[B][4] / wave_tau0[B] wrench, body_acc out, lines++), so that a device-resident
 *             closed loop is a sequence of identical calls. */
typedef struct br2_tick_io {
    const double *x0, *yref, *p, *thrusts;
    const int *lines;
    double *body_acc;            /* EKF input; also the plant's output when plant_h > 0 */
    double *u0, *thrust, *wf_dist;
    int *status;
    const double *wave_amp, *wave_tau0;
    double plant_h;
    int p_per_stage, ekf, compensate;
} br2_tick_io;
BR2_API int br2_batch_tick_device(br2_batch_solver *s, const br2_tick_io *io, void *stream);
BR2_API int br2_batch_tick_host(br2_batch_solver *s, const br2_tick_io *io);
/* Host path with an explicit reference (yref != NULL): a tracking controller knows its reference window one tick ahead -- the
 * measurement it does not.  Registering the NEXT tick's window (pinned host memory, not to be touched until that tick has returned)
 * before calling br2_batch_tick_host lets the library upload it while the current tick computes: behind the current tick's own
 * small uploads on the wire, beside its kernels, into the one of two device buffers the current tick does not read.  The tick whose
 * io.yref is that pointer then uploads only x0 / p; any other pointer takes the ordinary path.  The (N+1) x 16 doubles per instance
 * are 21.5 MB per tick at B = 4096, N = 40 -- longer on PCIe than the tick's kernels. */
BR2_API int br2_batch_set_next_yref_host(br2_batch_solver *s, const double *yref_next);
/* Pinned (page-locked, mapped) host memory for the buffers of the _host entry points; write_combined != 0 for buffers the CPU only
 * writes and the copy engine reads (large per-tick inputs): no cache snooping on the way to the device. */
BR2_API int br2_host_alloc(void **out, size_t bytes, int write_combined);
BR2_API int br2_host_free(void *p);
/* number of CUDA graphs instantiated so far (diagnostic: stays constant in a steady closed loop) */
BR2_API int br2_batch_graphs_built(const br2_batch_solver *s);
/* host path: replays of a cached graph on NEW (pinned) input buffers -- its upload nodes are re-pointed, nothing is re-instantiated */
BR2_API int br2_batch_graph_updates(const br2_batch_solver *s);
/* index the next tick's plant step uses for the wave phase (tau = tau0 + 0.125 * index, bluerov2_dob.cpp:774-797); the solver
 * counts ticks by itself from 0 at creation */
BR2_API int br2_batch_set_tick_index(br2_batch_solver *s, int next_tick);

/* Reference windowing on the device == BLUEROV2_DOB::ref_cb (bluerov2_dob.cpp:218-265) / BLUEROV2_PATH::read_N_pub
 * (bluerov2_path/src/bluerov2_path.cpp:79-118): the trajectory file's rows (16 columns each, readDataFromFile
 * bluerov2_dob.cpp:182-216) are uploaded once; per tick an instance only names its first row `line`
 * (line_number++ in the node) and stage k reads row min(line + k, rows - 1).  Saves the (N+1) x 16 yref upload per tick. */
BR2_API int br2_batch_set_trajectory(br2_batch_solver *s, const double *traj, int rows);
BR2_API int br2_batch_solve_windowed_device(br2_batch_solver *s, const double *d_x0, const int *d_lines, const double *d_p,
                                            int p_per_stage, double *d_u0, double *d_thrust, int *d_status, void *stream);
BR2_API int br2_batch_solve_windowed_host(br2_batch_solver *s, const double *x0, const int *lines, const double *p,
                                          int p_per_stage, double *u0, double *thrust, int *status);

/* statistics of the last solve: iters[B] (IPM iterations), info[B][4] = (mu, stationarity residual,
 * max |dynamics gap| at the linearisation point, stationarity scale) */
BR2_API int br2_batch_get_stats_host(br2_batch_solver *s, int *iters, double *info);
/* linearisation of the last solve (debug / tests): G[B][N][12][16] = [A|B] row-major, b[B][N][12] */
BR2_API int br2_batch_get_linearization_host(br2_batch_solver *s, double *AB, double *b);
/* device time of the last br2_batch_solve_* in seconds (CUDA events) == ocp_nlp_get(..., "time_tot", ...) */
BR2_API double br2_batch_last_solve_time(br2_batch_solver *s);
/* device time of the two kernels of the last solve (linearisation, Riccati IPM), seconds, CUDA events on the
 * launching stream */
BR2_API int br2_batch_last_kernel_times(br2_batch_solver *s, double *t_linearize, double *t_ipm);
/* IPM iterations executed, summed over all instances and all solves since creation / the last reset (the n_it of
 * the roofline formula bytes_sweep = 4384 * N * n_it) */
BR2_API long long br2_batch_ipm_iterations_total(br2_batch_solver *s, int reset);
/* instances that ended a solve with a non-zero status, summed over all solves since creation / the last reset */
BR2_API long long br2_batch_nonzero_status_total(br2_batch_solver *s, int reset);

/* Instrumentation (library built with -DBR2_PROFILE only; all zeros otherwise): SM cycles per phase of the IPM kernel, summed
 * over warps since the last reset -- [0] factor sweep (interior fast path), [1] factor sweep (pinned), [2] closed-loop roll-out,
 * [3] primal check, [4] costate check, [5] IPM start + roll-out, [6] IPM factor sweep, [7] affine forward sweep, [8] affine step
 * length / centring, [9] corrector backward sweep, [10] corrector forward sweep, [11] step lengths / update, [12] epilogue. */
BR2_API int br2_batch_phase_cycles(br2_batch_solver *s, unsigned long long *out16, int reset);
/* the same for the EKF kernel: [0] load, [1] RK4 + F, [2] P_pred, [3] h + H, [4] S, [5] inverse, [6] gain, [7] state + Joseph form,
 * [8] store */
BR2_API int br2_batch_ekf_phase_cycles(br2_batch_solver *s, unsigned long long *out12, int reset);

/* Sharding across GPUs: instances are independent, so the global batch is cut into contiguous blocks, one solver
 * (one process, one GPU) per block, and there is no data-path collective inside the solve.  What a tick exchanges is its output:
 * every rank ends the tick holding the thrust vectors of ALL instances.  The exchange is peer-to-peer and fused into the solve:
 * the QP epilogue stores each instance's six thrusts into the same row of every rank's gather buffer over NVLink (CUDA IPC
 * mappings between the processes), and the warp that finishes a rank's last instance publishes the tick index in every rank's
 * flag array -- no collective kernel, no extra launch.
 *   br2_batch_shard_init     allocate this rank's gather buffer [2][world * batch][6] (two tick parities) + flags; rank in [0, world)
 *   br2_batch_shard_handle   64-byte CUDA IPC handle of it: send it to the other ranks (MPI / torch.distributed / a pipe)
 *   br2_batch_shard_connect  map rank `peer`'s buffer from its handle; after all peers are connected the ticks deliver
 *   br2_batch_shard_wait     enqueue on `stream` a wait until every rank has published the current tick (tick = ticks solved so far)
 *   br2_batch_shard_gathered device pointer of the local gather buffer of tick parity `parity` ([world * batch][6], rank-major)
 * world <= 8 (one NVSwitch domain).  All ranks must run the same number of ticks. */
BR2_API int br2_batch_shard_init(br2_batch_solver *s, int rank, int world);
BR2_API int br2_batch_shard_handle(br2_batch_solver *s, void *handle64);
BR2_API int br2_batch_shard_connect(br2_batch_solver *s, int peer, const void *handle64);
BR2_API int br2_batch_shard_wait(br2_batch_solver *s, void *stream);
BR2_API int br2_batch_shard_gathered(br2_batch_solver *s, int parity, double **d_buf);
/* ticks solved so far (the index the NEXT tick will publish is this value + 1) */
BR2_API int br2_batch_tick_count(br2_batch_solver *s);

/* Nominal plant for device-resident closed-loop studies: one RK4 step of length h of the OCP model
 * (bluerov2_dobmpc/scripts/bluerov2.py:103-137) per instance, x[B][12] in place, inputs u[B][4], parameters p[B][16].
 * Optional (NULL to skip): d_dist[B][4] extra disturbance on p[0..3]; d_wave_amp[B][4] + d_wave_tau0[B] the wave wrench of
 * applyBodyWrench mode 0 (bluerov2_dob.cpp:774-797) at tick `tick`; d_body_acc[B][6] out = finite-differenced body
 * velocities (bluerov2_dob.cpp:148-153); d_lines[B] trajectory row counters, incremented (line_number++, :367).
 * Launches on the device that owns d_x (the caller's current device is left untouched); only enqueues on `stream`. */
BR2_API int br2_plant_step_device(int batch, double *d_x, const double *d_u, const double *d_p, const double *d_dist,
                                  const double *d_wave_amp, const double *d_wave_tau0, int tick, double h,
                                  double *d_body_acc, int *d_lines, void *stream);

/* Same plant step with the disturbance of applyBodyWrench mode 2 (bluerov2_dob.cpp:818-874): the wrench (fx, fy, fz, tz) is row
 * min(tick + phase[b], rows - 1) of d_table[rows][4] -- the reference's config/force{x,y,z}.txt and torquez.txt stacked column-wise,
 * all four read at the same counter.  d_phase[B] (int, may be NULL: 0) lets the instances start at different rows. */
BR2_API int br2_plant_step_replay_device(int batch, double *d_x, const double *d_u, const double *d_p, const double *d_table, int rows,
                                         const int *d_phase, int tick, double h, double *d_body_acc, int *d_lines, void *stream);

/* == BLUEROV2_DOB::EKF (bluerov2_dob.cpp:495-545) for every instance, on the solver's stream order.
 * esti_x[B][18], esti_P[B][18][18] live in the solver (br2_batch_ekf_reset sets x = (0,0,-20,0..,6,6,6,0,0,0),
 * P = I: bluerov2_dob.cpp:64-65).  thrusts[B][6] = measured thruster forces, meas[B][12] = pose + body velocities,
 * body_acc[B][6].  d_p_out[B][16] (may be NULL) receives the OCP parameter vector of bluerov2_dob.cpp:324-355
 * (disturbance estimates scaled by compensate_coef / rotor_constant when `compensate`, nominal hydrodynamics). */
BR2_API int br2_batch_ekf_reset(br2_batch_solver *s);
BR2_API int br2_batch_ekf_device(br2_batch_solver *s, const double *d_thrusts, const double *d_meas, const double *d_body_acc,
                         double *d_wf_dist, double *d_p_out, int compensate, void *stream);
BR2_API int br2_batch_ekf_host(br2_batch_solver *s, const double *thrusts, const double *meas, const double *body_acc,
                       double *wf_dist, double *p_out, int compensate);
BR2_API int br2_batch_ekf_get_state_host(br2_batch_solver *s, double *esti_x, double *esti_P);
BR2_API int br2_batch_ekf_set_state_host(br2_batch_solver *s, const double *esti_x, const double *esti_P);
/* The filter of the adaptive-MPC node (BLUEROV2_AMPC::EKF, bluerov2_ampc.cpp:518-545 with f/h :658-696) is the same
 * code with no damping in the model: select it with br2_batch_set_option_int(s, "ekf_model", 1) (0 = DOB, default). */

/* == BLUEROV2_AMPC::RLSFF (bluerov2_ampc.cpp:731-1004): recursive least squares with variable forgetting factor, four
 * estimators per instance (axes X, Y, Z, N) fed by the EKF state living in the solver (targets esti_x(12|13|14|17)),
 * regressors [body_acc, vel, 1, vel|vel|] from body_acc[B][6] and meas[B][12] (body velocities in 6..11).
 * State per axis, BR2_RLS_STRIDE doubles: theta[4] | P[4][4] | lambda | F-statistic | n_short | n_long |
 * short error window[5] | long error window[50] (oldest first) | pad; br2_batch_rls_reset = constructor values
 * (theta 0, P = I, lambda 0.9: :62-79).  p_out[B][16] (may be NULL) receives the OCP parameters as
 * BLUEROV2_AMPC::solve fills them (:340-382): p[0..3] = theta(2) / (compensate_coef | rotor_constant) plus the nominal
 * p[4..15] when `compensate`; otherwise p[0..3] = 0 and p[4..15] are left as they were (the reference's brace
 * placement, :345-379) -- so p_out is in/out.  AMPC tick = ekf (ekf_model 1) -> rls -> solve on one stream. */
#define BR2_RLS_STRIDE 80
BR2_API int br2_batch_rls_reset(br2_batch_solver *s);
BR2_API int br2_batch_rls_device(br2_batch_solver *s, const double *d_meas, const double *d_body_acc, double *d_p_out,
                         int compensate, void *stream);
BR2_API int br2_batch_rls_host(br2_batch_solver *s, const double *meas, const double *body_acc, double *p_out, int compensate);
BR2_API int br2_batch_rls_get_state_host(br2_batch_solver *s, double *state);
BR2_API int br2_batch_rls_set_state_host(br2_batch_solver *s, const double *state);

/* == the continuous-yaw accumulator at the top of BLUEROV2_DOB::solve (bluerov2_dob.cpp:272-304; same in
 * bluerov2_ampc.cpp:285-317): x0[B][12] arrives with the measured yaw in (-pi, pi] in column 5 (tf getRPY, pose_cb) and
 * leaves with yaw_sum, the accumulated shortest signed differences.  pre_yaw / yaw_sum are FLOATS in the reference
 * (bluerov2_dob.h:234-236) and so are they here: x0[psi] carries that float32 rounding.  State per instance:
 * (pre_yaw, yaw_sum), both 0 after br2_batch_yaw_reset / creation.  In place; the host variant round-trips x0. */
BR2_API int br2_batch_yaw_reset(br2_batch_solver *s);
BR2_API int br2_batch_yaw_unwrap_device(br2_batch_solver *s, double *d_x0, void *stream);
BR2_API int br2_batch_yaw_unwrap_host(br2_batch_solver *s, double *x0);
BR2_API int br2_batch_yaw_get_state_host(br2_batch_solver *s, float *state /* [B][2] */);
BR2_API int br2_batch_yaw_set_state_host(br2_batch_solver *s, const float *state);

/* == BLUEROV2_STATES::ImuDoNodelet::predict / update (bluerov2_states/src/Eskf.cpp:97-141, 197-331): the IMU error-state Kalman
 * filter that is the consolidated node's alternative disturbance source (ctrller_type DOMPC reads its /disturbance).  B independent
 * filters.  Nominal state per instance: p[3], v[3], R[9] (row-major rotation, body -> inertial), xi[3] (disturbance force, body
 * frame); error covariance P[21][21] over [dp, dv, dtheta, db_g, db_a, dg, dxi].  dt = 1/50 s (:103).
 *   br2_eskf_params   diagonals of Q_process / R_meas and the biases: q_p, q_v, q_r, q_q (the nine bias / gravity states), q_xi;
 *                     r_p, r_v, r_r, r_th; b_a[3], b_g[3] (Config.cpp:119-161; NULL = launch/config/imudo.yaml)
 *   predict           imu[B][6] = (specific force, angular rate) of the newest IMU sample
 *   update            gps_p[B][3], gps_v[B][3], R_meas[B][9] (the attitude measurement; ground truth in the reference, :170-176),
 *                     thrusts[B][6], imu_raw[B][6], R_gt[B][9] (attitude giving dynamics_Ma its gravity direction, Dynamics.cpp:180);
 *                     xi_world[B][3] (may be NULL) = R xi as published on /xi, innov[B][12] (may be NULL) = the innovation
 * Quirks kept: the velocity correction is injected twice (:320); biases and gravity are never injected. */
typedef struct br2_eskf br2_eskf;
typedef struct br2_eskf_params { double q_p, q_v, q_r, q_q, q_xi, r_p, r_v, r_r, r_th, b_a[3], b_g[3]; } br2_eskf_params;
BR2_API int br2_eskf_create(br2_eskf **out, int batch, const br2_eskf_params *prm, int device);
BR2_API int br2_eskf_free(br2_eskf *f);
/* state[B][18] = p, v, R, xi and P[B][441]; either may be NULL (left as is).  Creation leaves p = v = xi = 0, R = I, P = 0 */
BR2_API int br2_eskf_set_state_host(br2_eskf *f, const double *state, const double *P);
BR2_API int br2_eskf_get_state_host(br2_eskf *f, double *state, double *P);
BR2_API int br2_eskf_predict_device(br2_eskf *f, const double *d_imu, void *stream);
BR2_API int br2_eskf_update_device(br2_eskf *f, const double *d_gps_p, const double *d_gps_v, const double *d_R_meas, const double *d_thrusts,
                                   const double *d_imu_raw, const double *d_R_gt, double *d_xi_world, double *d_innov, void *stream);
BR2_API int br2_eskf_predict_host(br2_eskf *f, const double *imu);
BR2_API int br2_eskf_update_host(br2_eskf *f, const double *gps_p, const double *gps_v, const double *R_meas, const double *thrusts,
                                 const double *imu_raw, const double *R_gt, double *xi_world, double *innov);

#ifdef __cplusplus
}
#endif
#endif
