// engine.h -- internal interface between the host engine (engine.cu) and the kernels (kernels.cu).
#pragma once
#include <cuda_runtime.h>
#include "layout.h"
#include "model.cuh"

namespace br2 {

constexpr int NMAX = 256;   // maximum horizon supported by the engine

// Sharded batch (one solver per GPU / process): where the epilogue delivers the thrust vectors of this rank's instances
constexpr int MAX_SHARDS = 8;
struct ShardView {
    int world, rank;               // world <= 1: not sharded
    double* buf[MAX_SHARDS];       // gather buffer of every rank: [2 (tick parity)][world * B][6]; [rank] is the local one
    int* flag[MAX_SHARDS];         // flag array of every rank: [world] last tick index published by each rank
};

// Everything a launch needs; passed by value (fits the 4 KB kernel parameter space).
struct SolveArgs {
    int B, N;
    int lo, hi;            // instance range of this launch (lineariser / pdas_kernel): the whole batch, or one chunk of a pipelined tick
    int housekeeping;      // what block 0 of the lineariser resets: 2 = everything a tick needs (the normal case: this launch opens
                           // the tick), 0 = nothing (ranges of a pipelined tick: tick_begin_kernel has done it for all of them)
    int qidx;              // which work-queue counter (CTR_QUEUE + qidx) this range's pdas_kernel uses: ranges run concurrently
    // problem data (device)
    const double* x0;      // [B][12]
    const double* yref;    // [B][N+1][16], or null when the reference is windowed on the device:
    const double* traj;    // [traj_rows][16] reference trajectory (device), row = min(lines[b] + k, traj_rows - 1)
    const int* lines;      // [B] first trajectory row of each instance's horizon (ref_cb, bluerov2_dob.cpp:218-265)
    int traj_rows;
    const double* p;       // [B][p_inst_stride]: stage k reads p + k*p_stage_stride
    int p_inst_stride;     // 16 (one vector per instance) or (N+1)*16
    int p_stage_stride;    // 0 or 16
    const double* Ts;      // [N] (device)
    double W[16], We[12], lbu[4], ubu[4];
    // iterate (device, in/out)
    double* X;             // [B][N+1][12]
    double* U;             // [B][N][4]
    // workspaces (device)
    double* S;             // [B][N+1][SREC] stage records [V | G | F] (layout.h)
    // outputs (device)
    double* u0;            // [B][4]
    double* thrust;        // [B][6]  (may alias an NCCL / symmetric send buffer)
    int* status;           // [B]
    int* iters;            // [B]
    double* info;          // [B][4]: mu, res_stat, max|b| (dynamics gap at the linearisation point), max step
    int* ctr;              // [CTR_COUNT] CTR_QUEUE: persistent-kernel work queue; CTR_HARD / CTR_EASY: fill counts of the next solve's
                           //     visiting order; CTR_PARITY: which half of `order` is current (reset / flipped by the lineariser)
    int* fb;               // [B] fallback list: instances the fast paths handed to the interior-point kernel (count in ctr[CTR_FB])
    int* order;            // [2][B] visiting order of the instances (a permutation), double-buffered: hint = 1 instances first
    int* hint;             // [B] 1 = a bound was active at the previous solution (skip the interior fast path)
    int fast_path;         // try the interior-solution fast path (option "fast_path", default 1)
    int* aset;             // [B][N] guessed active set per stage, 2 bits per input (0 free, 1 at lbu, 2 at ubu)
    int active_set;        // primal-dual active-set iteration when the interior solution leaves the box (option "active_set_path")
    unsigned long long* iter_total;   // [1] IPM iterations executed, accumulated over instances and solves
    unsigned long long* bad_total;    // [1] instances that ended with a non-zero status, accumulated over solves
    unsigned long long* prof;         // [16] cycles per kernel phase (instrumentation build -DBR2_PROFILE only), else unused
    ShardView shard;
    // options
    int max_iter;          // qp_solver_iter_max (50)
    double tol;            // IPM tolerance on mu and on the scaled stationarity residual
};

// reference row of (instance, stage): explicit yref or the windowed trajectory (clamped to the last row)
__device__ __forceinline__ const double* yref_row(const SolveArgs& a, int inst, int k)
{
    if (a.yref) return a.yref + ((size_t)inst * (a.N + 1) + k) * NY;
    int r = a.lines[inst] + k;
    r = r < a.traj_rows - 1 ? r : a.traj_rows - 1;
    return a.traj + (size_t)(r < 0 ? 0 : r) * NY;
}

void launch_linearize(const SolveArgs& a, cudaStream_t s);
void launch_ipm(const SolveArgs& a, int sm_count, cudaStream_t s);          // = launch_pdas + launch_ipm_fallback
void launch_pdas(const SolveArgs& a, int sm_count, cudaStream_t s);
void launch_ipm_fallback(const SolveArgs& a, int sm_count, cudaStream_t s);
void launch_tick_begin(int* ctr, cudaStream_t s);
void launch_exchange(const ShardView& sh, int B, int* ctr, cudaStream_t s);
void launch_shard_wait(const int* flag, int world, int tick, int* timed_out, cudaStream_t s);     // per-tick housekeeping as a kernel of its own (pipelined ticks)
void configure_kernels();      // per-device function attributes (call with the solver's device current)
enum { CTR_HARD = 1, CTR_EASY = 2, CTR_PARITY = 3, CTR_FB = 4, CTR_FBQ = 5, CTR_TICK = 6, CTR_QUEUE = 8, CTR_XTICK = 12, CTR_XDONE = 13, CTR_COUNT = 16 };   // CTR_QUEUE .. +3: one per range

// EKF (bluerov2_dob.cpp:495-545), one warp per instance
struct EkfArgs {
    int B;
    double* esti_x;          // [B][18] in/out
    double* esti_P;          // [B][18][18] in/out
    const double* thrusts;   // [B][6]  measured thruster forces (meas_u)
    const double* meas;      // [B][12] pose + body velocities
    const double* body_acc;  // [B][6]
    double* wf_dist;         // [B][6] out (may be null)
    double* p_out;           // [B][16] out (may be null): OCP parameter vector per bluerov2_dob.cpp:324-355
    int compensate;          // COMPENSATE_D
    int model;               // 0 = BLUEROV2_DOB filter, 1 = BLUEROV2_AMPC filter (no damping in f / h)
};
void launch_ekf(const EkfArgs& a, cudaStream_t s);
void configure_ekf();          // per-device function attributes
void ekf_phase_cycles(unsigned long long* out12, int reset);   // -DBR2_PROFILE builds; zeros otherwise

// RLS with variable forgetting factor (rls.cu; BLUEROV2_AMPC::RLSFF, bluerov2_ampc.cpp:731-1004), one thread per
// (instance, axis); state layout RLS_* below == oracle ORC_RLS_STRIDE layout
constexpr int RLS_STRIDE = 80, RLS_THETA = 0, RLS_P = 4, RLS_LAMBDA = 20, RLS_F = 21, RLS_NN = 22, RLS_ND = 23, RLS_EN = 24,
              RLS_ED = 29, RLS_FFN = 5, RLS_FFD = 50;
struct RlsArgs {
    int B;
    double* state;           // [B][4][RLS_STRIDE] in/out
    const double* esti_x;    // [B][18] EKF estimate (targets: components 12, 13, 14, 17)
    const double* body_acc;  // [B][6]
    const double* meas;      // [B][12] (body velocities in 6..11)
    double* p_out;           // [B][16] out (may be null): parameters as BLUEROV2_AMPC::solve fills them (:340-382)
    int compensate;
};
void launch_rls(const RlsArgs& a, cudaStream_t s);

// IMU error-state Kalman filter (eskf.cu; bluerov2_states/src/Eskf.cpp:97-331), one warp per instance
struct EskfArgs {
    int B;
    double* state;           // [B][18] nominal state: p[3] v[3] R[9] (row-major) xi[3]
    double* P;               // [B][21][21] error covariance
    double Qd[21], Rd[12];   // diagonals of Q_process / R_meas (Config.cpp:128-155)
    double b_a[3], b_g[3];   // accelerometer / gyro bias (launch/config/imudo.yaml)
    const double* imu;       // [B][6] specific force + angular rate (predict: the IMU sample; update: imu_raw_B)
    const double* gps_p;     // [B][3] update: position measurement
    const double* gps_v;     // [B][3] update: velocity measurement
    const double* R_meas;    // [B][9] update: attitude measurement
    const double* R_gt;      // [B][9] update: attitude that gives dynamics_Ma its gravity direction
    const double* thrusts;   // [B][6] update: thruster forces
    double* xi_world;        // [B][3] out (may be null): R xi, what the nodelet publishes on /xi
    double* innov;           // [B][12] out (may be null): the innovation y
};
void launch_eskf_predict(const EskfArgs& a, cudaStream_t s);
void launch_eskf_update(const EskfArgs& a, cudaStream_t s);
void configure_eskf();

// continuous-yaw accumulator of the node glue (glue.cu; bluerov2_dob.cpp:272-304): state[B][2] = (pre_yaw, yaw_sum) floats,
// x0[B][12] in place on column 5
void launch_yaw_unwrap(int B, float* state, double* x0, cudaStream_t s);

// nominal plant (plant.cu): x <- RK4_h(x, u, p + disturbance); optional wave disturbance, body acceleration, line counter
struct PlantArgs {
    int B;
    double* x;               // [B][12] in/out
    const double* u;         // [B][4]
    const double* p;         // [B][16]
    const double* dist;      // [B][4] extra disturbance on p[0..3], or null
    const double* wave_amp;  // [B][4] wave amplitudes (A_x, A_y, A_z, A_y/3), or null
    const double* wave_tau0; // [B]    wave phases
    const double* table;     // [table_rows][4] wrench series (fx, fy, fz, tz) replayed row by row (mode 2 of applyBodyWrench), or null
    int table_rows;
    const int* table_phase;  // [B] first row of each instance (or null: 0)
    double* body_acc;        // [B][6] out, or null
    int* lines;              // [B] in/out trajectory row counters, or null
    double h;
    int tick;                // wave phase index; < 0: read it from *tick_ctr - 1 (the solver's tick counter: graph replays)
    const int* tick_ctr;
};
void launch_plant(const PlantArgs& a, cudaStream_t s);

}  // namespace br2
