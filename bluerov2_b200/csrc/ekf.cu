// ekf.cu -- batched 18-state disturbance-observer EKF, one warp per instance.
//
// Restates BLUEROV2_DOB::EKF and its helpers (bluerov2_dobmpc/src/bluerov2_dob.cpp:495-545, 621-752) with the
// constants of include/bluerov2_dobmpc/bluerov2_dob.h:170-208 and the constructor (bluerov2_dob.cpp:41-65):
//   state  [eta(6), nu(6), d(6)];  process model f (:637-695) with rigid-body Coriolis terms and invM(i,i) of the
//   COUPLED mass matrix;  RK4 with k3 = f(x + k2/3) (sic, :630);  F and H by forward differences, d = 1e-6
//   (:722-752);  gain through an explicit 18x18 inverse (:535);  Joseph-form covariance update (:537).
//
// Mapping.  (1) The 19 RK4 / 19 h() evaluations of the two finite-difference Jacobians run on lanes 0..18 in parallel
// (lane 0 = unperturbed, lane c + 1 perturbs component c); only the 12 components with dynamics are integrated (the
// disturbance states are constants of f), divisions by constants are multiplications.  (2) The ten 18 x 18 x 18 products
// of the covariance algebra run on the fp64 tensor-core instruction (DMMA m8n8k4) as four fused chains per tile row,
//      P_pred = (F P) F' + Q,   S = (H P_pred) H' + R,   K = (P_pred H') S^-1,  A = I - K H,   P = (A P_pred) A' + K R K',
// every operand a fragment read straight from the row-major matrix in shared memory: with the contraction index
// enumerated in accumulator-fragment order (even columns, then odd columns of an 8-tile) the accumulator of one product
// IS the A operand of the next, and X M' needs the same fragment of M as of X -- no re-layout anywhere.  18 = 8 + 8 + 2:
// fragments are zero-padded by predication; tile products whose operand tile is structurally zero (F = [* * *; * * *;
// 0 0 D], H = [D 0 0; * * 0; * * *] in 8-tiles) are skipped at compile time: 444 DMMA per instance instead of 540 (and
// instead of 2160 DFMA + 900 shared-memory loads of the round-1 kernel).  (3) The Gauss-Jordan inverse keeps one row of
// S per lane in registers, finds the pivot with three warp reductions and exchanges row POSITIONS instead of rows.
// 16 warps/SM (13.6 KB of shared memory per warp).  Output: esti_x, esti_P, world-frame disturbance (:540-545) and
// the OCP parameter vector handed to the solver (bluerov2_dob.cpp:324-355).
#include "engine.h"

namespace br2 {

#define FULL_MASK 0xffffffffu
constexpr int EN = 18;
constexpr int LD = 18;            // leading dimension of the shared-memory matrices (144-byte rows: 16-byte aligned)
#ifndef BR2_EKF_WARPS
#define BR2_EKF_WARPS 2
#endif
#ifndef BR2_EKF_MINB
#define BR2_EKF_MINB 8
#endif
constexpr int EKF_WARPS = BR2_EKF_WARPS;

namespace ekfc {
constexpr double DT = 0.05, M = 11.26, Ix = 0.3, Iy = 0.63, Iz = 0.58, Zg = 0.02, G = 9.81, EBUOY = 0.661618;
constexpr double COMP = 0.032546960744430276, RCK = 0.026546960744430276;
constexpr double AM0 = 1.7182, AM1 = 0, AM2 = 5.468, AM3 = 0, AM4 = 1.2481, AM5 = 0.4006;
// diagonal of M and of M^-1 for M = diag(m+am) with the four ZG couplings (bluerov2_dob.cpp:41-47):
// the (0,4) and (1,3) 2x2 blocks invert in closed form.
constexpr double M0 = M + AM0, M1 = M + AM1, M2 = M + AM2, M3 = Ix + AM3, M4 = Iy + AM4, M5 = Iz + AM5;
constexpr double C = M * Zg;
constexpr double IM0 = M4 / (M0 * M4 - C * C), IM4 = M0 / (M0 * M4 - C * C);
constexpr double IM1 = M3 / (M1 * M3 - C * C), IM3 = M1 / (M1 * M3 - C * C);
constexpr double IM2 = 1.0 / M2, IM5 = 1.0 / M5;
constexpr double QP = (DT * DT * DT * DT) / 4, QV = DT * DT;      // Q = diag(QP x 6, QV x 12), R = QP I (bluerov2_dob.cpp:49-65)
constexpr double FD = 1e-6, RFD = 1e6;                            // forward-difference step and its reciprocal
}  // namespace ekfc

__constant__ double c_K[36] = {
    0.7071067811847433, 0.7071067811847433, -0.7071067811919605, -0.7071067811919605, 0.0, 0.0,
    0.7071067811883519, -0.7071067811883519, 0.7071067811811348, -0.7071067811811348, 0.0, 0.0,
    0, 0, 0, 0, 1, 1,
    0.051265241636155506, -0.05126524163615552, 0.05126524163563227, -0.05126524163563227, -0.11050000000000001, 0.11050000000000003,
    -0.05126524163589389, -0.051265241635893896, 0.05126524163641713, 0.05126524163641713, -0.002499999999974481, -0.002499999999974481,
    0.16652364696949604, -0.16652364696949604, -0.17500892834341342, 0.17500892834341342, 0.0, 0.0};
// damping of the filter model, selected per launch (EkfArgs::model): 0 = BLUEROV2_DOB (bluerov2_dob.h:182-183),
// 1 = BLUEROV2_AMPC (bluerov2_ampc.cpp:41-42: Dl = 0; its f/h :658-696 have no quadratic damping)
__constant__ double c_DlM[2][6] = {{-11.7391, -20, -31.8678, -25, -44.9085, -5}, {0, 0, 0, 0, 0, 0}};
__constant__ double c_DnlM[2][6] = {{-18.18, -21.66, -36.99, -1.55, -1.55, -1.55}, {0, 0, 0, 0, 0, 0}};

// ---- models -------------------------------------------------------------------------------------------------------
// process model, bluerov2_dob.cpp:637-695, on the 9 components it reads (ang = x[3..5], nu = x[6..11]) plus the
// disturbance states dd = x[12..17] (constants of f: their derivative is zero, :689-694) and tau = K * thrusts
__device__ __forceinline__ void ekf_f12(const double (&ang)[3], const double (&nu)[6], const double (&dd)[6], const double* tau,
                                        double (&xd)[12], const double* c_Dl, const double* c_Dnl)
{
    using namespace ekfc;
    double s3, c3, s4, c4, s5, c5;
    br2_sincos(ang[0], &s3, &c3); br2_sincos(ang[1], &s4, &c4); br2_sincos(ang[2], &s5, &c5);      // (branch-free: fast_trig.h)
    const double r4 = br2_rcp(c4);
    xd[0] = (c5 * c4) * nu[0] + (-s5 * c3 + c5 * s4 * s3) * nu[1] + (s5 * s3 + c5 * c3 * s4) * nu[2];
    xd[1] = (s5 * c4) * nu[0] + (c5 * c3 + s3 * s4 * s5) * nu[1] + (-c5 * s3 + s4 * s5 * c3) * nu[2];
    xd[2] = (-s4) * nu[0] + (c4 * s3) * nu[1] + (c4 * c3) * nu[2];
    xd[3] = nu[3] + (s5 * s4 * r4) * nu[4] + c3 * s4 * r4 * nu[5];
    xd[4] = (c3) * nu[4] + (s3) * nu[5];
    xd[5] = (s3 * r4) * nu[4] + (c3 * r4) * nu[5];
    xd[6] = IM0 * (tau[0] + M * nu[5] * nu[1] - M * nu[4] * nu[2] - EBUOY * s4 + dd[0] + c_Dl[0] * nu[0] + c_Dnl[0] * fabs(nu[0]) * nu[0]);
    xd[7] = IM1 * (tau[1] - M * nu[5] * nu[0] + M * nu[3] * nu[2] + EBUOY * c4 * s3 + dd[1] + c_Dl[1] * nu[1] + c_Dnl[1] * fabs(nu[1]) * nu[1]);
    xd[8] = IM2 * (tau[2] + M * nu[4] * nu[0] - M * nu[3] * nu[1] + EBUOY * c4 * c3 + dd[2] + c_Dl[2] * nu[2] + c_Dnl[2] * fabs(nu[2]) * nu[2]);
    xd[9] = IM3 * (tau[3] + (Iy - Iz) * nu[4] * nu[5] - M * Zg * G * c4 * s3 + dd[3] + c_Dl[3] * nu[3] + c_Dnl[3] * fabs(nu[3]) * nu[3]);
    xd[10] = IM4 * (tau[4] + (Iz - Ix) * nu[3] * nu[5] - M * Zg * G * s4 + dd[4] + c_Dl[4] * nu[4] + c_Dnl[4] * fabs(nu[4]) * nu[4]);
    xd[11] = IM5 * (tau[5] - (Iy - Ix) * nu[3] * nu[4] + dd[5] + c_Dl[5] * nu[5] + c_Dnl[5] * fabs(nu[5]) * nu[5]);
}

// rows 12..17 of the measurement model, bluerov2_dob.cpp:698-719 (rows 0..11 are y = x)
__device__ __forceinline__ void ekf_h6(double a3, double a4, const double (&nu)[6], const double (&dd)[6], const double* acc,
                                       double (&y)[6], const double* c_Dl, const double* c_Dnl)
{
    using namespace ekfc;
    double s3, c3, s4, c4;
    br2_sincos(a3, &s3, &c3); br2_sincos(a4, &s4, &c4);
    y[0] = M0 * acc[0] - M * nu[5] * nu[1] + M * nu[4] * nu[2] + EBUOY * s4 - dd[0] - c_Dl[0] * nu[0] - c_Dnl[0] * fabs(nu[0]) * nu[0];
    y[1] = M1 * acc[1] + M * nu[5] * nu[0] - M * nu[3] * nu[2] - EBUOY * c4 * s3 - dd[1] - c_Dl[1] * nu[1] - c_Dnl[1] * fabs(nu[1]) * nu[1];
    y[2] = M2 * acc[2] - M * nu[4] * nu[0] + M * nu[3] * nu[1] - EBUOY * c4 * c3 - dd[2] - c_Dl[2] * nu[2] - c_Dnl[2] * fabs(nu[2]) * nu[2];
    y[3] = M3 * acc[3] - (Iy - Iz) * nu[4] * nu[5] + M * Zg * G * c4 * s3 - dd[3] - c_Dl[3] * nu[3] - c_Dnl[3] * fabs(nu[3]) * nu[3];
    y[4] = M4 * acc[4] - (Iz - Ix) * nu[3] * nu[5] + M * Zg * G * s4 - dd[4] - c_Dl[4] * nu[4] - c_Dnl[4] * fabs(nu[4]) * nu[4];
    y[5] = M5 * acc[5] + (Iy - Ix) * nu[3] * nu[4] - dd[5] - c_Dl[5] * nu[5] - c_Dnl[5] * fabs(nu[5]) * nu[5];
}

// ---- fp64 tensor-core products on row-major 18 x 18 matrices in shared memory -------------------------------------------
// Lane (q = lane >> 2, t = lane & 3).  Accumulator fragment of tile (I, J): {C[8I+q][8J+2t], C[8I+q][8J+2t+1]}.  With the
// contraction index of a k-tile enumerated as (0, 2, 4, 6) then (1, 3, 5, 7), the A operand of the two DMMA of k-tile K is
// the .x / .y of "row fragment" {X[8I+q][8K+2t], X[8I+q][8K+2t+1]} -- the accumulator layout --, and the B operand of
// C = X M' is the row fragment of M at (J, K).  C = X Y (no transpose) takes the "column fragment" of Y instead.
__device__ __forceinline__ void dmma(double (&c)[2], double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}
// {X[8R+q][8K+2t], X[8R+q][8K+2t+1]}, zero outside 18 x 18 (R, K compile-time after unrolling)
__device__ __forceinline__ double2 rowfrag(const double* X, int R, int K, int q, int t)
{
    double2 v = make_double2(0.0, 0.0);
    if ((R < 2 || q < 2) && (K < 2 || t == 0)) v = *reinterpret_cast<const double2*>(X + (8 * R + q) * LD + 8 * K + 2 * t);
    return v;
}
// {Y[8K+2t][8J+q], Y[8K+2t+1][8J+q]}: the row fragment of Y'
__device__ __forceinline__ double2 colfrag(const double* Y, int J, int K, int q, int t)
{
    double2 v = make_double2(0.0, 0.0);
    if ((J < 2 || q < 2) && (K < 2 || t == 0)) {
        v.x = Y[(8 * K + 2 * t) * LD + 8 * J + q];
        v.y = Y[(8 * K + 2 * t + 1) * LD + 8 * J + q];
    }
    return v;
}
// structural non-zero tiles, bit (R * 3 + K): F = [* * *; * * *; 0 0 D] (the disturbance rows of the transition matrix are
// the identity), H = [D 0 0; * * 0; * * *] (rows 0..11 of the measurement are the state)
constexpr unsigned TM_ALL = 0x1ff;
constexpr unsigned TM_F = 0x07 | (0x07 << 3) | (0x04 << 6);
constexpr unsigned TM_H = 0x01 | (0x03 << 3) | (0x07 << 6);
constexpr unsigned TM_HT = 0x07 | (0x06 << 3) | (0x04 << 6);      // column fragments of H: tile (J, K) <-> H tile (K, J)
// c[J] += sum_K a[K] * frag(M; J, K) for J = 0..2; amask = the non-zero k-tiles of a (3 bits), mmask = tiles of the fragment
template <bool TRANS, unsigned MMASK>
__device__ __forceinline__ void tile_row_product(double (&c)[3][2], const double2 (&a)[3], unsigned amask_const, const double* Mx, int q, int t)
{
#pragma unroll
    for (int J = 0; J < 3; J++)
#pragma unroll
        for (int K = 0; K < 3; K++)
            if (((MMASK >> (J * 3 + K)) & 1u) && ((amask_const >> K) & 1u)) {
                const double2 m = TRANS ? colfrag(Mx, J, K, q, t) : rowfrag(Mx, J, K, q, t);
                dmma(c[J], a[K].x, m.x);
                dmma(c[J], a[K].y, m.y);
            }
}
__device__ __forceinline__ void load_tile_row(double2 (&a)[3], const double* X, int I, int q, int t)
{
#pragma unroll
    for (int K = 0; K < 3; K++) a[K] = rowfrag(X, I, K, q, t);
}
__device__ __forceinline__ void acc_to_frag(double2 (&a)[3], const double (&c)[3][2])
{
#pragma unroll
    for (int K = 0; K < 3; K++) a[K] = make_double2(c[K][0], c[K][1]);
}
__device__ __forceinline__ void zero_acc(double (&c)[3][2])
{
#pragma unroll
    for (int J = 0; J < 3; J++) c[J][0] = c[J][1] = 0.0;
}
// tile row I of an accumulator -> row-major matrix (shared or global), entries outside 18 x 18 dropped
__device__ __forceinline__ void store_tile_row(double* X, const double (&c)[3][2], int I, int q, int t)
{
#pragma unroll
    for (int J = 0; J < 3; J++)
        if ((I < 2 || q < 2) && (J < 2 || t == 0))
            *reinterpret_cast<double2*>(X + (8 * I + q) * LD + 8 * J + 2 * t) = make_double2(c[J][0], c[J][1]);
}

// 1 / x to about an ulp without the division's special-case branch: hardware seed (~20 bits) and two Newton steps
__device__ __forceinline__ double rcp_full(double x)
{
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    y = fma(y, fma(-x, y, 1.0), y);
    return fma(y, fma(-x, y, 1.0), y);
}

__device__ __forceinline__ void cp16(void* smem, const void* gmem)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}

#ifdef BR2_PROFILE
// SM cycles per phase summed over warps (profile build only; scripts/ekf_phase_profile.py)
__device__ unsigned long long g_ekf_prof[12];
#define EPROF(i) do { const long long t_ = clock64(); if (lane == 0) atomicAdd(&g_ekf_prof[i], (unsigned long long)(t_ - eprof_t0)); eprof_t0 = clock64(); } while (0)
#else
#define EPROF(i) do { } while (0)
#endif

struct __align__(16) EkfSmem {
    double Fm[EN * LD];          // F; the gain once P_pred is formed
    double Pp[EN * LD];          // P_pred
    double Hm[EN * LD];          // H
    double Sm[EN * LD];          // S, then its inverse
    double Am[EN * LD];          // esti_P on entry; I - K H later
    double x[EN];                // esti_x on entry
    double xp[EN];               // x_pred
    double inn[EN];              // innovation
    double tau[6], acc[6];       // K * thrusts, body acceleration
    double prow[2][EN + 2];      // pivot row of the Gauss-Jordan inverse and the reciprocal of the pivot (double-buffered)
};

__global__ void __launch_bounds__(EKF_WARPS * 32, BR2_EKF_MINB) ekf_kernel(EkfArgs a)
{
    using namespace ekfc;
    extern __shared__ __align__(16) unsigned char smraw[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int q = lane >> 2, t = lane & 3;
    EkfSmem& sm = reinterpret_cast<EkfSmem*>(smraw)[wib];
    const int inst = blockIdx.x * EKF_WARPS + wib;
    if (inst >= a.B) return;
#ifdef BR2_PROFILE
    long long eprof_t0 = clock64();
#endif
    const double* c_Dl = c_DlM[a.model & 1];
    const double* c_Dnl = c_DnlM[a.model & 1];
    double* ex = a.esti_x + (size_t)inst * EN;
    double* eP = a.esti_P + (size_t)inst * EN * EN;

    // esti_P -> shared memory, asynchronously (162 chunks of 16 bytes): first read after the two Jacobians
    for (int ch = lane; ch < EN * EN / 2; ch += 32) cp16(sm.Am + 2 * ch, eP + 2 * ch);
    asm volatile("cp.async.commit_group;" ::: "memory");
    if (lane < EN) sm.x[lane] = ex[lane];
    if (lane < 6) {
        // tau = K * thrusts (meas_y rows 12..17, :499-504)
        double s = 0.0;
#pragma unroll
        for (int j = 0; j < 6; j++) s += c_K[lane * 6 + j] * a.thrusts[(size_t)inst * 6 + j];
        sm.tau[lane] = s;
        sm.acc[lane] = a.body_acc[(size_t)inst * 6 + lane];
    }
    __syncwarp();
    EPROF(0);

    // ---- F = d RK4 / dx by forward differences (:722-736): lane 0 unperturbed, lane c + 1 perturbs component c ----
    const int pc = lane - 1;                 // (lanes >= 19 repeat the unperturbed evaluation; nothing of theirs is stored)
    double dd[6];                            // disturbance states of this lane's evaluation
#pragma unroll
    for (int i = 0; i < 6; i++) dd[i] = sm.x[12 + i] + (pc == 12 + i ? FD : 0.0);
    double xn[12];                           // RK4(x (+ d e_c)), components 0..11
    {
        // x + (k1 + 2 k2 + 2 k3 + k4) / 6 with the sum accumulated left to right as the reference's expression evaluates it
        // (:621-635); k3 is evaluated at x + k2 / 3 (sic, :630)
        double x0[12], k[12], sum[12], ang[3], nu[6];
#pragma unroll
        for (int i = 0; i < 12; i++) x0[i] = sm.x[i] + (pc == i ? FD : 0.0);
#pragma unroll
        for (int i = 0; i < 3; i++) ang[i] = x0[3 + i];
#pragma unroll
        for (int i = 0; i < 6; i++) nu[i] = x0[6 + i];
        ekf_f12(ang, nu, dd, sm.tau, k, c_Dl, c_Dnl);
#pragma unroll
        for (int i = 0; i < 12; i++) { k[i] *= DT; sum[i] = k[i]; }
#pragma unroll
        for (int i = 0; i < 3; i++) ang[i] = x0[3 + i] + k[3 + i] * 0.5;
#pragma unroll
        for (int i = 0; i < 6; i++) nu[i] = x0[6 + i] + k[6 + i] * 0.5;
        ekf_f12(ang, nu, dd, sm.tau, k, c_Dl, c_Dnl);
#pragma unroll
        for (int i = 0; i < 12; i++) { k[i] *= DT; sum[i] = sum[i] + 2 * k[i]; }
#pragma unroll
        for (int i = 0; i < 3; i++) ang[i] = x0[3 + i] + k[3 + i] * (1.0 / 3.0);
#pragma unroll
        for (int i = 0; i < 6; i++) nu[i] = x0[6 + i] + k[6 + i] * (1.0 / 3.0);
        ekf_f12(ang, nu, dd, sm.tau, k, c_Dl, c_Dnl);
#pragma unroll
        for (int i = 0; i < 12; i++) { k[i] *= DT; sum[i] = sum[i] + 2 * k[i]; }
#pragma unroll
        for (int i = 0; i < 3; i++) ang[i] = x0[3 + i] + k[3 + i];
#pragma unroll
        for (int i = 0; i < 6; i++) nu[i] = x0[6 + i] + k[6 + i];
        ekf_f12(ang, nu, dd, sm.tau, k, c_Dl, c_Dnl);
#pragma unroll
        for (int i = 0; i < 12; i++) xn[i] = x0[i] + (sum[i] + k[i] * DT) * (1.0 / 6.0);
    }
    double xp[12];                           // x_pred = RK4(esti_x), components 0..11, on every lane
    const bool fdl = lane >= 1 && lane <= EN;
#pragma unroll
    for (int i = 0; i < 12; i++) {
        xp[i] = __shfl_sync(FULL_MASK, xn[i], 0);
        if (fdl) sm.Fm[i * LD + pc] = (xn[i] - xp[i]) * RFD;
    }
#pragma unroll
    for (int i = 0; i < 6; i++)
        if (fdl) sm.Fm[(12 + i) * LD + pc] = (dd[i] - sm.x[12 + i]) * RFD;      // disturbance rows: RK4 leaves them where they were
    if (lane >= 12 && lane < EN) sm.xp[lane] = sm.x[lane];
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < 12; i++) sm.xp[i] = xn[i];
    }
    __syncwarp();
    EPROF(1);
    // ---- H = dh/dx at x_pred by forward differences (:738-752), innovation ----
    {
        double nu[6], y6[6];
#pragma unroll
        for (int i = 0; i < 6; i++) nu[i] = xp[6 + i] + (pc == 6 + i ? FD : 0.0);
        const double a3 = xp[3] + (pc == 3 ? FD : 0.0), a4 = xp[4] + (pc == 4 ? FD : 0.0);
        ekf_h6(a3, a4, nu, dd, sm.acc, y6, c_Dl, c_Dnl);
#pragma unroll
        for (int i = 0; i < 12; i++)
            if (fdl) sm.Hm[i * LD + pc] = pc == i ? ((xp[i] + FD) - xp[i]) * RFD : 0.0;
#pragma unroll
        for (int i = 0; i < 6; i++) {
            const double yp = __shfl_sync(FULL_MASK, y6[i], 0);
            if (fdl) sm.Hm[(12 + i) * LD + pc] = (y6[i] - yp) * RFD;
            if (lane == 0) sm.inn[12 + i] = sm.tau[i] - y6[i];                  // innovation y - y_pred, rows 12..17
        }
        if (lane < 12) sm.inn[lane] = a.meas[(size_t)inst * 12 + lane] - sm.xp[lane];
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();
    EPROF(3);
    // ---- P_pred = (F P) F' + Q (:529), tile row by tile row ----
#pragma unroll
    for (int I = 0; I < 3; I++) {
        double2 af[3];
        double c1[3][2], c2[3][2];
        load_tile_row(af, sm.Fm, I, q, t);
        zero_acc(c1);
        tile_row_product<true, TM_ALL>(c1, af, (TM_F >> (3 * I)) & 7u, sm.Am, q, t);          // (F P)[I][*]: B fragment of X Y = column fragment of Y
        acc_to_frag(af, c1);
        zero_acc(c2);
        tile_row_product<false, TM_F>(c2, af, 7u, sm.Fm, q, t);
        const double qd = (I == 0 && q < 6) ? QP : QV;                                          // Q(r, r), r = 8 I + q
        if (q == 2 * t) c2[I][0] += qd;
        if (q == 2 * t + 1) c2[I][1] += qd;
        store_tile_row(sm.Pp, c2, I, q, t);
    }
    __syncwarp();
    EPROF(2);
    // ---- S = (H P_pred) H' (+ R, added when the rows are loaded for the inverse) (:535) ----
#pragma unroll
    for (int I = 0; I < 3; I++) {
        double2 af[3];
        double c1[3][2], c2[3][2];
        load_tile_row(af, sm.Hm, I, q, t);
        zero_acc(c1);
        tile_row_product<true, TM_ALL>(c1, af, (TM_H >> (3 * I)) & 7u, sm.Pp, q, t);
        acc_to_frag(af, c1);
        zero_acc(c2);
        tile_row_product<false, TM_H>(c2, af, 7u, sm.Hm, q, t);
        store_tile_row(sm.Sm, c2, I, q, t);
    }
    __syncwarp();
    EPROF(4);
    {
        // In-place Gauss-Jordan inverse of S (symmetric positive definite: H P_pred H' + R with R > 0, so the diagonal pivots
        // are safe; the reference's partial-pivot LU differs at rounding level).  Lane i < 18 keeps row i of the working matrix
        // in registers; step c: lane c publishes its row and 1 / pivot (formed one step ahead, as soon as that entry is final,
        // so the reciprocal's latency hides behind the rest of the elimination), every lane eliminates with g = row[c] / pivot.  Column c of the working matrix is dead once eliminated and takes the column of
        // the inverse that the step creates: the same multiplications as on the augmented [S | I] in half the registers.
        // The published row is double-buffered, so one warp barrier per step.
        const bool own = lane < EN;
        const int li = own ? lane : 0;
        double ra[EN];
#pragma unroll
        for (int j = 0; j < EN; j += 2) {
            const double2 v = *reinterpret_cast<const double2*>(sm.Sm + li * LD + j);
            ra[j] = v.x + (j == lane ? QP : 0.0);
            ra[j + 1] = v.y + (j + 1 == lane ? QP : 0.0);
        }
        double dn = rcp_full(ra[0]);          // reciprocal of the NEXT pivot, formed by its owner as soon as the entry is final
#pragma unroll
        for (int c = 0; c < EN; c++) {
            double* const pb = sm.prow[c & 1];
            const bool is_piv = lane == c;
            if (is_piv) {
#pragma unroll
                for (int j = 0; j < EN; j += 2) *reinterpret_cast<double2*>(pb + j) = make_double2(ra[j], ra[j + 1]);
                pb[EN] = dn;
#pragma unroll
                for (int j = 0; j < EN; j++) ra[j] = 0.0;
            }
            __syncwarp();
            double pr[EN];
#pragma unroll
            for (int j = 0; j < EN; j += 2) {
                const double2 v = *reinterpret_cast<const double2*>(pb + j);
                pr[j] = v.x; pr[j + 1] = v.y;
            }
            const double dinv = pb[EN];
            const double g = is_piv ? -dinv : ra[c] * dinv;          // pivot lane: row <- row / pivot, entry c <- 1 / pivot
            if (c + 1 < EN) {
                ra[c + 1] = fma(-g, pr[c + 1], ra[c + 1]);
                dn = rcp_full(ra[c + 1]);
            }
#pragma unroll
            for (int j = 0; j < EN; j++)
                if (j != c + 1) ra[j] = (j == c) ? -g : fma(-g, pr[j], ra[j]);
        }
        __syncwarp();
        if (own) {
#pragma unroll
            for (int j = 0; j < EN; j += 2) *reinterpret_cast<double2*>(sm.Sm + lane * LD + j) = make_double2(ra[j], ra[j + 1]);
        }
    }
    __syncwarp();
    EPROF(5);
    // ---- K = (P_pred H') S^-1 -> Fm (F is dead);  A = I - K H -> Am (esti_P is dead) ----
    double* const Kal = sm.Fm;
#pragma unroll
    for (int I = 0; I < 3; I++) {
        double2 af[3];
        double c1[3][2], c2[3][2];
        load_tile_row(af, sm.Pp, I, q, t);
        zero_acc(c1);
        tile_row_product<false, TM_H>(c1, af, 7u, sm.Hm, q, t);
        acc_to_frag(af, c1);
        zero_acc(c2);
        tile_row_product<true, TM_ALL>(c2, af, 7u, sm.Sm, q, t);
        store_tile_row(Kal, c2, I, q, t);
#pragma unroll
        for (int K = 0; K < 3; K++) af[K] = make_double2(-c2[K][0], -c2[K][1]);
        zero_acc(c1);
        if (q == 2 * t) c1[I][0] = 1.0;
        if (q == 2 * t + 1) c1[I][1] = 1.0;
        tile_row_product<true, TM_HT>(c1, af, 7u, sm.Hm, q, t);
        store_tile_row(sm.Am, c1, I, q, t);
    }
    __syncwarp();
    EPROF(6);
    // ---- esti_x = x_pred + K (y - y_pred) (:536) ----
    double exn = 0.0;
    if (lane < EN) {
        double s = sm.xp[lane];
#pragma unroll
        for (int j = 0; j < EN; j++) s += Kal[lane * LD + j] * sm.inn[j];
        exn = s;
        ex[lane] = s;
    }
    // ---- Joseph form (:537): P = (A P_pred) A' + K R K', R = QP I, straight to global memory ----
#pragma unroll
    for (int I = 0; I < 3; I++) {
        double2 af[3], kf[3];
        double c1[3][2], c2[3][2];
        load_tile_row(af, sm.Am, I, q, t);
        zero_acc(c1);
        tile_row_product<true, TM_ALL>(c1, af, 7u, sm.Pp, q, t);
        acc_to_frag(af, c1);
        zero_acc(c2);
        tile_row_product<false, TM_ALL>(c2, af, 7u, sm.Am, q, t);
        load_tile_row(kf, Kal, I, q, t);
#pragma unroll
        for (int K = 0; K < 3; K++) { kf[K].x *= QP; kf[K].y *= QP; }
        tile_row_product<false, TM_ALL>(c2, kf, 7u, Kal, q, t);
        store_tile_row(eP, c2, I, q, t);
    }
    EPROF(7);
    // ---- world-frame disturbance (:540-545) and OCP parameters (:324-355) ----
    const double e12 = __shfl_sync(FULL_MASK, exn, 12), e13 = __shfl_sync(FULL_MASK, exn, 13), e14 = __shfl_sync(FULL_MASK, exn, 14);
    const double e15 = __shfl_sync(FULL_MASK, exn, 15), e16 = __shfl_sync(FULL_MASK, exn, 16), e17 = __shfl_sync(FULL_MASK, exn, 17);
    double sa = 0.0, ca = 1.0;               // lanes 0..2: sin / cos of the measured roll, pitch, yaw
    if (a.wf_dist && lane < 3) br2_sincos(a.meas[(size_t)inst * 12 + 3 + lane], &sa, &ca);
    const double s3 = __shfl_sync(FULL_MASK, sa, 0), c3 = __shfl_sync(FULL_MASK, ca, 0);
    const double s4 = __shfl_sync(FULL_MASK, sa, 1), c4 = __shfl_sync(FULL_MASK, ca, 1);
    const double s5 = __shfl_sync(FULL_MASK, sa, 2), c5 = __shfl_sync(FULL_MASK, ca, 2);
    if (lane == 0) {
        if (a.wf_dist) {
            double* wf = a.wf_dist + (size_t)inst * 6;
            wf[0] = (c5 * c4) * e12 + (-s5 * c3 + c5 * s4 * s3) * e13 + (s5 * s3 + c5 * c3 * s4) * e14;
            wf[1] = (s5 * c4) * e12 + (c5 * c3 + s3 * s4 * s5) * e13 + (-c5 * s3 + s4 * s5 * c3) * e14;
            wf[2] = (-s4) * e12 + (c4 * s3) * e13 + (c4 * c3) * e14;
            const double r4 = br2_rcp(c4);          // (branch-free reciprocals: no out-of-line division code in this kernel)
            wf[3] = e15 + (s5 * s4 * r4) * e16 + c3 * s4 * r4 * e17;
            wf[4] = (c3) * e16 + (s3) * e17;
            wf[5] = (s3 * r4) * e16 + (c3 * r4) * e17;
        }
        if (a.p_out) {
            double* p = a.p_out + (size_t)inst * NP;
            p[0] = a.compensate ? e12 * (1.0 / COMP) : 0.0;
            p[1] = a.compensate ? e13 * (1.0 / COMP) : 0.0;
            p[2] = a.compensate ? e14 * (1.0 / RCK) : 0.0;
            p[3] = a.compensate ? e17 * (1.0 / RCK) : 0.0;
            p[4] = 1.7182; p[5] = 0; p[6] = 5.468; p[7] = 0.4006;
            p[8] = -11.7391; p[9] = -20; p[10] = -31.8678; p[11] = -5;
            p[12] = -18.18; p[13] = -21.66; p[14] = -36.99; p[15] = -1.55;
        }
    }
    EPROF(8);
}

// profile build: cycles per phase {load, rk4 + F, P_pred, h + H, S, inverse, gain, state + Joseph, store}; zeros otherwise
void ekf_phase_cycles(unsigned long long* out, int reset)
{
    for (int i = 0; i < 12; i++) out[i] = 0;
#ifdef BR2_PROFILE
    cudaMemcpyFromSymbol(out, g_ekf_prof, sizeof(unsigned long long) * 12);
    if (reset) { unsigned long long z[12] = {0}; cudaMemcpyToSymbol(g_ekf_prof, z, sizeof z); }
#endif
}

void configure_ekf()
{
    // function attributes are per device: called from br2_batch_create with the solver's device current
    cudaFuncSetAttribute(ekf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(EkfSmem) * EKF_WARPS));
}

void launch_ekf(const EkfArgs& a, cudaStream_t s)
{
    const size_t smem = sizeof(EkfSmem) * EKF_WARPS;
    const int grid = (a.B + EKF_WARPS - 1) / EKF_WARPS;
    ekf_kernel<<<grid, EKF_WARPS * 32, smem, s>>>(a);
}

}  // namespace br2
