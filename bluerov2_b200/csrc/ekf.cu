// ekf.cu -- batched 18-state disturbance-observer EKF, one warp per instance.
//
// Restates BLUEROV2_DOB::EKF and its helpers (bluerov2_dobmpc/src/bluerov2_dob.cpp:495-545, 621-752) with the
// constants of include/bluerov2_dobmpc/bluerov2_dob.h:170-208 and the constructor (bluerov2_dob.cpp:41-65):
//   state  [eta(6), nu(6), d(6)];  process model f (:637-695) with rigid-body Coriolis terms and invM(i,i) of the
//   COUPLED mass matrix;  RK4 with k3 = f(x + k2/3) (sic, :630);  F and H by forward differences, d = 1e-6
//   (:722-752);  gain through an explicit 18x18 inverse (:535);  Joseph-form covariance update (:537).
// Lane mapping: the 19 RK4 / 19 h() evaluations of the two finite-difference Jacobians run on lanes 0..18 in
// parallel (lane 0 = unperturbed); the ten 18x18x18 products use 3 x 4 register tiles on 30 lanes (operands in shared
// memory, five 18x18 buffers per warp); the Gauss-Jordan inverse keeps one row of [S | I] per lane in registers, finds
// the pivot with three warp reductions and exchanges row POSITIONS instead of rows.  16 warps/SM (128 registers,
// 13.4 KB of shared memory per warp): the kernel is latency-bound and its time falls with every resident warp
// (8 / 12 / 16 warps/SM: 0.24 / 0.18 / 0.15 ms at B = 4096, profiles/r01k_ekf_variants.txt).  Output: esti_x, esti_P, world-frame disturbance (:540-545) and the OCP parameter
// vector handed to the solver (bluerov2_dob.cpp:324-355).
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

// process model, bluerov2_dob.cpp:637-695 (tau = K * thrusts precomputed)
__device__ void ekf_f(const double* x, const double* tau, double* xd, const double* c_Dl, const double* c_Dnl)
{
    using namespace ekfc;
    double s3, c3, s4, c4, s5, c5;
    sincos(x[3], &s3, &c3); sincos(x[4], &s4, &c4); sincos(x[5], &s5, &c5);
    xd[0] = (c5 * c4) * x[6] + (-s5 * c3 + c5 * s4 * s3) * x[7] + (s5 * s3 + c5 * c3 * s4) * x[8];
    xd[1] = (s5 * c4) * x[6] + (c5 * c3 + s3 * s4 * s5) * x[7] + (-c5 * s3 + s4 * s5 * c3) * x[8];
    xd[2] = (-s4) * x[6] + (c4 * s3) * x[7] + (c4 * c3) * x[8];
    xd[3] = x[9] + (s5 * s4 / c4) * x[10] + c3 * s4 / c4 * x[11];
    xd[4] = (c3) * x[10] + (s3) * x[11];
    xd[5] = (s3 / c4) * x[10] + (c3 / c4) * x[11];
    xd[6] = IM0 * (tau[0] + M * x[11] * x[7] - M * x[10] * x[8] - EBUOY * s4 + x[12] + c_Dl[0] * x[6] + c_Dnl[0] * fabs(x[6]) * x[6]);
    xd[7] = IM1 * (tau[1] - M * x[11] * x[6] + M * x[9] * x[8] + EBUOY * c4 * s3 + x[13] + c_Dl[1] * x[7] + c_Dnl[1] * fabs(x[7]) * x[7]);
    xd[8] = IM2 * (tau[2] + M * x[10] * x[6] - M * x[9] * x[7] + EBUOY * c4 * c3 + x[14] + c_Dl[2] * x[8] + c_Dnl[2] * fabs(x[8]) * x[8]);
    xd[9] = IM3 * (tau[3] + (Iy - Iz) * x[10] * x[11] - M * Zg * G * c4 * s3 + x[15] + c_Dl[3] * x[9] + c_Dnl[3] * fabs(x[9]) * x[9]);
    xd[10] = IM4 * (tau[4] + (Iz - Ix) * x[9] * x[11] - M * Zg * G * s4 + x[16] + c_Dl[4] * x[10] + c_Dnl[4] * fabs(x[10]) * x[10]);
    xd[11] = IM5 * (tau[5] - (Iy - Ix) * x[9] * x[10] + x[17] + c_Dl[5] * x[11] + c_Dnl[5] * fabs(x[11]) * x[11]);
#pragma unroll
    for (int i = 12; i < EN; i++) xd[i] = 0.0;
}

// measurement model, bluerov2_dob.cpp:698-719
__device__ void ekf_h(const double* x, const double* acc, double* y, const double* c_Dl, const double* c_Dnl)
{
    using namespace ekfc;
    double s3, c3, s4, c4;
    sincos(x[3], &s3, &c3); sincos(x[4], &s4, &c4);
#pragma unroll
    for (int i = 0; i < 12; i++) y[i] = x[i];
    y[12] = M0 * acc[0] - M * x[11] * x[7] + M * x[10] * x[8] + EBUOY * s4 - x[12] - c_Dl[0] * x[6] - c_Dnl[0] * fabs(x[6]) * x[6];
    y[13] = M1 * acc[1] + M * x[11] * x[6] - M * x[9] * x[8] - EBUOY * c4 * s3 - x[13] - c_Dl[1] * x[7] - c_Dnl[1] * fabs(x[7]) * x[7];
    y[14] = M2 * acc[2] - M * x[10] * x[6] + M * x[9] * x[7] - EBUOY * c4 * c3 - x[14] - c_Dl[2] * x[8] - c_Dnl[2] * fabs(x[8]) * x[8];
    y[15] = M3 * acc[3] - (Iy - Iz) * x[10] * x[11] + M * Zg * G * c4 * s3 - x[15] - c_Dl[3] * x[9] - c_Dnl[3] * fabs(x[9]) * x[9];
    y[16] = M4 * acc[4] - (Iz - Ix) * x[9] * x[11] + M * Zg * G * s4 - x[16] - c_Dl[4] * x[10] - c_Dnl[4] * fabs(x[10]) * x[10];
    y[17] = M5 * acc[5] + (Iy - Ix) * x[9] * x[10] - x[17] - c_Dl[5] * x[11] - c_Dnl[5] * fabs(x[11]) * x[11];
}

__device__ void ekf_rk4(const double* x, const double* tau, double* xn, const double* c_Dl, const double* c_Dnl)
{
    using namespace ekfc;
    // x + (k1 + 2 k2 + 2 k3 + k4) / 6 with the sum accumulated left to right as the reference's expression evaluates it
    double k[EN], xs[EN], acc[EN];
    ekf_f(x, tau, k, c_Dl, c_Dnl);
#pragma unroll
    for (int i = 0; i < EN; i++) { k[i] *= DT; acc[i] = k[i]; xs[i] = x[i] + k[i] / 2; }
    ekf_f(xs, tau, k, c_Dl, c_Dnl);
#pragma unroll
    for (int i = 0; i < EN; i++) { k[i] *= DT; acc[i] = acc[i] + 2 * k[i]; xs[i] = x[i] + k[i] / 3; }   // sic: /3 (bluerov2_dob.cpp:630)
    ekf_f(xs, tau, k, c_Dl, c_Dnl);
#pragma unroll
    for (int i = 0; i < EN; i++) { k[i] *= DT; acc[i] = acc[i] + 2 * k[i]; xs[i] = x[i] + k[i]; }
    ekf_f(xs, tau, k, c_Dl, c_Dnl);
#pragma unroll
    for (int i = 0; i < EN; i++) { k[i] *= DT; xn[i] = x[i] + (acc[i] + k[i]) / 6; }
}

// C = A * B or A * B' (18x18, shared memory, leading dimension LD).  30 lanes each own a 3 x 4 tile of C (row group
// lane / 5, column group lane % 5; the last column group is two columns wide): 12 independent accumulators per lane,
// each summed over k = 0..17 in order with fma -- the order of the oracle's matmul.
__device__ __forceinline__ void mm18(const double* A, const double* B, double* C, bool transB, int lane)
{
    if (lane < 30) {
        const int r0 = 3 * (lane / 5), c0 = 4 * (lane % 5);
        const bool edge = c0 == 16;
        double acc[3][4];
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
            for (int j = 0; j < 4; j++) acc[i][j] = 0.0;
        if (!transB) {
#pragma unroll 6
            for (int k = 0; k < EN; k++) {
                const double2 b01 = *reinterpret_cast<const double2*>(B + k * LD + c0);
                const double2 b23 = edge ? make_double2(0.0, 0.0) : *reinterpret_cast<const double2*>(B + k * LD + c0 + 2);
#pragma unroll
                for (int i = 0; i < 3; i++) {
                    const double av = A[(r0 + i) * LD + k];
                    acc[i][0] = fma(av, b01.x, acc[i][0]); acc[i][1] = fma(av, b01.y, acc[i][1]);
                    acc[i][2] = fma(av, b23.x, acc[i][2]); acc[i][3] = fma(av, b23.y, acc[i][3]);
                }
            }
        } else {
            const int j2 = edge ? c0 : c0 + 2, j3 = edge ? c0 : c0 + 3;     // clamped: the extra products are discarded
#pragma unroll 6
            for (int k = 0; k < EN; k++) {
                const double b0 = B[c0 * LD + k], b1 = B[(c0 + 1) * LD + k], b2 = B[j2 * LD + k], b3 = B[j3 * LD + k];
#pragma unroll
                for (int i = 0; i < 3; i++) {
                    const double av = A[(r0 + i) * LD + k];
                    acc[i][0] = fma(av, b0, acc[i][0]); acc[i][1] = fma(av, b1, acc[i][1]);
                    acc[i][2] = fma(av, b2, acc[i][2]); acc[i][3] = fma(av, b3, acc[i][3]);
                }
            }
        }
#pragma unroll
        for (int i = 0; i < 3; i++) {
            *reinterpret_cast<double2*>(C + (r0 + i) * LD + c0) = make_double2(acc[i][0], acc[i][1]);
            if (!edge) *reinterpret_cast<double2*>(C + (r0 + i) * LD + c0 + 2) = make_double2(acc[i][2], acc[i][3]);
        }
    }
    __syncwarp();
}

struct __align__(16) EkfSmem {
    double Fm[EN * LD], Hm[EN * LD], Pp[EN * LD], T1[EN * LD], T2[EN * LD];   // Fm doubles as Kal once P_pred is formed
    double prow[EN + 2];         // scaled pivot row of the Gauss-Jordan inverse
    double vec[EN + 2];          // innovation
    double xpv[2 * EN];          // x_pred | measurement vector y = [pose, body velocity, tau]
};

__global__ void __launch_bounds__(EKF_WARPS * 32, BR2_EKF_MINB) ekf_kernel(EkfArgs a)
{
    using namespace ekfc;
    extern __shared__ __align__(16) unsigned char smraw[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    EkfSmem& sm = reinterpret_cast<EkfSmem*>(smraw)[wib];
    const int inst = blockIdx.x * EKF_WARPS + wib;
    if (inst >= a.B) return;
    const double d = 1e-6;
    const double* c_Dl = c_DlM[a.model & 1];
    const double* c_Dnl = c_DnlM[a.model & 1];
    double* ex = a.esti_x + (size_t)inst * EN;
    double* eP = a.esti_P + (size_t)inst * EN * EN;

    // meas_y = [pose, body velocity, tau = K * thrusts] (:499-504)
    double tau[6], acc[6];
#pragma unroll
    for (int i = 0; i < 6; i++) {
        double s = 0.0;
#pragma unroll
        for (int j = 0; j < 6; j++) s += c_K[i * 6 + j] * a.thrusts[(size_t)inst * 6 + j];
        tau[i] = s;
        acc[i] = a.body_acc[(size_t)inst * 6 + i];
    }
    double x[EN], f1[EN];
#pragma unroll
    for (int i = 0; i < EN; i++) x[i] = ex[i];
    // esti_P -> shared (T2 as staging)
    for (int idx = lane; idx < EN * EN; idx += 32) sm.T2[(idx / EN) * LD + idx % EN] = eP[idx];

    // ---- F = d RK4 / dx by forward differences: lane 0 unperturbed, lane c+1 perturbs component c ----
    if (lane >= 1 && lane <= EN) {
#pragma unroll
        for (int i = 0; i < EN; i++)
            if (i == lane - 1) x[i] += d;
    }
    ekf_rk4(x, tau, f1, c_Dl, c_Dnl);
    double xp[EN];   // x_pred = RK4(esti_x) on every lane
#pragma unroll
    for (int i = 0; i < EN; i++) {
        xp[i] = __shfl_sync(FULL_MASK, f1[i], 0);
        if (lane >= 1 && lane <= EN) sm.Fm[i * LD + lane - 1] = (f1[i] - xp[i]) / d;
        if (lane == 0) sm.xpv[i] = f1[i];          // x_pred, read back by component for the state update
    }
    if (lane < 6) {
#pragma unroll
        for (int i = 0; i < 6; i++)
            if (i == lane) sm.xpv[EN + 12 + i] = tau[i];     // measurement vector: tau part (pose / velocities added below)
    }
    if (lane < 12) sm.xpv[EN + lane] = a.meas[(size_t)inst * 12 + lane];
    __syncwarp();
    // ---- P_pred = F P F' + Q (:529) ----
    mm18(sm.Fm, sm.T2, sm.T1, false, lane);
    mm18(sm.T1, sm.Fm, sm.Pp, true, lane);
    if (lane < EN) sm.Pp[lane * LD + lane] += (lane < 6) ? (DT * DT * DT * DT) / 4 : DT * DT;
    __syncwarp();
    // ---- H = dh/dx at x_pred by forward differences (:738-752) ----
#pragma unroll
    for (int i = 0; i < EN; i++) {
        x[i] = xp[i];
        if (lane >= 1 && lane <= EN && i == lane - 1) x[i] += d;
    }
    {
        double y1[EN];
        ekf_h(x, acc, y1, c_Dl, c_Dnl);
#pragma unroll
        for (int i = 0; i < EN; i++) {
            const double yp = __shfl_sync(FULL_MASK, y1[i], 0);
            if (lane >= 1 && lane <= EN) sm.Hm[i * LD + lane - 1] = (y1[i] - yp) / d;
            if (lane == 0) sm.vec[i] = sm.xpv[EN + i] - y1[i];      // innovation y - y_pred
        }
    }
    __syncwarp();
    // ---- S = H Pp H' + R ; explicit inverse by Gauss-Jordan with partial pivoting (:535) ----
    mm18(sm.Hm, sm.Pp, sm.T1, false, lane);
    mm18(sm.T1, sm.Hm, sm.T2, true, lane);
    {
        // In-place Gauss-Jordan with partial pivoting.  Lane i < 18 keeps one row of the working matrix in registers.  Rows are
        // never moved: `myrow` is the row's position in the eliminated matrix, and a pivot exchange swaps positions.  Pivot =
        // largest |entry| of column c over positions >= c, lowest position on ties (what a sequential scan finds): three warp
        // reductions on the value's bit pattern.  Column c of the working matrix is dead once it has been eliminated, so it
        // takes the one column of the inverse that step c creates -- the column of the identity that belongs to the pivot
        // row, i.e. inverse column (lane id of the pivot lane): the same multiplications as on the augmented [S | I], half
        // the registers, shared-memory traffic and FMAs.
        const bool own = lane < EN;
        const int li = own ? lane : 0;
        double ra[EN];
#pragma unroll
        for (int j = 0; j < EN; j++) ra[j] = sm.T2[li * LD + j] + (j == lane ? (DT * DT * DT * DT) / 4 : 0.0);
        int myrow = own ? lane : 99;
        __syncwarp();
#pragma unroll
        for (int c = 0; c < EN; c++) {
            const double v = (own && myrow >= c) ? fabs(ra[c]) : -1.0;
            const long long bits = __double_as_longlong(v);
            const int hi = (int)(bits >> 32);
            const int mhi = __reduce_max_sync(FULL_MASK, hi);
            const unsigned lo = (hi == mhi) ? (unsigned)bits : 0u;
            const unsigned mlo = __reduce_max_sync(FULL_MASK, lo);
            const bool cand = (hi == mhi) && ((unsigned)bits == mlo);
            const int pr = __reduce_min_sync(FULL_MASK, cand ? myrow : 99);      // position of the pivot row
            const bool is_piv = own && (myrow == pr);
            // exchange positions c <-> pr
            if (is_piv) myrow = c;
            else if (myrow == c) myrow = pr;
            // pivot lane: scale its row (its own identity entry 1 becomes 1/pivot) and publish it
            const double dinv = 1.0 / ra[c];
            if (is_piv) {
#pragma unroll
                for (int j = 0; j < EN; j++) ra[j] = (j == c) ? dinv : ra[j] * dinv;
#pragma unroll
                for (int j = 0; j < EN; j += 2) *reinterpret_cast<double2*>(sm.prow + j) = make_double2(ra[j], ra[j + 1]);
            }
            __syncwarp();
            if (!is_piv) {
                const double f = ra[c];
                if (f != 0.0) {
#pragma unroll
                    for (int j = 0; j < EN; j += 2) {
                        const double2 pa = *reinterpret_cast<const double2*>(sm.prow + j);
                        ra[j] = (j == c) ? -(f * pa.x) : ra[j] - f * pa.x;
                        ra[j + 1] = (j + 1 == c) ? -(f * pa.y) : ra[j + 1] - f * pa.y;
                    }
                }
            }
            __syncwarp();
        }
        // Si -> T2: the lane at position r holds row r of the inverse; working column c is inverse column (lane at position c)
#pragma unroll
        for (int c = 0; c < EN; c++) {
            const int col = __ffs(__ballot_sync(FULL_MASK, myrow == c)) - 1;
            if (own) sm.T2[myrow * LD + col] = ra[c];
        }
    }
    __syncwarp();
    // ---- Kal = Pp H' Si ----
    mm18(sm.Pp, sm.Hm, sm.T1, true, lane);
    double* const Kal = sm.Fm;                     // F is dead: its buffer takes the gain
    mm18(sm.T1, sm.T2, Kal, false, lane);
    // ---- esti_x = x_pred + Kal (y - y_pred) (:536) ----
    double exn = 0.0;
    if (lane < EN) {
        double s = sm.xpv[lane];
        for (int j = 0; j < EN; j++) s += Kal[lane * LD + j] * sm.vec[j];
        exn = s;
        ex[lane] = s;
    }
    // ---- Joseph form (:537): P = (I - K H) Pp (I - K H)' + K R K' ----
    mm18(Kal, sm.Hm, sm.T1, false, lane);
    if (lane < EN)
        for (int j = 0; j < EN; j++) sm.T1[lane * LD + j] = (j == lane ? 1.0 : 0.0) - sm.T1[lane * LD + j];
    __syncwarp();
    mm18(sm.T1, sm.Pp, sm.T2, false, lane);
    mm18(sm.T2, sm.T1, sm.Hm, true, lane);          // Hm reused (H is dead): (I-KH) Pp (I-KH)'
    mm18(Kal, Kal, sm.T2, true, lane);              // K K'
    for (int idx = lane; idx < EN * EN; idx += 32) {
        const int i = idx / EN, j = idx % EN;
        eP[idx] = sm.Hm[i * LD + j] + sm.T2[i * LD + j] * ((DT * DT * DT * DT) / 4);
    }
    // ---- world-frame disturbance (:540-545) and OCP parameters (:324-355) ----
    const double e12 = __shfl_sync(FULL_MASK, exn, 12), e13 = __shfl_sync(FULL_MASK, exn, 13), e14 = __shfl_sync(FULL_MASK, exn, 14);
    const double e15 = __shfl_sync(FULL_MASK, exn, 15), e16 = __shfl_sync(FULL_MASK, exn, 16), e17 = __shfl_sync(FULL_MASK, exn, 17);
    if (lane == 0) {
        if (a.wf_dist) {
            double s3, c3, s4, c4, s5, c5;
            sincos(sm.xpv[EN + 3], &s3, &c3); sincos(sm.xpv[EN + 4], &s4, &c4); sincos(sm.xpv[EN + 5], &s5, &c5);
            double* wf = a.wf_dist + (size_t)inst * 6;
            wf[0] = (c5 * c4) * e12 + (-s5 * c3 + c5 * s4 * s3) * e13 + (s5 * s3 + c5 * c3 * s4) * e14;
            wf[1] = (s5 * c4) * e12 + (c5 * c3 + s3 * s4 * s5) * e13 + (-c5 * s3 + s4 * s5 * c3) * e14;
            wf[2] = (-s4) * e12 + (c4 * s3) * e13 + (c4 * c3) * e14;
            wf[3] = e15 + (s5 * s4 / c4) * e16 + c3 * s4 / c4 * e17;
            wf[4] = (c3) * e16 + (s3) * e17;
            wf[5] = (s3 / c4) * e16 + (c3 / c4) * e17;
        }
        if (a.p_out) {
            double* p = a.p_out + (size_t)inst * NP;
            p[0] = a.compensate ? e12 / COMP : 0.0;
            p[1] = a.compensate ? e13 / COMP : 0.0;
            p[2] = a.compensate ? e14 / RCK : 0.0;
            p[3] = a.compensate ? e17 / RCK : 0.0;
            p[4] = 1.7182; p[5] = 0; p[6] = 5.468; p[7] = 0.4006;
            p[8] = -11.7391; p[9] = -20; p[10] = -31.8678; p[11] = -5;
            p[12] = -18.18; p[13] = -21.66; p[14] = -36.99; p[15] = -1.55;
        }
    }
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
