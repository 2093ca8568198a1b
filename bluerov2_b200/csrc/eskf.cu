// eskf.cu -- batched 21-state IMU error-state Kalman filter, one warp per instance.
//
// Restates BLUEROV2_STATES::ImuDoNodelet::predict / set_F / update / set_H / inject (bluerov2_states/src/Eskf.cpp:97-331) with the
// dynamics terms of Dynamics.cpp:57-165, the noise / bias set-up of Config.cpp:82-161 and launch/config/imudo.yaml.  Error state
// [dp, dv, dtheta, db_g, db_a, dg, dxi]; nominal state p, v, R (row-major 3 x 3, body -> inertial), xi; quirks kept: the velocity
// correction is injected twice (:320), biases and g are never injected, attitude measurement and the gravity direction of
// dynamics_Ma are inputs (ground truth in the reference: :170-176, Dynamics.cpp:180-181).  SO(3) exp / log written out
// (Rodrigues / quaternion log) instead of Sophus; 12 x 12 innovation covariance inverted by Gauss-Jordan with partial pivoting.
// Lane mapping: the covariance (21 x 21) and the work matrices live in shared memory; a lane owns the entries idx = lane, lane +
// 32, ... of whatever matrix is being formed; the small vector algebra of the nominal state is done redundantly by every lane.
#include "engine.h"

namespace br2 {

#define FULL_MASK 0xffffffffu
constexpr int NE = 21, NM = 12, ESKF_WARPS = 2;

namespace eskc {
constexpr double DT = 1.0 / 50.0, M = 11.26, ZGE = 0.02, G = 9.81, EBUOY = 0.661618;
constexpr double AM0 = 1.7182, AM1 = 0.0, AM2 = 5.468;
constexpr double DL0 = -11.7391, DL1 = -20.0, DL2 = -31.8678, DNL0 = -18.18, DNL1 = -21.66, DNL2 = -36.99;
}  // namespace eskc

struct EskfSmem {
    double P[NE * NE], T[NE * NE], F[NE * NE];
    double S[NM * 2 * NM];      // [S | I] for the Gauss-Jordan inverse
    double K[NE * NM];
    double y[NM], dx[NE];
};

__device__ __forceinline__ void so3_exp(const double* w, double* E)
{
    const double th2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
    double a, b;
    if (th2 < 1e-8) { a = 1.0 - th2 / 6.0; b = 0.5 - th2 / 24.0; }
    else { const double th = sqrt(th2); a = sin(th) / th; b = (1.0 - cos(th)) / th2; }
    // I + a K + b K^2, K = hat(w)
    const double x = w[0], y = w[1], z = w[2];
    E[0] = 1.0 - b * (y * y + z * z); E[1] = -a * z + b * x * y;       E[2] = a * y + b * x * z;
    E[3] = a * z + b * x * y;         E[4] = 1.0 - b * (x * x + z * z); E[5] = -a * x + b * y * z;
    E[6] = -a * y + b * x * z;        E[7] = a * x + b * y * z;         E[8] = 1.0 - b * (x * x + y * y);
}
__device__ __forceinline__ void mat3_mul(const double* A, const double* B, double* C)
{
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) C[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
}
// rotation vector of R through its unit quaternion (w >= 0)
__device__ __forceinline__ void so3_log(const double* R, double* w)
{
    const double t = R[0] + R[4] + R[8];
    double q0, q1, q2, q3;
    if (t > 0) {
        const double s = sqrt(t + 1.0) * 2;
        q0 = 0.25 * s; q1 = (R[7] - R[5]) / s; q2 = (R[2] - R[6]) / s; q3 = (R[3] - R[1]) / s;
    } else if (R[0] > R[4] && R[0] > R[8]) {
        const double s = sqrt(1.0 + R[0] - R[4] - R[8]) * 2;
        q0 = (R[7] - R[5]) / s; q1 = 0.25 * s; q2 = (R[1] + R[3]) / s; q3 = (R[2] + R[6]) / s;
    } else if (R[4] > R[8]) {
        const double s = sqrt(1.0 + R[4] - R[0] - R[8]) * 2;
        q0 = (R[2] - R[6]) / s; q1 = (R[1] + R[3]) / s; q2 = 0.25 * s; q3 = (R[5] + R[7]) / s;
    } else {
        const double s = sqrt(1.0 + R[8] - R[0] - R[4]) * 2;
        q0 = (R[3] - R[1]) / s; q1 = (R[2] + R[6]) / s; q2 = (R[5] + R[7]) / s; q3 = 0.25 * s;
    }
    if (q0 < 0) { q0 = -q0; q1 = -q1; q2 = -q2; q3 = -q3; }
    const double n = sqrt(q1 * q1 + q2 * q2 + q3 * q3);
    const double k = n < 1e-6 ? 2.0 / q0 - 2.0 * n * n / (3.0 * q0 * q0 * q0) : 2.0 * atan2(n, q0) / n;
    w[0] = k * q1; w[1] = k * q2; w[2] = k * q3;
}

// nominal state record: p[3] v[3] R[9] xi[3]
constexpr int ES_P = 0, ES_V = 3, ES_R = 6, ES_XI = 15, ES_N = 18;

__global__ void __launch_bounds__(ESKF_WARPS * 32) eskf_predict_kernel(EskfArgs a)
{
    extern __shared__ __align__(16) unsigned char raw[];
    EskfSmem& sm = reinterpret_cast<EskfSmem*>(raw)[threadIdx.x >> 5];
    const int lane = threadIdx.x & 31;
    const int inst = blockIdx.x * ESKF_WARPS + (threadIdx.x >> 5);
    if (inst >= a.B) return;
    using namespace eskc;
    double* st = a.state + (size_t)inst * ES_N;
    double* Pg = a.P + (size_t)inst * NE * NE;
    const double* imu = a.imu + (size_t)inst * 6;
    for (int i = lane; i < NE * NE; i += 32) sm.P[i] = Pg[i];
    // ---- nominal state (every lane, redundantly): Eskf.cpp:106-133 ----
    double p[3], v[3], R[9], acc[3], w[3];
#pragma unroll
    for (int i = 0; i < 3; i++) { p[i] = st[ES_P + i]; v[i] = st[ES_V + i]; acc[i] = imu[i] - a.b_a[i]; w[i] = (imu[3 + i] - a.b_g[i]); }
#pragma unroll
    for (int i = 0; i < 9; i++) R[i] = st[ES_R + i];
    const double dt = DT;
    double Ra[3];
#pragma unroll
    for (int i = 0; i < 3; i++) Ra[i] = R[3 * i] * acc[0] + R[3 * i + 1] * acc[1] + R[3 * i + 2] * acc[2];
    const double g[3] = {0.0, 0.0, -G};
#pragma unroll
    for (int i = 0; i < 3; i++) {
        p[i] = p[i] + v[i] * dt + 0.5 * Ra[i] * dt * dt + 0.5 * g[i] * dt * dt;
        v[i] = v[i] + Ra[i] * dt + g[i] * dt;
    }
    double wd[3] = {w[0] * dt, w[1] * dt, w[2] * dt}, E[9], Rn[9];
    so3_exp(wd, E);
    mat3_mul(R, E, Rn);
    // ---- F (set_F, :143-158) with the propagated attitude ----
    for (int i = lane; i < NE * NE; i += 32) sm.F[i] = (i / NE == i % NE) ? 1.0 : 0.0;
    __syncwarp();
    if (lane == 0) {
        double wn[3] = {-wd[0], -wd[1], -wd[2]}, En[9];
        so3_exp(wn, En);
        // hat(acc): [0 -a2 a1; a2 0 -a0; -a1 a0 0];  block (3,6) = -R hat(acc) dt
        const double H[9] = {0, -acc[2], acc[1], acc[2], 0, -acc[0], -acc[1], acc[0], 0};
        double RH[9];
        mat3_mul(Rn, H, RH);
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++) {
                sm.F[(0 + i) * NE + 3 + j] = (i == j) ? dt : 0.0;
                sm.F[(3 + i) * NE + 6 + j] = -RH[3 * i + j] * dt;
                sm.F[(3 + i) * NE + 12 + j] = -Rn[3 * i + j] * dt;
                sm.F[(3 + i) * NE + 15 + j] = (i == j) ? dt : 0.0;
                sm.F[(6 + i) * NE + 6 + j] = En[3 * i + j];
                sm.F[(6 + i) * NE + 9 + j] = (i == j) ? -dt : 0.0;
            }
    }
    __syncwarp();
    // ---- P = F P F' + Q ----
    for (int idx = lane; idx < NE * NE; idx += 32) {
        const int r = idx / NE, c = idx % NE;
        double s = 0.0;
        for (int k = 0; k < NE; k++) s = fma(sm.F[r * NE + k], sm.P[k * NE + c], s);
        sm.T[idx] = s;
    }
    __syncwarp();
    for (int idx = lane; idx < NE * NE; idx += 32) {
        const int r = idx / NE, c = idx % NE;
        double s = 0.0;
        for (int k = 0; k < NE; k++) s = fma(sm.T[r * NE + k], sm.F[c * NE + k], s);
        if (r == c) s += a.Qd[r];
        Pg[idx] = s;
    }
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < 3; i++) { st[ES_P + i] = p[i]; st[ES_V + i] = v[i]; }
#pragma unroll
        for (int i = 0; i < 9; i++) st[ES_R + i] = Rn[i];
    }
}

__global__ void __launch_bounds__(ESKF_WARPS * 32) eskf_update_kernel(EskfArgs a)
{
    extern __shared__ __align__(16) unsigned char raw[];
    EskfSmem& sm = reinterpret_cast<EskfSmem*>(raw)[threadIdx.x >> 5];
    const int lane = threadIdx.x & 31;
    const int inst = blockIdx.x * ESKF_WARPS + (threadIdx.x >> 5);
    if (inst >= a.B) return;
    using namespace eskc;
    double* st = a.state + (size_t)inst * ES_N;
    double* Pg = a.P + (size_t)inst * NE * NE;
    for (int i = lane; i < NE * NE; i += 32) sm.P[i] = Pg[i];
    // ---- innovation (every lane redundantly; lane 0 publishes it): Eskf.cpp:205-272 ----
    double p[3], v[3], R[9], xi[3];
#pragma unroll
    for (int i = 0; i < 3; i++) { p[i] = st[ES_P + i]; v[i] = st[ES_V + i]; xi[i] = st[ES_XI + i]; }
#pragma unroll
    for (int i = 0; i < 9; i++) R[i] = st[ES_R + i];
    const double* pm = a.gps_p + (size_t)inst * 3;
    const double* vm = a.gps_v + (size_t)inst * 3;
    const double* Rm = a.R_meas + (size_t)inst * 9;
    const double* Rg = a.R_gt + (size_t)inst * 9;
    const double* th = a.thrusts + (size_t)inst * 6;
    const double* imu = a.imu + (size_t)inst * 6;
    double y[NM];
#pragma unroll
    for (int i = 0; i < 3; i++) { y[i] = pm[i] - p[i]; y[3 + i] = vm[i] - v[i]; }
    {
        double RtRm[9], lg[3];
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
            for (int j = 0; j < 3; j++) RtRm[3 * i + j] = R[i] * Rm[j] + R[3 + i] * Rm[3 + j] + R[6 + i] * Rm[6 + j];     // R' Rm
        so3_log(RtRm, lg);
        y[6] = lg[0]; y[7] = lg[1]; y[8] = lg[2];
    }
    {
        double vB[3], ic[6], gB[3];
#pragma unroll
        for (int i = 0; i < 3; i++) {
            vB[i] = R[i] * v[0] + R[3 + i] * v[1] + R[6 + i] * v[2];                 // R' v
            ic[i] = imu[i] - a.b_a[i]; ic[3 + i] = imu[3 + i] - a.b_g[i];
            gB[i] = Rg[6 + i] * (-G);                                                // R_gt' (0, 0, -g)
        }
        const double mrb[3] = {M * ic[0] + M * ZGE * ic[4], M * ic[1] - M * ZGE * ic[3], M * ic[2]};
        const double ma[3] = {AM0 * (ic[0] + gB[0]), AM1 * (ic[1] + gB[1]), AM2 * (ic[2] + gB[2])};
        const double d[3] = {(-DL0 - DNL0 * fabs(vB[0])) * vB[0], (-DL1 - DNL1 * fabs(vB[1])) * vB[1], (-DL2 - DNL2 * fabs(vB[2])) * vB[2]};
        // roll, pitch of the estimate: tf getEulerYPR (ros_utilities.cpp:32-45)
        double roll, pitch;
        if (fabs(R[6]) >= 1.0) {
            if (R[6] < 0) { roll = atan2(R[1], R[2]); pitch = 1.5707963267948966; }
            else { roll = atan2(-R[1], -R[2]); pitch = -1.5707963267948966; }
        } else {
            pitch = -asin(R[6]);
            const double c = cos(pitch);
            roll = atan2(R[7] / c, R[8] / c);
        }
        const double wb = M * G - EBUOY;
        const double gg[3] = {wb * sin(pitch), -wb * cos(pitch) * sin(roll), -wb * cos(pitch) * cos(roll)};
        const double tau[3] = {0.7071067811847433 * th[0] + 0.7071067811847433 * th[1] - 0.7071067811919605 * th[2] - 0.7071067811919605 * th[3],
                               0.7071067811883519 * th[0] - 0.7071067811883519 * th[1] + 0.7071067811811348 * th[2] - 0.7071067811811348 * th[3],
                               th[4] + th[5]};
#pragma unroll
        for (int i = 0; i < 3; i++) y[9 + i] = tau[i] - (mrb[i] - xi[i] + ma[i] + d[i] + gg[i]);
    }
    // ---- S = H P H' + R  with H = [I3 at (0,0); I3 at (3,3); I3 at (6,6); -I3 at (9,18)]  (set_H, :303-313) ----
    // row / column m of H picks state index hm = m < 9 ? m : m + 9 with sign sg = m < 9 ? +1 : -1
    for (int idx = lane; idx < NM * NM; idx += 32) {
        const int r = idx / NM, c = idx % NM;
        const int hr = r < 9 ? r : r + 9, hc = c < 9 ? c : c + 9;
        double s = sm.P[hr * NE + hc];
        if ((r < 9) != (c < 9)) s = -s;
        if (r == c) s += a.Rd[r];
        sm.S[r * 2 * NM + c] = s;
        sm.S[r * 2 * NM + NM + c] = (r == c) ? 1.0 : 0.0;
    }
    __syncwarp();
    // ---- S^-1: Gauss-Jordan with partial pivoting on [S | I] ----
    for (int k = 0; k < NM; k++) {
        int piv = k;
        double best = fabs(sm.S[k * 2 * NM + k]);
        for (int r = k + 1; r < NM; r++) {
            const double v_ = fabs(sm.S[r * 2 * NM + k]);
            if (v_ > best) { best = v_; piv = r; }
        }
        __syncwarp();
        if (piv != k && lane < 2 * NM) {
            const double t_ = sm.S[k * 2 * NM + lane];
            sm.S[k * 2 * NM + lane] = sm.S[piv * 2 * NM + lane];
            sm.S[piv * 2 * NM + lane] = t_;
        }
        __syncwarp();
        const double ip = 1.0 / sm.S[k * 2 * NM + k];
        __syncwarp();
        if (lane < 2 * NM) sm.S[k * 2 * NM + lane] *= ip;
        __syncwarp();
        for (int idx = lane; idx < NM * 2 * NM; idx += 32) {
            const int r = idx / (2 * NM), c = idx % (2 * NM);
            if (r != k) sm.T[idx] = sm.S[idx] - sm.S[r * 2 * NM + k] * sm.S[k * 2 * NM + c];
            else sm.T[idx] = sm.S[idx];
        }
        __syncwarp();
        for (int idx = lane; idx < NM * 2 * NM; idx += 32) sm.S[idx] = sm.T[idx];
        __syncwarp();
    }
    // ---- K = P H' S^-1 (21 x 12);  dx = K y ----
    for (int idx = lane; idx < NE * NM; idx += 32) {
        const int r = idx / NM, c = idx % NM;
        double s = 0.0;
        for (int m = 0; m < NM; m++) {
            const int hm = m < 9 ? m : m + 9;
            const double ph = m < 9 ? sm.P[r * NE + hm] : -sm.P[r * NE + hm];       // (P H')[r][m]
            s = fma(ph, sm.S[m * 2 * NM + NM + c], s);
        }
        sm.K[idx] = s;
    }
    __syncwarp();
    if (lane < NE) {
        double s = 0.0;
#pragma unroll
        for (int m = 0; m < NM; m++) s = fma(sm.K[lane * NM + m], y[m], s);
        sm.dx[lane] = s;
    }
    // ---- P = (I - K H) P ----
    for (int idx = lane; idx < NE * NE; idx += 32) {
        const int r = idx / NE, c = idx % NE;
        double s = sm.P[idx];
        for (int m = 0; m < NM; m++) {
            const int hm = m < 9 ? m : m + 9;
            const double hp = m < 9 ? sm.P[hm * NE + c] : -sm.P[hm * NE + c];       // (H P)[m][c]
            s = fma(-sm.K[r * NM + m], hp, s);
        }
        Pg[idx] = s;
    }
    __syncwarp();
    // ---- inject (:315-331): velocity correction twice; biases and g untouched ----
    if (lane == 0) {
        double dth[3] = {sm.dx[6], sm.dx[7], sm.dx[8]}, E[9], Rn[9];
        so3_exp(dth, E);
        mat3_mul(R, E, Rn);
#pragma unroll
        for (int i = 0; i < 3; i++) {
            st[ES_P + i] = p[i] + sm.dx[i];
            st[ES_V + i] = v[i] + sm.dx[3 + i] + sm.dx[3 + i];
            xi[i] = xi[i] + sm.dx[18 + i];
            st[ES_XI + i] = xi[i];
        }
#pragma unroll
        for (int i = 0; i < 9; i++) st[ES_R + i] = Rn[i];
        if (a.xi_world) {
#pragma unroll
            for (int i = 0; i < 3; i++) a.xi_world[(size_t)inst * 3 + i] = Rn[3 * i] * xi[0] + Rn[3 * i + 1] * xi[1] + Rn[3 * i + 2] * xi[2];
        }
        if (a.innov) {
#pragma unroll
            for (int i = 0; i < NM; i++) a.innov[(size_t)inst * NM + i] = y[i];
        }
    }
}

void configure_eskf()
{
    cudaFuncSetAttribute(eskf_predict_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(EskfSmem) * ESKF_WARPS));
    cudaFuncSetAttribute(eskf_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(EskfSmem) * ESKF_WARPS));
}
void launch_eskf_predict(const EskfArgs& a, cudaStream_t s)
{
    eskf_predict_kernel<<<(a.B + ESKF_WARPS - 1) / ESKF_WARPS, ESKF_WARPS * 32, sizeof(EskfSmem) * ESKF_WARPS, s>>>(a);
}
void launch_eskf_update(const EskfArgs& a, cudaStream_t s)
{
    eskf_update_kernel<<<(a.B + ESKF_WARPS - 1) / ESKF_WARPS, ESKF_WARPS * 32, sizeof(EskfSmem) * ESKF_WARPS, s>>>(a);
}

}  // namespace br2
