// rls.cu -- batched recursive least squares with variable forgetting factor, the parameter estimator of the
// adaptive-MPC node: BLUEROV2_AMPC::RLSFF (bluerov2_dobmpc/src/bluerov2_ampc.cpp:731-1004), initial values from the
// constructor (:62-79) and bluerov2_ampc.h:231,239-240 (numParams 4, FF_n 5, FF_d 50).
//
// Four independent estimators per instance (axes X, Y, Z, N), one THREAD per (instance, axis) -- the state of an axis is
// 80 doubles and an update is ~150 flops, so the kernel is a pure HBM stream (640 B read + 640 B written per axis).
//   x = [body_acc, vel, 1, vel |vel|],  y = esti_x(12 | 13 | 14 | 17)  (the EKF's disturbance estimate)
//   e = y - x.theta -> short (5) / long (50) error windows, F = var_short / var_long (0/0 = NaN on the first tick
//   compares false, as in the reference), F > 0.8 ? lambda = max(lambda - .01, .5) : lambda = min(lambda + .01, 1)
//   K = P x / (lambda + x'Px),  theta += K e,  P = (P - (K x') P) / lambda
// Epilogue (axis-0 thread): the OCP parameter vector BLUEROV2_AMPC::solve builds (:340-382) -- when `compensate` is
// false only p[0..3] = 0 is written and p[4..15] are left untouched, which is what the reference's brace placement does.
#include "engine.h"

namespace br2 {

__device__ __forceinline__ double rls_window_push_var(double* w, double* count, int cap, double e)
{
    int n = (int)*count;
    if (n < cap) {
        w[n++] = e;
    } else {                                   // push_back + erase(begin): the window stays oldest-first
        for (int i = 1; i < cap; i++) w[i - 1] = w[i];
        w[cap - 1] = e;
    }
    *count = n;
    double sum = 0.0;
    for (int i = 0; i < n; i++) sum += w[i];
    const double mean = sum / n;
    double var = 0.0;
    for (int i = 0; i < n; i++) var += (w[i] - mean) * (w[i] - mean);
    return var / n;
}

__global__ void __launch_bounds__(128) rls_kernel(RlsArgs a)
{
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    const int inst = gid >> 2, ax = gid & 3;
    if (inst >= a.B) return;
    const int yi = ax < 3 ? 12 + ax : 17, ai = ax < 3 ? ax : 5, vi = ax < 3 ? 6 + ax : 11;
    double* s = a.state + ((size_t)inst * 4 + ax) * RLS_STRIDE;
    const double y = a.esti_x[(size_t)inst * 18 + yi];
    const double acc = a.body_acc[(size_t)inst * 6 + ai];
    const double vel = a.meas[(size_t)inst * 12 + vi];
    const double x[4] = {acc, vel, 1.0, vel * fabs(vel)};
    double th[4], P[16];
#pragma unroll
    for (int i = 0; i < 4; i++) th[i] = s[RLS_THETA + i];
#pragma unroll
    for (int i = 0; i < 16; i++) P[i] = s[RLS_P + i];
    double pred = 0.0;
#pragma unroll
    for (int i = 0; i < 4; i++) pred += x[i] * th[i];
    const double e = y - pred;
    const double vn = rls_window_push_var(s + RLS_EN, s + RLS_NN, RLS_FFN, e);
    const double vd = rls_window_push_var(s + RLS_ED, s + RLS_ND, RLS_FFD, e);
    const double F = vn / vd;
    double lam = s[RLS_LAMBDA];
    if (F > 0.8) lam = (lam - 0.01 >= 0.5) ? lam - 0.01 : 0.5;
    else lam = (lam + 0.01 <= 1) ? lam + 0.01 : 1;
    double Px[4], K[4], xPx = 0.0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        double t = 0.0;
#pragma unroll
        for (int j = 0; j < 4; j++) t += P[i * 4 + j] * x[j];
        Px[i] = t;
    }
#pragma unroll
    for (int i = 0; i < 4; i++) xPx += x[i] * Px[i];
#pragma unroll
    for (int i = 0; i < 4; i++) { K[i] = Px[i] / (lam + xPx); th[i] += K[i] * e; }
    s[RLS_F] = F;
    s[RLS_LAMBDA] = lam;
#pragma unroll
    for (int i = 0; i < 4; i++) s[RLS_THETA + i] = th[i];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) {
            double t = 0.0;
#pragma unroll
            for (int k = 0; k < 4; k++) t += (K[i] * x[k]) * P[k * 4 + j];
            s[RLS_P + i * 4 + j] = (P[i * 4 + j] - t) / lam;
        }
    if (a.p_out) {
        // theta(2) of the four axes sits in four consecutive lanes
        const unsigned m = __activemask();
        const int l0 = (threadIdx.x & 31) & ~3;
        const double t0 = __shfl_sync(m, th[2], l0), t1 = __shfl_sync(m, th[2], l0 + 1), t2 = __shfl_sync(m, th[2], l0 + 2),
                     t3 = __shfl_sync(m, th[2], l0 + 3);
        if (ax == 0) {
            double* p = a.p_out + (size_t)inst * NP;
            if (!a.compensate) {
                p[0] = p[1] = p[2] = p[3] = 0.0;
            } else {
                const double comp = 0.032546960744430276;
                p[0] = t0 / comp; p[1] = t1 / comp; p[2] = t2 / RC; p[3] = t3 / RC;
                p[4] = 1.7182; p[5] = 0; p[6] = 5.468; p[7] = 0.4006;
                p[8] = -11.7391; p[9] = -20; p[10] = -31.8678; p[11] = -5;
                p[12] = -18.18; p[13] = -21.66; p[14] = -36.99; p[15] = -1.55;
            }
        }
    }
}

void launch_rls(const RlsArgs& a, cudaStream_t s)
{
    const int threads = a.B * 4;
    rls_kernel<<<(threads + 127) / 128, 128, 0, s>>>(a);
}

}  // namespace br2
