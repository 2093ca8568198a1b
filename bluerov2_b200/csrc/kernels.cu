// kernels.cu -- hand-written sm_100a kernels of the batched SQP-RTI step.
//
//   linearize_kernel : ERK4 + forward sensitivities of the 6-DOF model, one 16-lane team per (instance, stage)
//                      (replaces acados' ERK integrator driving bluerov2_expl_vde_forw, acados_solver_bluerov2.c:
//                      310-318,633-639) -> stage records G_k = [A_k | B_k], b_k in HBM.
//   ipm_kernel       : one warp per OCP instance, persistent over the batch.  Mehrotra predictor-corrector
//                      primal-dual IPM on the box-constrained OCP-QP of the RTI step; every Newton system is an
//                      LQR solved by a Riccati recursion over the horizon (replaces acados full condensing + HPIPM
//                      dense IPM, acados_solver_bluerov2.c:146,664-668).  Epilogue: full SQP step on (X,U),
//                      u0 and the 4->6 thrust allocation (bluerov2_dob.cpp:388-395).
//
// Arithmetic: fp64 throughout (casadi_real = double in the reference).  The algorithm is the one restated
// in oracle/bluerov2_oracle.c (feasible-start, residual-form Newton steps, split primal/dual step lengths);
// see DESIGN.md for the lane mapping and the per-stage byte/flop budget.
#include "engine.h"
#include <stdint.h>

namespace br2 {

#define FULL_MASK 0xffffffffu

// ------------------------------------------------------------------------------------------------------
// linearisation
// ------------------------------------------------------------------------------------------------------
// team lane c: 0 = state column (evolves by f); 1..9 = Sx columns 3..11; 10..13 = Su columns 0..3;
// 14, 15 = helpers that write the constant [I;0] columns 0, 1 of A (positions do not enter f).
__global__ void __launch_bounds__(128) linearize_kernel(SolveArgs a)
{
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const int c = lane & 15;
    const unsigned tmask = 0xffffu << (lane & 16);
    const int total = a.B * a.N;
    int team = gtid >> 4;
    const bool live = team < total;
    if (!live) team = total - 1;            // keep the lanes alive for the team shuffles; stores are predicated
    const int inst = team / a.N, k = team - inst * a.N;

    const double* p = a.p + (size_t)inst * a.p_inst_stride + (size_t)k * a.p_stage_stride;
    ModelConst mc;
    {
        double pl[NP];
#pragma unroll
        for (int i = 0; i < NP; i++) pl[i] = __ldg(p + i);
        mc.set(pl);
    }
    const double h = __ldg(a.Ts + k);
    const double* Xk = a.X + ((size_t)inst * (a.N + 1) + k) * NX;
    double u[NU];
#pragma unroll
    for (int i = 0; i < NU; i++) u[i] = __ldg(a.U + ((size_t)inst * a.N + k) * NU + i);

    double col0[NX], cur[NX], acc[NX], kk[NX];
#pragma unroll
    for (int i = 0; i < NX; i++) {
        col0[i] = (c == 0) ? __ldg(Xk + i) : ((c >= 1 && c <= 9 && i == c + 2) ? 1.0 : 0.0);
        cur[i] = col0[i];
        acc[i] = 0.0;
    }
    const int su_col = c - 10;   // valid for c in 10..13

#pragma unroll 1
    for (int s = 0; s < 4; s++) {
        // stage state (only components 3..11 enter f and J)
        double xs[NX];
        xs[0] = xs[1] = xs[2] = 0.0;
#pragma unroll
        for (int i = 3; i < NX; i++) xs[i] = __shfl_sync(tmask, cur[i], 0, 16);
        // one sincos per lane: lanes 0,1,2 of the team own phi, theta, psi
        double sn, cs;
        const int which = c % 3;
        sincos(which == 0 ? xs[3] : (which == 1 ? xs[4] : xs[5]), &sn, &cs);
        Trig t;
        t.sphi = __shfl_sync(tmask, sn, 0, 16); t.cphi = __shfl_sync(tmask, cs, 0, 16);
        t.sth = __shfl_sync(tmask, sn, 1, 16);  t.cth = __shfl_sync(tmask, cs, 1, 16);
        t.spsi = __shfl_sync(tmask, sn, 2, 16); t.cpsi = __shfl_sync(tmask, cs, 2, 16);
        if (c == 0) {
            ode(xs, u, mc, t, kk);
        } else {
            Jac J;
            jac_of(xs, mc, t, J);
            jac_mul(J, cur, kk);
            if (c >= 10 && c <= 13) ju_add(mc, su_col, kk);
        }
        const double bw = (s == 0 || s == 3) ? (1.0 / 6.0) : (1.0 / 3.0);
        const double cn = (s == 2) ? 1.0 : 0.5;     // c_{s+1} of the classical tableau
#pragma unroll
        for (int i = 0; i < NX; i++) {
            acc[i] += bw * kk[i];
            cur[i] = col0[i] + cn * h * kk[i];
        }
    }
#pragma unroll
    for (int i = 0; i < NX; i++) cur[i] = col0[i] + h * acc[i];

    if (!live) return;
    double* Gk = a.G + ((size_t)inst * a.N + k) * GREC;
    if (c == 0) {
        const double* Xn = Xk + NX;
#pragma unroll
        for (int i = 0; i < NX; i++) Gk[G_B_OFF + i] = cur[i] - __ldg(Xn + i);
#pragma unroll
        for (int l = 0; l < NX; l++) Gk[l * 16 + 2] = (l == 2) ? 1.0 : 0.0;
    } else if (c <= 13) {
#pragma unroll
        for (int l = 0; l < NX; l++) Gk[l * 16 + c + 2] = cur[l];
    } else {
        const int j = c - 14;
#pragma unroll
        for (int l = 0; l < NX; l++) Gk[l * 16 + j] = (l == j) ? 1.0 : 0.0;
        if (c == 14) {
#pragma unroll
            for (int i = 204; i < GREC; i++) Gk[i] = 0.0;
        }
    }
}

void launch_linearize(const SolveArgs& a, cudaStream_t s)
{
    const long long threads = (long long)a.B * a.N * 16;
    const int block = 128;
    const int grid = (int)((threads + block - 1) / block);
    linearize_kernel<<<grid, block, 0, s>>>(a);
}

// ------------------------------------------------------------------------------------------------------
// Riccati interior-point kernel
// ------------------------------------------------------------------------------------------------------
constexpr int IPM_WARPS = 4;

struct __align__(16) WarpSmem {
    double Gs[12 * 16];   // stage matrix [A|B]
    double Ws[12 * 16];   // W = P+ [A|B]; reused for the symmetrisation of the new P
    double Hu[4 * 16];    // rows 12..15 of H = [A|B]' W  (B'PA | B'PB)
    double Ks[4 * 16];    // feedback gain rows
    double vec[32];       // broadcast vectors (pi+, p+ / z)
    double gs[4];         // g = gh + B'p+
    double rt[4];         // barrier-augmented input Hessian diagonal
};

__device__ __forceinline__ double warp_min(double v)
{
#pragma unroll
    for (int o = 16; o; o >>= 1) v = fmin(v, __shfl_xor_sync(FULL_MASK, v, o));
    return v;
}
__device__ __forceinline__ double warp_max(double v)
{
#pragma unroll
    for (int o = 16; o; o >>= 1) v = fmax(v, __shfl_xor_sync(FULL_MASK, v, o));
    return v;
}
__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(FULL_MASK, v, o);
    return v;
}

// 4x4 Cholesky of the symmetric matrix with lower entries m (row-major lower: 00 10 11 20 21 22 30 31 32 33).
// Outputs strictly-lower entries and reciprocal diagonal.  Returns false when a pivot is not positive.
struct Chol4 { double l10, l20, l21, l30, l31, l32, i0, i1, i2, i3; };
__device__ __forceinline__ bool chol4(const double* m, Chol4& L)
{
    bool ok = true;
    double d = m[0];
    ok &= d > 0.0; L.i0 = rsqrt(d);
    L.l10 = m[1] * L.i0; L.l20 = m[3] * L.i0; L.l30 = m[6] * L.i0;
    d = m[2] - L.l10 * L.l10;
    ok &= d > 0.0; L.i1 = rsqrt(d);
    L.l21 = (m[4] - L.l20 * L.l10) * L.i1; L.l31 = (m[7] - L.l30 * L.l10) * L.i1;
    d = m[5] - L.l20 * L.l20 - L.l21 * L.l21;
    ok &= d > 0.0; L.i2 = rsqrt(d);
    L.l32 = (m[8] - L.l30 * L.l20 - L.l31 * L.l21) * L.i2;
    d = m[9] - L.l30 * L.l30 - L.l31 * L.l31 - L.l32 * L.l32;
    ok &= d > 0.0; L.i3 = rsqrt(d);
    return ok;
}
// v <- (L L')^-1 v
__device__ __forceinline__ void chol4_solve(const Chol4& L, double* v)
{
    v[0] = v[0] * L.i0;
    v[1] = (v[1] - L.l10 * v[0]) * L.i1;
    v[2] = (v[2] - L.l20 * v[0] - L.l21 * v[1]) * L.i2;
    v[3] = (v[3] - L.l30 * v[0] - L.l31 * v[1] - L.l32 * v[2]) * L.i3;
    v[3] = v[3] * L.i3;
    v[2] = (v[2] - L.l32 * v[3]) * L.i2;
    v[1] = (v[1] - L.l21 * v[2] - L.l31 * v[3]) * L.i1;
    v[0] = (v[0] - L.l10 * v[1] - L.l20 * v[2] - L.l30 * v[3]) * L.i0;
}
__device__ __forceinline__ void load_chol(const double* Fk, Chol4& L)
{
    L.l10 = Fk[F_L_OFF + 0]; L.l20 = Fk[F_L_OFF + 1]; L.l21 = Fk[F_L_OFF + 2];
    L.l30 = Fk[F_L_OFF + 3]; L.l31 = Fk[F_L_OFF + 4]; L.l32 = Fk[F_L_OFF + 5];
    L.i0 = Fk[F_ID_OFF + 0]; L.i1 = Fk[F_ID_OFF + 1]; L.i2 = Fk[F_ID_OFF + 2]; L.i3 = Fk[F_ID_OFF + 3];
}

// Lane roles inside the warp that owns one instance: r = lane & 15 is a row of the 16x16 stage matrix
// H = [A|B]' P+ [A|B] (+ diag(Q, R~)): r < 12 state rows, r >= 12 input rows; h = lane >> 4 selects the
// column half [8h, 8h+8) the lane accumulates.
struct Inst {
    const SolveArgs& a;
    int inst, lane, r, h, N;
    WarpSmem& sm;
    const double* G;
    double* F;
    double* V;
    const double* Xlin;
    const double* Ulin;
    const double* yref;
    __device__ Inst(const SolveArgs& a_, int inst_, int lane_, WarpSmem& sm_)
        : a(a_), inst(inst_), lane(lane_), r(lane_ & 15), h(lane_ >> 4), N(a_.N), sm(sm_)
    {
        G = a.G + (size_t)inst * N * GREC;
        F = a.F + (size_t)inst * N * FREC;
        V = a.V + (size_t)inst * (N + 1) * VREC;
        Xlin = a.X + (size_t)inst * (N + 1) * NX;
        Ulin = a.U + (size_t)inst * N * NU;
        yref = a.yref + (size_t)inst * (N + 1) * NY;
    }
};

// E0: cold start of the IPM iterate (qp_solver_warm_start 0): du = 0 pushed strictly inside the box, slacks
// exactly consistent, lam = mu0 / t.
__device__ void ipm_init(Inst& I)
{
    const double thr = 1e-1, mu0 = 1.0;
    for (int idx = I.lane; idx < 4 * I.N; idx += 32) {
        const int k = idx >> 2, e = idx & 3;
        const double uk = I.Ulin[k * NU + e];
        const double lb = I.a.lbu[e] - uk, ub = I.a.ubu[e] - uk;
        double v = fmin(fmax(0.0, lb + thr), ub - thr);
        if (ub - lb < 2 * thr) v = 0.5 * (lb + ub);
        const double tl = v - lb, tu = ub - v;
        double* Vk = I.V + (size_t)k * VREC;
        Vk[V_V + e] = v; Vk[V_TL + e] = tl; Vk[V_TU + e] = tu;
        Vk[V_LL + e] = mu0 / tl; Vk[V_LU + e] = mu0 / tu;
    }
    __syncwarp();
}

// Forward sweep.  mode 0: roll-out of the iterate  x+ = A x + B v + b, x_0 = x0 - X_0 (writes V_X);
// mode 1: Newton step  ddu = -K ddx - Lam^-1 g,  ddx+ = A ddx + B ddu, ddx_0 = 0 (writes V_DX, V_DV).
// Returns max |b| in mode 0.
__device__ double forward_sweep(Inst& I, int mode)
{
    const int r = I.r, h = I.h, N = I.N;
    WarpSmem& sm = I.sm;
    double xr = 0.0;          // component r of the propagated vector (r < 12), replicated in both halves
    double bmax = 0.0;
    if (mode == 0 && r < 12) xr = I.a.x0[(size_t)I.inst * NX + r] - I.Xlin[r];
    for (int k = 0; k < N; k++) {
        const double* Gk = I.G + (size_t)k * GREC;
        double* Vk = I.V + (size_t)k * VREC;
        // row r of [A|B], my half
        double g[8];
        if (r < 12) {
            const double2* g2 = reinterpret_cast<const double2*>(Gk + r * 16 + 8 * h);
#pragma unroll
            for (int j = 0; j < 4; j++) { double2 t = g2[j]; g[2 * j] = t.x; g[2 * j + 1] = t.y; }
        }
        if (r < 12 && h == 0) {
            sm.vec[r] = xr;
            Vk[mode ? V_DX + r : V_X + r] = xr;
        }
        __syncwarp();
        // input part z[12..15]
        if (r >= 12) {
            const int e = r - 12;
            double dv;
            if (mode == 0) {
                dv = Vk[V_V + e];
            } else {
                // ddu_e = -kff_e - sum_j K[e][j] ddx_j ; halves split j
                const double* Fk = I.F + (size_t)k * FREC;
                double s = 0.0;
#pragma unroll
                for (int j = 0; j < 6; j++) s = fma(Fk[(6 * h + j) * 4 + e], sm.vec[6 * h + j], s);
                s += __shfl_xor_sync(0xf000f000u, s, 16);
                Chol4 L;
                load_chol(Fk, L);
                double gg[4] = {Vk[V_G + 0], Vk[V_G + 1], Vk[V_G + 2], Vk[V_G + 3]};
                chol4_solve(L, gg);
                dv = -gg[e] - s;
                if (h == 0) Vk[V_DV + e] = dv;
            }
            if (h == 0) sm.vec[12 + e] = dv;
        }
        __syncwarp();
        double s = 0.0;
        if (r < 12) {
            const double2* z2 = reinterpret_cast<const double2*>(sm.vec + 8 * h);
#pragma unroll
            for (int j = 0; j < 4; j++) { double2 z = z2[j]; s = fma(g[2 * j], z.x, s); s = fma(g[2 * j + 1], z.y, s); }
        }
        s += __shfl_xor_sync(FULL_MASK, s, 16);
        if (mode == 0 && r < 12) {
            const double b = Gk[G_B_OFF + r];
            s += b;
            bmax = fmax(bmax, fabs(b));
        }
        xr = s;
        __syncwarp();
    }
    if (r < 12 && h == 0) I.V[(size_t)N * VREC + (mode ? V_DX + r : V_X + r)] = xr;
    __syncwarp();
    return mode == 0 ? warp_max(bmax) : 0.0;
}

// Backward sweep.  factor = true: costate recursion of the iterate (pi), reduced gradient gu, Riccati
// factorisation with the current barrier diagonal, and the vector recursion for the predictor rhs (gh = gu).
// factor = false: vector recursion only, rhs gh = gu - cl/tl + cu/tu (corrector).
// Returns false if a Cholesky pivot failed.
__device__ bool backward_sweep(Inst& I, bool factor)
{
    const int r = I.r, h = I.h, N = I.N, lane = I.lane;
    WarpSmem& sm = I.sm;
    const SolveArgs& a = I.a;
    double Prow[12];
    double pi_r = 0.0, pv_r = 0.0;
    bool ok = true;
    if (factor) {
        // terminal: P_N = diag(We), pi_N = We (x_N + X_N - yref_N)
#pragma unroll
        for (int l = 0; l < 12; l++) Prow[l] = (r < 12 && l == r) ? a.We[r < 12 ? r : 0] : 0.0;
        if (r < 12) pi_r = a.We[r] * (I.V[(size_t)N * VREC + V_X + r] + I.Xlin[N * NX + r] - I.yref[N * NY + r]);
    }
    for (int k = N - 1; k >= 0; k--) {
        const double* Gk = I.G + (size_t)k * GREC;
        double* Fk = I.F + (size_t)k * FREC;
        double* Vk = I.V + (size_t)k * VREC;
        const double tsk = a.Ts[k];
        // ---- stage scalars ----
        double qd = 0.0, qx = 0.0;          // state rows: Q_rr and  Q_rr x_r + q_r
        double rt = 0.0, gu = 0.0, gh = 0.0;  // input rows
        if (r < 12) {
            if (factor) {
                qd = tsk * a.W[r];
                qx = qd * (Vk[V_X + r] + I.Xlin[k * NX + r] - I.yref[k * NY + r]);
            }
        } else {
            const int e = r - 12;
            const double tl = Vk[V_TL + e], tu = Vk[V_TU + e], ll = Vk[V_LL + e], lu = Vk[V_LU + e];
            if (factor) {
                const double rd = tsk * a.W[12 + e];
                rt = rd + ll / tl + lu / tu;
                gu = rd * (Vk[V_V + e] + I.Ulin[k * NU + e] - I.yref[k * NY + 12 + e]);   // + B'pi+ below
            } else {
                gh = Vk[V_GU + e] - Vk[V_CL + e] / tl + Vk[V_CU + e] / tu;
            }
        }
        double gc[12];                       // column r of [A|B]
        double s4[4], kc[4];
        double at = 0.0, bt = 0.0;
        if (factor) {
            // ---- stage matrix to shared memory ----
            {
                const double2* src = reinterpret_cast<const double2*>(Gk);
                double2* dst = reinterpret_cast<double2*>(sm.Gs);
#pragma unroll
                for (int c = 0; c < 3; c++) dst[lane + 32 * c] = src[lane + 32 * c];
            }
            if (r < 12 && h == 0) { sm.vec[r] = pi_r; sm.vec[16 + r] = pv_r; }
            __syncwarp();
            // ---- W = P+ [A|B], row r, my column half ----
            double w[8];
#pragma unroll
            for (int j = 0; j < 8; j++) w[j] = 0.0;
            if (r < 12) {
#pragma unroll
                for (int l = 0; l < 12; l++) {
                    const double2* g2 = reinterpret_cast<const double2*>(sm.Gs + l * 16 + 8 * h);
                    const double pl = Prow[l];
#pragma unroll
                    for (int j = 0; j < 4; j++) { double2 t = g2[j]; w[2 * j] = fma(pl, t.x, w[2 * j]); w[2 * j + 1] = fma(pl, t.y, w[2 * j + 1]); }
                }
                double2* w2 = reinterpret_cast<double2*>(sm.Ws + r * 16 + 8 * h);
#pragma unroll
                for (int j = 0; j < 4; j++) w2[j] = make_double2(w[2 * j], w[2 * j + 1]);
            }
#pragma unroll
            for (int l = 0; l < 12; l++) gc[l] = sm.Gs[l * 16 + r];
            __syncwarp();
            // ---- H row r = sum_l G[l][r] W[l][:], my half; [A|B]'pi+, [A|B]'p+ ----
            double hr[8];
#pragma unroll
            for (int j = 0; j < 8; j++) hr[j] = 0.0;
#pragma unroll
            for (int l = 0; l < 12; l++) {
                const double2* w2 = reinterpret_cast<const double2*>(sm.Ws + l * 16 + 8 * h);
                const double gl = gc[l];
#pragma unroll
                for (int j = 0; j < 4; j++) { double2 t = w2[j]; hr[2 * j] = fma(gl, t.x, hr[2 * j]); hr[2 * j + 1] = fma(gl, t.y, hr[2 * j + 1]); }
                at = fma(gl, sm.vec[l], at);
                bt = fma(gl, sm.vec[16 + l], bt);
            }
            if (r >= 12) {
                const int e = r - 12;
                double2* h2 = reinterpret_cast<double2*>(sm.Hu + e * 16 + 8 * h);
#pragma unroll
                for (int j = 0; j < 4; j++) h2[j] = make_double2(hr[2 * j], hr[2 * j + 1]);
                gu += at;
                if (h == 0) { sm.gs[e] = gu + bt; sm.rt[e] = rt; Vk[V_GU + e] = gu; Vk[V_G + e] = gu + bt; }
            }
            __syncwarp();
            // ---- Lam = B'PB + R~, Cholesky (every lane, redundantly) ----
            double m[10];
            m[0] = sm.Hu[0 * 16 + 12] + sm.rt[0];
            m[1] = sm.Hu[1 * 16 + 12]; m[2] = sm.Hu[1 * 16 + 13] + sm.rt[1];
            m[3] = sm.Hu[2 * 16 + 12]; m[4] = sm.Hu[2 * 16 + 13]; m[5] = sm.Hu[2 * 16 + 14] + sm.rt[2];
            m[6] = sm.Hu[3 * 16 + 12]; m[7] = sm.Hu[3 * 16 + 13]; m[8] = sm.Hu[3 * 16 + 14]; m[9] = sm.Hu[3 * 16 + 15] + sm.rt[3];
            Chol4 L;
            ok &= chol4(m, L);
            if (lane == 0) {
                Fk[F_L_OFF + 0] = L.l10; Fk[F_L_OFF + 1] = L.l20; Fk[F_L_OFF + 2] = L.l21;
                Fk[F_L_OFF + 3] = L.l30; Fk[F_L_OFF + 4] = L.l31; Fk[F_L_OFF + 5] = L.l32;
                Fk[F_ID_OFF + 0] = L.i0; Fk[F_ID_OFF + 1] = L.i1; Fk[F_ID_OFF + 2] = L.i2; Fk[F_ID_OFF + 3] = L.i3;
            }
            // ---- K column r = Lam^-1 (B'PA)[:, r] ----
            const int rc = r < 12 ? r : 0;
#pragma unroll
            for (int e = 0; e < 4; e++) { s4[e] = sm.Hu[e * 16 + rc]; kc[e] = s4[e]; }
            chol4_solve(L, kc);
            double g0 = sm.gs[0], g1 = sm.gs[1], g2v = sm.gs[2], g3 = sm.gs[3];
            if (r < 12) {
                pv_r = bt - (kc[0] * g0 + kc[1] * g1 + kc[2] * g2v + kc[3] * g3);
                pi_r = qx + at;
                if (h == 0) {
#pragma unroll
                    for (int e = 0; e < 4; e++) sm.Ks[e * 16 + r] = kc[e];
                    *reinterpret_cast<double2*>(Fk + r * 4) = make_double2(kc[0], kc[1]);
                    *reinterpret_cast<double2*>(Fk + r * 4 + 2) = make_double2(kc[2], kc[3]);
                }
            }
            __syncwarp();
            // ---- P = Q + A'PA - (B'PA)' K ----
            if (r < 12) {
#pragma unroll
                for (int e = 0; e < 4; e++) {
                    const double2* k2 = reinterpret_cast<const double2*>(sm.Ks + e * 16 + 8 * h);
                    const double se = s4[e];
#pragma unroll
                    for (int j = 0; j < 4; j++) { double2 t = k2[j]; hr[2 * j] = fma(-se, t.x, hr[2 * j]); hr[2 * j + 1] = fma(-se, t.y, hr[2 * j + 1]); }
                }
#pragma unroll
                for (int j = 0; j < 8; j++)
                    if (8 * h + j == r) hr[j] += qd;
                double2* p2 = reinterpret_cast<double2*>(sm.Ws + r * 16 + 8 * h);
#pragma unroll
                for (int j = 0; j < 4; j++) p2[j] = make_double2(hr[2 * j], hr[2 * j + 1]);
            }
            __syncwarp();
            if (r < 12) {
#pragma unroll
                for (int l = 0; l < 12; l++) Prow[l] = 0.5 * (sm.Ws[r * 16 + l] + sm.Ws[l * 16 + r]);
            }
            __syncwarp();
        } else {
            // ---- vector recursion only: g = gh + B'p+,  p = A'p+ - K'g ----
            if (r < 12 && h == 0) sm.vec[r] = pv_r;
            __syncwarp();
            // column r of [A|B], halves split the 12 rows
            double sacc = 0.0;
#pragma unroll
            for (int l = 0; l < 6; l++) sacc = fma(Gk[(6 * h + l) * 16 + r], sm.vec[6 * h + l], sacc);
            bt = sacc + __shfl_xor_sync(FULL_MASK, sacc, 16);
            if (r >= 12 && h == 0) { const int e = r - 12; sm.gs[e] = gh + bt; Vk[V_G + e] = gh + bt; }
            __syncwarp();
            if (r < 12) {
                const double2 k01 = *reinterpret_cast<const double2*>(Fk + r * 4);
                const double2 k23 = *reinterpret_cast<const double2*>(Fk + r * 4 + 2);
                pv_r = bt - (k01.x * sm.gs[0] + k01.y * sm.gs[1] + k23.x * sm.gs[2] + k23.y * sm.gs[3]);
            }
            __syncwarp();
        }
    }
    return __all_sync(FULL_MASK, ok);
}

__device__ __forceinline__ double step_to_boundary(double v, double dv)
{
    return dv < 0.0 ? -v / dv : 2.0;   // 2 = "not blocking" (callers clamp at 1)
}

__global__ void __launch_bounds__(IPM_WARPS * 32) ipm_kernel(SolveArgs a)
{
    __shared__ WarpSmem smem[IPM_WARPS];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    WarpSmem& sm = smem[wib];
    const int N = a.N, nb = 4 * N;

    for (;;) {
        int inst = 0;
        if (lane == 0) inst = atomicAdd(a.work_counter, 1);
        inst = __shfl_sync(FULL_MASK, inst, 0);
        if (inst >= a.B) break;
        Inst I(a, inst, lane, sm);

        ipm_init(I);
        const double bmax = forward_sweep(I, 0);

        int status = 2, it = 0;
        double mu = 0.0, res_stat = 0.0, stat_scale = 1.0;
        // mu and stationarity of the starting point are produced by the first factor sweep's by-products
        for (it = 0; it < a.max_iter; it++) {
            // ---------- B1: factorisation + predictor rhs ----------
            if (!backward_sweep(I, true)) { status = 4; break; }
            if (it == 0) {
                // mu and stationarity residual of the starting point (later iterations get them from E2)
                double s = 0.0, rs = 0.0;
                for (int idx = lane; idx < nb; idx += 32) {
                    const double* Vk = I.V + (size_t)(idx >> 2) * VREC;
                    const int e = idx & 3;
                    const double tl = Vk[V_TL + e], tu = Vk[V_TU + e], ll = Vk[V_LL + e], lu = Vk[V_LU + e];
                    s += ll * tl + lu * tu;
                    rs = fmax(rs, fabs(Vk[V_GU + e] - ll + lu));
                }
                mu = warp_sum(s) / (2.0 * nb);
                res_stat = warp_max(rs);
                stat_scale = fmax(1.0, res_stat);
                if (mu < a.tol && res_stat < a.tol * stat_scale) { status = 0; break; }
            }
            // ---------- F1: affine step ----------
            forward_sweep(I, 1);
            // ---------- E1: affine step length, sigma, corrector rhs ----------
            double a_aff = 1.0;
            for (int idx = lane; idx < nb; idx += 32) {
                const double* Vk = I.V + (size_t)(idx >> 2) * VREC;
                const int e = idx & 3;
                const double tl = Vk[V_TL + e], tu = Vk[V_TU + e], ll = Vk[V_LL + e], lu = Vk[V_LU + e], dv = Vk[V_DV + e];
                const double dll = -ll - ll * dv / tl, dlu = -lu + lu * dv / tu;
                a_aff = fmin(a_aff, fmin(fmin(step_to_boundary(tl, dv), step_to_boundary(tu, -dv)),
                                         fmin(step_to_boundary(ll, dll), step_to_boundary(lu, dlu))));
            }
            a_aff = warp_min(a_aff);
            double mu_aff = 0.0;
            for (int idx = lane; idx < nb; idx += 32) {
                const double* Vk = I.V + (size_t)(idx >> 2) * VREC;
                const int e = idx & 3;
                const double tl = Vk[V_TL + e], tu = Vk[V_TU + e], ll = Vk[V_LL + e], lu = Vk[V_LU + e], dv = Vk[V_DV + e];
                const double dll = -ll - ll * dv / tl, dlu = -lu + lu * dv / tu;
                mu_aff += (ll + a_aff * dll) * (tl + a_aff * dv) + (lu + a_aff * dlu) * (tu - a_aff * dv);
            }
            mu_aff = warp_sum(mu_aff) / (2.0 * nb);
            double sigma = mu_aff / mu;
            sigma = sigma * sigma * sigma;
            for (int idx = lane; idx < nb; idx += 32) {
                double* Vk = I.V + (size_t)(idx >> 2) * VREC;
                const int e = idx & 3;
                const double tl = Vk[V_TL + e], tu = Vk[V_TU + e], ll = Vk[V_LL + e], lu = Vk[V_LU + e], dv = Vk[V_DV + e];
                const double dll = -ll - ll * dv / tl, dlu = -lu + lu * dv / tu;
                Vk[V_CL + e] = sigma * mu - dv * dll;
                Vk[V_CU + e] = sigma * mu + dv * dlu;
            }
            __syncwarp();
            // ---------- B2 / F2: corrector ----------
            backward_sweep(I, false);
            forward_sweep(I, 1);
            // ---------- E2: step lengths and update ----------
            double ap = 2.0, ad = 2.0;
            for (int idx = lane; idx < nb; idx += 32) {
                const double* Vk = I.V + (size_t)(idx >> 2) * VREC;
                const int e = idx & 3;
                const double tl = Vk[V_TL + e], tu = Vk[V_TU + e], ll = Vk[V_LL + e], lu = Vk[V_LU + e], dv = Vk[V_DV + e];
                const double dll = Vk[V_CL + e] / tl - ll - ll * dv / tl, dlu = Vk[V_CU + e] / tu - lu + lu * dv / tu;
                ap = fmin(ap, fmin(step_to_boundary(tl, dv), step_to_boundary(tu, -dv)));
                ad = fmin(ad, fmin(step_to_boundary(ll, dll), step_to_boundary(lu, dlu)));
            }
            ap = fmin(1.0, warp_min(ap));
            ad = fmin(1.0, warp_min(ad));
            const double tau = fmin(fmax(0.995, 1.0 - mu), 1.0 - 1e-8);
            ap = fmin(1.0, tau * ap);
            ad = fmin(1.0, tau * ad);
            // update; mu and the stationarity residual of the NEW iterate follow without another sweep: the reduced
            // gradient is affine in du and the Newton equation gives  d(gu) = -gh - (ll/tl + lu/tu) ddu  stage-locally.
            double s_mu = 0.0, s_rs = 0.0;
            for (int idx = lane; idx < nb; idx += 32) {
                double* Vk = I.V + (size_t)(idx >> 2) * VREC;
                const int e = idx & 3;
                const double tl = Vk[V_TL + e], tu = Vk[V_TU + e], ll = Vk[V_LL + e], lu = Vk[V_LU + e], dv = Vk[V_DV + e];
                const double cl = Vk[V_CL + e], cu = Vk[V_CU + e], gu = Vk[V_GU + e];
                const double dll = cl / tl - ll - ll * dv / tl, dlu = cu / tu - lu + lu * dv / tu;
                const double gh = gu - cl / tl + cu / tu;
                const double dgu = -gh - (ll / tl + lu / tu) * dv;
                const double tln = tl + ap * dv, tun = tu - ap * dv, lln = ll + ad * dll, lun = lu + ad * dlu;
                Vk[V_V + e] += ap * dv;
                Vk[V_TL + e] = tln;
                Vk[V_TU + e] = tun;
                Vk[V_LL + e] = lln;
                Vk[V_LU + e] = lun;
                s_mu += lln * tln + lun * tun;
                s_rs = fmax(s_rs, fabs(gu + ap * dgu - lln + lun));
            }
            mu = warp_sum(s_mu) / (2.0 * nb);
            res_stat = warp_max(s_rs);
            for (int idx = lane; idx < 12 * (N + 1); idx += 32) {
                double* Vk = I.V + (size_t)(idx / 12) * VREC;
                const int e = idx % 12;
                Vk[V_X + e] += ap * Vk[V_DX + e];
            }
            __syncwarp();
            if (mu < a.tol && res_stat < a.tol * stat_scale) { status = 0; it++; break; }
        }

        // ---------- epilogue: full SQP step, u0, thrust allocation ----------
        bool finite = true;
        for (int idx = lane; idx < nb; idx += 32) finite &= isfinite(I.V[(size_t)(idx >> 2) * VREC + V_V + (idx & 3)]);
        for (int idx = lane; idx < 12 * (N + 1); idx += 32) finite &= isfinite(I.V[(size_t)(idx / 12) * VREC + V_X + idx % 12]);
        finite = __all_sync(FULL_MASK, finite);
        double* Xo = a.X + (size_t)inst * (N + 1) * NX;
        double* Uo = a.U + (size_t)inst * N * NU;
        if (finite) {
            for (int idx = lane; idx < nb; idx += 32) Uo[idx] += I.V[(size_t)(idx >> 2) * VREC + V_V + (idx & 3)];
            for (int idx = lane; idx < 12 * (N + 1); idx += 32) Xo[idx] += I.V[(size_t)(idx / 12) * VREC + V_X + idx % 12];
        } else {
            status = 1;
        }
        __syncwarp();
        if (lane == 0) {
            double u0[4] = {Uo[0], Uo[1], Uo[2], Uo[3]};
            double th[6];
            thrust_alloc(u0, th);
#pragma unroll
            for (int i = 0; i < 4; i++) a.u0[(size_t)inst * 4 + i] = u0[i];
#pragma unroll
            for (int i = 0; i < 6; i++) a.thrust[(size_t)inst * 6 + i] = th[i];
            a.status[inst] = status;
            a.iters[inst] = it;
            atomicAdd(a.iter_total, (unsigned long long)it);
            a.info[(size_t)inst * 4 + 0] = mu;
            a.info[(size_t)inst * 4 + 1] = res_stat;
            a.info[(size_t)inst * 4 + 2] = bmax;
            a.info[(size_t)inst * 4 + 3] = stat_scale;
        }
        __syncwarp();
    }
}

void launch_ipm(const SolveArgs& a, int sm_count, cudaStream_t s)
{
    cudaMemsetAsync(a.work_counter, 0, sizeof(int), s);
    const int warps_needed = a.B;
    int blocks = (warps_needed + IPM_WARPS - 1) / IPM_WARPS;
    const int max_blocks = sm_count * 8;
    if (blocks > max_blocks) blocks = max_blocks;
    ipm_kernel<<<blocks, IPM_WARPS * 32, 0, s>>>(a);
}

}  // namespace br2
