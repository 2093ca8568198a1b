// kernels.cu -- hand-written sm_100a kernels of the batched SQP-RTI step.
//
//   linearize_kernel : ERK4 + forward sensitivities of the 6-DOF model, one warp per round of 16 (instance, stage) pairs
//                      in three phases (state trajectory / Jacobians / sensitivities); replaces acados' ERK integrator
//                      driving bluerov2_expl_vde_forw (acados_solver_bluerov2.c:310-318,633-639) -> stage records
//                      G_k = [A_k | B_k], b_k, cost gradients in HBM.
//   pdas_kernel      : the QP of the RTI step, one warp per OCP instance, persistent over the batch (atomic work queue in
//                      longest-first order).  Primal-dual active-set iteration on Riccati solves: the unconstrained LQR first (one
//                      backward factor sweep + one closed-loop roll-out that tests its own candidate against the box), inputs
//                      outside the box pinned and the LQR re-solved, a costate sweep checking the multiplier signs; a candidate
//                      that passes satisfies the KKT conditions of the strictly convex QP, i.e. is the minimiser HPIPM converges
//                      to (replaces acados full condensing + HPIPM dense IPM, acados_solver_bluerov2.c:146,664-668).
//   ipm_kernel       : the fallback list of pdas_kernel (normally empty): Mehrotra predictor-corrector primal-dual IPM, every
//                      Newton system an LQR solved by the same Riccati sweeps.
//                      Both: the two 12x16 stage products of the factorisation, the gain K = Lam^-1 [H_ux | g] and the Riccati
//                      update run on the fp64 tensor-core instruction (DMMA, mma.sync.m8n8k4.f64; on B200 the same 64 FMA/clk/SM
//                      as DFMA, reached from 4 warps/SM with 8x fewer issue slots and fragment-resident operands,
//                      profiles/r01_fp64_probe.txt); Lam^-1 is formed entry by entry, one 3x3 cofactor per lane; stage records
//                      are staged HBM -> shared memory by cp.async two stages ahead.  Epilogue: full SQP step on (X, U), u0 and
//                      the 4->6 thrust allocation (bluerov2_dob.cpp:388-395), hint and place in the next solve's visiting order.
//   exchange_kernel, shard_wait_kernel : sharded batches, peer-to-peer exchange of the thrust vectors.
//
// Arithmetic: fp64 throughout (casadi_real = double in the reference).  The interior-point iteration is the one restated
// in oracle/bluerov2_oracle.c (feasible-start, residual-form Newton steps, split primal/dual step lengths);
// see DESIGN.md for the lane mapping, the per-stage byte/flop budget and what bounds each kernel.
#include "engine.h"
#include <stdint.h>

// occupancy knobs (defaults = the measured best; scripts/gpu_tune.sh rebuilds with others)
#ifndef BR2_NSLOT
#define BR2_NSLOT 3
#endif
#ifndef BR2_IPM_WARPS
#define BR2_IPM_WARPS 4
#endif
#ifndef BR2_IPM_MINB
#define BR2_IPM_MINB (16 / BR2_IPM_WARPS)
#endif
#ifndef BR2_PDAS_MINB
#define BR2_PDAS_MINB (16 / BR2_IPM_WARPS)
#endif
#ifndef BR2_PF_DIST
#define BR2_PF_DIST 0
#endif
#ifndef BR2_PDL
#define BR2_PDL 0                   // programmatic dependent launch of the QP kernel behind the lineariser (experiment, see launch_pdas)
#endif
#ifndef BR2_LIN_MINB
#define BR2_LIN_MINB 12
#endif

namespace br2 {

#define FULL_MASK 0xffffffffu

// ------------------------------------------------------------------------------------------------------
// linearisation
// ------------------------------------------------------------------------------------------------------
// v2: one warp per round of LRND (instance, stage) pairs, three phases with three lane mappings so that nothing is
// computed redundantly:
//   A   lane per stage: the RK4 STATE trajectory only (x_1..x_4, 3 sincos + f per RK stage; the sensitivities do not
//       feed back into it) -> RK-stage states and trig values to shared memory, Phi -> b_k, cost gradients -> tail of G_k.
//   J   lane per (stage, RK stage), LSUB stages at a time: the 48 structural non-zeros of df/dx at x_s -> shared memory.
//   B   8-lane team per stage, a lane carries two of the 16 column slots of Z = [A|B] through the four RK stages with the
//       Jacobians read back from shared memory (LDS.128, broadcast inside the team): 48 FMAs per column and RK stage.
//       slots 0..12 = columns 3..15 of Z (Sx columns 3..11, Su); slots 13..15 = the constant columns 0..2 ([I;0]:
//       positions do not enter f, so J annihilates them and the same code writes them).
// (v1 formed the Jacobian, the sincos and f once per lane of the team, 8x redundantly: profiles/r01f_lin_ncu_summary.txt.)
#ifndef BR2_LIN_RND
#define BR2_LIN_RND 16
#endif
#ifndef BR2_LIN_SUB
#define BR2_LIN_SUB 4
#endif
constexpr int LRND = BR2_LIN_RND;              // stages per warp round (16 or 32)
constexpr int LSUB = BR2_LIN_SUB;              // stages per Jacobian sub-round (4 or 8)
#ifndef BR2_LIN_TRIG
#define BR2_LIN_TRIG 1                         // 1: phase A hands its sin / cos values to the Jacobian phase; 0: recomputed there (smaller xt)
#endif
constexpr int XT_F = BR2_LIN_TRIG ? 15 : 9;    // x[3..11] (, sphi cphi sth cth spsi cpsi)
constexpr int XT_SSTRIDE = XT_F * LRND + 8;    // +8: RK-stage planes land on different banks for the J-phase reads
constexpr int CS_F = 18;                       // imx imy imz imn | dl[4] | dnl[4] | ju[5] | h
constexpr int JREC = 50;                       // 48 + 2: 128-bit accesses of consecutive records are conflict-free
static_assert(LRND % LSUB == 0 && LSUB % 4 == 0 && LSUB * 4 <= 32 && LRND <= 32, "lineariser tiling");

struct __align__(16) LinSmem {
    double jb[LSUB * 4 * JREC];
    double xt[4 * XT_SSTRIDE];
    double cs[CS_F * LRND];
};
__device__ __forceinline__ int xt_idx(int s, int f, int slot) { return s * XT_SSTRIDE + f * LRND + slot; }

__device__ __forceinline__ void st256(double* p, double a, double b, double c, double d)
{
    asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}

#define BR2_JAC_FIELDS(X) \
    X(j03) X(j04) X(j05) X(j06) X(j07) X(j08) X(j13) X(j14) X(j15) X(j16) X(j17) X(j18) \
    X(j23) X(j24) X(j26) X(j27) X(j28) X(j33) X(j34) X(j35) X(j3a) X(j3b) X(j43) X(j4a) \
    X(j4b) X(j53) X(j54) X(j5a) X(j5b) X(j64) X(j66) X(j73) X(j74) X(j77) X(j83) X(j84) \
    X(j88) X(j93) X(j94) X(j9a) X(j9b) X(ja4) X(ja9) X(jab) X(jb9) X(jba) X(jbb)
// 47 named fields + one pad = 24 double2

__device__ __forceinline__ void jac_store(const Jac& J, double* dst)
{
    double v[48];
    int n = 0;
#define X(f) v[n++] = J.f;
    BR2_JAC_FIELDS(X)
#undef X
    v[47] = 0.0;
#pragma unroll
    for (int i = 0; i < 24; i++) reinterpret_cast<double2*>(dst)[i] = make_double2(v[2 * i], v[2 * i + 1]);
}
__device__ __forceinline__ void jac_load(Jac& J, const double* src)
{
    double v[48];
#pragma unroll
    for (int i = 0; i < 24; i++) {
        const double2 d = reinterpret_cast<const double2*>(src)[i];
        v[2 * i] = d.x; v[2 * i + 1] = d.y;
    }
    int n = 0;
#define X(f) J.f = v[n++];
    BR2_JAC_FIELDS(X)
#undef X
}

// phase A for one stage (lane-private)
__device__ __forceinline__ void lin_phase_a(const SolveArgs& a, LinSmem& sm, int slot, int gs, bool live)
{
    const int il = gs / a.N, k = gs - il * a.N, inst = a.lo + il;
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
    double x0v[NX], xs[NX], acc[NX], u[NU];
#pragma unroll
    for (int i = 0; i < NX; i++) { x0v[i] = __ldg(Xk + i); xs[i] = x0v[i]; acc[i] = 0.0; }
#pragma unroll
    for (int i = 0; i < NU; i++) u[i] = __ldg(a.U + ((size_t)inst * a.N + k) * NU + i);
#pragma unroll 1
    for (int s = 0; s < 4; s++) {
        Trig t;
        trig_of(xs, t);
        double f[NX];
        ode(xs, u, mc, t, f);
        double* xt = sm.xt + xt_idx(s, 0, slot);
#pragma unroll
        for (int i = 0; i < 9; i++) xt[i * LRND] = xs[3 + i];
        if (BR2_LIN_TRIG) {
            xt[9 * LRND] = t.sphi; xt[10 * LRND] = t.cphi; xt[11 * LRND] = t.sth;
            xt[12 * LRND] = t.cth; xt[13 * LRND] = t.spsi; xt[14 * LRND] = t.cpsi;
        }
        const double bw = (s == 0 || s == 3) ? (1.0 / 6.0) : (1.0 / 3.0);
        const double cn = (s == 2) ? 1.0 : 0.5;     // c_{s+1} of the classical tableau
#pragma unroll
        for (int i = 0; i < NX; i++) {
            acc[i] = fma(bw, f[i], acc[i]);
            xs[i] = fma(cn * h, f[i], x0v[i]);
        }
    }
    double* cs = sm.cs + slot;
    cs[0 * LRND] = mc.imx; cs[1 * LRND] = mc.imy; cs[2 * LRND] = mc.imz; cs[3 * LRND] = mc.imn;
#pragma unroll
    for (int i = 0; i < 4; i++) { cs[(4 + i) * LRND] = mc.dl[i]; cs[(8 + i) * LRND] = mc.dnl[i]; }
    cs[12 * LRND] = mc.ju_surge; cs[13 * LRND] = mc.ju_sway; cs[14 * LRND] = mc.ju_heave;
    cs[15 * LRND] = mc.ju_yaw2; cs[16 * LRND] = mc.ju_yaw4; cs[17 * LRND] = h;
    if (!live) return;
    // tail of the stage record: b_k, qlin_k, rlin_k, Ts_k, pad (32 doubles, 32-byte aligned)
    double* Gt = a.S + ((size_t)inst * (a.N + 1) + k) * SREC + S_G + G_B_OFF;
    const double* yr = yref_row(a, inst, k);
    const double* Xn = Xk + NX;
    double tl[32];
#pragma unroll
    for (int i = 0; i < NX; i++) {
        tl[i] = fma(h, acc[i], x0v[i]) - __ldg(Xn + i);
        tl[12 + i] = h * a.W[i] * (x0v[i] - __ldg(yr + i));
    }
#pragma unroll
    for (int i = 0; i < NU; i++) tl[24 + i] = h * a.W[NX + i] * (u[i] - __ldg(yr + NX + i));
    tl[28] = h; tl[29] = tl[30] = tl[31] = 0.0;
#pragma unroll
    for (int i = 0; i < 8; i++) st256(Gt + 4 * i, tl[4 * i], tl[4 * i + 1], tl[4 * i + 2], tl[4 * i + 3]);
}
static_assert(G_B_OFF == 192 && G_QLIN == 204 && G_RLIN == 216 && G_TS == 220 && GREC == 224, "tail layout");

// per-tick reset of the work queues, the order / fallback counters, the order parity and the tick index
__device__ __forceinline__ void tick_housekeeping(int* ctr)
{
    ctr[CTR_QUEUE] = 0; ctr[CTR_QUEUE + 1] = 0; ctr[CTR_QUEUE + 2] = 0; ctr[CTR_QUEUE + 3] = 0;
    ctr[CTR_HARD] = 0; ctr[CTR_EASY] = 0; ctr[CTR_PARITY] ^= 1; ctr[CTR_FB] = 0; ctr[CTR_FBQ] = 0; ctr[CTR_TICK] += 1;
}
__global__ void tick_begin_kernel(int* ctr) { tick_housekeeping(ctr); }
void launch_tick_begin(int* ctr, cudaStream_t s) { tick_begin_kernel<<<1, 1, 0, s>>>(ctr); }

__global__ void __launch_bounds__(32, BR2_LIN_MINB) linearize_kernel(SolveArgs a)
{
    __shared__ LinSmem sm;
    const int lane = threadIdx.x;
    const int total = (a.hi - a.lo) * a.N;
    const int base = blockIdx.x * LRND;
    if (blockIdx.x == 0 && lane == 0) {
        // housekeeping for the IPM kernel that follows in the stream: reset its work-queue counters, flip the order buffers
        if (a.housekeeping) tick_housekeeping(a.ctr);
    }
#if BR2_PDL
    // programmatic dependent launch: the QP kernel behind this one may be scheduled as soon as every block of this grid has got here
    // (its blocks then wait in griddepcontrol.wait for this grid to finish and flush); hides the launch gap behind the last wave
    asm volatile("griddepcontrol.launch_dependents;");
#endif
#ifdef BR2_PROFILE
    // cycles per phase of the lineariser into slots 13..15 (state trajectory / Jacobians / sensitivities + stores)
    long long lprof_t0 = clock64();
#define LPROF(i) do { const long long t_ = clock64(); if (lane == 0 && a.prof) atomicAdd(a.prof + 13 + (i), (unsigned long long)(t_ - lprof_t0)); lprof_t0 = clock64(); } while (0)
#else
#define LPROF(i) do { } while (0)
#endif
    if (lane < LRND) {
        const int gs = base + lane;
        lin_phase_a(a, sm, lane, gs < total ? gs : total - 1, gs < total);
    }
    __syncwarp();
    LPROF(0);

    const int tm = lane >> 3, l = lane & 7;
    // my two column slots: Z column, unit seed row (or none), Su column (or none)
    const int sa = 2 * l, sb = 2 * l + 1;
    const int zca = sa < 13 ? sa + 3 : sa - 13, zcb = sb < 13 ? sb + 3 : sb - 13;
    const int seed_a = zca < NX ? zca : -1, seed_b = zcb < NX ? zcb : -1;
    const int su_a = zca - NX, su_b = zcb - NX;

#pragma unroll 1
    for (int r = 0; r < LRND / LSUB; r++) {
        // ---- phase J: Jacobians of LSUB stages x 4 RK stages ----
        if (lane < LSUB * 4) {
            const int s = lane / LSUB, i = lane % LSUB;
            const int slot = r * LSUB + i;
            double xs[NX];
            xs[0] = xs[1] = xs[2] = 0.0;
            const double* xt = sm.xt + xt_idx(s, 0, slot);
#pragma unroll
            for (int f = 0; f < 9; f++) xs[3 + f] = xt[f * LRND];
            Trig t;
            if (BR2_LIN_TRIG) {
                t.sphi = xt[9 * LRND]; t.cphi = xt[10 * LRND]; t.sth = xt[11 * LRND];
                t.cth = xt[12 * LRND]; t.spsi = xt[13 * LRND]; t.cpsi = xt[14 * LRND];
            } else {
                trig_of(xs, t);
            }
            const double* cs = sm.cs + slot;
            ModelConst mc;
            mc.imx = cs[0 * LRND]; mc.imy = cs[1 * LRND]; mc.imz = cs[2 * LRND]; mc.imn = cs[3 * LRND];
#pragma unroll
            for (int j = 0; j < 4; j++) { mc.dl[j] = cs[(4 + j) * LRND]; mc.dnl[j] = cs[(8 + j) * LRND]; }
            Jac J;
            jac_of(xs, mc, t, J);
            jac_store(J, sm.jb + (i * 4 + s) * JREC);
        }
        __syncwarp();
        LPROF(1);
        // ---- phase B: sensitivities, 4 stages per pass ----
#pragma unroll 1
        for (int pp = 0; pp < LSUB / 4; pp++) {
            const int si = pp * 4 + tm;
            const int slot = r * LSUB + si;
            const int gs = base + slot;
            const double* cs = sm.cs + slot;
            const double h = cs[17 * LRND];
            const double ja6 = su_a == 0 ? cs[12 * LRND] : 0.0, jb6 = su_b == 0 ? cs[12 * LRND] : 0.0;
            const double ja7 = su_a == 1 ? cs[13 * LRND] : 0.0, jb7 = su_b == 1 ? cs[13 * LRND] : 0.0;
            const double ja8 = su_a == 2 ? cs[14 * LRND] : 0.0, jb8 = su_b == 2 ? cs[14 * LRND] : 0.0;
            const double jab_ = su_a == 1 ? cs[15 * LRND] : (su_a == 3 ? cs[16 * LRND] : 0.0);
            const double jbb_ = su_b == 1 ? cs[15 * LRND] : (su_b == 3 ? cs[16 * LRND] : 0.0);
            double ca[NX], cb[NX], aa[NX], ab[NX];
#pragma unroll
            for (int i = 0; i < NX; i++) {
                ca[i] = (i == seed_a) ? 1.0 : 0.0;
                cb[i] = (i == seed_b) ? 1.0 : 0.0;
                aa[i] = 0.0; ab[i] = 0.0;
            }
#pragma unroll 1
            for (int s = 0; s < 4; s++) {
                Jac J;
                jac_load(J, sm.jb + (si * 4 + s) * JREC);
                double ka[NX], kb[NX];
                jac_mul_add(J, ca, ja6, ja7, ja8, jab_, ka);
                jac_mul_add(J, cb, jb6, jb7, jb8, jbb_, kb);
                const double bw = (s == 0 || s == 3) ? (1.0 / 6.0) : (1.0 / 3.0);
                const double ch = ((s == 2) ? 1.0 : 0.5) * h;
#pragma unroll
                for (int i = 0; i < NX; i++) {
                    aa[i] = fma(bw, ka[i], aa[i]);
                    ab[i] = fma(bw, kb[i], ab[i]);
                }
                if (s < 3) {                        // the last stage's slope only enters the weighted sum
#pragma unroll
                    for (int i = 3; i < NX; i++) {  // rows 0..2 never feed back (columns 0..2 of J are zero)
                        ca[i] = fma(ch, ka[i], (i == seed_a) ? 1.0 : 0.0);
                        cb[i] = fma(ch, kb[i], (i == seed_b) ? 1.0 : 0.0);
                    }
                }
            }
            if (gs < total) {
                const int gi = gs / a.N;
                double* Gk = a.S + ((size_t)(a.lo + gi) * (a.N + 1) + (gs - gi * a.N)) * SREC + S_G;
                double* da = Gk + (((zca >> 3)) << 5) + ((zca & 7) << 2);     // g_off(0, zc); row block ki adds 64
                double* db = Gk + (((zcb >> 3)) << 5) + ((zcb & 7) << 2);
#pragma unroll
                for (int ki = 0; ki < 3; ki++) {
                    double oa[4], ob[4];
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        const int i = 4 * ki + j;
                        oa[j] = fma(h, aa[i], (i == seed_a) ? 1.0 : 0.0);
                        ob[j] = fma(h, ab[i], (i == seed_b) ? 1.0 : 0.0);
                    }
                    st256(da + 64 * ki, oa[0], oa[1], oa[2], oa[3]);
                    st256(db + 64 * ki, ob[0], ob[1], ob[2], ob[3]);
                }
            }
        }
        __syncwarp();
        LPROF(2);
    }
}

void launch_linearize(const SolveArgs& a, cudaStream_t s)
{
    const long long total = (long long)(a.hi - a.lo) * a.N;
    const int grid = (int)((total + LRND - 1) / LRND);
    linearize_kernel<<<grid, 32, 0, s>>>(a);
}


// ------------------------------------------------------------------------------------------------------
// Riccati interior-point kernel
// ------------------------------------------------------------------------------------------------------
// Lane coordinates (q, t) = (lane >> 2, lane & 3) are the DMMA thread coordinates (layout.h).  Recurring layouts of
// a 12- or 16-vector v over the warp:
//   "row layout"   lane (q,t) holds v[4 ki + t], ki = 0..3   (what a B fragment column / a dot over t needs)
//   "quad layout"  lane (q,*) holds v[q] and v[8 + q]        (what falls out of a reduction over t / a C fragment row)
constexpr int IPM_WARPS = BR2_IPM_WARPS;

__device__ __forceinline__ double shfl(double v, int src) { return __shfl_sync(FULL_MASK, v, src); }
__device__ __forceinline__ double shfl_x(double v, int m) { return __shfl_xor_sync(FULL_MASK, v, m); }

// D(8x8) += A(8x4) B(4x8), fp64 tensor-core instruction
__device__ __forceinline__ void dmma(double (&c)[2], double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

// -DBR2_PROFILE_STAGE (with -DBR2_PROFILE): cycles of the segments of ONE factor-sweep stage into slots 5..12 (the interior-point
// phases, unused on the fast path): scripts/stage_profile.py
#if defined(BR2_PROFILE) && defined(BR2_PROFILE_STAGE)
#define SPROF_START() long long sprof_t0 = clock64()
#define SPROF(i)                                                                                          \
    do {                                                                                                  \
        const long long sprof_t1 = clock64();                                                             \
        if (lane == 0 && a.prof) atomicAdd(a.prof + 5 + (i), (unsigned long long)(sprof_t1 - sprof_t0));  \
        sprof_t0 = sprof_t1;                                                                              \
    } while (0)
#else
#define SPROF_START()
#define SPROF(i)
#endif
// reciprocal good to ~1e-7 relative: the hardware seed (MUFU.RCP64H, ~20 bits) and one Newton step
__device__ __forceinline__ double rcp_approx(double x)
{
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    return y * fma(-x, y, 2.0);
}

// slot of entry (a, c), a <= c, of the symmetric 4x4 Lam^-1 in the interior-point F record (upper triangle, row by row: 10 doubles
// from F_L_OFF, where round 1 kept the Cholesky factor and its inverse diagonal)
__device__ __forceinline__ int linv_slot(int a, int c) { return 4 * a + c - (a * (a + 1)) / 2; }

// 1 / x to about an ulp without the division's special-case branch: the hardware seed and two Newton steps
__device__ __forceinline__ double rcp_newton(double x)
{
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    y = fma(y, fma(-x, y, 1.0), y);
    return fma(y, fma(-x, y, 1.0), y);
}

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


// ---- per-warp staging ring: the stage records of a sweep are copied HBM -> shared memory NSLOT-1 stages ahead of the compute ----
// cp.async (LDGSTS, 16 bytes per lane, L2 only) with commit / wait groups.  Round 1 used one-lane TMA bulk copies completing on
// mbarriers: ptxas wraps every cp.async.bulk issued from non-uniform registers in an elect / R2UR / branch loop and the parity
// bookkeeping came on top -- ~60 of the 150..600 instructions of EVERY stage of EVERY sweep were pipeline mechanics
// (profiles/r01o_ipm_ncu_summary.txt: IMAD 18.7 %, the kernel is issue-bound).  With the three records of a stage laid end to end
// (layout.h: S_k = [V | G | F]) a stage costs 4..6 copy instructions, one commit, one wait and one warp barrier.
constexpr int NSLOT = BR2_NSLOT;     // ring depth: records of NSLOT-1 stages are in flight ahead of the one being computed
constexpr int REC_BYTES = SREC * 8;
// xch: per-warp exchange buffer of the factor sweep: H[:, 12..15] (16 rows x 4) + g (4)
struct __align__(128) WarpSmem { double st[NSLOT][SREC]; double xch[80]; };
static_assert(sizeof(WarpSmem) % 128 == 0, "ring slots stay 128-byte aligned");
enum { P_V = 1, P_G = 2, P_F = 4 };  // which parts of a stage record a sweep reads

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp16(uint32_t dst, const char* src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
// the same under a predicate (no branch: a divergent-branch region around a copy is a scheduling barrier at the top of every stage)
__device__ __forceinline__ void cp16_if(bool pr, uint32_t dst, const char* src)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %2, 0;\n\t@p cp.async.cg.shared.global [%0], [%1], 16;\n\t}" ::"r"(dst), "l"(src), "r"((int)pr) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int PENDING>
__device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(PENDING) : "memory"); }

// One stage record, the parts named by PARTS: 16-byte chunk c = lane + 32 j sits at byte 512 j + 16 lane of the record, so
//   V = j 0,   G = j 1..3 and lanes < 16 of j 4,   F = lanes >= 16 of j 4 and lanes < 16 of j 5.
// dst / src are the lane's own addresses (slot base / record base + 16 lane).
template <int PARTS>
__device__ __forceinline__ void stage_copy(uint32_t dst, const char* src, int lane, bool on = true)
{
    if (PARTS & P_V) cp16_if(on, dst, src);
    if (PARTS & P_G) {
        cp16_if(on, dst + 512, src + 512);
        cp16_if(on, dst + 1024, src + 1024);
        cp16_if(on, dst + 1536, src + 1536);
    }
    if ((PARTS & P_G) && (PARTS & P_F)) cp16_if(on, dst + 2048, src + 2048);
    else if (PARTS & P_G) cp16_if(on && lane < 16, dst + 2048, src + 2048);
    else if (PARTS & P_F) cp16_if(on && lane >= 16, dst + 2048, src + 2048);
    if (PARTS & P_F) cp16_if(on && lane < 16, dst + 2560, src + 2560);
}

// Per-instance pointers and the sweep pipeline
struct Inst {
    const SolveArgs& a;
    WarpSmem& sm;
    int inst, lane, q, t, N;
    double* S;                  // the instance's stage records S_0 .. S_N (layout.h)
    double* V;                  // = S + S_V: V record of stage k at V + k * SREC
    const double* Xlin;
    const double* Ulin;
    const char* src;            // pipeline: the lane's global address inside the next record to copy
    uint32_t dst0;              // pipeline: the lane's shared address inside ring slot 0
    int rslot;                  // pipeline: ring slot of the stage being read (the next refill goes into the one read before it)
    __device__ Inst(const SolveArgs& a_, WarpSmem& sm_, int inst_, int lane_)
        : a(a_), sm(sm_), inst(inst_), lane(lane_), q(lane_ >> 2), t(lane_ & 3), N(a_.N)
    {
        S = a.S + (size_t)inst * (N + 1) * SREC;
        V = S + S_V;
        Xlin = a.X + (size_t)inst * (N + 1) * NX;
        Ulin = a.U + (size_t)inst * N * NU;
        dst0 = smem_u32(&sm.st[0][0]) + 16 * lane;
        src = nullptr;
    }
    // Stages are visited in sequence i = 0..N-1 (record k = i forward, k = N-1-i backward); ring slot = i % NSLOT.
    // begin(): everything this warp wrote to global so far is ordered before the copies (warp barrier), then fill the ring.
    template <int PARTS, bool BACKWARD>
    __device__ __forceinline__ void begin()
    {
        __syncwarp();
        src = reinterpret_cast<const char*>(S) + 16 * lane + (BACKWARD ? (size_t)(N - 1) * REC_BYTES : 0);
#pragma unroll
        for (int j = 0; j < NSLOT - 1; j++) {
            stage_copy<PARTS>(dst0 + j * REC_BYTES, src, lane, j < N);
            cp_commit();
            src += BACKWARD ? -REC_BYTES : REC_BYTES;
        }
        rslot = NSLOT - 1;      // (advance(0) steps it to slot 0 and refills slot NSLOT - 1)
    }
    // top of iteration i: wait for this iteration's records (all but the NSLOT-2 youngest groups complete), warp barrier (the
    // copies of the other lanes become visible, and every lane is past its reads of the slot consumed by iteration i-1), then
    // refill that slot with the records of stage i + NSLOT - 1.  Returns the slot to read.
    template <int PARTS, bool BACKWARD>
    __device__ __forceinline__ const double* advance(int i)
    {
        cp_wait<NSLOT - 2>();
        __syncwarp();
        // stage i + NSLOT - 1 goes into the slot read by iteration i - 1 (for i = 0: the last slot, still empty); no modulo, no branch
        const int wslot = rslot;
        rslot = rslot + 1 == NSLOT ? 0 : rslot + 1;
        stage_copy<PARTS>(dst0 + wslot * REC_BYTES, src, lane, i + NSLOT - 1 < N);
        cp_commit();
        if (BR2_PF_DIST > 0) {
            // experiment (off: BR2_PF_DIST = 0): the copy above runs NSLOT-1 stages ahead and 21 % of the stall samples are the wait
            // on it (profiles/r02c_pdas_ncu_summary.txt), so pull the record of stage i + BR2_PF_DIST into L2 first, one
            // 128-byte line per lane.  Measured SLOWER: 0.215 ms without, 0.222 / 0.225 / 0.231 ms at distance 5 / 8 / 12
            // (profiles/r02_pdas_occupancy.txt): the waits are not DRAM latency the L2 could hide
            constexpr int L0 = (PARTS & P_V) ? 0 : ((PARTS & P_G) ? 4 : 18), L1 = (PARTS & P_F) ? 22 : ((PARTS & P_G) ? 18 : 4);
            if (i + BR2_PF_DIST < N && lane >= L0 && lane < L1) {
                const char* pf = src + (BACKWARD ? -1 : 1) * (BR2_PF_DIST - (NSLOT - 1)) * REC_BYTES + 112 * lane;
                asm volatile("prefetch.global.L2 [%0];" ::"l"(pf));
            }
        }
        src += BACKWARD ? -REC_BYTES : REC_BYTES;
        return sm.st[rslot];
    }
};

// pull the instance's iterate (X, U rows) into L2 ahead of the epilogue that updates it
__device__ __forceinline__ void prefetch_iterate(const Inst& I)
{
    const char* xb = reinterpret_cast<const char*>(I.Xlin);
    const char* ub = reinterpret_cast<const char*>(I.Ulin);
    const int xbytes = (I.N + 1) * NX * 8, ubytes = I.N * NU * 8;
    for (int o = I.lane * 128; o < xbytes + 128; o += 32 * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(xb + (o < xbytes ? o : xbytes - 8)));
    for (int o = I.lane * 128; o < ubytes + 128; o += 32 * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(ub + (o < ubytes ? o : ubytes - 8)));
}

// E0: cold start of the IPM iterate (qp_solver_warm_start 0): du = 0 pushed strictly inside the box, slacks
// exactly consistent, lam = mu0 / t.
__device__ void ipm_init(Inst& I)
{
    const double thr = 1e-1, mu0 = 1.0;
#pragma unroll 5
    for (int idx = I.lane; idx < 4 * I.N; idx += 32) {
        const int k = idx >> 2, e = idx & 3;
        const double uk = I.Ulin[k * NU + e];
        const double lb = I.a.lbu[e] - uk, ub = I.a.ubu[e] - uk;
        double v = fmin(fmax(0.0, lb + thr), ub - thr);
        if (ub - lb < 2 * thr) v = 0.5 * (lb + ub);
        const double tl = v - lb, tu = ub - v;
        double* Vk = I.V + (size_t)k * SREC;
        Vk[V_V + e] = v; Vk[V_TL + e] = tl; Vk[V_TU + e] = tu;
        Vk[V_LL + e] = mu0 * rcp_newton(tl); Vk[V_LU + e] = mu0 * rcp_newton(tu);
    }
    __syncwarp();
}

// What the interior-point iteration accumulates while its sweeps run (round 1 made six extra passes over the V records per
// iteration for these -- "E1" / "E2", 23 % of the forced interior-point time, profiles/r2c_phase_profile.json).  Per bound
// (slack s, multiplier lam, slack step ds = +dv for the lower, -dv for the upper bound) the step lengths are kept as INVERSE
// step lengths (max instead of min, one division per instance instead of one per bound):
//   affine step     dlam = -lam (1 + ds/s):   1/alpha >= -ds/s, 1 + ds/s;   sum (lam + a dlam)(s + a ds) = (1 - a) sum lam s + a^2 S2
//   corrector step  dlam = (c - lam ds)/s - lam:  1/alpha_p >= -ds/s,  1/alpha_d >= -dlam/lam;
//                   sum (lam + ad dlam)(s + ap ds) = sum lam s + ap A1 + ad A2 + ap ad A3
//   stationarity residual after the step: r+ = (1 - ad) r + (ap - ad) dgu,  dgu = -gh - (ll/tl + lu/tu) dv  (Newton equation)
struct IpmAcc {
    double ia = 0.0, s2 = 0.0;                                  // affine sweep: inverse step length, S2
    double ia_p = 0.0, ia_d = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0, dgmax = 0.0;   // corrector sweep
};
// The update of the iterate is applied lazily: by the next factor sweep (which reads every V record anyway) or, after the last
// iteration, by the epilogue.  dlam of the step sits in V_CL / V_CU (the corrector forward sweep overwrites its rhs with it).
struct IpmUpd {
    bool pending = false;
    double ap = 0.0, ad = 0.0;
    double mu_sum = 0.0, res_max = 0.0;                         // out: complementarity sum and stationarity residual of the iterate
};

enum { PC_CHANGED = 1, PC_PINNED = 2, PC_ACTIVE = 4, PC_NAN = 8 };      // primal test of a candidate, see primal_check

// Forward sweep.
//   MODE 0: roll-out of the iterate            x+ = A x + B v + b,            x_0 = x0 - X_0   (reads V_V, writes V_X)
//   MODE 1: affine Newton step                 ddu = -kff - K ddx, ddx+ = A ddx + B ddu, ddx_0 = 0 (writes V_DV; accumulates acc.ia, s2)
//   MODE 2: closed-loop roll-out (fast path)   u = -kff - K x,     x+ = A x + B u + b,   x_0 = x0 - X_0 (writes V_X, V_V)
//   MODE 3: corrector Newton step              like MODE 1; writes V_DX, V_DV and dlam -> V_CL / V_CU; accumulates the rest of acc
// Z = [A|B] is read row-per-quad: lane (q,t) holds Z[q][4ki+t] and Z[8+q][4ki+t]; every product is 4 (3) local
// FMAs and a reduction over the 4 lanes of a quad.  Returns max |b| in MODE 0 / 2.
// MODE 2 with `ptest`: the roll-out also runs the primal test of the candidate on the fly -- the tests primal_check applies to an
// input that is not pinned (outside the box beyond the tolerance / within 1e-3 of a bound / not finite) -- and returns the PC_*
// flags in *ptest (PC_CHANGED = some input violates its box).  On the interior fast path, where nothing is pinned, this replaces
// the separate pass over the candidate unless an input has to be pinned.
template <int MODE>
__device__ double forward_sweep(Inst& I, IpmAcc* acc = nullptr, int* ptest = nullptr)
{
    const int q = I.q, t = I.t, N = I.N;
    const bool lo = q < 4;                      // quad owns a second state row 8 + q (else: no row 12..15)
    constexpr bool FEEDBACK = MODE != 0, AFFINE = (MODE == 0 || MODE == 2), NEWTON = (MODE == 1 || MODE == 3);
    constexpr int PARTS = NEWTON ? (P_V | P_G | P_F) : FEEDBACK ? (P_G | P_F) : (P_V | P_G);
    double ia = 0.0, ia_d = 0.0, c1 = 0.0, c2 = 0.0, c3 = 0.0, dgmax = 0.0;     // IpmAcc partials of this lane
    I.template begin<PARTS, false>();     // records in flight while the start values below are fetched
    double zr[3];                               // propagated vector, row layout: x[4ki + t]
    double bmax = 0.0;
    bool pviol = false, pact = false, pfin = true;      // MODE 2 primal test (lanes q < 4 test input q)
    const double lbq = I.a.lbu[q & 3], ubq = I.a.ubu[q & 3];
#pragma unroll
    for (int ki = 0; ki < 3; ki++)
        zr[ki] = AFFINE ? I.a.x0[(size_t)I.inst * NX + 4 * ki + t] - I.Xlin[4 * ki + t] : 0.0;
    constexpr int xoff = NEWTON ? V_DX : V_X;
    constexpr int uoff = NEWTON ? V_DV : V_V;
    int o0[4], o1[4];                           // shared-memory offsets of my rows of Z (constant over the sweep)
#pragma unroll
    for (int ki = 0; ki < 4; ki++) { o0[ki] = g_off(q, 4 * ki + t); o1[ki] = g_off(8 + (q & 3), 4 * ki + t); }
    for (int k = 0; k < N; k++) {
        const double* Ss = I.template advance<PARTS, false>(k);
        const double* Gs = Ss + S_G;
        const double* Fs = Ss + S_F;
        const double* Vs = Ss + S_V;
        double* Vk = I.V + (size_t)k * SREC;
        double z0[4], z1[4];
#pragma unroll
        for (int ki = 0; ki < 4; ki++) {
            z0[ki] = Gs[o0[ki]];
            z1[ki] = Gs[o1[ki]];                // (quads 4..7 read a valid row too: their p1 / uq are never consumed -- every
        }                                       // shuffle below reads from a lane of quads 0..3 -- and a select per load costs more)
        double ulin = 0.0;
        if (MODE == 2) ulin = I.Ulin[k * NU + (q & 3)];        // (for the primal test below; in flight during the gain product)
        if (q == 0 && MODE != 1) {              // (the state part of the affine step is never read)
#pragma unroll
            for (int ki = 0; ki < 3; ki++) Vk[xoff + 4 * ki + t] = zr[ki];
        }
        double ut;                              // u[t]
        if (!FEEDBACK) {
            ut = Vs[V_V + t];
        } else {
            // (K x)[q] for q < 4: K[q][4ki+t] -- MODE 1 (after an interior-point factorisation): Kt[4ki+t][q];  MODE 2 (after an
            // absolute-form factorisation): K in C-fragment order, kff = its column 12
            double part = 0.0;
#pragma unroll
            for (int ki = 0; ki < 3; ki++) part = fma(Fs[MODE == 2 ? f_kc(q & 3, 4 * ki + t) : (4 * ki + t) * 4 + (q & 3)], zr[ki], part);
            part += shfl_x(part, 1);
            part += shfl_x(part, 2);
            const double uq = -Fs[MODE == 2 ? f_kc(q & 3, 12) : F_KFF + (q & 3)] - part;
            if (lo && t == 0) Vk[uoff + q] = uq;
            ut = shfl(uq, 4 * t);
            if (MODE == 2) {
                const double lb = lbq - ulin, ub = ubq - ulin;
                const double tolu = 1e-12 * (ub - lb);
                pviol |= !(uq >= lb - tolu) || !(uq <= ub + tolu);     // (written so that a NaN fails)
                pact |= fmin(uq - lb, ub - uq) < 1e-3;
                pfin &= isfinite(uq);
            }
            if (NEWTON) {
                // Side computation on the lanes that own input q (off the recursion's critical path).  Lane t handles ONE bound of
                // input q: even t the lower (s = tl, lam = ll, ds = +du), odd t the upper (s = tu, lam = lu, ds = -du) -- the two are
                // mirror images, and TL/TU, LL/LU, CL/CU are adjacent fields.  Step-length tests use an approximate reciprocal
                // (rel. error < 1e-6, made conservative by the factor 1 + 2e-6 applied to the inverse step length afterwards).
                const int fo = (q & 3) + 4 * (t & 1);
                const double s_b = Vs[V_TL + fo], l_b = Vs[V_LL + fo];
                const double ds = (t & 1) ? -uq : uq;
                if (MODE == 1) {
                    const double r = ds * rcp_approx(s_b);
                    if (lo) ia = fmax(ia, fmax(-r, 1.0 + r));
                    if (lo && t < 2) c1 = fma(-l_b * (1.0 + r), ds, c1);                 // S2 += dlam ds
                } else {
                    const double is = rcp_newton(s_b), c_b = Vs[V_CL + fo];       // (branch-free: a division's special-case branch is a scheduling barrier)
                    const double dl = fma(fma(-l_b, ds, c_b), is, -l_b);                 // dlam = (c - lam ds) / s - lam
                    double piece = (t & 1) ? -(dl + l_b) : (dl + l_b);                     // dgu = -gu + (dll + ll) - (dlu + lu)
                    piece += shfl_x(piece, 1);
                    if (lo) {
                        ia = fmax(ia, -ds * is);
                        ia_d = fmax(ia_d, -dl * rcp_approx(l_b));
                        dgmax = fmax(dgmax, fabs(piece - Vs[V_GU + (q & 3)]));
                    }
                    if (lo && t < 2) {
                        c1 = fma(l_b, ds, c1); c2 = fma(dl, s_b, c2); c3 = fma(dl, ds, c3);
                        Vk[V_CL + fo] = dl;
                    }
                }
            }
        }
        // (the state part first: it does not wait for the gain product)
        double p0 = z0[0] * zr[0], p1 = z1[0] * zr[0];
#pragma unroll
        for (int ki = 1; ki < 3; ki++) { p0 = fma(z0[ki], zr[ki], p0); p1 = fma(z1[ki], zr[ki], p1); }
        p0 = fma(z0[3], ut, p0); p1 = fma(z1[3], ut, p1);
        p0 += shfl_x(p0, 1); p1 += shfl_x(p1, 1);
        p0 += shfl_x(p0, 2); p1 += shfl_x(p1, 2);
        if (AFFINE) {
            const double b0 = Gs[G_B_OFF + q], b1 = lo ? Gs[G_B_OFF + 8 + q] : 0.0;
            p0 += b0; p1 += b1;
            bmax = fmax(bmax, fmax(fabs(b0), fabs(b1)));
        }
        // quad layout -> row layout
        zr[0] = shfl(p0, 4 * t);
        zr[1] = shfl(p0, 4 * (4 + t));
        zr[2] = shfl(p1, 4 * t);
    }
    if (q == 0 && MODE != 1) {
#pragma unroll
        for (int ki = 0; ki < 3; ki++) I.V[(size_t)N * SREC + xoff + 4 * ki + t] = zr[ki];
    }
    if (MODE == 2) pfin &= isfinite(zr[0]) && isfinite(zr[1]) && isfinite(zr[2]);   // NaN / Inf anywhere in the data reaches x_N
    __syncwarp();
    constexpr double SAFE = 1.0 + 2e-6;         // covers the error of rcp_approx
    if (MODE == 1) { acc->ia = SAFE * warp_max(ia); acc->s2 = warp_sum(c1); }
    if (MODE == 3) {
        acc->ia_p = warp_max(ia); acc->ia_d = SAFE * warp_max(ia_d); acc->dgmax = warp_max(dgmax);
        acc->a1 = warp_sum(c1); acc->a2 = warp_sum(c2); acc->a3 = warp_sum(c3);
    }
    if (MODE == 2 && ptest) {
        *ptest = (__any_sync(FULL_MASK, lo && pviol) ? PC_CHANGED : 0) | (__any_sync(FULL_MASK, lo && pact) ? PC_ACTIVE : 0) |
                 (__any_sync(FULL_MASK, lo && !pfin) ? PC_NAN : 0);
    }
    return AFFINE ? warp_max(bmax) : 0.0;
}

// Backward factor sweep: Riccati factorisation of the stage LQR plus the vector recursion(s) that share it.
//   W' = Z' [P+ | v1 | v2]    (16 x 14, DMMA; the vectors ride in the otherwise padded columns 12, 13)
//   H  = W'[:, 0:12] Z        (16 x 16, DMMA)  = [A|B]' P+ [A|B]
//   Lam = H_uu + R~,  Lam^-1 entry by entry (one 3x3 cofactor per lane),  K = Lam^-1 [H_ux | g] (DMMA),  P = Q + H_xx - H_xu K (DMMA, k = 4)
// KIND = FS_IPM  (interior-point iteration, residual form around the iterate): R~ = R + lam_l/t_l + lam_u/t_u,
//                 v1 = pi+ (costate of the iterate -> reduced gradient gu), v2 = p+ (predictor rhs gh = gu).
// KIND = FS_ABS  (interior fast path, absolute form of the unconstrained LQR): R~ = R, v1 = s+ = P+ b_k + p+,
//                 g = rlin + B's+,  p = qlin + A's+ - K'g;  no roll-out of an iterate is needed beforehand.
// The contraction index of both products is enumerated in the order a C fragment holds it, so no fragment is re-laid-
// out between the products.  A C fragment gives lane (q,t) the columns 2t, 2t+1, 8+2t, 9+2t of rows q / 8+q.  The 12
// valid columns make three k-tiles:  kt0 = {2t},  kt1 = {2t+1},  kt2 = {8, 10, 9, 11}[t]  (lanes t = 2, 3 fetch their
// kt2 element, register 1 of lane t-2, with one shuffle).  Then
//   - the C registers of W' are the A fragments of the second product,
//   - the C registers of P+ (read through its symmetry, P[k][c] = P[c][k]) are the B fragments of the first product,
//   - and ONE gather of Z, z[kt][m] = Z[row(kt,t)][8m+q], is the A fragment of Z' (first product) and the B fragment
//     of Z (second product).
// Returns false unless every Lam is positive definite (determinant and principal minors).
enum { FS_IPM = 0, FS_ABS = 1, FS_AS = 2 };    // FS_AS: FS_ABS with the inputs of the guessed active set pinned at their bounds

template <int KIND>
__device__ bool factor_sweep(Inst& I, IpmUpd* upd = nullptr)
{
    const int q = I.q, t = I.t, N = I.N, lane = I.lane;
    const SolveArgs& a = I.a;
    constexpr bool ABSF = KIND != FS_IPM;       // absolute-form LQR sweep (fast paths)
    constexpr bool PIN = KIND == FS_AS;         // ... with pinned inputs
    const bool lo = q < 4;
    const int e = q & 3;                        // input index owned by quads 4..7
    const int qb = lane & ~3;                   // first lane of my quad
    const bool hi2 = t >= 2;
    const int src2 = hi2 ? lane - 2 : lane;     // where my kt2 element lives
    bool ok = true;
    double mu_sum = 0.0, res_max = 0.0;         // FS_IPM: accumulated on lanes (4 + e, 2)
    // contraction rows held for the three k-tiles, and their offsets inside a G record (column block m adds 32)
    const int row0 = 2 * t, row1 = 2 * t + 1, row2 = hi2 ? 2 * t + 5 : 8 + 2 * t;
    const int zo0 = g_off(row0, q), zo1 = g_off(row1, q), zo2 = g_off(row2, q);
    // cofactor (a = e, t) of the exchanged 4x4: rows {0..3} \ {e}, columns {0..3} \ {t}; pinned columns of lanes t >= 2
    const int ro0 = 4 * (0 + (0 >= e)), ro1 = 4 * (1 + (1 >= e)), ro2 = 4 * (2 + (2 >= e));
    const int co0 = 0 + (0 >= t), co1 = 1 + (1 >= t), co2 = 2 + (2 >= t);
    const int pc0 = 2 * (t & 1), pc1 = pc0 + 1;
    // P+ in C layout: h[m][n][j] = P[8m+q][8n+2t+j]; vin[kt] = (q == 4 ? v1 : q == 5 ? v2 : 0)[row(kt)]
    constexpr int PARTS = KIND == FS_IPM ? (P_V | P_G) : P_G;
    I.template begin<PARTS, true>();    // records in flight while the terminal values are fetched
    double h[2][2][2];
    double vin[3] = {0.0, 0.0, 0.0};
    double pq0 = 0.0, pq1 = 0.0;                // FS_ABS: p+ in quad layout (rows q, 8+q)
    {
        const double* VN = I.V + (size_t)N * SREC;
        const double* yN = yref_row(a, I.inst, N);
#pragma unroll
        for (int m = 0; m < 2; m++)
#pragma unroll
            for (int n = 0; n < 2; n++)
#pragma unroll
                for (int j = 0; j < 2; j++) {
                    const int r = 8 * m + q, c = 8 * n + 2 * t + j;
                    h[m][n][j] = (r == c && r < 12) ? a.We[r < 12 ? r : 0] : 0.0;
                }
        if (KIND == FS_IPM) {
            // pi_N = We (x_N + X_N - xref_N) for lanes q == 4; p_N = 0.  A pending update reaches x_N here: the three rows of the
            // four lanes of quad 4 are the twelve states, each exactly once
            if (q == 4) {
                double xn0 = VN[V_X + row0], xn1 = VN[V_X + row1], xn2 = VN[V_X + row2];
                if (upd->pending) {
                    double* VNw = I.V + (size_t)N * SREC;
                    xn0 = fma(upd->ap, VN[V_DX + row0], xn0); xn1 = fma(upd->ap, VN[V_DX + row1], xn1); xn2 = fma(upd->ap, VN[V_DX + row2], xn2);
                    VNw[V_X + row0] = xn0; VNw[V_X + row1] = xn1; VNw[V_X + row2] = xn2;
                }
                vin[0] = a.We[row0] * (xn0 + I.Xlin[N * NX + row0] - yN[row0]);
                vin[1] = a.We[row1] * (xn1 + I.Xlin[N * NX + row1] - yN[row1]);
                vin[2] = a.We[row2] * (xn2 + I.Xlin[N * NX + row2] - yN[row2]);
            }
        } else {
            // p_N = We (X_N - xref_N)
            pq0 = a.We[q] * (I.Xlin[N * NX + q] - yN[q]);
            pq1 = lo ? a.We[8 + e] * (I.Xlin[N * NX + 8 + e] - yN[8 + e]) : 0.0;
            if (q == 4) {                       // p_N in the rows the k-tiles of lanes (4, t) hold
                vin[0] = a.We[row0] * (I.Xlin[N * NX + row0] - yN[row0]);
                vin[1] = a.We[row1] * (I.Xlin[N * NX + row1] - yN[row1]);
                vin[2] = a.We[row2] * (I.Xlin[N * NX + row2] - yN[row2]);
            }
        }
    }
    SPROF_START();
    for (int k = N - 1, it = 0; k >= 0; k--, it++) {
        const double* Ss = I.template advance<PARTS, true>(it);
        SPROF(0);
        const double* Gs = Ss + S_G;
        const double* Vs = Ss + S_V;
        double* Fk = I.S + (size_t)k * SREC + S_F;
        double* Vk = I.V + (size_t)k * SREC;
        // FS_AS: the stage's guessed active set (2 bits per input: 1 = at the lower, 2 = at the upper bound; uniform over the
        // warp) and the pinned values du_c = bound_c - U_k[c] of the inputs this lane deals with: its two columns c0, c1 of
        // H[:, 12..15] (lanes t >= 2) and its own input e (quads 4..7)
        int code = 0;
        double ub0 = 0.0, ub1 = 0.0, ube = 0.0;
        if (PIN) {
            code = a.aset[(size_t)I.inst * N + k];
            if (code != 0) {
                const double* Uk = I.Ulin + k * NU;
                ub0 = (((code >> (2 * pc0)) & 3) == 2 ? a.ubu[pc0] : a.lbu[pc0]) - Uk[pc0];
                ub1 = (((code >> (2 * pc1)) & 3) == 2 ? a.ubu[pc1] : a.lbu[pc1]) - Uk[pc1];
                ube = (((code >> (2 * e)) & 3) == 2 ? a.ubu[e] : a.lbu[e]) - Uk[e];
            }
        }
        double z[3][2];
        z[0][0] = Gs[zo0]; z[0][1] = Gs[zo0 + 32];
        z[1][0] = Gs[zo1]; z[1][1] = Gs[zo1 + 32];
        z[2][0] = Gs[zo2]; z[2][1] = Gs[zo2 + 32];
        const double tsk = Gs[G_TS];
        // state rows q and 8+q
        // (stage weights from the warp's copy in shared memory: a.W sits in the constant bank, and a lane-dependent index there is replayed
        // once per distinct address -- three such loads per stage at the top of the dependent chain)
        const double* Wsm = I.sm.xch + 64;
        const double qd0 = tsk * Wsm[q];
        const double qd1 = lo ? tsk * Wsm[8 + e] : 0.0;
        double qx0 = Gs[G_QLIN + q], qx1 = lo ? Gs[G_QLIN + 8 + e] : 0.0;
        // input row e (meaningful in quads 4..7)
        const double rd = tsk * Wsm[12 + e];
        double rt = rd;
        double gu_loc = Gs[G_RLIN + e];
        double ll_e = 0.0, lu_e = 0.0, cmp_e = 0.0;     // FS_IPM: multipliers and complementarity products of input e
        if (KIND == FS_IPM) {
            // the iterate of this stage, with the previous iteration's step applied on the way (and written back)
            double xq0 = Vs[V_X + q], xq1 = lo ? Vs[V_X + 8 + e] : 0.0;
            double v = Vs[V_V + e], tl = Vs[V_TL + e], tu = Vs[V_TU + e];
            ll_e = Vs[V_LL + e]; lu_e = Vs[V_LU + e];
            if (upd->pending) {
                const double ap = upd->ap, ad = upd->ad, dv = Vs[V_DV + e];
                xq0 = fma(ap, Vs[V_DX + q], xq0);
                if (lo) xq1 = fma(ap, Vs[V_DX + 8 + e], xq1);
                v = fma(ap, dv, v); tl = fma(ap, dv, tl); tu = fma(-ap, dv, tu);
                ll_e = fma(ad, Vs[V_CL + e], ll_e); lu_e = fma(ad, Vs[V_CU + e], lu_e);
                if (t == 0) Vk[V_X + q] = xq0;
                if (lo && t == 1) Vk[V_X + 8 + q] = xq1;
                if (!lo && t == 2) { Vk[V_V + e] = v; Vk[V_TL + e] = tl; Vk[V_TU + e] = tu; Vk[V_LL + e] = ll_e; Vk[V_LU + e] = lu_e; }
            }
            // residual form around the iterate (x, v):  Q (x + X - xref) = Q x + qlin,  R (v + U - uref) = R v + rlin
            qx0 = fma(qd0, xq0, qx0);
            if (lo) qx1 = fma(qd1, xq1, qx1);
            rt += ll_e * rcp_newton(tl) + lu_e * rcp_newton(tu);
            gu_loc = fma(rd, v, gu_loc);
            cmp_e = ll_e * tl + lu_e * tu;
        }
        // kt2 elements of P+ that live in lane t-2
        const double hx0 = shfl(h[0][1][1], src2), hx1 = shfl(h[1][1][1], src2);
        const double b02 = hi2 ? hx0 : h[0][1][0];
        const double b12 = hi2 ? hx1 : h[1][1][0];
        // FS_ABS: vin = p+ (set at the end of the previous stage); the P+ b_k part of s+ = P+ b_k + p+ is formed AFTER the first
        // product as (Z'P+) b_k, off the path that leads into the DMMAs
        SPROF(1);
        // ---- W' = Z' [P+ | v1 | v2] ----
        double w[2][2][2] = {{{0.0, 0.0}, {0.0, 0.0}}, {{0.0, 0.0}, {0.0, 0.0}}};
        {
            const double b00 = h[0][0][0], b01 = h[0][0][1];
            const double b10 = lo ? h[1][0][0] : vin[0], b11 = lo ? h[1][0][1] : vin[1], b1x = lo ? b12 : vin[2];
#pragma unroll
            for (int m = 0; m < 2; m++) {
                dmma(w[m][0], z[0][m], b00); dmma(w[m][1], z[0][m], b10);
                dmma(w[m][0], z[1][m], b01); dmma(w[m][1], z[1][m], b11);
                dmma(w[m][0], z[2][m], b02); dmma(w[m][1], z[2][m], b1x);
            }
        }
        SPROF(2);
        // ---- H = W' Z ----
        const double wx0 = shfl(w[0][1][1], src2), wx1 = shfl(w[1][1][1], src2);
        const double a02 = hi2 ? wx0 : w[0][1][0];
        const double a12 = hi2 ? wx1 : w[1][1][0];
#pragma unroll
        for (int m = 0; m < 2; m++)
#pragma unroll
            for (int n = 0; n < 2; n++) { h[m][n][0] = 0.0; h[m][n][1] = 0.0; }
#pragma unroll
        for (int n = 0; n < 2; n++) {
            dmma(h[0][n], w[0][0][0], z[0][n]); dmma(h[1][n], w[1][0][0], z[0][n]);
            dmma(h[0][n], w[0][0][1], z[1][n]); dmma(h[1][n], w[1][0][1], z[1][n]);
            dmma(h[0][n], a02, z[2][n]);        dmma(h[1][n], a12, z[2][n]);
        }
        SPROF(3);
        // ---- the vector products sit in columns 12, 13 of W': lane (q,2) holds ([A|B]'v1)[8m+q], ([A|B]'v2)[8m+q] ----
        double* xch = I.sm.xch;
        double yg0, yg1;                         // (K'g)[q], (K'g)[8+q]
        double at0, at1, bt0 = 0.0, bt1 = 0.0;
        if (ABSF) {
            // ================= absolute-form LQR stage (fast paths): explicit inverse of Lam, distributed =================
            // Round 1 solved the 4x4 system redundantly on every lane (adjugate + three 4x4 applies: 133 of the stage's 154 fp64
            // instructions, 266 of its 756 fp64-pipe cycles -- and the factor sweep is fp64-pipe-bound: profiles/r2c_phase_profile.json).
            // Now lane (q,t) forms ONE entry of Lam^-1 (a 3x3 cofactor of the exchanged matrix), the determinant is a quad
            // reduction, and K = Lam^-1 [H_ux | g] is two DMMAs whose operands are single shared-memory loads:
            //   A fragment (row a = q < 4, col t)      = Lam^-1[a][t]
            //   B fragment (row t, col 8n+q)           = H_xu[8n+q][t] = xch[32 n + lane]   (column 12: g[t])
            // The same xch values are the A fragment of H_xu in the Riccati update, whose column 12 (K's column 12 = kff) returns
            // K'g for free.
            // ---- exchange: lanes t >= 2 hold H[8m+q][12 + c], c = 2(t-2) + j; rows 12..15 are Lam (R~ on its diagonal) ----
            double e0 = h[0][1][0], e1 = h[0][1][1], f0 = h[1][1][0], f1 = h[1][1][1];
            if (!lo && t == 2 + (e >> 1)) {
                if (e & 1) f1 += rt; else f0 += rt;
            }
            double pa0 = 0.0, pa1 = 0.0;
            bool pe = false;
            if (PIN && code != 0) {
                // inputs F pinned at ubar_F: their columns act on the states and on the free inputs through H[:, 12+F] ubar_F
                // (b_k <- b_k + B_F ubar_F), then row / column F leave the system (identity row, g_F = -ubar_F => u_F = ubar_F)
                const bool p0 = (code >> (2 * pc0)) & 3, p1 = (code >> (2 * pc1)) & 3;
                pe = (code >> (2 * e)) & 3;
                pa0 = (p0 ? e0 * ub0 : 0.0) + (p1 ? e1 * ub1 : 0.0);
                pa1 = (p0 ? f0 * ub0 : 0.0) + (p1 ? f1 * ub1 : 0.0);
                pa0 += shfl_x(pa0, 1); pa1 += shfl_x(pa1, 1);        // lanes 2 <-> 3 of the quad
                if (p0) { e0 = 0.0; f0 = lo ? 0.0 : ((e == pc0) ? 1.0 : 0.0); }
                if (p1) { e1 = 0.0; f1 = lo ? 0.0 : ((e == pc1) ? 1.0 : 0.0); }
                if (!lo && pe) { f0 = (e == pc0) ? 1.0 : 0.0; f1 = (e == pc1) ? 1.0 : 0.0; }
            }
            if (t >= 2) {
                *reinterpret_cast<double2*>(xch + q * 4 + 2 * (t - 2)) = make_double2(e0, e1);
                *reinterpret_cast<double2*>(xch + (8 + q) * 4 + 2 * (t - 2)) = make_double2(f0, f1);
            }
            __syncwarp();
            SPROF(4);
            // ---- Z's+ = Z'p+ + (Z'P+) b_k, behind the barrier: only g (two segments further down) and p need it.  Lane (q,t) holds
            // columns 2t, 2t+1, 8+2t, 9+2t of rows q / 8+q of Z'P+; the p+ part is column 12 of W', a register of lane (q,2) -- the only
            // lane that needs at0 / at1 from here on ----
            const double bb0 = Gs[G_B_OFF + row0], bb1 = Gs[G_B_OFF + row1];
            const double bb2 = hi2 ? 0.0 : Gs[G_B_OFF + 8 + 2 * t], bb3 = hi2 ? 0.0 : Gs[G_B_OFF + 9 + 2 * t];
            double s0 = w[0][0][0] * bb0 + w[0][0][1] * bb1 + w[0][1][0] * bb2 + w[0][1][1] * bb3;
            double s1 = w[1][0][0] * bb0 + w[1][0][1] * bb1 + w[1][1][0] * bb2 + w[1][1][1] * bb3;
            s0 += shfl_x(s0, 1); s1 += shfl_x(s1, 1);
            s0 += shfl_x(s0, 2); s1 += shfl_x(s1, 2);
            at0 = w[0][1][0] + s0; at1 = w[1][1][0] + s1;            // meaningful on lanes t == 2
            double gval = gu_loc + at1;                               // lane (4+e, 2): g_e = rlin_e + (B's+)_e
            if (PIN && code != 0) {
                at0 += pa0;
                if (lo) at1 += pa1; else gval = pe ? -ube : gval + pa1;
            }
            // ---- my entry of Lam^-1: cofactor (a, t) of the 4x4 at xch[48..63], a = q & 3 ----
            const double* LM = xch + 48;
            double cof;
            {
                const double m00 = LM[ro0 + co0], m01 = LM[ro0 + co1], m02 = LM[ro0 + co2];
                const double m10_ = LM[ro1 + co0], m11 = LM[ro1 + co1], m12 = LM[ro1 + co2];
                const double m20 = LM[ro2 + co0], m21 = LM[ro2 + co1], m22 = LM[ro2 + co2];
                const double d0 = m11 * m22 - m12 * m21, d1 = m10_ * m22 - m12 * m20, d2 = m10_ * m21 - m11 * m20;
                cof = m00 * d0 - m01 * d1 + m02 * d2;
                if ((e + t) & 1) cof = -cof;
            }
            double det = LM[4 * e + t] * cof;                         // expansion along row a
            det += shfl_x(det, 1);
            det += shfl_x(det, 2);
            ok &= (det > 0.0) && ((e != t) || (cof > 0.0));           // positive definite: det and the principal 3x3 minors
            const double linv = lo ? cof * rcp_newton(det) : 0.0;
            const double hx0 = xch[lane];
            // g reaches its place in the B fragment (lane (4,t): g[t]) by one shuffle from lane (4+t, 2): the products behind it
            // (Z's+ with its two reduction levels) stay off the path to the exchange barrier
            const double gsh = shfl(gval, 4 * (4 + t) + 2);
            const double hx1 = lo ? xch[32 + lane] : (q == 4 ? gsh : 0.0);
            // (xch is rewritten by the next stage only after the warp barrier at the top of its iteration)
            // ---- K = Lam^-1 [H_ux | g]: C layout, lane (a,t) holds K[a][8n+2t+j]; K[a][12] = kff[a] ----
            SPROF(5);
            double kc[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
            dmma(kc[0], linv, hx0);
            dmma(kc[1], linv, hx1);
            // F record = K in C-fragment order (layout.h): four coalesced 128-byte rows from lanes 0..15
            if (lane < 16) {
                Fk[lane] = kc[0][0]; Fk[16 + lane] = kc[0][1]; Fk[32 + lane] = kc[1][0]; Fk[48 + lane] = kc[1][1];
            }
            // ---- B fragment of K for the update: K[t][8n+q] sits in register q&1 of lane (t, q>>1) ----
            const int sl = 4 * t + (q >> 1);
            const double k00 = shfl(kc[0][0], sl), k01 = shfl(kc[0][1], sl), k10 = shfl(kc[1][0], sl), k11 = shfl(kc[1][1], sl);
            const double kb0 = (q & 1) ? k01 : k00, kb1 = (q & 1) ? k11 : k10;
            // ---- P = Q + H_xx - H_xu K; column 12 of the right tiles (lane t == 2, register 0) collects -(H_xu kff) = -K'g ----
            SPROF(6);
            if (t == 2) { h[0][1][0] = 0.0; h[1][1][0] = 0.0; }
            dmma(h[0][0], -hx0, kb0); dmma(h[0][1], -hx0, kb1);
            dmma(h[1][0], -hx1, kb0); dmma(h[1][1], -hx1, kb1);
            yg0 = -h[0][1][0]; yg1 = -h[1][1][0];                     // meaningful on lanes t == 2
        } else {
            // ================= interior-point stage: the same distributed inverse as the fast paths =================
            // (Round 1 / early round 2 factorised Lam redundantly on every lane -- Cholesky, five triangular solves, Y'Y update: a long
            // dependent chain of sqrt / divide per stage.)  Lane (a, t) forms one entry of Lam^-1 as a cofactor, K = Lam^-1 [H_ux | g]
            // and the update P = Q + H_xx - H_xu K are DMMAs, K'g falls out of the update's column 12.  The F record keeps the layout
            // the vector sweeps of the iteration read: K' at (4 ki + t) * 4 + a, kff, and -- for the corrector's second right-hand
            // side -- the upper triangle of Lam^-1 where the Cholesky factor used to be.
            at0 = shfl(w[0][1][0], qb | 2); at1 = shfl(w[1][1][0], qb | 2);
            bt0 = shfl(w[0][1][1], qb | 2); bt1 = shfl(w[1][1][1], qb | 2);
            const double gu = gu_loc + at1;          // quads 4..7: R du + r + B'pi+
            const double gval = gu + bt1;            // predictor rhs: gh = gu
            double e0 = h[0][1][0], e1 = h[0][1][1], f0 = h[1][1][0], f1 = h[1][1][1];
            if (!lo && t == 2 + (e >> 1)) {
                if (e & 1) f1 += rt; else f0 += rt;
            }
            if (t >= 2) {
                *reinterpret_cast<double2*>(xch + q * 4 + 2 * (t - 2)) = make_double2(e0, e1);
                *reinterpret_cast<double2*>(xch + (8 + q) * 4 + 2 * (t - 2)) = make_double2(f0, f1);
            }
            __syncwarp();
            if (!lo && t == 2) {
                Vk[V_GU + e] = gu;
                mu_sum += cmp_e;
                res_max = fmax(res_max, fabs(gu - ll_e + lu_e));
            }
            const double* LM = xch + 48;
            double cof;
            {
                const double m00 = LM[ro0 + co0], m01 = LM[ro0 + co1], m02 = LM[ro0 + co2];
                const double m10_ = LM[ro1 + co0], m11 = LM[ro1 + co1], m12 = LM[ro1 + co2];
                const double m20 = LM[ro2 + co0], m21 = LM[ro2 + co1], m22 = LM[ro2 + co2];
                const double d0 = m11 * m22 - m12 * m21, d1 = m10_ * m22 - m12 * m20, d2 = m10_ * m21 - m11 * m20;
                cof = m00 * d0 - m01 * d1 + m02 * d2;
                if ((e + t) & 1) cof = -cof;
            }
            double det = LM[4 * e + t] * cof;                         // expansion along row a
            det += shfl_x(det, 1);
            det += shfl_x(det, 2);
            ok &= (det > 0.0) && ((e != t) || (cof > 0.0));           // positive definite: det and the principal 3x3 minors
            const double linv = lo ? cof * rcp_newton(det) : 0.0;
            const double hx0 = xch[lane];
            const double gsh = shfl(gval, 4 * (4 + t) + 2);           // g[t] from lane (4+t, 2)
            const double hx1 = lo ? xch[32 + lane] : (q == 4 ? gsh : 0.0);
            double kc[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
            dmma(kc[0], linv, hx0);
            dmma(kc[1], linv, hx1);
            // F record (interior-point layout): K'[c][a] at c * 4 + a for c < 12, kff[a] = K[a][12], upper triangle of Lam^-1
            if (lo) {
                const int a4 = q;                                     // row a of K held by lanes (a, t)
                Fk[(2 * t) * 4 + a4] = kc[0][0]; Fk[(2 * t + 1) * 4 + a4] = kc[0][1];
                if (t < 2) { Fk[(8 + 2 * t) * 4 + a4] = kc[1][0]; Fk[(9 + 2 * t) * 4 + a4] = kc[1][1]; }
                if (t == 2) Fk[F_KFF + a4] = kc[1][0];
                if (t >= a4) Fk[F_L_OFF + linv_slot(a4, t)] = linv;
            }
            const int sl = 4 * t + (q >> 1);
            const double k00 = shfl(kc[0][0], sl), k01 = shfl(kc[0][1], sl), k10 = shfl(kc[1][0], sl), k11 = shfl(kc[1][1], sl);
            const double kb0 = (q & 1) ? k01 : k00, kb1 = (q & 1) ? k11 : k10;
            // P = Q + H_xx - H_xu K; column 12 of the right tiles (lane t == 2, register 0) collects -(H_xu kff) = -K'g
            if (t == 2) { h[0][1][0] = 0.0; h[1][1][0] = 0.0; }
            dmma(h[0][0], -hx0, kb0); dmma(h[0][1], -hx0, kb1);
            dmma(h[1][0], -hx1, kb0); dmma(h[1][1], -hx1, kb1);
            yg0 = shfl(-h[0][1][0], qb | 2); yg1 = shfl(-h[1][1][0], qb | 2);      // K'g lives on lanes t == 2: to the whole quad
        }
        if (t == (q >> 1)) {                     // diagonal element (8m+q, 8m+q) is C register q&1 of lane (q, q>>1)
            if (q & 1) { h[0][0][1] += qd0; h[1][1][1] += qd1; } else { h[0][0][0] += qd0; h[1][1][0] += qd1; }
        }
        if (KIND == FS_IPM) {
            // ---- pi, p for the next stage's vector columns: lane (4,t) needs pi[row(kt)], lane (5,t) p[row(kt)] ----
            const double pi0 = qx0 + at0, pi1 = qx1 + at1;
            const double pv0 = bt0 - yg0, pv1 = bt1 - yg1;
            const double vs0 = (t == 0) ? pi0 : pv0, vs1 = (t == 0) ? pi1 : pv1;    // lanes t = 0 serve pi, t = 1 serve p
            const int sel = q & 1;
            const double i0 = shfl(vs0, 8 * t + sel), i1 = shfl(vs0, 8 * t + 4 + sel);
            const double i2 = shfl(vs1, (hi2 ? 4 * (2 * t - 3) : 8 * t) + sel);
            const bool vq = (q == 4) || (q == 5);
            vin[0] = vq ? i0 : 0.0; vin[1] = vq ? i1 : 0.0; vin[2] = vq ? i2 : 0.0;
        } else {
            pq0 = qx0 + at0 - yg0;               // p = qlin + A's+ - K'g
            pq1 = qx1 + at1 - yg1;
            // (p lives on lane t == 2 of each quad)
            const double i0 = shfl(pq0, 8 * t + 2), i1 = shfl(pq0, 8 * t + 6), i2 = shfl(pq1, (hi2 ? 4 * (2 * t - 3) : 8 * t) + 2);
            vin[0] = (q == 4) ? i0 : 0.0; vin[1] = (q == 4) ? i1 : 0.0; vin[2] = (q == 4) ? i2 : 0.0;
        }
        SPROF(7);
    }
    __syncwarp();
    if (KIND == FS_IPM) { upd->mu_sum = warp_sum(mu_sum); upd->res_max = warp_max(res_max); upd->pending = false; }
    return __all_sync(FULL_MASK, ok);
}

// Backward vector sweep (corrector): rhs gh = gu - cl/tl + cu/tu with the second-order terms cl = sigma mu - dv_aff dll_aff,
// cu = sigma mu + dv_aff dlu_aff formed here from the affine step in V_DV (and stored for the corrector forward sweep);
// g = gh + B'p+,  kff = Lam^-1 g,  p = A'p+ - K'g.
// Z is read in fragment order; lane (q,t) forms its part of column q and column 8+q of Z'p and the quad reduces over t.
__device__ void backward_vec_sweep(Inst& I, double sigmu)
{
    const int q = I.q, t = I.t, N = I.N, lane = I.lane;
    const bool lo = q < 4;
    const int e = q & 3;
    double pr[3] = {0.0, 0.0, 0.0};             // p+ in row layout
    I.template begin<P_V | P_G | P_F, true>();
    for (int k = N - 1, it = 0; k >= 0; k--, it++) {
        const double* Ss = I.template advance<P_V | P_G | P_F, true>(it);
        const double* Gk = Ss + S_G;
        const double* Fk = Ss + S_F;
        const double* Vs = Ss + S_V;
        double* Vk = I.V + (size_t)k * SREC;
        double o0 = 0.0, o1 = 0.0;
#pragma unroll
        for (int ki = 0; ki < 3; ki++) {
            o0 = fma(Gk[((ki * 2 + 0) << 5) + lane], pr[ki], o0);
            o1 = fma(Gk[((ki * 2 + 1) << 5) + lane], pr[ki], o1);
        }
        const double itl = rcp_newton(Vs[V_TL + e]), itu = rcp_newton(Vs[V_TU + e]);
        const double ll = Vs[V_LL + e], lu = Vs[V_LU + e], dva = Vs[V_DV + e];
        const double cl = fma(dva, fma(ll * dva, itl, ll), sigmu);       // sigma mu - dv dll,  dll = -ll - ll dv / tl
        const double cu = fma(dva, fma(lu * dva, itu, -lu), sigmu);      // sigma mu + dv dlu,  dlu = -lu + lu dv / tu
        const double gh = Vs[V_GU + e] - cl * itl + cu * itu;
        if (!lo && t == 2) { Vk[V_CL + e] = cl; Vk[V_CU + e] = cu; }
        const double2 ka = *reinterpret_cast<const double2*>(Fk + q * 4), kb = *reinterpret_cast<const double2*>(Fk + q * 4 + 2);
        double2 kc = make_double2(0.0, 0.0), kd = make_double2(0.0, 0.0);
        if (lo) { kc = *reinterpret_cast<const double2*>(Fk + (8 + q) * 4); kd = *reinterpret_cast<const double2*>(Fk + (8 + q) * 4 + 2); }
        o0 += shfl_x(o0, 1); o1 += shfl_x(o1, 1);
        o0 += shfl_x(o0, 2); o1 += shfl_x(o1, 2);          // (Z'p+)[q], (Z'p+)[8+q]
        const double gval = gh + o1;                          // quads 4..7
        double gt[4];
#pragma unroll
        for (int c = 0; c < 4; c++) gt[c] = shfl(gval, 4 * (4 + c));
        const double pv0 = o0 - (ka.x * gt[0] + ka.y * gt[1] + kb.x * gt[2] + kb.y * gt[3]);
        const double pv1 = o1 - (kc.x * gt[0] + kc.y * gt[1] + kd.x * gt[2] + kd.y * gt[3]);
        pr[0] = shfl(pv0, 4 * t);
        pr[1] = shfl(pv0, 4 * (4 + t));
        pr[2] = shfl(pv1, 4 * t);
        // feed-forward for the forward sweep, kff = Lam^-1 g (off the recursion's critical path): lanes 0..3 one entry each, from the
        // upper triangle of Lam^-1 the factor sweep left in the F record
        if (lane < 4) {
            double kf = 0.0;
#pragma unroll
            for (int c = 0; c < 4; c++) kf = fma(Fk[F_L_OFF + linv_slot(lane < c ? lane : c, lane < c ? c : lane)], gt[c], kf);
            I.S[(size_t)k * SREC + S_F + F_KFF + lane] = kf;
        }
    }
    __syncwarp();
}

// Costate sweep over a candidate solution (dx, du in V_X / V_V) of the active-set fast path:
//   lam_N = We (dx_N + X_N - xref_N),  lam_k = qlin_k + Ts W_x dx_k + A_k' lam_{k+1},
//   mu_k  = Ts W_u du_k + rlin_k + B_k' lam_{k+1}          (gradient of the Lagrangian with respect to the inputs)
// The candidate solves the LQR with the guessed active set pinned, so mu = 0 on the free inputs; it is the minimiser of the
// box-constrained QP iff every free input lies inside its box and mu >= 0 (<= 0) on the inputs pinned at the lower (upper)
// bound -- the KKT conditions of a strictly convex QP.  Where the test fails the guess is repaired in place (a free input
// outside its box is pinned at the violated bound, a pinned input with the wrong multiplier sign is released): one step
// of a primal-dual active-set method.  Returns true iff the candidate passed.
__device__ bool costate_check(Inst& I)
{
    const int q = I.q, t = I.t, N = I.N, lane = I.lane;
    const SolveArgs& a = I.a;
    const bool lo = q < 4;
    const int e = q & 3;
    I.template begin<P_V | P_G, true>();
    double pr[3];                               // lam+ in row layout
    {
        const double* VN = I.V + (size_t)N * SREC;
        const double* yN = yref_row(a, I.inst, N);
#pragma unroll
        for (int ki = 0; ki < 3; ki++) {
            const int r = 4 * ki + t;
            pr[ki] = a.We[r] * (VN[V_X + r] + I.Xlin[N * NX + r] - yN[r]);
        }
    }
    bool good = true;
    for (int k = N - 1, it = 0; k >= 0; k--, it++) {
        const double* Ss = I.template advance<P_V | P_G, true>(it);
        const double* Gk = Ss + S_G;
        const double* Vs = Ss + S_V;
        double o0 = 0.0, o1 = 0.0;
#pragma unroll
        for (int ki = 0; ki < 3; ki++) {
            o0 = fma(Gk[((ki * 2 + 0) << 5) + lane], pr[ki], o0);
            o1 = fma(Gk[((ki * 2 + 1) << 5) + lane], pr[ki], o1);
        }
        const double tsk = Gk[G_TS];
        const int code = a.aset[(size_t)I.inst * N + k];
        const double uk = I.Ulin[k * NU + e];
        o0 += shfl_x(o0, 1); o1 += shfl_x(o1, 1);
        o0 += shfl_x(o0, 2); o1 += shfl_x(o1, 2);          // (Z'lam+)[q], (Z'lam+)[8+q]
        // input e of this stage lives in quad 4 + e
        const double du = Vs[V_V + e];
        const double mu = fma(tsk * a.W[12 + e], du, Gk[G_RLIN + e]) + o1;
        const double lb = a.lbu[e] - uk, ub = a.ubu[e] - uk;
        // tolerances in units of the answer: a free input may leave its box by 1e-12 of the box width (the epilogue clamps), and a
        // pinned input is released when doing so would move it by more than 1e-9 of the width (|mu| / Lam_ee, Lam_ee >= Ts R_ee)
        const double tolu = 1e-12 * (ub - lb), tolm = 1e-9 * (ub - lb) * tsk * a.W[12 + e];
        const int cc = (code >> (2 * e)) & 3;
        int nc = cc;
        if (cc == 0) {
            if (!(du >= lb - tolu)) nc = 1;                 // written so that a NaN fails
            else if (!(du <= ub + tolu)) nc = 2;
        } else if (cc == 1) {
            if (!(mu >= -tolm)) nc = 0;
        } else {
            if (!(mu <= tolm)) nc = 0;
        }
        if (!lo && (nc != cc || !isfinite(du) || !isfinite(mu))) good = false;
        const int c0 = __shfl_sync(FULL_MASK, nc, 16), c1 = __shfl_sync(FULL_MASK, nc, 20), c2 = __shfl_sync(FULL_MASK, nc, 24),
                  c3 = __shfl_sync(FULL_MASK, nc, 28);
        const int ncode = c0 | (c1 << 2) | (c2 << 4) | (c3 << 6);
        if (lane == 0 && ncode != code) a.aset[(size_t)I.inst * N + k] = ncode;
        // lam_k, then quad layout -> row layout
        const double pv0 = fma(tsk * a.W[q], Vs[V_X + q], Gk[G_QLIN + q]) + o0;
        const double pv1 = lo ? fma(tsk * a.W[8 + e], Vs[V_X + 8 + e], Gk[G_QLIN + 8 + e]) + o1 : 0.0;
        pr[0] = shfl(pv0, 4 * t);
        pr[1] = shfl(pv0, 4 * (4 + t));
        pr[2] = shfl(pv1, 4 * t);
    }
    __syncwarp();
    return __all_sync(FULL_MASK, good);
}

// Primal half of the active-set test on the candidate (dx, du) that forward_sweep<2> left in V_X / V_V: a free input outside its
// box is pinned at the violated bound (the other half, the multiplier signs of the pinned inputs, is costate_check).  `fresh`:
// the stage codes in a.aset are stale (first attempt of an instance without a guess): they count as 0 and are rewritten.
// Returns bit 0: a code changed, bit 1: some input is pinned, bit 2: some input ends within 1e-3 of a bound (hint for the
// next solve), bit 3: NaN / Inf in the candidate.
__device__ int primal_check(Inst& I, bool fresh)
{
    const SolveArgs& a = I.a;
    const int nb = 4 * I.N, lane = I.lane, e = lane & 3;
    int* as = a.aset + (size_t)I.inst * I.N;
    int flags = 0;
#pragma unroll 2
    for (int base = 0; base < nb; base += 32) {
        const int idx = base + lane;
        const bool valid = idx < nb;
        const int k = (valid ? idx : 0) >> 2;
        const int code = (valid && !fresh) ? as[k] : 0;
        int cc = (code >> (2 * e)) & 3;
        if (valid) {
            const double uk = I.Ulin[idx], du = I.V[(size_t)k * SREC + V_V + e];
            const double lb = a.lbu[e] - uk, ub = a.ubu[e] - uk;
            const double tolu = 1e-12 * (ub - lb);
            if (cc == 0) {
                if (!(du >= lb - tolu)) cc = 1;             // written so that a NaN fails
                else if (!(du <= ub + tolu)) cc = 2;
            }
            if (!isfinite(du)) flags |= PC_NAN;
            if (fmin(du - lb, ub - du) < 1e-3) flags |= PC_ACTIVE;
        }
        int word = cc << (2 * e);
        word |= __shfl_xor_sync(FULL_MASK, word, 1);
        word |= __shfl_xor_sync(FULL_MASK, word, 2);
        if (valid) {
            if (word != code) flags |= PC_CHANGED;
            if (word) flags |= PC_PINNED;
            if (e == 0 && (fresh || word != code)) as[k] = word;
        }
    }
    flags = __reduce_or_sync(FULL_MASK, flags);
    __syncwarp();
    return flags;
}

__device__ __forceinline__ double step_to_boundary(double v, double dv)
{
    return dv < 0.0 ? -v / dv : 2.0;   // 2 = "not blocking" (callers clamp at 1)
}

// Instrumentation build (-DBR2_PROFILE, scripts/phase_profile.py): cycles per phase of the kernel, summed over warps, in a.prof[]
#ifdef BR2_PROFILE
#define PROF_START() long long prof_t0 = clock64()
#define PROF(idx)                                                                              \
    do {                                                                                       \
        const long long prof_t1 = clock64();                                                   \
        if (lane == 0 && a.prof) atomicAdd(a.prof + (idx), (unsigned long long)(prof_t1 - prof_t0)); \
        prof_t0 = prof_t1;                                                                     \
    } while (0)
#else
#define PROF_START()
#define PROF(idx)
#endif
enum { PF_FACTOR_ABS = 0, PF_FACTOR_AS, PF_FWD_CL, PF_PRIMAL, PF_COSTATE, PF_IPM_INIT, PF_FACTOR_IPM, PF_FWD_AFF, PF_E1, PF_BVEC,
       PF_FWD_COR, PF_E2, PF_EPILOGUE, PF_COUNT };

// The interior-point iteration.  (Tried as a function that is never inlined, to shield the register allocation of the fast
// paths from this code: the generic-space accesses to the kernel parameters and the spills at the call cost 40 % on the forced
// interior-point run and gained nothing on the fast path -- gpurun_out r2g.)
struct IpmOut { int status, it; double mu, res_stat, stat_scale, bmax; };
__device__ __forceinline__ void ipm_solve(Inst& I, IpmOut& out)
{
    const SolveArgs& a = I.a;
    const int lane = I.lane, N = I.N, nb = 4 * N;
    int status = 2, it = 0;
    double mu = 0.0, res_stat = 0.0, stat_scale = 1.0, bmax = 0.0;
    PROF_START();
    // ---------- interior-point iteration (Mehrotra predictor-corrector), the fallback ----------
    // Four sweeps per iteration and nothing else: the factor sweep applies the previous iteration's step on the way and returns
    // mu and the stationarity residual of the iterate; the two forward sweeps accumulate the step lengths and the sums that
    // centring and the stopping test need (IpmAcc).
    IpmUpd upd;
    IpmAcc acc;
    ipm_init(I);
    bmax = forward_sweep<0>(I);
    PROF(PF_IPM_INIT);
    const double inv2nb = rcp_newton(2.0 * nb);
    for (it = 0; it < a.max_iter; it++) {
        // ---------- B1: (pending update,) factorisation, predictor rhs, mu and residual of the iterate ----------
        const bool okf = factor_sweep<FS_IPM>(I, &upd);
        PROF(PF_FACTOR_IPM);
        if (!okf) { status = 4; break; }
        mu = upd.mu_sum * inv2nb;
        res_stat = upd.res_max;
        if (it == 0) stat_scale = fmax(1.0, res_stat);
        if (mu < a.tol && res_stat < a.tol * stat_scale) { status = 0; break; }
        // ---------- F1: affine step, its step length and complementarity -> sigma ----------
        forward_sweep<1>(I, &acc);
        PROF(PF_FWD_AFF);
        const double a_aff = acc.ia > 1.0 ? rcp_newton(acc.ia) : 1.0;       // (branch-free reciprocals: no out-of-line division code)
        const double mu_aff = ((1.0 - a_aff) * upd.mu_sum + a_aff * a_aff * acc.s2) * inv2nb;
        double sigma = mu_aff * rcp_newton(mu);
        sigma = sigma * sigma * sigma;
        // ---------- B2 / F2: corrector ----------
        prefetch_iterate(I);            // this may be the last iteration: have X, U in L2 for the epilogue
        backward_vec_sweep(I, sigma * mu);
        PROF(PF_BVEC);
        forward_sweep<3>(I, &acc);
        PROF(PF_FWD_COR);
        // ---------- step lengths; the update itself is left to the next factor sweep / the epilogue ----------
        double ap = acc.ia_p > 1.0 ? rcp_newton(acc.ia_p) : 1.0, ad = acc.ia_d > 1.0 ? rcp_newton(acc.ia_d) : 1.0;
        const double tau = fmin(fmax(0.995, 1.0 - mu), 1.0 - 1e-8);
        ap = fmin(1.0, tau * ap);
        ad = fmin(1.0, tau * ad);
        upd.pending = true; upd.ap = ap; upd.ad = ad;
        // mu of the new iterate in closed form, its stationarity residual bounded by |1 - ad| |r| + |ap - ad| |dgu| (exact when
        // ap == ad): if that already meets the tolerance the next factor sweep is not needed
        const double mu_new = (upd.mu_sum + ap * acc.a1 + ad * acc.a2 + ap * ad * acc.a3) * inv2nb;
        const double res_new = (1.0 - ad) * res_stat + fabs(ap - ad) * acc.dgmax;
        if (mu_new < a.tol && res_new < a.tol * stat_scale) { mu = mu_new; res_stat = res_new; status = 0; it++; break; }
    }
    PROF(PF_E2);
    if (upd.pending) {
        // the last step is still pending: apply it to the iterate in place (primal part; the multipliers are not used again)
        const double app = upd.ap;
        for (int idx = lane; idx < nb; idx += 32) {
            double* Vk = I.V + (size_t)(idx >> 2) * SREC;
            const int e = idx & 3;
            const double dvp = app * Vk[V_DV + e];
            Vk[V_V + e] += dvp; Vk[V_TL + e] += dvp; Vk[V_TU + e] -= dvp;
        }
        for (int idx = lane; idx < 12 * (N + 1); idx += 32) {
            double* Vk = I.V + (size_t)(idx / 12) * SREC;
            Vk[V_X + idx % 12] = fma(app, Vk[V_DX + idx % 12], Vk[V_X + idx % 12]);
        }
        __syncwarp();
    }
    out.status = status; out.it = it; out.mu = mu; out.res_stat = res_stat; out.stat_scale = stat_scale; out.bmax = bmax;
}

// Maximum number of pinned-LQR solves of the primal-dual active-set iteration before the interior-point iteration takes over
#ifndef BR2_MAX_AS
#define BR2_MAX_AS 6
#endif

// Epilogue of an instance (both kernels): full SQP step on (X, U), u0, thrust allocation (bluerov2_dob.cpp:388-395), status,
// statistics, the hint and the place in the next solve's visiting order.  Both solution paths leave (dx, du) in V_X, V_V.
struct SolveOut { int status, it; double mu, res_stat, stat_scale, bmax; bool solved, active; };
__device__ __forceinline__ void finish_instance(Inst& I, const SolveOut& r, int* order_next)
{
    const SolveArgs& a = I.a;
    const int lane = I.lane, N = I.N, nb = 4 * N, inst = I.inst;
    int status = r.status;
    const int it = r.it;
    const bool solved = r.solved;
    bool active = r.active;
    const double mu = r.mu, res_stat = r.res_stat, stat_scale = r.stat_scale, bmax = r.bmax;
    // ---------- epilogue: full SQP step, u0, thrust allocation (both paths leave (dx, du) in V_X, V_V) ----------
    double* Xo = a.X + (size_t)inst * (N + 1) * NX;
    double* Uo = a.U + (size_t)inst * N * NU;
    bool finite = true;
    if (!solved) {
        // (a candidate the fast paths accepted has been tested by its roll-out: forward_sweep<2>, PC_NAN)
        bool act2 = false;
#pragma unroll 5
        for (int idx = lane; idx < nb; idx += 32) {
            const double* Vk = I.V + (size_t)(idx >> 2) * SREC;
            finite &= isfinite(Vk[V_V + (idx & 3)]);
            act2 |= fmin(Vk[V_TL + (idx & 3)], Vk[V_TU + (idx & 3)]) < 1e-3;
        }
        // the states are the exact roll-out of the inputs: NaN/Inf anywhere reaches x_N
        if (lane < 12) finite &= isfinite(I.V[(size_t)N * SREC + V_X + lane]);
        finite = __all_sync(FULL_MASK, finite);
        active = __any_sync(FULL_MASK, act2);
    }
    if (a.active_set && !solved && status == 0) {
        // the interior-point solution's active set (slack ~ mu / lam at an active bound) is the next solve's guess
        for (int base = 0; base < nb; base += 32) {
            const int idx = base + lane;
            const bool valid = idx < nb;
            const double* Vk = I.V + (size_t)((valid ? idx : 0) >> 2) * SREC;
            int cc = 0;
            if (valid) cc = Vk[V_TL + (idx & 3)] < 1e-7 ? 1 : (Vk[V_TU + (idx & 3)] < 1e-7 ? 2 : 0);
            int code = cc << (2 * (lane & 3));
            code |= __shfl_xor_sync(FULL_MASK, code, 1);
            code |= __shfl_xor_sync(FULL_MASK, code, 2);
            if (valid && (lane & 3) == 0) a.aset[(size_t)inst * N + (idx >> 2)] = code;
        }
    }
    int pos = 0;
    if (lane == 0) {
        const int hard = (active || status != 0) ? 1 : 0;
        a.hint[inst] = hard;
        // position in the next solve's visiting order: hard instances from the front, easy ones from the back (the counter's round
        // trip overlaps the update of the iterate; the position is stored at the end)
        pos = hard ? atomicAdd(a.ctr + CTR_HARD, 1) : a.B - 1 - atomicAdd(a.ctr + CTR_EASY, 1);
    }
    double unew0 = 0.0;                         // lanes 0..3: the new U_0
    if (finite) {
        // X / U were last touched by the lineariser, before ~300 MB of stage records went through L2: without care this is
        // a chain of DRAM round trips (it was 11 % of the kernel).  The lines are prefetched into L2 ahead of the forward
        // sweep (prefetch_iterate) and each batch issues all of its loads before its first store.
        for (int base = lane; base < nb; base += 32 * 5) {
            double d[5], u[5];
#pragma unroll
            for (int j = 0; j < 5; j++) {
                const int idx = base + 32 * j;
                const bool p = idx < nb;
                d[j] = p ? I.V[(size_t)(idx >> 2) * SREC + V_V + (idx & 3)] : 0.0;
                u[j] = p ? Uo[idx] : 0.0;
            }
#pragma unroll
            for (int j = 0; j < 5; j++) {
                const int idx = base + 32 * j;
                // (a pinned input lands on its bound up to one rounding of (bound - U) + U, an accepted free input within 1e-12
                // of the box width: clamp)
                if (idx < nb) {
                    const double un = fmin(fmax(u[j] + d[j], a.lbu[idx & 3]), a.ubu[idx & 3]);
                    Uo[idx] = un;
                    if (j == 0 && base == lane) unew0 = un;
                }
            }
        }
        const int nxs = 12 * (N + 1);
        for (int base = lane; base < nxs; base += 32 * 8) {
            double d[8], x[8];
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const int idx = base + 32 * j;
                const bool p = idx < nxs;
                d[j] = p ? I.V[(size_t)(idx / 12) * SREC + V_X + idx % 12] : 0.0;
                x[j] = p ? Xo[idx] : 0.0;
            }
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const int idx = base + 32 * j;
                if (idx < nxs) Xo[idx] = x[j] + d[j];
            }
        }
    } else {
        status = 1;
        if (lane < 4) unew0 = Uo[lane];
    }
    const double u00 = __shfl_sync(FULL_MASK, unew0, 0), u01 = __shfl_sync(FULL_MASK, unew0, 1);
    const double u02 = __shfl_sync(FULL_MASK, unew0, 2), u03 = __shfl_sync(FULL_MASK, unew0, 3);
    if (lane == 0) {
        order_next[pos] = inst;
        double u0[4] = {u00, u01, u02, u03};
        double th[6];
        thrust_alloc(u0, th);
#pragma unroll
        for (int i = 0; i < 4; i++) a.u0[(size_t)inst * 4 + i] = u0[i];
#pragma unroll
        for (int i = 0; i < 6; i++) a.thrust[(size_t)inst * 6 + i] = th[i];
        a.status[inst] = status;
        a.iters[inst] = it;
        atomicAdd(a.iter_total, (unsigned long long)it);
        if (status != 0) atomicAdd(a.bad_total, 1ULL);
        a.info[(size_t)inst * 4 + 0] = mu;
        a.info[(size_t)inst * 4 + 1] = res_stat;
        a.info[(size_t)inst * 4 + 2] = bmax;
        a.info[(size_t)inst * 4 + 3] = stat_scale;
        if (a.shard.world > 1) {
            // sharded batch: a second copy of the thrust vector goes into this rank's block of the LOCAL gather buffer (slot = tick
            // parity); exchange_kernel ships the block to the peers while the next tick is already linearising
            double* dst = a.shard.buf[a.shard.rank] + ((size_t)(a.ctr[CTR_TICK] & 1) * a.shard.world * a.B + (size_t)a.shard.rank * a.B + inst) * 6;
#pragma unroll
            for (int i = 0; i < 6; i++) dst[i] = th[i];
        }
    }
    __syncwarp();
}

// Kernel 1 of the QP solve: the fast paths.  One warp per instance at a time, persistent over the batch.  What it cannot solve
// (no acceptance after BR2_MAX_AS attempts, a failed factorisation, option fast_path = 0) goes to the fallback list of kernel 2.
// (Without the interior-point code this kernel fits 96 / 80 registers with 32 / 48 bytes of spills, i.e. 20 / 24 resident warps per
// SM instead of 16 -- and gets SLOWER, 0.240 / 0.248 ms against 0.216 ms: the factor sweep is bound by the fp64 pipe, not by
// latency, and the tighter allocation costs instructions.  gpurun_out r2h, profiles/r02_pdas_occupancy.txt.)
__global__ void __launch_bounds__(IPM_WARPS * 32, BR2_PDAS_MINB) pdas_kernel(const __grid_constant__ SolveArgs a)
{
    __shared__ WarpSmem smem[IPM_WARPS];
    const int lane = threadIdx.x & 31;
    const int N = a.N, nb = 4 * N;
    WarpSmem& sm = smem[threadIdx.x >> 5];
    if (lane < 16) sm.xch[64 + lane] = a.W[lane];           // the warp's copy of the stage weights (factor_sweep)
    __syncwarp();
    // Work distribution.  Instances are visited in the order the PREVIOUS solve left behind: those that ended with active bounds
    // (hint = 1: several factorisations, possibly interior-point iterations) first, so that the long jobs start at t = 0 and the
    // one-factorisation instances fill the tail (longest-processing-time-first).  queue position -> instance through order_cur;
    // a warp's first position is its global warp index, later ones come from an atomic counter that starts behind the statically
    // assigned block and is reserved one instance ahead (the atomic's round trip was ~4 % of the kernel).  Only lane 0 holds the
    // reservation.  The epilogue appends the instance to order_next (hard ones from the front, easy ones from the back).
#if BR2_PDL
    asm volatile("griddepcontrol.wait;" ::: "memory");      // (launched early: the lineariser's records and counters are complete from here on)
#endif
    const int par = a.ctr[CTR_PARITY] & 1;
    const int* order_cur = a.order + (size_t)par * a.B;
    int* order_next = a.order + (size_t)(par ^ 1) * a.B;
    const int nwarps = gridDim.x * IPM_WARPS;
    // queue position -> instance: through the visiting order for the whole batch, lo + position for an instance range
    const int nq = a.hi - a.lo;
    const bool ranged = nq != a.B;
    auto inst_of = [&](int qpos) { return qpos < nq ? (ranged ? a.lo + qpos : order_cur[qpos]) : a.B; };
    int reserved = blockIdx.x * IPM_WARPS + (threadIdx.x >> 5);
    if (lane == 0) reserved = inst_of(reserved);
    bool have = true;
    // (Tried: starting every other warp of a scheduler late by 15 .. 60 us so that one half of the resident warps rolls out while
    // the other half factorises, on the theory that the lockstep phases fight for the fp64 pipe and then idle it together.  Every
    // delay made the kernel slower, 0.217 -> 0.222 .. 0.265 ms at B = 4096: gpurun_out r2s)
    for (;;) {
        int inst = reserved;
        if (!have && lane == 0) {
            const int qpos = nwarps + atomicAdd(a.ctr + CTR_QUEUE + a.qidx, 1);
            inst = inst_of(qpos);
        }
        have = false;
        inst = __shfl_sync(FULL_MASK, inst, 0);
        if (inst >= a.B) break;
        Inst I(a, sm, inst, lane);
        PROF_START();

        int status = 2, it = 0;
        double mu = 0.0, res_stat = 0.0, stat_scale = 1.0, bmax = 0.0;
        bool solved = false;
        bool active = false;                    // a bound is (nearly) active at the solution -> hint for the next solve
        // ---------- primal-dual active-set iteration on Riccati solves ----------
        // Attempt 0 without a guess is the interior-solution fast path: the minimiser of the QP without its box, one factorisation
        // + one closed-loop roll-out; if it lies inside the box it IS the minimiser of the QP (convexity).  Otherwise the inputs
        // outside their box are pinned at the violated bound (primal_check) and the LQR is solved again with those inputs fixed
        // (factor_sweep<FS_AS>); once a candidate respects the box, a costate sweep checks the multiplier signs of the pinned inputs
        // (costate_check) and releases the wrong ones.  A candidate that passes both tests satisfies the KKT conditions of the strictly
        // convex QP, so it is its unique minimiser -- the same point the interior-point iteration converges to.  This is the primal-
        // dual active-set (semismooth Newton) method; on this problem class it needs 2 solves from scratch and 1 from the previous
        // tick's active set (hint = 1: a.aset holds it).  BR2_MAX_AS solves without acceptance -> interior-point iteration.
        if (a.fast_path) {
            const bool guess = a.active_set && a.hint[inst] == 1;
            const int max_att = a.active_set ? BR2_MAX_AS : 1;
            for (int att = 0; att < max_att; att++) {
                const bool okf = (att == 0 && !guess) ? factor_sweep<FS_ABS>(I) : factor_sweep<FS_AS>(I);
                PROF((att == 0 && !guess) ? PF_FACTOR_ABS : PF_FACTOR_AS);
                if (!okf) break;
                if (att == 0) {
                    prefetch_iterate(I);
                    if (lane == 0) {
                        const int qpos = nwarps + atomicAdd(a.ctr + CTR_QUEUE + a.qidx, 1);
                        reserved = inst_of(qpos);
                    }
                    have = true;
                }
                const bool fresh = att == 0 && !guess;      // nothing pinned, the stage codes in a.aset are stale
                int fl = 0;
                bmax = forward_sweep<2>(I, nullptr, &fl);   // leaves (dx, du) in V_X, V_V
                PROF(PF_FWD_CL);
                it = att + 1;
                if (fresh && !(fl & (PC_CHANGED | PC_NAN))) {
                    // the candidate of the unconstrained LQR lies inside the box (tested by the roll-out itself): accepted below.  Near
                    // a bound the next solve starts from a guess (hint = 1): leave it the empty active set
                    if (fl & PC_ACTIVE) {
                        for (int k = lane; k < N; k += 32) a.aset[(size_t)inst * N + k] = 0;
                    }
                } else {
                    fl = primal_check(I, fresh) | (fl & PC_NAN);
                }
                PROF(PF_PRIMAL);
                if (fl & PC_NAN) break;
                active = (fl & PC_ACTIVE) != 0;
                if (!(fl & PC_CHANGED)) {
                    if (!(fl & PC_PINNED)) { solved = true; break; }
                    const bool okc = costate_check(I);
                    PROF(PF_COSTATE);
                    if (okc) { solved = true; break; }
                }
            }
            if (solved) status = 0;
        }
        if (!solved) {
            // hand the instance to the interior-point kernel
            if (lane == 0) a.fb[atomicAdd(a.ctr + CTR_FB, 1)] = inst;
            continue;
        }
        SolveOut r;
        r.status = status; r.it = it; r.mu = mu; r.res_stat = res_stat; r.stat_scale = stat_scale; r.bmax = bmax;
        r.solved = true; r.active = active;
        finish_instance(I, r, order_next);
        PROF(PF_EPILOGUE);
    }
}

// Kernel 2: Mehrotra predictor-corrector interior-point iteration for the instances on the fallback list (normally empty: the
// blocks read the count and leave).
__global__ void __launch_bounds__(IPM_WARPS * 32, BR2_IPM_MINB) ipm_kernel(const __grid_constant__ SolveArgs a)
{
    __shared__ WarpSmem smem[IPM_WARPS];
    const int lane = threadIdx.x & 31;
    WarpSmem& sm = smem[threadIdx.x >> 5];
    if (lane < 16) sm.xch[64 + lane] = a.W[lane];           // the warp's copy of the stage weights (factor_sweep)
    __syncwarp();
    const int nfb = a.ctr[CTR_FB];
    int* order_next = a.order + (size_t)((a.ctr[CTR_PARITY] & 1) ^ 1) * a.B;
    int pos = blockIdx.x * IPM_WARPS + (threadIdx.x >> 5);
    const int nwarps = gridDim.x * IPM_WARPS;
    while (pos < nfb) {
        const int inst = a.fb[pos];
        Inst I(a, sm, inst, lane);
        PROF_START();
        prefetch_iterate(I);
        IpmOut o;
        ipm_solve(I, o);
#ifdef BR2_PROFILE
        prof_t0 = clock64();            // (ipm_solve accounts for its own phases)
#endif
        SolveOut r;
        r.status = o.status; r.it = o.it; r.mu = o.mu; r.res_stat = o.res_stat; r.stat_scale = o.stat_scale; r.bmax = o.bmax;
        r.solved = false; r.active = false;
        finish_instance(I, r, order_next);
        PROF(PF_EPILOGUE);
        if (lane == 0) pos = nwarps + atomicAdd(a.ctr + CTR_FBQ, 1);
        pos = __shfl_sync(FULL_MASK, pos, 0);
    }
}

// Sharded batch: ship this rank's block of the gather buffer (the thrust vectors of the tick that has just finished) to every
// peer's gather buffer over NVLink -- peer stores, no collective -- then publish the tick index in every rank's flag array.
// grid = (world - 1) x XCHG_CTAS blocks; block (p, j) copies slice j of the block to the p-th peer.  The exchange counter
// ctr[CTR_XTICK] counts the ticks shipped so far (tick index = count + 1, its parity selects the slot): the kernel of tick t
// runs as a side branch of tick t + 1's graph, concurrently with that tick's lineariser, so it cannot read the tick counter.
constexpr int XCHG_CTAS = 4;
__global__ void __launch_bounds__(256) exchange_kernel(ShardView sh, int B, int* ctr)
{
    const int tick = ctr[CTR_XTICK] + 1;
    int peer = blockIdx.x / XCHG_CTAS;
    if (peer >= sh.rank) peer++;                                         // skip myself
    const size_t row0 = ((size_t)(tick & 1) * sh.world * B + (size_t)sh.rank * B) * 6;
    const double2* src = reinterpret_cast<const double2*>(sh.buf[sh.rank] + row0);
    double2* dst = reinterpret_cast<double2*>(sh.buf[peer] + row0);
    const int n2 = B * 3;                                                // 6 doubles per instance = 3 double2
    for (int i = (blockIdx.x % XCHG_CTAS) * blockDim.x + threadIdx.x; i < n2; i += XCHG_CTAS * blockDim.x) dst[i] = src[i];
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const int done = atomicAdd(ctr + CTR_XDONE, 1);
        if (done == (int)gridDim.x - 1) {                                // the last block: every slice is on its way and fenced
            ctr[CTR_XDONE] = 0;
            ctr[CTR_XTICK] = tick;
            __threadfence_system();
            for (int r = 0; r < sh.world; r++) *reinterpret_cast<volatile int*>(sh.flag[r] + sh.rank) = tick;
        }
    }
}
void launch_exchange(const ShardView& sh, int B, int* ctr, cudaStream_t s)
{
    if (sh.world > 1) exchange_kernel<<<(sh.world - 1) * XCHG_CTAS, 256, 0, s>>>(sh, B, ctr);
}

// consumer side of the sharded exchange: returns when every rank has published tick index >= `tick` (bounded spin)
__global__ void shard_wait_kernel(const int* flag, int world, int tick, int* timed_out)
{
    const int r = threadIdx.x;
    if (r >= world) return;
    const volatile int* f = flag + r;
    long long spins = 0;
    while (*f < tick) {
        __nanosleep(200);
        if (++spins > 20000000LL) { *timed_out = 1; break; }      // ~4 s: a peer died
    }
    __threadfence_system();
}
void launch_shard_wait(const int* flag, int world, int tick, int* timed_out, cudaStream_t s)
{
    shard_wait_kernel<<<1, 32, 0, s>>>(flag, world, tick, timed_out);
}

void configure_kernels()
{
    // the resident blocks need MINB x WARPS x 9 KB of staging buffers: ask for the largest carve-out (function attributes
    // are per device: called from br2_batch_create with the solver's device current)
    cudaFuncSetAttribute(pdas_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(ipm_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
}

void launch_pdas(const SolveArgs& a, int sm_count, cudaStream_t s)
{
    // Persistent grid, one warp per instance at a time, instances handed out by an atomic queue (its counter is reset and the
    // order buffers flipped by block 0 of the lineariser that precedes this kernel in the stream).  (Sizing the resident
    // set for even waves -- 14 instead of 16 warps/SM at B = 4096 -- measured 8 % slower: throughput grows with the
    // number of resident warps and the queue already evens out the tail; profiles/r01h_ipm_variants.txt.)
    const int need = (a.hi - a.lo + IPM_WARPS - 1) / IPM_WARPS;
    const int blocks = need < sm_count * BR2_PDAS_MINB ? need : sm_count * BR2_PDAS_MINB;
#if BR2_PDL
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(blocks); cfg.blockDim = dim3(IPM_WARPS * 32); cfg.dynamicSmemBytes = 0; cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, pdas_kernel, a);
#else
    pdas_kernel<<<blocks, IPM_WARPS * 32, 0, s>>>(a);
#endif
}

void launch_ipm_fallback(const SolveArgs& a, int sm_count, cudaStream_t s)
{
    const int need = (a.B + IPM_WARPS - 1) / IPM_WARPS;
    const int blocks = need < sm_count * BR2_IPM_MINB ? need : sm_count * BR2_IPM_MINB;
    ipm_kernel<<<blocks, IPM_WARPS * 32, 0, s>>>(a);
}

void launch_ipm(const SolveArgs& a, int sm_count, cudaStream_t s)
{
    launch_pdas(a, sm_count, s);
    launch_ipm_fallback(a, sm_count, s);
}

}  // namespace br2
