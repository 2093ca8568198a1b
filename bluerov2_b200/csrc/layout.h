// layout.h -- HBM data layout of the batched SQP-RTI engine (see DESIGN.md "Data layout").
//
// One OCP instance is solved by one warp.  Lane l = 4 q + t  (q = l >> 2 in 0..7, t = l & 3), the thread
// coordinates of the fp64 tensor-core instruction mma.sync.m8n8k4 (DMMA): an A fragment holds element
// (row q, col t), a B fragment (row t, col q), a C fragment (row q, cols 2t, 2t+1).
#pragma once

namespace br2 {

// Stage record G_k written by the linearisation kernel, read by every sweep of the IPM.
//   Z = [A_k | B_k] is 12 x 16.  It is stored in FRAGMENT ORDER: six blocks of 32 doubles,
//   block (ki, mi), ki = 0..2, mi = 0..1, lane l holds Z[4 ki + t][8 mi + q].
//   -> the backward sweeps load block (ki, mi) as one fully coalesced 256-byte access and use it directly as the
//      DMMA A fragment of Z' (tile mi, ki) and as the B fragment of Z (tile ki, mi);
//   -> the forward sweeps read Z[8 mi + q][4 ki + t] (row-per-quad), which in this order is two full 128-byte lines.
//   [192..203] b_k = Phi(X_k,U_k) - X_{k+1}
//   [204..215] qlin_k = Ts_k W_x (X_k - xref_k)      gradient of the stage cost at the linearisation point
//   [216..219] rlin_k = Ts_k W_u (U_k - uref_k)
//   [220]      Ts_k;  [221..223] pad                  (record = 1792 B = 14 x 128 B)
// With these the IPM sweeps need nothing but the G, F and V records of a stage: each is one contiguous, 128-byte
// aligned block that a single lane prefetches into shared memory with cp.async.bulk (TMA bulk copy).
constexpr int GREC = 224;
constexpr int G_B_OFF = 192;
constexpr int G_QLIN = 204;
constexpr int G_RLIN = 216;
constexpr int G_TS = 220;
// offset of Z[row][col] inside a record
__host__ __device__ constexpr int g_off(int row, int col)
{
    return (((row >> 2) * 2 + (col >> 3)) << 5) + ((col & 7) << 2) + (row & 3);
}

// Factor record F_k written by the backward factorisation, read by the vector sweeps (record = 512 B).  Two layouts:
//  (a) after an interior-point factorisation (Lam = R~ + B'PB):
//   [0..47]  Kt[j][a] = K[a][j]  (feedback gain transposed: column j of K is 4 contiguous doubles)
//   [48..57] upper triangle of Lam^-1, row by row (the corrector's backward sweep forms its own feed-forward with it)
//   [58..61] feed-forward kff = Lam^-1 g of the most recent backward sweep (16-byte aligned);  [62..63] pad
//  (b) after an absolute-form factorisation (fast paths): [K | kff] = Lam^-1 [H_ux | g] (4 x 13, padded to 4 x 16) exactly as
//   the DMMA that forms it leaves it in registers -- C-fragment order: K[a][c] at f_kc(a, c); kff[a] = column 12
constexpr int FREC = 64;
constexpr int F_L_OFF = 48;
constexpr int F_KFF = 58;
// layout (b): C register j = c & 1 of tile n = c >> 3 of lane 4 a + ((c & 7) >> 1); register (n, j) of lanes 0..15 is one 128-byte row
__host__ __device__ constexpr int f_kc(int a, int c) { return (((c >> 3) * 2 + (c & 1)) << 4) + 4 * a + ((c & 7) >> 1); }

// Vector record V_k (IPM iterate + step, per stage; record = 512 B):
constexpr int VREC = 64;
constexpr int V_X = 0;     // dx_k   iterate state (roll-out of the current du)
constexpr int V_DX = 12;   // step in dx_k
constexpr int V_V = 24;    // du_k   iterate
constexpr int V_TL = 28;   // slack of  du - lb >= 0
constexpr int V_TU = 32;   // slack of  ub - du >= 0
constexpr int V_LL = 36;   // multiplier lower
constexpr int V_LU = 40;   // multiplier upper
constexpr int V_GU = 44;   // reduced input gradient R du + r + B'pi+
constexpr int V_DV = 48;   // step in du_k
constexpr int V_CL = 52;   // complementarity rhs lower: sigma*mu - dt_aff*dlam_aff
constexpr int V_CU = 56;   // complementarity rhs upper
// [60..63] pad

// Stage record S_k of an instance, k = 0..N (the terminal record only uses its V part): the three records of a stage laid
// end to end, S_k = [V_k | G_k | F_k], so that a sweep addresses everything of a stage from ONE running pointer and the
// staging copies of a stage are 16-byte chunks at fixed offsets of it.  352 doubles = 2816 B = 22 x 128 B.
constexpr int SREC = VREC + GREC + FREC;
constexpr int S_V = 0;
constexpr int S_G = VREC;
constexpr int S_F = VREC + GREC;

}  // namespace br2
