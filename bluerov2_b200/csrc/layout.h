// layout.h -- HBM data layout of the batched SQP-RTI engine (see DESIGN.md "Data layout").
//
// One OCP instance is solved by one warp, so "coalesced" means: every per-stage record is a contiguous,
// 128-byte-aligned block that the 32 lanes (or one bulk copy) fetch with 16-byte accesses.
#pragma once

namespace br2 {

// Stage record G_k written by the linearisation kernel, read by every sweep of the IPM:
//   [0..191]   G = [A_k | B_k], 12 x 16 row-major  (row l: A[l][0..11], B[l][0..3])
//   [192..203] b_k = Phi(X_k,U_k) - X_{k+1}
//   [204..207] pad (record = 1664 B = 13 x 128 B)
constexpr int GREC = 208;
constexpr int G_B_OFF = 192;

// Factor record F_k written by the backward factorisation, read by the vector sweeps:
//   [0..47]  Kt[j][a] = K[a][j]  (feedback gain, transposed so lane j owns 4 contiguous doubles)
//   [48..53] strictly-lower Cholesky entries of Lam = R~ + B'PB: l10 l20 l21 l30 l31 l32
//   [54..57] reciprocal diagonal 1/l00 .. 1/l33
//   [58..63] pad (record = 512 B)
constexpr int FREC = 64;
constexpr int F_L_OFF = 48;
constexpr int F_ID_OFF = 54;

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
constexpr int V_G = 60;    // g = gh + B'p+   (kff = Lam^-1 g)

}  // namespace br2
