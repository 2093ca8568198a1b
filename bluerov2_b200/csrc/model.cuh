// model.cuh -- BlueROV2 OCP model: dynamics f(x,u,p) and the action of its sparse Jacobian.
//
// Hand-derived (not generated) from the symbolic definition in the reference,
// bluerov2_dobmpc/scripts/bluerov2.py:77-137, whose CasADi expansion is
// c_generated_code/bluerov2_model/bluerov2_expl_ode_fun.c:66-321 (f) and
// bluerov2_expl_vde_forw.c:73 (f, Jx*Sx, Jx*Su + Ju).  Where CasADi's forward VDE spends ~4.35 kflop and
// 147 sin/cos per call on a dense 12x16 seed, this form needs 3 sincos + the 48 structural non-zeros of
// Jx per evaluation and 48 FMAs per propagated column.
//
// Reference quirks reproduced on purpose (parity): dphi uses sin(psi) (bluerov2.py:133); dtheta = cos(phi) q +
// sin(phi) r (:134); no Coriolis terms in du/dv/dw (:119-121); propulsion matrix rounded to 0.707/0.167/0.175
// with zero roll/pitch rows (:95-100); d(|v| v)/dv = 2|v| (CasADi: sign(v) v + |v|, sign(0) = 0).
//
// State  x = (x y z phi theta psi u v w p q r), control u = (u1..u4), parameters p[16] =
// (dist_x dist_y dist_z dist_psi | added mass x y z n | linear damping x y z n | quadratic damping x y z n).
#pragma once
#include <math.h>
#include "fast_trig.h"

#if defined(__CUDACC__)
#define BR2_HD __host__ __device__ __forceinline__
#else
#define BR2_HD inline
#endif

namespace br2 {

constexpr int NX = 12, NU = 4, NP = 16, NY = 16;

constexpr double MASS = 11.26, IX = 0.3, IY = 0.63, IZ = 0.58, ZG = 0.02, GRAV = 9.81;
constexpr double BUOY = 0.66;                       // bluerov2.py:83
constexpr double RC = 0.026546960744430276;         // rotor_constant, bluerov2.py:84
// K * t(u) collapsed (bluerov2.py:95-115): Kt0 = -4*0.707 u1/rc, Kt1 = +4*0.707 u2/rc, Kt2 = -2 u3/rc,
// Kt5 = (2*0.167 - 2*0.175) u2/rc + (2*0.167 + 2*0.175) u4/rc, Kt3 = Kt4 = 0.
constexpr double KT_SURGE = -4.0 * 0.707 / RC;
constexpr double KT_SWAY = 4.0 * 0.707 / RC;
constexpr double KT_HEAVE = -2.0 / RC;
constexpr double KT_YAW_U2 = (2.0 * 0.167 - 2.0 * 0.175) / RC;
constexpr double KT_YAW_U4 = (2.0 * 0.167 + 2.0 * 0.175) / RC;
constexpr double MZG = MASS * ZG * GRAV;

// reciprocal: branch-free on the device (fast_trig.h), IEEE division on the host (the CasADi-ABI exports)
BR2_HD double inv_of(double x)
{
#ifdef __CUDA_ARCH__
    return br2_rcp(x);
#else
    return 1.0 / x;
#endif
}

// Parameter-only constants, hoisted out of the RK stages.
struct ModelConst {
    double imx, imy, imz, imn;       // 1/(m + added mass), 1/(Iz + added mass n)
    double dist[4];                  // p0..p3
    double dl[4], dnl[4];            // p8..p11, p12..p15
    double ju_surge, ju_sway, ju_heave, ju_yaw2, ju_yaw4;   // the 5 constant non-zeros of df/du
    BR2_HD void set(const double* p)
    {
        imx = inv_of(MASS + p[4]); imy = inv_of(MASS + p[5]); imz = inv_of(MASS + p[6]); imn = inv_of(IZ + p[7]);
        for (int i = 0; i < 4; i++) { dist[i] = p[i]; dl[i] = p[8 + i]; dnl[i] = p[12 + i]; }
        ju_surge = KT_SURGE * imx; ju_sway = KT_SWAY * imy; ju_heave = KT_HEAVE * imz;
        ju_yaw2 = KT_YAW_U2 * imn; ju_yaw4 = KT_YAW_U4 * imn;
    }
};

struct Trig { double sphi, cphi, sth, cth, spsi, cpsi; };

// (device: branch-free sincos / reciprocal, fast_trig.h -- the library routines' slow-path branches are scheduling barriers that keep
// the three evaluations from overlapping; host -- the CasADi-ABI exports -- : libm)
BR2_HD void trig_of(const double* x, Trig& t)
{
#ifdef __CUDA_ARCH__
    br2_sincos(x[3], &t.sphi, &t.cphi);
    br2_sincos(x[4], &t.sth, &t.cth);
    br2_sincos(x[5], &t.spsi, &t.cpsi);
#else
    sincos(x[3], &t.sphi, &t.cphi);
    sincos(x[4], &t.sth, &t.cth);
    sincos(x[5], &t.spsi, &t.cpsi);
#endif
}


// f(x,u,p)
BR2_HD void ode(const double* x, const double* u, const ModelConst& c, const Trig& t, double* f)
{
    const double uu = x[6], v = x[7], w = x[8], pp = x[9], q = x[10], r = x[11];
#ifdef __CUDA_ARCH__
    const double icth = inv_of(t.cth), tth = t.sth * icth;
#else
    const double tth = t.sth / t.cth, icth = 1.0 / t.cth;
#endif
    f[0] = (t.cpsi * t.cth) * uu + (-t.spsi * t.cphi + t.cpsi * t.sth * t.sphi) * v + (t.spsi * t.sphi + t.cpsi * t.cphi * t.sth) * w;
    f[1] = (t.spsi * t.cth) * uu + (t.cpsi * t.cphi + t.sphi * t.sth * t.spsi) * v + (-t.cpsi * t.sphi + t.sth * t.spsi * t.cphi) * w;
    f[2] = (-t.sth) * uu + (t.cth * t.sphi) * v + (t.cth * t.cphi) * w;
    f[3] = pp + (t.spsi * tth) * q + (t.cphi * tth) * r;
    f[4] = t.cphi * q + t.sphi * r;
    f[5] = (t.sphi * icth) * q + (t.cphi * icth) * r;
    f[6] = c.imx * (KT_SURGE * u[0] - BUOY * t.sth + c.dist[0] + c.dl[0] * uu + c.dnl[0] * fabs(uu) * uu);
    f[7] = c.imy * (KT_SWAY * u[1] + BUOY * t.cth * t.sphi + c.dist[1] + c.dl[1] * v + c.dnl[1] * fabs(v) * v);
    f[8] = c.imz * (KT_HEAVE * u[2] + BUOY * t.cth * t.cphi + c.dist[2] + c.dl[2] * w + c.dnl[2] * fabs(w) * w);
    f[9] = (1.0 / IX) * ((IY - IZ) * q * r - MZG * t.cth * t.sphi);
    f[10] = (1.0 / IY) * ((IZ - IX) * pp * r - MZG * t.sth);
    f[11] = c.imn * (KT_YAW_U2 * u[1] + KT_YAW_U4 * u[3] - (IY - IX) * pp * q + c.dist[3] + c.dl[3] * r + c.dnl[3] * fabs(r) * r);
}

// The 48 structural non-zeros of Jx = df/dx at x (columns 0..2 are identically zero).
struct Jac {
    double j03, j04, j05, j06, j07, j08;
    double j13, j14, j15, j16, j17, j18;
    double j23, j24, j26, j27, j28;
    double j33, j34, j35, j3a, j3b;          // j39 = 1
    double j43, j4a, j4b;
    double j53, j54, j5a, j5b;
    double j64, j66;
    double j73, j74, j77;
    double j83, j84, j88;
    double j93, j94, j9a, j9b;
    double ja4, ja9, jab;
    double jb9, jba, jbb;
};

BR2_HD void jac_of(const double* x, const ModelConst& c, const Trig& t, Jac& J)
{
    const double uu = x[6], v = x[7], w = x[8], pp = x[9], q = x[10], r = x[11];
    const double icth = inv_of(t.cth), tth = t.sth * icth, sec2 = icth * icth;
    // rows 0..2: d(R v_b)/d(angles, v_b) with R = Rz(psi) Ry(theta) Rx(phi).  The angle columns follow from R and
    // w = R v_b:  d/dphi = v R[:,2] - w R[:,1],  d/dtheta = (cos psi, sin psi, .) * w_z,  d/dpsi = (-w_y, w_x, 0).
    const double cs = t.cpsi * t.sth, ss = t.spsi * t.sth;
    J.j06 = t.cpsi * t.cth;
    J.j07 = cs * t.sphi - t.spsi * t.cphi;
    J.j08 = cs * t.cphi + t.spsi * t.sphi;
    J.j16 = t.spsi * t.cth;
    J.j17 = ss * t.sphi + t.cpsi * t.cphi;
    J.j18 = ss * t.cphi - t.cpsi * t.sphi;
    J.j26 = -t.sth;
    J.j27 = t.cth * t.sphi;
    J.j28 = t.cth * t.cphi;
    const double wx = J.j06 * uu + J.j07 * v + J.j08 * w;
    const double wy = J.j16 * uu + J.j17 * v + J.j18 * w;
    const double wz = J.j26 * uu + J.j27 * v + J.j28 * w;
    J.j03 = J.j08 * v - J.j07 * w;
    J.j13 = J.j18 * v - J.j17 * w;
    J.j23 = J.j28 * v - J.j27 * w;
    J.j04 = t.cpsi * wz;
    J.j14 = t.spsi * wz;
    J.j24 = -(t.cth * uu + t.sth * (t.sphi * v + t.cphi * w));
    J.j05 = -wy;
    J.j15 = wx;
    // rows 3..5: Euler-angle rates (with the reference's sin(psi) in dphi)
    J.j3a = t.spsi * tth;
    J.j3b = t.cphi * tth;
    J.j5a = t.sphi * icth;
    J.j5b = t.cphi * icth;
    J.j33 = -t.sphi * tth * r;
    J.j34 = (t.spsi * q + t.cphi * r) * sec2;
    J.j35 = t.cpsi * tth * q;
    J.j43 = -t.sphi * q + t.cphi * r;
    J.j4a = t.cphi;
    J.j4b = t.sphi;
    J.j53 = J.j5b * q - J.j5a * r;
    J.j54 = (J.j5a * q + J.j5b * r) * tth;
    // rows 6..11: kinetics
    J.j64 = -BUOY * t.cth * c.imx;
    J.j66 = (c.dl[0] + 2.0 * c.dnl[0] * fabs(uu)) * c.imx;
    J.j73 = BUOY * J.j28 * c.imy;
    J.j74 = -BUOY * t.sth * t.sphi * c.imy;
    J.j77 = (c.dl[1] + 2.0 * c.dnl[1] * fabs(v)) * c.imy;
    J.j83 = -BUOY * J.j27 * c.imz;
    J.j84 = -BUOY * t.sth * t.cphi * c.imz;
    J.j88 = (c.dl[2] + 2.0 * c.dnl[2] * fabs(w)) * c.imz;
    J.j93 = -MZG * J.j28 * (1.0 / IX);
    J.j94 = MZG * t.sth * t.sphi * (1.0 / IX);
    J.j9a = (IY - IZ) * r * (1.0 / IX);
    J.j9b = (IY - IZ) * q * (1.0 / IX);
    J.ja4 = -MZG * t.cth * (1.0 / IY);
    J.ja9 = (IZ - IX) * r * (1.0 / IY);
    J.jab = (IZ - IX) * pp * (1.0 / IY);
    J.jb9 = -(IY - IX) * q * c.imn;
    J.jba = -(IY - IX) * pp * c.imn;
    J.jbb = (c.dl[3] + 2.0 * c.dnl[3] * fabs(r)) * c.imn;
}

// f(x,u,p) from the Jacobian's intermediates (rows 0..5 of f are linear in the body velocities with exactly the
// coefficients J holds), for kernels that need both: ~35 flops instead of a second trig expansion.
BR2_HD void ode_from_jac(const double* x, const double* u, const ModelConst& c, const Trig& t, const Jac& J, double* f)
{
    const double uu = x[6], v = x[7], w = x[8], pp = x[9], q = x[10], r = x[11];
    f[0] = J.j15;                                   // = R[0,:] v_b
    f[1] = -J.j05;                                  // = R[1,:] v_b
    f[2] = J.j26 * uu + J.j27 * v + J.j28 * w;
    f[3] = pp + J.j3a * q + J.j3b * r;
    f[4] = J.j4a * q + J.j4b * r;
    f[5] = J.j5a * q + J.j5b * r;
    f[6] = c.imx * (KT_SURGE * u[0] - BUOY * t.sth + c.dist[0] + c.dl[0] * uu + c.dnl[0] * fabs(uu) * uu);
    f[7] = c.imy * (KT_SWAY * u[1] + BUOY * J.j27 + c.dist[1] + c.dl[1] * v + c.dnl[1] * fabs(v) * v);
    f[8] = c.imz * (KT_HEAVE * u[2] + BUOY * J.j28 + c.dist[2] + c.dl[2] * w + c.dnl[2] * fabs(w) * w);
    f[9] = (1.0 / IX) * ((IY - IZ) * q * r - MZG * J.j27);
    f[10] = (1.0 / IY) * ((IZ - IX) * pp * r - MZG * t.sth);
    f[11] = c.imn * (KT_YAW_U2 * u[1] + KT_YAW_U4 * u[3] - (IY - IX) * pp * q + c.dist[3] + c.dl[3] * r + c.dnl[3] * fabs(r) * r);
}

// out = Jx * s for one 12-vector s (a column of Sx or Su); 48 multiply-adds.
BR2_HD void jac_mul(const Jac& J, const double* s, double* o)
{
    o[0] = J.j03 * s[3] + J.j04 * s[4] + J.j05 * s[5] + J.j06 * s[6] + J.j07 * s[7] + J.j08 * s[8];
    o[1] = J.j13 * s[3] + J.j14 * s[4] + J.j15 * s[5] + J.j16 * s[6] + J.j17 * s[7] + J.j18 * s[8];
    o[2] = J.j23 * s[3] + J.j24 * s[4] + J.j26 * s[6] + J.j27 * s[7] + J.j28 * s[8];
    o[3] = J.j33 * s[3] + J.j34 * s[4] + J.j35 * s[5] + s[9] + J.j3a * s[10] + J.j3b * s[11];
    o[4] = J.j43 * s[3] + J.j4a * s[10] + J.j4b * s[11];
    o[5] = J.j53 * s[3] + J.j54 * s[4] + J.j5a * s[10] + J.j5b * s[11];
    o[6] = J.j64 * s[4] + J.j66 * s[6];
    o[7] = J.j73 * s[3] + J.j74 * s[4] + J.j77 * s[7];
    o[8] = J.j83 * s[3] + J.j84 * s[4] + J.j88 * s[8];
    o[9] = J.j93 * s[3] + J.j94 * s[4] + J.j9a * s[10] + J.j9b * s[11];
    o[10] = J.ja4 * s[4] + J.ja9 * s[9] + J.jab * s[11];
    o[11] = J.jb9 * s[9] + J.jba * s[10] + J.jbb * s[11];
}

// out = Jx * s + c for a column of Su, where c = (df/du)[:, a] has non-zeros in rows 6, 7, 8, 11 only: the constant rides
// in the accumulator of those rows' first multiply-add.
BR2_HD void jac_mul_add(const Jac& J, const double* s, double c6, double c7, double c8, double c11, double* o)
{
    o[0] = J.j03 * s[3] + J.j04 * s[4] + J.j05 * s[5] + J.j06 * s[6] + J.j07 * s[7] + J.j08 * s[8];
    o[1] = J.j13 * s[3] + J.j14 * s[4] + J.j15 * s[5] + J.j16 * s[6] + J.j17 * s[7] + J.j18 * s[8];
    o[2] = J.j23 * s[3] + J.j24 * s[4] + J.j26 * s[6] + J.j27 * s[7] + J.j28 * s[8];
    o[3] = J.j33 * s[3] + J.j34 * s[4] + J.j35 * s[5] + s[9] + J.j3a * s[10] + J.j3b * s[11];
    o[4] = J.j43 * s[3] + J.j4a * s[10] + J.j4b * s[11];
    o[5] = J.j53 * s[3] + J.j54 * s[4] + J.j5a * s[10] + J.j5b * s[11];
    o[6] = fma(J.j64, s[4], c6) + J.j66 * s[6];
    o[7] = fma(J.j73, s[3], c7) + J.j74 * s[4] + J.j77 * s[7];
    o[8] = fma(J.j83, s[3], c8) + J.j84 * s[4] + J.j88 * s[8];
    o[9] = J.j93 * s[3] + J.j94 * s[4] + J.j9a * s[10] + J.j9b * s[11];
    o[10] = J.ja4 * s[4] + J.ja9 * s[9] + J.jab * s[11];
    o[11] = fma(J.jb9, s[9], c11) + J.jba * s[10] + J.jbb * s[11];
}

// o += (df/du)[:, a]  (constant in x and u)
BR2_HD void ju_add(const ModelConst& c, int a, double* o)
{
    if (a == 0) o[6] += c.ju_surge;
    else if (a == 1) { o[7] += c.ju_sway; o[11] += c.ju_yaw2; }
    else if (a == 2) o[8] += c.ju_heave;
    else o[11] += c.ju_yaw4;
}

// 4 -> 6 thrust allocation, bluerov2_dob.cpp:390-395 (== bluerov2_ctrl.cpp:256-266)
BR2_HD void thrust_alloc(const double* u, double* t)
{
#ifdef __CUDA_ARCH__
    // (device: times the reciprocal -- a division brings its out-of-line special-case code into the QP kernels' epilogue)
    constexpr double IRC = 1.0 / RC;
    t[0] = (-u[0] + u[1] + u[3]) * IRC;
    t[1] = (-u[0] - u[1] - u[3]) * IRC;
    t[2] = (u[0] + u[1] - u[3]) * IRC;
    t[3] = (u[0] - u[1] + u[3]) * IRC;
    t[4] = (-u[2]) * IRC;
    t[5] = (-u[2]) * IRC;
#else
    t[0] = (-u[0] + u[1] + u[3]) / RC;
    t[1] = (-u[0] - u[1] - u[3]) / RC;
    t[2] = (u[0] + u[1] - u[3]) / RC;
    t[3] = (u[0] - u[1] + u[3]) / RC;
    t[4] = (-u[2]) / RC;
    t[5] = (-u[2]) / RC;
#endif
}

}  // namespace br2
