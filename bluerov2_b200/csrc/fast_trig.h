// fast_trig.h -- branch-free double-precision sincos and reciprocal for the device code paths.
//
// CUDA's sincos() carries a slow path (Payne-Hanek reduction behind a divergent branch for |x| > 105615) and its IEEE division a
// special-case branch; inside the dependent chains of the lineariser a divergent-branch region is a scheduling barrier that keeps the
// three sincos of an Euler-angle triple from overlapping.  br2_sincos: Cody-Waite reduction with a three-part pi/2 (exact first
// step: the fused multiply-add keeps x - k HI without rounding for every |x| < 2^52), the fdlibm kernel polynomials on
// [-pi/4, pi/4], quadrant by selects.  Error <= 1.5 ulp for |x| < 1e5 (tests/test_fast_trig.py checks it against libm on the
// host build of this header); NaN / Inf in -> NaN out.  The angles of this model are Euler angles and an unwrapped yaw.
#pragma once
#include <cmath>

#if defined(__CUDACC__)
#define BR2_TRIG_HD __host__ __device__ __forceinline__
#else
#define BR2_TRIG_HD inline
#endif

BR2_TRIG_HD void br2_sincos(double x, double* sn, double* cs)
{
    const double TWO_OVER_PI = 0x1.45f306dc9c883p-1;
    const double PIO2_HI = 0x1.921fb54442d18p+0, PIO2_MID = 0x1.1a62633145c07p-54, PIO2_LO = -0x1.f1976b7ed8fbcp-110;
    const double MAGIC = 6755399441055744.0;                 // 1.5 * 2^52: adding it rounds to the nearest integer
    const double kb = fma(x, TWO_OVER_PI, MAGIC);
    const double k = kb - MAGIC;
#if defined(__CUDA_ARCH__)
    const int n = __double2loint(kb);                        // the integer sits in the low word of the biased sum
#else
    long long bits;
    __builtin_memcpy(&bits, &kb, sizeof bits);
    const int n = (int)(unsigned)bits;
#endif
    double r = fma(-k, PIO2_HI, x);
    r = fma(-k, PIO2_MID, r);
    r = fma(-k, PIO2_LO, r);
    const double z = r * r;
    // sin r = r + r^3 (S1 + z (S2 + ... )), cos r = 1 - z/2 + z^2 (C1 + z (C2 + ...))   (fdlibm k_sin.c / k_cos.c)
    double ps = fma(z, 1.58969099521155010221e-10, -2.50507602534068634195e-08);
    ps = fma(z, ps, 2.75573137070700676789e-06);
    ps = fma(z, ps, -1.98412698298579493134e-04);
    ps = fma(z, ps, 8.33333333332248946124e-03);
    ps = fma(z, ps, -1.66666666666666324348e-01);
    const double sr = fma(z * r, ps, r);
    double pc = fma(z, -1.13596475577881948265e-11, 2.08757232129817482790e-09);
    pc = fma(z, pc, -2.75573143513906633035e-07);
    pc = fma(z, pc, 2.48015872894767294178e-05);
    pc = fma(z, pc, -1.38888888888741095749e-03);
    pc = fma(z, pc, 4.16666666666666019037e-02);
    const double hz = 0.5 * z, w = 1.0 - hz;
    const double cr = w + (((1.0 - w) - hz) + (z * z) * pc);
    const bool swap = n & 1;
    double s0 = swap ? cr : sr, c0 = swap ? sr : cr;
    if (n & 2) s0 = -s0;
    if ((n + 1) & 2) c0 = -c0;
    *sn = s0; *cs = c0;
}

// 1 / x to about an ulp without the division's special-case branch: hardware seed (~20 bits) and two Newton steps on the device
BR2_TRIG_HD double br2_rcp(double x)
{
#if defined(__CUDA_ARCH__)
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    y = fma(y, fma(-x, y, 1.0), y);
    return fma(y, fma(-x, y, 1.0), y);
#else
    return 1.0 / x;
#endif
}
