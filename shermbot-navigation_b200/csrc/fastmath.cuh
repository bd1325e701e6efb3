// fastmath.cuh -- short-dependency-chain fp64 helpers for the scalar part of the FAST EKF update.
//
// The scalar part of an update (rsqrt, atan2, 2x2 inverse) is a pure latency chain: one lane per filter, nothing
// to overlap it with inside that lane. CUDA's library versions are accurate but long (IEEE division fix-ups,
// Horner polynomials, slow-path calls); these keep ~1-2 ulp accuracy with the minimum number of DEPENDENT steps:
// MUFU seed + two Newton steps, Estrin polynomials. They are used only where the FAST mode is allowed to differ
// from the oracle's libm by rounding (never on a landmark's first touch, see ekf_fast.cuh).
#pragma once
#include <cuda_runtime.h>

namespace nuslam
{

// 1/x: MUFU.RCP64H seed (~2^-20) + 2 Newton steps -> <= 1 ulp for normal x
__device__ __forceinline__ double rcp_fast(double x)
{
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double e = fma(-x, y, 1.0);
    y = fma(y, e, y);
    e = fma(-x, y, 1.0);
    y = fma(y, e, y);
    return y;
}

// 1/sqrt(x): MUFU.RSQ64H seed + 2 Newton steps
__device__ __forceinline__ double rsqrt_fast(double x)
{
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double e = fma(-x * y, y, 1.0);
    y = fma(0.5 * y, e, y);
    e = fma(-x * y, y, 1.0);
    y = fma(0.5 * y, e, y);
    return y;
}

// atan2(y, x) for finite arguments, not both zero. One reciprocal, range reduction to |t| <= tan(pi/8) by
// t = (mn - mx) / (mn + mx) when mn/mx > tan(pi/8), degree-10 minimax-like polynomial in t^2 (Chebyshev-node
// fit, max relative error 7e-18 before rounding), Estrin evaluation.
__device__ __forceinline__ double atan2_fast(double y, double x)
{
    constexpr double kTanPi8 = 4.14213562373095034e-01;
    constexpr double kPi4Hi = 7.85398163397448279e-01, kPi4Lo = 3.06161699786838302e-17;
    constexpr double kPi2Hi = 1.57079632679489656e+00, kPi2Lo = 6.12323399573676604e-17;
    constexpr double kPiHi = 3.14159265358979312e+00, kPiLo = 1.22464679914735321e-16;
    constexpr double c0 = -3.33333333333333315e-01, c1 = 1.99999999999955214e-01, c2 = -1.42857142846665425e-01,
                     c3 = 1.11111110152563614e-01, c4 = -9.09090457812390257e-02, c5 = 7.69218319082608654e-02,
                     c6 = -6.66451144738194751e-02, c7 = 5.85814891280221003e-02, c8 = -5.08544973794025981e-02,
                     c9 = 3.92316582955871893e-02, c10 = -1.91768871190622602e-02;
    const double ax = fabs(x), ay = fabs(y);
    const double mx = fmax(ax, ay), mn = fmin(ax, ay);
    const bool big = mn > kTanPi8 * mx;
    const double num = big ? (mn - mx) : mn;
    const double den = big ? (mn + mx) : mx;
    const double r = rcp_fast(den);
    double t = num * r;
    t = fma(fma(-den, t, num), r, t);   // one correction: t = num/den to ~1 ulp
    const double u = t * t;
    const double u2 = u * u, u4 = u2 * u2, u8 = u4 * u4;
    const double p01 = fma(c1, u, c0), p23 = fma(c3, u, c2), p45 = fma(c5, u, c4), p67 = fma(c7, u, c6), p89 = fma(c9, u, c8);
    const double q0 = fma(p23, u2, p01), q1 = fma(p67, u2, p45), q2 = fma(c10, u2, p89);
    const double p = fma(q2, u8, fma(q1, u4, q0));
    double a = fma(t * u, p, t);                       // atan(t), |t| <= tan(pi/8)
    a = big ? (kPi4Hi + (a + kPi4Lo)) : a;             // atan(mn/mx) in [0, pi/4]
    a = (ay > ax) ? (kPi2Hi - (a - kPi2Lo)) : a;       // first octant swap
    a = (x < 0.0) ? (kPiHi - (a - kPiLo)) : a;         // left half-plane
    return (y < 0.0) ? -a : a;
}

}   // namespace nuslam
