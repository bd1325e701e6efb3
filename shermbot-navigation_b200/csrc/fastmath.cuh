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
// t = (mn - mx) / (mn + mx) when mn/mx > tan(pi/8), degree-10 polynomial in t^2 (Chebyshev-node fit, max relative
// error 7e-18 before rounding) in Estrin form with the coefficients as constant-bank operands, and ONE table
// look-up for the octant: atan2 = A + S * atan(t) with A = k pi/4 (hi + lo) and S = +-1 chosen by
// (mn/mx > tan(pi/8), |y| > |x|, x < 0, y < 0).
__constant__ double kAtanC[11] = {-3.33333333333333315e-01, 1.99999999999955214e-01,  -1.42857142846665425e-01, 1.11111110152563614e-01,
                                  -9.09090457812390257e-02, 7.69218319082608654e-02,  -6.66451144738194751e-02, 5.85814891280221003e-02,
                                  -5.08544973794025981e-02, 3.92316582955871893e-02,  -1.91768871190622602e-02};
struct AtanOctant
{
    double hi, lo, s, pad;
};
#define NUSLAM_Q(k) ((k) * 7.85398163397448279e-01), ((k) * 3.06161699786838302e-17)
// index = big | swap << 1 | xneg << 2 | yneg << 3
__constant__ AtanOctant kAtanOct[16] = {
    {NUSLAM_Q(0.0), 1.0, 0.0},   {NUSLAM_Q(1.0), 1.0, 0.0},   {NUSLAM_Q(2.0), -1.0, 0.0},  {NUSLAM_Q(1.0), -1.0, 0.0},
    {NUSLAM_Q(4.0), -1.0, 0.0},  {NUSLAM_Q(3.0), -1.0, 0.0},  {NUSLAM_Q(2.0), 1.0, 0.0},   {NUSLAM_Q(3.0), 1.0, 0.0},
    {NUSLAM_Q(-0.0), -1.0, 0.0}, {NUSLAM_Q(-1.0), -1.0, 0.0}, {NUSLAM_Q(-2.0), 1.0, 0.0},  {NUSLAM_Q(-1.0), 1.0, 0.0},
    {NUSLAM_Q(-4.0), 1.0, 0.0},  {NUSLAM_Q(-3.0), 1.0, 0.0},  {NUSLAM_Q(-2.0), -1.0, 0.0}, {NUSLAM_Q(-3.0), -1.0, 0.0}};
#undef NUSLAM_Q

__device__ __forceinline__ double atan2_fast(double y, double x)
{
    constexpr double kTanPi8 = 4.14213562373095034e-01;
    const double ax = fabs(x), ay = fabs(y);
    const bool sw = ay > ax;
    const double mx = sw ? ay : ax, mn = sw ? ax : ay;
    const bool big = mn > kTanPi8 * mx;
    const int idx = (big ? 1 : 0) | (sw ? 2 : 0) | (x < 0.0 ? 4 : 0) | (y < 0.0 ? 8 : 0);
    const AtanOctant oc = kAtanOct[idx];
    const double num = big ? (mn - mx) : mn;
    const double den = big ? (mn + mx) : mx;
    const double r = rcp_fast(den);
    double t = num * r;
    t = fma(fma(-den, t, num), r, t);   // one correction: t = num/den to ~1 ulp
    const double u = t * t;
    const double u2 = u * u, u4 = u2 * u2, u8 = u4 * u4;
    const double p01 = fma(kAtanC[1], u, kAtanC[0]), p23 = fma(kAtanC[3], u, kAtanC[2]), p45 = fma(kAtanC[5], u, kAtanC[4]);
    const double p67 = fma(kAtanC[7], u, kAtanC[6]), p89 = fma(kAtanC[9], u, kAtanC[8]);
    const double q0 = fma(p23, u2, p01), q1 = fma(p67, u2, p45), q2 = fma(kAtanC[10], u2, p89);
    const double pp = fma(q2, u8, fma(q1, u4, q0));
    const double a = fma(t * u, pp, t);   // atan(t), |t| <= tan(pi/8)
    return oc.hi + fma(oc.s, a, oc.lo);
}

}   // namespace nuslam
