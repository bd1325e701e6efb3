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

// atan2(y, x) for finite arguments, not both zero, with ONE division. With mn = min(|x|, |y|), mx = max(|x|, |y|) and
// t_k = k / 64 the table point nearest to mn / mx (k from an fp32 estimate; any neighbouring k works),
//   atan(mn / mx) = atan(t_k) + atan(r),   r = (mn - t_k mx) / (mx + t_k mn),   |r| <= ~1 / 100,
// so a degree-7 odd polynomial (error r^9 / 9 < 1e-19) finishes it; atan(t_k) comes from a 65-entry (hi, lo) table and the
// octant (|y| > |x|, x < 0, y < 0) from an 8-entry table: atan2 = A + S (atan(t_k) + atan(r)) with A = k pi/4 (hi + lo).
struct AtanOctant
{
    double hi, lo, s, pad;
};
#define NUSLAM_Q(k) ((k) * 7.85398163397448279e-01), ((k) * 3.06161699786838302e-17)
// index = swap | xneg << 1 | yneg << 2
__constant__ AtanOctant kAtanOct[8] = {{NUSLAM_Q(0.0), 1.0, 0.0},   {NUSLAM_Q(2.0), -1.0, 0.0},  {NUSLAM_Q(4.0), -1.0, 0.0},  {NUSLAM_Q(2.0), 1.0, 0.0},
                                       {NUSLAM_Q(-0.0), -1.0, 0.0}, {NUSLAM_Q(-2.0), 1.0, 0.0},  {NUSLAM_Q(-4.0), 1.0, 0.0},  {NUSLAM_Q(-2.0), -1.0, 0.0}};
#undef NUSLAM_Q
__constant__ double2 kAtanTab[65] = {
    {0.0, 0.0},
    {0.015623728620476831, -4.913600136566304e-19},
    {0.031239833430268277, -1.188442711587748e-18},
    {0.046840712915969654, -1.655677442254952e-19},
    {0.06241880999595735, -1.5490756308295046e-18},
    {0.0779666338315423, 5.804551873143357e-18},
    {0.09347678115858947, -6.2844725995420954e-18},
    {0.10894195698986579, 6.8267122072409585e-18},
    {0.12435499454676144, -3.1253241424539383e-18},
    {0.13970887428916365, -2.9579864247315813e-18},
    {0.15499674192394097, 9.585415594114324e-18},
    {0.1702119252854744, -3.541164079802125e-18},
    {0.18534794999569476, 4.180692268843079e-18},
    {0.2003985538258785, 3.1399542871844493e-18},
    {0.21535769969773805, 4.738160130078733e-19},
    {0.23021958727684372, 1.2313404529142703e-17},
    {0.24497866312686414, 1.0698755618734451e-17},
    {0.2596296294082575, 1.9238754924615304e-17},
    {0.2741674511196588, 8.261353575163773e-18},
    {0.2885873618940774, -1.428369957377257e-17},
    {0.3028848683749714, -1.1010827903001369e-17},
    {0.31705575320914703, -1.893928924292642e-17},
    {0.3310960767041321, -7.952610375793799e-18},
    {0.34500217720710513, -2.2938804755578304e-17},
    {0.35877067027057225, -2.4623815582638635e-17},
    {0.3723984466767542, 1.9612311504845653e-17},
    {0.38588266939807375, 2.378822732491941e-17},
    {0.39922076957525254, 2.246598105617042e-17},
    {0.4124104415973873, -1.587652227770689e-17},
    {0.42544963737004227, 2.3315530741892885e-17},
    {0.43833655985795783, -2.494277030626541e-17},
    {0.4510696559885235, -2.2703795229420475e-17},
    {0.4636476090008061, 2.2698777452961687e-17},
    {0.4760693303227612, 1.4654487332256713e-17},
    {0.48833395105640554, -1.1373236189329585e-17},
    {0.5004408131472942, -4.7181675085518756e-17},
    {0.5123894603107377, -2.5462781472855804e-17},
    {0.5241796287829132, 5.520094119641666e-18},
    {0.5358112379604637, -4.0637956834825575e-18},
    {0.5472843809874369, 4.923709671396255e-17},
    {0.5585993153435624, -5.4556305485916264e-18},
    {0.5697564534829784, 1.2255062085054184e-17},
    {0.5807563535676704, -1.441464378193067e-17},
    {0.5915997103351114, 4.920495453686772e-17},
    {0.6022873461349642, 2.950430737228402e-17},
    {0.6128202021652414, -3.1552061848586226e-17},
    {0.6231993299340659, 2.672403885140095e-17},
    {0.6334258829691446, -2.7290767436015276e-17},
    {0.6435011087932844, 1.5834785051444286e-17},
    {0.6534263411807619, 3.5800634857340095e-17},
    {0.6632029927060933, -3.076054864429649e-17},
    {0.6728325475937632, -1.899315009714705e-17},
    {0.6823165548747481, 6.943223671560008e-18},
    {0.6916566218531999, -8.117151192285796e-18},
    {0.7008544078844502, -1.987626234335816e-17},
    {0.7099116184635249, -4.597166450584887e-17},
    {0.7188299996216245, -2.1478388444456983e-17},
    {0.7276113326265107, 2.569325697391839e-18},
    {0.7362574289814281, 3.473937648299457e-17},
    {0.7447701257160751, 3.708315849135547e-17},
    {0.7531512809621944, -2.4256934659182068e-17},
    {0.7614027698055784, 9.850030332752822e-18},
    {0.7695264804056583, -3.704991905602721e-17},
    {0.7775243103733478, -2.6676490951944502e-17},
    {0.7853981633974483, 3.061616997868383e-17}};

// fp64 literals of the per-update scalar chain, kept in the constant bank: an instruction reads them as a c[][] operand, where an
// immediate would be rebuilt in uniform registers (two UMOVs each) on every pass of the update loop
__constant__ double kFastK[8] = {-1.0 / 7.0, 0.2, -1.0 / 3.0, 6.28318530717958623200, 2.44929359829470641435e-16, 0.15915494309189533577,
                                 1.0e300, 3.14159265358979311600};

__device__ __forceinline__ double atan2_fast(double y, double x)
{
    const double ax = fabs(x), ay = fabs(y);
    const bool sw = ay > ax;
    const double mx = sw ? ay : ax, mn = sw ? ax : ay;
    const int idx = (sw ? 1 : 0) | (x < 0.0 ? 2 : 0) | (y < 0.0 ? 4 : 0);
    const AtanOctant oc = kAtanOct[idx];
    // table point: fp32 estimate of 64 mn / mx (the fp32 pipe is idle in this kernel)
    const float tf = __fdividef((float) mn, (float) mx);
    const int k = max(0, min(64, __float2int_rn(tf * 64.0f)));
    const double tk = (double) k * 0.015625;
    const double2 ak = kAtanTab[k];
    const double num = fma(-tk, mx, mn), den = fma(tk, mn, mx);
    const double rc = rcp_fast(den);
    double r = num * rc;
    r = fma(fma(-den, r, num), rc, r);   // r = num / den to ~1 ulp
    const double u = r * r;
    const double pl = fma(fma(fma(kFastK[0], u, kFastK[1]), u, kFastK[2]), u * r, r);   // atan(r)
    const double a = ak.x + (pl + ak.y);                                              // atan(mn / mx) in [0, pi/4]
    return oc.hi + fma(oc.s, a, oc.lo);
}

}   // namespace nuslam
