// ekf_misc.cuh -- small element-wise kernels behind the C ABI (constructor, getters of the measurement
// model, cartesian2polar, normalize_angle).
#pragma once
#include "ekf_common.cuh"

namespace nuslam
{

// ExtendedKalman::ExtendedKalman + initCov, slam_library.cpp:24-33,39-63: x = [robot, map], Sigma = 0 with
// INT_MAX on the landmark diagonal, seen = 0. One thread per Sigma element.
__global__ void k_ekf_init(int64_t batch, int len, const double * __restrict__ robot, const double * __restrict__ map,
                           double * __restrict__ x, double * __restrict__ sigma, int32_t * __restrict__ seen,
                           int32_t * __restrict__ status)
{
    const int64_t t = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t len2 = (int64_t) len * len;
    if (t >= batch * len2) return;
    const int64_t b = t / len2;
    const int e = (int) (t - b * len2);
    const int j = e / len, i = e - j * len;
    sigma[t] = (i == j && i >= 3) ? kLandmarkPrior : 0.0;
    if (j == 0)
    {
        double v;
        if (i < 3) v = robot[3 * b + i];
        else v = map ? map[(int64_t) (len - 3) * b + (i - 3)] : 0.0;
        x[b * len + i] = v;
        if (i == 0)
        {
            seen[b] = 0;
            status[b] = 0;
        }
    }
}

// computeTheoreticalMeasurement :150-160 and linearizedMeasurementModel :162-186 at the current state
__global__ void k_measurement_model(int64_t batch, int len, int n, const double * __restrict__ x, const int32_t * __restrict__ jj,
                                    double * __restrict__ zhat, double * __restrict__ Hout)
{
    const int64_t b = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= batch) return;
    const int j = jj[b];
    if (j < 1 || j > n) return;
    const int c = 3 + 2 * (j - 1);
    HEntries H;
    double zr, zb;
    measurement_model(x + b * len, c, H, zr, zb);
    if (zhat)
    {
        zhat[2 * b] = zr;
        zhat[2 * b + 1] = zb;
    }
    if (Hout)
    {
        double * Hb = Hout + b * 2 * (int64_t) len;   // 2 x len column-major
        for (int k = 0; k < 2 * len; ++k) Hb[k] = 0.0;
        Hb[1 + 2 * 0] = -1.0;
        Hb[0 + 2 * 1] = H.h01;
        Hb[1 + 2 * 1] = H.h11;
        Hb[0 + 2 * 2] = H.h02;
        Hb[1 + 2 * 2] = H.h12;
        Hb[0 + 2 * c] = H.h0c;
        Hb[1 + 2 * c] = H.h1c;
        Hb[0 + 2 * (c + 1)] = H.h0c1;
        Hb[1 + 2 * (c + 1)] = H.h1c1;
    }
}

// slam_library::cartesian2polar, slam_library.cpp:16-22
__global__ void k_cartesian2polar(const double * __restrict__ xy, double * __restrict__ rb, int64_t count)
{
    const int64_t t = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= count) return;
    const double x = xy[2 * t], y = xy[2 * t + 1];
    rb[2 * t] = sqrt(add_(mul_(x, x), mul_(y, y)));
    rb[2 * t + 1] = normalize_angle(atan2(y, x));
}

// rigid2d::normalize_angle, rigid2d.cpp:9-13
__global__ void k_normalize_angle(const double * __restrict__ in, double * __restrict__ out, int64_t count)
{
    const int64_t t = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= count) return;
    out[t] = normalize_angle(in[t]);
}

}   // namespace nuslam
