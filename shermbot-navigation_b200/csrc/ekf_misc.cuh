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
                           int32_t * __restrict__ status, double prior)
{
    const int64_t t = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t len2 = (int64_t) len * len;
    if (t >= batch * len2) return;
    const int64_t b = t / len2;
    const int e = (int) (t - b * len2);
    const int j = e / len, i = e - j * len;
    sigma[t] = (i == j && i >= 3) ? prior : 0.0;   // INT_MAX by default (slam_library.cpp:30)
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

// The hand-over between the two nodes, kept on the device: the markers the landmarks node publishes for scan b
// (landmarks.cpp:84-109, already filtered and in detection order in `circles`) become filter b's measurements
// z_i = cartesian2polar(marker.pose.position.{x, y}) (slam.cpp:282-286). circles: B x max_circles x 4 (cx, cy, R, cluster);
// z: B x m x 2; m_valid[b] = min(markers of scan b, m) (0 where the reference's clusterPoints is undefined, n_circles < 0).
__global__ void k_markers_to_measurements(const double * __restrict__ circles, const int32_t * __restrict__ n_circles, int64_t batch,
                                          int max_circles, int m, double * __restrict__ z, int32_t * __restrict__ m_valid)
{
    const int64_t t = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= batch * m) return;
    const int64_t b = t / m;
    const int i = (int) (t - b * m);
    int mv = n_circles[b];
    mv = mv < 0 ? 0 : (mv > m ? m : mv);
    if (i == 0) m_valid[b] = mv;
    double r = 0.0, bearing = 0.0;
    if (i < mv && i < max_circles)
    {
        const double x = circles[(b * max_circles + i) * 4], y = circles[(b * max_circles + i) * 4 + 1];
        r = sqrt(add_(mul_(x, x), mul_(y, y)));
        bearing = normalize_angle(atan2(y, x));
    }
    z[2 * t] = r;
    z[2 * t + 1] = bearing;
}

// rigid2d::normalize_angle, rigid2d.cpp:9-13
__global__ void k_normalize_angle(const double * __restrict__ in, double * __restrict__ out, int64_t count)
{
    const int64_t t = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= count) return;
    out[t] = normalize_angle(in[t]);
}

// rigid2d::DiffDrive, batched: getTwist (diff_drive.cpp:80-110) followed by operator() (:111-146), exactly as the slam node turns
// wheel angles into the EKF control (nuslam/src/slam.cpp:264-265). state: B x 7 = {wheelBase, wheelRad, x, y, th, thL, thR}
// (updated in place); twists: B x 3 (dth, dx, dy = 0). integrateTwist: rigid2d.cpp:294-328, Transform2D product / inverse
// :187-214 in their operation order (unfused); sin / cos / atan from the CUDA math library.
struct Tf2D
{
    double c, s, x, y;
};
__device__ __forceinline__ Tf2D tf_mul(const Tf2D & l, const Tf2D & r)
{
    Tf2D o;
    o.c = sub_(mul_(l.c, r.c), mul_(l.s, r.s));
    o.s = add_(mul_(l.s, r.c), mul_(l.c, r.s));
    o.x = add_(sub_(mul_(l.c, r.x), mul_(l.s, r.y)), l.x);
    o.y = add_(add_(mul_(l.s, r.x), mul_(l.c, r.y)), l.y);
    return o;
}
__device__ __forceinline__ Tf2D tf_inv(const Tf2D & t)
{
    Tf2D o;
    o.c = t.c;
    o.s = -t.s;
    o.x = add_(mul_(-t.x, t.c), mul_(-t.y, t.s));
    o.y = add_(mul_(t.x, t.s), mul_(-t.y, t.c));
    return o;
}
__device__ __forceinline__ Tf2D integrate_twist(double dth, double dx, double dy)
{
    if (dth == 0.0) return Tf2D{1.0, 0.0, dx, dy};
    const Tf2D T_sb = {1.0, 0.0, div_(dy, dth), -div_(dx, dth)};
    double sn, cs;
    sincos(dth, &sn, &cs);
    const Tf2D T_ss = {cs, sn, 0.0, 0.0};
    return tf_mul(tf_mul(tf_inv(T_sb), T_ss), T_sb);   // operator* is left-associative (rigid2d.cpp:211-214,325)
}

// rigid2d::integrateTwist (rigid2d.cpp:294-328), batched: twists count x 3 (dth, dx, dy) -> transforms count x 4 (cos, sin, x, y)
__global__ void k_integrate_twist(const double * __restrict__ twists, double * __restrict__ out, int64_t count)
{
    const int64_t b = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= count) return;
    const Tf2D T = integrate_twist(twists[3 * b], twists[3 * b + 1], twists[3 * b + 2]);
    out[4 * b] = T.c;
    out[4 * b + 1] = T.s;
    out[4 * b + 2] = T.x;
    out[4 * b + 3] = T.y;
}

__global__ void k_diffdrive_step(double * __restrict__ state, const double * __restrict__ thL, const double * __restrict__ thR,
                                 double * __restrict__ twists, int64_t count)
{
    const int64_t b = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= count) return;
    double * s = state + 7 * b;
    const double wheelBase = s[0], wheelRad = s[1];
    const double dUL = sub_(thL[b], s[5]), dUR = sub_(thR[b], s[6]);
    const double dth = mul_(div_(wheelRad, wheelBase), sub_(dUR, dUL));
    const double dx = mul_(div_(wheelRad, 2.0), add_(dUL, dUR));
    twists[3 * b] = dth;
    twists[3 * b + 1] = dx;
    twists[3 * b + 2] = 0.0;
    const Tf2D Tbb = integrate_twist(dth, dx, 0.0);
    const double dqb_th = atan(div_(Tbb.s, Tbb.c));   // :129
    double sn, cs;
    sincos(s[4], &sn, &cs);
    // adj = Transform2D(th); adj(dqb) = rigid2d.cpp:254-261 with x = y = 0
    const double dq_x = sub_(add_(mul_(0.0, dqb_th), mul_(cs, Tbb.x)), mul_(sn, Tbb.y));
    const double dq_y = add_(add_(-mul_(0.0, dqb_th), mul_(sn, Tbb.x)), mul_(cs, Tbb.y));
    s[4] = add_(s[4], dqb_th);
    s[2] = add_(s[2], dq_x);
    s[3] = add_(s[3], dq_y);
    s[5] = thL[b];
    s[6] = thR[b];
}

// EKFSlam::broadcast_map2odom_tf (nuslam/src/slam.cpp:175-210), batched: T_mo = T_mb * T_ob.inv() with T_ob from the odometry model
// (state7 rows {.., x, y, th, ..}) and T_mb from the filter's state estimate (theta, x, y = x[0..2], row stride len); out B x 3 =
// (translation x, y, yaw = normalize_angle(asin(sin))). Transform2D(v, rad), inv, operator*: rigid2d.cpp:166-214, unfused.
__global__ void k_map_to_odom(const double * __restrict__ odom_state7, const double * __restrict__ x, int len, double * __restrict__ out,
                              int64_t count)
{
    const int64_t b = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= count) return;
    const double * o = odom_state7 + 7 * b;
    const double * e = x + (int64_t) len * b;
    Tf2D T_ob, T_mb;
    sincos(o[4], &T_ob.s, &T_ob.c);
    T_ob.x = o[2];
    T_ob.y = o[3];
    sincos(e[0], &T_mb.s, &T_mb.c);
    T_mb.x = e[1];
    T_mb.y = e[2];
    const Tf2D T_mo = tf_mul(T_mb, tf_inv(T_ob));
    out[3 * b] = T_mo.x;
    out[3 * b + 1] = T_mo.y;
    out[3 * b + 2] = normalize_angle(asin(T_mo.s));
}

// DiffDrive::convertTwist (diff_drive.cpp:66-78), batched: twists B x 3 -> wheel velocities B x 2 (uL, uR)
__global__ void k_diffdrive_convert_twist(double wheel_base, double wheel_rad, const double * __restrict__ twists, double * __restrict__ u, int64_t count)
{
    const int64_t b = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= count) return;
    const double d = div_(wheel_base, 2.0), r = wheel_rad;
    const double omg = twists[3 * b], vbx = twists[3 * b + 1];
    u[2 * b] = add_(mul_(-div_(d, r), omg), div_(vbx, r));
    u[2 * b + 1] = add_(mul_(div_(d, r), omg), div_(vbx, r));
}

// K6: error statistics of a Monte-Carlo batch, reduced on the device -- one thread per filter, warp-shuffle sums,
// one atomicAdd per warp and statistic. The shard's eight sums are what the ranks all-reduce over NCCL at the end of a run.
//   out[0] sum of squared robot position errors (m^2)      out[1] sum of squared heading errors (rad^2, wrapped)
//   out[2] sum of NEES = e^T Sigma_rr^-1 e (3 dof)         out[3] filters counted in [0..2] (finite state, regular Sigma_rr)
//   out[4] sum of squared landmark position errors (m^2)   out[5] landmarks counted in [4] (the first `seen` of every counted filter)
//   out[6] filters with a non-zero status                  out[7] association ids differing from the expected ones
constexpr int kStatsCount = 8;
__global__ void __launch_bounds__(256)
k_error_stats(const double * __restrict__ x, const double * __restrict__ sigma, const int32_t * __restrict__ seen, const int32_t * __restrict__ status,
              int64_t batch, int len, int n, const double * __restrict__ truth_pose, const double * __restrict__ truth_map,
              const int32_t * __restrict__ ids_got, const int32_t * __restrict__ ids_want, int m, double * __restrict__ out)
{
    const int64_t b = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    double v[kStatsCount] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    if (b < batch)
    {
        const double * xb = x + b * len;
        const double * S = sigma + b * (int64_t) len * len;   // column-major
        if (status[b] != 0) v[6] = 1.0;
        if (truth_pose)
        {
            const double eth = normalize_angle(xb[0] - truth_pose[3 * b]), ex = xb[1] - truth_pose[3 * b + 1], ey = xb[2] - truth_pose[3 * b + 2];
            // 3 x 3 robot block (order theta, x, y), inverse by cofactors
            const double a00 = S[0], a10 = S[1], a20 = S[2], a01 = S[len], a11 = S[len + 1], a21 = S[len + 2], a02 = S[2 * len], a12 = S[2 * len + 1],
                         a22 = S[2 * len + 2];
            const double c00 = a11 * a22 - a12 * a21, c01 = a12 * a20 - a10 * a22, c02 = a10 * a21 - a11 * a20;
            const double det = a00 * c00 + a01 * c01 + a02 * c02;
            // solve Sigma_rr w = e by Cramer's rule
            const double e0 = eth, e1 = ex, e2 = ey;
            const double w0 = (e0 * c00 + a01 * (a12 * e2 - e1 * a22) + a02 * (e1 * a21 - a11 * e2)) / det;
            const double w1 = (a00 * (e1 * a22 - a12 * e2) + e0 * c01 + a02 * (a10 * e2 - e1 * a20)) / det;
            const double w2 = (a00 * (a11 * e2 - e1 * a21) + a01 * (e1 * a20 - a10 * e2) + e0 * c02) / det;
            const double nees = e0 * w0 + e1 * w1 + e2 * w2;
            if (fabs(nees) < 1e300 && fabs(ex) < 1e300 && fabs(ey) < 1e300)
            {
                v[0] = ex * ex + ey * ey;
                v[1] = eth * eth;
                v[2] = nees;
                v[3] = 1.0;
                if (truth_map)
                {
                    const int ns = min(n, max(0, seen[b]));
                    for (int j = 0; j < ns; ++j)
                    {
                        const double lx = xb[3 + 2 * j] - truth_map[2 * j], ly = xb[4 + 2 * j] - truth_map[2 * j + 1];
                        const double e = lx * lx + ly * ly;
                        if (e < 1e300)
                        {
                            v[4] += e;
                            v[5] += 1.0;
                        }
                    }
                }
            }
        }
        if (ids_got && ids_want)
            for (int i = 0; i < m; ++i)
                if (ids_got[b * m + i] != ids_want[b * m + i]) v[7] += 1.0;
    }
#pragma unroll
    for (int k = 0; k < kStatsCount; ++k)
    {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], d);
        if ((threadIdx.x & 31) == 0 && v[k] != 0.0) atomicAdd(out + k, v[k]);
    }
}

}   // namespace nuslam
