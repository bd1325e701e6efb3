// world_sim.cuh -- the simulator node's per-step arithmetic, batched: B independent simulated robots in one tube field
// (the step BEFORE the hot path, so that scan -> detect -> associate -> update never leaves the device).
//
// Reference: nuturtlesim/src/tube_world.cpp
//   :177-189  twist_callback        desired twist = commanded twist + gaussian noise      (the draws are inputs here)
//   :371-389  check_collision       slip along the tangent when closer than tube_rad + robot_rad
//   :512-529  main_loop body        wheel_vel = convertTwist; joints += wheel_vel dt; robot(joints + wheel_vel * slip draw)
//   :405-471  simulate_lidar_scanner 54 one-degree rays around each tube's bearing, ray / circle intersection, min per beam
// and rigid2d/src/diff_drive.cpp:66-78 (convertTwist), :111-146 (operator()), :154-159 (changeConfig).
//
// Two launches: the motion update with one thread per robot (it is a few dozen scalar operations with libm calls), then the scan
// with one warp per robot, 54 rays x T tubes spread over the lanes, reduced per beam with a shared-memory atomicMin on the float bit pattern (ranges are non-negative; the
// reference's `if (distance < ranges[ind]) ranges[ind] = distance` is order-independent: the result is float(min)), then written as
// one coalesced 1 440-byte row. Arithmetic is the reference's operation order, unfused; cos / sin of the integer-degree ray
// directions come from a host-libm table like the detector's (bit-identical to the oracle); atan2 / sincos of the pose from the
// CUDA math library (<= 2 ulp from glibc). HBM traffic per robot-step: 72 B + 24 B + 32 B in, 72 B + 16 B + 1 440 B out.
#pragma once
#include "ekf_misc.cuh"
#include <math.h>

namespace nuslam
{

constexpr int kWorldWarps = 4;          // robots per CTA
constexpr int kWorldMaxTubes = 64;
constexpr int kWorldRays = 54;          // tube_angle - 27 .. tube_angle + 26, tube_world.cpp:428
constexpr int kWorldDegMin = -180 - 27; // table covers every ray direction a finite pose can produce
constexpr int kWorldDegCount = 414;
constexpr double kWorldPi = 3.14159265358979323846;   // rigid2d.hpp:15

__device__ double g_world_cos[kWorldDegCount], g_world_sin[kWorldDegCount];

struct WorldParams
{
    int64_t count;
    double * world;          // count x 9 {wheelBase, wheelRad, x, y, th, thL, thR, jointL, jointR}
    const double * cmd;      // count x 3 (dth, dx, dy)
    const double * noise;    // count x 4 {twist dth, twist dx, slip L, slip R} or null
    const double * tubes;    // n_tubes x 2
    int n_tubes;
    double dt, tube_rad, robot_rad, max_range;
    float * ranges;          // count x 360
    double * joints;         // 2 x count (jointL[count], jointR[count]) or null: the encoder readings the odometry consumes
};

inline cudaError_t world_tables_init(int device)
{
    static bool done[64] = {false};
    if (device >= 0 && device < 64 && done[device]) return cudaSuccess;
    static double hc[kWorldDegCount], hs[kWorldDegCount];
    for (int k = 0; k < kWorldDegCount; ++k)
    {
        const double rad = (kWorldPi / (double) 180) * (double) (kWorldDegMin + k);   // rigid2d::deg2rad, rigid2d.hpp:40-44
        hc[k] = cos(rad);
        hs[k] = sin(rad);
    }
    cudaError_t e = cudaMemcpyToSymbol(g_world_cos, hc, sizeof(hc));
    if (e != cudaSuccess) return e;
    e = cudaMemcpyToSymbol(g_world_sin, hs, sizeof(hs));
    if (e != cudaSuccess) return e;
    if (device >= 0 && device < 64) done[device] = true;
    return cudaSuccess;
}

// world_ray's first decision on its own: true when the ray misses the tube for certain (world_ray then returns max_range + 1, the value
// the beams are initialised with: such a ray changes nothing). Same expressions, same values as in world_ray.
__device__ __forceinline__ bool world_ray_far(double x1, double y1, double c, double s, double tube_rad, double max_range)
{
    const double x2 = add_(x1, mul_(max_range, c));
    const double y2 = add_(y1, mul_(max_range, s));
    const double det = sub_(mul_(x1, y2), mul_(x2, y1));
    const double rr = mul_(tube_rad, tube_rad);
    return det * det > fma(rr * (max_range * max_range), 1.0 + 1e-9, 2e-5);
}

// one ray of simulate_lidar_scanner (tube_world.cpp:429-457): robot at (x1, y1) relative to the tube centre, direction (c, s)
__device__ __forceinline__ double world_ray(double x1, double y1, double c, double s, double tube_rad, double max_range)
{
    const double x2 = add_(x1, mul_(max_range, c));
    const double y2 = add_(y1, mul_(max_range, s));
    const double dx = sub_(x2, x1), dy = sub_(y2, y1);
    const double det = sub_(mul_(x1, y2), mul_(x2, y1));
    const double rr = mul_(tube_rad, tube_rad);
    // Most rays of the 54-degree window pass the tube far away. dr^2 equals max_range^2 to ~1e-15 (a unit direction times
    // max_range), so det^2 above r^2 max_range^2 with a 1e-9 margin plus the reference's 1e-5 tangency band means dis < -1e-5 for
    // certain: the reference's last branch, without the square root. Everything closer takes the reference's own expressions.
    if (det * det > fma(rr * (max_range * max_range), 1.0 + 1e-9, 2e-5)) return add_(max_range, 1.0);
    const double dr = sqrt(add_(mul_(dx, dx), mul_(dy, dy)));
    const double dr2 = mul_(dr, dr);
    const double dis = sub_(mul_(rr, dr2), mul_(det, det));
    if (fabs(dis) < 1e-5)
    {
        const double ix = div_(mul_(det, dy), dr2);
        const double iy = div_(-mul_(det, dx), dr2);
        const double ex = sub_(ix, x1), ey = sub_(iy, y1);
        return sqrt(add_(mul_(ex, ex), mul_(ey, ey)));
    }
    if (dis > 0)
    {
        const double root = sqrt(dis);
        // dy / fabs(dy) of the reference (:445): +-1 for a finite non-zero dy, NaN for a horizontal ray (0 / 0), an infinite or a NaN dy --
        // the same values without the IEEE division (18 % of this kernel's instructions were its five divisions per near ray)
        const double sg = (dy != 0.0 && fabs(dy) <= 1.7976931348623157e308) ? copysign(1.0, dy) : __longlong_as_double(0x7ff8000000000000LL);
        const double a = mul_(mul_(sg, dx), root), b = mul_(fabs(dy), root);
        const double ix1 = div_(add_(mul_(det, dy), a), dr2);
        const double iy1 = div_(add_(-mul_(det, dx), b), dr2);
        const double e1x = sub_(ix1, x1), e1y = sub_(iy1, y1);
        const double dist1 = sqrt(add_(mul_(e1x, e1x), mul_(e1y, e1y)));
        const double ix2 = div_(sub_(mul_(det, dy), a), dr2);
        const double iy2 = div_(sub_(-mul_(det, dx), b), dr2);
        const double e2x = sub_(ix2, x1), e2y = sub_(iy2, y1);
        const double dist2 = sqrt(add_(mul_(e2x, e2x), mul_(e2y, e2y)));
        return (dist2 < dist1) ? dist2 : dist1;   // std::min(dist1, dist2)
    }
    return add_(max_range, 1.0);
}

// stage 1: the motion update, one THREAD per robot (a few dozen scalar operations with libm calls: one lane's worth of work)
__global__ void __launch_bounds__(128) k_world_motion(const WorldParams p)
{
    const int64_t b = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= p.count) return;
    double * w = p.world + 9 * b;
    const double wheelBase = w[0], wheelRad = w[1];
    double x = w[2], y = w[3], th = w[4];
    const double thL = w[5], thR = w[6];
    double jL = w[7], jR = w[8];
    const double n_dth = p.noise ? p.noise[4 * b] : 0.0, n_dx = p.noise ? p.noise[4 * b + 1] : 0.0;
    const double slipL = p.noise ? p.noise[4 * b + 2] : 0.0, slipR = p.noise ? p.noise[4 * b + 3] : 0.0;
    const double tw_dth = add_(p.cmd[3 * b], n_dth), tw_dx = add_(p.cmd[3 * b + 1], n_dx);   // :181-183
    // check_collision :371-389 (sequential over the tubes: a slip changes the distance to the next one)
    for (int t = 0; t < p.n_tubes; ++t)
    {
        const double dx = sub_(p.tubes[2 * t], x), dy = sub_(p.tubes[2 * t + 1], y);
        const double dist = sqrt(add_(mul_(dx, dx), mul_(dy, dy)));
        if (dist <= add_(p.tube_rad, p.robot_rad))
        {
            x = add_(x, div_(div_(dy, dist), 50.0));
            y = add_(y, div_(div_(-dx, dist), 50.0));
        }
    }
    // convertTwist, diff_drive.cpp:66-78
    const double d = div_(wheelBase, 2.0), r = wheelRad;
    const double uL = add_(mul_(-div_(d, r), tw_dth), div_(tw_dx, r));
    const double uR = add_(mul_(div_(d, r), tw_dth), div_(tw_dx, r));
    jL = add_(jL, mul_(uL, p.dt));   // :522-523
    jR = add_(jR, mul_(uR, p.dt));
    // DiffDrive::operator()(jL + uL * slip, jR + uR * slip), diff_drive.cpp:111-146
    const double thLn = add_(jL, mul_(uL, slipL)), thRn = add_(jR, mul_(uR, slipR));
    {
        const double dUL = sub_(thLn, thL), dUR = sub_(thRn, thR);
        const double dth = mul_(div_(wheelRad, wheelBase), sub_(dUR, dUL));
        const double dxb = mul_(div_(wheelRad, 2.0), add_(dUL, dUR));
        const Tf2D Tbb = integrate_twist(dth, dxb, 0.0);
        const double dqb_th = atan(div_(Tbb.s, Tbb.c));
        double sn, cs;
        sincos(th, &sn, &cs);
        const double dq_x = sub_(add_(mul_(0.0, dqb_th), mul_(cs, Tbb.x)), mul_(sn, Tbb.y));
        const double dq_y = add_(add_(-mul_(0.0, dqb_th), mul_(sn, Tbb.x)), mul_(cs, Tbb.y));
        th = add_(th, dqb_th);
        x = add_(x, dq_x);
        y = add_(y, dq_y);
    }
    {
        w[2] = x;
        w[3] = y;
        w[4] = th;
        w[5] = thLn;
        w[6] = thRn;
        w[7] = jL;
        w[8] = jR;
        if (p.joints)
        {
            p.joints[b] = jL;
            p.joints[p.count + b] = jR;
        }
    }
}

// stage 2: simulate_lidar_scanner :405-471, one WARP per robot: 54 rays x T tubes spread over the lanes
__global__ void __launch_bounds__(32 * kWorldWarps, 8) k_world_scan(const WorldParams p)
{
    __shared__ __align__(16) int s_r[kWorldWarps][360];
    __shared__ int s_ta[kWorldWarps][kWorldMaxTubes];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t b = (int64_t) blockIdx.x * kWorldWarps + warp;
    if (b >= p.count) return;
    const double * w = p.world + 9 * b;
    const double x = w[2], y = w[3], th = w[4];   // the configuration after the motion update
    // simulate_lidar_scanner :405-471
    const float fill = (float) add_(p.max_range, 1.0);   // :416
    {
        const int fi = __float_as_int(fill);
        for (int k = lane; k < 90; k += 32) reinterpret_cast<int4 *>(s_r[warp])[k] = make_int4(fi, fi, fi, fi);   // 360 beams, four per store
    }
    for (int t = lane; t < p.n_tubes; t += 32)
    {
        const double xt = p.tubes[2 * t], yt = p.tubes[2 * t + 1];
        const double x1 = sub_(x, xt), y1 = sub_(y, yt);
        const double ang = round(mul_(180.0 / kWorldPi, atan2(sub_(yt, y1), sub_(xt, x1))));   // :426 (sic: relative coordinates)
        s_ta[warp][t] = (ang >= -180.0 && ang <= 180.0) ? (int) ang : 1000;                    // 1000: non-finite pose, no ray
    }
    __syncwarp();
    const int th_deg = (int) mul_(180.0 / kWorldPi, th);   // int(rad2deg(th)), :459
    // Two phases. Most of the 54 rays of a tube's window miss it for certain (world_ray_far: a dozen operations); the few that do not
    // take the reference's full expressions (five IEEE divisions, four square roots: ~200 instructions). Evaluated in place, almost every
    // group of 32 consecutive rays contains a near one and the whole warp walks the long path for a handful of lanes; so the near rays
    // are first compacted into a queue (ballot + popcount) and the long path runs on full warps of them.
    __shared__ unsigned short s_q[kWorldWarps][64];
    constexpr unsigned kFull = 0xffffffffu;
    const unsigned lt = (1u << lane) - 1u;
    auto ray = [&](const int item) {
        const int t = item / kWorldRays, k = item - t * kWorldRays;
        const int i = s_ta[warp][t] - 27 + k;
        const double x1 = sub_(x, p.tubes[2 * t]), y1 = sub_(y, p.tubes[2 * t + 1]);
        const double dist = world_ray(x1, y1, __ldg(&g_world_cos[i - kWorldDegMin]), __ldg(&g_world_sin[i - kWorldDegMin]), p.tube_rad, p.max_range);
        int ind = (i - th_deg) % 360;
        if (ind < 0) ind += 360;
        // `if (distance < ranges[ind]) ranges[ind] = distance` (:462-464): a NaN distance never stores; otherwise the minimum
        if (dist == dist) atomicMin(&s_r[warp][ind], __float_as_int((float) dist));
    };
    int nq = 0;
    // phase 1 tube by tube (warp-uniform tube: its centre and window are read once, no integer division per ray): rays k = lane and
    // lane + 32 of the tube's 54
    for (int t = 0; t < p.n_tubes; ++t)
    {
        const int ta = s_ta[warp][t];
        if (ta == 1000) continue;   // warp-uniform
        const double x1 = sub_(x, p.tubes[2 * t]), y1 = sub_(y, p.tubes[2 * t + 1]);
#pragma unroll
        for (int half = 0; half < 2; ++half)
        {
            const int k = lane + 32 * half;
            bool near = false;
            if (k < kWorldRays)
            {
                const int i = ta - 27 + k;
                near = !world_ray_far(x1, y1, __ldg(&g_world_cos[i - kWorldDegMin]), __ldg(&g_world_sin[i - kWorldDegMin]), p.tube_rad, p.max_range);
            }
            const unsigned mask = __ballot_sync(kFull, near);
            if (near) s_q[warp][nq + __popc(mask & lt)] = (unsigned short) (t * kWorldRays + k);
            nq += __popc(mask);
            __syncwarp();
            if (nq >= 32)
            {
                ray((int) s_q[warp][lane]);
                const int rest = nq - 32;
                const unsigned short moved = (lane < rest) ? s_q[warp][32 + lane] : (unsigned short) 0;
                __syncwarp();
                if (lane < rest) s_q[warp][lane] = moved;
                nq = rest;
                __syncwarp();
            }
        }
    }
    if (lane < nq) ray((int) s_q[warp][lane]);
    __syncwarp();
    float * out = p.ranges + 360 * b;
    if ((reinterpret_cast<uintptr_t>(p.ranges) & 15) == 0)   // a scan is 1 440 bytes: 16-byte aligned whenever the array is
        for (int k = lane; k < 90; k += 32) reinterpret_cast<int4 *>(out)[k] = reinterpret_cast<const int4 *>(s_r[warp])[k];
    else
        for (int k = lane; k < 360; k += 32) out[k] = __int_as_float(s_r[warp][k]);
}

inline cudaError_t launch_world_step(const WorldParams & p, int device, cudaStream_t stream)
{
    cudaError_t e = world_tables_init(device);
    if (e != cudaSuccess) return e;
    const int64_t blocks = (p.count + kWorldWarps - 1) / kWorldWarps;
    k_world_motion<<<(unsigned) ((p.count + 127) / 128), 128, 0, stream>>>(p);
    k_world_scan<<<(unsigned) blocks, 32 * kWorldWarps, 0, stream>>>(p);
    return cudaGetLastError();
}

}   // namespace nuslam
