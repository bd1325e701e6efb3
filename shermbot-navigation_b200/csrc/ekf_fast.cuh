// ekf_fast.cuh -- FAST arithmetic: fused predict + m sequential updates, one filter per warp, Sigma in REGISTERS.
//
// Roofs (MEASURED on B200, tools/ubench_fp64.cu, profiles/ubench_fp64_r01.txt): the fp64 pipe issues 64 FMA/clk/SM
// for DFMA and for DMMA alike (one pipe: mixing them adds nothing), a dependent DFMA takes 8.7 cycles, a
// broadcast LDS.128 occupies the shared-memory return path for ~2 cycles. A filter-step is ~21 k fp64 FMAs of
// rank-2 updates against 12 352 algorithmic bytes, so the fp64 pipe (~550 cycles/filter-step/SM) and HBM
// (~530 cycles/filter-step/SM at the measured 6.55 TB/s) are co-roofs; everything below is organised to keep the
// instruction count per update minimal and the per-update latency chain short.
//
// Layout of one filter inside its warp (state order [theta, x, y, m1x, m1y, ...], slam_library.cpp:46-59):
//   * landmark block Sigma(3.., 3..) (2N x 2N): fp64 tensor-core accumulator fragments of mma.m8n8k4 -- block
//     (br, bc), lane (g = lane / 4, t = lane % 4) holds Sigma(3 + 8 br + g, 3 + 8 bc + 2 t + {0, 1}); 18 doubles
//     per lane at N = 12 with no padding waste.
//   * robot rows / columns Sigma({th,x,y}, :) and Sigma(:, {th,x,y}) in VECTOR layout: lane i holds entry i of each
//     of the 6 vectors Rt, Rx, Ry (rows) and Ct, Cx, Cy (columns); the 3 x 3 robot block lives in both (updated
//     with bit-identical operations). predict (slam_library.cpp:65-108) touches only these vectors.
//   * the state x in vector layout too (lane i holds x_i).
// One update (slam_library.cpp:263-282), with H = D Ht, D = diag(1/sqrt d, 1/d), Ht = [0 -dx -dy dx dy; -d dy -dx
// -dy dx] free of divisions:   Sigma' = Sigma - Pt Minv Wt,  x' = x + Pt Minv (sqrt d dz0, d dz1),
//   Pt = Sigma Ht^T (lane i = row i), Wt = Ht Sigma (lane j = column j), M = Wt Ht^T + D^-1 R D^-1.
//   (A) the landmark's two rows and two columns are published from the fragments through shared memory into
//       vector layout; (B) lanes form Pt, Wt; the five lanes {th,x,y,c,c+1} hand Wt to the SCALAR WARP, which
//       evaluates M, Minv (one reciprocal), sqrt d, atan2 and the innovation for all filters of the CTA at once, one
//       filter per lane; (C) lanes form Kt = Pt Minv, update x and the 6 robot vectors with plain FMAs;
//   (D) updates are applied to the fragments LAZILY in chunks of CH = 2: the second update of a chunk takes its
//       landmark rows / columns from the stale fragments and corrects them in vector layout with the first update
//       (4 vectors x 2 FMAs), then ONE rank-4 DMMA pass (9 mma.m8n8k4 at N = 12, k = 4 fully used) applies both.
// CTA = 8 matrix warps + 1 scalar warp, 2 CTAs per SM, persistent over groups of 8 consecutive filters; the two CTAs
// of an SM run out of phase, so one's scalar phase hides under the other's matrix phase.
//
// Arithmetic: predict uses the oracle's operation order (it is O(len)). A filter-step that contains a landmark's
// FIRST TOUCH (INT_MAX prior, slam_library.cpp:28-31, where only the reference's own operation order reproduces its
// catastrophic cancellation) or an initializeLandmark is not evaluated here: the filter is
// appended to a work list that the STRICT kernel (ekf_strict.cuh) processes right after on the same stream.
// Reference: nuslam/src/slam_library.cpp:65-108 (predict), :263-282 (update), nuslam/src/slam.cpp:262-319.
#pragma once
#include "ekf_strict.cuh"
#include "fastmath.cuh"
#include <stdlib.h>

namespace nuslam
{

constexpr int kGroup = 8;                          // filters (= matrix warps) per CTA
constexpr int kFastThreads = 32 * (kGroup + 1);    // + the scalar warp
constexpr int kFastMMax = 16;                      // measurements per step handled by this kernel

template <int N>
struct FastGeom
{
    static constexpr int LEN = 3 + 2 * N;     // state length
    static constexpr int SIG = LEN * LEN;
    static constexpr int NB = (2 * N + 7) / 8;   // 8 x 8 fragment blocks per side of the landmark block
    static constexpr int TP = 8 * NB;         // padded landmark-block side
    static constexpr int VP = 3 + TP;         // padded vector length (state index space)
    static_assert(VP <= 32, "vector layout needs one lane per state index");
};

__device__ __forceinline__ void prefetch_l2_bulk(const void * src, uint32_t bytes)
{
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}

// D(8x8) += A(8x4) * B(4x8) on the fp64 tensor pipe; fragment layout in the header comment
__device__ __forceinline__ void dmma884(double & c0, double & c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// angle -> (-pi, pi]: what rigid2d::normalize_angle (rigid2d.cpp:9-13) returns, to ~1 ulp, without the
// sin/cos/atan2 round trip
__device__ __forceinline__ double wrap_angle(double a)
{
    constexpr double kPi = 3.14159265358979323846;
    constexpr double kTwoPiHi = 6.28318530717958623200, kTwoPiLo = 2.44929359829470641435e-16, kInvTwoPi = 0.15915494309189533577;
    if (a > kPi || a <= -kPi)
    {
        const double k = rint(a * kInvTwoPi);
        a = fma(-k, kTwoPiHi, a);
        a = fma(-k, kTwoPiLo, a);
        if (a > kPi) a -= kTwoPiHi;
        else if (a <= -kPi) a += kTwoPiHi;
    }
    return a;
}

// per-filter shared memory; the stride between filters is padded so that the scalar warp (lane = filter) reads the
// same field of 8 filters without bank conflicts
template <int N>
struct __align__(16) FastSmemFields
{
    using G = FastGeom<N>;
    static constexpr int PV = (G::VP + 2) & ~1;   // published row length (+1 shift so that index 3 is 16-byte aligned)
    static constexpr int KV = G::VP + 1;          // operand vector length (even)
    double2 kt[2][KV];        // -Kt of the chunk's two updates, kt[s][i] = (-k0, -k1) of state index i; DMMA A operand
    double2 wt[2][KV];        // Wt of the chunk's two updates; DMMA B operand
    double rho[2][2][PV];     // per chunk slot: landmark rows c, c+1 in vector layout, entry j at [j + 1]
    double2 kap[2][KV];       // per chunk slot: landmark columns (c, c+1) interleaved, entry i
    double xs[34];            // state broadcast copy, x_i at [i + 1] (pairs (x,y) and (mx,my) 16-byte aligned)
    double2 g[6];             // Wt at {th, x, y, c, c+1} and Pt at th, for the scalar warp
    double res[8];            // scalar warp results: Minv (4), y = Minv * scaled innovation (2), predict b10, b20
    double z[2 * kFastMMax];
    double tw[2];
    int ids[kFastMMax];
    int status;
};
template <int N>
struct __align__(16) FastSmem : FastSmemFields<N>
{
    static constexpr int kPad = (int) ((128 + 16 - sizeof(FastSmemFields<N>) % 128) % 128);
    unsigned char pad_[kPad == 0 ? 128 : kPad];   // filter stride = 16 (mod 128) bytes
};

// Scalar part of predict for one filter (one lane): predictEstimate :71-94 and the two Jacobian entries of
// getA :127-148 (theta read AFTER the motion update, :129), in the oracle's operation order.
template <int N>
__device__ __forceinline__ void predict_scalar(FastSmem<N> & f)
{
    const double dth = f.tw[0], dx = f.tw[1];
    const double theta = f.xs[1];
    double dq_th, dq_x, dq_y, s0, c0, b10, b20;
    sincos(theta, &s0, &c0);
    if (dth == 0.0)
    {
        dq_th = 0.0;
        dq_x = mul_(dx, c0);
        dq_y = mul_(dx, s0);
    }
    else
    {
        const double q = div_(dx, dth);
        double s1, c1;
        sincos(add_(theta, dth), &s1, &c1);
        dq_th = dth;
        dq_x = add_(mul_(-q, s0), mul_(q, s1));
        dq_y = sub_(mul_(q, c0), mul_(q, c1));
    }
    const double th1 = add_(theta, dq_th);
    f.xs[1] = th1;
    f.xs[2] = add_(f.xs[2], dq_x);
    f.xs[3] = add_(f.xs[3], dq_y);
    double s2, c2;
    sincos(th1, &s2, &c2);
    if (dth == 0.0)
    {
        b10 = mul_(-dx, s2);
        b20 = mul_(dx, c2);
    }
    else
    {
        const double q = div_(dx, dth);
        double s3, c3;
        sincos(add_(th1, dth), &s3, &c3);
        b10 = add_(mul_(-q, c2), mul_(q, c3));
        b20 = add_(mul_(-q, s2), mul_(q, s3));
    }
    f.res[6] = b10;
    f.res[7] = b20;
}

// Scalar part of one update for one filter (one lane): M = Wt Ht^T + D^-1 R D^-1, Minv, the innovation
// (slam_library.cpp:150-160 z_hat, :272 dz without wrap) scaled by D^-1, y = Minv * (sqrt d dz0, d dz1).
template <int N>
__device__ __forceinline__ void update_scalar(FastSmem<N> & f, int i, const double * R)
{
    const int id = f.ids[i];
    if (id < 1 || id > N)
    {
#pragma unroll
        for (int k = 0; k < 6; ++k) f.res[k] = 0.0;
        if (id > N) f.status |= kStatusBadId;
        return;
    }
    const int c = 3 + 2 * (id - 1);
    const double th = f.xs[1];
    const double dx = f.xs[c + 1] - f.xs[2], dy = f.xs[c + 2] - f.xs[3];
    const double d = fma(dx, dx, dy * dy);
    const double2 g0 = f.g[0], g1 = f.g[1], g2 = f.g[2], g3 = f.g[3], g4 = f.g[4];   // (Wt0, Wt1) at th, x, y, c, c+1
    const double e0 = g3.x - g1.x, f0 = g4.x - g2.x, e1 = g3.y - g1.y, f1 = g4.y - g2.y;
    const double s00 = fma(dx, e0, dy * f0), s01 = fma(dx, f0, fma(-dy, e0, -d * g0.x));
    const double s10 = fma(dx, e1, dy * f1), s11 = fma(dx, f1, fma(-dy, e1, -d * g0.y));
    const double rs = rsqrt_fast(d);
    double sq = d * rs;
    sq = fma(fma(-sq, sq, d), 0.5 * rs, sq);   // sqrt(d) to ~1 ulp
    const double dsq = d * sq;
    const double m00 = fma(d, R[0], s00), m10 = fma(dsq, R[1], s10), m01 = fma(dsq, R[2], s01), m11 = fma(d * d, R[3], s11);
    const double det = fma(m00, m11, -m01 * m10);
    const double idet = rcp_fast(det);
    const bool ok = (det != 0.0) && (fabs(idet) < 1.0e300) && (idet == idet);
    const double i00 = m11 * idet, i01 = -m01 * idet, i10 = -m10 * idet, i11 = m00 * idet;
    const double zb = wrap_angle(atan2_fast(dy, dx) - th);
    const double n0 = sq * (f.z[2 * i] - sq), n1 = d * (f.z[2 * i + 1] - zb);
    f.res[0] = ok ? i00 : 0.0;
    f.res[1] = ok ? i01 : 0.0;
    f.res[2] = ok ? i10 : 0.0;
    f.res[3] = ok ? i11 : 0.0;
    const double y0 = ok ? fma(i00, n0, i01 * n1) : 0.0, y1 = ok ? fma(i10, n0, i11 * n1) : 0.0;
    f.res[4] = y0;
    f.res[5] = y1;
    // theta' = normalize_angle(theta + K(0,:) dz) (slam_library.cpp:275-276); Pt(th,:) comes from lane 0 of the matrix warp
    const double2 pth = f.g[5];
    f.xs[1] = wrap_angle(fma(pth.x, y0, fma(pth.y, y1, th)));
    if (!ok) f.status |= kStatusSingular;   // arma::inv throws (slam_library.cpp:270); the update never happens
}

#ifdef NUSLAM_TIMING
#define NUSLAM_T(k) { const int probe_ = *reinterpret_cast<volatile int *>(&f.status); const long long now_ = clock64() + (probe_ & 0); tacc[k] += now_ - tlast; tlast = now_; }
__device__ long long g_fast_timing[16];
#else
#define NUSLAM_T(k)
#endif

template <int N>
__global__ void __launch_bounds__(kFastThreads, 2)
k_ekf_fast_step(const EkfParams p, const int do_predict, int32_t * __restrict__ worklist, int32_t * __restrict__ wl_count)
{
    using G = FastGeom<N>;
    using FS = FastSmem<N>;
    constexpr int LEN = G::LEN, SIG = G::SIG, NB = G::NB, VP = G::VP;
    constexpr unsigned kFull = 0xffffffffu;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    FS * fs = reinterpret_cast<FS *>(smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m = p.m;
    const int64_t ngroups = (p.batch + kGroup - 1) / kGroup;

    // ------------------------------------------------------------------ scalar warp: lane = filter of the group
    if (warp == kGroup)
    {
        FS & f = fs[lane < kGroup ? lane : 0];
        const bool mine = lane < kGroup;
        for (int64_t grp = blockIdx.x; grp < ngroups; grp += gridDim.x)
        {
            __syncthreads();   // B0: inputs of the group are in shared memory
            if (do_predict)
            {
                if (mine) predict_scalar<N>(f);
                __syncthreads();   // B1
            }
            for (int i = 0; i < m; ++i)
            {
                __syncthreads();   // Ba: Wt at the five H columns is published
#ifdef NUSLAM_TIMING
                const int probe0 = *reinterpret_cast<volatile int *>(&f.status);
                const long long s0 = clock64() + (probe0 & 0);
#endif
                if (mine) update_scalar<N>(f, i, p.R);
#ifdef NUSLAM_TIMING
                __syncwarp();
                if (lane == 0 && blockIdx.x == 0) atomicAdd((unsigned long long *) &g_fast_timing[8], (unsigned long long) (clock64() - s0));
#endif
                __syncthreads();   // Bb: Minv, y and the new theta are ready
            }
        }
        return;
    }

    // ------------------------------------------------------------------ matrix warps: one filter each
    FS & f = fs[warp];
    const int g = lane >> 2, t = lane & 3;
    const bool vlane = lane < LEN;      // lane owns a state index
    const bool vpad = lane < VP;        // lane owns a (possibly padded) vector slot
    const int lv = vpad ? lane : VP - 1;
    // zero the exchange buffers once (padding entries stay zero)
    for (int k = lane; k < (int) (sizeof(FS) / 8); k += 32) reinterpret_cast<double *>(&f)[k] = 0.0;
    __syncwarp();
#ifdef NUSLAM_TIMING
    long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long tlast = clock64();
#endif

    for (int64_t grp = blockIdx.x; grp < ngroups; grp += gridDim.x)
    {
        const int64_t bf = grp * kGroup + warp;
        const bool valid = bf < p.batch;
        const int64_t bl = valid ? bf : 0;
        // pull the Sigma of this CTA's next group towards L2 while the current one is computed
        {
            const int64_t nb = (grp + gridDim.x) * kGroup + warp;
            if (lane == 0 && nb < p.batch)
            {
                const uintptr_t a0 = reinterpret_cast<uintptr_t>(p.sigma + nb * SIG) & ~(uintptr_t) 15;
                prefetch_l2_bulk(reinterpret_cast<const void *>(a0), (uint32_t) ((sizeof(double) * SIG + 15) & ~15u));
            }
        }
        // ---- load: every global read of the group is issued before anything depends on one ----
        double C[NB][NB][2];
        double Rt = 0.0, Rx = 0.0, Ry = 0.0, Ct = 0.0, Cx = 0.0, Cy = 0.0, x = 0.0;
        const double * gs = p.sigma + bl * SIG;
#pragma unroll
        for (int br = 0; br < NB; ++br)
#pragma unroll
            for (int bc = 0; bc < NB; ++bc)
#pragma unroll
                for (int e = 0; e < 2; ++e)
                {
                    const int row = 3 + 8 * br + g, col = 3 + 8 * bc + 2 * t + e;
                    C[br][bc][e] = (row < LEN && col < LEN) ? __ldcs(gs + col * LEN + row) : 0.0;
                }
        if (vlane)
        {
            Ct = __ldcs(gs + lane);
            Cx = __ldcs(gs + LEN + lane);
            Cy = __ldcs(gs + 2 * LEN + lane);
            Rt = __ldcs(gs + lane * LEN);
            Rx = __ldcs(gs + lane * LEN + 1);
            Ry = __ldcs(gs + lane * LEN + 2);
            x = p.x[bl * LEN + lane];
        }
        const double diag = vlane ? gs[lane * (LEN + 1)] : 0.0;   // Sigma(lane, lane): first-touch detection
        const int st0 = p.status[bl], seen0 = p.seen[bl];
        const int my_id = (lane < m) ? p.ids[bl * m + lane] : 0;
        const double my_z = (lane < 2 * m) ? p.z[bl * m * 2 + lane] : 0.0;
        const double my_tw = (do_predict && lane < 2) ? p.twists[bl * 3 + lane] : 0.0;
        // ---- liveness ----
        bool dead = !valid || (st0 & (kStatusMapFull | kStatusSingular));   // the reference process died on an earlier scan
        {
            const bool idok = (unsigned) (my_id - 1) < (unsigned) N;
            const int c = idok ? 1 + 2 * my_id : 3;
            const double d0 = __shfl_sync(kFull, diag, c), d1 = __shfl_sync(kFull, diag, c + 1);
            // first touch (INT_MAX prior) or initializeLandmark (slam.cpp:295-297): the strict kernel takes this filter-step
            const bool need = idok && ((do_predict && my_id > seen0) || d0 > kFirstTouchVariance || d1 > kFirstTouchVariance);
            if (__any_sync(kFull, need) && !dead)
            {
                if (lane == 0) worklist[atomicAdd(wl_count, 1)] = (int32_t) bf;
                dead = true;
            }
        }
        if (dead)
        {
#pragma unroll
            for (int br = 0; br < NB; ++br)
#pragma unroll
                for (int bc = 0; bc < NB; ++bc) C[br][bc][0] = C[br][bc][1] = 0.0;
            Rt = Rx = Ry = Ct = Cx = Cy = x = 0.0;
        }
        if (lane < m) f.ids[lane] = dead ? 0 : my_id;
        if (lane < 2 * m) f.z[lane] = my_z;
        if (lane < 2) f.tw[lane] = dead ? 0.0 : my_tw;
        if (lane == 0) f.status = st0;
        f.xs[lane + 1] = x;
        NUSLAM_T(0)
        __syncthreads();   // B0

        // ---- predict (slam_library.cpp:65-108), oracle operation order, vector layout only ----
        if (do_predict)
        {
            __syncthreads();   // B1: the scalar warp has moved the pose and formed b10, b20
            const double b10 = f.res[6], b20 = f.res[7];
            if (lane < 3) x = f.xs[lane + 1];
            // T = A * Sigma: rows x, y += b * row theta
            Rx = add_(mul_(b10, Rt), Rx);
            Ry = add_(mul_(b20, Rt), Ry);
            {
                const double t0 = __shfl_sync(kFull, Ct, 0), t1 = __shfl_sync(kFull, Cx, 0), t2 = __shfl_sync(kFull, Cy, 0);
                const double bb = (lane == 1) ? b10 : b20;
                if (lane == 1 || lane == 2)
                {
                    Ct = add_(mul_(bb, t0), Ct);
                    Cx = add_(mul_(bb, t1), Cx);
                    Cy = add_(mul_(bb, t2), Cy);
                }
            }
            // U = T * A.t(): columns x, y += column theta * b
            Cx = add_(mul_(Ct, b10), Cx);
            Cy = add_(mul_(Ct, b20), Cy);
            {
                const double t0 = __shfl_sync(kFull, Rt, 0), t1 = __shfl_sync(kFull, Rx, 0), t2 = __shfl_sync(kFull, Ry, 0);
                const double bb = (lane == 1) ? b10 : b20;
                if (lane == 1 || lane == 2)
                {
                    Rt = add_(mul_(t0, bb), Rt);
                    Rx = add_(mul_(t1, bb), Rx);
                    Ry = add_(mul_(t2, bb), Ry);
                }
            }
            // + Q_bar on the robot block (expanded_process_noise :110-125); Q is column-major
            if (lane < 3 && !dead)
            {
                Rt = add_(Rt, p.Q[0 + 3 * lane]);
                Rx = add_(Rx, p.Q[1 + 3 * lane]);
                Ry = add_(Ry, p.Q[2 + 3 * lane]);
                Ct = add_(Ct, p.Q[lane + 3 * 0]);
                Cx = add_(Cx, p.Q[lane + 3 * 1]);
                Cy = add_(Cy, p.Q[lane + 3 * 2]);
            }
        }

        NUSLAM_T(1)
        // ---- m sequential updates in chunks of 2 (slam.cpp:279-319, known correspondence) ----
#pragma unroll 1
        for (int i0 = 0; i0 < m; i0 += 2)
        {
            // (A) publish the chunk's landmark rows / columns from the (stale) fragments into vector layout
            int cc[2];
#pragma unroll
            for (int s = 0; s < 2; ++s)
            {
                const int id = (i0 + s < m) ? f.ids[i0 + s] : 0;
                const bool live = (unsigned) (id - 1) < (unsigned) N;
                const int c = live ? 1 + 2 * id : 3;
                cc[s] = live ? c : -1;
                const int tau = c - 3;
                const int bsel = tau >> 3;
                const bool rsel = live && ((g >> 1) == ((tau & 7) >> 1));   // this lane holds row c or c+1
                const bool csel = live && (t == ((tau & 7) >> 1));          // this lane holds columns c, c+1
                double * const rdst = &f.rho[s][g & 1][4 + 2 * t];
                double2 * const cdst = &f.kap[s][3 + g];
#define NUSLAM_PUBLISH(b)                                                                                                  \
    if constexpr (NB > b)                                                                                                  \
    {                                                                                                                      \
        _Pragma("unroll") for (int q = 0; q < NB; ++q)                                                                     \
        {                                                                                                                  \
            if (rsel) *reinterpret_cast<double2 *>(rdst + 8 * q) = make_double2(C[b < NB ? b : 0][q][0], C[b < NB ? b : 0][q][1]); \
            if (csel) cdst[8 * q] = make_double2(C[q][b < NB ? b : 0][0], C[q][b < NB ? b : 0][1]);                        \
        }                                                                                                                  \
    }
                if (bsel == 0)
                {
                    NUSLAM_PUBLISH(0)
                }
                else if (bsel == 1)
                {
                    NUSLAM_PUBLISH(1)
                }
                else if (bsel == 2)
                {
                    NUSLAM_PUBLISH(2)
                }
                else
                {
                    NUSLAM_PUBLISH(3)
                }
#undef NUSLAM_PUBLISH
                const int e = lane - c;
                if (live && (e == 0 || e == 1))
                {
                    // robot part of row c+e: Sigma(c+e, {th,x,y}) is entry c+e of the column vectors; of column c+e: entry of the row vectors
                    f.rho[s][e][1] = Ct;
                    f.rho[s][e][2] = Cx;
                    f.rho[s][e][3] = Cy;
                    double * kd = reinterpret_cast<double *>(&f.kap[s][0]) + e;
                    kd[0] = Rt;
                    kd[2] = Rx;
                    kd[4] = Ry;
                }
            }
            __syncwarp();
            NUSLAM_T(2)
            double pW0 = 0.0, pW1 = 0.0, pK0 = 0.0, pK1 = 0.0;   // Wt and -Kt of the chunk's first update
#pragma unroll
            for (int s = 0; s < 2; ++s)
            {
                const int i = i0 + s;
                if (i < m)   // uniform over the CTA
                {
                    const bool live = cc[s] >= 0;
                    const int c = live ? cc[s] : 3;
                    // landmark rows c, c+1 (lane = column) and columns c, c+1 (lane = row)
                    double rho0 = f.rho[s][0][lv + 1], rho1 = f.rho[s][1][lv + 1];
                    const double2 kp = f.kap[s][lv];
                    double kap0 = kp.x, kap1 = kp.y;
                    if (s == 1)
                    {
                        // the fragments predate the chunk's first update: bring the four vectors up to date with it
                        const double2 ka = f.kt[0][c], kb = f.kt[0][c + 1], wa2 = f.wt[0][c], wb2 = f.wt[0][c + 1];
                        rho0 = fma(ka.x, pW0, fma(ka.y, pW1, rho0));
                        rho1 = fma(kb.x, pW0, fma(kb.y, pW1, rho1));
                        kap0 = fma(pK0, wa2.x, fma(pK1, wa2.y, kap0));
                        kap1 = fma(pK0, wb2.x, fma(pK1, wb2.y, kap1));
                    }
                    // (B) Pt (row role) and Wt (column role) of this lane
                    const double2 pxy = *reinterpret_cast<const double2 *>(&f.xs[2]);
                    const double2 mxy = *reinterpret_cast<const double2 *>(&f.xs[c + 1]);
                    const double dx = mxy.x - pxy.x, dy = mxy.y - pxy.y;
                    const double d = fma(dx, dx, dy * dy);
                    const double pa = kap0 - Cx, pb = kap1 - Cy;
                    const double wa = rho0 - Rx, wb = rho1 - Ry;
                    const bool on = live && vpad;
                    const double P0 = on ? fma(dx, pa, dy * pb) : 0.0, P1 = on ? fma(dx, pb, fma(-dy, pa, -d * Ct)) : 0.0;
                    const double W0 = on ? fma(dx, wa, dy * wb) : 0.0, W1 = on ? fma(dx, wb, fma(-dy, wa, -d * Rt)) : 0.0;
                    {
                        const int e = lane - c;
                        const int slot = (lane < 3) ? lane : ((e == 0 || e == 1) ? 3 + e : -1);
                        if (slot >= 0) f.g[slot] = make_double2(W0, W1);
                        if (lane == 0) f.g[5] = make_double2(P0, P1);
                    }
                    NUSLAM_T(3)
                    __syncthreads();   // Ba
                    __syncthreads();   // Bb
                    NUSLAM_T(4)
                    // (C) -Kt = -Pt Minv, x += Pt y
                    const double2 mi0 = *reinterpret_cast<const double2 *>(&f.res[0]);
                    const double2 mi1 = *reinterpret_cast<const double2 *>(&f.res[2]);
                    const double2 yy = *reinterpret_cast<const double2 *>(&f.res[4]);
                    const double th = f.xs[1];
                    const double nk0 = fma(-P0, mi0.x, -P1 * mi1.x), nk1 = fma(-P0, mi0.y, -P1 * mi1.y);
                    x = fma(P0, yy.x, fma(P1, yy.y, x));
                    x = (lane == 0) ? th : x;   // the scalar warp applied normalize_angle (slam_library.cpp:276)
                    if (vpad)
                    {
                        f.kt[s][lane] = make_double2(nk0, nk1);
                        f.wt[s][lane] = make_double2(W0, W1);
                        f.xs[lane + 1] = x;
                    }
                    __syncwarp();
                    // robot rows / columns: Sigma -= Kt Wt restricted to them
                    {
                        const double2 k0 = f.kt[s][0], k1 = f.kt[s][1], k2 = f.kt[s][2];
                        Rt = fma(k0.x, W0, fma(k0.y, W1, Rt));
                        Rx = fma(k1.x, W0, fma(k1.y, W1, Rx));
                        Ry = fma(k2.x, W0, fma(k2.y, W1, Ry));
                        const double2 w0 = f.wt[s][0], w1 = f.wt[s][1], w2 = f.wt[s][2];
                        Ct = fma(nk0, w0.x, fma(nk1, w0.y, Ct));
                        Cx = fma(nk0, w1.x, fma(nk1, w1.y, Cx));
                        Cy = fma(nk0, w2.x, fma(nk1, w2.y, Cy));
                    }
                    if (s == 0)
                    {
                        pW0 = W0;
                        pW1 = W1;
                        pK0 = nk0;
                        pK1 = nk1;
                    }
                    NUSLAM_T(5)
                }
                else
                {
                    // odd tail: the chunk's second slot contributes nothing to the rank-4 pass
                    if (vpad)
                    {
                        f.kt[s][lane] = make_double2(0.0, 0.0);
                        f.wt[s][lane] = make_double2(0.0, 0.0);
                    }
                    __syncwarp();
                }
            }
            // (D) one DMMA pass applies the chunk to the fragments: C += (-Kt) Wt, k = (u0, u1, v0, v1)
            {
                const double * ka = reinterpret_cast<const double *>(&f.kt[t >> 1][3 + g]) + (t & 1);
                const double * wa = reinterpret_cast<const double *>(&f.wt[t >> 1][3 + g]) + (t & 1);
                double a[NB], b[NB];
#pragma unroll
                for (int bb = 0; bb < NB; ++bb)
                {
                    a[bb] = ka[16 * bb];
                    b[bb] = wa[16 * bb];
                }
#pragma unroll
                for (int br = 0; br < NB; ++br)
#pragma unroll
                    for (int bc = 0; bc < NB; ++bc) dmma884(C[br][bc][0], C[br][bc][1], a[br], b[bc]);
            }
            __syncwarp();
            NUSLAM_T(6)
        }

        // ---- write back: registers -> HBM ----
        if (!dead)
        {
            double * gw = p.sigma + bf * SIG;
#pragma unroll
            for (int br = 0; br < NB; ++br)
#pragma unroll
                for (int bc = 0; bc < NB; ++bc)
#pragma unroll
                    for (int e = 0; e < 2; ++e)
                    {
                        const int row = 3 + 8 * br + g, col = 3 + 8 * bc + 2 * t + e;
                        if (row < LEN && col < LEN) __stcs(gw + col * LEN + row, C[br][bc][e]);
                    }
            if (vlane)
            {
                __stcs(gw + lane * LEN, Rt);
                __stcs(gw + lane * LEN + 1, Rx);
                __stcs(gw + lane * LEN + 2, Ry);
                if (lane >= 3)
                {
                    __stcs(gw + lane, Ct);
                    __stcs(gw + LEN + lane, Cx);
                    __stcs(gw + 2 * LEN + lane, Cy);
                }
                p.x[bf * LEN + lane] = x;
            }
            if (lane == 0 && f.status != st0) p.status[bf] = f.status;
        }
        NUSLAM_T(7)
    }
#ifdef NUSLAM_TIMING
    if (blockIdx.x == 0 && warp == 0 && lane == 0)
        for (int k = 0; k < 8; ++k) atomicAdd((unsigned long long *) &g_fast_timing[k], (unsigned long long) tacc[k]);
#endif
}

template <int N>
constexpr size_t fast_smem_bytes()
{
    return kGroup * sizeof(FastSmem<N>);
}

inline bool fast_supported(int n) { return n == 12 || n == 6; }

template <int N>
int launch_fast_n(const EkfParams & p, bool do_predict, int sm_count, int32_t * worklist, int32_t * wl_count, cudaStream_t stream)
{
    static_assert(sizeof(FastSmem<N>) % 128 == 16, "filter stride must be 16 (mod 128) bytes");
    static thread_local bool configured = false;
    constexpr size_t smem = fast_smem_bytes<N>();
    if (!configured)
    {
        cudaError_t e = cudaFuncSetAttribute(k_ekf_fast_step<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
        if (e != cudaSuccess) return (int) e;
        configured = true;
    }
    // persistent: two CTAs per SM, each looping over groups of kGroup consecutive filters
    int64_t blocks = (p.batch + kGroup - 1) / kGroup;
    int ctas_per_sm = 2;
#ifdef NUSLAM_TIMING
    if (const char * e = getenv("NUSLAM_FAST_CTAS_PER_SM")) ctas_per_sm = atoi(e);
#endif
    if (blocks > ctas_per_sm * (int64_t) sm_count) blocks = ctas_per_sm * (int64_t) sm_count;
    k_ekf_fast_step<N><<<(unsigned) blocks, kFastThreads, smem, stream>>>(p, do_predict ? 1 : 0, worklist, wl_count);
    return (int) cudaGetLastError();
}

// returns 0 on success, -1 when this configuration is not covered (caller falls back to the strict kernel), else a cudaError_t
inline int launch_fast(int n, const EkfParams & p, bool do_predict, int sm_count, int32_t * worklist, int32_t * wl_count, cudaStream_t stream)
{
    if (p.m > kFastMMax || p.m < 0 || p.ids == nullptr) return -1;
    if ((reinterpret_cast<uintptr_t>(p.sigma) & 7) || (reinterpret_cast<uintptr_t>(p.x) & 7)) return -1;
    if (n == 12) return launch_fast_n<12>(p, do_predict, sm_count, worklist, wl_count, stream);
    if (n == 6) return launch_fast_n<6>(p, do_predict, sm_count, worklist, wl_count, stream);
    return -1;
}

}   // namespace nuslam
