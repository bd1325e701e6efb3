// ekf_fast.cuh -- FAST arithmetic: fused predict + m sequential updates with Sigma held in REGISTERS.
//
// Why registers: a rank-2 update reads and writes every element of Sigma once; from shared memory that is
// 16 B x len^2 per update (11.6 KB at len 27), i.e. ~1100 shared-memory cycles per filter-step at 128 B/clk/SM,
// above the ~900 cycles/filter-step/SM that 60 % of the HBM roofline allows. Registers have no such limit.
//
// Layout (one filter per half-warp, two filters per warp, 2W filters per CTA of W warps):
//   * internal index = external index + 1 (slot 0 is a zero dummy), so every landmark occupies an aligned
//     (even, odd) pair and the padded length LP = 4T is exactly 28 for n = 12 (16 for n = 6)
//   * the 16 lanes of a half-warp form a 4 x 4 grid (a = row group, b = column group); lane (a,b) holds the
//     cyclic T x T tile  S[r][q] = Sigma(4r + a, 4q + b)  in registers (49 doubles at n = 12). In the
//     column-major HBM image a lane quartet (a = 0..3) owns 32 contiguous bytes, so tile loads and stores
//     move whole 32-byte sectors; the next group's Sigma is pulled into L2 by a bulk prefetch
//     (cp.async.bulk.prefetch.L2) while the current one is computed.
//   * per update: (A) every warp publishes the 5 rows and 5 columns of Sigma that H touches (static register
//     indices through a switch on the landmark's tile column); (B) ONE warp evaluates the scalar part
//     (H, S = H Sigma H^T + R, S^-1, innovation; rsqrt / atan2 / reciprocal) for all 2W filters of the CTA, one
//     filter per lane, instead of every half-warp repeating it 16-fold; (C) the 16 lanes of each filter form
//     K = Sigma H^T S^-1 and W = H Sigma row by row; (D) the tile update Sigma -= K W is two FMAs per element
//     with K / W operands fetched as 16-byte pairs.
//
// Arithmetic: predict always uses the oracle's operation order (it is O(len)). An update whose landmark still
// carries the INT_MAX prior (slam_library.cpp:28-31) is evaluated in the STRICT operation order inside the
// same tile framework (see ekf_strict.cuh for the order); every other update uses the fused rank-2 form.
// Reference: nuslam/src/slam_library.cpp:65-108 (predict), :263-282 (update), nuslam/src/slam.cpp:262-319.
#pragma once
#include "ekf_strict.cuh"

namespace nuslam
{

template <int N>
struct FastGeom
{
    static constexpr int LEN = 3 + 2 * N;   // external state length
    static constexpr int LI = LEN + 1;      // internal length (dummy slot 0)
    static constexpr int T = (LI + 3) / 4;  // tile edge
    static constexpr int LP = 4 * T;        // padded internal length
    static constexpr int SIG = LEN * LEN;
    static constexpr int M_MAX = 16;        // measurements per step handled by this kernel
};

constexpr int kFastWarps = 4;   // warps per CTA (8 filters)

__device__ __forceinline__ void prefetch_l2_bulk(const void * src, uint32_t bytes)
{
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}

// wrap an angle known to lie within (-3pi, 3pi) into (-pi, pi]: what normalize_angle returns, to ~1 ulp,
// without the sin/cos/atan2 round trip (used only by the non-first-touch path)
__device__ __forceinline__ double wrap_fast(double a)
{
    constexpr double kPi = 3.14159265358979323846, kTwoPi = 6.28318530717958647692;
    if (a > kPi) a -= kTwoPi;
    else if (a <= -kPi) a += kTwoPi;
    return a;
}

// per-filter shared memory
template <int N>
struct FastSmem
{
    using G = FastGeom<N>;
    double2 kt[G::LP];        // K = Sigma H^T S^-1, one (k0,k1) pair per row
    double2 wt[G::LP];        // W = H Sigma, one (w0,w1) pair per column
    double xs[G::LP];         // state, internal indexing (xs[0] dummy)
    double col[5 * G::LP];    // published columns {th,x,y,c,c+1}: col[k*LP + i] = Sigma(i, col_k); M columns on the strict path
    double row[5 * G::LP];    // published rows    {th,x,y,c,c+1}: row[k*LP + j] = Sigma(row_k, j)
    double sc[16];            // scalar-phase outputs: H (8), S^-1 (4), innovation (2), predict b10, b20 (2)
    double z[2 * G::M_MAX];   // this step's measurements (range, bearing)
    double tw[2];             // this step's twist (dth, dx)
    int ids[G::M_MAX];        // this step's landmark ids
    int flags[8];             // [0] update flags, [1] seen, [2] status, [3] seen snapshot, [4] theta owes a wrap, [5] frozen
};

constexpr int kFlagSkip = 1, kFlagStrict = 2;

template <int T>
__device__ __forceinline__ void publish_col(double * dst, const double (&S)[T][T], int q)
{
    // dst[4*r] = S[r][q] with static register indices
    switch (q)
    {
#define NUSLAM_PC(k)                                                                        \
    case k:                                                                                 \
        if constexpr (T > k)                                                                \
        {                                                                                   \
            _Pragma("unroll") for (int r = 0; r < T; ++r) dst[4 * r] = S[r][k < T ? k : 0]; \
        }                                                                                   \
        break;
        NUSLAM_PC(0)
        NUSLAM_PC(1)
        NUSLAM_PC(2)
        NUSLAM_PC(3)
        NUSLAM_PC(4)
        NUSLAM_PC(5)
        NUSLAM_PC(6)
        NUSLAM_PC(7)
        NUSLAM_PC(8)
#undef NUSLAM_PC
    default: break;
    }
}

template <int T>
__device__ __forceinline__ void publish_row(double * dst, const double (&S)[T][T], int r)
{
    switch (r)
    {
#define NUSLAM_PR(k)                                                                        \
    case k:                                                                                 \
        if constexpr (T > k)                                                                \
        {                                                                                   \
            _Pragma("unroll") for (int q = 0; q < T; ++q) dst[4 * q] = S[k < T ? k : 0][q]; \
        }                                                                                   \
        break;
        NUSLAM_PR(0)
        NUSLAM_PR(1)
        NUSLAM_PR(2)
        NUSLAM_PR(3)
        NUSLAM_PR(4)
        NUSLAM_PR(5)
        NUSLAM_PR(6)
        NUSLAM_PR(7)
        NUSLAM_PR(8)
#undef NUSLAM_PR
    default: break;
    }
}

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
    f.sc[14] = b10;
    f.sc[15] = b20;
}

// Scalar part of one update for one filter (one lane): H, S = H Sigma H^T + R, S^-1, innovation.
// Reads the 5x5 block of Sigma from the published columns. slam_library.cpp:265-272 (+ :255-261 when the
// landmark is new).
template <int N>
__device__ __forceinline__ void scalar_phase(FastSmem<N> & f, double z0, double z1, int id, bool do_init, const double * R)
{
    using G = FastGeom<N>;
    constexpr int LP = G::LP;
    const int cI = 4 + 2 * (id - 1);   // internal column of the landmark's x
    double th = f.xs[1];
    if (f.flags[4])
    {
        th = wrap_fast(th);   // normalize_angle owed by the previous fused update (slam_library.cpp:276)
        f.xs[1] = th;
        f.flags[4] = 0;
    }
    const double px = f.xs[2], py = f.xs[3];
    const bool strict = (f.col[3 * LP + cI] > kFirstTouchVariance) || (f.col[4 * LP + cI + 1] > kFirstTouchVariance);
    if (do_init)
    {
        // initializeLandmark, slam_library.cpp:255-261
        double s, c;
        sincos(add_(z1, th), &s, &c);
        f.xs[cI] = add_(px, mul_(z0, c));
        f.xs[cI + 1] = add_(py, mul_(z0, s));
    }
    HEntries H;
    double zr, zb, i00, i01, i10, i11;
    bool ok = true;
    // rows of the 5x5 block: Sigma(row_k, col_j) = col[j*LP + row_k], rows {1,2,3,cI,cI+1}
    const double * c0 = f.col;
    if (strict)
    {
        measurement_model(f.xs + 1, cI - 1, H, zr, zb);
        double g0[5], g1[5];
#pragma unroll
        for (int j = 0; j < 5; ++j)
        {
            const double s0 = c0[j * LP + 1], s1 = c0[j * LP + 2], s2 = c0[j * LP + 3], s3 = c0[j * LP + cI], s4 = c0[j * LP + cI + 1];
            double a0 = mul_(H.h01, s1);
            a0 = add_(a0, mul_(H.h02, s2));
            a0 = add_(a0, mul_(H.h0c, s3));
            a0 = add_(a0, mul_(H.h0c1, s4));
            double a1 = -s0;
            a1 = add_(a1, mul_(H.h11, s1));
            a1 = add_(a1, mul_(H.h12, s2));
            a1 = add_(a1, mul_(H.h1c, s3));
            a1 = add_(a1, mul_(H.h1c1, s4));
            g0[j] = a0;
            g1[j] = a1;
        }
        double p00 = mul_(g0[1], H.h01);
        p00 = add_(p00, mul_(g0[2], H.h02));
        p00 = add_(p00, mul_(g0[3], H.h0c));
        p00 = add_(p00, mul_(g0[4], H.h0c1));
        double p10 = mul_(g1[1], H.h01);
        p10 = add_(p10, mul_(g1[2], H.h02));
        p10 = add_(p10, mul_(g1[3], H.h0c));
        p10 = add_(p10, mul_(g1[4], H.h0c1));
        double p01 = -g0[0];
        p01 = add_(p01, mul_(g0[1], H.h11));
        p01 = add_(p01, mul_(g0[2], H.h12));
        p01 = add_(p01, mul_(g0[3], H.h1c));
        p01 = add_(p01, mul_(g0[4], H.h1c1));
        double p11 = -g1[0];
        p11 = add_(p11, mul_(g1[1], H.h11));
        p11 = add_(p11, mul_(g1[2], H.h12));
        p11 = add_(p11, mul_(g1[3], H.h1c));
        p11 = add_(p11, mul_(g1[4], H.h1c1));
        p00 = add_(p00, R[0]);
        p10 = add_(p10, R[1]);
        p01 = add_(p01, R[2]);
        p11 = add_(p11, R[3]);
        ok = inv2x2(p00, p01, p10, p11, i00, i01, i10, i11);
    }
    else
    {
        const double dx = f.xs[cI] - px, dy = f.xs[cI + 1] - py;
        const double d = dx * dx + dy * dy;
        const double rs = rsqrt(d);
        const double id2 = rs * rs;
        H.h0c = dx * rs;
        H.h0c1 = dy * rs;
        H.h01 = -H.h0c;
        H.h02 = -H.h0c1;
        H.h11 = dy * id2;
        H.h12 = -dx * id2;
        H.h1c = -H.h11;
        H.h1c1 = dx * id2;
        zr = d * rs;
        zb = wrap_fast(atan2(dy, dx) - th);
        double p00 = R[0], p10 = R[1], p01 = R[2], p11 = R[3];
        // psi = (H Sigma) H^T + R accumulated column by column of the 5x5 block (keeps 10 values live, not 25)
#pragma unroll
        for (int j = 0; j < 5; ++j)
        {
            const double s0 = c0[j * LP + 1], s1 = c0[j * LP + 2], s2 = c0[j * LP + 3], s3 = c0[j * LP + cI], s4 = c0[j * LP + cI + 1];
            const double g0 = H.h01 * s1 + H.h02 * s2 + H.h0c * s3 + H.h0c1 * s4;
            const double g1 = -s0 + H.h11 * s1 + H.h12 * s2 + H.h1c * s3 + H.h1c1 * s4;
            // column j of H^T: H(0,j), H(1,j) for j in {th, x, y, c, c+1}
            const double hj0 = (j == 0) ? 0.0 : (j == 1) ? H.h01 : (j == 2) ? H.h02 : (j == 3) ? H.h0c : H.h0c1;
            const double hj1 = (j == 0) ? -1.0 : (j == 1) ? H.h11 : (j == 2) ? H.h12 : (j == 3) ? H.h1c : H.h1c1;
            p00 += g0 * hj0;
            p10 += g1 * hj0;
            p01 += g0 * hj1;
            p11 += g1 * hj1;
        }
        const double det = p00 * p11 - p01 * p10;
        ok = det != 0.0;
        const double idet = 1.0 / det;
        i00 = p11 * idet;
        i01 = -p01 * idet;
        i10 = -p10 * idet;
        i11 = p00 * idet;
    }
    f.sc[0] = H.h01;
    f.sc[1] = H.h02;
    f.sc[2] = H.h0c;
    f.sc[3] = H.h0c1;
    f.sc[4] = H.h11;
    f.sc[5] = H.h12;
    f.sc[6] = H.h1c;
    f.sc[7] = H.h1c1;
    f.sc[8] = i00;
    f.sc[9] = i01;
    f.sc[10] = i10;
    f.sc[11] = i11;
    f.sc[12] = sub_(z0, zr);   // :272, no wrap
    f.sc[13] = sub_(z1, zb);
    int fl = strict ? kFlagStrict : 0;
    if (!ok)
    {
        fl |= kFlagSkip;
        f.flags[2] |= kStatusSingular;
    }
    f.flags[0] = fl;
    if (fl == 0) f.flags[4] = 1;   // the fused path wraps theta lazily
}

template <int N>
__global__ void __launch_bounds__(kFastWarps * 32, 3) k_ekf_fast_step(const EkfParams p, const int do_predict)
{
    using G = FastGeom<N>;
    constexpr int T = G::T, LP = G::LP, LEN = G::LEN, SIG = G::SIG;
    constexpr int F = 2 * kFastWarps;   // filters per CTA
    extern __shared__ __align__(16) unsigned char smem_raw[];
    FastSmem<N> * fs = reinterpret_cast<FastSmem<N> *>(smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int hw = lane >> 4;          // which filter of the warp's pair
    const int t16 = lane & 15;
    const int a = t16 >> 2, b = t16 & 3;
    const int m = p.m;
    const int fl_idx = 2 * warp + hw;   // filter slot within the CTA
    FastSmem<N> & f = fs[fl_idx];

    const int64_t ngroups = (p.batch + F - 1) / F;
    for (int64_t group = blockIdx.x; group < ngroups; group += gridDim.x)
    {
        const int64_t bf = group * F + fl_idx;
        const bool valid = bf < p.batch;
        // pull the next group's Sigma towards L2 while this group is computed
        {
            const int64_t nb = (group + gridDim.x) * F + 2 * warp;
            if (lane == 0 && nb + 1 < p.batch) prefetch_l2_bulk(p.sigma + nb * SIG, (uint32_t) (sizeof(double) * 2 * SIG));
        }
        // ---- load: Sigma tile straight into registers (32-byte runs per lane quartet), small inputs into shared memory ----
        double S[T][T];
        {
            const double * gs = p.sigma + (valid ? bf : 0) * SIG;
#pragma unroll
            for (int q = 0; q < T; ++q)
#pragma unroll
                for (int r = 0; r < T; ++r)
                {
                    const int i = 4 * r + a - 1, j = 4 * q + b - 1;   // external indices
                    S[r][q] = (valid && i >= 0 && j >= 0 && i < LEN && j < LEN) ? __ldcs(gs + j * LEN + i) : 0.0;
                }
            for (int e = t16; e < LP; e += 16) f.xs[e] = (valid && e >= 1 && e <= LEN) ? p.x[bf * LEN + e - 1] : 0.0;
            for (int e = t16; e < 2 * m; e += 16) f.z[e] = valid ? p.z[bf * m * 2 + e] : 0.0;
            for (int e = t16; e < m; e += 16) f.ids[e] = valid ? p.ids[bf * m + e] : 0;
            if (t16 < 2) f.tw[t16] = (valid && do_predict) ? p.twists[bf * 3 + t16] : 0.0;
            if (t16 == 0)
            {
                const int seen = valid ? p.seen[bf] : 0;
                const int st = valid ? p.status[bf] : 0;
                f.flags[1] = seen;
                f.flags[2] = st;
                f.flags[3] = seen;   // snapshot, slam.cpp:251
                f.flags[4] = 0;
                f.flags[5] = (!valid || (st & (kStatusMapFull | kStatusSingular))) ? 1 : 0;
            }
        }
        __syncthreads();
        const bool frozen = f.flags[5] != 0;

        // ---- predict (slam_library.cpp:65-108), always in the oracle's operation order ----
        if (do_predict)
        {
            if (warp == 0 && lane < F && !fs[lane].flags[5]) predict_scalar<N>(fs[lane]);
            __syncthreads();
            const double b10 = f.sc[14], b20 = f.sc[15];
            // T = A * Sigma: rows x (a = 2, r = 0) and y (a = 3, r = 0) += b * row theta (a = 1, r = 0)
            const int src_row = hw * 16 + 4 + b;
#pragma unroll
            for (int q = 0; q < T; ++q)
            {
                const double thv = __shfl_sync(0xffffffffu, S[0][q], src_row);
                if (!frozen)
                {
                    if (a == 2) S[0][q] = add_(mul_(b10, thv), S[0][q]);
                    if (a == 3) S[0][q] = add_(mul_(b20, thv), S[0][q]);
                }
            }
            // U = T * A.t(): columns x (b = 2, q = 0) and y (b = 3, q = 0) += b * column theta (b = 1, q = 0)
            const int src_col = hw * 16 + 4 * a + 1;
#pragma unroll
            for (int r = 0; r < T; ++r)
            {
                const double t0 = __shfl_sync(0xffffffffu, S[r][0], src_col);
                if (!frozen)
                {
                    if (b == 2) S[r][0] = add_(mul_(t0, b10), S[r][0]);
                    if (b == 3) S[r][0] = add_(mul_(t0, b20), S[r][0]);
                }
            }
            // + Q_bar on the robot block (internal rows/cols 1..3)
            if (!frozen && a >= 1 && b >= 1) S[0][0] = add_(S[0][0], p.Q[(a - 1) + 3 * (b - 1)]);
        }

        // ---- m sequential updates (slam.cpp:279-319, known correspondence) ----
        for (int i = 0; i < m; ++i)
        {
            const int id = frozen ? 0 : f.ids[i];
            const bool live = id >= 1 && id <= N;
            const int cI = live ? 4 + 2 * (id - 1) : 4;
            const int cq = cI >> 2, cb = cI & 3;   // tile column and lane group (0 or 2) of the landmark pair
            // A. publish the 5 columns and 5 rows of Sigma that H touches
            if (b >= 1)
            {
#pragma unroll
                for (int r = 0; r < T; ++r) f.col[(b - 1) * LP + 4 * r + a] = S[r][0];
            }
            if (a >= 1)
            {
#pragma unroll
                for (int q = 0; q < T; ++q) f.row[(a - 1) * LP + 4 * q + b] = S[0][q];
            }
            if ((b >> 1) == (cb >> 1)) publish_col<T>(f.col + (3 + (b & 1)) * LP + a, S, cq);
            if ((a >> 1) == (cb >> 1)) publish_row<T>(f.row + (3 + (a & 1)) * LP + b, S, cq);
            __syncthreads();
            // B. scalar part of all 2W filters of the CTA, one filter per lane, on one warp (rotating)
            if (warp == (i & (kFastWarps - 1)) && lane < F)
            {
                FastSmem<N> & g = fs[lane];
                const int gid = g.flags[5] ? 0 : g.ids[i];
                if (gid >= 1 && gid <= N)
                {
                    const bool do_init = do_predict && gid > g.flags[3];   // slam.cpp:295 (step protocol only)
                    if (do_predict && gid > g.flags[1]) g.flags[1] = gid;    // what associateLandmark would have done to `seen`
                    scalar_phase<N>(g, g.z[2 * i], g.z[2 * i + 1], gid, do_init, p.R);
                }
                else
                {
                    g.flags[0] = kFlagSkip;
                    if (gid > N) g.flags[2] |= kStatusBadId;
                }
            }
            __syncthreads();
            const int fl = f.flags[0];
            const bool skip = fl & kFlagSkip, strict = fl & kFlagStrict;
            // C. K = Sigma H^T S^-1 and W = H Sigma, one row / column per lane (two passes of 16)
            if (!skip)
            {
                const double h01 = f.sc[0], h02 = f.sc[1], h0c = f.sc[2], h0c1 = f.sc[3];
                const double h11 = f.sc[4], h12 = f.sc[5], h1c = f.sc[6], h1c1 = f.sc[7];
                const double i00 = f.sc[8], i01 = f.sc[9], i10 = f.sc[10], i11 = f.sc[11];
                const double dz0 = f.sc[12], dz1 = f.sc[13];
#pragma unroll
                for (int h2 = 0; h2 < 2; ++h2)
                {
                    const int rr = t16 + 16 * h2;
                    if (rr < LP)
                    {
                        const double c0 = f.col[0 * LP + rr], c1 = f.col[1 * LP + rr], c2 = f.col[2 * LP + rr];
                        const double c3 = f.col[3 * LP + rr], c4 = f.col[4 * LP + rr];
                        if (!strict)
                        {
                            const double r0 = f.row[0 * LP + rr], r1 = f.row[1 * LP + rr], r2 = f.row[2 * LP + rr];
                            const double r3 = f.row[3 * LP + rr], r4 = f.row[4 * LP + rr];
                            const double p0 = h01 * c1 + h02 * c2 + h0c * c3 + h0c1 * c4;
                            const double p1 = -c0 + h11 * c1 + h12 * c2 + h1c * c3 + h1c1 * c4;
                            const double k0 = p0 * i00 + p1 * i10, k1 = p0 * i01 + p1 * i11;
                            f.xs[rr] += k0 * dz0 + k1 * dz1;
                            const double w0 = h01 * r1 + h02 * r2 + h0c * r3 + h0c1 * r4;
                            const double w1 = -r0 + h11 * r1 + h12 * r2 + h1c * r3 + h1c1 * r4;
                            f.kt[rr] = make_double2(k0, k1);
                            f.wt[rr] = make_double2(w0, w1);
                        }
                        else
                        {
                            // oracle order: P = Sigma*H.t(), K = P*inv(psi), x += K*dz, M = eye - K*H
                            double pa = mul_(c1, h01);
                            pa = add_(pa, mul_(c2, h02));
                            pa = add_(pa, mul_(c3, h0c));
                            pa = add_(pa, mul_(c4, h0c1));
                            double pb = -c0;
                            pb = add_(pb, mul_(c1, h11));
                            pb = add_(pb, mul_(c2, h12));
                            pb = add_(pb, mul_(c3, h1c));
                            pb = add_(pb, mul_(c4, h1c1));
                            const double k0 = add_(mul_(pa, i00), mul_(pb, i10));
                            const double k1 = add_(mul_(pa, i01), mul_(pb, i11));
                            f.xs[rr] = add_(f.xs[rr], add_(mul_(k0, dz0), mul_(k1, dz1)));
                            // M columns overwrite this lane's own entries of col[] (row[] keeps the old rows of Sigma)
                            f.col[0 * LP + rr] = sub_((rr == 1) ? 1.0 : 0.0, -k1);
                            f.col[1 * LP + rr] = sub_((rr == 2) ? 1.0 : 0.0, add_(mul_(k0, h01), mul_(k1, h11)));
                            f.col[2 * LP + rr] = sub_((rr == 3) ? 1.0 : 0.0, add_(mul_(k0, h02), mul_(k1, h12)));
                            f.col[3 * LP + rr] = sub_((rr == cI) ? 1.0 : 0.0, add_(mul_(k0, h0c), mul_(k1, h1c)));
                            f.col[4 * LP + rr] = sub_((rr == cI + 1) ? 1.0 : 0.0, add_(mul_(k0, h0c1), mul_(k1, h1c1)));
                        }
                    }
                }
            }
            __syncwarp();
            // D. tile update
            if (!skip)
            {
                if (!strict)
                {
                    double2 w[T];
#pragma unroll
                    for (int q = 0; q < T; ++q) w[q] = f.wt[4 * q + b];
#pragma unroll
                    for (int r = 0; r < T; ++r)
                    {
                        const double2 k = f.kt[4 * r + a];
#pragma unroll
                        for (int q = 0; q < T; ++q)
                        {
                            S[r][q] = fma(-k.x, w[q].x, S[r][q]);
                            S[r][q] = fma(-k.y, w[q].y, S[r][q]);
                        }
                    }
                }
                else
                {
                    // Sigma = M * Sigma in the oracle's ascending-k order (see ekf_strict.cuh)
#pragma unroll
                    for (int r = 0; r < T; ++r)
                    {
                        const int ii = 4 * r + a;
                        const double m0 = f.col[0 * LP + ii], m1 = f.col[1 * LP + ii], m2 = f.col[2 * LP + ii];
                        const double m3 = f.col[3 * LP + ii], m4 = f.col[4 * LP + ii];
#pragma unroll
                        for (int q = 0; q < T; ++q)
                        {
                            const int jj = 4 * q + b;
                            double acc = mul_(m0, f.row[0 * LP + jj]);
                            acc = add_(acc, mul_(m1, f.row[1 * LP + jj]));
                            acc = add_(acc, mul_(m2, f.row[2 * LP + jj]));
                            if (ii >= 4 && ii < cI) acc = add_(acc, S[r][q]);
                            acc = add_(acc, mul_(m3, f.row[3 * LP + jj]));
                            acc = add_(acc, mul_(m4, f.row[4 * LP + jj]));
                            if (ii > cI + 1) acc = add_(acc, S[r][q]);
                            S[r][q] = acc;
                        }
                    }
                    if (t16 == 0) f.xs[1] = normalize_angle(f.xs[1]);   // :276, exact chain on the strict path
                }
            }
            __syncwarp();
        }

        // ---- write back: registers -> HBM (each lane quartet writes 32 contiguous bytes) ----
        if (!frozen)
        {
            if (t16 == 0 && f.flags[4]) f.xs[1] = wrap_fast(f.xs[1]);
            __syncwarp();
            double * gs = p.sigma + bf * SIG;
#pragma unroll
            for (int q = 0; q < T; ++q)
#pragma unroll
                for (int r = 0; r < T; ++r)
                {
                    const int i = 4 * r + a - 1, j = 4 * q + b - 1;
                    if (i >= 0 && j >= 0 && i < LEN && j < LEN) __stcs(gs + j * LEN + i, S[r][q]);
                }
            for (int e = t16 + 1; e <= LEN; e += 16) p.x[bf * LEN + e - 1] = f.xs[e];
            if (t16 == 0)
            {
                p.seen[bf] = f.flags[1];
                p.status[bf] = f.flags[2];
            }
        }
        __syncthreads();   // shared-memory inputs of this group are dead; the next group may overwrite them
    }
}

template <int N>
constexpr size_t fast_smem_bytes()
{
    return 2 * kFastWarps * sizeof(FastSmem<N>);
}

inline bool fast_supported(int n) { return n == 12 || n == 6; }

template <int N>
int launch_fast_n(const EkfParams & p, bool do_predict, int sm_count, cudaStream_t stream)
{
    static thread_local bool configured = false;
    static thread_local int ctas_per_sm = 1;
    constexpr size_t smem = fast_smem_bytes<N>();
    if (!configured)
    {
        cudaError_t e = cudaFuncSetAttribute(k_ekf_fast_step<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
        if (e != cudaSuccess) return (int) e;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, k_ekf_fast_step<N>, kFastWarps * 32, smem);
        if (e != cudaSuccess) return (int) e;
        if (ctas_per_sm < 1) ctas_per_sm = 1;
        configured = true;
    }
    const int64_t ngroups = (p.batch + 2 * kFastWarps - 1) / (2 * kFastWarps);
    int64_t blocks = ngroups;
    const int64_t resident = (int64_t) sm_count * ctas_per_sm;
    if (blocks > resident) blocks = resident;   // persistent CTAs stride over groups of 2W filters
    k_ekf_fast_step<N><<<(unsigned) blocks, kFastWarps * 32, smem, stream>>>(p, do_predict ? 1 : 0);
    return (int) cudaGetLastError();
}

// returns 0 on success, -1 when this configuration is not covered (caller falls back to the strict kernel), else a cudaError_t
inline int launch_fast(int n, const EkfParams & p, bool do_predict, int sm_count, cudaStream_t stream)
{
    if (p.m > FastGeom<12>::M_MAX || p.m < 0 || p.ids == nullptr) return -1;
    if ((reinterpret_cast<uintptr_t>(p.sigma) & 15) || (reinterpret_cast<uintptr_t>(p.x) & 7)) return -1;
    if (n == 12) return launch_fast_n<12>(p, do_predict, sm_count, stream);
    if (n == 6) return launch_fast_n<6>(p, do_predict, sm_count, stream);
    return -1;
}

}   // namespace nuslam
