// ekf_fast.cuh -- FAST arithmetic: fused predict + m sequential updates with Sigma held in REGISTERS.
//
// Why registers: a rank-2 update reads and writes every element of Sigma once; from shared memory that is
// 16 B x len^2 per update (11.6 KB at len 27), i.e. ~1100 shared-memory cycles per filter-step at 128 B/clk/SM,
// above the ~700-880 cycles/filter-step/SM that 60 % of the HBM roofline allows. Registers have no such limit.
//
// Layout (one filter per half-warp, two filters per warp):
//   * internal index = external index + 1 (slot 0 is a zero dummy), so every landmark occupies an aligned
//     (even, odd) pair and the padded length LP = 4T is exactly 28 for n = 12 (16 for n = 6)
//   * the 16 lanes of a half-warp form a 4 x 4 grid (a = row group, b = column group); lane (a,b) holds the
//     cyclic T x T tile  S[r][q] = Sigma(4r + a, 4q + b)  in registers (49 doubles at n = 12)
//   * per update the 5 rows and 5 columns that H touches are published to shared memory (static register
//     indices through a switch on the landmark's tile column), the scalar part (H, S, S^-1, innovation) runs
//     once per filter, the 16 lanes form K = Sigma H^T S^-1 and W = H Sigma row by row, and the tile update
//     Sigma -= K W is two FMAs per element with K/W operands fetched as 16-byte pairs.
//   * Sigma travels HBM -> shared memory by TMA bulk copies (cp.async.bulk + mbarrier), prefetched one pair
//     ahead of the computation, and goes back from registers with full 32-byte-sector stores.
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

// ---- PTX helpers: mbarrier + 1-D TMA bulk copy + cp.async ----
__device__ __forceinline__ uint32_t smem_u32(const void * p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t * bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t * bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t * bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(void * dst, const void * src, uint32_t bytes, uint64_t * bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
                 "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void cp_async8(void * dst, const void * src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async4(void * dst, const void * src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// wrap an angle known to lie within (-3pi, 3pi) into (-pi, pi]: what normalize_angle returns, to ~1 ulp,
// without the sin/cos/atan2 round trip (used only by the non-first-touch path)
__device__ __forceinline__ double wrap_fast(double a)
{
    constexpr double kPi = 3.14159265358979323846, kTwoPi = 6.28318530717958647692;
    if (a > kPi) a -= kTwoPi;
    else if (a <= -kPi) a += kTwoPi;
    return a;
}

// per-filter shared memory (doubles)
template <int N>
struct FastSmem
{
    using G = FastGeom<N>;
    double xs[G::LP];         // state, internal indexing (xs[0] dummy)
    double col[5 * G::LP];    // published columns {th,x,y,c,c+1}: col[k*LP + i] = Sigma(i, col_k); later K pairs / M columns
    double row[5 * G::LP];    // published rows    {th,x,y,c,c+1}: row[k*LP + j] = Sigma(row_k, j); later W pairs
    double sc[16];            // scalar-phase outputs
    int flags[4];             // [0] update flags, [1] seen, [2] status, [3] seen snapshot
};

constexpr int kFlagSkip = 1, kFlagStrict = 2;

template <int T>
__device__ __forceinline__ void publish_col(double * dst, const double (&S)[T][T], int q)
{
    // dst[4*r] = S[r][q] with static register indices
    switch (q)
    {
#define NUSLAM_PC(k)                                            \
    case k:                                                     \
        if constexpr (T > k)                                    \
        {                                                       \
            _Pragma("unroll") for (int r = 0; r < T; ++r) dst[4 * r] = S[r][k < T ? k : 0]; \
        }                                                       \
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
#define NUSLAM_PR(k)                                            \
    case k:                                                     \
        if constexpr (T > k)                                    \
        {                                                       \
            _Pragma("unroll") for (int q = 0; q < T; ++q) dst[4 * q] = S[k < T ? k : 0][q]; \
        }                                                       \
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

// Scalar part of one update for one filter (runs in one lane): H, S = H Sigma H^T + R, S^-1, innovation.
// Reads the 5x5 block of Sigma from the published columns. slam_library.cpp:265-272 (+ :255-261 when the
// landmark is new).
template <int N>
__device__ __forceinline__ void scalar_phase(FastSmem<N> & f, double z0, double z1, int id, bool do_init, bool & pend, const double * R)
{
    using G = FastGeom<N>;
    constexpr int LP = G::LP;
    const int cI = 4 + 2 * (id - 1);   // internal column of the landmark's x
    double th = f.xs[1];
    if (pend)
    {
        th = wrap_fast(th);   // normalize_angle owed by the previous fused update (slam_library.cpp:276)
        f.xs[1] = th;
        pend = false;
    }
    const double px = f.xs[2], py = f.xs[3];
    const double vcc = f.col[3 * LP + cI], vc1 = f.col[4 * LP + cI + 1];
    const bool strict = (vcc > kFirstTouchVariance) || (vc1 > kFirstTouchVariance);
    if (do_init)
    {
        // initializeLandmark, slam_library.cpp:255-261
        double s, c;
        sincos(add_(z1, th), &s, &c);
        f.xs[cI] = add_(px, mul_(z0, c));
        f.xs[cI + 1] = add_(py, mul_(z0, s));
    }
    const int ri[5] = {1, 2, 3, cI, cI + 1};
    double S5[5][5];   // S5[k][j] = Sigma(row_k, col_j)
#pragma unroll
    for (int j = 0; j < 5; ++j)
#pragma unroll
        for (int k = 0; k < 5; ++k) S5[k][j] = f.col[j * LP + ri[k]];
    HEntries H;
    double zr, zb, i00, i01, i10, i11;
    bool ok = true;
    if (strict)
    {
        measurement_model(f.xs + 1, cI - 1, H, zr, zb);
        double g0[5], g1[5];
#pragma unroll
        for (int j = 0; j < 5; ++j)
        {
            double a0 = mul_(H.h01, S5[1][j]);
            a0 = add_(a0, mul_(H.h02, S5[2][j]));
            a0 = add_(a0, mul_(H.h0c, S5[3][j]));
            a0 = add_(a0, mul_(H.h0c1, S5[4][j]));
            double a1 = -S5[0][j];
            a1 = add_(a1, mul_(H.h11, S5[1][j]));
            a1 = add_(a1, mul_(H.h12, S5[2][j]));
            a1 = add_(a1, mul_(H.h1c, S5[3][j]));
            a1 = add_(a1, mul_(H.h1c1, S5[4][j]));
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
        double g0[5], g1[5];
#pragma unroll
        for (int j = 0; j < 5; ++j)
        {
            g0[j] = H.h01 * S5[1][j] + H.h02 * S5[2][j] + H.h0c * S5[3][j] + H.h0c1 * S5[4][j];
            g1[j] = -S5[0][j] + H.h11 * S5[1][j] + H.h12 * S5[2][j] + H.h1c * S5[3][j] + H.h1c1 * S5[4][j];
        }
        const double p00 = g0[1] * H.h01 + g0[2] * H.h02 + g0[3] * H.h0c + g0[4] * H.h0c1 + R[0];
        const double p10 = g1[1] * H.h01 + g1[2] * H.h02 + g1[3] * H.h0c + g1[4] * H.h0c1 + R[1];
        const double p01 = -g0[0] + g0[1] * H.h11 + g0[2] * H.h12 + g0[3] * H.h1c + g0[4] * H.h1c1 + R[2];
        const double p11 = -g1[0] + g1[1] * H.h11 + g1[2] * H.h12 + g1[3] * H.h1c + g1[4] * H.h1c1 + R[3];
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
    if (!(fl & (kFlagSkip | kFlagStrict))) pend = true;   // the fused path wraps theta lazily
}

template <int N>
__global__ void __launch_bounds__(kFastWarps * 32, 3) k_ekf_fast_step(const EkfParams p, const int do_predict)
{
    using G = FastGeom<N>;
    constexpr int T = G::T, LP = G::LP, LEN = G::LEN, SIG = G::SIG;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int hw = lane >> 4;          // which filter of the pair
    const int t16 = lane & 15;
    const int a = t16 >> 2, b = t16 & 3;
    const int m = p.m;

    // ---- carve shared memory: per warp [stage 2*SIG | xstage 2*LEN | z 2 x 2*M_MAX*2 | ids 2 x 2*M_MAX | tw 2 x 8 | FastSmem x2 | mbar]
    constexpr size_t kStageB = sizeof(double) * 2 * SIG;                         // 11664 (16-aligned)
    constexpr size_t kXStageB = (sizeof(double) * 2 * LEN + 15) / 16 * 16;       // 432
    constexpr size_t kZB = sizeof(double) * 2 * 2 * G::M_MAX * 2;                // double-buffered z of 2 filters
    constexpr size_t kIdB = sizeof(int) * 2 * 2 * G::M_MAX;
    constexpr size_t kTwB = sizeof(double) * 2 * 8;
    constexpr size_t kFB = (sizeof(FastSmem<N>) + 15) / 16 * 16;
    constexpr size_t kWarpB = kStageB + kXStageB + kZB + kIdB + kTwB + 2 * kFB + 16;
    unsigned char * wbase = smem_raw + (size_t) warp * kWarpB;
    double * stage = reinterpret_cast<double *>(wbase);
    double * xstage = reinterpret_cast<double *>(wbase + kStageB);
    double * zbuf = reinterpret_cast<double *>(wbase + kStageB + kXStageB);
    int * idbuf = reinterpret_cast<int *>(wbase + kStageB + kXStageB + kZB);
    double * twbuf = reinterpret_cast<double *>(wbase + kStageB + kXStageB + kZB + kIdB);
    FastSmem<N> * fs = reinterpret_cast<FastSmem<N> *>(wbase + kStageB + kXStageB + kZB + kIdB + kTwB);
    uint64_t * mbar = reinterpret_cast<uint64_t *>(wbase + kStageB + kXStageB + kZB + kIdB + kTwB + 2 * kFB);
    FastSmem<N> & f = fs[hw];

    const int64_t npairs = (p.batch + 1) / 2;
    const int64_t pair_stride = (int64_t) gridDim.x * kFastWarps;
    int64_t pair = (int64_t) blockIdx.x * kFastWarps + warp;
    if (pair >= npairs) return;   // warps are independent: no CTA-wide barrier anywhere below

    if (lane == 0)
    {
        mbar_init(mbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();

    // prefetch of one pair: Sigma + x by TMA bulk copy (full pairs), small inputs by cp.async
    auto prefetch = [&](int64_t pr, int buf)
    {
        const int64_t b0 = 2 * pr;
        const int nvalid = (int) ((p.batch - b0) < 2 ? (p.batch - b0) : 2);
        if (nvalid == 2)
        {
            if (lane == 0)
            {
                mbar_expect_tx(mbar, (uint32_t) (kStageB + sizeof(double) * 2 * LEN));
                bulk_g2s(stage, p.sigma + b0 * SIG, (uint32_t) kStageB, mbar);
                bulk_g2s(xstage, p.x + b0 * LEN, (uint32_t) (sizeof(double) * 2 * LEN), mbar);
            }
        }
        else
        {
            for (int e = lane; e < SIG; e += 32) cp_async8(stage + e, p.sigma + b0 * SIG + e);
            for (int e = lane; e < LEN; e += 32) cp_async8(xstage + e, p.x + b0 * LEN + e);
        }
        double * zb_ = zbuf + buf * (2 * G::M_MAX * 2);
        int * ib_ = idbuf + buf * (2 * G::M_MAX);
        double * tb_ = twbuf + buf * 8;
        for (int e = lane; e < nvalid * m * 2; e += 32)
        {
            const int fi = e / (2 * m), r = e - fi * 2 * m;
            cp_async8(zb_ + fi * (G::M_MAX * 2) + r, p.z + (b0 + fi) * m * 2 + r);
        }
        for (int e = lane; e < nvalid * m; e += 32)
        {
            const int fi = e / m, r = e - fi * m;
            cp_async4(ib_ + fi * G::M_MAX + r, p.ids + (b0 + fi) * m + r);
        }
        if (do_predict && lane < nvalid * 3) cp_async8(tb_ + (lane / 3) * 4 + (lane % 3), p.twists + b0 * 3 + lane);
    };

    uint32_t parity = 0;
    int buf = 0;
    prefetch(pair, buf);

    for (; pair < npairs; pair += pair_stride, buf ^= 1)
    {
        const int64_t b0 = 2 * pair;
        const int nvalid = (int) ((p.batch - b0) < 2 ? (p.batch - b0) : 2);
        const bool valid = hw < nvalid;
        const int64_t bf = b0 + (valid ? hw : 0);

        // ---- wait for the staged pair, build the register tile ----
        cp_async_wait_all();
        if (nvalid == 2)
        {
            mbar_wait(mbar, parity);
            parity ^= 1;
        }
        __syncwarp();
        double S[T][T];
        {
            const double * sg = stage + hw * SIG;
#pragma unroll
            for (int r = 0; r < T; ++r)
#pragma unroll
                for (int q = 0; q < T; ++q)
                {
                    const int i = 4 * r + a - 1, j = 4 * q + b - 1;   // external indices
                    S[r][q] = (valid && i >= 0 && j >= 0 && i < LEN && j < LEN) ? sg[j * LEN + i] : 0.0;
                }
            for (int e = t16; e < LP; e += 16) f.xs[e] = (valid && e >= 1 && e <= LEN) ? xstage[hw * LEN + e - 1] : 0.0;
            if (t16 == 0)
            {
                f.flags[1] = valid ? p.seen[bf] : 0;
                f.flags[2] = valid ? p.status[bf] : 0;
            }
        }
        __syncwarp();
        // staging is free again: prefetch the next pair while this one is computed
        const int64_t next = pair + pair_stride;
        if (next < npairs)
        {
            fence_proxy_async();
            prefetch(next, buf ^ 1);
        }
        const double * zb_ = zbuf + buf * (2 * G::M_MAX * 2) + hw * (G::M_MAX * 2);
        const int * ib_ = idbuf + buf * (2 * G::M_MAX) + hw * G::M_MAX;
        const double * tb_ = twbuf + buf * 8 + hw * 4;
        const bool frozen = !valid || (f.flags[2] & (kStatusMapFull | kStatusSingular));
        bool pend = false;   // lane t16 == 0: theta still owes a wrap into (-pi, pi]
        if (t16 == 0) f.flags[3] = f.flags[1];   // seen snapshot, slam.cpp:251
        __syncwarp();

        // ---- predict (slam_library.cpp:65-108), always in the oracle's operation order ----
        if (do_predict)
        {
            double b10 = 0.0, b20 = 0.0;
            if (t16 == 0 && !frozen)
            {
                const double dth = tb_[0], dx = tb_[1];
                const double theta = f.xs[1];
                double dq_th, dq_x, dq_y;
                double s0, c0;
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
            }
            b10 = __shfl_sync(0xffffffffu, b10, hw * 16);
            b20 = __shfl_sync(0xffffffffu, b20, hw * 16);
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
            __syncwarp();
        }

        // ---- m sequential updates (slam.cpp:279-319, known correspondence) ----
        for (int i = 0; i < m; ++i)
        {
            const int id = frozen ? 0 : ib_[i];
            const bool live = id >= 1 && id <= N;
            if (!frozen && id > N && t16 == 0) f.flags[2] |= kStatusBadId;
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
            __syncwarp();
            // B. scalar part, once per filter
            if (t16 == 0)
            {
                if (live)
                {
                    const bool do_init = do_predict && id > f.flags[3];   // slam.cpp:295 (step protocol only)
                    if (do_predict && id > f.flags[1]) f.flags[1] = id;    // what associateLandmark would have done to `seen`
                    scalar_phase<N>(f, zb_[2 * i], zb_[2 * i + 1], id, do_init, pend, p.R);
                }
                else
                    f.flags[0] = kFlagSkip;
            }
            __syncwarp();
            const int fl = f.flags[0];
            const bool skip = fl & kFlagSkip, strict = fl & kFlagStrict;
            // C. K = Sigma H^T S^-1 and W = H Sigma, one row / column per lane (two passes of 16)
            double cv[2][5], rv[2][5];
#pragma unroll
            for (int h2 = 0; h2 < 2; ++h2)
            {
                const int rr = t16 + 16 * h2;
                if (rr < LP)
                {
#pragma unroll
                    for (int k = 0; k < 5; ++k)
                    {
                        cv[h2][k] = f.col[k * LP + rr];
                        rv[h2][k] = f.row[k * LP + rr];
                    }
                }
            }
            const double h01 = f.sc[0], h02 = f.sc[1], h0c = f.sc[2], h0c1 = f.sc[3];
            const double h11 = f.sc[4], h12 = f.sc[5], h1c = f.sc[6], h1c1 = f.sc[7];
            const double i00 = f.sc[8], i01 = f.sc[9], i10 = f.sc[10], i11 = f.sc[11];
            const double dz0 = f.sc[12], dz1 = f.sc[13];
            __syncwarp();   // col/row are about to be overwritten by the K / W (or M) tables
            double2 * Ktab = reinterpret_cast<double2 *>(f.col);
            double2 * Wtab = reinterpret_cast<double2 *>(f.row);
            if (!skip)
            {
#pragma unroll
                for (int h2 = 0; h2 < 2; ++h2)
                {
                    const int rr = t16 + 16 * h2;
                    if (rr < LP)
                    {
                        if (!strict)
                        {
                            const double p0 = h01 * cv[h2][1] + h02 * cv[h2][2] + h0c * cv[h2][3] + h0c1 * cv[h2][4];
                            const double p1 = -cv[h2][0] + h11 * cv[h2][1] + h12 * cv[h2][2] + h1c * cv[h2][3] + h1c1 * cv[h2][4];
                            const double k0 = p0 * i00 + p1 * i10, k1 = p0 * i01 + p1 * i11;
                            f.xs[rr] += k0 * dz0 + k1 * dz1;
                            const double w0 = h01 * rv[h2][1] + h02 * rv[h2][2] + h0c * rv[h2][3] + h0c1 * rv[h2][4];
                            const double w1 = -rv[h2][0] + h11 * rv[h2][1] + h12 * rv[h2][2] + h1c * rv[h2][3] + h1c1 * rv[h2][4];
                            Ktab[rr] = make_double2(k0, k1);
                            Wtab[rr] = make_double2(w0, w1);
                        }
                        else
                        {
                            // oracle order: P = Sigma*H.t(), K = P*inv(psi), x += K*dz, M = eye - K*H
                            double pa = mul_(cv[h2][1], h01);
                            pa = add_(pa, mul_(cv[h2][2], h02));
                            pa = add_(pa, mul_(cv[h2][3], h0c));
                            pa = add_(pa, mul_(cv[h2][4], h0c1));
                            double pb = -cv[h2][0];
                            pb = add_(pb, mul_(cv[h2][1], h11));
                            pb = add_(pb, mul_(cv[h2][2], h12));
                            pb = add_(pb, mul_(cv[h2][3], h1c));
                            pb = add_(pb, mul_(cv[h2][4], h1c1));
                            const double k0 = add_(mul_(pa, i00), mul_(pb, i10));
                            const double k1 = add_(mul_(pa, i01), mul_(pb, i11));
                            f.xs[rr] = add_(f.xs[rr], add_(mul_(k0, dz0), mul_(k1, dz1)));
                            const double kh0 = -k1;
                            const double kh1 = add_(mul_(k0, h01), mul_(k1, h11));
                            const double kh2 = add_(mul_(k0, h02), mul_(k1, h12));
                            const double khc = add_(mul_(k0, h0c), mul_(k1, h1c));
                            const double khc1 = add_(mul_(k0, h0c1), mul_(k1, h1c1));
                            // M columns overwrite col[] (row[] keeps the old rows of Sigma)
                            f.col[0 * LP + rr] = sub_((rr == 1) ? 1.0 : 0.0, kh0);
                            f.col[1 * LP + rr] = sub_((rr == 2) ? 1.0 : 0.0, kh1);
                            f.col[2 * LP + rr] = sub_((rr == 3) ? 1.0 : 0.0, kh2);
                            f.col[3 * LP + rr] = sub_((rr == cI) ? 1.0 : 0.0, khc);
                            f.col[4 * LP + rr] = sub_((rr == cI + 1) ? 1.0 : 0.0, khc1);
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
                    for (int q = 0; q < T; ++q) w[q] = Wtab[4 * q + b];
#pragma unroll
                    for (int r = 0; r < T; ++r)
                    {
                        const double2 k = Ktab[4 * r + a];
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

        if (t16 == 0 && pend) f.xs[1] = wrap_fast(f.xs[1]);
        __syncwarp();
        // ---- write back: registers -> HBM (each lane quartet writes 32 contiguous bytes) ----
        if (valid && !frozen)
        {
            double * gs = p.sigma + bf * SIG;
#pragma unroll
            for (int r = 0; r < T; ++r)
#pragma unroll
                for (int q = 0; q < T; ++q)
                {
                    const int i = 4 * r + a - 1, j = 4 * q + b - 1;
                    if (i >= 0 && j >= 0 && i < LEN && j < LEN) gs[j * LEN + i] = S[r][q];
                }
            for (int e = t16 + 1; e <= LEN; e += 16) p.x[bf * LEN + e - 1] = f.xs[e];
            if (t16 == 0)
            {
                p.seen[bf] = f.flags[1];
                p.status[bf] = f.flags[2];
            }
        }
        __syncwarp();
    }
}

template <int N>
constexpr size_t fast_smem_bytes()
{
    using G = FastGeom<N>;
    return kFastWarps * (sizeof(double) * 2 * G::SIG + (sizeof(double) * 2 * G::LEN + 15) / 16 * 16 + sizeof(double) * 2 * 2 * G::M_MAX * 2 +
                         sizeof(int) * 2 * 2 * G::M_MAX + sizeof(double) * 2 * 8 + 2 * ((sizeof(FastSmem<N>) + 15) / 16 * 16) + 16);
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
    const int64_t npairs = (p.batch + 1) / 2;
    int64_t blocks = (npairs + kFastWarps - 1) / kFastWarps;
    const int64_t resident = (int64_t) sm_count * ctas_per_sm;
    if (blocks > resident) blocks = resident;   // persistent: each warp strides over pairs, prefetching one ahead
    k_ekf_fast_step<N><<<(unsigned) blocks, kFastWarps * 32, smem, stream>>>(p, do_predict ? 1 : 0);
    return (int) cudaGetLastError();
}

// returns 0 on success, -1 when this configuration is not covered (caller falls back to the strict kernel), else a cudaError_t
inline int launch_fast(int n, const EkfParams & p, bool do_predict, int sm_count, cudaStream_t stream)
{
    if (p.m > FastGeom<12>::M_MAX || p.m < 0 || p.ids == nullptr) return -1;
    if ((reinterpret_cast<uintptr_t>(p.sigma) & 15) || (reinterpret_cast<uintptr_t>(p.x) & 15)) return -1;
    if ((reinterpret_cast<uintptr_t>(p.z) & 7) || (reinterpret_cast<uintptr_t>(p.ids) & 3)) return -1;
    if (n == 12) return launch_fast_n<12>(p, do_predict, sm_count, stream);
    if (n == 6) return launch_fast_n<6>(p, do_predict, sm_count, stream);
    return -1;
}

}   // namespace nuslam
