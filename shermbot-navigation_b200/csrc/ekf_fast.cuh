// ekf_fast.cuh -- FAST arithmetic: fused predict + m sequential updates with Sigma held in REGISTERS.
//
// Why registers: a rank-2 update reads and writes every element of Sigma once; from shared memory that is
// 16 B x len^2 per update (11.6 KB at len 27), i.e. ~1100 shared-memory cycles per filter-step at 128 B/clk/SM,
// above the ~900 cycles/filter-step/SM that 60 % of the HBM roofline allows. Registers have no such limit.
//
// Layout (one filter per warp, W filters per CTA of W warps):
//   * internal index = external index + 1 (slot 0 is a zero dummy), so every landmark occupies an aligned
//     (even, odd) pair; padded sizes 28 rows x 32 columns at n = 12 (16 x 16 at n = 6)
//   * the 32 lanes form a 4 x 8 grid (a = lane / 8 row group, b = lane % 8 column group); lane (a,b) holds the
//     cyclic TR x TC tile  S[r][q] = Sigma(4r + a, 8q + b)  in registers (28 doubles at n = 12), which leaves
//     room for ~18 resident warps per SM. In the column-major HBM image a lane quartet (a = 0..3) owns 32
//     contiguous bytes, so tile loads and stores move whole 32-byte sectors; the next group's Sigma is pulled
//     into L2 by a bulk prefetch (cp.async.bulk.prefetch.L2) while the current one is computed.
//   * per update: (A) every warp publishes the 5 rows and 5 columns of Sigma that H touches (static register
//     indices through one switch on the landmark's tile row); (B) ONE warp evaluates the scalar part
//     (H, S = H Sigma H^T + R, S^-1, innovation; rsqrt / atan2 / reciprocal) for all W filters of the CTA, one
//     filter per lane, instead of every warp repeating it 32-fold; (C) lane i forms row i of
//     K = Sigma H^T S^-1 and column i of W = H Sigma; (D) the tile update Sigma -= K W is two FMAs per element
//     with K / W operands fetched as 16-byte pairs.
//
// Arithmetic: predict always uses the oracle's operation order (it is O(len)). An update whose landmark still
// carries the INT_MAX prior (slam_library.cpp:28-31) is evaluated in the STRICT operation order inside the
// same tile framework (see ekf_strict.cuh for the order); every other update uses the fused rank-2 form.
// Reference: nuslam/src/slam_library.cpp:65-108 (predict), :263-282 (update), nuslam/src/slam.cpp:262-319.
#pragma once
#include "ekf_strict.cuh"
#include "fastmath.cuh"

namespace nuslam
{

template <int N>
struct FastGeom
{
    static constexpr int LEN = 3 + 2 * N;   // external state length
    static constexpr int LI = LEN + 1;      // internal length (dummy slot 0)
    static constexpr int TR = (LI + 3) / 4; // tile rows per lane
    static constexpr int TC = (LI + 7) / 8; // tile columns per lane
    static constexpr int LPR = 4 * TR;      // padded rows
    static constexpr int LPC = 8 * TC;      // padded columns
    static constexpr int SIG = LEN * LEN;
    static constexpr int M_MAX = 16;        // measurements per step handled by this kernel
};

#ifndef NUSLAM_MATRIX_WARPS
#define NUSLAM_MATRIX_WARPS 16
#endif
#ifndef NUSLAM_SCALAR_WARPS
#define NUSLAM_SCALAR_WARPS 2
#endif
constexpr int kMatrixWarps = NUSLAM_MATRIX_WARPS;   // one filter in flight per matrix warp
constexpr int kScalarWarps = NUSLAM_SCALAR_WARPS;   // scalar-server warps: lane l of server s serves matrix warp l * kScalarWarps + s
constexpr int kFastThreads = 32 * (kMatrixWarps + kScalarWarps);
static_assert(kMatrixWarps <= 32 * kScalarWarps, "a scalar server has 32 lanes");

__device__ __forceinline__ void prefetch_l2_bulk(const void * src, uint32_t bytes)
{
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}

__device__ __forceinline__ uint32_t smem_u32(const void * p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t * bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t * bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// hardware-suspended wait (try_wait sleeps in the barrier unit, it does not spin on the issue port)
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
__device__ __forceinline__ int ld_acquire_smem(const int * p)
{
    int v;
    asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_smem(int * p, int v)
{
    asm volatile("st.release.cta.shared.u32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory");
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
    double2 kt[G::LPR];       // K = Sigma H^T S^-1, one (k0,k1) pair per row
    double2 wt[G::LPC];       // W = H Sigma, one (w0,w1) pair per column
    double xs[G::LPC];        // state, internal indexing (xs[0] dummy)
    double col[5 * G::LPR];   // published columns {th,x,y,c,c+1}: col[k*LPR + i] = Sigma(i, col_k); M columns on the strict path
    double row[5 * G::LPC];   // published rows    {th,x,y,c,c+1}: row[k*LPC + j] = Sigma(row_k, j)
    double sc[16];            // scalar-phase outputs: H (8), S^-1 (4), innovation (2), predict b10, b20 (2)
    double z[2 * G::M_MAX];   // this step's measurements (range, bearing)
    double tw[2];             // this step's twist (dth, dx)
    int ids[G::M_MAX];        // this step's landmark ids
    int flags[8];             // [0] update flags, [1] seen, [2] status, [3] seen snapshot, [4] theta owes a wrap, [5] frozen
    int req;                  // mailbox to the scalar server: kReqNone / kReqPredict / kReqExit / kReqUpdate + measurement index
    int pad_;
    uint64_t done;            // mbarrier the scalar server arrives on when the request has been served
};
constexpr int kReqNone = 0, kReqPredict = 1, kReqExit = 2, kReqUpdate = 16;

constexpr int kFlagSkip = 1, kFlagStrict = 2;

// Publish the landmark's two rows and two columns. cr = cI / 4 is the tile row holding the pair; the tile
// column is cr / 2. Lanes whose row group (a) matches store their row slice, lanes whose column group (b)
// matches store their column slice; register indices are static inside every case.
template <int TR, int TC>
__device__ __forceinline__ void publish_landmark(double * rdst, double * cdst, bool rmatch, bool cmatch, const double (&S)[TR][TC], int cr)
{
    switch (cr)
    {
#define NUSLAM_PL(k)                                                                                          \
    case k:                                                                                                   \
        if constexpr (TR > k)                                                                                 \
        {                                                                                                     \
            if (rmatch)                                                                                       \
            {                                                                                                 \
                _Pragma("unroll") for (int q = 0; q < TC; ++q) rdst[8 * q] = S[k < TR ? k : 0][q];            \
            }                                                                                                 \
            if (cmatch)                                                                                       \
            {                                                                                                 \
                _Pragma("unroll") for (int r = 0; r < TR; ++r) cdst[4 * r] = S[r][(k / 2) < TC ? (k / 2) : 0]; \
            }                                                                                                 \
        }                                                                                                     \
        break;
        NUSLAM_PL(0)
        NUSLAM_PL(1)
        NUSLAM_PL(2)
        NUSLAM_PL(3)
        NUSLAM_PL(4)
        NUSLAM_PL(5)
        NUSLAM_PL(6)
        NUSLAM_PL(7)
        NUSLAM_PL(8)
#undef NUSLAM_PL
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

// Common case of the scalar part (landmark already initialised and past its first touch): straight-line,
// branch-free, short dependency chains. H's structure (h01 = -h0c, h02 = -h0c1, h1c = -h11, h1c1 = -h12) folds
// every 4- or 5-term contraction with H into two FMAs on pre-formed differences of Sigma entries.
template <int N>
__device__ __forceinline__ void scalar_phase_fast(FastSmem<N> & f, double z0, double z1, int cI, const double * R)
{
    constexpr int LP = FastGeom<N>::LPR;
    const double * c = f.col;
    double s0[5], A[5], B[5];   // per column j of the 5x5 block: Sigma(th,j), Sigma(c,j)-Sigma(x,j), Sigma(c+1,j)-Sigma(y,j)
#pragma unroll
    for (int j = 0; j < 5; ++j)
    {
        s0[j] = c[j * LP + 1];
        A[j] = c[j * LP + cI] - c[j * LP + 2];
        B[j] = c[j * LP + cI + 1] - c[j * LP + 3];
    }
    const double th_raw = f.xs[1];
    const double th = f.flags[4] ? wrap_fast(th_raw) : th_raw;   // normalize_angle owed by the previous fused update (:276)
    const double dx = f.xs[cI] - f.xs[2], dy = f.xs[cI + 1] - f.xs[3];
    const double d = fma(dx, dx, dy * dy);
    const double rs = rsqrt_fast(d);
    const double id2 = rs * rs;
    double sq = d * rs;
    sq = fma(fma(-sq, sq, d), 0.5 * rs, sq);   // sqrt(d) to ~1 ulp
    const double h0c = dx * rs, h0c1 = dy * rs, h11 = dy * id2, h12 = -dx * id2;
    const double zb = wrap_fast(atan2_fast(dy, dx) - th);
    double g0[5], g1[5];   // H * Sigma at the 5 columns
#pragma unroll
    for (int j = 0; j < 5; ++j)
    {
        g0[j] = fma(h0c, A[j], h0c1 * B[j]);
        g1[j] = -fma(h11, A[j], fma(h12, B[j], s0[j]));
    }
    // psi = (H Sigma) H^T + R
    const double p00 = fma(h0c, g0[3] - g0[1], fma(h0c1, g0[4] - g0[2], R[0]));
    const double p10 = fma(h0c, g1[3] - g1[1], fma(h0c1, g1[4] - g1[2], R[1]));
    const double p01 = fma(h11, g0[1] - g0[3], fma(h12, g0[2] - g0[4], R[2] - g0[0]));
    const double p11 = fma(h11, g1[1] - g1[3], fma(h12, g1[2] - g1[4], R[3] - g1[0]));
    const double det = fma(p00, p11, -p01 * p10);
    const double idet = rcp_fast(det);
    const bool ok = det != 0.0;
    f.xs[1] = th;
    f.sc[0] = -h0c;
    f.sc[1] = -h0c1;
    f.sc[2] = h0c;
    f.sc[3] = h0c1;
    f.sc[4] = h11;
    f.sc[5] = h12;
    f.sc[6] = -h11;
    f.sc[7] = -h12;
    f.sc[8] = ok ? p11 * idet : 0.0;
    f.sc[9] = ok ? -p01 * idet : 0.0;
    f.sc[10] = ok ? -p10 * idet : 0.0;
    f.sc[11] = ok ? p00 * idet : 0.0;
    f.sc[12] = ok ? z0 - sq : 0.0;   // :272, no wrap
    f.sc[13] = ok ? z1 - zb : 0.0;
    f.flags[0] = ok ? 0 : kFlagSkip;
    f.flags[4] = ok ? 1 : 0;         // the fused path wraps theta lazily
    if (!ok) f.flags[2] |= kStatusSingular;
}

// Scalar part of one update for one filter (one lane): H, S = H Sigma H^T + R, S^-1, innovation.
// Reads the 5x5 block of Sigma from the published columns. slam_library.cpp:265-272 (+ :255-261 when the
// landmark is new).
template <int N>
__device__ __forceinline__ void scalar_phase(FastSmem<N> & f, double z0, double z1, int id, bool do_init, const double * R)
{
    using G = FastGeom<N>;
    constexpr int LP = G::LPR;          // leading dimension of the published columns
    const int cI = 4 + 2 * (id - 1);   // internal column of the landmark's x
    if (!do_init && !((f.col[3 * LP + cI] > kFirstTouchVariance) || (f.col[4 * LP + cI + 1] > kFirstTouchVariance)))
    {
        scalar_phase_fast<N>(f, z0, z1, cI, R);
        return;
    }
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

// filter slot `g` gets a no-op update: K = 0, innovation 0 (branch-free skip in phases C and D)
template <int N>
__device__ __forceinline__ void scalar_skip(FastSmem<N> & g)
{
#pragma unroll
    for (int k = 0; k < 14; ++k) g.sc[k] = 0.0;
    g.flags[0] = kFlagSkip;
}

// Scalar server: each lane watches the mailbox of one matrix warp and serves whatever is posted there, so
// the rsqrt / atan2 / reciprocal chains of up to 32 filters advance together in one instruction stream.
template <int N>
__device__ __forceinline__ void scalar_server(FastSmem<N> * fs, int server, int lane, const EkfParams & p, int do_predict)
{
    const int w = lane * kScalarWarps + server;
    bool alive = w < kMatrixWarps;
    FastSmem<N> & g = fs[alive ? w : 0];
    while (__any_sync(0xffffffffu, alive))
    {
        const int r = alive ? ld_acquire_smem(&g.req) : kReqNone;
        if (r == kReqPredict)
        {
            predict_scalar<N>(g);
        }
        else if (r >= kReqUpdate)
        {
            const int i = r - kReqUpdate;
            const int gid = g.ids[i];
            if (gid >= 1 && gid <= N)
            {
                const bool do_init = do_predict && gid > g.flags[3];   // slam.cpp:295 (step protocol only)
                if (do_predict && gid > g.flags[1]) g.flags[1] = gid;    // what associateLandmark would have done to `seen`
                scalar_phase<N>(g, g.z[2 * i], g.z[2 * i + 1], gid, do_init, p.R);
            }
            else
            {
                scalar_skip<N>(g);
                if (gid > N) g.flags[2] |= kStatusBadId;
            }
        }
        if (r == kReqExit) alive = false;
        else if (r != kReqNone)
        {
            g.req = kReqNone;       // ordered before the arrive (release) below
            mbar_arrive(&g.done);
        }
        if (!__any_sync(0xffffffffu, r != kReqNone)) __nanosleep(40);
    }
}

template <int N>
__global__ void __launch_bounds__(kFastThreads, 1) k_ekf_fast_step(const EkfParams p, const int do_predict)
{
    using G = FastGeom<N>;
    constexpr int TR = G::TR, TC = G::TC, LPR = G::LPR, LPC = G::LPC, LEN = G::LEN, SIG = G::SIG;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    FastSmem<N> * fs = reinterpret_cast<FastSmem<N> *>(smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp < kMatrixWarps && lane == 0)
    {
        mbar_init(&fs[warp].done, 1);
        fs[warp].req = kReqNone;
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();   // the only CTA-wide barrier of the kernel
    if (warp >= kMatrixWarps)
    {
        scalar_server<N>(fs, warp - kMatrixWarps, lane, p, do_predict);
        return;
    }

    const int a = lane >> 3, b = lane & 7;
    const int m = p.m;
    FastSmem<N> & f = fs[warp];
    // lane-invariant addresses and predicates, hoisted out of every loop
    const int goff = (b - 1) * LEN + (a - 1);            // HBM offset of S[0][0] inside the filter's Sigma
    double * const col_robot = f.col + (b - 1) * LPR + a;   // valid for b in 1..3
    double * const row_robot = f.row + (a - 1) * LPC + b;   // valid for a in 1..3
    const bool pub_col = (b >= 1) && (b <= 3);
    const bool pub_row = (a >= 1);
    double * const rdst = f.row + (3 + (a & 1)) * LPC + b;
    double * const cdst = f.col + (3 + (b & 1)) * LPR + a;
    const bool lane_row = lane < LPR;
    uint32_t parity = 0;

    const int64_t stride = (int64_t) gridDim.x * kMatrixWarps;
    for (int64_t bf = (int64_t) blockIdx.x * kMatrixWarps + warp; bf < p.batch; bf += stride)
    {
        // pull this warp's next Sigma towards L2 while the current one is computed
        if (lane == 0)
        {
            const int64_t nb = bf + stride;
            if (nb + 1 < p.batch && ((nb & 1) == 0)) prefetch_l2_bulk(p.sigma + nb * SIG, (uint32_t) (sizeof(double) * 2 * SIG));
        }
        // ---- load: Sigma tile straight into registers (32-byte runs per lane quartet), small inputs into shared memory ----
        double S[TR][TC];
        {
            const double * gs = p.sigma + bf * SIG + goff;
#pragma unroll
            for (int q = 0; q < TC; ++q)
#pragma unroll
                for (int r = 0; r < TR; ++r)
                {
                    const int i = 4 * r + a - 1, j = 8 * q + b - 1;   // external indices
                    S[r][q] = (i >= 0 && j >= 0 && i < LEN && j < LEN) ? __ldcs(gs + (8 * q) * LEN + 4 * r) : 0.0;
                }
            f.xs[lane] = (lane >= 1 && lane <= LEN) ? p.x[bf * LEN + lane - 1] : 0.0;
            if (lane < 2 * m) f.z[lane] = p.z[bf * m * 2 + lane];
            if (lane < m) f.ids[lane] = p.ids[bf * m + lane];
            if (lane < 2) f.tw[lane] = do_predict ? p.twists[bf * 3 + lane] : 0.0;
        }
        const int seen0 = p.seen[bf], st0 = p.status[bf];
        if (lane == 0)
        {
            f.flags[1] = seen0;
            f.flags[2] = st0;
            f.flags[3] = seen0;   // snapshot, slam.cpp:251
            f.flags[4] = 0;
        }
        if (st0 & (kStatusMapFull | kStatusSingular)) continue;   // the reference process died on an earlier scan
        __syncwarp();

        // ---- predict (slam_library.cpp:65-108), always in the oracle's operation order ----
        if (do_predict)
        {
            if (lane == 0) st_release_smem(&f.req, kReqPredict);
            mbar_wait(&f.done, parity);
            parity ^= 1;
            // T = A * Sigma: rows x (a = 2, r = 0) and y (a = 3, r = 0) += b * row theta (a = 1, r = 0)
            const double brow = (a == 2) ? f.sc[14] : f.sc[15];
            const double bcol = (b == 2) ? f.sc[14] : f.sc[15];
            const bool do_row = a >= 2;
            const bool do_col = (b == 2 || b == 3);
#pragma unroll
            for (int q = 0; q < TC; ++q)
            {
                const double thv = __shfl_sync(0xffffffffu, S[0][q], 8 + b);
                const double v = add_(mul_(brow, thv), S[0][q]);
                S[0][q] = do_row ? v : S[0][q];
            }
            // U = T * A.t(): columns x (b = 2, q = 0) and y (b = 3, q = 0) += b * column theta (b = 1, q = 0)
#pragma unroll
            for (int r = 0; r < TR; ++r)
            {
                const double t0 = __shfl_sync(0xffffffffu, S[r][0], 8 * a + 1);
                const double v = add_(mul_(t0, bcol), S[r][0]);
                S[r][0] = do_col ? v : S[r][0];
            }
            // + Q_bar on the robot block (internal rows/cols 1..3)
            {
                const bool inq = a >= 1 && b >= 1 && b <= 3;
                const double qv = inq ? p.Q[(a - 1) + 3 * (b - 1)] : 0.0;
                const double v = add_(S[0][0], qv);
                S[0][0] = inq ? v : S[0][0];
            }
        }

        // ---- m sequential updates (slam.cpp:279-319, known correspondence) ----
        for (int i = 0; i < m; ++i)
        {
            const int id = f.ids[i];
            const bool live = id >= 1 && id <= N;
            const int cI = live ? 4 + 2 * (id - 1) : 4;
            // A. publish the 5 columns and 5 rows of Sigma that H touches
            if (pub_col)
            {
#pragma unroll
                for (int r = 0; r < TR; ++r) col_robot[4 * r] = S[r][0];
            }
            if (pub_row)
            {
#pragma unroll
                for (int q = 0; q < TC; ++q) row_robot[8 * q] = S[0][q];
            }
            publish_landmark<TR, TC>(rdst, cdst, (a >> 1) == ((cI & 3) >> 1), (b >> 1) == ((cI & 7) >> 1), S, cI >> 2);
            __syncwarp();
            // B. hand the scalar part (H, S, S^-1, innovation) to the scalar server and sleep until it is served
            if (lane == 0) st_release_smem(&f.req, kReqUpdate + i);
            mbar_wait(&f.done, parity);
            parity ^= 1;
            const bool strict = (f.flags[0] & kFlagStrict) != 0;
            // C. lane i forms row i of K = Sigma H^T S^-1 and column i of W = H Sigma
            {
                const double h01 = f.sc[0], h02 = f.sc[1], h0c = f.sc[2], h0c1 = f.sc[3];
                const double h11 = f.sc[4], h12 = f.sc[5], h1c = f.sc[6], h1c1 = f.sc[7];
                const double i00 = f.sc[8], i01 = f.sc[9], i10 = f.sc[10], i11 = f.sc[11];
                const double dz0 = f.sc[12], dz1 = f.sc[13];
                const int rr = lane_row ? lane : 0;
                const double c0 = f.col[0 * LPR + rr], c1 = f.col[1 * LPR + rr], c2 = f.col[2 * LPR + rr];
                const double c3 = f.col[3 * LPR + rr], c4 = f.col[4 * LPR + rr];
                if (!strict)
                {
                    const double r0 = f.row[0 * LPC + lane], r1 = f.row[1 * LPC + lane], r2 = f.row[2 * LPC + lane];
                    const double r3 = f.row[3 * LPC + lane], r4 = f.row[4 * LPC + lane];
                    const double p0 = h01 * c1 + h02 * c2 + h0c * c3 + h0c1 * c4;
                    const double p1 = -c0 + h11 * c1 + h12 * c2 + h1c * c3 + h1c1 * c4;
                    const double k0 = p0 * i00 + p1 * i10, k1 = p0 * i01 + p1 * i11;
                    const double w0 = h01 * r1 + h02 * r2 + h0c * r3 + h0c1 * r4;
                    const double w1 = -r0 + h11 * r1 + h12 * r2 + h1c * r3 + h1c1 * r4;
                    f.wt[lane] = make_double2(w0, w1);
                    if (lane_row)
                    {
                        f.xs[lane] += k0 * dz0 + k1 * dz1;
                        f.kt[lane] = make_double2(k0, k1);
                    }
                }
                else if (lane_row)
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
                    f.xs[lane] = add_(f.xs[lane], add_(mul_(k0, dz0), mul_(k1, dz1)));
                    // M columns overwrite this lane's own entries of col[] (row[] keeps the old rows of Sigma)
                    f.col[0 * LPR + lane] = sub_((lane == 1) ? 1.0 : 0.0, -k1);
                    f.col[1 * LPR + lane] = sub_((lane == 2) ? 1.0 : 0.0, add_(mul_(k0, h01), mul_(k1, h11)));
                    f.col[2 * LPR + lane] = sub_((lane == 3) ? 1.0 : 0.0, add_(mul_(k0, h02), mul_(k1, h12)));
                    f.col[3 * LPR + lane] = sub_((lane == cI) ? 1.0 : 0.0, add_(mul_(k0, h0c), mul_(k1, h1c)));
                    f.col[4 * LPR + lane] = sub_((lane == cI + 1) ? 1.0 : 0.0, add_(mul_(k0, h0c1), mul_(k1, h1c1)));
                }
            }
            __syncwarp();
            // D. tile update
            if (!strict)
            {
                double2 w[TC];
#pragma unroll
                for (int q = 0; q < TC; ++q) w[q] = f.wt[8 * q + b];
#pragma unroll
                for (int r = 0; r < TR; ++r)
                {
                    const double2 k = f.kt[4 * r + a];
#pragma unroll
                    for (int q = 0; q < TC; ++q) S[r][q] = fma(-k.y, w[q].y, fma(-k.x, w[q].x, S[r][q]));
                }
            }
            else if (!(f.flags[0] & kFlagSkip))
            {
                // Sigma = M * Sigma in the oracle's ascending-k order (see ekf_strict.cuh)
#pragma unroll
                for (int r = 0; r < TR; ++r)
                {
                    const int ii = 4 * r + a;
                    const double m0 = f.col[0 * LPR + ii], m1 = f.col[1 * LPR + ii], m2 = f.col[2 * LPR + ii];
                    const double m3 = f.col[3 * LPR + ii], m4 = f.col[4 * LPR + ii];
#pragma unroll
                    for (int q = 0; q < TC; ++q)
                    {
                        const int jj = 8 * q + b;
                        double acc = mul_(m0, f.row[0 * LPC + jj]);
                        acc = add_(acc, mul_(m1, f.row[1 * LPC + jj]));
                        acc = add_(acc, mul_(m2, f.row[2 * LPC + jj]));
                        if (ii >= 4 && ii < cI) acc = add_(acc, S[r][q]);
                        acc = add_(acc, mul_(m3, f.row[3 * LPC + jj]));
                        acc = add_(acc, mul_(m4, f.row[4 * LPC + jj]));
                        if (ii > cI + 1) acc = add_(acc, S[r][q]);
                        S[r][q] = acc;
                    }
                }
                if (lane == 0) f.xs[1] = normalize_angle(f.xs[1]);   // :276, exact chain on the strict path
            }
            __syncwarp();
        }

        // ---- write back: registers -> HBM (each lane quartet writes 32 contiguous bytes) ----
        if (lane == 0 && f.flags[4]) f.xs[1] = wrap_fast(f.xs[1]);
        __syncwarp();
        {
            double * gs = p.sigma + bf * SIG + goff;
#pragma unroll
            for (int q = 0; q < TC; ++q)
#pragma unroll
                for (int r = 0; r < TR; ++r)
                {
                    const int i = 4 * r + a - 1, j = 8 * q + b - 1;
                    if (i >= 0 && j >= 0 && i < LEN && j < LEN) __stcs(gs + (8 * q) * LEN + 4 * r, S[r][q]);
                }
            if (lane >= 1 && lane <= LEN) p.x[bf * LEN + lane - 1] = f.xs[lane];
            if (lane == 0)
            {
                p.seen[bf] = f.flags[1];
                p.status[bf] = f.flags[2];
            }
        }
        __syncwarp();
    }
    if (lane == 0) st_release_smem(&f.req, kReqExit);
}

template <int N>
constexpr size_t fast_smem_bytes()
{
    return kMatrixWarps * sizeof(FastSmem<N>);
}

inline bool fast_supported(int n) { return n == 12 || n == 6; }

template <int N>
int launch_fast_n(const EkfParams & p, bool do_predict, int sm_count, cudaStream_t stream)
{
    static thread_local bool configured = false;
    constexpr size_t smem = fast_smem_bytes<N>();
    if (!configured)
    {
        cudaError_t e = cudaFuncSetAttribute(k_ekf_fast_step<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
        if (e != cudaSuccess) return (int) e;
        configured = true;
    }
    // persistent: one CTA per SM (kMatrixWarps independent filters in flight + the scalar servers)
    int64_t blocks = (p.batch + kMatrixWarps - 1) / kMatrixWarps;
    if (blocks > sm_count) blocks = sm_count;
    k_ekf_fast_step<N><<<(unsigned) blocks, kFastThreads, smem, stream>>>(p, do_predict ? 1 : 0);
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
