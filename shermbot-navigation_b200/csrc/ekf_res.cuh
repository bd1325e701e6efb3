// ekf_res.cuh -- FAST arithmetic, one filter per warp, Sigma RESIDENT IN SHARED MEMORY: fused predict + m sequential updates with
// known correspondence (the headline kernel of BASELINE.json configs[1]).
//
// Why (round-1 / round-2 profiles, profiles/ncu_r01_ekf_fast_step_final.txt, profiles/r02_kernel_iterations.md): the register-fragment
// kernels (ekf_fast.cuh and its pair / static variants) hold the landmark block of Sigma as fp64 tensor-core accumulators -- 36
// registers that are live through every dependency chain of an update -- and must PUBLISH two rows and two columns of it through shared
// memory before every update (27 shared-memory wavefronts per update, the column half of it at a quarter of the store width because
// a fragment column is scattered over all four quarter-warps). At 126 registers an SM holds 16 warps = 16 filters; every warp is one
// serial instruction stream (~8 cycles per instruction), so the kernel sat at 0.30 of the HBM roof with no pipe above 65 %.
// This kernel changes the exchange instead of tuning it:
//   * the bulk async copy (TMA engine) lands the filter's Sigma image in shared memory and it STAYS there: one bulk copy in, one out,
//     in place -- no fragment load / store passes, no second staging buffer (8.4 KB of shared memory per warp instead of 11.8);
//   * an update reads its landmark's two rows and two columns STRAIGHT FROM THE IMAGE in vector layout (lane = state index; a column
//     is contiguous, a row has stride 27 doubles = conflict-free) -- the publish step is gone;
//   * updates are DELAYED in chunks of CH (2 or 4): inside a chunk the image is stale by the chunk's earlier updates u, and the four
//     vectors are corrected on the fly with Sigma_i(r, c) = Sigma_0(r, c) - sum_u Kt_u(r) Minv_u Wt_u(c) (eight FMAs per pending update,
//     the lane's own Kt_u / Wt_u from registers, the landmark's from two broadcast reads) -- the delayed-update algebra of
//     ekf_large.cuh; at the end of the chunk ONE pass applies the rank-2CH update to the landmark block IN PLACE on the fp64 tensor
//     pipe (mma.m8n8k4: accumulator tile read from the image, CH/2 DMMAs, written back). The tile rows are permuted
//     ({r, r+1, r+8, r+9} per half-warp) so that the tile accesses of two of the three row blocks are free of bank conflicts in the
//     packed 27 x 27 column-major image;
//   * robot rows / columns, the state and the pose live in registers exactly as in ekf_fast.cuh (vector layout; predict touches only
//     them); the image's copies of the robot rows / columns are refreshed only where the chunk's landmarks read them;
//   * nothing Sigma-sized is in registers any more: the kernel fits 80 (CH = 2) / 96 (CH = 4) registers, i.e. 24 / 20 resident
//     single-warp CTAs per SM instead of 16.
// The arithmetic of the vector / scalar part is ekf_fast.cuh's, statement for statement (division-free Ht, one reciprocal, one-Newton
// rsqrt, table atan2, replicated pose). A filter-step with a first touch / initializeLandmark goes to the strict work list as there.
// Reference: nuslam/src/slam_library.cpp:65-108 (predict), :263-282 (update), nuslam/src/slam.cpp:262-319 (caller protocol).
#pragma once
#include "ekf_fast.cuh"

namespace nuslam
{

#ifndef NUSLAM_RES_CH
#define NUSLAM_RES_CH 2        // updates per delayed chunk (2 or 4)
#endif
#ifndef NUSLAM_RES_SHFL
#define NUSLAM_RES_SHFL 0      // 1: the 2 x 2 part's inputs by register shuffles instead of broadcast reads (timing experiment)
#endif
#ifndef NUSLAM_RES_EXP
#define NUSLAM_RES_EXP 0       // timing what-ifs (WRONG results): 1 no in-place pass, 2 no 2 x 2 part, 3 no row / column reads, 4 no DMMA (loads / stores kept)
#endif
#ifndef NUSLAM_RES_CTAS
#define NUSLAM_RES_CTAS (NUSLAM_RES_CH == 2 ? 24 : 19)
#endif
constexpr int kResCtasPerSm = NUSLAM_RES_CTAS;

template <int CH>
struct __align__(16) ResSmem
{
    double2 kt[CH][36];        // -Kt of the chunk's updates, kt[s][i] = (-k0, -k1) of state index i; DMMA A operand, pending corrections
    double2 wt[CH][36];        // Wt of the chunk's updates; DMMA B operand (slot stride 72 doubles = 8 (mod 16): conflict-free operand loads)
    double z[2 * kFastMMax];   // this step's measurements
};

template <int N, int CH>
__global__ void __launch_bounds__(32, kResCtasPerSm)
k_ekf_res_step(const EkfParams p, const int do_predict, int32_t * __restrict__ worklist, int32_t * __restrict__ wl_count)
{
    using G = FastGeom<N>;
    static_assert(G::FIXED && 2 * N == 8 * G::NB && G::LEN <= 32, "resident layout: fixed map size, unpadded 8 x 8 tiles, one lane per state index");
    static_assert(CH == 2 || CH == 4, "chunks of 2 or 4 updates");
    constexpr int NB = G::NB, NL = N, LEN = G::LEN, SIG = G::SIG;
    constexpr unsigned kFull = 0xffffffffu;
    constexpr int kWin = ((SIG * 8 + 8 + 15) / 16) * 16;   // 16-byte aligned window that covers the image at either alignment
    __shared__ __align__(128) unsigned char stage[kWin];
    __shared__ __align__(16) ResSmem<CH> f;
    __shared__ uint64_t full_bar;
    const int lane = threadIdx.x;
    const int g = lane >> 2, t = lane & 3;
    const bool vlane = lane < LEN;
    // the lanes that own no state entry re-read what lanes 16..20 of their half-warp read (same address = broadcast: no bank conflict,
    // no access outside the image); their results are never used
    const int vl = vlane ? lane : lane - 11;
    // tile geometry of the in-place pass: rows of row block a for this lane's g (blocks 0, 1: {r, r+1, r+8, r+9} per half-warp, block 2
    // consecutive), columns consecutive
    const int rrow0 = 3 + (g & 1) + 8 * ((g >> 1) & 1) + 2 * (g >> 2);
    int rrow[NB];
#pragma unroll
    for (int a = 0; a < NB; ++a) rrow[a] = (a + 1 < NB) ? rrow0 + 4 * a : 3 + 8 * a + g;
    static_assert(NB == 3, "row permutation written for three row blocks");

    auto issue_load = [&](int64_t b) {   // lane 0 only: the 16-byte aligned window around filter b's Sigma
        const unsigned char * g0 = reinterpret_cast<const unsigned char *>(p.sigma + b * SIG);
        const uintptr_t lo = reinterpret_cast<uintptr_t>(g0) & ~(uintptr_t) 15;
        uint32_t bytes = kWin;
        const uintptr_t end = reinterpret_cast<uintptr_t>(p.sigma + p.batch * SIG);
        if (lo + bytes > end) bytes = (uint32_t) ((end - lo) & ~(uintptr_t) 15);   // never read past the array: the tail is fetched separately
        mbar_expect_tx(&full_bar, bytes);
        bulk_g2s(stage, reinterpret_cast<const void *>(lo), bytes, &full_bar);
    };
    uint32_t full_parity = 0;
    if (lane == 0)
    {
        mbar_init(&full_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if ((int64_t) blockIdx.x < p.batch) issue_load(blockIdx.x);
    }
    __syncwarp();
    const int m = p.m;

    for (int64_t bf = blockIdx.x; bf < p.batch; bf += gridDim.x)
    {
        const bool next = bf + gridDim.x < p.batch;
        // the image after next: towards L2 while this one is computed
        if (NUSLAM_RES_EXP != 7 && lane == 0 && bf + 2 * (int64_t) gridDim.x < p.batch)
        {
            const uintptr_t a0 = reinterpret_cast<uintptr_t>(p.sigma + (bf + 2 * (int64_t) gridDim.x) * SIG) & ~(uintptr_t) 15;
            prefetch_l2_bulk(reinterpret_cast<const void *>(a0), (uint32_t) ((sizeof(double) * SIG + 15) & ~15u));
        }
        // ---- small inputs: plain loads, issued before anything waits ----
        double x = vlane ? p.x[bf * LEN + lane] : 0.0;
        const int st0 = p.status[bf], seen0 = p.seen[bf];
        const int my_id = (lane < m) ? p.ids[bf * m + lane] : 0;
        const double my_z = (lane < 2 * m) ? p.z[bf * m * 2 + lane] : 0.0;
        const double my_tw = (do_predict && lane < 2) ? p.twists[bf * 3 + lane] : 0.0;
        double * const img = reinterpret_cast<double *>(stage + (reinterpret_cast<uintptr_t>(p.sigma + bf * SIG) & 15));
        if (NUSLAM_RES_EXP != 7 || bf == blockIdx.x)
        {
            mbar_wait(&full_bar, full_parity);
            full_parity ^= 1;
        }
        if (bf == p.batch - 1 && lane == 0 && ((reinterpret_cast<uintptr_t>(p.sigma + p.batch * SIG) & 15) != 0))
            img[SIG - 1] = p.sigma[bf * SIG + SIG - 1];   // tail the clamped window left out
        __syncwarp();
        // whoever leaves this iteration without storing starts the next image's copy (the buffer is free at once)
        auto leave = [&]() {
            __syncwarp();
            if (NUSLAM_RES_EXP != 7 && lane == 0 && next) issue_load(bf + gridDim.x);
        };
        // ---- liveness, first touch ----
        if (st0 & (kStatusMapFull | kStatusSingular))   // the reference process died on an earlier scan
        {
            if (p.ids_out && lane < m) p.ids_out[bf * m + lane] = 0;
            if (p.x_snap && vlane) p.x_snap[bf * LEN + lane] = x;
            leave();
            continue;
        }
        const bool idok = (unsigned) (my_id - 1) < (unsigned) NL;
        {
            const int c = idok ? 1 + 2 * my_id : 3;
            const double d0 = img[c * (LEN + 1)], d1 = img[(c + 1) * (LEN + 1)];
            // first touch (INT_MAX prior) or initializeLandmark (slam.cpp:295-297): the strict kernel takes this filter-step
            const bool need = idok && ((do_predict && my_id > seen0) || d0 > kFirstTouchVariance || d1 > kFirstTouchVariance);
            if (__any_sync(kFull, need))
            {
                if (lane == 0) worklist[atomicAdd(wl_count, 1)] = (int32_t) bf;
                leave();
                continue;
            }
        }
        int status = st0;
        if (p.ids_out && lane < m) p.ids_out[bf * m + lane] = my_id > 0 ? my_id : 0;
        // the m <= 16 ids as 4-bit codes (0 = no update in that slot) in two warp-uniform words; an id above N is flagged once
        unsigned idlo, idhi;
        {
            static_assert(G::NMAX <= 15 && kFastMMax <= 16, "4-bit id codes in two 32-bit words");
            const unsigned code = idok ? (unsigned) my_id : 0u;
            idlo = __reduce_or_sync(kFull, lane < 8 ? code << (4 * lane) : 0u);
            idhi = __reduce_or_sync(kFull, (lane >= 8 && lane < 16) ? code << (4 * (lane - 8)) : 0u);
            if (__any_sync(kFull, my_id > NL)) status |= kStatusBadId;
        }
        unsigned idw = idlo;
        if (lane < 2 * m) f.z[lane] = my_z;
        // ---- robot rows / columns: image -> registers (vector layout) ----
        double Ct = img[vl], Cx = img[LEN + vl], Cy = img[2 * LEN + vl];
        double Rt = img[vl * LEN], Rx = img[vl * LEN + 1], Ry = img[vl * LEN + 2];
        // robot pose, replicated in every lane; lanes 0..2 own the same values in x (bit-identical updates)
        double th = __shfl_sync(kFull, x, 0), px = __shfl_sync(kFull, x, 1), py = __shfl_sync(kFull, x, 2);

        // ---- predict (slam_library.cpp:65-108), oracle operation order, vector layout only ----
        if (do_predict && NUSLAM_RES_EXP != 6)
        {
            const double dth = __shfl_sync(kFull, my_tw, 0), dxx = __shfl_sync(kFull, my_tw, 1);
            double s0, c0, b10, b20;
            sincos_fast(th, &s0, &c0);
            if (dth == 0.0)
            {
                px = add_(px, mul_(dxx, c0));
                py = add_(py, mul_(dxx, s0));
                th = add_(th, 0.0);
                b10 = mul_(-dxx, s0);
                b20 = mul_(dxx, c0);
            }
            else
            {
                const double q = div_fast(dxx, dth);
                double sd, cd;
                sincos_small(dth, &sd, &cd);
                const double s1 = fma(s0, cd, c0 * sd), c1 = fma(c0, cd, -s0 * sd);
                const double s3 = fma(s1, cd, c1 * sd), c3 = fma(c1, cd, -s1 * sd);
                px = add_(px, add_(mul_(-q, s0), mul_(q, s1)));
                py = add_(py, sub_(mul_(q, c0), mul_(q, c1)));
                th = add_(th, dth);
                b10 = add_(mul_(-q, c1), mul_(q, c3));
                b20 = add_(mul_(-q, s1), mul_(q, s3));
            }
            x = (lane == 0) ? th : (lane == 1) ? px : (lane == 2) ? py : x;
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
            if (lane < 3)
            {
                Rt = add_(Rt, p.Q[0 + 3 * lane]);
                Rx = add_(Rx, p.Q[1 + 3 * lane]);
                Ry = add_(Ry, p.Q[2 + 3 * lane]);
                Ct = add_(Ct, p.Q[lane + 3 * 0]);
                Cx = add_(Cx, p.Q[lane + 3 * 1]);
                Cy = add_(Cy, p.Q[lane + 3 * 2]);
            }
        }

        // ---- m sequential updates in delayed chunks of CH (slam.cpp:279-319, known correspondence) ----
#pragma unroll 1
        for (int i0 = 0; i0 < (NUSLAM_RES_EXP == 5 ? 0 : m); i0 += CH)
        {
            if (i0 == 8) idw = idhi;
            int cc[CH];
            bool mine = false;
#pragma unroll
            for (int s = 0; s < CH; ++s)
            {
                const int id = (int) (idw & 15u);
                idw >>= 4;
                cc[s] = id ? 1 + 2 * id : -1;   // warp-uniform; -1: no measurement in this slot
                mine = mine || (cc[s] >= 0 && (unsigned) (lane - cc[s]) < 2u);
            }
            // the image's robot rows / columns are stale (they live in registers): refresh the entries this chunk's landmarks read
            if (mine)
            {
                img[lane] = Ct;
                img[LEN + lane] = Cx;
                img[2 * LEN + lane] = Cy;
                img[lane * LEN] = Rt;
                img[lane * LEN + 1] = Rx;
                img[lane * LEN + 2] = Ry;
            }
            __syncwarp();
            double pK0[CH - 1], pK1[CH - 1], pW0[CH - 1], pW1[CH - 1];   // this lane's -Kt and Wt of the chunk's earlier updates
#pragma unroll
            for (int s = 0; s < CH; ++s)
            {
                bool done = false;
                double W0 = 0.0, W1 = 0.0, nk0 = 0.0, nk1 = 0.0;
                if (cc[s] >= 0)   // warp-uniform
                {
                    const int c = cc[s];
                    // landmark rows c, c+1 (lane = column) and columns c, c+1 (lane = row) as of the chunk's start, from the image
#if NUSLAM_RES_EXP == 3
                    double rho0 = Rt * 0.5, rho1 = Rx * 0.5, kap0 = Ct * 0.5, kap1 = Cx * 0.5;
#else
                    double rho0 = img[vl * LEN + c], rho1 = img[vl * LEN + c + 1];
                    double kap0 = img[c * LEN + vl], kap1 = img[(c + 1) * LEN + vl];
#endif
                    const double mxv = __shfl_sync(kFull, x, c), myv = __shfl_sync(kFull, x, c + 1);
                    const double2 zz = *reinterpret_cast<const double2 *>(&f.z[2 * (i0 + s)]);
                    // bring the four vectors up to date with the chunk's earlier updates (delayed-update algebra)
#pragma unroll
                    for (int u = 0; u < s; ++u)
                    {
                        const double2 ka = f.kt[u][c], kb = f.kt[u][c + 1], wa2 = f.wt[u][c], wb2 = f.wt[u][c + 1];
                        rho0 = fma(ka.x, pW0[u], fma(ka.y, pW1[u], rho0));
                        rho1 = fma(kb.x, pW0[u], fma(kb.y, pW1[u], rho1));
                        kap0 = fma(pK0[u], wa2.x, fma(pK1[u], wa2.y, kap0));
                        kap1 = fma(pK0[u], wb2.x, fma(pK1[u], wb2.y, kap1));
                    }
                    // (B) Pt (row role) and Wt (column role) of this lane
                    const double dx = mxv - px, dy = myv - py;
                    const double d = fma(dx, dx, dy * dy);
                    const double pa = kap0 - Cx, pb = kap1 - Cy;
                    const double wa = rho0 - Rx, wb = rho1 - Ry;
                    const double P0 = fma(dx, pa, dy * pb), P1 = fma(dx, pb, fma(-dy, pa, -d * Ct));
                    W0 = fma(dx, wa, dy * wb);
                    W1 = fma(dx, wb, fma(-dy, wa, -d * Rt));
                    f.wt[s][lane] = make_double2(W0, W1);   // DMMA B operand of the chunk's pass, pending corrections of later slots
#if NUSLAM_RES_SHFL
                    // Wt at the five state indices of H by register shuffles (measured SLOWER than the broadcast reads: 443 vs 409 us)
                    const double2 g0 = make_double2(__shfl_sync(kFull, W0, 0), __shfl_sync(kFull, W1, 0));
                    const double2 g1 = make_double2(__shfl_sync(kFull, W0, 1), __shfl_sync(kFull, W1, 1));
                    const double2 g2 = make_double2(__shfl_sync(kFull, W0, 2), __shfl_sync(kFull, W1, 2));
                    const double2 g3 = make_double2(__shfl_sync(kFull, W0, c), __shfl_sync(kFull, W1, c));
                    const double2 g4 = make_double2(__shfl_sync(kFull, W0, c + 1), __shfl_sync(kFull, W1, c + 1));
#else
                    __syncwarp();
                    const double2 g0 = f.wt[s][0], g1 = f.wt[s][1], g2 = f.wt[s][2], g3 = f.wt[s][c], g4 = f.wt[s][c + 1];
#endif
                    // the 2 x 2 part, evaluated by every lane: M = Wt Ht^T + D^-1 R D^-1, Minv, innovation (:150-160, :272 no wrap)
#if NUSLAM_RES_EXP == 2
                    const double m00 = 1.0 + 1e-9 * g0.x, m01 = 1e-9 * g0.y, m10 = 1e-9 * g1.x, m11 = 1.0 + 1e-9 * g1.y;
                    const double idet = 1.0 + 1e-9 * g2.x;
                    const double n0 = 1e-6 * (zz.x - dx) + 1e-12 * (g3.x + g4.x), n1 = 1e-6 * (zz.y - dy) + 1e-12 * g2.y;
#else
                    const double e0 = g3.x - g1.x, f0 = g4.x - g2.x, e1 = g3.y - g1.y, f1 = g4.y - g2.y;
                    const double s00 = fma(dx, e0, dy * f0), s01 = fma(dx, f0, fma(-dy, e0, -d * g0.x));
                    const double s10 = fma(dx, e1, dy * f1), s11 = fma(dx, f1, fma(-dy, e1, -d * g0.y));
                    const double rs = rsqrt_1(d);
                    double sq = d * rs;
                    sq = fma(fma(-sq, sq, d), 0.5 * rs, sq);   // sqrt(d) to ~1 ulp
                    const double dsq = d * sq;
                    const double m00 = fma(d, p.R[0], s00), m10 = fma(dsq, p.R[1], s10), m01 = fma(dsq, p.R[2], s01), m11 = fma(d * d, p.R[3], s11);
                    const double det = fma(m00, m11, -m01 * m10);
                    const double idet = rcp_fast(det);
                    double zb = atan2_unit(dy, dx, rs) - th;
                    if (__any_sync(kFull, abs_ge_hi(zb, kHiPi))) zb = wrap_angle(zb);   // the identity inside [-pi, pi]; a real (uniform) branch
                    const double n0 = sq * (zz.x - sq), n1 = d * (zz.y - zb);
#endif
                    if (!abs_ge_hi(idet, kHi1e300))   // warp-uniform; false for det = 0, inf or nan, where arma::inv throws (slam_library.cpp:270)
                    {
                        done = true;
                        const double i00 = m11 * idet, i01 = -m01 * idet, i10 = -m10 * idet, i11 = m00 * idet;
                        // (C) -Kt = -Pt Minv, x += Kt n
                        nk0 = fma(-P0, i00, -P1 * i10);
                        nk1 = fma(-P0, i01, -P1 * i11);
                        f.kt[s][lane] = make_double2(nk0, nk1);   // DMMA A operand, pending corrections
#if NUSLAM_RES_SHFL
                        const double2 k0 = make_double2(__shfl_sync(kFull, nk0, 0), __shfl_sync(kFull, nk1, 0));
                        const double2 k1 = make_double2(__shfl_sync(kFull, nk0, 1), __shfl_sync(kFull, nk1, 1));
                        const double2 k2 = make_double2(__shfl_sync(kFull, nk0, 2), __shfl_sync(kFull, nk1, 2));
#else
                        __syncwarp();
                        const double2 k0 = f.kt[s][0], k1 = f.kt[s][1], k2 = f.kt[s][2];
#endif
                        // replicated pose: what lanes 0..2 compute for their own x, evaluated identically by every lane
                        th = fma(-k0.x, n0, fma(-k0.y, n1, th));
                        px = fma(-k1.x, n0, fma(-k1.y, n1, px));
                        py = fma(-k2.x, n0, fma(-k2.y, n1, py));
                        x = fma(-nk0, n0, fma(-nk1, n1, x));
                        if (__any_sync(kFull, abs_ge_hi(th, kHiPi))) th = wrap_angle(th);   // slam_library.cpp:275-276 (the identity inside [-pi, pi])
                        if (lane == 0) x = th;
                        // robot rows / columns: Sigma -= Kt Wt restricted to them
                        Rt = fma(k0.x, W0, fma(k0.y, W1, Rt));
                        Rx = fma(k1.x, W0, fma(k1.y, W1, Rx));
                        Ry = fma(k2.x, W0, fma(k2.y, W1, Ry));
                        Ct = fma(nk0, g0.x, fma(nk1, g0.y, Ct));
                        Cx = fma(nk0, g1.x, fma(nk1, g1.y, Cx));
                        Cy = fma(nk0, g2.x, fma(nk1, g2.y, Cy));
                    }
                    else
                        status |= kStatusSingular;
                }
                if (!done)
                {
                    // no measurement in this slot (or a singular one): it contributes nothing to the chunk's pass or to later corrections
                    W0 = W1 = nk0 = nk1 = 0.0;
                    f.kt[s][lane] = make_double2(0.0, 0.0);
                    f.wt[s][lane] = make_double2(0.0, 0.0);
                }
                if (s + 1 < CH) __syncwarp();   // later slots read this slot's Kt / Wt at their landmark's indices
                if (s + 1 < CH)
                {
                    pK0[s < CH - 1 ? s : 0] = nk0;
                    pK1[s < CH - 1 ? s : 0] = nk1;
                    pW0[s < CH - 1 ? s : 0] = W0;
                    pW1[s < CH - 1 ? s : 0] = W1;
                }
            }
            // (D) one pass applies the chunk to the landmark block of the image in place: tile += (-Kt) Wt, k = (u0, u1, v0, v1, ...)
            __syncwarp();
#if NUSLAM_RES_EXP != 1
            {
                const double * const ka = reinterpret_cast<const double *>(&f.kt[t >> 1][0]) + (t & 1);
                const double * const wa = reinterpret_cast<const double *>(&f.wt[t >> 1][3 + g]) + (t & 1);
                double a[CH / 2][NB], b[CH / 2][NB];
#pragma unroll
                for (int kk = 0; kk < CH / 2; ++kk)
#pragma unroll
                    for (int bb = 0; bb < NB; ++bb)
                    {
                        a[kk][bb] = ka[kk * 144 + 2 * rrow[bb]];
                        b[kk][bb] = wa[kk * 144 + 16 * bb];
                    }
                double * const tbase = img + (3 + 2 * t) * LEN;
                double c0[NB][NB], c1[NB][NB];
#pragma unroll
                for (int bc = 0; bc < NB; ++bc)
#pragma unroll
                    for (int br = 0; br < NB; ++br)
                    {
                        const double * const e0p = tbase + bc * (8 * LEN) + rrow[br];
                        c0[bc][br] = e0p[0];
                        c1[bc][br] = e0p[LEN];
                    }
#pragma unroll
                for (int kk = 0; kk < CH / 2; ++kk)
#pragma unroll
                    for (int bc = 0; bc < NB; ++bc)
#pragma unroll
                        for (int br = 0; br < NB; ++br)
                        {
#if NUSLAM_RES_EXP == 4
                            c0[bc][br] += a[kk][br];
                            c1[bc][br] += b[kk][bc];
#else
                            dmma884(c0[bc][br], c1[bc][br], a[kk][br], b[kk][bc]);
#endif
                        }
#pragma unroll
                for (int bc = 0; bc < NB; ++bc)
#pragma unroll
                    for (int br = 0; br < NB; ++br)
                    {
                        double * const e0p = tbase + bc * (8 * LEN) + rrow[br];
                        e0p[0] = c0[bc][br];
                        e0p[LEN] = c1[bc][br];
                    }
            }
#endif
            __syncwarp();
        }

        // ---- write back: robot rows / columns into the image, the image to HBM by one bulk store + one plain store ----
        if (vlane)
        {
            img[lane * LEN] = Rt;
            img[lane * LEN + 1] = Rx;
            img[lane * LEN + 2] = Ry;
            if (lane >= 3)
            {
                img[lane] = Ct;
                img[LEN + lane] = Cx;
                img[2 * LEN + lane] = Cy;
            }
            p.x[bf * LEN + lane] = x;
            if (p.x_snap) p.x_snap[bf * LEN + lane] = x;
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0)
        {
            double * gw = p.sigma + bf * SIG;
            const int odd = (int) ((reinterpret_cast<uintptr_t>(gw) >> 3) & 1);   // 1: the image starts 8 bytes past a 16-byte boundary
            static_assert((SIG & 1) == 1, "a filter's Sigma is an odd number of doubles");
            if (NUSLAM_RES_EXP != 7)
            {
                bulk_s2g(gw + odd, img + odd, (SIG - 1) * 8);   // the 16-byte aligned interior
                const int edge = odd ? 0 : SIG - 1;
                gw[edge] = img[edge];
            }
            if (status != st0) p.status[bf] = status;
            // the buffer receives the next image as soon as the store has read it
            if (NUSLAM_RES_EXP != 7)
            {
                bulk_wait_read();
                if (next) issue_load(bf + gridDim.x);
            }
        }
        __syncwarp();
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // shared memory must outlive the last bulk store
}

template <int N>
int launch_res_n(const EkfParams & p, bool do_predict, int sm_count, int32_t * worklist, int32_t * wl_count, cudaStream_t stream)
{
    int64_t blocks = p.batch;
    if (blocks > kResCtasPerSm * (int64_t) sm_count) blocks = kResCtasPerSm * (int64_t) sm_count;
    static bool configured_dev[kMaxDevices] = {false};
    bool & configured = configured_dev[device_slot()];
    if (!configured)
    {
        cudaFuncSetAttribute(k_ekf_res_step<N, NUSLAM_RES_CH>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
        configured = true;
    }
    k_ekf_res_step<N, NUSLAM_RES_CH><<<(unsigned) blocks, 32, 0, stream>>>(p, do_predict ? 1 : 0, worklist, wl_count);
    return (int) cudaGetLastError();
}

// known correspondence, 16-byte aligned Sigma array, the BASELINE map size; everything else stays with ekf_fast.cuh
inline bool res_supported(int n, const EkfParams & p)
{
    return n == 12 && p.ids != nullptr && p.m_valid == nullptr && p.m >= 0 && p.m <= kFastMMax && p.batch >= 1 &&
           (reinterpret_cast<uintptr_t>(p.sigma) & 15) == 0 && (reinterpret_cast<uintptr_t>(p.x) & 7) == 0;
}

}   // namespace nuslam
