// ekf_pair.cuh -- FAST arithmetic, TWO filters per warp: fused predict + m sequential updates, known correspondence.
//
// Why pairs (round-1 profile of the one-filter-per-warp kernel, profiles/ncu_r01_ekf_fast_step_final.txt): the kernel was bound
// by three half-used resources at once -- fp64 pipe 43 %, issue slots 54 %, shared-memory wavefronts 65 % -- at 16 filters in
// flight per SM, and a third of its fp64 instructions were the 2 x 2 part of an update (M, M^-1, sqrt d, atan2, innovation)
// evaluated redundantly by all 32 lanes of a warp: the fp64 pipe does not skip idle lanes (profiles/ubench_fp64_r01.txt), so a
// per-filter scalar costs a full warp instruction. Here a warp carries filters (2 pr, 2 pr + 1):
//   * VECTOR layout per HALF-warp: lane (h = lane / 16, q = lane % 16) serves filter h and owns state indices q and 16 + q
//     (two slots). A vector statement is issued once per slot = twice per warp and serves both filters: the same count per
//     filter as before. A scalar statement (the 2 x 2 part, the pose, predict's trigonometry) and every broadcast read from
//     shared memory is issued ONCE and serves both filters: each half evaluates its own filter's chain in the same instruction.
//   * FRAGMENT layout per warp: the landmark block of either filter as fp64 tensor-core accumulators (mma.m8n8k4), 18 doubles
//     per lane and filter, as in ekf_fast.cuh; publish / lazy rank-4 DMMA passes run once per filter.
//   * 24 filters in flight per SM (12 single-warp CTAs of 168 registers) instead of 16.
//   * Sigma of a pair is ONE 16-byte aligned 2 x 5 832 B block in HBM: one bulk async copy (TMA engine) in, one out; the staging
//     buffer is the input image, then the two exchange areas, then the output image.
// Per filter and update this removes half of the scalar fp64 instructions, half of the broadcast reads and half of the warp
// barriers; everything else is the arithmetic of ekf_fast.cuh, statement for statement (same rounding: a filter's results do
// not depend on which kernel ran it, tests/test_ekf_gpu.py::test_pair_kernel_equals_single).
// A filter-step with a first touch / initializeLandmark goes to the strict work list exactly as in ekf_fast.cuh.
// Reference: nuslam/src/slam_library.cpp:65-108 (predict), :263-282 (update), nuslam/src/slam.cpp:262-319.
#pragma once
#include "ekf_fast.cuh"

namespace nuslam
{

#ifndef NUSLAM_PAIR_CTAS
#define NUSLAM_PAIR_CTAS 8
#endif
#ifndef NUSLAM_PAIR_STAGES
#define NUSLAM_PAIR_STAGES 2   // 2: buffer A prefetches the next pair while buffer B is exchange area / output image; 1: one buffer
#endif
constexpr int kPairStages = NUSLAM_PAIR_STAGES;
constexpr int kPairCtasPerSm = NUSLAM_PAIR_CTAS;
// registers per thread: the 64 K registers of an SM over kPairCtasPerSm single-warp CTAs, in the allocation granule of 8
#ifndef NUSLAM_PAIR_REGS
#define NUSLAM_PAIR_REGS ((65536 / (NUSLAM_PAIR_CTAS * 32)) / 8 * 8 > 255 ? 255 : (65536 / (NUSLAM_PAIR_CTAS * 32)) / 8 * 8)
#endif

// exchange area of ONE filter; vectors are indexed by the state index i < 32 (entries i >= LEN stay zero)
struct __align__(16) PairSmem
{
    double2 kt[2][36];        // -Kt of the chunk's two updates (DMMA A operand)
    double2 wt[2][36];        // Wt of the chunk's two updates (DMMA B operand)
    double rho[2][2][40];     // per chunk slot: landmark rows c, c+1, entry j at [j + 1]
    double2 kap[2][32];       // per chunk slot: landmark columns (c, c+1) interleaved, entry i
    double z[2 * kFastMMax];  // this step's measurements
};

// shared-memory copy of the atan2 tables (fastmath.cuh): the two halves of a warp index them with different entries, which the
// constant cache serialises at a long latency -- and the bearing chain is the critical path of an update
struct PairTables
{
    AtanEntry unit[65];
    AtanOctant oct[8];
};

__device__ __forceinline__ double atan2_unit_tab(double y, double x, double rs, const PairTables & T)
{
    const double ax = abs_bits(x), ay = abs_bits(y);
    const bool sw = gt_nonneg(ay, ax);
    const double mx = sw ? ay : ax, mn = sw ? ax : ay;
    const int idx = (sw ? 1 : 0) | (sign_bit(x) ? 2 : 0) | (sign_bit(y) ? 4 : 0);
    const float tf = __fdividef((float) mn, (float) mx);
    const int k = max(0, min(64, __float2int_rn(tf * 64.0f)));
    const AtanEntry t = T.unit[k];
    const AtanOctant oc = T.oct[idx];
    const double u = mn * rs, v = mx * rs;
    const double e = fma(u, t.c, -(v * t.s));
    const double e2 = e * e;
    const double pl = fma(fma(fma(kFastK2[0], e2, kFastK2[1]), e2, kFastK2[2]), e2 * e, e);   // asin(e)
    const double a = t.hi + (pl + t.lo);
    return oc.hi + fma(oc.s, a, oc.lo);
}

// (A) publish landmark rows / columns c, c+1 (slot-space update i, chunk slot s) of either filter from its fragments into vector
// layout; after inlining into the unrolled update loop s, i and everything derived from them are constants
template <int NB>
__device__ __forceinline__ void pair_publish(const double (&C)[2][NB][NB][2], const double (&Rt)[2], const double (&Rx)[2], const double (&Ry)[2],
                                             const double (&Ct)[2], const double (&Cx)[2], const double (&Cy)[2], PairSmem * const (&EF)[2],
                                             PairSmem & E, const unsigned live_w, const int g, const int t, const int q, const int s, const int i)
{
    const int c = 3 + 2 * i;
    const int bsel = (2 * i) >> 3, sel = ((2 * i) & 7) >> 1;
    const bool rsel = (g >> 1) == sel;   // this lane holds row c or c+1
    const bool csel = t == sel;          // this lane holds columns c, c+1
#pragma unroll
    for (int f = 0; f < 2; ++f)
    {
        if ((live_w >> (16 * f + i)) & 1u)   // warp-uniform
        {
            PairSmem & F = *EF[f];
            double * const rdst = &F.rho[s][g & 1][4 + 2 * t];
            double2 * const cdst = &F.kap[s][3 + g];
#pragma unroll
            for (int qq = 0; qq < NB; ++qq)
            {
                if (rsel) *reinterpret_cast<double2 *>(rdst + 8 * qq) = make_double2(C[f][bsel][qq][0], C[f][bsel][qq][1]);
                if (csel) cdst[8 * qq] = make_double2(C[f][qq][bsel][0], C[f][qq][bsel][1]);
            }
        }
    }
    // robot part of rows / columns c, c+1: the lanes that own those indices, for their own filter (both filters at once)
#pragma unroll
    for (int e = 0; e < 2; ++e)
    {
        const int sl = (c + e) >> 4;
        if (q == ((c + e) & 15))
        {
            E.rho[s][e][1] = Ct[sl];
            E.rho[s][e][2] = Cx[sl];
            E.rho[s][e][3] = Cy[sl];
            double * kd = reinterpret_cast<double *>(&E.kap[s][0]) + e;
            kd[0] = Rt[sl];
            kd[2] = Rx[sl];
            kd[4] = Ry[sl];
        }
    }
}

// slot space: the pair kernel works on P Sigma P^T, x' = P x with the landmarks permuted so that measurement slot j of THIS step is
// landmark slot j (state indices 3 + 2 j, 4 + 2 j): the update loop is fully unrolled and every register index, shared-memory offset
// and lane predicate of update i is a compile-time constant (no id decode, no fragment-selection ladder, no address arithmetic).
// perm[j] = landmark (1-based) held by slot j; orig index of slot-space index i' >= 3: 1 + 2 perm[(i' - 3) / 2] + (i' - 3) % 2.
// The permutation exists when the step's ids are distinct valid landmarks (slots without a measurement -- id 0, or m < n -- take the
// unmeasured landmarks); a step with a repeated or out-of-range id goes to the strict work list.
template <int N>
__global__ void __maxnreg__(NUSLAM_PAIR_REGS)
k_ekf_pair_step(const EkfParams p, const int do_predict, int32_t * __restrict__ worklist, int32_t * __restrict__ wl_count)
{
    using G = FastGeom<N>;
    static_assert(G::FIXED && G::LEN > 16 && G::LEN <= 32 && 2 * N == 8 * G::NB, "pair layout: two slots of 16 state indices, unpadded fragments");
    constexpr int NB = G::NB, NL = N, LEN = G::LEN, SIG = G::SIG;
    constexpr unsigned kFull = 0xffffffffu;
    constexpr int kImg = SIG * 8;                 // bytes of one Sigma
    constexpr int kPairBytes = 2 * kImg;          // = 16 SIG: a pair is 16-byte aligned in HBM whenever the array is
    constexpr int kImg16 = (kImg + 15) / 16 * 16;
    constexpr int kEOff = kImg16 > (int) sizeof(PairSmem) ? kImg16 : (int) sizeof(PairSmem);   // exchange area of filter 1
    constexpr int kNeed = (kPairBytes > kEOff + (int) sizeof(PairSmem)) ? kPairBytes : kEOff + (int) sizeof(PairSmem);
    constexpr int kBuf = (kNeed + 127) / 128 * 128;
    __shared__ __align__(128) unsigned char stage[kPairStages][kBuf];
    __shared__ __align__(16) PairTables tabs;
    __shared__ uint64_t full_bar;
    __shared__ int perm_s[2][16];
    unsigned char * const buf = stage[0];                    // input images (bulk-loaded)
    unsigned char * const xbuf = stage[kPairStages - 1];     // exchange areas, then output images
    const int lane = threadIdx.x;
    const int h = lane >> 4, q = lane & 15;       // vector / scalar domain: filter of this lane, index inside the half
    const int g = lane >> 2, t = lane & 3;        // fragment domain
    PairSmem & E = *reinterpret_cast<PairSmem *>(xbuf + h * kEOff);                  // this lane's filter
    PairSmem * const EF[2] = {reinterpret_cast<PairSmem *>(xbuf), reinterpret_cast<PairSmem *>(xbuf + kEOff)};
    for (int k = lane; k < (int) (sizeof(PairTables) / 8); k += 32)
        reinterpret_cast<double *>(&tabs)[k] = (k < 260) ? reinterpret_cast<const double *>(kAtanUnit)[k] : reinterpret_cast<const double *>(kAtanOct)[k - 260];
    const int64_t npairs = (p.batch + 1) >> 1;
    const int m = p.m;

    auto issue_load = [&](int64_t pr) {   // lane 0 only
        const unsigned char * src = reinterpret_cast<const unsigned char *>(p.sigma + 2 * pr * SIG);
        // the last pair of an odd batch holds one filter: its 16-byte aligned part, the tail element is fetched separately
        const uint32_t bytes = (2 * pr + 1 < p.batch) ? (uint32_t) kPairBytes : (uint32_t) (kImg & ~15);
        mbar_expect_tx(&full_bar, bytes);
        bulk_g2s(buf, src, bytes, &full_bar);
    };
    uint32_t full_parity = 0;
#ifdef NUSLAM_TIMING
    long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long tlast = clock64();
#endif
    if (lane == 0)
    {
        mbar_init(&full_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if ((int64_t) blockIdx.x < npairs) issue_load(blockIdx.x);
    }
    __syncwarp();

    for (int64_t pr = blockIdx.x; pr < npairs; pr += gridDim.x)
    {
        const int64_t bf = 2 * pr + h;
        const bool has = bf < p.batch;
        const int64_t bfc = has ? bf : 2 * pr;   // safe addressing for the missing twin of an odd batch
        const bool next = pr + gridDim.x < npairs;
        // pull the next pair towards L2 while this one is computed (the staging buffer is busy until the output image has left)
        if (lane == 0 && next)
        {
            const int64_t pn = pr + gridDim.x;
            const uint32_t bytes = (2 * pn + 1 < p.batch) ? (uint32_t) kPairBytes : (uint32_t) (kImg & ~15);
            prefetch_l2_bulk(p.sigma + 2 * pn * SIG, bytes);
        }
        // ---- small inputs: plain loads, issued before anything waits ----
        const int st0 = p.status[bfc], seen0 = p.seen[bfc];
        const int my_id = (q < m) ? p.ids[bfc * m + q] : 0;
        const double my_z0 = (q < m) ? p.z[bfc * m * 2 + 2 * q] : 0.0;
        const double my_z1 = (q < m) ? p.z[bfc * m * 2 + 2 * q + 1] : 0.0;
        const double my_tw = (do_predict && q < 2) ? p.twists[bfc * 3 + q] : 0.0;
        // ---- the step's permutation: slot q <- landmark lq ----
        const bool idok = (unsigned) (my_id - 1) < (unsigned) NL;
        const bool meas = q < m && idok;                                    // slot q carries a measurement
        const unsigned bits = meas ? (1u << my_id) : 0u;
        const unsigned maskA = __reduce_or_sync(kFull, h ? 0u : bits), maskB = __reduce_or_sync(kFull, h ? bits : 0u);
        const unsigned mymask = h ? maskB : maskA;
        const unsigned meas_w = __ballot_sync(kFull, meas);
        const unsigned bad_w = __ballot_sync(kFull, q < m && my_id != 0 && !idok);   // an id outside 1..N (negative ids included)
        const unsigned meas_h = (meas_w >> (16 * h)) & 0xffffu;
        // a repeated landmark or a bad id: no permutation; the oracle-order kernel runs this filter-step (and flags the bad id)
        // (so is a measurement in a slot beyond the map size, m > n)
        const bool generic = __popc(meas_h) != __popc(mymask) || ((bad_w >> (16 * h)) & 0xffffu) != 0u || (meas_h >> NL) != 0u;
        int lq = my_id;
        if (!meas)
        {
            const unsigned unmeas = ~mymask & (((1u << NL) - 1u) << 1);
            const int before = __popc(~meas_h & ((1u << q) - 1u));          // slots without a measurement below this one
            lq = (int) __fns(unmeas, 0, before + 1);                        // the (before + 1)-th unmeasured landmark
        }
        if (q < NL) perm_s[h][q] = ((unsigned) (lq - 1) < (unsigned) NL) ? lq : 1;   // (garbage stays addressable when `generic`)

        // ---- Sigma: staging buffer -> registers ----
        mbar_wait(&full_bar, full_parity);
        full_parity ^= 1;
        if (2 * pr + 1 >= p.batch && lane == 0) reinterpret_cast<double *>(buf)[SIG - 1] = p.sigma[2 * pr * SIG + SIG - 1];
        __syncwarp();
        bool need;
        {
            // first touch (INT_MAX prior) or initializeLandmark (slam.cpp:295-297): the strict kernel takes this filter-step
            const double * img = reinterpret_cast<const double *>(buf + h * kImg);
            const int c = idok ? 1 + 2 * my_id : 3;
            const double d0 = img[c * (LEN + 1)], d1 = img[(c + 1) * (LEN + 1)];
            need = has && meas && ((do_predict && my_id > seen0) || d0 > kFirstTouchVariance || d1 > kFirstTouchVariance);
        }
        const unsigned need_w = __ballot_sync(kFull, need);
        const bool stat_dead = (st0 & (kStatusMapFull | kStatusSingular)) != 0;   // the reference process died on an earlier scan
        const bool to_strict = has && !stat_dead && (generic || ((need_w >> (16 * h)) & 0xffffu) != 0u);
        const bool dead = !has || stat_dead || to_strict;
        // orig state index of this lane's two vector slots
        int o[2];
#pragma unroll
        for (int sl = 0; sl < 2; ++sl)
        {
            const int i = 16 * sl + q;
            const int j = (i >= 3 && i < LEN) ? (i - 3) >> 1 : 0;
            o[sl] = (i < 3) ? i : (i < LEN) ? 1 + 2 * perm_s[h][j] + ((i - 3) & 1) : 0;
        }
        double x[2];
        x[0] = p.x[bfc * LEN + o[0]];
        x[1] = (16 + q < LEN) ? p.x[bfc * LEN + o[1]] : 0.0;
        if (has && stat_dead)
        {
            if (p.ids_out && q < m) p.ids_out[bf * m + q] = 0;
            if (p.x_snap)
            {
                p.x_snap[bf * LEN + o[0]] = x[0];
                if (16 + q < LEN) p.x_snap[bf * LEN + o[1]] = x[1];
            }
        }
        else if (to_strict && q == 0)
            worklist[atomicAdd(wl_count, 1)] = (int32_t) bf;
        const unsigned dead_w = __ballot_sync(kFull, dead);
        const bool deadA = (dead_w & 1u) != 0u, deadB = (dead_w & 0x10000u) != 0u;
        if (deadA && deadB)
        {
            __syncwarp();   // the images have been read (first-touch test)
            if (lane == 0 && next) issue_load(pr + gridDim.x);
            continue;
        }
        NUSLAM_T(0)
        double C[2][NB][NB][2];
        double Rt[2], Rx[2], Ry[2], Ct[2], Cx[2], Cy[2];
#pragma unroll
        for (int f = 0; f < 2; ++f)
        {
            const double * img = reinterpret_cast<const double *>(buf + f * kImg);
            int ro[NB], co[NB];
#pragma unroll
            for (int b = 0; b < NB; ++b)
            {
                ro[b] = 1 + 2 * perm_s[f][4 * b + (g >> 1)] + (g & 1);
                co[b] = (1 + 2 * perm_s[f][4 * b + t]) * LEN;
            }
#pragma unroll
            for (int br = 0; br < NB; ++br)
#pragma unroll
                for (int bc = 0; bc < NB; ++bc)
#pragma unroll
                    for (int e = 0; e < 2; ++e) C[f][br][bc][e] = img[co[bc] + e * LEN + ro[br]];
        }
        {
            const double * img = reinterpret_cast<const double *>(buf + h * kImg);
#pragma unroll
            for (int sl = 0; sl < 2; ++sl)
            {
                const bool v = 16 * sl + q < LEN;
                Ct[sl] = v ? img[o[sl]] : 0.0;
                Cx[sl] = v ? img[LEN + o[sl]] : 0.0;
                Cy[sl] = v ? img[2 * LEN + o[sl]] : 0.0;
                Rt[sl] = v ? img[o[sl] * LEN] : 0.0;
                Rx[sl] = v ? img[o[sl] * LEN + 1] : 0.0;
                Ry[sl] = v ? img[o[sl] * LEN + 2] : 0.0;
            }
        }
        __syncwarp();   // the images are dead from here on
        if (kPairStages > 1)
        {
            // buffer A is free again: prefetch this CTA's next pair while the current one is computed; buffer B held the previous
            // pair's output images: wait until the bulk store has read them
            if (lane == 0)
            {
                if (next) issue_load(pr + gridDim.x);
                bulk_wait_read();
            }
            __syncwarp();
        }
        int status = st0;
        if (!dead && p.ids_out && q < m) p.ids_out[bf * m + q] = my_id > 0 ? my_id : 0;
        // measurement slots of the live filters: bit i (filter 0), bit 16 + i (filter 1)
        const unsigned live_w = __ballot_sync(kFull, meas && !dead);
        // exchange entries that belong to no state index are zero
        if (16 + q >= LEN)
        {
#pragma unroll
            for (int s2 = 0; s2 < 2; ++s2)
            {
                E.rho[s2][0][16 + q + 1] = 0.0;
                E.rho[s2][1][16 + q + 1] = 0.0;
                E.kap[s2][16 + q] = make_double2(0.0, 0.0);
                E.kt[s2][16 + q] = make_double2(0.0, 0.0);
                E.wt[s2][16 + q] = make_double2(0.0, 0.0);
            }
        }
        if (q < m) *reinterpret_cast<double2 *>(&E.z[2 * q]) = make_double2(my_z0, my_z1);
        // robot pose of this lane's filter, replicated over its half; lanes q = 0..2 own the same values in x[0]
        double th = __shfl_sync(kFull, x[0], 16 * h), px = __shfl_sync(kFull, x[0], 16 * h + 1), py = __shfl_sync(kFull, x[0], 16 * h + 2);

        // ---- predict (slam_library.cpp:65-108), oracle operation order, vector layout only ----
        if (do_predict)
        {
            const double dth = __shfl_sync(kFull, my_tw, 16 * h), dxx = __shfl_sync(kFull, my_tw, 16 * h + 1);
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
                const double qq = div_fast(dxx, dth);
                double sd, cd;
                sincos_small(dth, &sd, &cd);
                const double s1 = fma(s0, cd, c0 * sd), c1 = fma(c0, cd, -s0 * sd);
                const double s3 = fma(s1, cd, c1 * sd), c3 = fma(c1, cd, -s1 * sd);
                px = add_(px, add_(mul_(-qq, s0), mul_(qq, s1)));
                py = add_(py, sub_(mul_(qq, c0), mul_(qq, c1)));
                th = add_(th, dth);
                b10 = add_(mul_(-qq, c1), mul_(qq, c3));
                b20 = add_(mul_(-qq, s1), mul_(qq, s3));
            }
            x[0] = (q == 0) ? th : (q == 1) ? px : (q == 2) ? py : x[0];
            // T = A * Sigma: rows x, y += b * row theta
#pragma unroll
            for (int sl = 0; sl < 2; ++sl)
            {
                Rx[sl] = add_(mul_(b10, Rt[sl]), Rx[sl]);
                Ry[sl] = add_(mul_(b20, Rt[sl]), Ry[sl]);
            }
            {
                const double t0 = __shfl_sync(kFull, Ct[0], 16 * h), t1 = __shfl_sync(kFull, Cx[0], 16 * h), t2 = __shfl_sync(kFull, Cy[0], 16 * h);
                const double bb = (q == 1) ? b10 : b20;
                if (q == 1 || q == 2)
                {
                    Ct[0] = add_(mul_(bb, t0), Ct[0]);
                    Cx[0] = add_(mul_(bb, t1), Cx[0]);
                    Cy[0] = add_(mul_(bb, t2), Cy[0]);
                }
            }
            // U = T * A.t(): columns x, y += column theta * b
#pragma unroll
            for (int sl = 0; sl < 2; ++sl)
            {
                Cx[sl] = add_(mul_(Ct[sl], b10), Cx[sl]);
                Cy[sl] = add_(mul_(Ct[sl], b20), Cy[sl]);
            }
            {
                const double t0 = __shfl_sync(kFull, Rt[0], 16 * h), t1 = __shfl_sync(kFull, Rx[0], 16 * h), t2 = __shfl_sync(kFull, Ry[0], 16 * h);
                const double bb = (q == 1) ? b10 : b20;
                if (q == 1 || q == 2)
                {
                    Rt[0] = add_(mul_(t0, bb), Rt[0]);
                    Rx[0] = add_(mul_(t1, bb), Rx[0]);
                    Ry[0] = add_(mul_(t2, bb), Ry[0]);
                }
            }
            // + Q_bar on the robot block (expanded_process_noise :110-125); Q is column-major
            if (q < 3)
            {
                Rt[0] = add_(Rt[0], p.Q[0 + 3 * q]);
                Rx[0] = add_(Rx[0], p.Q[1 + 3 * q]);
                Ry[0] = add_(Ry[0], p.Q[2 + 3 * q]);
                Ct[0] = add_(Ct[0], p.Q[q + 3 * 0]);
                Cx[0] = add_(Cx[0], p.Q[q + 3 * 1]);
                Cy[0] = add_(Cy[0], p.Q[q + 3 * 2]);
            }
        }
        __syncwarp();
        NUSLAM_T(1)

        // ---- m sequential updates in chunks of 2 (slam.cpp:279-319), slot space: update i works on state indices 3 + 2 i, 4 + 2 i ----
        // Shared-memory hand-overs per update: rows / columns -> (sync) -> Pt, Wt -> (sync) -> 2 x 2 part, -Kt -> (sync) -> pose, robot rows.
        // Everything that depends on the state alone (sqrt d, the bearing, the innovation) is evaluated BEFORE the first hand-over, and
        // the chunk's second publish is issued behind the first update's Wt hand-over, so that both overlap the waiting.
#pragma unroll
        for (int ch = 0; ch < (NL + 1) / 2; ++ch)
        {
            if (2 * ch >= m) break;   // warp-uniform
            double pW0[2], pW1[2], pK0[2], pK1[2];   // this lane's Wt and -Kt of the chunk's first update (lazy correction of the second)
#pragma unroll
            for (int s = 0; s < 2; ++s)
            {
                const int i = 2 * ch + s;
                const int c = 3 + 2 * i;
                const bool live = (live_w >> (16 * h + i)) & 1u;
                if (s == 0) pair_publish<NB>(C, Rt, Rx, Ry, Ct, Cx, Cy, EF, E, live_w, g, t, q, 0, i);
                // ---- state-only part: landmark position from the lanes that own it, sqrt d, bearing, innovation (:150-160, :272 no wrap) ----
                const double mxv = __shfl_sync(kFull, x[c >> 4], 16 * h + (c & 15));
                const double myv = __shfl_sync(kFull, x[(c + 1) >> 4], 16 * h + ((c + 1) & 15));
                const double2 zz = *reinterpret_cast<const double2 *>(&E.z[2 * (i < kFastMMax ? i : 0)]);
                const double dx = mxv - px, dy = myv - py;
                const double d = fma(dx, dx, dy * dy);
                const double rs = rsqrt_1(d);
                double sq = d * rs;
                sq = fma(fma(-sq, sq, d), 0.5 * rs, sq);   // sqrt(d) to ~1 ulp
                const double dsq = d * sq;
                double zb = atan2_unit_tab(dy, dx, rs, tabs) - th;
                if (__any_sync(kFull, abs_ge_hi(zb, kHiPi)))   // the wrap is the identity inside [-pi, pi]
                    if (abs_ge_hi(zb, kHiPi)) zb = wrap_angle(zb);
                double n0 = sq * (zz.x - sq), n1 = d * (zz.y - zb);
                const double r00 = d * p.R[0], r10 = dsq * p.R[1], r01 = dsq * p.R[2], r11 = (d * d) * p.R[3];   // D^-1 R D^-1
                if (s == 0) __syncwarp();   // the chunk's first publish (the second one is ordered by the first update's hand-overs)
                NUSLAM_T(2)
                double P0[2], P1[2], W0[2], W1[2];
                double2 ka, kb, wa2, wb2;
                if (s == 1)
                {
                    ka = E.kt[0][c];
                    kb = E.kt[0][c + 1];
                    wa2 = E.wt[0][c];
                    wb2 = E.wt[0][c + 1];
                }
#pragma unroll
                for (int sl = 0; sl < 2; ++sl)
                {
                    const int ii = 16 * sl + q;
                    // landmark rows c, c+1 (this lane = column ii) and columns c, c+1 (this lane = row ii)
                    double rho0 = E.rho[s][0][ii + 1], rho1 = E.rho[s][1][ii + 1];
                    const double2 kp = E.kap[s][ii];
                    double kap0 = kp.x, kap1 = kp.y;
                    if (s == 1)
                    {
                        // the fragments predate the chunk's first update: bring the four vectors up to date with it
                        rho0 = fma(ka.x, pW0[sl], fma(ka.y, pW1[sl], rho0));
                        rho1 = fma(kb.x, pW0[sl], fma(kb.y, pW1[sl], rho1));
                        kap0 = fma(pK0[sl], wa2.x, fma(pK1[sl], wa2.y, kap0));
                        kap1 = fma(pK0[sl], wb2.x, fma(pK1[sl], wb2.y, kap1));
                    }
                    // (B) Pt (row role) and Wt (column role)
                    const double pa = kap0 - Cx[sl], pb = kap1 - Cy[sl];
                    const double wa = rho0 - Rx[sl], wb = rho1 - Ry[sl];
                    P0[sl] = fma(dx, pa, dy * pb);
                    P1[sl] = fma(dx, pb, fma(-dy, pa, -d * Ct[sl]));
                    W0[sl] = fma(dx, wa, dy * wb);
                    W1[sl] = fma(dx, wb, fma(-dy, wa, -d * Rt[sl]));
                    E.wt[s][ii] = make_double2(W0[sl], W1[sl]);
                }
                __syncwarp();
                NUSLAM_T(3)
                if (s == 0 && 2 * ch + 1 < NL) pair_publish<NB>(C, Rt, Rx, Ry, Ct, Cx, Cy, EF, E, live_w, g, t, q, 1, i + 1);
                // the 2 x 2 part of this lane's filter: M = Wt Ht^T + D^-1 R D^-1, Minv
                const double2 g0 = E.wt[s][0], g1 = E.wt[s][1], g2 = E.wt[s][2], g3 = E.wt[s][c], g4 = E.wt[s][c + 1];
                const double e0 = g3.x - g1.x, f0 = g4.x - g2.x, e1 = g3.y - g1.y, f1 = g4.y - g2.y;
                const double m00 = fma(dx, e0, fma(dy, f0, r00)), m01 = fma(dx, f0, fma(-dy, e0, fma(-d, g0.x, r01)));
                const double m10 = fma(dx, e1, fma(dy, f1, r10)), m11 = fma(dx, f1, fma(-dy, e1, fma(-d, g0.y, r11)));
                const double det = fma(m00, m11, -m01 * m10);
                const double idet = rcp_fast(det);
                // |idet| < ~1e300: false for det = 0, inf or nan, where arma::inv throws (slam_library.cpp:270)
                const bool ok = live && !abs_ge_hi(idet, kHi1e300);
                const double i00 = m11 * idet, i01 = -m01 * idet, i10 = -m10 * idet, i11 = m00 * idet;
                // (C) -Kt = -Pt Minv
                double nk0[2], nk1[2];
#pragma unroll
                for (int sl = 0; sl < 2; ++sl)
                {
                    nk0[sl] = fma(-P0[sl], i00, -P1[sl] * i10);
                    nk1[sl] = fma(-P0[sl], i01, -P1[sl] * i11);
                }
                if (!__all_sync(kFull, ok))
                {
                    // a slot without a measurement (or a singular one) contributes nothing: its update is the identity
                    if (!ok)
                    {
                        if (live) status |= kStatusSingular;
                        n0 = 0.0;
                        n1 = 0.0;
#pragma unroll
                        for (int sl = 0; sl < 2; ++sl)
                        {
                            nk0[sl] = 0.0;
                            nk1[sl] = 0.0;
                            W0[sl] = 0.0;
                            W1[sl] = 0.0;
                            E.wt[s][16 * sl + q] = make_double2(0.0, 0.0);
                        }
                    }
                }
#pragma unroll
                for (int sl = 0; sl < 2; ++sl)
                {
                    E.kt[s][16 * sl + q] = make_double2(nk0[sl], nk1[sl]);
                    // robot columns: Sigma -= Kt Wt restricted to them (the last use of Wt(0..2): they leave the registers here)
                    Ct[sl] = fma(nk0[sl], g0.x, fma(nk1[sl], g0.y, Ct[sl]));
                    Cx[sl] = fma(nk0[sl], g1.x, fma(nk1[sl], g1.y, Cx[sl]));
                    Cy[sl] = fma(nk0[sl], g2.x, fma(nk1[sl], g2.y, Cy[sl]));
                    x[sl] = fma(-nk0[sl], n0, fma(-nk1[sl], n1, x[sl]));
                    if (s == 0)
                    {
                        pW0[sl] = W0[sl];
                        pW1[sl] = W1[sl];
                        pK0[sl] = nk0[sl];
                        pK1[sl] = nk1[sl];
                    }
                }
                __syncwarp();
                NUSLAM_T(4)
                const double2 k0 = E.kt[s][0], k1 = E.kt[s][1], k2 = E.kt[s][2];
                // replicated pose: what lanes q = 0..2 compute for their own x, evaluated identically by the whole half
                th = fma(-k0.x, n0, fma(-k0.y, n1, th));
                px = fma(-k1.x, n0, fma(-k1.y, n1, px));
                py = fma(-k2.x, n0, fma(-k2.y, n1, py));
                if (__any_sync(kFull, abs_ge_hi(th, kHiPi)))   // slam_library.cpp:275-276 (the identity inside [-pi, pi])
                    if (abs_ge_hi(th, kHiPi)) th = wrap_angle(th);
                if (q == 0) x[0] = th;
#pragma unroll
                for (int sl = 0; sl < 2; ++sl)
                {
                    // robot rows: Sigma -= Kt Wt restricted to them
                    Rt[sl] = fma(k0.x, W0[sl], fma(k0.y, W1[sl], Rt[sl]));
                    Rx[sl] = fma(k1.x, W0[sl], fma(k1.y, W1[sl], Rx[sl]));
                    Ry[sl] = fma(k2.x, W0[sl], fma(k2.y, W1[sl], Ry[sl]));
                }
                NUSLAM_T(5)
            }
            // (D) one DMMA pass per filter applies the chunk to its fragments: C += (-Kt) Wt, k = (u0, u1, v0, v1)
#pragma unroll
            for (int f = 0; f < 2; ++f)
            {
                if (f == 0 ? !deadA : !deadB)   // warp-uniform
                {
                    const PairSmem & F = *EF[f];
                    const double * ka = reinterpret_cast<const double *>(&F.kt[t >> 1][3 + g]) + (t & 1);
                    const double * wa = reinterpret_cast<const double *>(&F.wt[t >> 1][3 + g]) + (t & 1);
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
                        for (int bc = 0; bc < NB; ++bc) dmma884(C[f][br][bc][0], C[f][br][bc][1], a[br], b[bc]);
                }
            }
            NUSLAM_T(6)
        }
        __syncwarp();

        // ---- write back: registers -> output images (orig order) in the staging buffer -> bulk store ----
#pragma unroll
        for (int f = 0; f < 2; ++f)
        {
            if (f == 0 ? !deadA : !deadB)
            {
                double * img = reinterpret_cast<double *>(xbuf + f * kImg);
                int ro[NB], co[NB];
#pragma unroll
                for (int b = 0; b < NB; ++b)
                {
                    ro[b] = 1 + 2 * perm_s[f][4 * b + (g >> 1)] + (g & 1);
                    co[b] = (1 + 2 * perm_s[f][4 * b + t]) * LEN;
                }
#pragma unroll
                for (int br = 0; br < NB; ++br)
#pragma unroll
                    for (int bc = 0; bc < NB; ++bc)
#pragma unroll
                        for (int e = 0; e < 2; ++e) img[co[bc] + e * LEN + ro[br]] = C[f][br][bc][e];
            }
        }
        if (!dead)
        {
            double * img = reinterpret_cast<double *>(xbuf + h * kImg);
#pragma unroll
            for (int sl = 0; sl < 2; ++sl)
            {
                const int i = 16 * sl + q;
                if (i < LEN)
                {
                    img[o[sl] * LEN] = Rt[sl];
                    img[o[sl] * LEN + 1] = Rx[sl];
                    img[o[sl] * LEN + 2] = Ry[sl];
                    if (i >= 3)
                    {
                        img[o[sl]] = Ct[sl];
                        img[LEN + o[sl]] = Cx[sl];
                        img[2 * LEN + o[sl]] = Cy[sl];
                    }
                    p.x[bf * LEN + o[sl]] = x[sl];
                    if (p.x_snap) p.x_snap[bf * LEN + o[sl]] = x[sl];
                }
            }
            if (q == 0 && status != st0) p.status[bf] = status;
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0)
        {
            double * gw = p.sigma + 2 * pr * SIG;
            double * img = reinterpret_cast<double *>(xbuf);
            if (!deadA && !deadB)
                bulk_s2g(gw, img, kPairBytes);
            else if (!deadA)
            {
                // filter 2 pr alone: its 16-byte aligned part + the last element
                bulk_s2g(gw, img, kImg - 8);
                gw[SIG - 1] = img[SIG - 1];
            }
            else
            {
                // filter 2 pr + 1 alone: it starts 8 bytes past a 16-byte boundary, in HBM and in the buffer alike
                gw[SIG] = img[SIG];
                bulk_s2g(gw + SIG + 1, img + SIG + 1, kImg - 8);
            }
            if (kPairStages == 1)
            {
                // the buffer receives the next pair as soon as the store has read it
                bulk_wait_read();
                if (next) issue_load(pr + gridDim.x);
            }
        }
        __syncwarp();   // perm_s is rewritten by the next round
        NUSLAM_T(7)
    }
#ifdef NUSLAM_TIMING
    if (blockIdx.x == 0 && lane == 0)
        for (int k = 0; k < 8; ++k) atomicAdd((unsigned long long *) &g_fast_timing[k], (unsigned long long) tacc[k]);
#endif
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // shared memory must outlive the last bulk store
}

template <int N>
int launch_pair_n(const EkfParams & p, bool do_predict, int sm_count, int32_t * worklist, int32_t * wl_count, cudaStream_t stream)
{
    const int64_t npairs = (p.batch + 1) / 2;
    int64_t blocks = npairs;
    if (blocks > kPairCtasPerSm * (int64_t) sm_count) blocks = kPairCtasPerSm * (int64_t) sm_count;
    static bool configured_dev[kMaxDevices] = {false};
    bool & configured = configured_dev[device_slot()];
    if (!configured)
    {
        cudaFuncSetAttribute(k_ekf_pair_step<N>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
        configured = true;
    }
    k_ekf_pair_step<N><<<(unsigned) blocks, 32, 0, stream>>>(p, do_predict ? 1 : 0, worklist, wl_count);
    return (int) cudaGetLastError();
}

// known correspondence, 16-byte aligned Sigma, the BASELINE map size: the pair kernel; everything else stays with ekf_fast.cuh
inline bool pair_supported(int n, const EkfParams & p)
{
    const bool enabled = true;
    return enabled && n == 12 && p.ids != nullptr && p.m_valid == nullptr && p.m >= 0 && p.m <= kFastMMax && p.batch >= 1 &&
           (reinterpret_cast<uintptr_t>(p.sigma) & 15) == 0 && (reinterpret_cast<uintptr_t>(p.x) & 7) == 0;
}

}   // namespace nuslam
