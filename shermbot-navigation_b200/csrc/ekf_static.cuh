// ekf_static.cuh -- FAST arithmetic, one filter per warp, STATIC update schedule: fused predict + m sequential updates with known
// correspondence (the headline kernel of BASELINE.json configs[1]).
//
// Same arithmetic and register layout as ekf_fast.cuh (landmark block as fp64 tensor-core accumulator fragments, robot rows /
// columns and the state in vector layout, division-free H~, lazy rank-4 DMMA chunks, bulk async copies for Sigma). What changed is
// everything AROUND the arithmetic, after the round-1 profile showed the kernel latency-bound with 60 % of its instructions
// being address / id / selection logic (profiles/ncu_r01_ekf_fast_step_final.txt):
//   * SLOT SPACE. The filter is loaded as P Sigma P^T, x' = P x with the landmarks permuted so that measurement slot j of THIS step
//     is landmark slot j (state indices 3 + 2 j, 4 + 2 j). The update loop is fully unrolled and every register index, shared-
//     memory offset and lane predicate of update i is a compile-time constant: no id decode, no fragment-selection ladder, no
//     per-update address arithmetic. The permutation costs ~60 instructions per filter-step (address lookups at load / store).
//     It exists when the step's ids are distinct valid landmarks (slots without a measurement -- id 0, or m < n -- take the
//     unmeasured landmarks); a step with a repeated or out-of-range id goes to the strict work list, like a first touch.
//   * The part of an update that depends on the state alone (sqrt d, the bearing with its table atan2, the innovation,
//     D^-1 R D^-1) is evaluated BEFORE the first shared-memory hand-over of the update, and the chunk's second publish is issued
//     behind the first update's hand-over: both overlap the waiting instead of extending the dependency chain.
//   * The landmark position comes from the lanes that own it (two shuffles) instead of a shared-memory copy of the state.
// Reference: nuslam/src/slam_library.cpp:65-108 (predict), :263-282 (update), nuslam/src/slam.cpp:262-319.
#pragma once
#include "ekf_pair.cuh"

namespace nuslam
{

#ifndef NUSLAM_STATIC_CTAS
#define NUSLAM_STATIC_CTAS 16
#endif
constexpr int kStaticCtasPerSm = NUSLAM_STATIC_CTAS;

// exchange area of the filter; vectors are indexed by the (slot-space) state index i < 32 (entries i >= LEN stay zero)
struct __align__(16) StaticSmem
{
    double2 kt[2][36];      // -Kt of the chunk's two updates (DMMA A operand)
    double2 wt[2][36];      // Wt of the chunk's two updates (DMMA B operand)
    double rho[2][2][40];   // per chunk slot: landmark rows c, c+1, entry j at [j + 1]
    double2 kap[2][32];     // per chunk slot: landmark columns (c, c+1) interleaved, entry i
    double z[2 * kFastMMax];
};

// (A) publish landmark rows / columns c, c+1 (slot-space update i, chunk slot s) from the fragments into vector layout. The update
// loop is ROLLED (its fully unrolled form, 60 KB of code, lost more to instruction fetch than it saved: the instruction cache holds
// 32 KB): i is a loop value, so the fragment block is picked by a three-way switch; everything else is warp-uniform arithmetic on i.
template <int NB>
__device__ __forceinline__ void static_publish(const double (&C)[NB][NB][2], const double Rt, const double Rx, const double Ry, const double Ct,
                                               const double Cx, const double Cy, StaticSmem & F, const int g, const int t, const int lane,
                                               const int s, const int i)
{
    const int c = 3 + 2 * i;
    const int bsel = (2 * i) >> 3, sel = ((2 * i) & 7) >> 1;
    const bool rsel = (g >> 1) == sel;   // this lane holds row c or c+1
    const bool csel = t == sel;          // this lane holds columns c, c+1
    double * const rdst = &F.rho[s][g & 1][4 + 2 * t];
    double2 * const cdst = &F.kap[s][3 + g];
#define NUSLAM_SPUB(b)                                                                                                       \
    _Pragma("unroll") for (int qq = 0; qq < NB; ++qq)                                                                        \
    {                                                                                                                        \
        if (rsel) *reinterpret_cast<double2 *>(rdst + 8 * qq) = make_double2(C[b < NB ? b : 0][qq][0], C[b < NB ? b : 0][qq][1]); \
        if (csel) cdst[8 * qq] = make_double2(C[qq][b < NB ? b : 0][0], C[qq][b < NB ? b : 0][1]);                           \
    }
    if (bsel == 0)
    {
        NUSLAM_SPUB(0)
    }
    else if (bsel == 1)
    {
        NUSLAM_SPUB(1)
    }
    else
    {
        NUSLAM_SPUB(2)
    }
#undef NUSLAM_SPUB
    // robot part of rows / columns c, c+1: the two lanes that own those indices
    if (lane == c || lane == c + 1)
    {
        const int e = lane - c;
        F.rho[s][e][1] = Ct;
        F.rho[s][e][2] = Cx;
        F.rho[s][e][3] = Cy;
        double * kd = reinterpret_cast<double *>(&F.kap[s][0]) + e;
        kd[0] = Rt;
        kd[2] = Rx;
        kd[4] = Ry;
    }
}

template <int N>
__global__ void __launch_bounds__(32, kStaticCtasPerSm)
k_ekf_static_step(const EkfParams p, const int do_predict, int32_t * __restrict__ worklist, int32_t * __restrict__ wl_count)
{
    using G = FastGeom<N>;
    static_assert(G::FIXED && 2 * N == 8 * G::NB && G::LEN <= 32, "unpadded fragments, one lane per state index");
    constexpr int NB = G::NB, NL = N, LEN = G::LEN, SIG = G::SIG;
    constexpr unsigned kFull = 0xffffffffu;
    constexpr int kImg = SIG * 8;                                    // bytes of one Sigma
    constexpr int kWin = ((kImg + 8 + 15) / 16) * 16;                // 16-byte aligned window that covers it at either alignment
    constexpr int kStage = ((kWin > (int) sizeof(StaticSmem) ? kWin : (int) sizeof(StaticSmem)) + 127) / 128 * 128;
    __shared__ __align__(128) unsigned char stage[2][kStage];        // [0]: the NEXT filter's image; [1]: exchange area, then output image
    __shared__ uint64_t full_bar;
    __shared__ int perm_s[16];
    const int lane = threadIdx.x;
    const int g = lane >> 2, t = lane & 3;
    StaticSmem & E = *reinterpret_cast<StaticSmem *>(stage[1]);
    const int m = p.m;

    // issue the bulk load of filter b into stage 0 (lane 0 only): the 16-byte aligned window around its Sigma
    auto issue_load = [&](int64_t b) {
        const unsigned char * g0 = reinterpret_cast<const unsigned char *>(p.sigma + b * SIG);
        const uintptr_t lo = reinterpret_cast<uintptr_t>(g0) & ~(uintptr_t) 15;
        uint32_t bytes = kWin;
        // never read past the end of the array: the last filter's window loses its tail, fetched separately below
        const uintptr_t end = reinterpret_cast<uintptr_t>(p.sigma + p.batch * SIG);
        if (lo + bytes > end) bytes = (uint32_t) ((end - lo) & ~(uintptr_t) 15);
        mbar_expect_tx(&full_bar, bytes);
        bulk_g2s(stage[0], reinterpret_cast<const void *>(lo), bytes, &full_bar);
    };
    uint32_t full_parity = 0;
    if (lane == 0)
    {
        mbar_init(&full_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if ((int64_t) blockIdx.x < p.batch) issue_load(blockIdx.x);
    }
    __syncwarp();

    for (int64_t bf = blockIdx.x; bf < p.batch; bf += gridDim.x)
    {
        const bool next = bf + gridDim.x < p.batch;
        // ---- small inputs: plain loads, issued before anything waits ----
        const int st0 = p.status[bf], seen0 = p.seen[bf];
        const int my_id = (lane < m) ? p.ids[bf * m + lane] : 0;
        const double my_z = (lane < 2 * m) ? p.z[bf * m * 2 + lane] : 0.0;
        const double my_tw = (do_predict && lane < 2) ? p.twists[bf * 3 + lane] : 0.0;
        // ---- the step's permutation: slot `lane` <- landmark lq ----
        const bool idok = (unsigned) (my_id - 1) < (unsigned) NL;
        const bool meas = lane < m && idok;                                  // slot `lane` carries a measurement
        const unsigned mask = __reduce_or_sync(kFull, meas ? (1u << my_id) : 0u);
        const unsigned meas_w = __ballot_sync(kFull, meas);
        const unsigned bad_w = __ballot_sync(kFull, lane < m && my_id != 0 && !idok);   // an id outside 1..N (negative ids included)
        // a repeated landmark or a bad id: no permutation; the oracle-order kernel runs this filter-step (and flags the bad id)
        // (so is a measurement in a slot beyond the map size, m > n)
        const bool generic = __popc(meas_w) != __popc(mask) || bad_w != 0u || (meas_w >> NL) != 0u;
        int lq = my_id;
        if (!meas)
        {
            const unsigned unmeas = ~mask & (((1u << NL) - 1u) << 1);
            const int before = __popc(~meas_w & ((1u << lane) - 1u));         // slots without a measurement below this one
            lq = (int) __fns(unmeas, 0, before + 1);                          // the (before + 1)-th unmeasured landmark
        }
        if (lane < NL) perm_s[lane] = ((unsigned) (lq - 1) < (unsigned) NL) ? lq : 1;   // (garbage stays addressable when `generic`)

        // ---- Sigma: staging buffer -> registers ----
        // image of this filter inside stage 0: offset 0 or 8 (its alignment in HBM)
        const double * img = reinterpret_cast<const double *>(stage[0] + (reinterpret_cast<uintptr_t>(p.sigma + bf * SIG) & 15));
        mbar_wait(&full_bar, full_parity);
        full_parity ^= 1;
        if (bf == p.batch - 1 && lane == 0 && ((reinterpret_cast<uintptr_t>(p.sigma + p.batch * SIG) & 15) != 0))
            const_cast<double *>(img)[SIG - 1] = p.sigma[bf * SIG + SIG - 1];   // tail the clamped window left out
        __syncwarp();
        bool need;
        {
            // first touch (INT_MAX prior) or initializeLandmark (slam.cpp:295-297): the strict kernel takes this filter-step
            const int c = idok ? 1 + 2 * my_id : 3;
            const double d0 = img[c * (LEN + 1)], d1 = img[(c + 1) * (LEN + 1)];
            need = meas && ((do_predict && my_id > seen0) || d0 > kFirstTouchVariance || d1 > kFirstTouchVariance);
        }
        const bool stat_dead = (st0 & (kStatusMapFull | kStatusSingular)) != 0;   // the reference process died on an earlier scan
        const bool to_strict = !stat_dead && (generic || __any_sync(kFull, need));
        if (stat_dead || to_strict)
        {
            if (stat_dead)
            {
                if (p.ids_out && lane < m) p.ids_out[bf * m + lane] = 0;
                if (p.x_snap && lane < LEN) p.x_snap[bf * LEN + lane] = p.x[bf * LEN + lane];
            }
            else if (lane == 0)
                worklist[atomicAdd(wl_count, 1)] = (int32_t) bf;
            __syncwarp();   // the image has been read (first-touch test)
            if (lane == 0 && next) issue_load(bf + gridDim.x);
            continue;
        }
        // orig state index of this lane's slot-space index
        const int o = (lane < 3) ? lane : (lane < LEN) ? 1 + 2 * perm_s[(lane - 3) >> 1] + ((lane - 3) & 1) : 0;
        double x = (lane < LEN) ? p.x[bf * LEN + o] : 0.0;
        double C[NB][NB][2];
        double Rt = 0.0, Rx = 0.0, Ry = 0.0, Ct = 0.0, Cx = 0.0, Cy = 0.0;
        {
            int ro[NB], co[NB];
#pragma unroll
            for (int b = 0; b < NB; ++b)
            {
                ro[b] = 1 + 2 * perm_s[4 * b + (g >> 1)] + (g & 1);
                co[b] = (1 + 2 * perm_s[4 * b + t]) * LEN;
            }
#pragma unroll
            for (int br = 0; br < NB; ++br)
#pragma unroll
                for (int bc = 0; bc < NB; ++bc)
#pragma unroll
                    for (int e = 0; e < 2; ++e) C[br][bc][e] = img[co[bc] + e * LEN + ro[br]];
            if (lane < LEN)
            {
                Ct = img[o];
                Cx = img[LEN + o];
                Cy = img[2 * LEN + o];
                Rt = img[o * LEN];
                Rx = img[o * LEN + 1];
                Ry = img[o * LEN + 2];
            }
        }
        __syncwarp();   // the image is dead from here on
        // stage 0 is free again: prefetch this CTA's next filter while the current one is computed; stage 1 held the previous
        // filter's output image: wait until the bulk store has read it
        if (lane == 0)
        {
            if (next) issue_load(bf + gridDim.x);
            bulk_wait_read();
        }
        __syncwarp();
        int status = st0;
        if (p.ids_out && lane < m) p.ids_out[bf * m + lane] = my_id > 0 ? my_id : 0;
        const unsigned live_w = meas_w;   // measurement slots: bit i
        // exchange entries that belong to no state index are zero
        if (lane >= LEN)
        {
#pragma unroll
            for (int s2 = 0; s2 < 2; ++s2)
            {
                E.rho[s2][0][lane + 1] = 0.0;
                E.rho[s2][1][lane + 1] = 0.0;
                E.kap[s2][lane] = make_double2(0.0, 0.0);
                E.kt[s2][lane] = make_double2(0.0, 0.0);
                E.wt[s2][lane] = make_double2(0.0, 0.0);
            }
        }
        if (lane < 2 * m) E.z[lane] = my_z;
        // robot pose, replicated in every lane; lanes 0..2 own the same values in x (bit-identical updates)
        double th = __shfl_sync(kFull, x, 0), px = __shfl_sync(kFull, x, 1), py = __shfl_sync(kFull, x, 2);

        // ---- predict (slam_library.cpp:65-108), oracle operation order, vector layout only ----
        if (do_predict)
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
        __syncwarp();

        // ---- m sequential updates in chunks of 2 (slam.cpp:279-319), slot space: update i works on state indices 3 + 2 i, 4 + 2 i ----
#pragma unroll 1
        for (int ch = 0; ch < (NL + 1) / 2; ++ch)
        {
            if (2 * ch >= m) break;   // warp-uniform
            double pW0 = 0.0, pW1 = 0.0, pK0 = 0.0, pK1 = 0.0;   // this lane's Wt and -Kt of the chunk's first update (lazy correction of the second)
#pragma unroll
            for (int s = 0; s < 2; ++s)
            {
                const int i = 2 * ch + s;
                const int c = 3 + 2 * i;
                const bool live = (live_w >> i) & 1u;   // warp-uniform
                // ---- state-only part: landmark position from the lanes that own it, sqrt d, bearing, innovation (:150-160, :272 no wrap) ----
                const double mxv = __shfl_sync(kFull, x, c), myv = __shfl_sync(kFull, x, c + 1);
                const double2 zz = *reinterpret_cast<const double2 *>(&E.z[2 * (i < kFastMMax ? i : 0)]);
                const double dx = mxv - px, dy = myv - py;
                const double d = fma(dx, dx, dy * dy);
                const double rs = rsqrt_1(d);
                double sq = d * rs;
                sq = fma(fma(-sq, sq, d), 0.5 * rs, sq);   // sqrt(d) to ~1 ulp
                const double dsq = d * sq;
                double zb = atan2_unit(dy, dx, rs) - th;   // constant-bank tables: the index is warp-uniform here
                if (abs_ge_hi(zb, kHiPi)) zb = wrap_angle(zb);   // warp-uniform; the identity inside [-pi, pi]
                const double n0 = sq * (zz.x - sq), n1 = d * (zz.y - zb);
                const double r00 = d * p.R[0], r10 = dsq * p.R[1], r01 = dsq * p.R[2], r11 = (d * d) * p.R[3];   // D^-1 R D^-1
                if (s == 0)
                {
                    if (live) static_publish<NB>(C, Rt, Rx, Ry, Ct, Cx, Cy, E, g, t, lane, 0, i);
                    __syncwarp();   // (the chunk's second publish is ordered by the first update's hand-overs)
                }
                bool done = false;
                double W0 = 0.0, W1 = 0.0, nk0 = 0.0, nk1 = 0.0;
                if (live)
                {
                    // landmark rows c, c+1 (lane = column) and columns c, c+1 (lane = row)
                    double rho0 = E.rho[s][0][lane + 1], rho1 = E.rho[s][1][lane + 1];
                    const double2 kp = E.kap[s][lane];
                    double kap0 = kp.x, kap1 = kp.y;
                    if (s == 1)
                    {
                        // the fragments predate the chunk's first update: bring the four vectors up to date with it
                        const double2 ka = E.kt[0][c], kb = E.kt[0][c + 1], wa2 = E.wt[0][c], wb2 = E.wt[0][c + 1];
                        rho0 = fma(ka.x, pW0, fma(ka.y, pW1, rho0));
                        rho1 = fma(kb.x, pW0, fma(kb.y, pW1, rho1));
                        kap0 = fma(pK0, wa2.x, fma(pK1, wa2.y, kap0));
                        kap1 = fma(pK0, wb2.x, fma(pK1, wb2.y, kap1));
                    }
                    // (B) Pt (row role) and Wt (column role) of this lane
                    const double pa = kap0 - Cx, pb = kap1 - Cy;
                    const double wa = rho0 - Rx, wb = rho1 - Ry;
                    const double P0 = fma(dx, pa, dy * pb), P1 = fma(dx, pb, fma(-dy, pa, -d * Ct));
                    W0 = fma(dx, wa, dy * wb);
                    W1 = fma(dx, wb, fma(-dy, wa, -d * Rt));
                    E.wt[s][lane] = make_double2(W0, W1);
                    __syncwarp();
                    if (s == 0 && 2 * ch + 1 < NL && ((live_w >> (i + 1)) & 1u)) static_publish<NB>(C, Rt, Rx, Ry, Ct, Cx, Cy, E, g, t, lane, 1, i + 1);
                    // the 2 x 2 part, evaluated by every lane: M = Wt Ht^T + D^-1 R D^-1, Minv
                    const double2 g0 = E.wt[s][0], g1 = E.wt[s][1], g2 = E.wt[s][2], g3 = E.wt[s][c], g4 = E.wt[s][c + 1];
                    const double e0 = g3.x - g1.x, f0 = g4.x - g2.x, e1 = g3.y - g1.y, f1 = g4.y - g2.y;
                    const double m00 = fma(dx, e0, fma(dy, f0, r00)), m01 = fma(dx, f0, fma(-dy, e0, fma(-d, g0.x, r01)));
                    const double m10 = fma(dx, e1, fma(dy, f1, r10)), m11 = fma(dx, f1, fma(-dy, e1, fma(-d, g0.y, r11)));
                    const double det = fma(m00, m11, -m01 * m10);
                    const double idet = rcp_fast(det);
                    if (!abs_ge_hi(idet, kHi1e300))   // |idet| < ~1e300: warp-uniform; false for det = 0, inf or nan, where arma::inv throws (slam_library.cpp:270)
                    {
                        done = true;
                        const double i00 = m11 * idet, i01 = -m01 * idet, i10 = -m10 * idet, i11 = m00 * idet;
                        // (C) -Kt = -Pt Minv, x += Kt n
                        nk0 = fma(-P0, i00, -P1 * i10);
                        nk1 = fma(-P0, i01, -P1 * i11);
                        E.kt[s][lane] = make_double2(nk0, nk1);
                        // robot columns: Sigma -= Kt Wt restricted to them (the last use of Wt(0..2))
                        Ct = fma(nk0, g0.x, fma(nk1, g0.y, Ct));
                        Cx = fma(nk0, g1.x, fma(nk1, g1.y, Cx));
                        Cy = fma(nk0, g2.x, fma(nk1, g2.y, Cy));
                        x = fma(-nk0, n0, fma(-nk1, n1, x));
                        __syncwarp();
                        const double2 k0 = E.kt[s][0], k1 = E.kt[s][1], k2 = E.kt[s][2];
                        // replicated pose: what lanes 0..2 compute for their own x, evaluated identically by every lane
                        th = fma(-k0.x, n0, fma(-k0.y, n1, th));
                        px = fma(-k1.x, n0, fma(-k1.y, n1, px));
                        py = fma(-k2.x, n0, fma(-k2.y, n1, py));
                        if (abs_ge_hi(th, kHiPi)) th = wrap_angle(th);   // slam_library.cpp:275-276 (the identity inside [-pi, pi]); warp-uniform
                        if (lane == 0) x = th;
                        // robot rows: Sigma -= Kt Wt restricted to them
                        Rt = fma(k0.x, W0, fma(k0.y, W1, Rt));
                        Rx = fma(k1.x, W0, fma(k1.y, W1, Rx));
                        Ry = fma(k2.x, W0, fma(k2.y, W1, Ry));
                    }
                    else
                        status |= kStatusSingular;
                }
                else if (s == 0 && 2 * ch + 1 < NL && ((live_w >> (i + 1)) & 1u))
                    static_publish<NB>(C, Rt, Rx, Ry, Ct, Cx, Cy, E, g, t, lane, 1, i + 1);
                if (!done)
                {
                    // no measurement in this slot (or a singular one): it contributes nothing to the rank-4 pass
                    W0 = W1 = nk0 = nk1 = 0.0;
                    E.kt[s][lane] = make_double2(0.0, 0.0);
                    E.wt[s][lane] = make_double2(0.0, 0.0);
                    __syncwarp();
                }
                if (s == 0)
                {
                    pW0 = W0;
                    pW1 = W1;
                    pK0 = nk0;
                    pK1 = nk1;
                }
            }
            // (D) one DMMA pass applies the chunk to the fragments: C += (-Kt) Wt, k = (u0, u1, v0, v1)
            {
                const double * ka = reinterpret_cast<const double *>(&E.kt[t >> 1][3 + g]) + (t & 1);
                const double * wa = reinterpret_cast<const double *>(&E.wt[t >> 1][3 + g]) + (t & 1);
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
        }
        __syncwarp();

        // ---- write back: registers -> output image (orig order) in stage 1 -> one bulk store of the 16-byte aligned interior + one
        // plain store of the edge element ----
        double * gw = p.sigma + bf * SIG;
        const int odd = (int) ((reinterpret_cast<uintptr_t>(gw) >> 3) & 1);   // 1: HBM image starts 8 bytes past a 16-byte boundary
        double * oimg = reinterpret_cast<double *>(stage[1]) + odd;
        {
            int ro[NB], co[NB];
#pragma unroll
            for (int b = 0; b < NB; ++b)
            {
                ro[b] = 1 + 2 * perm_s[4 * b + (g >> 1)] + (g & 1);
                co[b] = (1 + 2 * perm_s[4 * b + t]) * LEN;
            }
#pragma unroll
            for (int br = 0; br < NB; ++br)
#pragma unroll
                for (int bc = 0; bc < NB; ++bc)
#pragma unroll
                    for (int e = 0; e < 2; ++e) oimg[co[bc] + e * LEN + ro[br]] = C[br][bc][e];
        }
        if (lane < LEN)
        {
            oimg[o * LEN] = Rt;
            oimg[o * LEN + 1] = Rx;
            oimg[o * LEN + 2] = Ry;
            if (lane >= 3)
            {
                oimg[o] = Ct;
                oimg[LEN + o] = Cx;
                oimg[2 * LEN + o] = Cy;
            }
            p.x[bf * LEN + o] = x;
            if (p.x_snap) p.x_snap[bf * LEN + o] = x;
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0)
        {
            constexpr int kInner = (SIG - 1) * 8;   // bytes of the aligned interior: SIG = (3 + 2n)^2 is odd, SIG - 1 elements = a multiple of 16 bytes
            static_assert((SIG & 1) == 1, "a filter's Sigma is an odd number of doubles");
            bulk_s2g(gw + odd, oimg + odd, kInner);
            const int edge = odd ? 0 : SIG - 1;
            gw[edge] = oimg[edge];
            if (status != st0) p.status[bf] = status;
        }
        __syncwarp();   // perm_s is rewritten by the next filter
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // shared memory must outlive the last bulk store
}

template <int N>
int launch_static_n(const EkfParams & p, bool do_predict, int sm_count, int32_t * worklist, int32_t * wl_count, cudaStream_t stream)
{
    int64_t blocks = p.batch;
    if (blocks > kStaticCtasPerSm * (int64_t) sm_count) blocks = kStaticCtasPerSm * (int64_t) sm_count;
    static bool configured_dev[kMaxDevices] = {false};
    bool & configured = configured_dev[device_slot()];
    if (!configured)
    {
        cudaFuncSetAttribute(k_ekf_static_step<N>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
        configured = true;
    }
    k_ekf_static_step<N><<<(unsigned) blocks, 32, 0, stream>>>(p, do_predict ? 1 : 0, worklist, wl_count);
    return (int) cudaGetLastError();
}

}   // namespace nuslam
