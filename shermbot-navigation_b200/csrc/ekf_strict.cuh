// ekf_strict.cuh -- STRICT arithmetic: one warp per filter, Sigma resident in shared memory.
//
// Reproduces, term by term, the operation order of the reference's dense Armadillo expressions
// (ascending k, one rounded multiply then one rounded add per term, chained products left to right;
// see oracle/shim/armadillo for the order the oracle defines) while visiting only the structurally
// non-zero terms:
//   predict : A = I + B has B(1,0), B(2,0) only (slam_library.cpp:133-146), so A*Sigma*A.t() + Q_bar
//             (:104) touches rows 1,2 then columns 1,2 then the 3x3 block            -> O(len)
//   update  : H_j has non-zeros in columns {0,1,2,c,c+1} (:175-183), so (I - K*H) has non-zeros on the
//             diagonal and in those 5 columns; (I-KH)*Sigma (:279) is 5 or 6 terms per element -> O(len^2)
// Skipped terms are exact zeros in the oracle (0 * finite), so every finite result is bit-identical
// up to the sign of zero.
#pragma once
#include "ekf_common.cuh"

namespace nuslam
{

// shared-memory footprint of one warp's filter, in doubles
__host__ __device__ constexpr int strict_smem_doubles(int len) { return len * len + 17 * len + 8; }

struct WarpFilter
{
    int len, n, lane;
    unsigned opt = 0u;   // NUSLAM_OPT_* (0 = the reference's behaviour)
    double * S;    // len x len, column-major
    double * x;    // len
    double * R5;   // 5 x len: rows {0,1,2,c,c+1} of Sigma before the update, R5[k*len + j]
    double * M5;   // 5 x len: columns {0,1,2,c,c+1} of (I - K*H), M5[k*len + i]
    double * G;    // 2 x len: H*Sigma, G[a*len + j]
    double * K;    // 2 x len (debug / future use)

    __device__ __forceinline__ void carve(double * base, int len_, int n_, int lane_)
    {
        len = len_;
        n = n_;
        lane = lane_;
        S = base;
        x = S + len * len;
        R5 = x + len;
        M5 = R5 + 5 * len;
        G = M5 + 5 * len;
        K = G + 2 * len;
    }

    __device__ __forceinline__ void load(const double * gx, const double * gS)
    {
        for (int e = lane; e < len * len; e += kWarp) S[e] = gS[e];
        for (int e = lane; e < len; e += kWarp) x[e] = gx[e];
        __syncwarp();
    }

    __device__ __forceinline__ void store(double * gx, double * gS, bool sigma)
    {
        __syncwarp();
        if (sigma)
            for (int e = lane; e < len * len; e += kWarp) gS[e] = S[e];
        for (int e = lane; e < len; e += kWarp) gx[e] = x[e];
    }

    // ExtendedKalman::predict, slam_library.cpp:65-69
    __device__ void predict(double dth, double dx, const double * Q)
    {
        // predictEstimate :71-94 (every lane evaluates the same scalars)
        const double theta = x[0];
        double dq_th, dq_x, dq_y;
        if (dth == 0.0)
        {
            dq_th = 0.0;
            dq_x = mul_(dx, cos(theta));
            dq_y = mul_(dx, sin(theta));
        }
        else
        {
            const double q = div_(dx, dth);
            dq_th = dth;
            dq_x = add_(mul_(-q, sin(theta)), mul_(q, sin(add_(theta, dth))));
            dq_y = sub_(mul_(q, cos(theta)), mul_(q, cos(add_(theta, dth))));
        }
        const double th1 = add_(theta, dq_th);
        const double x1 = add_(x[1], dq_x);
        const double y1 = add_(x[2], dq_y);
        __syncwarp();
        if (lane == 0)
        {
            x[0] = th1;
            x[1] = x1;
            x[2] = y1;
        }
        // getA :127-148 -- theta read AFTER predictEstimate (:129); the opt-in variant linearises where the motion started
        const double thJ = (opt & kOptPreMotionJacobian) ? theta : th1;
        double b10, b20;
        if (dth == 0.0)
        {
            b10 = mul_(-dx, sin(thJ));
            b20 = mul_(dx, cos(thJ));
        }
        else
        {
            const double q = div_(dx, dth);
            b10 = add_(mul_(-q, cos(thJ)), mul_(q, cos(add_(thJ, dth))));
            b20 = add_(mul_(-q, sin(thJ)), mul_(q, sin(add_(thJ, dth))));
        }
        // T = A * Sigma: rows 1 and 2 (k = 0 term first, then the unit diagonal term)
        for (int j = lane; j < len; j += kWarp)
        {
            const double s0 = S[0 + j * len];
            S[1 + j * len] = add_(mul_(b10, s0), S[1 + j * len]);
            S[2 + j * len] = add_(mul_(b20, s0), S[2 + j * len]);
        }
        __syncwarp();
        // U = T * A.t(): columns 1 and 2
        for (int i = lane; i < len; i += kWarp)
        {
            const double t0 = S[i + 0 * len];
            S[i + 1 * len] = add_(mul_(t0, b10), S[i + 1 * len]);
            S[i + 2 * len] = add_(mul_(t0, b20), S[i + 2 * len]);
        }
        __syncwarp();
        // + Q_bar (expanded_process_noise :110-125): only the robot block is non-zero
        if (lane < 9)
        {
            const int r = lane % 3, cc = lane / 3;
            S[r + cc * len] = add_(S[r + cc * len], Q[r + 3 * cc]);
        }
        __syncwarp();
    }

    // ExtendedKalman::initializeLandmark, slam_library.cpp:255-261
    __device__ void init_landmark(double z0, double z1, int id, int & status)
    {
        if (id < 1 || id > n)
        {
            status |= kStatusBadId;
            return;
        }
        const int c = 3 + 2 * (id - 1);
        const double a = add_(z1, x[0]);
        const double mx = add_(x[1], mul_(z0, cos(a)));
        const double my = add_(x[2], mul_(z0, sin(a)));
        __syncwarp();
        if (lane == 0)
        {
            x[c] = mx;
            x[c + 1] = my;
        }
        __syncwarp();
    }

    // Mahalanobis distance of z to landmark k at the current state: slam_library.cpp:212-232
    __device__ double mahalanobis(double z0, double z1, int k, const double * R, bool & singular)
    {
        const int c = 3 + 2 * (k - 1);
        HEntries H;
        double zr, zb;
        measurement_model(x, c, H, zr, zb);
        const int cols[5] = {0, 1, 2, c, c + 1};
        double g0[5], g1[5];   // (H*Sigma) at the 5 columns where H is non-zero
#pragma unroll
        for (int q = 0; q < 5; ++q)
        {
            const int j = cols[q];
            const double s0 = S[0 + j * len], s1 = S[1 + j * len], s2 = S[2 + j * len];
            const double sc = S[c + j * len], sc1 = S[c + 1 + j * len];
            double a0 = mul_(H.h01, s1);
            a0 = add_(a0, mul_(H.h02, s2));
            a0 = add_(a0, mul_(H.h0c, sc));
            a0 = add_(a0, mul_(H.h0c1, sc1));
            double a1 = -s0;
            a1 = add_(a1, mul_(H.h11, s1));
            a1 = add_(a1, mul_(H.h12, s2));
            a1 = add_(a1, mul_(H.h1c, sc));
            a1 = add_(a1, mul_(H.h1c1, sc1));
            g0[q] = a0;
            g1[q] = a1;
        }
        // psi = (H*Sigma)*H.t() + R
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
        double i00, i01, i10, i11;
        if (!inv2x2(p00, p01, p10, p11, i00, i01, i10, i11))
        {
            singular = true;
            return 0.0;
        }
        double dz1 = sub_(z1, zb);   // no angle wrap (:229-231) unless opted in
        if (opt & kOptWrapInnovation) dz1 = normalize_angle(dz1);
        const double dz0 = sub_(z0, zr);
        // (dz.t() * psi.i()) * dz
        const double t0 = add_(mul_(dz0, i00), mul_(dz1, i10));
        const double t1 = add_(mul_(dz0, i01), mul_(dz1, i11));
        return add_(mul_(t0, dz0), mul_(t1, dz1));
    }

    // ExtendedKalman::associateLandmark, slam_library.cpp:188-253. One candidate landmark per lane; the
    // reference's in-order early exit becomes "lowest lane whose distance decides".
    __device__ int associate(double z0, double z1, int & seen, int & status, const double * R, double amin, double amax)
    {
        if (seen == 0)
        {
            seen = 1;
            return 1;
        }
        if (3 + 2 * seen >= len)   // temp(3+2*seen) out of bounds: Armadillo throws (:206)
        {
            status |= kStatusMapFull;
            return kIdException;
        }
        for (int base = 0; base < seen; base += kWarp)
        {
            const int k = base + lane + 1;
            const bool active = k <= seen;
            bool singular = false;
            double d = 0.0;
            if (active) d = mahalanobis(z0, z1, k, R, singular);
            const unsigned sing = __ballot_sync(0xffffffffu, active && singular);
            const unsigned hitA = __ballot_sync(0xffffffffu, active && !singular && (d < amin));
            const unsigned hitB = __ballot_sync(0xffffffffu, active && !singular && (d > amin) && (d < amax));
            const unsigned any = hitA | hitB | sing;
            if (any)
            {
                const int first = __ffs(any) - 1;
                if ((sing >> first) & 1u)
                {
                    status |= kStatusSingular;
                    return kIdException;
                }
                if ((hitA >> first) & 1u) return base + first + 1;
                return -1;
            }
        }
        seen += 1;
        return seen;
    }

    // ExtendedKalman::update, slam_library.cpp:263-282
    __device__ void update(double z0, double z1, int id, int & status, const double * R)
    {
        if (id < 1 || id > n)
        {
            status |= kStatusBadId;
            return;
        }
        const int c = 3 + 2 * (id - 1);
        HEntries H;
        double zr, zb;
        measurement_model(x, c, H, zr, zb);
        // G = H * Sigma (2 x len); keep the 5 rows of Sigma the update reads
        for (int j = lane; j < len; j += kWarp)
        {
            const double s0 = S[0 + j * len], s1 = S[1 + j * len], s2 = S[2 + j * len];
            const double sc = S[c + j * len], sc1 = S[c + 1 + j * len];
            R5[0 * len + j] = s0;
            R5[1 * len + j] = s1;
            R5[2 * len + j] = s2;
            R5[3 * len + j] = sc;
            R5[4 * len + j] = sc1;
            double a0 = mul_(H.h01, s1);
            a0 = add_(a0, mul_(H.h02, s2));
            a0 = add_(a0, mul_(H.h0c, sc));
            a0 = add_(a0, mul_(H.h0c1, sc1));
            double a1 = -s0;
            a1 = add_(a1, mul_(H.h11, s1));
            a1 = add_(a1, mul_(H.h12, s2));
            a1 = add_(a1, mul_(H.h1c, sc));
            a1 = add_(a1, mul_(H.h1c1, sc1));
            G[0 * len + j] = a0;
            G[1 * len + j] = a1;
        }
        __syncwarp();
        // psi = G * H.t() + R, inverse
        double p00 = mul_(G[1], H.h01);
        p00 = add_(p00, mul_(G[2], H.h02));
        p00 = add_(p00, mul_(G[c], H.h0c));
        p00 = add_(p00, mul_(G[c + 1], H.h0c1));
        double p10 = mul_(G[len + 1], H.h01);
        p10 = add_(p10, mul_(G[len + 2], H.h02));
        p10 = add_(p10, mul_(G[len + c], H.h0c));
        p10 = add_(p10, mul_(G[len + c + 1], H.h0c1));
        double p01 = -G[0];
        p01 = add_(p01, mul_(G[1], H.h11));
        p01 = add_(p01, mul_(G[2], H.h12));
        p01 = add_(p01, mul_(G[c], H.h1c));
        p01 = add_(p01, mul_(G[c + 1], H.h1c1));
        double p11 = -G[len + 0];
        p11 = add_(p11, mul_(G[len + 1], H.h11));
        p11 = add_(p11, mul_(G[len + 2], H.h12));
        p11 = add_(p11, mul_(G[len + c], H.h1c));
        p11 = add_(p11, mul_(G[len + c + 1], H.h1c1));
        p00 = add_(p00, R[0]);
        p10 = add_(p10, R[1]);
        p01 = add_(p01, R[2]);
        p11 = add_(p11, R[3]);
        double i00, i01, i10, i11;
        if (!inv2x2(p00, p01, p10, p11, i00, i01, i10, i11))
        {
            status |= kStatusSingular;   // arma::inv throws; the update never happens
            __syncwarp();
            return;
        }
        double dz1 = sub_(z1, zb);   // :272, no wrap unless opted in
        if (opt & kOptWrapInnovation) dz1 = normalize_angle(dz1);
        const double dz0 = sub_(z0, zr);
        // P = Sigma * H.t(); K = P * inv(psi); x += K * dz; columns of M = I - K*H
        for (int i = lane; i < len; i += kWarp)
        {
            const double s0 = S[i + 0 * len], s1 = S[i + 1 * len], s2 = S[i + 2 * len];
            const double sc = S[i + c * len], sc1 = S[i + (c + 1) * len];
            double pa = mul_(s1, H.h01);
            pa = add_(pa, mul_(s2, H.h02));
            pa = add_(pa, mul_(sc, H.h0c));
            pa = add_(pa, mul_(sc1, H.h0c1));
            double pb = -s0;
            pb = add_(pb, mul_(s1, H.h11));
            pb = add_(pb, mul_(s2, H.h12));
            pb = add_(pb, mul_(sc, H.h1c));
            pb = add_(pb, mul_(sc1, H.h1c1));
            const double k0 = add_(mul_(pa, i00), mul_(pb, i10));
            const double k1 = add_(mul_(pa, i01), mul_(pb, i11));
            x[i] = add_(x[i], add_(mul_(k0, dz0), mul_(k1, dz1)));   // :275
            if (opt & kOptJoseph)
            {
                K[0 * len + i] = k0;
                K[1 * len + i] = k1;
            }
            // (K*H)(i,j) = K(i,0)*H(0,j) + K(i,1)*H(1,j);  M = eye - K*H
            const double kh0 = -k1;   // K(i,0)*0 + K(i,1)*(-1)
            const double kh1 = add_(mul_(k0, H.h01), mul_(k1, H.h11));
            const double kh2 = add_(mul_(k0, H.h02), mul_(k1, H.h12));
            const double khc = add_(mul_(k0, H.h0c), mul_(k1, H.h1c));
            const double khc1 = add_(mul_(k0, H.h0c1), mul_(k1, H.h1c1));
            M5[0 * len + i] = sub_((i == 0) ? 1.0 : 0.0, kh0);
            M5[1 * len + i] = sub_((i == 1) ? 1.0 : 0.0, kh1);
            M5[2 * len + i] = sub_((i == 2) ? 1.0 : 0.0, kh2);
            M5[3 * len + i] = sub_((i == c) ? 1.0 : 0.0, khc);
            M5[4 * len + i] = sub_((i == c + 1) ? 1.0 : 0.0, khc1);
        }
        __syncwarp();
        if (lane == 0) x[0] = normalize_angle(x[0]);   // :276
        // Sigma = M * Sigma (:279): ascending-k merge of {0,1,2,c,c+1} with the unit diagonal term k = i. Every element is an
        // independent expression, so the traversal is free: lane = row i (its five M entries stay in registers), loop over the
        // columns j (the five R5 entries of a column are one broadcast read each) -- a third of the shared-memory traffic of a
        // flat element loop, the same arithmetic per element.
        const int len2 = len * len;
        for (int i0 = 0; i0 < len; i0 += kWarp)
        {
            const int i = i0 + lane;
            if (i < len)
            {
                const double m0 = M5[0 * len + i], m1 = M5[1 * len + i], m2 = M5[2 * len + i], m3 = M5[3 * len + i], m4 = M5[4 * len + i];
                const bool mid = i >= 3 && i < c, hi = i > c + 1;
                for (int j = 0; j < len; ++j)
                {
                    const int e = i + j * len;
                    double acc = mul_(m0, R5[0 * len + j]);
                    acc = add_(acc, mul_(m1, R5[1 * len + j]));
                    acc = add_(acc, mul_(m2, R5[2 * len + j]));
                    if (mid) acc = add_(acc, S[e]);
                    acc = add_(acc, mul_(m3, R5[3 * len + j]));
                    acc = add_(acc, mul_(m4, R5[4 * len + j]));
                    if (hi) acc = add_(acc, S[e]);
                    S[e] = acc;
                }
            }
        }
        __syncwarp();
        if (opt & kOptJoseph)
        {
            // Joseph form on top of T = (I - K H) Sigma (now in S): Sigma' = T (I - K H)^T + K R K^T = T - (T H^T) K^T + K R K^T,
            // then the symmetric part. T H^T (len x 2) goes to the first two rows of R5 (dead after the pass above).
            for (int i = lane; i < len; i += kWarp)
            {
                const double t0 = S[i + 0 * len], t1 = S[i + 1 * len], t2 = S[i + 2 * len];
                const double tc = S[i + c * len], tc1 = S[i + (c + 1) * len];
                R5[0 * len + i] = fma(t1, H.h01, fma(t2, H.h02, fma(tc, H.h0c, tc1 * H.h0c1)));
                R5[1 * len + i] = fma(t1, H.h11, fma(t2, H.h12, fma(tc, H.h1c, fma(tc1, H.h1c1, -t0))));
            }
            __syncwarp();
            for (int e = lane; e < len2; e += kWarp)
            {
                const int j = e / len;
                const int i = e - j * len;
                const double k0j = K[0 * len + j], k1j = K[1 * len + j];
                const double rk0 = fma(R[0], k0j, R[2] * k1j), rk1 = fma(R[1], k0j, R[3] * k1j);   // (R K^T)(:, j), R column-major
                double acc = S[e];
                acc = fma(-R5[0 * len + i], k0j, acc);
                acc = fma(-R5[1 * len + i], k1j, acc);
                acc = fma(K[0 * len + i], rk0, acc);
                acc = fma(K[1 * len + i], rk1, acc);
                S[e] = acc;
            }
            __syncwarp();
            for (int e = lane; e < len2; e += kWarp)
            {
                const int j = e / len;
                const int i = e - j * len;
                if (i < j)
                {
                    const double a = 0.5 * (S[i + j * len] + S[j + i * len]);
                    S[i + j * len] = a;
                    S[j + i * len] = a;
                }
            }
            __syncwarp();
        }
    }
};

enum EkfOp
{
    kOpPredict = 0,
    kOpUpdate = 1,
    kOpAssociate = 2,
    kOpInit = 3,
    kOpStep = 4
};

struct EkfParams
{
    int64_t batch;
    int len, n, m;
    double * x;          // B x len
    double * sigma;      // B x len x len, column-major per filter
    int32_t * seen;      // B
    int32_t * status;    // B
    const double * twists;   // B x 3
    const double * z;        // B x m x 2
    const int32_t * ids;     // B x m or null
    unsigned options;        // NUSLAM_OPT_* (0 = the reference's behaviour)
    const int32_t * m_valid; // B or null: filter b uses only its first m_valid[b] (<= m) measurements (fused scan step)
    int32_t * ids_out;       // B x m or null
    double * x_snap;         // B x len or null: a second copy of the state vector after the step (pipelined host path: the snapshot the
                             // device -> host copy reads while the next step already runs)
    double Q[9], R[4];
    double amin, amax;
    // launch shape of k_ekf_strict_list for the filters a FAST kernel hands over (strict_tail below); tail_blocks == 0: the host launches it
    int tail_blocks, tail_threads, tail_smem;
};

// One warp replays reference call OP on filter b; Sigma staged through the warp's slice of shared memory.
template <int OP>
__device__ __forceinline__ void strict_filter(const EkfParams & p, const int64_t b, double * smem_warp, const int lane)
{
    WarpFilter f;
    f.carve(smem_warp, p.len, p.n, lane);
    f.opt = p.options;
    double * gx = p.x + b * p.len;
    double * gS = p.sigma + b * (int64_t) p.len * p.len;
    int status = p.status[b];
    int seen = p.seen[b];

    if (OP == kOpInit)
    {
        const int id = p.ids[b];
        if (id <= 0) return;
        for (int e = lane; e < p.len; e += kWarp) f.x[e] = gx[e];
        __syncwarp();
        f.init_landmark(p.z[2 * b], p.z[2 * b + 1], id, status);
        f.store(gx, gS, false);
        if (lane == 0) p.status[b] = status;
        return;
    }

    f.load(gx, gS);
    if (OP == kOpPredict)
    {
        f.predict(p.twists[3 * b], p.twists[3 * b + 1], p.Q);
        f.store(gx, gS, true);
    }
    else if (OP == kOpUpdate)
    {
        const int id = p.ids[b];
        if (id <= 0) return;
        f.update(p.z[2 * b], p.z[2 * b + 1], id, status, p.R);
        f.store(gx, gS, true);
        if (lane == 0) p.status[b] = status;
    }
    else if (OP == kOpAssociate)
    {
        const int id = f.associate(p.z[2 * b], p.z[2 * b + 1], seen, status, p.R, p.amin, p.amax);
        if (lane == 0)
        {
            p.ids_out[b] = id;
            p.seen[b] = seen;
            p.status[b] = status;
        }
    }
    else if (OP == kOpStep)
    {
        // EKFSlam::main_loop, nuslam/src/slam.cpp:262-319
        const int64_t mb = b * p.m;
        if (status & (kStatusMapFull | kStatusSingular))   // the reference process died on an earlier scan
        {
            if (p.ids_out)
                for (int i = lane; i < p.m; i += kWarp) p.ids_out[mb + i] = 0;
            if (p.x_snap)
                for (int e = lane; e < p.len; e += kWarp) p.x_snap[b * p.len + e] = f.x[e];
            return;
        }
        const int seen_snapshot = seen;                               // slam.cpp:251
        f.predict(p.twists[3 * b], p.twists[3 * b + 1], p.Q);         // slam.cpp:269
        const int mv = p.m_valid ? min(p.m, max(0, p.m_valid[b])) : p.m;   // markers this filter received (landmarks.cpp:84-109)
        if (p.ids_out)
            for (int i = mv + lane; i < p.m; i += kWarp) p.ids_out[mb + i] = 0;
        for (int i = 0; i < mv; ++i)                                  // slam.cpp:279
        {
            const double z0 = p.z[2 * (mb + i)], z1 = p.z[2 * (mb + i) + 1];
            int id;
            if (p.ids)
            {
                id = p.ids[mb + i];
                if (id <= 0)
                {
                    if (p.ids_out && lane == 0) p.ids_out[mb + i] = 0;
                    continue;
                }
                if (id > p.n)
                {
                    // not a landmark of this map: flagged and skipped like the FAST and LARGE paths do, `seen` untouched (the reference
                    // would index its state out of bounds)
                    status |= kStatusBadId;
                    if (p.ids_out && lane == 0) p.ids_out[mb + i] = id;
                    continue;
                }
                if (id > seen) seen = id;
            }
            else
            {
                id = f.associate(z0, z1, seen, status, p.R, p.amin, p.amax);   // slam.cpp:291
                if (id == kIdException)
                {
                    if (p.ids_out)
                        for (int r = i + lane; r < mv; r += kWarp) p.ids_out[mb + r] = (r == i) ? kIdException : 0;
                    break;
                }
            }
            if (p.ids_out && lane == 0) p.ids_out[mb + i] = id;
            if (id > seen_snapshot) f.init_landmark(z0, z1, id, status);   // slam.cpp:295-297
            else if (id < 0) continue;                                      // slam.cpp:298-300
            f.update(z0, z1, id, status, p.R);                              // slam.cpp:318
        }
        f.store(gx, gS, true);
        if (p.x_snap)
            for (int e = lane; e < p.len; e += kWarp) p.x_snap[b * p.len + e] = f.x[e];
        if (lane == 0)
        {
            p.seen[b] = seen;
            p.status[b] = status;
        }
    }
}

// One warp per filter over the whole batch.
template <int OP>
__global__ void __launch_bounds__(128) k_ekf_strict(const EkfParams p)
{
    extern __shared__ double smem[];
    const int warp = threadIdx.x / kWarp, lane = threadIdx.x % kWarp;
    const int64_t b = (int64_t) blockIdx.x * (blockDim.x / kWarp) + warp;
    if (b >= p.batch) return;
    strict_filter<OP>(p, b, smem + (size_t) warp * strict_smem_doubles(p.len), lane);
}

// One warp per filter over a device-side work list (the filters the FAST kernel handed over because their step
// contains a landmark's first touch). The last block to finish resets the list for the next launch.
// TU: the translation unit that instantiates it (0 nuslam_b200.cu, whole-program; 1 ekf_fast_tu.cu, relocatable): the two builds of the same
// kernel must not share a name, or the linker would merge their host stubs.
template <int OP, int TU>
__global__ void __launch_bounds__(128) k_ekf_strict_list(const EkfParams p, const int32_t * __restrict__ list, int32_t * count, int32_t * done)
{
    extern __shared__ double smem[];
    const int warps = blockDim.x / kWarp;
    const int warp = threadIdx.x / kWarp, lane = threadIdx.x % kWarp;
    const int n = *reinterpret_cast<volatile int32_t *>(count);
    for (int64_t k = (int64_t) blockIdx.x * warps + warp; k < n; k += (int64_t) gridDim.x * warps)
    {
        strict_filter<OP>(p, list[k], smem + (size_t) warp * strict_smem_doubles(p.len), lane);
        __syncwarp();
    }
    __syncthreads();
    if (threadIdx.x == 0)
    {
        __threadfence();
        if (atomicAdd(done, 1) == (int) gridDim.x - 1)
        {
            *count = 0;
            *done = 0;
            __threadfence();
        }
    }
}

#ifndef NUSLAM_TAIL_LAUNCH
#define NUSLAM_TAIL_LAUNCH 0    // 1: built with -rdc=true + cudadevrt; the FAST kernels launch the list kernel themselves, and only when needed
#endif

// Every warp of a FAST kernel calls this when it has finished its filters. The LAST warp of the grid looks at the work list and, if a
// filter was handed over, launches k_ekf_strict_list into the tail-launch stream: it runs after this grid has completed and before the
// next kernel of the host stream starts. A step without a first touch -- every step of a built map -- is ONE launch.
// wl_count: [0] entries, [1] finished blocks of the list kernel, [2] finished warps of the FAST kernel, [3] sticky device-launch error
__device__ __forceinline__ void strict_tail(const EkfParams & p, const int do_predict, int32_t * worklist, int32_t * wl_count, const int total_warps,
                                            const int lane)
{
#if NUSLAM_TAIL_LAUNCH && defined(NUSLAM_TU_FAST)   // (a device-side launch compiles only as relocatable device code)
    if (p.tail_blocks == 0) return;
    __syncwarp();
    if (lane == 0)
    {
        __threadfence();
        if (atomicAdd(wl_count + 2, 1) == total_warps - 1)
        {
            wl_count[2] = 0;
            __threadfence();
            if (*reinterpret_cast<volatile int32_t *>(wl_count) > 0)
            {
                // (inlined on purpose: an out-of-line launcher takes the parameter block through the stack and costs the kernels registers)
                if (do_predict)
                    k_ekf_strict_list<kOpStep, 1><<<p.tail_blocks, p.tail_threads, p.tail_smem, cudaStreamTailLaunch>>>(p, worklist, wl_count, wl_count + 1);
                else
                    k_ekf_strict_list<kOpUpdate, 1><<<p.tail_blocks, p.tail_threads, p.tail_smem, cudaStreamTailLaunch>>>(p, worklist, wl_count, wl_count + 1);
                const cudaError_t e = cudaGetLastError();
                if (e != cudaSuccess)
                {
                    wl_count[3] = (int32_t) e;   // reported by the next synchronize
                    wl_count[0] = 0;             // the next step must not replay this step's entries
                }
            }
        }
    }
#endif
}

}   // namespace nuslam
