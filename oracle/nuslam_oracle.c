/* oracle/nuslam_oracle.c -- TEST INFRASTRUCTURE ONLY (see oracle/oracle_api.h).
 *
 * Plain-C restatement of the reference's EKF-SLAM + circle-fit hot path. It is the CHECKER for the
 * CUDA path: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load it; nothing under shermbot-navigation_b200/ does.
 *
 * Pinning: the EKF functions have no golden vectors in the reference; this file is
 * pinned by bit-for-bit agreement with oracle/_ref/libnuslam_ref.so (the unmodified reference
 * sources compiled against oracle/shim/) in tests/test_oracle.py, and circleFit additionally by the
 * two known-answer tests of nuslam/tests/circle_tests.cpp:38-40,67-69.
 *
 * Dense products follow the arithmetic order documented in oracle/shim/armadillo (Armadillo is an
 * un-vendored, un-pinned dependency of the reference): ascending k, unfused multiply then add,
 * left-to-right evaluation of chained products. Build with -ffp-contract=off.
 *
 * Every function cites the reference lines it restates (paths relative to /root/reference).
 */
#include <limits.h>
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include "oracle_api.h"

#ifdef ORC_DETMATH
#include "detmath.h"
#define ORC_SIN(x) dm_sin(x)
#define ORC_COS(x) dm_cos(x)
#define ORC_ATAN2(y, x) dm_atan2(y, x)
#define ORC_FLAVOUR_STR "port+detmath"
#else
#define ORC_SIN(x) sin(x)
#define ORC_COS(x) cos(x)
#define ORC_ATAN2(y, x) atan2(y, x)
#define ORC_FLAVOUR_STR "port"
#endif

#define ORC_PI 3.14159265358979323846 /* rigid2d/include/rigid2d/rigid2d.hpp:16 */

const char * orc_flavour(void) { return ORC_FLAVOUR_STR; }
int orc_eig_sym_full_sweeps(int on) { (void) on; return 1; }   /* the restatement always runs every sweep */

/* ------------------------------------------------------------------ dense helpers (shim order) */

/* C(r x c) = A(r x k) * B(k x c), all column-major; C must not alias A or B */
static void mm(double * C, const double * A, const double * B, int r, int k, int c)
{
    for (int j = 0; j < c; ++j)
        for (int i = 0; i < r; ++i)
        {
            double acc = 0.0;
            for (int p = 0; p < k; ++p)
            {
                const double prod = A[i + p * r] * B[p + j * k];
                acc = acc + prod;
            }
            C[i + j * r] = acc;
        }
}

static void transpose(double * T, const double * A, int r, int c) /* T is c x r */
{
    for (int j = 0; j < c; ++j)
        for (int i = 0; i < r; ++i) T[j + i * c] = A[i + j * r];
}

/* closed-form 2x2 inverse (oracle/shim/armadillo inv()); returns 0 when singular */
static int inv2(double * out, const double * a)
{
    const double p = a[0], q = a[2], r = a[1], s = a[3];
    const double ps = p * s;
    const double qr = q * r;
    const double det = ps - qr;
    if (det == 0.0) return 0;
    out[0] = s / det;
    out[2] = -q / det;
    out[1] = -r / det;
    out[3] = p / det;
    return 1;
}

/* ------------------------------------------------------------------ rigid2d slice */

/* rigid2d/src/rigid2d.cpp:9-13 */
double orc_normalize_angle(double rad) { return ORC_ATAN2(ORC_SIN(rad), ORC_COS(rad)); }

static double deg2rad_(double deg) { return (ORC_PI / (double) 180) * deg; } /* rigid2d.hpp:40-44 */
static double rad2deg_(double rad) { return ((double) 180 / ORC_PI) * rad; } /* rigid2d.hpp:49-53 */

typedef struct
{
    double c, s, x, y;
} tf2d; /* rigid2d.hpp:165-171 */

/* rigid2d.cpp:198-209: lhs *= rhs */
static tf2d tf_mul(tf2d l, tf2d r)
{
    tf2d o;
    o.c = (l.c * r.c) - (l.s * r.s);
    o.s = (l.s * r.c) + (l.c * r.s);
    o.x = (l.c * r.x) - (l.s * r.y) + l.x;
    o.y = (l.s * r.x) + (l.c * r.y) + l.y;
    return o;
}

/* rigid2d.cpp:187-196 */
static tf2d tf_inv(tf2d t)
{
    tf2d o;
    o.c = t.c;
    o.s = -t.s;
    o.x = (-t.x * t.c) + (-t.y * t.s);
    o.y = (t.x * t.s) + (-t.y * t.c);
    return o;
}

/* rigid2d.cpp:294-328 */
static tf2d integrate_twist(double dth, double dx, double dy)
{
    tf2d out;
    if (dth == 0)
    {
        out.c = 1;
        out.s = 0;
        out.x = dx;
        out.y = dy;
        return out;
    }
    tf2d T_sb = {1, 0, dy / dth, -(dx / dth)};
    tf2d T_ss = {ORC_COS(dth), ORC_SIN(dth), 0, 0};
    tf2d T_bs = tf_inv(T_sb);
    return tf_mul(tf_mul(T_bs, T_ss), T_sb); /* operator* is left-associative (:211-214,325) */
}

void orc_integrate_twist(double dth, double dx, double dy, double * o)
{
    tf2d t = integrate_twist(dth, dx, dy);
    o[0] = t.c;
    o[1] = t.s;
    o[2] = t.x;
    o[3] = t.y;
}

/* rigid2d/src/diff_drive.cpp:66-78 */
void orc_diffdrive_convert_twist(double base, double rad, double dth, double dx, double * u)
{
    const double d = base / 2;
    const double r = rad;
    u[0] = (-(d / r) * dth) + (dx / r);
    u[1] = ((d / r) * dth) + (dx / r);
}

/* getTwist (diff_drive.cpp:80-110) then operator() (:111-146), as slam.cpp:264-265 calls them */
void orc_diffdrive_step(double * s, double thLnew, double thRnew, double * tw)
{
    const double wheelBase = s[0], wheelRad = s[1];
    const double dUL = thLnew - s[5];
    const double dUR = thRnew - s[6];
    tw[0] = (wheelRad / wheelBase) * (dUR - dUL);
    tw[1] = (wheelRad / 2) * (dUL + dUR);
    tw[2] = 0.0;
    tf2d Tbb = integrate_twist(tw[0], tw[1], tw[2]);
    const double dqb_th = atan(Tbb.s / Tbb.c); /* :129 */
    const double dqb_x = Tbb.x, dqb_y = Tbb.y;
    /* adj = Transform2D(th) (:134); adj(dqb) = rigid2d.cpp:254-261 with x = y = 0 */
    const double c = ORC_COS(s[4]), sn = ORC_SIN(s[4]);
    const double dq_th = dqb_th;
    const double dq_x = (0.0 * dqb_th) + (c * dqb_x) - (sn * dqb_y);
    const double dq_y = -(0.0 * dqb_th) + (sn * dqb_x) + (c * dqb_y);
    s[4] += dq_th;
    s[2] += dq_x;
    s[3] += dq_y;
    s[5] = thLnew;
    s[6] = thRnew;
}

/* ------------------------------------------------------------------ EKF (nuslam/src/slam_library.cpp) */

typedef struct
{
    int n, len, seen;
    double * x;   /* len */
    double * S;   /* len x len column-major */
    double Q[9];  /* 3x3 column-major */
    double R[4];  /* 2x2 column-major */
    double * w0, * w1, * w2, * w3; /* len x len scratch */
} okf;

/* slam_library.cpp:16-22 */
void orc_cartesian2polar(double x, double y, double * rb)
{
    rb[0] = sqrt(x * x + (y * y));
    rb[1] = orc_normalize_angle(ORC_ATAN2(y, x));
}

/* slam_library.cpp:39-63 + initCov :24-33 */
void * orc_ekf_new(int n, const double * robot3, const double * map2n, const double * Q9, const double * R4)
{
    okf * f = (okf *) calloc(1, sizeof(okf));
    f->n = n;
    f->len = 3 + 2 * n;
    f->seen = 0;
    const size_t l2 = (size_t) f->len * f->len;
    f->x = (double *) calloc(f->len, sizeof(double));
    f->S = (double *) calloc(l2, sizeof(double));
    f->w0 = (double *) calloc(l2, sizeof(double));
    f->w1 = (double *) calloc(l2, sizeof(double));
    f->w2 = (double *) calloc(l2, sizeof(double));
    f->w3 = (double *) calloc(l2, sizeof(double));
    memcpy(f->Q, Q9, sizeof(f->Q));
    memcpy(f->R, R4, sizeof(f->R));
    for (int i = 0; i < 3; ++i) f->x[i] = robot3[i];
    for (int i = 3; i < f->len; ++i) f->x[i] = map2n[i - 3];
    for (int i = 3; i < f->len; ++i) f->S[i + (size_t) i * f->len] = INT_MAX;
    return f;
}

void orc_ekf_free(void * h)
{
    okf * f = (okf *) h;
    if (!f) return;
    free(f->x);
    free(f->S);
    free(f->w0);
    free(f->w1);
    free(f->w2);
    free(f->w3);
    free(f);
}

void orc_ekf_get(void * h, double * x, double * sigma, int * seen)
{
    okf * f = (okf *) h;
    if (x) memcpy(x, f->x, sizeof(double) * f->len);
    if (sigma) memcpy(sigma, f->S, sizeof(double) * f->len * f->len);
    if (seen) *seen = f->seen;
}

void orc_ekf_set(void * h, const double * x, const double * sigma, int seen)
{
    okf * f = (okf *) h;
    if (x) memcpy(f->x, x, sizeof(double) * f->len);
    if (sigma) memcpy(f->S, sigma, sizeof(double) * f->len * f->len);
    f->seen = seen;
}

/* predict :65-69 = predictEstimate :71-94 then propagateUncertainty :96-108 */
void orc_ekf_predict(void * h, double dth, double dx, double dy)
{
    (void) dy; /* tw.dy never read by the filter */
    okf * f = (okf *) h;
    const int len = f->len;
    double dq_th, dq_x, dq_y;
    double theta = f->x[0];
    if (dth == 0.0) /* :77 */
    {
        dq_th = 0.0;
        dq_x = dx * ORC_COS(theta);
        dq_y = dx * ORC_SIN(theta);
    }
    else
    {
        dq_th = dth;
        dq_x = -(dx / dth) * ORC_SIN(theta) + (dx / dth) * ORC_SIN(theta + dth);
        dq_y = (dx / dth) * ORC_COS(theta) - (dx / dth) * ORC_COS(theta + dth);
    }
    f->x[0] += dq_th;
    f->x[1] += dq_x;
    f->x[2] += dq_y;

    /* getA :127-148 -- theta is read AFTER predictEstimate (:129) */
    theta = f->x[0];
    double * A = f->w0;
    memset(A, 0, sizeof(double) * len * len);
    double b10, b20;
    if (dth == 0)
    {
        b10 = -dx * ORC_SIN(theta);
        b20 = dx * ORC_COS(theta);
    }
    else
    {
        b10 = -(dx / dth) * ORC_COS(theta) + (dx / dth) * ORC_COS(theta + dth);
        b20 = -(dx / dth) * ORC_SIN(theta) + (dx / dth) * ORC_SIN(theta + dth);
    }
    for (int i = 0; i < len; ++i) A[i + i * len] = 1.0 + 0.0; /* I + B */
    A[1 + 0 * len] = 0.0 + b10;
    A[2 + 0 * len] = 0.0 + b20;

    /* :104  A * covariance * A.t() + Q_bar */
    double * AS = f->w1, * At = f->w2, * U = f->w3;
    mm(AS, A, f->S, len, len, len);
    transpose(At, A, len, len);
    mm(U, AS, At, len, len, len);
    for (int j = 0; j < len; ++j)
        for (int i = 0; i < len; ++i)
        {
            const double qb = (i < 3 && j < 3) ? f->Q[i + 3 * j] : 0.0; /* expanded_process_noise :110-125 */
            f->S[i + j * len] = U[i + j * len] + qb;
        }
}

/* computeTheoreticalMeasurement :150-160 (j is 1-based) */
static void zhat_(const okf * f, const double * sv, int j, double * z)
{
    (void) f;
    const double mx = sv[3 + 2 * (j - 1)] - sv[1];
    const double my = sv[4 + 2 * (j - 1)] - sv[2];
    orc_cartesian2polar(mx, my, z);
    z[1] = orc_normalize_angle(z[1] - sv[0]);
}

/* linearizedMeasurementModel :162-186; H is 2 x len column-major */
static void hmat_(const okf * f, const double * sv, int j, double * H)
{
    const int len = f->len;
    memset(H, 0, sizeof(double) * 2 * len);
    const double dx = sv[3 + 2 * (j - 1)] - sv[1];
    const double dy = sv[4 + 2 * (j - 1)] - sv[2];
    const double d = dx * dx + dy * dy;
    const int c = 3 + 2 * (j - 1);
    H[1 + 2 * 0] = -1;
    H[0 + 2 * 1] = -dx / sqrt(d);
    H[1 + 2 * 1] = dy / d;
    H[0 + 2 * 2] = -dy / sqrt(d);
    H[1 + 2 * 2] = -dx / d;
    H[0 + 2 * c] = dx / sqrt(d);
    H[1 + 2 * c] = -dy / d;
    H[0 + 2 * (c + 1)] = dy / sqrt(d);
    H[1 + 2 * (c + 1)] = dx / d;
}

void orc_ekf_zhat(void * h, int j, double * zhat2) { zhat_((okf *) h, ((okf *) h)->x, j, zhat2); }
void orc_ekf_H(void * h, int j, double * H) { hmat_((okf *) h, ((okf *) h)->x, j, H); }

/* psi = H * Sigma * H.t() + R  (:215, :270) */
static void innovation_cov(const okf * f, const double * H, double * HS, double * Ht, double * psi)
{
    const int len = f->len;
    mm(HS, H, f->S, 2, len, len);
    transpose(Ht, H, 2, len);
    mm(psi, HS, Ht, 2, len, 2);
    for (int k = 0; k < 4; ++k) psi[k] = psi[k] + f->R[k];
}

/* associateLandmark :188-253 */
int orc_ekf_associate(void * h, const double * z)
{
    okf * f = (okf *) h;
    const int len = f->len;
    const double min_threshold = 0.01, max_threshold = 60;
    if (f->seen == 0)
    {
        f->seen++;
        return f->seen;
    }
    /* temp = state with slot seen+1 written from z (:204-207); Armadillo bounds check throws when the map is full */
    if (3 + 2 * f->seen >= len) return ORC_EXC;
    double * temp = f->w3;
    memcpy(temp, f->x, sizeof(double) * len);
    temp[3 + 2 * f->seen] = temp[1] + z[0] * ORC_COS(z[1] + temp[0]);
    if (4 + 2 * f->seen >= len) return ORC_EXC;
    temp[4 + 2 * f->seen] = temp[2] + z[0] * ORC_SIN(z[1] + temp[0]);

    double * H = f->w0, * HS = f->w1, * Ht = f->w2;
    for (int k = 1; k < f->seen + 1; k++)
    {
        double psi[4], psi_i[4], zh[2];
        hmat_(f, temp, k, H);
        innovation_cov(f, H, HS, Ht, psi);
        zhat_(f, temp, k, zh);
        const double dz0 = z[0] - zh[0], dz1 = z[1] - zh[1]; /* no angle wrap (:229-231) */
        if (!inv2(psi_i, psi)) return ORC_EXC;
        /* (dz.t() * psi.i()) * dz */
        double t0 = 0.0, t1 = 0.0, d = 0.0;
        t0 = t0 + dz0 * psi_i[0];
        t0 = t0 + dz1 * psi_i[1];
        t1 = t1 + dz0 * psi_i[2];
        t1 = t1 + dz1 * psi_i[3];
        d = d + t0 * dz0;
        d = d + t1 * dz1;
        if (d < min_threshold) return k;
        else if ((d > min_threshold) && (d < max_threshold)) return -1;
    }
    f->seen++;
    return f->seen;
}

/* initializeLandmark :255-261 */
void orc_ekf_init_landmark(void * h, const double * z, int id)
{
    okf * f = (okf *) h;
    f->x[3 + 2 * (id - 1)] = f->x[1] + z[0] * ORC_COS(z[1] + f->x[0]);
    f->x[4 + 2 * (id - 1)] = f->x[2] + z[0] * ORC_SIN(z[1] + f->x[0]);
}

/* update :263-282 */
int orc_ekf_update(void * h, const double * z, int id)
{
    okf * f = (okf *) h;
    const int len = f->len;
    if (id < 1 || 4 + 2 * (id - 1) >= len) return ORC_EXC; /* Armadillo bounds check */
    double zh[2], psi[4], psi_i[4];
    zhat_(f, f->x, id, zh);
    double * H = f->w0, * HS = f->w1, * Ht = f->w2;
    hmat_(f, f->x, id, H);
    innovation_cov(f, H, HS, Ht, psi);
    if (!inv2(psi_i, psi)) return ORC_EXC;
    /* K = (Sigma * H.t()) * inv(...) (:270) */
    double * P = f->w1;       /* len x 2, HS no longer needed */
    double * K = f->w3;       /* len x 2 */
    mm(P, f->S, Ht, len, len, 2);
    mm(K, P, psi_i, len, 2, 2);
    const double dz[2] = {z[0] - zh[0], z[1] - zh[1]}; /* :272, no wrap */
    /* state += K * dz (:275) */
    for (int i = 0; i < len; ++i)
    {
        double acc = 0.0;
        acc = acc + K[i] * dz[0];
        acc = acc + K[i + len] * dz[1];
        f->x[i] = f->x[i] + acc;
    }
    f->x[0] = orc_normalize_angle(f->x[0]); /* :276 */
    /* Sigma = (I - K*H) * Sigma (:279) */
    double * KH = f->w2; /* Ht no longer needed */
    mm(KH, K, H, len, 2, len);
    double * M = f->w0;  /* H no longer needed after KH */
    for (int j = 0; j < len; ++j)
        for (int i = 0; i < len; ++i) M[i + j * len] = ((i == j) ? 1.0 : 0.0) - KH[i + j * len];
    double * NS = f->w1;
    mm(NS, M, f->S, len, len, len);
    memcpy(f->S, NS, sizeof(double) * len * len);
    return 0;
}

/* ------------------------------------------------------------------ batch driver (slam.cpp:262-319) */

typedef struct
{
    int n, T, m, use_initial_state;
    long B, b0, b1;
    const double * robot0, * map0, * Q9, * R4, * twists, * z;
    const int * ids;
    double * x_io, * sigma_io, * x_trace;
    int * seen_io, * status_out, * ids_out;
} run_args;

static void * run_worker(void * p)
{
    run_args * a = (run_args *) p;
    const int n = a->n, len = 3 + 2 * n, m = a->m;
    const long B = a->B;
    for (long b = a->b0; b < a->b1; ++b)
    {
        okf * f = (okf *) orc_ekf_new(n, a->robot0 + 3 * b, a->map0 + 2 * n * b, a->Q9, a->R4);
        if (a->use_initial_state) orc_ekf_set(f, a->x_io + (long) len * b, a->sigma_io + (long) len * len * b, a->seen_io[b]);
        int status = 0;
        for (int t = 0; t < a->T && status == 0; ++t)
        {
            const double * tw = a->twists + 3 * ((long) t * B + b);
            const int seen_snapshot = f->seen;               /* slam.cpp:251 */
            orc_ekf_predict(f, tw[0], tw[1], tw[2]);          /* slam.cpp:269 */
            for (int i = 0; i < m; ++i)                       /* slam.cpp:279 */
            {
                const long mi = ((long) t * B + b) * m + i;
                const double * zi = a->z + 2 * mi;
                int id;
                if (a->ids)
                {
                    id = a->ids[mi];
                    if (id <= 0)
                    {
                        if (a->ids_out) a->ids_out[mi] = 0;
                        continue;
                    }
                    if (id > f->seen) f->seen = id;
                }
                else
                {
                    id = orc_ekf_associate(f, zi);            /* slam.cpp:291 */
                    if (id == ORC_EXC)
                    {
                        status = 1;
                        if (a->ids_out) a->ids_out[mi] = ORC_EXC;
                        break;
                    }
                }
                if (a->ids_out) a->ids_out[mi] = id;
                if (id > seen_snapshot) orc_ekf_init_landmark(f, zi, id); /* slam.cpp:295-297 */
                else if (id < 0) continue;                                /* slam.cpp:298-300 */
                orc_ekf_update(f, zi, id);                                /* slam.cpp:318 */
            }
            if (a->x_trace) memcpy(a->x_trace + ((long) t * B + b) * len, f->x, sizeof(double) * len);
        }
        orc_ekf_get(f, a->x_io + (long) len * b, a->sigma_io + (long) len * len * b, a->seen_io + b);
        if (a->status_out) a->status_out[b] = status;
        orc_ekf_free(f);
    }
    return NULL;
}

int orc_ekf_run(int n, long B, int T, int m, const double * robot0, const double * map0, const double * Q9,
                const double * R4, const double * twists, const double * z, const int * ids, double * x_io,
                double * sigma_io, int * seen_io, int * status_out, int * ids_out, double * x_trace,
                int use_initial_state, int nthreads)
{
    if (nthreads <= 0) nthreads = (int) sysconf(_SC_NPROCESSORS_ONLN);
    if (nthreads < 1) nthreads = 1;
    if (nthreads > B) nthreads = (int) (B > 0 ? B : 1);
    run_args * args = (run_args *) calloc(nthreads, sizeof(run_args));
    pthread_t * th = (pthread_t *) calloc(nthreads, sizeof(pthread_t));
    for (int w = 0; w < nthreads; ++w)
    {
        run_args a = {n, T, m, use_initial_state, B, B * w / nthreads, B * (w + 1) / nthreads,
                      robot0, map0, Q9, R4, twists, z, ids, x_io, sigma_io, x_trace, seen_io, status_out, ids_out};
        args[w] = a;
        if (nthreads == 1) run_worker(&args[w]);
        else pthread_create(&th[w], NULL, run_worker, &args[w]);
    }
    if (nthreads > 1)
        for (int w = 0; w < nthreads; ++w) pthread_join(th[w], NULL);
    free(args);
    free(th);
    return 0;
}

/* ------------------------------------------------------------------ circle path (nuslam/src/circle_fit_library.cpp) */

/* clusterPoints :136-206. Returns clusters after the erase loop. */
int orc_cluster_points(const float * ranges, double minRange, double maxRange, int * offsets, int * beams,
                       double * px, double * py)
{
    /* pre-erase clusters: start/size in a flat list; wrap point appended to cluster 0 */
    int cstart[361], csize[361], nc = 0;
    int fb[361];
    double fx[361], fy[361];
    int nflat = 0;
    int wrap_beam = -1;
    double wrap_x = 0, wrap_y = 0;
    int cur_start = 0, cur_n = 0; /* current_cluster lives at the tail of the flat list */
    int curr_angle = 0;
    const double threshold = 0.04;
    while (curr_angle < 360)
    {
        if ((ranges[curr_angle] > maxRange) || (ranges[curr_angle] < minRange)) /* :149 */
        {
            curr_angle += 1;
            continue;
        }
        const int next_angle = (curr_angle + 1) % 360;
        const double curr_dist = ranges[curr_angle];
        const double next_dist = ranges[next_angle];
        const double x = ranges[curr_angle] * ORC_COS(deg2rad_(curr_angle)); /* :162 */
        const double y = ranges[curr_angle] * ORC_SIN(deg2rad_(curr_angle)); /* :163 */
        if (fabs(curr_dist - next_dist) < threshold)
        {
            if (next_angle < curr_angle)
            {
                if (nc == 0) return ORC_UB; /* clusters[0] on an empty vector (:173) */
                wrap_beam = curr_angle;
                wrap_x = x;
                wrap_y = y;
            }
            else
            {
                fb[nflat] = curr_angle;
                fx[nflat] = x;
                fy[nflat] = y;
                nflat++;
                cur_n++;
                curr_angle += 1;
            }
        }
        else
        {
            fb[nflat] = curr_angle;
            fx[nflat] = x;
            fy[nflat] = y;
            nflat++;
            cur_n++;
            cstart[nc] = cur_start;
            csize[nc] = cur_n;
            nc++;
            cur_start = nflat;
            cur_n = 0;
            curr_angle += 1;
        }
        if (next_angle < curr_angle) break; /* :192 */
    }
    /* erase loop :198-204 with its index-skipping behaviour */
    int keep[361], nk = 0;
    {
        int alive[361], na = nc;
        for (int i = 0; i < nc; ++i) alive[i] = i;
        for (int i = 0; i < na; i++)
        {
            const int c = alive[i];
            const int sz = csize[c] + ((c == 0 && wrap_beam >= 0) ? 1 : 0);
            if (sz < 3)
            {
                for (int k = i; k + 1 < na; ++k) alive[k] = alive[k + 1];
                na--;
            }
        }
        for (int i = 0; i < na; ++i) keep[nk++] = alive[i];
    }
    int off = 0;
    for (int q = 0; q < nk; ++q)
    {
        const int c = keep[q];
        offsets[q] = off;
        for (int k = 0; k < csize[c]; ++k)
        {
            beams[off] = fb[cstart[c] + k];
            px[off] = fx[cstart[c] + k];
            py[off] = fy[cstart[c] + k];
            off++;
        }
        if (c == 0 && wrap_beam >= 0)
        {
            beams[off] = wrap_beam;
            px[off] = wrap_x;
            py[off] = wrap_y;
            off++;
        }
    }
    offsets[nk] = off;
    return nk;
}

/* classifyCluster :208-250 */
int orc_classify_cluster(const double * px, const double * py, int N)
{
    const double p2x = px[0], p2y = py[0];
    const double p3x = px[N - 1], p3y = py[N - 1];
    double angles[362];
    int na = 0;
    for (int i = 1; i < N - 1; i++)
    {
        const double p1x = px[i], p1y = py[i];
        const double num = p2y * (p1x - p3x) + p1y * (p3x - p2x) + p3y * (p2x - p1x);
        const double den = (p2x - p1x) * (p1x - p3x) + (p2y - p1y) * (p1y - p3y);
        angles[na++] = rad2deg_(ORC_ATAN2(num, den));
    }
    double mean = 0;
    for (int k = 0; k < na; ++k) mean += angles[k] / na;
    double std_dev = 0;
    for (int k = 0; k < na; ++k) std_dev += (angles[k] - mean) * (angles[k] - mean);
    std_dev = sqrt(std_dev / na); /* na == 0 -> 0/0 = NaN -> false */
    return (std_dev < 10) ? 1 : 0;
}

/* one-sided Jacobi SVD of Z (N x 4): oracle/shim/armadillo svd(); s descending, V 4x4 col-major */
static void svd_n4(const double * Zin, int m, double * s, double * V)
{
    const int n = 4;
    double * A = (double *) malloc(sizeof(double) * m * n);
    memcpy(A, Zin, sizeof(double) * m * n);
    double W[16] = {0};
    for (int i = 0; i < n; ++i) W[i + i * n] = 1.0;
    for (int sweep = 0; sweep < 60; ++sweep)
    {
        int rotated = 0;
        for (int p = 0; p + 1 < n; ++p)
            for (int q = p + 1; q < n; ++q)
            {
                double alpha = 0.0, beta = 0.0, gamma = 0.0;
                for (int i = 0; i < m; ++i)
                {
                    const double ap = A[i + p * m], aq = A[i + q * m];
                    alpha = alpha + ap * ap;
                    beta = beta + aq * aq;
                    gamma = gamma + ap * aq;
                }
                if (gamma == 0.0) continue;
                if (fabs(gamma) <= 1e-300 || fabs(gamma) <= 2.220446049250313e-16 * sqrt(alpha * beta)) continue;
                rotated = 1;
                const double zeta = (beta - alpha) / (2.0 * gamma);
                const double tt = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                const double c = 1.0 / sqrt(1.0 + tt * tt);
                const double sn = c * tt;
                for (int i = 0; i < m; ++i)
                {
                    const double ap = A[i + p * m], aq = A[i + q * m];
                    A[i + p * m] = c * ap - sn * aq;
                    A[i + q * m] = sn * ap + c * aq;
                }
                for (int i = 0; i < n; ++i)
                {
                    const double wp = W[i + p * n], wq = W[i + q * n];
                    W[i + p * n] = c * wp - sn * wq;
                    W[i + q * n] = sn * wp + c * wq;
                }
            }
        if (!rotated) break;
    }
    double norms[4];
    int order[4] = {0, 1, 2, 3};
    for (int j = 0; j < n; ++j)
    {
        double acc = 0.0;
        for (int i = 0; i < m; ++i) acc = acc + A[i + j * m] * A[i + j * m];
        norms[j] = sqrt(acc);
    }
    for (int a = 1; a < n; ++a) /* stable insertion sort, descending */
    {
        const int o = order[a];
        int b = a - 1;
        while (b >= 0 && norms[order[b]] < norms[o])
        {
            order[b + 1] = order[b];
            b--;
        }
        order[b + 1] = o;
    }
    for (int jj = 0; jj < n; ++jj)
    {
        s[jj] = norms[order[jj]];
        for (int i = 0; i < n; ++i) V[i + jj * n] = W[i + order[jj] * n];
    }
    free(A);
}

/* cyclic Jacobi eigen-decomposition of a 4x4 (symmetrised): oracle/shim/armadillo eig_sym(); ascending */
static void eig_sym4(const double * X, double * val, double * vec)
{
    const int n = 4;
    double A[16], W[16] = {0};
    for (int j = 0; j < n; ++j)
        for (int i = 0; i < n; ++i) A[i + j * n] = 0.5 * (X[i + j * n] + X[j + i * n]);
    for (int i = 0; i < n; ++i) W[i + i * n] = 1.0;
    for (int sweep = 0; sweep < 100; ++sweep)
    {
        double off = 0.0, diag = 0.0;
        for (int j = 0; j < n; ++j)
            for (int i = 0; i < n; ++i)
            {
                if (i == j) diag = diag + A[i + j * n] * A[i + j * n];
                else off = off + A[i + j * n] * A[i + j * n];
            }
        if (off == 0.0 || off <= 1e-40 * diag) break;
        for (int p = 0; p + 1 < n; ++p)
            for (int q = p + 1; q < n; ++q)
            {
                const double apq = A[p + q * n];
                if (apq == 0.0) continue;
                const double theta = (A[q + q * n] - A[p + p * n]) / (2.0 * apq);
                const double tt = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(1.0 + theta * theta));
                const double c = 1.0 / sqrt(1.0 + tt * tt);
                const double sn = c * tt;
                for (int k = 0; k < n; ++k)
                {
                    const double akp = A[k + p * n], akq = A[k + q * n];
                    A[k + p * n] = c * akp - sn * akq;
                    A[k + q * n] = sn * akp + c * akq;
                }
                for (int k = 0; k < n; ++k)
                {
                    const double apk = A[p + k * n], aqk = A[q + k * n];
                    A[p + k * n] = c * apk - sn * aqk;
                    A[q + k * n] = sn * apk + c * aqk;
                }
                for (int k = 0; k < n; ++k)
                {
                    const double wp = W[k + p * n], wq = W[k + q * n];
                    W[k + p * n] = c * wp - sn * wq;
                    W[k + q * n] = sn * wp + c * wq;
                }
            }
    }
    int order[4] = {0, 1, 2, 3};
    for (int a = 1; a < n; ++a) /* stable insertion sort, ascending */
    {
        const int o = order[a];
        int b = a - 1;
        while (b >= 0 && A[order[b] + order[b] * n] > A[o + o * n])
        {
            order[b + 1] = order[b];
            b--;
        }
        order[b + 1] = o;
    }
    for (int jj = 0; jj < n; ++jj)
    {
        val[jj] = A[order[jj] + order[jj] * n];
        for (int i = 0; i < n; ++i) vec[i + jj * n] = W[i + order[jj] * n];
    }
}

/* Gaussian elimination with partial pivoting, 4x4: oracle/shim/armadillo solve() */
static int solve4(const double * Ain, const double * b, double * x)
{
    const int n = 4;
    double A[16];
    memcpy(A, Ain, sizeof(A));
    memcpy(x, b, sizeof(double) * n);
    for (int c = 0; c < n; ++c)
    {
        int piv = c;
        for (int r = c + 1; r < n; ++r)
            if (fabs(A[r + c * n]) > fabs(A[piv + c * n])) piv = r;
        if (A[piv + c * n] == 0.0) return 0;
        if (piv != c)
        {
            for (int k = 0; k < n; ++k)
            {
                const double t = A[piv + k * n];
                A[piv + k * n] = A[c + k * n];
                A[c + k * n] = t;
            }
            const double t = x[piv];
            x[piv] = x[c];
            x[c] = t;
        }
        for (int r = c + 1; r < n; ++r)
        {
            const double f = A[r + c * n] / A[c + c * n];
            if (f == 0.0) continue;
            for (int k = c; k < n; ++k) A[r + k * n] = A[r + k * n] - f * A[c + k * n];
            x[r] = x[r] - f * x[c];
        }
    }
    for (int ii = n; ii-- > 0;)
    {
        double acc = x[ii];
        for (int j = ii + 1; j < n; ++j) acc = acc - A[ii + j * n] * x[j];
        x[ii] = acc / A[ii + ii * n];
    }
    return 1;
}

/* circleFit :15-134 */
int orc_circle_fit(const double * pxi, const double * pyi, int N, double * out3)
{
    out3[0] = out3[1] = out3[2] = 0.0;
    if (N <= 0) return -1;
    double * dx = (double *) malloc(sizeof(double) * N);
    double * dy = (double *) malloc(sizeof(double) * N);
    double x_hat = 0, y_hat = 0;
    for (int i = 0; i < N; ++i) /* :21-25, size_t divisor converts to double */
    {
        x_hat += pxi[i] / (double) N;
        y_hat += pyi[i] / (double) N;
    }
    for (int i = 0; i < N; ++i)
    {
        dx[i] = pxi[i] - x_hat;
        dy[i] = pyi[i] - y_hat;
    }
    double z_bar = 0;
    double * Z = (double *) malloc(sizeof(double) * N * 4);
    for (int j = 0; j < N; ++j)
    {
        const double z = dx[j] * dx[j] + dy[j] * dy[j];
        z_bar += z / (double) N;
        Z[j + 0 * N] = z;
        Z[j + 1 * N] = dx[j];
        Z[j + 2 * N] = dy[j];
        Z[j + 3 * N] = 1.0;
    }
    int id = 0;
    double A[4];
    if (N < 4) /* s.size() < 4 (:72-76) */
    {
        id = -1;
    }
    else
    {
        double s[4], V[16];
        svd_n4(Z, N, s, V);
        if (s[3] < 1e-12) /* :78-80 */
        {
            for (int i = 0; i < 4; ++i) A[i] = V[i + 3 * 4];
        }
        else
        {
            /* Hinv :57-61 */
            double Hinv[16] = {0};
            Hinv[0 + 0 * 4] = 0.0;
            Hinv[1 + 1 * 4] = 1.0;
            Hinv[2 + 2 * 4] = 1.0;
            Hinv[0 + 3 * 4] = 0.5;
            Hinv[3 + 0 * 4] = 0.5;
            Hinv[3 + 3 * 4] = -2 * z_bar;
            /* Y = V * diagmat(s) * V.t() (:82) */
            double D[16] = {0}, VD[16], Vt[16], Y[16], YH[16], Q[16];
            for (int i = 0; i < 4; ++i) D[i + i * 4] = s[i];
            mm(VD, V, D, 4, 4, 4);
            transpose(Vt, V, 4, 4);
            mm(Y, VD, Vt, 4, 4, 4);
            mm(YH, Y, Hinv, 4, 4, 4); /* Q = Y * Hinv * Y (:83) */
            mm(Q, YH, Y, 4, 4, 4);
            double eigval[4], eigvec[16];
            eig_sym4(Q, eigval, eigvec);
            int eig_index = 0;
            double eig_max = INT_MAX; /* :92 */
            for (int i = 0; i < 4; i++)
                if (eigval[i] > 0 && eigval[i] < eig_max)
                {
                    eig_index = i;
                    eig_max = eigval[i];
                }
            double Astar[4];
            for (int i = 0; i < 4; ++i) Astar[i] = eigvec[i + eig_index * 4];
            if (!solve4(Y, Astar, A)) id = ORC_EXC;
        }
    }
    if (id == 0)
    {
        const double a = -A[1] / (2 * A[0]);
        const double b = -A[2] / (2 * A[0]);
        const double R2 = (A[1] * A[1] + A[2] * A[2] - 4 * A[0] * A[3]) / (4 * (A[0] * A[0]));
        const double tube_radius = sqrt(R2);
        out3[0] = a + x_hat;
        out3[1] = b + y_hat;
        out3[2] = (2 * tube_radius) / 2; /* scale.x = 2R (:124); callers read scale.x/2 (landmarks.cpp:95) */
    }
    free(dx);
    free(dy);
    free(Z);
    return id;
}

/* landmarks.cpp:84-109 */
int orc_scan_detect(const float * ranges, double minR, double maxR, int * cluster_of_beam, int * n_clusters,
                    double * circles, int max_circles)
{
    int offsets[362], beams[362];
    double px[362], py[362];
    for (int i = 0; i < 360; ++i) cluster_of_beam[i] = -1;
    *n_clusters = 0;
    const int nc = orc_cluster_points(ranges, minR, maxR, offsets, beams, px, py);
    if (nc < 0) return nc;
    *n_clusters = nc;
    int published = 0;
    for (int c = 0; c < nc; ++c)
    {
        const int o = offsets[c], N = offsets[c + 1] - offsets[c];
        for (int k = 0; k < N; ++k) cluster_of_beam[beams[o + k]] = c;
        if (!orc_classify_cluster(px + o, py + o, N)) continue;
        double fit[3];
        const int id = orc_circle_fit(px + o, py + o, N, fit);
        if (id < 0) continue;
        if (fit[2] > 1) continue;
        if (published < max_circles)
        {
            circles[4 * published + 0] = fit[0];
            circles[4 * published + 1] = fit[1];
            circles[4 * published + 2] = fit[2];
            circles[4 * published + 3] = (double) c;
        }
        ++published;
    }
    return published;
}

typedef struct
{
    long s0, s1;
    const float * ranges;
    double minR, maxR;
    short * cob;
    int * n_clusters, * n_circles;
    double * circles;
    int kmax;
} det_args;

static void * det_worker(void * p)
{
    det_args * a = (det_args *) p;
    int cob[360];
    for (long s = a->s0; s < a->s1; ++s)
    {
        int nc = 0;
        const int k = orc_scan_detect(a->ranges + 360 * s, a->minR, a->maxR, cob, &nc, a->circles + (long) 4 * a->kmax * s, a->kmax);
        a->n_clusters[s] = nc;
        a->n_circles[s] = k;
        if (a->cob)
            for (int i = 0; i < 360; ++i) a->cob[360 * s + i] = (short) cob[i];
    }
    return NULL;
}

int orc_scan_detect_batch(long S, const float * ranges, double minR, double maxR, short * cluster_of_beam,
                          int * n_clusters, int * n_circles, double * circles, int kmax, int nthreads)
{
    if (nthreads <= 0) nthreads = (int) sysconf(_SC_NPROCESSORS_ONLN);
    if (nthreads < 1) nthreads = 1;
    if (nthreads > S) nthreads = (int) (S > 0 ? S : 1);
    det_args * args = (det_args *) calloc(nthreads, sizeof(det_args));
    pthread_t * th = (pthread_t *) calloc(nthreads, sizeof(pthread_t));
    for (int w = 0; w < nthreads; ++w)
    {
        det_args a = {S * w / nthreads, S * (w + 1) / nthreads, ranges, minR, maxR, cluster_of_beam, n_clusters, n_circles, circles, kmax};
        args[w] = a;
        if (nthreads == 1) det_worker(&args[w]);
        else pthread_create(&th[w], NULL, det_worker, &args[w]);
    }
    if (nthreads > 1)
        for (int w = 0; w < nthreads; ++w) pthread_join(th[w], NULL);
    free(args);
    free(th);
    return 0;
}

/* simulator slice: restatement in oracle/world_oracle.h, DiffDrive through this flavour's own implementation */
#include "world_oracle.h"
void orc_world_step_batch(long B, double * world, const double * cmd, const double * noise, double dt, const double * tubes,
                          int n_tubes, double tube_rad, double robot_rad, double max_range, float * ranges)
{
    for (long b = 0; b < B; ++b)
        orc_w_step(world + 9 * b, cmd + 3 * b, noise ? noise + 4 * b : NULL, dt, tubes, n_tubes, tube_rad, robot_rad, max_range,
                   ranges + 360 * b, orc_diffdrive_convert_twist, orc_diffdrive_step);
}

/* nuslam/src/slam.cpp:175-210 with rigid2d.cpp:166-214 (Transform2D(v, rad), inv, operator*=) */
void orc_map_to_odom(const double * odom3, const double * est3, double * out3)
{
    tf2d T_ob = {ORC_COS(odom3[2]), ORC_SIN(odom3[2]), odom3[0], odom3[1]};
    tf2d T_mb = {ORC_COS(est3[0]), ORC_SIN(est3[0]), est3[1], est3[2]};
    tf2d T_mo = tf_mul(T_mb, tf_inv(T_ob));
    out3[0] = T_mo.x;
    out3[1] = T_mo.y;
    out3[2] = orc_normalize_angle(asin(T_mo.s));
}
