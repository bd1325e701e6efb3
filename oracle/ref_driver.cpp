// oracle/ref_driver.cpp -- TEST INFRASTRUCTURE ONLY (see oracle/oracle_api.h).
//
// C entry points over the UNMODIFIED reference classes. This file is linked with the reference's
// own translation units (compiled from /root/reference where they lie, never copied) and with
// oracle/shim/ standing in for Armadillo and the ROS message headers.
//
// Every function states the reference call it forwards to. The batch drivers replay the caller
// protocol of nuslam/src/slam.cpp:262-319 and nuslam/src/landmarks.cpp:84-109.
#include <cmath>
#include <cstring>
#include <iostream>
#include <sstream>
#include <thread>
#include <vector>
#include <algorithm>
#include <armadillo>

// teacher-forced tests must be able to overwrite (x, Sigma, seen); the class has no setters
// (slam_library.hpp:25-33), so the driver -- and only the driver -- sees the members.
#define private public
#include "nuslam/slam_library.hpp"
#undef private
#include "nuslam/circle_fit_library.hpp"
#include "rigid2d/rigid2d.hpp"
#include "rigid2d/diff_drive.hpp"

#include "oracle_api.h"

#ifndef ORC_FLAVOUR
#define ORC_FLAVOUR "reference"
#endif

namespace
{
    // associateLandmark / initializeLandmark print per call (slam_library.cpp:235-256): never time stream I/O
    struct Silence
    {
        Silence() { std::cout.setstate(std::ios::badbit); }
    } silence_instance;
#ifdef ORACLE_SHIM_BLAS
    // timing variant: one BLAS thread per calling thread (the driver's own std::thread pool partitions the filters)
    extern "C" void ORACLE_SHIM_SET_THREADS(int);
    struct BlasThreads
    {
        BlasThreads() { ORACLE_SHIM_SET_THREADS(1); }
    } blas_threads_instance;
#endif

    arma::mat to_mat(const double * p, int r, int c)
    {
        arma::mat out(r, c);
        for (int k = 0; k < r * c; ++k) out.mem[k] = p[k];
        return out;
    }
    arma::colvec to_vec(const double * p, int n)
    {
        arma::colvec out(n);
        for (int k = 0; k < n; ++k) out.mem[k] = p[k];
        return out;
    }
    inline slam_library::ExtendedKalman * ekf(void * h) { return static_cast<slam_library::ExtendedKalman *>(h); }
}

extern "C" {

const char * orc_flavour(void) { return ORC_FLAVOUR; }
int orc_eig_sym_full_sweeps(int on)
{
    const int was = arma::eig_sym_full_sweeps() ? 1 : 0;
    arma::eig_sym_full_sweeps() = on != 0;
    return was;
}

void * orc_ekf_new(int n, const double * robot3, const double * map2n, const double * Q9, const double * R4)
{
    // slam_library.cpp:39-63
    return new slam_library::ExtendedKalman(to_vec(robot3, 3), to_vec(map2n, 2 * n), to_mat(Q9, 3, 3), to_mat(R4, 2, 2));
}

void orc_ekf_free(void * h) { delete ekf(h); }

void orc_ekf_predict(void * h, double dth, double dx, double dy)
{
    rigid2d::Twist2D tw;
    tw.dth = dth;
    tw.dx = dx;
    tw.dy = dy;
    ekf(h)->predict(tw);   // slam_library.cpp:65-69
}

int orc_ekf_associate(void * h, const double * z2)
{
    try
    {
        return ekf(h)->associateLandmark(to_vec(z2, 2));   // slam_library.cpp:188-253
    }
    catch (const std::exception &)
    {
        return ORC_EXC;
    }
}

void orc_ekf_init_landmark(void * h, const double * z2, int id)
{
    ekf(h)->initializeLandmark(to_vec(z2, 2), id);   // slam_library.cpp:255-261
}

int orc_ekf_update(void * h, const double * z2, int id)
{
    rigid2d::Twist2D tw;
    tw.dth = 0.0;
    tw.dx = 0.0;
    tw.dy = 0.0;   // unused by update (slam_library.cpp:263)
    try
    {
        ekf(h)->update(tw, to_vec(z2, 2), id);   // slam_library.cpp:263-282
    }
    catch (const std::exception &)
    {
        return ORC_EXC;
    }
    return 0;
}

void orc_ekf_get(void * h, double * x, double * sigma, int * seen)
{
    const arma::colvec & s = ekf(h)->getStateVector();
    const arma::mat & c = ekf(h)->getCovariance();
    if (x) std::memcpy(x, s.memptr(), sizeof(double) * s.n_elem);
    if (sigma) std::memcpy(sigma, c.memptr(), sizeof(double) * c.n_elem);
    if (seen) *seen = ekf(h)->getSeenLandmarks();
}

void orc_ekf_set(void * h, const double * x, const double * sigma, int seen)
{
    slam_library::ExtendedKalman * e = ekf(h);
    if (x) std::memcpy(e->state_vector.memptr(), x, sizeof(double) * e->state_vector.n_elem);
    if (sigma) std::memcpy(e->covariance.memptr(), sigma, sizeof(double) * e->covariance.n_elem);
    e->seen_landmarks = seen;
}

void orc_ekf_zhat(void * h, int j, double * zhat2)
{
    arma::colvec z = ekf(h)->computeTheoreticalMeasurement(j, ekf(h)->getStateVector());   // :150-160
    zhat2[0] = z(0);
    zhat2[1] = z(1);
}

void orc_ekf_H(void * h, int j, double * H2xlen)
{
    arma::mat H = ekf(h)->linearizedMeasurementModel(j, ekf(h)->getStateVector());   // :162-186
    std::memcpy(H2xlen, H.memptr(), sizeof(double) * H.n_elem);
}

void orc_cartesian2polar(double x, double y, double * out2)
{
    arma::colvec rb = slam_library::cartesian2polar(x, y);   // slam_library.cpp:16-22
    out2[0] = rb(0);
    out2[1] = rb(1);
}

double orc_normalize_angle(double rad) { return rigid2d::normalize_angle(rad); }   // rigid2d.cpp:9-13

void orc_diffdrive_convert_twist(double base, double rad, double dth, double dx, double * uL_uR)
{
    rigid2d::DiffDrive dd(base, rad, 0.0, 0.0, 0.0, 0.0, 0.0);
    rigid2d::Twist2D tw;
    tw.dth = dth;
    tw.dx = dx;
    tw.dy = 0.0;
    rigid2d::wheelVel u = dd.convertTwist(tw);   // diff_drive.cpp:66-78
    uL_uR[0] = u.uL;
    uL_uR[1] = u.uR;
}

void orc_diffdrive_step(double * s, double thLnew, double thRnew, double * twist3)
{
    rigid2d::DiffDrive dd(s[0], s[1], s[2], s[3], s[4], s[5], s[6]);
    rigid2d::Twist2D tw = dd.getTwist(thLnew, thRnew);   // slam.cpp:264 -> diff_drive.cpp:80-110
    dd(thLnew, thRnew);                                  // slam.cpp:265 -> diff_drive.cpp:111-146
    twist3[0] = tw.dth;
    twist3[1] = tw.dx;
    twist3[2] = tw.dy;
    s[2] = dd.getX();
    s[3] = dd.getY();
    s[4] = dd.getTh();
    s[5] = dd.getThL();
    s[6] = dd.getThR();
}

void orc_integrate_twist(double dth, double dx, double dy, double * out4)
{
    rigid2d::Twist2D tw;
    tw.dth = dth;
    tw.dx = dx;
    tw.dy = dy;
    rigid2d::Transform2D T = rigid2d::integrateTwist(tw);   // rigid2d.cpp:294-328
    out4[0] = T.getCosTh();
    out4[1] = T.getSinTh();
    out4[2] = T.getX();
    out4[3] = T.getY();
}

int orc_ekf_run(int n, long B, int T, int m, const double * robot0, const double * map0, const double * Q9,
                const double * R4, const double * twists, const double * z, const int * ids, double * x_io,
                double * sigma_io, int * seen_io, int * status_out, int * ids_out, double * x_trace,
                int use_initial_state, int nthreads)
{
    const int len = 3 + 2 * n;
    if (nthreads <= 0) nthreads = (int) std::max(1u, std::thread::hardware_concurrency());
    nthreads = (int) std::min<long>(nthreads, std::max<long>(B, 1));
    auto work = [&](long b0, long b1)
    {
        for (long b = b0; b < b1; ++b)
        {
            slam_library::ExtendedKalman f(to_vec(robot0 + 3 * b, 3), to_vec(map0 + 2 * n * b, 2 * n),
                                           to_mat(Q9, 3, 3), to_mat(R4, 2, 2));
            if (use_initial_state)
                orc_ekf_set(&f, x_io + (long) len * b, sigma_io + (long) len * len * b, seen_io[b]);
            int status = 0;
            for (int t = 0; t < T && status == 0; ++t)
            {
                const double * tw3 = twists + 3 * ((long) t * B + b);
                rigid2d::Twist2D tw;
                tw.dth = tw3[0];
                tw.dx = tw3[1];
                tw.dy = tw3[2];
                const int seen_snapshot = f.getSeenLandmarks();   // slam.cpp:251
                f.predict(tw);                                    // slam.cpp:269
                for (int i = 0; i < m; ++i)                       // slam.cpp:279
                {
                    const long mi = ((long) t * B + b) * m + i;
                    arma::colvec z_i = to_vec(z + 2 * mi, 2);
                    int id;
                    if (ids)
                    {
                        id = ids[mi];
                        if (id <= 0)
                        {
                            if (ids_out) ids_out[mi] = 0;
                            continue;   // no measurement in this slot
                        }
                        if (id > f.seen_landmarks) f.seen_landmarks = id;   // what associateLandmark would have done
                    }
                    else
                    {
                        try
                        {
                            id = f.associateLandmark(z_i);   // slam.cpp:291
                        }
                        catch (const std::exception &)
                        {
                            status = 1;   // the node dies here; freeze the filter
                            if (ids_out) ids_out[mi] = ORC_EXC;
                            break;
                        }
                    }
                    if (ids_out) ids_out[mi] = id;
                    if (id > seen_snapshot) f.initializeLandmark(z_i, id);   // slam.cpp:295-297
                    else if (id < 0) continue;                               // slam.cpp:298-300
                    f.update(tw, z_i, id);                                   // slam.cpp:318
                }
                if (x_trace)
                    std::memcpy(x_trace + ((long) t * B + b) * len, f.getStateVector().memptr(), sizeof(double) * len);
            }
            orc_ekf_get(&f, x_io + (long) len * b, sigma_io + (long) len * len * b, seen_io + b);
            if (status_out) status_out[b] = status;
        }
    };
    if (nthreads == 1)
    {
        work(0, B);
        return 0;
    }
    std::vector<std::thread> pool;
    for (int w = 0; w < nthreads; ++w)
    {
        const long b0 = B * w / nthreads, b1 = B * (w + 1) / nthreads;
        pool.emplace_back(work, b0, b1);
    }
    for (auto & th : pool) th.join();
    return 0;
}

/* ---------------- circle path ---------------- */

namespace
{
    // true when the reference would index clusters[0] of an empty vector (circle_fit_library.cpp:173)
    bool wrap_append_on_empty(const float * r, double minR, double maxR)
    {
        if ((r[359] > maxR) || (r[359] < minR)) return false;
        const double c = r[359], nx = r[0];
        if (!(fabs(c - nx) < 0.04)) return false;
        for (int i = 0; i < 359; ++i)
        {
            if ((r[i] > maxR) || (r[i] < minR)) continue;
            const double a = r[i], b = r[i + 1];
            if (!(fabs(a - b) < 0.04)) return false;   // a cluster is closed before beam 359
        }
        return true;
    }

    int cluster_impl(const float * ranges360, double minR, double maxR, std::vector<std::vector<geometry_msgs::Point>> & clusters)
    {
        if (wrap_append_on_empty(ranges360, minR, maxR)) return ORC_UB;
        std::vector<float> ranges(ranges360, ranges360 + 360);
        clusters = circle_fit::clusterPoints(ranges, minR, maxR);   // circle_fit_library.cpp:136-206
        return (int) clusters.size();
    }

    // recover the beam index of every stored point: points are r*(cos,sin)(deg2rad(i)) evaluated by the
    // reference (circle_fit_library.cpp:161-163); within a cluster beams ascend, except a wrapped 359 at the end
    void beams_of(const float * r, const std::vector<geometry_msgs::Point> & cl, int start_hint, int * beams)
    {
        int i = start_hint;
        for (size_t k = 0; k < cl.size(); ++k)
        {
            int found = -1;
            for (int tries = 0; tries < 360; ++tries)
            {
                const int cand = (i + tries) % 360;
                const double px = r[cand] * cos(rigid2d::deg2rad(cand));
                const double py = r[cand] * sin(rigid2d::deg2rad(cand));
                // a NaN range counts as in range (circle_fit_library.cpp:149) and stores a NaN point: match it by NaN-ness
                if ((px == cl[k].x && py == cl[k].y) || (px != px && cl[k].x != cl[k].x))
                {
                    found = cand;
                    break;
                }
            }
            beams[k] = found;
            if (found >= 0) i = (found + 1) % 360;
        }
    }
}

int orc_cluster_points(const float * ranges360, double minR, double maxR, int * offsets, int * beams,
                       double * px, double * py)
{
    std::vector<std::vector<geometry_msgs::Point>> clusters;
    const int nc = cluster_impl(ranges360, minR, maxR, clusters);
    if (nc < 0) return nc;
    int off = 0, hint = 0;
    for (int c = 0; c < nc; ++c)
    {
        offsets[c] = off;
        beams_of(ranges360, clusters[c], hint, beams + off);
        for (size_t k = 0; k < clusters[c].size(); ++k)
        {
            px[off + k] = clusters[c][k].x;
            py[off + k] = clusters[c][k].y;
        }
        // next cluster starts after this cluster's largest non-wrapped beam
        int mx = -1;
        for (size_t k = 0; k < clusters[c].size(); ++k)
            if (!(c == 0 && k + 1 == clusters[c].size() && beams[off + k] == 359 && clusters[c].size() > 1 && beams[off + k - 1] != 358))
                mx = std::max(mx, beams[off + k]);
        hint = (mx + 1) % 360;
        off += (int) clusters[c].size();
    }
    offsets[nc] = off;
    return nc;
}

int orc_classify_cluster(const double * px, const double * py, int N)
{
    std::vector<geometry_msgs::Point> cl(N);
    for (int k = 0; k < N; ++k)
    {
        cl[k].x = px[k];
        cl[k].y = py[k];
    }
    return circle_fit::classifyCluster(cl) ? 1 : 0;   // circle_fit_library.cpp:208-250
}

int orc_circle_fit(const double * px, const double * py, int N, double * out3)
{
    std::vector<geometry_msgs::Point> cl(N);
    for (int k = 0; k < N; ++k)
    {
        cl[k].x = px[k];
        cl[k].y = py[k];
    }
    visualization_msgs::Marker mk = circle_fit::circleFit(cl);   // circle_fit_library.cpp:15-134
    out3[0] = mk.pose.position.x;
    out3[1] = mk.pose.position.y;
    out3[2] = mk.scale.x / 2;
    return mk.id;
}

int orc_scan_detect(const float * ranges360, double minR, double maxR, int * cluster_of_beam,
                    int * n_clusters, double * circles, int max_circles)
{
    std::vector<std::vector<geometry_msgs::Point>> clusters;
    for (int i = 0; i < 360; ++i) cluster_of_beam[i] = -1;
    *n_clusters = 0;
    const int nc = cluster_impl(ranges360, minR, maxR, clusters);   // landmarks.cpp:63
    if (nc < 0) return nc;
    *n_clusters = nc;
    std::vector<int> beams(361);
    int hint = 0, published = 0;
    for (int c = 0; c < nc; ++c)   // landmarks.cpp:84
    {
        const std::vector<geometry_msgs::Point> & cl = clusters[c];
        beams_of(ranges360, cl, hint, beams.data());
        int mx = -1;
        for (size_t k = 0; k < cl.size(); ++k)
        {
            if (beams[k] >= 0) cluster_of_beam[beams[k]] = c;
            if (!(c == 0 && k + 1 == cl.size() && beams[k] == 359 && cl.size() > 1 && beams[k - 1] != 358)) mx = std::max(mx, beams[k]);
        }
        hint = (mx + 1) % 360;
        if (!circle_fit::classifyCluster(cl)) continue;            // landmarks.cpp:86
        visualization_msgs::Marker mk = circle_fit::circleFit(cl);   // landmarks.cpp:89
        if (mk.id < 0) continue;                                   // landmarks.cpp:91
        if (mk.scale.x / 2 > 1) continue;                          // landmarks.cpp:95
        if (published < max_circles)
        {
            circles[4 * published + 0] = mk.pose.position.x;
            circles[4 * published + 1] = mk.pose.position.y;
            circles[4 * published + 2] = mk.scale.x / 2;
            circles[4 * published + 3] = (double) c;
        }
        ++published;   // marker.id = id++ (landmarks.cpp:104-105)
    }
    return published;
}

int orc_scan_detect_batch(long S, const float * ranges, double minR, double maxR, short * cluster_of_beam,
                          int * n_clusters, int * n_circles, double * circles, int kmax, int nthreads)
{
    if (nthreads <= 0) nthreads = (int) std::max(1u, std::thread::hardware_concurrency());
    nthreads = (int) std::min<long>(nthreads, std::max<long>(S, 1));
    auto work = [&](long s0, long s1)
    {
        int cob[360];
        for (long s = s0; s < s1; ++s)
        {
            int nc = 0;
            const int k = orc_scan_detect(ranges + 360 * s, minR, maxR, cob, &nc, circles + (long) 4 * kmax * s, kmax);
            n_clusters[s] = nc;
            n_circles[s] = k;
            if (cluster_of_beam)
                for (int i = 0; i < 360; ++i) cluster_of_beam[360 * s + i] = (short) cob[i];
        }
    };
    if (nthreads == 1)
    {
        work(0, S);
        return 0;
    }
    std::vector<std::thread> pool;
    for (int w = 0; w < nthreads; ++w) pool.emplace_back(work, S * w / nthreads, S * (w + 1) / nthreads);
    for (auto & th : pool) th.join();
    return 0;
}

}   // extern "C"

// simulator slice: tube_world.cpp is a roscpp node and cannot be compiled here; its per-step arithmetic is restated in
// oracle/world_oracle.h, with the DiffDrive calls going to the unmodified rigid2d sources through orc_diffdrive_*
#include "world_oracle.h"
extern "C" void orc_world_step_batch(long B, double * world, const double * cmd, const double * noise, double dt, const double * tubes,
                                     int n_tubes, double tube_rad, double robot_rad, double max_range, float * ranges)
{
    for (long b = 0; b < B; ++b)
        orc_w_step(world + 9 * b, cmd + 3 * b, noise ? noise + 4 * b : nullptr, dt, tubes, n_tubes, tube_rad, robot_rad, max_range,
                   ranges + 360 * b, orc_diffdrive_convert_twist, orc_diffdrive_step);
}


// nuslam/src/slam.cpp:175-210 (the node itself needs ROS): the same three rigid2d calls on the unmodified Transform2D
extern "C" void orc_map_to_odom(const double * odom3, const double * est3, double * out3)
{
    using namespace rigid2d;
    Vector2D v;
    v.x = odom3[0];
    v.y = odom3[1];
    Transform2D T_ob(v, odom3[2]);
    v.x = est3[1];
    v.y = est3[2];
    Transform2D T_mb(v, est3[0]);
    Transform2D T_mo = T_mb * T_ob.inv();
    out3[0] = T_mo.getX();
    out3[1] = T_mo.getY();
    out3[2] = normalize_angle(asin(T_mo.getSinTh()));
}
