/* oracle/oracle_api.h -- TEST INFRASTRUCTURE ONLY.
 *
 * One C API, implemented twice:
 *   oracle/ref_driver.cpp   -> oracle/_ref/libnuslam_ref*.so : the UNMODIFIED reference sources
 *                              (/root/reference/nuslam/src/{slam_library,circle_fit_library}.cpp,
 *                              /root/reference/rigid2d/src/{rigid2d,diff_drive}.cpp) compiled where
 *                              they lie against oracle/shim/, driven through their own C++ API.
 *   oracle/nuslam_oracle.c  -> oracle/libnuslam_oracle.so    : the plain-C restatement.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * either library. The product (shermbot-navigation_b200/) never does.
 */
#ifndef NUSLAM_ORACLE_API_H
#define NUSLAM_ORACLE_API_H

#ifdef __cplusplus
extern "C" {
#endif

/* returned where the reference throws (Armadillo bounds check, slam_library.cpp:206 with a full map) */
#define ORC_EXC (-1000)
/* returned by orc_cluster_points where the reference would execute clusters[0].push_back on an empty
   vector (circle_fit_library.cpp:173): undefined behaviour, never executed by the oracle */
#define ORC_UB (-2000)

const char * orc_flavour(void);
/* test hook ("ref" flavour; a no-op in the restatement, which always runs every sweep): 1 = the shim's eig_sym runs to its criterion / 100-sweep
 * cap instead of stopping at the fixed point of its outputs. Returns the previous setting. */
int orc_eig_sym_full_sweeps(int on);

/* ---- single-filter object API (slam_library.hpp:23-113); matrices column-major ---- */
void * orc_ekf_new(int n, const double * robot3, const double * map2n, const double * Q9, const double * R4);
void orc_ekf_free(void * h);
void orc_ekf_predict(void * h, double dth, double dx, double dy);
int orc_ekf_associate(void * h, const double * z2);
void orc_ekf_init_landmark(void * h, const double * z2, int id);
int orc_ekf_update(void * h, const double * z2, int id);
void orc_ekf_get(void * h, double * x, double * sigma, int * seen);
void orc_ekf_set(void * h, const double * x, const double * sigma, int seen);
void orc_ekf_zhat(void * h, int j, double * zhat2);          /* computeTheoreticalMeasurement(j, state) */
void orc_ekf_H(void * h, int j, double * H2xlen);            /* linearizedMeasurementModel(j, state)   */
void orc_cartesian2polar(double x, double y, double * out2);
double orc_normalize_angle(double rad);

/* ---- rigid2d slice used by the caller protocol (diff_drive.cpp:66-146) ---- */
void orc_diffdrive_convert_twist(double base, double rad, double dth, double dx, double * uL_uR);
/* state7 = {wheelBase, wheelRad, x, y, th, thL, thR}; getTwist then operator() as slam.cpp:264-265 */
void orc_diffdrive_step(double * state7, double thLnew, double thRnew, double * twist3);
void orc_integrate_twist(double dth, double dx, double dy, double * cos_sin_x_y);

/* ---- batch driver replaying nuslam/src/slam.cpp:262-319 for B independent filters ----
 * twists[T][B][3] (dth,dx,dy); z[T][B][m][2] polar (range, bearing); ids[T][B][m] (NULL => unknown
 * association through associateLandmark; else id >= 1 known correspondence, id <= 0 => no measurement).
 * Known-correspondence protocol: `if (id > seen_snapshot) initializeLandmark; seen = max(seen,id); update`.
 * status[b]: 0 OK, 1 map full (the reference threw; the filter is frozen at that point).
 * x_trace (optional) [T][B][len]: state after every step. ids_out (optional) [T][B][m].
 * Returns 0.  nthreads <= 0 => hardware concurrency. */
int orc_ekf_run(int n, long B, int T, int m, const double * robot0, const double * map0, const double * Q9,
                const double * R4, const double * twists, const double * z, const int * ids, double * x_io,
                double * sigma_io, int * seen_io, int * status_out, int * ids_out, double * x_trace,
                int use_initial_state, int nthreads);

/* ---- circle path (circle_fit_library.hpp:18-28) ----
 * orc_cluster_points: returns the number of clusters the reference returns (after its erase loop);
 * offsets[nc+1] index into beams/px/py (points in the reference's stored order). */
int orc_cluster_points(const float * ranges360, double minR, double maxR, int * offsets, int * beams,
                       double * px, double * py);
int orc_classify_cluster(const double * px, const double * py, int N);
/* returns marker.id (-1 when the reference rejects N < 4); out3 = {pose.x, pose.y, scale.x/2} */
int orc_circle_fit(const double * px, const double * py, int N, double * out3);
/* landmarks.cpp:84-109: per scan -> cluster index per beam (-1 none), circles (cx, cy, R, cluster).
 * Returns number of published markers, or ORC_UB. circles[k*4+{0,1,2,3}]. */
int orc_scan_detect(const float * ranges360, double minR, double maxR, int * cluster_of_beam,
                    int * n_clusters, double * circles, int max_circles);
/* batched, threaded: S scans. n_circles[S], cluster_of_beam[S][360] (int16), circles[S][kmax][4] */
int orc_scan_detect_batch(long S, const float * ranges, double minR, double maxR, short * cluster_of_beam,
                          int * n_clusters, int * n_circles, double * circles, int kmax, int nthreads);

/* ---- the step AFTER the path: EKFSlam::broadcast_map2odom_tf, nuslam/src/slam.cpp:175-210 ----
 * odom3 = (x, y, th) of the odometry model, est3 = (theta, x, y) = state_estimate(0..2); out3 = (translation x, y, yaw) of
 * T_mo = T_mb * T_ob.inv(), yaw = normalize_angle(asin(T_mo.getSinTh())). */
void orc_map_to_odom(const double * odom3, const double * est3, double * out3);

/* ---- simulator slice (nuturtlesim/src/tube_world.cpp:371-389, 405-471, 512-537; restated in oracle/world_oracle.h) ----
 * B robots, one step each: world[B][9] = {wheelBase, wheelRad, x, y, th, thL, thR, jointL, jointR} in/out, cmd[B][3] (dth, dx, dy),
 * noise[B][4] = {twist dth, twist dx, slip L, slip R} draws or NULL, tubes[n_tubes][2], ranges[B][360] out. */
void orc_world_step_batch(long B, double * world, const double * cmd, const double * noise, double dt, const double * tubes,
                          int n_tubes, double tube_rad, double robot_rad, double max_range, float * ranges);

#ifdef __cplusplus
}
#endif
#endif
