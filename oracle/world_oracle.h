/* oracle/world_oracle.h -- TEST INFRASTRUCTURE ONLY: plain-C restatement of the simulator node's per-step arithmetic,
 * /root/reference/nuturtlesim/src/tube_world.cpp (a roscpp node; the restatement serves BOTH oracle flavours and is PINNED bit
 * for bit to the unmodified node compiled against roscpp stand-ins: oracle/tube_world_driver.cpp + oracle/shim_ros ->
 * oracle/_ref/libtube_world_ref.so, checked by tests/test_oracle.py). The DiffDrive calls go
 * through the flavour's own orc_diffdrive_* (the unmodified rigid2d sources in oracle/_ref).
 *
 * One iteration of TubeWorld::main_loop (:512-537) for one robot:
 *   desired twist = cmd + gaussian twist noise                 twist_callback :177-189  (the two draws are INPUTS here)
 *   check_collision                                            :371-389
 *   wheel_vel = convertTwist; joints += wheel_vel * dt         :516-523
 *   robot(joints + wheel_vel * slip draw)                      :528-529                 (the two draws are INPUTS here)
 *   simulate_lidar_scanner                                     :405-471
 * world9 = {wheelBase, wheelRad, x, y, th, thL, thR, jointL, jointR}; noise4 = {twist dth, twist dx, slip L, slip R}.
 */
#ifndef NUSLAM_WORLD_ORACLE_H
#define NUSLAM_WORLD_ORACLE_H
#include <math.h>

#define ORC_W_PI 3.14159265358979323846 /* rigid2d.hpp:15 */
static double orc_w_deg2rad(double deg) { return (ORC_W_PI / (double) 180) * deg; } /* rigid2d.hpp:40-44 */
static double orc_w_rad2deg(double rad) { return ((double) 180 / ORC_W_PI) * rad; } /* rigid2d.hpp:49-53 */

/* tube_world.cpp:405-471; x, y, th = the simulated robot's configuration */
static void orc_w_lidar(double x, double y, double th, const double * tubes, int n_tubes, double tube_rad, double max_scan_range,
                        float * lidar_ranges)
{
    for (int k = 0; k < 360; ++k) lidar_ranges[k] = (float) (max_scan_range + 1); /* :416 */
    for (int t = 0; t < n_tubes; ++t)
    {
        const double xt = tubes[2 * t], yt = tubes[2 * t + 1];
        const double x1 = x - xt, y1 = y - yt;                                   /* :423-424 */
        const int tube_angle = (int) round(orc_w_rad2deg(atan2(yt - y1, xt - x1))); /* :426 (sic: relative coordinates) */
        for (int i = tube_angle - 27; i < tube_angle + 27; i++)
        {
            const double x2 = x1 + max_scan_range * cos(orc_w_deg2rad(i));
            const double y2 = y1 + max_scan_range * sin(orc_w_deg2rad(i));
            const double dx = x2 - x1, dy = y2 - y1;
            const double dr = sqrt(pow(dx, 2) + pow(dy, 2));
            const double det = x1 * y2 - x2 * y1;
            const double dis = (pow(tube_rad, 2) * pow(dr, 2)) - pow(det, 2);
            double distance;
            if (fabs(dis) < 1e-5)
            {
                const double inter_x = (det * dy) / pow(dr, 2);
                const double inter_y = -(det * dx) / pow(dr, 2);
                distance = sqrt(pow(inter_x - x1, 2) + pow(inter_y - y1, 2));
            }
            else if (dis > 0)
            {
                const double root = sqrt((pow(tube_rad, 2) * pow(dr, 2)) - pow(det, 2));
                const double inter_x1 = ((det * dy) + ((dy / fabs(dy)) * dx * root)) / pow(dr, 2);
                const double inter_y1 = (-(det * dx) + fabs(dy) * root) / pow(dr, 2);
                const double dist1 = sqrt(pow(inter_x1 - x1, 2) + pow(inter_y1 - y1, 2));
                const double inter_x2 = ((det * dy) - ((dy / fabs(dy)) * dx * root)) / pow(dr, 2);
                const double inter_y2 = (-(det * dx) - fabs(dy) * root) / pow(dr, 2);
                const double dist2 = sqrt(pow(inter_x2 - x1, 2) + pow(inter_y2 - y1, 2));
                distance = (dist2 < dist1) ? dist2 : dist1; /* std::min(dist1, dist2), :453 */
            }
            else
            {
                distance = max_scan_range + 1;
            }
            int ind = (i - (int) (orc_w_rad2deg(th))) % 360; /* :459 */
            if (ind < 0) ind += 360;
            if (distance < lidar_ranges[ind]) lidar_ranges[ind] = (float) distance; /* :462-464 */
        }
    }
}

/* diffdrive callbacks: the flavour's own implementation of rigid2d::DiffDrive */
typedef void (*orc_w_convert_fn)(double base, double rad, double dth, double dx, double * uL_uR);
typedef void (*orc_w_dd_step_fn)(double * state7, double thLnew, double thRnew, double * twist3);

static void orc_w_step(double * w, const double * cmd3, const double * noise4, double dt, const double * tubes, int n_tubes,
                       double tube_rad, double robot_rad, double max_scan_range, float * ranges360, orc_w_convert_fn convert,
                       orc_w_dd_step_fn dd_step)
{
    const double n_dth = noise4 ? noise4[0] : 0.0, n_dx = noise4 ? noise4[1] : 0.0;
    const double slipL = noise4 ? noise4[2] : 0.0, slipR = noise4 ? noise4[3] : 0.0;
    const double tw_dth = cmd3[0] + n_dth, tw_dx = cmd3[1] + n_dx; /* :181-183 */
    /* check_collision :371-389 */
    for (int t = 0; t < n_tubes; ++t)
    {
        const double dx = tubes[2 * t] - w[2];
        const double dy = tubes[2 * t + 1] - w[3];
        const double dist = sqrt(pow(dx, 2) + pow(dy, 2));
        if (dist <= (tube_rad + robot_rad))
        {
            const double move_x = dy / dist;
            const double move_y = -dx / dist;
            w[2] += move_x / 50; /* changeConfig, diff_drive.cpp:154-159 */
            w[3] += move_y / 50;
        }
    }
    double u[2];
    convert(w[0], w[1], tw_dth, tw_dx, u); /* :516 */
    w[7] += u[0] * dt;                     /* :522-523 */
    w[8] += u[1] * dt;
    double tw[3];
    dd_step(w, w[7] + u[0] * slipL, w[8] + u[1] * slipR, tw); /* :528-529, DiffDrive::operator() */
    orc_w_lidar(w[2], w[3], w[4], tubes, n_tubes, tube_rad, max_scan_range, ranges360);
}
#endif
