// oracle/tube_world_driver.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// Pins the simulator slice: the UNMODIFIED /root/reference/nuturtlesim/src/tube_world.cpp is compiled where it lies (it is textually
// included below; oracle/Makefile `tube_world` passes its path) against the roscpp / message stand-ins of oracle/shim_ros and the
// unmodified rigid2d sources, and its TubeWorld class is driven directly:
//   twref_lidar   TubeWorld::simulate_lidar_scanner (:405-471) at a given robot configuration
//   twref_run     TubeWorld::main_loop (:473-544) for T commanded twists from the origin, with the stub clock advancing one period per
//                 iteration: check_collision (:371-389), convertTwist, joint integration, DiffDrive::operator() with wheel slip,
//                 simulate_lidar_scanner -- everything the node publishes per iteration is recorded
// The node seeds its std::mt19937 from std::random_device (:57-58), so only noise-free runs are reproducible: twist_noise = 0 and
// slip_min = slip_max make both normal distributions degenerate (the slip factor is then the deterministic mean).
// `private` is redefined so that the driver -- and only the driver -- reaches the members; `main` so that the node's entry point
// does not clash.
#include <iostream>
#include <sstream>
#define private public
#define main tube_world_node_main
#include TUBE_WORLD_CPP
#undef main
#undef private

namespace
{
struct Silence
{
    Silence() { std::cout.setstate(std::ios::badbit); }
} silence_instance;

void set_params(double wheel_base, double wheel_rad, const double * tubes, double tube_rad, double robot_rad, double max_scan_range, double slip)
{
    ros::StubWorld & w = ros::StubWorld::get();
    w.num.clear();
    w.vec.clear();
    w.str.clear();
    w.num["max_range"] = max_scan_range;
    w.num["wheel_base"] = wheel_base;
    w.num["wheel_radius"] = wheel_rad;
    w.num["tube_radius"] = tube_rad;
    w.num["tube_var"] = 0.0;
    w.num["twist_noise"] = 0.0;
    w.num["slip_min"] = slip;
    w.num["slip_max"] = slip;
    w.num["robot_radius"] = robot_rad;
    w.num["maximum_range"] = max_scan_range;
    w.num["minimum_range"] = 0.12;
    w.num["angle_increment"] = 0.01745329238;
    w.num["sample_num"] = 360.0;
    w.num["resolution"] = 0.015;
    w.num["noise_level"] = 0.01;
    w.num["wall_width"] = 5.0;
    w.num["wall_height"] = 5.0;
    for (int t = 0; t < 6; ++t)
    {
        std::ostringstream key;
        key << "tube" << (t + 1) << "_location";
        w.vec[key.str()] = {tubes[2 * t], tubes[2 * t + 1]};
    }
    for (const char * k : {"odom_frame_id", "map_frame_id", "scanner_frame_id", "world_frame_id", "turtle_frame_id", "left_wheel_joint", "right_wheel_joint"})
        w.str[k] = k;
    w.cmd_vel.clear();
    w.scans.clear();
    w.joints.clear();
    w.poses.clear();
    ros::Time::clock() = 0;
}
}   // namespace

extern "C" const char * twref_flavour(void) { return "reference tube_world.cpp"; }

// simulate_lidar_scanner at robot configuration (x, y, th); tubes: 6 x 2; ranges_out: 360 floats
extern "C" int twref_lidar(double x, double y, double th, const double * tubes, double tube_rad, double max_scan_range, float * ranges_out)
{
    set_params(0.16, 0.033, tubes, tube_rad, 0.1, max_scan_range, 0.0);
    TubeWorld tw;
    tw.ninja_turtle = rigid2d::DiffDrive(0.16, 0.033, x, y, th, 0.0, 0.0);
    tw.simulate_lidar_scanner();
    const std::vector<float> & r = ros::StubWorld::get().scans.back().ranges;
    if (r.size() != 360) return -1;
    for (int i = 0; i < 360; ++i) ranges_out[i] = r[i];
    return 0;
}

// main_loop for T commanded twists (cmd: T x 3 = dth, dx, dy) from the origin. Outputs per processed twist: pose (x, y, th), joint
// positions (the published encoder readings), the 360 ranges; *dt_out = the loop period the node measured between iterations.
extern "C" int twref_run(int T, const double * cmd, double wheel_base, double wheel_rad, double slip, const double * tubes, double tube_rad,
                         double robot_rad, double max_scan_range, double * poses_out, double * joints_out, float * scans_out, double * dt_out)
{
    set_params(wheel_base, wheel_rad, tubes, tube_rad, robot_rad, max_scan_range, slip);
    ros::StubWorld & w = ros::StubWorld::get();
    for (int t = 0; t < T; ++t)
    {
        geometry_msgs::Twist m;
        m.angular.z = cmd[3 * t];
        m.linear.x = cmd[3 * t + 1];
        m.linear.y = cmd[3 * t + 2];
        w.cmd_vel.push_back(m);
    }
    TubeWorld tw;
    ros::stub_iterations() = T + 1;   // the first iteration only receives the first twist
    tw.main_loop();
    if ((int) w.scans.size() != T || (int) w.poses.size() != T || (int) w.joints.size() != T + 1) return -1;
    for (int t = 0; t < T; ++t)
    {
        poses_out[3 * t] = w.poses[t].pose.position.x;
        poses_out[3 * t + 1] = w.poses[t].pose.position.y;
        poses_out[3 * t + 2] = w.poses[t].pose.orientation.z;
        joints_out[2 * t] = w.joints[t + 1].position[0];
        joints_out[2 * t + 1] = w.joints[t + 1].position[1];
        for (int i = 0; i < 360; ++i) scans_out[360 * t + i] = w.scans[t].ranges[i];
    }
    ros::Duration d;
    d.ns = ros::Rate(50).period_ns;
    *dt_out = d.toSec();
    return 0;
}
