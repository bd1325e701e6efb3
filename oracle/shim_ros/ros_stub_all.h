// oracle/shim_ros/ros_stub_all.h -- TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// Stand-ins for the roscpp / message headers that /root/reference/nuturtlesim/src/tube_world.cpp includes (:21-42), just deep
// enough for that file to compile UNMODIFIED (oracle/Makefile `tube_world`) and for oracle/tube_world_driver.cpp to drive its
// TubeWorld class deterministically:
//   * the parameter server is a global map the driver fills (NodeHandle::getParam),
//   * ros::Time::now() is a global integer-nanosecond clock that ros::Rate::sleep() advances by exactly one period,
//   * ros::ok() counts down a global number of loop iterations,
//   * ros::spinOnce() delivers the next queued /cmd_vel message to the subscriber's callback,
//   * Publisher::publish hands sensor_msgs::LaserScan / JointState / nav_msgs::Path messages to the driver's recorder.
// Message structs carry the fields tube_world.cpp touches, with ROS's zero / empty defaults.
#ifndef ORACLE_SHIM_ROS_STUB_ALL_H
#define ORACLE_SHIM_ROS_STUB_ALL_H
#include <cmath>
#include <cstdint>
#include <deque>
#include <functional>
#include <map>
#include <string>
#include <vector>

namespace ros
{
struct Duration
{
    int64_t ns = 0;
    double toSec() const { return (double) (ns / 1000000000LL) + 1e-9 * (double) (ns % 1000000000LL); }   // ros::Duration::toSec: sec + 1e-9 nsec
};
struct Time
{
    int64_t ns = 0;
    static int64_t & clock()
    {
        static int64_t c = 0;
        return c;
    }
    static Time now()
    {
        Time t;
        t.ns = clock();
        return t;
    }
    Duration operator-(const Time & o) const
    {
        Duration d;
        d.ns = ns - o.ns;
        return d;
    }
};
struct Rate
{
    int64_t period_ns;
    explicit Rate(double hz) : period_ns((int64_t) std::llround(1e9 / hz)) {}
    bool sleep()
    {
        Time::clock() += period_ns;
        return true;
    }
};
inline int & stub_iterations()
{
    static int n = 0;
    return n;
}
inline bool ok() { return stub_iterations()-- > 0; }
inline void init(int &, char **, const std::string &) {}
}   // namespace ros

namespace std_msgs
{
struct Header
{
    unsigned int seq = 0;
    ros::Time stamp;
    std::string frame_id;
};
struct ColorRGBA
{
    float r = 0.f, g = 0.f, b = 0.f, a = 0.f;
};
}   // namespace std_msgs

namespace geometry_msgs
{
struct Point { double x = 0.0, y = 0.0, z = 0.0; };
struct Vector3 { double x = 0.0, y = 0.0, z = 0.0; };
struct Quaternion { double x = 0.0, y = 0.0, z = 0.0, w = 0.0; };
struct Pose { Point position; Quaternion orientation; };
struct PoseStamped { std_msgs::Header header; Pose pose; };
struct Twist { Vector3 linear, angular; };
struct Transform { Vector3 translation; Quaternion rotation; };
struct TransformStamped { std_msgs::Header header; std::string child_frame_id; Transform transform; };
}   // namespace geometry_msgs

namespace sensor_msgs
{
struct JointState { std_msgs::Header header; std::vector<std::string> name; std::vector<double> position, velocity, effort; };
struct LaserScan
{
    std_msgs::Header header;
    float angle_min = 0.f, angle_max = 0.f, angle_increment = 0.f, time_increment = 0.f, scan_time = 0.f, range_min = 0.f, range_max = 0.f;
    std::vector<float> ranges, intensities;
};
}   // namespace sensor_msgs

namespace visualization_msgs
{
struct Marker
{
    enum { ARROW = 0, CUBE = 1, SPHERE = 2, CYLINDER = 3, LINE_STRIP = 4, LINE_LIST = 5 };
    enum { ADD = 0, MODIFY = 0, DELETE = 2 };
    std_msgs::Header header;
    std::string ns;
    int id = 0, type = 0, action = 0;
    geometry_msgs::Pose pose;
    geometry_msgs::Vector3 scale;
    std_msgs::ColorRGBA color;
    ros::Duration lifetime;
    bool frame_locked = false;
    std::vector<geometry_msgs::Point> points;
};
struct MarkerArray { std::vector<Marker> markers; };
}   // namespace visualization_msgs

namespace nav_msgs
{
struct Path { std_msgs::Header header; std::vector<geometry_msgs::PoseStamped> poses; };
struct Odometry { std_msgs::Header header; std::string child_frame_id; };
}   // namespace nav_msgs

namespace tf2
{
class Quaternion
{
public:
    double x_ = 0.0, y_ = 0.0, z_ = 0.0, w_ = 1.0;
    void setRPY(double roll, double pitch, double yaw)
    {
        const double cr = std::cos(roll * 0.5), sr = std::sin(roll * 0.5);
        const double cp = std::cos(pitch * 0.5), sp = std::sin(pitch * 0.5);
        const double cy = std::cos(yaw * 0.5), sy = std::sin(yaw * 0.5);
        x_ = sr * cp * cy - cr * sp * sy;
        y_ = cr * sp * cy + sr * cp * sy;
        z_ = cr * cp * sy - sr * sp * cy;
        w_ = cr * cp * cy + sr * sp * sy;
    }
};
inline geometry_msgs::Quaternion toMsg(const Quaternion & q)
{
    geometry_msgs::Quaternion m;
    m.x = q.x_;
    m.y = q.y_;
    m.z = q.z_;
    m.w = q.w_;
    return m;
}
}   // namespace tf2

namespace tf2_ros
{
struct TransformBroadcaster
{
    void sendTransform(const geometry_msgs::TransformStamped &) {}
};
}   // namespace tf2_ros

namespace ros
{
// ---- what the driver sees of the node ----
struct StubWorld
{
    std::map<std::string, double> num;
    std::map<std::string, std::string> str;
    std::map<std::string, std::vector<double>> vec;
    std::deque<geometry_msgs::Twist> cmd_vel;                        // queued /cmd_vel messages, one delivered per spinOnce
    std::function<void(const geometry_msgs::Twist &)> cmd_vel_cb;
    std::vector<sensor_msgs::LaserScan> scans;                       // everything published on /scan
    std::vector<sensor_msgs::JointState> joints;                     // ... on /joint_states
    std::vector<geometry_msgs::PoseStamped> poses;                   // last pose of every /real_path message
    static StubWorld & get()
    {
        static StubWorld w;
        return w;
    }
};
inline void record(const sensor_msgs::LaserScan & m) { StubWorld::get().scans.push_back(m); }
inline void record(const sensor_msgs::JointState & m) { StubWorld::get().joints.push_back(m); }
inline void record(const nav_msgs::Path & m)
{
    if (!m.poses.empty()) StubWorld::get().poses.push_back(m.poses.back());
}
template <typename T>
inline void record(const T &) {}
struct Publisher
{
    template <typename T>
    void publish(const T & m) const { record(m); }
};
struct Subscriber
{
};
struct NodeHandle
{
    bool getParam(const std::string & k, double & v) const
    {
        auto it = StubWorld::get().num.find(k);
        if (it == StubWorld::get().num.end()) return false;
        v = it->second;
        return true;
    }
    bool getParam(const std::string & k, int & v) const
    {
        double d = 0.0;
        if (!getParam(k, d)) return false;
        v = (int) d;
        return true;
    }
    bool getParam(const std::string & k, std::string & v) const
    {
        auto it = StubWorld::get().str.find(k);
        if (it == StubWorld::get().str.end()) return false;
        v = it->second;
        return true;
    }
    bool getParam(const std::string & k, std::vector<double> & v) const
    {
        auto it = StubWorld::get().vec.find(k);
        if (it == StubWorld::get().vec.end()) return false;
        v = it->second;
        return true;
    }
    template <typename T>
    Publisher advertise(const std::string &, int, bool = false)
    {
        return Publisher();
    }
    template <typename C>
    Subscriber subscribe(const std::string &, int, void (C::*fn)(const geometry_msgs::Twist &), C * obj)
    {
        StubWorld::get().cmd_vel_cb = [obj, fn](const geometry_msgs::Twist & m) { (obj->*fn)(m); };
        return Subscriber();
    }
};
inline void spinOnce()
{
    StubWorld & w = StubWorld::get();
    if (!w.cmd_vel.empty() && w.cmd_vel_cb)
    {
        const geometry_msgs::Twist m = w.cmd_vel.front();
        w.cmd_vel.pop_front();
        w.cmd_vel_cb(m);
    }
}
}   // namespace ros
#endif
