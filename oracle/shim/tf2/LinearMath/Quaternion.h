// oracle/shim: TEST INFRASTRUCTURE ONLY. tf2::Quaternion::setRPY is the only member the
// reference calls (nuslam/src/circle_fit_library.cpp:112-114), always with (0,0,0).
#ifndef ORACLE_SHIM_TF2_QUATERNION_H
#define ORACLE_SHIM_TF2_QUATERNION_H
#include <cmath>
#include <geometry_msgs/Point.h>
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
        m.x = q.x_; m.y = q.y_; m.z = q.z_; m.w = q.w_;
        return m;
    }
}
#endif
