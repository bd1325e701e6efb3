// oracle/shim: TEST INFRASTRUCTURE ONLY. Field-for-field stand-in for the ROS 1 message
// types the reference circle-fit library names (nuslam/include/nuslam/circle_fit_library.hpp:8-9).
#ifndef ORACLE_SHIM_GEOMETRY_MSGS_POINT_H
#define ORACLE_SHIM_GEOMETRY_MSGS_POINT_H
namespace geometry_msgs
{
    struct Point { double x = 0.0; double y = 0.0; double z = 0.0; };
    struct Quaternion { double x = 0.0; double y = 0.0; double z = 0.0; double w = 0.0; };
    struct Vector3 { double x = 0.0; double y = 0.0; double z = 0.0; };
    struct Pose { Point position; Quaternion orientation; };
}
#endif
