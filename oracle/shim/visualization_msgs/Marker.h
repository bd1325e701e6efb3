// oracle/shim: TEST INFRASTRUCTURE ONLY. Stand-in for visualization_msgs/Marker (ROS 1);
// default-constructed values follow the ROS message defaults (all zero / empty).
#ifndef ORACLE_SHIM_VISUALIZATION_MSGS_MARKER_H
#define ORACLE_SHIM_VISUALIZATION_MSGS_MARKER_H
#include <string>
#include <geometry_msgs/Point.h>
namespace std_msgs
{
    struct ColorRGBA { float r = 0.f; float g = 0.f; float b = 0.f; float a = 0.f; };
    struct Header { unsigned int seq = 0; double stamp = 0.0; std::string frame_id; };
}
namespace visualization_msgs
{
    struct Marker
    {
        enum { ARROW = 0, CUBE = 1, SPHERE = 2, CYLINDER = 3 };
        enum { ADD = 0, MODIFY = 0, DELETE = 2 };
        std_msgs::Header header;
        std::string ns;
        int id = 0;
        int type = 0;
        int action = 0;
        geometry_msgs::Pose pose;
        geometry_msgs::Vector3 scale;
        std_msgs::ColorRGBA color;
        bool frame_locked = false;
    };
}
#endif
