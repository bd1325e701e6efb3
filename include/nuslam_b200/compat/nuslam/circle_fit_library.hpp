// include/nuslam_b200/compat/nuslam/circle_fit_library.hpp -- forwarding header: put include/nuslam_b200/compat on the include path
// (before the reference's nuslam/include) and `#include <nuslam/circle_fit_library.hpp>` of an unmodified caller -- the landmarks node
// (nuslam/src/landmarks.cpp:1-10) or the reference's own test (nuslam/tests/circle_tests.cpp:2) -- resolves to the B200 facade:
// circle_fit::clusterPoints / circleFit / classifyCluster with the reference's signatures (circle_fit_library.hpp:18-28) over the
// C ABI of libnuslam_b200.so. The ROS message types are the caller's own (<geometry_msgs/Point.h>, <visualization_msgs/Marker.h>).
#ifndef NUSLAM_B200_COMPAT_CIRCLE_FIT_LIBRARY_HPP
#define NUSLAM_B200_COMPAT_CIRCLE_FIT_LIBRARY_HPP
#ifndef NUSLAM_B200_USE_ROS
#define NUSLAM_B200_USE_ROS 1
#endif
#ifndef NUSLAM_B200_USE_RIGID2D
#define NUSLAM_B200_USE_RIGID2D 1   // the workspace's own rigid2d (the landmarks node includes it too)
#endif
#include "../../slam_library.hpp"
#endif
