// include/nuslam_b200/compat/nuslam/slam_library.hpp -- forwarding header: with include/nuslam_b200/compat on the include path,
// `#include <nuslam/slam_library.hpp>` of an unmodified caller (nuslam/src/slam.cpp:1-20) resolves to the B200 facade:
// slam_library::ExtendedKalman and slam_library::cartesian2polar with the reference's signatures (slam_library.hpp:18-108) over the
// C ABI of libnuslam_b200.so. Twist2D / normalize_angle come from the caller's own rigid2d (NUSLAM_B200_USE_RIGID2D).
#ifndef NUSLAM_B200_COMPAT_SLAM_LIBRARY_HPP
#define NUSLAM_B200_COMPAT_SLAM_LIBRARY_HPP
#ifndef NUSLAM_B200_USE_RIGID2D
#define NUSLAM_B200_USE_RIGID2D 1
#endif
#ifndef NUSLAM_B200_USE_ARMADILLO
#define NUSLAM_B200_USE_ARMADILLO 1   // colvec / mat are arma's, as in the reference's header
#endif
#include "../../slam_library.hpp"
#endif
