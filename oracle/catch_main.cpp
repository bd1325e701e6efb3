// oracle/catch_main.cpp -- TEST INFRASTRUCTURE ONLY. catch_ros normally supplies main() for
// nuslam/tests/circle_tests.cpp; this translation unit does it for the oracle build.
#define CATCH_CONFIG_MAIN
#include <catch_ros/catch.hpp>
