// oracle/shim: TEST INFRASTRUCTURE ONLY. Included but unused by nuslam/src/circle_fit_library.cpp:6.
#ifndef ORACLE_SHIM_TF2_ROS_TB_H
#define ORACLE_SHIM_TF2_ROS_TB_H
#endif
