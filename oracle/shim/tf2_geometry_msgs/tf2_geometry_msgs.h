// oracle/shim: TEST INFRASTRUCTURE ONLY. tf2::toMsg lives in the Quaternion stub.
#ifndef ORACLE_SHIM_TF2_GEOMETRY_MSGS_H
#define ORACLE_SHIM_TF2_GEOMETRY_MSGS_H
#include <tf2/LinearMath/Quaternion.h>
#endif
