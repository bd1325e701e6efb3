// oracle/shim_ros: TEST INFRASTRUCTURE ONLY -- see ros_stub_all.h
#include "../ros_stub_all.h"
