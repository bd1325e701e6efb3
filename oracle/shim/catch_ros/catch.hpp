// oracle/shim: TEST INFRASTRUCTURE ONLY. catch_ros normally forwards to Catch2; the reference
// vendors Catch v2.13.4 at rigid2d/include/rigid2d/catch.hpp, which the oracle build puts on
// the include path (it is never copied into this repo).
#include <rigid2d/catch.hpp>
