// ekf_fast.cuh -- FAST arithmetic (placeholder until the register-tile kernel lands).
#pragma once
#include "ekf_strict.cuh"
namespace nuslam
{
inline bool fast_supported(int) { return false; }
inline int launch_fast(int, const EkfParams &, bool, int, cudaStream_t) { return -1; }
}
