// scan_detect.cuh -- placeholder until the clustering / circle-fit kernels land.
#pragma once
#include "ekf_common.cuh"
namespace nuslam
{
inline cudaError_t launch_scan_detect(const float *, int64_t, double, double, int16_t *, int32_t *, int32_t *, double *, int32_t, cudaStream_t)
{
    return cudaErrorNotSupported;
}
__global__ void k_classify_and_fit(const double *, const double *, const int32_t *, int64_t, int32_t *, double *) {}
}
