// ekf_fast_api.cuh -- the seam between the library's two translation units.
//
// nuslam_b200.cu (the C ABI, the oracle-order, scan, simulator and large-map kernels) is compiled whole-program; ekf_fast_tu.cu (the FAST
// filter kernels and the list kernel they hand filters to) is compiled as relocatable device code and linked against the device runtime,
// because those kernels launch the list kernel themselves (ekf_strict.cuh strict_tail). Relocatable code costs the other kernels time
// (ABI calls into the math library's slow paths, constant tables behind relocations: closed loop +15 %, scan kernel +7 %, measured),
// so only the kernels that need it are built that way.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include "ekf_strict.cuh"

namespace nuslam
{

constexpr int kFastMMax = 16;   // measurements per step handled by the FAST kernels
inline bool fast_supported(int n) { return n >= 1 && n <= 12; }

#ifndef NUSLAM_DEFAULT_KERNEL
#define NUSLAM_DEFAULT_KERNEL 4   // 0 static, 1 pair, 2 fast, 3 resident (ekf_res.cuh), 4 resident pair (ekf_res2.cuh)
#endif
// which register kernel serves known correspondence at the BASELINE map size (NUSLAM_KERNEL, read at every call; A/B timing and
// the equality tests): "fast" (ekf_fast.cuh, which also serves everything else), "static" (this file), "pair" (ekf_pair.cuh).
// Measured on B200 (profiles/r02_kernel_iterations.md): the fully unrolled static schedules outgrow the instruction cache
// (92 KB / 60 KB of code against 32 KB) and lose more to instruction fetch than they save in address arithmetic.
inline int known_ids_kernel()
{
    const char * e = getenv("NUSLAM_KERNEL");
    if (e && e[0] == 'p') return 1;
    if (e && e[0] == 's') return 0;
    if (e && e[0] == 'f') return 2;
    // "res": ekf_res.cuh, "res2": ekf_res2.cuh, "res2a": ekf_res2.cuh + the resident pair kernel with on-device association (ekf_res2a.cuh)
    if (e && e[0] == 'r') return (e[1] && e[2] && e[3] == '2') ? (e[4] == 'a' ? 5 : 4) : 3;
    return NUSLAM_DEFAULT_KERNEL;
}

struct FastLaunch
{
    int n_landmarks, sm_count;
    int64_t batch;
    int32_t * worklist;
    int32_t * wl_count;   // [0] entries, [1] finished blocks of the list kernel, [2] finished warps of the FAST kernel, [3] device-launch error
    cudaStream_t stream;
    int strict_warps;     // launch shape of the list kernel
    size_t strict_smem;   // per warp, bytes
};

// One FAST-mode call: the register / resident kernel over the batch, then k_ekf_strict_list over the filters it handed over (launched by
// the kernel itself when the library is built with tail launches, else from here). op: kOpStep or kOpUpdate.
// Returns 0, or -1 when the configuration is not covered (the caller runs the oracle-order kernel over the batch), or 1 with *err / *where
// set when a CUDA call failed.
int fast_path_launch(const FastLaunch & fl, const EkfParams & p, bool do_predict, int op, cudaError_t * err, const char ** where);

// 1 when the FAST kernels launch the list kernel themselves (built with -rdc=true, not switched off by NUSLAM_NO_TAIL_LAUNCH)
int tail_launch_active();

// kernel-experiment builds (-DNUSLAM_TIMING): per-phase clock64 sums of block 0 / warp 0 of the FAST kernel
int fast_timing_read(long long * out16, int reset);

}   // namespace nuslam
