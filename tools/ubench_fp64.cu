// tools/ubench_fp64.cu -- fp64 micro-benchmarks on B200 that size the EKF kernels (DESIGN.md "fp64 pipe").
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o build/ubench_fp64 tools/ubench_fp64.cu
// Prints: dependent-issue latency of DFMA / MUFU.RCP64H / DMMA, per-SM throughput of DFMA, DMMA (m8n8k4,
// m16n8k8, m16n8k16) and of DFMA and DMMA issued together (are they one pipe or two?), LDS.128 broadcast rate.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define CK(x)                                                                                   \
    do                                                                                          \
    {                                                                                           \
        cudaError_t e_ = (x);                                                                   \
        if (e_ != cudaSuccess)                                                                  \
        {                                                                                       \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);    \
            return 1;                                                                           \
        }                                                                                       \
    } while (0)

__device__ __forceinline__ void dmma884(double & d0, double & d1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1688(double (&d)[4], const double (&a)[4], const double (&b)[2])
{
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void dmma1684(double (&d)[4], const double (&a)[2], double b)
{
    asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                 : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
                 : "d"(a[0]), "d"(a[1]), "d"(b));
}

// ---- latency: one warp, dependent chain, clock64 around it ----
__global__ void k_lat(double * out, long long * cyc, double seed)
{
    const int N = 1024;
    double x = seed + threadIdx.x * 1e-9, y = 1.0000001;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) x = fma(x, y, 1e-9);
    long long t1 = clock64();
    double r = seed + 2.0;
#pragma unroll 16
    for (int i = 0; i < N; ++i) asm volatile("rcp.approx.ftz.f64 %0, %0;" : "+d"(r));
    long long t2 = clock64();
    double d0 = x, d1 = r;
#pragma unroll 16
    for (int i = 0; i < N; ++i) dmma884(d0, d1, y, y);
    long long t3 = clock64();
    double s = seed + 3.0;
#pragma unroll 16
    for (int i = 0; i < N; ++i) s = __dadd_rn(s, y);
    long long t4 = clock64();
    double q = seed + 5.0;
#pragma unroll 16
    for (int i = 0; i < N; ++i) asm volatile("rsqrt.approx.ftz.f64 %0, %0;" : "+d"(q));
    long long t5 = clock64();
    double sh = seed;
#pragma unroll 16
    for (int i = 0; i < N; ++i) sh = __shfl_sync(0xffffffffu, sh, (threadIdx.x + 1) & 31);
    long long t6 = clock64();
    if (threadIdx.x == 0)
    {
        cyc[0] = (t1 - t0);
        cyc[1] = (t2 - t1);
        cyc[2] = (t3 - t2);
        cyc[3] = (t4 - t3);
        cyc[4] = (t5 - t4);
        cyc[5] = (t6 - t5);
    }
    out[threadIdx.x] = x + r + d0 + d1 + s + q + sh;
}

// ---- throughput kernels: persistent grid = SMs, W warps per CTA, ILP independent chains ----
template <int MODE>
__global__ void __launch_bounds__(1024) k_tp(double * out, int iters, double seed)
{
    // MODE 0: DFMA x16 chains; 1: DMMA884 x8; 2: DFMA x8 + DMMA884 x4 interleaved; 3: DMMA 16x8x8 x4; 4: DMMA 16x8x4 x4
    double acc[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) acc[k] = seed + k + threadIdx.x * 1e-6;
    const double y = 1.0000001 + seed, c = 1e-9;
    for (int it = 0; it < iters; ++it)
    {
        if (MODE == 0)
        {
#pragma unroll
            for (int k = 0; k < 16; ++k) acc[k] = fma(acc[k], y, c);
        }
        else if (MODE == 1)
        {
#pragma unroll
            for (int k = 0; k < 8; ++k) dmma884(acc[2 * k], acc[2 * k + 1], y, c);
        }
        else if (MODE == 2)
        {
#pragma unroll
            for (int k = 0; k < 4; ++k)
            {
                dmma884(acc[2 * k], acc[2 * k + 1], y, c);
                acc[8 + 2 * k] = fma(acc[8 + 2 * k], y, c);
                acc[9 + 2 * k] = fma(acc[9 + 2 * k], y, c);
            }
        }
        else if (MODE == 3)
        {
            const double a[4] = {y, c, y, c};
            const double b[2] = {c, y};
#pragma unroll
            for (int k = 0; k < 4; ++k) dmma1688(*reinterpret_cast<double(*)[4]>(&acc[4 * k]), a, b);
        }
        else if (MODE == 4)
        {
            const double a[2] = {y, c};
#pragma unroll
            for (int k = 0; k < 4; ++k) dmma1684(*reinterpret_cast<double(*)[4]>(&acc[4 * k]), a, c);
        }
        else if (MODE == 5 || MODE == 6 || MODE == 7)
        {
            // DFMA x16 chains with only part of the warp active: does the fp64 pipe skip inactive half / quarter warps?
            const int active = (MODE == 5) ? 16 : (MODE == 6) ? 8 : 1;
            if ((threadIdx.x & 31) < active)
            {
#pragma unroll
                for (int k = 0; k < 16; ++k) acc[k] = fma(acc[k], y, c);
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < 16; ++k) s += acc[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// LDS.128 broadcast (all lanes same address) and LDS.64 per-lane
template <int MODE>
__global__ void __launch_bounds__(1024) k_lds(double * out, int iters)
{
    __shared__ double2 buf[512];
    for (int i = threadIdx.x; i < 512; i += blockDim.x) buf[i] = make_double2(i, -i);
    __syncthreads();
    double s0 = 0, s1 = 0;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int it = 0; it < iters; ++it)
    {
#pragma unroll
        for (int k = 0; k < 16; ++k)
        {
            const int idx = MODE == 0 ? ((warp * 16 + k + it) & 511) : ((lane + 32 * k + it) & 511);
            const double2 v = buf[idx];
            s0 += v.x;
            s1 += v.y;
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s0 + s1;
}

template <typename F>
float time_ms(F launch)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    launch();
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    launch();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}

int main()
{
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    int clk_khz = 0;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    printf("device %s, %d SMs, clock %d kHz\n", prop.name, prop.multiProcessorCount, clk_khz);
    double * out;
    long long * cyc;
    CK(cudaMalloc(&out, sizeof(double) * 1024 * 1024));
    CK(cudaMalloc(&cyc, sizeof(long long) * 8));
    k_lat<<<1, 32>>>(out, cyc, 1.0);
    CK(cudaDeviceSynchronize());
    k_lat<<<1, 32>>>(out, cyc, 1.0);
    CK(cudaDeviceSynchronize());
    long long h[8];
    CK(cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost));
    printf("latency (cycles per dependent op, 1 warp): DFMA %.2f  MUFU.RCP64H %.2f  DMMA.884 %.2f  DADD %.2f  MUFU.RSQ64H %.2f  SHFL64 %.2f\n",
           h[0] / 1024.0, h[1] / 1024.0, h[2] / 1024.0, h[3] / 1024.0, h[4] / 1024.0, h[5] / 1024.0);

    const int sms = prop.multiProcessorCount;
    const int iters = 20000;
    for (int warps : {4, 8, 16, 32})
    {
        const int threads = warps * 32;
        float ms0 = time_ms([&] { k_tp<0><<<sms, threads>>>(out, iters, 0.0); });
        float ms1 = time_ms([&] { k_tp<1><<<sms, threads>>>(out, iters, 0.0); });
        float ms2 = time_ms([&] { k_tp<2><<<sms, threads>>>(out, iters, 0.0); });
        float ms3 = time_ms([&] { k_tp<3><<<sms, threads>>>(out, iters, 0.0); });
        float ms4 = time_ms([&] { k_tp<4><<<sms, threads>>>(out, iters, 0.0); });
        CK(cudaDeviceSynchronize());
        // per SM: FMAs per iteration
        const double f0 = 16.0 * 32 * warps * iters, f1 = 8.0 * 256 * warps * iters, f2 = (4.0 * 256 + 8.0 * 32) * warps * iters;
        const double f3 = 4.0 * 1024 * warps * iters, f4 = 4.0 * 512 * warps * iters;
        printf("warps/SM %2d: DFMA %.3f ms (%.1f FMA/ns/SM, %.2f TFLOP/s) | DMMA884 %.3f ms (%.1f FMA/ns/SM, %.2f TF) | mixed %.3f ms (%.1f FMA/ns/SM, %.2f TF) | DMMA1688 %.3f ms (%.2f TF) | DMMA1684 %.3f ms (%.2f TF)\n",
               warps, ms0, f0 / (ms0 * 1e6), 2 * f0 * sms / (ms0 * 1e9), ms1, f1 / (ms1 * 1e6), 2 * f1 * sms / (ms1 * 1e9), ms2,
               f2 / (ms2 * 1e6), 2 * f2 * sms / (ms2 * 1e9), ms3, 2 * f3 * sms / (ms3 * 1e9), ms4, 2 * f4 * sms / (ms4 * 1e9));
    }
    for (int warps : {8, 16, 32})
    {
        const int threads = warps * 32;
        float m0 = time_ms([&] { k_tp<0><<<sms, threads>>>(out, iters, 0.0); });
        float m5 = time_ms([&] { k_tp<5><<<sms, threads>>>(out, iters, 0.0); });
        float m6 = time_ms([&] { k_tp<6><<<sms, threads>>>(out, iters, 0.0); });
        float m7 = time_ms([&] { k_tp<7><<<sms, threads>>>(out, iters, 0.0); });
        CK(cudaDeviceSynchronize());
        const double wi = 16.0 * warps * iters;   // warp-level DFMA instructions per SM
        printf("warps/SM %2d: DFMA warp-instr/ns/SM with 32 / 16 / 8 / 1 active lanes: %.2f / %.2f / %.2f / %.2f\n", warps, wi / (m0 * 1e6),
               wi / (m5 * 1e6), wi / (m6 * 1e6), wi / (m7 * 1e6));
    }
    for (int warps : {8, 16, 32})
    {
        const int threads = warps * 32;
        const int it2 = 4000;
        float msa = time_ms([&] { k_lds<0><<<sms, threads>>>(out, it2); });
        float msb = time_ms([&] { k_lds<1><<<sms, threads>>>(out, it2); });
        CK(cudaDeviceSynchronize());
        const double n = 16.0 * warps * it2;   // warp-level LDS.128 instructions per SM
        printf("warps/SM %2d: LDS.128 broadcast %.2f warp-instr/ns/SM | LDS.128 per-lane-distinct %.2f warp-instr/ns/SM\n", warps, n / (msa * 1e6),
               n / (msb * 1e6));
    }
    return 0;
}
