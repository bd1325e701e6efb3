// tools/ubench_gridsync.cu -- latency of a grid-wide barrier with the geometry of k_large_updates_coop (129 blocks x 64 threads):
// cooperative_groups grid.sync() against a hand-rolled one (one atomic counter in L2, acquire polling).
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o build/ubench_gridsync tools/ubench_gridsync.cu && build/ubench_gridsync
#include <cooperative_groups.h>
#include <cstdio>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;

__global__ void k_cg(int n, double * out)
{
    cg::grid_group grid = cg::this_grid();
    double acc = threadIdx.x;
    for (int i = 0; i < n; ++i)
    {
        acc = acc * 1.0000001 + 1.0;
        grid.sync();
    }
    if (out) out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

// counter barrier: every block adds 1, everybody waits until the count reaches (generation + 1) * blocks
__device__ __forceinline__ void grid_barrier(unsigned * counter, unsigned target)
{
    __syncthreads();
    if (threadIdx.x == 0)
    {
        __threadfence();
        atomicAdd(counter, 1u);
        unsigned v;
        do
        {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
        } while (v < target);
    }
    __syncthreads();
}

__global__ void k_own(int n, double * out, unsigned * counter)
{
    double acc = threadIdx.x;
    for (int i = 0; i < n; ++i)
    {
        acc = acc * 1.0000001 + 1.0;
        grid_barrier(counter, (unsigned) (i + 1) * gridDim.x);
    }
    if (out) out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

int main()
{
    const int blocks = 129, threads = 64, n = 200;
    double * out;
    unsigned * counter;
    cudaMalloc(&out, sizeof(double) * blocks * threads);
    cudaMalloc(&counter, 4);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int rep = 0; rep < 3; ++rep)
    {
        int nn = n;
        void * args[] = {(void *) &nn, (void *) &out};
        cudaEventRecord(e0);
        cudaLaunchCooperativeKernel((const void *) k_cg, dim3(blocks), dim3(threads), args, 0, 0);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        printf("cooperative_groups grid.sync: %.2f us per barrier (%d blocks x %d threads)\n", ms * 1e3 / n, blocks, threads);
        cudaMemset(counter, 0, 4);
        void * args2[] = {(void *) &nn, (void *) &out, (void *) &counter};
        cudaEventRecord(e0);
        cudaLaunchCooperativeKernel((const void *) k_own, dim3(blocks), dim3(threads), args2, 0, 0);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        printf("counter barrier (atomicAdd + ld.acquire polling): %.2f us per barrier, %s\n", ms * 1e3 / n, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
