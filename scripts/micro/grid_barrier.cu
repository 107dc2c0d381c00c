// Microbenchmark: latency of a grid-wide barrier on B200 for a co-resident grid.
#include <cooperative_groups.h>
#include <cstdio>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;

__global__ void k_custom(unsigned* counter, int iters, int threads_poll) {
    unsigned target = 0;
    for (int it = 0; it < iters; ++it) {
        __syncthreads();
        if (threadIdx.x == 0) {
            target += gridDim.x;
            __threadfence();
            atomicAdd(counter, 1u);
            unsigned seen;
            do {
                asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(counter) : "memory");
            } while (seen < target);
        }
        __syncthreads();
    }
}
__global__ void k_relaxed(unsigned* counter, int iters) {
    unsigned target = 0;
    for (int it = 0; it < iters; ++it) {
        __syncthreads();
        if (threadIdx.x == 0) {
            target += gridDim.x;
            asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
            unsigned seen;
            do {
                asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(counter) : "memory");
            } while (seen < target);
            asm volatile("fence.acq_rel.gpu;" ::: "memory");
        }
        __syncthreads();
    }
}
__global__ void k_cg(int iters) {
    cg::grid_group g = cg::this_grid();
    for (int it = 0; it < iters; ++it) g.sync();
}
int main() {
    unsigned* c;
    cudaMalloc(&c, 4);
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    const int iters = 2000;
    for (int threads : {256, 1024}) {
        for (int per_sm : {1, 2, 4}) {
            if (threads * per_sm > 2048) continue;
            int grid = 148 * per_sm;
            float ms;
            for (int variant = 0; variant < 3; ++variant) {
                cudaMemset(c, 0, 4);
                int it = iters;
                void* args1[] = {&c, &it, &threads};
                void* args2[] = {&c, &it};
                void* args3[] = {&it};
                cudaEventRecord(a);
                cudaError_t e;
                if (variant == 0) e = cudaLaunchCooperativeKernel((void*)k_custom, dim3(grid), dim3(threads), args1, 0, 0);
                else if (variant == 1) e = cudaLaunchCooperativeKernel((void*)k_relaxed, dim3(grid), dim3(threads), args2, 0, 0);
                else e = cudaLaunchCooperativeKernel((void*)k_cg, dim3(grid), dim3(threads), args3, 0, 0);
                cudaEventRecord(b);
                cudaEventSynchronize(b);
                cudaEventElapsedTime(&ms, a, b);
                printf("threads=%4d ctas/sm=%d variant=%s  %.3f us/barrier (%s)\n", threads, per_sm,
                       variant == 0 ? "custom " : variant == 1 ? "relaxed" : "cg     ", 1e3 * ms / iters, cudaGetErrorString(e));
            }
        }
    }
    return 0;
}
