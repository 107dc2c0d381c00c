// Micro-benchmark: FP64 throughput of the vector pipe (DFMA) and of mma.sync m8n8k4 / m16n8k8 f64
// (DMMA) on sm_100a, alone and mixed (some warps DFMA, some DMMA), to decide whether the SYRK part
// of the Schur kernel should run on DMMA.   nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1688(double (&c)[4], const double (&a)[4], const double (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}

// mode 0: all warps DFMA; 1: all warps m8n8k4; 2: all warps m16n8k8; 3: even warps DFMA, odd m8n8k4;
// 4: even warps DFMA, odd idle; 5: odd warps m8n8k4, even idle (3 ~ max(4,5): separate pipes; ~ sum: shared)
template <int MODE>
__global__ void __launch_bounds__(256) k(double* out, int iters, double x) {
    const int warp = threadIdx.x >> 5;
    double acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = threadIdx.x * 1e-3 + i;
    if (MODE == 4 && (warp & 1)) return;
    if (MODE == 5 && !(warp & 1)) return;
    const bool use_mma = MODE == 1 || MODE == 2 || ((MODE == 3 || MODE == 5) && (warp & 1));
    if (!use_mma) {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 16; ++i) acc[i] = fma(acc[i], x, 1e-9);
        }
    } else if (MODE == 2) {
        double a[4] = {x, x + 1, x + 2, x + 3}, b[2] = {x, -x};
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 4; ++i) dmma1688(*reinterpret_cast<double(*)[4]>(&acc[4 * i]), a, b);
        }
    } else {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 8; ++i) dmma884(acc[2 * i], acc[2 * i + 1], x, x + i);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
static void run(const char* name, double fma_per_thread_iter_vec, double fma_per_thread_iter_mma) {
    double* out;
    const int grid = 148 * 4, iters = 20000;
    cudaMalloc(&out, grid * 256 * sizeof(double));
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    k<MODE><<<grid, 256>>>(out, 100, 1.0000001);
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    k<MODE><<<grid, 256>>>(out, iters, 1.0000001);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    double n_vec = MODE == 0 ? 1.0 : (MODE == 3 || MODE == 4) ? 0.5 : 0.0, n_mma = MODE == 1 || MODE == 2 ? 1.0 : (MODE == 3 || MODE == 5) ? 0.5 : 0.0;
    double fmas = double(grid) * 256 * iters * (n_vec * fma_per_thread_iter_vec + n_mma * fma_per_thread_iter_mma);
    printf("%-28s %8.3f ms  %7.2f TFLOP/s  (%s)\n", name, ms, 2 * fmas / ms * 1e-9, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out);
}

int main() {
    // per thread per iteration: DFMA loop 16 FMA; m8n8k4 x8: 8 * (8*8*4/32 = 8) = 64; m16n8k8 x4: 4 * (16*8*8/32 = 32) = 128
    run<0>("DFMA only", 16, 0);
    run<1>("DMMA m8n8k4 only", 0, 64);
    run<2>("DMMA m16n8k8 only", 0, 128);
    run<3>("half DFMA, half m8n8k4", 16, 64);
    run<4>("half DFMA, half idle", 16, 64);
    run<5>("half m8n8k4, half idle", 16, 64);
    return 0;
}
