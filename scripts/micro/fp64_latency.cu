// Microbenchmark (one warp): dependent-issue latency of the FP64 instructions that sit on the pivot
// chain of the banded / cyclic-reduction Cholesky kernels.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int MODE>
__global__ void k(double* out, long long* cyc, double seed, int n) {
    __shared__ double sm[64];
    __shared__ __align__(8) uint64_t bars[1024];
    const int lane = threadIdx.x & 31;
    if (threadIdx.x < 64) sm[threadIdx.x] = seed;
    if (MODE == 6)
        for (int i = threadIdx.x; i < 1024; i += blockDim.x)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bars[i])), "r"(32));
    __syncthreads();
    double x = seed + lane * 1e-3, y = 1.0000001;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
        if (MODE == 0) x = fma(x, y, 1e-9);                        // DFMA chain
        if (MODE == 1) x = x * y;                                  // DMUL chain
        if (MODE == 2) x = rsqrt(x) + 1.5;                         // rsqrt() + DADD
        if (MODE == 3) x = __shfl_sync(0xffffffffu, x, (lane + 1) & 31);  // 64-bit shuffle (2 SHFL)
        if (MODE == 4) { sm[lane] = x; __syncwarp(); x = sm[(lane + 1) & 31]; __syncwarp(); }  // STS -> LDS round trip
        if (MODE == 5) x = 1.0 / x + 0.5;                          // division
        if (MODE == 6) {                                           // 32-lane mbarrier arrive + dependent DFMA
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bars[i & 1023])) : "memory");
            x = fma(x, y, 1e-9);
        }
        if (MODE == 7) { float f = rsqrtf((float)x); x = (double)f + 1.5; }  // via FP32 MUFU
        if (MODE == 8) x = sqrt(x) + 0.5;
    }
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}

int main() {
    double* out;
    long long* cyc;
    cudaMalloc(&out, 1024 * 8);
    cudaMallocManaged(&cyc, 8);
    const int n = 4096;
    const char* names[] = {"DFMA chain", "DMUL chain", "rsqrt()+DADD", "SHFL 64-bit", "STS+syncwarp+LDS+syncwarp", "1.0/x + DADD",
                           "mbarrier.arrive x32 + DFMA", "cvt+rsqrtf+cvt+DADD", "sqrt()+DADD"};
    for (int mode = 0; mode < 9; ++mode) {
        for (int rep = 0; rep < 2; ++rep) {
            switch (mode) {
                case 0: k<0><<<1, 32>>>(out, cyc, 1.3, n); break;
                case 1: k<1><<<1, 32>>>(out, cyc, 1.3, n); break;
                case 2: k<2><<<1, 32>>>(out, cyc, 1.3, n); break;
                case 3: k<3><<<1, 32>>>(out, cyc, 1.3, n); break;
                case 4: k<4><<<1, 32>>>(out, cyc, 1.3, n); break;
                case 5: k<5><<<1, 32>>>(out, cyc, 1.3, n); break;
                case 6: k<6><<<1, 32>>>(out, cyc, 1.3, n); break;
                case 7: k<7><<<1, 32>>>(out, cyc, 1.3, n); break;
                case 8: k<8><<<1, 32>>>(out, cyc, 1.3, n); break;
            }
            cudaDeviceSynchronize();
        }
        printf("%-32s %.1f cycles / iteration (%s)\n", names[mode], double(*cyc) / n, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
