// Shared device / host utilities for the cslam_b200 CUDA back end (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdio>
#include <stdexcept>
#include <string>
#include <vector>

namespace cslam {

struct CudaError : std::runtime_error {
    explicit CudaError(const std::string& m) : std::runtime_error(m) {}
};

struct NotImplemented : std::logic_error {
    explicit NotImplemented(const std::string& m) : std::logic_error(m) {}
};

#define CSLAM_CUDA(expr)                                                                     \
    do {                                                                                     \
        cudaError_t _e = (expr);                                                             \
        if (_e != cudaSuccess)                                                               \
            throw ::cslam::CudaError(std::string(#expr) + " -> " + cudaGetErrorString(_e) +  \
                                     " (" __FILE__ ":" + std::to_string(__LINE__) + ")");    \
    } while (0)

// Function attributes (opt-in dynamic shared memory, occupancy) are PER DEVICE, and the ABI lets a
// process solve on several devices (cslam_options.device).  A call site keeps one PerDevice flag /
// value set: `first_use()` is true until `mark()` was called for the CURRENT device.  Setting an
// attribute twice from racing handles is harmless (idempotent), launching before it is set is not:
// so set first, mark after.
struct PerDevice {
    std::atomic<unsigned long long> done{0};
    int value[64] = {0};   // optional per-device cached integers (e.g. occupancy), guarded by `done`
    int value2[64] = {0};
    static int current() {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) dev = 0;
        return dev & 63;
    }
    bool first_use(int dev) const { return !(done.load(std::memory_order_acquire) & (1ull << dev)); }
    void mark(int dev) { done.fetch_or(1ull << dev, std::memory_order_release); }
};

// Device buffer with explicit lifetime; no implicit copies.
template <class T>
struct DBuf {
    T* p = nullptr;
    size_t n = 0;
    DBuf() = default;
    DBuf(const DBuf&) = delete;
    DBuf& operator=(const DBuf&) = delete;
    ~DBuf() { release(); }
    void release() {
        if (p) cudaFree(p);  // also valid for stream-ordered allocations (synchronises)
        p = nullptr;
        n = 0;
    }
    // stream-ordered release: no device synchronisation, memory returns to the pool
    void release_async(cudaStream_t s) {
        if (p) cudaFreeAsync(p, s);
        p = nullptr;
        n = 0;
    }
    // With a stream the memory comes from the device's stream-ordered pool (cudaMallocAsync): no
    // implicit synchronisation with in-flight copies / kernels and reuse across solves.
    // Returns true when new memory was handed out (same size: the old block is kept as it is).
    bool alloc_raw(size_t count, cudaStream_t s) {
        if (count == n && p) return false;
        if (s) {
            release_async(s);
            n = count;
            if (count) CSLAM_CUDA(cudaMallocAsync(&p, count * sizeof(T), s));
            return count != 0;
        }
        release();
        n = count;
        if (count) CSLAM_CUDA(cudaMalloc(&p, count * sizeof(T)));
        return count != 0;
    }
    // New memory is handed out zeroed: what a kernel reads before anything wrote it (a first-iteration
    // `0 * old` term, padding rows of a tile) must not depend on what the pool held before.  The fill runs at
    // device bandwidth at upload / structure time, not inside an LM iteration.
    void alloc(size_t count, cudaStream_t s = nullptr) {
        if (!alloc_raw(count, s)) return;
        CSLAM_CUDA(cudaMemsetAsync(p, 0, n * sizeof(T), s));
        // without a stream the fill ran on the legacy default stream, which the engines' non-blocking
        // streams do not wait for: finish it before a kernel of theirs can write the block
        if (!s) CSLAM_CUDA(cudaStreamSynchronize(nullptr));
    }
    void upload(const T* src, size_t count, cudaStream_t s) {
        alloc_raw(count, s);  // fully overwritten by the copy
        if (count) CSLAM_CUDA(cudaMemcpyAsync(p, src, count * sizeof(T), cudaMemcpyHostToDevice, s));
    }
    void upload(const std::vector<T>& v, cudaStream_t s) { upload(v.data(), v.size(), s); }
    void zero(cudaStream_t s) {
        if (n) CSLAM_CUDA(cudaMemsetAsync(p, 0, n * sizeof(T), s));
    }
    size_t bytes() const { return n * sizeof(T); }
};

#ifdef __CUDACC__
// ---- reductions --------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// Block-wide sum into one atomic; every thread must call it. `sh` holds >= 32 doubles.
__device__ __forceinline__ void block_atomic_sum(double v, double* dst, double* sh) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) sh[w] = v;
    __syncthreads();
    if (w == 0) {
        const int nw = (blockDim.x + 31) >> 5;
        double t = lane < nw ? sh[lane] : 0.0;
        t = warp_sum(t);
        if (lane == 0 && t != 0.0) atomicAdd(dst, t);
    }
    __syncthreads();
}
// max of non-negative doubles through their (order-preserving) bit patterns
__device__ __forceinline__ void atomic_max_nonneg(double* dst, double v) {
    atomicMax(reinterpret_cast<unsigned long long*>(dst),
              static_cast<unsigned long long>(__double_as_longlong(v)));
}
// fire-and-forget FP64 add (RED.E.ADD.F64 — no return value requested)
__device__ __forceinline__ void red_add(double* dst, double v) { atomicAdd(dst, v); }

// ---- TMA (bulk async copy) + mbarrier, 1-D flavour ------------------------------------------
// cp.async.bulk moves a contiguous, 16-byte aligned span between global and shared memory
// without tying up registers; completion is signalled on an mbarrier (loads) or through a
// bulk async-group (stores).  SASS: UBLKCP / SYNCS.
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)),
        "r"(phase)
        : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tma_store_1d(void* gmem_dst, const void* smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst),
                 "r"(smem_u32(smem_src)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void tma_store_wait() {
    asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}
// ---- cp.async (Ampere-style 16-byte async copies global -> shared) ----------------------------
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
#endif  // __CUDACC__

}  // namespace cslam
