// Structure analysis on the device (large problems): the O(n_obs) part of Engine::build_structure
// — observations per point, landmark-major observation lists sorted by (camera, block index), first
// camera / camera-list hash / grouping eligibility per landmark, the co-visibility pattern of the
// reduced camera system (dataset_vo.cpp:40-62 decides which blocks exist) — and, once the host has
// laid the landmarks out, the observation permutation.  The host keeps the O(n_landmarks) logic
// (bucket sorts, groups, work items), which is identical to the host-only analysis; the two paths
// produce the same layout (tests compare the layout hashes).
#include <cub/cub.cuh>

#include "kernels.cuh"

#define CSLAM_LAUNCHED(n) g_kernel_launches.fetch_add((n), std::memory_order_relaxed)

namespace cslam {

namespace {

constexpr int kSMs = 148;

inline int st_grid(size_t n, int block) {
    const size_t b = (n + size_t(block) - 1) / size_t(block);
    return int(std::max<size_t>(1, std::min<size_t>(b, size_t(32) * kSMs)));
}

// flags: [0] index out of range, [1] far pairs appended, [2] far list overflow, [3] group mismatch
__global__ void st_count_kernel(size_t n, const uint32_t* __restrict__ cam, const uint32_t* __restrict__ pt, uint32_t n_poses,
                                uint32_t n_points, uint32_t* __restrict__ cnt, uint8_t* __restrict__ used, int* flags) {
    for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) {
        const uint32_t c = cam[i], j = pt[i];
        if (c >= n_poses || j >= n_points) {
            flags[0] = 1;
            continue;
        }
        used[c] = 1;
        atomicAdd(&cnt[j], 1u);
    }
}

__global__ void st_fill_kernel(size_t n, const uint32_t* __restrict__ cam, const uint32_t* __restrict__ pt,
                               uint32_t* __restrict__ fill, unsigned long long* __restrict__ ck) {
    for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) {
        const uint32_t slot = atomicAdd(&fill[pt[i]], 1u);
        ck[slot] = (unsigned long long)cam[i] << 32 | (unsigned long long)i;
    }
}

__device__ void st_sift_down(unsigned long long* a, uint32_t start, uint32_t end) {
    uint32_t root = start;
    while (2 * root + 1 < end) {
        uint32_t child = 2 * root + 1;
        if (child + 1 < end && a[child] < a[child + 1]) ++child;
        if (a[root] >= a[child]) return;
        const unsigned long long t = a[root];
        a[root] = a[child];
        a[child] = t;
        root = child;
    }
}

// One thread per point: sort its (camera << 32 | block) list, then everything the host needs to know
// about the landmark, and its contribution to the reduced system's pattern.
__global__ void st_landmark_kernel(uint32_t n_points, const uint32_t* __restrict__ cnt, const uint32_t* __restrict__ ptr,
                                   unsigned long long* __restrict__ ck, const int* __restrict__ cam_free, int group_lmax,
                                   int allow_groups, uint32_t* __restrict__ mincam, unsigned long long* __restrict__ khash,
                                   uint8_t* __restrict__ kok, unsigned long long* __restrict__ mask, int2* __restrict__ far,
                                   int far_cap, int* flags) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_points) return;
    const uint32_t len = cnt[j];
    if (!len) {
        mincam[j] = 0xffffffffu;
        khash[j] = 0;
        kok[j] = 0;
        return;
    }
    unsigned long long* a = ck + ptr[j];
    if (len <= 64) {
        for (uint32_t x = 1; x < len; ++x) {
            const unsigned long long key = a[x];
            uint32_t y = x;
            while (y > 0 && key < a[y - 1]) {
                a[y] = a[y - 1];
                --y;
            }
            a[y] = key;
        }
    } else {
        for (uint32_t s = len / 2; s-- > 0;) st_sift_down(a, s, len);
        for (uint32_t e = len - 1; e > 0; --e) {
            const unsigned long long t = a[0];
            a[0] = a[e];
            a[e] = t;
            st_sift_down(a, 0, e);
        }
    }
    mincam[j] = uint32_t(a[0] >> 32);
    bool ok = allow_groups && len <= uint32_t(group_lmax);
    unsigned long long h = 1469598103934665603ull;
    uint32_t prev = 0xffffffffu;
    for (uint32_t k = 0; k < len; ++k) {
        const uint32_t c = uint32_t(a[k] >> 32);
        if (c == prev) ok = false;  // the same camera twice: generic kernel only
        prev = c;
        h = (h ^ c) * 1099511628211ull;
    }
    khash[j] = h;
    kok[j] = ok ? 1 : 0;
    // pairs of free cameras that share this landmark: bit (b - a) of row a for offsets below 64
    for (uint32_t x = 0; x + 1 < len; ++x) {
        const int fx = cam_free[uint32_t(a[x] >> 32)];
        if (fx < 0) continue;
        unsigned long long m = 0;
        for (uint32_t y = x + 1; y < len; ++y) {
            const int fy = cam_free[uint32_t(a[y] >> 32)];
            if (fy < 0 || fy == fx) continue;
            const int d = fy - fx;
            if (d < 64) {
                m |= 1ull << d;
            } else {
                const int slot = atomicAdd(&flags[1], 1);
                if (slot < far_cap)
                    far[slot] = make_int2(fx, fy);
                else
                    flags[2] = 1;
            }
        }
        if (m) atomicOr(&mask[fx], m);
    }
}

// camera list of each group = camera list of its first landmark
__global__ void st_group_cams_kernel(int n_groups, const uint32_t* __restrict__ g_first_user, const int* __restrict__ g_off,
                                     const int* __restrict__ g_L, const uint32_t* __restrict__ ptr,
                                     const unsigned long long* __restrict__ ck, int* __restrict__ g_cams) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_groups) return;
    const unsigned long long* a = ck + ptr[g_first_user[g]];
    for (int i = 0; i < g_L[g]; ++i) g_cams[g_off[g] + i] = int(a[i] >> 32);
}

// internal observation order <- caller's block index
__global__ void st_perm_kernel(int n_lm, const uint32_t* __restrict__ lm_user, const uint32_t* __restrict__ lm_base,
                               const uint32_t* __restrict__ lm_stride, const uint32_t* __restrict__ lm_cnt,
                               const uint32_t* __restrict__ ptr, const unsigned long long* __restrict__ ck,
                               uint32_t* __restrict__ obs_user) {
    const int li = blockIdx.x * blockDim.x + threadIdx.x;
    if (li >= n_lm) return;
    const unsigned long long* a = ck + ptr[lm_user[li]];
    const size_t b = lm_base[li], st = lm_stride[li];
    const uint32_t n = lm_cnt[li];
    for (uint32_t k = 0; k < n; ++k) obs_user[b + size_t(k) * st] = uint32_t(a[k]);
}

// The host formed the groups from (first camera, track length, 64-bit hash of the camera list) without
// looking at the lists: check every grouped landmark against its group's cameras (one warp per group).
__global__ void st_verify_kernel(int n_groups, const int* __restrict__ g_L, const int* __restrict__ g_G,
                                 const int* __restrict__ g_lm0, const int* __restrict__ g_off, const int* __restrict__ g_cams,
                                 const uint32_t* __restrict__ lm_user, const uint32_t* __restrict__ ptr,
                                 const unsigned long long* __restrict__ ck, int* flags) {
    const int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (g >= n_groups) return;
    const int L = g_L[g], G = g_G[g];
    const int* cams = g_cams + g_off[g];
    for (int jl = lane; jl < G; jl += 32) {
        const unsigned long long* a = ck + ptr[lm_user[g_lm0[g] + jl]];
        for (int i = 0; i < L; ++i)
            if (int(a[i] >> 32) != cams[i]) flags[3] = 1;
    }
}

}  // namespace

void launch_st_count(cudaStream_t s, size_t n, const uint32_t* cam, const uint32_t* pt, uint32_t n_poses, uint32_t n_points,
                     uint32_t* cnt, uint8_t* used, int* flags) {
    st_count_kernel<<<st_grid(n, 256), 256, 0, s>>>(n, cam, pt, n_poses, n_points, cnt, used, flags);
    CSLAM_LAUNCHED(1);
    CSLAM_CUDA(cudaGetLastError());
}

// ptr[0..n_points] = exclusive prefix sum of cnt[0..n_points] (cnt carries one trailing zero)
void launch_st_scan(cudaStream_t s, uint32_t n_points, const uint32_t* cnt, uint32_t* ptr, DBuf<uint8_t>& tmp) {
    size_t bytes = 0;
    CSLAM_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, bytes, cnt, ptr, size_t(n_points) + 1, s));
    tmp.alloc(std::max<size_t>(bytes, 1), s);
    CSLAM_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, bytes, cnt, ptr, size_t(n_points) + 1, s));
    CSLAM_LAUNCHED(2);
}

void launch_st_fill(cudaStream_t s, size_t n, const uint32_t* cam, const uint32_t* pt, uint32_t* fill, unsigned long long* ck) {
    st_fill_kernel<<<st_grid(n, 256), 256, 0, s>>>(n, cam, pt, fill, ck);
    CSLAM_LAUNCHED(1);
    CSLAM_CUDA(cudaGetLastError());
}

void launch_st_landmarks(cudaStream_t s, uint32_t n_points, const uint32_t* cnt, const uint32_t* ptr, unsigned long long* ck,
                         const int* cam_free, int group_lmax, int allow_groups, uint32_t* mincam, unsigned long long* khash,
                         uint8_t* kok, unsigned long long* mask, int2* far, int far_cap, int* flags) {
    st_landmark_kernel<<<(n_points + 127) / 128, 128, 0, s>>>(n_points, cnt, ptr, ck, cam_free, group_lmax, allow_groups, mincam,
                                                              khash, kok, mask, far, far_cap, flags);
    CSLAM_LAUNCHED(1);
    CSLAM_CUDA(cudaGetLastError());
}

void launch_st_group_cams(cudaStream_t s, int n_groups, const uint32_t* g_first_user, const int* g_off, const int* g_L,
                          const uint32_t* ptr, const unsigned long long* ck, int* g_cams) {
    if (n_groups <= 0) return;
    st_group_cams_kernel<<<(n_groups + 127) / 128, 128, 0, s>>>(n_groups, g_first_user, g_off, g_L, ptr, ck, g_cams);
    CSLAM_LAUNCHED(1);
    CSLAM_CUDA(cudaGetLastError());
}

void launch_st_perm(cudaStream_t s, int n_lm, const uint32_t* lm_user, const uint32_t* lm_base, const uint32_t* lm_stride,
                    const uint32_t* lm_cnt, const uint32_t* ptr, const unsigned long long* ck, uint32_t* obs_user) {
    if (n_lm <= 0) return;
    st_perm_kernel<<<(n_lm + 127) / 128, 128, 0, s>>>(n_lm, lm_user, lm_base, lm_stride, lm_cnt, ptr, ck, obs_user);
    CSLAM_LAUNCHED(1);
    CSLAM_CUDA(cudaGetLastError());
}

void launch_st_verify(cudaStream_t s, int n_groups, const int* g_L, const int* g_G, const int* g_lm0, const int* g_off,
                      const int* g_cams, const uint32_t* lm_user, const uint32_t* ptr, const unsigned long long* ck, int* flags) {
    if (n_groups <= 0) return;
    const long long threads = 32ll * n_groups;
    st_verify_kernel<<<int((threads + 127) / 128), 128, 0, s>>>(n_groups, g_L, g_G, g_lm0, g_off, g_cams, lm_user, ptr, ck, flags);
    CSLAM_LAUNCHED(1);
    CSLAM_CUDA(cudaGetLastError());
}

}  // namespace cslam
