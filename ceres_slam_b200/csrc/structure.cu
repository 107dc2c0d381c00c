// Structure analysis on the device (large problems): the O(n_obs) part of Engine::build_structure
// — observations per point, landmark-major observation lists sorted by (camera, block index), first
// camera / camera-list hash / grouping eligibility per landmark, the co-visibility pattern of the
// reduced camera system (dataset_vo.cpp:40-62 decides which blocks exist) — and, once the host has
// laid the landmarks out, the observation permutation.  The host keeps the O(n_landmarks) logic
// (bucket sorts, groups, work items), which is identical to the host-only analysis; the two paths
// produce the same layout (tests compare the layout hashes).
#include <cub/cub.cuh>
#include <thrust/iterator/counting_iterator.h>

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

// ---- landmark ordering and grouping on the device -------------------------------------------------
// key = first camera of the point (n_poses for points nobody observes), value = point index
__global__ void st_order_keys_kernel(uint32_t n, const uint32_t* __restrict__ cnt, const uint32_t* __restrict__ mincam,
                                     uint32_t n_poses, uint32_t* __restrict__ key, uint32_t* __restrict__ val) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    key[j] = cnt[j] ? mincam[j] : n_poses;
    val[j] = j;
}

// out[q] = first position in the sorted keys that is >= probe[q]
__global__ void st_lower_bound_kernel(uint32_t n, const uint32_t* __restrict__ keys, int n_probe, const uint32_t* __restrict__ probe,
                                      uint32_t* __restrict__ out) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n_probe) return;
    const uint32_t want = probe[q];
    uint32_t lo = 0, hi = n;
    while (lo < hi) {
        const uint32_t mid = lo + (hi - lo) / 2;
        if (keys[mid] < want)
            lo = mid + 1;
        else
            hi = mid;
    }
    out[q] = lo;
}

// per landmark in first-camera order: track length, grouping keys (key2 = first camera << 6 | length for
// eligible landmarks of this rank's shard, all ones otherwise; keyh = hash of the camera list)
__global__ void st_group_keys_kernel(uint32_t n, uint32_t lo, uint32_t hi, const uint32_t* __restrict__ all_lm,
                                     const uint32_t* __restrict__ mincam_s, const uint32_t* __restrict__ cnt,
                                     const uint8_t* __restrict__ kok, const unsigned long long* __restrict__ khash,
                                     uint32_t* __restrict__ len_a, uint32_t* __restrict__ key2,
                                     unsigned long long* __restrict__ keyh, uint32_t* __restrict__ val) {
    const uint32_t a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a > n) return;
    if (a == n) {
        len_a[a] = 0;  // the scan below runs over n + 1 entries
        return;
    }
    const uint32_t j = all_lm[a];
    const uint32_t len = cnt[j];
    len_a[a] = len;
    const bool ok = a >= lo && a < hi && kok[j];
    key2[a] = ok ? (mincam_s[a] << 6 | len) : 0xffffffffu;
    keyh[a] = khash[j];
    val[a] = a;
}

__global__ void st_gather_u32_kernel(uint32_t n, const uint32_t* __restrict__ idx, const uint32_t* __restrict__ src,
                                     uint32_t* __restrict__ dst) {
    const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x < n) dst[x] = src[idx[x]];
}

// a run starts where (key2, hash) changes; only the first n_ok (eligible) entries count
__global__ void st_run_flags_kernel(uint32_t n, const uint32_t* __restrict__ key2_s, const uint32_t* __restrict__ sorted_a,
                                    const unsigned long long* __restrict__ keyh, uint8_t* __restrict__ flag) {
    const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= n) return;
    const uint32_t k = key2_s[x];
    bool f = false;
    if (k != 0xffffffffu) f = x == 0 || key2_s[x - 1] != k || keyh[sorted_a[x - 1]] != keyh[sorted_a[x]];
    flag[x] = f ? 1 : 0;
}

__global__ void st_run_len_kernel(uint32_t n_runs, const uint32_t* __restrict__ run_pos, const uint32_t* __restrict__ key2_s,
                                  uint32_t* __restrict__ run_L) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < n_runs) run_L[r] = key2_s[run_pos[r]] & 63u;
}

// one warp per group: the layout rows of its landmarks, and (lane 0) the group's camera list
__global__ void st_group_layout_kernel(int n_groups, const uint32_t* __restrict__ g_x, const int* __restrict__ g_G,
                                       const int* __restrict__ g_L, const int* __restrict__ g_lm0,
                                       const uint32_t* __restrict__ g_obs0, const int* __restrict__ g_off,
                                       const uint32_t* __restrict__ sorted_a, const uint32_t* __restrict__ all_lm,
                                       const uint32_t* __restrict__ ptr, const unsigned long long* __restrict__ ck,
                                       uint32_t* __restrict__ lm_user, uint32_t* __restrict__ lm_base,
                                       uint32_t* __restrict__ lm_stride, uint32_t* __restrict__ lm_cnt,
                                       uint8_t* __restrict__ grouped, int* __restrict__ g_cams) {
    const int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (g >= n_groups) return;
    const uint32_t x = g_x[g];
    const int G = g_G[g], L = g_L[g];
    for (int jl = lane; jl < G; jl += 32) {
        const uint32_t a = sorted_a[x + jl];
        const size_t li = size_t(g_lm0[g]) + jl;
        lm_user[li] = all_lm[a];
        lm_base[li] = g_obs0[g] + uint32_t(jl);
        lm_stride[li] = uint32_t(G);
        lm_cnt[li] = uint32_t(L);
        grouped[a] = 1;
    }
    if (lane == 0) {
        const unsigned long long* o = ck + ptr[all_lm[sorted_a[x]]];
        for (int i = 0; i < L; ++i) g_cams[g_off[g] + i] = int(o[i] >> 32);
    }
}

// the landmarks of the shard that fell into no group, in first-camera order
__global__ void st_rest_flags_kernel(uint32_t n, uint32_t lo, uint32_t hi, const uint8_t* __restrict__ grouped,
                                     const uint32_t* __restrict__ len_a, uint32_t* __restrict__ rflag, uint32_t* __restrict__ rlen) {
    const uint32_t a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= n) return;
    const bool r = a >= lo && a < hi && !grouped[a];
    rflag[a] = r ? 1u : 0u;
    rlen[a] = r ? len_a[a] : 0u;
}

__global__ void st_rest_layout_kernel(uint32_t n, const uint32_t* __restrict__ rflag, const uint32_t* __restrict__ ridx,
                                      const uint32_t* __restrict__ roff, uint32_t lm0, uint32_t obs0,
                                      const uint32_t* __restrict__ all_lm, const uint32_t* __restrict__ len_a,
                                      uint32_t* __restrict__ lm_user, uint32_t* __restrict__ lm_base,
                                      uint32_t* __restrict__ lm_stride, uint32_t* __restrict__ lm_cnt) {
    const uint32_t a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= n || !rflag[a]) return;
    const size_t li = size_t(lm0) + ridx[a];
    lm_user[li] = all_lm[a];
    lm_base[li] = obs0 + roff[a];
    lm_stride[li] = 1;
    lm_cnt[li] = len_a[a];
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

// ---- StructureSorter: the cub calls of the ordering / grouping stage, with one scratch buffer ----
namespace {
template <class F>
void cub_call(cudaStream_t s, DBuf<uint8_t>& tmp, F&& f) {
    size_t bytes = 0;
    CSLAM_CUDA(f(static_cast<void*>(nullptr), bytes));
    if (tmp.n < bytes || !tmp.p) tmp.alloc(std::max<size_t>(bytes, 1), s);
    CSLAM_CUDA(f(static_cast<void*>(tmp.p), bytes));
    CSLAM_LAUNCHED(1);
}
inline int bits_for(uint32_t v) {
    int b = 1;
    while (b < 32 && (v >> b)) ++b;
    return b;
}
}  // namespace

void launch_st_order(cudaStream_t s, uint32_t n_points, uint32_t n_poses, const uint32_t* cnt, const uint32_t* mincam,
                     uint32_t* key_tmp, uint32_t* val_tmp, uint32_t* mincam_s, uint32_t* all_lm, uint32_t* n_active_out,
                     DBuf<uint8_t>& tmp) {
    const uint32_t n = n_points;
    if (!n) return;
    st_order_keys_kernel<<<(n + 255) / 256, 256, 0, s>>>(n, cnt, mincam, n_poses, key_tmp, val_tmp);
    CSLAM_LAUNCHED(1);
    cub_call(s, tmp, [&](void* t, size_t& b) {
        return cub::DeviceRadixSort::SortPairs(t, b, key_tmp, mincam_s, val_tmp, all_lm, int(n), 0, bits_for(n_poses), s);
    });
    // n_active = first position whose key is n_poses; the probe value is parked in key_tmp[0]
    CSLAM_CUDA(cudaMemcpyAsync(key_tmp, &n_poses, sizeof(uint32_t), cudaMemcpyHostToDevice, s));
    st_lower_bound_kernel<<<1, 32, 0, s>>>(n, mincam_s, 1, key_tmp, n_active_out);
    CSLAM_LAUNCHED(1);
    CSLAM_CUDA(cudaGetLastError());
}

void launch_st_group_sort(cudaStream_t s, uint32_t n_points, uint32_t lo, uint32_t hi, const uint32_t* all_lm,
                          const uint32_t* mincam_s, const uint32_t* cnt, const uint8_t* kok, const unsigned long long* khash,
                          uint32_t* len_a, uint32_t* all_ptr, uint32_t* key2, unsigned long long* keyh,
                          unsigned long long* keyh_tmp, uint32_t* val_a, uint32_t* val_b, uint32_t* key2_b, uint32_t* key2_s,
                          uint32_t* sorted_a, uint8_t* run_flag, uint32_t* run_pos, uint32_t* counts /* [n_ok, n_runs] */,
                          DBuf<uint8_t>& tmp) {
    const uint32_t n = n_points;
    if (!n) return;
    st_group_keys_kernel<<<(n + 1 + 255) / 256, 256, 0, s>>>(n, lo, hi, all_lm, mincam_s, cnt, kok, khash, len_a, key2, keyh, val_a);
    CSLAM_LAUNCHED(1);
    cub_call(s, tmp, [&](void* t, size_t& b) { return cub::DeviceScan::ExclusiveSum(t, b, len_a, all_ptr, size_t(n) + 1, s); });
    // stable LSD: by hash, then by (first camera, length); ties keep first-camera order
    cub_call(s, tmp, [&](void* t, size_t& b) {
        return cub::DeviceRadixSort::SortPairs(t, b, keyh, keyh_tmp, val_a, val_b, int(n), 0, 64, s);
    });
    st_gather_u32_kernel<<<(n + 255) / 256, 256, 0, s>>>(n, val_b, key2, key2_b);
    CSLAM_LAUNCHED(1);
    cub_call(s, tmp, [&](void* t, size_t& b) {
        return cub::DeviceRadixSort::SortPairs(t, b, key2_b, key2_s, val_b, sorted_a, int(n), 0, 32, s);
    });
    st_run_flags_kernel<<<(n + 255) / 256, 256, 0, s>>>(n, key2_s, sorted_a, keyh, run_flag);
    CSLAM_LAUNCHED(1);
    cub_call(s, tmp, [&](void* t, size_t& b) {
        return cub::DeviceSelect::Flagged(t, b, thrust::counting_iterator<uint32_t>(0), run_flag, run_pos, counts + 1, int(n), s);
    });
    const uint32_t all_ones = 0xffffffffu;
    CSLAM_CUDA(cudaMemcpyAsync(key2_b, &all_ones, sizeof(uint32_t), cudaMemcpyHostToDevice, s));
    st_lower_bound_kernel<<<1, 32, 0, s>>>(n, key2_s, 1, key2_b, counts);
    CSLAM_LAUNCHED(1);
    CSLAM_CUDA(cudaGetLastError());
}

// longest track (a single thread sorts a point's list and walks its camera pairs: very long tracks go to
// the host analysis instead)
void launch_st_max_track(cudaStream_t s, uint32_t n_points, const uint32_t* cnt, uint32_t* out, DBuf<uint8_t>& tmp) {
    if (!n_points) return;
    cub_call(s, tmp, [&](void* t, size_t& b) { return cub::DeviceReduce::Max(t, b, cnt, out, int(n_points), s); });
}

void launch_st_run_len(cudaStream_t s, uint32_t n_runs, const uint32_t* run_pos, const uint32_t* key2_s, uint32_t* run_L) {
    if (!n_runs) return;
    st_run_len_kernel<<<(n_runs + 255) / 256, 256, 0, s>>>(n_runs, run_pos, key2_s, run_L);
    CSLAM_LAUNCHED(1);
    CSLAM_CUDA(cudaGetLastError());
}

void launch_st_layout(cudaStream_t s, uint32_t n_points, uint32_t lo, uint32_t hi, int n_groups, const uint32_t* g_x,
                      const int* g_G, const int* g_L, const int* g_lm0, const uint32_t* g_obs0, const int* g_off,
                      const uint32_t* sorted_a, const uint32_t* all_lm, const uint32_t* len_a, const uint32_t* ptr,
                      const unsigned long long* ck, uint32_t n_lm_grouped, uint32_t obs_cursor, uint32_t* lm_user,
                      uint32_t* lm_base, uint32_t* lm_stride, uint32_t* lm_cnt, uint8_t* grouped, int* g_cams, uint32_t* rflag,
                      uint32_t* rlen, uint32_t* ridx, uint32_t* roff, DBuf<uint8_t>& tmp) {
    const uint32_t n = n_points;
    if (!n) return;
    CSLAM_CUDA(cudaMemsetAsync(grouped, 0, n, s));
    if (n_groups > 0) {
        const long long threads = 32ll * n_groups;
        st_group_layout_kernel<<<int((threads + 127) / 128), 128, 0, s>>>(n_groups, g_x, g_G, g_L, g_lm0, g_obs0, g_off, sorted_a,
                                                                         all_lm, ptr, ck, lm_user, lm_base, lm_stride, lm_cnt,
                                                                         grouped, g_cams);
        CSLAM_LAUNCHED(1);
    }
    st_rest_flags_kernel<<<(n + 255) / 256, 256, 0, s>>>(n, lo, hi, grouped, len_a, rflag, rlen);
    CSLAM_LAUNCHED(1);
    cub_call(s, tmp, [&](void* t, size_t& b) { return cub::DeviceScan::ExclusiveSum(t, b, rflag, ridx, int(n), s); });
    cub_call(s, tmp, [&](void* t, size_t& b) { return cub::DeviceScan::ExclusiveSum(t, b, rlen, roff, int(n), s); });
    st_rest_layout_kernel<<<(n + 255) / 256, 256, 0, s>>>(n, rflag, ridx, roff, n_lm_grouped, obs_cursor, all_lm, len_a, lm_user,
                                                          lm_base, lm_stride, lm_cnt);
    CSLAM_LAUNCHED(1);
    CSLAM_CUDA(cudaGetLastError());
}

}  // namespace cslam
