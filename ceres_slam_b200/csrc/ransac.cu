// KR — the reference's front end on the GPU: PointCloudAligner::compute_transformation_and_inliers
// (src/ceres_slam/point_cloud_aligner.cpp:64-136), the 3-point RANSAC `compute_initial_guess`
// (dataset_problem.cpp:179-270) runs on every consecutive pose pair (400 hypotheses, threshold 4).
// For a sliding window of 2 it dominates the wall time of the reference drivers.
//
// A batch of independent point-cloud pairs is one launch: one CTA per pair, one thread per
// hypothesis.  The points of the pair and their projections live in shared memory (or are read from
// global memory when a pair is too large); each hypothesis is the closed 3-point Kabsch alignment
// (point_cloud_aligner.cpp:12-62: centroids, W = 1/3 sum (q - qbar)(p - pbar)^T, rotation
// U diag(1, 1, det U det V) V^T from the SVD of W — here a one-sided Jacobi SVD in registers), then
// the inlier count over all points by squared reprojection distance (:117-123).  The best
// hypothesis is the largest inlier set, the EARLIEST on ties (`>` at :127).
//
// The index triples are the reference's own: std::mt19937 seeded with 42 (:70-72) through
// std::uniform_int_distribution<uint>(0, n-1) with re-draws on duplicates (:82-90).  They depend on
// n only and are generated on the host (ransac_triples): the draw loop is sequential by nature.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <map>
#include <mutex>
#include <random>
#include <tuple>
#include <vector>

#include "kernels.cuh"

namespace cslam {

// std::uniform_int_distribution<unsigned>(0, n - 1) applied to std::mt19937, restated so that the
// draws do not depend on the libstdc++ the library happens to be built with:
//   variant 0 — scaling + rejection (libstdc++ up to GCC 10, the reference's era):
//               scaling = 0xFFFFFFFF / n; past = n * scaling; redraw while r >= past; r / scaling
//   variant 1 — Lemire's nearly divisionless method (libstdc++ from GCC 11 for 32-bit generators)
static inline uint32_t draw_index(std::mt19937& rng, uint32_t n, int variant) {
    if (variant == 0) {
        const uint64_t urngrange = 0xFFFFFFFFull, uerange = n;
        if (urngrange > uerange - 1) {
            const uint64_t scaling = urngrange / uerange, past = uerange * scaling;
            uint64_t ret;
            do ret = uint64_t(rng());
            while (ret >= past);
            return uint32_t(ret / scaling);
        }
        return uint32_t(rng());  // n == 2^32: the raw draw
    }
    const uint32_t range = n;  // uerange as a 32-bit value
    uint64_t product = uint64_t(uint32_t(rng())) * uint64_t(range);
    uint32_t low = uint32_t(product);
    if (low < range) {
        const uint32_t threshold = uint32_t(-range) % range;
        while (low < threshold) {
            product = uint64_t(uint32_t(rng())) * uint64_t(range);
            low = uint32_t(product);
        }
    }
    return uint32_t(product >> 32);
}

void ransac_triples(uint32_t n, uint32_t num_iters, int variant, uint32_t* out) {
    std::mt19937 rng(42);  // point_cloud_aligner.cpp:72
    for (uint32_t it = 0; it < num_iters; ++it) {
        uint32_t a = draw_index(rng, n, variant);
        uint32_t b = draw_index(rng, n, variant);
        while (b == a) b = draw_index(rng, n, variant);
        uint32_t c = draw_index(rng, n, variant);
        while (c == a || c == b) c = draw_index(rng, n, variant);
        out[3 * it] = a;
        out[3 * it + 1] = b;
        out[3 * it + 2] = c;
    }
}

namespace {

constexpr int RS_THREADS = 128;

__device__ __forceinline__ void project3(const CameraIntrinsics& c, const double* p, double* o) {
    const double iz = 1.0 / p[2];  // stereo_camera.hpp:79-84
    o[0] = c.fu * p[0] * iz + c.cu;
    o[1] = c.fv * p[1] * iz + c.cv;
    o[2] = c.fu * c.b * iz;
}

// Rotation U diag(1,1,det U det V) V^T of the SVD of a 3x3 matrix W (row-major) by one-sided
// Jacobi: columns of G = W V are orthogonalised; sigma_i = |g_i|, u_i = g_i / sigma_i.  The column
// with the smallest singular value (zero for three points) is replaced by the cross product of the
// other two on both sides, which is what the determinant correction amounts to.
__device__ void kabsch_rotation(const double* W, double* C) {
    double G[9], V[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
#pragma unroll
    for (int i = 0; i < 9; ++i) G[i] = W[i];
    for (int sweep = 0; sweep < 30; ++sweep) {
        double off = 0.0;
#pragma unroll
        for (int pq = 0; pq < 3; ++pq) {
            const int p = pq == 2 ? 1 : 0, q = pq == 0 ? 1 : 2;
            double al = 0, be = 0, ga = 0;
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                al += G[3 * r + p] * G[3 * r + p];
                be += G[3 * r + q] * G[3 * r + q];
                ga += G[3 * r + p] * G[3 * r + q];
            }
            if (ga == 0.0 || fabs(ga) <= 1e-300) continue;
            const double lim = 1e-16 * sqrt(al * be);
            if (fabs(ga) <= lim) continue;
            off = fmax(off, fabs(ga) / fmax(sqrt(al * be), 1e-300));
            const double zeta = (be - al) / (2.0 * ga);
            const double t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
            const double cs = 1.0 / sqrt(1.0 + t * t), sn = cs * t;
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                const double gp = G[3 * r + p], gq = G[3 * r + q];
                G[3 * r + p] = cs * gp - sn * gq;
                G[3 * r + q] = sn * gp + cs * gq;
                const double vp = V[3 * r + p], vq = V[3 * r + q];
                V[3 * r + p] = cs * vp - sn * vq;
                V[3 * r + q] = sn * vp + cs * vq;
            }
        }
        if (off < 1e-15) break;
    }
    double s[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) s[c] = sqrt(G[c] * G[c] + G[3 + c] * G[3 + c] + G[6 + c] * G[6 + c]);
    // the two largest singular values keep their columns; the third is rebuilt
    int k = 0;
    if (s[1] < s[k]) k = 1;
    if (s[2] < s[k]) k = 2;
    const int a = (k + 1) % 3, b = (k + 2) % 3;
    double ua[3], ub[3], va[3], vb[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        ua[r] = G[3 * r + a] / s[a];
        ub[r] = G[3 * r + b] / s[b];
        va[r] = V[3 * r + a];
        vb[r] = V[3 * r + b];
    }
    const double uc[3] = {ua[1] * ub[2] - ua[2] * ub[1], ua[2] * ub[0] - ua[0] * ub[2], ua[0] * ub[1] - ua[1] * ub[0]};
    const double vc[3] = {va[1] * vb[2] - va[2] * vb[1], va[2] * vb[0] - va[0] * vb[2], va[0] * vb[1] - va[1] * vb[0]};
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) C[3 * r + c] = ua[r] * va[c] + ub[r] * vb[c] + uc[r] * vc[c];
}

// T_1_0 = (r, C) from three correspondences p (frame 0) -> q (frame 1)
__device__ void align3(const double* p0, const double* p1, const double* p2, const double* q0, const double* q1,
                       const double* q2, double* T12) {
    double pb[3], qb[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        pb[r] = ((0.0 + p0[r]) + p1[r] + p2[r]) / 3.0;  // Point::Zero() += ... then /= size (:27-37)
        qb[r] = ((0.0 + q0[r]) + q1[r] + q2[r]) / 3.0;
    }
    double W[9];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c)
            W[3 * r + c] = ((q0[r] - qb[r]) * (p0[c] - pb[c]) + (q1[r] - qb[r]) * (p1[c] - pb[c]) + (q2[r] - qb[r]) * (p2[c] - pb[c])) / 3.0;
    double* C = T12 + 3;
    kabsch_rotation(W, C);
#pragma unroll
    for (int r = 0; r < 3; ++r) T12[r] = qb[r] - (C[3 * r] * pb[0] + C[3 * r + 1] * pb[1] + C[3 * r + 2] * pb[2]);
}

struct RansacArgs {
    CameraIntrinsics cam;
    const uint32_t* offsets;   // [n_pairs + 1]
    const double* pts0;        // [total][3]
    const double* pts1;
    const uint32_t* triples;   // [n_tables][num_iters][3]
    const uint32_t* table_of;  // [n_pairs]
    uint32_t num_iters;
    double thresh;
    int smem_points;           // points that fit the shared staging area
    double* T_out;             // [n_pairs][12]
    uint8_t* inlier_out;       // [total]
    uint32_t* count_out;       // [n_pairs]
};

__global__ void __launch_bounds__(RS_THREADS) ransac_kernel(RansacArgs A) {
    extern __shared__ __align__(16) double smem_rs[];
    __shared__ unsigned long long s_best;
    __shared__ double s_T[12];
    const int pair = blockIdx.x, tid = threadIdx.x;
    const uint32_t o0 = A.offsets[pair], n = A.offsets[pair + 1] - o0;
    const double* P0 = A.pts0 + 3ull * o0;
    const double* P1 = A.pts1 + 3ull * o0;
    if (n < 3) {
        // the reference cannot form a hypothesis (its index draw does not terminate): identity, no inliers
        if (tid < 12) A.T_out[12ull * pair + tid] = (tid == 3 || tid == 7 || tid == 11) ? 1.0 : 0.0;
        for (uint32_t i = tid; i < n; i += RS_THREADS) A.inlier_out[o0 + i] = 0;
        if (tid == 0) A.count_out[pair] = 0;
        return;
    }
    const bool staged = int(n) <= A.smem_points;
    double* sp0 = smem_rs;           // [n][3] frame-0 points
    double* sz1 = smem_rs + 3 * n;   // [n][3] projections of the frame-1 points
    if (staged) {
        for (uint32_t i = tid; i < n; i += RS_THREADS) {
            sp0[3 * i] = P0[3 * i];
            sp0[3 * i + 1] = P0[3 * i + 1];
            sp0[3 * i + 2] = P0[3 * i + 2];
            project3(A.cam, P1 + 3 * i, sz1 + 3 * i);
        }
    }
    if (tid == 0) s_best = 0ull;
    __syncthreads();
    auto count_inliers = [&](const double* T, bool write) -> uint32_t {
        uint32_t cnt = 0;
        const uint32_t i0 = write ? tid : 0, di = write ? RS_THREADS : 1;
        for (uint32_t i = i0; i < n; i += di) {
            double p[3], z1[3];
            if (staged) {
                p[0] = sp0[3 * i], p[1] = sp0[3 * i + 1], p[2] = sp0[3 * i + 2];
                z1[0] = sz1[3 * i], z1[1] = sz1[3 * i + 1], z1[2] = sz1[3 * i + 2];
            } else {
                p[0] = P0[3 * i], p[1] = P0[3 * i + 1], p[2] = P0[3 * i + 2];
                project3(A.cam, P1 + 3 * i, z1);
            }
            const double* C = T + 3;
            const double x[3] = {C[0] * p[0] + C[1] * p[1] + C[2] * p[2] + T[0], C[3] * p[0] + C[4] * p[1] + C[5] * p[2] + T[1],
                                 C[6] * p[0] + C[7] * p[1] + C[8] * p[2] + T[2]};
            double z0[3];
            project3(A.cam, x, z0);
            const double e0 = z1[0] - z0[0], e1 = z1[1] - z0[1], e2 = z1[2] - z0[2];
            const bool in = (e0 * e0 + e1 * e1 + e2 * e2) < A.thresh;  // :119-123
            cnt += in ? 1u : 0u;
            if (write) A.inlier_out[o0 + i] = in ? 1 : 0;
        }
        return cnt;
    };
    const uint32_t* tri = A.triples + 3ull * A.num_iters * A.table_of[pair];
    for (uint32_t h = tid; h < A.num_iters; h += RS_THREADS) {
        const uint32_t a = tri[3 * h], b = tri[3 * h + 1], c = tri[3 * h + 2];
        double T[12];
        align3(P0 + 3 * a, P0 + 3 * b, P0 + 3 * c, P1 + 3 * a, P1 + 3 * b, P1 + 3 * c, T);
        const uint32_t cnt = count_inliers(T, false);
        // largest count, earliest hypothesis on ties: order by (count, ~h)
        const unsigned long long key = (static_cast<unsigned long long>(cnt) << 32) | (0xFFFFFFFFu - h);
        atomicMax(&s_best, key);
    }
    __syncthreads();
    const unsigned long long best = s_best;
    const uint32_t best_cnt = uint32_t(best >> 32), best_h = 0xFFFFFFFFu - uint32_t(best & 0xFFFFFFFFu);
    if (best_cnt == 0) {
        // no hypothesis had an inlier: best_T_1_0 stays default-constructed (identity) (:66-68,:127)
        if (tid < 12) A.T_out[12ull * pair + tid] = (tid == 3 || tid == 7 || tid == 11) ? 1.0 : 0.0;
        for (uint32_t i = tid; i < n; i += RS_THREADS) A.inlier_out[o0 + i] = 0;
        if (tid == 0) A.count_out[pair] = 0;
        return;
    }
    if (tid == 0) {
        const uint32_t a = tri[3 * best_h], b = tri[3 * best_h + 1], c = tri[3 * best_h + 2];
        double T[12];
        align3(P0 + 3 * a, P0 + 3 * b, P0 + 3 * c, P1 + 3 * a, P1 + 3 * b, P1 + 3 * c, T);
        for (int k = 0; k < 12; ++k) {
            s_T[k] = T[k];
            A.T_out[12ull * pair + k] = T[k];
        }
        A.count_out[pair] = best_cnt;
    }
    __syncthreads();
    double T[12];
    for (int k = 0; k < 12; ++k) T[k] = s_T[k];
    count_inliers(T, true);
}

}  // namespace

// Per-call resources, pooled process-wide like the window batch's arena (kernels_window.cu): a stream, a pinned
// staging block and a device block, all grow-only.  A sliding-window driver calls this once per window with a few
// KB of points: creating a stream and eight stream-ordered buffers per call cost more than the kernel.
struct RansacArena {
    int device = -1;
    cudaStream_t stream = nullptr;
    char* pinned = nullptr;
    char* dev = nullptr;
    size_t cap = 0;
};
static std::mutex g_rs_mu;
static std::vector<RansacArena*> g_rs_free;

static RansacArena* ransac_arena_take(int device, size_t bytes) {
    RansacArena* a = nullptr;
    {
        std::lock_guard<std::mutex> g(g_rs_mu);
        for (size_t i = 0; i < g_rs_free.size(); ++i)
            if (g_rs_free[i]->device == device) {
                a = g_rs_free[i];
                g_rs_free.erase(g_rs_free.begin() + long(i));
                break;
            }
    }
    if (!a) {
        a = new RansacArena;
        a->device = device;
        CSLAM_CUDA(cudaStreamCreateWithFlags(&a->stream, cudaStreamNonBlocking));
    }
    if (a->cap < bytes) {
        if (a->pinned) cudaFreeHost(a->pinned);
        if (a->dev) cudaFree(a->dev);
        a->pinned = a->dev = nullptr;
        a->cap = 0;
        const size_t want = std::max(bytes + bytes / 2, size_t(1) << 20);
        if (cudaMallocHost(reinterpret_cast<void**>(&a->pinned), want) != cudaSuccess ||
            cudaMalloc(reinterpret_cast<void**>(&a->dev), want) != cudaSuccess) {
            if (a->pinned) cudaFreeHost(a->pinned);
            a->pinned = nullptr;
            std::lock_guard<std::mutex> g(g_rs_mu);
            g_rs_free.push_back(a);
            throw CudaError("RANSAC staging allocation failed");
        }
        a->cap = want;
    }
    return a;
}

void ransac_align_batch(int device, uint32_t n_pairs, const uint32_t* offsets, const double* pts0, const double* pts1,
                        const double* intr5, uint32_t num_iters, double thresh, int rng_variant, double* T12_out,
                        uint8_t* inlier_out, uint32_t* n_inliers_out) {
    if (n_pairs == 0) return;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0)
        throw CudaError("no CUDA device: the cslam_b200 back end has no CPU fallback");
    CSLAM_CUDA(cudaSetDevice(device));
    const uint32_t total = offsets[n_pairs];
    // one table of index triples per distinct cloud size (they depend on n only: cached across calls)
    static std::mutex tri_mu;
    static std::map<std::tuple<uint32_t, uint32_t, int>, std::vector<uint32_t>> tri_cache;
    std::map<uint32_t, uint32_t> table;
    std::vector<uint32_t> table_of(n_pairs), triples;
    uint32_t n_max = 0;
    for (uint32_t p = 0; p < n_pairs; ++p) {
        const uint32_t n = offsets[p + 1] - offsets[p];
        n_max = std::max(n_max, n);
        if (n < 3) {
            table_of[p] = 0;
            continue;
        }
        auto it = table.find(n);
        if (it == table.end()) {
            const uint32_t id = uint32_t(triples.size() / (3ull * num_iters));
            triples.resize(triples.size() + 3ull * num_iters);
            {
                std::lock_guard<std::mutex> g(tri_mu);
                const auto key = std::make_tuple(n, num_iters, rng_variant);
                auto itc = tri_cache.find(key);
                if (itc == tri_cache.end()) {
                    if (tri_cache.size() > 4096) tri_cache.clear();  // bounded: a long track has a few hundred distinct sizes
                    itc = tri_cache.emplace(key, std::vector<uint32_t>(3ull * num_iters)).first;
                    ransac_triples(n, num_iters, rng_variant, itc->second.data());
                }
                std::memcpy(triples.data() + 3ull * num_iters * id, itc->second.data(), 12ull * num_iters);
            }
            it = table.emplace(n, id).first;
        }
        table_of[p] = it->second;
    }
    if (triples.empty()) triples.assign(3ull * std::max<uint32_t>(num_iters, 1), 0);

    // layout (256-byte aligned pieces): [offsets | triples | table_of | pts0 | pts1 || T | counts | inliers]
    size_t cursor = 0;
    auto place = [&](size_t bytes) {
        const size_t off = cursor;
        cursor = (cursor + bytes + 255) & ~size_t(255);
        return off;
    };
    const size_t npts = std::max<uint32_t>(total, 1);
    const size_t o_off = place((size_t(n_pairs) + 1) * 4), o_tri = place(triples.size() * 4), o_tab = place(size_t(n_pairs) * 4),
                 o_p0 = place(npts * 24), o_p1 = place(npts * 24);
    const size_t up_bytes = cursor;
    const size_t o_T = place(size_t(n_pairs) * 96), o_cnt = place(size_t(n_pairs) * 4), o_in = place(npts);
    const size_t all_bytes = cursor;
    RansacArena* arena = ransac_arena_take(device, all_bytes);
    struct Give {
        RansacArena* a;
        ~Give() {
            std::lock_guard<std::mutex> g(g_rs_mu);
            g_rs_free.push_back(a);
        }
    } give{arena};
    cudaStream_t s = arena->stream;
    char *hp = arena->pinned, *dp = arena->dev;
    std::memcpy(hp + o_off, offsets, (size_t(n_pairs) + 1) * 4);
    std::memcpy(hp + o_tri, triples.data(), triples.size() * 4);
    std::memcpy(hp + o_tab, table_of.data(), size_t(n_pairs) * 4);
    if (total) {
        std::memcpy(hp + o_p0, pts0, size_t(total) * 24);
        std::memcpy(hp + o_p1, pts1, size_t(total) * 24);
    }
    CSLAM_CUDA(cudaMemcpyAsync(dp, hp, up_bytes, cudaMemcpyHostToDevice, s));
    RansacArgs A;
    A.cam = CameraIntrinsics{intr5[0], intr5[1], intr5[2], intr5[3], intr5[4]};
    A.offsets = reinterpret_cast<const uint32_t*>(dp + o_off);
    A.pts0 = reinterpret_cast<const double*>(dp + o_p0);
    A.pts1 = reinterpret_cast<const double*>(dp + o_p1);
    A.triples = reinterpret_cast<const uint32_t*>(dp + o_tri);
    A.table_of = reinterpret_cast<const uint32_t*>(dp + o_tab);
    A.num_iters = num_iters;
    A.thresh = thresh;
    const int cap = 4096;  // points staged in shared memory: 48 B each
    A.smem_points = int(std::min<uint32_t>(n_max, cap));
    const size_t smem = size_t(A.smem_points) * 6 * sizeof(double);
    if (smem > 48 * 1024)
        CSLAM_CUDA(cudaFuncSetAttribute(ransac_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    A.T_out = reinterpret_cast<double*>(dp + o_T);
    A.inlier_out = reinterpret_cast<uint8_t*>(dp + o_in);
    A.count_out = reinterpret_cast<uint32_t*>(dp + o_cnt);
    ransac_kernel<<<n_pairs, RS_THREADS, smem, s>>>(A);
    g_kernel_launches.fetch_add(1, std::memory_order_relaxed);
    CSLAM_CUDA(cudaGetLastError());
    CSLAM_CUDA(cudaMemcpyAsync(hp + o_T, dp + o_T, all_bytes - o_T, cudaMemcpyDeviceToHost, s));
    CSLAM_CUDA(cudaStreamSynchronize(s));
    std::memcpy(T12_out, hp + o_T, size_t(n_pairs) * 96);
    if (total && inlier_out) std::memcpy(inlier_out, hp + o_in, total);
    if (n_inliers_out) std::memcpy(n_inliers_out, hp + o_cnt, size_t(n_pairs) * 4);
}

}  // namespace cslam
