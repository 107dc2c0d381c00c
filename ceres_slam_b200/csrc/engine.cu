// Host side of the cslam_b200 back end: structure analysis, device residency, and the
// Levenberg-Marquardt loop that drives the kernels (Ceres trust-region semantics, SURVEY.md
// App. B; the same rules the oracle restates in oracle/problem.hpp::solve).
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <limits>
#include <mutex>
#include <numeric>
#include <thread>

#include "comm.h"
#include "kernels.cuh"

namespace cslam {

namespace {
// CSLAM_TIMING=1 prints host-side phase times of upload() to stderr
struct PhaseTimer {
    bool on;
    std::chrono::steady_clock::time_point t0;
    PhaseTimer() : on(std::getenv("CSLAM_TIMING") != nullptr), t0(std::chrono::steady_clock::now()) {}
    void lap(const char* what) {
        if (!on) return;
        const auto t1 = std::chrono::steady_clock::now();
        std::fprintf(stderr, "[cslam timing] %-28s %8.2f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
        t0 = t1;
    }
};
}  // namespace

// The 1 KB pinned read-back block of a handle comes from a process-wide free list: cudaMallocHost was
// measured at 3..28 ms per call next to a busy device, which is most of a small solve.
namespace {
std::mutex g_pinned_mu;
std::vector<double*> g_pinned_free;
double* pinned_scalars_get() {
    {
        std::lock_guard<std::mutex> g(g_pinned_mu);
        if (!g_pinned_free.empty()) {
            double* p = g_pinned_free.back();
            g_pinned_free.pop_back();
            return p;
        }
    }
    double* p = nullptr;
    CSLAM_CUDA(cudaMallocHost(&p, 128 * sizeof(double)));
    return p;
}
void pinned_scalars_put(double* p) {
    std::lock_guard<std::mutex> g(g_pinned_mu);
    g_pinned_free.push_back(p);
}
}  // namespace

Engine::Engine(const cslam_options& o) : opt(o) {}

Engine::~Engine() {
    if (ev_a) cudaEventDestroy(ev_a);
    if (ev_b) cudaEventDestroy(ev_b);
    if (ev_c) cudaEventDestroy(ev_c);
    if (ev_d) cudaEventDestroy(ev_d);
    if (ev_fork) cudaEventDestroy(ev_fork);
    if (ph.fan_graph) cudaGraphExecDestroy(ph.fan_graph);
    band_sets.clear();
    if (h_pinned) {
        if (stream) cudaStreamSynchronize(stream);  // nothing of this handle may still write into the block
        pinned_scalars_put(h_pinned);
    }
    if (own_stream && stream) cudaStreamDestroy(stream);
    if (nccl_comm) comm_destroy(nccl_comm);
}

void Engine::set_stream(cudaStream_t s) {
    if (own_stream && stream) cudaStreamDestroy(stream);
    stream = s;
    own_stream = false;
}

static void ensure_device(Engine* e, cudaStream_t* stream, bool* own, cudaEvent_t* a, cudaEvent_t* b, cudaEvent_t* c,
                          cudaEvent_t* d, double** pinned) {
    PhaseTimer et;
    int count = 0;
    cudaError_t st = cudaGetDeviceCount(&count);
    if (st != cudaSuccess || count == 0)
        throw CudaError("no CUDA device: the cslam_b200 back end has no CPU fallback");
    CSLAM_CUDA(cudaSetDevice(e->opt.device));
    {
        // keep freed blocks in the stream-ordered pool: repeated solves reuse them
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, e->opt.device) == cudaSuccess) {
            unsigned long long keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
    }
    et.lap("  device / pool");
    if (!*stream) {
        CSLAM_CUDA(cudaStreamCreateWithFlags(stream, cudaStreamNonBlocking));
        *own = true;
    }
    et.lap("  stream");
    if (!*a) {
        CSLAM_CUDA(cudaEventCreate(a));
        CSLAM_CUDA(cudaEventCreate(b));
        CSLAM_CUDA(cudaEventCreate(c));
        CSLAM_CUDA(cudaEventCreate(d));
    }
    et.lap("  events");
    if (!*pinned) *pinned = pinned_scalars_get();
    et.lap("  pinned scalars");
}

void Engine::prof_begin(int) {
    if (opt.profile_kernels) CSLAM_CUDA(cudaEventRecord(ev_c, stream));
}
void Engine::prof_end(int k) {
    if (!opt.profile_kernels) return;
    CSLAM_CUDA(cudaEventRecord(ev_d, stream));
    CSLAM_CUDA(cudaEventSynchronize(ev_d));
    float ms = 0;
    CSLAM_CUDA(cudaEventElapsedTime(&ms, ev_c, ev_d));
    prof.ms[k] += ms;
    prof.launches[k] += 1;
}

void Engine::read_scalars(const double* dev, double* host, int n) {
    CSLAM_CUDA(cudaMemcpyAsync(h_pinned, dev, n * sizeof(double), cudaMemcpyDeviceToHost, stream));
    CSLAM_CUDA(cudaStreamSynchronize(stream));
    std::memcpy(host, h_pinned, n * sizeof(double));
}

DevView Engine::view(const double* poses, const double* points) const {
    DevView v;
    v.cam = cam;
    v.n_cams = int(n_poses);
    v.n_free = n_free;
    v.n_lm = n_lm;
    v.n_obs = n_obs;
    v.poses = poses;
    v.points = points;
    v.cam_free = d_cam_free.p;
    v.lm_base = d_lm_base.p;
    v.lm_stride = d_lm_stride.p;
    v.lm_cnt = d_lm_cnt.p;
    v.obs_cam = d_obs_cam.p;
    v.obs_u = d_obs_u.p;
    v.obs_v = d_obs_v.p;
    v.obs_d = d_obs_d.p;
    v.obs_W = d_obs_W.p;
    v.W_per_obs = st_W_per_obs;
    v.sc_p = d_sc_p.p;
    v.sc_l = d_sc_l.p;
    v.s_rowptr = d_s_rowptr.p;
    v.s_col = d_s_col.p;
    return v;
}

// -------------------------------------------------------------------------------------------------
// Structure analysis: which blocks exist (dataset_vo.cpp:40-62), landmark-major order, the
// co-visibility pattern of the reduced camera system, and this rank's landmark shard.
// -------------------------------------------------------------------------------------------------
// Run fn(t, lo, hi) over [0, n) split into contiguous chunks, one per host thread.
template <class F>
static void parallel_chunks(size_t n, size_t min_chunk, F&& fn) {
    static const size_t hw = [] {
        const char* e = std::getenv("CSLAM_THREADS");
        const unsigned v = e ? unsigned(std::atoi(e)) : std::thread::hardware_concurrency();
        return size_t(std::max(1u, std::min(32u, v)));
    }();
    const size_t nt = std::max<size_t>(1, std::min(hw, n / std::max<size_t>(1, min_chunk)));
    if (nt <= 1) {
        fn(0, size_t(0), n);
        return;
    }
    std::vector<std::thread> th;
    std::exception_ptr err;
    std::mutex mu;
    for (size_t t = 0; t < nt; ++t)
        th.emplace_back([&, t]() {
            try {
                fn(int(t), n * t / nt, n * (t + 1) / nt);
            } catch (...) {
                std::lock_guard<std::mutex> g(mu);
                err = std::current_exception();
            }
        });
    for (auto& x : th) x.join();
    if (err) std::rethrow_exception(err);
}

void Engine::build_structure() {
    if (!h_poses || n_poses == 0) throw std::invalid_argument("poses not set");
    if (n_st > 0 && (!h_points || n_points == 0)) throw std::invalid_argument("points not set");
    if (n_st >= (1ull << 32)) throw std::invalid_argument("more than 2^32 stereo blocks");
    PhaseTimer bt;
    const size_t kChunk = 1 << 16;
    const int max_threads = 32;

    // ---- which blocks exist; observations per point (dataset_vo.cpp:40-62) ----------------------
    std::vector<uint32_t> cnt(n_points, 0);
    std::vector<std::vector<uint8_t>> used_t(max_threads);
    parallel_chunks(n_st, kChunk, [&](int t, size_t lo, size_t hi) {
        std::vector<uint8_t>& used = used_t[t];
        used.assign(n_poses, 0);
        for (size_t i = lo; i < hi; ++i) {
            const uint32_t c = st_cam[i], j = st_pt[i];
            if (c >= n_poses || j >= n_points) throw std::invalid_argument("stereo block index out of range");
            used[c] = 1;
            __atomic_fetch_add(&cnt[j], 1u, __ATOMIC_RELAXED);
        }
    });
    std::vector<uint8_t> used(n_poses, 0);
    for (auto& u : used_t)
        for (size_t k = 0; k < u.size(); ++k) used[k] |= u[k];
    for (auto& s : suns) {
        if (s.cam >= n_poses) throw std::invalid_argument("sun block index out of range");
        used[s.cam] = 1;
    }
    for (auto& p : priors) {
        if (p.cam >= n_poses) throw std::invalid_argument("prior block index out of range");
        used[p.cam] = 1;
    }
    cam_free_h.assign(n_poses, -1);
    free_cams_h.clear();
    for (uint32_t k = 0; k < n_poses; ++k)
        if (used[k] && !pose_const[k]) {
            cam_free_h[k] = int(free_cams_h.size());
            free_cams_h.push_back(int(k));
        }
    n_free = int(free_cams_h.size());
    bt.lap("  count / free cams");

    // ---- observation lists per point, each sorted by (camera, block index) ---------------------
    std::vector<uint32_t> ptr_u(size_t(n_points) + 1, 0);
    for (uint32_t j = 0; j < n_points; ++j) ptr_u[j + 1] = ptr_u[j] + cnt[j];
    std::unique_ptr<uint32_t[]> obs_u_store(new uint32_t[std::max<size_t>(n_st, 1)]);  // filled in parallel below
    uint32_t* const obs_u = obs_u_store.get();
    {
        std::vector<uint32_t> fill(ptr_u.begin(), ptr_u.end() - 1);
        parallel_chunks(n_st, kChunk, [&](int, size_t lo, size_t hi) {
            for (size_t i = lo; i < hi; ++i) obs_u[__atomic_fetch_add(&fill[st_pt[i]], 1u, __ATOMIC_RELAXED)] = uint32_t(i);
        });
    }
    std::vector<uint32_t> mincam(n_points, 0xffffffffu);
    parallel_chunks(n_points, 1 << 12, [&](int, size_t lo, size_t hi) {
        for (size_t j = lo; j < hi; ++j) {
            uint32_t* first = obs_u + ptr_u[j];
            const uint32_t len = cnt[j];
            if (!len) continue;
            auto less = [&](uint32_t a, uint32_t b) { return st_cam[a] != st_cam[b] ? st_cam[a] < st_cam[b] : a < b; };
            if (len <= 32) {
                for (uint32_t x = 1; x < len; ++x) {
                    const uint32_t key = first[x];
                    uint32_t y = x;
                    while (y > 0 && less(key, first[y - 1])) {
                        first[y] = first[y - 1];
                        --y;
                    }
                    first[y] = key;
                }
            } else {
                std::sort(first, first + len, less);
            }
            mincam[j] = st_cam[first[0]];
        }
    });
    // landmarks in order of the first camera that sees them (counting sort, stable in user id)
    std::vector<uint32_t> bucket(size_t(n_poses) + 1, 0);
    uint32_t n_active = 0;
    for (uint32_t j = 0; j < n_points; ++j)
        if (cnt[j]) {
            bucket[mincam[j] + 1]++;
            n_active++;
        }
    for (uint32_t k = 0; k < n_poses; ++k) bucket[k + 1] += bucket[k];
    const std::vector<uint32_t> bstart(bucket);  // landmarks [bstart[c], bstart[c+1]) start at camera c
    std::vector<uint32_t> all_lm(n_active);
    for (uint32_t j = 0; j < n_points; ++j)
        if (cnt[j]) all_lm[bucket[mincam[j]]++] = j;
    std::vector<uint32_t> all_ptr(size_t(n_active) + 1, 0);
    for (uint32_t a = 0; a < n_active; ++a) all_ptr[a + 1] = all_ptr[a] + cnt[all_lm[a]];
    auto lm_obs = [&](uint32_t a) { return obs_u + ptr_u[all_lm[a]]; };  // cnt[all_lm[a]] entries
    auto lm_len = [&](uint32_t a) { return cnt[all_lm[a]]; };
    bt.lap("  landmark-major order");

    // ---- reduced camera system pattern (global: every rank derives the same one) ---------------
    {
        // pair (a, b), a <= b, of free cameras that share a landmark: bit (b - a) of row a's mask
        // for offsets below 64, an explicit list beyond
        std::vector<std::vector<uint64_t>> mask_t(max_threads);
        std::vector<std::vector<std::pair<int, int>>> far_t(max_threads);
        parallel_chunks(n_active, 1 << 12, [&](int t, size_t lo, size_t hi) {
            std::vector<uint64_t>& mask = mask_t[t];
            mask.assign(size_t(std::max(n_free, 1)), 0);
            std::vector<int> fr;
            for (size_t a = lo; a < hi; ++a) {
                const uint32_t* ob = lm_obs(uint32_t(a));
                const uint32_t len = lm_len(uint32_t(a));
                fr.clear();
                for (uint32_t k = 0; k < len; ++k) {
                    const int f = cam_free_h[st_cam[ob[k]]];
                    if (f >= 0 && (fr.empty() || fr.back() != f)) fr.push_back(f);
                }
                for (size_t x = 0; x < fr.size(); ++x)
                    for (size_t y = x + 1; y < fr.size(); ++y) {
                        const int d = fr[y] - fr[x];
                        if (d < 64)
                            mask[fr[x]] |= 1ull << d;
                        else
                            far_t[t].push_back({fr[x], fr[y]});
                    }
            }
        });
        std::vector<uint64_t> mask(size_t(std::max(n_free, 1)), 0);
        for (auto& m : mask_t)
            for (size_t k = 0; k < m.size(); ++k) mask[k] |= m[k];
        std::vector<std::pair<int, int>> far;
        for (auto& f : far_t) far.insert(far.end(), f.begin(), f.end());
        std::sort(far.begin(), far.end());
        far.erase(std::unique(far.begin(), far.end()), far.end());
        s_rowptr_h.assign(size_t(n_free) + 1, 0);
        s_col_h.clear();
        size_t fx = 0;
        for (int a = 0; a < n_free; ++a) {
            s_col_h.push_back(a);
            for (int d = 1; d < 64; ++d)
                if (mask[a] >> d & 1) s_col_h.push_back(a + d);
            for (; fx < far.size() && far[fx].first == a; ++fx) s_col_h.push_back(far[fx].second);
            s_rowptr_h[a + 1] = int(s_col_h.size());
        }
        nnzU = int(s_col_h.size());
    }
    bt.lap("  S pattern");

    // ---- this rank's shard: contiguous landmark range balanced by observation count ----
    uint32_t lo = 0, hi = n_active;
    if (n_ranks > 1) {
        auto cut = [&](int r) -> uint32_t {
            const uint64_t target = n_st * uint64_t(r) / uint64_t(n_ranks);
            return uint32_t(std::lower_bound(all_ptr.begin(), all_ptr.end(), uint32_t(target)) - all_ptr.begin());
        };
        lo = std::min(cut(rank), n_active);
        hi = rank == n_ranks - 1 ? n_active : std::min(cut(rank + 1), n_active);
    }
    n_lm = int(hi - lo);
    lm_lo = 0;
    lm_hi = n_lm;
    n_obs = (long long)all_ptr[hi] - (long long)all_ptr[lo];

    // ---- group landmarks with identical camera lists (grouped Schur kernel) ----
    // A landmark's sort key is (first camera, track length, hash of the camera list, index); the
    // landmark order is already first-camera major, so each camera's bucket is sorted on its own.
    const uint32_t nl = hi - lo;
    std::vector<uint64_t> khash(nl);
    std::vector<uint8_t> kok(nl);
    parallel_chunks(nl, 1 << 12, [&](int, size_t x0, size_t x1) {
        for (size_t x = x0; x < x1; ++x) {
            const uint32_t a = lo + uint32_t(x);
            const uint32_t* ob = lm_obs(a);
            const uint32_t len = lm_len(a);
            bool ok = opt.schur_path != 1 && len <= uint32_t(kGroupLmax);
            uint64_t h = 1469598103934665603ull;
            uint32_t prev = 0xffffffffu;
            for (uint32_t k = 0; k < len; ++k) {
                const uint32_t c = st_cam[ob[k]];
                if (c == prev) ok = false;  // the same camera twice: generic kernel only
                prev = c;
                h = (h ^ c) * 1099511628211ull;
            }
            khash[x] = h;
            kok[x] = ok ? 1 : 0;
        }
    });
    auto same_cams = [&](uint32_t a, uint32_t b) {
        const uint32_t len = lm_len(a);
        const uint32_t *oa = lm_obs(a), *ob = lm_obs(b);
        for (uint32_t k = 0; k < len; ++k)
            if (st_cam[oa[k]] != st_cam[ob[k]]) return false;
        return true;
    };
    // per camera bucket: sorted eligible landmarks and the runs of identical camera lists
    std::vector<uint32_t> sorted_a(nl);           // eligible landmarks, bucket by bucket
    std::vector<uint32_t> sorted_off(size_t(n_poses) + 1, 0);
    for (uint32_t c = 0; c < n_poses; ++c) {
        const uint32_t b0 = std::max(bstart[c], lo), b1 = std::min(bstart[c + 1], hi);
        uint32_t k = 0;
        for (uint32_t a = b0; a < b1; ++a) k += kok[a - lo];
        sorted_off[c + 1] = sorted_off[c] + k;
    }
    std::vector<uint8_t> run_start(nl, 0);
    parallel_chunks(n_poses, 16, [&](int, size_t c0, size_t c1) {
        for (size_t c = c0; c < c1; ++c) {
            const uint32_t b0 = std::max(bstart[c], lo), b1 = std::min(bstart[c + 1], hi);
            uint32_t* out = sorted_a.data() + sorted_off[c];
            uint32_t k = 0;
            for (uint32_t a = b0; a < b1; ++a)
                if (kok[a - lo]) out[k++] = a;
            std::sort(out, out + k, [&](uint32_t x, uint32_t y) {
                const uint32_t lx = lm_len(x), ly = lm_len(y);
                if (lx != ly) return lx < ly;
                if (khash[x - lo] != khash[y - lo]) return khash[x - lo] < khash[y - lo];
                return x < y;
            });
            for (uint32_t x = 0; x < k;) {
                run_start[sorted_off[c] + x] = 1;
                uint32_t y = x + 1;
                while (y < k && lm_len(out[y]) == lm_len(out[x]) && khash[out[y] - lo] == khash[out[x] - lo] &&
                       same_cams(out[x], out[y]))
                    ++y;
                x = y;
            }
        }
    });
    const uint32_t n_ok = sorted_off[n_poses];
    // Exact groups below 24 landmarks are left to the ragged pass, where their first-camera bucket's landmarks share ONE
    // group: every group is a work item with its own staging and pipeline fill, and real tracks produce thousands of
    // groups of 4 .. 20 (ragged 5 k-pose track: Schur build 2.90 -> 2.27 ms; CSLAM_EXACT_MIN_GROUP: A/B knob)
    static const int exact_min_env = [] {
        const char* e = std::getenv("CSLAM_EXACT_MIN_GROUP");
        return e ? std::atoi(e) : 24;
    }();
    const size_t min_group = opt.schur_path == 2 ? 1 : 4;
    const size_t min_exact = opt.schur_path == 2 ? 1 : (exact_min_env > 0 && want_ragged && !lighting_in_solve() ? size_t(exact_min_env) : min_group);
    g_L_h.clear(); g_G_h.clear(); g_lm0_h.clear(); g_obs0_h.clear(); g_off_h.clear(); g_cams_h.clear();
    g_blk_off_h.clear(); g_blk_h.clear(); item_group_h.clear(); item_j0_h.clear(); item_n_h.clear();
    std::vector<uint32_t> g_first;  // position of each group's first landmark in sorted_a
    std::vector<uint8_t> grouped(nl, 0);
    uint32_t obs_cursor = 0, lm_cursor = 0;
    int cams_cursor = 0, blk_cursor = 0;
    std::vector<std::pair<int, int>> items_small, items_large;  // (group, j0)
    for (uint32_t x = 0; x < n_ok;) {
        uint32_t y = x + 1;
        while (y < n_ok && !run_start[y]) ++y;
        const uint32_t G = y - x;
        if (G >= min_exact) {
            const int L = int(lm_len(sorted_a[x]));
            const int gid = int(g_L_h.size());
            g_L_h.push_back(L);
            g_G_h.push_back(int(G));
            g_lm0_h.push_back(int(lm_cursor));
            g_obs0_h.push_back(obs_cursor);
            g_off_h.push_back(cams_cursor);
            g_blk_off_h.push_back(blk_cursor);
            g_first.push_back(x);
            cams_cursor += L;
            blk_cursor += L * (L + 1) / 2;
            lm_cursor += G;
            obs_cursor += G * uint32_t(L);
            for (uint32_t j0 = 0; j0 < G; j0 += kItemMax) (L <= 10 ? items_small : items_large).push_back({gid, int(j0)});
            max_group_L = std::max(max_group_L, L);
        }
        x = y;
    }
    const size_t n_exact = g_L_h.size();
    // ---- second pass: "ragged" groups.  A landmark needs no identical twin: the still ungrouped landmarks
    // of a first-camera bucket whose cameras all lie within a window of 10 (the DMMA kernel's tile) form one
    // group whose camera list is the union of theirs; each landmark sees a subset (per-landmark slot maps).
    // Real stereo tracks (variable length, drop-outs) almost never share an exact list.
    struct RagGroup {
        std::vector<uint32_t> lms;
        std::vector<int> cams;
        int K = 0;
    };
    std::vector<RagGroup> rag;
    g_map_off_h.assign(n_exact, -1);
    std::vector<std::pair<int, int>> items_rag;
    long long map_cursor = 0;
    if (want_ragged && opt.schur_path != 1 && !lighting_in_solve()) {
        std::vector<uint8_t> primary(nl, 0);
        for (size_t g = 0; g < n_exact; ++g)
            for (int jl = 0; jl < g_G_h[g]; ++jl) primary[sorted_a[g_first[g] + jl] - lo] = 1;
        constexpr int kRagWindow = 10;
        for (uint32_t c = 0; c < n_poses; ++c) {
            const uint32_t b0 = std::max(bstart[c], lo), b1 = std::min(bstart[c + 1], hi);
            RagGroup rg;
            for (uint32_t a = b0; a < b1; ++a) {
                if (!kok[a - lo] || primary[a - lo]) continue;
                const uint32_t* ob = lm_obs(a);
                const uint32_t len = lm_len(a);
                if (st_cam[ob[len - 1]] - c >= uint32_t(kRagWindow)) continue;   // (camera lists are ascending)
                rg.lms.push_back(a);
                rg.K = std::max(rg.K, int(len));
                for (uint32_t k = 0; k < len; ++k) rg.cams.push_back(int(st_cam[ob[k]]));
            }
            if (rg.lms.size() < min_group) continue;
            std::sort(rg.cams.begin(), rg.cams.end());
            rg.cams.erase(std::unique(rg.cams.begin(), rg.cams.end()), rg.cams.end());
            const int L = int(rg.cams.size()), G = int(rg.lms.size());
            const int gid = int(g_L_h.size());
            g_L_h.push_back(L);
            g_G_h.push_back(G);
            g_lm0_h.push_back(int(lm_cursor));
            g_obs0_h.push_back(obs_cursor);
            g_off_h.push_back(cams_cursor);
            g_blk_off_h.push_back(blk_cursor);
            g_map_off_h.push_back(int(map_cursor));
            cams_cursor += L;
            blk_cursor += L * (L + 1) / 2;
            lm_cursor += uint32_t(G);
            obs_cursor += uint32_t(G) * uint32_t(rg.K);
            map_cursor += (long long)(L + rg.K) * G;
            for (int j0 = 0; j0 < G; j0 += kItemMax) items_rag.push_back({gid, j0});
            max_group_L = std::max(max_group_L, L);
            rag.push_back(std::move(rg));
        }
    }
    n_items_rag = int(items_rag.size());
    g_map_h.assign(size_t(std::max<long long>(map_cursor, 1)), 0xff);
    const size_t n_groups = n_exact;
    n_lm_grouped = int(lm_cursor);
    n_items_small = int(items_small.size());
    for (auto* lst : {&items_small, &items_large, &items_rag})
        for (auto& it : *lst) {
            item_group_h.push_back(it.first);
            item_j0_h.push_back(it.second);
            item_n_h.push_back(std::min(kItemMax, g_G_h[it.first] - it.second));
        }
    g_cams_h.assign(size_t(cams_cursor), 0);
    g_blk_h.assign(size_t(blk_cursor), -1);
    lm_user_h.assign(nl, 0);
    lm_base_h.assign(nl, 0);
    lm_stride_h.assign(nl, 0);
    lm_cnt_h.assign(nl, 0);
    // ragged groups store K rows per landmark column (K = the longest member): their padding counts as storage
    n_obs_true = n_obs;
    {
        long long rest_obs = 0;
        std::vector<uint8_t> in_group(nl, 0);
        for (size_t g = 0; g < n_exact; ++g)
            for (int jl = 0; jl < g_G_h[g]; ++jl) in_group[sorted_a[g_first[g] + jl] - lo] = 1;
        for (auto& rg : rag)
            for (uint32_t a : rg.lms) in_group[a - lo] = 1;
        for (uint32_t x = 0; x < nl; ++x)
            if (!in_group[x]) rest_obs += lm_len(lo + x);
        n_obs = (long long)obs_cursor + rest_obs;
    }
    obs_user_n = size_t(n_obs);
    obs_user_h.reset(new uint32_t[std::max<size_t>(obs_user_n, 1)]);
    // pair -> block table of a camera list (-1: a constant camera, or a pair no landmark co-observes)
    auto fill_blocks = [&](int t, const int* fr, int L) {
        for (int i = 0; i < L; ++i)
            for (int k = i; k < L; ++k) {
                int e = -1;
                if (fr[i] >= 0 && fr[k] >= 0) {
                    auto b0 = s_col_h.begin() + s_rowptr_h[fr[i]], b1 = s_col_h.begin() + s_rowptr_h[fr[i] + 1];
                    auto it = std::lower_bound(b0, b1, fr[k]);
                    if (it != b1 && *it == fr[k]) e = int(it - s_col_h.begin());
                }
                g_blk_h[size_t(t++)] = e;
            }
    };
    parallel_chunks(rag.size(), 8, [&](int, size_t r0, size_t r1) {
        for (size_t r = r0; r < r1; ++r) {
            const size_t g = n_exact + r;
            const RagGroup& rg = rag[r];
            const int L = g_L_h[g], G = g_G_h[g], K = rg.K;
            int fr[kGroupLmax];
            for (int i = 0; i < L; ++i) {
                g_cams_h[size_t(g_off_h[g]) + i] = rg.cams[i];
                fr[i] = cam_free_h[rg.cams[i]];
            }
            fill_blocks(g_blk_off_h[g], fr, L);
            const uint32_t base = g_obs0_h[g];
            unsigned char* inv = g_map_h.data() + g_map_off_h[g];
            unsigned char* fwd = inv + size_t(L) * G;
            for (int jl = 0; jl < G; ++jl) {
                const uint32_t a = rg.lms[jl];
                const size_t li = size_t(g_lm0_h[g]) + jl;
                grouped[a - lo] = 1;
                const uint32_t* ob = lm_obs(a);
                const int len = int(lm_len(a));
                lm_user_h[li] = all_lm[a];
                lm_base_h[li] = base + jl;
                lm_stride_h[li] = uint32_t(G);
                lm_cnt_h[li] = uint32_t(len);
                for (int k = 0; k < K; ++k) obs_user_h[size_t(base) + size_t(k) * G + jl] = ob[k < len ? k : 0];
                for (int k = 0; k < len; ++k) {
                    const int slot = int(std::lower_bound(rg.cams.begin(), rg.cams.end(), int(st_cam[ob[k]])) - rg.cams.begin());
                    fwd[size_t(k) * G + jl] = (unsigned char)slot;
                    inv[size_t(slot) * G + jl] = (unsigned char)k;
                }
            }
        }
    });
    parallel_chunks(n_groups, 8, [&](int, size_t g0, size_t g1) {
        for (size_t g = g0; g < g1; ++g) {
            const uint32_t x = g_first[g], G = uint32_t(g_G_h[g]);
            const int L = g_L_h[g];
            const uint32_t a0 = sorted_a[x];
            const uint32_t* ob0 = lm_obs(a0);
            int fr[kGroupLmax];
            for (int i = 0; i < L; ++i) {
                const int c = int(st_cam[ob0[i]]);
                g_cams_h[size_t(g_off_h[g]) + i] = c;
                fr[i] = cam_free_h[c];
            }
            int t = g_blk_off_h[g];
            for (int i = 0; i < L; ++i)
                for (int k = i; k < L; ++k) {
                    int e = -1;
                    if (fr[i] >= 0 && fr[k] >= 0) {
                        auto b0 = s_col_h.begin() + s_rowptr_h[fr[i]], b1 = s_col_h.begin() + s_rowptr_h[fr[i] + 1];
                        e = int(std::lower_bound(b0, b1, fr[k]) - s_col_h.begin());
                    }
                    g_blk_h[size_t(t++)] = e;
                }
            const uint32_t base = g_obs0_h[g];
            for (uint32_t jl = 0; jl < G; ++jl) {
                const uint32_t a = sorted_a[x + jl];
                const size_t li = size_t(g_lm0_h[g]) + jl;
                grouped[a - lo] = 1;
                lm_user_h[li] = all_lm[a];
                lm_base_h[li] = base + jl;
                lm_stride_h[li] = G;
                lm_cnt_h[li] = uint32_t(L);
                const uint32_t* ob = lm_obs(a);
                for (int i = 0; i < L; ++i) obs_user_h[size_t(base) + size_t(i) * G + jl] = ob[i];
            }
        }
    });
    bt.lap("  grouping");
    // the remaining landmarks, landmark-major, in first-camera order
    {
        std::vector<uint32_t> rest;
        for (uint32_t x = 0; x < nl; ++x)
            if (!grouped[x]) rest.push_back(lo + x);
        std::vector<uint32_t> roff(rest.size() + 1, obs_cursor);
        for (size_t r = 0; r < rest.size(); ++r) roff[r + 1] = roff[r] + lm_len(rest[r]);
        parallel_chunks(rest.size(), 1 << 12, [&](int, size_t r0, size_t r1) {
            for (size_t r = r0; r < r1; ++r) {
                const uint32_t a = rest[r], len = lm_len(a);
                const size_t li = size_t(n_lm_grouped) + r;
                lm_user_h[li] = all_lm[a];
                lm_base_h[li] = roff[r];
                lm_stride_h[li] = 1;
                lm_cnt_h[li] = len;
                const uint32_t* ob = lm_obs(a);
                for (uint32_t k = 0; k < len; ++k) obs_user_h[size_t(roff[r]) + k] = ob[k];
            }
        });
        // ---- wide-window slices (kernels.cu schur_wide_kernel): the landmarks no group took are mostly the long
        // tracks.  A run of them (they are in first-camera order) whose cameras all lie inside [c0, c0 + 32), c0 the
        // run's first camera, is eliminated as one S_win -= Z Z^T on the tensor cores instead of one RED per entry of
        // every camera pair.  Not eligible: a camera twice, more than 32 observations, a span beyond the window.
        wide_lo_h.clear(); wide_hi_h.clear(); wide_c0_h.clear();
        wide_flag_h.assign(rest.size(), 0);
        wide_obs0 = obs_cursor;
        n_wide_lm = 0;
        static const bool no_wide = std::getenv("CSLAM_NO_WIDE") != nullptr;   // A/B knob
        if (want_ragged && opt.schur_path != 1 && !lighting_in_solve() && !no_wide) {
            constexpr uint32_t kWin = 32;
            constexpr size_t kMinSlice = 8, kMaxSlice = 1 << 14;
            const size_t R = rest.size();
            std::vector<uint32_t> firstc(R), lastc(R);
            std::vector<uint8_t> elig(R);
            parallel_chunks(R, 1 << 12, [&](int, size_t r0, size_t r1) {
                for (size_t r = r0; r < r1; ++r) {
                    const uint32_t a = rest[r], len = lm_len(a);
                    const uint32_t* ob = lm_obs(a);
                    bool ok = len >= 1 && len <= kWin;
                    for (uint32_t k = 1; k < len && ok; ++k) ok = st_cam[ob[k]] > st_cam[ob[k - 1]];
                    firstc[r] = len ? st_cam[ob[0]] : 0;
                    lastc[r] = len ? st_cam[ob[len - 1]] : 0;
                    elig[r] = ok && lastc[r] - firstc[r] < kWin;
                }
            });
            for (size_t i = 0; i < R;) {
                if (!elig[i]) {
                    ++i;
                    continue;
                }
                const uint32_t c0 = firstc[i];
                size_t j = i;
                while (j < R && elig[j] && firstc[j] >= c0 && lastc[j] < c0 + kWin && j - i < kMaxSlice) ++j;
                if (j - i >= kMinSlice) {
                    wide_lo_h.push_back(n_lm_grouped + int(i));
                    wide_hi_h.push_back(n_lm_grouped + int(j));
                    wide_c0_h.push_back(int(c0));
                    for (size_t r = i; r < j; ++r) wide_flag_h[r] = 1;
                    n_wide_lm += (long long)(j - i);
                }
                i = j > i ? j : i + 1;
            }
        }
    }
    bt.lap("  remaining landmarks");
}

// Pageable host memory -> device.  cudaMemcpy from pageable memory goes through the driver's own
// staging and tops out near 10 GB/s however many threads issue it (measured); here a few host threads
// copy 4 MB pieces into their own pinned buffers (allocated once per process, double-buffered) and
// DMA them from there, which is bound by the host memcpy and PCIe instead.
namespace {
struct H2DPool {
    static constexpr int kThreads = 8;
    static constexpr size_t kPiece = size_t(4) << 20;
    std::mutex mu;  // one transfer at a time
    int device = -1;
    char* buf[kThreads][2] = {};
    cudaStream_t st[kThreads] = {};
    cudaEvent_t ev[kThreads][2] = {};
    bool ensure(int dev) {
        if (device == dev) return true;
        if (device != -1) return false;  // the pool belongs to another device of this process
        for (int t = 0; t < kThreads; ++t) {
            if (cudaStreamCreateWithFlags(&st[t], cudaStreamNonBlocking) != cudaSuccess) return false;
            for (int b = 0; b < 2; ++b) {
                if (cudaHostAlloc(reinterpret_cast<void**>(&buf[t][b]), kPiece, cudaHostAllocDefault) != cudaSuccess) return false;
                if (cudaEventCreateWithFlags(&ev[t][b], cudaEventDisableTiming) != cudaSuccess) return false;
            }
        }
        device = dev;
        return true;
    }
};
H2DPool g_h2d;
// staging threads per transfer: all of them for a single process, fewer when several ranks of one job share
// the host (8 ranks x 8 threads only fight over the cores and the memory bus)
std::atomic<int> g_stage_threads{H2DPool::kThreads};
}  // namespace

static void parallel_h2d(int device, void* dst, const void* src, size_t bytes) {
    if (!bytes) return;
    std::lock_guard<std::mutex> lock(g_h2d.mu);
    CSLAM_CUDA(cudaSetDevice(device));
    if (bytes < (size_t(1) << 20) || !g_h2d.ensure(device)) {
        CSLAM_CUDA(cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice));
        return;
    }
    const size_t n_piece = (bytes + H2DPool::kPiece - 1) / H2DPool::kPiece;
    const int nt = int(std::min<size_t>(size_t(g_stage_threads.load()), n_piece));
    std::string err;
    std::mutex mu;
    auto work = [&](int t) {
        cudaError_t e = cudaSetDevice(device);
        int b = 0;
        for (size_t pc = size_t(t); pc < n_piece && e == cudaSuccess; pc += size_t(nt), b ^= 1) {
            const size_t off = pc * H2DPool::kPiece, n = std::min(H2DPool::kPiece, bytes - off);
            e = cudaEventSynchronize(g_h2d.ev[t][b]);  // the DMA that last read this buffer
            if (e != cudaSuccess) break;
            std::memcpy(g_h2d.buf[t][b], static_cast<const char*>(src) + off, n);
            e = cudaMemcpyAsync(static_cast<char*>(dst) + off, g_h2d.buf[t][b], n, cudaMemcpyHostToDevice, g_h2d.st[t]);
            if (e == cudaSuccess) e = cudaEventRecord(g_h2d.ev[t][b], g_h2d.st[t]);
        }
        if (e == cudaSuccess) e = cudaStreamSynchronize(g_h2d.st[t]);
        if (e != cudaSuccess) {
            std::lock_guard<std::mutex> g(mu);
            err = cudaGetErrorString(e);
        }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < nt; ++t) th.emplace_back(work, t);
    work(0);
    for (auto& x : th) x.join();
    if (!err.empty()) throw CudaError("H2D copy failed: " + err);
}

// The way back: device -> pageable host memory through the same pinned pieces (DMA into a piece, then a
// host thread copies it out while the next piece is in flight).
static void parallel_d2h(int device, void* dst, const void* src, size_t bytes) {
    if (!bytes) return;
    std::lock_guard<std::mutex> lock(g_h2d.mu);
    CSLAM_CUDA(cudaSetDevice(device));
    if (bytes < (size_t(1) << 20) || !g_h2d.ensure(device)) {
        CSLAM_CUDA(cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost));
        return;
    }
    const size_t n_piece = (bytes + H2DPool::kPiece - 1) / H2DPool::kPiece;
    const int nt = int(std::min<size_t>(size_t(g_stage_threads.load()), n_piece));
    std::string err;
    std::mutex mu;
    auto work = [&](int t) {
        cudaError_t e = cudaSetDevice(device);
        int b = 0;
        size_t prev_off = 0, prev_n = 0;
        bool have_prev = false;
        for (size_t pc = size_t(t); pc < n_piece && e == cudaSuccess; pc += size_t(nt), b ^= 1) {
            const size_t off = pc * H2DPool::kPiece, n = std::min(H2DPool::kPiece, bytes - off);
            e = cudaMemcpyAsync(g_h2d.buf[t][b], static_cast<const char*>(src) + off, n, cudaMemcpyDeviceToHost, g_h2d.st[t]);
            if (e == cudaSuccess) e = cudaEventRecord(g_h2d.ev[t][b], g_h2d.st[t]);
            if (have_prev && e == cudaSuccess) {
                e = cudaEventSynchronize(g_h2d.ev[t][b ^ 1]);
                if (e == cudaSuccess) std::memcpy(static_cast<char*>(dst) + prev_off, g_h2d.buf[t][b ^ 1], prev_n);
            }
            prev_off = off;
            prev_n = n;
            have_prev = true;
        }
        if (have_prev && e == cudaSuccess) {
            e = cudaEventSynchronize(g_h2d.ev[t][b ^ 1]);
            if (e == cudaSuccess) std::memcpy(static_cast<char*>(dst) + prev_off, g_h2d.buf[t][b ^ 1], prev_n);
        }
        if (e != cudaSuccess) {
            std::lock_guard<std::mutex> g(mu);
            err = cudaGetErrorString(e);
        }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < nt; ++t) th.emplace_back(work, t);
    work(0);
    for (auto& x : th) x.join();
    if (!err.empty()) throw CudaError("D2H copy failed: " + err);
}

// Structure analysis with the O(n_obs) passes on the device (structure.cu); the host part below is the
// O(n_landmarks) logic of build_structure(), unchanged, fed from per-landmark arrays instead of the
// observation lists.  Same layout as the host analysis (CSLAM_VERIFY_STRUCTURE=1 checks the hashes).
bool Engine::build_structure_gpu(cudaEvent_t) {
    if (!h_poses || n_poses == 0) throw std::invalid_argument("poses not set");
    if (n_st > 0 && (!h_points || n_points == 0)) throw std::invalid_argument("points not set");
    if (n_st >= (1ull << 32)) throw std::invalid_argument("more than 2^32 stereo blocks");
    PhaseTimer bt;
    const int max_threads = 32;
    const int kFarCap = 1 << 22;
    DBuf<uint32_t> d_cnt, d_ptr, d_fill, d_mincam;
    DBuf<uint8_t> d_used, d_kok, d_tmp;
    DBuf<unsigned long long> d_ck, d_khash, d_mask;
    DBuf<int> d_flags;
    DBuf<int2> d_far;
    auto free_all = [&]() {
        d_cnt.release_async(stream); d_ptr.release_async(stream); d_fill.release_async(stream); d_mincam.release_async(stream);
        d_used.release_async(stream); d_kok.release_async(stream); d_tmp.release_async(stream); d_ck.release_async(stream);
        d_khash.release_async(stream); d_mask.release_async(stream); d_flags.release_async(stream); d_far.release_async(stream);
    };
    d_cnt.alloc(size_t(n_points) + 1, stream);
    d_cnt.zero(stream);
    d_used.alloc(n_poses, stream);
    d_used.zero(stream);
    d_flags.alloc(4, stream);
    d_flags.zero(stream);
    launch_st_count(stream, n_st, d_raw_cam.p, d_raw_pt.p, n_poses, n_points, d_cnt.p, d_used.p, d_flags.p);
    d_fill.alloc(std::max<size_t>(n_points, 1), stream);
    launch_st_max_track(stream, n_points, d_cnt.p, d_fill.p, d_tmp);  // d_fill[0] <- longest track (buffer reused below)
    std::vector<uint8_t> used(n_poses, 0);
    int flags[4] = {0, 0, 0, 0};
    uint32_t max_track = 0;
    CSLAM_CUDA(cudaMemcpyAsync(&max_track, d_fill.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
    CSLAM_CUDA(cudaMemcpyAsync(used.data(), d_used.p, n_poses, cudaMemcpyDeviceToHost, stream));
    CSLAM_CUDA(cudaMemcpyAsync(flags, d_flags.p, sizeof(flags), cudaMemcpyDeviceToHost, stream));
    CSLAM_CUDA(cudaStreamSynchronize(stream));
    if (flags[0]) {
        free_all();
        throw std::invalid_argument("stereo block index out of range");
    }
    const uint32_t max_track_gpu = [] {  // (the knob exists for the test of this hand-over)
        const char* e = std::getenv("CSLAM_GPU_STRUCTURE_MAX_TRACK");
        return e ? uint32_t(std::atoll(e)) : 4096u;
    }();
    if (max_track > max_track_gpu) {
        free_all();
        return false;  // one thread per point sorts its list and walks its camera pairs: host analysis
    }
    for (auto& s : suns) {
        if (s.cam >= n_poses) throw std::invalid_argument("sun block index out of range");
        used[s.cam] = 1;
    }
    for (auto& p : priors) {
        if (p.cam >= n_poses) throw std::invalid_argument("prior block index out of range");
        used[p.cam] = 1;
    }
    cam_free_h.assign(n_poses, -1);
    free_cams_h.clear();
    for (uint32_t k = 0; k < n_poses; ++k)
        if (used[k] && !pose_const[k]) {
            cam_free_h[k] = int(free_cams_h.size());
            free_cams_h.push_back(int(k));
        }
    n_free = int(free_cams_h.size());
    d_cam_free.upload(cam_free_h, stream);
    bt.lap("  [gpu] count / free cams");

    // observation lists per point, sorted by (camera, block index); per-landmark facts; S pattern masks
    d_ptr.alloc(size_t(n_points) + 1, stream);
    launch_st_scan(stream, n_points, d_cnt.p, d_ptr.p, d_tmp);
    d_fill.alloc(std::max<size_t>(n_points, 1), stream);
    CSLAM_CUDA(cudaMemcpyAsync(d_fill.p, d_ptr.p, size_t(n_points) * sizeof(uint32_t), cudaMemcpyDeviceToDevice, stream));
    d_ck.alloc(std::max<size_t>(n_st, 1), stream);
    launch_st_fill(stream, n_st, d_raw_cam.p, d_raw_pt.p, d_fill.p, d_ck.p);
    d_mincam.alloc(std::max<size_t>(n_points, 1), stream);
    d_khash.alloc(std::max<size_t>(n_points, 1), stream);
    d_kok.alloc(std::max<size_t>(n_points, 1), stream);
    d_mask.alloc(size_t(std::max(n_free, 1)), stream);
    d_mask.zero(stream);
    d_far.alloc(kFarCap, stream);
    launch_st_landmarks(stream, n_points, d_cnt.p, d_ptr.p, d_ck.p, d_cam_free.p, kGroupLmax, opt.schur_path != 1, d_mincam.p,
                        d_khash.p, d_kok.p, d_mask.p, d_far.p, kFarCap, d_flags.p);
    if (n_poses >= (1u << 26)) {
        free_all();
        return false;  // the grouping key packs the first camera into 26 bits
    }
    // landmarks in order of the first camera that sees them (stable in the point index), on the device
    const size_t np1 = size_t(n_points) + 1;
    DBuf<uint32_t> d_keyt, d_valt, d_mincam_s, d_all_lm, d_counts, d_len_a, d_all_ptr, d_key2, d_key2b, d_key2s, d_val_a, d_val_b,
        d_sorted_a, d_run_pos, d_run_L, d_rflag, d_rlen, d_ridx, d_roff, d_gx, d_gobs0;
    DBuf<unsigned long long> d_keyh, d_keyh2;
    DBuf<uint8_t> d_run_flag, d_grouped;
    DBuf<int> d_goff, d_gL, d_gG, d_glm0, d_gcams;
    auto free_more = [&]() {
        for (DBuf<uint32_t>* d : {&d_keyt, &d_valt, &d_mincam_s, &d_all_lm, &d_counts, &d_len_a, &d_all_ptr, &d_key2, &d_key2b, &d_key2s,
                                  &d_val_a, &d_val_b, &d_sorted_a, &d_run_pos, &d_run_L, &d_rflag, &d_rlen, &d_ridx, &d_roff, &d_gx, &d_gobs0})
            d->release_async(stream);
        d_keyh.release_async(stream); d_keyh2.release_async(stream); d_run_flag.release_async(stream); d_grouped.release_async(stream);
        for (DBuf<int>* d : {&d_goff, &d_gL, &d_gG, &d_glm0, &d_gcams}) d->release_async(stream);
        free_all();
    };
    for (DBuf<uint32_t>* d : {&d_keyt, &d_valt, &d_mincam_s, &d_all_lm, &d_key2, &d_key2b, &d_key2s, &d_val_a, &d_val_b, &d_sorted_a,
                              &d_run_pos, &d_rflag, &d_rlen, &d_ridx, &d_roff})
        d->alloc(std::max<size_t>(n_points, 1), stream);
    d_len_a.alloc(np1, stream);
    d_all_ptr.alloc(np1, stream);
    d_counts.alloc(4, stream);
    d_keyh.alloc(std::max<size_t>(n_points, 1), stream);
    d_keyh2.alloc(std::max<size_t>(n_points, 1), stream);
    d_run_flag.alloc(std::max<size_t>(n_points, 1), stream);
    d_grouped.alloc(std::max<size_t>(n_points, 1), stream);
    launch_st_order(stream, n_points, n_poses, d_cnt.p, d_mincam.p, d_keyt.p, d_valt.p, d_mincam_s.p, d_all_lm.p, d_counts.p + 2, d_tmp);
    std::vector<unsigned long long> mask(size_t(std::max(n_free, 1)));
    uint32_t counts[4] = {0, 0, 0, 0};  // n_ok, n_runs, n_active
    CSLAM_CUDA(cudaMemcpyAsync(mask.data(), d_mask.p, mask.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream));
    CSLAM_CUDA(cudaMemcpyAsync(flags, d_flags.p, sizeof(flags), cudaMemcpyDeviceToHost, stream));
    CSLAM_CUDA(cudaMemcpyAsync(counts + 2, d_counts.p + 2, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
    CSLAM_CUDA(cudaStreamSynchronize(stream));
    if (flags[2]) {
        free_more();
        return false;  // more far pairs than the list holds: host analysis
    }
    std::vector<std::pair<int, int>> far;
    far.resize(size_t(flags[1]));
    if (!far.empty()) {
        static_assert(sizeof(std::pair<int, int>) == sizeof(int2), "pair layout");
        CSLAM_CUDA(cudaMemcpyAsync(far.data(), d_far.p, far.size() * sizeof(int2), cudaMemcpyDeviceToHost, stream));
        CSLAM_CUDA(cudaStreamSynchronize(stream));
    }
    const uint32_t n_active = counts[2];
    bt.lap("  [gpu] lists / facts / order");

    // reduced camera system pattern
    {
        std::sort(far.begin(), far.end());
        far.erase(std::unique(far.begin(), far.end()), far.end());
        s_rowptr_h.assign(size_t(n_free) + 1, 0);
        s_col_h.clear();
        size_t fx = 0;
        for (int a = 0; a < n_free; ++a) {
            s_col_h.push_back(a);
            for (int d = 1; d < 64; ++d)
                if (mask[a] >> d & 1) s_col_h.push_back(a + d);
            for (; fx < far.size() && far[fx].first == a; ++fx) s_col_h.push_back(far[fx].second);
            s_rowptr_h[a + 1] = int(s_col_h.size());
        }
        nnzU = int(s_col_h.size());
    }

    // this rank's shard: contiguous landmark range balanced by observation count (needs the prefix sums
    // of the track lengths on the host; a single rank takes everything)
    uint32_t lo = 0, hi = n_active;
    std::vector<uint32_t> all_ptr;
    if (n_ranks > 1) {
        // track lengths and their prefix sums come out of the grouping pass; run it once over everything to get them
        launch_st_group_sort(stream, n_points, 0, n_active, d_all_lm.p, d_mincam_s.p, d_cnt.p, d_kok.p, d_khash.p, d_len_a.p,
                             d_all_ptr.p, d_key2.p, d_keyh.p, d_keyh2.p, d_val_a.p, d_val_b.p, d_key2b.p, d_key2s.p, d_sorted_a.p,
                             d_run_flag.p, d_run_pos.p, d_counts.p, d_tmp);
        all_ptr.resize(size_t(n_active) + 1);
        CSLAM_CUDA(cudaMemcpyAsync(all_ptr.data(), d_all_ptr.p, all_ptr.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
        CSLAM_CUDA(cudaStreamSynchronize(stream));
        auto cut = [&](int r) -> uint32_t {
            const uint64_t target = n_st * uint64_t(r) / uint64_t(n_ranks);
            return uint32_t(std::lower_bound(all_ptr.begin(), all_ptr.end(), uint32_t(target)) - all_ptr.begin());
        };
        lo = std::min(cut(rank), n_active);
        hi = rank == n_ranks - 1 ? n_active : std::min(cut(rank + 1), n_active);
    }
    // groups: runs of equal (first camera, track length, camera-list hash) among the eligible landmarks of
    // the shard, sorted on the device; the lists themselves are compared afterwards (st_verify_kernel)
    launch_st_group_sort(stream, n_points, lo, hi, d_all_lm.p, d_mincam_s.p, d_cnt.p, d_kok.p, d_khash.p, d_len_a.p, d_all_ptr.p,
                         d_key2.p, d_keyh.p, d_keyh2.p, d_val_a.p, d_val_b.p, d_key2b.p, d_key2s.p, d_sorted_a.p, d_run_flag.p,
                         d_run_pos.p, d_counts.p, d_tmp);
    uint32_t ptr_lo_hi[2] = {0, 0};
    CSLAM_CUDA(cudaMemcpyAsync(counts, d_counts.p, 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
    CSLAM_CUDA(cudaMemcpyAsync(&ptr_lo_hi[0], d_all_ptr.p + lo, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
    CSLAM_CUDA(cudaMemcpyAsync(&ptr_lo_hi[1], d_all_ptr.p + hi, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
    CSLAM_CUDA(cudaStreamSynchronize(stream));
    const uint32_t n_ok = counts[0], n_runs = counts[1];
    n_lm = int(hi - lo);
    lm_lo = 0;
    lm_hi = n_lm;
    n_obs = (long long)ptr_lo_hi[1] - (long long)ptr_lo_hi[0];
    const uint32_t nl = hi - lo;
    std::vector<uint32_t> run_pos(n_runs), run_L(n_runs);
    if (n_runs) {
        d_run_L.alloc(n_runs, stream);
        launch_st_run_len(stream, n_runs, d_run_pos.p, d_key2s.p, d_run_L.p);
        CSLAM_CUDA(cudaMemcpyAsync(run_pos.data(), d_run_pos.p, n_runs * sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
        CSLAM_CUDA(cudaMemcpyAsync(run_L.data(), d_run_L.p, n_runs * sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
        CSLAM_CUDA(cudaStreamSynchronize(stream));
    }
    bt.lap("  [gpu] grouping sort");

    const size_t min_group = opt.schur_path == 2 ? 1 : 4;
    g_L_h.clear(); g_G_h.clear(); g_lm0_h.clear(); g_obs0_h.clear(); g_off_h.clear(); g_cams_h.clear();
    g_blk_off_h.clear(); g_blk_h.clear(); item_group_h.clear(); item_j0_h.clear(); item_n_h.clear();
    std::vector<uint32_t> g_first;  // position of each group's first landmark in the sorted order
    uint32_t obs_cursor = 0, lm_cursor = 0;
    int cams_cursor = 0, blk_cursor = 0;
    std::vector<std::pair<int, int>> items_small, items_large;
    for (uint32_t r = 0; r < n_runs; ++r) {
        const uint32_t x = run_pos[r];
        const uint32_t G = (r + 1 < n_runs ? run_pos[r + 1] : n_ok) - x;
        if (G >= min_group) {
            const int L = int(run_L[r]);
            const int gid = int(g_L_h.size());
            g_L_h.push_back(L);
            g_G_h.push_back(int(G));
            g_lm0_h.push_back(int(lm_cursor));
            g_obs0_h.push_back(obs_cursor);
            g_off_h.push_back(cams_cursor);
            g_blk_off_h.push_back(blk_cursor);
            g_first.push_back(x);
            cams_cursor += L;
            blk_cursor += L * (L + 1) / 2;
            lm_cursor += G;
            obs_cursor += G * uint32_t(L);
            for (uint32_t j0 = 0; j0 < G; j0 += kItemMax) (L <= 10 ? items_small : items_large).push_back({gid, int(j0)});
            max_group_L = std::max(max_group_L, L);
        }
    }
    const size_t n_groups = g_L_h.size();
    n_lm_grouped = int(lm_cursor);
    n_items_small = int(items_small.size());
    for (auto* lst : {&items_small, &items_large})
        for (auto& it : *lst) {
            item_group_h.push_back(it.first);
            item_j0_h.push_back(it.second);
            item_n_h.push_back(std::min(kItemMax, g_G_h[it.first] - it.second));
        }
    bt.lap("    runs -> groups (host)");
    g_map_off_h.assign(g_L_h.size(), -1);   // the device analysis forms exact groups only
    g_map_h.assign(1, 0xff);
    n_items_rag = 0;
    n_obs_true = n_obs;
    // layout rows of every landmark and the groups' camera lists, on the device
    g_cams_h.assign(size_t(cams_cursor), 0);
    d_lm_user.alloc(std::max<size_t>(nl, 1), stream);
    d_lm_base.alloc(std::max<size_t>(nl, 1), stream);
    d_lm_stride.alloc(std::max<size_t>(nl, 1), stream);
    d_lm_cnt.alloc(std::max<size_t>(nl, 1), stream);
    if (n_groups) {
        d_gx.upload(g_first, stream);
        d_gobs0.upload(g_obs0_h, stream);
        d_goff.upload(g_off_h, stream);
        d_gL.upload(g_L_h, stream);
        d_gG.upload(g_G_h, stream);
        d_glm0.upload(g_lm0_h, stream);
    }
    d_gcams.alloc(std::max<size_t>(size_t(cams_cursor), 1), stream);
    launch_st_layout(stream, n_points, lo, hi, int(n_groups), d_gx.p, d_gG.p, d_gL.p, d_glm0.p, d_gobs0.p, d_goff.p, d_sorted_a.p,
                     d_all_lm.p, d_len_a.p, d_ptr.p, d_ck.p, uint32_t(n_lm_grouped), obs_cursor, d_lm_user.p, d_lm_base.p,
                     d_lm_stride.p, d_lm_cnt.p, d_grouped.p, d_gcams.p, d_rflag.p, d_rlen.p, d_ridx.p, d_roff.p, d_tmp);
    if (n_groups) {
        CSLAM_CUDA(cudaMemcpyAsync(g_cams_h.data(), d_gcams.p, g_cams_h.size() * sizeof(int), cudaMemcpyDeviceToHost, stream));
        CSLAM_CUDA(cudaStreamSynchronize(stream));
    }
    bt.lap("    layout kernels + camera lists back");
    // pair -> block tables of the groups (host: binary searches in the pattern)
    g_blk_h.assign(size_t(blk_cursor), -1);
    parallel_chunks(n_groups, 8, [&](int, size_t g0, size_t g1) {
        for (size_t g = g0; g < g1; ++g) {
            const int L = g_L_h[g];
            int fr[kGroupLmax];
            for (int i = 0; i < L; ++i) fr[i] = cam_free_h[g_cams_h[size_t(g_off_h[g]) + i]];
            int t = g_blk_off_h[g];
            for (int i = 0; i < L; ++i)
                for (int k = i; k < L; ++k) {
                    int e = -1;
                    if (fr[i] >= 0 && fr[k] >= 0) {
                        auto b0 = s_col_h.begin() + s_rowptr_h[fr[i]], b1 = s_col_h.begin() + s_rowptr_h[fr[i] + 1];
                        e = int(std::lower_bound(b0, b1, fr[k]) - s_col_h.begin());
                    }
                    g_blk_h[size_t(t++)] = e;
                }
        }
    });
    bt.lap("  [gpu] groups / layout tables");

    // the observation permutation and the check of the groups; host copies of the layout rows only on request
    obs_user_n = size_t(n_obs);
    obs_user_h.reset();
    lm_user_h.clear(); lm_base_h.clear(); lm_stride_h.clear(); lm_cnt_h.clear();
    d_obs_user.alloc(std::max<size_t>(obs_user_n, 1), stream);
    launch_st_perm(stream, n_lm, d_lm_user.p, d_lm_base.p, d_lm_stride.p, d_lm_cnt.p, d_ptr.p, d_ck.p, d_obs_user.p);
    if (n_groups)
        launch_st_verify(stream, int(n_groups), d_gL.p, d_gG.p, d_glm0.p, d_goff.p, d_gcams.p, d_lm_user.p, d_ptr.p, d_ck.p, d_flags.p);
    CSLAM_CUDA(cudaMemcpyAsync(flags, d_flags.p, sizeof(flags), cudaMemcpyDeviceToHost, stream));
    if (std::getenv("CSLAM_VERIFY_STRUCTURE")) {
        lm_user_h.resize(nl); lm_base_h.resize(nl); lm_stride_h.resize(nl); lm_cnt_h.resize(nl);
        if (nl) {
            CSLAM_CUDA(cudaMemcpyAsync(lm_user_h.data(), d_lm_user.p, nl * sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
            CSLAM_CUDA(cudaMemcpyAsync(lm_base_h.data(), d_lm_base.p, nl * sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
            CSLAM_CUDA(cudaMemcpyAsync(lm_stride_h.data(), d_lm_stride.p, nl * sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
            CSLAM_CUDA(cudaMemcpyAsync(lm_cnt_h.data(), d_lm_cnt.p, nl * sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
        }
    }
    CSLAM_CUDA(cudaStreamSynchronize(stream));
    free_more();
    bt.lap("  [gpu] permutation + group check");
    if (flags[3]) return false;  // two camera lists with one 64-bit hash: host analysis
    structure_on_device = true;
    return true;
}

GroupView Engine::group_view() const {
    GroupView g;
    g.n_items = int(item_group_h.size());
    g.item_group = d_item_group.p;
    g.item_j0 = d_item_j0.p;
    g.item_n = d_item_n.p;
    g.g_L = d_g_L.p;
    g.g_G = d_g_G.p;
    g.g_lm0 = d_g_lm0.p;
    g.g_obs0 = d_g_obs0.p;
    g.g_off = d_g_off.p;
    g.g_cams = d_g_cams.p;
    g.g_blk_off = d_g_blk_off.p;
    g.g_blk = d_g_blk.p;
    g.g_map_off = d_g_map_off.p;
    g.g_map = d_g_map.p;
    g.g_long = bandpc_active ? d_g_long.p : nullptr;
    g.S_long = bandpc_active ? d_S2.p : nullptr;
    g.Bdiag_long = bandpc_active ? d_Bdiag2.p : nullptr;
    return g;
}

void Engine::launch_schur(const DevView& v, const LmDiag& dg) {
    if (!item_group_h.empty())
        launch_schur_grouped(stream, v, group_view(), n_items_small, n_items_rag, dg, d_S, d_Bdiag, d_bp, d_gp, d_gl.p, d_scal);
    // (the landmarks no group takes are mostly the long tracks: beyond the banded preconditioner's window)
    if (n_lm > n_lm_grouped) {
        double* St = bandpc_active ? d_S2.p : d_S;
        double* Bt = bandpc_active ? d_Bdiag2.p : d_Bdiag;
        const bool wide = !wide_lo_h.empty();
        if (wide)
            launch_schur_wide(stream, v, n_lm_grouped, n_lm, d_wide_flag.p, int(wide_lo_h.size()), d_wide_lo.p, d_wide_hi.p,
                              d_wide_c0.p, wide_obs0, d_wide_Z.p, dg, St, Bt, d_bp, d_gp, d_gl.p, d_scal);
        if (!wide || n_wide_lm < (long long)(n_lm - n_lm_grouped))
            launch_schur_generic(stream, v, n_lm_grouped, n_lm, wide ? d_wide_flag.p : nullptr, dg, St, Bt, d_bp, d_gp, d_gl.p, d_scal);
    }
}

void Engine::upload() {
    PhaseTimer pt;
    const auto t_upload = std::chrono::steady_clock::now();
    ensure_device(this, &stream, &own_stream, &ev_a, &ev_b, &ev_c, &ev_d, &h_pinned);
    pt.lap("ensure_device");
    // Large problems: the index arrays go up first (several host threads), the structure is analysed on
    // the device while the measurements follow.  Otherwise (and for the lighting solve, whose setup reads
    // the host-side permutation) the host analyses the structure while a second thread uploads.
    structure_on_device = false;
    g_stage_threads.store(std::max(2, H2DPool::kThreads / std::max(1, n_ranks)));
    const size_t gpu_min = [] {  // CSLAM_HOST_STRUCTURE=1 keeps the host analysis; ..._MIN moves the size threshold
        if (std::getenv("CSLAM_HOST_STRUCTURE")) return ~size_t(0);
        const char* e = std::getenv("CSLAM_GPU_STRUCTURE_MIN");
        return e ? size_t(std::atoll(e)) : (size_t(1) << 20);
    }();
    const bool gpu_structure = n_st >= gpu_min && n_st > 0 && !lighting_in_solve();
    if (gpu_structure) {
        const int dev = opt.device;
        // Several ranks of one job share the host and its PCIe root: every rank uploading the whole problem
        // multiplies the host-side traffic by the rank count (e2e at 8 GPUs was half of 1 GPU).  Instead rank r
        // uploads the r-th slice of every array and the slices are exchanged over NVLink (in-place all-gather):
        // chunk sizes are rounded up, so the device buffers carry a little padding.
        const size_t R = size_t(std::max(1, n_ranks));
        auto chunk_of = [&](size_t bytes) { return ((bytes + R - 1) / R + 255) & ~size_t(255); };
        auto padded = [&](size_t bytes, size_t elem) { return R > 1 ? (chunk_of(bytes) * R + elem - 1) / elem : bytes / elem; };
        auto h2d_slice = [&](void* dst, const void* src, size_t bytes) {
            if (R == 1) {
                parallel_h2d(dev, dst, src, bytes);
                return;
            }
            const size_t ch = chunk_of(bytes), lo = std::min(bytes, size_t(rank) * ch), hi = std::min(bytes, lo + ch);
            if (hi > lo) parallel_h2d(dev, static_cast<char*>(dst) + lo, static_cast<const char*>(src) + lo, hi - lo);
        };
        auto gather = [&](void* buf, size_t bytes) {
            if (R > 1) comm_allgather_bytes(nccl_comm, buf, chunk_of(bytes), rank, stream);
        };
        const size_t w_count = st_W_per_obs ? 9 * n_st : 9;
        d_raw_cam.alloc(padded(n_st * sizeof(uint32_t), sizeof(uint32_t)), stream);
        d_raw_pt.alloc(padded(n_st * sizeof(uint32_t), sizeof(uint32_t)), stream);
        d_raw_uvd.alloc(padded(3 * n_st * sizeof(double), sizeof(double)), stream);
        d_raw_W.alloc(st_W_per_obs ? padded(w_count * sizeof(double), sizeof(double)) : 9, stream);
        d_raw_pts.alloc(padded(3 * std::max<size_t>(n_points, 1) * sizeof(double), sizeof(double)), stream);
        CSLAM_CUDA(cudaStreamSynchronize(stream));
        h2d_slice(d_raw_cam.p, st_cam, n_st * sizeof(uint32_t));
        h2d_slice(d_raw_pt.p, st_pt, n_st * sizeof(uint32_t));
        gather(d_raw_cam.p, n_st * sizeof(uint32_t));
        gather(d_raw_pt.p, n_st * sizeof(uint32_t));
        pt.lap("index H2D");
        std::exception_ptr rest_err;
        std::thread rest([&, dev]() {
            try {
                h2d_slice(d_raw_uvd.p, st_uvd, 3 * n_st * sizeof(double));
                if (st_W_per_obs)
                    h2d_slice(d_raw_W.p, st_W, w_count * sizeof(double));
                else
                    parallel_h2d(dev, d_raw_W.p, st_W, 9 * sizeof(double));
                if (n_points) h2d_slice(d_raw_pts.p, h_points, 3 * size_t(n_points) * sizeof(double));
                if (pt.on)
                    std::fprintf(stderr, "[cslam timing] %-28s %8.2f ms (second thread, from the start of upload)\n", "  raw H2D done",
                                 std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_upload).count());
            } catch (...) {
                rest_err = std::current_exception();
            }
        });
        bool ok = false;
        try {
            ok = build_structure_gpu(nullptr);
            // the device analysis only forms groups of identical camera lists; when more than 5 % of the landmarks
            // stay outside them (ragged tracks), the host analysis — which also forms ragged groups — takes over
            if (ok && !std::getenv("CSLAM_VERIFY_STRUCTURE") && (n_lm - n_lm_grouped) > n_lm / 20) {
                structure_on_device = false;
                ok = false;
            }
            if (!ok) build_structure();
        } catch (...) {
            rest.join();
            throw;
        }
        pt.lap(ok ? "build_structure_gpu" : "build_structure (host fallback)");
        rest.join();
        if (rest_err) std::rethrow_exception(rest_err);
        // the other ranks' slices of the measurements and points (NCCL is driven from this thread only)
        gather(d_raw_uvd.p, 3 * n_st * sizeof(double));
        if (st_W_per_obs) gather(d_raw_W.p, w_count * sizeof(double));
        if (n_points) gather(d_raw_pts.p, 3 * size_t(n_points) * sizeof(double));
        pt.lap("wait for raw H2D");
        if (ok && std::getenv("CSLAM_VERIFY_STRUCTURE")) {
            // debugging / tests: the host analysis must produce exactly the layout the device built
            obs_user_h.reset(new uint32_t[std::max<size_t>(obs_user_n, 1)]);
            CSLAM_CUDA(cudaMemcpyAsync(obs_user_h.get(), d_obs_user.p, obs_user_n * sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
            CSLAM_CUDA(cudaStreamSynchronize(stream));
            const unsigned long long h_dev = layout_hash();
            const std::vector<int> rp(s_rowptr_h), cl(s_col_h);
            want_ragged = false;   // (the device analysis forms no ragged groups)
            build_structure();
            want_ragged = true;
            if (layout_hash() != h_dev || rp != s_rowptr_h || cl != s_col_h)
                throw std::runtime_error("device and host structure analyses disagree");
        }
    } else {
        // The caller's arrays go to the device as they are, from a second host thread, while this
        // thread analyses the structure; the landmark-major SoA layout is then gathered on the GPU.
        cudaStream_t copy_stream = nullptr;
        CSLAM_CUDA(cudaStreamCreateWithFlags(&copy_stream, cudaStreamNonBlocking));
        d_raw_cam.alloc(std::max<size_t>(n_st, 1), stream);
        d_raw_uvd.alloc(3 * std::max<size_t>(n_st, 1), stream);
        d_raw_W.alloc(st_W_per_obs ? 9 * std::max<size_t>(n_st, 1) : 9, stream);
        d_raw_pts.alloc(3 * std::max<size_t>(n_points, 1), stream);
        CSLAM_CUDA(cudaStreamSynchronize(stream));  // the copy stream may touch them from here on
        std::string copy_err;
        const int dev = opt.device;
        std::thread copier([&, dev]() {
            auto chk = [&](cudaError_t e) {
                if (e != cudaSuccess && copy_err.empty()) copy_err = cudaGetErrorString(e);
            };
            chk(cudaSetDevice(dev));
            if (n_st) {
                chk(cudaMemcpyAsync(d_raw_cam.p, st_cam, n_st * sizeof(uint32_t), cudaMemcpyHostToDevice, copy_stream));
                chk(cudaMemcpyAsync(d_raw_uvd.p, st_uvd, 3 * n_st * sizeof(double), cudaMemcpyHostToDevice, copy_stream));
                chk(cudaMemcpyAsync(d_raw_W.p, st_W, (st_W_per_obs ? 9 * n_st : 9) * sizeof(double), cudaMemcpyHostToDevice,
                                    copy_stream));
            }
            if (n_points)
                chk(cudaMemcpyAsync(d_raw_pts.p, h_points, 3 * size_t(n_points) * sizeof(double), cudaMemcpyHostToDevice, copy_stream));
            chk(cudaStreamSynchronize(copy_stream));
            if (pt.on)
                std::fprintf(stderr, "[cslam timing] %-28s %8.2f ms (second thread, from the start of upload)\n", "  raw H2D done",
                             std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_upload).count());
        });
        try {
            build_structure();
        } catch (...) {
            copier.join();
            cudaStreamDestroy(copy_stream);
            throw;
        }
        pt.lap("build_structure");
        copier.join();
        cudaStreamDestroy(copy_stream);
        if (!copy_err.empty()) throw CudaError("raw H2D copy failed: " + copy_err);
        pt.lap("wait for raw H2D");
    }
    if (!structure_on_device) d_cam_free.upload(cam_free_h, stream);
    auto up_i = [&](DBuf<int>& d, const std::vector<int>& h) { d.upload(h.empty() ? std::vector<int>(1, 0) : h, stream); };
    auto up_u = [&](DBuf<uint32_t>& d, const std::vector<uint32_t>& h) {
        d.upload(h.empty() ? std::vector<uint32_t>(1, 0) : h, stream);
    };
    if (!structure_on_device) {
        up_u(d_lm_base, lm_base_h);
        up_u(d_lm_stride, lm_stride_h);
        up_u(d_lm_cnt, lm_cnt_h);
    }
    up_i(d_item_group, item_group_h);
    up_i(d_item_j0, item_j0_h);
    up_i(d_item_n, item_n_h);
    up_i(d_g_L, g_L_h);
    up_i(d_g_G, g_G_h);
    up_i(d_g_lm0, g_lm0_h);
    up_u(d_g_obs0, g_obs0_h);
    up_i(d_g_off, g_off_h);
    up_i(d_g_cams, g_cams_h);
    up_i(d_g_blk_off, g_blk_off_h);
    up_i(d_g_blk, g_blk_h);
    up_i(d_g_map_off, g_map_off_h);
    if (structure_on_device) {   // (the device analysis forms no wide-window slices)
        wide_lo_h.clear(); wide_hi_h.clear(); wide_c0_h.clear(); wide_flag_h.clear();
        n_wide_lm = 0;
    }
    if (!wide_lo_h.empty()) {
        up_i(d_wide_lo, wide_lo_h);
        up_i(d_wide_hi, wide_hi_h);
        up_i(d_wide_c0, wide_c0_h);
        d_wide_flag.upload(wide_flag_h, stream);
        d_wide_Z.alloc(18 * size_t(std::max<long long>(n_obs - wide_obs0, 1)), stream);
    }
    d_g_map.upload(g_map_h.empty() ? std::vector<unsigned char>(1, 0xff) : g_map_h, stream);
    // internal order <- caller's order, on the device
    const size_t no = size_t(std::max<long long>(n_obs, 1));
    if (!structure_on_device) {
        d_obs_user.alloc(std::max<size_t>(obs_user_n, 1), stream);
        CSLAM_CUDA(cudaStreamSynchronize(stream));
        // pageable H2D is bound by the host-side staging copy: split it over a few threads / streams
        std::string perr;
        const int dev = opt.device;
        parallel_chunks(obs_user_n, size_t(1) << 21, [&](int, size_t c0, size_t c1) {
            cudaStream_t cs = nullptr;
            cudaError_t e = cudaSetDevice(dev);
            if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking);
            if (e == cudaSuccess)
                e = cudaMemcpyAsync(d_obs_user.p + c0, obs_user_h.get() + c0, (c1 - c0) * sizeof(uint32_t), cudaMemcpyHostToDevice, cs);
            if (e == cudaSuccess) e = cudaStreamSynchronize(cs);
            if (cs) cudaStreamDestroy(cs);
            if (e != cudaSuccess) perr = cudaGetErrorString(e);
        });
        if (!perr.empty()) throw CudaError("layout H2D failed: " + perr);
        d_lm_user.upload(lm_user_h.empty() ? std::vector<uint32_t>(1, 0) : lm_user_h, stream);
    }
    pt.lap("  small arrays + perm H2D (enqueue)");
    d_obs_cam.alloc(no, stream);
    d_obs_u.alloc(no, stream);
    d_obs_v.alloc(no, stream);
    d_obs_d.alloc(no, stream);
    d_obs_W.alloc(st_W_per_obs ? 9 * no : 9, stream);
    d_points_init.alloc(3 * size_t(std::max(n_lm, 1)), stream);
    launch_gather_layout(stream, n_obs, d_obs_user.p, d_raw_cam.p, d_raw_uvd.p, d_raw_W.p, st_W_per_obs, d_obs_cam.p,
                         d_obs_u.p, d_obs_v.p, d_obs_d.p, d_obs_W.p, n_lm, d_lm_user.p, d_raw_pts.p, d_points_init.p);
    pt.lap("  obs alloc + gather launch");
    d_poses_init.upload(h_poses, 12 * size_t(n_poses), stream);
    d_poses.alloc(12 * size_t(n_poses), stream);
    d_poses_cand.alloc(12 * size_t(n_poses), stream);
    d_poses_best.alloc(12 * size_t(n_poses), stream);
    d_points.alloc(3 * size_t(n_lm), stream);
    d_points_cand.alloc(3 * size_t(n_lm), stream);
    d_points_best.alloc(3 * size_t(n_lm), stream);
    d_sc_p.alloc(6 * size_t(std::max(n_free, 1)), stream);
    d_sc_l.alloc(3 * size_t(std::max(n_lm, 1)), stream);
    d_cn_l.alloc(3 * size_t(std::max(n_lm, 1)), stream);
    d_gl.alloc(3 * size_t(std::max(n_lm, 1)), stream);
    d_yl.alloc(3 * size_t(std::max(n_lm, 1)), stream);
    d_s_rowptr.upload(s_rowptr_h, stream);
    d_s_col.upload(s_col_h.empty() ? std::vector<int>(1, 0) : s_col_h, stream);
    pt.lap("  state allocs");
    // mirrored lists for the symmetric SpMV: row b lists (a, block) for every stored upper block
    // (a, b) with a < b, i.e. the blocks that act on it transposed
    {
        std::vector<int> ep(size_t(n_free) + 1, 0);
        for (int a = 0; a < n_free; ++a)
            for (int e = s_rowptr_h[a]; e < s_rowptr_h[a + 1]; ++e)
                if (s_col_h[e] != a) ep[s_col_h[e] + 1]++;
        for (int a = 0; a < n_free; ++a) ep[a + 1] += ep[a];
        std::vector<int> ecb(2 * size_t(std::max(1, ep[n_free])));
        std::vector<int> fill(ep.begin(), ep.end() - 1);
        for (int a = 0; a < n_free; ++a)
            for (int e = s_rowptr_h[a]; e < s_rowptr_h[a + 1]; ++e) {
                const int b = s_col_h[e];
                if (b == a) continue;
                ecb[2 * size_t(fill[b])] = a;
                ecb[2 * size_t(fill[b]++) + 1] = e;
            }
        d_lt_rowptr.upload(ep, stream);
        d_lt_col.upload(ecb, stream);
    }
    pt.lap("  mirrored lists");
    // the joint lighting solve appends its dense border [S_cg | S_gg | b_g | g_g | h_g] to the same buffer
    ph.active = lighting_in_solve();
    size_t gext = 0;
    if (ph.active) {
        check_phong_solve();
        gext = size_t(ph.n_g) * 6 * size_t(n_free) + size_t(ph.n_g) * ph.n_g + 3 * size_t(ph.n_g);
    }
    red_count = 36 * size_t(nnzU) + 36 * size_t(n_free) + 6 * size_t(n_free) + 6 * size_t(n_free) + gext + SC_COUNT;
    d_red.alloc(red_count, stream);
    d_S = d_red.p;
    d_Bdiag = d_S + 36 * size_t(nnzU);
    d_bp = d_Bdiag + 36 * size_t(n_free);
    d_gp = d_bp + 6 * size_t(n_free);
    d_scal = d_gp + 6 * size_t(n_free);
    if (ph.active) {
        ph.Scg = d_scal;
        ph.Sgg = ph.Scg + size_t(ph.n_g) * 6 * size_t(n_free);
        ph.bg = ph.Sgg + size_t(ph.n_g) * ph.n_g;
        ph.gg = ph.bg + ph.n_g;
        ph.hg = ph.gg + ph.n_g;
        d_scal = ph.hg + ph.n_g;
    }
    d_Minv.alloc(36 * size_t(std::max(n_free, 1)), stream);
    d_diag_p.alloc(6 * size_t(std::max(n_free, 1)), stream);
    const size_t nv = 6 * size_t(std::max(n_free, 1));
    d_yp.alloc(nv, stream);
    d_pr.alloc(nv, stream);
    d_pz.alloc(nv, stream);
    d_pp.alloc(nv, stream);
    d_pq.alloc(nv, stream);
    d_pp2.alloc(nv, stream);
    d_prec.alloc(16, stream);
    d_pscal.alloc(PS_COUNT, stream);
    pt.lap("  reduced system allocs");
    plan_band_solver();
    plan_dense_solver();
    pt.lap("  plan_band_solver");
    d_scal2.alloc(SC_COUNT, stream);
    d_dmax_tmp.alloc(1, stream);
    if (!suns.empty()) d_suns.upload(suns, stream);
    if (!priors.empty()) d_priors.upload(priors, stream);
    if (ph.active) setup_phong_solve();
    CSLAM_CUDA(cudaStreamSynchronize(stream));
    pt.lap("  stream sync");
    d_raw_cam.release_async(stream);  // d_raw_pts stays: download() scatters the result into it
    d_raw_pt.release_async(stream);
    d_raw_uvd.release_async(stream);
    d_raw_W.release_async(stream);
    d_obs_user.release_async(stream);
    pt.lap("alloc + layout H2D + gather");
    uploaded = true;
    begun = false;
    user_copy_ready = false;
    reset_state();
}

// Decide whether the exact solve can use the banded direct solver, and lay out its buffers.
void Engine::plan_band_solver() {
    band_active = false;
    if (opt.linear_solver != 0 || n_free <= 0) return;
    int w = 0;
    for (int a = 0; a < n_free; ++a)
        if (s_rowptr_h[a + 1] > s_rowptr_h[a]) w = std::max(w, s_col_h[s_rowptr_h[a + 1] - 1] - a);
    bandpc_active = false;
    wband_active = false;
    if (w > kBandWmax) {
        // Not a narrow band.  A wide band (tracks of 14 .. 65 frames) on a problem that is long compared with it: the
        // chunked bordered-band factorisation (kernels_wband.cu) — exact, and far cheaper than the O(n^3) dense one.
        if (opt.dense_solver <= 0 && opt.bandpc_solver != 1 && opt.bandpc_solver >= 0 && !ph.active && plan_wband_solver(w)) return;
        // Too large for the dense factorisation as well: conjugate gradients preconditioned
        // with the banded solve of the short-track landmarks' part of the system (kernels_bandpcg.cu).
        if (6ll * n_free <= kDenseMaxN || opt.dense_solver > 0 || opt.bandpc_solver < 0 || n_free < 64 * kBandPcW || ph.active)
            return;
        bandpc_active = true;
        w = kBandPcW;
    }
    if (w == 0) w = 1;  // a single camera: treat as bandwidth 1 (absent blocks read as zero)
    const int n = n_free;
    const int W = band_storage_width(w);  // separators and the shared-memory window use this width
    // separator system: block cyclic reduction (log depth) once there are enough separators for it to
    // beat the sequential band factorisation
    band_sep = opt.band_separator_solver == 0 ? (n >= 400 * W ? 2 : 1) : opt.band_separator_solver;
    // leaves cost ~3 us per block row; the separator chain ~2.5 us per separator row when factored
    // as a band on one CTA, ~45 us per level of the cyclic reduction
    int P = opt.band_leaves > 0 ? opt.band_leaves
                                : (band_sep >= 2 ? 148 : int(std::lround(std::sqrt(0.8 * double(n) / W))));
    P = std::min(P, band_sep >= 2 ? 444 : 148);
    P = std::min(P, n / (4 * W));
    P = std::max(P, (n + 3499) / 3500);
    P = std::max(P, 1);
    int m = (n + P - 1) / P;
    P = (n + m - 1) / m;
    if (P > 1 && (m < 2 * W || n - (P - 1) * m < 1)) {
        bandpc_active = false;
        return;
    }
    band_w = w;
    band_P = P;
    band_m = m;
    std::vector<int> idx(size_t(n) * (w + 1), -1);
    if (bandpc_active) {
        // the preconditioner's matrix lives in dense band storage [n][w + 1][36] (rebuilt from S before every solve)
        for (int a = 0; a < n; ++a)
            for (int e = s_rowptr_h[a]; e < s_rowptr_h[a + 1]; ++e) {
                const int d = s_col_h[e] - a;
                if (d <= w) idx[size_t(a) * (w + 1) + d] = a * (w + 1) + d;
            }
        d_bpc_S.alloc(size_t(n) * (w + 1) * 36, stream);
        d_bpc_scal.alloc(4, stream);
        // which groups reach beyond the window (distance between their first and last FREE camera)
        std::vector<int> glong(std::max<size_t>(g_L_h.size(), 1), 0);
        for (size_t g = 0; g < g_L_h.size(); ++g) {
            int f0 = -1, f1 = -1;
            for (int i = 0; i < g_L_h[g]; ++i) {
                const int f = cam_free_h[g_cams_h[size_t(g_off_h[g]) + i]];
                if (f < 0) continue;
                if (f0 < 0) f0 = f;
                f1 = f;
            }
            glong[g] = (f0 >= 0 && f1 - f0 > w) ? 1 : 0;
        }
        d_g_long.upload(glong, stream);
        d_S2.alloc(36 * size_t(std::max(nnzU, 1)), stream);
        d_Bdiag2.alloc(36 * size_t(n), stream);
    } else {
        for (int a = 0; a < n; ++a)
            for (int e = s_rowptr_h[a]; e < s_rowptr_h[a + 1]; ++e) idx[size_t(a) * (w + 1) + (s_col_h[e] - a)] = e;
    }
    d_band_idx.upload(idx, stream);
    d_band_fail.alloc(1, stream);
    const size_t b = 6 * size_t(W), GC = P > 1 ? 1 + b : 1;
    d_Lbuf.alloc(size_t(n) * (W + 1) * 36, stream);
    d_Xbuf.alloc(size_t(n) * 6 * GC, stream);
    d_Ta.alloc(size_t(P) * b * b, stream);
    d_Ca.alloc(size_t(P) * b * b, stream);
    d_Tb.alloc(size_t(P) * b * b, stream);
    d_fa.alloc(size_t(P) * b, stream);
    d_fb.alloc(size_t(P) * b, stream);
    const size_t n2 = size_t(std::max(P - 1, 1)) * W, W2 = 2 * size_t(W) - 1;
    d_T2.alloc(n2 * (W2 + 1) * 36, stream);
    d_L2.alloc(n2 * (W2 + 1) * 36, stream);
    d_rhs2.alloc(6 * n2, stream);
    d_X2.alloc(6 * n2, stream);
    d_y2.alloc(6 * n2, stream);
    band_active = !bandpc_active;
}

// The exact solve of a wide block-banded reduced system (half-bandwidth w of 13 .. 64 blocks), kernels_wband.cu:
// C chunks separated by separators of w poses.  Taken automatically (bandpc_solver = 0) when the trajectory is long
// compared with the band; bandpc_solver = 2 takes it whenever the layout is possible (tests on small problems).
// One level of the layout: n_units blocks of `unit` scalars, half-bandwidth w units, C chunks.
bool Engine::plan_wband_level(int lvl, int n_units, int unit, int w, int C) {
    WbandLevel& L = wb[lvl];
    const int n = n_units;
    const int inter = n - (C - 1) * w;
    if (C < 1 || inter < C * (w + 1)) return false;
    const int base = inter / C, extra = inter % C;
    std::vector<int> owner(n), local(n), p0(C), len(C);
    int pos = 0;
    for (int c = 0; c < C; ++c) {
        len[c] = base + (c < extra ? 1 : 0);
        p0[c] = pos;
        for (int a = 0; a < len[c]; ++a) {
            owner[pos] = c;
            local[pos++] = a;
        }
        if (c + 1 < C)
            for (int a = 0; a < w; ++a) {
                owner[pos] = -(c + 1);
                local[pos++] = a;
            }
    }
    if (lvl == 0) {
        // every stored block must fit the layout (interior x interior of one chunk, interior x adjacent separator, or
        // inside one separator): guaranteed by the band width, checked because a miss would be a silent wrong answer
        for (int a = 0; a < n; ++a)
            for (int e = s_rowptr_h[a]; e < s_rowptr_h[a + 1]; ++e) {
                const int oa = owner[a], ob = owner[s_col_h[e]];
                const bool ok = (oa >= 0 && (ob == oa || ob == -(oa + 1))) || (oa < 0 && (ob == -oa || ob == oa));
                if (!ok) return false;
            }
    }
    const int nb = dense_panel_width();
    const int m_pad = (unit * (base + (extra ? 1 : 0)) + nb - 1) / nb * nb;
    const int sepw = C > 1 ? unit * w : 0, nbr = 2 * sepw + 1, ldB = nbr + (nbr & 1);
    const int bwr = (unit * (w + 1) - 1 + 7) / 8 * 8, ld = bwr + nb;
    const size_t a_stride = size_t(m_pad) * (ld + 1) + 2, b_stride = size_t(m_pad + ldB) * ldB;
    const size_t ns = size_t(C - 1) * sepw, ns_pad = (ns + nb - 1) / nb * nb;
    const size_t bytes = 8 * (C * (a_stride + b_stride) + (ns_pad + 8) * (ns_pad + 1));
    if (bytes > (size_t(12) << 30)) return false;
    L.n_units = n;
    L.unit = unit;
    L.w = w;
    L.C = C;
    L.m_pad = m_pad;
    L.r_start = C > 1 ? unit * (base - w) / nb * nb : 0;   // the right separator couples to a chunk's last w units only
    L.owner.upload(owner, stream);
    L.local.upload(local, stream);
    L.p0.upload(p0, stream);
    L.len.upload(len, stream);
    L.A.alloc(a_stride * C, stream);
    L.Bd.alloc(b_stride * C, stream);
    L.Ld.alloc(size_t(C) * (m_pad / nb) * nb * nb, stream);
    L.xw.alloc(size_t(C) * m_pad, stream);
    if (C > 1) {
        L.xsep.alloc(ns, stream);
        L.T.alloc((ns_pad + 8) * (ns_pad + 1), stream);
        L.TLd.alloc((ns_pad / nb) * nb * nb, stream);
        L.Txw.alloc(ns_pad, stream);
    }
    if (std::getenv("CSLAM_DEBUG_SOLVER"))
        std::fprintf(stderr, "[solver] wide-band Cholesky%s: %d blocks of %d, half-bandwidth %d, %d chunks of <= %d unknowns, %zu separator unknowns\n",
                     lvl ? " (separator level)" : "", n, unit, w, C, m_pad, ns);
    return true;
}

bool Engine::plan_wband_solver(int w) {
    if (w > kWbandMaxW) return false;
    const int n = n_free;
    const bool forced = opt.bandpc_solver == 2;
    if (!forced && (n < 256 || n < 8 * (w + 1))) return false;
    // Cost model (us, measured on the ragged 5 k-pose track): a chunk panel of 48 columns costs ~19 + 0.3 C (factor + trailing
    // update — its tensor-core work grows with the number of chunks in the launch — + its share of the one-launch
    // back-substitution), a panel of a dense separator system ~26.  One level: C chunks + the dense solve of (C - 1) 6w
    // separator unknowns.  Two levels: the separator system is block tridiagonal with blocks of 6w and is chunked again
    // (separators of ONE block; ~19 + 1.2 C2 per panel: three times the tiles per chunk), its own separators go to the
    // dense solver.  The optimum is flat (20 .. 27 chunks + 4 .. 5 measured within 3 %).
    const double td = 26.0, sb = 6.0 * w / 48.0;   // sb: panels per separator block
    auto panels = [](double scalars) { return std::ceil(scalars / 48.0); };
    const int Cmax = std::max(1, std::min(1 + kDenseMaxN / (6 * w), (n + w) / (3 * w + 1)));
    static const int force_levels = [] {
        const char* e = std::getenv("CSLAM_WBAND_LEVELS");   // A/B knob: 1 = never chunk the separator system
        return e ? std::atoi(e) : 0;
    }();
    double best = 1e300;
    int bestC = 1, bestC2 = 0;
    for (int C = 1; C <= Cmax; ++C) {
        const double ns = (C - 1) * 6.0 * w;
        const double chunk = panels(6.0 * (n - (C - 1) * w) / C) * (19.0 + 0.3 * C) + ns * ns * 8.0 / 6e6;   // (+ clearing the dense T)
        const double one = chunk + panels(ns) * td;
        if (one < best) {
            best = one;
            bestC = C;
            bestC2 = 0;
        }
        const int B = C - 1;
        if (force_levels == 1) continue;
        for (int C2 = 2; B >= 8 && C2 <= (B + 1) / 3; ++C2) {
            const double two = chunk + panels(6.0 * w * std::ceil(double(B - (C2 - 1)) / C2)) * (19.0 + 1.2 * C2) +
                               std::ceil((C2 - 1) * sb) * td + 40.0;
            if (two < best) {
                best = two;
                bestC = C;
                bestC2 = C2;
            }
        }
    }
    if (const char* e = std::getenv("CSLAM_WBAND_CHUNKS")) {   // A/B knob: "C0" or "C0,C1" chunk counts per level
        int c0 = 0, c1 = 0;
        if (std::sscanf(e, "%d,%d", &c0, &c1) >= 1 && c0 >= 1 && c0 <= Cmax) {
            bestC = c0;
            bestC2 = c1;
        }
    }
    wb_levels = 0;
    if (!plan_wband_level(0, n, 6, w, bestC)) return false;
    wb_levels = 1;
    if (bestC2 >= 2 && plan_wband_level(1, bestC - 1, 6 * w, 1, bestC2)) wb_levels = 2;
    if (!d_band_fail.p) d_band_fail.alloc(1, stream);
    wband_active = true;
    return true;
}

WbandView Engine::wband_view(int lvl) {
    const int nb = dense_panel_width();
    WbandLevel& L = wb[lvl];
    WbandView V;
    V.n_free = L.n_units;
    V.unit = L.unit;
    V.next = nullptr;
    V.w = L.w;
    V.C = L.C;
    V.m_pad = L.m_pad;
    V.sepw = L.C > 1 ? L.unit * L.w : 0;
    V.nbr = 2 * V.sepw + 1;
    V.bwr = (L.unit * (L.w + 1) - 1 + 7) / 8 * 8;
    V.ld = V.bwr + nb;
    V.ldB = V.nbr + (V.nbr & 1);
    V.r_start = L.r_start;
    V.a_stride = (long long)V.m_pad * (V.ld + 1) + 2;
    V.b_stride = (long long)(V.m_pad + V.ldB) * V.ldB;
    V.rowptr = d_s_rowptr.p;
    V.col = d_s_col.p;
    V.S = d_S;
    V.rhs = nullptr;
    V.owner = L.owner.p;
    V.local = L.local.p;
    V.chunk_p0 = L.p0.p;
    V.chunk_len = L.len.p;
    V.A = L.A.p;
    V.Bd = L.Bd.p;
    V.Ldiag = L.Ld.p;
    V.xw = L.xw.p;
    V.xsep = L.xsep.p;
    V.T.n = (L.C - 1) * V.sepw;
    V.T.n_pad = (V.T.n + nb - 1) / nb * nb;
    V.T.ld = V.T.n_pad + 8;
    V.T.rowptr = nullptr;
    V.T.col = nullptr;
    V.T.S = nullptr;
    V.T.rhs = nullptr;
    V.T.A = L.T.p;
    V.T.Ldiag = L.TLd.p;
    V.T.y = L.xsep.p;
    V.T.fail = d_band_fail.p;
    V.Txw = L.Txw.p;
    V.y = nullptr;
    V.fail = d_band_fail.p;
    return V;
}

// Conjugate gradients on S y = rhs preconditioned with the banded direct solve (kernels_bandpcg.cu).  Host-driven:
// three scalar read-backs per iteration, ~10 iterations.  Status in d_pscal like the other solvers.
void Engine::bandpc_solve(const double* rhs, double* y) {
    const int n = n_free;
    const long long n6 = 6ll * n;
    BandView V;   // (the preconditioner's matrix d_bpc_S was formed in schur_pass)
    V.n = n;
    V.w = band_w;
    V.P = band_P;
    V.m = band_m;
    V.sep_solver = band_sep;
    V.band_idx = d_band_idx.p;
    V.S = d_bpc_S.p;
    V.Lbuf = d_Lbuf.p;
    V.Xbuf = d_Xbuf.p;
    V.Ta = d_Ta.p;
    V.Ca = d_Ca.p;
    V.fa = d_fa.p;
    V.Tb = d_Tb.p;
    V.fb = d_fb.p;
    V.fail = d_band_fail.p;
    const BandScratch K{d_T2.p, d_rhs2.p, d_L2.p, d_X2.p, d_y2.p};
    double *r = d_pr.p, *z = d_pz.p, *pv = d_pp.p, *q = d_pq.p;
    CSLAM_CUDA(cudaMemsetAsync(y, 0, n6 * sizeof(double), stream));
    CSLAM_CUDA(cudaMemcpyAsync(r, rhs, n6 * sizeof(double), cudaMemcpyDeviceToDevice, stream));
    auto dot = [&](const double* a, const double* b) {
        CSLAM_CUDA(cudaMemsetAsync(d_bpc_scal.p, 0, sizeof(double), stream));
        launch_dot(stream, a, b, n6, d_bpc_scal.p);
        double v;
        read_scalars(d_bpc_scal.p, &v, 1);
        return v;
    };
    const double normb2 = dot(rhs, rhs);
    if (std::getenv("CSLAM_BPC_DEBUG")) std::fprintf(stderr, "[bandpc] |b|^2 %.6e (n %d, w %d, P %d)\n", normb2, n, band_w, band_P);
    const double tol2 = 1e-15 * 1e-15 * normb2;
    double rho_old = 0.0;
    int it = 0;
    bool fail = false;
    const int max_it = 2000;
    if (normb2 > 0.0)
        for (; it < max_it;) {
            V.rhs = r;
            V.y = z;
            launch_band_solve(stream, V, K, d_pscal.p);
            const double rho = dot(r, z);
            double ps[PS_COUNT];
            read_scalars(d_pscal.p, ps, PS_COUNT);
            if (ps[PS_FAIL] == 2.0 || !std::isfinite(rho) || !(rho > 0.0)) {
                fail = !(rho == 0.0);
                if (std::getenv("CSLAM_BPC_DEBUG"))
                    std::fprintf(stderr, "[bandpc] stopped at it %d: band factorisation %s, rho %.3e, |b|^2 %.3e\n", it,
                                 ps[PS_FAIL] == 2.0 ? "FAILED" : "ok", rho, normb2);
                break;
            }
            launch_bpc_xpby(stream, n6, z, it == 0 ? 0.0 : rho / rho_old, pv);
            launch_bpc_spmv(stream, n, d_s_rowptr.p, d_s_col.p, d_lt_rowptr.p, d_lt_col.p, d_S, pv, q);
            const double pq = dot(pv, q);
            if (!std::isfinite(pq) || !(pq > 0.0)) {
                fail = true;
                if (std::getenv("CSLAM_BPC_DEBUG"))
                    std::fprintf(stderr, "[bandpc] stopped at it %d: p.Sp = %.3e (rho %.3e): the system is not positive definite along p\n", it, pq, rho);
                break;
            }
            launch_bpc_update(stream, n6, rho / pq, pv, q, y, r);
            rho_old = rho;
            ++it;
            const double nr2 = dot(r, r);
            static const bool dbg = std::getenv("CSLAM_BPC_DEBUG") != nullptr;
            if (dbg) std::fprintf(stderr, "[bandpc] it %d  |r|/|b| %.3e  rho %.3e  pq %.3e\n", it, std::sqrt(nr2 / normb2), rho, pq);
            if (nr2 <= tol2) break;
        }
    double ps[PS_COUNT] = {0};
    ps[PS_ITERS] = double(std::max(it, 1));
    ps[PS_FAIL] = fail ? 2.0 : 0.0;
    CSLAM_CUDA(cudaMemcpyAsync(d_pscal.p, ps, sizeof(ps), cudaMemcpyHostToDevice, stream));
    CSLAM_CUDA(cudaStreamSynchronize(stream));
}

// The exact solve of a reduced system that is not a narrow band: dense Cholesky when it is small (PCG run
// to 1e-15 needs hundreds of iterations there and is not exact) or dense enough to be a real contraction.
void Engine::plan_dense_solver() {
    dense_active = false;
    if (opt.linear_solver != 0 || n_free <= 0 || opt.dense_solver < 0 || ph.active || wband_active) return;
    const long long n = 6ll * n_free;
    if (n > kDenseMaxN) return;
    // auto: whenever the band solver does not apply and the factor fits — measured on a closed 500-pose loop
    // (n = 2994): 2.8 ms against 81 ms of PCG (6020 iterations to 1e-15); at n = 12 k PCG needs 24 k iterations
    if (opt.dense_solver == 0 && band_active) return;
    const int nb = dense_panel_width();
    dense_npad = int((n + nb - 1) / nb) * nb;
    dense_ld = dense_npad + 8;
    d_dense_A.alloc(size_t(dense_ld) * size_t(dense_npad + 1), stream);
    d_dense_Ld.alloc(size_t(dense_npad / nb) * nb * nb, stream);
    d_dense_xw.alloc(size_t(dense_npad), stream);
    if (!d_band_fail.p) d_band_fail.alloc(1, stream);
    dense_active = true;
    band_active = false;
}

void Engine::alloc_band_sets(int count) {
    if (ph.fan_graph) {
        cudaGraphExecDestroy(ph.fan_graph);  // captured over the old buffers
        ph.fan_graph = nullptr;
    }
    ph.fan_calls = 0;
    band_sets.clear();
    if (!band_active || count <= 0) return;
    const int n = n_free, W = band_storage_width(band_w), P = band_P;
    const size_t b = 6 * size_t(W), GC = P > 1 ? 1 + b : 1;
    const size_t n2 = size_t(std::max(P - 1, 1)) * W, W2 = 2 * size_t(W) - 1;
    const size_t per_set = (size_t(n) * (W + 1) * 36 + size_t(n) * 6 * GC + 3 * size_t(P) * b * b + 2 * n2 * (W2 + 1) * 36) * 8;
    count = int(std::min<size_t>(size_t(count), std::max<size_t>(1, (size_t(2) << 30) / std::max<size_t>(per_set, 1))));
    if (!ev_fork) CSLAM_CUDA(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
    // one allocation for every set (hundreds of cudaMalloc calls cost seconds)
    const size_t sizes[13] = {size_t(n) * (W + 1) * 36, size_t(n) * 6 * GC, size_t(P) * b * b, size_t(P) * b * b,
                              size_t(P) * b * b,        size_t(P) * b,      size_t(P) * b,     n2 * (W2 + 1) * 36,
                              n2 * (W2 + 1) * 36,       6 * n2,             6 * n2,            6 * n2,
                              size_t(PS_COUNT)};
    size_t set_doubles = 2;  // the fail flag
    for (size_t z : sizes) set_doubles += (z + 1) & ~size_t(1);
    band_pool.alloc(set_doubles * size_t(count), stream);
    for (int i = 0; i < count; ++i) {
        auto bs = std::make_unique<BandSet>();
        CSLAM_CUDA(cudaStreamCreateWithFlags(&bs->stream, cudaStreamNonBlocking));
        CSLAM_CUDA(cudaEventCreateWithFlags(&bs->done, cudaEventDisableTiming));
        double* base = band_pool.p + set_doubles * size_t(i);
        bs->fail = reinterpret_cast<int*>(base);
        base += 2;
        double** slots[13] = {&bs->Lbuf, &bs->Xbuf, &bs->Ta, &bs->Ca, &bs->Tb, &bs->fa, &bs->fb, &bs->T2,
                              &bs->L2,   &bs->rhs2, &bs->X2, &bs->y2, &bs->ps};
        for (int k = 0; k < 13; ++k) {
            *slots[k] = base;
            base += (sizes[k] + 1) & ~size_t(1);
        }
        band_sets.push_back(std::move(bs));
    }
}

// One banded solve S_cc y = rhs on the set's own stream and scratch (enqueue only)
void Engine::solve_reduced_on(BandSet& bs, const double* rhs, double* y) {
    BandView V;
    V.n = n_free;
    V.w = band_w;
    V.P = band_P;
    V.m = band_m;
    V.sep_solver = band_sep;
    V.band_idx = d_band_idx.p;
    V.S = d_S;
    V.rhs = rhs;
    V.Lbuf = bs.Lbuf;
    V.Xbuf = bs.Xbuf;
    V.Ta = bs.Ta;
    V.Ca = bs.Ca;
    V.fa = bs.fa;
    V.Tb = bs.Tb;
    V.fb = bs.fb;
    V.y = y;
    V.fail = bs.fail;
    const BandScratch K{bs.T2, bs.rhs2, bs.L2, bs.X2, bs.y2};
    launch_band_solve(bs.stream, V, K, bs.ps);
}

void Engine::reset_state() {
    if (!uploaded) throw std::invalid_argument("reset_state before upload");
    CSLAM_CUDA(cudaMemcpyAsync(d_poses.p, d_poses_init.p, d_poses.bytes(), cudaMemcpyDeviceToDevice, stream));
    CSLAM_CUDA(cudaMemcpyAsync(d_poses_best.p, d_poses_init.p, d_poses.bytes(), cudaMemcpyDeviceToDevice, stream));
    if (n_lm) {
        CSLAM_CUDA(cudaMemcpyAsync(d_points.p, d_points_init.p, d_points.bytes(), cudaMemcpyDeviceToDevice, stream));
        CSLAM_CUDA(cudaMemcpyAsync(d_points_best.p, d_points_init.p, d_points.bytes(), cudaMemcpyDeviceToDevice, stream));
    }
    if (ph.active) {
        auto cp = [&](DBuf<double>& dst, const DBuf<double>& src) {
            if (src.n) CSLAM_CUDA(cudaMemcpyAsync(dst.p, src.p, src.bytes(), cudaMemcpyDeviceToDevice, stream));
        };
        cp(ph.normals, ph.normals_init);
        cp(ph.normals_best, ph.normals_init);
        cp(ph.gx, ph.gx_init);
        cp(ph.gx_best, ph.gx_init);
    }
    begun = false;
}

void Engine::download() {
    if (!uploaded) throw std::invalid_argument("download before upload");
    std::vector<double> pos(12 * size_t(n_poses));
    CSLAM_CUDA(cudaMemcpyAsync(pos.data(), d_poses_best.p, d_poses_best.bytes(), cudaMemcpyDeviceToHost, stream));
    if (n_lm || (n_ranks > 1 && n_points)) {
        // best landmarks back into the caller's point order on the device, then one copy straight
        // into the caller's array
        if (n_ranks > 1) {
            // every rank returns the COMPLETE solution: the shards' landmarks are summed over the ranks
            // (collective: all ranks of the communicator call download / cslam_solve together)
            DBuf<double> z4;
            z4.alloc(4 * size_t(n_points), stream);
            CSLAM_CUDA(cudaMemsetAsync(z4.p, 0, z4.bytes(), stream));
            launch_scatter_points4(stream, n_lm, d_lm_user.p, d_points_best.p, z4.p);
            comm_allreduce_sum(nccl_comm, z4.p, 4 * size_t(n_points), stream);
            launch_merge_points4(stream, (long long)n_points, z4.p, d_raw_pts.p);
            z4.release_async(stream);
        } else {
            launch_scatter_points(stream, n_lm, d_lm_user.p, d_points_best.p, d_raw_pts.p);
        }
        CSLAM_CUDA(cudaStreamSynchronize(stream));
        parallel_d2h(opt.device, h_points, d_raw_pts.p, 3 * size_t(n_points) * sizeof(double));
    }
    std::vector<double> nrm, gxh, nrm_all;
    if (ph.active) {
        nrm.resize(3 * size_t(n_lm));
        gxh.resize(ph.n_g);
        if (n_ranks > 1) {
            // the shards' normals are summed over the ranks into the caller's order, like the points above
            DBuf<double> z4;
            z4.alloc(4 * size_t(n_points), stream);
            CSLAM_CUDA(cudaMemsetAsync(z4.p, 0, z4.bytes(), stream));
            launch_scatter_points4(stream, n_lm, d_lm_user.p, ph.normals_best.p, z4.p);
            comm_allreduce_sum(nccl_comm, z4.p, 4 * size_t(n_points), stream);
            nrm_all.resize(4 * size_t(n_points));
            CSLAM_CUDA(cudaMemcpyAsync(nrm_all.data(), z4.p, nrm_all.size() * sizeof(double), cudaMemcpyDeviceToHost, stream));
            CSLAM_CUDA(cudaStreamSynchronize(stream));
            z4.release_async(stream);
        }
        if (n_lm) CSLAM_CUDA(cudaMemcpyAsync(nrm.data(), ph.normals_best.p, nrm.size() * sizeof(double), cudaMemcpyDeviceToHost, stream));
        CSLAM_CUDA(cudaMemcpyAsync(gxh.data(), ph.gx_best.p, gxh.size() * sizeof(double), cudaMemcpyDeviceToHost, stream));
    }
    CSLAM_CUDA(cudaStreamSynchronize(stream));
    for (int k : free_cams_h) std::memcpy(h_poses + 12 * size_t(k), &pos[12 * size_t(k)], 96);
    if (ph.active) {
        if (n_ranks > 1) {
            for (size_t j = 0; j < size_t(n_points); ++j)
                if (nrm_all[4 * j + 3] != 0.0) std::memcpy(h_normals + 3 * j, &nrm_all[4 * j], 24);  // weight slot: owned by some rank
        } else {
            for (int j = 0; j < n_lm; ++j) std::memcpy(h_normals + 3 * size_t(lm_user_h[j]), &nrm[3 * size_t(j)], 24);
        }
        const int t0 = 3 * ph.n_mat, l0 = t0 + ph.n_tex;
        for (int k = 0; k < ph.n_g; ++k) {
            if (!ph.g_used_h[k]) continue;
            if (k < t0)
                h_phong[k] = gxh[k];
            else if (k < l0)
                h_tex_shared[k - t0] = gxh[k];
            else
                h_light[k - l0] = gxh[k];
        }
    }
}

// -------------------------------------------------------------------------------------------------
// Joint lighting solve (dataset_ba_phong.cpp:100-252): vertex = position + normal, shared blocks
// -------------------------------------------------------------------------------------------------
void Engine::check_phong_solve() {
    auto not_impl = [](const char* m) { throw NotImplemented(m); };
    if (!h_normals || !h_material_id || !h_phong || !h_light) throw std::invalid_argument("vertices / materials / light not set");
    if (n_vertices != n_points) throw std::invalid_argument("one normal / material id per point expected");
    if (!h_tex_shared) not_impl("lighting solve: textures must be shared blocks (cslam_set_textures)");
    if (n_ph != n_st || std::memcmp(ph_cam, st_cam, n_st * sizeof(uint32_t)) != 0 ||
        std::memcmp(ph_vertex, st_pt, n_st * sizeof(uint32_t)) != 0)
        not_impl("lighting solve: lighting blocks must pair one-to-one with the stereo blocks");
    if (st_W_per_obs || !suns.empty() || !priors.empty()) not_impl("lighting solve: shared stereo stiffness, no sun / prior blocks");
    for (uint32_t j = 0; j < n_vertices; ++j)
        if (h_material_id[j] >= n_materials) throw std::invalid_argument("material id out of range");
    ph.n_mat = int(n_materials);
    ph.n_tex = int(n_tex_shared);
    ph.n_g = 3 * ph.n_mat + ph.n_tex + 3;
    // the dense border system [n_g][n_g + 1] is factored inside one CTA's shared memory up to 160 columns and in
    // global memory beyond; the shared blocks' own kernels are one CTA with a thread per column
    if (ph.n_g > 1024) not_impl("lighting solve: at most 1024 shared columns (3 per material + 1 per texture + 3)");
    ph.max_track = 0;
    // (vertices observed more than 32 times take the chunked kernels of kernels_phong_long.cu)
    for (int j = 0; j < n_lm; ++j) ph.max_track = std::max(ph.max_track, int(lm_cnt_h[j]));
}

void Engine::setup_phong_solve() {
    const size_t nl = size_t(std::max(n_lm, 1)), no = size_t(std::max<long long>(n_obs, 1));
    std::vector<double> nrm(3 * nl, 0.0), gxh(ph.n_g), oI(no, 0.0), on(3 * no, 0.0);
    std::vector<int> vm(nl, 0), vt(nl, 0);
    ph.g_used_h.assign(ph.n_g, 0);
    const int t0 = 3 * ph.n_mat, l0 = t0 + ph.n_tex;
    for (int j = 0; j < n_lm; ++j) {
        const uint32_t u = lm_user_h[j];
        std::memcpy(&nrm[3 * size_t(j)], h_normals + 3 * size_t(u), 24);
        vm[j] = int(h_material_id[u]);
        vt[j] = int(h_texture_id[u]);
        for (int k = 0; k < 3; ++k) ph.g_used_h[3 * vm[j] + k] = 1;
        ph.g_used_h[t0 + vt[j]] = 1;
    }
    if (n_ranks > 1) {
        // the set of referenced shared columns must be the same on every rank (it decides which columns of the
        // all-reduced border exist): take it from every observed vertex, not from this rank's shard
        for (uint64_t i = 0; i < n_st; ++i) {
            const uint32_t u = st_pt[i];
            for (int k = 0; k < 3; ++k) ph.g_used_h[3 * int(h_material_id[u]) + k] = 1;
            ph.g_used_h[t0 + int(h_texture_id[u])] = 1;
        }
    }
    if (n_lm || (n_ranks > 1 && n_st))
        for (int k = 0; k < 3; ++k) ph.g_used_h[l0 + k] = 1;
    parallel_chunks(size_t(n_obs), size_t(1) << 16, [&](int, size_t c0, size_t c1) {
        for (size_t e = c0; e < c1; ++e) {
            const uint32_t u = obs_user_h[e];
            oI[e] = ph_intensity[u];
            for (int k = 0; k < 3; ++k) on[size_t(k) * no + e] = ph_normal_obs[3 * size_t(u) + k];
        }
    });
    std::memcpy(gxh.data(), h_phong, 3 * size_t(ph.n_mat) * sizeof(double));
    std::memcpy(gxh.data() + t0, h_tex_shared, size_t(ph.n_tex) * sizeof(double));
    std::memcpy(gxh.data() + l0, h_light, 24);
    ph.normals_init.upload(nrm, stream);
    ph.gx_init.upload(gxh, stream);
    ph.v_mat.upload(vm, stream);
    ph.v_tex.upload(vt, stream);
    ph.g_used.upload(ph.g_used_h, stream);
    ph.obs_I.upload(oI, stream);
    ph.obs_n.upload(on, stream);
    for (DBuf<double>* b : {&ph.normals, &ph.normals_cand, &ph.normals_best, &ph.sc_n, &ph.cn_n}) b->alloc(3 * nl, stream);
    for (DBuf<double>* b : {&ph.gx, &ph.gx_cand, &ph.gx_best, &ph.sc_g, &ph.yg, &ph.diag_g, &ph.zero_g}) b->alloc(ph.n_g, stream);
    ph.zero_g.zero(stream);
    ph.gv.alloc(6 * nl, stream);
    ph.yv.alloc(6 * nl, stream);
    ph.X.alloc(size_t(ph.n_g + 1) * 6 * size_t(std::max(n_free, 1)), stream);
    ph.T.alloc(size_t(ph.n_g) * (ph.n_g + 1), stream);
    alloc_band_sets(ph.n_g + 1);
    CSLAM_CUDA(cudaStreamSynchronize(stream));  // the host staging vectors go out of scope
}

PhongSolveView Engine::phong_solve_view(const double* normals, const double* gx) const {
    PhongSolveView q;
    q.normals = normals;
    q.v_mat = ph.v_mat.p;
    q.v_tex = ph.v_tex.p;
    q.gx = gx;
    q.obs_I = ph.obs_I.p;
    q.obs_n = ph.obs_n.p;
    q.sc_n = ph.sc_n.p;
    q.sc_g = ph.sc_g.p;
    q.g_used = ph.g_used.p;
    std::memcpy(q.Wn, ph_W_normal, sizeof(q.Wn));
    q.int_stiffness = ph_int_stiffness;
    q.directional = light_directional;
    q.n_mat = ph.n_mat;
    q.n_tex = ph.n_tex;
    q.n_g = ph.n_g;
    q.hold_positions = hold_positions ? 1 : 0;
    for (int k = 0; k < 3; ++k) q.mat_lo[k] = mat_lo[k], q.mat_hi[k] = mat_hi[k];
    q.tex_lo = tex_lo;
    q.tex_hi = tex_hi;
    return q;
}

PhongSystem Engine::phong_system() {
    PhongSystem o;
    o.S = d_S;
    o.Bdiag = d_Bdiag;
    o.bp = d_bp;
    o.gp = d_gp;
    o.Scg = ph.Scg;
    o.Sgg = ph.Sgg;
    o.bg = ph.bg;
    o.gg = ph.gg;
    o.hg = ph.hg;
    o.gv = ph.gv.p;
    o.cn_l = d_cn_l.p;
    o.cn_n = ph.cn_n.p;
    o.scal = d_scal;
    return o;
}

// The arrowhead system [S_cc S_cg; S_cg^T S_gg]: n_g + 1 solves against S_cc (border columns and
// b_c), then the dense n_g x n_g border system, then y_c = x_b - X_g y_g.
void Engine::phong_linear_solve(int* iters, bool* ok) {
    *iters = 1;
    *ok = true;
    prof_begin(CSLAM_K_PCG);
    const size_t nf6 = 6 * size_t(n_free);
    bool side_fail = false;
    if (n_free > 0 && band_active && !band_sets.empty()) {
        // the n_g + 1 banded solves are independent: fan them out over the scratch sets' streams.
        // ~300 small enqueues per LM iteration are host-bound, so from the second call on the
        // fan-out is replayed as ONE captured CUDA graph (same buffers every iteration).
        auto fan_out = [&]() {
            CSLAM_CUDA(cudaEventRecord(ev_fork, stream));
            for (auto& bs : band_sets) CSLAM_CUDA(cudaStreamWaitEvent(bs->stream, ev_fork, 0));
            size_t slot = 0;
            for (int k = 0; k <= ph.n_g; ++k) {
                double* y = ph.X.p + size_t(k) * nf6;
                if (k < ph.n_g && !ph.g_used_h[k]) {
                    CSLAM_CUDA(cudaMemsetAsync(y, 0, nf6 * sizeof(double), stream));
                    continue;
                }
                const double* rhs = k < ph.n_g ? ph.Scg + size_t(k) * nf6 : d_bp;
                solve_reduced_on(*band_sets[slot++ % band_sets.size()], rhs, y);
            }
            for (auto& bs : band_sets) {
                CSLAM_CUDA(cudaEventRecord(bs->done, bs->stream));
                CSLAM_CUDA(cudaStreamWaitEvent(stream, bs->done, 0));
            }
        };
        if (ph.fan_graph) {
            CSLAM_CUDA(cudaGraphLaunch(ph.fan_graph, stream));
            g_kernel_launches.fetch_add(ph.fan_graph_kernels, std::memory_order_relaxed);
        } else if (ph.fan_calls++ == 0) {
            fan_out();  // plain the first time (function attributes, lazy module loading)
        } else {
            const unsigned long long k0 = g_kernel_launches.load();
            CSLAM_CUDA(cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal));
            cudaGraph_t graph = nullptr;
            try {
                fan_out();
            } catch (...) {
                cudaStreamEndCapture(stream, &graph);
                if (graph) cudaGraphDestroy(graph);
                throw;
            }
            CSLAM_CUDA(cudaStreamEndCapture(stream, &graph));
            ph.fan_graph_kernels = g_kernel_launches.load() - k0;
            CSLAM_CUDA(cudaGraphInstantiate(&ph.fan_graph, graph, 0));
            CSLAM_CUDA(cudaGraphDestroy(graph));
            CSLAM_CUDA(cudaGraphLaunch(ph.fan_graph, stream));
        }
        launch_fill(stream, d_pscal.p, PS_COUNT, 0.0);
        // every solve factors the same S: one status tells them all
        CSLAM_CUDA(cudaMemcpyAsync(h_pinned + 64, band_sets[0]->ps, PS_COUNT * sizeof(double), cudaMemcpyDeviceToHost, stream));
        side_fail = true;  // checked after the synchronising read below
    } else if (n_free > 0) {
        for (int k = 0; k < ph.n_g; ++k) {
            if (ph.g_used_h[k])
                solve_reduced(ph.Scg + size_t(k) * nf6, ph.X.p + size_t(k) * nf6);
            else
                CSLAM_CUDA(cudaMemsetAsync(ph.X.p + size_t(k) * nf6, 0, nf6 * sizeof(double), stream));
        }
        solve_reduced(d_bp, ph.X.p + size_t(ph.n_g) * nf6);
    } else {
        launch_fill(stream, d_pscal.p, PS_COUNT, 0.0);
    }
    launch_phong_border_solve(stream, ph.n_g, int(nf6), ph.Scg, ph.X.p, ph.Sgg, ph.bg, ph.T.p, ph.yg.p, d_yp.p, d_pscal.p);
    if (n_ranks > 1) {
        // every rank solved the same (all-reduced) system; rank 0's iterate is adopted everywhere so that the
        // ranks stay bit-identical (same accept / reject decisions), as in run_pcg
        if (nf6) comm_broadcast(nccl_comm, d_yp.p, nf6, 0, stream);
        comm_broadcast(nccl_comm, ph.yg.p, size_t(ph.n_g), 0, stream);
        comm_broadcast(nccl_comm, d_pscal.p, PS_COUNT, 0, stream);
    }
    double ps[PS_COUNT];
    read_scalars(d_pscal.p, ps, PS_COUNT);
    prof_end(CSLAM_K_PCG);
    *ok = ps[PS_FAIL] != 2.0;
    if (side_fail && h_pinned[64 + PS_FAIL] == 2.0) *ok = false;
}

void Engine::allreduce_system() {
    if (n_ranks <= 1) return;
    prof_begin(CSLAM_K_ALLREDUCE);
    comm_allreduce_sum(nccl_comm, d_red.p, red_count, stream);
    prof_end(CSLAM_K_ALLREDUCE);
}
void Engine::allreduce_small(double* dev, int n) {
    if (n_ranks <= 1) return;
    comm_allreduce_sum(nccl_comm, dev, size_t(n), stream);
}

// -------------------------------------------------------------------------------------------------
// One Schur build at (x, radius): S, rhs, gradient, cost — leaves them in d_red
// -------------------------------------------------------------------------------------------------
void Engine::schur_pass() {
    const LmDiag dg = current_diag();
    DevView v = view(d_poses.p, d_points.p);
    prof_begin(CSLAM_K_SCHUR);
    d_red.zero(stream);
    if (ph.active) {
        // grouped vertices: one CTA per work item, the 6L x 6L tile of the group in DMMA accumulators
        // (CSLAM_PHONG_GROUPED=0: every vertex through the warp-per-vertex kernel, for A/B runs and tests)
        static const bool use_grouped = [] {
            const char* e = std::getenv("CSLAM_PHONG_GROUPED");
            return !(e && std::atoi(e) == 0);
        }();
        const PhongSolveView pq = phong_solve_view(ph.normals.p, ph.gx.p);
        int first_rest = 0;
        if (use_grouped && !item_group_h.empty()) {
            launch_phong_build_grouped(stream, v, pq, group_view(), n_items_small, dg, phong_system());
            first_rest = n_lm_grouped;
        }
        launch_phong_build(stream, v, pq, first_rest, n_lm, dg, phong_system(), true, ph.max_track);
    } else {
        if (bandpc_active) {
            d_S2.zero(stream);
            d_Bdiag2.zero(stream);
        }
        launch_schur(v, dg);
        if (rank == 0)
            launch_camonly_build(stream, v, d_suns.p, int(suns.size()), d_priors.p, int(priors.size()), d_Bdiag, d_bp, d_gp,
                                 d_scal);
    }
    prof_end(CSLAM_K_SCHUR);
    allreduce_system();
    if (bandpc_active) {
        if (n_ranks > 1) {
            comm_allreduce_sum(nccl_comm, d_S2.p, d_S2.n, stream);
            comm_allreduce_sum(nccl_comm, d_Bdiag2.p, d_Bdiag2.n, stream);
        }
        // preconditioner M = [short-track landmarks' Schur terms] + damping, in dense band storage; then S, Bdiag get
        // the long tracks' terms
        launch_bpc_build(stream, n_free, band_w, d_s_rowptr.p, d_s_col.p, d_S, d_Bdiag, d_Bdiag2.p, dg, d_bpc_S.p);
        launch_bpc_add(stream, 36ll * nnzU, d_S2.p, d_S);
        launch_bpc_add(stream, 36ll * n_free, d_Bdiag2.p, d_Bdiag);
    }
    prof_begin(CSLAM_K_FINALIZE);
    launch_finalize(stream, v, dg, opt.preconditioner, d_S, d_Bdiag, d_diag_p.p, d_Minv.p, d_scal);
    if (ph.active)
        launch_phong_gfinalize(stream, phong_solve_view(ph.normals.p, ph.gx.p), dg, ph.Sgg, ph.bg, ph.hg, ph.diag_g.p);
    prof_end(CSLAM_K_FINALIZE);
    lm.have_system = true;
}

void Engine::gradient_norm_pass() {
    DevView v = view(d_poses.p, d_points.p);
    // SC_GRADMAX / SC_XNORM2_CUR are not touched by the Schur kernels
    if (ph.active) {
        launch_gradnorm(stream, v, 0, 0, d_gp, d_gl.p, d_scal, rank == 0);
        launch_phong_gradnorm(stream, v, phong_solve_view(ph.normals.p, ph.gx.p), 0, n_lm, ph.gv.p, ph.gg, d_scal, rank == 0);
    } else {
        launch_gradnorm(stream, v, 0, n_lm, d_gp, d_gl.p, d_scal, rank == 0);
    }
    double loc[SC_COUNT];
    if (n_ranks > 1) {
        // max over ranks == max of the per-rank maxima: all-reduce the squared-norm slot with sum
        // and the max slot with max
        comm_allreduce_max(nccl_comm, d_scal + SC_GRADMAX, 1, stream);
        comm_allreduce_sum(nccl_comm, d_scal + SC_XNORM2_CUR, 1, stream);
    }
    read_scalars(d_scal, loc, SC_COUNT);
    lm.gradient_max_norm = loc[SC_GRADMAX];
    lm.x_norm = std::sqrt(loc[SC_XNORM2_CUR]);
    lm.grad_fresh = true;
}

// One solve S_cc y = rhs with the solver the options select (enqueue only; status in d_pscal)
void Engine::solve_reduced(const double* rhs, double* y) {
    PcgBufs B;
    B.rowptr = d_s_rowptr.p;
    B.col = d_s_col.p;
    B.ent_ptr = d_lt_rowptr.p;
    B.ent_cb = d_lt_col.p;
    B.S = d_S;
    B.Minv = d_Minv.p;
    B.b = rhs;
    B.x = y;
    B.r = d_pr.p;
    B.z = d_pz.p;
    B.p = d_pp.p;
    B.q = d_pq.p;
    B.ps = d_pscal.p;
    B.nf = n_free;
    if (n_free == 0) return;
    double q_tol, r_tol;
    int max_it, min_it;
    if (opt.linear_solver == 0) {
        // exact: run CG to the attainable residual; a 6x6 system converges in one step because
        // the block-Jacobi preconditioner is then the exact inverse
        q_tol = -1.0;
        r_tol = 1e-15;
        max_it = std::max(64, 12 * n_free + 32);
        min_it = 0;
    } else {
        q_tol = opt.eta;
        r_tol = -1.0;
        max_it = opt.max_linear_solver_iterations;
        min_it = opt.min_linear_solver_iterations;
    }
    if (wband_active) {
        WbandView V0 = wband_view(0), V1;
        V0.rhs = rhs;
        V0.y = y;
        if (wb_levels == 2) {
            V1 = wband_view(1);
            V1.y = V0.xsep;   // the separator level solves level 0's separator system in place of the dense solver
            V0.next = &V1;
        }
        launch_wband_solve(stream, V0, d_pscal.p);
    } else if (bandpc_active) {
        bandpc_solve(rhs, y);
    } else if (dense_active) {
        DenseView V;
        V.n = 6 * n_free;
        V.n_pad = dense_npad;
        V.ld = dense_ld;
        V.rowptr = d_s_rowptr.p;
        V.col = d_s_col.p;
        V.S = d_S;
        V.rhs = rhs;
        V.A = d_dense_A.p;
        V.Ldiag = d_dense_Ld.p;
        V.y = y;
        V.fail = d_band_fail.p;
        launch_dense_solve(stream, V, d_dense_xw.p, d_pscal.p);
    } else if (band_active) {
        BandView V;
        V.n = n_free;
        V.w = band_w;
        V.P = band_P;
        V.m = band_m;
        V.sep_solver = band_sep;
        V.band_idx = d_band_idx.p;
        V.S = d_S;
        V.rhs = rhs;
        V.Lbuf = d_Lbuf.p;
        V.Xbuf = d_Xbuf.p;
        V.Ta = d_Ta.p;
        V.Ca = d_Ca.p;
        V.fa = d_fa.p;
        V.Tb = d_Tb.p;
        V.fb = d_fb.p;
        V.y = y;
        V.fail = d_band_fail.p;
        const BandScratch K{d_T2.p, d_rhs2.p, d_L2.p, d_X2.p, d_y2.p};
        launch_band_solve(stream, V, K, d_pscal.p);
    } else {
        launch_pcg_persistent(stream, B, d_pp2.p, d_prec.p, q_tol, r_tol, min_it, max_it, 10);
    }
}

void Engine::run_pcg(int* iters, bool* ok) {
    *iters = 0;
    *ok = true;
    if (n_free == 0) return;
    prof_begin(CSLAM_K_PCG);
    solve_reduced(d_bp, d_yp.p);
    if (n_ranks > 1) {
        // every rank solved the same system; rank 0's iterate is adopted everywhere so that all
        // ranks keep bit-identical poses (and therefore take identical accept/reject decisions)
        comm_broadcast(nccl_comm, d_yp.p, 6 * size_t(n_free), 0, stream);
        comm_broadcast(nccl_comm, d_pscal.p, PS_COUNT, 0, stream);
    }
    double ps[PS_COUNT];
    read_scalars(d_pscal.p, ps, PS_COUNT);
    prof_end(CSLAM_K_PCG);
    *iters = int(ps[PS_ITERS]);
    *ok = ps[PS_FAIL] != 2.0;
}

void Engine::copy_phong_best() {
    if (!ph.active) return;
    if (n_lm) CSLAM_CUDA(cudaMemcpyAsync(ph.normals_best.p, ph.normals.p, ph.normals.bytes(), cudaMemcpyDeviceToDevice, stream));
    CSLAM_CUDA(cudaMemcpyAsync(ph.gx_best.p, ph.gx.p, ph.gx.bytes(), cudaMemcpyDeviceToDevice, stream));
}

// Back-substitution of the vertex blocks, model cost change, candidate point and its cost; for a
// bounded problem the Armijo search of TrustRegionMinimizer::DoLineSearch along the step (the
// candidate is then x (+) alpha * delta).  sc2 receives the scalars of the accepted candidate.
void Engine::phong_step(const LmDiag& dg, double* sc2) {
    DevView v = view(d_poses.p, d_points.p);
    const PhongSolveView q = phong_solve_view(ph.normals.p, ph.gx.p);
    prof_begin(CSLAM_K_BACKSUB);
    d_scal2.zero(stream);
    launch_phong_backsub(stream, v, q, 0, n_lm, dg, d_yp.p, ph.yg.p, ph.gv.p, ph.yv.p, d_scal2.p, ph.max_track);
    if (bounded && rank == 0) {
        launch_dot(stream, d_gp, d_yp.p, 6ll * n_free, d_scal2.p + SC_LS_GY);
        launch_dot(stream, ph.gg, ph.yg.p, ph.n_g, d_scal2.p + SC_LS_GY);
        launch_absmax_scaled(stream, d_yp.p, d_sc_p.p, 6ll * n_free, d_scal2.p + SC_LS_DMAX);
        launch_absmax_scaled(stream, ph.yg.p, ph.sc_g.p, ph.n_g, d_scal2.p + SC_LS_DMAX);
    }
    launch_phong_candidate(stream, v, q, 0, n_lm, 1.0, d_yp.p, ph.yg.p, ph.yv.p, d_poses_cand.p, ph.gx_cand.p, d_points_cand.p,
                           ph.normals_cand.p, d_scal2.p, rank == 0, ph.max_track);
    phong_reduce_scal2();
    read_scalars(d_scal2.p, sc2, SC_COUNT);
    if (bounded && sc2[SC_NONFINITE] == 0.0 && sc2[SC_MODEL] > 0.0) phong_line_search(d_yp.p, ph.yg.p, ph.yv.p, sc2);
    prof_end(CSLAM_K_BACKSUB);
}

// Candidate / model scalars over the ranks: everything is a sum except the line search's |delta|_inf
void Engine::phong_reduce_scal2() {
    if (n_ranks <= 1) return;
    // the max slot is reduced on its own copy, then written back over the (meaningless) sum
    CSLAM_CUDA(cudaMemcpyAsync(d_dmax_tmp.p, d_scal2.p + SC_LS_DMAX, sizeof(double), cudaMemcpyDeviceToDevice, stream));
    comm_allreduce_max(nccl_comm, d_dmax_tmp.p, 1, stream);
    comm_allreduce_sum(nccl_comm, d_scal2.p, SC_COUNT, stream);
    CSLAM_CUDA(cudaMemcpyAsync(d_scal2.p + SC_LS_DMAX, d_dmax_tmp.p, sizeof(double), cudaMemcpyDeviceToDevice, stream));
}

// TrustRegionMinimizer::DoLineSearch for a bounded problem: Armijo search along the step (yp, yg, yv), starting
// from the full step whose scalars are in sc2 (SC_LS_GY = g . y, SC_LS_DMAX = |delta|_inf, SC_MODEL, SC_CAND_COST).
void Engine::phong_line_search(const double* yp, const double* yg, const double* yv, double* sc2) {
    DevView v = view(d_poses.p, d_points.p);
    const PhongSolveView q = phong_solve_view(ph.normals.p, ph.gx.p);
    {
        const double g0 = -sc2[SC_LS_GY], dmax = sc2[SC_LS_DMAX];
        const double model = sc2[SC_MODEL];
        double a_cur = 1.0, f_cur = sc2[SC_CAND_COST];
        double first[SC_COUNT];
        std::memcpy(first, sc2, sizeof(first));
        bool success = true;
        int ls_it = 0;
        const double armijo = opt.line_search_sufficient_function_decrease;
        while (!std::isfinite(f_cur) || f_cur > lm.x_cost + armijo * g0 * a_cur) {
            if (++ls_it >= 20) {
                success = false;
                break;
            }
            const double lo = 1e-3 * a_cur, hi = 0.6 * a_cur;
            double a_new;
            if (!std::isfinite(f_cur)) {
                a_new = std::min(std::max(0.5 * a_cur, lo), hi);
            } else {
                const double c2 = (f_cur - lm.x_cost - g0 * a_cur) / (a_cur * a_cur);
                a_new = c2 > 0.0 ? -g0 / (2.0 * c2) : hi;
                a_new = std::min(std::max(a_new, lo), hi);
            }
            if (a_new * dmax < 1e-9) {
                success = false;
                break;
            }
            a_cur = a_new;
            d_scal2.zero(stream);
            launch_phong_candidate(stream, v, q, 0, n_lm, a_cur, yp, yg, yv, d_poses_cand.p, ph.gx_cand.p,
                                   d_points_cand.p, ph.normals_cand.p, d_scal2.p, rank == 0, ph.max_track);
            phong_reduce_scal2();
            read_scalars(d_scal2.p, sc2, SC_COUNT);
            f_cur = sc2[SC_CAND_COST];
        }
        if (!success && a_cur != 1.0) {
            // the search failed: Ceres keeps the full step
            d_scal2.zero(stream);
            launch_phong_candidate(stream, v, q, 0, n_lm, 1.0, yp, yg, yv, d_poses_cand.p, ph.gx_cand.p,
                                   d_points_cand.p, ph.normals_cand.p, d_scal2.p, rank == 0, ph.max_track);
            phong_reduce_scal2();
            read_scalars(d_scal2.p, sc2, SC_COUNT);
        }
        sc2[SC_MODEL] = model;
        sc2[SC_NONFINITE] = first[SC_NONFINITE];
    }
}

// DOGLEG for the lighting solve (dataset_ba_phong.cpp:88-89): DoglegStrategy::ComputeStep over
// [poses | vertices | shared blocks] and the candidate evaluation (same structure as dogleg_step).
void Engine::phong_dogleg_step(int* lin_iters, bool* valid, double* sc2) {
    DevView v = view(d_poses.p, d_points.p);
    const PhongSolveView q = phong_solve_view(ph.normals.p, ph.gx.p);
    *lin_iters = 0;
    if (!dl.reuse) {
        bool solved = false;
        while (dl.mu < 1.0) {
            if (!lm.have_system) schur_pass();
            double sc1[SC_COUNT];
            read_scalars(d_scal, sc1, SC_COUNT);
            int it = 0;
            bool ok = sc1[SC_INVALID] == 0.0;
            if (ok) phong_linear_solve(&it, &ok);
            if (ok) {
                solved = true;
                break;
            }
            dl.mu *= 10.0;
            lm.have_system = false;
        }
        if (!solved) {
            *valid = false;
            return;
        }
        *lin_iters = 1;
        const LmDiag dg = current_diag();
        prof_begin(CSLAM_K_BACKSUB);
        d_scal2.zero(stream);  // (the back-substitution's model / line-search outputs are not used here)
        launch_phong_backsub(stream, v, q, 0, n_lm, dg, d_yp.p, ph.yg.p, ph.gv.p, ph.yv.p, d_scal2.p, ph.max_track);
        d_dsums.zero(stream);
        launch_dogleg_products(stream, v, 0, 0, dg, nullptr, 0, nullptr, 0, d_gp, d_diag_p.p, d_yp.p, nullptr, nullptr, nullptr,
                               d_dsums.p, rank == 0);
        launch_phong_dogleg_products(stream, v, q, 0, n_lm, dg, d_gp, d_diag_p.p, d_yp.p, ph.gg, ph.diag_g.p, ph.yg.p, ph.gv.p,
                                     ph.yv.p, ph.diag_v.p, ph.sc_v.p, d_dsums.p, ph.max_track, rank == 0);
        prof_end(CSLAM_K_BACKSUB);
        allreduce_small(d_dsums.p, DG_COUNT);
        double sm[DG_COUNT];
        read_scalars(d_dsums.p, sm, DG_COUNT);
        DoglegModel& m = dl.model;
        m.G11 = sm[DG_G11], m.G12 = sm[DG_G12], m.G22 = sm[DG_G22];
        m.JGG = sm[DG_JGG], m.JGY = sm[DG_JGY], m.JYY = sm[DG_JYY], m.JGR = sm[DG_JGR], m.JYR = sm[DG_JYR];
        bool finite = true;
        for (double x : sm) finite = finite && std::isfinite(x);
        if (!finite || !m.prepare(opt.dogleg_type == 1)) {
            *valid = false;
            return;
        }
        dl.reuse = true;
    }
    double c1, c2;
    if (opt.dogleg_type == 1)
        dl.model.subspace(lm.radius, &c1, &c2, &dl.step_norm);
    else
        dl.model.traditional(lm.radius, &c1, &c2, &dl.step_norm);
    prof_begin(CSLAM_K_BACKSUB);
    launch_dogleg_combine(stream, 6ll * n_free, c1, c2, d_gp, d_diag_p.p, d_yp.p, d_Yp.p);
    launch_dogleg_combine(stream, ph.n_g, c1, c2, ph.gg, ph.diag_g.p, ph.yg.p, ph.Yg.p);
    launch_dogleg_combine(stream, 6ll * n_lm, c1, c2, ph.gv.p, ph.diag_v.p, ph.yv.p, ph.Yv.p);
    d_scal2.zero(stream);
    if (bounded) {
        if (rank == 0) {
            launch_dot(stream, d_gp, d_Yp.p, 6ll * n_free, d_scal2.p + SC_LS_GY);
            launch_dot(stream, ph.gg, ph.Yg.p, ph.n_g, d_scal2.p + SC_LS_GY);
            launch_absmax_scaled(stream, d_Yp.p, d_sc_p.p, 6ll * n_free, d_scal2.p + SC_LS_DMAX);
            launch_absmax_scaled(stream, ph.Yg.p, ph.sc_g.p, ph.n_g, d_scal2.p + SC_LS_DMAX);
        }
        launch_dot(stream, ph.gv.p, ph.Yv.p, 6ll * n_lm, d_scal2.p + SC_LS_GY);
        launch_absmax_scaled(stream, ph.Yv.p, ph.sc_v.p, 6ll * n_lm, d_scal2.p + SC_LS_DMAX);
    }
    launch_phong_candidate(stream, v, q, 0, n_lm, 1.0, d_Yp.p, ph.Yg.p, ph.Yv.p, d_poses_cand.p, ph.gx_cand.p, d_points_cand.p,
                           ph.normals_cand.p, d_scal2.p, rank == 0, ph.max_track);
    phong_reduce_scal2();
    read_scalars(d_scal2.p, sc2, SC_COUNT);
    sc2[SC_MODEL] = dl.model.model_cost_change(c1, c2);
    *valid = sc2[SC_NONFINITE] == 0.0 && sc2[SC_MODEL] > 0.0;
    if (bounded && *valid) phong_line_search(d_Yp.p, ph.Yg.p, ph.Yv.p, sc2);
    prof_end(CSLAM_K_BACKSUB);
}

// DoglegStrategy::ComputeStep + the candidate evaluation of the trust-region loop.  Per Jacobian: the
// Gauss-Newton solve (J^T J + mu D^2) y = J^T r through the same Schur path as LM (mu in place of
// 1/radius, raised tenfold while the factorisation fails), the back-substitution for the landmark
// part, then ONE pass for the eight inner products the strategy needs.  Per trial radius (also
// after a rejected step, when everything above is reused): two coefficients on the host, the
// combined step, the candidate and its cost.
void Engine::dogleg_step(int* lin_iters, bool* valid, double* sc2) {
    DevView v = view(d_poses.p, d_points.p);
    *lin_iters = 0;
    if (!dl.reuse) {
        bool solved = false;
        while (dl.mu < 1.0) {
            if (!lm.have_system) schur_pass();
            double sc1[SC_COUNT];
            read_scalars(d_scal, sc1, SC_COUNT);
            int it = 0;
            bool ok = sc1[SC_INVALID] == 0.0;
            if (ok) run_pcg(&it, &ok);
            if (ok) {
                solved = true;
                break;
            }
            dl.mu *= 10.0;
            lm.have_system = false;
        }
        if (!solved) {
            *valid = false;
            return;
        }
        *lin_iters = 1;
        const LmDiag dg = current_diag();
        prof_begin(CSLAM_K_BACKSUB);
        // landmark part of y (the kernel's candidate / model outputs are not used here)
        d_scal2.zero(stream);
        launch_pose_plus(stream, v, d_yp.p, d_poses_cand.p, d_scal2.p, rank == 0);
        launch_backsub(stream, v, 0, n_lm, dg, d_yp.p, d_poses_cand.p, d_points_cand.p, d_yl.p, d_scal2.p);
        d_dsums.zero(stream);
        launch_dogleg_products(stream, v, 0, n_lm, dg, d_suns.p, int(suns.size()), d_priors.p, int(priors.size()), d_gp,
                               d_diag_p.p, d_yp.p, d_gl.p, d_yl.p, d_diag_l.p, d_dsums.p, rank == 0);
        prof_end(CSLAM_K_BACKSUB);
        allreduce_small(d_dsums.p, DG_COUNT);
        double sm[DG_COUNT];
        read_scalars(d_dsums.p, sm, DG_COUNT);
        DoglegModel& m = dl.model;
        m.G11 = sm[DG_G11], m.G12 = sm[DG_G12], m.G22 = sm[DG_G22];
        m.JGG = sm[DG_JGG], m.JGY = sm[DG_JGY], m.JYY = sm[DG_JYY], m.JGR = sm[DG_JGR], m.JYR = sm[DG_JYR];
        bool finite = true;
        for (double x : sm) finite = finite && std::isfinite(x);
        if (!finite || !m.prepare(opt.dogleg_type == 1)) {
            *valid = false;
            return;
        }
        dl.reuse = true;
    }
    double c1, c2;
    if (opt.dogleg_type == 1)
        dl.model.subspace(lm.radius, &c1, &c2, &dl.step_norm);
    else
        dl.model.traditional(lm.radius, &c1, &c2, &dl.step_norm);
    prof_begin(CSLAM_K_BACKSUB);
    launch_dogleg_combine(stream, 6ll * n_free, c1, c2, d_gp, d_diag_p.p, d_yp.p, d_Yp.p);
    launch_dogleg_combine(stream, 3ll * n_lm, c1, c2, d_gl.p, d_diag_l.p, d_yl.p, d_Yl.p);
    d_scal2.zero(stream);
    launch_pose_plus(stream, v, d_Yp.p, d_poses_cand.p, d_scal2.p, rank == 0);
    launch_points_apply(stream, v, 0, n_lm, d_Yl.p, d_poses_cand.p, d_points_cand.p, d_scal2.p);
    if (rank == 0)
        launch_camonly_step(stream, v, d_suns.p, int(suns.size()), d_priors.p, int(priors.size()), d_Yp.p, d_poses_cand.p,
                            d_scal2.p);
    prof_end(CSLAM_K_BACKSUB);
    allreduce_small(d_scal2.p, SC_COUNT);
    read_scalars(d_scal2.p, sc2, SC_COUNT);
    sc2[SC_MODEL] = dl.model.model_cost_change(c1, c2);
    *valid = sc2[SC_NONFINITE] == 0.0 && sc2[SC_MODEL] > 0.0;
}

// -------------------------------------------------------------------------------------------------
// LM loop
// -------------------------------------------------------------------------------------------------
void Engine::lm_begin() {
    if (!uploaded) throw std::invalid_argument("lm_begin before upload");
    lm = Lm();
    dl = Dogleg();
    if (dogleg()) {
        if (ph.active) {
            ph.diag_v.alloc(6 * size_t(std::max(n_lm, 1)), stream);
            ph.sc_v.alloc(6 * size_t(std::max(n_lm, 1)), stream);
            ph.Yv.alloc(6 * size_t(std::max(n_lm, 1)), stream);
            ph.Yg.alloc(size_t(std::max(ph.n_g, 1)), stream);
        }
        d_diag_l.alloc(3 * size_t(std::max(n_lm, 1)), stream);
        d_Yl.alloc(3 * size_t(std::max(n_lm, 1)), stream);
        d_Yp.alloc(6 * size_t(std::max(n_free, 1)), stream);
        d_dsums.alloc(DG_COUNT, stream);
    }
    log.clear();
    // unit scaling for the initial pass
    launch_fill(stream, d_sc_p.p, d_sc_p.n, 1.0);
    launch_fill(stream, d_sc_l.p, d_sc_l.n, 1.0);
    if (ph.active) {
        launch_fill(stream, ph.sc_n.p, ph.sc_n.n, 1.0);
        launch_fill(stream, ph.sc_g.p, ph.sc_g.n, 1.0);
        if (bounded) {
            // IterationZero of a bounded problem: x <- Plus(x, 0), i.e. projected onto the box (and
            // unit vectors re-normalised); poses and positions are unchanged by a zero step
            launch_phong_project_initial(stream, phong_solve_view(ph.normals.p, ph.gx.p), n_lm, ph.normals.p, ph.gx_cand.p,
                                         ph.zero_g.p);
            CSLAM_CUDA(cudaMemcpyAsync(ph.gx.p, ph.gx_cand.p, ph.gx.bytes(), cudaMemcpyDeviceToDevice, stream));
            CSLAM_CUDA(cudaMemcpyAsync(ph.gx_best.p, ph.gx.p, ph.gx.bytes(), cudaMemcpyDeviceToDevice, stream));
            CSLAM_CUDA(cudaMemcpyAsync(ph.normals_best.p, ph.normals.p, ph.normals.bytes(), cudaMemcpyDeviceToDevice, stream));
        }
    }
    DevView v = view(d_poses.p, d_points.p);
    prof_begin(CSLAM_K_COLNORM);
    d_red.zero(stream);
    if (ph.active) {
        const LmDiag unit{1.0, opt.min_lm_diagonal, opt.max_lm_diagonal};
        launch_phong_build(stream, v, phong_solve_view(ph.normals.p, ph.gx.p), 0, n_lm, unit, phong_system(), false, ph.max_track);
    } else {
        launch_colnorm(stream, v, 0, n_lm, d_Bdiag, d_cn_l.p, d_gp, d_gl.p, d_scal);
        if (rank == 0)
            launch_camonly_build(stream, v, d_suns.p, int(suns.size()), d_priors.p, int(priors.size()), d_Bdiag, d_bp, d_gp,
                                 d_scal);
    }
    prof_end(CSLAM_K_COLNORM);
    allreduce_system();
    gradient_norm_pass();  // sc == 1: gp, gl are the unscaled gradient
    double loc[SC_COUNT];
    read_scalars(d_scal, loc, SC_COUNT);
    launch_jacobi_scale_cams(stream, d_Bdiag, d_sc_p.p, n_free, opt.jacobi_scaling);
    launch_jacobi_scale(stream, d_cn_l.p, d_sc_l.p, 3ll * n_lm, opt.jacobi_scaling);
    if (ph.active) {
        launch_jacobi_scale(stream, ph.cn_n.p, ph.sc_n.p, 3ll * n_lm, opt.jacobi_scaling);
        launch_jacobi_scale(stream, ph.hg, ph.sc_g.p, ph.n_g, opt.jacobi_scaling);
    }
    if (!std::isfinite(loc[SC_COST])) {
        lm.finished = true;
        lm.termination_type = 2;
        lm.termination_reason = 8;
        begun = true;
        throw std::domain_error("non-finite cost at the initial point");
    }
    lm.x_cost = lm.initial_cost = lm.minimum_cost = loc[SC_COST];
    lm.fixed_cost = ph.active ? loc[SC_FIXED] : 0.0;
    lm.se_minimum = lm.se_current = lm.se_reference = lm.se_candidate = lm.x_cost;
    lm.radius = opt.initial_trust_region_radius;
    lm.decrease_factor = 2.0;
    lm.have_system = false;
    LmRow row{};
    row.v[0] = 0;
    row.v[1] = lm.x_cost;
    row.v[3] = lm.gradient_max_norm;
    row.v[6] = lm.radius;
    log.push_back(row);
    begun = true;
}

void Engine::fill_summary(cslam_summary* s) const {
    if (!s) return;
    std::memset(s, 0, sizeof(*s));
    s->initial_cost = lm.initial_cost + lm.fixed_cost;
    s->final_cost = lm.minimum_cost + lm.fixed_cost;
    s->num_iterations = lm.iteration;
    s->num_successful_steps = lm.num_successful;
    s->num_unsuccessful_steps = lm.num_unsuccessful;
    s->termination_type = lm.termination_type;
    s->termination_reason = lm.termination_reason;
    s->final_radius = lm.radius;
    s->total_linear_iterations = lm.total_linear;
    s->device_ms = lm.device_ms;
}

void Engine::lm_iterate(int n, bool ignore_convergence, cslam_summary* s) {
    if (!begun) throw std::invalid_argument("lm_iterate before lm_begin");
    CSLAM_CUDA(cudaEventRecord(ev_a, stream));
    const int max_nonmono = opt.use_nonmonotonic_steps ? opt.max_consecutive_nonmonotonic_steps : 0;
    auto finish = [&](int type, int reason) {
        lm.termination_type = type;
        lm.termination_reason = reason;
        lm.finished = !ignore_convergence || type == 2;
    };
    for (int it = 0; it < n && !lm.finished; ++it) {
        // ---- FinalizeIterationAndCheckIfMinimizerCanContinue ----
        if (lm.iteration > 0) {
            if (lm.step_ok_prev) {
                ++lm.num_successful;
                if (lm.x_cost < lm.minimum_cost) {
                    lm.minimum_cost = lm.x_cost;
                    CSLAM_CUDA(cudaMemcpyAsync(d_poses_best.p, d_poses.p, d_poses.bytes(), cudaMemcpyDeviceToDevice, stream));
                    if (n_lm)
                        CSLAM_CUDA(cudaMemcpyAsync(d_points_best.p, d_points.p, d_points.bytes(), cudaMemcpyDeviceToDevice, stream));
                    copy_phong_best();
                }
            } else {
                ++lm.num_unsuccessful;
            }
            lm.step_ok_prev = false;
        }
        if (lm.iteration >= opt.max_num_iterations && !ignore_convergence) {
            finish(1, 4);
            break;
        }
        if (!lm.have_system) schur_pass();
        if (!lm.grad_fresh) {
            gradient_norm_pass();
            if (!log.empty()) log.back().v[3] = lm.gradient_max_norm;
        }
        if (lm.gradient_max_norm <= opt.gradient_tolerance) {
            finish(0, 1);
            if (lm.finished) break;
        }
        if (lm.radius < opt.min_trust_region_radius) {
            finish(0, 5);
            if (lm.finished) break;
        }
        ++lm.iteration;
        LmRow row{};
        row.v[0] = lm.iteration;
        // ---- LevenbergMarquardtStrategy::ComputeStep ----
        double sc1[SC_COUNT];
        read_scalars(d_scal, sc1, SC_COUNT);
        int lin_iters = 0;
        bool lin_ok = true;
        bool valid = sc1[SC_INVALID] == 0.0;
        const bool use_dogleg = dogleg();
        if (valid && !use_dogleg) {
            if (ph.active)
                phong_linear_solve(&lin_iters, &lin_ok);
            else
                run_pcg(&lin_iters, &lin_ok);
        }
        valid = valid && lin_ok;
        row.v[7] = lin_iters;
        lm.total_linear += lin_iters;
        double sc2[SC_COUNT] = {0};
        const LmDiag dg = current_diag();
        if (use_dogleg && ph.active) {
            if (valid) phong_dogleg_step(&lin_iters, &valid, sc2);
            row.v[7] = lin_iters;
            lm.total_linear += lin_iters;
        } else if (use_dogleg) {
            dogleg_step(&lin_iters, &valid, sc2);
            row.v[7] = lin_iters;
            lm.total_linear += lin_iters;
        } else if (valid && ph.active) {
            phong_step(dg, sc2);
            if (sc2[SC_NONFINITE] != 0.0) valid = false;
            if (!(sc2[SC_MODEL] > 0.0)) valid = false;
        } else if (valid) {
            DevView v = view(d_poses.p, d_points.p);
            prof_begin(CSLAM_K_BACKSUB);
            d_scal2.zero(stream);
            launch_pose_plus(stream, v, d_yp.p, d_poses_cand.p, d_scal2.p, rank == 0);
            launch_backsub(stream, v, 0, n_lm, dg, d_yp.p, d_poses_cand.p, d_points_cand.p, d_yl.p, d_scal2.p);
            if (rank == 0)
                launch_camonly_step(stream, v, d_suns.p, int(suns.size()), d_priors.p, int(priors.size()), d_yp.p,
                                    d_poses_cand.p, d_scal2.p);
            prof_end(CSLAM_K_BACKSUB);
            allreduce_small(d_scal2.p, SC_COUNT);
            read_scalars(d_scal2.p, sc2, SC_COUNT);
            if (sc2[SC_NONFINITE] != 0.0) valid = false;
            if (!(sc2[SC_MODEL] > 0.0)) valid = false;
        }
        if (!valid) {
            row.v[1] = lm.x_cost;
            row.v[3] = lm.gradient_max_norm;
            if (++lm.invalid_steps >= opt.max_num_consecutive_invalid_steps) {
                row.v[6] = lm.radius;
                log.push_back(row);
                lm.termination_type = 2;
                lm.termination_reason = 6;
                lm.finished = true;
                break;
            }
            if (use_dogleg) {
                dl.mu *= 10.0;  // DoglegStrategy::StepIsInvalid
                dl.reuse = false;
            } else {
                lm.radius = lm.radius / lm.decrease_factor;
                lm.decrease_factor *= 2.0;
            }
            lm.have_system = false;
            row.v[6] = lm.radius;
            log.push_back(row);
            continue;
        }
        lm.invalid_steps = 0;
        row.v[8] = 1;
        const double model_cost_change = sc2[SC_MODEL];
        double cand_cost = sc2[SC_CAND_COST];
        if (!std::isfinite(cand_cost)) cand_cost = std::numeric_limits<double>::max();
        row.v[4] = std::sqrt(sc2[SC_STEP_NORM2]);
        row.v[2] = lm.x_cost - cand_cost;
        if (!ignore_convergence) {
            if (row.v[4] <= opt.parameter_tolerance * (lm.x_norm + opt.parameter_tolerance)) {
                row.v[1] = lm.x_cost;
                row.v[3] = lm.gradient_max_norm;
                row.v[6] = lm.radius;
                log.push_back(row);
                finish(0, 2);
                break;
            }
            if (std::fabs(row.v[2]) <= opt.function_tolerance * lm.x_cost) {
                row.v[1] = lm.x_cost;
                row.v[3] = lm.gradient_max_norm;
                row.v[6] = lm.radius;
                log.push_back(row);
                finish(0, 3);
                break;
            }
        }
        const double rel = (lm.se_current - cand_cost) / model_cost_change;
        const double hist = (lm.se_reference - cand_cost) / (lm.se_acc_ref + model_cost_change);
        row.v[5] = std::max(rel, hist);
        if (row.v[5] > opt.min_relative_decrease) {
            std::swap(d_poses.p, d_poses_cand.p);
            std::swap(d_points.p, d_points_cand.p);
            if (ph.active) {
                std::swap(ph.normals.p, ph.normals_cand.p);
                std::swap(ph.gx.p, ph.gx_cand.p);
            }
            lm.x_cost = cand_cost;
            lm.x_norm = std::sqrt(sc2[SC_XNORM2]);
            lm.step_ok_prev = true;
            lm.have_system = false;
            lm.grad_fresh = false;
            row.v[9] = 1;
            if (use_dogleg) {
                // DoglegStrategy::StepAccepted
                if (row.v[5] < 0.25) lm.radius *= 0.5;
                if (row.v[5] > 0.75) {
                    lm.radius = std::max(lm.radius, 3.0 * dl.step_norm);
                    lm.radius = std::min(lm.radius, opt.max_trust_region_radius);
                }
                dl.mu = std::max(1e-8, 2.0 * dl.mu / 10.0);
                dl.reuse = false;
            } else {
                lm.radius = lm.radius / std::max(1.0 / 3.0, 1.0 - std::pow(2.0 * row.v[5] - 1.0, 3));
                lm.radius = std::min(opt.max_trust_region_radius, lm.radius);
                lm.decrease_factor = 2.0;
            }
            lm.se_current = cand_cost;
            lm.se_acc_cand += model_cost_change;
            lm.se_acc_ref += model_cost_change;
            if (lm.se_current < lm.se_minimum) {
                lm.se_minimum = lm.se_current;
                lm.se_nonmono = 0;
                lm.se_candidate = lm.se_current;
                lm.se_acc_cand = 0;
            } else {
                ++lm.se_nonmono;
                if (lm.se_current > lm.se_candidate) {
                    lm.se_candidate = lm.se_current;
                    lm.se_acc_cand = 0;
                }
            }
            if (lm.se_nonmono == max_nonmono) {
                lm.se_reference = lm.se_candidate;
                lm.se_acc_ref = lm.se_acc_cand;
            }
        } else if (use_dogleg) {
            lm.radius *= 0.5;  // DoglegStrategy::StepRejected: the Gauss-Newton and gradient vectors stay valid
            dl.reuse = true;
        } else {
            lm.radius = lm.radius / lm.decrease_factor;
            lm.decrease_factor *= 2.0;
            lm.have_system = false;
        }
        row.v[1] = lm.x_cost;
        row.v[3] = lm.gradient_max_norm;
        row.v[6] = lm.radius;
        log.push_back(row);
    }
    // bookkeeping for a step accepted in the last iteration of this call
    if (lm.step_ok_prev && (lm.finished || true)) {
        if (lm.x_cost < lm.minimum_cost) {
            lm.minimum_cost = lm.x_cost;
            CSLAM_CUDA(cudaMemcpyAsync(d_poses_best.p, d_poses.p, d_poses.bytes(), cudaMemcpyDeviceToDevice, stream));
            if (n_lm)
                CSLAM_CUDA(cudaMemcpyAsync(d_points_best.p, d_points.p, d_points.bytes(), cudaMemcpyDeviceToDevice, stream));
            copy_phong_best();
        }
    }
    CSLAM_CUDA(cudaEventRecord(ev_b, stream));
    CSLAM_CUDA(cudaEventSynchronize(ev_b));
    float ms = 0;
    CSLAM_CUDA(cudaEventElapsedTime(&ms, ev_a, ev_b));
    lm.device_ms += ms;
    fill_summary(s);
}

// -------------------------------------------------------------------------------------------------
// ceres::Problem::Evaluate at the caller's current parameter values (materialised r, J)
// -------------------------------------------------------------------------------------------------
void Engine::ensure_user_copy() {
    ensure_device(this, &stream, &own_stream, &ev_a, &ev_b, &ev_c, &ev_d, &h_pinned);
    if (user_copy_ready) return;
    if (cam_free_h.size() != n_poses) build_structure();
    const size_t n = n_st;
    std::vector<double> u(n), v(n), d(n);
    for (size_t i = 0; i < n; ++i) {
        u[i] = st_uvd[3 * i];
        v[i] = st_uvd[3 * i + 1];
        d[i] = st_uvd[3 * i + 2];
    }
    const size_t tiles = (n + 127) / 128;
    std::vector<int> tlo(std::max<size_t>(tiles, 1), 0), tn(std::max<size_t>(tiles, 1), 0);
    for (size_t t = 0; t < tiles; ++t) {
        uint32_t lo = 0xffffffffu, hi = 0;
        for (size_t i = 128 * t; i < std::min(n, 128 * (t + 1)); ++i) {
            lo = std::min(lo, st_cam[i]);
            hi = std::max(hi, st_cam[i]);
        }
        tlo[t] = int(lo);
        tn[t] = (hi - lo + 1 <= 16) ? int(hi - lo + 1) : 0;  // 0: gather poses from global memory
    }
    d_u_cam.upload(st_cam, n, stream);
    d_u_pt.upload(st_pt, n, stream);
    d_u_u.upload(u, stream);
    d_u_v.upload(v, stream);
    d_u_d.upload(d, stream);
    if (st_W && n)
        d_u_W.upload(st_W, st_W_per_obs ? 9 * n : 9, stream);
    else
        d_u_W.alloc(9);
    d_u_tile_lo.upload(tlo, stream);
    d_u_tile_n.upload(tn, stream);
    d_cam_free.upload(cam_free_h, stream);
    d_o_r.alloc(3 * n);
    d_o_Jc.alloc(18 * n);
    d_o_Jp.alloc(9 * n);
    if (!d_scal2.p) d_scal2.alloc(SC_COUNT);
    if (!suns.empty()) d_suns.upload(suns, stream);
    if (!priors.empty()) d_priors.upload(priors, stream);
    CSLAM_CUDA(cudaStreamSynchronize(stream));
    user_copy_ready = true;
}

void Engine::evaluate(int apply_loss, double* cost, double* r_st, double* Jc_st, double* Jp_st, double* r_sun,
                      double* J_sun, double* r_pr, double* J_pr) {
    ensure_user_copy();
    // parameter values as they are in the caller's arrays right now
    d_poses_cand.upload(h_poses, 12 * size_t(n_poses), stream);
    d_u_points.upload(h_points, 3 * size_t(n_points), stream);
    d_scal2.zero(stream);
    launch_resjac(stream, cam, (long long)n_st, d_u_cam.p, d_u_pt.p, d_u_u.p, d_u_v.p, d_u_d.p, d_u_W.p, st_W_per_obs,
                  d_poses_cand.p, d_u_points.p, d_cam_free.p, d_u_tile_lo.p, d_u_tile_n.p, d_o_r.p, d_o_Jc.p, d_o_Jp.p,
                  d_scal2.p);
    const int ns = int(suns.size()), np = int(priors.size());
    DBuf<double> o_rs, o_Js, o_rp, o_Jp2;
    if (ns + np > 0) {
        o_rs.alloc(2 * size_t(std::max(ns, 1)));
        o_Js.alloc(12 * size_t(std::max(ns, 1)));
        o_rp.alloc(6 * size_t(std::max(np, 1)));
        o_Jp2.alloc(36 * size_t(std::max(np, 1)));
        DevView v{};
        v.cam = cam;
        v.n_cams = int(n_poses);
        v.poses = d_poses_cand.p;
        v.cam_free = d_cam_free.p;
        launch_camonly_eval(stream, v, d_suns.p, ns, d_priors.p, np, apply_loss, o_rs.p, o_Js.p, o_rp.p, o_Jp2.p,
                            d_scal2.p);
    }
    auto d2h = [&](double* dst, const double* src, size_t count) {
        if (dst && count) CSLAM_CUDA(cudaMemcpyAsync(dst, src, count * sizeof(double), cudaMemcpyDeviceToHost, stream));
    };
    d2h(r_st, d_o_r.p, 3 * size_t(n_st));
    d2h(Jc_st, d_o_Jc.p, 18 * size_t(n_st));
    d2h(Jp_st, d_o_Jp.p, 9 * size_t(n_st));
    d2h(r_sun, o_rs.p, 2 * size_t(ns));
    d2h(J_sun, o_Js.p, 12 * size_t(ns));
    d2h(r_pr, o_rp.p, 6 * size_t(np));
    d2h(J_pr, o_Jp2.p, 36 * size_t(np));
    double c = 0;
    read_scalars(d_scal2.p, &c, 1);
    if (cost) *cost = c;
}

// -------------------------------------------------------------------------------------------------
// lighting blocks: ceres::Problem::Evaluate over the intensity and normal residual blocks
// -------------------------------------------------------------------------------------------------
void Engine::ensure_phong() {
    ensure_device(this, &stream, &own_stream, &ev_a, &ev_b, &ev_c, &ev_d, &h_pinned);
    if (!h_poses || !h_points) throw std::invalid_argument("poses / points not set");
    if (!h_normals || !h_textures || !h_material_id || !h_phong || !h_light)
        throw std::invalid_argument("vertices / materials / light not set");
    if (n_vertices != n_points) throw std::invalid_argument("one normal / texture / material id per point expected");
    for (uint32_t j = 0; j < n_vertices; ++j)
        if (h_material_id[j] >= n_materials) throw std::invalid_argument("material id out of range");
    if (!phong_ready) {
        for (uint64_t i = 0; i < n_ph; ++i)
            if (ph_cam[i] >= n_poses || ph_vertex[i] >= n_points) throw std::invalid_argument("lighting block index out of range");
        std::vector<int> cf(n_poses);
        for (uint32_t k = 0; k < n_poses; ++k) cf[k] = pose_const[k] ? -1 : int(k);
        d_ph_cam_free.upload(cf, stream);
        d_ph_cam.upload(ph_cam, n_ph, stream);
        d_ph_vertex.upload(ph_vertex, n_ph, stream);
        d_ph_int.upload(ph_intensity, n_ph, stream);
        d_ph_nobs.upload(ph_normal_obs, 3 * n_ph, stream);
        d_ph_mat.upload(h_material_id, n_vertices, stream);
        d_ph_W.upload(ph_W_normal, 9, stream);
        d_ph_rI.alloc(n_ph, stream);
        d_ph_JI.alloc(19 * n_ph, stream);
        d_ph_rN.alloc(3 * n_ph, stream);
        d_ph_JNc.alloc(18 * n_ph, stream);
        d_ph_JNn.alloc(9 * n_ph, stream);
        if (!d_scal2.p) d_scal2.alloc(SC_COUNT, stream);
        phong_ready = true;
    }
    // parameter values as they are in the caller's arrays right now
    d_ph_poses.upload(h_poses, 12 * size_t(n_poses), stream);
    d_ph_points.upload(h_points, 3 * size_t(n_points), stream);
    d_ph_normals.upload(h_normals, 3 * size_t(n_vertices), stream);
    if (h_tex_shared) {
        std::vector<double> tex(n_vertices);
        for (uint32_t j = 0; j < n_vertices; ++j) tex[j] = h_tex_shared[h_texture_id[j]];
        d_ph_tex.upload(tex, stream);
        CSLAM_CUDA(cudaStreamSynchronize(stream));
    } else {
        d_ph_tex.upload(h_textures, n_vertices, stream);
    }
    d_ph_phong.upload(h_phong, 3 * size_t(n_materials), stream);
    d_ph_light.upload(h_light, 3, stream);
}

PhongView Engine::phong_view() {
    PhongView v;
    v.n = (long long)n_ph;
    v.cam = d_ph_cam.p;
    v.vertex = d_ph_vertex.p;
    v.material_id = d_ph_mat.p;
    v.intensity = d_ph_int.p;
    v.normal_obs = d_ph_nobs.p;
    v.poses = d_ph_poses.p;
    v.points = d_ph_points.p;
    v.normals = d_ph_normals.p;
    v.texture = d_ph_tex.p;
    v.phong = d_ph_phong.p;
    v.light = d_ph_light.p;
    v.W_normal = d_ph_W.p;
    v.cam_free = d_ph_cam_free.p;
    v.int_stiffness = ph_int_stiffness;
    v.directional = light_directional;
    return v;
}

void Engine::evaluate_phong(double* cost, double* r_int, double* J_int, double* r_n, double* Jc_n, double* Jn_n) {
    ensure_phong();
    d_scal2.zero(stream);
    launch_phong_eval(stream, phong_view(), d_ph_rI.p, d_ph_JI.p, d_ph_rN.p, d_ph_JNc.p, d_ph_JNn.p, d_scal2.p);
    auto d2h = [&](double* dst, const double* src, size_t count) {
        if (dst && count) CSLAM_CUDA(cudaMemcpyAsync(dst, src, count * sizeof(double), cudaMemcpyDeviceToHost, stream));
    };
    d2h(r_int, d_ph_rI.p, n_ph);
    d2h(J_int, d_ph_JI.p, 19 * n_ph);
    d2h(r_n, d_ph_rN.p, 3 * n_ph);
    d2h(Jc_n, d_ph_JNc.p, 18 * n_ph);
    d2h(Jn_n, d_ph_JNn.p, 9 * n_ph);
    double c = 0;
    read_scalars(d_scal2.p, &c, 1);
    if (cost) *cost = c;
}

double Engine::time_phong(int reps) {
    ensure_phong();
    d_scal2.zero(stream);
    const PhongView v = phong_view();
    auto go = [&]() { launch_phong_eval(stream, v, d_ph_rI.p, d_ph_JI.p, d_ph_rN.p, d_ph_JNc.p, d_ph_JNn.p, d_scal2.p); };
    for (int i = 0; i < 3; ++i) go();
    CSLAM_CUDA(cudaEventRecord(ev_a, stream));
    for (int i = 0; i < reps; ++i) go();
    CSLAM_CUDA(cudaEventRecord(ev_b, stream));
    CSLAM_CUDA(cudaEventSynchronize(ev_b));
    float ms = 0;
    CSLAM_CUDA(cudaEventElapsedTime(&ms, ev_a, ev_b));
    return double(ms) / reps;
}

// -------------------------------------------------------------------------------------------------
// marginal covariance of one pose block (dataset_vo_sun.cpp:159-183)
// -------------------------------------------------------------------------------------------------
void Engine::covariance_block(uint32_t cam, double* cov36) {
    if (cam >= n_poses) throw std::invalid_argument("covariance: pose index out of range");
    if (n_ranks > 1) throw std::invalid_argument("covariance: single-GPU problems only");
    if (window_eligible(true)) {
        // a sliding window (<= 8 poses): one launch of the one-CTA covariance kernel instead of upload + structure
        // analysis + six host-driven solves (dataset_vo_sun asks for this block after every window)
        Engine* self = this;
        solve_window_batch(&self, 1, nullptr, int(cam), cov36);
        return;
    }
    // the reduced system at the caller's current values, undamped, exact solve
    const cslam_options keep = opt;
    opt.linear_solver = 0;
    try {
        upload();
        lm_begin();
    } catch (...) {
        opt = keep;
        throw;
    }
    opt = keep;
    const int f = cam_free_h[cam];
    if (f < 0) throw std::invalid_argument("covariance: the pose block is constant");
    lm.radius = std::numeric_limits<double>::infinity();  // D = 0
    dl.mu = 0.0;                                          // (a DOGLEG problem damps with mu, not 1 / radius)
    schur_pass();
    double sc1[SC_COUNT];
    read_scalars(d_scal, sc1, SC_COUNT);
    if (sc1[SC_INVALID] != 0.0) throw std::domain_error("covariance: a landmark block is rank deficient");
    double sp[6];
    read_scalars(d_sc_p.p + 6ll * f, sp, 6);
    // S^-1 e_c for the six columns of the block; un-scale: J = J_s diag(s)^-1  =>  cov = diag(s) cov_s diag(s)
    for (int c = 0; c < 6; ++c) {
        CSLAM_CUDA(cudaMemsetAsync(d_bp, 0, 6 * size_t(n_free) * sizeof(double), stream));
        const double one = 1.0;
        CSLAM_CUDA(cudaMemcpyAsync(d_bp + 6ll * f + c, &one, sizeof(double), cudaMemcpyHostToDevice, stream));
        CSLAM_CUDA(cudaStreamSynchronize(stream));
        int iters = 0;
        bool ok = true;
        const cslam_options keep2 = opt;
        opt.linear_solver = 0;
        run_pcg(&iters, &ok);
        opt = keep2;
        if (!ok) throw std::domain_error("covariance: the reduced camera system is not positive definite");
        double col[6];
        read_scalars(d_yp.p + 6ll * f, col, 6);
        for (int r = 0; r < 6; ++r) cov36[6 * r + c] = sp[r] * col[r] * sp[c];
    }
    lm.have_system = false;
    uploaded = begun = false;  // the device state no longer mirrors a solve in progress
}

double Engine::time_resjac(int reps) {
    ensure_user_copy();
    d_poses_cand.upload(h_poses, 12 * size_t(n_poses), stream);
    d_u_points.upload(h_points, 3 * size_t(n_points), stream);
    d_scal2.zero(stream);
    auto go = [&]() {
        launch_resjac(stream, cam, (long long)n_st, d_u_cam.p, d_u_pt.p, d_u_u.p, d_u_v.p, d_u_d.p, d_u_W.p,
                      st_W_per_obs, d_poses_cand.p, d_u_points.p, d_cam_free.p, d_u_tile_lo.p, d_u_tile_n.p, d_o_r.p,
                      d_o_Jc.p, d_o_Jp.p, d_scal2.p);
    };
    for (int i = 0; i < 3; ++i) go();
    CSLAM_CUDA(cudaEventRecord(ev_a, stream));
    for (int i = 0; i < reps; ++i) go();
    CSLAM_CUDA(cudaEventRecord(ev_b, stream));
    CSLAM_CUDA(cudaEventSynchronize(ev_b));
    float ms = 0;
    CSLAM_CUDA(cudaEventElapsedTime(&ms, ev_a, ev_b));
    return double(ms) / reps;
}

double Engine::time_schur(int reps) {
    if (!begun) throw std::invalid_argument("time_schur before lm_begin");
    const LmDiag dg{1.0 / lm.radius, opt.min_lm_diagonal, opt.max_lm_diagonal};
    DevView v = view(d_poses.p, d_points.p);
    auto go = [&]() {
        d_red.zero(stream);
        launch_schur(v, dg);
    };
    for (int i = 0; i < 2; ++i) go();
    float total = 0;
    for (int i = 0; i < reps; ++i) {
        d_red.zero(stream);
        CSLAM_CUDA(cudaEventRecord(ev_a, stream));
        launch_schur(v, dg);
        CSLAM_CUDA(cudaEventRecord(ev_b, stream));
        CSLAM_CUDA(cudaEventSynchronize(ev_b));
        float ms = 0;
        CSLAM_CUDA(cudaEventElapsedTime(&ms, ev_a, ev_b));
        total += ms;
    }
    lm.have_system = false;
    return double(total) / reps;
}

void Engine::get_reduced_sizes(int* nf, int* nnz) const {
    *nf = n_free;
    *nnz = nnzU;
}

void Engine::get_reduced_system(int* rowptr, int* col, double* values, double* rhs, int* ids) {
    if (!begun) throw std::invalid_argument("no reduced system yet");
    if (!lm.have_system) schur_pass();
    std::memcpy(rowptr, s_rowptr_h.data(), s_rowptr_h.size() * sizeof(int));
    std::memcpy(col, s_col_h.data(), s_col_h.size() * sizeof(int));
    std::memcpy(ids, free_cams_h.data(), free_cams_h.size() * sizeof(int));
    CSLAM_CUDA(cudaMemcpyAsync(values, d_S, 36 * size_t(nnzU) * sizeof(double), cudaMemcpyDeviceToHost, stream));
    CSLAM_CUDA(cudaMemcpyAsync(rhs, d_bp, 6 * size_t(n_free) * sizeof(double), cudaMemcpyDeviceToHost, stream));
    CSLAM_CUDA(cudaStreamSynchronize(stream));
}

unsigned long long Engine::layout_hash() const {
    unsigned long long lh = 1469598103934665603ull;
    auto mix = [&](const auto& vec) {
        for (auto v : vec) lh = (lh ^ (unsigned long long)(unsigned)v) * 1099511628211ull;
        lh = (lh ^ 0xffull) * 1099511628211ull;
    };
    mix(cam_free_h); mix(free_cams_h); mix(lm_user_h); mix(lm_base_h); mix(lm_stride_h); mix(lm_cnt_h);
    for (size_t i = 0; i < obs_user_n; ++i) lh = (lh ^ (unsigned long long)obs_user_h[i]) * 1099511628211ull;
    lh = (lh ^ 0xffull) * 1099511628211ull;
    mix(item_group_h); mix(item_j0_h); mix(item_n_h); mix(g_L_h); mix(g_G_h); mix(g_lm0_h);
    mix(g_obs0_h); mix(g_off_h); mix(g_cams_h); mix(g_blk_off_h); mix(g_blk_h); mix(g_map_off_h);
    lh = (lh ^ (unsigned long long)n_lm_grouped) * 1099511628211ull;
    lh = (lh ^ (unsigned long long)n_items_small) * 1099511628211ull;
    lh = (lh ^ (unsigned long long)max_group_L) * 1099511628211ull;
    return lh;
}

void Engine::analyze(int n_ranks_, int rank_, cslam_structure_info* out) {
    const int keep_n = n_ranks, keep_r = rank;
    n_ranks = n_ranks_;
    rank = rank_;
    try {
        build_structure();
    } catch (...) {
        n_ranks = keep_n;
        rank = keep_r;
        throw;
    }
    n_ranks = keep_n;
    rank = keep_r;
    uploaded = begun = false;
    out->n_free_cams = n_free;
    out->n_landmarks = n_lm;
    out->n_observations = n_obs_true;
    out->nnz_blocks = nnzU;
    unsigned long long h = 1469598103934665603ull;
    for (int v : s_rowptr_h) h = (h ^ (unsigned long long)(unsigned)v) * 1099511628211ull;
    for (int v : s_col_h) h = (h ^ (unsigned long long)(unsigned)v) * 1099511628211ull;
    out->pattern_hash = h;
    out->n_groups = int(g_L_h.size());
    out->n_grouped_landmarks = n_lm_grouped;
    out->n_work_items = int(item_group_h.size());
    unsigned long long s = 0;
    for (uint32_t id : lm_user_h) s += id;
    out->landmark_id_sum = s;
    const unsigned long long lh = layout_hash();
    out->layout_hash = lh;
}

void Engine::set_window_summary(const cslam_summary& s) {
    lm = Lm();
    lm.initial_cost = s.initial_cost;
    lm.minimum_cost = s.final_cost;
    lm.iteration = s.num_iterations;
    lm.num_successful = s.num_successful_steps;
    lm.num_unsuccessful = s.num_unsuccessful_steps;
    lm.termination_type = s.termination_type;
    lm.termination_reason = s.termination_reason;
    lm.radius = s.final_radius;
    lm.total_linear = s.total_linear_iterations;
    lm.device_ms = s.device_ms;
    lm.finished = true;
    uploaded = begun = false;
}

}  // namespace cslam
