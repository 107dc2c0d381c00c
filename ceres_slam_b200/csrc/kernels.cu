// sm_100a kernels of the cslam_b200 bundle-adjustment back end (generic paths).
//
//   resjac_kernel        K1  materialised residual + Jacobian (ceres::Problem::Evaluate)
//   colnorm_kernel           squared column norms + gradient + cost at the initial point
//   schur_generic_kernel K2  fused residual/Jacobian + Schur elimination, warp per landmark
//   camonly_*                sun-sensor / pose-prior blocks
//   finalize_kernel          LM diagonal on the camera blocks + block-Jacobi inverse
//   (K3a, the persistent PCG kernel, lives in kernels_pcg.cu; the grouped K2 in kernels_grouped.cu)
//   pose_plus / backsub  K4  Plus, back-substitution, model cost change, candidate cost
//
// Everything is FP64, no fast-math.  Reference semantics: SURVEY.md App. A / App. B.
#include <atomic>

#include "kernels.cuh"

namespace cslam {

std::atomic<unsigned long long> g_kernel_launches{0};

namespace {

constexpr int kSMs = 148;
#define CSLAM_LAUNCHED(n) g_kernel_launches.fetch_add((n), std::memory_order_relaxed)

__device__ __forceinline__ const double* obs_W_ptr(const DevView& v, long long e) {
    return v.W_per_obs ? v.obs_W + 9 * e : v.obs_W;
}

__device__ __forceinline__ int find_block(const int* __restrict__ rowptr, const int* __restrict__ col, int a, int b) {
    int lo = rowptr[a], hi = rowptr[a + 1] - 1;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (col[mid] < b)
            lo = mid + 1;
        else
            hi = mid;
    }
    return lo;  // pattern is built from co-visibility, so the block exists
}

// =============================================================================================
// K1 — materialised residual + Jacobian
// =============================================================================================
constexpr int RJ_TILE = 128;
constexpr int RJ_PMAX = 16;
constexpr int RJ_STAGE_DOUBLES = RJ_TILE * 30;
constexpr size_t RJ_SMEM = 2 * RJ_STAGE_DOUBLES * sizeof(double) + RJ_PMAX * 12 * sizeof(double) + 16;

template <bool kWPerObs>
__global__ void __launch_bounds__(RJ_TILE)
    resjac_kernel(CameraIntrinsics cam, long long n, const uint32_t* __restrict__ cam_idx,
                  const uint32_t* __restrict__ pt_idx, const double* __restrict__ ou,
                  const double* __restrict__ ov, const double* __restrict__ od, const double* __restrict__ Wg,
                  const double* __restrict__ poses, const double* __restrict__ points,
                  const int* __restrict__ cam_free, const int* __restrict__ tile_lo,
                  const int* __restrict__ tile_n, double* __restrict__ out_r, double* __restrict__ out_Jc,
                  double* __restrict__ out_Jp, double* __restrict__ cost_out) {
    extern __shared__ __align__(128) unsigned char smem_rj[];
    double* s_out = reinterpret_cast<double*>(smem_rj);
    double* s_pose = s_out + 2 * RJ_STAGE_DOUBLES;
    uint64_t* bar = reinterpret_cast<uint64_t*>(s_pose + RJ_PMAX * 12);
    __shared__ double s_red[32];

    const int tid = threadIdx.x;
    if (tid == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
    }
    __syncthreads();
    const long long n_tiles = (n + RJ_TILE - 1) / RJ_TILE;
    uint32_t phase = 0;
    int it = 0;
    double cost = 0.0;
    double Wl[9];
    if (!kWPerObs) {
#pragma unroll
        for (int k = 0; k < 9; ++k) Wl[k] = Wg[k];
    }
    for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
        const long long base = t * RJ_TILE;
        const long long i = base + tid;
        const bool active = i < n;
        const int lo = tile_lo[t], np = tile_n[t];
        if (np > 0 && tid == 0) {
            // camera poses of this tile: one bulk async copy global -> shared, completion on mbarrier
            mbar_expect_tx(bar, np * 96);
            tma_load_1d(s_pose, poses + 12ll * lo, np * 96, bar);
        }
        uint32_t c = 0;
        double u = 0, vv = 0, d = 0, p[3] = {0, 0, 1};
        if (active) {
            c = cam_idx[i];
            const uint32_t j = pt_idx[i];
            u = ou[i];
            vv = ov[i];
            d = od[i];
            p[0] = points[3ll * j];
            p[1] = points[3ll * j + 1];
            p[2] = points[3ll * j + 2];
            if (kWPerObs) {
#pragma unroll
                for (int k = 0; k < 9; ++k) Wl[k] = Wg[9 * i + k];
            }
        }
        double pose[12];
        if (np > 0) {
            mbar_wait(bar, phase);
            phase ^= 1;
            const double* sp = s_pose + 12 * (active ? int(c) - lo : 0);
#pragma unroll
            for (int k = 0; k < 12; ++k) pose[k] = sp[k];
        } else {
            const double* gp = poses + 12ll * c;
#pragma unroll
            for (int k = 0; k < 12; ++k) pose[k] = gp[k];
        }
        double r[3] = {0, 0, 0}, Jc[18], Jp[9];
        if (active) {
            stereo_block<true>(cam, pose, p, u, vv, d, Wl, r, Jc, Jp);
            if (cam_free[c] < 0) {
                // constant pose block: Ceres drops its columns (dataset_vo.cpp:62)
#pragma unroll
                for (int k = 0; k < 18; ++k) Jc[k] = 0.0;
            }
            cost += 0.5 * (r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
        }
        const bool full = base + RJ_TILE <= n;
        if (full) {
            double* st = s_out + (it & 1) * RJ_STAGE_DOUBLES;
            // the bulk store that read this stage two tiles ago must have drained
            if (tid == 0) tma_store_wait_read<1>();
            __syncthreads();
#pragma unroll
            for (int k = 0; k < 3; ++k) st[3 * tid + k] = r[k];
#pragma unroll
            for (int k = 0; k < 18; ++k) st[RJ_TILE * 3 + 18 * tid + k] = Jc[k];
#pragma unroll
            for (int k = 0; k < 9; ++k) st[RJ_TILE * 21 + 9 * tid + k] = Jp[k];
            fence_proxy_async();
            __syncthreads();
            if (tid == 0) {
                if (out_r) tma_store_1d(out_r + 3 * base, st, RJ_TILE * 24);
                if (out_Jc) tma_store_1d(out_Jc + 18 * base, st + RJ_TILE * 3, RJ_TILE * 144);
                if (out_Jp) tma_store_1d(out_Jp + 9 * base, st + RJ_TILE * 21, RJ_TILE * 72);
                tma_store_commit();
            }
        } else {
            __syncthreads();  // keep s_pose alive until every thread has read it
            if (active) {
                if (out_r)
                    for (int k = 0; k < 3; ++k) out_r[3 * i + k] = r[k];
                if (out_Jc)
                    for (int k = 0; k < 18; ++k) out_Jc[18 * i + k] = Jc[k];
                if (out_Jp)
                    for (int k = 0; k < 9; ++k) out_Jp[9 * i + k] = Jp[k];
            }
        }
    }
    if (tid == 0) tma_store_wait<0>();
    block_atomic_sum(cost, cost_out, s_red);
}

// =============================================================================================
// initial pass: cost, squared column norms and gradient of the unscaled Jacobian
// =============================================================================================
__global__ void __launch_bounds__(256)
    colnorm_kernel(DevView v, int lm_lo, int lm_hi, double* __restrict__ cn_p, double* __restrict__ cn_l,
                   double* __restrict__ gp, double* __restrict__ gl, double* __restrict__ scal) {
    __shared__ double s_red[32];
    double cost = 0.0;
    const int lane = threadIdx.x & 31;
    // whole warps walk the landmarks together (a lane past the end idles) so that the camera sums can be
    // reduced in the warp: inside a group the 32 landmarks of a warp see the same camera at step k, and
    // one RED per value instead of 32 to the same address is what this pass is bound by
    const int n_lm = lm_hi - lm_lo;
    const int n_round = (n_lm + int(gridDim.x * blockDim.x) - 1) / int(gridDim.x * blockDim.x);
    for (int rd = 0; rd < n_round; ++rd) {
        const int j = lm_lo + rd * int(gridDim.x * blockDim.x) + int(blockIdx.x * blockDim.x + threadIdx.x);
        const bool valid = j < lm_hi;
        const int jj = valid ? j : lm_hi - 1;
        const double p[3] = {v.points[3ll * jj], v.points[3ll * jj + 1], v.points[3ll * jj + 2]};
        double cl[3] = {0, 0, 0}, g[3] = {0, 0, 0};
        const uint32_t cnt = valid ? v.lm_cnt[jj] : 0u, stride = v.lm_stride[jj];
        uint32_t e = v.lm_base[jj];
        const uint32_t cmax = __reduce_max_sync(0xffffffffu, cnt);
        for (uint32_t k = 0; k < cmax; ++k, e += stride) {
            const bool act = k < cnt;
            int f = -1;
            double cn[6] = {0, 0, 0, 0, 0, 0}, gc[6] = {0, 0, 0, 0, 0, 0};
            if (act) {
                const uint32_t c = v.obs_cam[e];
                double r[3], Jc[18], Jp[9];
                stereo_block<true>(v.cam, v.poses + 12ll * c, p, v.obs_u[e], v.obs_v[e], v.obs_d[e], obs_W_ptr(v, e),
                                   r, Jc, Jp);
                cost += 0.5 * (r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
#pragma unroll
                for (int q = 0; q < 3; ++q) {
                    cl[q] += Jp[q] * Jp[q] + Jp[3 + q] * Jp[3 + q] + Jp[6 + q] * Jp[6 + q];
                    g[q] += Jp[q] * r[0] + Jp[3 + q] * r[1] + Jp[6 + q] * r[2];
                }
                f = v.cam_free[c];
                if (f >= 0) {
#pragma unroll
                    for (int q = 0; q < 6; ++q) {
                        cn[q] = Jc[q] * Jc[q] + Jc[6 + q] * Jc[6 + q] + Jc[12 + q] * Jc[12 + q];
                        gc[q] = Jc[q] * r[0] + Jc[6 + q] * r[1] + Jc[12 + q] * r[2];
                    }
                }
            }
            // do all lanes that contribute name the same camera?
            const unsigned contrib = __ballot_sync(0xffffffffu, f >= 0);
            if (contrib == 0u) continue;
            const int leader = __ffs(int(contrib)) - 1;
            const int f0 = __shfl_sync(0xffffffffu, f, leader);
            if (__all_sync(0xffffffffu, f < 0 || f == f0)) {
#pragma unroll
                for (int q = 0; q < 6; ++q) {
                    cn[q] = warp_sum(cn[q]);
                    gc[q] = warp_sum(gc[q]);
                }
                if (lane == 0) {
#pragma unroll
                    for (int q = 0; q < 6; ++q) {
                        red_add(&cn_p[36ll * f0 + 7 * q], cn[q]);
                        red_add(&gp[6ll * f0 + q], gc[q]);
                    }
                }
            } else if (f >= 0) {
#pragma unroll
                for (int q = 0; q < 6; ++q) {
                    red_add(&cn_p[36ll * f + 7 * q], cn[q]);
                    red_add(&gp[6ll * f + q], gc[q]);
                }
            }
        }
        if (valid) {
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                cn_l[3ll * j + q] = cl[q];
                gl[3ll * j + q] = g[q];
            }
        }
    }
    block_atomic_sum(cost, &scal[SC_COST], s_red);
}

__global__ void jacobi_scale_kernel(const double* __restrict__ cn, double* __restrict__ sc, long long n, int enabled) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) sc[i] = enabled ? 1.0 / (1.0 + sqrt(cn[i])) : 1.0;
}

// =============================================================================================
// K2 — generic fused Schur build: one warp per landmark, lanes over its observations
// =============================================================================================
constexpr int SG_WARPS = 4;
constexpr int SG_WARP_DOUBLES = 3 * 18 * 32;  // sW, sY, sW2 stored [k][lane]
constexpr size_t SG_SMEM = SG_WARPS * (SG_WARP_DOUBLES * sizeof(double) + 2 * 32 * sizeof(int));

struct ObsEval {
    double r[3], Jc[18], Jp[9];
    int f;
};

__device__ __forceinline__ void eval_obs_scaled(const DevView& v, long long e, const double* p, const double* sl,
                                                ObsEval& o) {
    const uint32_t c = v.obs_cam[e];
    stereo_block<true>(v.cam, v.poses + 12ll * c, p, v.obs_u[e], v.obs_v[e], v.obs_d[e], obs_W_ptr(v, e), o.r, o.Jc,
                       o.Jp);
    o.f = v.cam_free[c];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
#pragma unroll
        for (int q = 0; q < 3; ++q) o.Jp[3 * k + q] *= sl[q];
    }
    if (o.f >= 0) {
        const double* sp = v.sc_p + 6ll * o.f;
#pragma unroll
        for (int q = 0; q < 6; ++q) {
            const double s = sp[q];
            o.Jc[q] *= s;
            o.Jc[6 + q] *= s;
            o.Jc[12 + q] *= s;
        }
    }
}

// the same with the observation's camera index and measurement already loaded (software prefetch)
__device__ __forceinline__ void eval_obs_scaled_pre(const DevView& v, long long e, uint32_t c, double ou, double ov, double od,
                                                    const double* p, const double* sl, ObsEval& o) {
    stereo_block<true>(v.cam, v.poses + 12ll * c, p, ou, ov, od, obs_W_ptr(v, e), o.r, o.Jc, o.Jp);
    o.f = v.cam_free[c];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
#pragma unroll
        for (int q = 0; q < 3; ++q) o.Jp[3 * k + q] *= sl[q];
    }
    if (o.f >= 0) {
        const double* sp = v.sc_p + 6ll * o.f;
#pragma unroll
        for (int q = 0; q < 6; ++q) {
            const double s = sp[q];
            o.Jc[q] *= s;
            o.Jc[6 + q] *= s;
            o.Jc[12 + q] *= s;
        }
    }
}

// W = Jc^T Jp (6x3 row-major)
__device__ __forceinline__ void form_W(const ObsEval& o, double* W) {
#pragma unroll
    for (int p = 0; p < 6; ++p)
#pragma unroll
        for (int q = 0; q < 3; ++q)
            W[3 * p + q] = o.Jc[p] * o.Jp[q] + o.Jc[6 + p] * o.Jp[3 + q] + o.Jc[12 + p] * o.Jp[6 + q];
}
// Y = W Vi, Vi symmetric 3x3 given by its 6 unique entries
__device__ __forceinline__ void form_Y(const double* W, const double* Vi, double* Y) {
#pragma unroll
    for (int p = 0; p < 6; ++p) {
        const double w0 = W[3 * p], w1 = W[3 * p + 1], w2 = W[3 * p + 2];
        Y[3 * p + 0] = w0 * Vi[0] + w1 * Vi[1] + w2 * Vi[2];
        Y[3 * p + 1] = w0 * Vi[1] + w1 * Vi[3] + w2 * Vi[4];
        Y[3 * p + 2] = w0 * Vi[2] + w1 * Vi[4] + w2 * Vi[5];
    }
}

__device__ __forceinline__ void pair_to_S(const DevView& v, double* __restrict__ S, const double* sY, const double* sW,
                                          int x, int y, int fx, int fy, bool same_obs) {
    // P = Y_x W_y^T ; block(fx,fy) -= P (fx <= fy) else block(fy,fx) -= P^T ;
    // two observations from one camera (fx == fy, x != y): diagonal block -= P + P^T
    double Yx[18], Wy[18];
#pragma unroll
    for (int k = 0; k < 18; ++k) {
        Yx[k] = sY[k * 32 + x];
        Wy[k] = sW[k * 32 + y];
    }
    const bool swap = fx > fy;
    const int a = swap ? fy : fx, b = swap ? fx : fy;
    double* B = S + 36ll * find_block(v.s_rowptr, v.s_col, a, b);
    if (fx == fy) {
        // diagonal block: only its upper triangle is accumulated (finalize mirrors it)
        const bool dup = !same_obs;
#pragma unroll
        for (int p = 0; p < 6; ++p)
#pragma unroll
            for (int q = p; q < 6; ++q) {
                double val = Yx[3 * p] * Wy[3 * q] + Yx[3 * p + 1] * Wy[3 * q + 1] + Yx[3 * p + 2] * Wy[3 * q + 2];
                if (dup) val += Yx[3 * q] * Wy[3 * p] + Yx[3 * q + 1] * Wy[3 * p + 1] + Yx[3 * q + 2] * Wy[3 * p + 2];
                red_add(&B[6 * p + q], -val);
            }
        return;
    }
#pragma unroll
    for (int p = 0; p < 6; ++p)
#pragma unroll
        for (int q = 0; q < 6; ++q) {
            const double val = Yx[3 * p] * Wy[3 * q] + Yx[3 * p + 1] * Wy[3 * q + 1] + Yx[3 * p + 2] * Wy[3 * q + 2];
            red_add(swap ? &B[6 * q + p] : &B[6 * p + q], -val);
        }
}

__global__ void __launch_bounds__(SG_WARPS * 32)
    schur_generic_kernel(DevView v, int lm_lo, int lm_hi, const uint8_t* __restrict__ skip, LmDiag dg, double* __restrict__ S,
                         double* __restrict__ Bdiag, double* __restrict__ bp, double* __restrict__ gp,
                         double* __restrict__ gl, double* __restrict__ scal) {
    extern __shared__ __align__(128) unsigned char smem_sg[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    double* sW = reinterpret_cast<double*>(smem_sg) + wib * SG_WARP_DOUBLES;
    double* sY = sW + 18 * 32;
    double* sW2 = sY + 18 * 32;
    int* sF = reinterpret_cast<int*>(reinterpret_cast<double*>(smem_sg) + SG_WARPS * SG_WARP_DOUBLES) + wib * 64;
    int* sF2 = sF + 32;
    __shared__ double s_red[32];

    double cost = 0.0;
    const int warps_total = gridDim.x * SG_WARPS;
    for (int j = lm_lo + blockIdx.x * SG_WARPS + wib; j < lm_hi; j += warps_total) {
        if (skip && skip[j - lm_lo]) continue;  // taken by the wide-window kernel
        const long long e0 = v.lm_base[j];
        const long long es = v.lm_stride[j];
        const int L = int(v.lm_cnt[j]);
        const double p[3] = {v.points[3ll * j], v.points[3ll * j + 1], v.points[3ll * j + 2]};
        const double sl[3] = {v.sc_l[3ll * j], v.sc_l[3ll * j + 1], v.sc_l[3ll * j + 2]};
        // ---- pass 1: V = sum Jp^T Jp, g = sum Jp^T r --------------------------------------
        double V[6] = {0, 0, 0, 0, 0, 0}, g[3] = {0, 0, 0};
        ObsEval o;
        o.f = -1;
        for (int i = lane; i < L; i += 32) {
            eval_obs_scaled(v, e0 + i * es, p, sl, o);
            cost += 0.5 * (o.r[0] * o.r[0] + o.r[1] * o.r[1] + o.r[2] * o.r[2]);
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const double a = o.Jp[3 * k], b = o.Jp[3 * k + 1], c = o.Jp[3 * k + 2];
                V[0] += a * a;
                V[1] += a * b;
                V[2] += a * c;
                V[3] += b * b;
                V[4] += b * c;
                V[5] += c * c;
                g[0] += a * o.r[k];
                g[1] += b * o.r[k];
                g[2] += c * o.r[k];
            }
        }
#pragma unroll
        for (int k = 0; k < 6; ++k) V[k] = warp_sum(V[k]);
#pragma unroll
        for (int k = 0; k < 3; ++k) g[k] = warp_sum(g[k]);
        // LM diagonal on the point block: clamp(diag) / radius (levenberg_marquardt_strategy)
        V[0] += fmin(fmax(V[0], dg.min_diag), dg.max_diag) * dg.inv_radius;
        V[3] += fmin(fmax(V[3], dg.min_diag), dg.max_diag) * dg.inv_radius;
        V[5] += fmin(fmax(V[5], dg.min_diag), dg.max_diag) * dg.inv_radius;
        double Vi[6];
        const bool pd = invert_sym3(V, Vi);
        if (lane == 0) {
            gl[3ll * j] = g[0];
            gl[3ll * j + 1] = g[1];
            gl[3ll * j + 2] = g[2];
            if (!pd) red_add(&scal[SC_INVALID], 1.0);
        }
        if (!pd) continue;
        // ---- pass 2: camera blocks, in chunks of 32 observations ----------------------------
        for (int c0 = 0; c0 < L; c0 += 32) {
            const int n = min(32, L - c0);
            if (L > 32 && lane < n) eval_obs_scaled(v, e0 + (c0 + lane) * es, p, sl, o);
            __syncwarp();
            if (lane < n) {
                sF[lane] = o.f;
                if (o.f >= 0) {
                    double W[18], Y[18];
                    form_W(o, W);
                    form_Y(W, Vi, Y);
#pragma unroll
                    for (int k = 0; k < 18; ++k) {
                        sW[k * 32 + lane] = W[k];
                        sY[k * 32 + lane] = Y[k];
                    }
                    // U_aa (upper 21) and gradients
                    double* Bd = Bdiag + 36ll * o.f;
#pragma unroll
                    for (int a = 0; a < 6; ++a) {
#pragma unroll
                        for (int b = a; b < 6; ++b)
                            red_add(&Bd[6 * a + b], o.Jc[a] * o.Jc[b] + o.Jc[6 + a] * o.Jc[6 + b] + o.Jc[12 + a] * o.Jc[12 + b]);
                        const double ga = o.Jc[a] * o.r[0] + o.Jc[6 + a] * o.r[1] + o.Jc[12 + a] * o.r[2];
                        const double yg = Y[3 * a] * g[0] + Y[3 * a + 1] * g[1] + Y[3 * a + 2] * g[2];
                        red_add(&gp[6ll * o.f + a], ga);
                        red_add(&bp[6ll * o.f + a], ga - yg);
                    }
                }
            }
            __syncwarp();
            // pairs inside the chunk (x <= y)
            const int npairs = n * (n + 1) / 2;
            for (int pidx = lane; pidx < npairs; pidx += 32) {
                // row x of the upper triangle holds n - x entries
                int x = int((2.0 * n + 1.0 - sqrt((2.0 * n + 1.0) * (2.0 * n + 1.0) - 8.0 * pidx)) * 0.5);
                while (x > 0 && x * n - x * (x - 1) / 2 > pidx) --x;
                while ((x + 1) * n - (x + 1) * x / 2 <= pidx) ++x;
                const int y = x + (pidx - (x * n - x * (x - 1) / 2));
                const int fx = sF[x], fy = sF[y];
                if (fx < 0 || fy < 0) continue;
                pair_to_S(v, S, sY, sW, x, y, fx, fy, x == y);
            }
            // pairs against later chunks (tracks longer than 32 observations)
            for (int d0 = c0 + 32; d0 < L; d0 += 32) {
                const int m = min(32, L - d0);
                __syncwarp();
                if (lane < m) {
                    ObsEval o2;
                    eval_obs_scaled(v, e0 + (d0 + lane) * es, p, sl, o2);
                    sF2[lane] = o2.f;
                    if (o2.f >= 0) {
                        double W[18];
                        form_W(o2, W);
#pragma unroll
                        for (int k = 0; k < 18; ++k) sW2[k * 32 + lane] = W[k];
                    }
                }
                __syncwarp();
                for (int pidx = lane; pidx < n * m; pidx += 32) {
                    const int x = pidx / m, y = pidx - x * m;
                    const int fx = sF[x], fy = sF2[y];
                    if (fx < 0 || fy < 0) continue;
                    pair_to_S(v, S, sY, sW2, x, y, fx, fy, false);
                }
            }
            __syncwarp();
        }
    }
    block_atomic_sum(cost, &scal[SC_COST], s_red);
}

// =============================================================================================
// K2w — Schur build for LONG, ragged tracks: landmarks no group takes (more than 10 frames, or no companions) but
// whose cameras fit a window of 32 consecutive poses.  The per-landmark kernel above issues one RED.ADD.F64 per
// entry of every camera pair (7 560 for a track of 20 frames) and is bound by L2 atomics.  Here the elimination is
// S_win -= Z Z^T per SLICE of such landmarks (consecutive in first-camera order, all inside [c0, c0 + 32)):
//   schur_wide_produce_kernel  warp per landmark, lane per observation: pass 1 with warp sums, U_aa / gradient /
//                              right-hand-side REDs as above, and Z_obs = W A (6 x 3, A A^T = V^-1) to global
//                              memory, 18 doubles per observation;
//   schur_wide_kernel          one CTA (10 warps) per slice: the whole 192 x 192 window tile stays in tensor-core
//                              accumulators (300 upper MMA tiles: a 6 x 6 super-block per warp, 12 fragment loads
//                              per 36 MMAs); batches of 20 landmarks: their Z
//                              rows go to shared memory laid out [column][row] (6 rows per camera slot, rows of
//                              cameras a landmark does not see are zero), DMMA m8n8k4 applies the batch's 60
//                              columns; one flush per slice (entries that stayed zero are skipped).
// Two kernels because the closed-form evaluation needs ~240 registers and the accumulators 76: together they spill.
// =============================================================================================
constexpr int WD_WIN = 32;                  // cameras in the window
constexpr int WD_ROWS = 6 * WD_WIN;         // 192
constexpr int WD_WARPS = 10;                // one per 6 x 6 super-block of the upper triangle of the 24 x 24 MMA tiles
constexpr int WD_BATCH = 2 * WD_WARPS;      // landmarks per batch: two per warp
constexpr int WD_ZS = WD_ROWS + 4;          // row stride of Z[column][row]: 4 mod 16 -> conflict-free fragment loads
constexpr size_t WD_SMEM = sizeof(double) * 3 * WD_BATCH * WD_ZS;

__global__ void __launch_bounds__(SG_WARPS * 32)
    schur_wide_produce_kernel(DevView v, int lm_lo, int lm_hi, const uint8_t* __restrict__ wide, long long obs0, LmDiag dg,
                              double* __restrict__ Zg, double* __restrict__ Bdiag, double* __restrict__ bp,
                              double* __restrict__ gp, double* __restrict__ gl, double* __restrict__ scal) {
    __shared__ double s_red[32];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    double cost = 0.0;
    const int warps_total = gridDim.x * SG_WARPS;
    for (int j = lm_lo + blockIdx.x * SG_WARPS + wib; j < lm_hi; j += warps_total) {
        if (!wide[j - lm_lo]) continue;
        const long long e0 = v.lm_base[j];
        const long long es = v.lm_stride[j];
        const int L = int(v.lm_cnt[j]);   // <= 32 (one observation per camera of the window)
        const double p[3] = {v.points[3ll * j], v.points[3ll * j + 1], v.points[3ll * j + 2]};
        const double sl[3] = {v.sc_l[3ll * j], v.sc_l[3ll * j + 1], v.sc_l[3ll * j + 2]};
        double V[6] = {0, 0, 0, 0, 0, 0}, gq[3] = {0, 0, 0};
        ObsEval o;
        o.f = -1;
        const long long e = e0 + lane * es;
        if (lane < L) {
            eval_obs_scaled(v, e, p, sl, o);
            cost += 0.5 * (o.r[0] * o.r[0] + o.r[1] * o.r[1] + o.r[2] * o.r[2]);
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const double a = o.Jp[3 * k], b = o.Jp[3 * k + 1], c = o.Jp[3 * k + 2];
                V[0] += a * a;
                V[1] += a * b;
                V[2] += a * c;
                V[3] += b * b;
                V[4] += b * c;
                V[5] += c * c;
                gq[0] += a * o.r[k];
                gq[1] += b * o.r[k];
                gq[2] += c * o.r[k];
            }
        }
#pragma unroll
        for (int k = 0; k < 6; ++k) V[k] = warp_sum(V[k]);
#pragma unroll
        for (int k = 0; k < 3; ++k) gq[k] = warp_sum(gq[k]);
        V[0] += fmin(fmax(V[0], dg.min_diag), dg.max_diag) * dg.inv_radius;
        V[3] += fmin(fmax(V[3], dg.min_diag), dg.max_diag) * dg.inv_radius;
        V[5] += fmin(fmax(V[5], dg.min_diag), dg.max_diag) * dg.inv_radius;
        double Vi[6];
        const bool pd = invert_sym3(V, Vi);
        if (lane == 0) {
            gl[3ll * j] = gq[0];
            gl[3ll * j + 1] = gq[1];
            gl[3ll * j + 2] = gq[2];
            if (!pd) red_add(&scal[SC_INVALID], 1.0);
        }
        if (lane >= L) continue;
        double Z[18];
#pragma unroll
        for (int k = 0; k < 18; ++k) Z[k] = 0.0;
        if (pd && o.f >= 0) {
            double W[18], Y[18];
            form_W(o, W);
            form_Y(W, Vi, Y);
            double* Bd = Bdiag + 36ll * o.f;
#pragma unroll
            for (int a = 0; a < 6; ++a) {
#pragma unroll
                for (int b = a; b < 6; ++b)
                    red_add(&Bd[6 * a + b], o.Jc[a] * o.Jc[b] + o.Jc[6 + a] * o.Jc[6 + b] + o.Jc[12 + a] * o.Jc[12 + b]);
                const double ga = o.Jc[a] * o.r[0] + o.Jc[6 + a] * o.r[1] + o.Jc[12 + a] * o.r[2];
                const double yg = Y[3 * a] * gq[0] + Y[3 * a + 1] * gq[1] + Y[3 * a + 2] * gq[2];
                red_add(&gp[6ll * o.f + a], ga);
                red_add(&bp[6ll * o.f + a], ga - yg);
            }
            // A A^T = V^-1 (lower Cholesky factor of the inverse), Z = W A, stored [column][row]
            const double a00 = sqrt(fmax(Vi[0], 0.0)), i00 = a00 > 0.0 ? 1.0 / a00 : 0.0;
            const double a10 = Vi[1] * i00, a20 = Vi[2] * i00;
            const double a11 = sqrt(fmax(Vi[3] - a10 * a10, 0.0)), i11 = a11 > 0.0 ? 1.0 / a11 : 0.0;
            const double a21 = (Vi[4] - a20 * a10) * i11;
            const double a22 = sqrt(fmax(Vi[5] - a20 * a20 - a21 * a21, 0.0));
#pragma unroll
            for (int r = 0; r < 6; ++r) {
                const double w0 = W[3 * r], w1 = W[3 * r + 1], w2 = W[3 * r + 2];
                Z[r] = w0 * a00 + w1 * a10 + w2 * a20;
                Z[6 + r] = w1 * a11 + w2 * a21;
                Z[12 + r] = w2 * a22;
            }
        }
        double2* out = reinterpret_cast<double2*>(Zg + 18 * (e - obs0));
#pragma unroll
        for (int k = 0; k < 9; ++k) out[k] = make_double2(Z[2 * k], Z[2 * k + 1]);
    }
    block_atomic_sum(cost, &scal[SC_COST], s_red);
}

__device__ __forceinline__ void wd_dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// warp -> 6 x 6 super-block (wr, wc), wr <= wc, of the 4 x 4 grid of super-blocks over the 24 x 24 MMA tiles; ordered so
// that the four schedulers (warp % 4) get 36 + 21 + 21, 36 + 21 + 21, 36 + 36, 36 + 36 tiles
__device__ __forceinline__ void wd_super_block(int warp, int& wr, int& wc) {
    constexpr unsigned WR = 0u | 0u << 2 | 0u << 4 | 1u << 6 | 0u << 8 | 2u << 10 | 1u << 12 | 2u << 14 | 1u << 16 | 3u << 18;
    constexpr unsigned WC = 1u | 2u << 2 | 3u << 4 | 2u << 6 | 0u << 8 | 2u << 10 | 3u << 12 | 3u << 14 | 1u << 16 | 3u << 18;
    wr = (WR >> (2 * warp)) & 3;
    wc = (WC >> (2 * warp)) & 3;
}

__global__ void __launch_bounds__(WD_WARPS * 32, 1)
    schur_wide_kernel(DevView v, int n_slices, const int* __restrict__ sl_lo, const int* __restrict__ sl_hi,
                      const int* __restrict__ sl_c0, long long obs0, const double* __restrict__ Zg, double* __restrict__ S,
                      double* __restrict__ scal) {
    extern __shared__ __align__(16) double s_Z[];   // [3 * WD_BATCH][WD_ZS]
    __shared__ int s_free[WD_WIN];
    __shared__ int s_blk[WD_WIN * (WD_WIN + 1) / 2];   // pair (a <= b) -> block of S, -1: none
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, q = lane & 3;
    int wr, wc;
    wd_super_block(warp, wr, wc);
    const bool diag = wr == wc;
    for (int sidx = blockIdx.x; sidx < n_slices; sidx += gridDim.x) {
        const int lo = sl_lo[sidx], hi = sl_hi[sidx], c0 = sl_c0[sidx];
        __syncthreads();   // (the previous slice's flush has read s_blk)
        if (tid < WD_WIN) s_free[tid] = c0 + tid < v.n_cams ? v.cam_free[c0 + tid] : -1;
        __syncthreads();
        for (int pidx = tid; pidx < WD_WIN * (WD_WIN + 1) / 2; pidx += WD_WARPS * 32) {
            // pair index -> (a, b), a <= b, row-major upper triangle of the 32 x 32 slot pairs
            int a = 0, rem = pidx;
            while (rem >= WD_WIN - a) {
                rem -= WD_WIN - a;
                ++a;
            }
            const int b = a + rem, fa = s_free[a], fb = s_free[b];
            int e = -1;
            if (fa >= 0 && fb >= 0) {
                int l0 = v.s_rowptr[fa], h0 = v.s_rowptr[fa + 1];
                while (l0 < h0) {
                    const int mid = (l0 + h0) >> 1;
                    if (v.s_col[mid] < fb)
                        l0 = mid + 1;
                    else
                        h0 = mid;
                }
                if (l0 < v.s_rowptr[fa + 1] && v.s_col[l0] == fb) e = l0;
            }
            s_blk[pidx] = e;
        }
        // tile (i, jx) of the warp's super-block: MMA row tile 6 wr + i, column tile 6 wc + jx
        double acc[6][6][2];
#pragma unroll
        for (int i = 0; i < 6; ++i)
#pragma unroll
            for (int jx = 0; jx < 6; ++jx) acc[i][jx][0] = acc[i][jx][1] = 0.0;

        for (int b0 = lo; b0 < hi; b0 += WD_BATCH) {
            const int nb = min(WD_BATCH, hi - b0);
            __syncthreads();   // (the previous batch's MMAs have read Z)
            for (int idx = tid; idx < 3 * WD_BATCH * WD_ZS; idx += WD_WARPS * 32) s_Z[idx] = 0.0;
            __syncthreads();
            // ---- stage: warp = landmark, lane = observation; its 18 doubles go to rows 6 slot .. + 5 of 3 columns ----
            for (int l = warp; l < nb; l += WD_WARPS) {
                const int j = b0 + l;
                const int L = int(v.lm_cnt[j]);
                if (lane < L) {
                    const long long e = (long long)v.lm_base[j] + lane * (long long)v.lm_stride[j];
                    const int slot = int(v.obs_cam[e]) - c0;
                    const double2* in = reinterpret_cast<const double2*>(Zg + 18 * (e - obs0));
                    if (slot >= 0 && slot < WD_WIN) {
                        double* Zc = s_Z + (3 * l) * WD_ZS + 6 * slot;
                        // entries 2k, 2k + 1 of [column][row 0..5]: column (2k) / 6, row (2k) % 6 (even: both in one column)
                        double2 z[9];
#pragma unroll
                        for (int k = 0; k < 9; ++k) z[k] = in[k];
#pragma unroll
                        for (int k = 0; k < 9; ++k)
                            *reinterpret_cast<double2*>(Zc + ((2 * k) / 6) * WD_ZS + (2 * k) % 6) = z[k];
                    } else {
                        red_add(&scal[SC_INVALID], 1.0);   // (the slicing guarantees the window)
                    }
                }
            }
            __syncthreads();
            // ---- consume: the batch's 3 nb columns, four per k-step; 12 fragments feed the warp's 36 (21) MMAs ----
            const int ksteps = (3 * nb + 3) / 4;
            for (int ks = 0; ks < ksteps; ++ks) {
                const double* Zk = s_Z + (4 * ks + q) * WD_ZS + g;
                double fa[6], fb[6];
#pragma unroll
                for (int i = 0; i < 6; ++i) {
                    fa[i] = Zk[8 * (6 * wr + i)];
                    fb[i] = Zk[8 * (6 * wc + i)];
                }
#pragma unroll
                for (int i = 0; i < 6; ++i)
#pragma unroll
                    for (int jx = 0; jx < 6; ++jx) {
                        if (jx < i && diag) continue;   // below the diagonal of a diagonal super-block
                        wd_dmma(acc[i][jx][0], acc[i][jx][1], fa[i], fb[jx]);
                    }
            }
        }
        // ---- flush: lane (g, q) holds entries (row 8 I + g, columns 8 J + 2 q, + 1) of its tiles ----
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            const int r = 8 * (6 * wr + i) + g;
            const int a = r / 6, pr = r - 6 * a;
#pragma unroll
            for (int jx = 0; jx < 6; ++jx) {
                if (jx < i && diag) continue;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int c = 8 * (6 * wc + jx) + 2 * q + h;
                    const double val = acc[i][jx][h];
                    if (c < r || val == 0.0) continue;   // lower part of a diagonal MMA tile / a pair nobody observes
                    const int b = c / 6, pc = c - 6 * b;
                    // pair (a, b), a <= b: index in the row-major upper triangle
                    const int e = s_blk[a * WD_WIN - a * (a - 1) / 2 + (b - a)];
                    if (e >= 0) red_add(&S[36ll * e + 6 * pr + pc], -val);
                }
            }
        }
    }
}

// =============================================================================================
// camera-only blocks (sun sensor, pose prior)
// =============================================================================================
__device__ __forceinline__ void camonly_eval(const DevView& v, const SunBlockData* suns, int n_sun,
                                             const PriorBlockData* priors, int i, const double* poses, bool want_J,
                                             double* r, double* J, int* rows, int* cam, double* cost) {
    if (i < n_sun) {
        const SunBlockData& s = suns[i];
        *cam = int(s.cam);
        *rows = 2;
        sun_block(poses + 12ll * s.cam, s.obs_c, s.ref_g, s.W, s.az_thresh, s.zen_thresh, r, want_J ? J : nullptr);
        const double sq = r[0] * r[0] + r[1] * r[1];
        double rho0 = sq, sr = 1.0;
        if (s.huber > 0.0) huber_rho(s.huber, sq, &rho0, &sr);
        *cost = 0.5 * rho0;
        r[0] *= sr;
        r[1] *= sr;
        if (want_J)
            for (int k = 0; k < 12; ++k) J[k] *= sr;
    } else {
        const PriorBlockData& p = priors[i - n_sun];
        *cam = int(p.cam);
        *rows = 6;
        prior_block(poses + 12ll * p.cam, p.Tref, p.W, r, want_J ? J : nullptr);
        double c = 0;
        for (int k = 0; k < 6; ++k) c += 0.5 * r[k] * r[k];
        *cost = c;
    }
}

__global__ void camonly_build_kernel(DevView v, const SunBlockData* suns, int n_sun, const PriorBlockData* priors,
                                     int n_prior, double* Bdiag, double* bp, double* gp, double* scal) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_sun + n_prior) return;
    double r[6], J[36], cost;
    int rows, cam;
    camonly_eval(v, suns, n_sun, priors, i, v.poses, true, r, J, &rows, &cam, &cost);
    red_add(&scal[SC_COST], cost);
    const int f = v.cam_free[cam];
    if (f < 0) return;
    const double* sp = v.sc_p + 6ll * f;
    for (int a = 0; a < 6; ++a) {
        for (int b = a; b < 6; ++b) {
            double s = 0;
            for (int k = 0; k < rows; ++k) s += J[6 * k + a] * J[6 * k + b];
            red_add(&Bdiag[36ll * f + 6 * a + b], s * sp[a] * sp[b]);
        }
        double g = 0;
        for (int k = 0; k < rows; ++k) g += J[6 * k + a] * r[k];
        red_add(&gp[6ll * f + a], g * sp[a]);
        red_add(&bp[6ll * f + a], g * sp[a]);
    }
}

// model cost change and candidate cost of the camera-only blocks
__global__ void camonly_step_kernel(DevView v, const SunBlockData* suns, int n_sun, const PriorBlockData* priors,
                                    int n_prior, const double* yp, const double* poses_cand, double* scal2) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_sun + n_prior) return;
    double r[6], J[36], cost;
    int rows, cam;
    camonly_eval(v, suns, n_sun, priors, i, v.poses, true, r, J, &rows, &cam, &cost);
    const int f = v.cam_free[cam];
    double acc = 0;
    for (int k = 0; k < rows; ++k) {
        double m = 0;
        if (f >= 0)
            for (int a = 0; a < 6; ++a) m -= J[6 * k + a] * v.sc_p[6ll * f + a] * yp[6ll * f + a];
        acc -= m * (r[k] + 0.5 * m);
    }
    red_add(&scal2[SC_MODEL], acc);
    double rc[6], cc;
    camonly_eval(v, suns, n_sun, priors, i, poses_cand, false, rc, nullptr, &rows, &cam, &cc);
    red_add(&scal2[SC_CAND_COST], cc);
}

// =============================================================================================
// finalize: S_aa += U_aa + D^2, symmetrise, block-Jacobi inverse
// =============================================================================================
__device__ __forceinline__ bool chol6_inverse(const double* A, double* Ainv) {
    double L[36];
    for (int i = 0; i < 36; ++i) L[i] = A[i];
    for (int j = 0; j < 6; ++j) {
        double d = L[6 * j + j];
        for (int k = 0; k < j; ++k) d -= L[6 * j + k] * L[6 * j + k];
        if (!(d > 0.0)) return false;
        d = sqrt(d);
        L[6 * j + j] = d;
        for (int i = j + 1; i < 6; ++i) {
            double s = L[6 * i + j];
            for (int k = 0; k < j; ++k) s -= L[6 * i + k] * L[6 * j + k];
            L[6 * i + j] = s / d;
        }
    }
    for (int c = 0; c < 6; ++c) {
        double e[6];
        for (int i = 0; i < 6; ++i) {
            double s = (i == c) ? 1.0 : 0.0;
            for (int k = 0; k < i; ++k) s -= L[6 * i + k] * e[k];
            e[i] = s / L[6 * i + i];
        }
        for (int i = 5; i >= 0; --i) {
            double s = e[i];
            for (int k = i + 1; k < 6; ++k) s -= L[6 * k + i] * e[k];
            e[i] = s / L[6 * i + i];
        }
        for (int r = 0; r < 6; ++r) Ainv[6 * r + c] = e[r];
    }
    return true;
}

__global__ void finalize_kernel(DevView v, LmDiag dg, int preconditioner, double* S, double* Bdiag, double* diag_p,
                                double* Minv, double* scal) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= v.n_free) return;
    double U[36];
    double* Bd = Bdiag + 36ll * f;
    for (int a = 0; a < 6; ++a)
        for (int b = a; b < 6; ++b) U[6 * a + b] = U[6 * b + a] = Bd[6 * a + b];
    for (int a = 0; a < 6; ++a) {
        const double dd = fmin(fmax(U[7 * a], dg.min_diag), dg.max_diag);
        diag_p[6ll * f + a] = dd;
        U[7 * a] += dd * dg.inv_radius;
    }
    double* Sd = S + 36ll * v.s_rowptr[f];  // diagonal block is the first block of the row
    double A[36];
    for (int a = 0; a < 6; ++a)
        for (int b = a; b < 6; ++b) {
            const double sv = Sd[6 * a + b] + U[6 * a + b];  // Schur part arrives as an upper triangle
            A[6 * a + b] = A[6 * b + a] = sv;
        }
    for (int k = 0; k < 36; ++k) {
        Sd[k] = A[k];
        Bd[k] = U[k];
    }
    double Mi[36];
    if (!chol6_inverse(preconditioner == 0 ? U : A, Mi)) {
        red_add(&scal[SC_INVALID], 1.0);
        for (int k = 0; k < 36; ++k) Mi[k] = (k % 7 == 0) ? 1.0 : 0.0;
    }
    for (int k = 0; k < 36; ++k) Minv[36ll * f + k] = Mi[k];
}

// =============================================================================================
// K4 — Plus, back-substitution, model cost change, candidate cost
// =============================================================================================
__global__ void pose_plus_kernel(DevView v, const double* __restrict__ yp, double* __restrict__ poses_cand,
                                 double* __restrict__ scal2, int count_cams) {
    __shared__ double s_red[32];
    const int c = blockIdx.x * blockDim.x + threadIdx.x;  // thread per camera
    double sn = 0, xn = 0, bad = 0;
    if (c < v.n_cams) {
        const int ff = v.cam_free[c];
        const double* x = v.poses + 12ll * c;
        double* y = poses_cand + 12ll * c;
        if (ff >= 0) {
            double eps[6];
            for (int k = 0; k < 6; ++k) {
                eps[k] = -yp[6ll * ff + k] * v.sc_p[6ll * ff + k];
                if (isnan(eps[k]) || isinf(eps[k])) bad = 1;
            }
            double out[12];
            se3_plus(x, eps, out);
            for (int k = 0; k < 12; ++k) {
                y[k] = out[k];
                sn += (x[k] - out[k]) * (x[k] - out[k]);
                xn += out[k] * out[k];
            }
        } else {
            for (int k = 0; k < 12; ++k) y[k] = x[k];
        }
    }
    if (!count_cams) {
        sn = 0;
        xn = 0;
    }
    block_atomic_sum(sn, &scal2[SC_STEP_NORM2], s_red);
    block_atomic_sum(xn, &scal2[SC_XNORM2], s_red);
    block_atomic_sum(bad, &scal2[SC_NONFINITE], s_red);
}

__global__ void __launch_bounds__(128, 4)
    backsub_kernel(DevView v, int lm_lo, int lm_hi, LmDiag dg, const double* __restrict__ yp,
                   const double* __restrict__ poses_cand, double* __restrict__ points_cand,
                   double* __restrict__ yl_out, double* __restrict__ scal2) {
    __shared__ double s_red[32];
    double model = 0, ccost = 0, sn = 0, xn = 0, bad = 0;
    for (int j = lm_lo + blockIdx.x * blockDim.x + threadIdx.x; j < lm_hi; j += gridDim.x * blockDim.x) {
        const long long e0 = v.lm_base[j], es = v.lm_stride[j];
        const long long e1 = e0 + es * v.lm_cnt[j];
        const double p[3] = {v.points[3ll * j], v.points[3ll * j + 1], v.points[3ll * j + 2]};
        const double sl[3] = {v.sc_l[3ll * j], v.sc_l[3ll * j + 1], v.sc_l[3ll * j + 2]};
        // One Jacobian evaluation per observation.  With Jy_i = Jc_i y_p the pass accumulates
        //   V0 = sum Jp^T Jp,  gl = sum Jp^T r,  q = sum Jp^T Jy (= W^T y_p),  a = sum Jy.r,  b = sum |Jy|^2
        // so that, once y_l = V^-1 (gl - q) is known, the landmark's share of the model cost change
        //   sum_i -(J s).(r + J s / 2),  s = -y,  J s = -(Jp y_l + Jy)
        // is  y_l.gl + a - (y_l^T V0 y_l + 2 y_l.q + b) / 2  without a second pass over the Jacobians.
        double V[6] = {0, 0, 0, 0, 0, 0}, tg[3] = {0, 0, 0}, tq[3] = {0, 0, 0}, sa = 0, sb = 0;
        ObsEval o;
        // the next observation's camera index and measurement are requested one iteration ahead
        uint32_t nc = 0;
        double nu = 0, nv = 0, nd = 0;
        if (e0 < e1) {
            nc = v.obs_cam[e0];
            nu = v.obs_u[e0];
            nv = v.obs_v[e0];
            nd = v.obs_d[e0];
        }
        for (long long e = e0; e < e1; e += es) {
            const uint32_t cc = nc;
            const double cu = nu, cv = nv, cd = nd;
            if (e + es < e1) {
                nc = v.obs_cam[e + es];
                nu = v.obs_u[e + es];
                nv = v.obs_v[e + es];
                nd = v.obs_d[e + es];
            }
            eval_obs_scaled_pre(v, e, cc, cu, cv, cd, p, sl, o);
            double Jy[3] = {0, 0, 0};
            if (o.f >= 0) {
                const double* y = yp + 6ll * o.f;
#pragma unroll
                for (int k = 0; k < 3; ++k)
#pragma unroll
                    for (int a = 0; a < 6; ++a) Jy[k] += o.Jc[6 * k + a] * y[a];
            }
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const double a = o.Jp[3 * k], b = o.Jp[3 * k + 1], c = o.Jp[3 * k + 2];
                V[0] += a * a;
                V[1] += a * b;
                V[2] += a * c;
                V[3] += b * b;
                V[4] += b * c;
                V[5] += c * c;
                tg[0] += a * o.r[k];
                tg[1] += b * o.r[k];
                tg[2] += c * o.r[k];
                tq[0] += a * Jy[k];
                tq[1] += b * Jy[k];
                tq[2] += c * Jy[k];
                sa += Jy[k] * o.r[k];
                sb += Jy[k] * Jy[k];
            }
        }
        const double V0[6] = {V[0], V[1], V[2], V[3], V[4], V[5]};
        V[0] += fmin(fmax(V[0], dg.min_diag), dg.max_diag) * dg.inv_radius;
        V[3] += fmin(fmax(V[3], dg.min_diag), dg.max_diag) * dg.inv_radius;
        V[5] += fmin(fmax(V[5], dg.min_diag), dg.max_diag) * dg.inv_radius;
        double Vi[6];
        double yl[3] = {0, 0, 0};
        if (invert_sym3(V, Vi)) {
            // t = g_l - W^T yp
            const double t[3] = {tg[0] - tq[0], tg[1] - tq[1], tg[2] - tq[2]};
            yl[0] = Vi[0] * t[0] + Vi[1] * t[1] + Vi[2] * t[2];
            yl[1] = Vi[1] * t[0] + Vi[3] * t[1] + Vi[4] * t[2];
            yl[2] = Vi[2] * t[0] + Vi[4] * t[1] + Vi[5] * t[2];
        }
        {
            const double yVy = yl[0] * (V0[0] * yl[0] + V0[1] * yl[1] + V0[2] * yl[2]) +
                               yl[1] * (V0[1] * yl[0] + V0[3] * yl[1] + V0[4] * yl[2]) +
                               yl[2] * (V0[2] * yl[0] + V0[4] * yl[1] + V0[5] * yl[2]);
            const double yg = yl[0] * tg[0] + yl[1] * tg[1] + yl[2] * tg[2];
            const double yq = yl[0] * tq[0] + yl[1] * tq[1] + yl[2] * tq[2];
            model += yg + sa - 0.5 * (yVy + 2.0 * yq + sb);
        }
        double pn[3];
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            const double dl = -yl[q] * sl[q];
            if (isnan(dl) || isinf(dl)) bad = 1;
            pn[q] = p[q] + dl;
            points_cand[3ll * j + q] = pn[q];
            yl_out[3ll * j + q] = yl[q];
            sn += (p[q] - pn[q]) * (p[q] - pn[q]);
            xn += pn[q] * pn[q];
        }
        // the cost at the candidate
        if (e0 < e1) {
            nc = v.obs_cam[e0];
            nu = v.obs_u[e0];
            nv = v.obs_v[e0];
            nd = v.obs_d[e0];
        }
        for (long long e = e0; e < e1; e += es) {
            double rc[3];
            const uint32_t c = nc;
            const double cu = nu, cv = nv, cd = nd;
            if (e + es < e1) {
                nc = v.obs_cam[e + es];
                nu = v.obs_u[e + es];
                nv = v.obs_v[e + es];
                nd = v.obs_d[e + es];
            }
            stereo_block<false>(v.cam, poses_cand + 12ll * c, pn, cu, cv, cd, obs_W_ptr(v, e), rc,
                                nullptr, nullptr);
            ccost += 0.5 * (rc[0] * rc[0] + rc[1] * rc[1] + rc[2] * rc[2]);
        }
    }
    block_atomic_sum(model, &scal2[SC_MODEL], s_red);
    block_atomic_sum(ccost, &scal2[SC_CAND_COST], s_red);
    block_atomic_sum(sn, &scal2[SC_STEP_NORM2], s_red);
    block_atomic_sum(xn, &scal2[SC_XNORM2], s_red);
    block_atomic_sum(bad, &scal2[SC_NONFINITE], s_red);
}


// =============================================================================================
// DOGLEG — DoglegStrategy of Ceres works in the coordinates step' = D step with
// D = sqrt(clamp(diag(J^T J))).  Every vector it forms (scaled gradient g' = D^-1 g, Gauss-Newton
// point gn' = -D y, the Cauchy point, the subspace basis, the final step) is a combination of g' and
// gn', so the device only has to provide, once per Jacobian, eight sums:
//   DG_G11 = g'.g'   DG_G12 = g'.gn'   DG_G22 = gn'.gn'
//   DG_JGG = |J D^-2 g|^2   DG_JGY = (J D^-2 g).(J y)   DG_JYY = |J y|^2   DG_JGR, DG_JYR = their dots with r
// and then the step Y = -c1 D^-2 g + c2 y for the coefficients the host picks (engine.cu).
// =============================================================================================
__global__ void __launch_bounds__(128)
    dogleg_products_kernel(DevView v, int lm_lo, int lm_hi, LmDiag dg, const double* __restrict__ gp,
                           const double* __restrict__ diag_p, const double* __restrict__ yp, const double* __restrict__ gl,
                           const double* __restrict__ yl, double* __restrict__ diag_l_out, double* __restrict__ sums) {
    __shared__ double s_red[32];
    double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int j = lm_lo + blockIdx.x * blockDim.x + threadIdx.x; j < lm_hi; j += gridDim.x * blockDim.x) {
        const long long e0 = v.lm_base[j], es = v.lm_stride[j];
        const long long e1 = e0 + es * v.lm_cnt[j];
        const double p[3] = {v.points[3ll * j], v.points[3ll * j + 1], v.points[3ll * j + 2]};
        const double sl[3] = {v.sc_l[3ll * j], v.sc_l[3ll * j + 1], v.sc_l[3ll * j + 2]};
        double d2[3] = {0, 0, 0};
        ObsEval o;
        for (long long e = e0; e < e1; e += es) {
            eval_obs_scaled(v, e, p, sl, o);
#pragma unroll
            for (int q = 0; q < 3; ++q) d2[q] += o.Jp[q] * o.Jp[q] + o.Jp[3 + q] * o.Jp[3 + q] + o.Jp[6 + q] * o.Jp[6 + q];
        }
        double tg[3], ty[3];
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            d2[q] = fmin(fmax(d2[q], dg.min_diag), dg.max_diag);
            diag_l_out[3ll * j + q] = d2[q];
            const double g = gl[3ll * j + q], y = yl[3ll * j + q];
            tg[q] = g / d2[q];
            ty[q] = y;
            acc[0] += g * g / d2[q];
            acc[1] -= g * y;
            acc[2] += d2[q] * y * y;
        }
        for (long long e = e0; e < e1; e += es) {
            eval_obs_scaled(v, e, p, sl, o);
            const double* gpf = o.f >= 0 ? gp + 6ll * o.f : nullptr;
            const double* dpf = o.f >= 0 ? diag_p + 6ll * o.f : nullptr;
            const double* ypf = o.f >= 0 ? yp + 6ll * o.f : nullptr;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                double jg = o.Jp[3 * k] * tg[0] + o.Jp[3 * k + 1] * tg[1] + o.Jp[3 * k + 2] * tg[2];
                double jy = o.Jp[3 * k] * ty[0] + o.Jp[3 * k + 1] * ty[1] + o.Jp[3 * k + 2] * ty[2];
                if (o.f >= 0) {
#pragma unroll
                    for (int a = 0; a < 6; ++a) {
                        jg += o.Jc[6 * k + a] * (gpf[a] / dpf[a]);
                        jy += o.Jc[6 * k + a] * ypf[a];
                    }
                }
                acc[3] += jg * jg;
                acc[4] += jg * jy;
                acc[5] += jy * jy;
                acc[6] += jg * o.r[k];
                acc[7] += jy * o.r[k];
            }
        }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) block_atomic_sum(acc[k], &sums[k], s_red);
}

// camera part of the three inner products and the camera-only blocks' rows (thread per camera,
// then thread per block)
__global__ void dogleg_products_cam_kernel(DevView v, const SunBlockData* suns, int n_sun, const PriorBlockData* priors,
                                           int n_prior, const double* __restrict__ gp, const double* __restrict__ diag_p,
                                           const double* __restrict__ yp, double* __restrict__ sums) {
    __shared__ double s_red[32];
    double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < v.n_free) {
        for (int a = 0; a < 6; ++a) {
            const double g = gp[6ll * i + a], d2 = diag_p[6ll * i + a], y = yp[6ll * i + a];
            acc[0] += g * g / d2;
            acc[1] -= g * y;
            acc[2] += d2 * y * y;
        }
    } else if (i - v.n_free < n_sun + n_prior) {
        double r[6], J[36], cost;
        int rows, cam;
        camonly_eval(v, suns, n_sun, priors, i - v.n_free, v.poses, true, r, J, &rows, &cam, &cost);
        const int f = v.cam_free[cam];
        if (f >= 0)
            for (int k = 0; k < rows; ++k) {
                double jg = 0, jy = 0;
                for (int a = 0; a < 6; ++a) {
                    const double js = J[6 * k + a] * v.sc_p[6ll * f + a];
                    jg += js * (gp[6ll * f + a] / diag_p[6ll * f + a]);
                    jy += js * yp[6ll * f + a];
                }
                acc[3] += jg * jg;
                acc[4] += jg * jy;
                acc[5] += jy * jy;
                acc[6] += jg * r[k];
                acc[7] += jy * r[k];
            }
    }
    for (int k = 0; k < 8; ++k) block_atomic_sum(acc[k], &sums[k], s_red);
}

// Y = -c1 g / D^2 + c2 y
__global__ void dogleg_combine_kernel(long long n, double c1, double c2, const double* __restrict__ g,
                                      const double* __restrict__ d2, const double* __restrict__ y, double* __restrict__ out) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) out[i] = -c1 * g[i] / d2[i] + c2 * y[i];
}

// candidate landmarks p - Y sl and the cost there
__global__ void __launch_bounds__(128)
    points_apply_kernel(DevView v, int lm_lo, int lm_hi, const double* __restrict__ Yl, const double* __restrict__ poses_cand,
                        double* __restrict__ points_cand, double* __restrict__ scal2) {
    __shared__ double s_red[32];
    double ccost = 0, sn = 0, xn = 0, bad = 0;
    for (int j = lm_lo + blockIdx.x * blockDim.x + threadIdx.x; j < lm_hi; j += gridDim.x * blockDim.x) {
        const long long e0 = v.lm_base[j], es = v.lm_stride[j];
        const long long e1 = e0 + es * v.lm_cnt[j];
        double p[3], pn[3];
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            p[q] = v.points[3ll * j + q];
            const double dl = -Yl[3ll * j + q] * v.sc_l[3ll * j + q];
            if (isnan(dl) || isinf(dl)) bad = 1;
            pn[q] = p[q] + dl;
            points_cand[3ll * j + q] = pn[q];
            sn += (p[q] - pn[q]) * (p[q] - pn[q]);
            xn += pn[q] * pn[q];
        }
        for (long long e = e0; e < e1; e += es) {
            double rc[3];
            const uint32_t c = v.obs_cam[e];
            stereo_block<false>(v.cam, poses_cand + 12ll * c, pn, v.obs_u[e], v.obs_v[e], v.obs_d[e], obs_W_ptr(v, e), rc,
                                nullptr, nullptr);
            ccost += 0.5 * (rc[0] * rc[0] + rc[1] * rc[1] + rc[2] * rc[2]);
        }
    }
    block_atomic_sum(ccost, &scal2[SC_CAND_COST], s_red);
    block_atomic_sum(sn, &scal2[SC_STEP_NORM2], s_red);
    block_atomic_sum(xn, &scal2[SC_XNORM2], s_red);
    block_atomic_sum(bad, &scal2[SC_NONFINITE], s_red);
}

// |x - Plus(x, -g)|_inf (trust_region_minimizer: gradient norm in ambient coordinates) and |x|^2
__global__ void gradnorm_kernel(DevView v, int lm_lo, int lm_hi, const double* __restrict__ gp_s,
                                const double* __restrict__ gl_s, double* __restrict__ scal, int count_cams) {
    __shared__ double s_red[32];
    // grid-stride: a few CTAs per SM, so that the two reductions end in ~1 k atomics instead of one per 256 elements
    const long long n = (long long)v.n_cams + (lm_hi - lm_lo);
    double m = 0, xn = 0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        if (i < v.n_cams) {
            const int f = v.cam_free[i];
            if (f >= 0 && count_cams) {
                const double* x = v.poses + 12ll * i;
                double eps[6], out[12];
                for (int k = 0; k < 6; ++k) eps[k] = -gp_s[6ll * f + k] / v.sc_p[6ll * f + k];
                se3_plus(x, eps, out);
                for (int k = 0; k < 12; ++k) {
                    m = fmax(m, fabs(x[k] - out[k]));
                    xn += x[k] * x[k];
                }
            }
        } else {
            const long long j = lm_lo + (i - v.n_cams);
            for (int q = 0; q < 3; ++q) {
                const double x = v.points[3 * j + q];
                const double g = gl_s[3 * j + q] / v.sc_l[3 * j + q];
                m = fmax(m, fabs(x - (x - g)));
                xn += x * x;
            }
        }
    }
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0) atomic_max_nonneg(&scal[SC_GRADMAX], m);
    block_atomic_sum(xn, &scal[SC_XNORM2_CUR], s_red);
}

// FP64 FMA microbenchmark: 8 independent chains per thread, register resident
__global__ void fp64_peak_kernel(double* out, int iters) {
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6,
           a7 = a0 + 7;
    const double b = 1.0000001, c = 1e-9;
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, b, c);
        a1 = fma(a1, b, c);
        a2 = fma(a2, b, c);
        a3 = fma(a3, b, c);
        a4 = fma(a4, b, c);
        a5 = fma(a5, b, c);
        a6 = fma(a6, b, c);
        a7 = fma(a7, b, c);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

// camera scaling from the diagonal of Bdiag (squared column norms of the unscaled Jacobian)
__global__ void jacobi_scale_cams_kernel(const double* __restrict__ Bdiag, double* __restrict__ sc, int nf, int enabled) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 6 * nf) {
        const int f = i / 6, a = i - 6 * f;
        sc[i] = enabled ? 1.0 / (1.0 + sqrt(Bdiag[36ll * f + 7 * a])) : 1.0;
    }
}

// materialised residuals / Jacobians of the camera-only blocks (cslam_evaluate)
__global__ void camonly_eval_kernel(DevView v, const SunBlockData* suns, int n_sun, const PriorBlockData* priors,
                                    int n_prior, int apply_loss, double* r_sun, double* J_sun, double* r_pr,
                                    double* J_pr, double* cost_out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_sun + n_prior) return;
    double r[6], J[36], cost;
    int rows, cam;
    if (i < n_sun && !apply_loss) {
        const SunBlockData& s = suns[i];
        cam = int(s.cam);
        rows = 2;
        sun_block(v.poses + 12ll * s.cam, s.obs_c, s.ref_g, s.W, s.az_thresh, s.zen_thresh, r, J);
        cost = 0.5 * (r[0] * r[0] + r[1] * r[1]);
    } else {
        camonly_eval(v, suns, n_sun, priors, i, v.poses, true, r, J, &rows, &cam, &cost);
    }
    red_add(cost_out, cost);
    const bool is_const = v.cam_free[cam] < 0;
    if (i < n_sun) {
        for (int k = 0; k < 2; ++k) r_sun[2 * i + k] = r[k];
        for (int k = 0; k < 12; ++k) J_sun[12 * i + k] = is_const ? 0.0 : J[k];
    } else {
        const int q = i - n_sun;
        for (int k = 0; k < 6; ++k) r_pr[6 * q + k] = r[k];
        for (int k = 0; k < 36; ++k) J_pr[36 * q + k] = is_const ? 0.0 : J[k];
    }
}

__global__ void gather_obs_kernel(long long n_obs, const uint32_t* __restrict__ obs_user, const uint32_t* __restrict__ raw_cam,
                                  const double* __restrict__ raw_uvd, const double* __restrict__ raw_W, int W_per_obs,
                                  uint32_t* __restrict__ obs_cam, double* __restrict__ u, double* __restrict__ v,
                                  double* __restrict__ d, double* __restrict__ W) {
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n_obs; e += (long long)gridDim.x * blockDim.x) {
        const long long i = obs_user[e];
        obs_cam[e] = raw_cam[i];
        u[e] = raw_uvd[3 * i];
        v[e] = raw_uvd[3 * i + 1];
        d[e] = raw_uvd[3 * i + 2];
        if (W_per_obs) {
#pragma unroll
            for (int k = 0; k < 9; ++k) W[9 * e + k] = raw_W[9 * i + k];
        }
    }
    if (!W_per_obs && blockIdx.x == 0 && threadIdx.x < 9) W[threadIdx.x] = raw_W[threadIdx.x];
}
__global__ void gather_points_kernel(int n_lm, const uint32_t* __restrict__ lm_user, const double* __restrict__ raw_pts,
                                     double* __restrict__ points) {
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a < n_lm) {
        const long long j = lm_user[a];
        points[3ll * a] = raw_pts[3 * j];
        points[3ll * a + 1] = raw_pts[3 * j + 1];
        points[3ll * a + 2] = raw_pts[3 * j + 2];
    }
}
__global__ void scatter_points_kernel(int n_lm, const uint32_t* __restrict__ lm_user, const double* __restrict__ points,
                                      double* __restrict__ raw_pts) {
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a < n_lm) {
        const long long j = lm_user[a];
        raw_pts[3 * j] = points[3ll * a];
        raw_pts[3 * j + 1] = points[3ll * a + 1];
        raw_pts[3 * j + 2] = points[3ll * a + 2];
    }
}
// multi-GPU download: owned landmarks go into a zeroed [x y z flag] buffer that is summed over the
// ranks (every point has at most one owner, and x + 0 = x exactly), then merged where flag > 0
__global__ void scatter_points4_kernel(int n_lm, const uint32_t* __restrict__ lm_user, const double* __restrict__ points,
                                       double* __restrict__ z4) {
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a < n_lm) {
        const long long j = lm_user[a];
        z4[4 * j] = points[3ll * a];
        z4[4 * j + 1] = points[3ll * a + 1];
        z4[4 * j + 2] = points[3ll * a + 2];
        z4[4 * j + 3] = 1.0;
    }
}
__global__ void merge_points4_kernel(long long n_points, const double* __restrict__ z4, double* __restrict__ raw_pts) {
    const long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (j < n_points && z4[4 * j + 3] > 0.0) {
        raw_pts[3 * j] = z4[4 * j];
        raw_pts[3 * j + 1] = z4[4 * j + 1];
        raw_pts[3 * j + 2] = z4[4 * j + 2];
    }
}
__global__ void fill_kernel(double* p, size_t n, double value) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = value;
}

inline int grid_for(long long n, int block, int max_blocks) {
    long long g = (n + block - 1) / block;
    if (g < 1) g = 1;
    if (g > max_blocks) g = max_blocks;
    return int(g);
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------
void launch_resjac(cudaStream_t s, const CameraIntrinsics& cam, long long n, const uint32_t* cam_idx,
                   const uint32_t* pt_idx, const double* u, const double* v, const double* d, const double* W,
                   int W_per_obs, const double* poses, const double* points, const int* cam_free, const int* tile_lo,
                   const int* tile_n, double* r, double* Jc, double* Jp, double* cost) {
    if (n <= 0) return;
    static PerDevice attr_done;
    const int dev_ = PerDevice::current();
    if (attr_done.first_use(dev_)) {
        CSLAM_CUDA(cudaFuncSetAttribute(resjac_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(RJ_SMEM)));
        CSLAM_CUDA(cudaFuncSetAttribute(resjac_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(RJ_SMEM)));
        attr_done.mark(dev_);
    }
    const long long tiles = (n + RJ_TILE - 1) / RJ_TILE;
    const int grid = int(tiles < 3ll * kSMs ? tiles : 3ll * kSMs);  // persistent: 3 CTAs per SM
    if (W_per_obs)
        resjac_kernel<true><<<grid, RJ_TILE, RJ_SMEM, s>>>(cam, n, cam_idx, pt_idx, u, v, d, W, poses, points, cam_free,
                                                           tile_lo, tile_n, r, Jc, Jp, cost);
    else
        resjac_kernel<false><<<grid, RJ_TILE, RJ_SMEM, s>>>(cam, n, cam_idx, pt_idx, u, v, d, W, poses, points, cam_free,
                                                            tile_lo, tile_n, r, Jc, Jp, cost);
    CSLAM_LAUNCHED(1);
    CSLAM_CUDA(cudaGetLastError());
}

void launch_colnorm(cudaStream_t s, const DevView& v, int lm_lo, int lm_hi, double* cn_p, double* cn_l, double* gp,
                    double* gl, double* scal) {
    if (lm_hi <= lm_lo) return;
    colnorm_kernel<<<grid_for(lm_hi - lm_lo, 256, 8 * kSMs), 256, 0, s>>>(v, lm_lo, lm_hi, cn_p, cn_l, gp, gl, scal);
    CSLAM_LAUNCHED(1);
    CSLAM_CUDA(cudaGetLastError());
}

void launch_jacobi_scale(cudaStream_t s, const double* cn, double* sc, long long n, int enabled) {
    if (n <= 0) return;
    jacobi_scale_kernel<<<int((n + 255) / 256), 256, 0, s>>>(cn, sc, n, enabled);
    CSLAM_LAUNCHED(1);
    CSLAM_CUDA(cudaGetLastError());
}

void launch_jacobi_scale_cams(cudaStream_t s, const double* Bdiag, double* sc, int nf, int enabled) {
    if (nf <= 0) return;
    jacobi_scale_cams_kernel<<<(6 * nf + 255) / 256, 256, 0, s>>>(Bdiag, sc, nf, enabled);
    CSLAM_LAUNCHED(1);
    CSLAM_CUDA(cudaGetLastError());
}

void launch_camonly_eval(cudaStream_t s, const DevView& v, const SunBlockData* suns, int n_sun,
                         const PriorBlockData* priors, int n_prior, int apply_loss, double* r_sun, double* J_sun,
                         double* r_pr, double* J_pr, double* cost) {
    const int n = n_sun + n_prior;
    if (n <= 0) return;
    camonly_eval_kernel<<<(n + 63) / 64, 64, 0, s>>>(v, suns, n_sun, priors, n_prior, apply_loss, r_sun, J_sun, r_pr,
                                                    J_pr, cost);
    CSLAM_LAUNCHED(1);
    CSLAM_CUDA(cudaGetLastError());
}

void launch_schur_wide(cudaStream_t s, const DevView& v, int lm_lo, int lm_hi, const uint8_t* wide, int n_slices, const int* sl_lo,
                       const int* sl_hi, const int* sl_c0, long long obs0, double* Zg, LmDiag dg, double* S, double* Bdiag,
                       double* bp, double* gp, double* gl, double* scal) {
    if (n_slices <= 0 || lm_hi <= lm_lo) return;
    schur_wide_produce_kernel<<<grid_for(lm_hi - lm_lo, SG_WARPS, 8 * kSMs), SG_WARPS * 32, 0, s>>>(v, lm_lo, lm_hi, wide, obs0, dg, Zg,
                                                                                                Bdiag, bp, gp, gl, scal);
    // (the opt-in is per device and costs a microsecond: set at every launch rather than cached process-wide)
    CSLAM_CUDA(cudaFuncSetAttribute(schur_wide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(WD_SMEM)));
    const int grid = n_slices < kSMs ? n_slices : kSMs;
    schur_wide_kernel<<<grid, WD_WARPS * 32, WD_SMEM, s>>>(v, n_slices, sl_lo, sl_hi, sl_c0, obs0, Zg, S, scal);
    CSLAM_LAUNCHED(2);
    CSLAM_CUDA(cudaGetLastError());
}

void launch_schur_generic(cudaStream_t s, const DevView& v, int lm_lo, int lm_hi, const uint8_t* skip, LmDiag dg, double* S,
                          double* Bdiag, double* bp, double* gp, double* gl, double* scal) {
    if (lm_hi <= lm_lo) return;
    static PerDevice attr_done;
    const int dev_ = PerDevice::current();
    if (attr_done.first_use(dev_)) {
        CSLAM_CUDA(cudaFuncSetAttribute(schur_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(SG_SMEM)));
        attr_done.mark(dev_);
    }
    const int grid = grid_for(lm_hi - lm_lo, SG_WARPS, 4 * kSMs);
    schur_generic_kernel<<<grid, SG_WARPS * 32, SG_SMEM, s>>>(v, lm_lo, lm_hi, skip, dg, S, Bdiag, bp, gp, gl, scal);
    CSLAM_LAUNCHED(1);
    CSLAM_CUDA(cudaGetLastError());
}

void launch_camonly_build(cudaStream_t s, const DevView& v, const SunBlockData* suns, int n_sun,
                          const PriorBlockData* priors, int n_prior, double* Bdiag, double* bp, double* gp, double* scal) {
    const int n = n_sun + n_prior;
    if (n <= 0) return;
    camonly_build_kernel<<<(n + 63) / 64, 64, 0, s>>>(v, suns, n_sun, priors, n_prior, Bdiag, bp, gp, scal);
    CSLAM_LAUNCHED(1);
    CSLAM_CUDA(cudaGetLastError());
}

void launch_finalize(cudaStream_t s, const DevView& v, LmDiag dg, int preconditioner, double* S, double* Bdiag,
                     double* diag_p, double* Minv, double* scal) {
    if (v.n_free <= 0) return;
    finalize_kernel<<<(v.n_free + 63) / 64, 64, 0, s>>>(v, dg, preconditioner, S, Bdiag, diag_p, Minv, scal);
    CSLAM_LAUNCHED(1);
    CSLAM_CUDA(cudaGetLastError());
}

void launch_pose_plus(cudaStream_t s, const DevView& v, const double* yp, double* poses_cand, double* scal2,
                      int count_cams) {
    pose_plus_kernel<<<(v.n_cams + 127) / 128, 128, 0, s>>>(v, yp, poses_cand, scal2, count_cams);
    CSLAM_LAUNCHED(1);
    CSLAM_CUDA(cudaGetLastError());
}

void launch_backsub(cudaStream_t s, const DevView& v, int lm_lo, int lm_hi, LmDiag dg, const double* yp,
                    const double* poses_cand, double* points_cand, double* yl, double* scal2) {
    if (lm_hi <= lm_lo) return;
    backsub_kernel<<<grid_for(lm_hi - lm_lo, 128, 16 * kSMs), 128, 0, s>>>(v, lm_lo, lm_hi, dg, yp, poses_cand,
                                                                          points_cand, yl, scal2);
    CSLAM_LAUNCHED(1);
    CSLAM_CUDA(cudaGetLastError());
}

void launch_camonly_step(cudaStream_t s, const DevView& v, const SunBlockData* suns, int n_sun,
                         const PriorBlockData* priors, int n_prior, const double* yp, const double* poses_cand,
                         double* scal2) {
    const int n = n_sun + n_prior;
    if (n <= 0) return;
    camonly_step_kernel<<<(n + 63) / 64, 64, 0, s>>>(v, suns, n_sun, priors, n_prior, yp, poses_cand, scal2);
    CSLAM_LAUNCHED(1);
    CSLAM_CUDA(cudaGetLastError());
}

void launch_dogleg_products(cudaStream_t s, const DevView& v, int lm_lo, int lm_hi, LmDiag dg, const SunBlockData* suns,
                            int n_sun, const PriorBlockData* priors, int n_prior, const double* gp, const double* diag_p,
                            const double* yp, const double* gl, const double* yl, double* diag_l, double* sums, int count_cams) {
    if (lm_hi > lm_lo) {
        dogleg_products_kernel<<<grid_for(lm_hi - lm_lo, 128, 8 * kSMs), 128, 0, s>>>(v, lm_lo, lm_hi, dg, gp, diag_p, yp, gl, yl,
                                                                                     diag_l, sums);
        CSLAM_LAUNCHED(1);
    }
    const int n = v.n_free + n_sun + n_prior;
    if (count_cams && n > 0) {
        dogleg_products_cam_kernel<<<(n + 127) / 128, 128, 0, s>>>(v, suns, n_sun, priors, n_prior, gp, diag_p, yp, sums);
        CSLAM_LAUNCHED(1);
    }
    CSLAM_CUDA(cudaGetLastError());
}
void launch_dogleg_combine(cudaStream_t s, long long n, double c1, double c2, const double* g, const double* d2, const double* y,
                           double* out) {
    if (n <= 0) return;
    dogleg_combine_kernel<<<int((n + 255) / 256), 256, 0, s>>>(n, c1, c2, g, d2, y, out);
    CSLAM_LAUNCHED(1);
    CSLAM_CUDA(cudaGetLastError());
}
void launch_points_apply(cudaStream_t s, const DevView& v, int lm_lo, int lm_hi, const double* Yl, const double* poses_cand,
                         double* points_cand, double* scal2) {
    if (lm_hi <= lm_lo) return;
    points_apply_kernel<<<grid_for(lm_hi - lm_lo, 128, 8 * kSMs), 128, 0, s>>>(v, lm_lo, lm_hi, Yl, poses_cand, points_cand, scal2);
    CSLAM_LAUNCHED(1);
    CSLAM_CUDA(cudaGetLastError());
}

void launch_gradnorm(cudaStream_t s, const DevView& v, int lm_lo, int lm_hi, const double* gp_scaled,
                     const double* gl_scaled, double* scal, int count_cams) {
    const long long n = (long long)v.n_cams + (lm_hi - lm_lo);
    if (n <= 0) return;
    gradnorm_kernel<<<int(std::min<long long>((n + 255) / 256, 8 * kSMs)), 256, 0, s>>>(v, lm_lo, lm_hi, gp_scaled, gl_scaled, scal,
                                                                                         count_cams);
    CSLAM_LAUNCHED(1);
    CSLAM_CUDA(cudaGetLastError());
}

void launch_gather_layout(cudaStream_t s, long long n_obs, const uint32_t* obs_user, const uint32_t* raw_cam,
                          const double* raw_uvd, const double* raw_W, int W_per_obs, uint32_t* obs_cam, double* u,
                          double* v, double* d, double* W, int n_lm, const uint32_t* lm_user, const double* raw_pts,
                          double* points) {
    gather_obs_kernel<<<grid_for(std::max<long long>(n_obs, 1), 256, 16 * kSMs), 256, 0, s>>>(n_obs, obs_user, raw_cam, raw_uvd, raw_W,
                                                                                   W_per_obs, obs_cam, u, v, d, W);
    CSLAM_LAUNCHED(1);
    CSLAM_CUDA(cudaGetLastError());
    if (n_lm > 0) {
        gather_points_kernel<<<(n_lm + 255) / 256, 256, 0, s>>>(n_lm, lm_user, raw_pts, points);
        CSLAM_LAUNCHED(1);
        CSLAM_CUDA(cudaGetLastError());
    }
}

void launch_scatter_points(cudaStream_t s, int n_lm, const uint32_t* lm_user, const double* points, double* raw_pts) {
    if (n_lm <= 0) return;
    scatter_points_kernel<<<(n_lm + 255) / 256, 256, 0, s>>>(n_lm, lm_user, points, raw_pts);
    CSLAM_LAUNCHED(1);
    CSLAM_CUDA(cudaGetLastError());
}

void launch_scatter_points4(cudaStream_t s, int n_lm, const uint32_t* lm_user, const double* points, double* z4) {
    if (n_lm <= 0) return;
    scatter_points4_kernel<<<(n_lm + 255) / 256, 256, 0, s>>>(n_lm, lm_user, points, z4);
    CSLAM_LAUNCHED(1);
    CSLAM_CUDA(cudaGetLastError());
}

void launch_merge_points4(cudaStream_t s, long long n_points, const double* z4, double* raw_pts) {
    if (n_points <= 0) return;
    merge_points4_kernel<<<int((n_points + 255) / 256), 256, 0, s>>>(n_points, z4, raw_pts);
    CSLAM_LAUNCHED(1);
    CSLAM_CUDA(cudaGetLastError());
}

void launch_fill(cudaStream_t s, double* p, size_t n, double value) {
    if (!n) return;
    fill_kernel<<<grid_for((long long)n, 256, 8 * kSMs), 256, 0, s>>>(p, n, value);
    CSLAM_LAUNCHED(1);
    CSLAM_CUDA(cudaGetLastError());
}

double measure_fp64_peak_tflops(int device) {
    CSLAM_CUDA(cudaSetDevice(device));
    const int blocks = kSMs * 8, threads = 256, iters = 1 << 15;
    double* out = nullptr;
    CSLAM_CUDA(cudaMalloc(&out, size_t(blocks) * threads * sizeof(double)));
    cudaEvent_t a, b;
    CSLAM_CUDA(cudaEventCreate(&a));
    CSLAM_CUDA(cudaEventCreate(&b));
    fp64_peak_kernel<<<blocks, threads>>>(out, 1024);  // warm-up
    double best = 0;
    for (int rep = 0; rep < 5; ++rep) {
        CSLAM_CUDA(cudaEventRecord(a));
        fp64_peak_kernel<<<blocks, threads>>>(out, iters);
        CSLAM_CUDA(cudaEventRecord(b));
        CSLAM_CUDA(cudaEventSynchronize(b));
        float ms = 0;
        CSLAM_CUDA(cudaEventElapsedTime(&ms, a, b));
        const double flops = 2.0 * 8.0 * iters * double(blocks) * threads;
        const double tf = flops / (ms * 1e-3) / 1e12;
        if (tf > best) best = tf;
    }
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    cudaFree(out);
    return best;
}

}  // namespace cslam
