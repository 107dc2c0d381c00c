// Batched independent sliding-window problems (BASELINE.json configs 1, 2 and 4: `dataset_vo
// --window N`, `dataset_vo_sun --window 2`, and scripts/ba_all_*.sh running many tracks).
//
// A window has a handful of poses and a few hundred landmarks: far too little work for a
// host-driven loop of kernel launches (one LM iteration of the generic engine is ~10 launches and
// ~6 host synchronisations).  Here ONE CTA owns ONE window and runs the whole
// Levenberg-Marquardt loop on the device: residual/Jacobian evaluation, Schur elimination of the
// landmark blocks into a dense reduced camera system held in shared memory, in-CTA Cholesky
// (the SPARSE_SCHUR-equivalent exact solve), back-substitution, step acceptance and trust-region
// bookkeeping (Ceres rules; the same sequence Engine::lm_iterate drives from
// the host).  A batch of windows is one launch; nothing returns to the host until every window
// has terminated.  The kernel is instantiated twice: Levenberg-Marquardt, and DOGLEG (TRADITIONAL /
// SUBSPACE, dataset_vo_sun.cpp:142-143 — the default of the workload scripts/ba_all_*.sh runs): the
// Gauss-Newton solve goes through the same in-CTA Schur path with mu in place of 1/radius, the
// eight inner products of dogleg_products_kernel are formed by the CTA, and thread 0 runs the
// scalar strategy of dogleg_host.h (the functions the host-driven engine runs on the CPU).
#include <algorithm>
#include <cmath>
#include <cstring>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <mutex>

#include "kernels.cuh"

namespace cslam {

namespace {

constexpr int WIN_PMAX = 8;             // poses per window
constexpr int WIN_NMAX = 6 * WIN_PMAX;  // reduced system dimension
constexpr int WIN_THREADS = 128;
constexpr int WIN_CAMV = 33;            // per camera: U upper triangle (21) | g (6) | rhs (6)

struct WinOpts {
    int max_iter, nonmono, max_nonmono, max_invalid, jacobi, dogleg_type;
    double r0, rmax, rmin, min_rel, dmin, dmax, ftol, gtol, ptol;
};

struct WinDesc {
    CameraIntrinsics cam;
    WinOpts o;
    int n_poses, n_free, n_lm, n_sun, n_prior, W_per_obs, log_cap, pad;
    long long pose_off, lm_off, lmptr_off, obs_off, W_off, sun_off, prior_off, log_off;
};

struct WinBufs {
    const WinDesc* desc;
    double* poses;             // [sum n_poses][12]  in: initial, out: best
    const int* cam_free;       // [sum n_poses]      local free index or -1
    double* pts;               // [sum n_lm][3]      in: initial
    double* pts_cand;
    double* pts_best;          // out: best
    const uint32_t* lm_ptr;    // per window n_lm + 1 offsets into its observation range
    const uint8_t* obs_cam;    // local pose index
    const double *obs_u, *obs_v, *obs_d, *obs_W;
    double* sc_l;              // [sum n_lm][3] Jacobi scaling of the point columns
    double* gl;                // [sum n_lm][3] scaled point gradient of the last Schur pass
    double* yl;                // [sum n_lm][3] DOGLEG: landmark part of the Gauss-Newton solve (scaled)
    double* diag_l;            // [sum n_lm][3] DOGLEG: clamp(diag(J^T J)) of the point columns (scaled)
    const SunBlockData* suns;
    const PriorBlockData* priors;
    cslam_summary* summaries;
    double* logs;
    int* log_rows;
};

// Per-CTA view of one window.
struct Win {
    CameraIntrinsics cam;
    int n_poses, n_free, n_lm, n_sun, n_prior, W_per_obs;
    const int* free_idx;       // shared: local pose -> free index
    const double* scp;         // shared: [6 n_free] Jacobi scaling of the pose columns
    const uint32_t* lm_ptr;
    const uint8_t* obs_cam;
    const double *obs_u, *obs_v, *obs_d, *obs_W;
    double* sc_l;
    double* gl;
    const SunBlockData* suns;
    const PriorBlockData* priors;
};

struct WObs {
    double r[3], Jc[18], Jp[9];
    int f;
};

__device__ __forceinline__ void w_eval(const Win& w, uint32_t e, const double* poses, const double* p, const double* sl,
                                       WObs& o) {
    const int c = w.obs_cam[e];
    const double* Wm = w.W_per_obs ? w.obs_W + 9ll * e : w.obs_W;
    stereo_block<true>(w.cam, poses + 12 * c, p, w.obs_u[e], w.obs_v[e], w.obs_d[e], Wm, o.r, o.Jc, o.Jp);
    o.f = w.free_idx[c];
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
        for (int q = 0; q < 3; ++q) o.Jp[3 * k + q] *= sl[q];
    if (o.f >= 0) {
        const double* sp = w.scp + 6 * o.f;
#pragma unroll
        for (int q = 0; q < 6; ++q) {
            const double s = sp[q];
            o.Jc[q] *= s;
            o.Jc[6 + q] *= s;
            o.Jc[12 + q] *= s;
        }
    }
}

// Sum over the CTA, result to every thread, fixed summation order.
__device__ __forceinline__ double cta_sum(double v, double* sh) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
    if (lane == 0) sh[wp] = v;
    __syncthreads();
    double t = 0.0;
#pragma unroll
    for (int i = 0; i < WIN_THREADS / 32; ++i) t += sh[i];
    __syncthreads();
    return t;
}
__device__ __forceinline__ double cta_max(double v, double* sh) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
    if (lane == 0) sh[wp] = v;
    __syncthreads();
    double t = 0.0;
#pragma unroll
    for (int i = 0; i < WIN_THREADS / 32; ++i) t = fmax(t, sh[i]);
    __syncthreads();
    return t;
}

// Add `vals` of every valid lane into dst(slot)[k].  When the valid lanes of the warp agree on
// the slot (the usual case: the k-th observation of every landmark of a window is taken by the
// same pose) the warp reduces with shuffles and issues one shared-memory add per value.
template <int NV>
__device__ __forceinline__ void warp_accumulate(bool valid, int slot, const double* vals, double* base, int stride) {
    const unsigned m = __ballot_sync(0xffffffffu, valid);
    if (m == 0) return;
    const int leader = __ffs(m) - 1;
    const int s0 = __shfl_sync(0xffffffffu, slot, leader);
    const bool uniform = __all_sync(0xffffffffu, !valid || slot == s0);
    if (uniform) {
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            const double t = warp_sum(valid ? vals[k] : 0.0);
            if ((threadIdx.x & 31) == 0) atomicAdd(&base[s0 * stride + k], t);
        }
    } else if (valid) {
#pragma unroll
        for (int k = 0; k < NV; ++k) atomicAdd(&base[slot * stride + k], vals[k]);
    }
}

// S block (fx, fy) -= Zx Zy^T on the dense reduced system (row-major, leading dimension n)
__device__ __forceinline__ void warp_accumulate_pair(bool valid, int fx, int fy, const double* Zx, const double* Zy,
                                                     double* S, int n) {
    const unsigned m = __ballot_sync(0xffffffffu, valid);
    if (m == 0) return;
    const int key = fx * WIN_PMAX + fy;
    const int leader = __ffs(m) - 1;
    const int k0 = __shfl_sync(0xffffffffu, key, leader);
    const bool uniform = __all_sync(0xffffffffu, !valid || key == k0);
    if (uniform) {
        double* B = S + 6 * (k0 / WIN_PMAX) * n + 6 * (k0 % WIN_PMAX);
#pragma unroll
        for (int p = 0; p < 6; ++p)
#pragma unroll
            for (int q = 0; q < 6; ++q) {
                double val = 0.0;
                if (valid) val = Zx[3 * p] * Zy[3 * q] + Zx[3 * p + 1] * Zy[3 * q + 1] + Zx[3 * p + 2] * Zy[3 * q + 2];
                val = warp_sum(val);
                if ((threadIdx.x & 31) == 0) atomicAdd(&B[p * n + q], -val);
            }
    } else if (valid) {
        double* B = S + 6 * fx * n + 6 * fy;
#pragma unroll
        for (int p = 0; p < 6; ++p)
#pragma unroll
            for (int q = 0; q < 6; ++q)
                atomicAdd(&B[p * n + q], -(Zx[3 * p] * Zy[3 * q] + Zx[3 * p + 1] * Zy[3 * q + 1] + Zx[3 * p + 2] * Zy[3 * q + 2]));
    }
}

// Z = (Jc^T Jp) C^-T with Ci = C^-1 (lower: 00,10,11,20,21,22)
__device__ __forceinline__ void form_Z(const WObs& o, const double* Ci, double* Z) {
#pragma unroll
    for (int p = 0; p < 6; ++p) {
        const double w0 = o.Jc[p] * o.Jp[0] + o.Jc[6 + p] * o.Jp[3] + o.Jc[12 + p] * o.Jp[6];
        const double w1 = o.Jc[p] * o.Jp[1] + o.Jc[6 + p] * o.Jp[4] + o.Jc[12 + p] * o.Jp[7];
        const double w2 = o.Jc[p] * o.Jp[2] + o.Jc[6 + p] * o.Jp[5] + o.Jc[12 + p] * o.Jp[8];
        // (W C^-T)[p][q] = sum_k W[p][k] Ci[q][k]
        Z[3 * p + 0] = w0 * Ci[0];
        Z[3 * p + 1] = w0 * Ci[1] + w1 * Ci[2];
        Z[3 * p + 2] = w0 * Ci[3] + w1 * Ci[4] + w2 * Ci[5];
    }
}

// V (00,01,02,11,12,22) = C C^T; returns C^-1 (lower) — false when V is not positive definite
__device__ __forceinline__ bool chol3_inverse(const double* V, double* Ci) {
    if (!(V[0] > 0.0)) return false;
    const double c00 = sqrt(V[0]);
    const double c10 = V[1] / c00, c20 = V[2] / c00;
    const double d1 = V[3] - c10 * c10;
    if (!(d1 > 0.0)) return false;
    const double c11 = sqrt(d1);
    const double c21 = (V[4] - c20 * c10) / c11;
    const double d2 = V[5] - c20 * c20 - c21 * c21;
    if (!(d2 > 0.0) || !(d2 < 1.7976931348623157e308)) return false;
    const double c22 = sqrt(d2);
    const double i00 = 1.0 / c00, i11 = 1.0 / c11, i22 = 1.0 / c22;
    Ci[0] = i00;
    Ci[1] = -c10 * i00 * i11;
    Ci[2] = i11;
    Ci[3] = -(c20 * Ci[0] + c21 * Ci[1]) * i22;
    Ci[4] = -c21 * i11 * i22;
    Ci[5] = i22;
    return true;
}

// One pass over the window's landmarks at `poses`/`pts`:
//   kColnorm: cost, squared column norms (camera: diagonal of U, landmark: cn_out), gradient
//   else    : cost, dense S (Schur part), U, g, rhs with the LM diagonal of `dg`
// s_cam is [n_free][WIN_CAMV]; S is [n][n]; both zeroed here.  Returns cost and the number of
// landmark blocks whose V was not positive definite.
template <bool kColnorm>
__device__ void window_schur_pass(const Win& w, const double* poses, const double* pts, LmDiag dg, double* S, int n,
                                  double* s_cam, double* cn_out, double* s_red, double* cost_out, double* invalid_out) {
    const int tid = threadIdx.x;
    for (int i = tid; i < n * n; i += WIN_THREADS) S[i] = 0.0;
    for (int i = tid; i < w.n_free * WIN_CAMV; i += WIN_THREADS) s_cam[i] = 0.0;
    __syncthreads();
    double cost = 0.0, invalid = 0.0;
    for (int j0 = 0; j0 < w.n_lm; j0 += WIN_THREADS) {
        const int j = j0 + tid;
        const bool active = j < w.n_lm;
        double p[3] = {0, 0, 1}, sl[3] = {1, 1, 1};
        uint32_t e0 = 0;
        int cnt = 0;
        if (active) {
            e0 = w.lm_ptr[j];
            cnt = int(w.lm_ptr[j + 1] - e0);
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                p[q] = pts[3 * j + q];
                sl[q] = w.sc_l[3 * j + q];
            }
        }
        double V[6] = {0, 0, 0, 0, 0, 0}, g[3] = {0, 0, 0};
        WObs o;
        for (int k = 0; k < cnt; ++k) {
            w_eval(w, e0 + k, poses, p, sl, o);
            cost += 0.5 * (o.r[0] * o.r[0] + o.r[1] * o.r[1] + o.r[2] * o.r[2]);
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                const double a = o.Jp[3 * i], b = o.Jp[3 * i + 1], c = o.Jp[3 * i + 2];
                V[0] += a * a; V[1] += a * b; V[2] += a * c; V[3] += b * b; V[4] += b * c; V[5] += c * c;
                g[0] += a * o.r[i]; g[1] += b * o.r[i]; g[2] += c * o.r[i];
            }
        }
        double Ci[6] = {1, 0, 1, 0, 0, 1}, h[3] = {0, 0, 0};
        bool pd = true;
        if (active) {
#pragma unroll
            for (int q = 0; q < 3; ++q) w.gl[3 * j + q] = g[q];
            if (kColnorm) {
                cn_out[3 * j] = V[0];
                cn_out[3 * j + 1] = V[3];
                cn_out[3 * j + 2] = V[5];
            } else {
                V[0] += fmin(fmax(V[0], dg.min_diag), dg.max_diag) * dg.inv_radius;
                V[3] += fmin(fmax(V[3], dg.min_diag), dg.max_diag) * dg.inv_radius;
                V[5] += fmin(fmax(V[5], dg.min_diag), dg.max_diag) * dg.inv_radius;
                pd = chol3_inverse(V, Ci);
                if (!pd) invalid += 1.0;
                h[0] = Ci[0] * g[0];
                h[1] = Ci[1] * g[0] + Ci[2] * g[1];
                h[2] = Ci[3] * g[0] + Ci[4] * g[1] + Ci[5] * g[2];
            }
        }
        int maxcnt = cnt;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) maxcnt = max(maxcnt, __shfl_xor_sync(0xffffffffu, maxcnt, off));
        for (int x = 0; x < maxcnt; ++x) {
            bool vx = active && pd && x < cnt;
            double Zx[18];
            int fx = 0;
            if (vx) {
                w_eval(w, e0 + x, poses, p, sl, o);
                fx = o.f;
                vx = fx >= 0;
            }
            double vals[WIN_CAMV];
            if (vx) {
                if (!kColnorm) form_Z(o, Ci, Zx);
                int t = 0;
#pragma unroll
                for (int a = 0; a < 6; ++a)
#pragma unroll
                    for (int b = a; b < 6; ++b)
                        vals[t++] = o.Jc[a] * o.Jc[b] + o.Jc[6 + a] * o.Jc[6 + b] + o.Jc[12 + a] * o.Jc[12 + b];
#pragma unroll
                for (int a = 0; a < 6; ++a) {
                    const double ga = o.Jc[a] * o.r[0] + o.Jc[6 + a] * o.r[1] + o.Jc[12 + a] * o.r[2];
                    vals[21 + a] = ga;
                    vals[27 + a] = kColnorm ? ga : ga - (Zx[3 * a] * h[0] + Zx[3 * a + 1] * h[1] + Zx[3 * a + 2] * h[2]);
                }
            }
            warp_accumulate<WIN_CAMV>(vx, fx, vals, s_cam, WIN_CAMV);
            if (kColnorm) continue;
            for (int y = 0; y < maxcnt; ++y) {
                bool vy = vx && y < cnt;
                double Zy[18];
                int fy = 0;
                if (vy) {
                    if (y == x) {
                        fy = fx;
#pragma unroll
                        for (int k = 0; k < 18; ++k) Zy[k] = Zx[k];
                    } else {
                        WObs o2;
                        w_eval(w, e0 + y, poses, p, sl, o2);
                        fy = o2.f;
                        vy = fy >= 0;
                        if (vy) form_Z(o2, Ci, Zy);
                    }
                }
                warp_accumulate_pair(vy, fx, fy, Zx, Zy, S, n);
            }
        }
    }
    // sun-sensor and pose-prior blocks: camera-only contributions (camonly_build_kernel)
    if (tid < w.n_sun + w.n_prior) {
        double r[6], J[36];
        int rows, cam;
        if (tid < w.n_sun) {
            const SunBlockData& s = w.suns[tid];
            cam = int(s.cam);
            rows = 2;
            sun_block(poses + 12 * cam, s.obs_c, s.ref_g, s.W, s.az_thresh, s.zen_thresh, r, J);
            const double sq = r[0] * r[0] + r[1] * r[1];
            double rho0 = sq, sr = 1.0;
            if (s.huber > 0.0) huber_rho(s.huber, sq, &rho0, &sr);
            cost += 0.5 * rho0;
            r[0] *= sr;
            r[1] *= sr;
            for (int k = 0; k < 12; ++k) J[k] *= sr;
        } else {
            const PriorBlockData& pr = w.priors[tid - w.n_sun];
            cam = int(pr.cam);
            rows = 6;
            prior_block(poses + 12 * cam, pr.Tref, pr.W, r, J);
            for (int k = 0; k < 6; ++k) cost += 0.5 * r[k] * r[k];
        }
        const int f = w.free_idx[cam];
        if (f >= 0) {
            const double* sp = w.scp + 6 * f;
            int t = 0;
            for (int a = 0; a < 6; ++a)
                for (int b = a; b < 6; ++b) {
                    double s = 0;
                    for (int k = 0; k < rows; ++k) s += J[6 * k + a] * J[6 * k + b];
                    atomicAdd(&s_cam[f * WIN_CAMV + t++], s * sp[a] * sp[b]);
                }
            for (int a = 0; a < 6; ++a) {
                double gg = 0;
                for (int k = 0; k < rows; ++k) gg += J[6 * k + a] * r[k];
                atomicAdd(&s_cam[f * WIN_CAMV + 21 + a], gg * sp[a]);
                atomicAdd(&s_cam[f * WIN_CAMV + 27 + a], gg * sp[a]);
            }
        }
    }
    *cost_out = cta_sum(cost, s_red);
    *invalid_out = cta_sum(invalid, s_red);
}

// In-place Cholesky of the dense SPD matrix S (lower triangle used) and solve S y = b.
// Returns false (to every thread) when a pivot is not positive.
// factor == false: S already holds the factor of an earlier call (further right-hand sides).
__device__ bool window_cholesky_solve(double* S, int n, const double* b, double* y, int* s_flag, bool factor = true) {
    const int tid = threadIdx.x;
    if (tid == 0) *s_flag = 1;
    __syncthreads();
    for (int j = 0; j < n && factor; ++j) {
        if (tid == 0) {
            const double d = S[j * n + j];
            if (!(d > 0.0) || !(d < 1.7976931348623157e308))
                *s_flag = 0;
            else
                S[j * n + j] = sqrt(d);
        }
        __syncthreads();
        if (!*s_flag) break;
        const double dj = S[j * n + j];
        for (int i = j + 1 + tid; i < n; i += WIN_THREADS) S[i * n + j] /= dj;
        __syncthreads();
        const int m = n - j - 1;
        for (int idx = tid; idx < m * m; idx += WIN_THREADS) {
            const int i = j + 1 + idx / m, k = j + 1 + idx % m;
            if (k <= i) S[i * n + k] -= S[i * n + j] * S[k * n + j];
        }
        __syncthreads();
    }
    const bool ok = *s_flag != 0;
    __syncthreads();
    if (!ok) return false;
    if (tid < 32) {
        // forward substitution L z = b (column oriented), then L^T y = z
        for (int i = tid; i < n; i += 32) y[i] = b[i];
        __syncwarp();
        for (int j = 0; j < n; ++j) {
            if (tid == 0) y[j] /= S[j * n + j];
            __syncwarp();
            const double yj = y[j];
            for (int i = j + 1 + tid; i < n; i += 32) y[i] -= S[i * n + j] * yj;
            __syncwarp();
        }
        for (int j = n - 1; j >= 0; --j) {
            if (tid == 0) y[j] /= S[j * n + j];
            __syncwarp();
            const double yj = y[j];
            for (int i = tid; i < j; i += 32) y[i] -= S[j * n + i] * yj;
            __syncwarp();
        }
    }
    __syncthreads();
    return true;
}

enum { CTL_STOP = 0, CTL_COPY_BEST, CTL_NEED_GRAD, CTL_ACCEPT, CTL_COUNT };

template <bool DOGLEG>
__global__ void __launch_bounds__(WIN_THREADS) window_lm_kernel(WinBufs B, int win0) {
    __shared__ double sS[WIN_NMAX * WIN_NMAX];
    __shared__ double s_cam[WIN_PMAX * WIN_CAMV];
    __shared__ double s_bp[WIN_NMAX], s_y[WIN_NMAX], s_scp[WIN_NMAX], s_gp[WIN_NMAX];
    __shared__ double s_pose[3][WIN_PMAX * 12];
    __shared__ int s_free[WIN_PMAX];
    __shared__ double s_red[WIN_THREADS / 32];
    __shared__ int s_ctl[CTL_COUNT];
    __shared__ int s_flag;
    __shared__ double s_radius;
    __shared__ WinDesc D;
    // DOGLEG: clamp(diag(J^T J)) of the pose columns, {c1, c2, model cost change} of the trial step, the model
    __shared__ double s_dp[DOGLEG ? WIN_NMAX : 1], s_c[3];
    __shared__ __align__(8) unsigned char s_model_raw[DOGLEG ? sizeof(DoglegModel) : 8];

    const int tid = threadIdx.x;
    const int win = win0 + blockIdx.x;
    if (tid == 0) D = B.desc[win];
    __syncthreads();
    const WinOpts& O = D.o;
    const int n = 6 * D.n_free;
    Win w;
    w.cam = D.cam;
    w.n_poses = D.n_poses;
    w.n_free = D.n_free;
    w.n_lm = D.n_lm;
    w.n_sun = D.n_sun;
    w.n_prior = D.n_prior;
    w.W_per_obs = D.W_per_obs;
    w.free_idx = s_free;
    w.scp = s_scp;
    w.lm_ptr = B.lm_ptr + D.lmptr_off;
    w.obs_cam = B.obs_cam + D.obs_off;
    w.obs_u = B.obs_u + D.obs_off;
    w.obs_v = B.obs_v + D.obs_off;
    w.obs_d = B.obs_d + D.obs_off;
    w.obs_W = B.obs_W + D.W_off;
    w.sc_l = B.sc_l + 3 * D.lm_off;
    w.gl = B.gl + 3 * D.lm_off;
    w.suns = B.suns + D.sun_off;
    w.priors = B.priors + D.prior_off;
    double* g_poses = B.poses + 12 * D.pose_off;
    double* pts_cur = B.pts + 3 * D.lm_off;
    double* pts_cand = B.pts_cand + 3 * D.lm_off;
    double* pts_best = B.pts_best + 3 * D.lm_off;
    double* yl_g = B.yl + 3 * D.lm_off;
    double* dl_g = B.diag_l + 3 * D.lm_off;
    double* logp = B.logs + D.log_off;
    int cur = 0, cand = 1;  // indices into s_pose; 2 = best

    for (int i = tid; i < 12 * D.n_poses; i += WIN_THREADS) s_pose[0][i] = s_pose[1][i] = s_pose[2][i] = g_poses[i];
    for (int i = tid; i < D.n_poses; i += WIN_THREADS) s_free[i] = B.cam_free[D.pose_off + i];
    for (int i = tid; i < WIN_NMAX; i += WIN_THREADS) s_scp[i] = 1.0;
    for (int i = tid; i < 3 * D.n_lm; i += WIN_THREADS) {
        w.sc_l[i] = 1.0;
        pts_best[i] = pts_cur[i];
    }
    __syncthreads();

    // gradient max norm |x - Plus(x, -g)|_inf and |x| (gradnorm_kernel), from s_gp / gl (scaled)
    auto gradnorm = [&](double* gmax, double* xnorm) {
        double m = 0.0, xn = 0.0;
        if (tid < D.n_poses) {
            const int f = s_free[tid];
            if (f >= 0) {
                const double* x = s_pose[cur] + 12 * tid;
                double eps[6], out[12];
                for (int k = 0; k < 6; ++k) eps[k] = -s_gp[6 * f + k] / s_scp[6 * f + k];
                se3_plus(x, eps, out);
                for (int k = 0; k < 12; ++k) {
                    m = fmax(m, fabs(x[k] - out[k]));
                    xn += x[k] * x[k];
                }
            }
        }
        for (int i = tid; i < 3 * D.n_lm; i += WIN_THREADS) {
            const double x = pts_cur[i];
            const double g = w.gl[i] / w.sc_l[i];
            m = fmax(m, fabs(x - (x - g)));
            xn += x * x;
        }
        *gmax = cta_max(m, s_red);
        *xnorm = sqrt(cta_sum(xn, s_red));
    };
    auto unpack_cam = [&]() {
        for (int i = tid; i < n; i += WIN_THREADS) {
            s_gp[i] = s_cam[(i / 6) * WIN_CAMV + 21 + i % 6];
            s_bp[i] = s_cam[(i / 6) * WIN_CAMV + 27 + i % 6];
        }
        __syncthreads();
    };

    // ---- lm_begin: cost, column norms, gradient at the initial point; Jacobi scaling ----------
    double cost0, inv0;
    window_schur_pass<true>(w, s_pose[cur], pts_cur, LmDiag{0, 0, 0}, sS, n, s_cam, pts_cand, s_red, &cost0, &inv0);
    unpack_cam();
    double gmax, xnorm;
    gradnorm(&gmax, &xnorm);
    if (O.jacobi) {
        for (int i = tid; i < n; i += WIN_THREADS) {
            const int f = i / 6, a = i % 6;
            int t = 0;
            for (int q = 0; q < a; ++q) t += 6 - q;  // index of (a, a) in the packed upper triangle
            s_scp[i] = 1.0 / (1.0 + sqrt(s_cam[f * WIN_CAMV + t]));
        }
        for (int i = tid; i < 3 * D.n_lm; i += WIN_THREADS) w.sc_l[i] = 1.0 / (1.0 + sqrt(pts_cand[i]));
    }
    __syncthreads();

    // ---- LM state (thread 0 owns it; decisions are broadcast through s_ctl) --------------------
    double x_cost = cost0, x_norm = xnorm, gradient_max_norm = gmax, radius = O.r0, decrease_factor = 2.0;
    double minimum_cost = cost0, initial_cost = cost0;
    double se_minimum = cost0, se_current = cost0, se_reference = cost0, se_candidate = cost0, se_acc_ref = 0,
           se_acc_cand = 0;
    int se_nonmono = 0, invalid_steps = 0, iteration = 0, num_successful = 0, num_unsuccessful = 0, total_linear = 0;
    int termination_type = 1, termination_reason = 0, n_rows = 0;
    bool step_ok_prev = false, grad_fresh = true;
    const int max_nonmono = O.nonmono ? O.max_nonmono : 0;
    // DOGLEG (DoglegStrategy): mu and `reuse` are uniform over the CTA (they change on decisions every thread sees);
    // the model and the norm of the last step belong to thread 0
    double mu = 1e-8, dl_step_norm = 0.0;
    bool reuse = false;
    DoglegModel& dmodel = *reinterpret_cast<DoglegModel*>(s_model_raw);
    if (DOGLEG && tid == 0) dmodel = DoglegModel();
    auto push_row = [&](const double* row) {
        if (n_rows < D.log_cap) {
            for (int k = 0; k < CSLAM_LOG_COLS; ++k) logp[CSLAM_LOG_COLS * n_rows + k] = row[k];
            ++n_rows;
        }
    };
    bool failed_start = false;
    if (tid == 0) {
        double row[CSLAM_LOG_COLS] = {0, x_cost, 0, gradient_max_norm, 0, 0, radius, 0, 0, 0};
        push_row(row);
        if (!(fabs(cost0) <= 1.7976931348623157e308)) {  // non-finite cost at the initial point
            termination_type = 2;
            termination_reason = 8;
            s_ctl[CTL_STOP] = 1;
        } else {
            s_ctl[CTL_STOP] = 0;
        }
    }
    __syncthreads();
    failed_start = s_ctl[CTL_STOP] != 0;
    __syncthreads();

    while (!failed_start) {
        // ---- FinalizeIterationAndCheckIfMinimizerCanContinue ---------------------------------
        if (tid == 0) {
            int stop = 0, copy_best = 0;
            if (iteration > 0) {
                if (step_ok_prev) {
                    ++num_successful;
                    if (x_cost < minimum_cost) {
                        minimum_cost = x_cost;
                        copy_best = 1;
                    }
                } else {
                    ++num_unsuccessful;
                }
                step_ok_prev = false;
            }
            if (iteration >= O.max_iter) {
                termination_type = 1;
                termination_reason = 4;
                stop = 1;
            }
            s_ctl[CTL_STOP] = stop;
            s_ctl[CTL_COPY_BEST] = copy_best;
            s_ctl[CTL_NEED_GRAD] = grad_fresh ? 0 : 1;
            s_radius = radius;
        }
        __syncthreads();
        if (s_ctl[CTL_COPY_BEST]) {
            for (int i = tid; i < 12 * D.n_poses; i += WIN_THREADS) s_pose[2][i] = s_pose[cur][i];
            for (int i = tid; i < 3 * D.n_lm; i += WIN_THREADS) pts_best[i] = pts_cur[i];
        }
        if (s_ctl[CTL_STOP]) break;
        const bool need_grad = s_ctl[CTL_NEED_GRAD] != 0;
        LmDiag dg{DOGLEG ? mu : 1.0 / s_radius, O.dmin, O.dmax};
        // ---- Schur build at (x, radius); DOGLEG: at (x, mu), kept while rejected steps shrink the radius ----
        double pass_cost, n_invalid = 0.0;
        auto build_system = [&]() {
            window_schur_pass<false>(w, s_pose[cur], pts_cur, dg, sS, n, s_cam, nullptr, s_red, &pass_cost, &n_invalid);
            unpack_cam();
            // finalize: S_aa += U_aa + D^2 (finalize_kernel)
            for (int i = tid; i < D.n_free * 36; i += WIN_THREADS) {
                const int f = i / 36, a = (i % 36) / 6, b = i % 6;
                const int lo = min(a, b), hi = max(a, b);
                int t = 0;
                for (int q = 0; q < lo; ++q) t += 6 - q;
                double u = s_cam[f * WIN_CAMV + t + (hi - lo)];
                if (a == b) {
                    const double dd = fmin(fmax(u, dg.min_diag), dg.max_diag);
                    if (DOGLEG) s_dp[6 * f + a] = dd;
                    u += dd * dg.inv_radius;
                }
                sS[(6 * f + a) * n + 6 * f + b] += u;
            }
            __syncthreads();
        };
        if (!DOGLEG || !reuse) build_system();
        if (need_grad) gradnorm(&gmax, &xnorm);
        if (tid == 0) {
            int stop = 0;
            if (need_grad) {
                gradient_max_norm = gmax;
                x_norm = xnorm;
                grad_fresh = true;
                if (n_rows > 0 && n_rows <= D.log_cap) logp[CSLAM_LOG_COLS * (n_rows - 1) + 3] = gradient_max_norm;
            }
            if (gradient_max_norm <= O.gtol) {
                termination_type = 0;
                termination_reason = 1;
                stop = 1;
            } else if (radius < O.rmin) {
                termination_type = 0;
                termination_reason = 5;
                stop = 1;
            } else {
                ++iteration;
            }
            s_ctl[CTL_STOP] = stop;
        }
        __syncthreads();
        if (s_ctl[CTL_STOP]) break;
        // ---- LevenbergMarquardtStrategy::ComputeStep: exact solve of the reduced system --------
        bool valid = n_invalid == 0.0;
        int lin_iters = 0;
        double model = 0, ccost = 0, sn = 0, xn = 0, bad = 0;
        if (DOGLEG) {
            // ---- DoglegStrategy::ComputeStep (Engine::dogleg_step) ------------------------------------
            valid = true;
            if (!reuse) {
                // Gauss-Newton solve (J^T J + mu D^2) y = J^T r, mu raised tenfold while the factorisation fails
                bool solved = false, have = true;
                while (mu < 1.0) {
                    if (!have) {
                        dg.inv_radius = mu;
                        build_system();
                        have = true;
                    }
                    bool ok = n_invalid == 0.0;
                    if (ok && n > 0) ok = window_cholesky_solve(sS, n, s_bp, s_y, &s_flag);
                    if (ok) {
                        solved = true;
                        break;
                    }
                    mu *= 10.0;
                    have = false;
                }
                valid = solved;
                if (solved) {
                    lin_iters = 1;
                    // landmark part of y (backsub_kernel), clamp(diag(J^T J)), and the eight sums of
                    // dogleg_products_kernel / dogleg_products_cam_kernel
                    double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
                    for (int j = tid; j < w.n_lm; j += WIN_THREADS) {
                        const uint32_t e0 = w.lm_ptr[j], e1 = w.lm_ptr[j + 1];
                        const double p[3] = {pts_cur[3 * j], pts_cur[3 * j + 1], pts_cur[3 * j + 2]};
                        const double sl[3] = {w.sc_l[3 * j], w.sc_l[3 * j + 1], w.sc_l[3 * j + 2]};
                        double V[6] = {0, 0, 0, 0, 0, 0}, t[3] = {0, 0, 0};
                        WObs o;
                        for (uint32_t e = e0; e < e1; ++e) {
                            w_eval(w, e, s_pose[cur], p, sl, o);
                            double Jy[3] = {0, 0, 0};
                            if (o.f >= 0) {
                                const double* y = s_y + 6 * o.f;
#pragma unroll
                                for (int k = 0; k < 3; ++k)
#pragma unroll
                                    for (int a = 0; a < 6; ++a) Jy[k] += o.Jc[6 * k + a] * y[a];
                            }
#pragma unroll
                            for (int k = 0; k < 3; ++k) {
                                const double a = o.Jp[3 * k], b = o.Jp[3 * k + 1], c = o.Jp[3 * k + 2];
                                V[0] += a * a; V[1] += a * b; V[2] += a * c; V[3] += b * b; V[4] += b * c; V[5] += c * c;
                                const double wv = o.r[k] - Jy[k];
                                t[0] += a * wv; t[1] += b * wv; t[2] += c * wv;
                            }
                        }
                        double d2[3] = {V[0], V[3], V[5]};
                        V[0] += fmin(fmax(V[0], dg.min_diag), dg.max_diag) * dg.inv_radius;
                        V[3] += fmin(fmax(V[3], dg.min_diag), dg.max_diag) * dg.inv_radius;
                        V[5] += fmin(fmax(V[5], dg.min_diag), dg.max_diag) * dg.inv_radius;
                        double Vi[6], yl[3] = {0, 0, 0}, tg[3];
                        if (invert_sym3(V, Vi)) {
                            yl[0] = Vi[0] * t[0] + Vi[1] * t[1] + Vi[2] * t[2];
                            yl[1] = Vi[1] * t[0] + Vi[3] * t[1] + Vi[4] * t[2];
                            yl[2] = Vi[2] * t[0] + Vi[4] * t[1] + Vi[5] * t[2];
                        }
#pragma unroll
                        for (int q = 0; q < 3; ++q) {
                            d2[q] = fmin(fmax(d2[q], dg.min_diag), dg.max_diag);
                            dl_g[3 * j + q] = d2[q];
                            yl_g[3 * j + q] = yl[q];
                            const double g = w.gl[3 * j + q];
                            tg[q] = g / d2[q];
                            acc[0] += g * g / d2[q];
                            acc[1] -= g * yl[q];
                            acc[2] += d2[q] * yl[q] * yl[q];
                        }
                        for (uint32_t e = e0; e < e1; ++e) {
                            w_eval(w, e, s_pose[cur], p, sl, o);
#pragma unroll
                            for (int k = 0; k < 3; ++k) {
                                double jg = o.Jp[3 * k] * tg[0] + o.Jp[3 * k + 1] * tg[1] + o.Jp[3 * k + 2] * tg[2];
                                double jy = o.Jp[3 * k] * yl[0] + o.Jp[3 * k + 1] * yl[1] + o.Jp[3 * k + 2] * yl[2];
                                if (o.f >= 0) {
#pragma unroll
                                    for (int a = 0; a < 6; ++a) {
                                        jg += o.Jc[6 * k + a] * (s_gp[6 * o.f + a] / s_dp[6 * o.f + a]);
                                        jy += o.Jc[6 * k + a] * s_y[6 * o.f + a];
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
                    if (tid < n) {
                        const double g = s_gp[tid], d2 = s_dp[tid], y = s_y[tid];
                        acc[0] += g * g / d2;
                        acc[1] -= g * y;
                        acc[2] += d2 * y * y;
                    }
                    if (tid < w.n_sun + w.n_prior) {
                        double r[6], J[36];
                        int rows, cam;
                        if (tid < w.n_sun) {
                            const SunBlockData& sb = w.suns[tid];
                            cam = int(sb.cam);
                            rows = 2;
                            sun_block(s_pose[cur] + 12 * cam, sb.obs_c, sb.ref_g, sb.W, sb.az_thresh, sb.zen_thresh, r, J);
                            const double sq = r[0] * r[0] + r[1] * r[1];
                            double rho0 = sq, sr = 1.0;
                            if (sb.huber > 0.0) huber_rho(sb.huber, sq, &rho0, &sr);
                            r[0] *= sr;
                            r[1] *= sr;
                            for (int k = 0; k < 12; ++k) J[k] *= sr;
                        } else {
                            const PriorBlockData& pr = w.priors[tid - w.n_sun];
                            cam = int(pr.cam);
                            rows = 6;
                            prior_block(s_pose[cur] + 12 * cam, pr.Tref, pr.W, r, J);
                        }
                        const int f = s_free[cam];
                        if (f >= 0)
                            for (int k = 0; k < rows; ++k) {
                                double jg = 0, jy = 0;
                                for (int a = 0; a < 6; ++a) {
                                    const double js = J[6 * k + a] * s_scp[6 * f + a];
                                    jg += js * (s_gp[6 * f + a] / s_dp[6 * f + a]);
                                    jy += js * s_y[6 * f + a];
                                }
                                acc[3] += jg * jg;
                                acc[4] += jg * jy;
                                acc[5] += jy * jy;
                                acc[6] += jg * r[k];
                                acc[7] += jy * r[k];
                            }
                    }
#pragma unroll
                    for (int k = 0; k < 8; ++k) acc[k] = cta_sum(acc[k], s_red);
                    if (tid == 0) {
                        DoglegModel& m = dmodel;
                        m.G11 = acc[0], m.G12 = acc[1], m.G22 = acc[2];
                        m.JGG = acc[3], m.JGY = acc[4], m.JYY = acc[5], m.JGR = acc[6], m.JYR = acc[7];
                        bool finite = true;
#pragma unroll
                        for (int k = 0; k < 8; ++k) finite = finite && isfinite(acc[k]);
                        s_flag = (finite && m.prepare(O.dogleg_type == 1)) ? 1 : 0;
                    }
                    __syncthreads();
                    valid = s_flag != 0;
                    __syncthreads();
                    if (valid) reuse = true;
                }
            }
            if (valid) {
                // ---- the step for this radius: two coefficients, Y = -c1 g / D^2 + c2 y (dogleg_combine_kernel) ----
                if (tid == 0) {
                    double c1, c2;
                    if (O.dogleg_type == 1)
                        dmodel.subspace(radius, &c1, &c2, &dl_step_norm);
                    else
                        dmodel.traditional(radius, &c1, &c2, &dl_step_norm);
                    s_c[0] = c1;
                    s_c[1] = c2;
                    s_c[2] = dmodel.model_cost_change(c1, c2);
                }
                __syncthreads();
                const double c1 = s_c[0], c2 = s_c[1];
                if (tid < D.n_poses) {
                    const int ff = s_free[tid];
                    const double* x = s_pose[cur] + 12 * tid;
                    double* yv = s_pose[cand] + 12 * tid;
                    if (ff >= 0) {
                        double eps[6], out[12];
                        for (int k = 0; k < 6; ++k) {
                            const double Y = -c1 * s_gp[6 * ff + k] / s_dp[6 * ff + k] + c2 * s_y[6 * ff + k];
                            eps[k] = -Y * s_scp[6 * ff + k];
                            if (isnan(eps[k]) || isinf(eps[k])) bad = 1;
                        }
                        se3_plus(x, eps, out);
                        for (int k = 0; k < 12; ++k) {
                            yv[k] = out[k];
                            sn += (x[k] - out[k]) * (x[k] - out[k]);
                            xn += out[k] * out[k];
                        }
                    } else {
                        for (int k = 0; k < 12; ++k) yv[k] = x[k];
                    }
                }
                __syncthreads();
                // candidate landmarks and the cost there (points_apply_kernel)
                for (int j = tid; j < w.n_lm; j += WIN_THREADS) {
                    const uint32_t e0 = w.lm_ptr[j], e1 = w.lm_ptr[j + 1];
                    double pn[3];
#pragma unroll
                    for (int q = 0; q < 3; ++q) {
                        const double pq = pts_cur[3 * j + q];
                        const double Y = -c1 * w.gl[3 * j + q] / dl_g[3 * j + q] + c2 * yl_g[3 * j + q];
                        const double dl = -Y * w.sc_l[3 * j + q];
                        if (isnan(dl) || isinf(dl)) bad = 1;
                        pn[q] = pq + dl;
                        pts_cand[3 * j + q] = pn[q];
                        sn += (pq - pn[q]) * (pq - pn[q]);
                        xn += pn[q] * pn[q];
                    }
                    for (uint32_t e = e0; e < e1; ++e) {
                        double rc[3];
                        const int c = w.obs_cam[e];
                        const double* Wm = w.W_per_obs ? w.obs_W + 9ll * e : w.obs_W;
                        stereo_block<false>(w.cam, s_pose[cand] + 12 * c, pn, w.obs_u[e], w.obs_v[e], w.obs_d[e], Wm, rc, nullptr,
                                            nullptr);
                        ccost += 0.5 * (rc[0] * rc[0] + rc[1] * rc[1] + rc[2] * rc[2]);
                    }
                }
                // camera-only blocks at the candidate (camonly_step_kernel)
                if (tid < w.n_sun + w.n_prior) {
                    double r[6];
                    if (tid < w.n_sun) {
                        const SunBlockData& sb = w.suns[tid];
                        sun_block(s_pose[cand] + 12 * int(sb.cam), sb.obs_c, sb.ref_g, sb.W, sb.az_thresh, sb.zen_thresh, r, nullptr);
                        const double sq = r[0] * r[0] + r[1] * r[1];
                        double rho0 = sq, sr = 1.0;
                        if (sb.huber > 0.0) huber_rho(sb.huber, sq, &rho0, &sr);
                        ccost += 0.5 * rho0;
                    } else {
                        const PriorBlockData& pr = w.priors[tid - w.n_sun];
                        prior_block(s_pose[cand] + 12 * int(pr.cam), pr.Tref, pr.W, r, nullptr);
                        for (int k = 0; k < 6; ++k) ccost += 0.5 * r[k] * r[k];
                    }
                }
            }
        } else if (valid && n > 0) {
            valid = window_cholesky_solve(sS, n, s_bp, s_y, &s_flag);
            lin_iters = 1;
        }
        // ---- candidate: Plus, back-substitution, model cost change, candidate cost ---------------
        if (!DOGLEG && valid) {
            if (tid < D.n_poses) {
                const int ff = s_free[tid];
                const double* x = s_pose[cur] + 12 * tid;
                double* yv = s_pose[cand] + 12 * tid;
                if (ff >= 0) {
                    double eps[6], out[12];
                    for (int k = 0; k < 6; ++k) {
                        eps[k] = -s_y[6 * ff + k] * s_scp[6 * ff + k];
                        if (isnan(eps[k]) || isinf(eps[k])) bad = 1;
                    }
                    se3_plus(x, eps, out);
                    for (int k = 0; k < 12; ++k) {
                        yv[k] = out[k];
                        sn += (x[k] - out[k]) * (x[k] - out[k]);
                        xn += out[k] * out[k];
                    }
                } else {
                    for (int k = 0; k < 12; ++k) yv[k] = x[k];
                }
            }
            __syncthreads();
            for (int j = tid; j < w.n_lm; j += WIN_THREADS) {
                const uint32_t e0 = w.lm_ptr[j], e1 = w.lm_ptr[j + 1];
                const double p[3] = {pts_cur[3 * j], pts_cur[3 * j + 1], pts_cur[3 * j + 2]};
                const double sl[3] = {w.sc_l[3 * j], w.sc_l[3 * j + 1], w.sc_l[3 * j + 2]};
                double V[6] = {0, 0, 0, 0, 0, 0}, t[3] = {0, 0, 0};
                WObs o;
                for (uint32_t e = e0; e < e1; ++e) {
                    w_eval(w, e, s_pose[cur], p, sl, o);
                    double Jy[3] = {0, 0, 0};
                    if (o.f >= 0) {
                        const double* y = s_y + 6 * o.f;
#pragma unroll
                        for (int k = 0; k < 3; ++k)
#pragma unroll
                            for (int a = 0; a < 6; ++a) Jy[k] += o.Jc[6 * k + a] * y[a];
                    }
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        const double a = o.Jp[3 * k], b = o.Jp[3 * k + 1], c = o.Jp[3 * k + 2];
                        V[0] += a * a; V[1] += a * b; V[2] += a * c; V[3] += b * b; V[4] += b * c; V[5] += c * c;
                        const double wv = o.r[k] - Jy[k];
                        t[0] += a * wv; t[1] += b * wv; t[2] += c * wv;
                    }
                }
                V[0] += fmin(fmax(V[0], dg.min_diag), dg.max_diag) * dg.inv_radius;
                V[3] += fmin(fmax(V[3], dg.min_diag), dg.max_diag) * dg.inv_radius;
                V[5] += fmin(fmax(V[5], dg.min_diag), dg.max_diag) * dg.inv_radius;
                double Vi[6], yl[3] = {0, 0, 0};
                if (invert_sym3(V, Vi)) {
                    yl[0] = Vi[0] * t[0] + Vi[1] * t[1] + Vi[2] * t[2];
                    yl[1] = Vi[1] * t[0] + Vi[3] * t[1] + Vi[4] * t[2];
                    yl[2] = Vi[2] * t[0] + Vi[4] * t[1] + Vi[5] * t[2];
                }
                double pn[3];
#pragma unroll
                for (int q = 0; q < 3; ++q) {
                    const double dl = -yl[q] * sl[q];
                    if (isnan(dl) || isinf(dl)) bad = 1;
                    pn[q] = p[q] + dl;
                    pts_cand[3 * j + q] = pn[q];
                    sn += (p[q] - pn[q]) * (p[q] - pn[q]);
                    xn += pn[q] * pn[q];
                }
                for (uint32_t e = e0; e < e1; ++e) {
                    w_eval(w, e, s_pose[cur], p, sl, o);
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        double m = -(o.Jp[3 * k] * yl[0] + o.Jp[3 * k + 1] * yl[1] + o.Jp[3 * k + 2] * yl[2]);
                        if (o.f >= 0) {
                            const double* y = s_y + 6 * o.f;
#pragma unroll
                            for (int a = 0; a < 6; ++a) m -= o.Jc[6 * k + a] * y[a];
                        }
                        model -= m * (o.r[k] + 0.5 * m);
                    }
                    double rc[3];
                    const int c = w.obs_cam[e];
                    const double* Wm = w.W_per_obs ? w.obs_W + 9ll * e : w.obs_W;
                    stereo_block<false>(w.cam, s_pose[cand] + 12 * c, pn, w.obs_u[e], w.obs_v[e], w.obs_d[e], Wm, rc, nullptr,
                                        nullptr);
                    ccost += 0.5 * (rc[0] * rc[0] + rc[1] * rc[1] + rc[2] * rc[2]);
                }
            }
            // camera-only blocks (camonly_step_kernel)
            if (tid < w.n_sun + w.n_prior) {
                for (int which = 0; which < 2; ++which) {
                    const double* poses = which == 0 ? s_pose[cur] : s_pose[cand];
                    double r[6], J[36], cc = 0;
                    int rows, cam;
                    if (tid < w.n_sun) {
                        const SunBlockData& s = w.suns[tid];
                        cam = int(s.cam);
                        rows = 2;
                        sun_block(poses + 12 * cam, s.obs_c, s.ref_g, s.W, s.az_thresh, s.zen_thresh, r, which == 0 ? J : nullptr);
                        const double sq = r[0] * r[0] + r[1] * r[1];
                        double rho0 = sq, sr = 1.0;
                        if (s.huber > 0.0) huber_rho(s.huber, sq, &rho0, &sr);
                        cc = 0.5 * rho0;
                        r[0] *= sr;
                        r[1] *= sr;
                        if (which == 0)
                            for (int k = 0; k < 12; ++k) J[k] *= sr;
                    } else {
                        const PriorBlockData& pr = w.priors[tid - w.n_sun];
                        cam = int(pr.cam);
                        rows = 6;
                        prior_block(poses + 12 * cam, pr.Tref, pr.W, r, which == 0 ? J : nullptr);
                        for (int k = 0; k < 6; ++k) cc += 0.5 * r[k] * r[k];
                    }
                    if (which == 0) {
                        const int f = s_free[cam];
                        for (int k = 0; k < rows; ++k) {
                            double m = 0;
                            if (f >= 0)
                                for (int a = 0; a < 6; ++a) m -= J[6 * k + a] * s_scp[6 * f + a] * s_y[6 * f + a];
                            model -= m * (r[k] + 0.5 * m);
                        }
                    } else {
                        ccost += cc;
                    }
                }
            }
        }
        // block-wide reductions happen unconditionally so that every thread reaches the barriers
        model = cta_sum(model, s_red);
        ccost = cta_sum(ccost, s_red);
        sn = cta_sum(sn, s_red);
        xn = cta_sum(xn, s_red);
        bad = cta_sum(bad, s_red);
        if (DOGLEG && valid) model = s_c[2];
        if (valid && (bad != 0.0 || !(model > 0.0))) valid = false;
        // ---- step evaluation (thread 0) ---------------------------------------------------------
        if (tid == 0) {
            double row[CSLAM_LOG_COLS] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
            row[0] = iteration;
            row[7] = lin_iters;
            total_linear += lin_iters;
            int stop = 0, accept = 0;
            if (!valid) {
                row[1] = x_cost;
                row[3] = gradient_max_norm;
                if (++invalid_steps >= O.max_invalid) {
                    row[6] = radius;
                    push_row(row);
                    termination_type = 2;
                    termination_reason = 6;
                    stop = 1;
                } else {
                    if (!DOGLEG) {  // (DoglegStrategy::StepIsInvalid raises mu, below)
                        radius = radius / decrease_factor;
                        decrease_factor *= 2.0;
                    }
                    row[6] = radius;
                    push_row(row);
                }
            } else {
                invalid_steps = 0;
                row[8] = 1;
                double cand_cost = ccost;
                if (!(fabs(cand_cost) <= 1.7976931348623157e308)) cand_cost = 1.7976931348623157e308;
                row[4] = sqrt(sn);
                row[2] = x_cost - cand_cost;
                if (row[4] <= O.ptol * (x_norm + O.ptol)) {
                    row[1] = x_cost;
                    row[3] = gradient_max_norm;
                    row[6] = radius;
                    push_row(row);
                    termination_type = 0;
                    termination_reason = 2;
                    stop = 1;
                } else if (fabs(row[2]) <= O.ftol * x_cost) {
                    row[1] = x_cost;
                    row[3] = gradient_max_norm;
                    row[6] = radius;
                    push_row(row);
                    termination_type = 0;
                    termination_reason = 3;
                    stop = 1;
                } else {
                    const double rel = (se_current - cand_cost) / model;
                    const double hist = (se_reference - cand_cost) / (se_acc_ref + model);
                    row[5] = fmax(rel, hist);
                    if (row[5] > O.min_rel) {
                        accept = 1;
                        x_cost = cand_cost;
                        x_norm = sqrt(xn);
                        step_ok_prev = true;
                        grad_fresh = false;
                        row[9] = 1;
                        if (DOGLEG) {
                            // DoglegStrategy::StepAccepted
                            if (row[5] < 0.25) radius *= 0.5;
                            if (row[5] > 0.75) {
                                radius = fmax(radius, 3.0 * dl_step_norm);
                                radius = fmin(radius, O.rmax);
                            }
                        } else {
                            radius = radius / fmax(1.0 / 3.0, 1.0 - pow(2.0 * row[5] - 1.0, 3.0));
                            radius = fmin(O.rmax, radius);
                            decrease_factor = 2.0;
                        }
                        se_current = cand_cost;
                        se_acc_cand += model;
                        se_acc_ref += model;
                        if (se_current < se_minimum) {
                            se_minimum = se_current;
                            se_nonmono = 0;
                            se_candidate = se_current;
                            se_acc_cand = 0;
                        } else {
                            ++se_nonmono;
                            if (se_current > se_candidate) {
                                se_candidate = se_current;
                                se_acc_cand = 0;
                            }
                        }
                        if (se_nonmono == max_nonmono) {
                            se_reference = se_candidate;
                            se_acc_ref = se_acc_cand;
                        }
                    } else if (DOGLEG) {
                        radius *= 0.5;  // DoglegStrategy::StepRejected: the Gauss-Newton and gradient vectors stay valid
                    } else {
                        radius = radius / decrease_factor;
                        decrease_factor *= 2.0;
                    }
                    row[1] = x_cost;
                    row[3] = gradient_max_norm;
                    row[6] = radius;
                    push_row(row);
                }
            }
            s_ctl[CTL_STOP] = stop;
            s_ctl[CTL_ACCEPT] = accept;
        }
        __syncthreads();
        if (s_ctl[CTL_ACCEPT]) {
            const int t = cur;
            cur = cand;
            cand = t;
            double* tp = pts_cur;
            pts_cur = pts_cand;
            pts_cand = tp;
        }
        if (DOGLEG) {
            if (!valid) {
                mu *= 10.0;  // DoglegStrategy::StepIsInvalid
                reuse = false;
            } else if (s_ctl[CTL_ACCEPT]) {
                mu = fmax(1e-8, 2.0 * mu / 10.0);
                reuse = false;
            } else {
                reuse = true;
            }
        }
        const bool stop_now = s_ctl[CTL_STOP] != 0;
        __syncthreads();
        if (stop_now) break;
    }
    // bookkeeping for a step accepted in the last iteration (Engine::lm_iterate tail)
    if (tid == 0) {
        int copy_best = 0;
        if (step_ok_prev && x_cost < minimum_cost) {
            minimum_cost = x_cost;
            copy_best = 1;
        }
        s_ctl[CTL_COPY_BEST] = copy_best;
    }
    __syncthreads();
    if (s_ctl[CTL_COPY_BEST]) {
        for (int i = tid; i < 12 * D.n_poses; i += WIN_THREADS) s_pose[2][i] = s_pose[cur][i];
        for (int i = tid; i < 3 * D.n_lm; i += WIN_THREADS) pts_best[i] = pts_cur[i];
    }
    __syncthreads();
    for (int i = tid; i < 12 * D.n_poses; i += WIN_THREADS) g_poses[i] = s_pose[2][i];
    if (tid == 0) {
        cslam_summary& s = B.summaries[win];
        s.initial_cost = initial_cost;
        s.final_cost = minimum_cost;
        s.num_iterations = iteration;
        s.num_successful_steps = num_successful;
        s.num_unsuccessful_steps = num_unsuccessful;
        s.termination_type = termination_type;
        s.termination_reason = termination_reason;
        s.final_radius = radius;
        s.total_linear_iterations = total_linear;
        s.device_ms = 0.0;
        B.log_rows[win] = n_rows;
    }
}

// cslam_covariance_block for a window: the block of pose `cam` of the inverse of the UNDAMPED reduced camera system
// at the given values (what ceres::Covariance returns for a pose block after marginalising the landmarks;
// dataset_vo_sun.cpp:159-183 feeds it to the next window's prior).  Same passes as the LM kernel — column norms for
// the Jacobi scaling, Schur build with D = 0, in-CTA Cholesky — then six solves against the unit vectors of the block;
// un-scaled on the way out: cov = diag(s) cov_s diag(s).  status: 0 ok, 1 a landmark block is rank deficient, 2 the
// reduced system is not positive definite.
__global__ void __launch_bounds__(WIN_THREADS) window_cov_kernel(WinBufs B, int cam, double* out) {
    __shared__ double sS[WIN_NMAX * WIN_NMAX];
    __shared__ double s_cam[WIN_PMAX * WIN_CAMV];
    __shared__ double s_bp[WIN_NMAX], s_y[WIN_NMAX], s_scp[WIN_NMAX];
    __shared__ double s_pose[WIN_PMAX * 12];
    __shared__ int s_free[WIN_PMAX];
    __shared__ double s_red[WIN_THREADS / 32];
    __shared__ int s_flag;
    __shared__ WinDesc D;
    const int tid = threadIdx.x;
    if (tid == 0) D = B.desc[0];
    __syncthreads();
    const WinOpts& O = D.o;
    const int n = 6 * D.n_free;
    Win w;
    w.cam = D.cam;
    w.n_poses = D.n_poses;
    w.n_free = D.n_free;
    w.n_lm = D.n_lm;
    w.n_sun = D.n_sun;
    w.n_prior = D.n_prior;
    w.W_per_obs = D.W_per_obs;
    w.free_idx = s_free;
    w.scp = s_scp;
    w.lm_ptr = B.lm_ptr + D.lmptr_off;
    w.obs_cam = B.obs_cam + D.obs_off;
    w.obs_u = B.obs_u + D.obs_off;
    w.obs_v = B.obs_v + D.obs_off;
    w.obs_d = B.obs_d + D.obs_off;
    w.obs_W = B.obs_W + D.W_off;
    w.sc_l = B.sc_l + 3 * D.lm_off;
    w.gl = B.gl + 3 * D.lm_off;
    w.suns = B.suns + D.sun_off;
    w.priors = B.priors + D.prior_off;
    const double* pts = B.pts + 3 * D.lm_off;
    double* cn = B.pts_cand + 3 * D.lm_off;
    for (int i = tid; i < 12 * D.n_poses; i += WIN_THREADS) s_pose[i] = B.poses[12 * D.pose_off + i];
    for (int i = tid; i < D.n_poses; i += WIN_THREADS) s_free[i] = B.cam_free[D.pose_off + i];
    for (int i = tid; i < WIN_NMAX; i += WIN_THREADS) s_scp[i] = 1.0;
    for (int i = tid; i < 3 * D.n_lm; i += WIN_THREADS) w.sc_l[i] = 1.0;
    __syncthreads();
    double cost, n_invalid;
    window_schur_pass<true>(w, s_pose, pts, LmDiag{0, 0, 0}, sS, n, s_cam, cn, s_red, &cost, &n_invalid);
    if (O.jacobi) {
        for (int i = tid; i < n; i += WIN_THREADS) {
            const int f = i / 6, a = i % 6;
            int t = 0;
            for (int q = 0; q < a; ++q) t += 6 - q;
            s_scp[i] = 1.0 / (1.0 + sqrt(s_cam[f * WIN_CAMV + t]));
        }
        for (int i = tid; i < 3 * D.n_lm; i += WIN_THREADS) w.sc_l[i] = 1.0 / (1.0 + sqrt(cn[i]));
    }
    __syncthreads();
    const LmDiag dg{0.0, O.dmin, O.dmax};  // radius = infinity: D = 0
    window_schur_pass<false>(w, s_pose, pts, dg, sS, n, s_cam, nullptr, s_red, &cost, &n_invalid);
    for (int i = tid; i < D.n_free * 36; i += WIN_THREADS) {
        const int f = i / 36, a = (i % 36) / 6, b = i % 6;
        const int lo = min(a, b), hi = max(a, b);
        int t = 0;
        for (int q = 0; q < lo; ++q) t += 6 - q;
        sS[(6 * f + a) * n + 6 * f + b] += s_cam[f * WIN_CAMV + t + (hi - lo)];
    }
    __syncthreads();
    int status = n_invalid != 0.0 ? 1 : 0;
    const int f = s_free[cam];
    for (int c = 0; c < 6 && status == 0; ++c) {
        for (int i = tid; i < n; i += WIN_THREADS) s_bp[i] = i == 6 * f + c ? 1.0 : 0.0;
        __syncthreads();
        if (!window_cholesky_solve(sS, n, s_bp, s_y, &s_flag, c == 0)) {
            status = 2;
            break;
        }
        if (tid < 6) out[6 * tid + c] = s_scp[6 * f + tid] * s_y[6 * f + tid] * s_scp[6 * f + c];
        __syncthreads();
    }
    if (tid == 0) out[36] = double(status);
}

template <class T>
void append(std::vector<T>& dst, const T* src, size_t n) {
    dst.insert(dst.end(), src, src + n);
}

}  // namespace

bool Engine::window_eligible(bool any_strategy) const {
    if (opt.window_path == 1) return false;
    if (!any_strategy && opt.trust_region_strategy != 0 && opt.trust_region_strategy != 1) return false;
    // the one-CTA window kernel has no lighting terms, no box projection and no held positions:
    // such problems take the generic engine (also inside cslam_solve_batch)
    if (lighting_in_solve() || bounded || hold_positions) return false;
    if (n_ranks > 1 || opt.linear_solver != 0) return false;
    if (n_poses == 0 || n_poses > uint32_t(WIN_PMAX)) return false;
    if (suns.size() + priors.size() > size_t(WIN_THREADS)) return false;
    if (n_st >= (1ull << 31)) return false;
    return true;
}

// Pack the windows into flat batch arrays, one launch, unpack.  Results are written into each
// engine's caller-owned pose / point arrays, its iteration log and `summaries`.
// Per-call resources of solve_window_batch, pooled process-wide: a pinned staging block (grow-only), a stream
// and two events.  Concurrent callers get different arenas.
struct WindowArena {
    int device = -1;
    char* pinned = nullptr;
    size_t cap = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
};
static std::mutex g_arena_mu;
static std::vector<WindowArena*> g_arena_free;

static WindowArena* window_arena_take(int device, size_t bytes) {
    WindowArena* a = nullptr;
    {
        std::lock_guard<std::mutex> g(g_arena_mu);
        for (size_t i = 0; i < g_arena_free.size(); ++i)
            if (g_arena_free[i]->device == device) {
                a = g_arena_free[i];
                g_arena_free.erase(g_arena_free.begin() + long(i));
                break;
            }
    }
    if (!a) {
        a = new WindowArena;
        a->device = device;
        CSLAM_CUDA(cudaStreamCreateWithFlags(&a->stream, cudaStreamNonBlocking));
        CSLAM_CUDA(cudaEventCreate(&a->ev0));
        CSLAM_CUDA(cudaEventCreate(&a->ev1));
    }
    if (a->cap < bytes) {
        if (a->pinned) cudaFreeHost(a->pinned);
        a->pinned = nullptr;
        a->cap = 0;
        const size_t want = std::max(bytes + bytes / 2, size_t(1) << 20);
        if (cudaMallocHost(reinterpret_cast<void**>(&a->pinned), want) != cudaSuccess) {
            std::lock_guard<std::mutex> g(g_arena_mu);
            g_arena_free.push_back(a);
            throw CudaError("pinned staging allocation failed");
        }
        a->cap = want;
    }
    return a;
}
static void window_arena_give(WindowArena* a) {
    std::lock_guard<std::mutex> g(g_arena_mu);
    g_arena_free.push_back(a);
}

void solve_window_batch(Engine** engines, int n, cslam_summary* summaries, int cov_cam, double* cov_out) {
    // windows the kernel does not take go through the generic engine
    std::vector<int> take;
    const bool cov_mode = cov_cam >= 0;
    if (cov_mode && (n != 1 || !cov_out || !engines[0]->window_eligible(true)))
        throw std::invalid_argument("window covariance: one window-eligible problem expected");
    for (int i = 0; i < n; ++i) {
        Engine& e = *engines[i];
        if (cov_mode || e.window_eligible()) {
            take.push_back(i);
        } else {
            if (e.opt.window_path == 2) throw std::invalid_argument("window_path = 2 but the problem is not window-eligible");
            e.upload();
            e.lm_begin();
            e.lm_iterate(e.opt.max_num_iterations + 1, false, summaries ? &summaries[i] : nullptr);
            e.download();
        }
    }
    if (take.empty()) return;
    // Levenberg-Marquardt windows first, DOGLEG windows behind them: one launch per strategy present
    std::stable_sort(take.begin(), take.end(), [&](int a, int b) {
        return engines[a]->opt.trust_region_strategy < engines[b]->opt.trust_region_strategy;
    });
    int n_lm_windows = 0;
    for (int i : take) n_lm_windows += engines[i]->opt.trust_region_strategy == 0 ? 1 : 0;
    Engine& first = *engines[take[0]];
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0)
        throw CudaError("no CUDA device: the cslam_b200 back end has no CPU fallback");
    CSLAM_CUDA(cudaSetDevice(first.opt.device));

    // CSLAM_WINDOW_TIMING=1: host phases of this call to stderr (packing / staging / device / unpacking)
    static const bool timing = [] {
        const char* e = std::getenv("CSLAM_WINDOW_TIMING");
        return e && std::atoi(e) != 0;
    }();
    using clk = std::chrono::steady_clock;
    const auto t_begin = clk::now();
    auto since = [&](clk::time_point t) { return std::chrono::duration<double, std::micro>(clk::now() - t).count(); };
    const int nw = int(take.size());
    // ---- pass 1: sizes and offsets from the counts alone (a window's landmark arrays are sized by its point count,
    // an upper bound of the landmarks it observes), so that pass 2 can write every window straight into the pinned
    // staging block: no intermediate vectors, no second copy ----
    std::vector<WinDesc> desc(nw);
    size_t n_pose_t = 0, n_pt_t = 0, n_lmptr_t = 0, n_obs_t = 0, n_W_t = 0, n_sun_t = 0, n_prior_t = 0;
    long long log_total = 0;
    for (int wi = 0; wi < nw; ++wi) {
        Engine& e = *engines[take[wi]];
        if (!e.h_poses) throw std::invalid_argument("poses not set");
        if (e.n_st > 0 && (!e.h_points || e.n_points == 0)) throw std::invalid_argument("points not set");
        WinDesc& d = desc[wi];
        std::memset(&d, 0, sizeof(d));
        d.cam = e.cam;
        const cslam_options& o = e.opt;
        d.o = WinOpts{o.max_num_iterations, o.use_nonmonotonic_steps, o.max_consecutive_nonmonotonic_steps,
                      o.max_num_consecutive_invalid_steps, o.jacobi_scaling, o.dogleg_type, o.initial_trust_region_radius,
                      o.max_trust_region_radius, o.min_trust_region_radius, o.min_relative_decrease, o.min_lm_diagonal,
                      o.max_lm_diagonal, o.function_tolerance, o.gradient_tolerance, o.parameter_tolerance};
        d.n_poses = int(e.n_poses);
        d.pose_off = (long long)n_pose_t;
        d.lm_off = (long long)n_pt_t;
        d.lmptr_off = (long long)n_lmptr_t;
        d.obs_off = (long long)n_obs_t;
        d.W_per_obs = e.st_W_per_obs;
        d.W_off = (long long)n_W_t;
        d.n_sun = int(e.suns.size());
        d.n_prior = int(e.priors.size());
        d.sun_off = (long long)n_sun_t;
        d.prior_off = (long long)n_prior_t;
        d.log_cap = std::min(std::max(o.max_num_iterations, 0), 254) + 2;
        d.log_off = log_total * CSLAM_LOG_COLS;
        log_total += d.log_cap;
        n_pose_t += e.n_poses;
        n_pt_t += e.n_points;
        n_lmptr_t += size_t(e.n_points) + 1;
        n_obs_t += e.n_st;
        n_W_t += e.st_W_per_obs ? 9 * size_t(e.n_st) : 9;
        n_sun_t += e.suns.size();
        n_prior_t += e.priors.size();
    }
    const double us_sizes = since(t_begin);

    // One pinned staging block, one device block, ONE host-to-device copy and ONE device-to-host copy per
    // batch (a window is a few KB: a dozen separate copies and allocations cost more than the kernel).
    // Layout (256-byte aligned pieces):  [inputs ... | poses | best points | summaries | logs | log rows | scratch]
    // The copy up covers inputs + poses, the copy back poses .. log rows.
    size_t cursor = 0;
    auto place = [&](size_t bytes) {
        const size_t off = cursor;
        cursor = (cursor + std::max<size_t>(bytes, 8) + 255) & ~size_t(255);
        return off;
    };
    const size_t pts_bytes = std::max<size_t>(n_pt_t, 1) * 24;
    const size_t o_desc = place(desc.size() * sizeof(WinDesc)), o_camfree = place(n_pose_t * sizeof(int)), o_pts = place(pts_bytes),
                 o_lmptr = place(n_lmptr_t * sizeof(uint32_t)), o_ocam = place(n_obs_t), o_ou = place(n_obs_t * 8),
                 o_ov = place(n_obs_t * 8), o_od = place(n_obs_t * 8), o_oW = place(n_W_t * 8),
                 o_suns = place(n_sun_t * sizeof(SunBlockData)), o_priors = place(n_prior_t * sizeof(PriorBlockData));
    const size_t o_poses = place(n_pose_t * 96);
    const size_t up_bytes = cursor;
    const size_t o_best = place(pts_bytes), o_sum = place(size_t(nw) * sizeof(cslam_summary)),
                 o_logs = place(size_t(log_total) * CSLAM_LOG_COLS * 8), o_rows = place(size_t(nw) * sizeof(int));
    const size_t o_cov = place(37 * sizeof(double));  // covariance mode: the block and a status word
    const size_t down_end = cursor;
    const size_t o_cand = place(pts_bytes), o_scl = place(pts_bytes), o_gl = place(pts_bytes);
    const bool any_dogleg = n_lm_windows < nw;
    const size_t o_yl = place(any_dogleg ? pts_bytes : 8), o_dl = place(any_dogleg ? pts_bytes : 8);
    const size_t total = cursor;

    WindowArena* arena = window_arena_take(first.opt.device, down_end);
    struct Give {
        WindowArena* a;
        ~Give() { window_arena_give(a); }
    } give{arena};
    cudaStream_t stream = arena->stream;
    cudaEvent_t ev0 = arena->ev0, ev1 = arena->ev1;
    char* hp = arena->pinned;

    // ---- pass 2: every window written in place ----
    int* cam_free = reinterpret_cast<int*>(hp + o_camfree);
    double* h_pts = reinterpret_cast<double*>(hp + o_pts);
    uint32_t* h_lmptr = reinterpret_cast<uint32_t*>(hp + o_lmptr);
    uint8_t* h_ocam = reinterpret_cast<uint8_t*>(hp + o_ocam);
    double* h_ou = reinterpret_cast<double*>(hp + o_ou);
    double* h_ov = reinterpret_cast<double*>(hp + o_ov);
    double* h_od = reinterpret_cast<double*>(hp + o_od);
    double* h_oW = reinterpret_cast<double*>(hp + o_oW);
    SunBlockData* h_suns = reinterpret_cast<SunBlockData*>(hp + o_suns);
    PriorBlockData* h_priors = reinterpret_cast<PriorBlockData*>(hp + o_priors);
    double* h_poses = reinterpret_cast<double*>(hp + o_poses);
    std::vector<uint32_t> lm_user(std::max<size_t>(n_pt_t, 1));  // [lm_off + a] -> the caller's point index
    auto pack_range = [&](int w_lo, int w_hi) {
    std::vector<uint32_t> cnt, slot;                              // per-window scratch, reused
    std::vector<uint8_t> used;
    for (int wi = w_lo; wi < w_hi; ++wi) {
        Engine& e = *engines[take[wi]];
        WinDesc& d = desc[wi];
        // which blocks exist and which are free (dataset_vo.cpp:40-62)
        used.assign(e.n_poses, 0);
        cnt.assign(e.n_points, 0);
        for (uint64_t i = 0; i < e.n_st; ++i) {
            if (e.st_cam[i] >= e.n_poses || e.st_pt[i] >= e.n_points)
                throw std::invalid_argument("stereo block index out of range");
            used[e.st_cam[i]] = 1;
            cnt[e.st_pt[i]]++;
        }
        for (auto& sb : e.suns) used[sb.cam] = 1;
        for (auto& pb : e.priors) used[pb.cam] = 1;
        int nf = 0;
        int* cf = cam_free + d.pose_off;
        for (uint32_t k = 0; k < e.n_poses; ++k) cf[k] = (used[k] && !e.pose_const[k]) ? nf++ : -1;
        d.n_free = nf;
        std::memcpy(h_poses + 12 * size_t(d.pose_off), e.h_poses, 96 * size_t(e.n_poses));
        // landmark-major observation lists, landmarks in the caller's point order; `slot[j]` becomes the write cursor
        // of point j's list
        slot.resize(e.n_points);
        uint32_t* lp = h_lmptr + d.lmptr_off;
        uint32_t* lu = lm_user.data() + d.lm_off;
        double* wp = h_pts + 3 * size_t(d.lm_off);
        uint32_t acc = 0, n_lm = 0;
        for (uint32_t j = 0; j < e.n_points; ++j)
            if (cnt[j]) {
                lu[n_lm] = j;
                lp[n_lm] = acc;
                slot[j] = acc;
                acc += cnt[j];
                wp[3 * n_lm] = e.h_points[3 * size_t(j)];
                wp[3 * n_lm + 1] = e.h_points[3 * size_t(j) + 1];
                wp[3 * n_lm + 2] = e.h_points[3 * size_t(j) + 2];
                ++n_lm;
            }
        lp[n_lm] = acc;
        d.n_lm = int(n_lm);
        const size_t base = size_t(d.obs_off);
        double* Wdst = h_oW + d.W_off;
        if (!e.st_W_per_obs) {
            if (e.n_st)
                std::memcpy(Wdst, e.st_W, 72);
            else
                std::memset(Wdst, 0, 72);
        }
        for (uint64_t i = 0; i < e.n_st; ++i) {
            const size_t rel = slot[e.st_pt[i]]++;
            const size_t pos = base + rel;
            h_ocam[pos] = uint8_t(e.st_cam[i]);
            h_ou[pos] = e.st_uvd[3 * i];
            h_ov[pos] = e.st_uvd[3 * i + 1];
            h_od[pos] = e.st_uvd[3 * i + 2];
            if (e.st_W_per_obs) std::memcpy(Wdst + 9 * rel, e.st_W + 9 * i, 72);
        }
        if (!e.suns.empty()) std::memcpy(h_suns + d.sun_off, e.suns.data(), e.suns.size() * sizeof(SunBlockData));
        if (!e.priors.empty()) std::memcpy(h_priors + d.prior_off, e.priors.data(), e.priors.size() * sizeof(PriorBlockData));
    }
    };
    // (packing a 256-window batch on four threads was measured slower than this loop — 1.42 vs 1.03 ms per call:
    // spawning the threads costs more than the 0.35 ms they share)
    pack_range(0, nw);
    std::memcpy(hp + o_desc, desc.data(), desc.size() * sizeof(WinDesc));
    const double us_pack = since(t_begin);
    const double us_stage = since(t_begin);
    {
        DBuf<uint8_t> d_all;
        d_all.alloc(total, stream);
        struct Free {
            DBuf<uint8_t>& d;
            cudaStream_t s;
            ~Free() { d.release_async(s); }
        } free_all{d_all, stream};
        uint8_t* dp = d_all.p;
        CSLAM_CUDA(cudaMemcpyAsync(dp, hp, up_bytes, cudaMemcpyHostToDevice, stream));
        WinBufs B;
        B.desc = reinterpret_cast<WinDesc*>(dp + o_desc);
        B.poses = reinterpret_cast<double*>(dp + o_poses);
        B.cam_free = reinterpret_cast<int*>(dp + o_camfree);
        B.pts = reinterpret_cast<double*>(dp + o_pts);
        B.pts_cand = reinterpret_cast<double*>(dp + o_cand);
        B.pts_best = reinterpret_cast<double*>(dp + o_best);
        B.lm_ptr = reinterpret_cast<uint32_t*>(dp + o_lmptr);
        B.obs_cam = dp + o_ocam;
        B.obs_u = reinterpret_cast<double*>(dp + o_ou);
        B.obs_v = reinterpret_cast<double*>(dp + o_ov);
        B.obs_d = reinterpret_cast<double*>(dp + o_od);
        B.obs_W = reinterpret_cast<double*>(dp + o_oW);
        B.sc_l = reinterpret_cast<double*>(dp + o_scl);
        B.gl = reinterpret_cast<double*>(dp + o_gl);
        B.yl = reinterpret_cast<double*>(dp + o_yl);
        B.diag_l = reinterpret_cast<double*>(dp + o_dl);
        B.suns = reinterpret_cast<SunBlockData*>(dp + o_suns);
        B.priors = reinterpret_cast<PriorBlockData*>(dp + o_priors);
        B.summaries = reinterpret_cast<cslam_summary*>(dp + o_sum);
        B.logs = reinterpret_cast<double*>(dp + o_logs);
        B.log_rows = reinterpret_cast<int*>(dp + o_rows);
        if (cov_mode) {
            const Engine& e = *engines[take[0]];
            if (uint32_t(cov_cam) >= e.n_poses) throw std::invalid_argument("covariance: pose index out of range");
            if (cam_free[size_t(cov_cam)] < 0) throw std::invalid_argument("covariance: the pose block is constant");
            window_cov_kernel<<<1, WIN_THREADS, 0, stream>>>(B, cov_cam, reinterpret_cast<double*>(dp + o_cov));
            g_kernel_launches.fetch_add(1, std::memory_order_relaxed);
            CSLAM_CUDA(cudaGetLastError());
            CSLAM_CUDA(cudaMemcpyAsync(hp + o_cov, dp + o_cov, 37 * sizeof(double), cudaMemcpyDeviceToHost, stream));
            CSLAM_CUDA(cudaStreamSynchronize(stream));
            const double* r = reinterpret_cast<const double*>(hp + o_cov);
            if (r[36] == 1.0) throw std::domain_error("covariance: a landmark block is rank deficient");
            if (r[36] != 0.0) throw std::domain_error("covariance: the reduced camera system is not positive definite");
            std::memcpy(cov_out, r, 36 * sizeof(double));
            return;
        }
        CSLAM_CUDA(cudaEventRecord(ev0, stream));
        if (n_lm_windows > 0) {
            window_lm_kernel<false><<<n_lm_windows, WIN_THREADS, 0, stream>>>(B, 0);
            g_kernel_launches.fetch_add(1, std::memory_order_relaxed);
        }
        if (any_dogleg) {
            window_lm_kernel<true><<<nw - n_lm_windows, WIN_THREADS, 0, stream>>>(B, n_lm_windows);
            g_kernel_launches.fetch_add(1, std::memory_order_relaxed);
        }
        CSLAM_CUDA(cudaGetLastError());
        CSLAM_CUDA(cudaEventRecord(ev1, stream));
        CSLAM_CUDA(cudaMemcpyAsync(hp + o_poses, dp + o_poses, down_end - o_poses, cudaMemcpyDeviceToHost, stream));
        CSLAM_CUDA(cudaStreamSynchronize(stream));
        const double us_device = since(t_begin);
        float ms = 0;
        CSLAM_CUDA(cudaEventElapsedTime(&ms, ev0, ev1));
        const double* poses_out = reinterpret_cast<const double*>(hp + o_poses);
        const double* pts_out = reinterpret_cast<const double*>(hp + o_best);
        const cslam_summary* sums_in = reinterpret_cast<const cslam_summary*>(hp + o_sum);
        const double* logs = reinterpret_cast<const double*>(hp + o_logs);
        const int* log_rows = reinterpret_cast<const int*>(hp + o_rows);
        std::vector<cslam_summary> sums(sums_in, sums_in + nw);
        for (int wi = 0; wi < nw; ++wi) {
            Engine& e = *engines[take[wi]];
            const WinDesc& d = desc[wi];
            for (uint32_t k = 0; k < e.n_poses; ++k)
                if (cam_free[size_t(d.pose_off) + k] >= 0)
                    std::memcpy(e.h_poses + 12 * size_t(k), &poses_out[12 * (size_t(d.pose_off) + k)], 96);
            for (int a = 0; a < d.n_lm; ++a)
                std::memcpy(e.h_points + 3 * size_t(lm_user[size_t(d.lm_off) + a]), &pts_out[3 * (size_t(d.lm_off) + a)], 24);
            e.log.clear();
            for (int r = 0; r < log_rows[wi]; ++r) {
                LmRow row;
                std::memcpy(row.v, &logs[size_t(d.log_off) + size_t(r) * CSLAM_LOG_COLS], sizeof(row.v));
                e.log.push_back(row);
            }
            sums[wi].device_ms = ms;  // the whole batch is one launch
            e.set_window_summary(sums[wi]);
            e.prof.ms[CSLAM_K_WINDOW] += wi == 0 ? ms : 0.0;
            e.prof.launches[CSLAM_K_WINDOW] += wi == 0 ? 1 : 0;
            if (summaries) summaries[take[wi]] = sums[wi];
        }
        if (timing)
            std::fprintf(stderr, "[cslam window batch] %d windows: sizes %.0f us, pack (in place) %.0f us, alloc+H2D+kernel(%.0f us)+D2H "
                                 "%.0f us, unpack %.0f us\n", nw, us_sizes, us_pack - us_sizes, ms * 1e3, us_device - us_stage,
                         since(t_begin) - us_device);
    }
}

}  // namespace cslam
