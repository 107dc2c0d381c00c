// K2 fast path — fused residual/Jacobian + Schur elimination for landmarks grouped by identical
// camera list (sm_100a, FP64).
//
// For a group of G landmarks that are all seen by the same L cameras, the landmark elimination
//      S_ab -= sum_j W_aj V_j^-1 W_bj^T ,      b_a = g_a - sum_j W_aj V_j^-1 g_lj
// is a rank-3G symmetric update of one (6L x 6L) tile:  with V_j = C_j C_j^T (Cholesky) and
// Z_aj = W_aj C_j^-T it reads  S_tile -= Z Z^T  (SYRK-shaped).  One CTA owns a slice of a group:
//   * warp 0 (producer): lane = (landmark-in-round q, camera slot i); the slot's pose and Jacobi
//     scaling stay in registers; per observation it evaluates r, Jc, Jp in closed form, forms
//     W = Jc^T Jp and Z = W C^-T, writes Z to a double-buffered shared tile, and accumulates the
//     slot's U_aa = sum Jc^T Jc, gradient and reduced right-hand side in registers;
//   * warps 1.. (consumers): thread = camera-slot pair (a <= b); its 6x6 block of the tile lives
//     in 36 registers (output stationary) and receives  Z_a Z_b^T  for every landmark of the
//     slice; it is flushed to the global block-sparse S with one RED per entry per slice.
// Camera poses of the slice are staged in shared memory with one TMA bulk copy (contiguous camera
// lists) or L bulk copies, completion on an mbarrier.  Observations are stored slot-major inside
// a group, so every global load of u, v, d is coalesced.
#include <cstdlib>

#include "kernels.cuh"

namespace cslam {

namespace {

constexpr int kSMs = 148;
constexpr int GZ_DOUBLES = 32 * 18;  // one Z tile: up to 32 (landmark, slot) lanes x 18

__device__ __forceinline__ const double* gW_ptr(const DevView& v, long long e) {
    return v.W_per_obs ? v.obs_W + 9 * e : v.obs_W;
}

template <int CW>
__global__ void __launch_bounds__(32 * (1 + CW))
    schur_grouped_kernel(DevView v, GroupView gv, int item_lo, int item_hi, LmDiag dg, double* __restrict__ S,
                         double* __restrict__ Bdiag, double* __restrict__ bp, double* __restrict__ gp,
                         double* __restrict__ gl, double* __restrict__ scal) {
    __shared__ __align__(128) double s_pose[kGroupLmax * 12];
    __shared__ double s_sp[kGroupLmax * 6];
    __shared__ double s_A[kItemMax * 9];
    __shared__ __align__(16) double s_Z[2 * GZ_DOUBLES];
    __shared__ double s_redsum[32];
    __shared__ int s_free[kGroupLmax];
    __shared__ int s_blk[kGroupLmax * (kGroupLmax + 1) / 2];
    __shared__ __align__(8) uint64_t s_bar;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool is_producer = warp == 0;
    const int cid = tid - 32;  // consumer index
    if (tid == 0) {
        mbar_init(&s_bar, 1);
        fence_mbar_init();
    }
    __syncthreads();
    uint32_t phase = 0;
    double cost = 0.0;
    double Wsh[9];
    if (!v.W_per_obs) {
#pragma unroll
        for (int k = 0; k < 9; ++k) Wsh[k] = v.obs_W[k];
    }

    for (int w = item_lo + blockIdx.x; w < item_hi; w += gridDim.x) {
        const int g = gv.item_group[w];
        const int L = gv.g_L[g], G = gv.g_G[g];
        const int j0 = gv.item_j0[w], nj = gv.item_n[w];
        const int lm0 = gv.g_lm0[g] + j0;
        const long long obs0 = (long long)gv.g_obs0[g] + j0;
        const int* __restrict__ cams = gv.g_cams + gv.g_off[g];
        const int* __restrict__ blk = gv.g_blk + gv.g_blk_off[g];
        const int P = L * (L + 1) / 2;

        // ---- stage the slice's cameras: free index, scaling, pair->block table, poses (TMA) ----
        if (tid < L) {
            const int f = v.cam_free[cams[tid]];
            s_free[tid] = f;
#pragma unroll
            for (int k = 0; k < 6; ++k) s_sp[6 * tid + k] = f >= 0 ? v.sc_p[6ll * f + k] : 0.0;
        }
        for (int k = tid; k < P; k += blockDim.x) s_blk[k] = blk[k];
        if (tid == 0) {
            const int c0 = cams[0];
            mbar_expect_tx(&s_bar, L * 96);
            if (cams[L - 1] - c0 == L - 1) {
                tma_load_1d(s_pose, v.poses + 12ll * c0, L * 96, &s_bar);
            } else {
                for (int i = 0; i < L; ++i) tma_load_1d(s_pose + 12 * i, v.poses + 12ll * cams[i], 96, &s_bar);
            }
        }
        mbar_wait(&s_bar, phase);
        phase ^= 1;
        __syncthreads();

        // ---- pass 1: V_j = sum Jp^T Jp + D^2, Cholesky, A = C^-1, t = A g_l ----
        for (int jl = tid; jl < nj; jl += blockDim.x) {
            const long long j = lm0 + jl;
            const double p[3] = {v.points[3 * j], v.points[3 * j + 1], v.points[3 * j + 2]};
            const double sl[3] = {v.sc_l[3 * j], v.sc_l[3 * j + 1], v.sc_l[3 * j + 2]};
            double V[6] = {0, 0, 0, 0, 0, 0}, gq[3] = {0, 0, 0};
            for (int i = 0; i < L; ++i) {
                const long long e = obs0 + (long long)i * G + jl;
                double r[3], Jp[9];
                stereo_block_point(v.cam, s_pose + 12 * i, p, v.obs_u[e], v.obs_v[e], v.obs_d[e],
                                   v.W_per_obs ? v.obs_W + 9 * e : Wsh, r, Jp);
                cost += 0.5 * (r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const double a = Jp[3 * k] * sl[0], b = Jp[3 * k + 1] * sl[1], c = Jp[3 * k + 2] * sl[2];
                    V[0] += a * a;
                    V[1] += a * b;
                    V[2] += a * c;
                    V[3] += b * b;
                    V[4] += b * c;
                    V[5] += c * c;
                    gq[0] += a * r[k];
                    gq[1] += b * r[k];
                    gq[2] += c * r[k];
                }
            }
            V[0] += fmin(fmax(V[0], dg.min_diag), dg.max_diag) * dg.inv_radius;
            V[3] += fmin(fmax(V[3], dg.min_diag), dg.max_diag) * dg.inv_radius;
            V[5] += fmin(fmax(V[5], dg.min_diag), dg.max_diag) * dg.inv_radius;
            // V = C C^T, C lower triangular
            double A[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
            bool pd = V[0] > 0.0;
            if (pd) {
                const double c00 = sqrt(V[0]);
                const double c10 = V[1] / c00, c20 = V[2] / c00;
                const double d1 = V[3] - c10 * c10;
                pd = d1 > 0.0;
                if (pd) {
                    const double c11 = sqrt(d1);
                    const double c21 = (V[4] - c20 * c10) / c11;
                    const double d2 = V[5] - c20 * c20 - c21 * c21;
                    pd = d2 > 0.0 && d2 < 1.7976931348623157e308;
                    if (pd) {
                        const double c22 = sqrt(d2);
                        const double a00 = 1.0 / c00, a11 = 1.0 / c11, a22 = 1.0 / c22;
                        const double a10 = -c10 * a00 * a11;
                        const double a21 = -c21 * a11 * a22;
                        const double a20 = -(c20 * a00 + c21 * a10) * a22;
                        A[0] = a00;
                        A[1] = a10;
                        A[2] = a11;
                        A[3] = a20;
                        A[4] = a21;
                        A[5] = a22;
                        A[6] = a00 * gq[0];
                        A[7] = a10 * gq[0] + a11 * gq[1];
                        A[8] = a20 * gq[0] + a21 * gq[1] + a22 * gq[2];
                    }
                }
            }
            if (!pd) red_add(&scal[SC_INVALID], 1.0);
#pragma unroll
            for (int k = 0; k < 9; ++k) s_A[9 * jl + k] = A[k];
            gl[3 * j] = gq[0];
            gl[3 * j + 1] = gq[1];
            gl[3 * j + 2] = gq[2];
        }
        __syncthreads();

        // ---- main loop: producer forms Z, consumers accumulate Z_a Z_b^T ----
        const int R = 32 / L;  // landmarks per round
        const int rounds = (nj + R - 1) / R;
        // producer lane state
        const int pi = lane / R, pq = lane - pi * R;
        const bool p_active = is_producer && lane < R * L;
        int pf = -1;
        double pose[12], sp[6];
        double U[21], gpa[6], bpa[6];
        if (p_active) {
            pf = s_free[pi];
#pragma unroll
            for (int k = 0; k < 12; ++k) pose[k] = s_pose[12 * pi + k];
#pragma unroll
            for (int k = 0; k < 6; ++k) sp[k] = s_sp[6 * pi + k];
        }
#pragma unroll
        for (int k = 0; k < 21; ++k) U[k] = 0.0;
#pragma unroll
        for (int k = 0; k < 6; ++k) gpa[k] = bpa[k] = 0.0;
        // consumer thread state: pair (a <= b), row-major enumeration of the upper triangle
        int ca = 0, cb = 0;
        const bool c_active = !is_producer && cid < P;
        if (c_active) {
            int rem = cid;
            while (rem >= L - ca) {
                rem -= L - ca;
                ++ca;
            }
            cb = ca + rem;
        }
        double M[36];
#pragma unroll
        for (int k = 0; k < 36; ++k) M[k] = 0.0;

        for (int rd = 0; rd <= rounds; ++rd) {
            if (is_producer) {
                if (rd < rounds && p_active) {
                    double* zt = s_Z + (rd & 1) * GZ_DOUBLES + (pq * L + pi) * 18;
                    const int jl = rd * R + pq;
                    double Z[18];
#pragma unroll
                    for (int k = 0; k < 18; ++k) Z[k] = 0.0;
                    if (jl < nj && pf >= 0) {
                        const long long j = lm0 + jl;
                        const long long e = obs0 + (long long)pi * G + jl;
                        const double p[3] = {v.points[3 * j], v.points[3 * j + 1], v.points[3 * j + 2]};
                        const double sl[3] = {v.sc_l[3 * j], v.sc_l[3 * j + 1], v.sc_l[3 * j + 2]};
                        double r[3], Jc[18], Jp[9];
                        stereo_block<true>(v.cam, pose, p, v.obs_u[e], v.obs_v[e], v.obs_d[e],
                                           v.W_per_obs ? v.obs_W + 9 * e : Wsh, r, Jc, Jp);
#pragma unroll
                        for (int k = 0; k < 3; ++k) {
#pragma unroll
                            for (int q = 0; q < 3; ++q) Jp[3 * k + q] *= sl[q];
#pragma unroll
                            for (int q = 0; q < 6; ++q) Jc[6 * k + q] *= sp[q];
                        }
                        const double* A = s_A + 9 * jl;
                        const double a00 = A[0], a10 = A[1], a11 = A[2], a20 = A[3], a21 = A[4], a22 = A[5];
                        const double t0 = A[6], t1 = A[7], t2 = A[8];
                        int u = 0;
#pragma unroll
                        for (int a = 0; a < 6; ++a) {
                            const double w0 = Jc[a] * Jp[0] + Jc[6 + a] * Jp[3] + Jc[12 + a] * Jp[6];
                            const double w1 = Jc[a] * Jp[1] + Jc[6 + a] * Jp[4] + Jc[12 + a] * Jp[7];
                            const double w2 = Jc[a] * Jp[2] + Jc[6 + a] * Jp[5] + Jc[12 + a] * Jp[8];
                            // Z = W A^T (A lower triangular)
                            const double z0 = w0 * a00;
                            const double z1 = w0 * a10 + w1 * a11;
                            const double z2 = w0 * a20 + w1 * a21 + w2 * a22;
                            Z[3 * a] = z0;
                            Z[3 * a + 1] = z1;
                            Z[3 * a + 2] = z2;
                            const double ga = Jc[a] * r[0] + Jc[6 + a] * r[1] + Jc[12 + a] * r[2];
                            gpa[a] += ga;
                            bpa[a] += ga - (z0 * t0 + z1 * t1 + z2 * t2);
#pragma unroll
                            for (int b = a; b < 6; ++b, ++u)
                                U[u] += Jc[a] * Jc[b] + Jc[6 + a] * Jc[6 + b] + Jc[12 + a] * Jc[12 + b];
                        }
                    }
#pragma unroll
                    for (int k = 0; k < 18; k += 2) *reinterpret_cast<double2*>(zt + k) = make_double2(Z[k], Z[k + 1]);
                }
            } else if (rd > 0 && c_active) {
                const double* zt = s_Z + ((rd - 1) & 1) * GZ_DOUBLES;
                const int nval = min(R, nj - (rd - 1) * R);
                for (int jj = 0; jj < nval; ++jj) {
                    const double* za = zt + (jj * L + ca) * 18;
                    const double* zb = zt + (jj * L + cb) * 18;
                    double B[18];
#pragma unroll
                    for (int k = 0; k < 18; k += 2) {
                        const double2 t = *reinterpret_cast<const double2*>(zb + k);
                        B[k] = t.x;
                        B[k + 1] = t.y;
                    }
#pragma unroll
                    for (int p = 0; p < 6; ++p) {
                        const double x0 = za[3 * p], x1 = za[3 * p + 1], x2 = za[3 * p + 2];
#pragma unroll
                        for (int q = 0; q < 6; ++q)
                            M[6 * p + q] += x0 * B[3 * q] + x1 * B[3 * q + 1] + x2 * B[3 * q + 2];
                    }
                }
            }
            __syncthreads();
        }

        // ---- flush: pair blocks (consumers), camera diagonal / gradients (producer) ----
        if (c_active) {
            const int e = s_blk[cid];
            if (e >= 0) {
                double* Bk = S + 36ll * e;
                if (ca == cb) {
                    // diagonal block: upper triangle only, finalize mirrors it
#pragma unroll
                    for (int p = 0; p < 6; ++p)
#pragma unroll
                        for (int q = p; q < 6; ++q) red_add(&Bk[6 * p + q], -M[6 * p + q]);
                } else {
#pragma unroll
                    for (int k = 0; k < 36; ++k) red_add(&Bk[k], -M[k]);
                }
            }
        }
        if (is_producer) {
            // reduce over the R lanes that share a slot (consecutive lanes), then one RED each
            double* red = s_Z;  // all consumer reads of s_Z are behind the last barrier
            if (p_active) {
#pragma unroll
                for (int k = 0; k < 21; ++k) red[lane * 33 + k] = U[k];
#pragma unroll
                for (int k = 0; k < 6; ++k) {
                    red[lane * 33 + 21 + k] = gpa[k];
                    red[lane * 33 + 27 + k] = bpa[k];
                }
            }
            __syncwarp();
            if (p_active && pq == 0 && pf >= 0) {
                double acc[33];
#pragma unroll
                for (int k = 0; k < 33; ++k) acc[k] = red[lane * 33 + k];
                for (int q = 1; q < R; ++q)
#pragma unroll
                    for (int k = 0; k < 33; ++k) acc[k] += red[(lane + q) * 33 + k];
                double* Bd = Bdiag + 36ll * pf;
                int u = 0;
#pragma unroll
                for (int a = 0; a < 6; ++a) {
#pragma unroll
                    for (int b = a; b < 6; ++b, ++u) red_add(&Bd[6 * a + b], acc[u]);
                    red_add(&gp[6ll * pf + a], acc[21 + a]);
                    red_add(&bp[6ll * pf + a], acc[27 + a]);
                }
            }
        }
        __syncthreads();
    }
    block_atomic_sum(cost, &scal[SC_COST], s_redsum);
}


// ---------------------------------------------------------------------------------------------
// Variant for track length L <= 10.
// Every thread is producer AND consumer: in one step of the software pipeline a thread first
// evaluates one observation of batch n+1 (camera slot fixed per thread: pose and scaling stay in
// registers) and writes its Z = W C^-T to one half of a double-buffered shared tile; then the
// CTA's four warps apply batch n to the tile with FP64 tensor-core MMAs:
//      S_tile (6L x 6L, upper triangle) -= Z Z^T ,   Z = [6L rows] x [3 columns per landmark]
// as DMMA m8n8k4 (mma.sync.aligned.m8n8k4.row.col.f64): the tile is cut into ceil(6L/8)^2 8x8
// MMA tiles (they straddle the 6x6 camera blocks; the flush maps entries back), the upper
// triangle of MMA tiles is dealt to the warps as two triangular and two rectangular sets so a
// warp reuses each fragment 2..4 times, and because the update is symmetric the A fragment of
// row tile i IS the B fragment of column tile i (one shared-memory word per thread per tile).
// DMMA has the DFMA peak on B200 (scripts/micro/dmma_micro.cu: 36.7 vs 36.1 TFLOP/s, one pipe),
// so the gain is issue slots and shared-memory traffic: 1 LDS feeds 8 FMAs per thread instead of 2.
// One CTA barrier per batch.
// ---------------------------------------------------------------------------------------------
constexpr int G2_NT = 128;
constexpr int G2_LMAX = 10;  // longest camera list this variant takes
constexpr int G2_ZB = G2_NT * 18;  // doubles per Z buffer: one (landmark, slot) entry per thread
constexpr int G2_ZPAD = 32;  // fragment rows of the last (partial) MMA row tile read past the batch
// Z tile layout.  0: [landmark][slot][row 6][col 3] (a thread's 18 values contiguous; fragment loads have 2-way bank
// conflicts).  1: [k = 3 landmark + col][row], row stride = the smallest value >= 8 T8 that is 4 mod 16 — the
// fragment load of lane (g, t) reads word (4 ks + t) * stride + 8 tile + g: 4 t + g mod 16 is distinct over a
// half-warp, conflict-free — with the producer threads slot-fastest (poses / scaling in padded shared arrays) so that
// their 16-byte stores are conflict-free too; the buffers are dynamic shared memory (2 x G2_ZCAP doubles).
#ifndef CSLAM_G2_KROW
#define CSLAM_G2_KROW 1
#endif
constexpr bool G2_KROW = CSLAM_G2_KROW != 0;

constexpr int G2_ZCAP = 3456;  // doubles per Z buffer, KROW layout (L >= 4: every thread produces)
constexpr int G2_PSTR = 13, G2_SSTR = 7;  // KROW: row strides of the padded pose / scaling copies (odd: conflict-free)
__device__ __forceinline__ int g2_zstride(int T8) { return ((8 * T8 + 11) / 16) * 16 + 4; }

// D(8x8) += A(8x4) B(4x8): lane (g = lane/4, t = lane%4) holds A[g][t], B[t][g], D[g][2t], D[g][2t+1]
__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

// RAG: a "ragged" group — its landmarks see SUBSETS of the group's camera list (variable track lengths, drop-outs):
// observation k of landmark j sits in camera slot fwd[k][j], slot i of landmark j holds observation inv[i][j] (0xff:
// the landmark is not seen from that camera; its rows of Z are zero).  Everything downstream of Z is unchanged.
template <bool WPO, int MINB, bool RAG = false>  // WPO: one 3x3 weight per observation (dataset_vo_sun.cpp:57-59) instead of a shared one
__global__ void __launch_bounds__(G2_NT, MINB)
    schur_grouped2_kernel(DevView v, GroupView gv, int item_lo, int item_hi, LmDiag dg, double* __restrict__ S,
                          double* __restrict__ Bdiag, double* __restrict__ bp, double* __restrict__ gp,
                          double* __restrict__ gl, double* __restrict__ scal) {
    __shared__ __align__(128) double s_pose[G2_LMAX * 12];
    __shared__ double s_sp[G2_LMAX * 6];
    __shared__ double s_A[kItemMax * 9];
    __shared__ __align__(16) double s_Zst[G2_KROW ? 2 : 2 * G2_ZB + G2_ZPAD];
    extern __shared__ __align__(16) double s_Zdyn[];  // KROW: 2 * G2_ZCAP doubles
    double* const s_Z = G2_KROW ? s_Zdyn : s_Zst;
    __shared__ double s_poseP[G2_KROW ? G2_LMAX * G2_PSTR : 1], s_spP[G2_KROW ? G2_LMAX * G2_SSTR : 1];
    __shared__ double s_redsum[32];
    __shared__ int s_free[G2_LMAX];
    __shared__ int s_blk[G2_LMAX * (G2_LMAX + 1) / 2];
    __shared__ double s_W[9];
    __shared__ __align__(8) uint64_t s_bar;

    const int tid = threadIdx.x;
    if (tid == 0) {
        mbar_init(&s_bar, 1);
        fence_mbar_init();
    }
    if (tid < 9 && !WPO) s_W[tid] = v.obs_W[tid];
    __syncthreads();
    uint32_t phase = 0;
    double cost = 0.0;

    for (int w = item_lo + blockIdx.x; w < item_hi; w += gridDim.x) {
        const int g = gv.item_group[w];
        const int L = gv.g_L[g], G = gv.g_G[g];
        const int j0 = gv.item_j0[w], nj = gv.item_n[w];
        const int lm0 = gv.g_lm0[g] + j0;
        const long long obs0 = (long long)gv.g_obs0[g] + j0;
        const int* __restrict__ cams = gv.g_cams + gv.g_off[g];
        const int* __restrict__ blk = gv.g_blk + gv.g_blk_off[g];
        const int P = L * (L + 1) / 2;
        // a group whose cameras span more than the banded preconditioner's window accumulates into the second pair
        // of buffers (engine.cu bandpc_*): S = S_short + S_long, the preconditioner is built from S_short alone
        const bool to_long = gv.g_long != nullptr && gv.g_long[g] != 0;
        double* __restrict__ Sg = to_long ? gv.S_long : S;
        double* __restrict__ Bg = to_long ? gv.Bdiag_long : Bdiag;
        const unsigned char* __restrict__ inv = RAG ? gv.g_map + gv.g_map_off[g] + j0 : nullptr;   // [i * G + jl]
        const unsigned char* __restrict__ fwd = RAG ? inv + (long long)L * G : nullptr;            // [k * G + jl]

        // ---- stage the slice's cameras: free index, scaling, pair->block table, poses (TMA) ----
        if (tid < L) {
            const int f = v.cam_free[cams[tid]];
            s_free[tid] = f;
#pragma unroll
            for (int k = 0; k < 6; ++k) s_sp[6 * tid + k] = f >= 0 ? v.sc_p[6ll * f + k] : 0.0;
        }
        for (int k = tid; k < P; k += G2_NT) s_blk[k] = blk[k];
        if (tid == 0) {
            const int c0 = cams[0];
            mbar_expect_tx(&s_bar, L * 96);
            if (cams[L - 1] - c0 == L - 1) {
                tma_load_1d(s_pose, v.poses + 12ll * c0, L * 96, &s_bar);
            } else {
                for (int i = 0; i < L; ++i) tma_load_1d(s_pose + 12 * i, v.poses + 12ll * cams[i], 96, &s_bar);
            }
        }
        mbar_wait(&s_bar, phase);
        phase ^= 1;
        __syncthreads();
        if (G2_KROW) {
            // (visible to the producers after pass 1's barrier)
            for (int k = tid; k < 12 * L; k += G2_NT) s_poseP[(k / 12) * G2_PSTR + k % 12] = s_pose[k];
            for (int k = tid; k < 6 * L; k += G2_NT) s_spP[(k / 6) * G2_SSTR + k % 6] = s_sp[k];
        }

        // ---- roles of the pipeline below; the first observation is requested before pass 1 ----
        // landmarks per batch (one observation per thread), a multiple of 4 so that a batch's
        // 3 TL columns are whole k-steps of 4
        const int T8 = (6 * L + 7) >> 3;
        const int ZS = g2_zstride(T8);
        const int TL = G2_KROW ? min((G2_NT / L) & ~3, (G2_ZCAP / (3 * ZS)) & ~3) : (G2_NT / L) & ~3;
        const int ZBUF = G2_KROW ? G2_ZCAP : G2_ZB;
        const int nbatch = (nj + TL - 1) / TL;
        const int pq = G2_KROW ? tid / L : tid % TL, pi = G2_KROW ? (tid < TL * L ? tid % L : L) : tid / TL;
        const bool p_active = pi < L;
        const int pf = p_active ? s_free[pi] : -1;
        const bool producing = p_active && pf >= 0;
        // this thread's observation of the NEXT batch (u, v, d, point, point scaling), loaded one
        // pipeline step ahead so the global-load latency hides behind the MMA stage
        double nx[9];
        int nx_k = 0;  // RAG: which observation of the landmark this slot holds (0xff: none)
        auto prefetch = [&](int jl) {
            if (producing && jl < nj) {
                const long long j = lm0 + jl;
                if (RAG) nx_k = inv[(long long)pi * G + jl];
                const long long e = obs0 + (long long)(RAG ? (nx_k == 0xff ? 0 : nx_k) : pi) * G + jl;
                nx[0] = v.obs_u[e]; nx[1] = v.obs_v[e]; nx[2] = v.obs_d[e];
                nx[3] = v.points[3 * j]; nx[4] = v.points[3 * j + 1]; nx[5] = v.points[3 * j + 2];
                nx[6] = v.sc_l[3 * j]; nx[7] = v.sc_l[3 * j + 1]; nx[8] = v.sc_l[3 * j + 2];
            }
        };
#pragma unroll
        for (int k = 0; k < 9; ++k) nx[k] = 0.0;
        prefetch(pq);

        // ---- pass 1: V_j = sum Jp^T Jp + D^2, Cholesky, A = C^-1, t = A g_l ----
        for (int jl = tid; jl < nj; jl += G2_NT) {
            const long long j = lm0 + jl;
            const double p[3] = {v.points[3 * j], v.points[3 * j + 1], v.points[3 * j + 2]};
            const double sl[3] = {v.sc_l[3 * j], v.sc_l[3 * j + 1], v.sc_l[3 * j + 2]};
            double V[6] = {0, 0, 0, 0, 0, 0}, gq[3] = {0, 0, 0};
            // the L observations of a landmark are a chain of dependent global loads: keep three in
            // flight.  The queue is three named slots used in rotation by a loop unrolled three times —
            // rotating it with register moves would make each move wait for the load it forwards.
            double qu[3], qv[3], qd[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const long long e = obs0 + (long long)(RAG ? 0 : min(k, L - 1)) * G + jl;  // (RAG: the queue is not used)
                qu[k] = v.obs_u[e];
                qv[k] = v.obs_v[e];
                qd[k] = v.obs_d[e];
            }
            double Wl[9];
            if (!WPO) {
#pragma unroll
                for (int k = 0; k < 9; ++k) Wl[k] = s_W[k];
            }
            if (RAG) {
                // the landmark's own observations, each in the camera slot the map names
                const int cnt = int(v.lm_cnt[j]);
                for (int k = 0; k < cnt; ++k) {
                    const long long e = obs0 + (long long)k * G + jl;
                    const int ci = fwd[(long long)k * G + jl];
                    if (WPO) {
#pragma unroll
                        for (int q = 0; q < 9; ++q) Wl[q] = v.obs_W[9 * e + q];
                    }
                    double r[3], Jp[9];
                    stereo_block_point(v.cam, s_pose + 12 * ci, p, v.obs_u[e], v.obs_v[e], v.obs_d[e], Wl, r, Jp);
                    cost += 0.5 * (r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
#pragma unroll
                    for (int q = 0; q < 3; ++q) {
                        const double a = Jp[3 * q] * sl[0], b = Jp[3 * q + 1] * sl[1], c = Jp[3 * q + 2] * sl[2];
                        V[0] += a * a; V[1] += a * b; V[2] += a * c; V[3] += b * b; V[4] += b * c; V[5] += c * c;
                        gq[0] += a * r[q]; gq[1] += b * r[q]; gq[2] += c * r[q];
                    }
                }
            }
            auto step = [&](double& su, double& sv, double& sd, int i) {
                const double ou = su, ov = sv, od = sd;
                {
                    // unconditional (index clamped): a predicated load goes to a temporary and is forwarded
                    // into the slot by a move that waits for it
                    const long long e3 = obs0 + (long long)min(i + 3, L - 1) * G + jl;
                    su = v.obs_u[e3];
                    sv = v.obs_v[e3];
                    sd = v.obs_d[e3];
                }
                double r[3], Jp[9];
                if (WPO) {
                    // (a predicated-off load in this loop would share a scoreboard with the prefetch
                    // above and make every iteration wait for it: hence the template parameter)
                    const double* Wg = v.obs_W + 9 * (obs0 + (long long)i * G + jl);
#pragma unroll
                    for (int k = 0; k < 9; ++k) Wl[k] = Wg[k];
                }
                stereo_block_point(v.cam, s_pose + 12 * i, p, ou, ov, od, Wl, r, Jp);
                cost += 0.5 * (r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const double a = Jp[3 * k] * sl[0], b = Jp[3 * k + 1] * sl[1], c = Jp[3 * k + 2] * sl[2];
                    V[0] += a * a; V[1] += a * b; V[2] += a * c; V[3] += b * b; V[4] += b * c; V[5] += c * c;
                    gq[0] += a * r[k]; gq[1] += b * r[k]; gq[2] += c * r[k];
                }
            };
            for (int i = 0; i < L && !RAG; i += 3) {
                step(qu[0], qv[0], qd[0], i);
                if (i + 1 < L) step(qu[1], qv[1], qd[1], i + 1);
                if (i + 2 < L) step(qu[2], qv[2], qd[2], i + 2);
            }
            V[0] += fmin(fmax(V[0], dg.min_diag), dg.max_diag) * dg.inv_radius;
            V[3] += fmin(fmax(V[3], dg.min_diag), dg.max_diag) * dg.inv_radius;
            V[5] += fmin(fmax(V[5], dg.min_diag), dg.max_diag) * dg.inv_radius;
            double A[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
            bool pd = V[0] > 0.0;
            if (pd) {
                const double c00 = sqrt(V[0]);
                const double c10 = V[1] / c00, c20 = V[2] / c00;
                const double d1 = V[3] - c10 * c10;
                pd = d1 > 0.0;
                if (pd) {
                    const double c11 = sqrt(d1);
                    const double c21 = (V[4] - c20 * c10) / c11;
                    const double d2 = V[5] - c20 * c20 - c21 * c21;
                    pd = d2 > 0.0 && d2 < 1.7976931348623157e308;
                    if (pd) {
                        const double c22 = sqrt(d2);
                        const double a00 = 1.0 / c00, a11 = 1.0 / c11, a22 = 1.0 / c22;
                        const double a10 = -c10 * a00 * a11;
                        const double a21 = -c21 * a11 * a22;
                        const double a20 = -(c20 * a00 + c21 * a10) * a22;
                        A[0] = a00; A[1] = a10; A[2] = a11; A[3] = a20; A[4] = a21; A[5] = a22;
                        A[6] = a00 * gq[0];
                        A[7] = a10 * gq[0] + a11 * gq[1];
                        A[8] = a20 * gq[0] + a21 * gq[1] + a22 * gq[2];
                    }
                }
            }
            if (!pd) red_add(&scal[SC_INVALID], 1.0);
#pragma unroll
            for (int k = 0; k < 9; ++k) s_A[9 * jl + k] = A[k];
            gl[3 * j] = gq[0];
            gl[3 * j + 1] = gq[1];
            gl[3 * j + 2] = gq[2];
        }
        __syncthreads();

        // ---- pipeline: produce Z of batch n+1, apply batch n with DMMA ----
        double U[21], gpa[6], bpa[6];
#pragma unroll
        for (int k = 0; k < 21; ++k) U[k] = 0.0;
#pragma unroll
        for (int k = 0; k < 6; ++k) gpa[k] = bpa[k] = 0.0;
        // consumer role (per warp): MMA row tiles [r0, r0 + nr) x column tiles [c0, c0 + nc); the
        // triangular warps (0 and 3) own the tiles i <= j of a diagonal square, the rectangular
        // warps (1 and 2) split the off-diagonal rectangle by rows.  nr <= 4 (2 for rectangles), nc <= 4.
        const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);  // tells the compiler the role is warp-uniform
        const int lane = tid & 31, fg = lane >> 2, ft = lane & 3;
        const int Th = (T8 + 1) >> 1, Th1 = (Th + 1) >> 1;
        const bool tri = warp == 0 || warp == 3;
        int r0, nr, c0, nc;
        if (warp == 0) {
            r0 = 0; nr = Th; c0 = 0; nc = Th;
        } else if (warp == 3) {
            r0 = Th; nr = T8 - Th; c0 = Th; nc = T8 - Th;
        } else if (warp == 1) {
            r0 = 0; nr = Th1; c0 = Th; nc = T8 - Th;
        } else {
            r0 = Th1; nr = Th - Th1; c0 = Th; nc = T8 - Th;
        }
        const int zrow_r = 3 * (8 * r0 + fg), zrow_c = 3 * (8 * c0 + fg);  // + 24 per further tile
        const int zlm = 18 * L;                                           // doubles per landmark
        double M[20];  // triangular: tile (i, j >= i) at 4i - i(i-1)/2 + j - i; rectangular: 4i + j
#pragma unroll
        for (int k = 0; k < 20; ++k) M[k] = 0.0;

        for (int bt = 0; bt <= nbatch; ++bt) {
            if (bt < nbatch && producing) {
                const int jl = bt * TL + pq;
                // this thread's 6 rows x 3 columns of the tile
                double* zt = G2_KROW ? s_Z + (bt & 1) * ZBUF + 3 * pq * ZS + 6 * pi : s_Z + (bt & 1) * ZBUF + (pq * L + pi) * 18;
                auto zero_rows = [&]() {
                    if (G2_KROW) {
#pragma unroll
                        for (int c = 0; c < 3; ++c)
#pragma unroll
                            for (int a = 0; a < 6; a += 2) *reinterpret_cast<double2*>(zt + c * ZS + a) = make_double2(0.0, 0.0);
                    } else {
#pragma unroll
                        for (int k = 0; k < 18; k += 2) *reinterpret_cast<double2*>(zt + k) = make_double2(0.0, 0.0);
                    }
                };
                if (RAG && jl < nj && nx_k == 0xff) {
                    // the landmark is not seen from this camera: zero rows
                    zero_rows();
                    prefetch(jl + TL);
                } else if (jl < nj) {
                    const double p[3] = {nx[3], nx[4], nx[5]};
                    const double sl[3] = {nx[6], nx[7], nx[8]};
                    const double ou = nx[0], ov = nx[1], od = nx[2];
                    const long long e = obs0 + (long long)(RAG ? nx_k : pi) * G + jl;
                    double r[3], Jc[18], Jp[9];
                    {
                        double pose[12], Wl[9];
#pragma unroll
                        for (int k = 0; k < 12; ++k) pose[k] = G2_KROW ? s_poseP[G2_PSTR * pi + k] : s_pose[12 * pi + k];
#pragma unroll
                        for (int k = 0; k < 9; ++k) Wl[k] = WPO ? v.obs_W[9 * e + k] : s_W[k];
                        stereo_block<true>(v.cam, pose, p, ou, ov, od, Wl, r, Jc, Jp);
                    }
                    prefetch(jl + TL);
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
#pragma unroll
                        for (int q = 0; q < 3; ++q) Jp[3 * k + q] *= sl[q];
#pragma unroll
                        for (int q = 0; q < 6; ++q) Jc[6 * k + q] *= G2_KROW ? s_spP[G2_SSTR * pi + q] : s_sp[6 * pi + q];
                    }
                    const double* A = s_A + 9 * jl;
                    const double a00 = A[0], a10 = A[1], a11 = A[2], a20 = A[3], a21 = A[4], a22 = A[5];
                    const double t0 = A[6], t1 = A[7], t2 = A[8];
                    int u = 0;
                    double Zr[G2_KROW ? 18 : 1];  // KROW: [col][row], stored as 16-byte pairs below
#pragma unroll
                    for (int a = 0; a < 6; ++a) {
                        const double w0 = Jc[a] * Jp[0] + Jc[6 + a] * Jp[3] + Jc[12 + a] * Jp[6];
                        const double w1 = Jc[a] * Jp[1] + Jc[6 + a] * Jp[4] + Jc[12 + a] * Jp[7];
                        const double w2 = Jc[a] * Jp[2] + Jc[6 + a] * Jp[5] + Jc[12 + a] * Jp[8];
                        const double z0 = w0 * a00;
                        const double z1 = w0 * a10 + w1 * a11;
                        const double z2 = w0 * a20 + w1 * a21 + w2 * a22;
                        if (G2_KROW) {
                            Zr[a] = z0, Zr[6 + a] = z1, Zr[12 + a] = z2;
                        } else {
                            zt[3 * a] = z0;
                            zt[3 * a + 1] = z1;
                            zt[3 * a + 2] = z2;
                        }
                        const double ga = Jc[a] * r[0] + Jc[6 + a] * r[1] + Jc[12 + a] * r[2];
                        gpa[a] += ga;
                        bpa[a] = fma(-z0, t0, fma(-z1, t1, fma(-z2, t2, bpa[a] + ga)));
#pragma unroll
                        for (int b = a; b < 6; ++b, ++u)
                            U[u] = fma(Jc[a], Jc[b], fma(Jc[6 + a], Jc[6 + b], fma(Jc[12 + a], Jc[12 + b], U[u])));
                    }
                    if (G2_KROW) {
#pragma unroll
                        for (int c = 0; c < 3; ++c)
#pragma unroll
                            for (int a = 0; a < 6; a += 2)
                                *reinterpret_cast<double2*>(zt + c * ZS + a) = make_double2(Zr[6 * c + a], Zr[6 * c + a + 1]);
                    }
                } else {
                    // a k-step of the last batch may reach one landmark past nj: zero columns
                    zero_rows();
                }
            }
            if (bt > 0) {
                const double* zt = s_Z + ((bt - 1) & 1) * ZBUF;
                const int nval = min(TL, nj - (bt - 1) * TL);
                const int nk = (3 * nval + 3) >> 2;
                // fragment word of (k, row): KROW k * ZS + row, else landmark * zlm + col + 3 row; rstep = one tile down
                const int rstep = G2_KROW ? 8 : 24;
                const int frow_r = G2_KROW ? 8 * r0 + fg : zrow_r, frow_c = G2_KROW ? 8 * c0 + fg : zrow_c;
                // The fragments of k-step ks + 1 are requested right after the MMAs of k-step ks have been issued (they read
                // their operands at issue), so the loads fly while the tensor pipe drains: 2.39 -> 2.33 ms on C5.  Loading
                // them BEFORE the MMAs into a second register set and moving them over measured no better than the plain
                // loop (2.39 ms: +4 registers, and the moves wait for the loads).
                if (tri) {
                    auto ld = [&](int ks, double* f) {
                        const int kk = 4 * ks + ft, jj = kk / 3;
                        const double* zb = (G2_KROW ? zt + kk * ZS : zt + jj * zlm + (kk - 3 * jj)) + frow_r;
#pragma unroll
                        for (int i = 0; i < 4; ++i) f[i] = i < nr ? zb[rstep * i] : 0.0;
                    };
                    double f[4];
                    ld(0, f);
                    for (int ks = 0; ks < nk; ++ks) {
#pragma unroll
                        for (int i = 0; i < 4; ++i)
#pragma unroll
                            for (int j = i; j < 4; ++j)
                                if (j < nr) dmma884(M[2 * (4 * i - i * (i - 1) / 2 + j - i)], M[2 * (4 * i - i * (i - 1) / 2 + j - i) + 1], f[i], f[j]);
                        if (ks + 1 < nk) ld(ks + 1, f);
                    }
                } else {
                    auto ld = [&](int ks, double* fr, double* fc) {
                        const int kk = 4 * ks + ft, jj = kk / 3;
                        const double* zb = G2_KROW ? zt + kk * ZS : zt + jj * zlm + (kk - 3 * jj);
#pragma unroll
                        for (int i = 0; i < 2; ++i) fr[i] = i < nr ? zb[frow_r + rstep * i] : 0.0;
#pragma unroll
                        for (int j = 0; j < 4; ++j) fc[j] = j < nc ? zb[frow_c + rstep * j] : 0.0;
                    };
                    double fr[2], fc[4];
                    ld(0, fr, fc);
                    for (int ks = 0; ks < nk; ++ks) {
#pragma unroll
                        for (int i = 0; i < 2; ++i)
#pragma unroll
                            for (int j = 0; j < 4; ++j)
                                if (i < nr && j < nc) dmma884(M[2 * (4 * i + j)], M[2 * (4 * i + j) + 1], fr[i], fc[j]);
                        if (ks + 1 < nk) ld(ks + 1, fr, fc);
                    }
                }
            }
            __syncthreads();
        }

        // ---- flush: MMA tile entries back to the 6x6 pair blocks, camera diagonal / gradients ----
        {
            // entry (row, col) of the 6L x 6L tile, row <= col: block (row/6, col/6), diagonal blocks
            // upper triangle only (finalize mirrors it); one RED per entry per slice as before
            auto flush_tile = [&](int mt, int nt, double m0, double m1) {
                const int row = 8 * mt + fg;
                if (row >= 6 * L) return;
                const int a = row / 6, ra = row - 6 * a;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int col = 8 * nt + 2 * ft + h;
                    if (col >= 6 * L || col < row) continue;
                    const int b = col / 6, rb = col - 6 * b;
                    const int e = s_blk[a * L - a * (a - 1) / 2 + (b - a)];
                    if (e >= 0) red_add(&Sg[36ll * e + 6 * ra + rb], h ? -m1 : -m0);
                }
            };
            if (tri) {
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = i; j < 4; ++j)
                        if (j < nr) flush_tile(r0 + i, r0 + j, M[2 * (4 * i - i * (i - 1) / 2 + j - i)], M[2 * (4 * i - i * (i - 1) / 2 + j - i) + 1]);
            } else {
#pragma unroll
                for (int i = 0; i < 2; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (i < nr && j < nc) flush_tile(r0 + i, c0 + j, M[2 * (4 * i + j)], M[2 * (4 * i + j) + 1]);
            }
        }
        {
            // reduce U, g, rhs over the TL threads that share a slot, then one RED per value
            double* red = s_Z;  // every consumer read of s_Z is behind the last barrier
            if (producing) {
#pragma unroll
                for (int k = 0; k < 21; ++k) red[tid * 33 + k] = U[k];
#pragma unroll
                for (int k = 0; k < 6; ++k) {
                    red[tid * 33 + 21 + k] = gpa[k];
                    red[tid * 33 + 27 + k] = bpa[k];
                }
            }
            __syncthreads();
            // 33 values per slot: thread (slot, value)
            for (int idx = tid; idx < L * 33; idx += G2_NT) {
                const int i = idx / 33, k = idx - 33 * i;
                const int f = s_free[i];
                if (f < 0) continue;
                double acc = 0.0;
                const int nl = min(TL, nj);  // lanes beyond nj never produced
                for (int q = 0; q < nl; ++q) acc += red[(G2_KROW ? q * L + i : i * TL + q) * 33 + k];
                if (k < 21) {
                    int a = 0, rem = k;
                    while (rem >= 6 - a) {
                        rem -= 6 - a;
                        ++a;
                    }
                    red_add(&Bg[36ll * f + 6 * a + a + rem], acc);
                } else if (k < 27) {
                    red_add(&gp[6ll * f + (k - 21)], acc);
                } else {
                    red_add(&bp[6ll * f + (k - 27)], acc);
                }
            }
        }
        __syncthreads();
    }
    block_atomic_sum(cost, &scal[SC_COST], s_redsum);
}


}  // namespace

namespace {
constexpr int G2_DYN = G2_KROW ? 2 * G2_ZCAP * int(sizeof(double)) : 0;  // dynamic shared memory of schur_grouped2_kernel
template <class K>
void g2_launch(K kernel, int grid, cudaStream_t s, const DevView& v, const GroupView& g, int lo, int hi, LmDiag dg, double* S,
               double* Bdiag, double* bp, double* gp, double* gl, double* scal) {
    // (the opt-in is per device and costs a microsecond: set at every launch rather than cached process-wide)
    if (G2_DYN > 48 * 1024) CSLAM_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, G2_DYN));
    kernel<<<grid, G2_NT, G2_DYN, s>>>(v, g, lo, hi, dg, S, Bdiag, bp, gp, gl, scal);
}
}  // namespace

void launch_schur_grouped(cudaStream_t s, const DevView& v, const GroupView& g, int n_items_small, int n_items_rag, LmDiag dg,
                          double* S, double* Bdiag, double* bp, double* gp, double* gl, double* scal) {
    // items: [0, n_items_small) exact groups with L <= 10, then exact groups with 10 < L <= 16, then the last
    // n_items_rag items: ragged groups (camera window <= 10)
    if (n_items_rag > 0) {
        const int lo = g.n_items - n_items_rag;
        const int grid = n_items_rag < 2 * kSMs ? n_items_rag : 2 * kSMs;
        if (v.W_per_obs)
            g2_launch(schur_grouped2_kernel<true, 2, true>, grid, s, v, g, lo, g.n_items, dg, S, Bdiag, bp, gp, gl, scal);
        else
            g2_launch(schur_grouped2_kernel<false, 2, true>, grid, s, v, g, lo, g.n_items, dg, S, Bdiag, bp, gp, gl, scal);
        g_kernel_launches.fetch_add(1, std::memory_order_relaxed);
    }
    // items [0, n_items_small) have L <= 10 (two consumer warps), the rest 10 < L <= 16 (five)
    if (n_items_small > 0) {
        // Two resident CTAs per SM.  Measured alternatives on C5 (this kernel: 2.47 ms): three CTAs per SM
        // (CSLAM_G2_OCC=3: 168 registers, spills) 2.78 ms; a warp-specialised 384-thread CTA (4 DMMA warps +
        // 8 producer warps, Z ring with full/empty mbarriers; commit "K2 experiment") 2.63 ms — it hides the MMA
        // stage completely but leaves the producers' staging / pass 1 / reduction bubbles of a slice exposed,
        // which a second resident CTA covers for free; pass 1 as a full-occupancy kernel of its own (factors
        // through global memory) 2.60 ms in total — inside this kernel it hides behind the other CTA's MMA stage.
        static const int occ = [] {
            const char* e = std::getenv("CSLAM_G2_OCC");  // A/B knob: resident CTAs per SM
            return e && std::atoi(e) == 3 ? 3 : 2;
        }();
        const int grid = n_items_small < occ * kSMs ? n_items_small : occ * kSMs;
        // (a single-evaluation variant — V_j reduced through shared memory inside the pipeline, no pass 1 —
        // was measured slower, 2.67 vs 2.50 ms on C5: its per-batch Cholesky chain is exposed latency)
        if (v.W_per_obs)
            g2_launch(schur_grouped2_kernel<true, 2>, grid, s, v, g, 0, n_items_small, dg, S, Bdiag, bp, gp, gl, scal);
        else if (occ == 3)
            g2_launch(schur_grouped2_kernel<false, 3>, grid, s, v, g, 0, n_items_small, dg, S, Bdiag, bp, gp, gl, scal);
        else
            g2_launch(schur_grouped2_kernel<false, 2>, grid, s, v, g, 0, n_items_small, dg, S, Bdiag, bp, gp, gl, scal);
        g_kernel_launches.fetch_add(1, std::memory_order_relaxed);
    }
    if (g.n_items - n_items_rag > n_items_small) {
        const int n = g.n_items - n_items_rag - n_items_small;
        const int grid = n < 2 * kSMs ? n : 2 * kSMs;
        // (more than 10 cameras: always beyond the preconditioner's window)
        schur_grouped_kernel<5><<<grid, 192, 0, s>>>(v, g, n_items_small, g.n_items - n_items_rag, dg, g.S_long ? g.S_long : S,
                                                      g.S_long ? g.Bdiag_long : Bdiag, bp, gp, gl, scal);
        g_kernel_launches.fetch_add(1, std::memory_order_relaxed);
    }
    CSLAM_CUDA(cudaGetLastError());
}

}  // namespace cslam
