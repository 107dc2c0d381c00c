// Shared by kernels_phong_solve.cu (K2p / K4p, one warp per vertex) and kernels_phong_grouped.cu (K2p for grouped
// vertices, fused DMMA tile update): the per-observation lighting blocks in tangent coordinates and the small
// dense helpers around the 6x6 vertex block.  Everything is `static`-like (anonymous namespace) per translation unit.
#pragma once
#include "kernels.cuh"

namespace cslam {
namespace {

constexpr int PB_WARPS = 4;

__device__ __forceinline__ int find_block(const int* __restrict__ rowptr, const int* __restrict__ col, int a, int b) {
    int lo = rowptr[a], hi = rowptr[a + 1] - 1;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (col[mid] < b)
            lo = mid + 1;
        else
            hi = mid;
    }
    return lo;  // the pattern is built from co-visibility, so the block exists
}

// sum over the LW lanes of the caller's segment of the warp (LW = 32: the whole warp)
template <int LW>
__device__ __forceinline__ double seg_sum(double v) {
#pragma unroll
    for (int o = LW / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Cholesky-based inverse of a symmetric positive definite 6x6 (full storage in, full out).
__device__ __forceinline__ bool spd6_inverse(const double* A, double* Ainv) {
    double L[36];
#pragma unroll
    for (int i = 0; i < 36; ++i) L[i] = A[i];
    bool ok = true;
#pragma unroll
    for (int j = 0; j < 6; ++j) {
        double d = L[6 * j + j];
#pragma unroll
        for (int k = 0; k < j; ++k) d -= L[6 * j + k] * L[6 * j + k];
        if (!(d > 0.0) || !(d < 1.7976931348623157e308)) ok = false;
        d = sqrt(d);
        L[6 * j + j] = d;
        const double id = 1.0 / d;
#pragma unroll
        for (int i = j + 1; i < 6; ++i) {
            double s = L[6 * i + j];
#pragma unroll
            for (int k = 0; k < j; ++k) s -= L[6 * i + k] * L[6 * j + k];
            L[6 * i + j] = s * id;
        }
    }
    // Li = L^-1 (lower), Ainv = Li^T Li
    double Li[36];
#pragma unroll
    for (int c = 0; c < 6; ++c) {
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            if (i < c) {
                Li[6 * i + c] = 0.0;
            } else if (i == c) {
                Li[6 * i + c] = 1.0 / L[6 * i + i];
            } else {
                double s = 0.0;
#pragma unroll
                for (int k = c; k < i; ++k) s -= L[6 * i + k] * Li[6 * k + c];
                Li[6 * i + c] = s / L[6 * i + i];
            }
        }
    }
#pragma unroll
    for (int a = 0; a < 6; ++a)
#pragma unroll
        for (int b = a; b < 6; ++b) {
            double s = 0.0;
#pragma unroll
            for (int k = b; k < 6; ++k) s += Li[6 * k + a] * Li[6 * k + b];
            Ainv[6 * a + b] = Ainv[6 * b + a] = s;
        }
    return ok;
}

// The three blocks of one observation in tangent coordinates, columns scaled.
struct PhObs {
    double rs[3], rI, rN[3];
    double Jcs[18], JIc[6], JNc[18];  // pose columns: stereo 3x6, intensity 1x6, normal 3x6
    double S[9], ip[3];               // position columns: stereo 3x3, intensity 1x3
    double in[3], N[9];               // normal columns: intensity 1x3, normal 3x3
    double ag[7];                     // material 3 | texture 1 | light 3 (intensity row)
    int f;
};

struct VertexCtx {
    double p[3], n[3], phong[3], kd, light[3];
    double sl[3], sn[3], sg[7];
    int gi[7];
};

__device__ __forceinline__ void load_vertex(const DevView& v, const PhongSolveView& q, int j, VertexCtx& c) {
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        c.p[k] = v.points[3ll * j + k];
        c.n[k] = q.normals[3ll * j + k];
        c.sl[k] = v.sc_l[3ll * j + k];
        c.sn[k] = q.sc_n[3ll * j + k];
    }
    const int m = q.v_mat[j], t = q.v_tex[j];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        c.gi[k] = 3 * m + k;
        c.gi[4 + k] = 3 * q.n_mat + q.n_tex + k;
    }
    c.gi[3] = 3 * q.n_mat + t;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        c.phong[k] = q.gx[c.gi[k]];
        c.light[k] = q.gx[c.gi[4 + k]];
    }
    c.kd = q.gx[c.gi[3]];
#pragma unroll
    for (int k = 0; k < 7; ++k) c.sg[k] = q.sc_g[c.gi[k]];
}

__device__ __forceinline__ void eval_phong_obs(const DevView& v, const PhongSolveView& q, long long e,
                                               const VertexCtx& c, PhObs& o) {
    const uint32_t cam = v.obs_cam[e];
    const double* pose = v.poses + 12ll * cam;
    o.f = v.cam_free[cam];
    stereo_block<true>(v.cam, pose, c.p, v.obs_u[e], v.obs_v[e], v.obs_d[e], v.obs_W, o.rs, o.Jcs, o.S);
    double Jk[3], Jt[1], Jl[3];
    intensity_block(pose, c.p, c.n, c.phong, c.kd, c.light, q.obs_I[e], q.int_stiffness, q.directional != 0, &o.rI,
                    o.JIc, o.ip, o.in, Jk, Jt, Jl);
    const double nobs[3] = {q.obs_n[e], q.obs_n[v.n_obs + e], q.obs_n[2 * v.n_obs + e]};
    normal_block(pose, c.n, nobs, q.Wn, o.rN, o.JNc, o.N);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
#pragma unroll
        for (int b = 0; b < 3; ++b) {
            o.S[3 * k + b] *= c.sl[b];
            o.N[3 * k + b] *= c.sn[b];
        }
        o.ip[k] *= c.sl[k];
        o.in[k] *= c.sn[k];
        o.ag[k] = Jk[k] * c.sg[k];
        o.ag[4 + k] = Jl[k] * c.sg[4 + k];
    }
    o.ag[3] = Jt[0] * c.sg[3];
    if (q.hold_positions) {
#pragma unroll
        for (int k = 0; k < 9; ++k) o.S[k] = 0.0;
#pragma unroll
        for (int k = 0; k < 3; ++k) o.ip[k] = 0.0;
        if (o.f < 0) o.rs[0] = o.rs[1] = o.rs[2] = 0.0;  // dropped block: no variable parameter left
    }
    if (o.f >= 0) {
        const double* sp = v.sc_p + 6ll * o.f;
#pragma unroll
        for (int a = 0; a < 6; ++a) {
            const double s = sp[a];
            o.Jcs[a] *= s;
            o.Jcs[6 + a] *= s;
            o.Jcs[12 + a] *= s;
            o.JIc[a] *= s;
            o.JNc[a] *= s;
            o.JNc[6 + a] *= s;
            o.JNc[12 + a] *= s;
        }
    }
}

// residuals only, at arbitrary state arrays (candidate evaluation)
__device__ __forceinline__ double phong_obs_cost(const DevView& v, const PhongSolveView& q, long long e,
                                                 const double* poses, const double* p, const double* n,
                                                 const double* phong, double kd, const double* light) {
    const uint32_t cam = v.obs_cam[e];
    const double* pose = poses + 12ll * cam;
    double rs[3], rI, rN[3];
    stereo_block<false>(v.cam, pose, p, v.obs_u[e], v.obs_v[e], v.obs_d[e], v.obs_W, rs, nullptr, nullptr);
    if (q.hold_positions && v.cam_free[cam] < 0) rs[0] = rs[1] = rs[2] = 0.0;
    intensity_block(pose, p, n, phong, kd, light, q.obs_I[e], q.int_stiffness, q.directional != 0, &rI, nullptr,
                    nullptr, nullptr, nullptr, nullptr, nullptr);
    const double nobs[3] = {q.obs_n[e], q.obs_n[v.n_obs + e], q.obs_n[2 * v.n_obs + e]};
    normal_block(pose, n, nobs, q.Wn, rN, nullptr, nullptr);
    return 0.5 * (rs[0] * rs[0] + rs[1] * rs[1] + rs[2] * rs[2] + rI * rI + rN[0] * rN[0] + rN[1] * rN[1] + rN[2] * rN[2]);
}

// A_v^T A_v (upper 21, row-major order of the upper triangle) and A_v^T w for a 7-vector w
__device__ __forceinline__ void vertex_normal_eq(const PhObs& o, const double* w, double* V21, double* g6) {
    int idx = 0;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
#pragma unroll
        for (int b = a; b < 3; ++b)
            V21[idx++] = o.S[a] * o.S[b] + o.S[3 + a] * o.S[3 + b] + o.S[6 + a] * o.S[6 + b] + o.ip[a] * o.ip[b];
#pragma unroll
        for (int b = 0; b < 3; ++b) V21[idx++] = o.ip[a] * o.in[b];
    }
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = a; b < 3; ++b)
            V21[idx++] = o.in[a] * o.in[b] + o.N[a] * o.N[b] + o.N[3 + a] * o.N[3 + b] + o.N[6 + a] * o.N[6 + b];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        g6[a] = o.S[a] * w[0] + o.S[3 + a] * w[1] + o.S[6 + a] * w[2] + o.ip[a] * w[3];
        g6[3 + a] = o.in[a] * w[3] + o.N[a] * w[4] + o.N[3 + a] * w[5] + o.N[6 + a] * w[6];
    }
}

__device__ __forceinline__ void unpack_sym6(const double* V21, double* V) {
    int idx = 0;
#pragma unroll
    for (int a = 0; a < 6; ++a)
#pragma unroll
        for (int b = a; b < 6; ++b) {
            V[6 * a + b] = V[6 * b + a] = V21[idx++];
        }
}

__device__ __forceinline__ void add_lm_diag(double* V, const LmDiag& dg) {
#pragma unroll
    for (int a = 0; a < 6; ++a) V[7 * a] += fmin(fmax(V[7 * a], dg.min_diag), dg.max_diag) * dg.inv_radius;
}

}  // namespace
}  // namespace cslam
