// Closed-form FP64 residuals and SE(3)-tangent Jacobians for ceres-slam's residual blocks.
//
// What Ceres obtains by running the reference's templated functors on Jets and multiplying by
// the autodiff Jacobian of SE3Perturbation (perturbations.hpp:45-65, T' = exp(eps) T with the
// decoupled exp of se3group.hpp:313-325) is written here in closed form:
//     d(Rp+t)/d eps = [ I | -[p_c]x ],   d(Rv)/d eps = [ 0 | -[v_c]x ].
// The same formula shape as the reference is kept where rounding matters (one reciprocal of z,
// then products — stereo_camera.hpp:79-84) and every branch condition is the reference's.
// Functions are __host__ __device__ so the CPU test-suite can check them against the Jet oracle
// without a GPU; the product only ever calls them from kernels.
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define CSLAM_HD __host__ __device__ __forceinline__
#else
#define CSLAM_HD inline
#endif

namespace cslam {

struct CameraIntrinsics {
    double fu, fv, cu, cv, b;
};

// p_c = R p + t for a pose stored [t | R row-major] (se3group.hpp:191-193)
CSLAM_HD void transform_point(const double* pose, const double* p, double* pc) {
    pc[0] = pose[3] * p[0] + pose[4] * p[1] + pose[5] * p[2] + pose[0];
    pc[1] = pose[6] * p[0] + pose[7] * p[1] + pose[8] * p[2] + pose[1];
    pc[2] = pose[9] * p[0] + pose[10] * p[1] + pose[11] * p[2] + pose[2];
}

// Stereo reprojection block (stereo_reprojection_error.hpp:27-55).
//   r  = W (pi(p_c) - z)                       3
//   Jc = W Pi [ I | -[p_c]x ]                  3x6 row-major (pose tangent)
//   Jp = W Pi R                                3x3 row-major (point)
// Pi as in stereo_camera.hpp:88-104.
template <bool kJac>
CSLAM_HD void stereo_block(const CameraIntrinsics& c, const double* pose, const double* p,
                           double u, double v, double d, const double* W, double* r, double* Jc,
                           double* Jp) {
    double pc[3];
    transform_point(pose, p, pc);
    const double iz = 1.0 / pc[2];
    const double e0 = c.fu * pc[0] * iz + c.cu - u;
    const double e1 = c.fv * pc[1] * iz + c.cv - v;
    const double e2 = c.fu * c.b * iz - d;
    r[0] = W[0] * e0 + W[1] * e1 + W[2] * e2;
    r[1] = W[3] * e0 + W[4] * e1 + W[5] * e2;
    r[2] = W[6] * e0 + W[7] * e1 + W[8] * e2;
    if (!kJac) return;
    const double iz2 = iz * iz;
    const double p00 = c.fu * iz, p02 = -c.fu * pc[0] * iz2;
    const double p11 = c.fv * iz, p12 = -c.fv * pc[1] * iz2;
    const double p22 = -c.fu * c.b * iz2;
    const double* R = pose + 3;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        // A = W Pi, row i
        const double a0 = W[3 * i] * p00;
        const double a1 = W[3 * i + 1] * p11;
        const double a2 = W[3 * i] * p02 + W[3 * i + 1] * p12 + W[3 * i + 2] * p22;
        Jc[6 * i + 0] = a0;
        Jc[6 * i + 1] = a1;
        Jc[6 * i + 2] = a2;
        Jc[6 * i + 3] = a2 * pc[1] - a1 * pc[2];
        Jc[6 * i + 4] = a0 * pc[2] - a2 * pc[0];
        Jc[6 * i + 5] = a1 * pc[0] - a0 * pc[1];
        Jp[3 * i + 0] = a0 * R[0] + a1 * R[3] + a2 * R[6];
        Jp[3 * i + 1] = a0 * R[1] + a1 * R[4] + a2 * R[7];
        Jp[3 * i + 2] = a0 * R[2] + a1 * R[5] + a2 * R[8];
    }
}

// Same block, residual and point Jacobian only (first pass of the Schur elimination, where only
// V = sum Jp^T Jp and g = sum Jp^T r are needed).
CSLAM_HD void stereo_block_point(const CameraIntrinsics& c, const double* pose, const double* p,
                                 double u, double v, double d, const double* W, double* r,
                                 double* Jp) {
    double pc[3];
    transform_point(pose, p, pc);
    const double iz = 1.0 / pc[2];
    const double e0 = c.fu * pc[0] * iz + c.cu - u;
    const double e1 = c.fv * pc[1] * iz + c.cv - v;
    const double e2 = c.fu * c.b * iz - d;
    r[0] = W[0] * e0 + W[1] * e1 + W[2] * e2;
    r[1] = W[3] * e0 + W[4] * e1 + W[5] * e2;
    r[2] = W[6] * e0 + W[7] * e1 + W[8] * e2;
    const double iz2 = iz * iz;
    const double p00 = c.fu * iz, p02 = -c.fu * pc[0] * iz2;
    const double p11 = c.fv * iz, p12 = -c.fv * pc[1] * iz2;
    const double p22 = -c.fu * c.b * iz2;
    const double* R = pose + 3;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const double a0 = W[3 * i] * p00;
        const double a1 = W[3 * i + 1] * p11;
        const double a2 = W[3 * i] * p02 + W[3 * i + 1] * p12 + W[3 * i + 2] * p22;
        Jp[3 * i + 0] = a0 * R[0] + a1 * R[3] + a2 * R[6];
        Jp[3 * i + 1] = a0 * R[1] + a1 * R[4] + a2 * R[7];
        Jp[3 * i + 2] = a0 * R[2] + a1 * R[5] + a2 * R[8];
    }
}

// utils.hpp:28-31
CSLAM_HD double t_fabs(double a) { return (a >= 0.0) ? a : -a; }

// Sun-sensor block (sun_sensor_error.hpp:35-104).  obs_c / ref_g already unit length
// (constructor, :30-31).  r 2, J 2x6 row-major (translation columns are zero).
CSLAM_HD void sun_block(const double* pose, const double* obs_c, const double* ref_g,
                        const double* W2, double az_thresh, double zen_thresh, double* r,
                        double* J) {
    const double pi = atan(1.) * 4.;  // utils.hpp:13
    const double* R = pose + 3;
    const double x = R[0] * ref_g[0] + R[1] * ref_g[1] + R[2] * ref_g[2];
    const double y = R[3] * ref_g[0] + R[4] * ref_g[1] + R[5] * ref_g[2];
    const double z = R[6] * ref_g[0] + R[7] * ref_g[1] + R[8] * ref_g[2];
    const double expected_zen = acos(-y);
    const double expected_az = atan2(x, z);
    const double observed_zen = acos(-obs_c[1]);
    const double observed_az = atan2(obs_c[0], obs_c[2]);
    double res_az = expected_az - observed_az;
    double res_zen = expected_zen - observed_zen;
    if (res_az > pi) {
        res_az = res_az - 2 * pi;
    } else if (res_az < -pi) {
        res_az = res_az + 2 * pi;
    }
    // d az / d phi and d zen / d phi with d e / d phi = -[e]x
    const double h = 1.0 / (x * x + z * z);
    double daz[3] = {-x * y * h, 1.0, -z * y * h};
    const double s = 1.0 / sqrt(1.0 - y * y);
    double dzen[3] = {-z * s, 0.0, x * s};
    if (t_fabs(res_az) > az_thresh) {
        res_az = 0.;
        daz[0] = daz[1] = daz[2] = 0.0;
    }
    if (t_fabs(res_zen) > zen_thresh) {
        res_zen = 0.;
        dzen[0] = dzen[1] = dzen[2] = 0.0;
    }
    r[0] = W2[0] * res_az + W2[1] * res_zen;
    r[1] = W2[2] * res_az + W2[3] * res_zen;
    if (!J) return;
    for (int i = 0; i < 2; ++i) {
        J[6 * i + 0] = J[6 * i + 1] = J[6 * i + 2] = 0.0;
        for (int k = 0; k < 3; ++k) J[6 * i + 3 + k] = W2[2 * i] * daz[k] + W2[2 * i + 1] * dzen[k];
    }
}

// SO(3) log exactly as so3group.hpp:299-349 (atan2 form, first-order branch on |angle| <= eps)
CSLAM_HD void so3_log(const double* C, double* phi) {
    double axis[3] = {C[7] - C[5], C[2] - C[6], C[3] - C[1]};
    const double sin_angle = 0.5 * sqrt(axis[0] * axis[0] + axis[1] * axis[1] + axis[2] * axis[2]);
    const double cos_angle = 0.5 * ((C[0] + C[4] + C[8]) - 1.0);
    const double angle = atan2(sin_angle, cos_angle);
    if (t_fabs(angle) <= 2.220446049250313e-16) {
        phi[0] = 0.5 * (C[7] - C[5]);
        phi[1] = 0.5 * (C[2] - C[6]);
        phi[2] = 0.5 * (C[3] - C[1]);
        return;
    }
    for (int i = 0; i < 3; ++i) phi[i] = 0.5 * angle * axis[i] / sin_angle;
}

// Pose-prior block (pose_error.hpp:22-55): r = W6 [t_res ; Log(R_res)], T_res = T_ref T^-1.
//   d t_res / d eps = [ -R_res | 0 ],   d Log / d eps = [ 0 | -Jr^-1(theta) ]
// Jr^-1(theta) = I + 1/2 [theta]x + c(|theta|) [theta]x^2,
// c = 1/|theta|^2 - (1 + cos|theta|) / (2 |theta| sin|theta|)   (-> 1/12 as theta -> 0).
// On the first-order branch of the log the derivative of vee(R_res Exp(-phi) - I) is used,
// which is what differentiating that branch gives.
CSLAM_HD void prior_block(const double* pose, const double* Tref, const double* W6, double* r,
                          double* J) {
    const double* R = pose + 3;
    const double* Rr = Tref + 3;
    double Rres[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            Rres[3 * i + j] = Rr[3 * i] * R[3 * j] + Rr[3 * i + 1] * R[3 * j + 1] + Rr[3 * i + 2] * R[3 * j + 2];
    // T^-1 = (R^T, -(R^T t)) then compose: t_res = R_ref * (-(R^T t)) + t_ref  (se3group.hpp:152-183)
    double ti[3];
    for (int i = 0; i < 3; ++i) ti[i] = -(R[i] * pose[0] + R[3 + i] * pose[1] + R[6 + i] * pose[2]);
    double xi[6];
    for (int i = 0; i < 3; ++i) xi[i] = Rr[3 * i] * ti[0] + Rr[3 * i + 1] * ti[1] + Rr[3 * i + 2] * ti[2] + Tref[i];
    so3_log(Rres, xi + 3);
    for (int i = 0; i < 6; ++i) {
        double a = W6[6 * i] * xi[0];
        for (int k = 1; k < 6; ++k) a = a + W6[6 * i + k] * xi[k];
        r[i] = a;
    }
    if (!J) return;
    double D[36];
    for (int i = 0; i < 36; ++i) D[i] = 0.0;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) D[6 * i + j] = -Rres[3 * i + j];
    // rotation part
    double axis[3] = {Rres[7] - Rres[5], Rres[2] - Rres[6], Rres[3] - Rres[1]};
    const double sin_angle = 0.5 * sqrt(axis[0] * axis[0] + axis[1] * axis[1] + axis[2] * axis[2]);
    const double cos_angle = 0.5 * ((Rres[0] + Rres[4] + Rres[8]) - 1.0);
    const double angle = atan2(sin_angle, cos_angle);
    double M[9];
    if (t_fabs(angle) <= 2.220446049250313e-16) {
        // d vee(Rres (I - [phi]x) - I) / d phi_k = -vee(Rres [e_k]x)
        // Rres [e_k]x columns: [e_0]x = [[0,0,0],[0,0,-1],[0,1,0]] etc.
        for (int k = 0; k < 3; ++k) {
            double B[9];
            for (int i = 0; i < 3; ++i) {
                const double a = Rres[3 * i], b = Rres[3 * i + 1], c = Rres[3 * i + 2];
                if (k == 0) { B[3 * i] = 0.0; B[3 * i + 1] = c; B[3 * i + 2] = -b; }
                if (k == 1) { B[3 * i] = -c; B[3 * i + 1] = 0.0; B[3 * i + 2] = a; }
                if (k == 2) { B[3 * i] = b; B[3 * i + 1] = -a; B[3 * i + 2] = 0.0; }
            }
            M[0 + k] = 0.5 * (B[7] - B[5]);
            M[3 + k] = 0.5 * (B[2] - B[6]);
            M[6 + k] = 0.5 * (B[3] - B[1]);
        }
    } else {
        const double* th = xi + 3;
        const double t2 = th[0] * th[0] + th[1] * th[1] + th[2] * th[2];
        const double t = sqrt(t2);
        double c;
        if (t < 1e-2) {
            c = 1.0 / 12.0 + t2 / 720.0 + t2 * t2 / 30240.0;
        } else {
            c = 1.0 / t2 - (1.0 + cos(t)) / (2.0 * t * sin(t));
        }
        const double K[9] = {0.0, -th[2], th[1], th[2], 0.0, -th[0], -th[1], th[0], 0.0};
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) {
                const double k2 = K[3 * i] * K[j] + K[3 * i + 1] * K[3 + j] + K[3 * i + 2] * K[6 + j];
                M[3 * i + j] = ((i == j) ? 1.0 : 0.0) + 0.5 * K[3 * i + j] + c * k2;
            }
    }
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) D[6 * (i + 3) + 3 + j] = -M[3 * i + j];
    for (int i = 0; i < 6; ++i)
        for (int j = 0; j < 6; ++j) {
            double a = 0.0;
            for (int k = 0; k < 6; ++k) a += W6[6 * i + k] * D[6 * k + j];
            J[6 * i + j] = a;
        }
}

// SO(3) exp exactly as so3group.hpp:273-291
CSLAM_HD void so3_exp(const double* phi, double* R) {
    const double angle = sqrt(phi[0] * phi[0] + phi[1] * phi[1] + phi[2] * phi[2]);
    if (angle <= 2.220446049250313e-16) {
        R[0] = 1.0; R[1] = -phi[2]; R[2] = phi[1];
        R[3] = phi[2]; R[4] = 1.0; R[5] = -phi[0];
        R[6] = -phi[1]; R[7] = phi[0]; R[8] = 1.0;
        return;
    }
    const double ax[3] = {phi[0] / angle, phi[1] / angle, phi[2] / angle};
    const double cp = cos(angle), sp = sin(angle), omc = 1.0 - cp;
    const double K[9] = {0.0, -ax[2], ax[1], ax[2], 0.0, -ax[0], -ax[1], ax[0], 0.0};
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            R[3 * i + j] = cp * ((i == j) ? 1.0 : 0.0) + omc * ax[i] * ax[j] + sp * K[3 * i + j];
}

// SE3Perturbation on plain doubles: T' = exp(eps) T (perturbations.hpp:62, se3group.hpp:176-183)
CSLAM_HD void se3_plus(const double* pose, const double* eps, double* out) {
    double E[9];
    so3_exp(eps + 3, E);
    double o[12];
    for (int i = 0; i < 3; ++i) {
        o[i] = E[3 * i] * pose[0] + E[3 * i + 1] * pose[1] + E[3 * i + 2] * pose[2] + eps[i];
        for (int j = 0; j < 3; ++j)
            o[3 + 3 * i + j] = E[3 * i] * pose[3 + j] + E[3 * i + 1] * pose[6 + j] + E[3 * i + 2] * pose[9 + j];
    }
    for (int i = 0; i < 12; ++i) out[i] = o[i];
}

// UnitVectorPerturbation on plain doubles (perturbations.hpp:97-104):
//   x' = normalize(x + delta - (delta . x / |x|^2) x)
CSLAM_HD void unit_plus(const double* x, const double* dl, double* out) {
    const double s = (dl[0] * x[0] + dl[1] * x[1] + dl[2] * x[2]) / (x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);
    const double y0 = x[0] + dl[0] - s * x[0], y1 = x[1] + dl[1] - s * x[1], y2 = x[2] + dl[2] - s * x[2];
    const double n = sqrt(y0 * y0 + y1 * y1 + y2 * y2);
    out[0] = y0 / n;
    out[1] = y1 / n;
    out[2] = y2 / n;
}
// Its Jacobian at delta = 0 (what AutoDiffLocalParameterization yields, perturbations.hpp:110-111,
// declared 3 -> 3, rank 2):  (I - x x^T / |x|^2) / |x|.   Row-vector product g^T J:
CSLAM_HD void unit_plus_pullback(const double* x, const double* g, double* out) {
    const double n2 = x[0] * x[0] + x[1] * x[1] + x[2] * x[2];
    const double in = 1.0 / sqrt(n2);
    const double gx = (g[0] * x[0] + g[1] * x[1] + g[2] * x[2]) / n2;
    out[0] = (g[0] - gx * x[0]) * in;
    out[1] = (g[1] - gx * x[1]) * in;
    out[2] = (g[2] - gx * x[2]) * in;
}

// Normal block (normal_error.hpp:20-41): r = W (R n - n_obs).
//   Jc = W [ 0 | -[n_c]x ]   3x6,   Jn = W R (I - n n^T/|n|^2)/|n|   3x3 (through UnitVectorPerturbation)
CSLAM_HD void normal_block(const double* pose, const double* n, const double* obs, const double* W,
                           double* r, double* Jc, double* Jn) {
    const double* R = pose + 3;
    double nc[3];
    for (int i = 0; i < 3; ++i) nc[i] = R[3 * i] * n[0] + R[3 * i + 1] * n[1] + R[3 * i + 2] * n[2];
    const double e0 = nc[0] - obs[0], e1 = nc[1] - obs[1], e2 = nc[2] - obs[2];
    for (int i = 0; i < 3; ++i) r[i] = W[3 * i] * e0 + W[3 * i + 1] * e1 + W[3 * i + 2] * e2;
    if (!Jc) return;
    for (int i = 0; i < 3; ++i) {
        const double a0 = W[3 * i], a1 = W[3 * i + 1], a2 = W[3 * i + 2];
        Jc[6 * i + 0] = Jc[6 * i + 1] = Jc[6 * i + 2] = 0.0;
        Jc[6 * i + 3] = a2 * nc[1] - a1 * nc[2];
        Jc[6 * i + 4] = a0 * nc[2] - a2 * nc[0];
        Jc[6 * i + 5] = a1 * nc[0] - a0 * nc[1];
        const double g[3] = {a0 * R[0] + a1 * R[3] + a2 * R[6], a0 * R[1] + a1 * R[4] + a2 * R[7],
                             a0 * R[2] + a1 * R[5] + a2 * R[8]};
        unit_plus_pullback(n, g, Jn + 3 * i);
    }
}

// Intensity block (intensity_error_point_light.hpp:24-90 / intensity_error_directional_light.hpp:24-90
// -> point_light.hpp:76-90 / directional_light.hpp:82-91 -> phong.hpp:25-139):
//   I = clamp(kd max(0, l.n_c) + ks (m.c)^alpha, 0, 1),  r = w (I - I_obs)
// with p_c = R p + t, n_c = R n, light vector lv = R (l - p) (point light; t cancels) or R d
// (directional, normalised by the DirectionalLight constructor), l = lv/|lv|, camera direction
// c = -p_c/|p_c| (camera at the origin, intensity_error_point_light.hpp:83), m = 2 (n_c.l) n_c - l
// normalised.  Ambient is disabled (phong.hpp:31-33).  Branches as in the reference: diffuse 0 when
// l is not finite or l.n_c <= 0; specular 0 when |m|^2 <= 0 or m.c <= 0; clamp with the constant as
// first argument of fmax/fmin (utils.hpp:16-25) so ties take the constant (zero derivative).
// Jacobians (one row): pose tangent 6, point 3, normal 3 (through UnitVectorPerturbation),
// phong [ka, ks, alpha] 3, texture kd 1, light 3 (through UnitVectorPerturbation when directional).
CSLAM_HD void intensity_block(const double* pose, const double* p, const double* n, const double* phong,
                              double kd, const double* light, double colour, double w, bool directional,
                              double* r, double* Jc, double* Jp, double* Jn, double* Jk, double* Jt,
                              double* Jl) {
    const double* R = pose + 3;
    double pc[3], nc[3], lv[3];
    transform_point(pose, p, pc);
    for (int i = 0; i < 3; ++i) nc[i] = R[3 * i] * n[0] + R[3 * i + 1] * n[1] + R[3 * i + 2] * n[2];
    if (directional) {
        for (int i = 0; i < 3; ++i) lv[i] = R[3 * i] * light[0] + R[3 * i + 1] * light[1] + R[3 * i + 2] * light[2];
    } else {
        double lc[3];
        transform_point(pose, light, lc);
        for (int i = 0; i < 3; ++i) lv[i] = lc[i] - pc[i];
    }
    const double L = sqrt(lv[0] * lv[0] + lv[1] * lv[1] + lv[2] * lv[2]);
    const double l[3] = {lv[0] / L, lv[1] / L, lv[2] / L};
    const double cn = sqrt(pc[0] * pc[0] + pc[1] * pc[1] + pc[2] * pc[2]);
    const double c[3] = {-pc[0] / cn, -pc[1] / cn, -pc[2] / cn};
    const double ks = phong[1], alpha = phong[2];
    // gradients of the colour w.r.t. camera-frame quantities
    double g_nc[3] = {0, 0, 0}, g_l[3] = {0, 0, 0}, g_c[3] = {0, 0, 0};
    double d_kd = 0.0, d_ks = 0.0, d_alpha = 0.0;
    const double a = l[0] * nc[0] + l[1] * nc[1] + l[2] * nc[2];
    const bool finite_l = (l[0] - l[0] == 0.0) && (l[1] - l[1] == 0.0) && (l[2] - l[2] == 0.0);
    double col = 0.0;
    if (finite_l && !(a <= 0.0)) {
        col = kd * a;
        d_kd = a;
        for (int i = 0; i < 3; ++i) {
            g_nc[i] = kd * l[i];
            g_l[i] = kd * nc[i];
        }
    }
    double m[3] = {2.0 * a * nc[0] - l[0], 2.0 * a * nc[1] - l[1], 2.0 * a * nc[2] - l[2]};
    const double m2 = m[0] * m[0] + m[1] * m[1] + m[2] * m[2];
    if (!(m2 <= 0.0)) {
        const double mn = sqrt(m2);
        m[0] /= mn; m[1] /= mn; m[2] /= mn;
        const double s = m[0] * c[0] + m[1] * c[1] + m[2] * c[2];
        if (!(s <= 0.0)) {
            const double sa = pow(s, alpha);
            col += ks * sa;
            d_ks = sa;
            d_alpha = ks * sa * log(s);
            const double dS = ks * alpha * pow(s, alpha - 1.0);  // d spec / d s
            double u[3];                                          // d spec / d m (unnormalised m)
            for (int i = 0; i < 3; ++i) u[i] = dS * (c[i] - s * m[i]) / mn;
            const double un = u[0] * nc[0] + u[1] * nc[1] + u[2] * nc[2];
            for (int i = 0; i < 3; ++i) {
                g_nc[i] += 2.0 * a * u[i] + 2.0 * un * l[i];
                g_l[i] += 2.0 * un * nc[i] - u[i];
                g_c[i] = dS * m[i];
            }
        }
    }
    bool flat = false;  // clamp (phong.hpp:136-139)
    if (0.0 >= col) { col = 0.0; flat = true; }
    if (1.0 <= col) { col = 1.0; flat = true; }
    r[0] = w * (col - colour);
    if (!Jc) return;
    if (flat) {
        for (int i = 0; i < 6; ++i) Jc[i] = 0.0;
        for (int i = 0; i < 3; ++i) Jp[i] = Jn[i] = Jk[i] = Jl[i] = 0.0;
        Jt[0] = 0.0;
        return;
    }
    // through the normalisations: l = lv/|lv|, c = -p_c/|p_c|
    const double gl_l = g_l[0] * l[0] + g_l[1] * l[1] + g_l[2] * l[2];
    const double gc_c = g_c[0] * c[0] + g_c[1] * c[1] + g_c[2] * c[2];
    double g_lv[3], g_pc[3];
    for (int i = 0; i < 3; ++i) {
        g_lv[i] = (g_l[i] - gl_l * l[i]) / L;
        g_pc[i] = -(g_c[i] - gc_c * c[i]) / cn;
    }
    // pose tangent: translation g_pc; rotation p_c x g_pc + n_c x g_nc + lv x g_lv
    Jc[0] = w * g_pc[0];
    Jc[1] = w * g_pc[1];
    Jc[2] = w * g_pc[2];
    Jc[3] = w * ((pc[1] * g_pc[2] - pc[2] * g_pc[1]) + (nc[1] * g_nc[2] - nc[2] * g_nc[1]) + (lv[1] * g_lv[2] - lv[2] * g_lv[1]));
    Jc[4] = w * ((pc[2] * g_pc[0] - pc[0] * g_pc[2]) + (nc[2] * g_nc[0] - nc[0] * g_nc[2]) + (lv[2] * g_lv[0] - lv[0] * g_lv[2]));
    Jc[5] = w * ((pc[0] * g_pc[1] - pc[1] * g_pc[0]) + (nc[0] * g_nc[1] - nc[1] * g_nc[0]) + (lv[0] * g_lv[1] - lv[1] * g_lv[0]));
    double gp[3], gn[3], gl[3];
    for (int j = 0; j < 3; ++j) {
        const double dp0 = directional ? g_pc[0] : g_pc[0] - g_lv[0];
        const double dp1 = directional ? g_pc[1] : g_pc[1] - g_lv[1];
        const double dp2 = directional ? g_pc[2] : g_pc[2] - g_lv[2];
        gp[j] = dp0 * R[j] + dp1 * R[3 + j] + dp2 * R[6 + j];
        gn[j] = g_nc[0] * R[j] + g_nc[1] * R[3 + j] + g_nc[2] * R[6 + j];
        gl[j] = g_lv[0] * R[j] + g_lv[1] * R[3 + j] + g_lv[2] * R[6 + j];
    }
    double t3[3];
    unit_plus_pullback(n, gn, t3);
    for (int j = 0; j < 3; ++j) {
        Jp[j] = w * gp[j];
        Jn[j] = w * t3[j];
    }
    if (directional) {
        unit_plus_pullback(light, gl, t3);
        for (int j = 0; j < 3; ++j) Jl[j] = w * t3[j];
    } else {
        for (int j = 0; j < 3; ++j) Jl[j] = w * gl[j];
    }
    Jk[0] = 0.0;
    Jk[1] = w * d_ks;
    Jk[2] = w * d_alpha;
    Jt[0] = w * d_kd;
}

// ceres::HuberLoss + Corrector (rho'' <= 0 branch): residual and Jacobian scale sqrt(rho')
CSLAM_HD void huber_rho(double a, double s, double* rho0, double* sqrt_rho1) {
    const double b = a * a;
    if (s > b) {
        const double r = sqrt(s);
        *rho0 = 2.0 * a * r - b;
        const double r1 = a / r;
        *sqrt_rho1 = sqrt(r1 > 2.2250738585072014e-308 ? r1 : 2.2250738585072014e-308);
    } else {
        *rho0 = s;
        *sqrt_rho1 = 1.0;
    }
}

// Symmetric 3x3 inverse by adjugate; V given as 6 unique entries (00,01,02,11,12,22).
// Returns false when det is not positive/finite.
CSLAM_HD bool invert_sym3(const double* V, double* Vi) {
    const double a = V[0], b = V[1], c = V[2], d = V[3], e = V[4], f = V[5];
    const double c00 = d * f - e * e, c01 = c * e - b * f, c02 = b * e - c * d;
    const double det = a * c00 + b * c01 + c * c02;
    if (!(det > 0.0) || !(det < 1.7976931348623157e308)) return false;
    const double id = 1.0 / det;
    Vi[0] = c00 * id;
    Vi[1] = c01 * id;
    Vi[2] = c02 * id;
    Vi[3] = (a * f - c * c) * id;
    Vi[4] = (b * c - a * e) * id;
    Vi[5] = (a * d - b * b) * id;
    return true;
}

}  // namespace cslam
