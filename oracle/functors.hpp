// ORACLE — TEST INFRASTRUCTURE ONLY (parity PINNED: checked block by block against the reference's own unmodified
// headers compiled in oracle/_ref, tests/test_ref_pin.py, 1e-14).
//
// Restatement of the reference's sensor / lighting models and Ceres cost functors as templates on
// the scalar type.  Running them on oracle::Jet<N> reproduces what ceres::AutoDiffCostFunction
// computes for the reference (exact forward-mode chain rule), which is the parity target
// BASELINE.json names ("match Ceres autodiff within 1e-10 relative").
#pragma once
#include "geometry.hpp"

namespace oracle {

// stereo_camera.hpp:159-163
struct Camera {
    double fu, fv, cu, cv, b;
};

// stereo_camera.hpp:77-84 — one reciprocal, then products
template <class T>
inline void camera_project(const Camera& c, const T* pt_c, T* obs) {
    T one_over_z = T(1.0) / pt_c[2];
    obs[0] = T(c.fu) * pt_c[0] * one_over_z + T(c.cu);
    obs[1] = T(c.fv) * pt_c[1] * one_over_z + T(c.cv);
    obs[2] = T(c.fu) * T(c.b) * one_over_z;
}
// stereo_camera.hpp:112-120
template <class T>
inline void camera_triangulate(const Camera& c, const T* obs, T* pt_c) {
    T b_over_d = T(c.b) / obs[2];
    T fu_over_fv = T(c.fu) / T(c.fv);
    pt_c[0] = (obs[0] - T(c.cu)) * b_over_d;
    pt_c[1] = (obs[1] - T(c.cv)) * b_over_d * fu_over_fv;
    pt_c[2] = T(c.fu) * b_over_d;
}

// ---------------------------------------------------------------------------------------------
// Phong shading — lighting/phong.hpp:25-104,136-139
// ---------------------------------------------------------------------------------------------
template <class T>
inline bool all_finite3(const T* v) {
    return std::isfinite(value_of(v[0])) && std::isfinite(value_of(v[1])) &&
           std::isfinite(value_of(v[2]));
}
// phong.hpp:59-74
template <class T>
inline T phong_diffuse(const T* normal, const T& texture_col, const T* light_dir) {
    if (!all_finite3(light_dir)) return T(0.0);
    T ldn = dot3(light_dir, normal);
    if (ldn <= T(0.0)) return T(0.0);
    return texture_col * ldn;
}
// phong.hpp:77-104
template <class T>
inline T phong_specular(const T* normal, const T& ks, const T& alpha, const T* light_dir,
                        const T* camera_dir) {
    T ndl = dot3(normal, light_dir);
    T mirror[3];
    for (int i = 0; i < 3; ++i) mirror[i] = T(2.0) * ndl * normal[i] - light_dir[i];
    if (dot3(mirror, mirror) <= T(0.0)) return T(0.0);
    T mn = norm3(mirror);
    for (int i = 0; i < 3; ++i) mirror[i] = mirror[i] / mn;
    T mdc = dot3(mirror, camera_dir);
    if (mdc <= T(0.0)) return T(0.0);
    return ks * pow(mdc, alpha);
}
// phong.hpp:25-51 — ambient term hard-disabled (:31-33); clamp via utils fmax/fmin with the
// constant as first argument (:136-139)
template <class T>
inline T phong_shade(const T* normal, const T* phong_params, const T& texture_col,
                     const T* light_dir, const T* camera_dir, const T& light_colour) {
    T ambient = T(0.0);
    T diffuse = phong_diffuse(normal, texture_col, light_dir);
    T specular = phong_specular(normal, phong_params[1], phong_params[2], light_dir, camera_dir);
    T col = light_colour * (ambient + diffuse + specular);
    col = t_fmax(T(0.0), col);
    col = t_fmin(T(1.0), col);
    return col;
}
template <class T>
inline void normalized3(const T* v, T* out) {
    T n = norm3(v);
    for (int i = 0; i < 3; ++i) out[i] = v[i] / n;
}
// lighting/point_light.hpp:76-90
template <class T>
inline T point_light_shade(const T* light_pos, const T* vpos, const T* vnormal,
                           const T* phong_params, const T& texture_col, const T* camera_position,
                           const T& light_colour) {
    T light_vec[3], camera_vec[3], light_dir[3], camera_dir[3];
    for (int i = 0; i < 3; ++i) {
        light_vec[i] = light_pos[i] - vpos[i];
        camera_vec[i] = camera_position[i] - vpos[i];
    }
    normalized3(light_vec, light_dir);
    normalized3(camera_vec, camera_dir);
    return phong_shade(vnormal, phong_params, texture_col, light_dir, camera_dir, light_colour);
}
// lighting/directional_light.hpp:51-54 (ctor normalises), :82-91
template <class T>
inline T directional_light_shade(const T* light_direction, const T* vpos, const T* vnormal,
                                 const T* phong_params, const T& texture_col,
                                 const T* camera_position, const T& light_colour) {
    T dir[3], camera_vec[3], camera_dir[3];
    normalized3(light_direction, dir);
    for (int i = 0; i < 3; ++i) camera_vec[i] = camera_position[i] - vpos[i];
    normalized3(camera_vec, camera_dir);
    return phong_shade(vnormal, phong_params, texture_col, dir, camera_dir, light_colour);
}

// ---------------------------------------------------------------------------------------------
// Cost functors.  Each exposes kNumResiduals, kNumBlocks, kSizes[] and
//   template <class T> bool operator()(T const* const* params, T* residuals) const
// ---------------------------------------------------------------------------------------------

// stereo_reprojection_error.hpp:27-55 ; blocks: pose(12), point(3) ; 3 residuals
struct StereoReprojectionError {
    static constexpr int kNumResiduals = 3;
    static constexpr int kNumBlocks = 2;
    static constexpr int kSizes[2] = {12, 3};
    static constexpr int kTotal = 15;
    Camera camera;
    double observation[3];
    double stiffness[9];  // row-major 3x3
    template <class T>
    bool operator()(T const* const* params, T* residuals) const {
        SE3<T> T_c_g = SE3<T>::from(params[0]);
        const T* pt_g = params[1];
        T pt_c[3];
        se3_transform_point(T_c_g, pt_g, pt_c);
        T predicted[3];
        camera_project(camera, pt_c, predicted);
        T e[3];
        for (int i = 0; i < 3; ++i) e[i] = predicted[i] - T(observation[i]);
        for (int i = 0; i < 3; ++i)
            residuals[i] = T(stiffness[3 * i]) * e[0] + T(stiffness[3 * i + 1]) * e[1] +
                           T(stiffness[3 * i + 2]) * e[2];
        return true;
    }
};

// sun_sensor_error.hpp:20-104 ; block: pose(12) ; 2 residuals
struct SunSensorError {
    static constexpr int kNumResiduals = 2;
    static constexpr int kNumBlocks = 1;
    static constexpr int kSizes[1] = {12};
    static constexpr int kTotal = 12;
    double observed_sun_dir_c[3];  // normalised at construction (:30)
    double expected_sun_dir_g[3];  // normalised at construction (:31)
    double stiffness[4];           // row-major 2x2
    double az_err_thresh, zen_err_thresh;
    void normalize_inputs() {
        double n = std::sqrt(dot3(observed_sun_dir_c, observed_sun_dir_c));
        for (int i = 0; i < 3; ++i) observed_sun_dir_c[i] /= n;
        n = std::sqrt(dot3(expected_sun_dir_g, expected_sun_dir_g));
        for (int i = 0; i < 3; ++i) expected_sun_dir_g[i] /= n;
    }
    template <class T>
    bool operator()(T const* const* params, T* residuals) const {
        const double pi = std::atan(1.) * 4.;  // utils.hpp:13
        SE3<T> T_c_g = SE3<T>::from(params[0]);
        T eg[3] = {T(expected_sun_dir_g[0]), T(expected_sun_dir_g[1]), T(expected_sun_dir_g[2])};
        T ec[3];
        se3_transform_vector(T_c_g, eg, ec);
        T oc[3] = {T(observed_sun_dir_c[0]), T(observed_sun_dir_c[1]), T(observed_sun_dir_c[2])};
        T expected_zen = acos(-ec[1]);
        T expected_az = atan2(ec[0], ec[2]);
        T observed_zen = acos(-oc[1]);
        T observed_az = atan2(oc[0], oc[2]);
        T residual_az = expected_az - observed_az;
        T residual_zen = expected_zen - observed_zen;
        if (residual_az > T(pi)) {
            residual_az = residual_az - T(2 * pi);
        } else if (residual_az < -T(pi)) {
            residual_az = residual_az + T(2 * pi);
        }
        if (t_fabs(residual_az) > T(az_err_thresh)) residual_az = T(0.);
        if (t_fabs(residual_zen) > T(zen_err_thresh)) residual_zen = T(0.);
        residuals[0] = T(stiffness[0]) * residual_az + T(stiffness[1]) * residual_zen;
        residuals[1] = T(stiffness[2]) * residual_az + T(stiffness[3]) * residual_zen;
        return true;
    }
};

// pose_error.hpp:17-55 ; block: pose(12) ; 6 residuals
struct PoseError {
    static constexpr int kNumResiduals = 6;
    static constexpr int kNumBlocks = 1;
    static constexpr int kSizes[1] = {12};
    static constexpr int kTotal = 12;
    double T_ref[12];
    double stiffness[36];  // row-major 6x6
    template <class T>
    bool operator()(T const* const* params, T* residuals) const {
        SE3<T> T_k_0 = SE3<T>::from(params[0]);
        SE3<T> Tref = SE3<T>::from(T_ref);
        SE3<T> T_residual = se3_mul(Tref, se3_inverse(T_k_0));
        T xi[6];
        se3_log(T_residual, xi);
        for (int i = 0; i < 6; ++i) {
            T acc = T(stiffness[6 * i]) * xi[0];
            for (int k = 1; k < 6; ++k) acc = acc + T(stiffness[6 * i + k]) * xi[k];
            residuals[i] = acc;
        }
        return true;
    }
};

// normal_error.hpp:16-42 ; blocks: pose(12), normal(3) ; 3 residuals
struct NormalError {
    static constexpr int kNumResiduals = 3;
    static constexpr int kNumBlocks = 2;
    static constexpr int kSizes[2] = {12, 3};
    static constexpr int kTotal = 15;
    double obs_normal_c[3];
    double stiffness[9];
    template <class T>
    bool operator()(T const* const* params, T* residuals) const {
        SE3<T> T_c_g = SE3<T>::from(params[0]);
        T normal_c[3];
        se3_transform_vector(T_c_g, params[1], normal_c);
        T e[3];
        for (int i = 0; i < 3; ++i) e[i] = normal_c[i] - T(obs_normal_c[i]);
        for (int i = 0; i < 3; ++i)
            residuals[i] = T(stiffness[3 * i]) * e[0] + T(stiffness[3 * i + 1]) * e[1] +
                           T(stiffness[3 * i + 2]) * e[2];
        return true;
    }
};

// intensity_error_point_light.hpp:24-96 / intensity_error_directional_light.hpp:24-96
// blocks: pose(12), point(3), normal(3), phong(3), texture(1), light(3) ; 1 residual
struct IntensityError {
    static constexpr int kNumResiduals = 1;
    static constexpr int kNumBlocks = 6;
    static constexpr int kSizes[6] = {12, 3, 3, 3, 1, 3};
    static constexpr int kTotal = 25;
    double colour;
    double stiffness;
    bool directional;
    template <class T>
    bool operator()(T const* const* params, T* residuals) const {
        SE3<T> T_c_g = SE3<T>::from(params[0]);
        T pt_c[3], normal_c[3];
        se3_transform_point(T_c_g, params[1], pt_c);
        se3_transform_vector(T_c_g, params[2], normal_c);
        const T* phong_params = params[3];
        T texture_col = params[4][0];
        T campos_c[3] = {T(0.0), T(0.0), T(0.0)};  // :83 camera sits at the origin
        T light_colour = T(1.0);
        T predicted;
        if (directional) {
            T lightdir_c[3];
            se3_transform_vector(T_c_g, params[5], lightdir_c);
            predicted = directional_light_shade(lightdir_c, pt_c, normal_c, phong_params,
                                                texture_col, campos_c, light_colour);
        } else {
            T lightpos_c[3];
            se3_transform_point(T_c_g, params[5], lightpos_c);
            predicted = point_light_shade(lightpos_c, pt_c, normal_c, phong_params, texture_col,
                                          campos_c, light_colour);
        }
        residuals[0] = T(stiffness) * (predicted - T(colour));
        return true;
    }
};

// ---------------------------------------------------------------------------------------------
// Manifold "plus" operations — perturbations.hpp
// ---------------------------------------------------------------------------------------------
// perturbations.hpp:45-65 — T' = exp(eps) * T
struct SE3Perturbation {
    static constexpr int kGlobal = 12;
    static constexpr int kLocal = 6;
    template <class T>
    bool operator()(const T* x, const T* delta, T* x_plus_delta) const {
        SE3<T> T_op = SE3<T>::from(x);
        SE3<T> T_new = se3_mul(se3_exp(delta), T_op);
        for (int i = 0; i < 12; ++i) x_plus_delta[i] = T_new.d[i];
        return true;
    }
};
// perturbations.hpp:87-107 — normalize(x + delta - (delta.x/|x|^2) x); declared 3 -> 3 (:110-111)
struct UnitVectorPerturbation {
    static constexpr int kGlobal = 3;
    static constexpr int kLocal = 3;
    template <class T>
    bool operator()(const T* x, const T* delta, T* x_plus_delta) const {
        T s = dot3(delta, x) / dot3(x, x);
        T y[3];
        for (int i = 0; i < 3; ++i) y[i] = x[i] + (delta[i] - s * x[i]);
        T n = norm3(y);
        for (int i = 0; i < 3; ++i) x_plus_delta[i] = y[i] / n;
        return true;
    }
};

// ---------------------------------------------------------------------------------------------
// Autodiff drivers (what AutoDiffCostFunction::Evaluate and
// AutoDiffLocalParameterization::ComputeJacobian do [Ceres 1.x])
// ---------------------------------------------------------------------------------------------
// jacobians[k] is row-major kNumResiduals x kSizes[k] (ambient), may be nullptr.
template <class F>
inline bool autodiff_cost(const F& f, double const* const* params, double* residuals,
                          double** jacobians) {
    constexpr int N = F::kTotal;
    using J = Jet<N>;
    J x[N];
    const J* ptrs[F::kNumBlocks];
    int off = 0;
    for (int k = 0; k < F::kNumBlocks; ++k) {
        ptrs[k] = x + off;
        for (int i = 0; i < F::kSizes[k]; ++i) x[off + i] = J(params[k][i], off + i);
        off += F::kSizes[k];
    }
    J out[F::kNumResiduals];
    if (!f(ptrs, out)) return false;
    for (int r = 0; r < F::kNumResiduals; ++r) residuals[r] = out[r].a;
    if (jacobians) {
        off = 0;
        for (int k = 0; k < F::kNumBlocks; ++k) {
            if (jacobians[k])
                for (int r = 0; r < F::kNumResiduals; ++r)
                    for (int i = 0; i < F::kSizes[k]; ++i)
                        jacobians[k][r * F::kSizes[k] + i] = out[r].v[off + i];
            off += F::kSizes[k];
        }
    }
    return true;
}
template <class F>
inline bool eval_cost(const F& f, double const* const* params, double* residuals) {
    return f(params, residuals);
}
// d Plus(x, delta) / d delta at delta = 0, row-major kGlobal x kLocal
template <class P>
inline void autodiff_plus_jacobian(const P& plus, const double* x, double* jac) {
    using J = Jet<P::kLocal>;
    J xj[P::kGlobal], dj[P::kLocal], out[P::kGlobal];
    for (int i = 0; i < P::kGlobal; ++i) xj[i] = J(x[i]);
    for (int i = 0; i < P::kLocal; ++i) dj[i] = J(0.0, i);
    plus(xj, dj, out);
    for (int i = 0; i < P::kGlobal; ++i)
        for (int k = 0; k < P::kLocal; ++k) jac[i * P::kLocal + k] = out[i].v[k];
}

}  // namespace oracle
