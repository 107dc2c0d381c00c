// TEST INFRASTRUCTURE — oracle/_ref: the REFERENCE'S OWN, UNMODIFIED headers behind a C API.
//
// Compiled by `make -C oracle ref` with -I/root/reference/include (the sources stay where they
// lie; nothing of them is copied into this repo) and -Ioracle/ref_standin, which supplies
// stand-ins for the two libraries the image lacks: the subset of Eigen the headers use and
// Ceres' public Jet / AutoDiffCostFunction / AutoDiffLocalParameterization (see those files for
// what they restate).  Everything below goes through the reference's own factory functions —
// `XxxErrorAutomatic::Create(...)->Evaluate(...)`, `XxxPerturbation::Create()->Plus /
// ComputeJacobian(...)` — i.e. the calls ceres::Problem makes for the blocks the drivers add
// (dataset_vo.cpp:51-58, dataset_vo_sun.cpp:68-129, dataset_ba_phong.cpp:65-205).
//
// Used by tests/test_ref_pin.py to pin oracle/ (the restatement) and, through the committed
// vectors it generates (tests/golden/ref_blocks.json), the CUDA path.  Never linked or loaded by
// the product.
#include <cstdint>
#include <cstring>
#include <memory>

#include <ceres_slam/geometry/geometry.hpp>
#include <ceres_slam/intensity_error_directional_light.hpp>
#include <ceres_slam/intensity_error_point_light.hpp>
#include <ceres_slam/lighting/lighting.hpp>
#include <ceres_slam/normal_error.hpp>
#include <ceres_slam/perturbations.hpp>
#include <ceres_slam/pose_error.hpp>
#include <ceres_slam/stereo_camera.hpp>
#include <ceres_slam/stereo_reprojection_error.hpp>
#include <ceres_slam/sun_sensor_error.hpp>

namespace cs = ceres_slam;

namespace {
typedef cs::StereoCamera<double> Camera;
typedef cs::SE3Group<double> SE3;
typedef cs::SO3Group<double> SO3;
typedef cs::Vector3D<double> Vec;
typedef cs::Point3D<double> Pt;

// J_tangent (rows x local) = J_ambient (rows x global, row-major) * dPlus/ddelta (global x local):
// what ceres::Problem forms from a cost function and the block's LocalParameterization.
void chain(const double* Ja, const double* P, int rows, int global, int local, double* Jt) {
    for (int r = 0; r < rows; ++r)
        for (int c = 0; c < local; ++c) {
            double s = 0;
            for (int k = 0; k < global; ++k) s += Ja[r * global + k] * P[k * local + c];
            Jt[r * local + c] = s;
        }
}
void se3_plus_jacobian(const double* pose12, double* P72) {
    std::unique_ptr<ceres::LocalParameterization> lp(cs::SE3Perturbation::Create());  // perturbations.hpp:67-73
    lp->ComputeJacobian(pose12, P72);
}
void unit_plus_jacobian(const double* x3, double* U9) {
    std::unique_ptr<ceres::LocalParameterization> lp(cs::UnitVectorPerturbation::Create());  // :108-112
    lp->ComputeJacobian(x3, U9);
}
template <typename M>
void load(M& m, const double* p) {  // row-major fill
    for (int i = 0; i < m.rows(); ++i)
        for (int j = 0; j < m.cols(); ++j) m(i, j) = p[i * m.cols() + j];
}
}  // namespace

extern "C" {

const char* cslam_ref_describe() {
    return "ceres-slam reference headers (unmodified, /root/reference/include) + Eigen/Jet stand-ins";
}

// ---- geometry (so3group.hpp, se3group.hpp) ---------------------------------------------------
void cslam_ref_so3_exp(const double* phi3, double* R9) {
    SO3::TangentVector phi(phi3[0], phi3[1], phi3[2]);
    SO3 C = SO3::exp(phi);
    std::memcpy(R9, C.data(), 72);
}
void cslam_ref_so3_log(const double* R9, double* phi3) {
    Eigen::Map<const SO3> C(R9);
    SO3::TangentVector phi = SO3::log(SO3(C));
    for (int i = 0; i < 3; ++i) phi3[i] = phi(i);
}
void cslam_ref_so3_wedge(const double* phi3, double* M9) {
    SO3::TransformationMatrix m = SO3::wedge(SO3::TangentVector(phi3[0], phi3[1], phi3[2]));
    std::memcpy(M9, m.data(), 72);
}
void cslam_ref_se3_exp(const double* xi6, double* T12) {
    SE3::TangentVector xi;
    for (int i = 0; i < 6; ++i) xi(i) = xi6[i];
    SE3 T = SE3::exp(xi);
    std::memcpy(T12, T.data(), 24);
    std::memcpy(T12 + 3, T.rotation().data(), 72);
}
void cslam_ref_se3_log(const double* T12, double* xi6) {
    Eigen::Map<const SE3> T(T12);
    SE3::TangentVector xi = SE3::log(SE3(T));
    for (int i = 0; i < 6; ++i) xi6[i] = xi(i);
}
void cslam_ref_se3_compose(const double* A12, const double* B12, double* C12) {
    Eigen::Map<const SE3> A(A12), B(B12);
    SE3 C = A * SE3(B);
    std::memcpy(C12, C.data(), 24);
    std::memcpy(C12 + 3, C.rotation().data(), 72);
}
void cslam_ref_se3_inverse(const double* A12, double* C12) {
    Eigen::Map<const SE3> A(A12);
    SE3 C = A.inverse();
    std::memcpy(C12, C.data(), 24);
    std::memcpy(C12 + 3, C.rotation().data(), 72);
}
void cslam_ref_se3_transform_point(const double* T12, const double* p3, double* out3, double* J36) {
    Eigen::Map<const SE3> T(T12);
    Eigen::Map<const Pt> p(p3);
    SE3::TransformJacobian J;
    Pt q = T.transform(p, J36 ? &J : nullptr);
    for (int i = 0; i < 3; ++i) out3[i] = q(i);
    if (J36) std::memcpy(J36, J.data(), 36 * 8);
}
void cslam_ref_se3_adjoint(const double* T12, double* Ad36) {
    Eigen::Map<const SE3> T(T12);
    SE3::AdjointMatrix Ad = T.adjoint();
    std::memcpy(Ad36, Ad.data(), 36 * 8);
}
// ---- camera (stereo_camera.hpp:79-140) -------------------------------------------------------
void cslam_ref_camera_project(const double* intr5, const double* pt_c, double* uvd, double* J9) {
    Camera cam(intr5[0], intr5[1], intr5[2], intr5[3], intr5[4]);
    Camera::ProjectionJacobian J;
    Camera::Observation o = cam.project(Pt(Eigen::Map<const Pt>(pt_c)), J9 ? &J : nullptr);
    for (int i = 0; i < 3; ++i) uvd[i] = o(i);
    if (J9) std::memcpy(J9, J.data(), 72);
}
void cslam_ref_camera_triangulate(const double* intr5, const double* uvd, double* pt_c, double* J9) {
    Camera cam(intr5[0], intr5[1], intr5[2], intr5[3], intr5[4]);
    Camera::TriangulationJacobian J;
    Pt p = cam.triangulate(Camera::Observation(uvd[0], uvd[1], uvd[2]), J9 ? &J : nullptr);
    for (int i = 0; i < 3; ++i) pt_c[i] = p(i);
    if (J9) std::memcpy(J9, J.data(), 72);
}
// ---- plus operations (perturbations.hpp) -----------------------------------------------------
void cslam_ref_se3_plus(const double* x12, const double* d6, double* out12) {
    std::unique_ptr<ceres::LocalParameterization> lp(cs::SE3Perturbation::Create());
    lp->Plus(x12, d6, out12);
}
void cslam_ref_se3_plus_jacobian(const double* x12, double* J72) { se3_plus_jacobian(x12, J72); }
void cslam_ref_so3_plus(const double* x9, const double* d3, double* out9) {
    std::unique_ptr<ceres::LocalParameterization> lp(cs::SO3Perturbation::Create());
    lp->Plus(x9, d3, out9);
}
void cslam_ref_so3_plus_jacobian(const double* x9, double* J27) {
    std::unique_ptr<ceres::LocalParameterization> lp(cs::SO3Perturbation::Create());
    lp->ComputeJacobian(x9, J27);
}
void cslam_ref_unit_plus(const double* x3, const double* d3, double* out3) {
    std::unique_ptr<ceres::LocalParameterization> lp(cs::UnitVectorPerturbation::Create());
    lp->Plus(x3, d3, out3);
}
void cslam_ref_unit_plus_jacobian(const double* x3, double* J9) { unit_plus_jacobian(x3, J9); }

// ---- residual blocks, tangent-space Jacobians ------------------------------------------------
// n stereo blocks as dataset_vo.cpp:40-58 / dataset_vo_sun.cpp:52-70 add them.
// r[3n]; Jpose[18n] (3x6 per block); Jpoint[9n] (3x3); optional ambient Jpose_amb[36n] (3x12).
int cslam_ref_stereo_blocks(uint64_t n, const double* intr5, const uint32_t* cam, const uint32_t* pt,
                            const double* uvd, const double* W, int W_per_obs, const double* poses12,
                            const double* points3, double* r, double* Jpose, double* Jpoint, double* Jpose_amb) {
    Camera::ConstPtr camera = std::make_shared<const Camera>(intr5[0], intr5[1], intr5[2], intr5[3], intr5[4]);
    for (uint64_t i = 0; i < n; ++i) {
        Camera::Observation obs(uvd[3 * i], uvd[3 * i + 1], uvd[3 * i + 2]);
        Camera::ObservationCovariance stiffness;
        load(stiffness, W_per_obs ? W + 9 * i : W);
        std::unique_ptr<ceres::CostFunction> f(cs::StereoReprojectionErrorAutomatic::Create(camera, obs, stiffness));
        const double* params[2] = {poses12 + 12 * cam[i], points3 + 3 * pt[i]};
        double Ja[36], P[72];
        double* jac[2] = {Ja, Jpoint + 9 * i};
        if (!f->Evaluate(params, r + 3 * i, jac)) return 1;
        se3_plus_jacobian(params[0], P);
        chain(Ja, P, 3, 12, 6, Jpose + 18 * i);
        if (Jpose_amb) std::memcpy(Jpose_amb + 36 * i, Ja, sizeof(Ja));
    }
    return 0;
}
// n sun blocks (dataset_vo_sun.cpp:78-99; the Huber loss is Ceres', not the functor's). r[2n], J[12n].
int cslam_ref_sun_blocks(uint32_t n, const uint32_t* cam, const double* obs_c, const double* ref_g,
                         const double* W4, double az_thresh, double zen_thresh, const double* poses12, double* r,
                         double* J) {
    cs::SunSensorErrorAutomatic::ResidualCovariance stiffness;
    load(stiffness, W4);
    for (uint32_t i = 0; i < n; ++i) {
        Vec o = Vec(Eigen::Map<const Vec>(obs_c + 3 * i)), e = Vec(Eigen::Map<const Vec>(ref_g + 3 * i));
        std::unique_ptr<ceres::CostFunction> f(
            cs::SunSensorErrorAutomatic::Create(o, e, stiffness, az_thresh, zen_thresh));
        const double* params[1] = {poses12 + 12 * cam[i]};
        double Ja[24], P[72];
        double* jac[1] = {Ja};
        if (!f->Evaluate(params, r + 2 * i, jac)) return 1;
        se3_plus_jacobian(params[0], P);
        chain(Ja, P, 2, 12, 6, J + 12 * i);
    }
    return 0;
}
// the pose prior of dataset_vo_sun.cpp:109-124. r[6], J[36].
int cslam_ref_prior_block(const double* pose12, const double* Tref12, const double* W36, double* r, double* J) {
    SE3 Tref = SE3(Eigen::Map<const SE3>(Tref12));
    SE3::AdjointMatrix stiffness;
    load(stiffness, W36);
    std::unique_ptr<ceres::CostFunction> f(cs::PoseErrorAutomatic::Create(Tref, stiffness));
    const double* params[1] = {pose12};
    double Ja[72], P[72];
    double* jac[1] = {Ja};
    if (!f->Evaluate(params, r, jac)) return 1;
    se3_plus_jacobian(pose12, P);
    chain(Ja, P, 6, 12, 6, J);
    return 0;
}
// dataset_ba_phong.cpp:100-127 (intensity) — same argument order as cslam_oracle_intensity_block
int cslam_ref_intensity_block(const double* pose12, const double* pt3, const double* n3, const double* phong3,
                              const double* tex1, const double* light3, double colour, double stiffness,
                              int directional, double* r, double* J_pose, double* J_point, double* J_normal,
                              double* J_phong, double* J_tex, double* J_light) {
    std::unique_ptr<ceres::CostFunction> f(
        directional ? cs::IntensityErrorDirectionalLightAutomatic::Create(colour, stiffness)
                    : cs::IntensityErrorPointLightAutomatic::Create(colour, stiffness));
    const double* params[6] = {pose12, pt3, n3, phong3, tex1, light3};
    double Ja[12], Jn[3], Jl[3], P[72], Un[9], Ul[9];
    double* jac[6] = {Ja, J_point, Jn, J_phong, J_tex, Jl};
    if (!f->Evaluate(params, r, jac)) return 1;
    se3_plus_jacobian(pose12, P);
    chain(Ja, P, 1, 12, 6, J_pose);
    unit_plus_jacobian(n3, Un);                        // dataset_ba_phong.cpp:193
    chain(Jn, Un, 1, 3, 3, J_normal);
    if (directional) {
        unit_plus_jacobian(light3, Ul);                // dataset_ba_phong.cpp:202
        chain(Jl, Ul, 1, 3, 3, J_light);
    } else {
        std::memcpy(J_light, Jl, 24);
    }
    return 0;
}
// dataset_ba_phong.cpp:128-135 (normal) — same argument order as cslam_oracle_normal_block
int cslam_ref_normal_block(const double* pose12, const double* n3, const double* obs3, const double* W9, double* r,
                           double* J_pose, double* J_normal) {
    Vec obs = Vec(Eigen::Map<const Vec>(obs3));
    Vec::Covariance stiffness;
    load(stiffness, W9);
    std::unique_ptr<ceres::CostFunction> f(cs::NormalErrorAutomatic::Create(obs, stiffness));
    const double* params[2] = {pose12, n3};
    double Ja[36], Jn[9], P[72], Un[9];
    double* jac[2] = {Ja, Jn};
    if (!f->Evaluate(params, r, jac)) return 1;
    se3_plus_jacobian(pose12, P);
    chain(Ja, P, 3, 12, 6, J_pose);
    unit_plus_jacobian(n3, Un);
    chain(Jn, Un, 3, 3, 3, J_normal);
    return 0;
}
// PointLight::shade with the camera at `campos` (light_test.cpp:65-68)
double cslam_ref_point_light_shade(const double* light_pos, const double* vpos, const double* vnormal,
                                   const double* phong3, double texture, const double* campos) {
    typedef cs::Material<double> Mat;
    typedef cs::Texture<double> Tex;
    Mat::PhongParams pp;
    load(pp, phong3);
    Mat::Ptr m = std::make_shared<Mat>(pp);
    Tex::Ptr t = std::make_shared<Tex>(texture);
    cs::Vertex3D<double> v(Pt(Eigen::Map<const Pt>(vpos)), Vec(Eigen::Map<const Vec>(vnormal)), m, t);
    Pt lp = Pt(Eigen::Map<const Pt>(light_pos));
    double col = 1.0;
    cs::PointLight<double> light(lp, col);
    return light.shade(v, Vec(Eigen::Map<const Vec>(campos)));
}

}  // extern "C"
