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
#include <iostream>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

#include <ceres_slam/dataset_problem.hpp>
#include <ceres_slam/dataset_problem_phong.hpp>
#include <ceres_slam/dataset_problem_sun.hpp>
#include <ceres_slam/geometry/geometry.hpp>
#include <ceres_slam/intensity_error_directional_light.hpp>
#include <ceres_slam/intensity_error_point_light.hpp>
#include <ceres_slam/lighting/lighting.hpp>
#include <ceres_slam/normal_error.hpp>
#include <ceres_slam/perturbations.hpp>
#include <ceres_slam/point_cloud_aligner.hpp>
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
// ---- front end: PointCloudAligner (src/ceres_slam/point_cloud_aligner.cpp, compiled unmodified from
// where it lies as a second translation unit: oracle/Makefile) ---------------------------------------
// compute_transformation_and_inliers (:64-136) per pose pair, as compute_initial_guess calls it
// (dataset_problem.cpp:232-234); same argument order as cslam_oracle_ransac_align / cslam_ransac_align.
// The draws are the real std::mt19937(42) + std::uniform_int_distribution<uint> of this compiler, the
// SVD is the stand-in's (ref_standin/Eigen/Core); `rng_variant` is ignored.
int cslam_ref_ransac_align(int, uint32_t n_pairs, const uint32_t* offsets, const double* pts0, const double* pts1,
                           const double* intr5, uint32_t num_iters, double thresh, int, double* T12_out,
                           uint8_t* inlier_out, uint32_t* n_inliers_out) {
    Camera::ConstPtr cam = std::make_shared<const Camera>(intr5[0], intr5[1], intr5[2], intr5[3], intr5[4]);
    cs::PointCloudAligner aligner;
    for (uint32_t p = 0; p < n_pairs; ++p) {
        const uint32_t o = offsets[p], n = offsets[p + 1] - o;
        std::vector<Pt> a, b;
        for (uint32_t i = 0; i < n; ++i) {
            a.push_back(Pt(Eigen::Map<const Pt>(pts0 + 3 * size_t(o + i))));
            b.push_back(Pt(Eigen::Map<const Pt>(pts1 + 3 * size_t(o + i))));
        }
        SE3 T;
        std::vector<uint> in = aligner.compute_transformation_and_inliers(T, a, b, cam, num_iters, thresh);
        std::memcpy(T12_out + 12 * size_t(p), T.data(), 24);
        std::memcpy(T12_out + 12 * size_t(p) + 3, T.rotation().data(), 72);
        if (inlier_out) {
            for (uint32_t i = 0; i < n; ++i) inlier_out[o + i] = 0;
            for (uint i : in) inlier_out[o + i] = 1;
        }
        if (n_inliers_out) n_inliers_out[p] = uint32_t(in.size());
    }
    return 0;
}
// compute_transformation (:12-62) on n >= 3 correspondences: T_1_0 as [t | R row-major]
void cslam_ref_kabsch(uint32_t n, const double* pts0, const double* pts1, double* T12) {
    cs::PointCloudAligner aligner;
    std::vector<Pt> a, b;
    for (uint32_t i = 0; i < n; ++i) {
        a.push_back(Pt(Eigen::Map<const Pt>(pts0 + 3 * size_t(i))));
        b.push_back(Pt(Eigen::Map<const Pt>(pts1 + 3 * size_t(i))));
    }
    SE3 T = aligner.compute_transformation(a, b);
    std::memcpy(T12, T.data(), 24);
    std::memcpy(T12 + 3, T.rotation().data(), 72);
}
// the stand-in SVD on its own: A (row-major 3x3) -> U, s, V (row-major), for the test of the stand-in
void cslam_ref_svd3(const double* A9, double* U9, double* s3, double* V9) {
    Eigen::Matrix3d A;
    load(A, A9);
    Eigen::JacobiSVD<Eigen::Matrix3d> svd(A, Eigen::ComputeThinU | Eigen::ComputeThinV);
    for (int i = 0; i < 3; ++i) {
        s3[i] = svd.singularValue(i);
        for (int j = 0; j < 3; ++j) {
            U9[3 * i + j] = svd.matrixU()(i, j);
            V9[3 * i + j] = svd.matrixV()(i, j);
        }
    }
}

// ---- the three DatasetProblem classes (src/ceres_slam/dataset_problem{,_sun,_phong}.cpp, compiled unmodified
// as further translation units): CSV readers, compute_initial_guess(k1, k2), CSV writers — everything the
// drivers do around solveWindow (dataset_vo.cpp:107-136, dataset_vo_sun.cpp:232-290, dataset_ba_phong.cpp:300-340).
// kind: 0 = DatasetProblem, 1 = DatasetProblemSun, 2 = DatasetProblemPhong.
namespace {
struct RefDataset {
    int kind = 0;
    cs::DatasetProblem vo;
    cs::DatasetProblemSun sun;
    cs::DatasetProblemPhong phong;
    explicit RefDataset(int k, bool dir_light) : kind(k), phong(dir_light) {}
};
// the readers narrate on std::cerr / std::cout; keep the test output clean
struct Quiet {
    std::ostringstream sink;   // (declared first: the stream buffers below are taken from it)
    std::streambuf *o, *e;
    Quiet() : o(std::cout.rdbuf(sink.rdbuf())), e(std::cerr.rdbuf(sink.rdbuf())) {}
    ~Quiet() {
        std::cout.rdbuf(o);
        std::cerr.rdbuf(e);
    }
};
void put_pose(const SE3& T, double* P12) {
    std::memcpy(P12, T.data(), 24);
    std::memcpy(P12 + 3, T.rotation().data(), 72);
}
}  // namespace

void* cslam_ref_dataset_open(int kind, const char* f1, const char* f2, const char* f3, int dir_light) {
    Quiet q;
    std::unique_ptr<RefDataset> d(new RefDataset(kind, dir_light != 0));
    bool ok = false;
    try {
        if (kind == 0) ok = d->vo.read_csv(f1);
        if (kind == 1) ok = d->sun.read_csv(f1, f2, f3);
        if (kind == 2) ok = d->phong.read_csv(f1);
    } catch (...) {
        ok = false;
    }
    return ok ? d.release() : nullptr;
}
void cslam_ref_dataset_close(void* h) { delete static_cast<RefDataset*>(h); }
void cslam_ref_dataset_dims(void* h, uint32_t* n_states, uint32_t* n_points, uint64_t* n_obs, uint32_t* n_materials) {
    RefDataset& d = *static_cast<RefDataset*>(h);
    *n_materials = 0;
    if (d.kind == 0) { *n_states = d.vo.num_states; *n_points = d.vo.num_points; *n_obs = d.vo.stereo_obs_list.size(); }
    if (d.kind == 1) { *n_states = d.sun.num_states; *n_points = d.sun.num_points; *n_obs = d.sun.stereo_obs_list.size(); }
    if (d.kind == 2) {
        *n_states = d.phong.num_states; *n_points = d.phong.num_vertices; *n_obs = d.phong.stereo_obs_list.size();
        *n_materials = d.phong.num_materials;
    }
}
// what the readers stored: per observation the state index (the run of equal ids / timestamps it belongs to, as
// obs_indices_at_state groups them), the point id and (u, v, d); intrinsics; the shared variance (kinds 0, 2)
void cslam_ref_dataset_observations(void* h, uint32_t* state_of_obs, uint32_t* point_ids, double* uvd, double* intr5,
                                    double* var3) {
    RefDataset& d = *static_cast<RefDataset*>(h);
    auto fill = [&](auto& p, const std::vector<uint>& ids, uint n_states) {
        for (size_t i = 0; i < p.stereo_obs_list.size(); ++i) {
            point_ids[i] = ids[i];
            for (int c = 0; c < 3; ++c) uvd[3 * i + c] = p.stereo_obs_list[i](c);
        }
        for (uint k = 0; k < n_states; ++k) {
            std::vector<uint> idx;
            try { idx = p.obs_indices_at_state(k); } catch (...) { break; }
            for (uint i : idx) state_of_obs[i] = k;
        }
        intr5[0] = p.camera->fu(); intr5[1] = p.camera->fv(); intr5[2] = p.camera->cu(); intr5[3] = p.camera->cv();
        intr5[4] = p.camera->b();
    };
    if (d.kind == 0) { fill(d.vo, d.vo.point_ids, d.vo.num_states); for (int c = 0; c < 3; ++c) var3[c] = d.vo.stereo_obs_var(c); }
    if (d.kind == 1) fill(d.sun, d.sun.point_ids, d.sun.num_states);
    if (d.kind == 2) { fill(d.phong, d.phong.vertex_ids, d.phong.num_states); for (int c = 0; c < 3; ++c) var3[c] = d.phong.stereo_obs_var(c); }
}
// kind 1: per-observation stereo covariances (9 each, row-major), per state: has-sun flag, observed sun direction
// (camera frame), its 2x2 covariance, reference direction (global frame)
void cslam_ref_dataset_sun_data(void* h, double* stereo_covars9, uint8_t* has_sun, double* sun_obs3, double* sun_covar4,
                                double* sun_dir_g3) {
    cs::DatasetProblemSun& p = static_cast<RefDataset*>(h)->sun;
    for (size_t i = 0; i < p.stereo_obs_covars.size(); ++i)
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) stereo_covars9[9 * i + 3 * r + c] = p.stereo_obs_covars[i](r, c);
    for (uint k = 0; k < p.num_states; ++k) {
        has_sun[k] = p.state_has_sun_obs[k] ? 1 : 0;
        for (int c = 0; c < 3; ++c) {
            sun_obs3[3 * k + c] = has_sun[k] ? p.sun_obs_list[k](c) : 0.0;
            sun_dir_g3[3 * k + c] = has_sun[k] ? p.sun_dir_g[k](c) : 0.0;
        }
        for (int r = 0; r < 2; ++r)
            for (int c = 0; c < 2; ++c) sun_covar4[4 * k + 2 * r + c] = has_sun[k] ? p.sun_obs_covars[k](r, c) : 0.0;
    }
}
// kind 2: per observation material id, intensity, observed normal; the shared normal variance, intensity variance,
// the light (position, or unit direction with dir_light)
void cslam_ref_dataset_phong_data(void* h, uint32_t* material_ids, double* intensities, double* normal_obs3, double* normal_var3,
                                  double* int_var, double* light3) {
    cs::DatasetProblemPhong& p = static_cast<RefDataset*>(h)->phong;
    for (size_t i = 0; i < p.int_list.size(); ++i) {
        material_ids[i] = p.material_ids[i];
        intensities[i] = p.int_list[i];
        for (int c = 0; c < 3; ++c) normal_obs3[3 * i + c] = p.normal_obs_list[i](c);
    }
    for (int c = 0; c < 3; ++c) {
        normal_var3[c] = p.normal_obs_var(c);
        light3[c] = p.directional_light ? p.light_dir(c) : p.light_pos(c);
    }
    *int_var = p.int_var;
}
// compute_initial_guess(k1, k2); returns 0 when the sun variant gave up (fewer than 3 inliers), 1 otherwise
int cslam_ref_dataset_initial_guess(void* h, uint32_t k1, uint32_t k2) {
    Quiet q;
    RefDataset& d = *static_cast<RefDataset*>(h);
    if (d.kind == 0) d.vo.compute_initial_guess(k1, k2);
    if (d.kind == 1) return d.sun.compute_initial_guess(k1, k2) ? 1 : 0;
    if (d.kind == 2) d.phong.compute_initial_guess(k1, k2);
    return 1;
}
void cslam_ref_dataset_reset_points(void* h) {
    RefDataset& d = *static_cast<RefDataset*>(h);
    if (d.kind == 0) d.vo.reset_points();
    if (d.kind == 1) d.sun.reset_points();
}
// current state: poses [t | R] per state, point positions and initialised flags; kind 2 also the vertex normals,
// Phong parameters (ka, ks, exponent — Material::phong_params order) and texture per INITIALISED vertex
void cslam_ref_dataset_state(void* h, double* poses12, double* points3, uint8_t* initialized, double* normals3,
                             double* phong3, double* texture1) {
    RefDataset& d = *static_cast<RefDataset*>(h);
    auto put = [&](auto& p, auto& pts, const std::vector<bool>& init) {
        for (size_t k = 0; k < p.poses.size(); ++k) put_pose(p.poses[k], poses12 + 12 * k);
        for (size_t j = 0; j < init.size(); ++j) initialized[j] = init[j] ? 1 : 0;
        (void)pts;
    };
    if (d.kind == 0) {
        put(d.vo, d.vo.map_points, d.vo.initialized_point);
        for (size_t j = 0; j < d.vo.map_points.size(); ++j)
            for (int c = 0; c < 3; ++c) points3[3 * j + c] = d.vo.initialized_point[j] ? d.vo.map_points[j](c) : 0.0;
    }
    if (d.kind == 1) {
        put(d.sun, d.sun.map_points, d.sun.initialized_point);
        for (size_t j = 0; j < d.sun.map_points.size(); ++j)
            for (int c = 0; c < 3; ++c) points3[3 * j + c] = d.sun.initialized_point[j] ? d.sun.map_points[j](c) : 0.0;
    }
    if (d.kind == 2) {
        put(d.phong, d.phong.map_vertices, d.phong.initialized_vertex);
        for (size_t j = 0; j < d.phong.map_vertices.size(); ++j) {
            const bool in = d.phong.initialized_vertex[j];
            for (int c = 0; c < 3; ++c) {
                points3[3 * j + c] = in ? d.phong.map_vertices[j].position()(c) : 0.0;
                if (normals3) normals3[3 * j + c] = in ? d.phong.map_vertices[j].normal()(c) : 0.0;
                if (phong3) phong3[3 * j + c] = in ? d.phong.map_vertices[j].material()->phong_params()(c) : 0.0;
            }
            if (texture1) texture1[j] = in ? d.phong.map_vertices[j].texture()->col() : 0.0;
        }
    }
}
// overwrite the poses (what solveWindow leaves behind is outside this library; lets a test write a known state)
void cslam_ref_dataset_set_pose(void* h, uint32_t k, const double* P12) {
    RefDataset& d = *static_cast<RefDataset*>(h);
    SE3 T = SE3(Eigen::Map<const SE3>(P12));
    if (d.kind == 0) d.vo.poses[k] = T;
    if (d.kind == 1) d.sun.poses[k] = T;
    if (d.kind == 2) d.phong.poses[k] = T;
}
int cslam_ref_dataset_write(void* h, const char* filename) {
    Quiet q;
    RefDataset& d = *static_cast<RefDataset*>(h);
    if (d.kind == 0) return d.vo.write_csv(filename) ? 1 : 0;
    if (d.kind == 1) return d.sun.write_csv(filename) ? 1 : 0;
    return d.phong.write_csv(filename) ? 1 : 0;
}

}  // extern "C"
