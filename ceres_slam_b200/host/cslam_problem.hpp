// C++ host mirror of the seam the reference drivers use: the subset of ceres::Problem /
// ceres::Solve that `solveWindow` needs (tests/dataset_vo.cpp:22-85,
// tests/dataset_vo_sun.cpp:25-187), forwarding to the C ABI of the B200 back end
// (include/cslam_b200.h).  Parameter blocks are raw double* owned by the caller and updated in
// place, exactly as with Ceres; blocks are recognised by pointer identity.
#pragma once
#include <cstdint>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/cslam_b200.h"

namespace cslam_b200 {

struct Summary {
    cslam_summary s{};
    std::string BriefReport() const {  // ceres::Solver::Summary::BriefReport (dataset_vo.cpp:82)
        const char* term = s.termination_type == 0 ? "CONVERGENCE" : s.termination_type == 1 ? "NO_CONVERGENCE" : "FAILURE";
        char buf[256];
        std::snprintf(buf, sizeof(buf), "cslam_b200 Report: Iterations: %d, Initial cost: %e, Final cost: %e, Termination: %s",
                      s.num_iterations, s.initial_cost, s.final_cost, term);
        return buf;
    }
};

class Problem {
   public:
    Problem() {
        cslam_options_init(&options);
    }
    cslam_options options;  // same fields the drivers set on ceres::Solver::Options

    void SetCamera(double fu, double fv, double cu, double cv, double b) { cam_ = {fu, fv, cu, cv, b}; }

    // problem.AddResidualBlock(StereoReprojectionErrorAutomatic::Create(camera, obs, W), NULL, pose, point)
    void AddStereoBlock(double* pose12, double* point3, const double obs[3], const double W[9]) {
        st_cam_.push_back(pose_index(pose12));
        st_pt_.push_back(point_index(point3));
        st_uvd_.insert(st_uvd_.end(), obs, obs + 3);
        st_W_.insert(st_W_.end(), W, W + 9);
    }
    // problem.AddResidualBlock(SunSensorErrorAutomatic::Create(obs_c, ref_g, W2, az, zen), loss, pose)
    void AddSunBlock(double* pose12, const double obs_c[3], const double ref_g[3], const double W2[4], double az_thresh,
                     double zen_thresh, double huber) {
        sun_cam_.push_back(pose_index(pose12));
        sun_obs_.insert(sun_obs_.end(), obs_c, obs_c + 3);
        sun_ref_.insert(sun_ref_.end(), ref_g, ref_g + 3);
        sun_W_.insert(sun_W_.end(), W2, W2 + 4);
        az_ = az_thresh;
        zen_ = zen_thresh;
        huber_ = huber;
    }
    // problem.AddResidualBlock(PoseErrorAutomatic::Create(T_ref, W6), NULL, pose)
    void AddPosePrior(double* pose12, const double Tref[12], const double W6[36]) {
        Prior p;
        p.cam = pose_index(pose12);
        std::memcpy(p.Tref, Tref, 96);
        std::memcpy(p.W, W6, 288);
        priors_.push_back(p);
    }
    // One observation of dataset_ba_phong (tests/dataset_ba_phong.cpp:103-190): the intensity block
    //   AddResidualBlock(IntensityError*LightAutomatic::Create(I, w), NULL, pose, position, normal, phong, texture, light)
    // and the normal block
    //   AddResidualBlock(NormalErrorAutomatic::Create(n_obs, W), NULL, pose, normal)
    // of the same (pose, vertex).  Must follow the AddStereoBlock of that observation: the back end
    // pairs them one-to-one.  Shared blocks (material, texture, light) are recognised by pointer.
    void AddLightingBlocks(double* pose12, double* position3, double* normal3, double* phong3, double* texture1,
                           double* light3, double intensity, double int_stiffness, const double normal_obs[3],
                           const double W_normal[9]) {
        ph_cam_.push_back(pose_index(pose12));
        const uint32_t j = point_index(position3);
        ph_vtx_.push_back(j);
        if (normal_ptr_.size() <= j) {
            normal_ptr_.resize(j + 1, nullptr);
            vertex_mat_.resize(j + 1, 0);
            vertex_tex_.resize(j + 1, 0);
        }
        normal_ptr_[j] = normal3;
        vertex_mat_[j] = shared_index(phong3, mat_id_, mat_ptr_);
        vertex_tex_[j] = shared_index(texture1, tex_id_, tex_ptr_);
        light_ptr_ = light3;
        ph_int_.push_back(intensity);
        ph_nobs_.insert(ph_nobs_.end(), normal_obs, normal_obs + 3);
        int_stiffness_ = int_stiffness;
        std::memcpy(Wn_, W_normal, 72);
    }
    // problem.SetParameterBlockConstant / Variable on every vertex position (dataset_ba_phong.cpp:215-220, :236-239)
    void SetPointsConstant(bool constant) { points_constant_ = constant; }
    // problem.SetParameterBlockVariable(pose)
    void SetParameterBlockVariable(double* pose12) { constant_[pose_index(pose12)] = 0; }
    // drop the lighting blocks again (the stereo-only solve of stage 1 on the same Problem object)
    bool HasLighting() const { return !ph_cam_.empty(); }
    // problem.SetParameterization(light_dir, UnitVectorPerturbation) (dataset_ba_phong.cpp:199-203)
    void SetLightDirectional(bool directional) { directional_ = directional; }
    // problem.SetParameterLowerBound / UpperBound on every material / texture block (:143-181)
    void SetMaterialBounds(const double lo[3], const double hi[3]) {
        std::memcpy(mat_lo_, lo, 24);
        std::memcpy(mat_hi_, hi, 24);
        mat_bounded_ = true;
    }
    void SetTextureBounds(double lo, double hi) {
        tex_lo_ = lo;
        tex_hi_ = hi;
        tex_bounded_ = true;
    }
    // ceres::Covariance::Compute + GetCovarianceBlockInTangentSpace(pose, pose, cov) after Solve
    // (dataset_vo_sun.cpp:159-183); false when the computation failed (rank-deficient Jacobian)
    bool GetCovarianceBlockInTangentSpace(double* pose12, double* cov36) {
        if (!handle_) return false;
        return cslam_covariance_block(handle_, pose_index(pose12), cov36) == CSLAM_OK;
    }
    ~Problem() {
        if (handle_) cslam_problem_destroy(handle_);
    }
    Problem(const Problem&) = delete;
    Problem& operator=(const Problem&) = delete;

    // problem.SetParameterization(pose, SE3Perturbation) is implied for every pose block
    void AddPoseBlock(double* pose12) { pose_index(pose12); }
    // problem.SetParameterBlockConstant(pose)
    void SetParameterBlockConstant(double* pose12) { constant_[pose_index(pose12)] = 1; }

    // ceres::Solve(options, &problem, &summary)
    void Solve(Summary* summary) {
        Prepare();
        cslam_summary s{};
        check(cslam_solve(handle_, &s), handle_);
        if (summary) summary->s = s;
        Finish();
    }

    // Many independent problems in one call (scripts/ba_all_*.sh run many (trajectory x sun file) jobs; window w
    // of every job is independent of the others): cslam_solve_batch packs the window-eligible ones into one
    // launch.  Same effect on every problem as its own Solve().
    static void SolveBatch(const std::vector<Problem*>& problems, std::vector<Summary>* summaries) {
        if (problems.empty()) return;
        std::vector<cslam_problem*> handles;
        for (Problem* q : problems) {
            q->Prepare();
            handles.push_back(q->handle_);
        }
        std::vector<cslam_summary> sums(problems.size());
        check(cslam_solve_batch(handles.data(), int(handles.size()), sums.data()), handles[0]);
        if (summaries) {
            summaries->resize(problems.size());
            for (size_t i = 0; i < problems.size(); ++i) (*summaries)[i].s = sums[i];
        }
        for (Problem* q : problems) q->Finish();
    }

   private:
    // gather the blocks' current values and state the problem through the C ABI
    void Prepare() {
        const uint32_t nc = uint32_t(pose_ptr_.size()), np = uint32_t(point_ptr_.size());
        std::vector<double>&poses = poses_, &points = points_;
        poses.assign(12 * size_t(nc), 0.0);
        points.assign(3 * size_t(np), 0.0);
        for (uint32_t k = 0; k < nc; ++k) std::memcpy(&poses[12 * size_t(k)], pose_ptr_[k], 96);
        for (uint32_t j = 0; j < np; ++j) std::memcpy(&points[3 * size_t(j)], point_ptr_[j], 24);
        if (handle_) cslam_problem_destroy(handle_);
        handle_ = nullptr;
        cslam_problem* p = nullptr;
        check(cslam_problem_create(&p, &options), p);
        handle_ = p;
        const bool lighting = !ph_cam_.empty();
        check(cslam_set_camera(p, cam_.fu, cam_.fv, cam_.cu, cam_.cv, cam_.b), p);
        check(cslam_set_poses(p, nc, poses.data(), constant_.data()), p);
        check(cslam_set_points(p, np, points.data()), p);
        // one shared W if all blocks carry the same matrix (dataset_vo.cpp:29-32)
        bool shared = true;
        for (size_t i = 1; i < st_cam_.size() && shared; ++i)
            shared = std::memcmp(&st_W_[0], &st_W_[9 * i], 72) == 0;
        check(cslam_add_stereo(p, st_cam_.size(), st_cam_.data(), st_pt_.data(), st_uvd_.data(), st_W_.data(), shared ? 0 : 1), p);
        if (!sun_cam_.empty())
            check(cslam_add_sun(p, uint32_t(sun_cam_.size()), sun_cam_.data(), sun_obs_.data(), sun_ref_.data(), sun_W_.data(),
                                az_, zen_, huber_), p);
        for (auto& pr : priors_) check(cslam_add_pose_prior(p, pr.cam, pr.Tref, pr.W), p);
        if (lighting) {
            if (normal_ptr_.size() != np) throw std::runtime_error("cslam_b200: every vertex needs lighting blocks");
            normals_.assign(3 * size_t(np), 0.0);
            for (uint32_t j = 0; j < np; ++j) {
                if (!normal_ptr_[j]) throw std::runtime_error("cslam_b200: vertex without a normal block");
                std::memcpy(&normals_[3 * size_t(j)], normal_ptr_[j], 24);
            }
            mats_.assign(3 * mat_ptr_.size(), 0.0);
            for (size_t m = 0; m < mat_ptr_.size(); ++m) std::memcpy(&mats_[3 * m], mat_ptr_[m], 24);
            texs_.assign(tex_ptr_.size(), 0.0);
            for (size_t t = 0; t < tex_ptr_.size(); ++t) texs_[t] = *tex_ptr_[t];
            std::memcpy(light_, light_ptr_, 24);
            std::vector<double> kd_unused(np, 0.0);
            check(cslam_set_vertices(p, np, normals_.data(), kd_unused.data(), vertex_mat_.data()), p);
            check(cslam_set_textures(p, uint32_t(texs_.size()), texs_.data(), vertex_tex_.data()), p);
            check(cslam_set_materials(p, uint32_t(mat_ptr_.size()), mats_.data()), p);
            check(cslam_set_light(p, light_, directional_ ? 1 : 0), p);
            check(cslam_add_phong(p, ph_cam_.size(), ph_cam_.data(), ph_vtx_.data(), ph_int_.data(), int_stiffness_,
                                  ph_nobs_.data(), Wn_), p);
            check(cslam_set_points_constant(p, points_constant_ ? 1 : 0), p);
            if (mat_bounded_) check(cslam_set_bounds(p, 0, mat_lo_, mat_hi_), p);
            if (tex_bounded_) check(cslam_set_bounds(p, 1, &tex_lo_, &tex_hi_), p);
        }
    }
    // the library wrote the solution into poses_ / points_ / ...: back into the caller's blocks, like Ceres
    void Finish() {
        const uint32_t nc = uint32_t(pose_ptr_.size()), np = uint32_t(point_ptr_.size());
        for (uint32_t k = 0; k < nc; ++k) std::memcpy(pose_ptr_[k], &poses_[12 * size_t(k)], 96);
        for (uint32_t j = 0; j < np; ++j) std::memcpy(point_ptr_[j], &points_[3 * size_t(j)], 24);
        if (!ph_cam_.empty()) {
            for (uint32_t j = 0; j < np; ++j) std::memcpy(normal_ptr_[j], &normals_[3 * size_t(j)], 24);
            for (size_t m = 0; m < mat_ptr_.size(); ++m) std::memcpy(mat_ptr_[m], &mats_[3 * m], 24);
            for (size_t t = 0; t < tex_ptr_.size(); ++t) *tex_ptr_[t] = texs_[t];
            std::memcpy(light_ptr_, light_, 24);
        }
    }

    struct Cam {
        double fu = 1, fv = 1, cu = 0, cv = 0, b = 1;
    } cam_;
    struct Prior {
        uint32_t cam;
        double Tref[12], W[36];
    };
    static void check(cslam_status st, cslam_problem* p) {
        if (st != CSLAM_OK) throw std::runtime_error(std::string("cslam_b200: ") + (p ? cslam_last_error(p) : "create failed"));
    }
    uint32_t pose_index(double* ptr) {
        auto it = pose_id_.find(ptr);
        if (it != pose_id_.end()) return it->second;
        const uint32_t id = uint32_t(pose_ptr_.size());
        pose_id_[ptr] = id;
        pose_ptr_.push_back(ptr);
        constant_.push_back(0);
        return id;
    }
    uint32_t point_index(double* ptr) {
        auto it = point_id_.find(ptr);
        if (it != point_id_.end()) return it->second;
        const uint32_t id = uint32_t(point_ptr_.size());
        point_id_[ptr] = id;
        point_ptr_.push_back(ptr);
        return id;
    }
    static uint32_t shared_index(double* ptr, std::map<double*, uint32_t>& ids, std::vector<double*>& ptrs) {
        auto it = ids.find(ptr);
        if (it != ids.end()) return it->second;
        const uint32_t id = uint32_t(ptrs.size());
        ids[ptr] = id;
        ptrs.push_back(ptr);
        return id;
    }
    cslam_problem* handle_ = nullptr;      // kept after Solve for the covariance query
    std::vector<double> poses_, points_, normals_, mats_, texs_;
    double light_[3] = {0, 0, 0};
    // lighting blocks
    std::vector<uint32_t> ph_cam_, ph_vtx_, vertex_mat_, vertex_tex_;
    std::vector<double> ph_int_, ph_nobs_;
    std::vector<double*> normal_ptr_, mat_ptr_, tex_ptr_;
    std::map<double*, uint32_t> mat_id_, tex_id_;
    double* light_ptr_ = nullptr;
    double int_stiffness_ = 1.0, Wn_[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    bool directional_ = false, mat_bounded_ = false, tex_bounded_ = false, points_constant_ = false;
    double mat_lo_[3] = {0, 0, 0}, mat_hi_[3] = {0, 0, 0}, tex_lo_ = 0, tex_hi_ = 0;
    std::map<double*, uint32_t> pose_id_, point_id_;
    std::vector<double*> pose_ptr_, point_ptr_;
    std::vector<uint8_t> constant_;
    std::vector<uint32_t> st_cam_, st_pt_, sun_cam_;
    std::vector<double> st_uvd_, st_W_, sun_obs_, sun_ref_, sun_W_;
    std::vector<Prior> priors_;
    double az_ = 1000., zen_ = 1000., huber_ = 0.;
};

}  // namespace cslam_b200
