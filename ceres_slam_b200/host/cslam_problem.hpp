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
    // problem.SetParameterization(pose, SE3Perturbation) is implied for every pose block
    void AddPoseBlock(double* pose12) { pose_index(pose12); }
    // problem.SetParameterBlockConstant(pose)
    void SetParameterBlockConstant(double* pose12) { constant_[pose_index(pose12)] = 1; }

    // ceres::Solve(options, &problem, &summary)
    void Solve(Summary* summary) {
        const uint32_t nc = uint32_t(pose_ptr_.size()), np = uint32_t(point_ptr_.size());
        std::vector<double> poses(12 * size_t(nc)), points(3 * size_t(np));
        for (uint32_t k = 0; k < nc; ++k) std::memcpy(&poses[12 * size_t(k)], pose_ptr_[k], 96);
        for (uint32_t j = 0; j < np; ++j) std::memcpy(&points[3 * size_t(j)], point_ptr_[j], 24);
        cslam_problem* p = nullptr;
        check(cslam_problem_create(&p, &options), p);
        try {
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
            cslam_summary s{};
            check(cslam_solve(p, &s), p);
            if (summary) summary->s = s;
        } catch (...) {
            cslam_problem_destroy(p);
            throw;
        }
        cslam_problem_destroy(p);
        for (uint32_t k = 0; k < nc; ++k) std::memcpy(pose_ptr_[k], &poses[12 * size_t(k)], 96);
        for (uint32_t j = 0; j < np; ++j) std::memcpy(point_ptr_[j], &points[3 * size_t(j)], 24);
    }

   private:
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
    std::map<double*, uint32_t> pose_id_, point_id_;
    std::vector<double*> pose_ptr_, point_ptr_;
    std::vector<uint8_t> constant_;
    std::vector<uint32_t> st_cam_, st_pt_, sun_cam_;
    std::vector<double> st_uvd_, st_W_, sun_obs_, sun_ref_, sun_W_;
    std::vector<Prior> priors_;
    double az_ = 1000., zen_ = 1000., huber_ = 0.;
};

}  // namespace cslam_b200
