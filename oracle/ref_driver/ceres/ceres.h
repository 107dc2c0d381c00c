// TEST INFRASTRUCTURE — a Ceres-API FACADE over the C ABI, so that the reference's own driver sources
// (/root/reference/tests/dataset_vo.cpp) compile UNMODIFIED and run their `solveWindow` — ceres::Problem,
// AddResidualBlock, SetParameterization, SetParameterBlockConstant, ceres::Solver::Options, ceres::Solve —
// on this repo's back end.  It is what INTEGRATION.md §2 describes as the zero-change binding: every Ceres call of the
// path is mapped onto `cslam_b200::Problem` (ceres_slam_b200/host/cslam_problem.hpp), which states the problem through
// include/cslam_b200.h.  Which library answers those calls is decided by the build: the product library, or — with
// -DCSLAM_FACADE_ORACLE, the only variant that can run without a GPU — the CPU oracle (liboracle.so, same signatures
// under the prefix `cslam_oracle_`).
//
// Found as <ceres/ceres.h> ahead of oracle/ref_standin/ceres/ceres.h (include order: -Ioracle/ref_driver first),
// whose Jet / AutoDiffCostFunction / AutoDiffLocalParameterization stand-ins it re-exports.  The cost functors keep
// their data in private members; this header includes the functor headers with those members visible
// (`#define private public` around the include, nothing else sees it).  A maintainer's binding would add accessors.
// Covers the plain stereo driver (config 1).  Never part of the product.
#ifndef CSLAM_REF_DRIVER_CERES_FACADE
#define CSLAM_REF_DRIVER_CERES_FACADE

#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <typeinfo>
#include <vector>

#include "../../ref_standin/ceres/ceres.h"

#include <Eigen/Core>
#include <ceres_slam/geometry/geometry.hpp>
#include <ceres_slam/stereo_camera.hpp>
#include <ceres_slam/utils/utils.hpp>
#define private public
#include <ceres_slam/stereo_reprojection_error.hpp>
#undef private

#ifdef CSLAM_FACADE_ORACLE
// the oracle answers the C ABI calls cslam_problem.hpp makes (same signatures, oracle/oracle_capi.cpp)
#define cslam_options_init cslam_oracle_options_init
#define cslam_problem_create cslam_oracle_problem_create
#define cslam_problem_destroy cslam_oracle_problem_destroy
#define cslam_last_error cslam_oracle_last_error
#define cslam_set_camera cslam_oracle_set_camera
#define cslam_set_poses cslam_oracle_set_poses
#define cslam_set_points cslam_oracle_set_points
#define cslam_add_stereo cslam_oracle_add_stereo
#define cslam_add_sun cslam_oracle_add_sun
#define cslam_add_pose_prior cslam_oracle_add_pose_prior
#define cslam_solve cslam_oracle_solve
#define cslam_set_points_constant cslam_oracle_set_points_constant
#define cslam_add_phong cslam_oracle_add_phong
#define cslam_set_bounds cslam_oracle_set_bounds
#define cslam_set_light cslam_oracle_set_light
#define cslam_set_materials cslam_oracle_set_materials
#define cslam_set_textures cslam_oracle_set_textures
#define cslam_set_vertices cslam_oracle_set_vertices
#endif
#include "../../../ceres_slam_b200/host/cslam_problem.hpp"

namespace ceres {

class LossFunction {
   public:
    virtual ~LossFunction() {}
};
typedef void* ResidualBlockId;

struct Solver {
    struct Options {
        bool minimizer_progress_to_stdout = false;
        int num_threads = 1;
        int num_linear_solver_threads = 1;
        int max_num_iterations = 50;        // Ceres' default; the drivers set 1000
        bool use_nonmonotonic_steps = false;
    };
    struct Summary {
        cslam_b200::Summary inner;
        std::string BriefReport() const { return inner.BriefReport(); }
        std::string FullReport() const { return inner.BriefReport(); }
    };
};

class Problem {
   public:
    // AddResidualBlock(StereoReprojectionErrorAutomatic::Create(camera, obs, W), NULL, pose, point): dataset_vo.cpp:51
    ResidualBlockId AddResidualBlock(CostFunction* cost, LossFunction* loss, double* x0, double* x1) {
        owned_.emplace_back(cost);
        if (cost->functor_type() == typeid(ceres_slam::StereoReprojectionErrorAutomatic)) {
            if (loss) throw std::runtime_error("facade: a loss on a stereo block is not part of the path");
            const auto* f = static_cast<const ceres_slam::StereoReprojectionErrorAutomatic*>(cost->functor_ptr());
            inner_.SetCamera(f->camera_->fu(), f->camera_->fv(), f->camera_->cu(), f->camera_->cv(), f->camera_->b());
            double obs[3], W[9];
            for (int r = 0; r < 3; ++r) {
                obs[r] = f->observation_(r);
                for (int c = 0; c < 3; ++c) W[3 * r + c] = f->stiffness_(r, c);
            }
            note_pose(x0);
            inner_.AddStereoBlock(x0, x1, obs, W);
            ++n_stereo_;
            return cost;
        }
        throw std::runtime_error(std::string("facade: cost functor not on the path: ") + cost->functor_type().name());
    }
    void SetParameterization(double* x, LocalParameterization* lp) {
        if (lp->GlobalSize() != 12 || lp->LocalSize() != 6) throw std::runtime_error("facade: only SE3Perturbation on 12-blocks");
        if (!lp_owned_ || lp_owned_.get() != lp) lp_owned_.reset(lp);   // Ceres takes ownership (one object, many blocks)
        note_pose(x);
        inner_.AddPoseBlock(x);
    }
    void SetParameterBlockConstant(double* x) {
        note_pose(x);
        inner_.SetParameterBlockConstant(x);
    }

    cslam_b200::Problem inner_;
    std::vector<double*> poses_;   // pose blocks in order of first appearance
    size_t n_stereo_ = 0;

   private:
    void note_pose(double* x) {
        for (double* p : poses_)
            if (p == x) return;
        poses_.push_back(x);
    }
    std::vector<std::unique_ptr<CostFunction>> owned_;
    std::unique_ptr<LocalParameterization> lp_owned_;
};

// One line per solve into $CSLAM_FACADE_TRACE (full precision): what was solved and what came back
inline void facade_trace(const Problem& p, const cslam_summary& s) {
    const char* path = std::getenv("CSLAM_FACADE_TRACE");
    if (!path) return;
    std::ofstream f(path, std::ios::app);
    f << std::setprecision(17) << "{\"n_poses\": " << p.poses_.size() << ", \"n_stereo\": " << p.n_stereo_
      << ", \"iterations\": " << s.num_iterations << ", \"initial_cost\": " << s.initial_cost << ", \"final_cost\": " << s.final_cost
      << ", \"termination\": " << s.termination_type << ", \"poses\": [";
    for (size_t i = 0; i < p.poses_.size(); ++i)
        for (int k = 0; k < 12; ++k) f << (i + k ? ", " : "") << p.poses_[i][k];
    f << "]}\n";
}

inline void Solve(const Solver::Options& options, Problem* problem, Solver::Summary* summary) {
    cslam_options& o = problem->inner_.options;          // defaults = Ceres' (cslam_options_init)
    o.max_num_iterations = options.max_num_iterations;   // dataset_vo.cpp:69
    o.use_nonmonotonic_steps = options.use_nonmonotonic_steps ? 1 : 0;  // :70
    o.num_threads = options.num_threads;                 // :67 (the GPU back end ignores it)
    if (problem->n_stereo_ == 0) {
        // Ceres solves an empty problem trivially; the C ABI wants at least a camera
        cslam_summary s{};
        if (summary) summary->inner.s = s;
        facade_trace(*problem, s);
        return;
    }
    cslam_b200::Summary inner;
    problem->inner_.Solve(&inner);
    if (summary) summary->inner = inner;
    facade_trace(*problem, inner.s);
}

}  // namespace ceres
#endif  // CSLAM_REF_DRIVER_CERES_FACADE
