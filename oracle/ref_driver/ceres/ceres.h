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
// Covers the plain stereo driver (config 1) and the sun-sensor driver (config 2: HuberLoss, pose prior, DOGLEG options,
// ceres::Covariance).  Never part of the product.
#ifndef CSLAM_REF_DRIVER_CERES_FACADE
#define CSLAM_REF_DRIVER_CERES_FACADE

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <limits>
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
#include <ceres_slam/intensity_error_directional_light.hpp>
#include <ceres_slam/intensity_error_point_light.hpp>
#include <ceres_slam/normal_error.hpp>
#include <ceres_slam/pose_error.hpp>
#include <ceres_slam/stereo_reprojection_error.hpp>
#include <ceres_slam/sun_sensor_error.hpp>
#undef private

#ifdef CSLAM_FACADE_ORACLE
// the oracle answers the C ABI calls cslam_problem.hpp makes (same signatures, oracle/oracle_capi.cpp); with
// CSLAM_FACADE_ABI_TRACE through the recording layer of ../abi_trace.cpp
#ifdef CSLAM_FACADE_ABI_TRACE
#define CSLAM_REMAP_PREFIX cslam_trace_
#else
#define CSLAM_REMAP_PREFIX cslam_oracle_
#endif
#include "../abi_remap.h"
#endif
#include "../../../ceres_slam_b200/host/cslam_problem.hpp"

namespace ceres {

class LossFunction {
   public:
    virtual ~LossFunction() {}
    virtual void Evaluate(double s, double rho[3]) const = 0;   // rho(s), rho'(s), rho''(s), s = |r|^2
};
// rho(s) = s for s <= a^2, 2 a sqrt(s) - a^2 beyond (Ceres' documented HuberLoss)
class HuberLoss : public LossFunction {
   public:
    explicit HuberLoss(double a) : a_(a), b_(a * a) {}
    void Evaluate(double s, double rho[3]) const override {
        if (s > b_) {
            const double r = std::sqrt(s);
            rho[0] = 2.0 * a_ * r - b_;
            rho[1] = std::max(std::numeric_limits<double>::min(), a_ / r);
            rho[2] = -rho[1] / (2.0 * s);
        } else {
            rho[0] = s;
            rho[1] = 1.0;
            rho[2] = 0.0;
        }
    }
    double a_, b_;
};
typedef void* ResidualBlockId;

enum TrustRegionStrategyType { LEVENBERG_MARQUARDT = 0, DOGLEG = 1 };
enum DoglegType { TRADITIONAL_DOGLEG = 0, SUBSPACE_DOGLEG = 1 };
enum SparseLinearAlgebraLibraryType { SUITE_SPARSE, CX_SPARSE, EIGEN_SPARSE, NO_SPARSE };
enum CovarianceAlgorithmType { DENSE_SVD, SPARSE_QR };
enum LinearSolverType { DENSE_NORMAL_CHOLESKY, DENSE_QR, SPARSE_NORMAL_CHOLESKY, DENSE_SCHUR, SPARSE_SCHUR, ITERATIVE_SCHUR, CGNR };

struct Solver {
    struct Options {
        bool minimizer_progress_to_stdout = false;
        int num_threads = 1;
        int num_linear_solver_threads = 1;
        int max_num_iterations = 50;        // Ceres' default; the drivers set 1000
        bool use_nonmonotonic_steps = false;
        TrustRegionStrategyType trust_region_strategy_type = LEVENBERG_MARQUARDT;
        DoglegType dogleg_type = TRADITIONAL_DOGLEG;
        LinearSolverType linear_solver_type = SPARSE_NORMAL_CHOLESKY;
    };
    struct Summary {
        cslam_b200::Summary inner;
        std::string BriefReport() const { return inner.BriefReport(); }
        std::string FullReport() const { return inner.BriefReport(); }
    };
};

// Records what the driver states; every Solve builds a fresh cslam_b200::Problem from the record as it stands then
// (dataset_ba_phong's multi-stage option solves, adds the lighting blocks, freezes and thaws blocks, solves again —
// :94-98, :207-252).
class Problem {
   public:
    struct Block {
        CostFunction* cost;
        LossFunction* loss;
        std::vector<double*> x;
    };
    // AddResidualBlock(StereoReprojectionErrorAutomatic::Create(camera, obs, W), NULL, pose, point): dataset_vo.cpp:51;
    // NormalErrorAutomatic::Create(n_obs, W), NULL, pose, normal: dataset_ba_phong.cpp:128-135
    ResidualBlockId AddResidualBlock(CostFunction* cost, LossFunction* loss, double* x0, double* x1) {
        return add(cost, loss, {x0, x1});
    }
    // one-block residuals: SunSensorErrorAutomatic (+ HuberLoss) dataset_vo_sun.cpp:83-99, PoseErrorAutomatic :113-117
    ResidualBlockId AddResidualBlock(CostFunction* cost, LossFunction* loss, double* x0) { return add(cost, loss, {x0}); }
    // IntensityError{Point,Directional}LightAutomatic: pose, position, normal, phong, texture, light (dataset_ba_phong.cpp:103-126)
    ResidualBlockId AddResidualBlock(CostFunction* cost, LossFunction* loss, double* x0, double* x1, double* x2, double* x3,
                                     double* x4, double* x5) {
        return add(cost, loss, {x0, x1, x2, x3, x4, x5});
    }
    void SetParameterization(double* x, LocalParameterization* lp) {
        if (lp->GlobalSize() == 12 && lp->LocalSize() == 6) {
            if (se3_.get() != lp) se3_.reset(lp);        // Ceres takes ownership (one object, many blocks)
            note_pose(x);
        } else if (lp->GlobalSize() == 3 && lp->LocalSize() == 3) {
            if (unit_.get() != lp) unit_.reset(lp);
            unit_blocks_.push_back(x);
        } else {
            throw std::runtime_error("facade: parameterization not on the path");
        }
    }
    void SetParameterBlockConstant(double* x) {
        if (!is_constant(x)) constant_.push_back(x);
    }
    void SetParameterBlockVariable(double* x) { constant_.erase(std::remove(constant_.begin(), constant_.end(), x), constant_.end()); }
    void SetParameterLowerBound(double* x, int i, double v) { lower_[x][i] = v; }
    void SetParameterUpperBound(double* x, int i, double v) { upper_[x][i] = v; }

    bool is_constant(const double* x) const { return std::find(constant_.begin(), constant_.end(), x) != constant_.end(); }
    bool has_unit_parameterization(const double* x) const {
        return std::find(unit_blocks_.begin(), unit_blocks_.end(), x) != unit_blocks_.end();
    }
    const LocalParameterization* lp() const { return se3_.get(); }
    static bool is(const Block& b, const std::type_info& t) { return b.cost->functor_type() == t; }

    // State the recorded problem through cslam_b200::Problem.  Returns false when there is nothing to solve.
    bool build(cslam_b200::Problem& q) const {
        using namespace ceres_slam;
        size_t n_res = 0;
        for (double* x : poses_) q.AddPoseBlock(x);
        // normal block of a (pose, normal) pair, to go with the intensity block of the same observation
        std::map<std::pair<double*, double*>, const Block*> normal_of;
        for (const Block& b : blocks_)
            if (is(b, typeid(NormalErrorAutomatic))) normal_of[{b.x[0], b.x[1]}] = &b;
        bool lighting = false, directional = false;
        std::vector<double*> positions;
        for (const Block& b : blocks_) {
            if (is(b, typeid(StereoReprojectionErrorAutomatic))) {
                if (b.loss) throw std::runtime_error("facade: a loss on a stereo block is not part of the path");
                const auto* f = static_cast<const StereoReprojectionErrorAutomatic*>(b.cost->functor_ptr());
                q.SetCamera(f->camera_->fu(), f->camera_->fv(), f->camera_->cu(), f->camera_->cv(), f->camera_->b());
                double obs[3], W[9];
                for (int r = 0; r < 3; ++r) {
                    obs[r] = f->observation_(r);
                    for (int c = 0; c < 3; ++c) W[3 * r + c] = f->stiffness_(r, c);
                }
                q.AddStereoBlock(b.x[0], b.x[1], obs, W);
                positions.push_back(b.x[1]);
                ++n_res;
            } else if (is(b, typeid(SunSensorErrorAutomatic))) {
                const auto* f = static_cast<const SunSensorErrorAutomatic*>(b.cost->functor_ptr());
                double obs[3], ref[3], W2[4];
                for (int r = 0; r < 3; ++r) obs[r] = f->observed_sun_dir_c_(r), ref[r] = f->expected_sun_dir_g_(r);
                for (int r = 0; r < 2; ++r)
                    for (int c = 0; c < 2; ++c) W2[2 * r + c] = f->stiffness_(r, c);
                double huber = 0.0;
                if (b.loss) {
                    const HuberLoss* h = dynamic_cast<const HuberLoss*>(b.loss);
                    if (!h) throw std::runtime_error("facade: only HuberLoss is on the path");
                    huber = h->a_;
                }
                q.AddSunBlock(b.x[0], obs, ref, W2, f->az_err_thresh_, f->zen_err_thresh_, huber);
                ++n_res;
            } else if (is(b, typeid(PoseErrorAutomatic))) {
                if (b.loss) throw std::runtime_error("facade: a loss on the pose prior is not part of the path");
                const auto* f = static_cast<const PoseErrorAutomatic*>(b.cost->functor_ptr());
                double Tref[12], W6[36];
                for (int r = 0; r < 3; ++r) {
                    Tref[r] = f->T_k_0_ref_.translation()(r);
                    for (int c = 0; c < 3; ++c) Tref[3 + 3 * r + c] = f->T_k_0_ref_.rotation().matrix()(r, c);
                }
                for (int r = 0; r < 6; ++r)
                    for (int c = 0; c < 6; ++c) W6[6 * r + c] = f->stiffness_(r, c);
                q.AddPosePrior(b.x[0], Tref, W6);
                ++n_res;
            }
        }
        // lighting blocks after the stereo blocks, in the same observation order (the back end pairs them one to one)
        for (const Block& b : blocks_) {
            const bool point = is(b, typeid(IntensityErrorPointLightAutomatic));
            const bool dir = is(b, typeid(IntensityErrorDirectionalLightAutomatic));
            if (!point && !dir) continue;
            double intensity, stiffness;
            if (point) {
                const auto* f = static_cast<const IntensityErrorPointLightAutomatic*>(b.cost->functor_ptr());
                intensity = f->colour_, stiffness = f->stiffness_;
            } else {
                const auto* f = static_cast<const IntensityErrorDirectionalLightAutomatic*>(b.cost->functor_ptr());
                intensity = f->colour_, stiffness = f->stiffness_;
            }
            auto it = normal_of.find({b.x[0], b.x[2]});
            if (it == normal_of.end()) throw std::runtime_error("facade: intensity block without its normal block");
            const auto* fn = static_cast<const NormalErrorAutomatic*>(it->second->cost->functor_ptr());
            double nobs[3], Wn[9];
            for (int r = 0; r < 3; ++r) {
                nobs[r] = fn->obs_normal_c_(r);
                for (int c = 0; c < 3; ++c) Wn[3 * r + c] = fn->stiffness_(r, c);
            }
            if (!has_unit_parameterization(b.x[2])) throw std::runtime_error("facade: a normal without UnitVectorPerturbation");
            q.AddLightingBlocks(b.x[0], b.x[1], b.x[2], b.x[3], b.x[4], b.x[5], intensity, stiffness, nobs, Wn);
            lighting = true;
            directional = dir;
            if (dir != has_unit_parameterization(b.x[5])) throw std::runtime_error("facade: light parameterization mismatch");
            // the bounds the driver sets on this observation's material and texture (:143-181)
            const auto lo = lower_.find(b.x[3]), hi = upper_.find(b.x[3]);
            if (lo != lower_.end()) {
                double l[3] = {-HUGE_VAL, -HUGE_VAL, -HUGE_VAL}, h[3] = {HUGE_VAL, HUGE_VAL, HUGE_VAL};
                for (const auto& kv : lo->second) l[kv.first] = kv.second;
                if (hi != upper_.end())
                    for (const auto& kv : hi->second) h[kv.first] = kv.second;
                q.SetMaterialBounds(l, h);
            }
            const auto tl = lower_.find(b.x[4]), th = upper_.find(b.x[4]);
            if (tl != lower_.end() && th != upper_.end()) q.SetTextureBounds(tl->second.at(0), th->second.at(0));
        }
        if (lighting) q.SetLightDirectional(directional);
        for (double* x : poses_)
            if (is_constant(x)) q.SetParameterBlockConstant(x);
        size_t n_const_pos = 0;
        for (double* x : positions) n_const_pos += is_constant(x) ? 1 : 0;
        if (n_const_pos != 0 && n_const_pos != positions.size()) throw std::runtime_error("facade: some positions constant, some not");
        if (n_const_pos) q.SetPointsConstant(true);
        return n_res != 0;
    }

    std::vector<double*> poses_;   // pose blocks in order of first appearance
    std::vector<double*> constant_;
    std::vector<Block> blocks_;

   private:
    ResidualBlockId add(CostFunction* cost, LossFunction* loss, std::vector<double*> x) {
        owned_.emplace_back(cost);
        if (loss) loss_owned_.emplace_back(loss);
        if (x.size() <= 2 || x.size() == 6) {
            const std::type_info& t = cost->functor_type();
            using namespace ceres_slam;
            if (t != typeid(StereoReprojectionErrorAutomatic) && t != typeid(SunSensorErrorAutomatic) && t != typeid(PoseErrorAutomatic) &&
                t != typeid(NormalErrorAutomatic) && t != typeid(IntensityErrorPointLightAutomatic) &&
                t != typeid(IntensityErrorDirectionalLightAutomatic))
                throw std::runtime_error(std::string("facade: cost functor not on the path: ") + t.name());
        }
        note_pose(x[0]);
        blocks_.push_back({cost, loss, x});
        return cost;
    }
    void note_pose(double* x) {
        if (std::find(poses_.begin(), poses_.end(), x) == poses_.end()) poses_.push_back(x);
    }
    std::vector<std::unique_ptr<CostFunction>> owned_;
    std::vector<std::unique_ptr<LossFunction>> loss_owned_;
    std::unique_ptr<LocalParameterization> se3_, unit_;
    std::vector<double*> unit_blocks_;
    std::map<double*, std::map<int, double>> lower_, upper_;

   public:
    std::unique_ptr<cslam_b200::Problem> solved_;   // the problem of the last Solve (ceres::Covariance asks it)
};

// One line per solve into $CSLAM_FACADE_TRACE (full precision): what was solved and what came back — the pose blocks,
// then every other parameter block of the problem in order of first appearance (positions, normals, materials, ...)
inline void facade_trace(const Problem& p, const cslam_summary& s) {
    const char* path = std::getenv("CSLAM_FACADE_TRACE");
    if (!path) return;
    std::ofstream f(path, std::ios::app);
    size_t n_stereo = 0;
    for (const auto& b : p.blocks_) n_stereo += Problem::is(b, typeid(ceres_slam::StereoReprojectionErrorAutomatic)) ? 1 : 0;
    f << std::setprecision(17) << "{\"n_poses\": " << p.poses_.size() << ", \"n_stereo\": " << n_stereo << ", \"n_blocks\": " << p.blocks_.size()
      << ", \"iterations\": " << s.num_iterations << ", \"initial_cost\": " << s.initial_cost << ", \"final_cost\": " << s.final_cost
      << ", \"termination\": " << s.termination_type << ", \"poses\": [";
    for (size_t i = 0; i < p.poses_.size(); ++i)
        for (int k = 0; k < 12; ++k) f << (i + k ? ", " : "") << p.poses_[i][k];
    f << "]";
    // other blocks by parameter slot of the 2- and 6-parameter residuals
    const char* names[6] = {nullptr, "positions", "normals", "materials", "textures", "light"};
    const int sizes[6] = {12, 3, 3, 3, 1, 3};
    for (int slot = 1; slot < 6; ++slot) {
        std::vector<const double*> seen;
        for (const auto& b : p.blocks_) {
            const bool six = b.x.size() == 6;
            const bool stereo = Problem::is(b, typeid(ceres_slam::StereoReprojectionErrorAutomatic));
            if (!(six || (stereo && slot == 1))) continue;
            const double* x = b.x[slot];
            if (std::find(seen.begin(), seen.end(), x) == seen.end()) seen.push_back(x);
        }
        if (seen.empty()) continue;
        f << ", \"" << names[slot] << "\": [";
        for (size_t i = 0; i < seen.size(); ++i)
            for (int k = 0; k < sizes[slot]; ++k) f << (i + k ? ", " : "") << seen[i][k];
        f << "]";
    }
    f << "}\n";
}

inline void Solve(const Solver::Options& options, Problem* problem, Solver::Summary* summary) {
    problem->solved_.reset(new cslam_b200::Problem());
    cslam_b200::Problem& q = *problem->solved_;
    cslam_options& o = q.options;                         // defaults = Ceres' (cslam_options_init)
    o.max_num_iterations = options.max_num_iterations;   // dataset_vo.cpp:69
    o.use_nonmonotonic_steps = options.use_nonmonotonic_steps ? 1 : 0;  // :70
    o.num_threads = options.num_threads;                 // :67 (the GPU back end ignores it)
    o.trust_region_strategy = int(options.trust_region_strategy_type);   // dataset_vo_sun.cpp:142
    o.dogleg_type = int(options.dogleg_type);                            // :143
    // linear_solver_type: every exact choice (SPARSE_NORMAL_CHOLESKY of dataset_ba_phong.cpp:87, Ceres' default
    // otherwise) is the back end's exact solve
    if (!problem->build(q)) {
        // Ceres solves an empty problem trivially; the C ABI wants at least a camera
        cslam_summary s{};
        if (summary) summary->inner.s = s;
        facade_trace(*problem, s);
        return;
    }
    cslam_b200::Summary inner;
    q.Solve(&inner);
    if (summary) summary->inner = inner;
    facade_trace(*problem, inner.s);
}

// ceres::Covariance as dataset_vo_sun.cpp:159-183 uses it: Compute({(pose, pose)}, &problem), then
// GetCovarianceBlockInTangentSpace(pose, pose, out 6x6 row-major).
//  * product build (and the ABI-trace build, where the oracle's covariance entry answers): cslam_covariance_block of the
//    solved problem.
//  * oracle build — deliberately NOT through the oracle's covariance entry, as an independent route: (J^T J)^-1 formed HERE from the problem's own cost
//    functions — the reference's functors differentiated by the AutoDiffCostFunction stand-in, chained with the
//    SE3Perturbation plus-Jacobian, loss-corrected like Ceres' Corrector does for rho'' <= 0 (scale by sqrt(rho')) — with a
//    dense Cholesky factorisation; what SPARSE_QR computes for a full-rank Jacobian.
class Covariance {
   public:
    struct Options {
        int num_threads = 1;
        SparseLinearAlgebraLibraryType sparse_linear_algebra_library_type = SUITE_SPARSE;
        CovarianceAlgorithmType algorithm_type = SPARSE_QR;
    };
    explicit Covariance(const Options&) {}
    bool Compute(const std::vector<std::pair<const double*, const double*>>& blocks, Problem* problem) {
        problem_ = problem;
        cov_.clear();
        for (const auto& b : blocks) {
            if (b.first != b.second) throw std::runtime_error("facade: only diagonal covariance blocks are on the path");
            std::vector<double> c(36);
            if (!block(const_cast<double*>(b.first), c.data())) return false;
            cov_[b.first] = c;
        }
        return true;
    }
    bool GetCovarianceBlockInTangentSpace(const double* x0, const double* x1, double* out) const {
        auto it = cov_.find(x0);
        if (x0 != x1 || it == cov_.end()) return false;
        for (int i = 0; i < 36; ++i) out[i] = it->second[i];
        if (const char* path = std::getenv("CSLAM_FACADE_TRACE")) {
            std::ofstream f(path, std::ios::app);
            f << std::setprecision(17) << "{\"covariance\": [";
            for (int i = 0; i < 36; ++i) f << (i ? ", " : "") << out[i];
            f << "]}\n";
        }
        return true;
    }

   private:
    bool block(double* pose, double* out36) {
#if !defined(CSLAM_FACADE_ORACLE) || defined(CSLAM_FACADE_ABI_TRACE)
        return problem_->solved_ && problem_->solved_->GetCovarianceBlockInTangentSpace(pose, out36);
#else
        Problem& p = *problem_;
        // variable blocks: non-constant poses (6 tangent columns each), then points (3 each) in order of appearance
        std::map<const double*, int> col;
        int n = 0;
        auto is_const = [&](const double* x) { return p.is_constant(x); };
        for (double* x : p.poses_)
            if (!is_const(x)) col[x] = n, n += 6;
        for (const auto& b : p.blocks_)
            if (b.x.size() == 2 && !col.count(b.x[1])) col[b.x[1]] = n, n += 3;
        if (!col.count(pose)) return false;
        std::vector<double> H(size_t(n) * n, 0.0);
        for (const auto& b : p.blocks_) {
            const int m = b.cost->num_residuals();
            std::vector<double> r(m), Ja(size_t(m) * 12), Jb(size_t(m) * 3), P(72);
            double* jac[2] = {Ja.data(), Jb.data()};
            const double* params[2] = {b.x[0], b.x.size() > 1 ? b.x[1] : nullptr};
            if (!b.cost->Evaluate(params, r.data(), jac)) return false;
            double scale = 1.0;
            if (b.loss) {
                double s = 0, rho[3];
                for (double v : r) s += v * v;
                b.loss->Evaluate(s, rho);
                scale = std::sqrt(rho[1]);   // Corrector with rho'' <= 0 (Huber): alpha = 0
            }
            // tangent-space Jacobian of the block, m x w, and where its columns live
            std::vector<std::pair<int, std::vector<double>>> parts;   // (first column, m x width row-major)
            if (col.count(b.x[0])) {
                p.lp()->ComputeJacobian(b.x[0], P.data());
                std::vector<double> Jt(size_t(m) * 6, 0.0);
                for (int i = 0; i < m; ++i)
                    for (int c = 0; c < 6; ++c) {
                        double a = 0;
                        for (int k = 0; k < 12; ++k) a += Ja[size_t(i) * 12 + k] * P[k * 6 + c];
                        Jt[size_t(i) * 6 + c] = scale * a;
                    }
                parts.push_back({col[b.x[0]], Jt});
            }
            if (b.x.size() > 1) {
                std::vector<double> Jt(Jb);
                for (double& v : Jt) v *= scale;
                parts.push_back({col[b.x[1]], Jt});
            }
            for (const auto& A : parts)
                for (const auto& B : parts) {
                    const int wa = int(A.second.size()) / m, wb = int(B.second.size()) / m;
                    for (int i = 0; i < wa; ++i)
                        for (int j = 0; j < wb; ++j) {
                            double a = 0;
                            for (int q = 0; q < m; ++q) a += A.second[size_t(q) * wa + i] * B.second[size_t(q) * wb + j];
                            H[size_t(A.first + i) * n + B.first + j] += a;
                        }
                }
        }
        // Cholesky H = L L^T, then the six columns of H^-1 that belong to the pose
        for (int j = 0; j < n; ++j) {
            double d = H[size_t(j) * n + j];
            for (int k = 0; k < j; ++k) d -= H[size_t(j) * n + k] * H[size_t(j) * n + k];
            if (!(d > 0.0)) return false;   // rank deficient: Ceres' Compute fails
            d = std::sqrt(d);
            H[size_t(j) * n + j] = d;
            for (int i = j + 1; i < n; ++i) {
                double a = H[size_t(i) * n + j];
                for (int k = 0; k < j; ++k) a -= H[size_t(i) * n + k] * H[size_t(j) * n + k];
                H[size_t(i) * n + j] = a / d;
            }
        }
        const int c0 = col[pose];
        for (int c = 0; c < 6; ++c) {
            std::vector<double> y(n, 0.0);
            y[c0 + c] = 1.0;
            for (int i = 0; i < n; ++i) {
                double a = y[i];
                for (int k = 0; k < i; ++k) a -= H[size_t(i) * n + k] * y[k];
                y[i] = a / H[size_t(i) * n + i];
            }
            for (int i = n - 1; i >= 0; --i) {
                double a = y[i];
                for (int k = i + 1; k < n; ++k) a -= H[size_t(k) * n + i] * y[k];
                y[i] = a / H[size_t(i) * n + i];
            }
            for (int r = 0; r < 6; ++r) out36[6 * r + c] = y[c0 + r];
        }
        return true;
#endif
    }
    Problem* problem_ = nullptr;
    std::map<const double*, std::vector<double>> cov_;
};

}  // namespace ceres
#endif  // CSLAM_REF_DRIVER_CERES_FACADE
