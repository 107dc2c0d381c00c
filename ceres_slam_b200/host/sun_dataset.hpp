// Shared by dataset_vo_sun_b200 and ba_all_b200: the sun dataset (dataset_problem_sun.cpp:33-175) and the
// window problem of tests/dataset_vo_sun.cpp:25-187, split into build / solve / absorb so that the windows of
// several independent jobs can go through one cslam_solve_batch call.
#pragma once
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>

#include "cslam_problem.hpp"
#include "dataset.hpp"

using namespace cslam_b200;

struct SunDataset {
    unsigned num_states = 0, num_points = 0;
    double intr[5];
    std::vector<double> poses, points, pose_covars;  // 12 / 3 / 36 per entry
    std::vector<char> initialized;
    ObservationTable obs;
    std::vector<double> obs_covars;                  // 9 per observation
    std::vector<double> sun_dir_g, sun_obs, sun_covars;  // 3 / 3 / 4 per state
    std::vector<char> has_sun;
};

static bool read_csv(const std::string& track, const std::string& ref_sun, const std::string& obs_sun, SunDataset& d) {
    std::ifstream in(track);
    if (!in.is_open()) return false;
    std::string line;
    std::getline(in, line);
    auto v = parse_csv_line(line);
    d.num_states = unsigned(v.at(0));
    d.num_points = unsigned(v.at(1));
    std::getline(in, line);
    v = parse_csv_line(line);
    for (int i = 0; i < 5; ++i) d.intr[i] = v.at(i);
    std::getline(in, line);
    v = parse_csv_line(line);  // first pose (no variance line in this format)
    d.poses.assign(12 * size_t(d.num_states), 0.0);
    for (unsigned s = 0; s < d.num_states; ++s) pose_from_matrix16(v, &d.poses[12 * size_t(s)]);
    d.pose_covars.assign(36 * size_t(d.num_states), 0.0);
    for (int i = 0; i < 6; ++i) d.pose_covars[7 * i] = 1e-12;  // dataset_problem_sun.cpp:80
    d.points.assign(3 * size_t(d.num_points), 0.0);
    d.initialized.assign(d.num_points, 0);
    d.obs.state_obs.assign(d.num_states, {});
    while (std::getline(in, line)) {
        if (line.empty()) continue;
        v = parse_csv_line(line);
        d.obs.state_obs.at(unsigned(v.at(0))).push_back(unsigned(d.obs.k.size()));
        d.obs.k.push_back(unsigned(v.at(0)));
        d.obs.j.push_back(unsigned(v.at(1)));
        d.obs.uvd.insert(d.obs.uvd.end(), {v.at(2), v.at(3), v.at(4)});
        for (int c = 0; c < 9; ++c) d.obs_covars.push_back(v.at(5 + c));
    }
    d.sun_dir_g.assign(3 * size_t(d.num_states), 0.0);
    d.sun_obs.assign(3 * size_t(d.num_states), 0.0);
    d.sun_covars.assign(4 * size_t(d.num_states), 0.0);
    d.has_sun.assign(d.num_states, 0);
    std::ifstream in2(ref_sun);
    if (!in2.is_open()) return false;
    while (std::getline(in2, line)) {
        if (line.empty()) continue;
        v = parse_csv_line(line);
        const unsigned k = unsigned(v.at(0));
        for (int c = 0; c < 3; ++c) d.sun_dir_g.at(3 * size_t(k) + c) = v.at(1 + c);
    }
    std::ifstream in3(obs_sun);
    if (!in3.is_open()) return false;
    while (std::getline(in3, line)) {
        if (line.empty()) continue;
        v = parse_csv_line(line);
        const unsigned k = unsigned(v.at(0));
        for (int c = 0; c < 3; ++c) d.sun_obs.at(3 * size_t(k) + c) = v.at(1 + c);
        for (int c = 0; c < 4; ++c) d.sun_covars.at(4 * size_t(k) + c) = v.at(4 + c);
        d.has_sun[k] = 1;
    }
    return true;
}

static bool g_dogleg = true;

// The problem of window [k1, k2) (dataset_vo_sun.cpp:49-143), ready to solve.
static void buildWindow(SunDataset& d, unsigned k1, unsigned k2, bool use_sun, double huber, double az, double zen,
                        int max_iters, Problem& problem) {
    problem.SetCamera(d.intr[0], d.intr[1], d.intr[2], d.intr[3], d.intr[4]);
    for (unsigned k = k1; k < k2; ++k) {
        double* P = &d.poses[12 * size_t(k)];
        problem.AddPoseBlock(P);
        for (unsigned i : d.obs.state_obs[k]) {
            const unsigned j = d.obs.j[i];
            if (!d.initialized[j]) continue;                                             // :54
            // the reference indexes the per-observation covariances by POINT id (:58);
            // kept, guarded against the out-of-range read it would make on short tables
            const size_t ci = size_t(j) < d.obs_covars.size() / 9 ? size_t(j) : size_t(i);
            double W[9];
            sym_inverse_sqrt(&d.obs_covars[9 * ci], 3, W);                               // :57-59
            problem.AddStereoBlock(P, &d.points[3 * size_t(j)], &d.obs.uvd[3 * size_t(i)], W);  // :62-70
        }
        if (use_sun && d.has_sun[k]) {                                                   // :75
            double W2[4];
            sym_inverse_sqrt(&d.sun_covars[4 * size_t(k)], 2, W2);                       // :78-80
            problem.AddSunBlock(P, &d.sun_obs[3 * size_t(k)], &d.sun_dir_g[3 * size_t(k)], W2, az, zen, huber);  // :83-99
        }
    }
    // prior on the first pose of the window from the previous window's covariance (:109-124)
    double W6[36];
    sym_inverse_sqrt(&d.pose_covars[36 * size_t(k1)], 6, W6);
    double Tref[12];
    std::memcpy(Tref, &d.poses[12 * size_t(k1)], 96);
    problem.AddPosePrior(&d.poses[12 * size_t(k1)], Tref, W6);
    problem.options.max_num_iterations = max_iters;  // :140
    problem.options.use_nonmonotonic_steps = 1;      // :141
    problem.options.trust_region_strategy = g_dogleg ? 1 : 0;  // :142 ceres::DOGLEG
    problem.options.dogleg_type = 1;                 // :143 ceres::SUBSPACE_DOGLEG
}

// After the solve: covariance of the second pose of the window -> prior of the next window (:159-183)
static void absorbWindow(SunDataset& d, unsigned k1, Problem& problem) {
    if (k1 + 1 < d.num_states) {
        if (!problem.GetCovarianceBlockInTangentSpace(&d.poses[12 * size_t(k1 + 1)], &d.pose_covars[36 * size_t(k1 + 1)])) {
            std::cout << "WARNING: Covariance computation failed! Using previous state covariance." << std::endl;
            std::memcpy(&d.pose_covars[36 * size_t(k1 + 1)], &d.pose_covars[36 * size_t(k1)], 288);
        }
    }
}

[[maybe_unused]] static void solveWindow(SunDataset& d, unsigned k1, unsigned k2, bool use_sun, double huber, double az, double zen,
                        int max_iters) {
    std::cerr << "Working on interval [" << k1 << "," << k2 << ")/" << d.num_states << ": ";
    Problem problem;
    buildWindow(d, k1, k2, use_sun, huber, az, zen, max_iters, problem);
    Summary summary;
    problem.Solve(&summary);
    std::cout << summary.BriefReport() << std::endl;
    absorbWindow(d, k1, problem);
}

