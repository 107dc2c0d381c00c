// dataset_vo_sun restated over the B200 back end (tests/dataset_vo_sun.cpp:189-323): sliding-window
// stereo VO with sun-sensor blocks, a pose prior on the first pose of each window whose covariance
// is the marginal covariance of that pose from the previous window, per-observation stereo
// covariances.  Files: sun track CSV (dataset_problem_sun.cpp:33-103), ephemeris `k,e,n,u`
// (:139-146), observed sun `k,x,y,z,c00,c01,c10,c11` (:162-175).
// Trust-region strategy: SUBSPACE_DOGLEG as the reference sets it (:142-143); `--strategy lm` runs
// Levenberg-Marquardt instead (then the 2-pose windows take the one-CTA-per-window kernel).
//
//   usage: dataset_vo_sun_b200 <track_file> <ref_sun_file> <obs_sun_file> [--window (2)]
//          [--huber-param (0)] [--az-err-thresh (1000)] [--zen-err-thresh (1000)] [--sun-only]
//          [--max-iters (1000)] [--strategy dogleg|lm]
#include <cmath>
#include <cstdio>
#include <cstdlib>
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

static void solveWindow(SunDataset& d, unsigned k1, unsigned k2, bool use_sun, double huber, double az, double zen,
                        int max_iters) {
    std::cerr << "Working on interval [" << k1 << "," << k2 << ")/" << d.num_states << ": ";
    Problem problem;
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
    Summary summary;
    problem.Solve(&summary);
    std::cout << summary.BriefReport() << std::endl;
    // covariance of the second pose of the window -> prior of the next window (:159-183)
    if (k1 + 1 < d.num_states) {
        if (!problem.GetCovarianceBlockInTangentSpace(&d.poses[12 * size_t(k1 + 1)], &d.pose_covars[36 * size_t(k1 + 1)])) {
            std::cout << "WARNING: Covariance computation failed! Using previous state covariance." << std::endl;
            std::memcpy(&d.pose_covars[36 * size_t(k1 + 1)], &d.pose_covars[36 * size_t(k1)], 288);
        }
    }
}

static void run_pass(SunDataset& d, unsigned window, bool use_sun, double huber, double az, double zen, int max_iters) {
    for (unsigned k1 = 0; k1 + window <= d.num_states; ++k1) {
        const unsigned k2 = k1 + window;
        const InitialGuessStats st = compute_initial_guess(d.obs, d.intr, d.num_states, k1, k2, 4.0, true, d.poses, d.points,
                                                           d.initialized, [](unsigned, unsigned, const double*, unsigned) {});
        if (st.ok) {
            solveWindow(d, k1, k2, use_sun, huber, az, zen, max_iters);
        } else {
            std::cerr << "WARNING: Initial guess failed. Copying previous pose and covariance." << std::endl;  // :283-288
            std::memcpy(&d.poses[12 * size_t(k2 - 1)], &d.poses[12 * size_t(k1)], 96);
            std::memcpy(&d.pose_covars[36 * size_t(k2 - 1)], &d.pose_covars[36 * size_t(k1)], 288);
        }
        std::fill(d.initialized.begin(), d.initialized.end(), 0);  // reset_points
    }
}

int main(int argc, char** argv) {
    const std::string usage(
        "usage: dataset_vo_sun_b200 <track_file> <ref_sun_file> <obs_sun_file> [--window (2)] [--huber-param (0)] "
        "[--az-err-thresh (1000)] [--zen-err-thresh (1000)] [--sun-only] [--max-iters (1000)] [--strategy dogleg|lm]");
    if (argc < 4) {
        std::cerr << usage << std::endl;
        return EXIT_FAILURE;
    }
    unsigned window = 2;
    bool sun_only = false;
    double huber = 0., az = 1000., zen = 1000.;
    int max_iters = 1000;
    const double pi = 3.14159265358979323846;
    const std::string track(argv[1]), ref_sun(argv[2]), obs_sun(argv[3]);
    for (int a = 4; a < argc; ++a) {
        const std::string flag(argv[a]);
        if (flag == "--window" && argc > a + 1) window = unsigned(std::stoi(argv[++a]));
        else if (flag == "--huber-param" && argc > a + 1) huber = std::stod(argv[++a]);
        else if (flag == "--az-err-thresh" && argc > a + 1) az = std::stod(argv[++a]) * pi / 180.;  // degrees in
        else if (flag == "--zen-err-thresh" && argc > a + 1) zen = std::stod(argv[++a]) * pi / 180.;
        else if (flag == "--sun-only") sun_only = true;
        else if (flag == "--max-iters" && argc > a + 1) max_iters = std::stoi(argv[++a]);
        else if (flag == "--strategy" && argc > a + 1) g_dogleg = std::string(argv[++a]) != "lm";
        else {
            std::cerr << usage << std::endl;
            return EXIT_FAILURE;
        }
    }
    SunDataset d;
    if (!read_csv(track, ref_sun, obs_sun, d)) return EXIT_FAILURE;
    if (window == 0 || window > d.num_states) window = d.num_states;
    if (!sun_only) {
        std::cerr << "Computing VO without sun measurements" << std::endl;        // :271-296
        run_pass(d, window, false, 0., 1000., 1000., max_iters);
        write_poses_csv(file_stem(track) + "_poses.csv", d.poses, d.num_states);
    }
    std::cerr << "Computing VO with sun measurements" << std::endl;                // :298-311
    run_pass(d, window, true, huber, az, zen, max_iters);
    // <track stem>_<last '_' token of the observed-sun stem>_poses.csv (:313-320)
    std::string os = file_stem(obs_sun);
    const size_t us = os.rfind('_');
    if (us != std::string::npos) os = os.substr(us + 1);
    write_poses_csv(file_stem(track) + "_" + os + "_poses.csv", d.poses, d.num_states);
    return EXIT_SUCCESS;
}
