// dataset_vo restated over the B200 back end: sliding-window stereo VO / BA on the reference's
// plain track CSV (format: src/ceres_slam/dataset_problem.cpp:27-83; driver:
// tests/dataset_vo.cpp:87-138).  The reference's front end (RANSAC point-cloud alignment,
// dataset_problem.cpp:179-270) is out of scope for this build, so the initial guess is the
// constant-pose model: pose k starts at the optimised pose k-1 and every point is triangulated
// from its first observation in the window (stereo_camera.hpp:112-120).
//
//   usage: dataset_vo_b200 <input_file> [--window N=0] [--max-iters M]
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

#include "cslam_problem.hpp"

struct Track {
    unsigned num_states = 0, num_points = 0;
    double fu, fv, cu, cv, b, var[3];
    std::vector<double> poses;   // 12 per state, [t | R row-major]
    std::vector<double> points;  // 3 per point
    std::vector<unsigned> k, j;
    std::vector<double> uvd;
    std::vector<std::vector<unsigned>> state_obs;
};

static std::vector<double> parse_line(const std::string& line) {
    std::vector<double> v;
    std::stringstream ss(line);
    std::string tok;
    while (std::getline(ss, tok, ',')) v.push_back(std::stod(tok));
    return v;
}

static bool read_csv(const std::string& file, Track& t) {
    std::ifstream in(file);
    if (!in.is_open()) return false;
    std::string line;
    std::getline(in, line);
    auto v = parse_line(line);
    t.num_states = unsigned(v.at(0));
    t.num_points = unsigned(v.at(1));
    std::getline(in, line);
    v = parse_line(line);
    t.fu = v.at(0); t.fv = v.at(1); t.cu = v.at(2); t.cv = v.at(3); t.b = v.at(4);
    std::getline(in, line);
    v = parse_line(line);
    for (int i = 0; i < 3; ++i) t.var[i] = v.at(i);
    std::getline(in, line);
    v = parse_line(line);  // first pose, 4x4 row-major
    t.poses.assign(12 * size_t(t.num_states), 0.0);
    for (unsigned s = 0; s < t.num_states; ++s) {
        double* P = &t.poses[12 * size_t(s)];
        for (int r = 0; r < 3; ++r) {
            P[r] = v.at(4 * r + 3);
            for (int c = 0; c < 3; ++c) P[3 + 3 * r + c] = v.at(4 * r + c);
        }
    }
    t.points.assign(3 * size_t(t.num_points), 0.0);
    t.state_obs.assign(t.num_states, {});
    while (std::getline(in, line)) {
        if (line.empty()) continue;
        v = parse_line(line);
        t.state_obs.at(unsigned(v.at(0))).push_back(unsigned(t.k.size()));
        t.k.push_back(unsigned(v.at(0)));
        t.j.push_back(unsigned(v.at(1)));
        t.uvd.insert(t.uvd.end(), {v.at(2), v.at(3), v.at(4)});
    }
    return true;
}

static void solveWindow(Track& t, unsigned k1, unsigned k2, int max_iters) {
    std::cerr << "Working on interval [" << k1 << "," << k2 << ")" << std::endl;
    cslam_b200::Problem problem;
    problem.SetCamera(t.fu, t.fv, t.cu, t.cv, t.b);
    // stiffness = diag(var)^-1/2  (dataset_vo.cpp:29-32)
    double W[9] = {1 / std::sqrt(t.var[0]), 0, 0, 0, 1 / std::sqrt(t.var[1]), 0, 0, 0, 1 / std::sqrt(t.var[2])};
    std::vector<unsigned> seen(t.num_points, 0);
    for (unsigned k = k1; k < k2; ++k)
        for (unsigned i : t.state_obs[k]) seen[t.j[i]]++;
    std::vector<char> init(t.num_points, 0);
    for (unsigned k = k1; k < k2; ++k) {
        if (k > k1) std::memcpy(&t.poses[12 * size_t(k)], &t.poses[12 * size_t(k - 1)], 96);  // constant-pose guess
        double* P = &t.poses[12 * size_t(k)];
        problem.AddPoseBlock(P);
        for (unsigned i : t.state_obs[k]) {
            const unsigned j = t.j[i];
            if (seen[j] < 2 && k2 - k1 > 1) continue;  // only points shared inside the window
            double* X = &t.points[3 * size_t(j)];
            if (!init[j]) {
                // triangulate in camera k, move to the base frame with T^-1 = (R^T, -R^T t)
                const double* z = &t.uvd[3 * size_t(i)];
                const double bod = t.b / z[2];
                const double pc[3] = {(z[0] - t.cu) * bod, (z[1] - t.cv) * bod * t.fu / t.fv, t.fu * bod};
                for (int c = 0; c < 3; ++c)
                    X[c] = P[3 + c] * (pc[0] - P[0]) + P[6 + c] * (pc[1] - P[1]) + P[9 + c] * (pc[2] - P[2]);
                init[j] = 1;
            }
            problem.AddStereoBlock(P, X, &t.uvd[3 * size_t(i)], W);
        }
    }
    problem.SetParameterBlockConstant(&t.poses[12 * size_t(k1)]);
    problem.options.max_num_iterations = max_iters;  // dataset_vo.cpp:69 uses 1000
    problem.options.use_nonmonotonic_steps = 1;      // dataset_vo.cpp:70
    cslam_b200::Summary summary;
    problem.Solve(&summary);
    std::cout << summary.BriefReport() << std::endl << std::endl;
}

int main(int argc, char** argv) {
    const std::string usage("usage: dataset_vo_b200 <input_file> [--window N=0] [--max-iters M=1000]");
    if (argc < 2) {
        std::cerr << usage << std::endl;
        return EXIT_FAILURE;
    }
    unsigned window = 0;
    int max_iters = 1000;
    const std::string filename(argv[1]);
    for (int a = 2; a < argc; ++a) {
        const std::string flag(argv[a]);
        if (flag == "--window" && argc > a + 1) window = unsigned(std::atoi(argv[++a]));
        else if (flag == "--max-iters" && argc > a + 1) max_iters = std::atoi(argv[++a]);
        else {
            std::cerr << usage << std::endl;
            return EXIT_FAILURE;
        }
    }
    Track t;
    if (!read_csv(filename, t)) return EXIT_FAILURE;
    if (window == 0 || window > t.num_states) window = t.num_states;  // 0 = full batch (dataset_vo.cpp:118-121)
    for (unsigned k1 = 0; k1 + window <= t.num_states; ++k1) solveWindow(t, k1, k1 + window, max_iters);
    // <stem>_poses.csv with 16 values per row (dataset_problem.cpp:139-150); full precision here
    const std::string stem = filename.substr(0, filename.find('.'));
    std::ofstream out(stem + "_poses.csv");
    out << "T_00, T_01, T_02, T_03,T_10, T_11, T_12, T_13,T_20, T_21, T_22, T_23,T_30, T_31, T_32, T_33\n";
    out.precision(17);
    for (unsigned s = 0; s < t.num_states; ++s) {
        const double* P = &t.poses[12 * size_t(s)];
        for (int r = 0; r < 3; ++r) out << P[3 + 3 * r] << "," << P[4 + 3 * r] << "," << P[5 + 3 * r] << "," << P[r] << ",";
        out << "0,0,0,1\n";
    }
    return EXIT_SUCCESS;
}
