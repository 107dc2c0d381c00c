// dataset_vo restated over the B200 back end: sliding-window stereo VO / BA on the reference's
// plain track CSV (format: src/ceres_slam/dataset_problem.cpp:27-83; driver:
// tests/dataset_vo.cpp:87-138).  The initial guess is the reference's: per window
// `compute_initial_guess(k1, k2)` (RANSAC point-cloud alignment of consecutive poses on the GPU,
// dataset_problem.cpp:179-270), residual blocks only for the points it initialised
// (dataset_vo.cpp:44), `reset_points()` after every window (:130).  `--init constant` keeps the
// earlier front-end-free start (pose k starts at pose k-1, points triangulated from their first
// observation in the window).
//
//   usage: dataset_vo_b200 <input_file> [--window N=0] [--max-iters M=1000] [--init ransac|constant]
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <iostream>

#include "cslam_problem.hpp"
#include "dataset.hpp"

using namespace cslam_b200;

struct Track {
    unsigned num_states = 0, num_points = 0;
    double intr[5], var[3];
    std::vector<double> poses;   // 12 per state, [t | R row-major]
    std::vector<double> points;  // 3 per point
    std::vector<char> initialized;
    ObservationTable obs;
};

static bool read_csv(const std::string& file, Track& t) {
    std::ifstream in(file);
    if (!in.is_open()) return false;
    std::string line;
    std::getline(in, line);
    auto v = parse_csv_line(line);
    t.num_states = unsigned(v.at(0));
    t.num_points = unsigned(v.at(1));
    std::getline(in, line);
    v = parse_csv_line(line);
    for (int i = 0; i < 5; ++i) t.intr[i] = v.at(i);
    std::getline(in, line);
    v = parse_csv_line(line);
    for (int i = 0; i < 3; ++i) t.var[i] = v.at(i);
    std::getline(in, line);
    v = parse_csv_line(line);  // first pose, 4x4 row-major; every pose starts there
    t.poses.assign(12 * size_t(t.num_states), 0.0);
    for (unsigned s = 0; s < t.num_states; ++s) pose_from_matrix16(v, &t.poses[12 * size_t(s)]);
    t.points.assign(3 * size_t(t.num_points), 0.0);
    t.initialized.assign(t.num_points, 0);
    t.obs.state_obs.assign(t.num_states, {});
    while (std::getline(in, line)) {
        if (line.empty()) continue;
        v = parse_csv_line(line);
        t.obs.state_obs.at(unsigned(v.at(0))).push_back(unsigned(t.obs.k.size()));
        t.obs.k.push_back(unsigned(v.at(0)));
        t.obs.j.push_back(unsigned(v.at(1)));
        t.obs.uvd.insert(t.obs.uvd.end(), {v.at(2), v.at(3), v.at(4)});
    }
    return true;
}

static void constant_pose_guess(Track& t, unsigned k1, unsigned k2) {
    std::vector<unsigned> seen(t.num_points, 0);
    for (unsigned k = k1; k < k2; ++k)
        for (unsigned i : t.obs.state_obs[k]) seen[t.obs.j[i]]++;
    for (unsigned k = k1; k < k2; ++k) {
        if (k > k1) std::memcpy(&t.poses[12 * size_t(k)], &t.poses[12 * size_t(k - 1)], 96);
        const double* P = &t.poses[12 * size_t(k)];
        for (unsigned i : t.obs.state_obs[k]) {
            const unsigned j = t.obs.j[i];
            if ((seen[j] < 2 && k2 - k1 > 1) || t.initialized[j]) continue;  // only points shared inside the window
            double pc[3];
            triangulate(t.intr, &t.obs.uvd[3 * size_t(i)], pc);
            pose_inverse_apply(P, pc, false, &t.points[3 * size_t(j)]);
            t.initialized[j] = 1;
        }
    }
}

static void solveWindow(Track& t, unsigned k1, unsigned k2, int max_iters) {
    std::cerr << "Working on interval [" << k1 << "," << k2 << ")" << std::endl;
    Problem problem;
    problem.SetCamera(t.intr[0], t.intr[1], t.intr[2], t.intr[3], t.intr[4]);
    // stiffness = diag(var)^-1/2  (dataset_vo.cpp:29-32)
    const double cov[9] = {t.var[0], 0, 0, 0, t.var[1], 0, 0, 0, t.var[2]};
    double W[9];
    sym_inverse_sqrt(cov, 3, W);
    for (unsigned k = k1; k < k2; ++k) {
        double* P = &t.poses[12 * size_t(k)];
        problem.AddPoseBlock(P);                                                  // :58
        for (unsigned i : t.obs.state_obs[k]) {
            const unsigned j = t.obs.j[i];
            if (t.initialized[j])                                                // :44
                problem.AddStereoBlock(P, &t.points[3 * size_t(j)], &t.obs.uvd[3 * size_t(i)], W);  // :46-53
        }
    }
    problem.SetParameterBlockConstant(&t.poses[12 * size_t(k1)]);                // :62
    problem.options.max_num_iterations = max_iters;  // dataset_vo.cpp:69 uses 1000
    problem.options.use_nonmonotonic_steps = 1;      // dataset_vo.cpp:70
    Summary summary;
    problem.Solve(&summary);
    std::cout << summary.BriefReport() << std::endl << std::endl;
}

// DatasetProblem::write_csv (dataset_problem.cpp:125-166): <stem>_poses.csv and <stem>_map.csv, the latter one row per
// point that is initialised at the time of the call (none after the last window's reset_points: a header-only file,
// as the reference leaves it)
static void write_outputs(const Track& t, const std::string& filename) {
    const std::string stem = file_stem(filename);
    write_poses_csv(stem + "_poses.csv", t.poses, t.num_states);
    std::ofstream map_file(stem + "_map.csv");
    map_file.precision(17);
    map_file << "point_id, x, y, z\n";
    for (unsigned j = 0; j < t.num_points; ++j)
        if (t.initialized[j])
            map_file << j << "," << t.points[3 * size_t(j)] << "," << t.points[3 * size_t(j) + 1] << ","
                     << t.points[3 * size_t(j) + 2] << "\n";
}

int main(int argc, char** argv) {
    const std::string usage("usage: dataset_vo_b200 <input_file> [--window N=0] [--max-iters M=1000] [--init ransac|constant]");
    if (argc < 2) {
        std::cerr << usage << std::endl;
        return EXIT_FAILURE;
    }
    unsigned window = 0;
    int max_iters = 1000;
    bool ransac = true;
    const std::string filename(argv[1]);
    for (int a = 2; a < argc; ++a) {
        const std::string flag(argv[a]);
        if (flag == "--window" && argc > a + 1) window = unsigned(std::atoi(argv[++a]));
        else if (flag == "--max-iters" && argc > a + 1) max_iters = std::atoi(argv[++a]);
        else if (flag == "--init" && argc > a + 1) ransac = std::string(argv[++a]) != "constant";
        else {
            std::cerr << usage << std::endl;
            return EXIT_FAILURE;
        }
    }
    Track t;
    if (!read_csv(filename, t)) return EXIT_FAILURE;
    if (window == 0 || window > t.num_states) window = t.num_states;  // 0 = full batch (dataset_vo.cpp:118-121)
    // Warm-up, untimed and on a copy: the first window once, so that CUDA context creation and kernel module loading
    // (0.3 - 2 s, once per process) are not billed to the window loop below
    const auto t_warm = std::chrono::steady_clock::now();
    if (t.num_states >= window) {
        Track w = t;
        if (ransac)
            compute_initial_guess(w.obs, w.intr, w.num_states, 0, window, 4.0, false, w.poses, w.points, w.initialized,
                                  [](unsigned, unsigned, const double*, unsigned) {});
        else
            constant_pose_guess(w, 0, window);
        std::cout.setstate(std::ios_base::failbit);
        solveWindow(w, 0, window, max_iters);
        std::cout.clear();
    }
    const double warm_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_warm).count();
    std::cerr << "Computing VO" << std::endl;
    const auto t_loop = std::chrono::steady_clock::now();
    unsigned n_windows = 0;
    double guess_s = 0.0, solve_s = 0.0;
    for (unsigned k1 = 0; k1 + window <= t.num_states; ++k1, ++n_windows) {
        const unsigned k2 = k1 + window;
        const auto t_w0 = std::chrono::steady_clock::now();
        if (ransac)
            compute_initial_guess(t.obs, t.intr, t.num_states, k1, k2, 4.0, false, t.poses, t.points, t.initialized,
                                  [](unsigned, unsigned, const double*, unsigned) {});    // :127
        else
            constant_pose_guess(t, k1, k2);
        const auto t_w1 = std::chrono::steady_clock::now();
        solveWindow(t, k1, k2, max_iters);                                                  // :128
        guess_s += std::chrono::duration<double>(t_w1 - t_w0).count();
        solve_s += std::chrono::duration<double>(std::chrono::steady_clock::now() - t_w1).count();
        std::fill(t.initialized.begin(), t.initialized.end(), 0);                           // reset_points, :130
    }
    // front end + solve of every window, without the CSV input / output (bench.py's C1 line reads this)
    std::cerr << "cslam_b200 timing: windows=" << n_windows << " loop_s="
              << std::chrono::duration<double>(std::chrono::steady_clock::now() - t_loop).count()
              << " initial_guess_s=" << guess_s << " solve_s=" << solve_s << " warmup_s=" << warm_s << std::endl;
    write_outputs(t, filename);                                                             // :134
    return EXIT_SUCCESS;
}
