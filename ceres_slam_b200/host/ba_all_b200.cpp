// ba_all_b200 — scripts/ba_all_sims.sh:29-57 / ba_all_icra.sh:75-125 run dataset_vo_sun over many (trajectory x
// sun file) pairs, one process after the other.  The jobs are independent and window w of a job depends only
// on window w-1 of the same job, so this runner advances all jobs in lock-step and hands window w of every job
// to ONE cslam_solve_batch call (BASELINE.json config 4: many windows packed into one launch per GPU).  Per job
// it does exactly what dataset_vo_sun_b200 does (same files in, same `_poses.csv` files out).
//
//   usage: ba_all_b200 <jobs_file> [--window (2)] [--huber-param (0)] [--az-err-thresh (1000)]
//          [--zen-err-thresh (1000)] [--sun-only] [--max-iters (1000)] [--strategy lm|dogleg]
//   jobs_file: one job per line, `<track_file> <ref_sun_file> <obs_sun_file>`
// With --strategy lm (default here) the 2-pose windows take the one-CTA-per-window kernel, all jobs in one launch;
// DOGLEG windows are solved one after the other by the generic engine inside the same call.
#include <memory>
#include <sstream>

#include "sun_dataset.hpp"

struct Job {
    std::string track, ref_sun, obs_sun;
    SunDataset d;
};

static void run_pass_all(std::vector<Job>& jobs, unsigned window, bool use_sun, double huber, double az, double zen,
                         int max_iters) {
    unsigned max_states = 0;
    for (auto& j : jobs) max_states = std::max(max_states, j.d.num_states);
    for (unsigned k1 = 0; k1 + window <= max_states; ++k1) {
        const unsigned k2 = k1 + window;
        std::vector<std::unique_ptr<Problem>> problems;
        std::vector<Problem*> batch;
        std::vector<Job*> owner;
        // the RANSAC front end of every job's window in one cslam_ransac_align call per set of intrinsics
        std::vector<Job*> live;
        for (auto& j : jobs)
            if (k2 <= j.d.num_states) live.push_back(&j);
        std::vector<char> done(live.size(), 0);
        for (size_t a = 0; a < live.size(); ++a) {
            if (done[a]) continue;
            std::vector<InitialGuessJob> ig;
            std::vector<size_t> who;
            for (size_t b = a; b < live.size(); ++b)
                if (!done[b] && std::memcmp(live[b]->d.intr, live[a]->d.intr, sizeof(live[a]->d.intr)) == 0) {
                    SunDataset& d = live[b]->d;
                    ig.push_back(InitialGuessJob{&d.obs, d.num_states, k1, k2, &d.poses, &d.points, &d.initialized, {}});
                    who.push_back(b);
                    done[b] = 1;
                }
            compute_initial_guess_batch(ig, live[a]->d.intr, 4.0, true);
            for (size_t i = 0; i < ig.size(); ++i) {
                Job& j = *live[who[i]];
                SunDataset& d = j.d;
                if (ig[i].stats.ok) {
                    problems.emplace_back(new Problem);
                    buildWindow(d, k1, k2, use_sun, huber, az, zen, max_iters, *problems.back());
                    batch.push_back(problems.back().get());
                    owner.push_back(&j);
                } else {
                    std::cerr << "WARNING: Initial guess failed. Copying previous pose and covariance." << std::endl;
                    std::memcpy(&d.poses[12 * size_t(k2 - 1)], &d.poses[12 * size_t(k1)], 96);
                    std::memcpy(&d.pose_covars[36 * size_t(k2 - 1)], &d.pose_covars[36 * size_t(k1)], 288);
                }
            }
        }
        std::vector<Summary> sums;
        Problem::SolveBatch(batch, &sums);
        for (size_t i = 0; i < batch.size(); ++i) {
            std::cout << "[" << k1 << "," << k2 << ") " << owner[i]->track << ": " << sums[i].BriefReport() << std::endl;
            absorbWindow(owner[i]->d, k1, *batch[i]);
        }
        for (auto& j : jobs) std::fill(j.d.initialized.begin(), j.d.initialized.end(), 0);  // reset_points
    }
}

int main(int argc, char** argv) {
    const std::string usage(
        "usage: ba_all_b200 <jobs_file> [--window (2)] [--huber-param (0)] [--az-err-thresh (1000)] "
        "[--zen-err-thresh (1000)] [--sun-only] [--max-iters (1000)] [--strategy lm|dogleg]");
    if (argc < 2) {
        std::cerr << usage << std::endl;
        return EXIT_FAILURE;
    }
    unsigned window = 2;
    bool sun_only = false;
    double huber = 0., az = 1000., zen = 1000.;
    int max_iters = 1000;
    const double pi = 3.14159265358979323846;
    g_dogleg = false;
    for (int a = 2; a < argc; ++a) {
        const std::string flag(argv[a]);
        if (flag == "--window" && argc > a + 1) window = unsigned(std::stoi(argv[++a]));
        else if (flag == "--huber-param" && argc > a + 1) huber = std::stod(argv[++a]);
        else if (flag == "--az-err-thresh" && argc > a + 1) az = std::stod(argv[++a]) * pi / 180.;
        else if (flag == "--zen-err-thresh" && argc > a + 1) zen = std::stod(argv[++a]) * pi / 180.;
        else if (flag == "--sun-only") sun_only = true;
        else if (flag == "--max-iters" && argc > a + 1) max_iters = std::stoi(argv[++a]);
        else if (flag == "--strategy" && argc > a + 1) g_dogleg = std::string(argv[++a]) != "lm";
        else {
            std::cerr << usage << std::endl;
            return EXIT_FAILURE;
        }
    }
    std::vector<Job> jobs;
    {
        std::ifstream in(argv[1]);
        if (!in.is_open()) {
            std::cerr << "cannot open " << argv[1] << std::endl;
            return EXIT_FAILURE;
        }
        std::string line;
        while (std::getline(in, line)) {
            std::istringstream ls(line);
            Job j;
            if (!(ls >> j.track >> j.ref_sun >> j.obs_sun)) continue;
            jobs.push_back(std::move(j));
        }
    }
    for (auto& j : jobs)
        if (!read_csv(j.track, j.ref_sun, j.obs_sun, j.d)) {
            std::cerr << "cannot read job " << j.track << std::endl;
            return EXIT_FAILURE;
        }
    if (jobs.empty()) return EXIT_SUCCESS;
    if (window == 0) window = 2;
    if (!sun_only) {
        std::cerr << "Computing VO without sun measurements (" << jobs.size() << " jobs)" << std::endl;
        run_pass_all(jobs, window, false, 0., 1000., 1000., max_iters);
        for (auto& j : jobs) write_poses_csv(file_stem(j.track) + "_poses.csv", j.d.poses, j.d.num_states);
    }
    std::cerr << "Computing VO with sun measurements (" << jobs.size() << " jobs)" << std::endl;
    run_pass_all(jobs, window, true, huber, az, zen, max_iters);
    for (auto& j : jobs) {
        std::string os = file_stem(j.obs_sun);
        const size_t us = os.rfind('_');
        if (us != std::string::npos) os = os.substr(us + 1);
        write_poses_csv(file_stem(j.track) + "_" + os + "_poses.csv", j.d.poses, j.d.num_states);
    }
    return EXIT_SUCCESS;
}
