// dataset_ba_phong restated over the B200 back end (tests/dataset_ba_phong.cpp:257-349): stereo BA
// with Phong lighting — vertex positions and normals, one material [ka, ks, alpha] and one texture
// kd per material id shared by its vertices, a point light or (--dirlight) a directional light.
// File: Phong CSV (dataset_problem_phong.cpp:29-117): `num_states,num_vertices,num_materials`;
// intrinsics; `var_u,var_v,var_d,var_nx,var_ny,var_nz,var_I`; light position or direction; first
// pose; rows `t,j,mat_id,u,v,d,I,nx,ny,nz` grouped by timestamp.
// Trust-region strategy: SUBSPACE_DOGLEG with an exact (SPARSE_NORMAL_CHOLESKY-equivalent) linear solve, as the
// reference sets it (:85-87); `--strategy lm` runs Levenberg-Marquardt instead.
// --multistage runs the reference's three solves per window: stage 1 poses and points without
// lighting (:94-98), stage 2 lighting only with every pose and position constant (:207-231), stage 3
// everything jointly (:249-252).
//
//   usage: dataset_ba_phong_b200 <input_file> [--nolight | --dirlight] [--window N] [--multistage]
//          [--max-iters M] [--material-by-observation] [--strategy dogleg|lm]
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <iostream>

#include "cslam_problem.hpp"
#include "dataset.hpp"

using namespace cslam_b200;

static bool g_dogleg = true;

struct PhongDataset {
    unsigned num_states = 0, num_vertices = 0, num_materials = 0;
    bool directional = false;
    double intr[5], stereo_var[3], normal_var[3], int_var = 1;
    double light[3];
    std::vector<double> poses, positions, normals;   // 12 / 3 / 3
    std::vector<unsigned> vertex_material;           // material block a vertex was given
    std::vector<double> materials, textures;         // 3 / 1 per material id
    std::vector<char> initialized;
    bool materials_started = false;
    ObservationTable obs;                            // obs.k holds the index of the timestamp run
    std::vector<unsigned> material_ids;              // per observation
    std::vector<double> intensity, normal_obs;       // 1 / 3 per observation
};

static bool read_csv(const std::string& file, PhongDataset& d) {
    std::ifstream in(file);
    if (!in.is_open()) return false;
    std::string line;
    std::getline(in, line);
    auto v = parse_csv_line(line);
    d.num_states = unsigned(v.at(0));
    d.num_vertices = unsigned(v.at(1));
    d.num_materials = unsigned(v.at(2));
    std::getline(in, line);
    v = parse_csv_line(line);
    for (int i = 0; i < 5; ++i) d.intr[i] = v.at(i);
    std::getline(in, line);
    v = parse_csv_line(line);
    for (int i = 0; i < 3; ++i) d.stereo_var[i] = v.at(i), d.normal_var[i] = v.at(3 + i);
    d.int_var = v.at(6);
    std::getline(in, line);
    v = parse_csv_line(line);
    for (int i = 0; i < 3; ++i) d.light[i] = v.at(i);
    if (d.directional) {  // light_dir.normalize() (:76)
        const double n = std::sqrt(d.light[0] * d.light[0] + d.light[1] * d.light[1] + d.light[2] * d.light[2]);
        for (double& x : d.light) x /= n;
    }
    std::getline(in, line);
    v = parse_csv_line(line);
    d.poses.assign(12 * size_t(d.num_states), 0.0);
    for (unsigned s = 0; s < d.num_states; ++s) pose_from_matrix16(v, &d.poses[12 * size_t(s)]);
    d.positions.assign(3 * size_t(d.num_vertices), 0.0);
    d.normals.assign(3 * size_t(d.num_vertices), 0.0);
    d.vertex_material.assign(d.num_vertices, 0);
    d.initialized.assign(d.num_vertices, 0);
    d.materials.assign(3 * size_t(d.num_materials), 0.0);
    d.textures.assign(d.num_materials, 0.0);
    double t_prev = 0;
    while (std::getline(in, line)) {
        if (line.empty()) continue;
        v = parse_csv_line(line);
        // states are RUNS of equal timestamps (:121-131)
        if (d.obs.k.empty() || v.at(0) != t_prev) d.obs.state_obs.emplace_back();
        t_prev = v.at(0);
        d.obs.state_obs.back().push_back(unsigned(d.obs.k.size()));
        d.obs.k.push_back(unsigned(d.obs.state_obs.size() - 1));
        d.obs.j.push_back(unsigned(v.at(1)));
        d.material_ids.push_back(unsigned(v.at(2)));
        d.obs.uvd.insert(d.obs.uvd.end(), {v.at(3), v.at(4), v.at(5)});
        d.intensity.push_back(v.at(6));
        d.normal_obs.insert(d.normal_obs.end(), {v.at(7), v.at(8), v.at(9)});
    }
    d.obs.state_obs.resize(d.num_states);
    return true;
}

// DatasetProblemPhong::compute_initial_guess (dataset_problem_phong.cpp:248-390)
static void initial_guess(PhongDataset& d, unsigned k1, unsigned k2, bool material_by_observation) {
    if (!d.materials_started) {
        // materials (0, 0, 1), textures the median observed intensity of the material (:262-279).  The
        // reference re-creates these objects on EVERY call while the vertices keep the shared_ptr they
        // were given at their own initialisation; with vertices that are never reset only the objects
        // of the first call are ever used, which is what a single initialisation reproduces.
        for (unsigned m = 0; m < d.num_materials; ++m) {
            d.materials[3 * size_t(m)] = 0., d.materials[3 * size_t(m) + 1] = 0., d.materials[3 * size_t(m) + 2] = 1.;
            std::vector<double> ints;
            for (size_t i = 0; i < d.material_ids.size(); ++i)
                if (d.material_ids[i] == m) ints.push_back(d.intensity[i]);
            if (ints.empty()) continue;
            std::nth_element(ints.begin(), ints.begin() + ints.size() / 2, ints.end());
            d.textures[m] = ints[ints.size() / 2];
        }
        d.materials_started = true;
    }
    compute_initial_guess(d.obs, d.intr, d.num_states, k1, k2, 9.0, false, d.poses, d.positions, d.initialized,
                          [&](unsigned i_obs, unsigned j, const double* Pm1, unsigned i_in_pair) {
                              // normal of the first cloud, rotated into the base frame (:356-357)
                              pose_inverse_apply(Pm1, &d.normal_obs[3 * size_t(i_obs)], true, &d.normals[3 * size_t(j)]);
                              // the reference reads material_ids at the INLIER index, not at the observation
                              // (:369-370); kept unless --material-by-observation
                              const size_t mi = material_by_observation ? size_t(i_obs) : size_t(i_in_pair);
                              d.vertex_material[j] = d.material_ids.at(mi);
                          });
}

// stage: 0 = the problem as the flags say (joint solve when use_light), 1 = stereo blocks only,
// 2 = lighting only (every pose and vertex position constant)
static void solve_stage(PhongDataset& d, unsigned k1, unsigned k2, bool use_light, int max_iters, int stage) {
    Problem problem;
    problem.SetCamera(d.intr[0], d.intr[1], d.intr[2], d.intr[3], d.intr[4]);
    const double cs[9] = {d.stereo_var[0], 0, 0, 0, d.stereo_var[1], 0, 0, 0, d.stereo_var[2]};
    const double cn[9] = {d.normal_var[0], 0, 0, 0, d.normal_var[1], 0, 0, 0, d.normal_var[2]};
    double Ws[9], Wn[9];
    sym_inverse_sqrt(cs, 3, Ws);                      // :31-35
    sym_inverse_sqrt(cn, 3, Wn);                      // :37-41
    const double int_stiffness = 1. / std::sqrt(d.int_var);  // :43
    const bool light = use_light && stage != 1;
    for (unsigned k = k1; k < k2; ++k) {
        double* P = &d.poses[12 * size_t(k)];
        problem.AddPoseBlock(P);                      // :72
        for (unsigned i : d.obs.state_obs[k]) {
            const unsigned j = d.obs.j[i];
            if (!d.initialized[j]) continue;          // :57
            double* X = &d.positions[3 * size_t(j)];
            problem.AddStereoBlock(P, X, &d.obs.uvd[3 * size_t(i)], Ws);  // :59-67
            if (light) {
                const unsigned m = d.vertex_material[j];
                problem.AddLightingBlocks(P, X, &d.normals[3 * size_t(j)], &d.materials[3 * size_t(m)], &d.textures[m], d.light,
                                          d.intensity[i], int_stiffness, &d.normal_obs[3 * size_t(i)], Wn);  // :103-190
            }
        }
    }
    problem.SetParameterBlockConstant(&d.poses[12 * size_t(k1)]);  // :76
    if (stage == 2) {
        for (unsigned k = k1; k < k2; ++k) problem.SetParameterBlockConstant(&d.poses[12 * size_t(k)]);  // :222
        problem.SetPointsConstant(true);                                                                // :215-220
    }
    if (light) {
        const double lo[3] = {0., 0., 1.}, hi[3] = {1., 1., HUGE_VAL};
        problem.SetMaterialBounds(lo, hi);            // :143-172
        problem.SetTextureBounds(0., 1.);             // :177-181
        problem.SetLightDirectional(d.directional);   // :199-203
    }
    problem.options.max_num_iterations = max_iters;   // :83 (1000)
    problem.options.use_nonmonotonic_steps = 1;       // :84
    problem.options.trust_region_strategy = g_dogleg ? 1 : 0;  // :85 ceres::DOGLEG
    problem.options.dogleg_type = 1;                  // :86 ceres::SUBSPACE_DOGLEG
    Summary summary;
    problem.Solve(&summary);
    std::cout << summary.BriefReport() << std::endl << std::endl;
}

static void solveWindow(PhongDataset& d, unsigned k1, unsigned k2, bool use_light, bool multi_stage, int max_iters) {
    std::cerr << "Working on interval [" << k1 << "," << k2 << ")" << std::endl;
    if (multi_stage) {
        std::cerr << "Solving stage 1: poses and points" << std::endl;   // :94-98
        solve_stage(d, k1, k2, use_light, max_iters, 1);
        std::cerr << "Solving stage 2: lighting" << std::endl;            // :226-229
        solve_stage(d, k1, k2, use_light, max_iters, 2);
    }
    std::cerr << "Solving SLAM and lighting jointly" << std::endl;       // :249-252
    solve_stage(d, k1, k2, use_light, max_iters, 0);
}

static void write_outputs(const PhongDataset& d, const std::string& filename) {
    const std::string stem = file_stem(filename);
    write_poses_csv(stem + "_poses.csv", d.poses, d.num_states);
    std::ofstream map_file(stem + "_map.csv"), light_file(stem + "_lights.csv");
    map_file.precision(17);
    light_file.precision(17);
    map_file << "point_id, x, y, z, nx, ny, nz, ka, ks, exponent, kd\n";  // :212-218
    for (unsigned j = 0; j < d.num_vertices; ++j) {
        if (!d.initialized[j]) continue;
        const unsigned m = d.vertex_material[j];
        map_file << j;
        for (int c = 0; c < 3; ++c) map_file << "," << d.positions[3 * size_t(j) + c];
        for (int c = 0; c < 3; ++c) map_file << "," << d.normals[3 * size_t(j) + c];
        for (int c = 0; c < 3; ++c) map_file << "," << d.materials[3 * size_t(m) + c];
        map_file << "," << d.textures[m] << "\n";
    }
    light_file << (d.directional ? "i, j, k\n" : "x, y, z\n");            // :221-227
    light_file << d.light[0] << "," << d.light[1] << "," << d.light[2] << "\n";
}

int main(int argc, char** argv) {
    const std::string usage(
        "usage: dataset_ba_phong_b200 <input_file> [--nolight | --dirlight] [--window N] [--multistage] [--max-iters M] "
        "[--material-by-observation] [--strategy dogleg|lm]");
    if (argc < 2) {
        std::cerr << usage << std::endl;
        return EXIT_FAILURE;
    }
    bool use_light = true, directional = false, use_window = false, by_obs = false, multi_stage = false;
    unsigned window = 0;
    int max_iters = 1000;
    const std::string filename(argv[1]);
    for (int a = 2; a < argc; ++a) {
        const std::string flag(argv[a]);
        if (flag == "--nolight") use_light = false, directional = false;
        else if (flag == "--dirlight") use_light = true, directional = true;
        else if (flag == "--multistage") multi_stage = true, use_light = true;   // :283-285
        else if (flag == "--window" && argc > a + 1) use_window = true, window = unsigned(std::atoi(argv[++a]));
        else if (flag == "--max-iters" && argc > a + 1) max_iters = std::atoi(argv[++a]);
        else if (flag == "--material-by-observation") by_obs = true;
        else if (flag == "--strategy" && argc > a + 1) g_dogleg = std::string(argv[++a]) != "lm";
        else {
            std::cerr << usage << std::endl;
            return EXIT_FAILURE;
        }
    }
    PhongDataset d;
    d.directional = directional;
    if (!read_csv(filename, d)) return EXIT_FAILURE;
    std::cerr << "Computing VO initial guess" << std::endl;
    initial_guess(d, 0, d.num_states, by_obs);                      // :309
    write_poses_csv(file_stem(filename) + "_initial_poses.csv", d.poses, d.num_states);  // :312-314
    if (!use_window || window == 0 || window > d.num_states) window = d.num_states;
    for (unsigned k1 = 0; k1 + window <= d.num_states; ++k1) {
        const unsigned k2 = k1 + window;
        if (k1 > 0) initial_guess(d, k2 - 1, k2, by_obs);           // :324-328 (a single pose: nothing to align)
        else initial_guess(d, k1, k2, by_obs);
        solveWindow(d, k1, k2, use_light, multi_stage, max_iters);
    }
    std::cerr << "Outputting to file " << std::endl;
    write_outputs(d, filename);
    return EXIT_SUCCESS;
}
