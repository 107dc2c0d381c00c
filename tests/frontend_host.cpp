// TEST SHIM — the product's host front end (ceres_slam_b200/host: the CSV readers of the three restated
// drivers, `compute_initial_guess` with its matching / chaining / vertex initialisation) as a library a CPU test
// can drive, WITHOUT a GPU: every call of `cslam_ransac_align` inside the included driver is redirected to an
// alignment entry the test injects (the reference's own point_cloud_aligner.cpp from oracle/_ref, or the oracle).
// tests/test_ref_frontend.py compares the result with the reference's own DatasetProblem* classes.
//
// Compiled three times (-DKIND=0 dataset_vo_b200, 1 dataset_vo_sun_b200, 2 dataset_ba_phong_b200): the driver
// source is included as it is, its `main` renamed.  With -DFH_ABI_TRACE also the C ABI calls of the driver's solves are
// redirected (recorded, then answered by the CPU oracle): the whole driver runs here.  Nothing here is shipped.
#include <cstdint>
#include <cstring>

#ifdef FH_ABI_TRACE
// the whole driver on the CPU: its C ABI calls go through the recording layer oracle/ref_driver/abi_trace.cpp to the oracle
#define CSLAM_REMAP_PREFIX cslam_trace_
#include "../oracle/ref_driver/abi_remap.h"
#endif
#include "../include/cslam_b200.h"

typedef cslam_status (*ransac_entry_t)(int, uint32_t, const uint32_t*, const double*, const double*, const double*, uint32_t,
                                       double, int, double*, uint8_t*, uint32_t*);
static ransac_entry_t g_entry = nullptr;
static cslam_status fh_ransac_hook(int device, uint32_t n_pairs, const uint32_t* offsets, const double* p0, const double* p1,
                                   const double* intr5, uint32_t num_iters, double thresh, int rng_variant, double* T12,
                                   uint8_t* inl, uint32_t* cnt) {
    if (!g_entry) return CSLAM_ERR_INVALID;
    return g_entry(device, n_pairs, offsets, p0, p1, intr5, num_iters, thresh, rng_variant, T12, inl, cnt);
}
#define cslam_ransac_align fh_ransac_hook
#define main driver_main

#if KIND == 0
#include "../ceres_slam_b200/host/dataset_vo_b200.cpp"
typedef Track Data;
#elif KIND == 1
#include "../ceres_slam_b200/host/dataset_vo_sun_b200.cpp"
typedef SunDataset Data;
#else
#include "../ceres_slam_b200/host/dataset_ba_phong_b200.cpp"
typedef PhongDataset Data;
#endif
#undef main

extern "C" {

// the driver's own main (argument parsing, window loop, output files)
int fh_driver_main(int argc, char** argv) {
    try {
        return driver_main(argc, argv);
    } catch (const std::exception& e) {
        std::cerr << "driver failed: " << e.what() << std::endl;
        return 70;
    }
}

void fh_set_ransac_entry(void* fn) { g_entry = reinterpret_cast<ransac_entry_t>(fn); }

void* fh_open(const char* f1, const char* f2, const char* f3, int dir_light) {
    Data* d = new Data();
    bool ok = false;
    try {
#if KIND == 0
        ok = read_csv(f1, *d);
#elif KIND == 1
        ok = read_csv(f1, f2, f3, *d);
#else
        d->directional = dir_light != 0;
        ok = read_csv(f1, *d);
#endif
    } catch (...) {
        ok = false;
    }
    (void)f2; (void)f3; (void)dir_light;
    if (!ok) {
        delete d;
        return nullptr;
    }
    return d;
}
void fh_close(void* h) { delete static_cast<Data*>(h); }

void fh_dims(void* h, uint32_t* n_states, uint32_t* n_points, uint64_t* n_obs, uint32_t* n_materials) {
    Data& d = *static_cast<Data*>(h);
    *n_states = d.num_states;
    *n_obs = d.obs.k.size();
#if KIND == 2
    *n_points = d.num_vertices;
    *n_materials = d.num_materials;
#else
    *n_points = d.num_points;
    *n_materials = 0;
#endif
}
void fh_observations(void* h, uint32_t* state_of_obs, uint32_t* point_ids, double* uvd, double* intr5, double* var3) {
    Data& d = *static_cast<Data*>(h);
    for (size_t k = 0; k < d.obs.state_obs.size(); ++k)
        for (unsigned i : d.obs.state_obs[k]) state_of_obs[i] = uint32_t(k);
    for (size_t i = 0; i < d.obs.j.size(); ++i) point_ids[i] = d.obs.j[i];
    std::memcpy(uvd, d.obs.uvd.data(), d.obs.uvd.size() * sizeof(double));
    std::memcpy(intr5, d.intr, 5 * sizeof(double));
#if KIND == 0
    std::memcpy(var3, d.var, 3 * sizeof(double));
#elif KIND == 2
    std::memcpy(var3, d.stereo_var, 3 * sizeof(double));
#else
    (void)var3;
#endif
}
#if KIND == 1
void fh_sun_data(void* h, double* stereo_covars9, uint8_t* has_sun, double* sun_obs3, double* sun_covar4, double* sun_dir_g3) {
    Data& d = *static_cast<Data*>(h);
    std::memcpy(stereo_covars9, d.obs_covars.data(), d.obs_covars.size() * sizeof(double));
    for (unsigned k = 0; k < d.num_states; ++k) has_sun[k] = uint8_t(d.has_sun[k]);
    std::memcpy(sun_obs3, d.sun_obs.data(), d.sun_obs.size() * sizeof(double));
    std::memcpy(sun_covar4, d.sun_covars.data(), d.sun_covars.size() * sizeof(double));
    std::memcpy(sun_dir_g3, d.sun_dir_g.data(), d.sun_dir_g.size() * sizeof(double));
}
#endif
#if KIND == 2
void fh_phong_data(void* h, uint32_t* material_ids, double* intensities, double* normal_obs3, double* normal_var3, double* int_var,
                   double* light3) {
    Data& d = *static_cast<Data*>(h);
    for (size_t i = 0; i < d.material_ids.size(); ++i) material_ids[i] = d.material_ids[i];
    std::memcpy(intensities, d.intensity.data(), d.intensity.size() * sizeof(double));
    std::memcpy(normal_obs3, d.normal_obs.data(), d.normal_obs.size() * sizeof(double));
    std::memcpy(normal_var3, d.normal_var, 3 * sizeof(double));
    *int_var = d.int_var;
    std::memcpy(light3, d.light, 3 * sizeof(double));
}
#endif

// the driver's own front-end call for window [k1, k2) (dataset_vo_b200.cpp main, sun_dataset.hpp, dataset_ba_phong_b200.cpp
// initial_guess); returns 0 when the sun variant gives up
int fh_initial_guess(void* h, uint32_t k1, uint32_t k2) {
    Data& d = *static_cast<Data*>(h);
#if KIND == 0
    compute_initial_guess(d.obs, d.intr, d.num_states, k1, k2, 4.0, false, d.poses, d.points, d.initialized,
                          [](unsigned, unsigned, const double*, unsigned) {});
    return 1;
#elif KIND == 1
    // the call of run_pass (dataset_vo_sun_b200.cpp): threshold 4, gives up below 3 inliers
    return compute_initial_guess(d.obs, d.intr, d.num_states, k1, k2, 4.0, true, d.poses, d.points, d.initialized,
                                 [](unsigned, unsigned, const double*, unsigned) {}).ok ? 1 : 0;
#else
    initial_guess(d, k1, k2, false);
    return 1;
#endif
}
void fh_reset_points(void* h) {
    Data& d = *static_cast<Data*>(h);
    std::fill(d.initialized.begin(), d.initialized.end(), 0);
}
// the driver's output files for `filename` (<stem>_poses.csv, and what else the driver writes)
void fh_write(void* h, const char* filename) {
    Data& d = *static_cast<Data*>(h);
#if KIND == 1
    write_poses_csv(file_stem(filename) + "_poses.csv", d.poses, d.num_states);   // dataset_vo_sun_b200.cpp main
#else
    write_outputs(d, filename);
#endif
}
void fh_state(void* h, double* poses12, double* points3, uint8_t* initialized, double* normals3, double* phong3, double* texture1) {
    Data& d = *static_cast<Data*>(h);
    std::memcpy(poses12, d.poses.data(), d.poses.size() * sizeof(double));
    for (size_t j = 0; j < d.initialized.size(); ++j) initialized[j] = uint8_t(d.initialized[j] != 0);
#if KIND == 2
    const std::vector<double>& pts = d.positions;
#else
    const std::vector<double>& pts = d.points;
#endif
    for (size_t j = 0; j < d.initialized.size(); ++j)
        for (int c = 0; c < 3; ++c) points3[3 * j + c] = d.initialized[j] ? pts[3 * j + c] : 0.0;
#if KIND == 2
    for (size_t j = 0; j < d.initialized.size(); ++j) {
        const unsigned m = d.vertex_material[j];
        for (int c = 0; c < 3; ++c) {
            if (normals3) normals3[3 * j + c] = d.initialized[j] ? d.normals[3 * j + c] : 0.0;
            if (phong3) phong3[3 * j + c] = d.initialized[j] ? d.materials[3 * size_t(m) + c] : 0.0;
        }
        if (texture1) texture1[j] = d.initialized[j] ? d.textures[m] : 0.0;
    }
#else
    (void)normals3; (void)phong3; (void)texture1;
#endif
}

}  // extern "C"
