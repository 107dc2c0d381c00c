// TEST INFRASTRUCTURE — the reference's OWN driver, /root/reference/tests/dataset_vo.cpp, compiled unmodified
// (its `main` renamed) against the Ceres-API facade of ceres/ceres.h in this directory: read_csv, the window loop,
// compute_initial_guess, solveWindow's problem assembly and options are the reference's code; ceres::Solve lands on
// this repo's C ABI.  Built by `make -C oracle ref` into oracle/_ref/ together with the reference's
// dataset_problem.cpp / point_cloud_aligner.cpp / utils.cpp.
#define main cslam_ref_dataset_vo_main_impl
#include "dataset_vo.cpp"   // found through -I$(REFERENCE)/tests
#undef main

extern "C" int cslam_ref_dataset_vo_main(int argc, char** argv) {
    try {
        return cslam_ref_dataset_vo_main_impl(argc, argv);
    } catch (const std::exception& e) {
        std::cerr << "reference driver failed: " << e.what() << std::endl;
        return 70;
    }
}
