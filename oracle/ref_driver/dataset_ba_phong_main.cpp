// TEST INFRASTRUCTURE — the reference's OWN driver /root/reference/tests/dataset_ba_phong.cpp, compiled unmodified (its
// `main` renamed) against the Ceres-API facade of ceres/ceres.h in this directory; see dataset_vo_main.cpp.
#define main cslam_ref_dataset_ba_phong_main_impl
#include "dataset_ba_phong.cpp"   // found through -I$(REFERENCE)/tests
#undef main

extern "C" int cslam_ref_dataset_ba_phong_main(int argc, char** argv) {
    try {
        return cslam_ref_dataset_ba_phong_main_impl(argc, argv);
    } catch (const std::exception& e) {
        std::cerr << "reference driver failed: " << e.what() << std::endl;
        return 70;
    }
}
