// TEST INFRASTRUCTURE — sends the C ABI calls that ceres_slam_b200/host/cslam_problem.hpp makes to another library
// with the same signatures.  Define CSLAM_REMAP_PREFIX before including:
//   cslam_oracle_  the CPU oracle (oracle/oracle_capi.cpp)
//   cslam_trace_   the recording layer of abi_trace.cpp (logs every call, forwards to the oracle)
// Must come before include/cslam_b200.h is first included (the declarations are renamed with the calls).
#ifndef CSLAM_REF_DRIVER_ABI_REMAP
#define CSLAM_REF_DRIVER_ABI_REMAP
#define CSLAM_REMAP_CAT2(a, b) a##b
#define CSLAM_REMAP_CAT(a, b) CSLAM_REMAP_CAT2(a, b)
#define CSLAM_REMAP(name) CSLAM_REMAP_CAT(CSLAM_REMAP_PREFIX, name)
#define cslam_options_init CSLAM_REMAP(options_init)
#define cslam_problem_create CSLAM_REMAP(problem_create)
#define cslam_problem_destroy CSLAM_REMAP(problem_destroy)
#define cslam_last_error CSLAM_REMAP(last_error)
#define cslam_set_camera CSLAM_REMAP(set_camera)
#define cslam_set_poses CSLAM_REMAP(set_poses)
#define cslam_set_points CSLAM_REMAP(set_points)
#define cslam_add_stereo CSLAM_REMAP(add_stereo)
#define cslam_add_sun CSLAM_REMAP(add_sun)
#define cslam_add_pose_prior CSLAM_REMAP(add_pose_prior)
#define cslam_solve CSLAM_REMAP(solve)
#define cslam_set_points_constant CSLAM_REMAP(set_points_constant)
#define cslam_add_phong CSLAM_REMAP(add_phong)
#define cslam_set_bounds CSLAM_REMAP(set_bounds)
#define cslam_set_light CSLAM_REMAP(set_light)
#define cslam_set_materials CSLAM_REMAP(set_materials)
#define cslam_set_textures CSLAM_REMAP(set_textures)
#define cslam_set_vertices CSLAM_REMAP(set_vertices)
#define cslam_covariance_block CSLAM_REMAP(covariance_block)
#endif
