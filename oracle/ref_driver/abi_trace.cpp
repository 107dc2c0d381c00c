// TEST INFRASTRUCTURE — a recording layer with the signatures of include/cslam_b200.h (prefix `cslam_trace_`): every
// call is appended to $CSLAM_ABI_TRACE as one JSON line with its arguments at full precision, then forwarded to the CPU
// oracle (`cslam_oracle_*`); after a solve the in-out arrays the caller registered (poses, points, normals, textures,
// materials, light) are logged as the solver left them.  Two drivers that state the same problems through the C ABI
// produce the same stream: tests/test_ref_driver.py compares the reference's own dataset_ba_phong.cpp (over the
// Ceres-API facade) with this repo's restated driver that way.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iomanip>
#include <map>
#include <string>

#define CSLAM_REMAP_PREFIX cslam_oracle_
#include "abi_remap.h"
#include "../../include/cslam_b200.h"   // declares cslam_oracle_* through the remap

namespace {
struct Arrays {
    double *poses = nullptr, *points = nullptr, *normals = nullptr, *textures = nullptr, *materials = nullptr, *light = nullptr;
    uint32_t n_poses = 0, n_points = 0, n_textures = 0, n_materials = 0;
};
std::map<const cslam_problem*, Arrays> g_arrays;

struct Line {
    std::ofstream f;
    bool first = true;
    explicit Line(const char* call) {
        const char* path = std::getenv("CSLAM_ABI_TRACE");
        if (path) f.open(path, std::ios::app);
        f << std::setprecision(17) << "{\"call\": \"" << call << "\"";
    }
    ~Line() { f << "}\n"; }
    template <class T>
    Line& num(const char* k, T v) {
        f << ", \"" << k << "\": " << v;
        return *this;
    }
    template <class T>
    Line& arr(const char* k, const T* p, size_t n) {
        f << ", \"" << k << "\": [";
        for (size_t i = 0; p && i < n; ++i) f << (i ? ", " : "") << +p[i];
        f << "]";
        return *this;
    }
};
}  // namespace

extern "C" {
void cslam_trace_options_init(cslam_options* opt) { cslam_oracle_options_init(opt); }
cslam_status cslam_trace_problem_create(cslam_problem** out, const cslam_options* o) {
    Line("problem_create").num("max_num_iterations", o->max_num_iterations).num("use_nonmonotonic_steps", o->use_nonmonotonic_steps)
        .num("trust_region_strategy", o->trust_region_strategy).num("dogleg_type", o->dogleg_type).num("linear_solver", o->linear_solver)
        .num("function_tolerance", o->function_tolerance).num("gradient_tolerance", o->gradient_tolerance)
        .num("parameter_tolerance", o->parameter_tolerance).num("initial_trust_region_radius", o->initial_trust_region_radius);
    return cslam_oracle_problem_create(out, o);
}
void cslam_trace_problem_destroy(cslam_problem* p) {
    g_arrays.erase(p);
    cslam_oracle_problem_destroy(p);
}
const char* cslam_trace_last_error(const cslam_problem* p) { return cslam_oracle_last_error(p); }
cslam_status cslam_trace_set_camera(cslam_problem* p, double fu, double fv, double cu, double cv, double b) {
    Line("set_camera").num("fu", fu).num("fv", fv).num("cu", cu).num("cv", cv).num("b", b);
    return cslam_oracle_set_camera(p, fu, fv, cu, cv, b);
}
cslam_status cslam_trace_set_poses(cslam_problem* p, uint32_t n, double* poses12, const uint8_t* constant) {
    Line("set_poses").arr("poses", poses12, 12 * size_t(n)).arr("constant", constant, n);
    g_arrays[p].poses = poses12, g_arrays[p].n_poses = n;
    return cslam_oracle_set_poses(p, n, poses12, constant);
}
cslam_status cslam_trace_set_points(cslam_problem* p, uint32_t n, double* xyz) {
    Line("set_points").arr("points", xyz, 3 * size_t(n));
    g_arrays[p].points = xyz, g_arrays[p].n_points = n;
    return cslam_oracle_set_points(p, n, xyz);
}
cslam_status cslam_trace_add_stereo(cslam_problem* p, uint64_t n, const uint32_t* cam, const uint32_t* pt, const double* uvd,
                                    const double* W, int W_per_obs) {
    Line("add_stereo").arr("cam", cam, n).arr("pt", pt, n).arr("uvd", uvd, 3 * n).arr("W", W, W_per_obs ? 9 * n : 9).num("W_per_obs", W_per_obs);
    return cslam_oracle_add_stereo(p, n, cam, pt, uvd, W, W_per_obs);
}
cslam_status cslam_trace_add_sun(cslam_problem* p, uint32_t n, const uint32_t* cam, const double* obs_c, const double* ref_g,
                                 const double* W2x2, double az, double zen, double huber) {
    Line("add_sun").arr("cam", cam, n).arr("obs", obs_c, 3 * size_t(n)).arr("ref", ref_g, 3 * size_t(n)).arr("W", W2x2, 4 * size_t(n))
        .num("az", az).num("zen", zen).num("huber", huber);
    return cslam_oracle_add_sun(p, n, cam, obs_c, ref_g, W2x2, az, zen, huber);
}
cslam_status cslam_trace_add_pose_prior(cslam_problem* p, uint32_t cam, const double* Tref12, const double* W6x6) {
    Line("add_pose_prior").num("cam", cam).arr("Tref", Tref12, 12).arr("W", W6x6, 36);
    return cslam_oracle_add_pose_prior(p, cam, Tref12, W6x6);
}
cslam_status cslam_trace_set_points_constant(cslam_problem* p, int constant) {
    Line("set_points_constant").num("constant", constant);
    return cslam_oracle_set_points_constant(p, constant);
}
cslam_status cslam_trace_add_phong(cslam_problem* p, uint64_t n, const uint32_t* cam, const uint32_t* vertex, const double* intensity,
                                   double int_stiffness, const double* normal_obs3, const double* W_normal9) {
    Line("add_phong").arr("cam", cam, n).arr("vertex", vertex, n).arr("intensity", intensity, n).num("int_stiffness", int_stiffness)
        .arr("normal_obs", normal_obs3, 3 * n).arr("W_normal", W_normal9, 9);
    return cslam_oracle_add_phong(p, n, cam, vertex, intensity, int_stiffness, normal_obs3, W_normal9);
}
cslam_status cslam_trace_set_bounds(cslam_problem* p, int kind, const double* lower, const double* upper) {
    const size_t n = kind == 0 ? 3 : 1;
    // (infinite bounds are not JSON numbers)
    double lo[3], hi[3];
    for (size_t i = 0; i < n; ++i) lo[i] = lower[i] < -1e300 ? -1e300 : lower[i], hi[i] = upper[i] > 1e300 ? 1e300 : upper[i];
    Line("set_bounds").num("kind", kind).arr("lower", lo, n).arr("upper", hi, n);
    return cslam_oracle_set_bounds(p, kind, lower, upper);
}
cslam_status cslam_trace_set_light(cslam_problem* p, double* light3, int directional) {
    Line("set_light").arr("light", light3, 3).num("directional", directional);
    g_arrays[p].light = light3;
    return cslam_oracle_set_light(p, light3, directional);
}
cslam_status cslam_trace_set_materials(cslam_problem* p, uint32_t n, double* phong3) {
    Line("set_materials").arr("materials", phong3, 3 * size_t(n));
    g_arrays[p].materials = phong3, g_arrays[p].n_materials = n;
    return cslam_oracle_set_materials(p, n, phong3);
}
cslam_status cslam_trace_set_textures(cslam_problem* p, uint32_t n, double* kd, const uint32_t* vertex_texture_id) {
    Line("set_textures").arr("textures", kd, n).arr("vertex_texture_id", vertex_texture_id, g_arrays[p].n_points);
    g_arrays[p].textures = kd, g_arrays[p].n_textures = n;
    return cslam_oracle_set_textures(p, n, kd, vertex_texture_id);
}
cslam_status cslam_trace_set_vertices(cslam_problem* p, uint32_t n, double* normals3, double* textures, const uint32_t* material_id) {
    Line("set_vertices").arr("normals", normals3, 3 * size_t(n)).arr("material_id", material_id, n);
    g_arrays[p].normals = normals3;
    return cslam_oracle_set_vertices(p, n, normals3, textures, material_id);
}
cslam_status cslam_trace_covariance_block(cslam_problem* p, uint32_t cam, double* cov6x6) {
    const cslam_status st = cslam_oracle_covariance_block(p, cam, cov6x6);
    Line("covariance_block").num("cam", cam).num("status", int(st)).arr("covariance", cov6x6, st == CSLAM_OK ? 36 : 0);
    return st;
}
cslam_status cslam_trace_solve(cslam_problem* p, cslam_summary* s) {
    const cslam_status st = cslam_oracle_solve(p, s);
    const Arrays& a = g_arrays[p];
    Line("solve").num("status", int(st)).num("iterations", s->num_iterations).num("initial_cost", s->initial_cost)
        .num("final_cost", s->final_cost).num("termination", s->termination_type).arr("poses", a.poses, 12 * size_t(a.n_poses))
        .arr("points", a.points, 3 * size_t(a.n_points)).arr("normals", a.normals, a.normals ? 3 * size_t(a.n_points) : 0)
        .arr("textures", a.textures, a.n_textures).arr("materials", a.materials, 3 * size_t(a.n_materials)).arr("light", a.light, a.light ? 3 : 0);
    return st;
}
}  // extern "C"
