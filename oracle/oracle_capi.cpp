// ORACLE — TEST INFRASTRUCTURE ONLY.  Pinned to the reference's own unmodified headers / sources compiled in
// oracle/_ref: every residual / Jacobian / plus block and the RANSAC front end (tests/test_ref_pin.py,
// tests/test_ref_frontend.py).  UNPINNED: the solver rules of problem.hpp / phong_problem.hpp (Ceres is absent).
// C entry points over the CPU restatement, shaped like include/cslam_b200.h (prefix
// `cslam_oracle_`) so the same Python harness can drive either library.  Only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs load this.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>

#include "../include/cslam_b200.h"
#include "phong_problem.hpp"
#include "problem.hpp"
#include "ransac.hpp"

using namespace oracle;

struct cslam_oracle_problem {
    Problem prob;
    PhongProblem ph;            // lighting blocks of dataset_ba_phong (filled by add_phong)
    Options opt;
    Summary last;
    std::string err;
};

static Options to_options(const cslam_options* o) {
    Options r;
    if (!o) return r;
    r.max_num_iterations = o->max_num_iterations;
    r.use_nonmonotonic_steps = o->use_nonmonotonic_steps;
    r.max_consecutive_nonmonotonic_steps = o->max_consecutive_nonmonotonic_steps;
    r.initial_trust_region_radius = o->initial_trust_region_radius;
    r.max_trust_region_radius = o->max_trust_region_radius;
    r.min_trust_region_radius = o->min_trust_region_radius;
    r.min_relative_decrease = o->min_relative_decrease;
    r.min_lm_diagonal = o->min_lm_diagonal;
    r.max_lm_diagonal = o->max_lm_diagonal;
    r.max_num_consecutive_invalid_steps = o->max_num_consecutive_invalid_steps;
    r.function_tolerance = o->function_tolerance;
    r.gradient_tolerance = o->gradient_tolerance;
    r.parameter_tolerance = o->parameter_tolerance;
    r.jacobi_scaling = o->jacobi_scaling;
    r.linear_solver = o->linear_solver;
    r.preconditioner = o->preconditioner;
    r.eta = o->eta;
    r.max_linear_solver_iterations = o->max_linear_solver_iterations;
    r.min_linear_solver_iterations = o->min_linear_solver_iterations;
    r.num_threads = o->num_threads > 0 ? o->num_threads : 1;
    r.trust_region_strategy = o->trust_region_strategy;
    r.dogleg_type = o->dogleg_type;
    if (o->line_search_sufficient_function_decrease > 0.0)
        r.line_search_sufficient_function_decrease = o->line_search_sufficient_function_decrease;
    return r;
}

extern "C" {

void cslam_oracle_options_init(cslam_options* o) {
    Options d;
    std::memset(o, 0, sizeof(*o));
    o->max_num_iterations = d.max_num_iterations;
    o->use_nonmonotonic_steps = d.use_nonmonotonic_steps;
    o->max_consecutive_nonmonotonic_steps = d.max_consecutive_nonmonotonic_steps;
    o->initial_trust_region_radius = d.initial_trust_region_radius;
    o->max_trust_region_radius = d.max_trust_region_radius;
    o->min_trust_region_radius = d.min_trust_region_radius;
    o->min_relative_decrease = d.min_relative_decrease;
    o->min_lm_diagonal = d.min_lm_diagonal;
    o->max_lm_diagonal = d.max_lm_diagonal;
    o->max_num_consecutive_invalid_steps = d.max_num_consecutive_invalid_steps;
    o->function_tolerance = d.function_tolerance;
    o->gradient_tolerance = d.gradient_tolerance;
    o->parameter_tolerance = d.parameter_tolerance;
    o->jacobi_scaling = d.jacobi_scaling;
    o->linear_solver = d.linear_solver;
    o->preconditioner = d.preconditioner;
    o->eta = d.eta;
    o->max_linear_solver_iterations = d.max_linear_solver_iterations;
    o->min_linear_solver_iterations = d.min_linear_solver_iterations;
    o->num_threads = d.num_threads;
    o->trust_region_strategy = d.trust_region_strategy;
    o->dogleg_type = d.dogleg_type;
    o->line_search_sufficient_function_decrease = d.line_search_sufficient_function_decrease;
}

int cslam_oracle_problem_create(cslam_oracle_problem** out, const cslam_options* opt) {
    if (!out) return CSLAM_ERR_INVALID;
    *out = new cslam_oracle_problem();
    (*out)->opt = to_options(opt);
    return CSLAM_OK;
}
void cslam_oracle_problem_destroy(cslam_oracle_problem* p) { delete p; }
const char* cslam_oracle_last_error(const cslam_oracle_problem* p) { return p ? p->err.c_str() : "null handle"; }
int cslam_oracle_set_options(cslam_oracle_problem* p, const cslam_options* opt) {
    p->opt = to_options(opt);
    return CSLAM_OK;
}
int cslam_oracle_set_camera(cslam_oracle_problem* p, double fu, double fv, double cu, double cv, double b) {
    p->prob.camera = Camera{fu, fv, cu, cv, b};
    return CSLAM_OK;
}
int cslam_oracle_set_poses(cslam_oracle_problem* p, uint32_t n, double* poses12, const uint8_t* constant) {
    p->prob.poses = poses12;
    p->prob.n_poses = int(n);
    p->prob.pose_const.assign(n, 0);
    if (constant) std::memcpy(p->prob.pose_const.data(), constant, n);
    return CSLAM_OK;
}
int cslam_oracle_set_points(cslam_oracle_problem* p, uint32_t n, double* xyz) {
    p->prob.points = xyz;
    p->prob.n_points = int(n);
    return CSLAM_OK;
}
int cslam_oracle_add_stereo(cslam_oracle_problem* p, uint64_t n, const uint32_t* cam, const uint32_t* pt,
                            const double* uvd, const double* W, int W_per_obs) {
    Problem& q = p->prob;
    for (uint64_t i = 0; i < n; ++i)
        if (int(cam[i]) >= q.n_poses || int(pt[i]) >= q.n_points) {
            p->err = "stereo block index out of range";
            return CSLAM_ERR_INVALID;
        }
    q.st_cam.assign(cam, cam + n);
    q.st_pt.assign(pt, pt + n);
    q.st_uvd.assign(uvd, uvd + 3 * n);
    q.st_W_per_obs = W_per_obs != 0;
    q.st_W.assign(W, W + (W_per_obs ? 9 * n : 9));
    return CSLAM_OK;
}
int cslam_oracle_add_sun(cslam_oracle_problem* p, uint32_t n, const uint32_t* cam, const double* obs_c,
                         const double* ref_g, const double* W2x2, double az_thresh, double zen_thresh,
                         double huber) {
    for (uint32_t i = 0; i < n; ++i) {
        Problem::Sun s;
        s.cam = cam[i];
        if (int(s.cam) >= p->prob.n_poses) {
            p->err = "sun block index out of range";
            return CSLAM_ERR_INVALID;
        }
        std::memcpy(s.f.observed_sun_dir_c, obs_c + 3 * i, 24);
        std::memcpy(s.f.expected_sun_dir_g, ref_g + 3 * i, 24);
        std::memcpy(s.f.stiffness, W2x2 + 4 * i, 32);
        s.f.az_err_thresh = az_thresh;
        s.f.zen_err_thresh = zen_thresh;
        s.f.normalize_inputs();
        s.huber = huber;
        p->prob.suns.push_back(s);
    }
    return CSLAM_OK;
}
int cslam_oracle_add_pose_prior(cslam_oracle_problem* p, uint32_t cam, const double* Tref12, const double* W6x6) {
    if (int(cam) >= p->prob.n_poses) {
        p->err = "prior block index out of range";
        return CSLAM_ERR_INVALID;
    }
    Problem::Prior pr;
    pr.cam = cam;
    std::memcpy(pr.f.T_ref, Tref12, 96);
    std::memcpy(pr.f.stiffness, W6x6, 288);
    p->prob.priors.push_back(pr);
    return CSLAM_OK;
}
int cslam_oracle_evaluate(cslam_oracle_problem* p, int apply_loss, double* cost, double* r_stereo,
                          double* Jpose_stereo, double* Jpoint_stereo, double* r_sun, double* J_sun,
                          double* r_prior, double* J_prior) {
    Problem::Eval ev;
    Problem& q = p->prob;
    if (!q.evaluate(q.poses, q.points, true, apply_loss != 0, ev, p->opt.num_threads)) {
        p->err = "evaluation failed";
        if (cost) *cost = ev.cost;
        return CSLAM_ERR_NUMERIC;
    }
    if (cost) *cost = ev.cost;
    // constant parameter blocks have no Jacobian columns in ceres::Problem::Evaluate
    // (dataset_vo.cpp:62): report them as zeros
    for (size_t i = 0; i < q.n_stereo(); ++i)
        if (q.pose_const[q.st_cam[i]]) std::fill(&ev.Jc_st[18 * i], &ev.Jc_st[18 * i] + 18, 0.0);
    for (size_t i = 0; i < q.suns.size(); ++i)
        if (q.pose_const[q.suns[i].cam]) std::fill(&ev.J_sun[12 * i], &ev.J_sun[12 * i] + 12, 0.0);
    for (size_t i = 0; i < q.priors.size(); ++i)
        if (q.pose_const[q.priors[i].cam]) std::fill(&ev.J_pr[36 * i], &ev.J_pr[36 * i] + 36, 0.0);
    auto cp = [](double* dst, const std::vector<double>& src) {
        if (dst && !src.empty()) std::memcpy(dst, src.data(), src.size() * sizeof(double));
    };
    cp(r_stereo, ev.r_st);
    cp(Jpose_stereo, ev.Jc_st);
    cp(Jpoint_stereo, ev.Jp_st);
    cp(r_sun, ev.r_sun);
    cp(J_sun, ev.J_sun);
    cp(r_prior, ev.r_pr);
    cp(J_prior, ev.J_pr);
    return CSLAM_OK;
}
// ceres::Covariance::Compute + GetCovarianceBlockInTangentSpace for one pose block (dataset_vo_sun.cpp:159-183), same
// contract as cslam_covariance_block: the (cam, cam) 6x6 block of (J^T J)^-1 in tangent coordinates at the current
// parameter values, loss-corrected Jacobians, no damping.  Dense normal equations over the free poses and the points,
// Cholesky, six solves (a window has a few hundred unknowns).
int cslam_oracle_covariance_block(cslam_oracle_problem* p, uint32_t cam, double* cov36) {
    Problem& q = p->prob;
    if (cam >= q.pose_const.size() || q.pose_const[cam]) {
        p->err = "covariance of a constant or unknown pose";
        return CSLAM_ERR_INVALID;
    }
    Problem::Eval ev;
    if (!q.evaluate(q.poses, q.points, true, true, ev, p->opt.num_threads)) {
        p->err = "evaluation failed";
        return CSLAM_ERR_NUMERIC;
    }
    // columns only for parameter blocks some residual block touches (others are not part of a ceres::Problem)
    const size_t nc = q.pose_const.size();
    std::vector<char> pose_used(nc, 0);
    for (uint32_t c : q.st_cam) pose_used[c] = 1;
    for (const auto& sb : q.suns) pose_used[sb.cam] = 1;
    for (const auto& pb : q.priors) pose_used[pb.cam] = 1;
    std::vector<long> pcol(nc, -1);
    long n = 0;
    for (size_t k = 0; k < nc; ++k)
        if (!q.pose_const[k] && pose_used[k]) pcol[k] = n, n += 6;
    if (pcol[cam] < 0) {
        p->err = "covariance of a pose no residual block touches";
        return CSLAM_ERR_INVALID;
    }
    uint32_t n_pts = 0;
    for (uint32_t j : q.st_pt) n_pts = std::max(n_pts, j + 1);
    std::vector<long> xcol(n_pts, -1);
    for (uint32_t j : q.st_pt)
        if (xcol[j] < 0) xcol[j] = n, n += 3;
    std::vector<double> H(size_t(n) * size_t(n), 0.0);
    // H += A^T B for an m-row block pair with column offsets ca, cb and widths wa, wb (row-major inputs)
    auto acc = [&](const double* A, long ca, int wa, const double* B, long cb, int wb, int m) {
        for (int i = 0; i < wa; ++i)
            for (int j = 0; j < wb; ++j) {
                double a = 0;
                for (int r = 0; r < m; ++r) a += A[r * wa + i] * B[r * wb + j];
                H[size_t(ca + i) * n + size_t(cb + j)] += a;
            }
    };
    for (size_t i = 0; i < q.n_stereo(); ++i) {
        const double *Jc = &ev.Jc_st[18 * i], *Jp = &ev.Jp_st[9 * i];
        const long cc = pcol[q.st_cam[i]], cp = xcol[q.st_pt[i]];
        if (cc >= 0) {
            acc(Jc, cc, 6, Jc, cc, 6, 3);
            acc(Jc, cc, 6, Jp, cp, 3, 3);
            acc(Jp, cp, 3, Jc, cc, 6, 3);
        }
        acc(Jp, cp, 3, Jp, cp, 3, 3);
    }
    for (size_t i = 0; i < q.suns.size(); ++i)
        if (pcol[q.suns[i].cam] >= 0) acc(&ev.J_sun[12 * i], pcol[q.suns[i].cam], 6, &ev.J_sun[12 * i], pcol[q.suns[i].cam], 6, 2);
    for (size_t i = 0; i < q.priors.size(); ++i)
        if (pcol[q.priors[i].cam] >= 0) acc(&ev.J_pr[36 * i], pcol[q.priors[i].cam], 6, &ev.J_pr[36 * i], pcol[q.priors[i].cam], 6, 6);
    for (long j = 0; j < n; ++j) {
        double d = H[size_t(j) * n + j];
        for (long k = 0; k < j; ++k) d -= H[size_t(j) * n + k] * H[size_t(j) * n + k];
        if (!(d > 0.0)) {
            p->err = "covariance: the Jacobian is rank deficient";
            return CSLAM_ERR_NUMERIC;
        }
        d = std::sqrt(d);
        H[size_t(j) * n + j] = d;
        for (long i = j + 1; i < n; ++i) {
            double a = H[size_t(i) * n + j];
            for (long k = 0; k < j; ++k) a -= H[size_t(i) * n + k] * H[size_t(j) * n + k];
            H[size_t(i) * n + j] = a / d;
        }
    }
    const long c0 = pcol[cam];
    std::vector<double> y(n);
    for (int c = 0; c < 6; ++c) {
        std::fill(y.begin(), y.end(), 0.0);
        y[c0 + c] = 1.0;
        for (long i = 0; i < n; ++i) {
            double a = y[i];
            for (long k = 0; k < i; ++k) a -= H[size_t(i) * n + k] * y[k];
            y[i] = a / H[size_t(i) * n + i];
        }
        for (long i = n - 1; i >= 0; --i) {
            double a = y[i];
            for (long k = i + 1; k < n; ++k) a -= H[size_t(k) * n + i] * y[k];
            y[i] = a / H[size_t(i) * n + i];
        }
        for (int r = 0; r < 6; ++r) cov36[6 * r + c] = y[c0 + r];
    }
    return CSLAM_OK;
}
// ---- lighting blocks (dataset_ba_phong.cpp:100-205) ------------------------------------------------
int cslam_oracle_set_vertices(cslam_oracle_problem* p, uint32_t n, double* normals3, double* textures,
                              const uint32_t* material_id) {
    PhongProblem& q = p->ph;
    q.n_vertices = int(n);
    q.normals = normals3;
    q.v_mat.assign(material_id, material_id + n);
    // per-vertex texture values unless cslam_oracle_set_textures shares them
    q.textures = textures;
    q.n_tex = int(n);
    q.v_tex.resize(n);
    for (uint32_t j = 0; j < n; ++j) q.v_tex[j] = j;
    return CSLAM_OK;
}
int cslam_oracle_set_textures(cslam_oracle_problem* p, uint32_t n_textures, double* kd, const uint32_t* vertex_texture_id) {
    PhongProblem& q = p->ph;
    for (int j = 0; j < q.n_vertices; ++j)
        if (vertex_texture_id[j] >= n_textures) {
            p->err = "texture index out of range";
            return CSLAM_ERR_INVALID;
        }
    q.textures = kd;
    q.n_tex = int(n_textures);
    q.v_tex.assign(vertex_texture_id, vertex_texture_id + q.n_vertices);
    return CSLAM_OK;
}
int cslam_oracle_set_materials(cslam_oracle_problem* p, uint32_t n, double* phong3) {
    p->ph.materials = phong3;
    p->ph.n_mat = int(n);
    return CSLAM_OK;
}
int cslam_oracle_set_light(cslam_oracle_problem* p, double* light3, int directional) {
    p->ph.light = light3;
    p->ph.directional = directional != 0;
    return CSLAM_OK;
}
int cslam_oracle_set_bounds(cslam_oracle_problem* p, int block_kind, const double* lower, const double* upper) {
    PhongProblem& q = p->ph;
    if (block_kind == 0) {
        for (int k = 0; k < 3; ++k) q.mat_lo[k] = lower[k], q.mat_hi[k] = upper[k];
    } else if (block_kind == 1) {
        q.tex_lo = lower[0];
        q.tex_hi = upper[0];
    } else {
        p->err = "set_bounds: block_kind must be 0 (material) or 1 (texture)";
        return CSLAM_ERR_INVALID;
    }
    q.bounded = true;
    return CSLAM_OK;
}
int cslam_oracle_set_points_constant(cslam_oracle_problem* p, int constant) {
    p->ph.hold_positions = constant != 0;
    return CSLAM_OK;
}
int cslam_oracle_add_phong(cslam_oracle_problem* p, uint64_t n, const uint32_t* cam, const uint32_t* vertex,
                           const double* intensity, double int_stiffness, const double* normal_obs3,
                           const double* W_normal9) {
    PhongProblem& q = p->ph;
    q.cam.assign(cam, cam + n);
    q.vtx.assign(vertex, vertex + n);
    q.intensity.assign(intensity, intensity + n);
    q.normal_obs.assign(normal_obs3, normal_obs3 + 3 * n);
    q.int_stiffness = int_stiffness;
    std::memcpy(q.Wn, W_normal9, sizeof(q.Wn));
    return CSLAM_OK;
}
// joins the stereo blocks (same (pose, vertex) list, dataset_ba_phong.cpp:55-69 and :100-190 walk the
// same observations) with the lighting blocks
static int bind_phong(cslam_oracle_problem* p) {
    PhongProblem& q = p->ph;
    const Problem& s = p->prob;
    if (s.st_cam != q.cam || s.st_pt != q.vtx) {
        p->err = "lighting blocks must pair one-to-one with the stereo blocks (same pose / vertex lists)";
        return CSLAM_ERR_INVALID;
    }
    if (s.st_W_per_obs || !s.suns.empty() || !s.priors.empty()) {
        p->err = "lighting solve: shared stereo stiffness, no sun / prior blocks";
        return CSLAM_ERR_NOT_IMPL;
    }
    if (!q.normals || !q.materials || !q.light || q.n_vertices != s.n_points) {
        p->err = "lighting solve: vertices / materials / light not set";
        return CSLAM_ERR_INVALID;
    }
    q.camera = s.camera;
    q.poses = s.poses;
    q.n_poses = s.n_poses;
    q.pose_const = s.pose_const;
    q.positions = s.points;
    q.uvd = s.st_uvd;
    std::memcpy(q.W, s.st_W.data(), sizeof(q.W));
    return CSLAM_OK;
}
int cslam_oracle_solve(cslam_oracle_problem* p, cslam_summary* s) {
    bool ok;
    if (!p->ph.cam.empty()) {
        const int st = bind_phong(p);
        if (st != CSLAM_OK) return st;
        ok = p->ph.solve(p->opt, p->last);
    } else {
        ok = p->prob.solve(p->opt, p->last);
    }
    if (s) {
        std::memset(s, 0, sizeof(*s));
        s->initial_cost = p->last.initial_cost;
        s->final_cost = p->last.final_cost;
        s->num_iterations = p->last.num_iterations;
        s->num_successful_steps = p->last.num_successful_steps;
        s->num_unsuccessful_steps = p->last.num_unsuccessful_steps;
        s->termination_type = p->last.termination_type;
        s->termination_reason = p->last.termination_reason;
        s->final_radius = p->last.final_radius;
        s->total_linear_iterations = p->last.total_linear_iterations;
    }
    return ok ? CSLAM_OK : CSLAM_ERR_NUMERIC;
}
int cslam_oracle_get_iteration_log(const cslam_oracle_problem* p, double* rows, int max_rows, int* n_rows) {
    const int n = int(p->last.rows.size());
    if (n_rows) *n_rows = n;
    for (int i = 0; i < n && i < max_rows; ++i) std::memcpy(rows + CSLAM_LOG_COLS * i, &p->last.rows[i], sizeof(IterationRow));
    return CSLAM_OK;
}

// wall-clock stamp (seconds, steady clock) of every logged iteration row of the last solve
int cslam_oracle_get_iteration_seconds(const cslam_oracle_problem* p, double* t, int max_rows, int* n_rows) {
    const int n = int(p->last.row_seconds.size());
    if (n_rows) *n_rows = n;
    for (int i = 0; i < n && i < max_rows; ++i) t[i] = p->last.row_seconds[i];
    return CSLAM_OK;
}

// ---- front end: 3-point RANSAC point-cloud alignment (point_cloud_aligner.cpp:64-136) ---------------
int cslam_oracle_ransac_align(int, uint32_t n_pairs, const uint32_t* offsets, const double* pts0, const double* pts1,
                              const double* intr5, uint32_t num_iters, double thresh, int rng_variant, double* T12_out,
                              uint8_t* inlier_out, uint32_t* n_inliers_out) {
    const Camera cam{intr5[0], intr5[1], intr5[2], intr5[3], intr5[4]};
    for (uint32_t p = 0; p < n_pairs; ++p) {
        const uint32_t o = offsets[p], n = offsets[p + 1] - o;
        std::vector<uint32_t> in = ransac_align(cam, pts0 + 3 * size_t(o), pts1 + 3 * size_t(o), n, num_iters, thresh,
                                                rng_variant, T12_out + 12 * size_t(p));
        if (inlier_out) {
            std::fill(inlier_out + o, inlier_out + o + n, uint8_t(0));
            for (uint32_t i : in) inlier_out[o + i] = 1;
        }
        if (n_inliers_out) n_inliers_out[p] = uint32_t(in.size());
    }
    return CSLAM_OK;
}
// the draws of the restated distribution next to std::uniform_int_distribution of this compiler
void cslam_oracle_ransac_draws(uint32_t n, uint32_t count, int variant, uint32_t* restated, uint32_t* libstdcxx) {
    std::mt19937 a(42), b(42);
    std::uniform_int_distribution<unsigned> d(0, n - 1);
    for (uint32_t i = 0; i < count; ++i) {
        restated[i] = ransac_draw(a, n, variant);
        libstdcxx[i] = d(b);
    }
}
void cslam_oracle_kabsch(uint32_t n, const double* pts0, const double* pts1, double* T12) {
    std::vector<const double*> a, b;
    for (uint32_t i = 0; i < n; ++i) {
        a.push_back(pts0 + 3 * size_t(i));
        b.push_back(pts1 + 3 * size_t(i));
    }
    kabsch(a, b, T12);
}

// ---- DOGLEG scalar pieces, for the known-answer tests ---------------------------------------------
int cslam_oracle_poly_root_real_parts(const double* coeffs, int n_coeffs, double* out) {
    const std::vector<double> r = polynomial_root_real_parts(std::vector<double>(coeffs, coeffs + n_coeffs));
    for (size_t i = 0; i < r.size(); ++i) out[i] = r[i];
    return int(r.size());
}
int cslam_oracle_dogleg_boundary_minimum(const double* B4, const double* g2, double radius, double* x2) {
    return dogleg_boundary_minimum(B4, g2, radius, x2) ? 1 : 0;
}

// ---- direct access to the restated geometry / models, for the known-answer tests ----------
void cslam_oracle_so3_exp(const double* phi, double* R) { so3_exp(phi, R); }
void cslam_oracle_so3_log(const double* R, double* phi) { so3_log(R, phi); }
void cslam_oracle_se3_exp(const double* xi, double* T12) {
    SE3<double> X = se3_exp(xi);
    std::memcpy(T12, X.d, 96);
}
void cslam_oracle_se3_log(const double* T12, double* xi) { se3_log(SE3<double>::from(T12), xi); }
void cslam_oracle_se3_mul(const double* A12, const double* B12, double* C12) {
    SE3<double> C = se3_mul(SE3<double>::from(A12), SE3<double>::from(B12));
    std::memcpy(C12, C.d, 96);
}
void cslam_oracle_se3_inverse(const double* A12, double* C12) {
    SE3<double> C = se3_inverse(SE3<double>::from(A12));
    std::memcpy(C12, C.d, 96);
}
void cslam_oracle_se3_adjoint(const double* A12, double* Ad36) { se3_adjoint(SE3<double>::from(A12), Ad36); }
void cslam_oracle_se3_transform(const double* A12, const double* p, int is_vector, double* out) {
    SE3<double> A = SE3<double>::from(A12);
    if (is_vector)
        se3_transform_vector(A, p, out);
    else
        se3_transform_point(A, p, out);
}
void cslam_oracle_se3_plus(const double* T12, const double* eps6, double* out12) {
    SE3Perturbation P;
    P(T12, eps6, out12);
}
void cslam_oracle_se3_plus_jacobian(const double* T12, double* J72) {
    autodiff_plus_jacobian(SE3Perturbation(), T12, J72);
}
void cslam_oracle_unit_plus(const double* x3, const double* d3, double* out3) {
    UnitVectorPerturbation P;
    P(x3, d3, out3);
}
void cslam_oracle_unit_plus_jacobian(const double* x3, double* J9) {
    autodiff_plus_jacobian(UnitVectorPerturbation(), x3, J9);
}
void cslam_oracle_camera_project(const double* intr5, const double* pt_c, double* uvd) {
    Camera c{intr5[0], intr5[1], intr5[2], intr5[3], intr5[4]};
    camera_project(c, pt_c, uvd);
}
void cslam_oracle_camera_triangulate(const double* intr5, const double* uvd, double* pt_c) {
    Camera c{intr5[0], intr5[1], intr5[2], intr5[3], intr5[4]};
    camera_triangulate(c, uvd, pt_c);
}
// PointLight::shade with the camera at `campos` (light_test.cpp:65-68)
double cslam_oracle_point_light_shade(const double* light_pos, const double* vpos, const double* vnormal,
                                      const double* phong3, double texture, const double* campos) {
    return point_light_shade<double>(light_pos, vpos, vnormal, phong3, texture, campos, 1.0);
}
// One intensity block: residual and tangent-space Jacobians
//   J_pose 1x6, J_point 1x3, J_normal 1x3 (through UnitVectorPerturbation), J_phong 1x3,
//   J_tex 1x1, J_light 1x3 (through UnitVectorPerturbation when directional)
int cslam_oracle_intensity_block(const double* pose12, const double* pt3, const double* n3,
                                 const double* phong3, const double* tex1, const double* light3,
                                 double colour, double stiffness, int directional, double* r,
                                 double* J_pose, double* J_point, double* J_normal, double* J_phong,
                                 double* J_tex, double* J_light) {
    IntensityError f;
    f.colour = colour;
    f.stiffness = stiffness;
    f.directional = directional != 0;
    const double* params[6] = {pose12, pt3, n3, phong3, tex1, light3};
    double Ja[12], Jp[3], Jn[3], Jk[3], Jt[1], Jl[3];
    double* jac[6] = {Ja, Jp, Jn, Jk, Jt, Jl};
    if (!autodiff_cost(f, params, r, jac)) return CSLAM_ERR_NUMERIC;
    double P[72], Un[9], Ul[9];
    autodiff_plus_jacobian(SE3Perturbation(), pose12, P);
    autodiff_plus_jacobian(UnitVectorPerturbation(), n3, Un);
    for (int c = 0; c < 6; ++c) {
        double s = 0;
        for (int k = 0; k < 12; ++k) s += Ja[k] * P[6 * k + c];
        J_pose[c] = s;
    }
    for (int c = 0; c < 3; ++c) {
        J_point[c] = Jp[c];
        J_phong[c] = Jk[c];
        J_normal[c] = Jn[0] * Un[c] + Jn[1] * Un[3 + c] + Jn[2] * Un[6 + c];
    }
    J_tex[0] = Jt[0];
    if (directional) {
        autodiff_plus_jacobian(UnitVectorPerturbation(), light3, Ul);
        for (int c = 0; c < 3; ++c) J_light[c] = Jl[0] * Ul[c] + Jl[1] * Ul[3 + c] + Jl[2] * Ul[6 + c];
    } else {
        for (int c = 0; c < 3; ++c) J_light[c] = Jl[c];
    }
    return CSLAM_OK;
}
// One normal block: residual(3), J_pose 3x6, J_normal 3x3 (through UnitVectorPerturbation)
int cslam_oracle_normal_block(const double* pose12, const double* n3, const double* obs3, const double* W9,
                              double* r, double* J_pose, double* J_normal) {
    NormalError f;
    std::memcpy(f.obs_normal_c, obs3, 24);
    std::memcpy(f.stiffness, W9, 72);
    const double* params[2] = {pose12, n3};
    double Ja[36], Jn[9];
    double* jac[2] = {Ja, Jn};
    if (!autodiff_cost(f, params, r, jac)) return CSLAM_ERR_NUMERIC;
    double P[72], Un[9];
    autodiff_plus_jacobian(SE3Perturbation(), pose12, P);
    autodiff_plus_jacobian(UnitVectorPerturbation(), n3, Un);
    for (int rr = 0; rr < 3; ++rr) {
        for (int c = 0; c < 6; ++c) {
            double s = 0;
            for (int k = 0; k < 12; ++k) s += Ja[12 * rr + k] * P[6 * k + c];
            J_pose[6 * rr + c] = s;
        }
        for (int c = 0; c < 3; ++c)
            J_normal[3 * rr + c] = Jn[3 * rr] * Un[c] + Jn[3 * rr + 1] * Un[3 + c] + Jn[3 * rr + 2] * Un[6 + c];
    }
    return CSLAM_OK;
}

}  // extern "C"
