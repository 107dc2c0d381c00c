// C ABI of the cslam_b200 back end (include/cslam_b200.h).  Nothing throws across the boundary:
// every entry point returns a status and records a message for cslam_last_error().
#include <cmath>
#include <cstring>
#include <stdexcept>

#include "comm.h"
#include "kernels.cuh"

using cslam::Engine;

struct cslam_problem {
    Engine* e;
};

namespace {
template <class F>
cslam_status guarded(cslam_problem* p, F&& f) {
    if (!p || !p->e) return CSLAM_ERR_INVALID;
    try {
        f(*p->e);
        return CSLAM_OK;
    } catch (const cslam::CudaError& ex) {
        p->e->err = ex.what();
        return CSLAM_ERR_CUDA;
    } catch (const cslam::NotImplemented& ex) {
        p->e->err = ex.what();
        return CSLAM_ERR_NOT_IMPL;
    } catch (const std::invalid_argument& ex) {
        p->e->err = ex.what();
        return CSLAM_ERR_INVALID;
    } catch (const std::domain_error& ex) {
        p->e->err = ex.what();
        return CSLAM_ERR_NUMERIC;
    } catch (const std::exception& ex) {
        p->e->err = ex.what();
        return CSLAM_ERR_COMM;
    }
}
}  // namespace

extern "C" {

void cslam_options_init(cslam_options* o) {
    std::memset(o, 0, sizeof(*o));
    o->max_num_iterations = 1000;
    o->use_nonmonotonic_steps = 1;
    o->max_consecutive_nonmonotonic_steps = 5;
    o->initial_trust_region_radius = 1e4;
    o->max_trust_region_radius = 1e16;
    o->min_trust_region_radius = 1e-32;
    o->min_relative_decrease = 1e-3;
    o->min_lm_diagonal = 1e-6;
    o->max_lm_diagonal = 1e32;
    o->max_num_consecutive_invalid_steps = 5;
    o->function_tolerance = 1e-6;
    o->gradient_tolerance = 1e-10;
    o->parameter_tolerance = 1e-8;
    o->jacobi_scaling = 1;
    o->linear_solver = 0;
    o->preconditioner = 1;
    o->eta = 0.1;
    o->max_linear_solver_iterations = 500;
    o->min_linear_solver_iterations = 0;
    o->num_threads = 8;
    o->device = 0;
    o->profile_kernels = 0;
    o->schur_path = 0;
    o->band_leaves = 0;
    o->window_path = 0;
    o->band_separator_solver = 0;
    o->trust_region_strategy = 0;
    o->dogleg_type = 1;
    o->line_search_sufficient_function_decrease = 1e-4;
    o->dense_solver = 0;
    o->bandpc_solver = 0;
}

cslam_status cslam_problem_create(cslam_problem** out, const cslam_options* opt) {
    if (!out) return CSLAM_ERR_INVALID;
    cslam_options o;
    if (opt)
        o = *opt;
    else
        cslam_options_init(&o);
    try {
        *out = new cslam_problem{new Engine(o)};
    } catch (...) {
        *out = nullptr;
        return CSLAM_ERR_INVALID;
    }
    return CSLAM_OK;
}

void cslam_problem_destroy(cslam_problem* p) {
    if (!p) return;
    delete p->e;
    delete p;
}

const char* cslam_last_error(const cslam_problem* p) { return (p && p->e) ? p->e->err.c_str() : "null handle"; }

cslam_status cslam_set_options(cslam_problem* p, const cslam_options* opt) {
    return guarded(p, [&](Engine& e) {
        if (!opt) throw std::invalid_argument("null options");
        e.opt = *opt;
    });
}

cslam_status cslam_set_camera(cslam_problem* p, double fu, double fv, double cu, double cv, double b) {
    return guarded(p, [&](Engine& e) { e.cam = cslam::CameraIntrinsics{fu, fv, cu, cv, b}; });
}

cslam_status cslam_set_poses(cslam_problem* p, uint32_t n, double* poses12, const uint8_t* constant) {
    return guarded(p, [&](Engine& e) {
        if (!poses12 || n == 0) throw std::invalid_argument("poses: null or empty");
        e.h_poses = poses12;
        e.n_poses = n;
        e.pose_const.assign(n, 0);
        if (constant) std::memcpy(e.pose_const.data(), constant, n);
        e.uploaded = e.begun = false;
    });
}

cslam_status cslam_set_points(cslam_problem* p, uint32_t n, double* xyz) {
    return guarded(p, [&](Engine& e) {
        if (!xyz && n) throw std::invalid_argument("points: null");
        e.h_points = xyz;
        e.n_points = n;
        e.uploaded = e.begun = false;
    });
}

cslam_status cslam_add_stereo(cslam_problem* p, uint64_t n, const uint32_t* cam, const uint32_t* pt, const double* uvd,
                              const double* W, int W_per_obs) {
    return guarded(p, [&](Engine& e) {
        if (n && (!cam || !pt || !uvd || !W)) throw std::invalid_argument("stereo: null array");
        for (uint64_t i = 0; i < n; ++i)
            if (cam[i] >= e.n_poses || pt[i] >= e.n_points) throw std::invalid_argument("stereo block index out of range");
        e.n_st = n;
        e.st_cam = cam;
        e.st_pt = pt;
        e.st_uvd = uvd;
        e.st_W = W;
        e.st_W_per_obs = W_per_obs ? 1 : 0;
        e.uploaded = e.begun = false;
    });
}

cslam_status cslam_add_sun(cslam_problem* p, uint32_t n, const uint32_t* cam, const double* obs_c, const double* ref_g,
                           const double* W2x2, double az_thresh, double zen_thresh, double huber) {
    return guarded(p, [&](Engine& e) {
        for (uint32_t i = 0; i < n; ++i) {
            if (cam[i] >= e.n_poses) throw std::invalid_argument("sun block index out of range");
            cslam::SunBlockData s;
            s.cam = cam[i];
            // the functor's constructor normalises both directions (sun_sensor_error.hpp:30-31)
            double no = 0, nr = 0;
            for (int k = 0; k < 3; ++k) {
                no += obs_c[3 * i + k] * obs_c[3 * i + k];
                nr += ref_g[3 * i + k] * ref_g[3 * i + k];
            }
            no = std::sqrt(no);
            nr = std::sqrt(nr);
            for (int k = 0; k < 3; ++k) {
                s.obs_c[k] = obs_c[3 * i + k] / no;
                s.ref_g[k] = ref_g[3 * i + k] / nr;
            }
            std::memcpy(s.W, W2x2 + 4 * i, 32);
            s.az_thresh = az_thresh;
            s.zen_thresh = zen_thresh;
            s.huber = huber;
            e.suns.push_back(s);
        }
        e.uploaded = e.begun = false;
    });
}

cslam_status cslam_add_pose_prior(cslam_problem* p, uint32_t cam, const double* Tref12, const double* W6x6) {
    return guarded(p, [&](Engine& e) {
        if (cam >= e.n_poses) throw std::invalid_argument("prior block index out of range");
        cslam::PriorBlockData d;
        d.cam = cam;
        std::memcpy(d.Tref, Tref12, 96);
        std::memcpy(d.W, W6x6, 288);
        e.priors.push_back(d);
        e.uploaded = e.begun = false;
    });
}

cslam_status cslam_evaluate(cslam_problem* p, int apply_loss, double* cost, double* r_stereo, double* Jpose_stereo,
                            double* Jpoint_stereo, double* r_sun, double* J_sun, double* r_prior, double* J_prior) {
    return guarded(p, [&](Engine& e) {
        e.evaluate(apply_loss, cost, r_stereo, Jpose_stereo, Jpoint_stereo, r_sun, J_sun, r_prior, J_prior);
    });
}

cslam_status cslam_set_vertices(cslam_problem* p, uint32_t n, double* normals3, double* textures,
                                const uint32_t* material_id) {
    return guarded(p, [&](Engine& e) {
        if (n && (!normals3 || !textures || !material_id)) throw std::invalid_argument("vertices: null array");
        e.h_normals = normals3;
        e.h_textures = textures;
        e.h_material_id = material_id;
        e.n_vertices = n;
        e.h_tex_shared = nullptr;
        e.h_texture_id = nullptr;
        e.n_tex_shared = 0;
        e.phong_ready = false;
        e.uploaded = e.begun = false;
    });
}
cslam_status cslam_set_textures(cslam_problem* p, uint32_t n_textures, double* kd, const uint32_t* vertex_texture_id) {
    return guarded(p, [&](Engine& e) {
        if (!kd || !vertex_texture_id || n_textures == 0) throw std::invalid_argument("textures: null or empty");
        if (e.n_vertices == 0) throw std::invalid_argument("textures: call cslam_set_vertices first");
        for (uint32_t j = 0; j < e.n_vertices; ++j)
            if (vertex_texture_id[j] >= n_textures) throw std::invalid_argument("texture index out of range");
        e.h_tex_shared = kd;
        e.h_texture_id = vertex_texture_id;
        e.n_tex_shared = n_textures;
        e.phong_ready = false;
        e.uploaded = e.begun = false;
    });
}
cslam_status cslam_set_bounds(cslam_problem* p, int block_kind, const double* lower, const double* upper) {
    return guarded(p, [&](Engine& e) {
        if (!lower || !upper) throw std::invalid_argument("bounds: null");
        if (block_kind == 0) {
            for (int k = 0; k < 3; ++k) e.mat_lo[k] = lower[k], e.mat_hi[k] = upper[k];
        } else if (block_kind == 1) {
            e.tex_lo = lower[0];
            e.tex_hi = upper[0];
        } else {
            throw std::invalid_argument("bounds: block_kind must be 0 (material) or 1 (texture)");
        }
        e.bounded = true;
        e.uploaded = e.begun = false;
    });
}
cslam_status cslam_set_points_constant(cslam_problem* p, int constant) {
    return guarded(p, [&](Engine& e) {
        e.hold_positions = constant != 0;
        e.uploaded = e.begun = false;
    });
}
cslam_status cslam_set_materials(cslam_problem* p, uint32_t n_materials, double* phong3) {
    return guarded(p, [&](Engine& e) {
        if (!phong3 || n_materials == 0) throw std::invalid_argument("materials: null or empty");
        e.h_phong = phong3;
        e.n_materials = n_materials;
        e.uploaded = e.begun = false;
    });
}
cslam_status cslam_set_light(cslam_problem* p, double* light3, int directional) {
    return guarded(p, [&](Engine& e) {
        if (!light3) throw std::invalid_argument("light: null");
        e.h_light = light3;
        e.light_directional = directional ? 1 : 0;
        e.uploaded = e.begun = false;
    });
}
cslam_status cslam_add_phong(cslam_problem* p, uint64_t n, const uint32_t* cam, const uint32_t* vertex,
                             const double* intensity, double int_stiffness, const double* normal_obs3,
                             const double* W_normal9) {
    return guarded(p, [&](Engine& e) {
        if (n && (!cam || !vertex || !intensity || !normal_obs3 || !W_normal9)) throw std::invalid_argument("phong: null array");
        for (uint64_t i = 0; i < n; ++i)
            if (cam[i] >= e.n_poses || vertex[i] >= e.n_points) throw std::invalid_argument("lighting block index out of range");
        e.n_ph = n;
        e.ph_cam = cam;
        e.ph_vertex = vertex;
        e.ph_intensity = intensity;
        e.ph_normal_obs = normal_obs3;
        e.ph_int_stiffness = int_stiffness;
        std::memcpy(e.ph_W_normal, W_normal9, 72);
        e.phong_ready = false;
        e.uploaded = e.begun = false;
    });
}
cslam_status cslam_evaluate_phong(cslam_problem* p, double* cost, double* r_int, double* J_int, double* r_normal,
                                  double* Jpose_normal, double* Jn_normal) {
    return guarded(p, [&](Engine& e) { e.evaluate_phong(cost, r_int, J_int, r_normal, Jpose_normal, Jn_normal); });
}
cslam_status cslam_time_phong(cslam_problem* p, int reps, double* ms_per_launch) {
    return guarded(p, [&](Engine& e) {
        if (!ms_per_launch || reps <= 0) throw std::invalid_argument("time_phong: bad arguments");
        *ms_per_launch = e.time_phong(reps);
    });
}

cslam_status cslam_upload(cslam_problem* p) {
    return guarded(p, [&](Engine& e) { e.upload(); });
}
cslam_status cslam_lm_begin(cslam_problem* p) {
    return guarded(p, [&](Engine& e) { e.lm_begin(); });
}
cslam_status cslam_lm_iterate(cslam_problem* p, int n, int ignore_convergence, cslam_summary* summary) {
    return guarded(p, [&](Engine& e) { e.lm_iterate(n, ignore_convergence != 0, summary); });
}
cslam_status cslam_download(cslam_problem* p) {
    return guarded(p, [&](Engine& e) { e.download(); });
}
cslam_status cslam_reset_state(cslam_problem* p) {
    return guarded(p, [&](Engine& e) { e.reset_state(); });
}

cslam_status cslam_solve(cslam_problem* p, cslam_summary* summary) {
    return guarded(p, [&](Engine& e) {
        if (e.window_eligible() || (e.opt.window_path == 2 && !e.lighting_in_solve())) {
            // configs 1/2: a sliding window is one CTA with the LM loop on the device
            Engine* one = &e;
            cslam::solve_window_batch(&one, 1, summary);
            return;
        }
        e.upload();
        e.lm_begin();
        e.lm_iterate(e.opt.max_num_iterations + 1, false, summary);
        e.download();
    });
}

cslam_status cslam_solve_batch(cslam_problem** problems, int n, cslam_summary* summaries) {
    if (!problems || n <= 0) return CSLAM_ERR_INVALID;
    std::vector<Engine*> es(n);
    for (int i = 0; i < n; ++i) {
        if (!problems[i] || !problems[i]->e) return CSLAM_ERR_INVALID;
        es[i] = problems[i]->e;
    }
    return guarded(problems[0], [&](Engine&) { cslam::solve_window_batch(es.data(), n, summaries); });
}

cslam_status cslam_covariance_block(cslam_problem* p, uint32_t cam, double* cov6x6) {
    return guarded(p, [&](Engine& e) {
        if (!cov6x6) throw std::invalid_argument("covariance: null output");
        e.covariance_block(cam, cov6x6);
    });
}

cslam_status cslam_get_iteration_log(const cslam_problem* p, double* rows, int max_rows, int* n_rows) {
    if (!p || !p->e) return CSLAM_ERR_INVALID;
    const int n = int(p->e->log.size());
    if (n_rows) *n_rows = n;
    for (int i = 0; i < n && i < max_rows && rows; ++i) {
        std::memcpy(rows + CSLAM_LOG_COLS * i, p->e->log[i].v, CSLAM_LOG_COLS * sizeof(double));
        rows[CSLAM_LOG_COLS * i + 1] += p->e->fixed_cost();  // IterationSummary::cost includes Ceres' fixed_cost
    }
    return CSLAM_OK;
}

cslam_status cslam_get_reduced_sizes(const cslam_problem* p, int* n_free_cams, int* nnz_blocks) {
    if (!p || !p->e || !n_free_cams || !nnz_blocks) return CSLAM_ERR_INVALID;
    p->e->get_reduced_sizes(n_free_cams, nnz_blocks);
    return CSLAM_OK;
}
cslam_status cslam_get_reduced_system(const cslam_problem* p, int* rowptr, int* col, double* values, double* rhs,
                                      int* free_cam_ids) {
    return guarded(const_cast<cslam_problem*>(p),
                   [&](Engine& e) { e.get_reduced_system(rowptr, col, values, rhs, free_cam_ids); });
}

cslam_status cslam_get_profile(const cslam_problem* p, cslam_profile* out) {
    if (!p || !p->e || !out) return CSLAM_ERR_INVALID;
    *out = p->e->prof;
    return CSLAM_OK;
}
cslam_status cslam_reset_profile(cslam_problem* p) {
    if (!p || !p->e) return CSLAM_ERR_INVALID;
    std::memset(&p->e->prof, 0, sizeof(p->e->prof));
    return CSLAM_OK;
}
cslam_status cslam_set_stream(cslam_problem* p, void* cuda_stream) {
    return guarded(p, [&](Engine& e) { e.set_stream(static_cast<cudaStream_t>(cuda_stream)); });
}
cslam_status cslam_time_resjac(cslam_problem* p, int reps, double* ms) {
    return guarded(p, [&](Engine& e) { *ms = e.time_resjac(reps); });
}
cslam_status cslam_time_schur(cslam_problem* p, int reps, double* ms) {
    return guarded(p, [&](Engine& e) { *ms = e.time_schur(reps); });
}
cslam_status cslam_measure_fp64_peak(int device, double* tflops) {
    try {
        *tflops = cslam::measure_fp64_peak_tflops(device);
        return CSLAM_OK;
    } catch (...) {
        return CSLAM_ERR_CUDA;
    }
}

cslam_status cslam_analyze(cslam_problem* p, int n_ranks, int rank, cslam_structure_info* out) {
    return guarded(p, [&](Engine& e) {
        if (!out || n_ranks < 1 || rank < 0 || rank >= n_ranks) throw std::invalid_argument("bad arguments");
        e.analyze(n_ranks, rank, out);
    });
}

cslam_status cslam_ransac_align(int device, uint32_t n_pairs, const uint32_t* offsets, const double* pts0,
                                const double* pts1, const double* intr5, uint32_t num_iters, double thresh,
                                int rng_variant, double* T12_out, uint8_t* inlier_out, uint32_t* n_inliers_out) {
    if (!offsets || !pts0 || !pts1 || !intr5 || !T12_out || num_iters == 0 || rng_variant < 0 || rng_variant > 1)
        return CSLAM_ERR_INVALID;
    for (uint32_t p = 0; p < n_pairs; ++p)
        if (offsets[p + 1] < offsets[p]) return CSLAM_ERR_INVALID;
    try {
        cslam::ransac_align_batch(device, n_pairs, offsets, pts0, pts1, intr5, num_iters, thresh, rng_variant, T12_out,
                                  inlier_out, n_inliers_out);
        return CSLAM_OK;
    } catch (...) {
        return CSLAM_ERR_CUDA;
    }
}
cslam_status cslam_ransac_triples(uint32_t n, uint32_t num_iters, int rng_variant, uint32_t* triples) {
    if (n < 3 || !triples || rng_variant < 0 || rng_variant > 1) return CSLAM_ERR_INVALID;
    cslam::ransac_triples(n, num_iters, rng_variant, triples);
    return CSLAM_OK;
}

cslam_status cslam_get_launch_count(uint64_t* count) {
    if (!count) return CSLAM_ERR_INVALID;
    *count = cslam::g_kernel_launches.load();
    return CSLAM_OK;
}

cslam_status cslam_comm_unique_id(uint8_t id[128]) {
    try {
        cslam::comm_unique_id(id);
        return CSLAM_OK;
    } catch (...) {
        return CSLAM_ERR_COMM;
    }
}
cslam_status cslam_attach_comm(cslam_problem* p, int n_ranks, int rank, const uint8_t id[128]) {
    return guarded(p, [&](Engine& e) {
        if (n_ranks < 1 || rank < 0 || rank >= n_ranks) throw std::invalid_argument("bad rank / n_ranks");
        CSLAM_CUDA(cudaSetDevice(e.opt.device));
        e.n_ranks = n_ranks;
        e.rank = rank;
        if (n_ranks > 1) {
            e.nccl_comm = cslam::comm_create(n_ranks, rank, id);
            // NCCL connects its channels lazily at the first collective of each kind: do that here,
            // not inside the first solve
            // (with messages of the size a solve sends: large collectives use more channels and other algorithms
            // than small ones, and each of those is connected at its own first use — ~80 ms at 8 ranks)
            const size_t nw = size_t(4) << 20;   // 32 MB of doubles
            double* tmp = nullptr;
            CSLAM_CUDA(cudaMalloc(&tmp, nw * sizeof(double)));
            CSLAM_CUDA(cudaMemset(tmp, 0, nw * sizeof(double)));
            cslam::comm_allreduce_sum(e.nccl_comm, tmp, 64, nullptr);
            cslam::comm_allreduce_sum(e.nccl_comm, tmp, nw, nullptr);
            cslam::comm_allreduce_max(e.nccl_comm, tmp, 1, nullptr);
            cslam::comm_broadcast(e.nccl_comm, tmp, 64, 0, nullptr);
            cslam::comm_broadcast(e.nccl_comm, tmp, nw / 8, 0, nullptr);
            cslam::comm_allgather_bytes(e.nccl_comm, tmp, (nw * sizeof(double) / size_t(n_ranks)) & ~size_t(255), rank, nullptr);
            CSLAM_CUDA(cudaDeviceSynchronize());
            cudaFree(tmp);
        }
        e.uploaded = e.begun = false;
    });
}

}  // extern "C"
