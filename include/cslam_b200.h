/*
 * cslam_b200 — C ABI of the B200-native bundle-adjustment back end for ceres-slam's hot path.
 *
 * The reference (utiasSTARS/ceres-slam) has no FFI: its seam is the ceres::Problem built inside
 * `solveWindow` and the `ceres::Solve` call that follows (tests/dataset_vo.cpp:22-85,
 * tests/dataset_vo_sun.cpp:25-187, tests/dataset_ba_phong.cpp:26-255).  Every entry point below
 * replaces one group of Ceres calls at that seam; the citation names the reference lines.
 * Plain pointers and sizes only.  All arrays are caller-owned; parameter arrays (poses, points)
 * are updated IN PLACE by cslam_solve, exactly like Ceres writes into `dataset.poses[k].data()`.
 * No function throws; each returns a status and leaves a message for cslam_last_error().
 * A handle is single-threaded; distinct handles may be used from distinct threads.
 * There is no CPU fallback: without a CUDA device every compute entry point fails with
 * CSLAM_ERR_CUDA.
 */
#ifndef CSLAM_B200_H_
#define CSLAM_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct cslam_problem cslam_problem;

typedef enum cslam_status {
    CSLAM_OK = 0,
    CSLAM_ERR_INVALID = 1,   /* bad argument / index out of range / call order */
    CSLAM_ERR_CUDA = 2,      /* CUDA runtime error or no device */
    CSLAM_ERR_NOT_IMPL = 3,
    CSLAM_ERR_NUMERIC = 4,   /* non-finite cost at the initial point */
    CSLAM_ERR_COMM = 5       /* NCCL failure */
} cslam_status;

/* Mirrors the ceres::Solver::Options fields the drivers set (dataset_vo.cpp:65-74) plus the
 * Ceres 1.x defaults the LM loop depends on. */
typedef struct cslam_options {
    int max_num_iterations;                 /* 1000  dataset_vo.cpp:69 */
    int use_nonmonotonic_steps;             /* 1     dataset_vo.cpp:70 */
    int max_consecutive_nonmonotonic_steps; /* 5 */
    double initial_trust_region_radius;     /* 1e4 */
    double max_trust_region_radius;         /* 1e16 */
    double min_trust_region_radius;         /* 1e-32 */
    double min_relative_decrease;           /* 1e-3 */
    double min_lm_diagonal;                 /* 1e-6 */
    double max_lm_diagonal;                 /* 1e32 */
    int max_num_consecutive_invalid_steps;  /* 5 */
    double function_tolerance;              /* 1e-6 */
    double gradient_tolerance;              /* 1e-10 */
    double parameter_tolerance;             /* 1e-8 */
    int jacobi_scaling;                     /* 1 */
    int linear_solver;      /* 0 = exact Schur solve (SPARSE_SCHUR-equivalent): block-banded Cholesky cut
                               into leaves + separators when the reduced system is banded (half-bandwidth
                               <= 12 blocks), chunked bordered-band / dense Cholesky beyond (dense_solver,
                               bandpc_solver), in-kernel dense Cholesky for windows, else PCG run to 1e-15;
                               1 = ITERATIVE_SCHUR (PCG, Ceres Q-rule, eta) */
    int preconditioner;     /* 0 = JACOBI (block diag of B), 1 = SCHUR_JACOBI (block diag of S) */
    double eta;                             /* 0.1 */
    int max_linear_solver_iterations;       /* 500 */
    int min_linear_solver_iterations;       /* 0 */
    int num_threads;        /* ignored by the GPU back end (dataset_vo.cpp:67); used by the oracle */
    int device;             /* CUDA device ordinal */
    int profile_kernels;    /* 1 = bracket every kernel class with CUDA events (cslam_get_profile) */
    int schur_path;         /* 0 = auto, 1 = force generic per-landmark kernel, 2 = force grouped */
    int band_leaves;        /* exact solve of a banded reduced system: number of leaves the band is cut
                               into (0 = auto) */
    int window_path;        /* small problems (<= 8 poses, exact solve): 0 = auto (one-CTA-per-window kernel
                               with the trust-region loop — LEVENBERG_MARQUARDT or DOGLEG — on the device; also
                               cslam_covariance_block as one launch), 1 = never, 2 = require it */
    int band_separator_solver; /* exact solve of a banded reduced system, separator system between the
                               leaves: 0 = auto, 1 = banded Cholesky on one CTA, 2 = block cyclic reduction */
    int trust_region_strategy;  /* 0 = LEVENBERG_MARQUARDT (Ceres default, dataset_vo.cpp),
                               1 = DOGLEG (dataset_vo_sun.cpp:142, dataset_ba_phong.cpp:88) */
    int dogleg_type;            /* 0 = TRADITIONAL_DOGLEG, 1 = SUBSPACE_DOGLEG (dataset_vo_sun.cpp:143) */
    double line_search_sufficient_function_decrease; /* 1e-4: Armijo constant of the line search a
                               bounded problem runs along the trust-region step */
    int dense_solver;       /* exact solve (linear_solver 0) of a reduced system that is NOT a narrow band (loop
                               closures, tracks longer than 13 frames): 0 = auto (dense FP64 Cholesky on the
                               DMMA tensor cores when 6 x free poses <= 12288, i.e. a factor of at most 1.2 GB;
                               PCG run to 1e-15 beyond that), 1 = dense whenever it fits (also instead of the
                               banded solver), -1 = never */
    int bandpc_solver;      /* exact solve of a reduced system that is not a narrow band (tracks longer than 13 frames):
                               0 = auto: a WIDE band (half-bandwidth 13 .. 64 blocks) on a trajectory long compared with it
                               is factored exactly as chunks of bordered bands + a dense separator system (this also
                               replaces the dense factorisation there); otherwise, when too large for the dense
                               factorisation, conjugate gradients preconditioned with the banded direct solve of the
                               short-track landmarks' part of the system, run to a 1e-15 residual;
                               1 = the preconditioned conjugate gradients, never the wide-band factorisation;
                               2 = the wide-band factorisation whenever the layout allows it (also on small problems);
                               -1 = neither (dense factorisation if it fits, else block-Jacobi PCG run to 1e-15) */
} cslam_options;

typedef struct cslam_summary {
    double initial_cost;
    double final_cost;
    int num_iterations;            /* LM iterations attempted (successful + unsuccessful) */
    int num_successful_steps;
    int num_unsuccessful_steps;
    int termination_type;          /* 0 CONVERGENCE, 1 NO_CONVERGENCE, 2 FAILURE */
    int termination_reason;        /* 1 gradient tol, 2 parameter tol, 3 function tol,
                                      4 max iterations, 5 min radius, 6 invalid steps,
                                      7 linear solver, 8 evaluation failed */
    double final_radius;
    int total_linear_iterations;
    double device_ms;              /* CUDA-event time of the LM loop on the problem's stream */
} cslam_summary;

/* One row per LM iteration (row 0 = initial point), 10 doubles each:
 * iteration, cost, cost_change, gradient_max_norm, step_norm, relative_decrease, radius,
 * linear_iterations, step_is_valid, step_is_successful.  Same columns Ceres prints with
 * minimizer_progress_to_stdout (dataset_vo.cpp:66). */
#define CSLAM_LOG_COLS 10

/* Accumulated CUDA-event times per kernel class since the last cslam_reset_profile(). */
typedef struct cslam_profile {
    double ms[16];
    int launches[16];
} cslam_profile;
enum {
    CSLAM_K_RESJAC = 0,      /* materialised residual + Jacobian (cslam_evaluate) */
    CSLAM_K_COLNORM = 1,     /* column norms + gradient (Jacobi scaling, LM diagonal) */
    CSLAM_K_SCHUR = 2,       /* fused residual/Jacobian + Schur elimination */
    CSLAM_K_FINALIZE = 3,    /* LM diagonal on S, block-Jacobi inverse */
    CSLAM_K_PCG = 4,         /* all PCG kernels */
    CSLAM_K_BACKSUB = 5,     /* back-substitution + Plus + model/candidate cost */
    CSLAM_K_ALLREDUCE = 6,   /* NCCL all-reduce of [S | g] */
    CSLAM_K_WINDOW = 7,      /* batched sliding-window LM kernel */
    CSLAM_K_OTHER = 8
};

void cslam_options_init(cslam_options* opt);

/* ceres::Problem problem;  (dataset_vo.cpp:26) */
cslam_status cslam_problem_create(cslam_problem** out, const cslam_options* opt);
void cslam_problem_destroy(cslam_problem* p);
const char* cslam_last_error(const cslam_problem* p);
cslam_status cslam_set_options(cslam_problem* p, const cslam_options* opt);

/* StereoCamera(fu, fv, cu, cv, b)  (stereo_camera.hpp:159-163, dataset_problem.cpp:42-48) */
cslam_status cslam_set_camera(cslam_problem* p, double fu, double fv, double cu, double cv, double b);

/* Pose parameter blocks: n x 12 doubles [t | R row-major] (se3group.hpp:115-118), all with
 * SE3Perturbation (dataset_vo.cpp:58); constant[k] != 0 == SetParameterBlockConstant
 * (dataset_vo.cpp:62).  In-out. `constant` may be NULL. */
cslam_status cslam_set_poses(cslam_problem* p, uint32_t n, double* poses12, const uint8_t* constant);

/* Map point parameter blocks: n x 3 doubles (dataset_vo.cpp:52-53).  In-out. */
cslam_status cslam_set_points(cslam_problem* p, uint32_t n, double* xyz);

/* n StereoReprojectionErrorAutomatic blocks (dataset_vo.cpp:46-53): cam[i], pt[i], uvd[3i..],
 * stiffness W row-major 3x3: 9 doubles shared (dataset_vo.cpp:29-32) or 9 per observation when
 * W_per_obs != 0 (dataset_vo_sun.cpp:57-59).  Replaces earlier stereo blocks. */
cslam_status cslam_add_stereo(cslam_problem* p, uint64_t n, const uint32_t* cam, const uint32_t* pt,
                              const double* uvd, const double* W, int W_per_obs);

/* n SunSensorErrorAutomatic blocks (dataset_vo_sun.cpp:75-101): observed direction (camera
 * frame) and ephemeris direction (global frame), 3 doubles each, normalised by the callee like
 * the functor's constructor (sun_sensor_error.hpp:30-31); W2x2 row-major per block; thresholds
 * in radians; huber <= 0 means no loss (dataset_vo_sun.cpp:89-99). */
cslam_status cslam_add_sun(cslam_problem* p, uint32_t n, const uint32_t* cam, const double* obs_c,
                           const double* ref_g, const double* W2x2, double az_thresh,
                           double zen_thresh, double huber);

/* One PoseErrorAutomatic block (dataset_vo_sun.cpp:109-124). */
cslam_status cslam_add_pose_prior(cslam_problem* p, uint32_t cam, const double* Tref12,
                                  const double* W6x6);

/* ceres::Problem::Evaluate at the current parameter values: cost and, per residual block, the
 * residuals and the Jacobians in tangent coordinates (ambient Jacobian times the Plus Jacobian).
 * Outputs are in the caller's block order; any pointer may be NULL.
 *   r_stereo 3n, Jpose_stereo 18n (3x6 row-major), Jpoint_stereo 9n (3x3 row-major)
 *   r_sun 2m, J_sun 12m (2x6), r_prior 6q, J_prior 36q (6x6)
 * Loss-corrected like Ceres when apply_loss != 0. */
cslam_status cslam_evaluate(cslam_problem* p, int apply_loss, double* cost, double* r_stereo,
                            double* Jpose_stereo, double* Jpoint_stereo, double* r_sun,
                            double* J_sun, double* r_prior, double* J_prior);

/* ---- lighting blocks of dataset_ba_phong (tests/dataset_ba_phong.cpp:100-205) --------------------
 * A map vertex is a point (cslam_set_points) plus a unit normal, a diffuse texture value kd and a
 * material id (map_vertex / material / texture, dataset_problem_phong.cpp:335-376); a material is
 * the Phong parameter block [ka, ks, alpha] (material.hpp); the light is a position or, when
 * `directional` != 0, a direction (dataset_ba_phong.cpp:103-123).  All arrays are caller-owned.
 * The normal (and a directional light) carry UnitVectorPerturbation (perturbations.hpp:87-113). */
cslam_status cslam_set_vertices(cslam_problem* p, uint32_t n, double* normals3, double* textures,
                                const uint32_t* material_id);
/* Texture blocks shared between vertices: vertex j uses kd[vertex_texture_id[j]].  The reference
 * creates one Texture per material and hands the same shared_ptr to every vertex of that material
 * (dataset_problem_phong.cpp:262-279, :342-343), so Ceres sees ONE parameter block per material;
 * this call states that sharing (it replaces the per-vertex values of cslam_set_vertices).  In-out. */
cslam_status cslam_set_textures(cslam_problem* p, uint32_t n_textures, double* kd, const uint32_t* vertex_texture_id);
cslam_status cslam_set_materials(cslam_problem* p, uint32_t n_materials, double* phong3);
/* SetParameterLowerBound / SetParameterUpperBound on every block of one kind
 * (dataset_ba_phong.cpp:143-181): block_kind 0 = material [ka, ks, alpha] (3 lower, 3 upper),
 * 1 = texture kd (1, 1).  +-HUGE_VAL leaves a side open.  A bounded problem is solved the way Ceres
 * does it: every Plus is projected onto the box and the trust-region step goes through an Armijo
 * line search (see oracle/phong_problem.hpp for the stated deviation in the interpolation). */
cslam_status cslam_set_bounds(cslam_problem* p, int block_kind, const double* lower, const double* upper);
cslam_status cslam_set_light(cslam_problem* p, double* light3, int directional);
/* SetParameterBlockConstant / SetParameterBlockVariable on EVERY vertex position block (stage 2 of
 * dataset_ba_phong --multistage, tests/dataset_ba_phong.cpp:209-220 and :236-239).  Lighting solves
 * only.  With the poses constant as well (cslam_set_poses) the stereo blocks have no variable left:
 * like Ceres, the solve drops them and carries their cost as a constant that is reported with every
 * cost but takes no part in the minimiser's decisions. */
cslam_status cslam_set_points_constant(cslam_problem* p, int constant);
/* n observations, each adding one IntensityError{Point,Directional}LightAutomatic block
 * (dataset_ba_phong.cpp:103-123: pose, position, normal, phong params, texture, light) and one
 * NormalErrorAutomatic block (dataset_ba_phong.cpp:183-190: pose, normal).  int_stiffness =
 * 1/sqrt(int_var) (dataset_ba_phong.cpp:43); W_normal9 row-major 3x3 (dataset_ba_phong.cpp:38-41). */
cslam_status cslam_add_phong(cslam_problem* p, uint64_t n, const uint32_t* cam, const uint32_t* vertex,
                             const double* intensity, double int_stiffness, const double* normal_obs3,
                             const double* W_normal9);
/* ceres::Problem::Evaluate restricted to the lighting blocks, caller's block order; any output
 * may be NULL.  Per block: r_int 1; J_int 19 = [pose 6 | point 3 | normal 3 | phong 3 | texture 1 |
 * light 3]; r_normal 3; Jpose_normal 3x6; Jn_normal 3x3.  cost = 1/2 sum of squares. */
cslam_status cslam_evaluate_phong(cslam_problem* p, double* cost, double* r_int, double* J_int,
                                  double* r_normal, double* Jpose_normal, double* Jn_normal);
/* `reps` launches of the lighting-block kernel on device-resident data: ms per launch. */
cslam_status cslam_time_phong(cslam_problem* p, int reps, double* ms_per_launch);

/* ceres::Solve (dataset_vo.cpp:81): uploads the problem, runs the LM loop on the device, writes
 * the best parameters back into the caller's pose / point arrays.
 * With lighting blocks (cslam_add_phong) this is the joint solve of dataset_ba_phong.cpp:249-252:
 * poses, vertex positions and normals, materials, textures and the light are optimised together and
 * all of those arrays are updated in place.  The lighting blocks must pair one-to-one with the stereo
 * blocks (the driver adds both while walking the same observations, :55-69 / :100-190), textures
 * must be shared blocks (cslam_set_textures) and there are at most 1024 shared columns (3 per material
 * + 1 per texture + 3); anything else returns CSLAM_ERR_NOT_IMPL.  A track may be of any length.
 * With a communicator attached (cslam_attach_comm) the vertices are sharded over the ranks. */
cslam_status cslam_solve(cslam_problem* p, cslam_summary* summary);

/* The same, split so a caller can keep the problem resident in HBM:
 *   upload   : host arrays -> device, structure analysis (landmark-major order, tiles, S pattern)
 *   lm_begin : evaluate at the initial point, set up the trust-region state
 *   lm_iterate: run up to n LM iterations (stops early on convergence unless
 *               ignore_convergence != 0); may be called repeatedly
 *   download : best parameters -> host arrays
 *   reset_state: restore the uploaded initial parameter values on the device */
cslam_status cslam_upload(cslam_problem* p);
cslam_status cslam_lm_begin(cslam_problem* p);
cslam_status cslam_lm_iterate(cslam_problem* p, int n, int ignore_convergence, cslam_summary* summary);
cslam_status cslam_download(cslam_problem* p);
cslam_status cslam_reset_state(cslam_problem* p);

/* Config 4 (scripts/ba_all_*.sh: many independent tracks): solve n independent small problems
 * in one launch per GPU and trust-region strategy present (the scripts' default, SUBSPACE_DOGLEG of
 * dataset_vo_sun.cpp:142-143, included).  Each problem must already hold its data. */
cslam_status cslam_solve_batch(cslam_problem** problems, int n, cslam_summary* summaries);

/* ceres::Covariance::Compute + GetCovarianceBlockInTangentSpace for one pose block
 * (dataset_vo_sun.cpp:159-183: the covariance of pose k1+1 becomes the prior of the next window):
 * the (cam, cam) 6x6 block of (J^T J)^-1 in tangent coordinates at the caller's current parameter
 * values, loss-corrected Jacobians, no LM damping; computed as the (cam, cam) block of the inverse of
 * the undamped reduced camera system.  Row-major 6x6.  Fails with CSLAM_ERR_INVALID for a constant
 * pose and CSLAM_ERR_NUMERIC when the reduced system is not positive definite (rank-deficient J). */
cslam_status cslam_covariance_block(cslam_problem* p, uint32_t cam, double* cov6x6);

/* Iteration log of the last solve: up to max_rows rows of CSLAM_LOG_COLS doubles; returns the
 * number of rows available in *n_rows. */
cslam_status cslam_get_iteration_log(const cslam_problem* p, double* rows, int max_rows, int* n_rows);

/* Reduced camera system of the last Schur build (diagnostics / parity): sizes first, then data.
 * S is returned as upper block-CSR with 6x6 row-major blocks; rhs is 6 per free camera. */
cslam_status cslam_get_reduced_sizes(const cslam_problem* p, int* n_free_cams, int* nnz_blocks);
cslam_status cslam_get_reduced_system(const cslam_problem* p, int* rowptr, int* col, double* values,
                                      double* rhs, int* free_cam_ids);

/* Kernel-class timings (CUDA events on the problem's stream). */
cslam_status cslam_get_profile(const cslam_problem* p, cslam_profile* out);
cslam_status cslam_reset_profile(cslam_problem* p);

/* Run the problem's work on a caller-owned stream (cudaStream_t passed as void*). */
cslam_status cslam_set_stream(cslam_problem* p, void* cuda_stream);

/* Time the materialised residual+Jacobian kernel on device-resident data: `reps` launches,
 * average milliseconds per launch from CUDA events. */
cslam_status cslam_time_resjac(cslam_problem* p, int reps, double* ms_per_launch);
/* Time `reps` launches of the fused Schur-build kernel at the current state. */
cslam_status cslam_time_schur(cslam_problem* p, int reps, double* ms_per_launch);
/* FP64 FMA microbenchmark (register-resident DFMA chains on every SM): TFLOP/s. */
cslam_status cslam_measure_fp64_peak(int device, double* tflops);

/* Host-only structure analysis (no CUDA needed): which blocks are free, this rank's landmark
 * shard, the grouping used by the fused Schur kernel and the pattern of the reduced camera
 * system.  Lets the sharding logic be tested without a GPU. */
typedef struct cslam_structure_info {
    int n_free_cams;
    int n_landmarks;          /* landmarks in this rank's shard */
    long long n_observations; /* observations in this rank's shard */
    int nnz_blocks;           /* upper 6x6 blocks of S (global pattern, identical on all ranks) */
    unsigned long long pattern_hash; /* FNV-1a over (rowptr, col) of that pattern */
    int n_groups;             /* camera-list groups in the shard */
    int n_grouped_landmarks;
    int n_work_items;
    unsigned long long landmark_id_sum; /* sum of the caller's point indices in the shard */
    unsigned long long layout_hash;     /* FNV-1a over the whole internal layout (orders, groups, work items) */
} cslam_structure_info;
cslam_status cslam_analyze(cslam_problem* p, int n_ranks, int rank, cslam_structure_info* out);

/* PointCloudAligner::compute_transformation_and_inliers (src/ceres_slam/point_cloud_aligner.cpp:64-136),
 * the 3-point RANSAC `compute_initial_guess` runs on every consecutive pose pair
 * (dataset_problem.cpp:239-242: 400 iterations, threshold 4; dataset_problem_phong.cpp:323-326:
 * threshold 9), for a BATCH of independent pairs in one launch.  Pair p owns the correspondences
 * [offsets[p], offsets[p+1]) of pts0 / pts1 (3 doubles each: the triangulated points in camera
 * frame k-1 and k, index-aligned).  Every hypothesis aligns three correspondences whose indices are
 * drawn exactly like the reference draws them: std::mt19937 seeded with 42 (:70-72) through
 * std::uniform_int_distribution<uint>(0, n-1), re-drawing duplicates (:82-90).  rng_variant selects
 * the libstdc++ algorithm of that distribution: 0 = scaling + rejection (GCC <= 10, the reference's
 * era), 1 = Lemire (GCC >= 11).  The best hypothesis is the one with the most inliers (squared
 * reprojection distance < thresh, :117-123), the earliest on ties (:127).
 * Out: T12_out 12 doubles per pair [t | R row-major] (T_1_0), inlier_out one byte per correspondence
 * (may be NULL), n_inliers_out per pair (may be NULL).  A pair with fewer than 3 correspondences, or
 * whose hypotheses all have no inlier, returns the identity and no inliers.  intr5 = fu, fv, cu, cv, b. */
cslam_status cslam_ransac_align(int device, uint32_t n_pairs, const uint32_t* offsets, const double* pts0,
                                const double* pts1, const double* intr5, uint32_t num_iters, double thresh,
                                int rng_variant, double* T12_out, uint8_t* inlier_out, uint32_t* n_inliers_out);
/* The index triples of those hypotheses for a cloud of n correspondences (host only; n >= 3). */
cslam_status cslam_ransac_triples(uint32_t n, uint32_t num_iters, int rng_variant, uint32_t* triples);

/* Number of kernels this library has launched in this process (all handles). */
cslam_status cslam_get_launch_count(uint64_t* count);

/* Multi-GPU (config 5): every rank holds all poses and its own shard of landmarks and
 * observations; each Schur build is followed by one all-reduce of [S | rhs | scalars].
 * cslam_solve / cslam_download are then COLLECTIVE calls (every rank of the communicator makes
 * them) and every rank's pose and point arrays receive the complete solution: the landmarks a
 * rank does not own are gathered from their owners before the write-back.
 * The 128-byte id comes from rank 0 and is distributed by the launcher (torch.distributed). */
cslam_status cslam_comm_unique_id(uint8_t id[128]);
cslam_status cslam_attach_comm(cslam_problem* p, int n_ranks, int rank, const uint8_t id[128]);

#ifdef __cplusplus
}
#endif
#endif /* CSLAM_B200_H_ */
