// Kernel launch surface of the cslam_b200 back end.  Every launcher enqueues on `s` and returns.
#pragma once
#include <atomic>

#include "engine.h"

namespace cslam {

// K1 — materialised residual + Jacobian, caller's block order, TMA-staged poses / TMA bulk stores
void launch_resjac(cudaStream_t s, const CameraIntrinsics& cam, long long n, const uint32_t* cam_idx,
                   const uint32_t* pt_idx, const double* u, const double* v, const double* d,
                   const double* W, int W_per_obs, const double* poses, const double* points,
                   const int* cam_free, const int* tile_lo, const int* tile_n, double* r, double* Jc,
                   double* Jp, double* cost);

// initial pass: cost, squared column norms, gradient (unscaled J); camera column norms go to
// the diagonal of Bdiag (stride 36, offset 7q) so one all-reduce covers them
void launch_colnorm(cudaStream_t s, const DevView& v, int lm_lo, int lm_hi, double* cn_p, double* cn_l,
                    double* gp, double* gl, double* scal);
void launch_jacobi_scale(cudaStream_t s, const double* cn, double* sc, long long n, int enabled);
void launch_jacobi_scale_cams(cudaStream_t s, const double* Bdiag, double* sc, int nf, int enabled);
void launch_camonly_eval(cudaStream_t s, const DevView& v, const SunBlockData* suns, int n_sun,
                         const PriorBlockData* priors, int n_prior, int apply_loss, double* r_sun, double* J_sun,
                         double* r_pr, double* J_pr, double* cost);

// K2 — fused residual/Jacobian + Schur elimination, one warp per landmark (any track length)
// skip: optional flags [lm_hi - lm_lo], 1 = the landmark belongs to a slice of the wide-window kernel below
void launch_schur_generic(cudaStream_t s, const DevView& v, int lm_lo, int lm_hi, const uint8_t* skip, LmDiag dg, double* S,
                          double* Bdiag, double* bp, double* gp, double* gl, double* scal);
// K2w — long ragged tracks: slices of ungrouped landmarks whose cameras fit a window of 32 consecutive poses
// [c0, c0 + 32): Z = W chol(V^-1) per observation to Zg (18 doubles per observation, index e - obs0), then per slice the
// 192 x 192 window tile of S in DMMA accumulators, one flush per slice.  wide: flags [lm_hi - lm_lo]
void launch_schur_wide(cudaStream_t s, const DevView& v, int lm_lo, int lm_hi, const uint8_t* wide, int n_slices, const int* sl_lo,
                       const int* sl_hi, const int* sl_c0, long long obs0, double* Zg, LmDiag dg, double* S, double* Bdiag,
                       double* bp, double* gp, double* gl, double* scal);
// K2 (fast path) — landmarks grouped by identical camera list: one CTA per slice of a group,
// a producer warp forms Z = W chol(V)^-T per observation, consumer warps keep the 6x6 pair
// blocks of the slice in registers and flush them once (SYRK-shaped, output stationary)
void launch_schur_grouped(cudaStream_t s, const DevView& v, const GroupView& g, int n_items_small, int n_items_rag, LmDiag dg,
                          double* S, double* Bdiag, double* bp, double* gp, double* gl, double* scal);
// sun-sensor and pose-prior blocks (camera-only): adds to Bdiag, bp, gp and the cost
void launch_camonly_build(cudaStream_t s, const DevView& v, const SunBlockData* suns, int n_sun,
                          const PriorBlockData* priors, int n_prior, double* Bdiag, double* bp, double* gp,
                          double* scal);
// S_aa += U_aa + D_p^2, preconditioner block inverse
void launch_finalize(cudaStream_t s, const DevView& v, LmDiag dg, int preconditioner, double* S, double* Bdiag,
                     double* diag_p, double* Minv, double* scal);

// K3a — block-Jacobi PCG on the block-sparse reduced system
struct PcgBufs {
    const int *rowptr, *col;              // upper block-CSR of S
    const int* ent_ptr;                   // mirrored lists: row b -> packed (a, block) pairs of the
    const int* ent_cb;                    //   stored upper blocks (a, b), a < b
    const double *S, *Minv, *b;
    double *x, *r, *z, *p, *q, *ps;
    int nf;
};
// One cooperative launch runs the whole solve; results: x, and ps[PS_ITERS], ps[PS_FAIL]
// (1 = indefinite / NO_CONVERGENCE, 2 = FAILURE).  r_tol < 0 disables the residual rule.
void launch_pcg_persistent(cudaStream_t s, const PcgBufs& B, double* pbuf2, double* rec3, double q_tol, double r_tol,
                           int min_iters, int max_iters, int reset_period);

// K3b — direct solve of a block-banded reduced system (leaves + separators, kernels_band.cu);
// writes ps[PS_ITERS] = 1 and ps[PS_FAIL] = 2 when a pivot is not positive
void launch_band_solve(cudaStream_t s, const BandView& B, const BandScratch& K, double* ps);
// K3c — CG preconditioned with the banded direct solver (kernels_bandpcg.cu); the loop itself is in engine.cu
void launch_bpc_build(cudaStream_t s, int n, int W, const int* rowptr, const int* col, const double* S1, const double* U1,
                      const double* U2, LmDiag dg, double* Sband);
void launch_bpc_add(cudaStream_t s, long long n, const double* a, double* b);
void launch_bpc_spmv(cudaStream_t s, int n, const int* rowptr, const int* col, const int* ent_ptr, const int* ent_cb,
                     const double* S, const double* p, double* q);
void launch_bpc_xpby(cudaStream_t s, long long n, const double* z, double beta, double* p);
void launch_bpc_update(cudaStream_t s, long long n, double alpha, const double* p, const double* q, double* x, double* r);
// K3d — dense Cholesky of a reduced system that is not a narrow band (kernels_dense.cu); xw: [n_pad] work vector
void launch_dense_solve(cudaStream_t s, const DenseView& V, double* xw, double* ps);
// the factorisation and both substitutions of a dense system the caller has already written into V.A (lower triangle,
// identity padding, rhs as row n_pad); does not reset V.fail and writes no status
void launch_dense_factor(cudaStream_t s, const DenseView& V, double* xw);
// K3e — wide block-banded reduced system: chunks as bordered bands + dense separator system (kernels_wband.cu)
void launch_wband_solve(cudaStream_t s, const WbandView& V, double* ps);

// K4 — Plus on the poses, back-substitution, model cost change, candidate cost
void launch_pose_plus(cudaStream_t s, const DevView& v, const double* yp, double* poses_cand, double* scal2,
                      int count_cams);
void launch_backsub(cudaStream_t s, const DevView& v, int lm_lo, int lm_hi, LmDiag dg, const double* yp,
                    const double* poses_cand, double* points_cand, double* yl, double* scal2);
void launch_camonly_step(cudaStream_t s, const DevView& v, const SunBlockData* suns, int n_sun,
                         const PriorBlockData* priors, int n_prior, const double* yp, const double* poses_cand,
                         double* scal2);
// |x - Plus(x, -g)|_inf over free cameras and this rank's landmarks; also |x|^2
void launch_gradnorm(cudaStream_t s, const DevView& v, int lm_lo, int lm_hi, const double* gp_scaled,
                     const double* gl_scaled, double* scal, int count_cams);

// DOGLEG: the eight sums of DoglegStrategy (see kernels.cu), the combined step, the candidate landmarks
enum DoglegSum { DG_G11 = 0, DG_G12, DG_G22, DG_JGG, DG_JGY, DG_JYY, DG_JGR, DG_JYR, DG_COUNT };
void launch_dogleg_products(cudaStream_t s, const DevView& v, int lm_lo, int lm_hi, LmDiag dg, const SunBlockData* suns,
                            int n_sun, const PriorBlockData* priors, int n_prior, const double* gp, const double* diag_p,
                            const double* yp, const double* gl, const double* yl, double* diag_l, double* sums, int count_cams);
void launch_dogleg_combine(cudaStream_t s, long long n, double c1, double c2, const double* g, const double* d2, const double* y,
                           double* out);
void launch_points_apply(cudaStream_t s, const DevView& v, int lm_lo, int lm_hi, const double* Yl, const double* poses_cand,
                         double* points_cand, double* scal2);

// layout: internal (landmark-major / slot-major) observation and landmark arrays gathered from the
// caller-order arrays on the device; the reverse for the landmarks at download
void launch_gather_layout(cudaStream_t s, long long n_obs, const uint32_t* obs_user, const uint32_t* raw_cam,
                          const double* raw_uvd, const double* raw_W, int W_per_obs, uint32_t* obs_cam, double* u,
                          double* v, double* d, double* W, int n_lm, const uint32_t* lm_user, const double* raw_pts,
                          double* points);
void launch_scatter_points(cudaStream_t s, int n_lm, const uint32_t* lm_user, const double* points, double* raw_pts);
void launch_scatter_points4(cudaStream_t s, int n_lm, const uint32_t* lm_user, const double* points, double* z4);
void launch_merge_points4(cudaStream_t s, long long n_points, const double* z4, double* raw_pts);
void launch_fill(cudaStream_t s, double* p, size_t n, double value);

// K1-phong — materialised residuals / Jacobians of the intensity and normal blocks
void launch_phong_eval(cudaStream_t s, const PhongView& v, double* r_int, double* J_int, double* r_normal,
                       double* Jpose_normal, double* Jn_normal, double* cost);

// K2p / K4p — joint lighting solve (kernels_phong_solve.cu)
// K2p for grouped vertices (kernels_phong_grouped.cu): items [0, n_items_small) have L <= 10, the rest L <= 16
void launch_phong_build_grouped(cudaStream_t s, const DevView& v, const PhongSolveView& q, const GroupView& g, int n_items_small,
                                LmDiag dg, const PhongSystem& o);
// vertices observed more than 32 times (kernels_phong_long.cu): called by the launchers below when max_track_len > 32
void launch_phong_build_long(cudaStream_t s, const DevView& v, const PhongSolveView& q, int lm_lo, int lm_hi, LmDiag dg,
                             const PhongSystem& o, bool schur);
void launch_phong_backsub_long(cudaStream_t s, const DevView& v, const PhongSolveView& q, int lm_lo, int lm_hi, LmDiag dg,
                               const double* yp, const double* yg, const double* gv, double* yv, double* scal2);
void launch_phong_dogleg_products_long(cudaStream_t s, const DevView& v, const PhongSolveView& q, int lm_lo, int lm_hi, LmDiag dg,
                                       const double* gp, const double* diag_p, const double* yp, const double* gg,
                                       const double* diag_g, const double* yg, const double* gv, const double* yv, double* diag_v,
                                       double* sc_v, double* sums);
void launch_phong_candidate_long(cudaStream_t s, const DevView& v, const PhongSolveView& q, int lm_lo, int lm_hi, double alpha,
                                 const double* yv, const double* poses_cand, const double* gx_cand, double* points_cand,
                                 double* normals_cand, double* scal2);
void launch_phong_build(cudaStream_t s, const DevView& v, const PhongSolveView& q, int lm_lo, int lm_hi, LmDiag dg,
                        const PhongSystem& o, bool schur, int max_track_len);
void launch_phong_gfinalize(cudaStream_t s, const PhongSolveView& q, LmDiag dg, double* Sgg, double* bg, const double* hg,
                            double* diag_g);
// border elimination: X holds n_g + 1 solves against S_cc (border columns, then b_c)
void launch_phong_border_solve(cudaStream_t s, int n_g, int nf6, const double* Scg, const double* X, const double* Sgg,
                               const double* bg, double* T, double* yg, double* yc, double* ps);
void launch_phong_backsub(cudaStream_t s, const DevView& v, const PhongSolveView& q, int lm_lo, int lm_hi, LmDiag dg,
                          const double* yp, const double* yg, const double* gv, double* yv, double* scal2, int max_track_len);
// DOGLEG for the lighting solve: the strategy's eight sums (vertex and observation terms + shared-block terms;
// the pose terms come from launch_dogleg_products with an empty landmark range), diag_v / sc_v as by-products
void launch_phong_dogleg_products(cudaStream_t s, const DevView& v, const PhongSolveView& q, int lm_lo, int lm_hi, LmDiag dg,
                                  const double* gp, const double* diag_p, const double* yp, const double* gg, const double* diag_g,
                                  const double* yg, const double* gv, const double* yv, double* diag_v, double* sc_v, double* sums,
                                  int max_track_len, int count_shared);
void launch_phong_candidate(cudaStream_t s, const DevView& v, const PhongSolveView& q, int lm_lo, int lm_hi, double alpha,
                            const double* yp, const double* yg, const double* yv, double* poses_cand, double* gx_cand,
                            double* points_cand, double* normals_cand, double* scal2, int count_shared, int max_track_len);
void launch_phong_gradnorm(cudaStream_t s, const DevView& v, const PhongSolveView& q, int lm_lo, int lm_hi, const double* gv,
                           const double* gg, double* scal, int count_shared);
void launch_phong_project_initial(cudaStream_t s, const PhongSolveView& q, int n_lm, double* normals, double* gx_out,
                                  double* zero_g);
void launch_dot(cudaStream_t s, const double* a, const double* b, long long n, double* dst);
void launch_absmax_scaled(cudaStream_t s, const double* y, const double* sc, long long n, double* dst);

// KR — batched 3-point RANSAC point-cloud alignment (ransac.cu); host arrays in and out
void ransac_align_batch(int device, uint32_t n_pairs, const uint32_t* offsets, const double* pts0, const double* pts1,
                        const double* intr5, uint32_t num_iters, double thresh, int rng_variant, double* T12_out,
                        uint8_t* inlier_out, uint32_t* n_inliers_out);
void ransac_triples(uint32_t n, uint32_t num_iters, int variant, uint32_t* out);

// structure analysis on the device (structure.cu): see Engine::build_structure_gpu
void launch_st_count(cudaStream_t s, size_t n, const uint32_t* cam, const uint32_t* pt, uint32_t n_poses, uint32_t n_points,
                     uint32_t* cnt, uint8_t* used, int* flags);
void launch_st_scan(cudaStream_t s, uint32_t n_points, const uint32_t* cnt, uint32_t* ptr, DBuf<uint8_t>& tmp);
void launch_st_fill(cudaStream_t s, size_t n, const uint32_t* cam, const uint32_t* pt, uint32_t* fill, unsigned long long* ck);
void launch_st_landmarks(cudaStream_t s, uint32_t n_points, const uint32_t* cnt, const uint32_t* ptr, unsigned long long* ck,
                         const int* cam_free, int group_lmax, int allow_groups, uint32_t* mincam, unsigned long long* khash,
                         uint8_t* kok, unsigned long long* mask, int2* far, int far_cap, int* flags);
void launch_st_group_cams(cudaStream_t s, int n_groups, const uint32_t* g_first_user, const int* g_off, const int* g_L,
                          const uint32_t* ptr, const unsigned long long* ck, int* g_cams);
void launch_st_perm(cudaStream_t s, int n_lm, const uint32_t* lm_user, const uint32_t* lm_base, const uint32_t* lm_stride,
                    const uint32_t* lm_cnt, const uint32_t* ptr, const unsigned long long* ck, uint32_t* obs_user);
void launch_st_order(cudaStream_t s, uint32_t n_points, uint32_t n_poses, const uint32_t* cnt, const uint32_t* mincam,
                     uint32_t* key_tmp, uint32_t* val_tmp, uint32_t* mincam_s, uint32_t* all_lm, uint32_t* n_active_out,
                     DBuf<uint8_t>& tmp);
void launch_st_group_sort(cudaStream_t s, uint32_t n_points, uint32_t lo, uint32_t hi, const uint32_t* all_lm,
                          const uint32_t* mincam_s, const uint32_t* cnt, const uint8_t* kok, const unsigned long long* khash,
                          uint32_t* len_a, uint32_t* all_ptr, uint32_t* key2, unsigned long long* keyh,
                          unsigned long long* keyh_tmp, uint32_t* val_a, uint32_t* val_b, uint32_t* key2_b, uint32_t* key2_s,
                          uint32_t* sorted_a, uint8_t* run_flag, uint32_t* run_pos, uint32_t* counts, DBuf<uint8_t>& tmp);
void launch_st_max_track(cudaStream_t s, uint32_t n_points, const uint32_t* cnt, uint32_t* out, DBuf<uint8_t>& tmp);
void launch_st_run_len(cudaStream_t s, uint32_t n_runs, const uint32_t* run_pos, const uint32_t* key2_s, uint32_t* run_L);
void launch_st_layout(cudaStream_t s, uint32_t n_points, uint32_t lo, uint32_t hi, int n_groups, const uint32_t* g_x,
                      const int* g_G, const int* g_L, const int* g_lm0, const uint32_t* g_obs0, const int* g_off,
                      const uint32_t* sorted_a, const uint32_t* all_lm, const uint32_t* len_a, const uint32_t* ptr,
                      const unsigned long long* ck, uint32_t n_lm_grouped, uint32_t obs_cursor, uint32_t* lm_user,
                      uint32_t* lm_base, uint32_t* lm_stride, uint32_t* lm_cnt, uint8_t* grouped, int* g_cams, uint32_t* rflag,
                      uint32_t* rlen, uint32_t* ridx, uint32_t* roff, DBuf<uint8_t>& tmp);
void launch_st_verify(cudaStream_t s, int n_groups, const int* g_L, const int* g_G, const int* g_lm0, const int* g_off,
                      const int* g_cams, const uint32_t* lm_user, const uint32_t* ptr, const unsigned long long* ck, int* flags);

double measure_fp64_peak_tflops(int device);
extern std::atomic<unsigned long long> g_kernel_launches;  // every kernel this library launches

}  // namespace cslam
