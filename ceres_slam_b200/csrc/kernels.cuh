// Kernel launch surface of the cslam_b200 back end.  Every launcher enqueues on `s` and returns.
#pragma once
#include <atomic>

#include "engine.h"

namespace cslam {

struct LmDiag {
    double inv_radius, min_diag, max_diag;
};

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
void launch_schur_generic(cudaStream_t s, const DevView& v, int lm_lo, int lm_hi, LmDiag dg, double* S,
                          double* Bdiag, double* bp, double* gp, double* gl, double* scal);
// sun-sensor and pose-prior blocks (camera-only): adds to Bdiag, bp, gp and the cost
void launch_camonly_build(cudaStream_t s, const DevView& v, const SunBlockData* suns, int n_sun,
                          const PriorBlockData* priors, int n_prior, double* Bdiag, double* bp, double* gp,
                          double* scal);
// S_aa += U_aa + D_p^2, preconditioner block inverse
void launch_finalize(cudaStream_t s, const DevView& v, LmDiag dg, int preconditioner, double* S, double* Bdiag,
                     double* diag_p, double* Minv, double* scal);

// K3a — block-Jacobi PCG on the block-sparse reduced system
struct PcgBufs {
    const int *rowptr, *col, *lt_rowptr, *lt_col, *lt_blk;
    const double *S, *Minv, *b;
    double *x, *r, *z, *p, *q, *ps;
    int nf;
};
void launch_pcg_init(cudaStream_t s, const PcgBufs& B);
void launch_pcg_iteration(cudaStream_t s, const PcgBufs& B, int iteration, double q_tol, double r_tol2,
                          int min_iters, int max_iters, int reset_period);

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

double measure_fp64_peak_tflops(int device);
extern std::atomic<unsigned long long> g_kernel_launches;  // every kernel this library launches

}  // namespace cslam
