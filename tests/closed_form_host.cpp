// Test-only host build of ceres_slam_b200/csrc/closed_form.h: lets the CPU test-suite check the
// device formulas against the Jet oracle without a GPU.  Never loaded by the product.
#include "../ceres_slam_b200/csrc/closed_form.h"
using namespace cslam;
extern "C" {
void cf_stereo_block(const double* intr, const double* pose, const double* p, const double* uvd,
                     const double* W, double* r, double* Jc, double* Jp) {
    CameraIntrinsics c{intr[0], intr[1], intr[2], intr[3], intr[4]};
    stereo_block<true>(c, pose, p, uvd[0], uvd[1], uvd[2], W, r, Jc, Jp);
}
void cf_sun_block(const double* pose, const double* obs_c, const double* ref_g, const double* W2,
                  double az, double zen, double* r, double* J) {
    sun_block(pose, obs_c, ref_g, W2, az, zen, r, J);
}
void cf_prior_block(const double* pose, const double* Tref, const double* W6, double* r, double* J) {
    prior_block(pose, Tref, W6, r, J);
}
void cf_se3_plus(const double* pose, const double* eps, double* out) { se3_plus(pose, eps, out); }
void cf_so3_log(const double* C, double* phi) { so3_log(C, phi); }
void cf_unit_plus(const double* x, const double* dl, double* out) { unit_plus(x, dl, out); }
void cf_normal_block(const double* pose, const double* n, const double* obs, const double* W, double* r,
                     double* Jc, double* Jn) {
    normal_block(pose, n, obs, W, r, Jc, Jn);
}
void cf_intensity_block(const double* pose, const double* p, const double* n, const double* phong,
                        const double* tex, const double* light, double colour, double w, int directional,
                        double* r, double* Jc, double* Jp, double* Jn, double* Jk, double* Jt, double* Jl) {
    intensity_block(pose, p, n, phong, tex[0], light, colour, w, directional != 0, r, Jc, Jp, Jn, Jk, Jt, Jl);
}
}
