// ORACLE — TEST INFRASTRUCTURE ONLY (residual blocks PINNED via oracle/_ref; the solver rules below are Ceres' and
// Ceres is absent: that part is parity UNPINNED).
//
// CPU restatement of what the reference's `solveWindow` hands to Ceres and what Ceres then does
// with it: problem assembly (tests/dataset_vo.cpp:22-85, tests/dataset_vo_sun.cpp:25-187),
// Jet-autodiff evaluation chained with the local-parameterisation Jacobian, Schur elimination of
// the point blocks and a trust-region Levenberg-Marquardt loop.  Ceres is an un-vendored,
// un-pinned dependency (CMakeLists.txt:17); the solver rules are restated from the published
// Ceres 1.x algorithm (trust_region_minimizer, levenberg_marquardt_strategy,
// trust_region_step_evaluator, conjugate_gradients_solver, corrector) — SURVEY.md App. B.
// Nothing here is shipped or measured as the product; it is the checker and the CPU baseline.
#pragma once
#include <algorithm>
#include <atomic>
#include <chrono>
#include <complex>
#include <cstdint>
#include <cstring>
#include <functional>
#include <string>
#include <thread>
#include <vector>

#include "functors.hpp"

namespace oracle {

struct Options {
    int max_num_iterations = 1000;                  // dataset_vo.cpp:69
    int use_nonmonotonic_steps = 1;                 // dataset_vo.cpp:70
    int max_consecutive_nonmonotonic_steps = 5;     // Ceres default
    double initial_trust_region_radius = 1e4;       // Ceres defaults below
    double max_trust_region_radius = 1e16;
    double min_trust_region_radius = 1e-32;
    double min_relative_decrease = 1e-3;
    double min_lm_diagonal = 1e-6;
    double max_lm_diagonal = 1e32;
    int max_num_consecutive_invalid_steps = 5;
    double function_tolerance = 1e-6;
    double gradient_tolerance = 1e-10;
    double parameter_tolerance = 1e-8;
    int jacobi_scaling = 1;
    int linear_solver = 0;      // 0 exact Schur (SPARSE_SCHUR-equivalent), 1 ITERATIVE_SCHUR (PCG)
    int preconditioner = 1;     // 0 JACOBI (diag of B), 1 SCHUR_JACOBI (diag of S)
    double eta = 0.1;
    int max_linear_solver_iterations = 500;
    int min_linear_solver_iterations = 0;
    int num_threads = 8;        // dataset_vo.cpp:67
    double line_search_sufficient_function_decrease = 1e-4;  // Ceres default (bounded problems)
    int trust_region_strategy = 0;  // 0 LEVENBERG_MARQUARDT, 1 DOGLEG (dataset_vo_sun.cpp:142)
    int dogleg_type = 1;            // 0 TRADITIONAL_DOGLEG, 1 SUBSPACE_DOGLEG (dataset_vo_sun.cpp:143)
};

enum Termination { CONVERGENCE = 0, NO_CONVERGENCE = 1, FAILURE = 2 };
enum Reason {
    R_NONE = 0,
    R_GRADIENT_TOL = 1,
    R_PARAMETER_TOL = 2,
    R_FUNCTION_TOL = 3,
    R_MAX_ITERATIONS = 4,
    R_MIN_RADIUS = 5,
    R_INVALID_STEPS = 6,
    R_LINEAR_SOLVER = 7,
    R_INITIAL_EVAL = 8
};

struct IterationRow {
    double iteration, cost, cost_change, gradient_max_norm, step_norm, relative_decrease, radius,
        linear_iterations, step_is_valid, step_is_successful;
};

struct Summary {
    double initial_cost = 0, final_cost = 0;
    int num_iterations = 0, num_successful_steps = 0, num_unsuccessful_steps = 0;
    int termination_type = NO_CONVERGENCE, termination_reason = R_NONE;
    double final_radius = 0;
    int total_linear_iterations = 0;
    std::vector<IterationRow> rows;
    std::vector<double> row_seconds;  // wall clock (steady) when each row was logged: bench.py's CPU arm times
                                      // K iterations after W warm-up iterations inside ONE solve
    void push_row(const IterationRow& r) {
        rows.push_back(r);
        row_seconds.push_back(std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count());
    }
};

inline void parallel_for(size_t n, int nthreads, const std::function<void(size_t, size_t, int)>& fn) {
    if (nthreads <= 1 || n < 2048) {
        fn(0, n, 0);
        return;
    }
    std::vector<std::thread> th;
    size_t chunk = (n + nthreads - 1) / nthreads;
    for (int t = 0; t < nthreads; ++t) {
        size_t b = std::min(n, chunk * t), e = std::min(n, chunk * (t + 1));
        if (b >= e) break;
        th.emplace_back(fn, b, e, t);
    }
    for (auto& x : th) x.join();
}
inline void atomic_add(double& dst, double v) {
    std::atomic_ref<double> r(dst);
    r.fetch_add(v, std::memory_order_relaxed);
}

// Symmetric block matrix, upper block-row storage, 6x6 blocks.
struct BlockSym {
    int n = 0;
    std::vector<int> rowptr, col;
    std::vector<double> val;
    int find(int a, int b) const {  // a <= b
        auto lo = col.begin() + rowptr[a], hi = col.begin() + rowptr[a + 1];
        auto it = std::lower_bound(lo, hi, b);
        return (it != hi && *it == b) ? int(it - col.begin()) : -1;
    }
    void multiply(const double* x, double* y) const {
        std::fill(y, y + 6 * n, 0.0);
        for (int a = 0; a < n; ++a)
            for (int e = rowptr[a]; e < rowptr[a + 1]; ++e) {
                const int b = col[e];
                const double* B = &val[36 * size_t(e)];
                for (int i = 0; i < 6; ++i)
                    for (int j = 0; j < 6; ++j) y[6 * a + i] += B[6 * i + j] * x[6 * b + j];
                if (b != a)
                    for (int i = 0; i < 6; ++i)
                        for (int j = 0; j < 6; ++j) y[6 * b + j] += B[6 * i + j] * x[6 * a + i];
            }
    }
};

// In-place Cholesky of a dense symmetric n x n (row-major, lower used). Returns false if not PD.
inline bool dense_cholesky(double* A, int n) {
    for (int j = 0; j < n; ++j) {
        double d = A[size_t(j) * n + j];
        for (int k = 0; k < j; ++k) d -= A[size_t(j) * n + k] * A[size_t(j) * n + k];
        if (!(d > 0.0)) return false;
        d = std::sqrt(d);
        A[size_t(j) * n + j] = d;
        for (int i = j + 1; i < n; ++i) {
            double s = A[size_t(i) * n + j];
            const double* ri = A + size_t(i) * n;
            const double* rj = A + size_t(j) * n;
            for (int k = 0; k < j; ++k) s -= ri[k] * rj[k];
            A[size_t(i) * n + j] = s / d;
        }
    }
    return true;
}
inline void dense_cholesky_solve(const double* L, int n, double* b) {
    for (int i = 0; i < n; ++i) {
        double s = b[i];
        for (int k = 0; k < i; ++k) s -= L[size_t(i) * n + k] * b[k];
        b[i] = s / L[size_t(i) * n + i];
    }
    for (int i = n - 1; i >= 0; --i) {
        double s = b[i];
        for (int k = i + 1; k < n; ++k) s -= L[size_t(k) * n + i] * b[k];
        b[i] = s / L[size_t(i) * n + i];
    }
}
// Banded Cholesky: lower band storage B[i*(w+1) + (j - i + w)] for i-w <= j <= i.
inline bool band_cholesky(std::vector<double>& B, int n, int w) {
    const int ld = w + 1;
    auto at = [&](int i, int j) -> double& { return B[size_t(i) * ld + (j - i + w)]; };
    for (int j = 0; j < n; ++j) {
        double d = at(j, j);
        for (int k = std::max(0, j - w); k < j; ++k) d -= at(j, k) * at(j, k);
        if (!(d > 0.0)) return false;
        d = std::sqrt(d);
        at(j, j) = d;
        for (int i = j + 1; i <= std::min(n - 1, j + w); ++i) {
            double s = at(i, j);
            for (int k = std::max(0, i - w); k < j; ++k) s -= at(i, k) * at(j, k);
            at(i, j) = s / d;
        }
    }
    return true;
}
inline void band_cholesky_solve(const std::vector<double>& B, int n, int w, double* b) {
    const int ld = w + 1;
    auto at = [&](int i, int j) -> double { return B[size_t(i) * ld + (j - i + w)]; };
    for (int i = 0; i < n; ++i) {
        double s = b[i];
        for (int k = std::max(0, i - w); k < i; ++k) s -= at(i, k) * b[k];
        b[i] = s / at(i, i);
    }
    for (int i = n - 1; i >= 0; --i) {
        double s = b[i];
        for (int k = i + 1; k <= std::min(n - 1, i + w); ++k) s -= at(k, i) * b[k];
        b[i] = s / at(i, i);
    }
}
inline bool invert_spd6(const double* A, double* Ainv) {
    double L[36];
    std::memcpy(L, A, sizeof(L));
    if (!dense_cholesky(L, 6)) return false;
    for (int c = 0; c < 6; ++c) {
        double e[6] = {0, 0, 0, 0, 0, 0};
        e[c] = 1.0;
        dense_cholesky_solve(L, 6, e);
        for (int r = 0; r < 6; ++r) Ainv[6 * r + c] = e[r];
    }
    return true;
}
inline bool invert_spd3(const double* V, double* Vi) {
    // closed-form adjugate inverse of a symmetric 3x3 (row-major, full storage)
    const double a = V[0], b = V[1], c = V[2], d = V[4], e = V[5], f = V[8];
    const double c00 = d * f - e * e, c01 = c * e - b * f, c02 = b * e - c * d;
    const double det = a * c00 + b * c01 + c * c02;
    if (!(det > 0.0) || !std::isfinite(det)) return false;
    const double id = 1.0 / det;
    Vi[0] = c00 * id;
    Vi[1] = Vi[3] = c01 * id;
    Vi[2] = Vi[6] = c02 * id;
    Vi[4] = (a * f - c * c) * id;
    Vi[5] = Vi[7] = (b * c - a * e) * id;
    Vi[8] = (a * d - b * b) * id;
    return true;
}

// Real parts of the roots of c[0] y^n + ... + c[n] (Ceres FindPolynomialRoots returns the
// eigenvalues of the companion matrix; here Durand-Kerner in complex arithmetic — same roots).
inline std::vector<double> polynomial_root_real_parts(std::vector<double> c) {
    while (!c.empty() && c.front() == 0.0) c.erase(c.begin());
    const int n = int(c.size()) - 1;
    std::vector<double> out;
    if (n < 1) return out;
    using cd = std::complex<double>;
    std::vector<cd> a(n + 1), z(n);
    for (int i = 0; i <= n; ++i) a[i] = c[i] / c[0];
    double rad = 0;
    for (int i = 1; i <= n; ++i) rad = std::max(rad, std::pow(std::abs(a[i]), 1.0 / i));
    rad = 2.0 * rad + 1e-300;
    for (int i = 0; i < n; ++i) z[i] = std::polar(rad, 2.0 * 3.14159265358979323846 * i / n + 0.4);
    for (int it = 0; it < 500; ++it) {
        double delta = 0;
        for (int i = 0; i < n; ++i) {
            cd pv = a[0];
            for (int k = 1; k <= n; ++k) pv = pv * z[i] + a[k];
            cd den = 1.0;
            for (int j = 0; j < n; ++j)
                if (j != i) den *= (z[i] - z[j]);
            if (std::abs(den) == 0.0) den = 1e-300;
            const cd dz = pv / den;
            z[i] -= dz;
            delta = std::max(delta, std::abs(dz) / std::max(std::abs(z[i]), 1e-300));
        }
        if (delta < 1e-15) break;
    }
    for (int i = 0; i < n; ++i) out.push_back(z[i].real());
    return out;
}

// DoglegStrategy::FindMinimumOnTrustRegionBoundary [Ceres 1.x dogleg_strategy.cc, from memory]:
// stationary points of 1/2 x^T B x + g^T x on |x| = r through the quartic in the multiplier.
inline bool dogleg_boundary_minimum(const double B[4], const double g[2], double r, double x_out[2]) {
    const double detB = B[0] * B[3] - B[1] * B[2], trB = B[0] + B[3], r2 = r * r;
    const double Ba[4] = {B[3], -B[1], -B[2], B[0]};  // adjugate
    const double gg = g[0] * g[0] + g[1] * g[1];
    const double Bag[2] = {Ba[0] * g[0] + Ba[1] * g[1], Ba[2] * g[0] + Ba[3] * g[1]};
    std::vector<double> poly(5);
    poly[0] = r2;
    poly[1] = 2.0 * r2 * trB;
    poly[2] = r2 * (trB * trB + 2.0 * detB) - gg;
    poly[3] = -2.0 * ((g[0] * Bag[0] + g[1] * Bag[1]) - r2 * detB * trB);
    poly[4] = r2 * detB * detB - (Bag[0] * Bag[0] + Bag[1] * Bag[1]);
    const std::vector<double> roots = polynomial_root_real_parts(poly);
    x_out[0] = x_out[1] = 0.0;
    double best = std::numeric_limits<double>::max();
    bool found = false;
    for (double y : roots) {
        // x(y) = -(B + y I)^-1 g
        const double a = B[0] + y, b = B[1], c2 = B[2], d = B[3] + y;
        const double det = a * d - b * c2;
        if (det == 0.0 || !std::isfinite(det)) continue;
        const double x[2] = {-(d * g[0] - b * g[1]) / det, -(-c2 * g[0] + a * g[1]) / det};
        const double nx = std::sqrt(x[0] * x[0] + x[1] * x[1]);
        if (!(nx > 0.0) || !std::isfinite(nx)) continue;
        const double p[2] = {r / nx * x[0], r / nx * x[1]};
        const double f = 0.5 * (p[0] * (B[0] * p[0] + B[1] * p[1]) + p[1] * (B[2] * p[0] + B[3] * p[1])) + g[0] * p[0] + g[1] * p[1];
        found = true;
        if (f < best) {
            best = f;
            x_out[0] = x[0];
            x_out[1] = x[1];
        }
    }
    return found;
}

class Problem {
   public:
    Camera camera{1, 1, 0, 0, 1};
    double* poses = nullptr;  // n_poses x 12, in-out (user memory)
    int n_poses = 0;
    std::vector<uint8_t> pose_const;
    double* points = nullptr;  // n_points x 3, in-out
    int n_points = 0;

    // stereo blocks — dataset_vo.cpp:46-53
    std::vector<uint32_t> st_cam, st_pt;
    std::vector<double> st_uvd;  // 3 per obs
    std::vector<double> st_W;    // 9 shared or 9 per obs (dataset_vo_sun.cpp:57-59)
    bool st_W_per_obs = false;
    // sun blocks — dataset_vo_sun.cpp:75-101
    struct Sun {
        uint32_t cam;
        SunSensorError f;
        double huber;  // 0 = no loss
    };
    std::vector<Sun> suns;
    // pose priors — dataset_vo_sun.cpp:109-124
    struct Prior {
        uint32_t cam;
        PoseError f;
    };
    std::vector<Prior> priors;

    std::string error;

    size_t n_stereo() const { return st_cam.size(); }
    const double* W_of(size_t i) const { return st_W.data() + (st_W_per_obs ? 9 * i : 0); }

    // ----------------------------------------------------------------------------------------
    // Evaluation (Jet autodiff chained with the Plus Jacobian, like ceres::Problem::Evaluate)
    // ----------------------------------------------------------------------------------------
    struct Eval {
        double cost = 0;
        std::vector<double> r_st, Jc_st, Jp_st;  // 3, 18 (3x6), 9 (3x3) per stereo block
        std::vector<double> r_sun, J_sun;        // 2, 12 (2x6)
        std::vector<double> r_pr, J_pr;          // 6, 36
    };

    StereoReprojectionError make_stereo(size_t i) const {
        StereoReprojectionError f;
        f.camera = camera;
        for (int k = 0; k < 3; ++k) f.observation[k] = st_uvd[3 * i + k];
        std::memcpy(f.stiffness, W_of(i), 9 * sizeof(double));
        return f;
    }

    // Huber — ceres::HuberLoss::Evaluate + ceres Corrector (rho'' <= 0 branch)
    static void huber(double a, double s, double rho[3]) {
        const double b = a * a;
        if (s > b) {
            const double r = std::sqrt(s);
            rho[0] = 2.0 * a * r - b;
            rho[1] = std::max(std::numeric_limits<double>::min(), a / r);
            rho[2] = -rho[1] / (2.0 * s);
        } else {
            rho[0] = s;
            rho[1] = 1.0;
            rho[2] = 0.0;
        }
    }

    // x_poses / x_points: state to evaluate at.  jac=false → cost only (plain doubles).
    bool evaluate(const double* x_poses, const double* x_points, bool jac, bool apply_loss,
                  Eval& ev, int nthreads) const {
        const size_t ns = n_stereo();
        std::vector<double> plusjac;
        if (jac) {
            plusjac.resize(size_t(n_poses) * 72);
            SE3Perturbation plus;
            for (int k = 0; k < n_poses; ++k)
                autodiff_plus_jacobian(plus, x_poses + 12 * k, &plusjac[72 * size_t(k)]);
            ev.r_st.assign(3 * ns, 0.0);
            ev.Jc_st.assign(18 * ns, 0.0);
            ev.Jp_st.assign(9 * ns, 0.0);
            ev.r_sun.assign(2 * suns.size(), 0.0);
            ev.J_sun.assign(12 * suns.size(), 0.0);
            ev.r_pr.assign(6 * priors.size(), 0.0);
            ev.J_pr.assign(36 * priors.size(), 0.0);
        }
        std::vector<double> partial(std::max(1, nthreads), 0.0);
        std::atomic<bool> ok{true};
        parallel_for(ns, nthreads, [&](size_t b, size_t e, int tid) {
            double c = 0;
            for (size_t i = b; i < e; ++i) {
                StereoReprojectionError f = make_stereo(i);
                const double* params[2] = {x_poses + 12 * size_t(st_cam[i]),
                                           x_points + 3 * size_t(st_pt[i])};
                double r[3];
                if (jac) {
                    double Ja[36], Jp[9];
                    double* jp[2] = {Ja, Jp};
                    if (!autodiff_cost(f, params, r, jp)) ok = false;
                    const double* P = &plusjac[72 * size_t(st_cam[i])];
                    double* Jc = &ev.Jc_st[18 * i];
                    for (int rr = 0; rr < 3; ++rr)
                        for (int cc = 0; cc < 6; ++cc) {
                            double s = 0;
                            for (int k = 0; k < 12; ++k) s += Ja[12 * rr + k] * P[6 * k + cc];
                            Jc[6 * rr + cc] = s;
                        }
                    std::memcpy(&ev.Jp_st[9 * i], Jp, sizeof(Jp));
                    std::memcpy(&ev.r_st[3 * i], r, sizeof(r));
                } else {
                    if (!eval_cost(f, params, r)) ok = false;
                }
                c += 0.5 * (r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
            }
            partial[tid] += c;
        });
        double cost = 0;
        for (double c : partial) cost += c;
        for (size_t i = 0; i < suns.size(); ++i) {
            const Sun& s = suns[i];
            const double* params[1] = {x_poses + 12 * size_t(s.cam)};
            double r[2], Ja[24];
            double* jp[1] = {Ja};
            if (jac)
                autodiff_cost(s.f, params, r, jp);
            else
                eval_cost(s.f, params, r);
            const double sq = r[0] * r[0] + r[1] * r[1];
            double rho[3] = {sq, 1.0, 0.0};
            if (s.huber > 0.0 && apply_loss) huber(s.huber, sq, rho);
            cost += 0.5 * rho[0];
            if (jac) {
                const double sc = std::sqrt(rho[1]);
                const double* P = &plusjac[72 * size_t(s.cam)];
                for (int rr = 0; rr < 2; ++rr) {
                    for (int cc = 0; cc < 6; ++cc) {
                        double a = 0;
                        for (int k = 0; k < 12; ++k) a += Ja[12 * rr + k] * P[6 * k + cc];
                        ev.J_sun[12 * i + 6 * rr + cc] = sc * a;
                    }
                    ev.r_sun[2 * i + rr] = sc * r[rr];
                }
            }
        }
        for (size_t i = 0; i < priors.size(); ++i) {
            const Prior& p = priors[i];
            const double* params[1] = {x_poses + 12 * size_t(p.cam)};
            double r[6], Ja[72];
            double* jp[1] = {Ja};
            if (jac)
                autodiff_cost(p.f, params, r, jp);
            else
                eval_cost(p.f, params, r);
            for (int k = 0; k < 6; ++k) cost += 0.5 * r[k] * r[k];
            if (jac) {
                const double* P = &plusjac[72 * size_t(p.cam)];
                for (int rr = 0; rr < 6; ++rr) {
                    for (int cc = 0; cc < 6; ++cc) {
                        double a = 0;
                        for (int k = 0; k < 12; ++k) a += Ja[12 * rr + k] * P[6 * k + cc];
                        ev.J_pr[36 * i + 6 * rr + cc] = a;
                    }
                    ev.r_pr[6 * i + rr] = r[rr];
                }
            }
        }
        ev.cost = cost;
        return ok && std::isfinite(cost);
    }

    // ----------------------------------------------------------------------------------------
    // Solve: trust-region LM with Schur elimination of the point blocks
    // ----------------------------------------------------------------------------------------
    struct Structure {
        std::vector<int> cam_free;   // pose -> free index or -1 (constant or unused)
        std::vector<int> free_cams;  // free index -> pose
        std::vector<int> pt_active;  // point -> active index or -1
        std::vector<int> active_pts;
        std::vector<size_t> pt_ptr;  // CSR over active points -> stereo obs
        std::vector<uint32_t> pt_obs;
        BlockSym S;
    };

    void build_structure(Structure& st) const {
        st.cam_free.assign(n_poses, -1);
        std::vector<uint8_t> used(n_poses, 0);
        for (uint32_t c : st_cam) used[c] = 1;
        for (auto& s : suns) used[s.cam] = 1;
        for (auto& p : priors) used[p.cam] = 1;
        for (int k = 0; k < n_poses; ++k)
            if (used[k] && !pose_const[k]) {
                st.cam_free[k] = int(st.free_cams.size());
                st.free_cams.push_back(k);
            }
        st.pt_active.assign(n_points, -1);
        std::vector<size_t> cnt(n_points, 0);
        for (uint32_t j : st_pt) cnt[j]++;
        for (int j = 0; j < n_points; ++j)
            if (cnt[j]) {
                st.pt_active[j] = int(st.active_pts.size());
                st.active_pts.push_back(j);
            }
        const size_t na = st.active_pts.size();
        st.pt_ptr.assign(na + 1, 0);
        for (size_t a = 0; a < na; ++a) st.pt_ptr[a + 1] = st.pt_ptr[a] + cnt[st.active_pts[a]];
        st.pt_obs.resize(n_stereo());
        std::vector<size_t> fill(st.pt_ptr.begin(), st.pt_ptr.end() - 1);
        for (size_t i = 0; i < n_stereo(); ++i) st.pt_obs[fill[st.pt_active[st_pt[i]]]++] = uint32_t(i);
        // co-visibility pattern of the reduced camera system
        const int nf = int(st.free_cams.size());
        std::vector<std::vector<int>> rows(nf);
        for (int a = 0; a < nf; ++a) rows[a].push_back(a);
        std::vector<int> fc;
        for (size_t a = 0; a < na; ++a) {
            fc.clear();
            for (size_t e = st.pt_ptr[a]; e < st.pt_ptr[a + 1]; ++e) {
                int f = st.cam_free[st_cam[st.pt_obs[e]]];
                if (f >= 0) fc.push_back(f);
            }
            std::sort(fc.begin(), fc.end());
            fc.erase(std::unique(fc.begin(), fc.end()), fc.end());
            for (size_t x = 0; x < fc.size(); ++x)
                for (size_t y = x + 1; y < fc.size(); ++y) rows[fc[x]].push_back(fc[y]);
        }
        st.S.n = nf;
        st.S.rowptr.assign(nf + 1, 0);
        for (int a = 0; a < nf; ++a) {
            std::sort(rows[a].begin(), rows[a].end());
            rows[a].erase(std::unique(rows[a].begin(), rows[a].end()), rows[a].end());
            st.S.rowptr[a + 1] = st.S.rowptr[a] + int(rows[a].size());
        }
        st.S.col.reserve(st.S.rowptr[nf]);
        for (int a = 0; a < nf; ++a) st.S.col.insert(st.S.col.end(), rows[a].begin(), rows[a].end());
        st.S.val.assign(36 * st.S.col.size(), 0.0);
    }

    // SE3Perturbation on plain doubles for the poses, Euclidean plus for points (Evaluator::Plus)
    void plus(const Structure& st, const std::vector<double>& xp, const std::vector<double>& xl,
              const double* dp, const double* dl, std::vector<double>& yp,
              std::vector<double>& yl) const {
        yp = xp;
        yl = xl;
        SE3Perturbation P;
        for (size_t f = 0; f < st.free_cams.size(); ++f) {
            const int k = st.free_cams[f];
            P(&xp[12 * size_t(k)], dp + 6 * f, &yp[12 * size_t(k)]);
        }
        for (size_t a = 0; a < st.active_pts.size(); ++a) {
            const int j = st.active_pts[a];
            for (int c = 0; c < 3; ++c) yl[3 * size_t(j) + c] = xl[3 * size_t(j) + c] + dl[3 * a + c];
        }
    }

    struct Linear {
        int iterations = 0;
        bool ok = true;
    };

    // Solve  min |J y - r|^2 + |D y|^2  by eliminating the points; J already column-scaled.
    // Outputs y (cams then points).  Dp/Dl are the LM diagonals (sqrt(diag/radius)).
    Linear schur_solve(const Structure& st, const Eval& ev, const std::vector<double>& sc_p,
                       const std::vector<double>& sc_l, const std::vector<double>& Dp,
                       const std::vector<double>& Dl, const Options& opt, std::vector<double>& yp,
                       std::vector<double>& yl, BlockSym& S) const {
        const int nf = int(st.free_cams.size());
        const size_t na = st.active_pts.size();
        std::fill(S.val.begin(), S.val.end(), 0.0);
        std::vector<double> bp(6 * size_t(nf), 0.0);       // reduced rhs
        std::vector<double> Bdiag(36 * size_t(nf), 0.0);   // U + Dp^2 (JACOBI preconditioner)
        std::vector<double> Vinv(9 * na), gl(3 * na);
        // camera-only residuals (sun, prior) and the LM diagonal on the camera blocks
        auto add_cam_block = [&](int f, const double* J, const double* r, int rows) {
            double* Saa = &S.val[36 * size_t(S.rowptr[f])];
            const double* s = &sc_p[6 * size_t(f)];
            for (int i = 0; i < 6; ++i) {
                for (int j = 0; j < 6; ++j) {
                    double a = 0;
                    for (int k = 0; k < rows; ++k) a += J[6 * k + i] * J[6 * k + j];
                    Saa[6 * i + j] += a * s[i] * s[j];
                    Bdiag[36 * size_t(f) + 6 * i + j] += a * s[i] * s[j];
                }
                double g = 0;
                for (int k = 0; k < rows; ++k) g += J[6 * k + i] * r[k];
                bp[6 * size_t(f) + i] += g * s[i];
            }
        };
        for (size_t i = 0; i < suns.size(); ++i) {
            int f = st.cam_free[suns[i].cam];
            if (f >= 0) add_cam_block(f, &ev.J_sun[12 * i], &ev.r_sun[2 * i], 2);
        }
        for (size_t i = 0; i < priors.size(); ++i) {
            int f = st.cam_free[priors[i].cam];
            if (f >= 0) add_cam_block(f, &ev.J_pr[36 * i], &ev.r_pr[6 * i], 6);
        }
        for (int f = 0; f < nf; ++f)
            for (int i = 0; i < 6; ++i) {
                const double d2 = Dp[6 * size_t(f) + i] * Dp[6 * size_t(f) + i];
                S.val[36 * size_t(S.rowptr[f]) + 7 * i] += d2;
                Bdiag[36 * size_t(f) + 7 * i] += d2;
            }
        std::atomic<bool> ok{true};
        parallel_for(na, opt.num_threads, [&](size_t b, size_t e, int) {
            std::vector<double> Wl, Yl;
            std::vector<int> fl;
            for (size_t a = b; a < e; ++a) {
                const size_t o0 = st.pt_ptr[a], o1 = st.pt_ptr[a + 1];
                const double* sl = &sc_l[3 * a];
                double V[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, g[3] = {0, 0, 0};
                for (size_t x = o0; x < o1; ++x) {
                    const size_t i = st.pt_obs[x];
                    const double* Jp = &ev.Jp_st[9 * i];
                    const double* r = &ev.r_st[3 * i];
                    for (int p = 0; p < 3; ++p) {
                        for (int q = 0; q < 3; ++q) {
                            double s = 0;
                            for (int k = 0; k < 3; ++k) s += Jp[3 * k + p] * Jp[3 * k + q];
                            V[3 * p + q] += s * sl[p] * sl[q];
                        }
                        double s = 0;
                        for (int k = 0; k < 3; ++k) s += Jp[3 * k + p] * r[k];
                        g[p] += s * sl[p];
                    }
                }
                for (int p = 0; p < 3; ++p) V[4 * p] += Dl[3 * a + p] * Dl[3 * a + p];
                double Vi[9];
                if (!invert_spd3(V, Vi)) {
                    ok = false;
                    continue;
                }
                std::memcpy(&Vinv[9 * a], Vi, sizeof(Vi));
                std::memcpy(&gl[3 * a], g, sizeof(g));
                const size_t L = o1 - o0;
                Wl.assign(18 * L, 0.0);
                Yl.assign(18 * L, 0.0);
                fl.assign(L, -1);
                for (size_t x = o0; x < o1; ++x) {
                    const size_t i = st.pt_obs[x];
                    const int f = st.cam_free[st_cam[i]];
                    fl[x - o0] = f;
                    if (f < 0) continue;
                    const double* Jc = &ev.Jc_st[18 * i];
                    const double* Jp = &ev.Jp_st[9 * i];
                    const double* r = &ev.r_st[3 * i];
                    const double* sp = &sc_p[6 * size_t(f)];
                    double* W = &Wl[18 * (x - o0)];
                    double* Y = &Yl[18 * (x - o0)];
                    double U[36], ga[6];
                    for (int p = 0; p < 6; ++p) {
                        for (int q = 0; q < 6; ++q) {
                            double s = 0;
                            for (int k = 0; k < 3; ++k) s += Jc[6 * k + p] * Jc[6 * k + q];
                            U[6 * p + q] = s * sp[p] * sp[q];
                        }
                        for (int q = 0; q < 3; ++q) {
                            double s = 0;
                            for (int k = 0; k < 3; ++k) s += Jc[6 * k + p] * Jp[3 * k + q];
                            W[3 * p + q] = s * sp[p] * sl[q];
                        }
                        double s = 0;
                        for (int k = 0; k < 3; ++k) s += Jc[6 * k + p] * r[k];
                        ga[p] = s * sp[p];
                    }
                    for (int p = 0; p < 6; ++p)
                        for (int q = 0; q < 3; ++q)
                            Y[3 * p + q] = W[3 * p] * Vi[q] + W[3 * p + 1] * Vi[3 + q] +
                                           W[3 * p + 2] * Vi[6 + q];
                    double* Saa = &S.val[36 * size_t(S.rowptr[f])];
                    for (int p = 0; p < 6; ++p) {
                        for (int q = 0; q < 6; ++q) {
                            atomic_add(Saa[6 * p + q], U[6 * p + q]);
                            atomic_add(Bdiag[36 * size_t(f) + 6 * p + q], U[6 * p + q]);
                        }
                        const double yg = Y[3 * p] * g[0] + Y[3 * p + 1] * g[1] + Y[3 * p + 2] * g[2];
                        atomic_add(bp[6 * size_t(f) + p], ga[p] - yg);
                    }
                }
                for (size_t x = 0; x < L; ++x) {
                    if (fl[x] < 0) continue;
                    for (size_t y = 0; y < L; ++y) {
                        if (fl[y] < 0 || fl[y] < fl[x]) continue;
                        if (fl[y] == fl[x] && y < x) continue;  // same camera twice: keep x<=y once
                        const int e2 = S.find(fl[x], fl[y]);
                        double* B = &S.val[36 * size_t(e2)];
                        const double* Y = &Yl[18 * x];
                        const double* W = &Wl[18 * y];
                        for (int p = 0; p < 6; ++p)
                            for (int q = 0; q < 6; ++q) {
                                double v = Y[3 * p] * W[3 * q] + Y[3 * p + 1] * W[3 * q + 1] +
                                           Y[3 * p + 2] * W[3 * q + 2];
                                if (fl[y] == fl[x] && y != x) {
                                    // two observations of one point from the same camera: the
                                    // diagonal block receives both cross terms
                                    atomic_add(B[6 * p + q], -v);
                                    atomic_add(B[6 * q + p], -v);
                                } else {
                                    atomic_add(B[6 * p + q], -v);
                                }
                            }
                    }
                }
            }
        });
        Linear lin;
        if (!ok) {
            lin.ok = false;
            return lin;
        }
        // ---- reduced system S yp = bp ----
        const int n = 6 * nf;
        yp.assign(n, 0.0);
        if (n > 0) {
            if (opt.linear_solver == 0) {
                // exact: band Cholesky over the block pattern (dense when the band is full)
                int wb = 0;
                for (int a = 0; a < nf; ++a)
                    if (S.rowptr[a + 1] > S.rowptr[a]) wb = std::max(wb, S.col[S.rowptr[a + 1] - 1] - a);
                const int w = std::min(n - 1, 6 * wb + 5);
                std::vector<double> B(size_t(n) * (w + 1), 0.0);
                for (int a = 0; a < nf; ++a)
                    for (int e2 = S.rowptr[a]; e2 < S.rowptr[a + 1]; ++e2) {
                        const int bb = S.col[e2];
                        const double* blk = &S.val[36 * size_t(e2)];
                        for (int p = 0; p < 6; ++p)
                            for (int q = 0; q < 6; ++q) {
                                const int i = 6 * bb + q, j = 6 * a + p;  // lower: i >= j
                                if (i < j) continue;
                                B[size_t(i) * (w + 1) + (j - i + w)] = blk[6 * p + q];
                            }
                    }
                if (!band_cholesky(B, n, w)) {
                    lin.ok = false;
                    return lin;
                }
                yp = bp;
                band_cholesky_solve(B, n, w, yp.data());
                lin.iterations = 1;
            } else {
                lin = pcg(S, Bdiag, bp, opt, yp);
                if (!lin.ok) return lin;
            }
        }
        // ---- back-substitution: yl = Vinv (gl - W^T yp) ----
        yl.assign(3 * na, 0.0);
        parallel_for(na, opt.num_threads, [&](size_t b, size_t e, int) {
            for (size_t a = b; a < e; ++a) {
                const double* sl = &sc_l[3 * a];
                double t[3] = {gl[3 * a], gl[3 * a + 1], gl[3 * a + 2]};
                for (size_t x = st.pt_ptr[a]; x < st.pt_ptr[a + 1]; ++x) {
                    const size_t i = st.pt_obs[x];
                    const int f = st.cam_free[st_cam[i]];
                    if (f < 0) continue;
                    const double* Jc = &ev.Jc_st[18 * i];
                    const double* Jp = &ev.Jp_st[9 * i];
                    const double* sp = &sc_p[6 * size_t(f)];
                    double Jy[3] = {0, 0, 0};
                    for (int k = 0; k < 3; ++k)
                        for (int p = 0; p < 6; ++p) Jy[k] += Jc[6 * k + p] * sp[p] * yp[6 * size_t(f) + p];
                    for (int q = 0; q < 3; ++q)
                        for (int k = 0; k < 3; ++k) t[q] -= Jp[3 * k + q] * sl[q] * Jy[k];
                }
                const double* Vi = &Vinv[9 * a];
                for (int q = 0; q < 3; ++q) yl[3 * a + q] = Vi[3 * q] * t[0] + Vi[3 * q + 1] * t[1] + Vi[3 * q + 2] * t[2];
            }
        });
        return lin;
    }

    // Block-Jacobi PCG with the Ceres conjugate_gradients_solver termination rule
    // (Q-based, q_tolerance = eta, r_tolerance = -1, residual reset every 10 iterations).
    Linear pcg(const BlockSym& S, const std::vector<double>& Bdiag, const std::vector<double>& b,
               const Options& opt, std::vector<double>& x) const {
        Linear lin;
        const int nf = S.n, n = 6 * nf;
        std::vector<double> Minv(36 * size_t(nf));
        for (int f = 0; f < nf; ++f) {
            const double* blk = opt.preconditioner == 0 ? &Bdiag[36 * size_t(f)]
                                                         : &S.val[36 * size_t(S.rowptr[f])];
            if (!invert_spd6(blk, &Minv[36 * size_t(f)])) {
                lin.ok = false;
                return lin;
            }
        }
        auto dot = [&](const std::vector<double>& u, const std::vector<double>& v) {
            double s = 0;
            for (int i = 0; i < n; ++i) s += u[i] * v[i];
            return s;
        };
        std::fill(x.begin(), x.end(), 0.0);
        const double norm_b = std::sqrt(dot(b, b));
        if (norm_b == 0.0) return lin;
        std::vector<double> r = b, z(n), p(n), q(n), tmp(n);
        double rho = 1.0;
        double Q0 = 0.0;  // -x.(b + r) with x = 0
        for (lin.iterations = 1;; ++lin.iterations) {
            for (int f = 0; f < nf; ++f)
                for (int i = 0; i < 6; ++i) {
                    double s = 0;
                    for (int j = 0; j < 6; ++j) s += Minv[36 * size_t(f) + 6 * i + j] * r[6 * f + j];
                    z[6 * f + i] = s;
                }
            const double last_rho = rho;
            rho = dot(r, z);
            if (rho == 0.0 || !std::isfinite(rho)) {
                lin.ok = false;
                break;
            }
            if (lin.iterations == 1) {
                p = z;
            } else {
                const double beta = rho / last_rho;
                if (beta == 0.0 || !std::isfinite(beta)) {
                    lin.ok = false;
                    break;
                }
                for (int i = 0; i < n; ++i) p[i] = z[i] + beta * p[i];
            }
            S.multiply(p.data(), q.data());
            const double pq = dot(p, q);
            if (pq <= 0.0 || !std::isfinite(pq)) break;  // NO_CONVERGENCE: keep current x
            const double alpha = rho / pq;
            if (!std::isfinite(alpha)) {
                lin.ok = false;
                break;
            }
            for (int i = 0; i < n; ++i) x[i] += alpha * p[i];
            if (lin.iterations % 10 == 0) {
                S.multiply(x.data(), tmp.data());
                for (int i = 0; i < n; ++i) r[i] = b[i] - tmp[i];
            } else {
                for (int i = 0; i < n; ++i) r[i] -= alpha * q[i];
            }
            double Q1 = 0;
            for (int i = 0; i < n; ++i) Q1 -= x[i] * (b[i] + r[i]);
            const double zeta = lin.iterations * (Q1 - Q0) / Q1;
            if (zeta < opt.eta && lin.iterations >= opt.min_linear_solver_iterations) break;
            Q0 = Q1;
            if (lin.iterations >= opt.max_linear_solver_iterations) break;
        }
        return lin;
    }

    bool solve(const Options& opt, Summary& sum) {
        Structure st;
        build_structure(st);
        const int nf = int(st.free_cams.size());
        const size_t na = st.active_pts.size();
        std::vector<double> xp(poses, poses + 12 * size_t(n_poses));
        std::vector<double> xl(points, points + 3 * size_t(n_points));
        std::vector<double> cand_p, cand_l;
        Eval ev, evc;
        sum = Summary();
        if (!evaluate(xp.data(), xl.data(), true, true, ev, opt.num_threads)) {
            sum.termination_type = FAILURE;
            sum.termination_reason = R_INITIAL_EVAL;
            return false;
        }
        double x_cost = ev.cost;
        sum.initial_cost = x_cost;

        std::vector<double> sc_p(6 * size_t(nf), 1.0), sc_l(3 * na, 1.0);
        std::vector<double> gp(6 * size_t(nf)), gl(3 * na);
        std::vector<double> cn_p(6 * size_t(nf)), cn_l(3 * na);  // squared column norms (unscaled J)
        double gradient_max_norm = 0;

        auto column_norms_and_gradient = [&]() {
            std::fill(cn_p.begin(), cn_p.end(), 0.0);
            std::fill(cn_l.begin(), cn_l.end(), 0.0);
            std::fill(gp.begin(), gp.end(), 0.0);
            std::fill(gl.begin(), gl.end(), 0.0);
            for (size_t i = 0; i < n_stereo(); ++i) {
                const int f = st.cam_free[st_cam[i]];
                const int a = st.pt_active[st_pt[i]];
                const double* Jc = &ev.Jc_st[18 * i];
                const double* Jp = &ev.Jp_st[9 * i];
                const double* r = &ev.r_st[3 * i];
                for (int k = 0; k < 3; ++k) {
                    if (f >= 0)
                        for (int p = 0; p < 6; ++p) {
                            cn_p[6 * size_t(f) + p] += Jc[6 * k + p] * Jc[6 * k + p];
                            gp[6 * size_t(f) + p] += Jc[6 * k + p] * r[k];
                        }
                    for (int q = 0; q < 3; ++q) {
                        cn_l[3 * size_t(a) + q] += Jp[3 * k + q] * Jp[3 * k + q];
                        gl[3 * size_t(a) + q] += Jp[3 * k + q] * r[k];
                    }
                }
            }
            auto cam_only = [&](int f, const double* J, const double* r, int rows) {
                if (f < 0) return;
                for (int k = 0; k < rows; ++k)
                    for (int p = 0; p < 6; ++p) {
                        cn_p[6 * size_t(f) + p] += J[6 * k + p] * J[6 * k + p];
                        gp[6 * size_t(f) + p] += J[6 * k + p] * r[k];
                    }
            };
            for (size_t i = 0; i < suns.size(); ++i)
                cam_only(st.cam_free[suns[i].cam], &ev.J_sun[12 * i], &ev.r_sun[2 * i], 2);
            for (size_t i = 0; i < priors.size(); ++i)
                cam_only(st.cam_free[priors[i].cam], &ev.J_pr[36 * i], &ev.r_pr[6 * i], 6);
            // |x - Plus(x, -g)|_inf in ambient coordinates (trust_region_minimizer.cc)
            std::vector<double> ng_p(gp), ng_l(gl), yp2, yl2;
            for (auto& v : ng_p) v = -v;
            for (auto& v : ng_l) v = -v;
            plus(st, xp, xl, ng_p.data(), ng_l.data(), yp2, yl2);
            double m = 0;
            for (int f = 0; f < nf; ++f)
                for (int c = 0; c < 12; ++c) {
                    const size_t idx = 12 * size_t(st.free_cams[f]) + c;
                    m = std::max(m, std::fabs(xp[idx] - yp2[idx]));
                }
            for (size_t a = 0; a < na; ++a)
                for (int c = 0; c < 3; ++c) {
                    const size_t idx = 3 * size_t(st.active_pts[a]) + c;
                    m = std::max(m, std::fabs(xl[idx] - yl2[idx]));
                }
            gradient_max_norm = m;
        };
        column_norms_and_gradient();
        if (opt.jacobi_scaling) {
            for (size_t i = 0; i < sc_p.size(); ++i) sc_p[i] = 1.0 / (1.0 + std::sqrt(cn_p[i]));
            for (size_t i = 0; i < sc_l.size(); ++i) sc_l[i] = 1.0 / (1.0 + std::sqrt(cn_l[i]));
        }
        auto x_norm_of = [&](const std::vector<double>& P, const std::vector<double>& Lm) {
            double s = 0;
            for (int f = 0; f < nf; ++f)
                for (int c = 0; c < 12; ++c) {
                    const double v = P[12 * size_t(st.free_cams[f]) + c];
                    s += v * v;
                }
            for (size_t a = 0; a < na; ++a)
                for (int c = 0; c < 3; ++c) {
                    const double v = Lm[3 * size_t(st.active_pts[a]) + c];
                    s += v * v;
                }
            return std::sqrt(s);
        };
        double x_norm = x_norm_of(xp, xl);

        // trust_region_step_evaluator state
        const int max_nonmono = opt.use_nonmonotonic_steps ? opt.max_consecutive_nonmonotonic_steps : 0;
        double se_minimum = x_cost, se_current = x_cost, se_reference = x_cost, se_candidate = x_cost;
        double se_acc_ref = 0, se_acc_cand = 0;
        int se_nonmono = 0;
        // minimizer state
        double minimum_cost = x_cost;
        double radius = opt.initial_trust_region_radius, decrease_factor = 2.0;
        bool reuse_diagonal = false;
        std::vector<double> diag_p(6 * size_t(nf)), diag_l(3 * na), Dp(6 * size_t(nf)), Dl(3 * na);
        int invalid_steps = 0;
        IterationRow row{};
        row.iteration = 0;
        row.cost = x_cost;
        row.gradient_max_norm = gradient_max_norm;
        row.radius = radius;
        row.step_is_valid = row.step_is_successful = 0;
        sum.push_row(row);
        bool step_ok_prev = false;

        auto finish = [&](int type, int reason) {
            sum.termination_type = type;
            sum.termination_reason = reason;
        };
        int iteration = 0;
        BlockSym S = st.S;
        std::vector<double> yp, yl, dp(6 * size_t(nf)), dl(3 * na);

        // ---- DoglegStrategy [Ceres 1.x dogleg_strategy.cc, restated from memory] -------------------
        // Works in the coordinates step' = D step, D = sqrt(clamp(diag(J^T J))) of the (Jacobi-scaled)
        // Jacobian, where the trust region is a ball of radius `radius`.
        double dl_mu = 1e-8, dl_alpha = 0.0, dl_step_norm = 0.0;
        bool dl_reuse = false, dl_1d = false;
        std::vector<double> dgp, dgl, ggp, ggl, nnp, nnl, u0p, u0l, u1p, u1l;
        double sub_B[4] = {0, 0, 0, 0}, sub_g[2] = {0, 0};
        // |J v|^2-type products with the scaled Jacobian: rows = stereo 3, sun 2, prior 6
        auto jv = [&](const std::vector<double>& vp, const std::vector<double>& vl, std::vector<double>& out) {
            out.assign(3 * n_stereo() + 2 * suns.size() + 6 * priors.size(), 0.0);
            for (size_t i = 0; i < n_stereo(); ++i) {
                const int f = st.cam_free[st_cam[i]];
                const int a = st.pt_active[st_pt[i]];
                for (int k = 0; k < 3; ++k) {
                    double m = 0;
                    if (f >= 0)
                        for (int p = 0; p < 6; ++p) m += ev.Jc_st[18 * i + 6 * k + p] * sc_p[6 * size_t(f) + p] * vp[6 * size_t(f) + p];
                    for (int q = 0; q < 3; ++q) m += ev.Jp_st[9 * i + 3 * k + q] * sc_l[3 * size_t(a) + q] * vl[3 * size_t(a) + q];
                    out[3 * i + k] = m;
                }
            }
            size_t o = 3 * n_stereo();
            auto cam_only = [&](int f, const double* J, int rows) {
                for (int k = 0; k < rows; ++k, ++o)
                    if (f >= 0)
                        for (int p = 0; p < 6; ++p) out[o] += J[6 * k + p] * sc_p[6 * size_t(f) + p] * vp[6 * size_t(f) + p];
            };
            for (size_t i = 0; i < suns.size(); ++i) cam_only(st.cam_free[suns[i].cam], &ev.J_sun[12 * i], 2);
            for (size_t i = 0; i < priors.size(); ++i) cam_only(st.cam_free[priors[i].cam], &ev.J_pr[36 * i], 6);
        };
        auto dot2 = [](const std::vector<double>& a, const std::vector<double>& b) {
            double v = 0;
            for (size_t i = 0; i < a.size(); ++i) v += a[i] * b[i];
            return v;
        };
        auto traditional_step = [&](std::vector<double>& sp, std::vector<double>& sl) {
            // DoglegStrategy::ComputeTraditionalDoglegStep; result in D-scaled coordinates
            const double gnorm = std::sqrt(dot2(ggp, ggp) + dot2(ggl, ggl));
            const double nnorm = std::sqrt(dot2(nnp, nnp) + dot2(nnl, nnl));
            double cg = 0, cn = 0;
            if (nnorm <= radius) {
                cn = 1.0;
                dl_step_norm = nnorm;
            } else if (gnorm * dl_alpha >= radius) {
                cg = -(radius / gnorm);
                dl_step_norm = radius;
            } else {
                const double b_dot_a = -dl_alpha * (dot2(ggp, nnp) + dot2(ggl, nnl));
                const double a2 = std::pow(dl_alpha * gnorm, 2.0);
                const double bma2 = a2 - 2 * b_dot_a + std::pow(nnorm, 2);
                const double c = b_dot_a - a2;
                const double d = std::sqrt(c * c + bma2 * (std::pow(radius, 2.0) - a2));
                const double beta = (c <= 0) ? (d - c) / bma2 : (radius * radius - a2) / (d + c);
                cg = -dl_alpha * (1.0 - beta);
                cn = beta;
                dl_step_norm = -1.0;  // norm of the combination, below
            }
            sp.resize(ggp.size());
            sl.resize(ggl.size());
            for (size_t i = 0; i < sp.size(); ++i) sp[i] = cg * ggp[i] + cn * nnp[i];
            for (size_t i = 0; i < sl.size(); ++i) sl[i] = cg * ggl[i] + cn * nnl[i];
            if (dl_step_norm < 0.0) dl_step_norm = std::sqrt(dot2(sp, sp) + dot2(sl, sl));
        };
        auto subspace_step = [&](std::vector<double>& sp, std::vector<double>& sl) {
            // DoglegStrategy::ComputeSubspaceDoglegStep
            const double nnorm = std::sqrt(dot2(nnp, nnp) + dot2(nnl, nnl));
            if (nnorm <= radius) {
                sp = nnp;
                sl = nnl;
                dl_step_norm = nnorm;
                return;
            }
            if (dl_1d) {
                const double gnorm = std::sqrt(dot2(ggp, ggp) + dot2(ggl, ggl));
                sp.resize(ggp.size());
                sl.resize(ggl.size());
                for (size_t i = 0; i < sp.size(); ++i) sp[i] = -(radius / gnorm) * ggp[i];
                for (size_t i = 0; i < sl.size(); ++i) sl[i] = -(radius / gnorm) * ggl[i];
                dl_step_norm = radius;
                return;
            }
            double x[2];
            if (!dogleg_boundary_minimum(sub_B, sub_g, radius, x)) {
                traditional_step(sp, sl);
                return;
            }
            sp.resize(ggp.size());
            sl.resize(ggl.size());
            for (size_t i = 0; i < sp.size(); ++i) sp[i] = x[0] * u0p[i] + x[1] * u1p[i];
            for (size_t i = 0; i < sl.size(); ++i) sl[i] = x[0] * u0l[i] + x[1] * u1l[i];
            dl_step_norm = radius;
        };
        auto dogleg_compute_step = [&]() -> Linear {
            Linear lin;
            std::vector<double> sp, sl;
            if (!dl_reuse) {
                dl_reuse = true;
                const size_t np6 = 6 * size_t(nf), nl3 = 3 * na;
                dgp.resize(np6);
                dgl.resize(nl3);
                ggp.resize(np6);
                ggl.resize(nl3);
                for (size_t i = 0; i < np6; ++i) {
                    dgp[i] = std::sqrt(std::min(std::max(cn_p[i] * sc_p[i] * sc_p[i], opt.min_lm_diagonal), opt.max_lm_diagonal));
                    ggp[i] = gp[i] * sc_p[i] / dgp[i];  // ComputeGradient: D^-1 J^T r
                }
                for (size_t i = 0; i < nl3; ++i) {
                    dgl[i] = std::sqrt(std::min(std::max(cn_l[i] * sc_l[i] * sc_l[i], opt.min_lm_diagonal), opt.max_lm_diagonal));
                    ggl[i] = gl[i] * sc_l[i] / dgl[i];
                }
                // ComputeCauchyPoint: alpha = |g'|^2 / |J D^-1 g'|^2
                std::vector<double> tp(np6), tl(nl3), Jg;
                for (size_t i = 0; i < np6; ++i) tp[i] = ggp[i] / dgp[i];
                for (size_t i = 0; i < nl3; ++i) tl[i] = ggl[i] / dgl[i];
                jv(tp, tl, Jg);
                dl_alpha = (dot2(ggp, ggp) + dot2(ggl, ggl)) / dot2(Jg, Jg);
                // ComputeGaussNewtonStep: (J^T J + mu D^2) y = J^T r, mu raised on failure
                lin.ok = false;
                while (dl_mu < 1.0) {
                    std::vector<double> Dp2(np6), Dl2(nl3);
                    for (size_t i = 0; i < np6; ++i) Dp2[i] = dgp[i] * std::sqrt(dl_mu);
                    for (size_t i = 0; i < nl3; ++i) Dl2[i] = dgl[i] * std::sqrt(dl_mu);
                    lin = schur_solve(st, ev, sc_p, sc_l, Dp2, Dl2, opt, yp, yl, S);
                    bool fin = lin.ok;
                    if (fin)
                        for (double v : yp) fin = fin && std::isfinite(v);
                    if (fin)
                        for (double v : yl) fin = fin && std::isfinite(v);
                    if (!fin) {
                        dl_mu *= 10.0;
                        lin.ok = false;
                        continue;
                    }
                    break;
                }
                if (!lin.ok) return lin;
                nnp.resize(np6);
                nnl.resize(nl3);
                for (size_t i = 0; i < np6; ++i) nnp[i] = -dgp[i] * yp[i];
                for (size_t i = 0; i < nl3; ++i) nnl[i] = -dgl[i] * yl[i];
                if (opt.dogleg_type == 1) {
                    // ComputeSubspaceModel: orthonormal basis of span{g', gn'} (column-pivoted QR: the
                    // longer column first), rank test 2 eps |R00|
                    const double ng2 = dot2(ggp, ggp) + dot2(ggl, ggl), nn2 = dot2(nnp, nnp) + dot2(nnl, nnl);
                    const bool g_first = ng2 >= nn2;
                    const std::vector<double>&ap = g_first ? ggp : nnp, &al = g_first ? ggl : nnl;
                    const std::vector<double>&bp = g_first ? nnp : ggp, &bl = g_first ? nnl : ggl;
                    const double r00 = std::sqrt(std::max(ng2, nn2));
                    if (!(r00 > 0.0)) {
                        lin.ok = false;
                        return lin;
                    }
                    u0p.resize(np6);
                    u0l.resize(nl3);
                    u1p.resize(np6);
                    u1l.resize(nl3);
                    for (size_t i = 0; i < np6; ++i) u0p[i] = ap[i] / r00;
                    for (size_t i = 0; i < nl3; ++i) u0l[i] = al[i] / r00;
                    const double pr = dot2(u0p, bp) + dot2(u0l, bl);
                    for (size_t i = 0; i < np6; ++i) u1p[i] = bp[i] - pr * u0p[i];
                    for (size_t i = 0; i < nl3; ++i) u1l[i] = bl[i] - pr * u0l[i];
                    const double r11 = std::sqrt(dot2(u1p, u1p) + dot2(u1l, u1l));
                    dl_1d = !(r11 > 2.0 * std::numeric_limits<double>::epsilon() * r00);
                    if (!dl_1d) {
                        for (auto& v : u1p) v /= r11;
                        for (auto& v : u1l) v /= r11;
                        sub_g[0] = dot2(u0p, ggp) + dot2(u0l, ggl);
                        sub_g[1] = dot2(u1p, ggp) + dot2(u1l, ggl);
                        std::vector<double> J0, J1;
                        for (size_t i = 0; i < np6; ++i) tp[i] = u0p[i] / dgp[i];
                        for (size_t i = 0; i < nl3; ++i) tl[i] = u0l[i] / dgl[i];
                        jv(tp, tl, J0);
                        for (size_t i = 0; i < np6; ++i) tp[i] = u1p[i] / dgp[i];
                        for (size_t i = 0; i < nl3; ++i) tl[i] = u1l[i] / dgl[i];
                        jv(tp, tl, J1);
                        sub_B[0] = dot2(J0, J0);
                        sub_B[1] = sub_B[2] = dot2(J0, J1);
                        sub_B[3] = dot2(J1, J1);
                    }
                }
                lin.iterations = 1;
            } else {
                lin.iterations = 0;
            }
            if (opt.dogleg_type == 1)
                subspace_step(sp, sl);
            else
                traditional_step(sp, sl);
            // step = D^-1 step'; the minimiser's convention here is delta = -y (scaled coordinates)
            yp.resize(sp.size());
            yl.resize(sl.size());
            for (size_t i = 0; i < sp.size(); ++i) yp[i] = -sp[i] / dgp[i];
            for (size_t i = 0; i < sl.size(); ++i) yl[i] = -sl[i] / dgl[i];
            return lin;
        };
        for (;;) {
            // FinalizeIterationAndCheckIfMinimizerCanContinue
            if (iteration > 0) {
                if (step_ok_prev) {
                    ++sum.num_successful_steps;
                    if (x_cost < minimum_cost) {
                        minimum_cost = x_cost;
                        write_back(st, xp, xl);
                    }
                } else {
                    ++sum.num_unsuccessful_steps;
                }
            }
            if (iteration >= opt.max_num_iterations) {
                finish(NO_CONVERGENCE, R_MAX_ITERATIONS);
                break;
            }
            if (gradient_max_norm <= opt.gradient_tolerance) {
                finish(CONVERGENCE, R_GRADIENT_TOL);
                break;
            }
            if (radius < opt.min_trust_region_radius) {
                finish(CONVERGENCE, R_MIN_RADIUS);
                break;
            }
            ++iteration;
            step_ok_prev = false;
            row = IterationRow{};
            row.iteration = iteration;
            Linear lin;
            if (opt.trust_region_strategy == 1) {
                lin = dogleg_compute_step();
            } else {
                // ---- LevenbergMarquardtStrategy::ComputeStep ----
                if (!reuse_diagonal) {
                    for (size_t i = 0; i < diag_p.size(); ++i)
                        diag_p[i] = std::min(std::max(cn_p[i] * sc_p[i] * sc_p[i], opt.min_lm_diagonal),
                                             opt.max_lm_diagonal);
                    for (size_t i = 0; i < diag_l.size(); ++i)
                        diag_l[i] = std::min(std::max(cn_l[i] * sc_l[i] * sc_l[i], opt.min_lm_diagonal),
                                             opt.max_lm_diagonal);
                }
                for (size_t i = 0; i < Dp.size(); ++i) Dp[i] = std::sqrt(diag_p[i] / radius);
                for (size_t i = 0; i < Dl.size(); ++i) Dl[i] = std::sqrt(diag_l[i] / radius);
                lin = schur_solve(st, ev, sc_p, sc_l, Dp, Dl, opt, yp, yl, S);
            }
            reuse_diagonal = true;
            row.linear_iterations = lin.iterations;
            sum.total_linear_iterations += lin.iterations;
            bool valid = lin.ok;
            if (valid)
                for (double v : yp) valid = valid && std::isfinite(v);
            if (valid)
                for (double v : yl) valid = valid && std::isfinite(v);
            double model_cost_change = 0;
            if (valid) {
                // step (scaled coordinates) = -y ; model_cost_change = -(J s).(r + J s / 2)
                double acc = 0;
                std::vector<double> partial(std::max(1, opt.num_threads), 0.0);
                parallel_for(n_stereo(), opt.num_threads, [&](size_t b, size_t e, int tid) {
                    double a2 = 0;
                    for (size_t i = b; i < e; ++i) {
                        const int f = st.cam_free[st_cam[i]];
                        const int a = st.pt_active[st_pt[i]];
                        const double* Jc = &ev.Jc_st[18 * i];
                        const double* Jp = &ev.Jp_st[9 * i];
                        const double* r = &ev.r_st[3 * i];
                        for (int k = 0; k < 3; ++k) {
                            double m = 0;
                            if (f >= 0)
                                for (int p = 0; p < 6; ++p)
                                    m -= Jc[6 * k + p] * sc_p[6 * size_t(f) + p] * yp[6 * size_t(f) + p];
                            for (int q = 0; q < 3; ++q)
                                m -= Jp[3 * k + q] * sc_l[3 * size_t(a) + q] * yl[3 * size_t(a) + q];
                            a2 -= m * (r[k] + 0.5 * m);
                        }
                    }
                    partial[tid] += a2;
                });
                for (double v : partial) acc += v;
                auto cam_only = [&](int f, const double* J, const double* r, int rows) {
                    for (int k = 0; k < rows; ++k) {
                        double m = 0;
                        if (f >= 0)
                            for (int p = 0; p < 6; ++p)
                                m -= J[6 * k + p] * sc_p[6 * size_t(f) + p] * yp[6 * size_t(f) + p];
                        acc -= m * (r[k] + 0.5 * m);
                    }
                };
                for (size_t i = 0; i < suns.size(); ++i)
                    cam_only(st.cam_free[suns[i].cam], &ev.J_sun[12 * i], &ev.r_sun[2 * i], 2);
                for (size_t i = 0; i < priors.size(); ++i)
                    cam_only(st.cam_free[priors[i].cam], &ev.J_pr[36 * i], &ev.r_pr[6 * i], 6);
                model_cost_change = acc;
                if (!(model_cost_change > 0.0)) valid = false;
            }
            if (!valid) {
                row.step_is_valid = 0;
                row.cost = x_cost;
                row.gradient_max_norm = gradient_max_norm;
                if (++invalid_steps >= opt.max_num_consecutive_invalid_steps) {
                    row.radius = radius;
                    sum.push_row(row);
                    finish(FAILURE, R_INVALID_STEPS);
                    break;
                }
                if (opt.trust_region_strategy == 1) {
                    dl_mu *= 10.0;  // DoglegStrategy::StepIsInvalid
                    dl_reuse = false;
                } else {
                    radius = radius / decrease_factor;  // StepIsInvalid == StepRejected(0)
                    decrease_factor *= 2.0;
                    reuse_diagonal = true;
                }
                row.radius = radius;
                sum.push_row(row);
                continue;
            }
            invalid_steps = 0;
            row.step_is_valid = 1;
            for (size_t i = 0; i < dp.size(); ++i) dp[i] = -yp[i] * sc_p[i];
            for (size_t i = 0; i < dl.size(); ++i) dl[i] = -yl[i] * sc_l[i];
            plus(st, xp, xl, dp.data(), dl.data(), cand_p, cand_l);
            double cand_cost = std::numeric_limits<double>::max();
            if (evaluate(cand_p.data(), cand_l.data(), false, true, evc, opt.num_threads))
                cand_cost = evc.cost;
            // ParameterToleranceReached
            double sn = 0;
            for (int f = 0; f < nf; ++f)
                for (int c = 0; c < 12; ++c) {
                    const size_t idx = 12 * size_t(st.free_cams[f]) + c;
                    sn += (xp[idx] - cand_p[idx]) * (xp[idx] - cand_p[idx]);
                }
            for (size_t a = 0; a < na; ++a)
                for (int c = 0; c < 3; ++c) {
                    const size_t idx = 3 * size_t(st.active_pts[a]) + c;
                    sn += (xl[idx] - cand_l[idx]) * (xl[idx] - cand_l[idx]);
                }
            row.step_norm = std::sqrt(sn);
            row.cost_change = x_cost - cand_cost;
            if (row.step_norm <= opt.parameter_tolerance * (x_norm + opt.parameter_tolerance)) {
                row.cost = x_cost;
                row.gradient_max_norm = gradient_max_norm;
                row.radius = radius;
                sum.push_row(row);
                finish(CONVERGENCE, R_PARAMETER_TOL);
                break;
            }
            // FunctionToleranceReached
            if (std::fabs(row.cost_change) <= opt.function_tolerance * x_cost) {
                row.cost = x_cost;
                row.gradient_max_norm = gradient_max_norm;
                row.radius = radius;
                sum.push_row(row);
                finish(CONVERGENCE, R_FUNCTION_TOL);
                break;
            }
            // IsStepSuccessful — TrustRegionStepEvaluator::StepQuality
            const double rel = (se_current - cand_cost) / model_cost_change;
            const double hist = (se_reference - cand_cost) / (se_acc_ref + model_cost_change);
            row.relative_decrease = std::max(rel, hist);
            if (row.relative_decrease > opt.min_relative_decrease) {
                xp = cand_p;
                xl = cand_l;
                x_norm = x_norm_of(xp, xl);
                if (!evaluate(xp.data(), xl.data(), true, true, ev, opt.num_threads)) {
                    finish(FAILURE, R_INITIAL_EVAL);
                    break;
                }
                x_cost = ev.cost;
                column_norms_and_gradient();
                step_ok_prev = true;
                row.step_is_successful = 1;
                // strategy StepAccepted
                if (opt.trust_region_strategy == 1) {
                    if (row.relative_decrease < 0.25) radius *= 0.5;
                    if (row.relative_decrease > 0.75) {
                        radius = std::max(radius, 3.0 * dl_step_norm);
                        radius = std::min(radius, opt.max_trust_region_radius);
                    }
                    dl_mu = std::max(1e-8, 2.0 * dl_mu / 10.0);
                    dl_reuse = false;
                } else {
                    radius = radius / std::max(1.0 / 3.0, 1.0 - std::pow(2.0 * row.relative_decrease - 1.0, 3));
                    radius = std::min(opt.max_trust_region_radius, radius);
                    decrease_factor = 2.0;
                    reuse_diagonal = false;
                }
                // step evaluator StepAccepted(candidate_cost, model_cost_change)
                se_current = cand_cost;
                se_acc_cand += model_cost_change;
                se_acc_ref += model_cost_change;
                if (se_current < se_minimum) {
                    se_minimum = se_current;
                    se_nonmono = 0;
                    se_candidate = se_current;
                    se_acc_cand = 0;
                } else {
                    ++se_nonmono;
                    if (se_current > se_candidate) {
                        se_candidate = se_current;
                        se_acc_cand = 0;
                    }
                }
                if (se_nonmono == max_nonmono) {
                    se_reference = se_candidate;
                    se_acc_ref = se_acc_cand;
                }
            } else if (opt.trust_region_strategy == 1) {
                radius *= 0.5;  // DoglegStrategy::StepRejected
                dl_reuse = true;
            } else {
                radius = radius / decrease_factor;
                decrease_factor *= 2.0;
                reuse_diagonal = true;
            }
            row.cost = x_cost;
            row.gradient_max_norm = gradient_max_norm;
            row.radius = radius;
            sum.push_row(row);
        }
        sum.num_iterations = iteration;
        sum.final_cost = minimum_cost;
        sum.final_radius = radius;
        return sum.termination_type != FAILURE;
    }

    void write_back(const Structure& st, const std::vector<double>& xp, const std::vector<double>& xl) {
        for (int k : st.free_cams) std::memcpy(poses + 12 * size_t(k), &xp[12 * size_t(k)], 12 * sizeof(double));
        for (int j : st.active_pts) std::memcpy(points + 3 * size_t(j), &xl[3 * size_t(j)], 3 * sizeof(double));
    }
};

}  // namespace oracle
