// Scalar part of the DOGLEG trust-region strategy (Ceres DoglegStrategy, restated: SURVEY.md 8f-3).
// The device supplies eight inner products per Jacobian (kernels.cuh, DoglegSum); everything the
// strategy decides — Cauchy point, traditional interpolation, the two-dimensional subspace model
// and the quartic of its boundary problem — is arithmetic on those scalars, and the step is
// returned as two coefficients: step' = c1 g' + c2 gn' in D-scaled coordinates.
// Host AND device code: the host-driven engine (engine.cu) runs it on the CPU between launches,
// the one-CTA-per-window kernel (kernels_window.cu) runs the same functions on thread 0.
#pragma once
#include <cfloat>
#include <cmath>

#if defined(__CUDACC__)
#define CSLAM_HD __host__ __device__
#else
#define CSLAM_HD
#endif

namespace cslam {

struct Cplx {
    double re, im;
};
CSLAM_HD inline Cplx c_add(Cplx a, Cplx b) { return {a.re + b.re, a.im + b.im}; }
CSLAM_HD inline Cplx c_sub(Cplx a, Cplx b) { return {a.re - b.re, a.im - b.im}; }
CSLAM_HD inline Cplx c_mul(Cplx a, Cplx b) { return {a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; }
CSLAM_HD inline double c_abs(Cplx a) { return hypot(a.re, a.im); }
// Smith's division (no overflow of |b|^2)
CSLAM_HD inline Cplx c_div(Cplx a, Cplx b) {
    if (fabs(b.re) >= fabs(b.im)) {
        const double r = b.im / b.re, d = b.re + b.im * r;
        return {(a.re + a.im * r) / d, (a.im - a.re * r) / d};
    }
    const double r = b.re / b.im, d = b.re * r + b.im;
    return {(a.re * r + a.im) / d, (a.im * r - a.re) / d};
}

// real parts of the roots of c[0] y^n + ... + c[n], n <= 4 (Durand-Kerner); returns the root count
CSLAM_HD inline int poly_root_real_parts(const double* c_in, int len, double* out) {
    int lo = 0;
    while (lo < len && c_in[lo] == 0.0) ++lo;
    const int n = len - lo - 1;
    if (n < 1) return 0;
    double a[5];
    Cplx z[4];
    for (int i = 0; i <= n; ++i) a[i] = c_in[lo + i] / c_in[lo];
    double rad = 0;
    for (int i = 1; i <= n; ++i) rad = fmax(rad, pow(fabs(a[i]), 1.0 / i));
    rad = 2.0 * rad + 1e-300;
    for (int i = 0; i < n; ++i) {
        const double th = 2.0 * 3.14159265358979323846 * i / n + 0.4;
        z[i] = {rad * cos(th), rad * sin(th)};
    }
    for (int it = 0; it < 500; ++it) {
        double delta = 0;
        for (int i = 0; i < n; ++i) {
            Cplx pv = {a[0], 0.0};
            for (int k = 1; k <= n; ++k) pv = c_add(c_mul(pv, z[i]), Cplx{a[k], 0.0});
            Cplx den = {1.0, 0.0};
            for (int j = 0; j < n; ++j)
                if (j != i) den = c_mul(den, c_sub(z[i], z[j]));
            if (c_abs(den) == 0.0) den = {1e-300, 0.0};
            const Cplx dz = c_div(pv, den);
            z[i] = c_sub(z[i], dz);
            delta = fmax(delta, c_abs(dz) / fmax(c_abs(z[i]), 1e-300));
        }
        if (delta < 1e-15) break;
    }
    for (int i = 0; i < n; ++i) out[i] = z[i].re;
    return n;
}

struct DoglegModel {
    // inputs: the eight sums
    double G11 = 0, G12 = 0, G22 = 0, JGG = 0, JGY = 0, JYY = 0, JGR = 0, JYR = 0;
    // derived once per Jacobian
    double alpha = 0;
    bool one_d = false;
    double u0[2] = {0, 0}, u1[2] = {0, 0};  // subspace basis as coefficients over (g', gn')
    double B[4] = {0, 0, 0, 0}, g[2] = {0, 0};

    CSLAM_HD double gram(const double a[2], const double b[2]) const { return a[0] * b[0] * G11 + (a[0] * b[1] + a[1] * b[0]) * G12 + a[1] * b[1] * G22; }
    // (J D^-1 a) . (J D^-1 b):  J D^-1 (c1 g' + c2 gn') = c1 Jg - c2 Jy
    CSLAM_HD double jgram(const double a[2], const double b[2]) const {
        return a[0] * b[0] * JGG - (a[0] * b[1] + a[1] * b[0]) * JGY + a[1] * b[1] * JYY;
    }
    // false: both vectors vanish
    CSLAM_HD bool prepare(bool subspace) {
        alpha = G11 / JGG;
        if (!subspace) return true;
        const bool g_first = G11 >= G22;
        const double a[2] = {g_first ? 1.0 : 0.0, g_first ? 0.0 : 1.0}, b[2] = {g_first ? 0.0 : 1.0, g_first ? 1.0 : 0.0};
        const double r00 = sqrt(fmax(G11, G22));
        if (!(r00 > 0.0)) return false;
        u0[0] = a[0] / r00;
        u0[1] = a[1] / r00;
        const double pr = gram(u0, b);
        double raw[2] = {b[0] - pr * u0[0], b[1] - pr * u0[1]};
        const double r11 = sqrt(fmax(gram(raw, raw), 0.0));
        one_d = !(r11 > 2.0 * DBL_EPSILON * r00);
        if (one_d) return true;
        u1[0] = raw[0] / r11;
        u1[1] = raw[1] / r11;
        const double gv[2] = {1.0, 0.0};
        g[0] = gram(u0, gv);
        g[1] = gram(u1, gv);
        B[0] = jgram(u0, u0);
        B[1] = B[2] = jgram(u0, u1);
        B[3] = jgram(u1, u1);
        return true;
    }
    CSLAM_HD double norm(double c1, double c2) const {
        const double c[2] = {c1, c2};
        return sqrt(fmax(gram(c, c), 0.0));
    }
    CSLAM_HD void traditional(double radius, double* c1, double* c2, double* step_norm) const {
        const double gnorm = sqrt(G11), nnorm = sqrt(G22);
        if (nnorm <= radius) {
            *c1 = 0.0, *c2 = 1.0, *step_norm = nnorm;
            return;
        }
        if (gnorm * alpha >= radius) {
            *c1 = -(radius / gnorm), *c2 = 0.0, *step_norm = radius;
            return;
        }
        const double b_dot_a = -alpha * G12, a2 = pow(alpha * gnorm, 2.0);
        const double bma2 = a2 - 2 * b_dot_a + pow(nnorm, 2);
        const double c = b_dot_a - a2;
        const double d = sqrt(c * c + bma2 * (pow(radius, 2.0) - a2));
        const double beta = (c <= 0) ? (d - c) / bma2 : (radius * radius - a2) / (d + c);
        *c1 = -alpha * (1.0 - beta);
        *c2 = beta;
        *step_norm = norm(*c1, *c2);
    }
    CSLAM_HD bool boundary_minimum(double r, double x_out[2]) const {
        const double detB = B[0] * B[3] - B[1] * B[2], trB = B[0] + B[3], r2 = r * r;
        const double Ba[4] = {B[3], -B[1], -B[2], B[0]};
        const double gg = g[0] * g[0] + g[1] * g[1];
        const double Bag[2] = {Ba[0] * g[0] + Ba[1] * g[1], Ba[2] * g[0] + Ba[3] * g[1]};
        double poly[5], roots[4];
        poly[0] = r2;
        poly[1] = 2.0 * r2 * trB;
        poly[2] = r2 * (trB * trB + 2.0 * detB) - gg;
        poly[3] = -2.0 * ((g[0] * Bag[0] + g[1] * Bag[1]) - r2 * detB * trB);
        poly[4] = r2 * detB * detB - (Bag[0] * Bag[0] + Bag[1] * Bag[1]);
        const int n_roots = poly_root_real_parts(poly, 5, roots);
        x_out[0] = x_out[1] = 0.0;
        double best = DBL_MAX;
        bool found = false;
        for (int ri = 0; ri < n_roots; ++ri) {
            const double y = roots[ri];
            const double a = B[0] + y, b = B[1], c2 = B[2], d = B[3] + y;
            const double det = a * d - b * c2;
            if (det == 0.0 || !isfinite(det)) continue;
            const double x[2] = {-(d * g[0] - b * g[1]) / det, -(-c2 * g[0] + a * g[1]) / det};
            const double nx = sqrt(x[0] * x[0] + x[1] * x[1]);
            if (!(nx > 0.0) || !isfinite(nx)) continue;
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
    CSLAM_HD void subspace(double radius, double* c1, double* c2, double* step_norm) const {
        const double nnorm = sqrt(G22);
        if (nnorm <= radius) {
            *c1 = 0.0, *c2 = 1.0, *step_norm = nnorm;
            return;
        }
        if (one_d) {
            *c1 = -(radius / sqrt(G11)), *c2 = 0.0, *step_norm = radius;
            return;
        }
        double x[2];
        if (!boundary_minimum(radius, x)) {
            traditional(radius, c1, c2, step_norm);
            return;
        }
        *c1 = x[0] * u0[0] + x[1] * u1[0];
        *c2 = x[0] * u0[1] + x[1] * u1[1];
        *step_norm = radius;
    }
    // -(J s).(r + J s / 2) for the step with coefficients (c1, c2): J s = c1 Jg - c2 Jy
    CSLAM_HD double model_cost_change(double c1, double c2) const {
        const double js_r = c1 * JGR - c2 * JYR;
        const double js2 = c1 * c1 * JGG - 2.0 * c1 * c2 * JGY + c2 * c2 * JYY;
        return -(js_r + 0.5 * js2);
    }
};

}  // namespace cslam
