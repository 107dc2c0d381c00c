// ORACLE — TEST INFRASTRUCTURE ONLY (the functors evaluated with these dual numbers are PINNED against the reference's
// own headers on oracle/_ref's stand-in for ceres::Jet, tests/test_ref_pin.py; the reference ships no golden vectors).
//
// Forward-mode dual number with N partials. This is the arithmetic that
// `ceres::AutoDiffCostFunction` / `ceres::AutoDiffLocalParameterization` run the reference's
// templated functors on (reference call sites: stereo_reprojection_error.hpp:62,
// sun_sensor_error.hpp:114, pose_error.hpp:61, normal_error.hpp:48,
// intensity_error_point_light.hpp:102, perturbations.hpp:36,70,110).  Ceres itself is not in
// /root/reference (un-vendored, un-pinned `find_package(Ceres)`, CMakeLists.txt:17); the rules
// below are the published chain rules of ceres/jet.h (1.x): value part drives every comparison,
// each elementary function carries its exact derivative.
#pragma once
#include <cmath>
#include <limits>

namespace oracle {

template <int N>
struct Jet {
    double a;
    double v[N];

    Jet() : a(0.0) {
        for (int i = 0; i < N; ++i) v[i] = 0.0;
    }
    Jet(double value) : a(value) {  // NOLINT (implicit on purpose, like ceres::Jet)
        for (int i = 0; i < N; ++i) v[i] = 0.0;
    }
    Jet(double value, int k) : a(value) {
        for (int i = 0; i < N; ++i) v[i] = 0.0;
        v[k] = 1.0;
    }
    Jet& operator+=(const Jet& y) {
        a += y.a;
        for (int i = 0; i < N; ++i) v[i] += y.v[i];
        return *this;
    }
    Jet& operator-=(const Jet& y) {
        a -= y.a;
        for (int i = 0; i < N; ++i) v[i] -= y.v[i];
        return *this;
    }
};

template <int N>
inline Jet<N> operator-(const Jet<N>& f) {
    Jet<N> h;
    h.a = -f.a;
    for (int i = 0; i < N; ++i) h.v[i] = -f.v[i];
    return h;
}
template <int N>
inline Jet<N> operator+(const Jet<N>& f, const Jet<N>& g) {
    Jet<N> h;
    h.a = f.a + g.a;
    for (int i = 0; i < N; ++i) h.v[i] = f.v[i] + g.v[i];
    return h;
}
template <int N>
inline Jet<N> operator+(const Jet<N>& f, double s) {
    Jet<N> h = f;
    h.a += s;
    return h;
}
template <int N>
inline Jet<N> operator+(double s, const Jet<N>& f) {
    return f + s;
}
template <int N>
inline Jet<N> operator-(const Jet<N>& f, const Jet<N>& g) {
    Jet<N> h;
    h.a = f.a - g.a;
    for (int i = 0; i < N; ++i) h.v[i] = f.v[i] - g.v[i];
    return h;
}
template <int N>
inline Jet<N> operator-(const Jet<N>& f, double s) {
    Jet<N> h = f;
    h.a -= s;
    return h;
}
template <int N>
inline Jet<N> operator-(double s, const Jet<N>& f) {
    Jet<N> h;
    h.a = s - f.a;
    for (int i = 0; i < N; ++i) h.v[i] = -f.v[i];
    return h;
}
template <int N>
inline Jet<N> operator*(const Jet<N>& f, const Jet<N>& g) {
    Jet<N> h;
    h.a = f.a * g.a;
    for (int i = 0; i < N; ++i) h.v[i] = f.a * g.v[i] + f.v[i] * g.a;
    return h;
}
template <int N>
inline Jet<N> operator*(const Jet<N>& f, double s) {
    Jet<N> h;
    h.a = f.a * s;
    for (int i = 0; i < N; ++i) h.v[i] = f.v[i] * s;
    return h;
}
template <int N>
inline Jet<N> operator*(double s, const Jet<N>& f) {
    return f * s;
}
template <int N>
inline Jet<N> operator/(const Jet<N>& f, const Jet<N>& g) {
    // ceres/jet.h: one reciprocal, then products.
    const double g_a_inverse = 1.0 / g.a;
    const double f_a_by_g_a = f.a * g_a_inverse;
    Jet<N> h;
    h.a = f_a_by_g_a;
    for (int i = 0; i < N; ++i) h.v[i] = (f.v[i] - f_a_by_g_a * g.v[i]) * g_a_inverse;
    return h;
}
template <int N>
inline Jet<N> operator/(const Jet<N>& f, double s) {
    const double s_inverse = 1.0 / s;
    return f * s_inverse;
}
template <int N>
inline Jet<N> operator/(double s, const Jet<N>& g) {
    const double minus_s_g_a_inverse2 = -s / (g.a * g.a);
    Jet<N> h;
    h.a = s / g.a;
    for (int i = 0; i < N; ++i) h.v[i] = g.v[i] * minus_s_g_a_inverse2;
    return h;
}

// Comparisons look at the value only (ceres/jet.h CERES_DEFINE_JET_COMPARISON_OPERATOR).
#define ORACLE_JET_CMP(op)                                        \
    template <int N>                                              \
    inline bool operator op(const Jet<N>& f, const Jet<N>& g) {   \
        return f.a op g.a;                                        \
    }                                                             \
    template <int N>                                              \
    inline bool operator op(const Jet<N>& f, double s) {          \
        return f.a op s;                                          \
    }                                                             \
    template <int N>                                              \
    inline bool operator op(double s, const Jet<N>& g) {          \
        return s op g.a;                                          \
    }
ORACLE_JET_CMP(<)
ORACLE_JET_CMP(<=)
ORACLE_JET_CMP(>)
ORACLE_JET_CMP(>=)
ORACLE_JET_CMP(==)
ORACLE_JET_CMP(!=)
#undef ORACLE_JET_CMP

template <int N>
inline Jet<N> sqrt(const Jet<N>& f) {
    const double tmp = std::sqrt(f.a);
    const double two_a_inverse = 1.0 / (2.0 * tmp);
    Jet<N> h;
    h.a = tmp;
    for (int i = 0; i < N; ++i) h.v[i] = f.v[i] * two_a_inverse;
    return h;
}
template <int N>
inline Jet<N> sin(const Jet<N>& f) {
    const double c = std::cos(f.a);
    Jet<N> h;
    h.a = std::sin(f.a);
    for (int i = 0; i < N; ++i) h.v[i] = c * f.v[i];
    return h;
}
template <int N>
inline Jet<N> cos(const Jet<N>& f) {
    const double ms = -std::sin(f.a);
    Jet<N> h;
    h.a = std::cos(f.a);
    for (int i = 0; i < N; ++i) h.v[i] = ms * f.v[i];
    return h;
}
template <int N>
inline Jet<N> acos(const Jet<N>& f) {
    const double tmp = -1.0 / std::sqrt(1.0 - f.a * f.a);
    Jet<N> h;
    h.a = std::acos(f.a);
    for (int i = 0; i < N; ++i) h.v[i] = tmp * f.v[i];
    return h;
}
// atan2(g, f) = atan(g / f): d = (f dg - g df) / (f^2 + g^2)
template <int N>
inline Jet<N> atan2(const Jet<N>& g, const Jet<N>& f) {
    const double tmp = 1.0 / (f.a * f.a + g.a * g.a);
    Jet<N> h;
    h.a = std::atan2(g.a, f.a);
    for (int i = 0; i < N; ++i) h.v[i] = tmp * (-g.a * f.v[i] + f.a * g.v[i]);
    return h;
}
// pow(jet, jet) as in ceres 1.x: d = g f^(g-1) df + f^g log(f) dg
template <int N>
inline Jet<N> pow(const Jet<N>& f, const Jet<N>& g) {
    const double tmp1 = std::pow(f.a, g.a);
    const double tmp2 = g.a * std::pow(f.a, g.a - 1.0);
    const double tmp3 = tmp1 * std::log(f.a);
    Jet<N> h;
    h.a = tmp1;
    for (int i = 0; i < N; ++i) h.v[i] = tmp2 * f.v[i] + tmp3 * g.v[i];
    return h;
}

// Value accessors so templated code can branch on the scalar part (Eigen's allFinite() on a
// Jet matrix only ever looks at `.a` through operator==).
inline double value_of(double x) { return x; }
template <int N>
inline double value_of(const Jet<N>& x) {
    return x.a;
}

// Plain-double overloads picked up by ADL-free unqualified calls in the templated functors.
using std::acos;
using std::atan2;
using std::cos;
using std::pow;
using std::sin;
using std::sqrt;

}  // namespace oracle
