// TEST INFRASTRUCTURE — stand-in for the part of Ceres Solver's PUBLIC API that the reference's
// hot-path headers name: ceres::Jet, AutoDiffCostFunction, AutoDiffLocalParameterization and their
// base classes.  Ceres is not in the image (and there is no network); this header lets
// oracle/ref_capi.cpp instantiate the reference's own, unmodified cost functors and plus
// operations so that `Create()->Evaluate(...)` and `Create()->ComputeJacobian(...)` run the
// reference's templates exactly as `ceres::Problem` would.  Written from Ceres' documented
// behaviour (jet.h's chain rules, autodiff_cost_function.h's calling convention); no Ceres code.
//
// What is restated here and therefore NOT pinned by the reference: the chain rule of each
// elementary function (listed beside it) and the order in which AutoDiff seeds the partials.  The
// solver itself (Problem, Solve, Covariance, loss functions) is NOT provided: the trust-region
// rules stay "restated from the Ceres 1.x sources" (SURVEY.md App. B).
#ifndef CSLAM_REF_STANDIN_CERES_H
#define CSLAM_REF_STANDIN_CERES_H

#include <cmath>
#include <limits>
#include <memory>
#include <typeinfo>
#include <utility>
#include <vector>

#include <Eigen/Core>

namespace ceres {

template <typename T, int N>
struct Jet {
    T a;
    T v[N];
    Jet() : a() { for (int i = 0; i < N; ++i) v[i] = T(); }
    Jet(const T& value) : a(value) { for (int i = 0; i < N; ++i) v[i] = T(); }   // NOLINT (implicit, as in Ceres)
    Jet(const T& value, int k) : a(value) { for (int i = 0; i < N; ++i) v[i] = T(); v[k] = T(1.0); }
    Jet& operator+=(const Jet& y) { *this = *this + y; return *this; }
    Jet& operator-=(const Jet& y) { *this = *this - y; return *this; }
    Jet& operator*=(const Jet& y) { *this = *this * y; return *this; }
    Jet& operator/=(const Jet& y) { *this = *this / y; return *this; }
};

#define CSLAM_JET_LOOP for (int i = 0; i < N; ++i)
template <typename T, int N> inline Jet<T, N> operator+(const Jet<T, N>& f) { return f; }
template <typename T, int N> inline Jet<T, N> operator-(const Jet<T, N>& f) {
    Jet<T, N> h; h.a = -f.a; CSLAM_JET_LOOP h.v[i] = -f.v[i]; return h; }
template <typename T, int N> inline Jet<T, N> operator+(const Jet<T, N>& f, const Jet<T, N>& g) {
    Jet<T, N> h; h.a = f.a + g.a; CSLAM_JET_LOOP h.v[i] = f.v[i] + g.v[i]; return h; }
template <typename T, int N> inline Jet<T, N> operator+(const Jet<T, N>& f, T s) { Jet<T, N> h = f; h.a = f.a + s; return h; }
template <typename T, int N> inline Jet<T, N> operator+(T s, const Jet<T, N>& f) { Jet<T, N> h = f; h.a = f.a + s; return h; }
template <typename T, int N> inline Jet<T, N> operator-(const Jet<T, N>& f, const Jet<T, N>& g) {
    Jet<T, N> h; h.a = f.a - g.a; CSLAM_JET_LOOP h.v[i] = f.v[i] - g.v[i]; return h; }
template <typename T, int N> inline Jet<T, N> operator-(const Jet<T, N>& f, T s) { Jet<T, N> h = f; h.a = f.a - s; return h; }
template <typename T, int N> inline Jet<T, N> operator-(T s, const Jet<T, N>& f) {
    Jet<T, N> h; h.a = s - f.a; CSLAM_JET_LOOP h.v[i] = -f.v[i]; return h; }
// d(fg) = f dg + g df
template <typename T, int N> inline Jet<T, N> operator*(const Jet<T, N>& f, const Jet<T, N>& g) {
    Jet<T, N> h; h.a = f.a * g.a; CSLAM_JET_LOOP h.v[i] = f.a * g.v[i] + f.v[i] * g.a; return h; }
template <typename T, int N> inline Jet<T, N> operator*(const Jet<T, N>& f, T s) {
    Jet<T, N> h; h.a = f.a * s; CSLAM_JET_LOOP h.v[i] = f.v[i] * s; return h; }
template <typename T, int N> inline Jet<T, N> operator*(T s, const Jet<T, N>& f) {
    Jet<T, N> h; h.a = f.a * s; CSLAM_JET_LOOP h.v[i] = f.v[i] * s; return h; }
// d(f/g) = (df - (f/g) dg) / g, evaluated with one reciprocal of g
template <typename T, int N> inline Jet<T, N> operator/(const Jet<T, N>& f, const Jet<T, N>& g) {
    const T g_a_inverse = T(1.0) / g.a;
    const T f_a_by_g_a = f.a * g_a_inverse;
    Jet<T, N> h; h.a = f_a_by_g_a; CSLAM_JET_LOOP h.v[i] = (f.v[i] - f_a_by_g_a * g.v[i]) * g_a_inverse; return h; }
template <typename T, int N> inline Jet<T, N> operator/(T s, const Jet<T, N>& g) {
    const T minus_s_g_a_inverse2 = -s / (g.a * g.a);
    Jet<T, N> h; h.a = s / g.a; CSLAM_JET_LOOP h.v[i] = g.v[i] * minus_s_g_a_inverse2; return h; }
template <typename T, int N> inline Jet<T, N> operator/(const Jet<T, N>& f, T s) {
    const T s_inverse = T(1.0) / s;
    Jet<T, N> h; h.a = f.a * s_inverse; CSLAM_JET_LOOP h.v[i] = f.v[i] * s_inverse; return h; }

// comparisons look at the value only
#define CSLAM_JET_CMP(op)                                                                                   \
    template <typename T, int N> inline bool operator op(const Jet<T, N>& f, const Jet<T, N>& g) { return f.a op g.a; } \
    template <typename T, int N> inline bool operator op(const T& s, const Jet<T, N>& g) { return s op g.a; }          \
    template <typename T, int N> inline bool operator op(const Jet<T, N>& f, const T& s) { return f.a op s; }
CSLAM_JET_CMP(<)
CSLAM_JET_CMP(<=)
CSLAM_JET_CMP(>)
CSLAM_JET_CMP(>=)
CSLAM_JET_CMP(==)
CSLAM_JET_CMP(!=)
#undef CSLAM_JET_CMP

// elementary functions (found by ADL from the reference's unqualified calls)
template <typename T, int N> inline Jet<T, N> abs(const Jet<T, N>& f) { return f.a < T(0.0) ? -f : f; }
template <typename T, int N> inline Jet<T, N> sqrt(const Jet<T, N>& f) {        // d sqrt = df / (2 sqrt f)
    const T tmp = std::sqrt(f.a); const T two_a_inverse = T(1.0) / (T(2.0) * tmp);
    Jet<T, N> h; h.a = tmp; CSLAM_JET_LOOP h.v[i] = f.v[i] * two_a_inverse; return h; }
template <typename T, int N> inline Jet<T, N> cos(const Jet<T, N>& f) {         // -sin f df
    const T s = -std::sin(f.a); Jet<T, N> h; h.a = std::cos(f.a); CSLAM_JET_LOOP h.v[i] = s * f.v[i]; return h; }
template <typename T, int N> inline Jet<T, N> sin(const Jet<T, N>& f) {         // cos f df
    const T c = std::cos(f.a); Jet<T, N> h; h.a = std::sin(f.a); CSLAM_JET_LOOP h.v[i] = c * f.v[i]; return h; }
template <typename T, int N> inline Jet<T, N> acos(const Jet<T, N>& f) {        // -df / sqrt(1 - f^2)
    const T tmp = -T(1.0) / std::sqrt(T(1.0) - f.a * f.a);
    Jet<T, N> h; h.a = std::acos(f.a); CSLAM_JET_LOOP h.v[i] = tmp * f.v[i]; return h; }
template <typename T, int N> inline Jet<T, N> asin(const Jet<T, N>& f) {
    const T tmp = T(1.0) / std::sqrt(T(1.0) - f.a * f.a);
    Jet<T, N> h; h.a = std::asin(f.a); CSLAM_JET_LOOP h.v[i] = tmp * f.v[i]; return h; }
template <typename T, int N> inline Jet<T, N> log(const Jet<T, N>& f) {
    const T a_inverse = T(1.0) / f.a; Jet<T, N> h; h.a = std::log(f.a); CSLAM_JET_LOOP h.v[i] = f.v[i] * a_inverse; return h; }
template <typename T, int N> inline Jet<T, N> exp(const Jet<T, N>& f) {
    const T tmp = std::exp(f.a); Jet<T, N> h; h.a = tmp; CSLAM_JET_LOOP h.v[i] = tmp * f.v[i]; return h; }
// atan2(g, f): d = (f dg - g df) / (f^2 + g^2)
template <typename T, int N> inline Jet<T, N> atan2(const Jet<T, N>& g, const Jet<T, N>& f) {
    const T tmp = T(1.0) / (f.a * f.a + g.a * g.a);
    Jet<T, N> h; h.a = std::atan2(g.a, f.a); CSLAM_JET_LOOP h.v[i] = tmp * (-g.a * f.v[i] + f.a * g.v[i]); return h; }
// pow: d(f^g) = g f^(g-1) df + f^g log(f) dg   (Ceres 1.x form; f > 0 on every call the path makes)
template <typename T, int N> inline Jet<T, N> pow(const Jet<T, N>& f, const Jet<T, N>& g) {
    const T tmp1 = std::pow(f.a, g.a);
    const T tmp2 = g.a * std::pow(f.a, g.a - T(1.0));
    const T tmp3 = tmp1 * std::log(f.a);
    Jet<T, N> h; h.a = tmp1; CSLAM_JET_LOOP h.v[i] = tmp2 * f.v[i] + tmp3 * g.v[i]; return h; }
template <typename T, int N> inline Jet<T, N> pow(const Jet<T, N>& f, double g) {
    const T tmp = g * std::pow(f.a, g - T(1.0));
    Jet<T, N> h; h.a = std::pow(f.a, g); CSLAM_JET_LOOP h.v[i] = tmp * f.v[i]; return h; }
template <typename T, int N> inline Jet<T, N> pow(double f, const Jet<T, N>& g) {
    const T tmp = std::pow(f, g.a); const T l = tmp * std::log(f);
    Jet<T, N> h; h.a = tmp; CSLAM_JET_LOOP h.v[i] = l * g.v[i]; return h; }
template <typename T, int N> inline bool IsFinite(const Jet<T, N>& f) {
    if (!std::isfinite(f.a)) return false;
    CSLAM_JET_LOOP if (!std::isfinite(f.v[i])) return false;
    return true; }
template <typename T, int N> inline bool isfinite(const Jet<T, N>& f) { return IsFinite(f); }
#undef CSLAM_JET_LOOP

// ------------------------------------------------------------------------------------------------
class CostFunction {
   public:
    virtual ~CostFunction() {}
    // jacobians[i] (may be null) is num_residuals x parameter_block_sizes[i], row-major
    virtual bool Evaluate(double const* const* parameters, double* residuals, double** jacobians) const = 0;
    const std::vector<int>& parameter_block_sizes() const { return sizes_; }
    int num_residuals() const { return num_residuals_; }
    // stand-in only: the functor an AutoDiffCostFunction owns (for oracle/ref_driver's Problem facade)
    virtual const void* functor_ptr() const { return nullptr; }
    virtual const std::type_info& functor_type() const { return typeid(void); }

   protected:
    std::vector<int> sizes_;
    int num_residuals_ = 0;
};

namespace internal {
template <int... Ns> struct Sum;
template <> struct Sum<> { enum { value = 0 }; };
template <int N, int... Ns> struct Sum<N, Ns...> { enum { value = N + Sum<Ns...>::value }; };

template <typename Functor, typename J, int... I>
inline bool call(const Functor& f, J* const* blocks, J* out, std::integer_sequence<int, I...>) {
    return f(blocks[I]..., out);
}
}  // namespace internal

// One Jet of width sum(Ns...) per scalar parameter: block b's k-th coordinate is seeded with the
// unit partial at offset(b) + k; all blocks are differentiated in one functor call.
template <typename CostFunctor, int kNumResiduals, int... Ns>
class AutoDiffCostFunction : public CostFunction {
   public:
    explicit AutoDiffCostFunction(CostFunctor* functor) : functor_(functor) {
        num_residuals_ = kNumResiduals;
        sizes_ = {Ns...};
    }
    bool Evaluate(double const* const* parameters, double* residuals, double** jacobians) const override {
        constexpr int kBlocks = sizeof...(Ns);
        const int sizes[kBlocks] = {Ns...};
        if (!jacobians) {
            double* blocks[kBlocks];
            for (int b = 0; b < kBlocks; ++b) blocks[b] = const_cast<double*>(parameters[b]);
            return internal::call(*functor_, blocks, residuals, std::make_integer_sequence<int, kBlocks>());
        }
        constexpr int kWidth = internal::Sum<Ns...>::value;
        typedef Jet<double, kWidth> JetT;
        std::vector<JetT> x(kWidth), out(kNumResiduals);
        JetT* blocks[kBlocks];
        int off = 0;
        for (int b = 0; b < kBlocks; ++b) {
            blocks[b] = x.data() + off;
            for (int k = 0; k < sizes[b]; ++k) x[off + k] = JetT(parameters[b][k], off + k);
            off += sizes[b];
        }
        if (!internal::call(*functor_, blocks, out.data(), std::make_integer_sequence<int, kBlocks>())) return false;
        for (int r = 0; r < kNumResiduals; ++r) residuals[r] = out[r].a;
        off = 0;
        for (int b = 0; b < kBlocks; ++b) {
            if (jacobians[b])
                for (int r = 0; r < kNumResiduals; ++r)
                    for (int k = 0; k < sizes[b]; ++k) jacobians[b][r * sizes[b] + k] = out[r].v[off + k];
            off += sizes[b];
        }
        return true;
    }

    const void* functor_ptr() const override { return functor_.get(); }
    const std::type_info& functor_type() const override { return typeid(CostFunctor); }

   private:
    std::unique_ptr<CostFunctor> functor_;
};

class LocalParameterization {
   public:
    virtual ~LocalParameterization() {}
    virtual bool Plus(const double* x, const double* delta, double* x_plus_delta) const = 0;
    virtual bool ComputeJacobian(const double* x, double* jacobian) const = 0;  // GlobalSize x LocalSize, row-major
    virtual int GlobalSize() const = 0;
    virtual int LocalSize() const = 0;
};

// Plus(x, delta) = functor(x, delta); Jacobian = d functor / d delta at delta = 0, taken with Jets
// of width kGlobalSize + kLocalSize (x seeded too, its partials dropped), as AutoDiff does.
template <typename Functor, int kGlobalSize, int kLocalSize>
class AutoDiffLocalParameterization : public LocalParameterization {
   public:
    AutoDiffLocalParameterization() : functor_(new Functor()) {}
    explicit AutoDiffLocalParameterization(Functor* f) : functor_(f) {}
    bool Plus(const double* x, const double* delta, double* x_plus_delta) const override {
        return (*functor_)(x, delta, x_plus_delta);
    }
    bool ComputeJacobian(const double* x, double* jacobian) const override {
        typedef Jet<double, kGlobalSize + kLocalSize> JetT;
        JetT xj[kGlobalSize], dj[kLocalSize], out[kGlobalSize];
        for (int k = 0; k < kGlobalSize; ++k) xj[k] = JetT(x[k], k);
        for (int k = 0; k < kLocalSize; ++k) dj[k] = JetT(0.0, kGlobalSize + k);
        if (!(*functor_)(static_cast<const JetT*>(xj), static_cast<const JetT*>(dj), static_cast<JetT*>(out))) return false;
        for (int r = 0; r < kGlobalSize; ++r)
            for (int k = 0; k < kLocalSize; ++k) jacobian[r * kLocalSize + k] = out[r].v[kGlobalSize + k];
        return true;
    }
    int GlobalSize() const override { return kGlobalSize; }
    int LocalSize() const override { return kLocalSize; }

   private:
    std::unique_ptr<Functor> functor_;
};

}  // namespace ceres

namespace std {
template <typename T, int N>
struct numeric_limits<ceres::Jet<T, N> > {
    static constexpr bool is_specialized = true;
    static ceres::Jet<T, N> epsilon() { return ceres::Jet<T, N>(numeric_limits<T>::epsilon()); }
    static ceres::Jet<T, N> min() { return ceres::Jet<T, N>(numeric_limits<T>::min()); }
    static ceres::Jet<T, N> max() { return ceres::Jet<T, N>(numeric_limits<T>::max()); }
    static ceres::Jet<T, N> infinity() { return ceres::Jet<T, N>(numeric_limits<T>::infinity()); }
    static ceres::Jet<T, N> quiet_NaN() { return ceres::Jet<T, N>(numeric_limits<T>::quiet_NaN()); }
};
}  // namespace std

#endif  // CSLAM_REF_STANDIN_CERES_H
