// ORACLE — TEST INFRASTRUCTURE ONLY (parity PINNED: checked against the reference's own unmodified headers compiled
// in oracle/_ref, tests/test_ref_pin.py, 1e-14).
//
// CPU restatement, templated on the scalar (double or oracle::Jet<N>), of the reference's
// header-only geometry: SO(3) as row-major 3x3, SE(3) as the 12-vector [t | R row-major],
// Rodrigues exp with the first-order branch, atan2-form log, and the *decoupled* SE(3)
// exp/log.  Every function cites the reference lines whose behaviour (branch conditions,
// order of operations) it follows.  No Eigen: the container has none.
#pragma once
#include <limits>

#include "jet.hpp"

namespace oracle {

// utils.hpp:28-31 — templated fabs with `a >= 0 ? a : -a`
template <class T>
inline T t_fabs(const T& a) {
    return (a >= T(0.0)) ? a : -a;
}
// utils.hpp:16-19 / :22-25 — ties return the FIRST argument
template <class T>
inline T t_fmax(const T& a, const T& b) {
    return (a >= b) ? a : b;
}
template <class T>
inline T t_fmin(const T& a, const T& b) {
    return (a <= b) ? a : b;
}

template <class T>
inline T dot3(const T* a, const T* b) {
    return a[0] * b[0] + a[1] * b[1] + a[2] * b[2];
}
template <class T>
inline T norm3(const T* a) {
    return sqrt(dot3(a, a));
}
// y = M x, M row-major 3x3
template <class T>
inline void mat3_vec(const T* M, const T* x, T* y) {
    for (int i = 0; i < 3; ++i) y[i] = M[3 * i] * x[0] + M[3 * i + 1] * x[1] + M[3 * i + 2] * x[2];
}
template <class T>
inline void mat3_mul(const T* A, const T* B, T* C) {
    T tmp[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            tmp[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
    for (int i = 0; i < 9; ++i) C[i] = tmp[i];
}
template <class T>
inline void mat3_transpose(const T* A, T* At) {
    T tmp[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) tmp[3 * j + i] = A[3 * i + j];
    for (int i = 0; i < 9; ++i) At[i] = tmp[i];
}

// so3group.hpp:248-254
template <class T>
inline void so3_wedge(const T* phi, T* Phi) {
    Phi[0] = T(0.0);
    Phi[1] = -phi[2];
    Phi[2] = phi[1];
    Phi[3] = phi[2];
    Phi[4] = T(0.0);
    Phi[5] = -phi[0];
    Phi[6] = -phi[1];
    Phi[7] = phi[0];
    Phi[8] = T(0.0);
}
// so3group.hpp:260-265
template <class T>
inline void so3_vee(const T* Phi, T* phi) {
    phi[0] = T(0.5) * (Phi[7] - Phi[5]);
    phi[1] = T(0.5) * (Phi[2] - Phi[6]);
    phi[2] = T(0.5) * (Phi[3] - Phi[1]);
}

// so3group.hpp:273-291 — Rodrigues; `angle <= eps` takes I + wedge(phi)
template <class T>
inline void so3_exp(const T* phi, T* R) {
    T angle = norm3(phi);
    if (angle <= std::numeric_limits<double>::epsilon()) {
        T W[9];
        so3_wedge(phi, W);
        for (int i = 0; i < 9; ++i) R[i] = W[i];
        R[0] = R[0] + T(1.0);
        R[4] = R[4] + T(1.0);
        R[8] = R[8] + T(1.0);
        return;
    }
    T axis[3] = {phi[0] / angle, phi[1] / angle, phi[2] / angle};
    T cp = cos(angle);
    T sp = sin(angle);
    T W[9];
    so3_wedge(axis, W);
    T omc = T(1.0) - cp;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            T id = (i == j) ? T(1.0) : T(0.0);
            R[3 * i + j] = cp * id + omc * axis[i] * axis[j] + sp * W[3 * i + j];
        }
}

// so3group.hpp:299-349 — atan2 form, first-order branch on |angle| <= eps (templated fabs)
template <class T>
inline void so3_log(const T* C, T* phi) {
    T axis[3];
    axis[0] = C[7] - C[5];
    axis[1] = C[2] - C[6];
    axis[2] = C[3] - C[1];
    T sin_angle = T(0.5) * norm3(axis);
    T cos_angle = T(0.5) * ((C[0] + C[4] + C[8]) - T(1.0));
    T angle = atan2(sin_angle, cos_angle);
    if (t_fabs(angle) <= std::numeric_limits<double>::epsilon()) {
        T Phi[9];
        for (int i = 0; i < 9; ++i) Phi[i] = C[i];
        Phi[0] = Phi[0] - T(1.0);
        Phi[4] = Phi[4] - T(1.0);
        Phi[8] = Phi[8] - T(1.0);
        so3_vee(Phi, phi);
        return;
    }
    // 0.5 * angle * axis / sin_angle, evaluated left to right as written
    for (int i = 0; i < 3; ++i) phi[i] = T(0.5) * angle * axis[i] / sin_angle;
}

// SE(3) stored as 12 scalars [t(3) | R(9) row-major] — se3group.hpp:115-118, :479, :543
template <class T>
struct SE3 {
    T d[12];
    T* t() { return d; }
    const T* t() const { return d; }
    T* R() { return d + 3; }
    const T* R() const { return d + 3; }
    static SE3 identity() {
        SE3 I;
        for (int i = 0; i < 12; ++i) I.d[i] = T(0.0);
        I.d[3] = T(1.0);
        I.d[7] = T(1.0);
        I.d[11] = T(1.0);
        return I;
    }
    template <class U>
    static SE3 from(const U* src) {
        SE3 X;
        for (int i = 0; i < 12; ++i) X.d[i] = T(src[i]);
        return X;
    }
};

// se3group.hpp:176-183 — (R1 R2, R1 t2 + t1)
template <class T>
inline SE3<T> se3_mul(const SE3<T>& A, const SE3<T>& B) {
    SE3<T> C;
    mat3_mul(A.R(), B.R(), C.R());
    T Rt[3];
    mat3_vec(A.R(), B.t(), Rt);
    for (int i = 0; i < 3; ++i) C.t()[i] = Rt[i] + A.t()[i];
    return C;
}
// se3group.hpp:152-158 — (R^T, -(R^T t))
template <class T>
inline SE3<T> se3_inverse(const SE3<T>& A) {
    SE3<T> C;
    mat3_transpose(A.R(), C.R());
    T Rt[3];
    mat3_vec(C.R(), A.t(), Rt);
    for (int i = 0; i < 3; ++i) C.t()[i] = -Rt[i];
    return C;
}
// se3group.hpp:191-193 — point: R p + t
template <class T>
inline void se3_transform_point(const SE3<T>& A, const T* p, T* out) {
    T Rp[3];
    mat3_vec(A.R(), p, Rp);
    for (int i = 0; i < 3; ++i) out[i] = Rp[i] + A.t()[i];
}
// se3group.hpp:242-244 — vector: R v
template <class T>
inline void se3_transform_vector(const SE3<T>& A, const T* v, T* out) {
    mat3_vec(A.R(), v, out);
}
// se3group.hpp:161-170 — [[R, [t]x R],[0, R]] row-major 6x6
template <class T>
inline void se3_adjoint(const SE3<T>& A, T* Ad) {
    T W[9], WR[9];
    so3_wedge(A.t(), W);
    mat3_mul(W, A.R(), WR);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            Ad[6 * i + j] = A.R()[3 * i + j];
            Ad[6 * i + 3 + j] = WR[3 * i + j];
            Ad[6 * (i + 3) + j] = T(0.0);
            Ad[6 * (i + 3) + 3 + j] = A.R()[3 * i + j];
        }
}
// se3group.hpp:323-325 — decoupled: (rho, Exp(phi))
template <class T>
inline SE3<T> se3_exp(const T* xi) {
    SE3<T> X;
    for (int i = 0; i < 3; ++i) X.t()[i] = xi[i];
    so3_exp(xi + 3, X.R());
    return X;
}
// se3group.hpp:337-342 — decoupled: [t ; Log(R)]
template <class T>
inline void se3_log(const SE3<T>& X, T* xi) {
    for (int i = 0; i < 3; ++i) xi[i] = X.t()[i];
    so3_log(X.R(), xi + 3);
}

}  // namespace oracle
