// ORACLE — TEST INFRASTRUCTURE ONLY.
//
// CPU restatement of PointCloudAligner (src/ceres_slam/point_cloud_aligner.cpp): the N-point
// Kabsch alignment of compute_transformation (:12-62) and the 3-point RANSAC of
// compute_transformation_and_inliers (:64-136).  The index draws are the reference's:
// std::mt19937 seeded with 42 (:70-72) through std::uniform_int_distribution<uint>(0, n-1) — both
// libstdc++ algorithms are restated (variant 0: scaling + rejection, GCC <= 10; variant 1: Lemire,
// GCC >= 11) and tests/test_ransac.py pins variant 1 and the Mersenne twister against the
// std:: classes of the compiler in this image.
// PINNED: oracle/_ref compiles the reference's own point_cloud_aligner.cpp (unmodified) and tests/test_ref_pin.py
// checks this restatement against it (identical inlier index lists, transformations to 2e-13).
// Eigen::JacobiSVD is not available; the rotation U diag(1,1,det U det V) V^T is computed with a
// one-sided Jacobi SVD in long double (the result is unique whenever the two largest singular
// values are distinct and non-zero, so any accurate SVD gives the same rotation).
#pragma once
#include <cmath>
#include <cstdint>
#include <random>
#include <vector>

#include "functors.hpp"

namespace oracle {

inline uint32_t ransac_draw(std::mt19937& rng, uint32_t n, int variant) {
    if (variant == 0) {
        const unsigned long urngrange = 0xFFFFFFFFul, uerange = n;
        const unsigned long scaling = urngrange / uerange, past = uerange * scaling;
        unsigned long ret;
        do {
            ret = (unsigned long)(rng());
        } while (ret >= past);
        return uint32_t(ret / scaling);
    }
    uint64_t product = uint64_t(uint32_t(rng())) * uint64_t(n);
    uint32_t low = uint32_t(product);
    if (low < n) {
        const uint32_t threshold = uint32_t(0u - n) % n;
        while (low < threshold) {
            product = uint64_t(uint32_t(rng())) * uint64_t(n);
            low = uint32_t(product);
        }
    }
    return uint32_t(product >> 32);
}

// compute_transformation for any number of correspondences: T_1_0 as [t | R row-major]
inline void kabsch(const std::vector<const double*>& p0, const std::vector<const double*>& p1, double* T12) {
    const size_t n = p0.size();
    double pb[3] = {0, 0, 0}, qb[3] = {0, 0, 0};
    for (size_t i = 0; i < n; ++i)
        for (int r = 0; r < 3; ++r) pb[r] += p0[i][r];
    for (int r = 0; r < 3; ++r) pb[r] /= double(n);
    for (size_t i = 0; i < n; ++i)
        for (int r = 0; r < 3; ++r) qb[r] += p1[i][r];
    for (int r = 0; r < 3; ++r) qb[r] /= double(n);
    double W[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (size_t i = 0; i < n; ++i)
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) W[3 * r + c] += (p1[i][r] - qb[r]) * (p0[i][c] - pb[c]);
    for (double& w : W) w /= double(n);
    long double G[9], V[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    for (int i = 0; i < 9; ++i) G[i] = W[i];
    for (int sweep = 0; sweep < 60; ++sweep) {
        long double off = 0;
        for (int p = 0; p < 2; ++p)
            for (int q = p + 1; q < 3; ++q) {
                long double al = 0, be = 0, ga = 0;
                for (int r = 0; r < 3; ++r) {
                    al += G[3 * r + p] * G[3 * r + p];
                    be += G[3 * r + q] * G[3 * r + q];
                    ga += G[3 * r + p] * G[3 * r + q];
                }
                if (ga == 0 || al == 0 || be == 0) continue;
                off = std::max(off, std::fabs(ga) / std::sqrt(al * be));
                const long double zeta = (be - al) / (2 * ga);
                const long double t = (zeta >= 0 ? 1.0L : -1.0L) / (std::fabs(zeta) + std::sqrt(1 + zeta * zeta));
                const long double cs = 1 / std::sqrt(1 + t * t), sn = cs * t;
                for (int r = 0; r < 3; ++r) {
                    const long double gp = G[3 * r + p], gq = G[3 * r + q], vp = V[3 * r + p], vq = V[3 * r + q];
                    G[3 * r + p] = cs * gp - sn * gq;
                    G[3 * r + q] = sn * gp + cs * gq;
                    V[3 * r + p] = cs * vp - sn * vq;
                    V[3 * r + q] = sn * vp + cs * vq;
                }
            }
        if (off < 1e-19L) break;
    }
    long double s[3], U[9];
    for (int c = 0; c < 3; ++c) s[c] = std::sqrt(G[c] * G[c] + G[3 + c] * G[3 + c] + G[6 + c] * G[6 + c]);
    int k = 0;
    for (int c = 1; c < 3; ++c)
        if (s[c] < s[k]) k = c;
    const int a = (k + 1) % 3, b = (k + 2) % 3;
    for (int r = 0; r < 3; ++r) {
        U[3 * r + a] = G[3 * r + a] / s[a];
        U[3 * r + b] = G[3 * r + b] / s[b];
    }
    auto cross_into = [&](long double* M) {
        M[0 + k] = M[3 + a] * M[6 + b] - M[6 + a] * M[3 + b];
        M[3 + k] = M[6 + a] * M[0 + b] - M[0 + a] * M[6 + b];
        M[6 + k] = M[0 + a] * M[3 + b] - M[3 + a] * M[0 + b];
    };
    // third column of U: orthogonal complement; the sign that the determinant correction
    // diag(1, 1, det U det V) would undo is fixed by taking both third columns as cross products
    cross_into(U);
    cross_into(V);
    double* C = T12 + 3;
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) {
            long double v = 0;
            for (int m = 0; m < 3; ++m) v += U[3 * r + m] * V[3 * c + m];
            C[3 * r + c] = double(v);
        }
    for (int r = 0; r < 3; ++r) T12[r] = qb[r] - (C[3 * r] * pb[0] + C[3 * r + 1] * pb[1] + C[3 * r + 2] * pb[2]);
}

// compute_transformation_and_inliers for one pair; returns the inlier indices (ascending)
inline std::vector<uint32_t> ransac_align(const Camera& cam, const double* pts0, const double* pts1, uint32_t n,
                                          uint32_t num_iters, double thresh, int variant, double* T12) {
    for (int k = 0; k < 12; ++k) T12[k] = (k == 3 || k == 7 || k == 11) ? 1.0 : 0.0;
    std::vector<uint32_t> best, cur;
    if (n < 3) return best;
    std::mt19937 rng(42);
    for (uint32_t it = 0; it < num_iters; ++it) {
        uint32_t idx[3];
        idx[0] = ransac_draw(rng, n, variant);
        idx[1] = ransac_draw(rng, n, variant);
        while (idx[1] == idx[0]) idx[1] = ransac_draw(rng, n, variant);
        idx[2] = ransac_draw(rng, n, variant);
        while (idx[2] == idx[0] || idx[2] == idx[1]) idx[2] = ransac_draw(rng, n, variant);
        std::vector<const double*> a, b;
        for (int k = 0; k < 3; ++k) {
            a.push_back(pts0 + 3 * size_t(idx[k]));
            b.push_back(pts1 + 3 * size_t(idx[k]));
        }
        double T[12];
        kabsch(a, b, T);
        cur.clear();
        for (uint32_t i = 0; i < n; ++i) {
            const double* p = pts0 + 3 * size_t(i);
            const double x[3] = {T[3] * p[0] + T[4] * p[1] + T[5] * p[2] + T[0], T[6] * p[0] + T[7] * p[1] + T[8] * p[2] + T[1],
                                 T[9] * p[0] + T[10] * p[1] + T[11] * p[2] + T[2]};
            double z0[3], z1[3];
            camera_project(cam, x, z0);
            camera_project(cam, pts1 + 3 * size_t(i), z1);
            double e = 0;
            for (int k = 0; k < 3; ++k) e += (z1[k] - z0[k]) * (z1[k] - z0[k]);
            if (e < thresh) cur.push_back(i);
        }
        if (cur.size() > best.size()) {
            best = cur;
            std::memcpy(T12, T, sizeof(T));
        }
    }
    return best;
}

}  // namespace oracle
