// Host-side pieces the restated drivers share: the reference's CSV rows, small dense helpers that
// stand in for the few Eigen calls the drivers make around the solve
// (SelfAdjointEigenSolver::operatorInverseSqrt, tests/dataset_vo.cpp:29-32), SE(3) on the 12-double
// pose blocks, and `compute_initial_guess` (src/ceres_slam/dataset_problem.cpp:179-270,
// dataset_problem_sun.cpp:248-355, dataset_problem_phong.cpp:248-390) over the batched RANSAC
// entry point of the C ABI.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstring>
#include <fstream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <unordered_set>
#include <vector>

#include "../../include/cslam_b200.h"

namespace cslam_b200 {

inline std::vector<double> parse_csv_line(const std::string& line) {
    std::vector<double> v;
    std::stringstream ss(line);
    std::string tok;
    while (std::getline(ss, tok, ',')) v.push_back(std::stod(tok));
    return v;
}

// [t | R row-major] from a 4x4 row-major matrix (se3group.hpp:115-118)
inline void pose_from_matrix16(const std::vector<double>& m, double* P) {
    for (int r = 0; r < 3; ++r) {
        P[r] = m.at(4 * r + 3);
        for (int c = 0; c < 3; ++c) P[3 + 3 * r + c] = m.at(4 * r + c);
    }
}
// C = A * B (se3group.hpp:176-183)
inline void pose_mul(const double* A, const double* B, double* C) {
    double out[12];
    for (int r = 0; r < 3; ++r) {
        out[r] = A[3 + 3 * r] * B[0] + A[4 + 3 * r] * B[1] + A[5 + 3 * r] * B[2] + A[r];
        for (int c = 0; c < 3; ++c)
            out[3 + 3 * r + c] = A[3 + 3 * r] * B[3 + c] + A[4 + 3 * r] * B[6 + c] + A[5 + 3 * r] * B[9 + c];
    }
    std::memcpy(C, out, sizeof(out));
}
// x_g = T^-1 x_c with T^-1 = (-R^T t, R^T) (se3group.hpp:152-157)
inline void pose_inverse_apply(const double* P, const double* xc, bool is_vector, double* xg) {
    double ti[3];
    for (int c = 0; c < 3; ++c) ti[c] = is_vector ? 0.0 : -(P[3 + c] * P[0] + P[6 + c] * P[1] + P[9 + c] * P[2]);
    for (int c = 0; c < 3; ++c) xg[c] = P[3 + c] * xc[0] + P[6 + c] * xc[1] + P[9 + c] * xc[2] + ti[c];
}
// StereoCamera::triangulate (stereo_camera.hpp:112-120)
inline void triangulate(const double intr[5], const double* uvd, double* pc) {
    const double b_over_d = intr[4] / uvd[2], fu_over_fv = intr[0] / intr[1];
    pc[0] = (uvd[0] - intr[2]) * b_over_d;
    pc[1] = (uvd[1] - intr[3]) * b_over_d * fu_over_fv;
    pc[2] = intr[0] * b_over_d;
}

// A^(-1/2) of a symmetric positive definite n x n matrix (row-major), by cyclic Jacobi: what
// Eigen::SelfAdjointEigenSolver(A).operatorInverseSqrt() returns.
inline void sym_inverse_sqrt(const double* A, int n, double* out) {
    std::vector<double> M(A, A + n * n), V(n * n, 0.0);
    for (int i = 0; i < n; ++i) V[i * n + i] = 1.0;
    for (int sweep = 0; sweep < 60; ++sweep) {
        double off = 0;
        for (int p = 0; p < n; ++p)
            for (int q = p + 1; q < n; ++q) off += M[p * n + q] * M[p * n + q];
        if (off < 1e-300) break;
        for (int p = 0; p < n; ++p)
            for (int q = p + 1; q < n; ++q) {
                const double apq = M[p * n + q];
                if (apq == 0.0) continue;
                const double theta = (M[q * n + q] - M[p * n + p]) / (2.0 * apq);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(1.0 + theta * theta));
                const double c = 1.0 / std::sqrt(1.0 + t * t), s = c * t;
                for (int k = 0; k < n; ++k) {
                    const double mkp = M[k * n + p], mkq = M[k * n + q];
                    M[k * n + p] = c * mkp - s * mkq;
                    M[k * n + q] = s * mkp + c * mkq;
                }
                for (int k = 0; k < n; ++k) {
                    const double mpk = M[p * n + k], mqk = M[q * n + k];
                    M[p * n + k] = c * mpk - s * mqk;
                    M[q * n + k] = s * mpk + c * mqk;
                }
                for (int k = 0; k < n; ++k) {
                    const double vkp = V[k * n + p], vkq = V[k * n + q];
                    V[k * n + p] = c * vkp - s * vkq;
                    V[k * n + q] = s * vkp + c * vkq;
                }
            }
    }
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) {
            double s = 0;
            for (int k = 0; k < n; ++k) s += V[i * n + k] * V[j * n + k] / std::sqrt(M[k * n + k]);
            out[i * n + j] = s;
        }
}

// What every DatasetProblem* keeps per observation, in file order
struct ObservationTable {
    std::vector<unsigned> k, j;                    // state id / timestamp run, point id
    std::vector<double> uvd;                       // 3 per observation
    std::vector<std::vector<unsigned>> state_obs;  // obs_indices_at_state
};

struct InitialGuessStats {
    unsigned pairs = 0, failed_pair = 0;
    bool ok = true;
};

// compute_initial_guess(k1, k2).  `poses` 12 per state, `points` 3 per point, `initialized` one
// flag per point; `on_init(i_obs_km1, j, pose_km1)` lets the Phong variant initialise the normal
// and material of a vertex next to its position (dataset_problem_phong.cpp:352-379).
// `require_three_inliers`: the sun variant gives up (returns ok = false) when a pair has fewer
// than 3 inliers (dataset_problem_sun.cpp:323-326); the others carry on.
// All pose pairs of the call go to the GPU as one batch; the chaining is sequential afterwards.
// The three stages are separate functions so that a batch runner can put the pairs of MANY jobs into one
// cslam_ransac_align call (ba_all_b200): matches -> alignment (GPU) -> chaining.
struct InitialGuessPairs {
    unsigned k1 = 0, k2 = 0, n_pairs = 0;
    std::vector<uint32_t> offsets;   // n_pairs + 1, in correspondences
    std::vector<double> p0, p1;      // triangulated points of the two states, 3 per correspondence
    std::vector<unsigned> idx0, jid; // per correspondence: observation index in state k-1, point id
};

inline InitialGuessPairs initial_guess_pairs(const ObservationTable& obs, const double intr[5], unsigned num_states, unsigned k1,
                                             unsigned k2) {
    InitialGuessPairs q;
    if (k1 >= k2) {
        k1 = 0;
        k2 = num_states;
    }
    q.k1 = k1;
    q.k2 = k2;
    if (k2 - k1 < 2) return q;
    q.n_pairs = k2 - k1 - 1;
    q.offsets.assign(q.n_pairs + 1, 0);
    for (unsigned k = k1 + 1; k < k2; ++k) {
        // reciprocal matches, each list in its own order, paired by position (:207-229)
        std::vector<unsigned> a = obs.state_obs[k - 1], b = obs.state_obs[k];
        std::unordered_set<unsigned> ids_a, ids_b;
        for (unsigned i : a) ids_a.insert(obs.j[i]);
        for (unsigned i : b) ids_b.insert(obs.j[i]);
        std::vector<unsigned> ka, kb;
        for (unsigned i : a)
            if (ids_b.count(obs.j[i])) ka.push_back(i);
        for (unsigned i : b)
            if (ids_a.count(obs.j[i])) kb.push_back(i);
        const size_t n = std::min(ka.size(), kb.size());
        for (size_t i = 0; i < n; ++i) {
            double x[3];
            triangulate(intr, &obs.uvd[3 * size_t(ka[i])], x);
            q.p0.insert(q.p0.end(), x, x + 3);
            triangulate(intr, &obs.uvd[3 * size_t(kb[i])], x);
            q.p1.insert(q.p1.end(), x, x + 3);
            q.idx0.push_back(ka[i]);
            q.jid.push_back(obs.j[ka[i]]);
        }
        q.offsets[k - k1] = uint32_t(q.p0.size() / 3);
    }
    return q;
}

// T: 12 per pair, inl: one flag per correspondence, cnt: inliers per pair — this job's slices of the alignment's output
template <class OnInit>
inline InitialGuessStats initial_guess_chain(const InitialGuessPairs& q, const double* T, const uint8_t* inl, const uint32_t* cnt,
                                             bool require_three_inliers, std::vector<double>& poses, std::vector<double>& points,
                                             std::vector<char>& initialized, OnInit on_init) {
    InitialGuessStats st;
    st.pairs = q.n_pairs;
    for (unsigned k = q.k1 + 1; k < q.k2; ++k) {
        const unsigned p = k - q.k1 - 1;
        if (require_three_inliers && cnt[p] < 3) {
            st.ok = false;
            st.failed_pair = k;
            return st;
        }
        const double* Pm1 = &poses[12 * size_t(k - 1)];
        pose_mul(&T[12 * size_t(p)], Pm1, &poses[12 * size_t(k)]);  // poses[k] = T_k_km1 * poses[k-1]
        for (uint32_t c = q.offsets[p]; c < q.offsets[p + 1]; ++c) {
            if (!inl[c] || initialized[q.jid[c]]) continue;
            pose_inverse_apply(Pm1, &q.p0[3 * size_t(c)], false, &points[3 * size_t(q.jid[c])]);
            initialized[q.jid[c]] = 1;
            on_init(q.idx0[c], q.jid[c], Pm1, c - q.offsets[p]);
        }
    }
    return st;
}

template <class OnInit>
inline InitialGuessStats compute_initial_guess(const ObservationTable& obs, const double intr[5], unsigned num_states,
                                               unsigned k1, unsigned k2, double thresh, bool require_three_inliers,
                                               std::vector<double>& poses, std::vector<double>& points,
                                               std::vector<char>& initialized, OnInit on_init, int rng_variant = 0) {
    InitialGuessPairs q = initial_guess_pairs(obs, intr, num_states, k1, k2);
    if (q.n_pairs == 0) return InitialGuessStats();
    if (q.p0.empty()) {
        q.p0.assign(3, 0.0);
        q.p1.assign(3, 0.0);
    }
    std::vector<double> T(12 * size_t(q.n_pairs));
    std::vector<uint8_t> inl(std::max<size_t>(q.offsets[q.n_pairs], 1));
    std::vector<uint32_t> cnt(q.n_pairs);
    if (cslam_ransac_align(0, q.n_pairs, q.offsets.data(), q.p0.data(), q.p1.data(), intr, 400, thresh, rng_variant, T.data(),
                           inl.data(), cnt.data()) != CSLAM_OK)
        throw std::runtime_error("cslam_ransac_align failed (no CUDA device?)");
    return initial_guess_chain(q, T.data(), inl.data(), cnt.data(), require_three_inliers, poses, points, initialized, on_init);
}

// The same for several independent jobs that share the intrinsics: the pairs of all of them in ONE alignment call.
struct InitialGuessJob {
    const ObservationTable* obs;
    unsigned num_states, k1, k2;
    std::vector<double>* poses;
    std::vector<double>* points;
    std::vector<char>* initialized;
    InitialGuessStats stats;
};
inline void compute_initial_guess_batch(std::vector<InitialGuessJob>& jobs, const double intr[5], double thresh,
                                        bool require_three_inliers, int rng_variant = 0) {
    std::vector<InitialGuessPairs> qs;
    std::vector<uint32_t> offsets(1, 0);
    std::vector<double> p0, p1;
    for (auto& j : jobs) {
        qs.push_back(initial_guess_pairs(*j.obs, intr, j.num_states, j.k1, j.k2));
        const InitialGuessPairs& q = qs.back();
        const uint32_t base = uint32_t(p0.size() / 3);
        for (unsigned p = 0; p < q.n_pairs; ++p) offsets.push_back(base + q.offsets[p + 1]);
        p0.insert(p0.end(), q.p0.begin(), q.p0.end());
        p1.insert(p1.end(), q.p1.begin(), q.p1.end());
    }
    const uint32_t n_pairs = uint32_t(offsets.size() - 1);
    if (n_pairs == 0) return;
    if (p0.empty()) {
        p0.assign(3, 0.0);
        p1.assign(3, 0.0);
    }
    std::vector<double> T(12 * size_t(n_pairs));
    std::vector<uint8_t> inl(std::max<size_t>(offsets[n_pairs], 1));
    std::vector<uint32_t> cnt(n_pairs);
    if (cslam_ransac_align(0, n_pairs, offsets.data(), p0.data(), p1.data(), intr, 400, thresh, rng_variant, T.data(), inl.data(),
                           cnt.data()) != CSLAM_OK)
        throw std::runtime_error("cslam_ransac_align failed (no CUDA device?)");
    uint32_t pair0 = 0;
    for (size_t i = 0; i < jobs.size(); ++i) {
        const InitialGuessPairs& q = qs[i];
        if (q.n_pairs == 0) continue;
        jobs[i].stats = initial_guess_chain(q, &T[12 * size_t(pair0)], &inl[offsets[pair0]], &cnt[pair0], require_three_inliers,
                                            *jobs[i].poses, *jobs[i].points, *jobs[i].initialized,
                                            [](unsigned, unsigned, const double*, unsigned) {});
        pair0 += q.n_pairs;
    }
}

inline void write_poses_csv(const std::string& path, const std::vector<double>& poses, unsigned num_states) {
    // <stem>_poses.csv, 16 values per row (dataset_problem.cpp:139-150); full precision here — the
    // reference prints 4 significant digits (utils.hpp:34)
    std::ofstream out(path);
    out << "T_00, T_01, T_02, T_03,T_10, T_11, T_12, T_13,T_20, T_21, T_22, T_23,T_30, T_31, T_32, T_33\n";
    out.precision(17);
    for (unsigned s = 0; s < num_states; ++s) {
        const double* P = &poses[12 * size_t(s)];
        for (int r = 0; r < 3; ++r) out << P[3 + 3 * r] << "," << P[4 + 3 * r] << "," << P[5 + 3 * r] << "," << P[r] << ",";
        out << "0,0,0,1\n";
    }
}

inline std::string file_stem(const std::string& filename) { return filename.substr(0, filename.find('.')); }  // split(filename, '.').at(0)

}  // namespace cslam_b200
