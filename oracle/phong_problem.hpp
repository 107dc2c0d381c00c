// ORACLE — TEST INFRASTRUCTURE ONLY (residual blocks PINNED via oracle/_ref; the solver rules below are Ceres' and
// Ceres is absent: that part is parity UNPINNED).
//
// CPU restatement of the problem `solveWindow` of tests/dataset_ba_phong.cpp:26-255 hands to Ceres
// when lighting is enabled (default / --dirlight, the joint solve of stage 3) and of what Ceres then
// does with it.  Parameter blocks: pose (SE3Perturbation, :72), vertex position (:67), vertex normal
// (UnitVectorPerturbation, :191-194), material [ka, ks, alpha] shared by every vertex of a material
// (:117-121, bounds :143-172), texture kd shared the same way (:122, bounds :177-181), light
// position (:123) or direction (UnitVectorPerturbation, :199-203).  Residual blocks per observation:
// StereoReprojectionError (:59-67), IntensityError{Point,Directional}Light (:103-131), NormalError
// (:183-190).  Jacobians come from Jets chained with the autodiff Plus Jacobians, like
// AutoDiffCostFunction + AutoDiffLocalParameterization.
//
// Solver: the trust-region Levenberg-Marquardt loop of problem.hpp (same rules, same option
// names), plus the two things Ceres adds for a bounded problem [Ceres 1.x, from memory]:
//   * ParameterBlock::Plus projects x (+) delta onto the box, so every candidate, the initial
//     point (IterationZero) and the gradient-norm point x (+) (-g) are projected;
//   * TrustRegionMinimizer::DoLineSearch: an Armijo search along the trust-region step starting at
//     step 1 (sufficient decrease 1e-4, contraction to [1e-3, 0.6] of the step, 20 iterations,
//     min step 1e-9).  DEVIATION, stated: Ceres interpolates with a cubic that also uses the
//     directional derivative at the trial point; this restatement uses the quadratic through
//     phi(0), phi'(0), phi(step).  It only matters on steps whose projected candidate fails the
//     Armijo test at step 1.
// The linear solve eliminates each vertex's (position, normal) 6x6 block exactly and factors the
// dense reduced system over [free poses | materials | textures | light] with Cholesky — an exact
// solve of the normal equations, i.e. what SPARSE_SCHUR / SPARSE_NORMAL_CHOLESKY return up to
// rounding.  trust_region_strategy = 1: Ceres' DoglegStrategy (what the driver sets, :88-89), the same
// restatement as problem.hpp's over [poses | vertices | shared blocks].
#pragma once
#include "problem.hpp"

namespace oracle {

struct PhongState {
    std::vector<double> poses, pos, nrm, g;  // g = [materials 3 n_mat | textures n_tex | light 3]
};

class PhongProblem {
   public:
    Camera camera{1, 1, 0, 0, 1};
    double* poses = nullptr;
    int n_poses = 0;
    std::vector<uint8_t> pose_const;
    double* positions = nullptr;
    double* normals = nullptr;
    int n_vertices = 0;
    std::vector<uint32_t> v_mat, v_tex;
    double* materials = nullptr;
    int n_mat = 0;
    double* textures = nullptr;
    int n_tex = 0;
    double* light = nullptr;
    bool directional = false;
    double mat_lo[3], mat_hi[3], tex_lo, tex_hi;
    bool bounded = false;
    // SetParameterBlockConstant on every vertex position (stage 2 of --multistage,
    // dataset_ba_phong.cpp:209-220): the position columns disappear, and a stereo block whose pose is
    // constant too has no variable left — Ceres removes it from the reduced program and carries its
    // cost as `fixed_cost`, which is reported but takes no part in the minimiser's decisions.
    bool hold_positions = false;
    double fixed_cost = 0.0;

    std::vector<uint32_t> cam, vtx;
    std::vector<double> uvd, intensity, normal_obs;
    double W[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, Wn[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    double int_stiffness = 1.0;
    std::string error;

    PhongProblem() {
        const double inf = std::numeric_limits<double>::infinity();
        for (int k = 0; k < 3; ++k) mat_lo[k] = -inf, mat_hi[k] = inf;
        tex_lo = -inf;
        tex_hi = inf;
    }

    size_t n_obs() const { return cam.size(); }
    int n_g() const { return 3 * n_mat + n_tex + 3; }
    int g_tex0() const { return 3 * n_mat; }
    int g_light0() const { return 3 * n_mat + n_tex; }

    // ---- structure -------------------------------------------------------------------------------
    struct Structure {
        std::vector<int> cam_free, free_cams, v_active, active_v;
        std::vector<size_t> v_ptr;
        std::vector<uint32_t> v_obs;
        std::vector<uint8_t> g_used;
    };
    void build_structure(Structure& st) const {
        st.cam_free.assign(n_poses, -1);
        std::vector<uint8_t> used(n_poses, 0);
        for (uint32_t c : cam) used[c] = 1;
        for (int k = 0; k < n_poses; ++k)
            if (used[k] && !pose_const[k]) {
                st.cam_free[k] = int(st.free_cams.size());
                st.free_cams.push_back(k);
            }
        st.v_active.assign(n_vertices, -1);
        std::vector<size_t> cnt(n_vertices, 0);
        for (uint32_t j : vtx) cnt[j]++;
        for (int j = 0; j < n_vertices; ++j)
            if (cnt[j]) {
                st.v_active[j] = int(st.active_v.size());
                st.active_v.push_back(j);
            }
        const size_t na = st.active_v.size();
        st.v_ptr.assign(na + 1, 0);
        for (size_t a = 0; a < na; ++a) st.v_ptr[a + 1] = st.v_ptr[a] + cnt[st.active_v[a]];
        st.v_obs.resize(n_obs());
        std::vector<size_t> fill(st.v_ptr.begin(), st.v_ptr.end() - 1);
        for (size_t i = 0; i < n_obs(); ++i) st.v_obs[fill[st.v_active[vtx[i]]]++] = uint32_t(i);
        st.g_used.assign(n_g(), 0);
        for (int j : st.active_v) {
            for (int k = 0; k < 3; ++k) st.g_used[3 * v_mat[j] + k] = 1;
            st.g_used[g_tex0() + v_tex[j]] = 1;
        }
        if (na)
            for (int k = 0; k < 3; ++k) st.g_used[g_light0() + k] = 1;
    }
    void gidx_of(int j, int* gi) const {
        for (int k = 0; k < 3; ++k) gi[k] = 3 * int(v_mat[j]) + k;
        gi[3] = g_tex0() + int(v_tex[j]);
        for (int k = 0; k < 3; ++k) gi[4 + k] = g_light0() + k;
    }

    // ---- evaluation ------------------------------------------------------------------------------
    // Per observation, tangent coordinates: r 7 = [stereo 3 | intensity 1 | normal 3],
    // Ac 7x6 (pose), Av 7x6 (position 3 | normal 3), ag 7 (material 3 | texture 1 | light 3; the
    // intensity row is the only one that sees them).
    struct Lin {
        std::vector<double> r, Ac, Av, ag;
        double cost = 0;
    };
    bool evaluate(const PhongState& x, bool jac, Lin& L, int nthreads) const {
        const size_t n = n_obs();
        L.r.assign(7 * n, 0.0);
        std::vector<double> Ppose, Pn;
        double Pl[9];
        if (jac) {
            L.Ac.assign(42 * n, 0.0);
            L.Av.assign(42 * n, 0.0);
            L.ag.assign(7 * n, 0.0);
            Ppose.resize(72 * size_t(n_poses));
            SE3Perturbation sp;
            for (int k = 0; k < n_poses; ++k) autodiff_plus_jacobian(sp, &x.poses[12 * size_t(k)], &Ppose[72 * size_t(k)]);
            Pn.resize(9 * size_t(n_vertices));
            UnitVectorPerturbation up;
            for (int j = 0; j < n_vertices; ++j) autodiff_plus_jacobian(up, &x.nrm[3 * size_t(j)], &Pn[9 * size_t(j)]);
            if (directional) autodiff_plus_jacobian(up, &x.g[g_light0()], Pl);
        }
        std::vector<double> partial(std::max(1, nthreads), 0.0);
        std::atomic<bool> ok{true};
        parallel_for(n, nthreads, [&](size_t b, size_t e, int tid) {
            double c = 0;
            for (size_t i = b; i < e; ++i) {
                const size_t k = cam[i], j = vtx[i];
                const double* pose = &x.poses[12 * k];
                const double* pos = &x.pos[3 * j];
                const double* nr = &x.nrm[3 * j];
                const double* ph = &x.g[3 * size_t(v_mat[j])];
                const double* tx = &x.g[g_tex0() + v_tex[j]];
                const double* lt = &x.g[g_light0()];
                StereoReprojectionError fs;
                fs.camera = camera;
                for (int q = 0; q < 3; ++q) fs.observation[q] = uvd[3 * i + q];
                std::memcpy(fs.stiffness, W, sizeof(W));
                IntensityError fi;
                fi.colour = intensity[i];
                fi.stiffness = int_stiffness;
                fi.directional = directional;
                NormalError fn;
                for (int q = 0; q < 3; ++q) fn.obs_normal_c[q] = normal_obs[3 * i + q];
                std::memcpy(fn.stiffness, Wn, sizeof(Wn));
                const double* ps[2] = {pose, pos};
                const double* pi[6] = {pose, pos, nr, ph, tx, lt};
                const double* pn[2] = {pose, nr};
                double* r = &L.r[7 * i];
                if (!jac) {
                    if (!eval_cost(fs, ps, r) || !eval_cost(fi, pi, r + 3) || !eval_cost(fn, pn, r + 4)) ok = false;
                } else {
                    double Ja_s[36], Jp_s[9], Ja_i[12], Jp_i[3], Jn_i[3], Jm_i[3], Jt_i[1], Jl_i[3], Ja_n[36], Jn_n[9];
                    double* js[2] = {Ja_s, Jp_s};
                    double* ji[6] = {Ja_i, Jp_i, Jn_i, Jm_i, Jt_i, Jl_i};
                    double* jn[2] = {Ja_n, Jn_n};
                    if (!autodiff_cost(fs, ps, r, js) || !autodiff_cost(fi, pi, r + 3, ji) ||
                        !autodiff_cost(fn, pn, r + 4, jn))
                        ok = false;
                    const double* P = &Ppose[72 * k];
                    const double* Q = &Pn[9 * j];
                    double* Ac = &L.Ac[42 * i];
                    double* Av = &L.Av[42 * i];
                    double* ag = &L.ag[7 * i];
                    auto pose_row = [&](const double* Ja, double* out) {
                        for (int cc = 0; cc < 6; ++cc) {
                            double s = 0;
                            for (int q = 0; q < 12; ++q) s += Ja[q] * P[6 * q + cc];
                            out[cc] = s;
                        }
                    };
                    auto unit_row = [&](const double* Jn, const double* Pm, double* out) {
                        for (int cc = 0; cc < 3; ++cc) out[cc] = Jn[0] * Pm[cc] + Jn[1] * Pm[3 + cc] + Jn[2] * Pm[6 + cc];
                    };
                    for (int rr = 0; rr < 3; ++rr) {
                        pose_row(Ja_s + 12 * rr, Ac + 6 * rr);
                        for (int cc = 0; cc < 3; ++cc) Av[6 * rr + cc] = Jp_s[3 * rr + cc];
                        pose_row(Ja_n + 12 * rr, Ac + 6 * (4 + rr));
                        unit_row(Jn_n + 3 * rr, Q, Av + 6 * (4 + rr) + 3);
                    }
                    pose_row(Ja_i, Ac + 18);
                    for (int cc = 0; cc < 3; ++cc) Av[18 + cc] = Jp_i[cc];
                    unit_row(Jn_i, Q, Av + 18 + 3);
                    for (int cc = 0; cc < 3; ++cc) ag[cc] = Jm_i[cc];
                    ag[3] = Jt_i[0];
                    if (directional)
                        unit_row(Jl_i, Pl, ag + 4);
                    else
                        for (int cc = 0; cc < 3; ++cc) ag[4 + cc] = Jl_i[cc];
                }
                if (hold_positions) {
                    if (jac)
                        for (int rr = 0; rr < 7; ++rr)
                            for (int cc = 0; cc < 3; ++cc) L.Av[42 * i + 6 * rr + cc] = 0.0;
                    if (pose_const[k]) r[0] = r[1] = r[2] = 0.0;  // dropped block (its cost is fixed_cost)
                }
                for (int q = 0; q < 7; ++q) c += 0.5 * r[q] * r[q];
            }
            partial[tid] += c;
        });
        L.cost = 0;
        for (double c : partial) L.cost += c;
        return ok && std::isfinite(L.cost);
    }

    // Evaluator::Plus with the projection of ParameterBlock::Plus.  d = [poses 6 nf | vertices 6 na | g]
    void plus(const Structure& st, const PhongState& x, const double* dc, const double* dv, const double* dg,
              PhongState& y) const {
        y = x;
        SE3Perturbation sp;
        UnitVectorPerturbation up;
        for (size_t f = 0; f < st.free_cams.size(); ++f) {
            const size_t k = st.free_cams[f];
            sp(&x.poses[12 * k], dc + 6 * f, &y.poses[12 * k]);
        }
        for (size_t a = 0; a < st.active_v.size(); ++a) {
            const size_t j = st.active_v[a];
            if (!hold_positions)
                for (int c = 0; c < 3; ++c) y.pos[3 * j + c] = x.pos[3 * j + c] + dv[6 * a + c];
            up(&x.nrm[3 * j], dv + 6 * a + 3, &y.nrm[3 * j]);
        }
        for (int m = 0; m < n_mat; ++m)
            for (int c = 0; c < 3; ++c) {
                const int q = 3 * m + c;
                if (!st.g_used[q]) continue;
                y.g[q] = std::min(std::max(x.g[q] + dg[q], mat_lo[c]), mat_hi[c]);
            }
        for (int t = 0; t < n_tex; ++t) {
            const int q = g_tex0() + t;
            if (!st.g_used[q]) continue;
            y.g[q] = std::min(std::max(x.g[q] + dg[q], tex_lo), tex_hi);
        }
        const int l0 = g_light0();
        if (st.g_used[l0]) {
            if (directional)
                up(&x.g[l0], dg + l0, &y.g[l0]);
            else
                for (int c = 0; c < 3; ++c) y.g[l0 + c] = x.g[l0 + c] + dg[l0 + c];
        }
    }

    struct Scaling {
        std::vector<double> c, v, g;  // column scaling (cams 6 nf, vertices 6 na, globals n_g)
    };

    // (J^T J + D^2) y = J^T r with J column-scaled; eliminates the vertex blocks.
    bool schur_solve(const Structure& st, const Lin& L, const Scaling& sc, const std::vector<double>& Dc,
                     const std::vector<double>& Dv, const std::vector<double>& Dg, int nthreads,
                     std::vector<double>& yc, std::vector<double>& yv, std::vector<double>& yg) const {
        const int nf = int(st.free_cams.size()), ng = n_g();
        const size_t na = st.active_v.size();
        const int nr = 6 * nf + ng;
        std::vector<double> R(size_t(nr) * nr, 0.0), b(nr, 0.0);
        std::vector<double> Vinv(36 * na), tv(6 * na);
        for (int i = 0; i < 6 * nf; ++i) R[size_t(i) * nr + i] = Dc[i] * Dc[i];
        for (int q = 0; q < ng; ++q) {
            const size_t i = 6 * size_t(nf) + q;
            R[i * nr + i] = st.g_used[q] ? Dg[q] * Dg[q] : 1.0;
        }
        std::atomic<bool> ok{true};
        parallel_for(na, nthreads, [&](size_t b0, size_t e0, int) {
            std::vector<double> Ws, Es;
            std::vector<int> fl;
            for (size_t a = b0; a < e0; ++a) {
                const int j = st.active_v[a];
                int gi[7];
                gidx_of(j, gi);
                const size_t o0 = st.v_ptr[a], o1 = st.v_ptr[a + 1], Lk = o1 - o0;
                const double* sv = &sc.v[6 * a];
                double sg[7];
                for (int q = 0; q < 7; ++q) sg[q] = sc.g[gi[q]];
                double V[36] = {0}, gv[6] = {0}, G[42] = {0}, Hgg[49] = {0}, gg[7] = {0};
                Ws.assign(36 * Lk, 0.0);
                Es.assign(42 * Lk, 0.0);
                fl.assign(Lk, -1);
                std::vector<double> Us(36 * Lk, 0.0), gcs(6 * Lk, 0.0);
                for (size_t x = o0; x < o1; ++x) {
                    const size_t i = st.v_obs[x];
                    const int f = st.cam_free[cam[i]];
                    fl[x - o0] = f;
                    const double* r = &L.r[7 * i];
                    double Av[42], ag[7], Ac[42];
                    for (int rr = 0; rr < 7; ++rr)
                        for (int c = 0; c < 6; ++c) Av[6 * rr + c] = L.Av[42 * i + 6 * rr + c] * sv[c];
                    for (int q = 0; q < 7; ++q) ag[q] = L.ag[7 * i + q] * sg[q];
                    for (int p = 0; p < 6; ++p) {
                        for (int q = 0; q < 6; ++q) {
                            double s = 0;
                            for (int rr = 0; rr < 7; ++rr) s += Av[6 * rr + p] * Av[6 * rr + q];
                            V[6 * p + q] += s;
                        }
                        double s = 0;
                        for (int rr = 0; rr < 7; ++rr) s += Av[6 * rr + p] * r[rr];
                        gv[p] += s;
                    }
                    for (int q = 0; q < 7; ++q) {
                        for (int p = 0; p < 6; ++p) G[6 * q + p] += ag[q] * Av[18 + p];
                        for (int q2 = 0; q2 < 7; ++q2) Hgg[7 * q + q2] += ag[q] * ag[q2];
                        gg[q] += ag[q] * r[3];
                    }
                    if (f < 0) continue;
                    const double* scf = &sc.c[6 * size_t(f)];
                    for (int rr = 0; rr < 7; ++rr)
                        for (int c = 0; c < 6; ++c) Ac[6 * rr + c] = L.Ac[42 * i + 6 * rr + c] * scf[c];
                    double* Wk = &Ws[36 * (x - o0)];
                    double* Ek = &Es[42 * (x - o0)];
                    double* Uk = &Us[36 * (x - o0)];
                    double* gc = &gcs[6 * (x - o0)];
                    for (int p = 0; p < 6; ++p) {
                        for (int q = 0; q < 6; ++q) {
                            double s = 0, u = 0;
                            for (int rr = 0; rr < 7; ++rr) {
                                s += Ac[6 * rr + p] * Av[6 * rr + q];
                                u += Ac[6 * rr + p] * Ac[6 * rr + q];
                            }
                            Wk[6 * p + q] = s;
                            Uk[6 * p + q] = u;
                        }
                        for (int q = 0; q < 7; ++q) Ek[7 * p + q] = Ac[18 + p] * ag[q];
                        double s = 0;
                        for (int rr = 0; rr < 7; ++rr) s += Ac[6 * rr + p] * r[rr];
                        gc[p] = s;
                    }
                }
                for (int p = 0; p < 6; ++p) V[7 * p] += Dv[6 * a + p] * Dv[6 * a + p];
                double Vi[36];
                if (!invert_spd6(V, Vi)) {
                    ok = false;
                    continue;
                }
                std::memcpy(&Vinv[36 * a], Vi, sizeof(Vi));
                std::memcpy(&tv[6 * a], gv, sizeof(gv));
                // global-global: Hgg - G Vi G^T ; rhs gg - G Vi gv
                double GV[42];
                for (int q = 0; q < 7; ++q)
                    for (int p = 0; p < 6; ++p) {
                        double s = 0;
                        for (int k = 0; k < 6; ++k) s += G[6 * q + k] * Vi[6 * k + p];
                        GV[6 * q + p] = s;
                    }
                for (int q = 0; q < 7; ++q) {
                    const size_t iq = 6 * size_t(nf) + gi[q];
                    for (int q2 = 0; q2 < 7; ++q2) {
                        double s = Hgg[7 * q + q2];
                        for (int k = 0; k < 6; ++k) s -= GV[6 * q + k] * G[6 * q2 + k];
                        atomic_add(R[iq * nr + 6 * size_t(nf) + gi[q2]], s);
                    }
                    double s = gg[q];
                    for (int k = 0; k < 6; ++k) s -= GV[6 * q + k] * gv[k];
                    atomic_add(b[iq], s);
                }
                for (size_t x = 0; x < Lk; ++x) {
                    const int f = fl[x];
                    if (f < 0) continue;
                    const double* Wk = &Ws[36 * x];
                    double Y[36];
                    for (int p = 0; p < 6; ++p)
                        for (int q = 0; q < 6; ++q) {
                            double s = 0;
                            for (int k = 0; k < 6; ++k) s += Wk[6 * p + k] * Vi[6 * k + q];
                            Y[6 * p + q] = s;
                        }
                    for (int p = 0; p < 6; ++p) {
                        const size_t ip = 6 * size_t(f) + p;
                        // camera-global: E - Y G^T (both triangles of R are kept)
                        for (int q = 0; q < 7; ++q) {
                            double s = Es[42 * x + 7 * p + q];
                            for (int k = 0; k < 6; ++k) s -= Y[6 * p + k] * G[6 * q + k];
                            const size_t iq = 6 * size_t(nf) + gi[q];
                            atomic_add(R[ip * nr + iq], s);
                            atomic_add(R[iq * nr + ip], s);
                        }
                        double s = gcs[6 * x + p];
                        for (int k = 0; k < 6; ++k) s -= Y[6 * p + k] * gv[k];
                        atomic_add(b[ip], s);
                        for (int q = 0; q < 6; ++q) atomic_add(R[ip * nr + 6 * size_t(f) + q], Us[36 * x + 6 * p + q]);
                    }
                    for (size_t y = 0; y < Lk; ++y) {
                        const int f2 = fl[y];
                        if (f2 < 0) continue;
                        const double* W2 = &Ws[36 * y];
                        for (int p = 0; p < 6; ++p)
                            for (int q = 0; q < 6; ++q) {
                                double s = 0;
                                for (int k = 0; k < 6; ++k) s += Y[6 * p + k] * W2[6 * q + k];
                                atomic_add(R[(6 * size_t(f) + p) * nr + 6 * size_t(f2) + q], -s);
                            }
                    }
                }
            }
        });
        if (!ok) return false;
        if (!dense_cholesky(R.data(), nr)) return false;
        dense_cholesky_solve(R.data(), nr, b.data());
        yc.assign(b.begin(), b.begin() + 6 * nf);
        yg.assign(b.begin() + 6 * nf, b.end());
        for (int q = 0; q < ng; ++q)
            if (!st.g_used[q]) yg[q] = 0.0;
        // back-substitution: yv = Vi (gv - sum_k W_k^T yc_k - G^T yg)
        yv.assign(6 * na, 0.0);
        parallel_for(na, nthreads, [&](size_t b0, size_t e0, int) {
            for (size_t a = b0; a < e0; ++a) {
                const int j = st.active_v[a];
                int gi[7];
                gidx_of(j, gi);
                const double* sv = &sc.v[6 * a];
                double t[6];
                for (int p = 0; p < 6; ++p) t[p] = tv[6 * a + p];
                for (size_t x = st.v_ptr[a]; x < st.v_ptr[a + 1]; ++x) {
                    const size_t i = st.v_obs[x];
                    const int f = st.cam_free[cam[i]];
                    // J y restricted to the camera and global columns of this observation
                    double Jy[7] = {0, 0, 0, 0, 0, 0, 0};
                    if (f >= 0)
                        for (int rr = 0; rr < 7; ++rr)
                            for (int c = 0; c < 6; ++c)
                                Jy[rr] += L.Ac[42 * i + 6 * rr + c] * sc.c[6 * size_t(f) + c] * yc[6 * size_t(f) + c];
                    for (int q = 0; q < 7; ++q) Jy[3] += L.ag[7 * i + q] * sc.g[gi[q]] * yg[gi[q]];
                    for (int p = 0; p < 6; ++p)
                        for (int rr = 0; rr < 7; ++rr) t[p] -= L.Av[42 * i + 6 * rr + p] * sv[p] * Jy[rr];
                }
                const double* Vi = &Vinv[36 * a];
                for (int p = 0; p < 6; ++p) {
                    double s = 0;
                    for (int k = 0; k < 6; ++k) s += Vi[6 * p + k] * t[k];
                    yv[6 * a + p] = s;
                }
            }
        });
        return true;
    }

    void load_state(PhongState& x) const {
        x.poses.assign(poses, poses + 12 * size_t(n_poses));
        x.pos.assign(positions, positions + 3 * size_t(n_vertices));
        x.nrm.assign(normals, normals + 3 * size_t(n_vertices));
        x.g.assign(n_g(), 0.0);
        std::copy(materials, materials + 3 * n_mat, x.g.begin());
        std::copy(textures, textures + n_tex, x.g.begin() + g_tex0());
        std::copy(light, light + 3, x.g.begin() + g_light0());
    }
    void write_back(const Structure& st, const PhongState& x) {
        for (int k : st.free_cams) std::memcpy(poses + 12 * size_t(k), &x.poses[12 * size_t(k)], 12 * sizeof(double));
        for (int j : st.active_v) {
            std::memcpy(positions + 3 * size_t(j), &x.pos[3 * size_t(j)], 3 * sizeof(double));
            std::memcpy(normals + 3 * size_t(j), &x.nrm[3 * size_t(j)], 3 * sizeof(double));
        }
        for (int q = 0; q < n_g(); ++q) {
            if (!st.g_used[q]) continue;
            if (q < g_tex0())
                materials[q] = x.g[q];
            else if (q < g_light0())
                textures[q - g_tex0()] = x.g[q];
            else
                light[q - g_light0()] = x.g[q];
        }
    }

    bool solve(const Options& opt, Summary& sum) {
        Structure st;
        build_structure(st);
        const int nf = int(st.free_cams.size()), ng = n_g();
        const size_t na = st.active_v.size();
        PhongState x, cand, tmp;
        load_state(x);
        sum = Summary();
        if (bounded) {
            // IterationZero: project the initial point onto the feasible set (Plus with delta = 0)
            std::vector<double> zc(6 * size_t(nf), 0.0), zv(6 * na, 0.0), zg(ng, 0.0);
            plus(st, x, zc.data(), zv.data(), zg.data(), tmp);
            x = tmp;
        }
        Lin L, Lc;
        if (!evaluate(x, true, L, opt.num_threads)) {
            sum.termination_type = FAILURE;
            sum.termination_reason = R_INITIAL_EVAL;
            return false;
        }
        double x_cost = L.cost;
        fixed_cost = 0.0;
        if (hold_positions) {
            // cost of the removed blocks at the (never changing) constant parameters
            const bool saved = hold_positions;
            hold_positions = false;
            Lin Lall;
            evaluate(x, false, Lall, opt.num_threads);
            hold_positions = saved;
            fixed_cost = Lall.cost - L.cost;
        }
        sum.initial_cost = x_cost + fixed_cost;
        Scaling sc;
        sc.c.assign(6 * size_t(nf), 1.0);
        sc.v.assign(6 * na, 1.0);
        sc.g.assign(ng, 1.0);
        std::vector<double> cn_c(6 * size_t(nf)), cn_v(6 * na), cn_g(ng), gc(6 * size_t(nf)), gv(6 * na), gg(ng);
        double gradient_max_norm = 0;
        auto ambient_diff_max = [&](const PhongState& a, const PhongState& b2, bool squared_sum) {
            double m = 0, s = 0;
            auto acc = [&](double d) {
                m = std::max(m, std::fabs(d));
                s += d * d;
            };
            for (int f = 0; f < nf; ++f)
                for (int c = 0; c < 12; ++c) acc(a.poses[12 * size_t(st.free_cams[f]) + c] - b2.poses[12 * size_t(st.free_cams[f]) + c]);
            for (size_t v = 0; v < na; ++v)
                for (int c = 0; c < 3; ++c) {
                    const size_t idx = 3 * size_t(st.active_v[v]) + c;
                    if (!hold_positions) acc(a.pos[idx] - b2.pos[idx]);
                    acc(a.nrm[idx] - b2.nrm[idx]);
                }
            for (int q = 0; q < ng; ++q)
                if (st.g_used[q]) acc(a.g[q] - b2.g[q]);
            return squared_sum ? s : m;
        };
        auto column_norms_and_gradient = [&]() {
            std::fill(cn_c.begin(), cn_c.end(), 0.0);
            std::fill(cn_v.begin(), cn_v.end(), 0.0);
            std::fill(cn_g.begin(), cn_g.end(), 0.0);
            std::fill(gc.begin(), gc.end(), 0.0);
            std::fill(gv.begin(), gv.end(), 0.0);
            std::fill(gg.begin(), gg.end(), 0.0);
            for (size_t i = 0; i < n_obs(); ++i) {
                const int f = st.cam_free[cam[i]], a = st.v_active[vtx[i]];
                int gi[7];
                gidx_of(vtx[i], gi);
                const double* r = &L.r[7 * i];
                for (int rr = 0; rr < 7; ++rr) {
                    for (int c = 0; c < 6; ++c) {
                        if (f >= 0) {
                            const double v = L.Ac[42 * i + 6 * rr + c];
                            cn_c[6 * size_t(f) + c] += v * v;
                            gc[6 * size_t(f) + c] += v * r[rr];
                        }
                        const double v = L.Av[42 * i + 6 * rr + c];
                        cn_v[6 * size_t(a) + c] += v * v;
                        gv[6 * size_t(a) + c] += v * r[rr];
                    }
                }
                for (int q = 0; q < 7; ++q) {
                    const double v = L.ag[7 * i + q];
                    cn_g[gi[q]] += v * v;
                    gg[gi[q]] += v * r[3];
                }
            }
            std::vector<double> nc(gc), nv(gv), ngv(gg);
            for (auto& v : nc) v = -v;
            for (auto& v : nv) v = -v;
            for (auto& v : ngv) v = -v;
            plus(st, x, nc.data(), nv.data(), ngv.data(), tmp);
            gradient_max_norm = ambient_diff_max(x, tmp, false);
        };
        column_norms_and_gradient();
        if (opt.jacobi_scaling) {
            for (size_t i = 0; i < sc.c.size(); ++i) sc.c[i] = 1.0 / (1.0 + std::sqrt(cn_c[i]));
            for (size_t i = 0; i < sc.v.size(); ++i) sc.v[i] = 1.0 / (1.0 + std::sqrt(cn_v[i]));
            for (size_t i = 0; i < sc.g.size(); ++i) sc.g[i] = 1.0 / (1.0 + std::sqrt(cn_g[i]));
        }
        auto x_norm_of = [&](const PhongState& a) {
            double s = 0;
            for (int f = 0; f < nf; ++f)
                for (int c = 0; c < 12; ++c) {
                    const double v = a.poses[12 * size_t(st.free_cams[f]) + c];
                    s += v * v;
                }
            for (size_t v = 0; v < na; ++v)
                for (int c = 0; c < 3; ++c) {
                    const size_t idx = 3 * size_t(st.active_v[v]) + c;
                    s += (hold_positions ? 0.0 : a.pos[idx] * a.pos[idx]) + a.nrm[idx] * a.nrm[idx];
                }
            for (int q = 0; q < ng; ++q)
                if (st.g_used[q]) s += a.g[q] * a.g[q];
            return std::sqrt(s);
        };
        double x_norm = x_norm_of(x);
        const int max_nonmono = opt.use_nonmonotonic_steps ? opt.max_consecutive_nonmonotonic_steps : 0;
        double se_minimum = x_cost, se_current = x_cost, se_reference = x_cost, se_candidate = x_cost;
        double se_acc_ref = 0, se_acc_cand = 0;
        int se_nonmono = 0;
        double minimum_cost = x_cost;
        double radius = opt.initial_trust_region_radius, decrease_factor = 2.0;
        bool reuse_diagonal = false;
        std::vector<double> diag_c(6 * size_t(nf)), diag_v(6 * na), diag_g(ng), Dc(6 * size_t(nf)), Dv(6 * na), Dg(ng);
        int invalid_steps = 0;
        IterationRow row{};
        row.cost = x_cost;
        row.gradient_max_norm = gradient_max_norm;
        row.radius = radius;
        sum.push_row(row);
        bool step_ok_prev = false;
        auto finish = [&](int type, int reason) {
            sum.termination_type = type;
            sum.termination_reason = reason;
        };
        auto clampd = [&](double v) { return std::min(std::max(v, opt.min_lm_diagonal), opt.max_lm_diagonal); };
        int iteration = 0;
        std::vector<double> yc, yv, yg, dc(6 * size_t(nf)), dv(6 * na), dg(ng), sdc, sdv, sdg;

        // ---- DoglegStrategy (dataset_ba_phong.cpp:88-89 sets DOGLEG / SUBSPACE_DOGLEG) ------------------
        // The same restatement as problem.hpp's, over the flat vector [poses | vertices | shared blocks] in
        // the coordinates step' = D step, D = sqrt(clamp(diag(J^T J))) of the Jacobi-scaled Jacobian.
        const size_t NC = 6 * size_t(nf), NV = 6 * na, NT = NC + NV + size_t(ng);
        double dl_mu = 1e-8, dl_alpha = 0.0, dl_step_norm = 0.0;
        bool dl_reuse = false, dl_1d = false;
        std::vector<double> dlD, dlg, dln, dlu0, dlu1;
        double sub_B[4] = {0, 0, 0, 0}, sub_g[2] = {0, 0};
        auto jv_flat = [&](const std::vector<double>& t, std::vector<double>& out) {
            out.assign(7 * n_obs(), 0.0);
            for (size_t i = 0; i < n_obs(); ++i) {
                const int f = st.cam_free[cam[i]], a = st.v_active[vtx[i]];
                int gi[7];
                gidx_of(vtx[i], gi);
                for (int rr = 0; rr < 7; ++rr) {
                    double m = 0;
                    for (int c = 0; c < 6; ++c) {
                        if (f >= 0) m += L.Ac[42 * i + 6 * rr + c] * sc.c[6 * size_t(f) + c] * t[6 * size_t(f) + c];
                        m += L.Av[42 * i + 6 * rr + c] * sc.v[6 * size_t(a) + c] * t[NC + 6 * size_t(a) + c];
                    }
                    if (rr == 3)
                        for (int q = 0; q < 7; ++q) m += L.ag[7 * i + q] * sc.g[gi[q]] * t[NC + NV + gi[q]];
                    out[7 * i + rr] = m;
                }
            }
        };
        auto dotf = [](const std::vector<double>& a, const std::vector<double>& b2) {
            double v = 0;
            for (size_t i = 0; i < a.size(); ++i) v += a[i] * b2[i];
            return v;
        };
        auto traditional_step = [&](std::vector<double>& sp) {
            const double gnorm = std::sqrt(dotf(dlg, dlg)), nnorm = std::sqrt(dotf(dln, dln));
            double cg = 0, cn = 0;
            if (nnorm <= radius) {
                cn = 1.0;
                dl_step_norm = nnorm;
            } else if (gnorm * dl_alpha >= radius) {
                cg = -(radius / gnorm);
                dl_step_norm = radius;
            } else {
                const double b_dot_a = -dl_alpha * dotf(dlg, dln);
                const double a2 = std::pow(dl_alpha * gnorm, 2.0);
                const double bma2 = a2 - 2 * b_dot_a + std::pow(nnorm, 2);
                const double c = b_dot_a - a2;
                const double d = std::sqrt(c * c + bma2 * (std::pow(radius, 2.0) - a2));
                const double beta = (c <= 0) ? (d - c) / bma2 : (radius * radius - a2) / (d + c);
                cg = -dl_alpha * (1.0 - beta);
                cn = beta;
                dl_step_norm = -1.0;
            }
            sp.resize(NT);
            for (size_t i = 0; i < NT; ++i) sp[i] = cg * dlg[i] + cn * dln[i];
            if (dl_step_norm < 0.0) dl_step_norm = std::sqrt(dotf(sp, sp));
        };
        auto subspace_step = [&](std::vector<double>& sp) {
            const double nnorm = std::sqrt(dotf(dln, dln));
            if (nnorm <= radius) {
                sp = dln;
                dl_step_norm = nnorm;
                return;
            }
            if (dl_1d) {
                const double gnorm = std::sqrt(dotf(dlg, dlg));
                sp.resize(NT);
                for (size_t i = 0; i < NT; ++i) sp[i] = -(radius / gnorm) * dlg[i];
                dl_step_norm = radius;
                return;
            }
            double xb[2];
            if (!dogleg_boundary_minimum(sub_B, sub_g, radius, xb)) {
                traditional_step(sp);
                return;
            }
            sp.resize(NT);
            for (size_t i = 0; i < NT; ++i) sp[i] = xb[0] * dlu0[i] + xb[1] * dlu1[i];
            dl_step_norm = radius;
        };
        // returns false when no valid step exists; sets lin_it to the number of linear solves
        auto dogleg_compute_step = [&](int& lin_it) -> bool {
            lin_it = 0;
            if (!dl_reuse) {
                dl_reuse = true;
                dlD.resize(NT);
                dlg.resize(NT);
                for (size_t i = 0; i < NC; ++i) {
                    dlD[i] = std::sqrt(clampd(cn_c[i] * sc.c[i] * sc.c[i]));
                    dlg[i] = gc[i] * sc.c[i] / dlD[i];
                }
                for (size_t i = 0; i < NV; ++i) {
                    dlD[NC + i] = std::sqrt(clampd(cn_v[i] * sc.v[i] * sc.v[i]));
                    dlg[NC + i] = gv[i] * sc.v[i] / dlD[NC + i];
                }
                for (int q = 0; q < ng; ++q) {
                    dlD[NC + NV + q] = std::sqrt(clampd(cn_g[q] * sc.g[q] * sc.g[q]));
                    dlg[NC + NV + q] = gg[q] * sc.g[q] / dlD[NC + NV + q];
                }
                std::vector<double> t(NT), Jg;
                for (size_t i = 0; i < NT; ++i) t[i] = dlg[i] / dlD[i];
                jv_flat(t, Jg);
                dl_alpha = dotf(dlg, dlg) / dotf(Jg, Jg);
                bool ok = false;
                while (dl_mu < 1.0) {
                    for (size_t i = 0; i < NC; ++i) Dc[i] = dlD[i] * std::sqrt(dl_mu);
                    for (size_t i = 0; i < NV; ++i) Dv[i] = dlD[NC + i] * std::sqrt(dl_mu);
                    for (int q = 0; q < ng; ++q) Dg[q] = dlD[NC + NV + q] * std::sqrt(dl_mu);
                    ok = schur_solve(st, L, sc, Dc, Dv, Dg, opt.num_threads, yc, yv, yg);
                    if (ok) {
                        for (double v : yc) ok = ok && std::isfinite(v);
                        for (double v : yv) ok = ok && std::isfinite(v);
                        for (double v : yg) ok = ok && std::isfinite(v);
                    }
                    if (ok) break;
                    dl_mu *= 10.0;
                }
                if (!ok) return false;
                dln.resize(NT);
                for (size_t i = 0; i < NC; ++i) dln[i] = -dlD[i] * yc[i];
                for (size_t i = 0; i < NV; ++i) dln[NC + i] = -dlD[NC + i] * yv[i];
                for (int q = 0; q < ng; ++q) dln[NC + NV + q] = -dlD[NC + NV + q] * yg[q];
                if (opt.dogleg_type == 1) {
                    const double ng2 = dotf(dlg, dlg), nn2 = dotf(dln, dln);
                    const bool g_first = ng2 >= nn2;
                    const std::vector<double>&av = g_first ? dlg : dln, &bv = g_first ? dln : dlg;
                    const double r00 = std::sqrt(std::max(ng2, nn2));
                    if (!(r00 > 0.0)) return false;
                    dlu0.resize(NT);
                    dlu1.resize(NT);
                    for (size_t i = 0; i < NT; ++i) dlu0[i] = av[i] / r00;
                    const double pr = dotf(dlu0, bv);
                    for (size_t i = 0; i < NT; ++i) dlu1[i] = bv[i] - pr * dlu0[i];
                    const double r11 = std::sqrt(dotf(dlu1, dlu1));
                    dl_1d = !(r11 > 2.0 * std::numeric_limits<double>::epsilon() * r00);
                    if (!dl_1d) {
                        for (auto& v : dlu1) v /= r11;
                        sub_g[0] = dotf(dlu0, dlg);
                        sub_g[1] = dotf(dlu1, dlg);
                        std::vector<double> J0, J1;
                        for (size_t i = 0; i < NT; ++i) t[i] = dlu0[i] / dlD[i];
                        jv_flat(t, J0);
                        for (size_t i = 0; i < NT; ++i) t[i] = dlu1[i] / dlD[i];
                        jv_flat(t, J1);
                        sub_B[0] = dotf(J0, J0);
                        sub_B[1] = sub_B[2] = dotf(J0, J1);
                        sub_B[3] = dotf(J1, J1);
                    }
                }
                lin_it = 1;
            }
            std::vector<double> sp;
            if (opt.dogleg_type == 1)
                subspace_step(sp);
            else
                traditional_step(sp);
            yc.resize(NC);
            yv.resize(NV);
            yg.resize(size_t(ng));
            for (size_t i = 0; i < NC; ++i) yc[i] = -sp[i] / dlD[i];
            for (size_t i = 0; i < NV; ++i) yv[i] = -sp[NC + i] / dlD[NC + i];
            for (int q = 0; q < ng; ++q) yg[q] = -sp[NC + NV + q] / dlD[NC + NV + q];
            return true;
        };
        const bool use_dogleg = opt.trust_region_strategy == 1;
        for (;;) {
            if (iteration > 0) {
                if (step_ok_prev) {
                    ++sum.num_successful_steps;
                    if (x_cost < minimum_cost) {
                        minimum_cost = x_cost;
                        write_back(st, x);
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
            bool valid;
            if (use_dogleg) {
                int lin_it = 0;
                valid = dogleg_compute_step(lin_it);
                row.linear_iterations = lin_it;
                sum.total_linear_iterations += lin_it;
            } else {
                if (!reuse_diagonal) {
                    for (size_t i = 0; i < diag_c.size(); ++i) diag_c[i] = clampd(cn_c[i] * sc.c[i] * sc.c[i]);
                    for (size_t i = 0; i < diag_v.size(); ++i) diag_v[i] = clampd(cn_v[i] * sc.v[i] * sc.v[i]);
                    for (size_t i = 0; i < diag_g.size(); ++i) diag_g[i] = clampd(cn_g[i] * sc.g[i] * sc.g[i]);
                }
                for (size_t i = 0; i < Dc.size(); ++i) Dc[i] = std::sqrt(diag_c[i] / radius);
                for (size_t i = 0; i < Dv.size(); ++i) Dv[i] = std::sqrt(diag_v[i] / radius);
                for (size_t i = 0; i < Dg.size(); ++i) Dg[i] = std::sqrt(diag_g[i] / radius);
                valid = schur_solve(st, L, sc, Dc, Dv, Dg, opt.num_threads, yc, yv, yg);
                row.linear_iterations = 1;
                sum.total_linear_iterations += 1;
            }
            reuse_diagonal = true;
            if (valid) {
                for (double v : yc) valid = valid && std::isfinite(v);
                for (double v : yv) valid = valid && std::isfinite(v);
                for (double v : yg) valid = valid && std::isfinite(v);
            }
            double model_cost_change = 0;
            if (valid) {
                double acc = 0;
                for (size_t i = 0; i < n_obs(); ++i) {
                    const int f = st.cam_free[cam[i]], a = st.v_active[vtx[i]];
                    int gi[7];
                    gidx_of(vtx[i], gi);
                    for (int rr = 0; rr < 7; ++rr) {
                        double m = 0;
                        for (int c = 0; c < 6; ++c) {
                            if (f >= 0) m -= L.Ac[42 * i + 6 * rr + c] * sc.c[6 * size_t(f) + c] * yc[6 * size_t(f) + c];
                            m -= L.Av[42 * i + 6 * rr + c] * sc.v[6 * size_t(a) + c] * yv[6 * size_t(a) + c];
                        }
                        if (rr == 3)
                            for (int q = 0; q < 7; ++q) m -= L.ag[7 * i + q] * sc.g[gi[q]] * yg[gi[q]];
                        acc -= m * (L.r[7 * i + rr] + 0.5 * m);
                    }
                }
                model_cost_change = acc;
                if (!(model_cost_change > 0.0)) valid = false;
            }
            if (!valid) {
                row.cost = x_cost;
                row.gradient_max_norm = gradient_max_norm;
                if (++invalid_steps >= opt.max_num_consecutive_invalid_steps) {
                    row.radius = radius;
                    sum.push_row(row);
                    finish(FAILURE, R_INVALID_STEPS);
                    break;
                }
                if (use_dogleg) {
                    dl_mu *= 10.0;  // DoglegStrategy::StepIsInvalid
                    dl_reuse = false;
                } else {
                    radius = radius / decrease_factor;
                    decrease_factor *= 2.0;
                }
                row.radius = radius;
                sum.push_row(row);
                continue;
            }
            invalid_steps = 0;
            row.step_is_valid = 1;
            for (size_t i = 0; i < dc.size(); ++i) dc[i] = -yc[i] * sc.c[i];
            for (size_t i = 0; i < dv.size(); ++i) dv[i] = -yv[i] * sc.v[i];
            for (size_t i = 0; i < dg.size(); ++i) dg[i] = -yg[i] * sc.g[i];
            auto cost_at = [&](double alpha, PhongState& out) {
                sdc = dc, sdv = dv, sdg = dg;
                if (alpha != 1.0) {
                    for (auto& v : sdc) v *= alpha;
                    for (auto& v : sdv) v *= alpha;
                    for (auto& v : sdg) v *= alpha;
                }
                plus(st, x, sdc.data(), sdv.data(), sdg.data(), out);
                if (evaluate(out, false, Lc, opt.num_threads)) return Lc.cost;
                return std::numeric_limits<double>::quiet_NaN();
            };
            double alpha = 1.0;
            double cand_cost = cost_at(1.0, cand);
            if (bounded) {
                // TrustRegionMinimizer::DoLineSearch (Armijo, see the header for the interpolation)
                double g0 = 0;
                for (size_t i = 0; i < dc.size(); ++i) g0 += gc[i] * dc[i];
                for (size_t i = 0; i < dv.size(); ++i) g0 += gv[i] * dv[i];
                for (int q = 0; q < ng; ++q)
                    if (st.g_used[q]) g0 += gg[q] * dg[q];
                double dmax = 0;
                for (double v : dc) dmax = std::max(dmax, std::fabs(v));
                for (double v : dv) dmax = std::max(dmax, std::fabs(v));
                for (double v : dg) dmax = std::max(dmax, std::fabs(v));
                double a_cur = 1.0, f_cur = cand_cost;
                bool success = true;
                int ls_it = 0;
                while (!std::isfinite(f_cur) || f_cur > x_cost + opt.line_search_sufficient_function_decrease * g0 * a_cur) {
                    if (++ls_it >= 20) {
                        success = false;
                        break;
                    }
                    const double lo = 1e-3 * a_cur, hi = 0.6 * a_cur;
                    double a_new;
                    if (!std::isfinite(f_cur)) {
                        a_new = std::min(std::max(0.5 * a_cur, lo), hi);
                    } else {
                        const double c2 = (f_cur - x_cost - g0 * a_cur) / (a_cur * a_cur);
                        a_new = c2 > 0.0 ? -g0 / (2.0 * c2) : hi;
                        a_new = std::min(std::max(a_new, lo), hi);
                    }
                    if (a_new * dmax < 1e-9) {
                        success = false;
                        break;
                    }
                    a_cur = a_new;
                    f_cur = cost_at(a_cur, tmp);
                }
                if (success && a_cur != 1.0) {
                    alpha = a_cur;
                    cand = tmp;
                    cand_cost = f_cur;
                }
            }
            (void)alpha;
            if (!std::isfinite(cand_cost)) cand_cost = std::numeric_limits<double>::max();
            row.step_norm = std::sqrt(ambient_diff_max(x, cand, true));
            row.cost_change = x_cost - cand_cost;
            if (row.step_norm <= opt.parameter_tolerance * (x_norm + opt.parameter_tolerance)) {
                row.cost = x_cost;
                row.gradient_max_norm = gradient_max_norm;
                row.radius = radius;
                sum.push_row(row);
                finish(CONVERGENCE, R_PARAMETER_TOL);
                break;
            }
            if (std::fabs(row.cost_change) <= opt.function_tolerance * x_cost) {
                row.cost = x_cost;
                row.gradient_max_norm = gradient_max_norm;
                row.radius = radius;
                sum.push_row(row);
                finish(CONVERGENCE, R_FUNCTION_TOL);
                break;
            }
            const double rel = (se_current - cand_cost) / model_cost_change;
            const double hist = (se_reference - cand_cost) / (se_acc_ref + model_cost_change);
            row.relative_decrease = std::max(rel, hist);
            if (row.relative_decrease > opt.min_relative_decrease) {
                x = cand;
                x_norm = x_norm_of(x);
                if (!evaluate(x, true, L, opt.num_threads)) {
                    finish(FAILURE, R_INITIAL_EVAL);
                    break;
                }
                x_cost = L.cost;
                column_norms_and_gradient();
                step_ok_prev = true;
                row.step_is_successful = 1;
                if (use_dogleg) {
                    // DoglegStrategy::StepAccepted
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
            } else if (use_dogleg) {
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
        sum.final_cost = minimum_cost + fixed_cost;
        sum.final_radius = radius;
        for (auto& rw : sum.rows) rw.cost += fixed_cost;  // IterationSummary::cost includes the fixed cost
        return sum.termination_type != FAILURE;
    }
};

}  // namespace oracle
