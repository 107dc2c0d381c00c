// K2p / K4p — the joint solve of dataset_ba_phong (tests/dataset_ba_phong.cpp:26-255, stage 3): stereo,
// intensity and normal blocks of every observation; parameter blocks pose (6, SE3Perturbation),
// vertex = position + normal (3 + 3, UnitVectorPerturbation on the normal, perturbations.hpp:87-113)
// and the blocks every vertex of a material shares: material [ka, ks, alpha], texture kd, light.
//
// The vertex blocks are eliminated exactly like the landmark blocks of the stereo path, only 6x6:
//   V = sum A_v^T A_v + D^2,  W_k = A_c^T A_v (6x6),  G = a_g (x) A_v[intensity row]  (7x6)
//   S_cc[a,b] -= W_a V^-1 W_b^T      block-sparse reduced camera system (same pattern as stereo)
//   S_cg[a,:] += A_c[int]^T a_g - W_a V^-1 G^T      dense border, n_g = 3 n_mat + n_tex + 3 columns
//   S_gg      += a_g a_g^T - G V^-1 G^T
// which leaves an ARROWHEAD system [S_cc S_cg; S_cg^T S_gg]: banded camera part + a dense border of
// a few dozen columns.  The border is eliminated with n_g + 1 solves against S_cc (engine.cu).
//
// One warp per vertex, one lane per observation (track length <= 32); the per-observation blocks
// are the closed forms of closed_form.h.  FP64, no fast-math.
#include "kernels.cuh"
#include "phong_common.cuh"

namespace cslam {

namespace {

// =============================================================================================
// K2p — fused residual/Jacobian + elimination of the vertex blocks.
// kSchur == false: the initial pass (cost, squared column norms, gradient) with unit scaling.
// =============================================================================================
// LW lanes per vertex: 32 (tracks up to 32 observations) or 16 (two vertices per warp, one per half).
template <bool kSchur, int LW>
__global__ void __launch_bounds__(PB_WARPS * 32)
    phong_build_kernel(DevView v, PhongSolveView q, int lm_lo, int lm_hi, LmDiag dg, PhongSystem o) {
    __shared__ double sW[PB_WARPS][36 * 32];  // W of every lane, [k][lane]
    __shared__ double sG[PB_WARPS][32 / LW][72];  // the vertex's global contributions, staged for the atomics
    __shared__ int sF[PB_WARPS][32];
    __shared__ double s_red[32];
    constexpr int NSUB = 32 / LW;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, lv = lane % LW, sub = lane / LW, lbase = sub * LW;
    double* myW = sW[wib];
    double* myG = sG[wib][sub];
    int* myF = sF[wib];
    const int nf6 = 6 * v.n_free;
    double cost = 0.0, fixed = 0.0;
    const int warps_total = gridDim.x * PB_WARPS;
    for (int jb = lm_lo + (blockIdx.x * PB_WARPS + wib) * NSUB; jb < lm_hi; jb += warps_total * NSUB) {
        const bool vok = jb + sub < lm_hi;
        const int j = vok ? jb + sub : lm_hi - 1;  // an idle half re-reads the last vertex and writes nothing
        const long long e0 = v.lm_base[j], es = v.lm_stride[j];
        const int L = vok ? int(v.lm_cnt[j]) : 0;
        if (LW == 32 && L > 32) continue;  // (warp-uniform) a long vertex: kernels_phong_long.cu
        VertexCtx c;
        load_vertex(v, q, j, c);
        PhObs ob;
        ob.f = -1;
        const bool act = lv < L;
        double V21[21], gv[6], r7[7];
        if (act) {
            eval_phong_obs(v, q, e0 + lv * es, c, ob);
            r7[0] = ob.rs[0], r7[1] = ob.rs[1], r7[2] = ob.rs[2], r7[3] = ob.rI;
            r7[4] = ob.rN[0], r7[5] = ob.rN[1], r7[6] = ob.rN[2];
#pragma unroll
            for (int k = 0; k < 7; ++k) cost += 0.5 * r7[k] * r7[k];
            if (!kSchur && q.hold_positions && ob.f < 0) {
                // the dropped stereo block's cost, once (initial pass): Ceres' fixed_cost
                const long long e = e0 + lv * es;
                double rs[3];
                stereo_block<false>(v.cam, v.poses + 12ll * v.obs_cam[e], c.p, v.obs_u[e], v.obs_v[e], v.obs_d[e], v.obs_W, rs,
                                    nullptr, nullptr);
                fixed += 0.5 * (rs[0] * rs[0] + rs[1] * rs[1] + rs[2] * rs[2]);
            }
            vertex_normal_eq(ob, r7, V21, gv);
        } else {
#pragma unroll
            for (int k = 0; k < 21; ++k) V21[k] = 0.0;
#pragma unroll
            for (int k = 0; k < 6; ++k) gv[k] = 0.0;
#pragma unroll
            for (int k = 0; k < 7; ++k) ob.ag[k] = 0.0;
#pragma unroll
            for (int k = 0; k < 3; ++k) ob.ip[k] = ob.in[k] = 0.0;
            ob.rI = 0.0;
        }
#pragma unroll
        for (int k = 0; k < 21; ++k) V21[k] = seg_sum<LW>(V21[k]);
#pragma unroll
        for (int k = 0; k < 6; ++k) gv[k] = seg_sum<LW>(gv[k]);
        // globals: gradient and diagonal (always), G and H_gg (Schur pass)
        double ggl[7], hd[7];
#pragma unroll
        for (int k = 0; k < 7; ++k) {
            ggl[k] = seg_sum<LW>(ob.ag[k] * ob.rI);
            hd[k] = seg_sum<LW>(ob.ag[k] * ob.ag[k]);
        }
        if (!kSchur) {
            if (lv == 0 && vok) {
                int idx = 0;
#pragma unroll
                for (int a = 0; a < 6; ++a) {
                    if (a < 3)
                        o.cn_l[3ll * j + a] = V21[idx];
                    else
                        o.cn_n[3ll * j + a - 3] = V21[idx];
                    idx += 6 - a;
                    o.gv[6ll * j + a] = gv[a];
                }
#pragma unroll
                for (int k = 0; k < 7; ++k) {
                    red_add(&o.gg[c.gi[k]], ggl[k]);
                    red_add(&o.hg[c.gi[k]], hd[k]);
                }
            }
            if (act && ob.f >= 0) {
#pragma unroll
                for (int a = 0; a < 6; ++a) {
                    const double cn = ob.Jcs[a] * ob.Jcs[a] + ob.Jcs[6 + a] * ob.Jcs[6 + a] + ob.Jcs[12 + a] * ob.Jcs[12 + a] +
                                      ob.JIc[a] * ob.JIc[a] + ob.JNc[a] * ob.JNc[a] + ob.JNc[6 + a] * ob.JNc[6 + a] +
                                      ob.JNc[12 + a] * ob.JNc[12 + a];
                    const double ga = ob.Jcs[a] * r7[0] + ob.Jcs[6 + a] * r7[1] + ob.Jcs[12 + a] * r7[2] + ob.JIc[a] * r7[3] +
                                      ob.JNc[a] * r7[4] + ob.JNc[6 + a] * r7[5] + ob.JNc[12 + a] * r7[6];
                    red_add(&o.Bdiag[36ll * ob.f + 7 * a], cn);
                    red_add(&o.gp[6ll * ob.f + a], ga);
                }
            }
            continue;
        }
        // ---- Schur pass ---------------------------------------------------------------------
        double G[42];  // G[q][p] = sum_obs ag[q] * Av[intensity row][p]
#pragma unroll
        for (int k = 0; k < 7; ++k) {
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                G[6 * k + a] = seg_sum<LW>(ob.ag[k] * ob.ip[a]);
                G[6 * k + 3 + a] = seg_sum<LW>(ob.ag[k] * ob.in[a]);
            }
        }
        double V[36], Vi[36];
        unpack_sym6(V21, V);
        add_lm_diag(V, dg);
        const bool pd = spd6_inverse(V, Vi) && vok;
        if (lv == 0 && vok) {
#pragma unroll
            for (int a = 0; a < 6; ++a) o.gv[6ll * j + a] = gv[a];
            if (!pd) red_add(&o.scal[SC_INVALID], 1.0);
        }
        // (no early exit: the other half of the warp may hold a valid vertex; `pd` guards every write)
        // GV = G Vi (7x6)
        double GV[42];
#pragma unroll
        for (int k = 0; k < 7; ++k)
#pragma unroll
            for (int a = 0; a < 6; ++a) {
                double s = 0.0;
#pragma unroll
                for (int b = 0; b < 6; ++b) s += G[6 * k + b] * Vi[6 * b + a];
                GV[6 * k + a] = s;
            }
        // H_gg off-diagonal needs the pair sums ag_k ag_k2: 21 more reductions
        __syncwarp();
        {
            int idx = 0;
#pragma unroll
            for (int k = 0; k < 7; ++k)
#pragma unroll
                for (int k2 = k; k2 < 7; ++k2) {
                    double h = (k2 == k) ? hd[k] : seg_sum<LW>(ob.ag[k] * ob.ag[k2]);
#pragma unroll
                    for (int a = 0; a < 6; ++a) h -= GV[6 * k + a] * G[6 * k2 + a];
                    if (lv == 0) myG[idx] = h;
                    ++idx;
                }
#pragma unroll
            for (int k = 0; k < 7; ++k) {
                double bb = ggl[k];
#pragma unroll
                for (int a = 0; a < 6; ++a) bb -= GV[6 * k + a] * gv[a];
                if (lv == 0) {
                    myG[28 + k] = bb;
                    myG[35 + k] = ggl[k];
                    myG[42 + k] = hd[k];
                }
            }
        }
        __syncwarp();
        // 28 upper entries of the 7x7 block (mirrored), then rhs / gradient / diagonal
        {
            // entry -> (k, k2) of the upper triangle
            for (int en = lv; en < 28 && pd; en += LW) {
                int k = 0, rem = en;
                while (rem >= 7 - k) {
                    rem -= 7 - k;
                    ++k;
                }
                const int k2 = k + rem;
                const double h = myG[en];
                const int gk = c.gi[k], gk2 = c.gi[k2];
                red_add(&o.Sgg[(long long)gk * q.n_g + gk2], h);
                if (gk != gk2) red_add(&o.Sgg[(long long)gk2 * q.n_g + gk], h);
            }
            if (lv < 7 && pd) {
                red_add(&o.bg[c.gi[lv]], myG[28 + lv]);
                red_add(&o.gg[c.gi[lv]], myG[35 + lv]);
                red_add(&o.hg[c.gi[lv]], myG[42 + lv]);
            }
        }
        // ---- camera part --------------------------------------------------------------------
        double Y[36];
        myF[lane] = (act && pd) ? ob.f : -1;
        if (act && pd && ob.f >= 0) {
            double W[36];
#pragma unroll
            for (int a = 0; a < 6; ++a) {
#pragma unroll
                for (int b = 0; b < 3; ++b) {
                    W[6 * a + b] = ob.Jcs[a] * ob.S[b] + ob.Jcs[6 + a] * ob.S[3 + b] + ob.Jcs[12 + a] * ob.S[6 + b] +
                                   ob.JIc[a] * ob.ip[b];
                    W[6 * a + 3 + b] = ob.JIc[a] * ob.in[b] + ob.JNc[a] * ob.N[b] + ob.JNc[6 + a] * ob.N[3 + b] +
                                       ob.JNc[12 + a] * ob.N[6 + b];
                }
            }
#pragma unroll
            for (int k = 0; k < 36; ++k) myW[k * 32 + lane] = W[k];
#pragma unroll
            for (int a = 0; a < 6; ++a)
#pragma unroll
                for (int b = 0; b < 6; ++b) {
                    double s = 0.0;
#pragma unroll
                    for (int k = 0; k < 6; ++k) s += W[6 * a + k] * Vi[6 * k + b];
                    Y[6 * a + b] = s;
                }
            double* Bd = o.Bdiag + 36ll * ob.f;
#pragma unroll
            for (int a = 0; a < 6; ++a) {
#pragma unroll
                for (int b = a; b < 6; ++b)
                    red_add(&Bd[6 * a + b], ob.Jcs[a] * ob.Jcs[b] + ob.Jcs[6 + a] * ob.Jcs[6 + b] + ob.Jcs[12 + a] * ob.Jcs[12 + b] +
                                                ob.JIc[a] * ob.JIc[b] + ob.JNc[a] * ob.JNc[b] + ob.JNc[6 + a] * ob.JNc[6 + b] +
                                                ob.JNc[12 + a] * ob.JNc[12 + b]);
                const double ga = ob.Jcs[a] * r7[0] + ob.Jcs[6 + a] * r7[1] + ob.Jcs[12 + a] * r7[2] + ob.JIc[a] * r7[3] +
                                  ob.JNc[a] * r7[4] + ob.JNc[6 + a] * r7[5] + ob.JNc[12 + a] * r7[6];
                double yg = 0.0;
#pragma unroll
                for (int k = 0; k < 6; ++k) yg += Y[6 * a + k] * gv[k];
                red_add(&o.gp[6ll * ob.f + a], ga);
                red_add(&o.bp[6ll * ob.f + a], ga - yg);
                // border: E - Y G^T
#pragma unroll
                for (int k = 0; k < 7; ++k) {
                    double s = ob.JIc[a] * ob.ag[k];
#pragma unroll
                    for (int b = 0; b < 6; ++b) s -= Y[6 * a + b] * G[6 * k + b];
                    red_add(&o.Scg[(long long)c.gi[k] * nf6 + 6 * ob.f + a], s);
                }
            }
        }
        __syncwarp();
        // camera pairs: lane x takes the pairs (x, (x + s) mod L), s = 0 .. L/2 — every unordered
        // pair once, all lanes busy
        if (act && pd && ob.f >= 0) {
            const int fx = ob.f;
            for (int s = 0; s <= L / 2; ++s) {
                if (2 * s == L && lv >= s) break;  // even L: the antipodal pairs appear twice
                int y = lv + s;
                if (y >= L) y -= L;
                y += lbase;
                const int fy = myF[y];
                if (fy < 0) continue;
                double Wy[36];
#pragma unroll
                for (int k = 0; k < 36; ++k) Wy[k] = myW[k * 32 + y];
                const bool swap = fx > fy;
                const int a = swap ? fy : fx, b = swap ? fx : fy;
                double* B = o.S + 36ll * find_block(v.s_rowptr, v.s_col, a, b);
                if (fx == fy) {
                    const bool dup = s != 0;  // the same camera observing the vertex twice
#pragma unroll
                    for (int pp = 0; pp < 6; ++pp)
#pragma unroll
                        for (int qq = pp; qq < 6; ++qq) {
                            double val = 0.0;
#pragma unroll
                            for (int k = 0; k < 6; ++k) val += Y[6 * pp + k] * Wy[6 * qq + k];
                            if (dup)
#pragma unroll
                                for (int k = 0; k < 6; ++k) val += Y[6 * qq + k] * Wy[6 * pp + k];
                            red_add(&B[6 * pp + qq], -val);
                        }
                } else {
#pragma unroll
                    for (int pp = 0; pp < 6; ++pp)
#pragma unroll
                        for (int qq = 0; qq < 6; ++qq) {
                            double val = 0.0;
#pragma unroll
                            for (int k = 0; k < 6; ++k) val += Y[6 * pp + k] * Wy[6 * qq + k];
                            red_add(swap ? &B[6 * qq + pp] : &B[6 * pp + qq], -val);
                        }
                }
            }
        }
        __syncwarp();
    }
    block_atomic_sum(cost, &o.scal[SC_COST], s_red);
    if (!kSchur) block_atomic_sum(fixed, &o.scal[SC_FIXED], s_red);
}

// =============================================================================================
// K4p — back-substitution of the vertex blocks and the model cost change
// =============================================================================================
template <int LW>
__global__ void __launch_bounds__(PB_WARPS * 32)
    phong_backsub_kernel(DevView v, PhongSolveView q, int lm_lo, int lm_hi, LmDiag dg, const double* __restrict__ yp,
                         const double* __restrict__ yg, const double* __restrict__ gv, double* __restrict__ yv_out,
                         double* __restrict__ scal2) {
    __shared__ double s_red[32];
    constexpr int NSUB = 32 / LW;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, lv = lane % LW, sub = lane / LW;
    double model = 0.0, bad = 0.0, gy = 0.0, dmax = 0.0;
    const int warps_total = gridDim.x * PB_WARPS;
    for (int jb = lm_lo + (blockIdx.x * PB_WARPS + wib) * NSUB; jb < lm_hi; jb += warps_total * NSUB) {
        const bool vok = jb + sub < lm_hi;
        const int j = vok ? jb + sub : lm_hi - 1;
        const long long e0 = v.lm_base[j], es = v.lm_stride[j];
        const int L = vok ? int(v.lm_cnt[j]) : 0;
        if (LW == 32 && L > 32) continue;  // (warp-uniform) a long vertex: kernels_phong_long.cu
        VertexCtx c;
        load_vertex(v, q, j, c);
        PhObs ob;
        ob.f = -1;
        const bool act = lv < L;
        double V21[21], t6[6], r7[7], Jy[7];
        if (act) {
            eval_phong_obs(v, q, e0 + lv * es, c, ob);
            r7[0] = ob.rs[0], r7[1] = ob.rs[1], r7[2] = ob.rs[2], r7[3] = ob.rI;
            r7[4] = ob.rN[0], r7[5] = ob.rN[1], r7[6] = ob.rN[2];
            // J y restricted to the camera and global columns
#pragma unroll
            for (int k = 0; k < 7; ++k) Jy[k] = 0.0;
            if (ob.f >= 0) {
                const double* y = yp + 6ll * ob.f;
#pragma unroll
                for (int a = 0; a < 6; ++a) {
                    const double ya = y[a];
                    Jy[0] += ob.Jcs[a] * ya;
                    Jy[1] += ob.Jcs[6 + a] * ya;
                    Jy[2] += ob.Jcs[12 + a] * ya;
                    Jy[3] += ob.JIc[a] * ya;
                    Jy[4] += ob.JNc[a] * ya;
                    Jy[5] += ob.JNc[6 + a] * ya;
                    Jy[6] += ob.JNc[12 + a] * ya;
                }
            }
#pragma unroll
            for (int k = 0; k < 7; ++k) Jy[3] += ob.ag[k] * yg[c.gi[k]];
            double w[7];
#pragma unroll
            for (int k = 0; k < 7; ++k) w[k] = r7[k] - Jy[k];
            vertex_normal_eq(ob, w, V21, t6);
        } else {
#pragma unroll
            for (int k = 0; k < 21; ++k) V21[k] = 0.0;
#pragma unroll
            for (int k = 0; k < 6; ++k) t6[k] = 0.0;
        }
#pragma unroll
        for (int k = 0; k < 21; ++k) V21[k] = seg_sum<LW>(V21[k]);
#pragma unroll
        for (int k = 0; k < 6; ++k) t6[k] = seg_sum<LW>(t6[k]);
        double V[36], Vi[36], yv[6];
        unpack_sym6(V21, V);
        add_lm_diag(V, dg);
        const bool pd = spd6_inverse(V, Vi);
#pragma unroll
        for (int a = 0; a < 6; ++a) {
            double s = 0.0;
#pragma unroll
            for (int b = 0; b < 6; ++b) s += Vi[6 * a + b] * t6[b];
            yv[a] = pd ? s : 0.0;
        }
        if (lv == 0 && vok) {
#pragma unroll
            for (int a = 0; a < 6; ++a) {
                yv_out[6ll * j + a] = yv[a];
                if (isnan(yv[a]) || isinf(yv[a])) bad = 1.0;
                // line search of a bounded problem: g . y and |delta|_inf over the vertex blocks
                gy += gv[6ll * j + a] * yv[a];
                dmax = fmax(dmax, fabs(yv[a] * (a < 3 ? c.sl[a] : c.sn[a - 3])));
            }
        }
        if (act) {
            // m = -(J y) over all columns; model -= m . (r + m / 2)
            double m[7];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                m[k] = -(Jy[k] + ob.S[3 * k] * yv[0] + ob.S[3 * k + 1] * yv[1] + ob.S[3 * k + 2] * yv[2]);
                m[4 + k] = -(Jy[4 + k] + ob.N[3 * k] * yv[3] + ob.N[3 * k + 1] * yv[4] + ob.N[3 * k + 2] * yv[5]);
            }
            m[3] = -(Jy[3] + ob.ip[0] * yv[0] + ob.ip[1] * yv[1] + ob.ip[2] * yv[2] + ob.in[0] * yv[3] + ob.in[1] * yv[4] +
                     ob.in[2] * yv[5]);
#pragma unroll
            for (int k = 0; k < 7; ++k) model -= m[k] * (r7[k] + 0.5 * m[k]);
        }
    }
    for (int o = 16; o > 0; o >>= 1) dmax = fmax(dmax, __shfl_xor_sync(0xffffffffu, dmax, o));
    if (lane == 0 && dmax > 0) atomic_max_nonneg(&scal2[SC_LS_DMAX], dmax);
    block_atomic_sum(model, &scal2[SC_MODEL], s_red);
    block_atomic_sum(bad, &scal2[SC_NONFINITE], s_red);
    block_atomic_sum(gy, &scal2[SC_LS_GY], s_red);
}

// =============================================================================================
// DOGLEG for the lighting solve (dataset_ba_phong.cpp:88-89): the eight inner products of
// DoglegStrategy over [poses | vertices | shared blocks] — see dogleg_products_kernel in kernels.cu
// for the stereo path — with the vertex part of D^2 = clamp(diag(J^T J)) formed on the fly and
// written out (diag_v, and the interleaved column scaling sc_v) for the combine / line-search passes.
//   sums: 0 g'.g'  1 g'.gn'  2 gn'.gn'  3 |J D^-2 g|^2  4 (J D^-2 g).(J y)  5 |J y|^2  6 (J D^-2 g).r  7 (J y).r
// The pose and shared-block terms of sums 0..2 are added by their own small kernels.
// =============================================================================================
template <int LW>
__global__ void __launch_bounds__(PB_WARPS * 32)
    phong_dogleg_products_kernel(DevView v, PhongSolveView q, int lm_lo, int lm_hi, LmDiag dg, const double* __restrict__ gp,
                                 const double* __restrict__ diag_p, const double* __restrict__ yp,
                                 const double* __restrict__ gg, const double* __restrict__ diag_g,
                                 const double* __restrict__ yg, const double* __restrict__ gv, const double* __restrict__ yv,
                                 double* __restrict__ diag_v_out, double* __restrict__ sc_v_out, double* __restrict__ sums) {
    __shared__ double s_red[32];
    constexpr int NSUB = 32 / LW;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, lv = lane % LW, sub = lane / LW;
    double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const int warps_total = gridDim.x * PB_WARPS;
    for (int jb = lm_lo + (blockIdx.x * PB_WARPS + wib) * NSUB; jb < lm_hi; jb += warps_total * NSUB) {
        const bool vok = jb + sub < lm_hi;
        const int j = vok ? jb + sub : lm_hi - 1;
        const long long e0 = v.lm_base[j], es = v.lm_stride[j];
        const int L = vok ? int(v.lm_cnt[j]) : 0;
        if (LW == 32 && L > 32) continue;  // (warp-uniform) a long vertex: kernels_phong_long.cu
        VertexCtx c;
        load_vertex(v, q, j, c);
        PhObs ob;
        ob.f = -1;
        const bool act = lv < L;
        double V21[21], t6[6], r7[7];
        if (act) {
            eval_phong_obs(v, q, e0 + lv * es, c, ob);
            r7[0] = ob.rs[0], r7[1] = ob.rs[1], r7[2] = ob.rs[2], r7[3] = ob.rI;
            r7[4] = ob.rN[0], r7[5] = ob.rN[1], r7[6] = ob.rN[2];
            vertex_normal_eq(ob, r7, V21, t6);
        } else {
#pragma unroll
            for (int k = 0; k < 21; ++k) V21[k] = 0.0;
        }
        // diagonal of V = sum A_v^T A_v: packed entries 0, 6, 11, 15, 18, 20
        double d2[6];
        d2[0] = seg_sum<LW>(V21[0]);
        d2[1] = seg_sum<LW>(V21[6]);
        d2[2] = seg_sum<LW>(V21[11]);
        d2[3] = seg_sum<LW>(V21[15]);
        d2[4] = seg_sum<LW>(V21[18]);
        d2[5] = seg_sum<LW>(V21[20]);
        double tg[6], ty[6];
#pragma unroll
        for (int a = 0; a < 6; ++a) {
            d2[a] = fmin(fmax(d2[a], dg.min_diag), dg.max_diag);
            const double g = gv[6ll * j + a], y = yv[6ll * j + a];
            tg[a] = g / d2[a];
            ty[a] = y;
            if (lv == 0 && vok) {
                diag_v_out[6ll * j + a] = d2[a];
                sc_v_out[6ll * j + a] = a < 3 ? c.sl[a] : c.sn[a - 3];
                acc[0] += g * g / d2[a];
                acc[1] -= g * y;
                acc[2] += d2[a] * y * y;
            }
        }
        if (act) {
            double jg[7], jy[7];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                jg[k] = ob.S[3 * k] * tg[0] + ob.S[3 * k + 1] * tg[1] + ob.S[3 * k + 2] * tg[2];
                jy[k] = ob.S[3 * k] * ty[0] + ob.S[3 * k + 1] * ty[1] + ob.S[3 * k + 2] * ty[2];
                jg[4 + k] = ob.N[3 * k] * tg[3] + ob.N[3 * k + 1] * tg[4] + ob.N[3 * k + 2] * tg[5];
                jy[4 + k] = ob.N[3 * k] * ty[3] + ob.N[3 * k + 1] * ty[4] + ob.N[3 * k + 2] * ty[5];
            }
            jg[3] = ob.ip[0] * tg[0] + ob.ip[1] * tg[1] + ob.ip[2] * tg[2] + ob.in[0] * tg[3] + ob.in[1] * tg[4] + ob.in[2] * tg[5];
            jy[3] = ob.ip[0] * ty[0] + ob.ip[1] * ty[1] + ob.ip[2] * ty[2] + ob.in[0] * ty[3] + ob.in[1] * ty[4] + ob.in[2] * ty[5];
            if (ob.f >= 0) {
                const double* g6 = gp + 6ll * ob.f;
                const double* d6 = diag_p + 6ll * ob.f;
                const double* y6 = yp + 6ll * ob.f;
#pragma unroll
                for (int a = 0; a < 6; ++a) {
                    const double ga = g6[a] / d6[a], ya = y6[a];
                    jg[0] += ob.Jcs[a] * ga, jy[0] += ob.Jcs[a] * ya;
                    jg[1] += ob.Jcs[6 + a] * ga, jy[1] += ob.Jcs[6 + a] * ya;
                    jg[2] += ob.Jcs[12 + a] * ga, jy[2] += ob.Jcs[12 + a] * ya;
                    jg[3] += ob.JIc[a] * ga, jy[3] += ob.JIc[a] * ya;
                    jg[4] += ob.JNc[a] * ga, jy[4] += ob.JNc[a] * ya;
                    jg[5] += ob.JNc[6 + a] * ga, jy[5] += ob.JNc[6 + a] * ya;
                    jg[6] += ob.JNc[12 + a] * ga, jy[6] += ob.JNc[12 + a] * ya;
                }
            }
#pragma unroll
            for (int k = 0; k < 7; ++k) {
                const int gi = c.gi[k];
                jg[3] += ob.ag[k] * (gg[gi] / diag_g[gi]);
                jy[3] += ob.ag[k] * yg[gi];
            }
#pragma unroll
            for (int k = 0; k < 7; ++k) {
                acc[3] += jg[k] * jg[k];
                acc[4] += jg[k] * jy[k];
                acc[5] += jy[k] * jy[k];
                acc[6] += jg[k] * r7[k];
                acc[7] += jy[k] * r7[k];
            }
        }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) block_atomic_sum(acc[k], &sums[k], s_red);
}

// shared-block terms of g'.g', g'.gn', gn'.gn'
__global__ void phong_dogleg_global_kernel(int n_g, const int* __restrict__ g_used, const double* __restrict__ gg,
                                           const double* __restrict__ diag_g, const double* __restrict__ yg,
                                           double* __restrict__ sums) {
    double a0 = 0, a1 = 0, a2 = 0;
    for (int k = threadIdx.x; k < n_g; k += blockDim.x)
        if (g_used[k]) {
            const double g = gg[k], d2 = diag_g[k], y = yg[k];
            a0 += g * g / d2;
            a1 -= g * y;
            a2 += d2 * y * y;
        }
    a0 = warp_sum(a0), a1 = warp_sum(a1), a2 = warp_sum(a2);
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&sums[0], a0);
        atomicAdd(&sums[1], a1);
        atomicAdd(&sums[2], a2);
    }
}

// Candidate vertices x (+) alpha * delta and the cost there (poses_cand / gx_cand already formed).
template <int LW>
__global__ void __launch_bounds__(PB_WARPS * 32)
    phong_candidate_kernel(DevView v, PhongSolveView q, int lm_lo, int lm_hi, double alpha, const double* __restrict__ yv,
                           const double* __restrict__ poses_cand, const double* __restrict__ gx_cand,
                           double* __restrict__ points_cand, double* __restrict__ normals_cand, double* __restrict__ scal2) {
    __shared__ double s_red[32];
    constexpr int NSUB = 32 / LW;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, lv = lane % LW, sub = lane / LW;
    double ccost = 0.0, sn = 0.0, xn = 0.0;
    const int warps_total = gridDim.x * PB_WARPS;
    for (int jb = lm_lo + (blockIdx.x * PB_WARPS + wib) * NSUB; jb < lm_hi; jb += warps_total * NSUB) {
        const bool vok = jb + sub < lm_hi;
        const int j = vok ? jb + sub : lm_hi - 1;
        const long long e0 = v.lm_base[j], es = v.lm_stride[j];
        const int L = vok ? int(v.lm_cnt[j]) : 0;
        if (LW == 32 && L > 32) continue;  // (warp-uniform) a long vertex: kernels_phong_long.cu
        VertexCtx c;
        load_vertex(v, q, j, c);
        double pn[3], dn[3], nn[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            pn[k] = c.p[k] + alpha * (-yv[6ll * j + k] * c.sl[k]);
            dn[k] = alpha * (-yv[6ll * j + 3 + k] * c.sn[k]);
        }
        unit_plus(c.n, dn, nn);
        if (lv == 0 && vok) {
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                points_cand[3ll * j + k] = pn[k];
                normals_cand[3ll * j + k] = nn[k];
                sn += (c.p[k] - pn[k]) * (c.p[k] - pn[k]) + (c.n[k] - nn[k]) * (c.n[k] - nn[k]);
                xn += (q.hold_positions ? 0.0 : pn[k] * pn[k]) + nn[k] * nn[k];
            }
        }
        if (lv < L) {
            double phong[3], light[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                phong[k] = gx_cand[c.gi[k]];
                light[k] = gx_cand[c.gi[4 + k]];
            }
            ccost += phong_obs_cost(v, q, e0 + lv * es, poses_cand, pn, nn, phong, gx_cand[c.gi[3]], light);
        }
    }
    block_atomic_sum(ccost, &scal2[SC_CAND_COST], s_red);
    block_atomic_sum(sn, &scal2[SC_STEP_NORM2], s_red);
    block_atomic_sum(xn, &scal2[SC_XNORM2], s_red);
}

// Poses x (+) alpha * delta (launch_pose_plus with a step length)
__global__ void phong_pose_plus_kernel(DevView v, double alpha, const double* __restrict__ yp, double* __restrict__ poses_cand,
                                       double* __restrict__ scal2, int count) {
    __shared__ double s_red[32];
    const int cidx = blockIdx.x * blockDim.x + threadIdx.x;
    double sn = 0, xn = 0, bad = 0;
    if (cidx < v.n_cams) {
        const int ff = v.cam_free[cidx];
        const double* x = v.poses + 12ll * cidx;
        double* y = poses_cand + 12ll * cidx;
        if (ff >= 0) {
            double eps[6], out[12];
            for (int k = 0; k < 6; ++k) {
                eps[k] = alpha * (-yp[6ll * ff + k] * v.sc_p[6ll * ff + k]);
                if (isnan(eps[k]) || isinf(eps[k])) bad = 1;
            }
            se3_plus(x, eps, out);
            for (int k = 0; k < 12; ++k) {
                y[k] = out[k];
                sn += (x[k] - out[k]) * (x[k] - out[k]);
                xn += out[k] * out[k];
            }
        } else {
            for (int k = 0; k < 12; ++k) y[k] = x[k];
        }
    }
    if (!count) sn = xn = 0;
    block_atomic_sum(sn, &scal2[SC_STEP_NORM2], s_red);
    block_atomic_sum(xn, &scal2[SC_XNORM2], s_red);
    block_atomic_sum(bad, &scal2[SC_NONFINITE], s_red);
}

// ParameterBlock::Plus on the shared blocks: x + alpha * delta projected onto the box (materials,
// textures), UnitVectorPerturbation for a directional light.  One block.
__device__ __forceinline__ double project_global(const PhongSolveView& q, int k, double x) {
    if (k < 3 * q.n_mat) return fmin(fmax(x, q.mat_lo[k % 3]), q.mat_hi[k % 3]);
    if (k < 3 * q.n_mat + q.n_tex) return fmin(fmax(x, q.tex_lo), q.tex_hi);
    return x;
}
__global__ void phong_global_plus_kernel(PhongSolveView q, double alpha, double sign, const double* __restrict__ step,
                                         int step_is_gradient, double* __restrict__ gx_out, double* __restrict__ scal,
                                         int slot_step, int slot_xnorm, int slot_bad, int slot_max, int count) {
    // step_is_gradient: delta = -g / sc (gradient-norm point); else delta = alpha * (-y * sc)
    __shared__ double s_red[32];
    const int k = threadIdx.x;
    const int l0 = 3 * q.n_mat + q.n_tex;
    double sn = 0, xn = 0, bad = 0, mx = 0;
    if (k < q.n_g) {
        const double x = q.gx[k];
        double out = x;
        if (q.g_used[k]) {
            const double d = step_is_gradient ? sign * step[k] / q.sc_g[k] : alpha * (sign * step[k] * q.sc_g[k]);
            if (isnan(d) || isinf(d)) bad = 1;
            if (k >= l0 && q.directional) {
                double dl[3], o3[3];
                for (int a = 0; a < 3; ++a)
                    dl[a] = step_is_gradient ? sign * step[l0 + a] / q.sc_g[l0 + a] : alpha * (sign * step[l0 + a] * q.sc_g[l0 + a]);
                unit_plus(q.gx + l0, dl, o3);
                out = o3[k - l0];
            } else {
                out = project_global(q, k, x + d);
            }
            sn = (x - out) * (x - out);
            xn = step_is_gradient ? x * x : out * out;
            mx = fabs(x - out);
        }
        if (gx_out) gx_out[k] = out;
    }
    if (!count) sn = xn = mx = 0;
    if (slot_max >= 0) {
        for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        if ((threadIdx.x & 31) == 0 && mx > 0) atomic_max_nonneg(&scal[slot_max], mx);
    }
    if (slot_step >= 0) block_atomic_sum(sn, &scal[slot_step], s_red);
    if (slot_xnorm >= 0) block_atomic_sum(xn, &scal[slot_xnorm], s_red);
    if (slot_bad >= 0) block_atomic_sum(bad, &scal[slot_bad], s_red);
}

// |x - Plus(x, -g)|_inf and |x|^2 over the vertex blocks (thread per vertex)
__global__ void phong_gradnorm_kernel(DevView v, PhongSolveView q, int lm_lo, int lm_hi, const double* __restrict__ gv,
                                      double* __restrict__ scal) {
    __shared__ double s_red[32];
    const int j = lm_lo + blockIdx.x * blockDim.x + threadIdx.x;
    double m = 0, xn = 0;
    if (j < lm_hi) {
        double n[3], dn[3], nn[3];
        for (int k = 0; k < 3; ++k) {
            const double x = v.points[3ll * j + k];
            const double g = gv[6ll * j + k] / v.sc_l[3ll * j + k];
            m = fmax(m, fabs(x - (x - g)));
            n[k] = q.normals[3ll * j + k];
            dn[k] = -gv[6ll * j + 3 + k] / q.sc_n[3ll * j + k];
            xn += (q.hold_positions ? 0.0 : x * x) + n[k] * n[k];
        }
        unit_plus(n, dn, nn);
        for (int k = 0; k < 3; ++k) m = fmax(m, fabs(n[k] - nn[k]));
    }
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0) atomic_max_nonneg(&scal[SC_GRADMAX], m);
    block_atomic_sum(xn, &scal[SC_XNORM2_CUR], s_red);
}

// IterationZero of a bounded problem: x <- Plus(x, 0) (normals re-normalised; globals projected by
// phong_global_plus_kernel)
__global__ void phong_renormalize_kernel(int n_lm, double* __restrict__ normals) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_lm) return;
    const double z[3] = {0, 0, 0};
    double n[3] = {normals[3ll * j], normals[3ll * j + 1], normals[3ll * j + 2]}, o[3];
    unit_plus(n, z, o);
    for (int k = 0; k < 3; ++k) normals[3ll * j + k] = o[k];
}

// sum_i a[i] * b[i] into dst (grid-stride)
__global__ void dot_kernel(const double* __restrict__ a, const double* __restrict__ b, long long n, double* dst) {
    __shared__ double s_red[32];
    double s = 0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        s += a[i] * b[i];
    block_atomic_sum(s, dst, s_red);
}
// max_i |y[i] * sc[i]| into dst (bit-pattern max)
__global__ void absmax_scaled_kernel(const double* __restrict__ y, const double* __restrict__ sc, long long n, double* dst) {
    double m = 0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        m = fmax(m, fabs(y[i] * sc[i]));
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0) atomic_max_nonneg(dst, m);
}

// ---- the dense border --------------------------------------------------------------------------
// LM diagonal on S_gg; unused columns become identity rows.  One block.
__global__ void phong_gfinalize_kernel(PhongSolveView q, LmDiag dg, double* __restrict__ Sgg, double* __restrict__ bg,
                                       const double* __restrict__ hg, double* __restrict__ diag_g) {
    const int k = threadIdx.x;
    if (k >= q.n_g) return;
    if (!q.g_used[k]) {
        Sgg[(long long)k * q.n_g + k] = 1.0;
        bg[k] = 0.0;
        diag_g[k] = 1.0;
        return;
    }
    const double dd = fmin(fmax(hg[k], dg.min_diag), dg.max_diag);
    diag_g[k] = dd;
    Sgg[(long long)k * q.n_g + k] += dd * dg.inv_radius;
}

// T[q][q2] = S_gg[q][q2] - S_cg[:,q] . X[:,q2]   (q2 == n_g: the right-hand side column).
// X holds n_g + 1 solves against S_cc: columns 0..n_g-1 for the border, column n_g for b_c.
__global__ void __launch_bounds__(256)
    phong_border_reduce_kernel(int n_g, int nf6, const double* __restrict__ Scg, const double* __restrict__ X,
                               const double* __restrict__ Sgg, const double* __restrict__ bg, double* __restrict__ T) {
    __shared__ double s_red[32];
    const int qa = blockIdx.x / (n_g + 1), qb = blockIdx.x % (n_g + 1);
    const double* a = Scg + (long long)qa * nf6;
    const double* b = X + (long long)qb * nf6;
    double s = 0;
    for (int i = threadIdx.x; i < nf6; i += blockDim.x) s += a[i] * b[i];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0;
        for (int w = 0; w < (blockDim.x >> 5); ++w) t += s_red[w];  // fixed order: deterministic
        const double base = qb < n_g ? Sgg[(long long)qa * n_g + qb] : bg[qa];
        T[(long long)qa * (n_g + 1) + qb] = base - t;
    }
}

// Dense Cholesky solve of the n_g x n_g border system (one block); yg out, fail flag.  The system lives in the
// CTA's shared memory when it fits (n_g <= 160); a larger border (hundreds of materials / textures) is factored in
// place in global memory — L2-resident, one CTA of 1024 threads — which is slower per column but has no size limit.
__global__ void __launch_bounds__(1024) phong_border_solve_kernel(int n_g, double* __restrict__ T, double* __restrict__ yg,
                                                                  double* __restrict__ ps, int in_shared) {
    extern __shared__ double sT_sh[];  // [n_g][n_g + 1]
    const int ld = n_g + 1;
    double* sT = in_shared ? sT_sh : T;
    if (in_shared)
        for (int i = threadIdx.x; i < n_g * ld; i += blockDim.x) sT[i] = T[i];
    __shared__ int ok;
    if (threadIdx.x == 0) ok = 1;
    __syncthreads();
    // symmetrise from the accumulated (numerically almost symmetric) matrix: use the lower triangle
    for (int j = 0; j < n_g; ++j) {
        if (threadIdx.x == 0) {
            const double d = sT[j * ld + j];
            if (!(d > 0.0) || !(d < 1.7976931348623157e308)) ok = 0;
            sT[j * ld + j] = sqrt(d);
        }
        __syncthreads();
        const double djj = sT[j * ld + j];
        for (int i = j + 1 + threadIdx.x; i < n_g; i += blockDim.x) sT[i * ld + j] /= djj;
        __syncthreads();
        for (int idx = threadIdx.x; idx < (n_g - j - 1) * (n_g - j - 1); idx += blockDim.x) {
            const int i = j + 1 + idx / (n_g - j - 1), k = j + 1 + idx % (n_g - j - 1);
            if (k <= i) sT[i * ld + k] -= sT[i * ld + j] * sT[k * ld + j];
        }
        __syncthreads();
    }
    // forward / backward substitution on the rhs column (index n_g), column oriented: one pivot, then every row
    for (int j = 0; j < n_g; ++j) {
        if (threadIdx.x == 0) sT[j * ld + n_g] /= sT[j * ld + j];
        __syncthreads();
        const double yj = sT[j * ld + n_g];
        for (int i = j + 1 + threadIdx.x; i < n_g; i += blockDim.x) sT[i * ld + n_g] -= sT[i * ld + j] * yj;
        __syncthreads();
    }
    for (int j = n_g - 1; j >= 0; --j) {
        if (threadIdx.x == 0) sT[j * ld + n_g] /= sT[j * ld + j];
        __syncthreads();
        const double yj = sT[j * ld + n_g];
        for (int i = threadIdx.x; i < j; i += blockDim.x) sT[i * ld + n_g] -= sT[j * ld + i] * yj;
        __syncthreads();
    }
    for (int i = threadIdx.x; i < n_g; i += blockDim.x) yg[i] = sT[i * ld + n_g];
    if (threadIdx.x == 0 && !ok) ps[PS_FAIL] = 2.0;
}

// y_c = x_b - X_g y_g
__global__ void phong_border_backsub_kernel(int n_g, int nf6, const double* __restrict__ X, const double* __restrict__ yg,
                                            double* __restrict__ yc) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nf6) return;
    double s = X[(long long)n_g * nf6 + i];
    for (int k = 0; k < n_g; ++k) s -= X[(long long)k * nf6 + i] * yg[k];
    yc[i] = s;
}

inline void count_launch() { g_kernel_launches.fetch_add(1, std::memory_order_relaxed); }
inline int vertex_grid(int n) {
    const int blocks = (n + PB_WARPS - 1) / PB_WARPS;
    return blocks < 1 ? 1 : (blocks < 148 * 8 ? blocks : 148 * 8);
}

}  // namespace

void launch_phong_build(cudaStream_t s, const DevView& v, const PhongSolveView& q, int lm_lo, int lm_hi, LmDiag dg,
                        const PhongSystem& o, bool schur, int max_track_len) {
    if (lm_hi <= lm_lo) return;
    // tracks of at most 16 observations: two vertices per warp
    const bool packed = max_track_len <= 16;
    const int grid = vertex_grid(packed ? (lm_hi - lm_lo + 1) / 2 : lm_hi - lm_lo);
    if (schur) {
        if (packed)
            phong_build_kernel<true, 16><<<grid, PB_WARPS * 32, 0, s>>>(v, q, lm_lo, lm_hi, dg, o);
        else
            phong_build_kernel<true, 32><<<grid, PB_WARPS * 32, 0, s>>>(v, q, lm_lo, lm_hi, dg, o);
    } else if (packed) {
        phong_build_kernel<false, 16><<<grid, PB_WARPS * 32, 0, s>>>(v, q, lm_lo, lm_hi, dg, o);
    } else {
        phong_build_kernel<false, 32><<<grid, PB_WARPS * 32, 0, s>>>(v, q, lm_lo, lm_hi, dg, o);
    }
    count_launch();
    CSLAM_CUDA(cudaGetLastError());
    if (max_track_len > 32) launch_phong_build_long(s, v, q, lm_lo, lm_hi, dg, o, schur);
}

void launch_phong_gfinalize(cudaStream_t s, const PhongSolveView& q, LmDiag dg, double* Sgg, double* bg, const double* hg,
                            double* diag_g) {
    phong_gfinalize_kernel<<<1, ((q.n_g + 31) / 32) * 32, 0, s>>>(q, dg, Sgg, bg, hg, diag_g);
    count_launch();
    CSLAM_CUDA(cudaGetLastError());
}

void launch_phong_border_solve(cudaStream_t s, int n_g, int nf6, const double* Scg, const double* X, const double* Sgg,
                               const double* bg, double* T, double* yg, double* yc, double* ps) {
    if (nf6 > 0) {
        phong_border_reduce_kernel<<<n_g * (n_g + 1), 256, 0, s>>>(n_g, nf6, Scg, X, Sgg, bg, T);
        count_launch();
    } else {
        // no free camera: T = [S_gg | b_g]
        CSLAM_CUDA(cudaMemcpy2DAsync(T, (n_g + 1) * sizeof(double), Sgg, n_g * sizeof(double), n_g * sizeof(double), n_g,
                                     cudaMemcpyDeviceToDevice, s));
        CSLAM_CUDA(cudaMemcpy2DAsync(T + n_g, (n_g + 1) * sizeof(double), bg, sizeof(double), sizeof(double), n_g,
                                     cudaMemcpyDeviceToDevice, s));
    }
    const bool in_shared = n_g <= 160;
    const size_t smem = in_shared ? size_t(n_g) * (n_g + 1) * sizeof(double) : 0;
    if (smem > 48 * 1024)   // per device: no process-wide cache
        CSLAM_CUDA(cudaFuncSetAttribute(phong_border_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    phong_border_solve_kernel<<<1, in_shared ? 128 : 1024, smem, s>>>(n_g, T, yg, ps, in_shared ? 1 : 0);
    count_launch();
    if (nf6 > 0) {
        phong_border_backsub_kernel<<<(nf6 + 255) / 256, 256, 0, s>>>(n_g, nf6, X, yg, yc);
        count_launch();
    }
    CSLAM_CUDA(cudaGetLastError());
}

void launch_phong_backsub(cudaStream_t s, const DevView& v, const PhongSolveView& q, int lm_lo, int lm_hi, LmDiag dg,
                          const double* yp, const double* yg, const double* gv, double* yv, double* scal2, int max_track_len) {
    if (lm_hi <= lm_lo) return;
    if (max_track_len <= 16)
        phong_backsub_kernel<16><<<vertex_grid((lm_hi - lm_lo + 1) / 2), PB_WARPS * 32, 0, s>>>(v, q, lm_lo, lm_hi, dg, yp, yg, gv, yv, scal2);
    else
        phong_backsub_kernel<32><<<vertex_grid(lm_hi - lm_lo), PB_WARPS * 32, 0, s>>>(v, q, lm_lo, lm_hi, dg, yp, yg, gv, yv, scal2);
    count_launch();
    CSLAM_CUDA(cudaGetLastError());
    if (max_track_len > 32) launch_phong_backsub_long(s, v, q, lm_lo, lm_hi, dg, yp, yg, gv, yv, scal2);
}

void launch_phong_dogleg_products(cudaStream_t s, const DevView& v, const PhongSolveView& q, int lm_lo, int lm_hi, LmDiag dg,
                                  const double* gp, const double* diag_p, const double* yp, const double* gg, const double* diag_g,
                                  const double* yg, const double* gv, const double* yv, double* diag_v, double* sc_v, double* sums,
                                  int max_track_len, int count_shared) {
    if (lm_hi > lm_lo) {
        if (max_track_len <= 16)
            phong_dogleg_products_kernel<16><<<vertex_grid((lm_hi - lm_lo + 1) / 2), PB_WARPS * 32, 0, s>>>(
                v, q, lm_lo, lm_hi, dg, gp, diag_p, yp, gg, diag_g, yg, gv, yv, diag_v, sc_v, sums);
        else
            phong_dogleg_products_kernel<32><<<vertex_grid(lm_hi - lm_lo), PB_WARPS * 32, 0, s>>>(
                v, q, lm_lo, lm_hi, dg, gp, diag_p, yp, gg, diag_g, yg, gv, yv, diag_v, sc_v, sums);
        count_launch();
        if (max_track_len > 32)
            launch_phong_dogleg_products_long(s, v, q, lm_lo, lm_hi, dg, gp, diag_p, yp, gg, diag_g, yg, gv, yv, diag_v, sc_v, sums);
    }
    if (count_shared) {
        phong_dogleg_global_kernel<<<1, 128, 0, s>>>(q.n_g, q.g_used, gg, diag_g, yg, sums);
        count_launch();
    }
    CSLAM_CUDA(cudaGetLastError());
}

void launch_phong_candidate(cudaStream_t s, const DevView& v, const PhongSolveView& q, int lm_lo, int lm_hi, double alpha,
                            const double* yp, const double* yg, const double* yv, double* poses_cand, double* gx_cand,
                            double* points_cand, double* normals_cand, double* scal2, int count_shared, int max_track_len) {
    phong_pose_plus_kernel<<<(v.n_cams + 127) / 128, 128, 0, s>>>(v, alpha, yp, poses_cand, scal2, count_shared);
    count_launch();
    phong_global_plus_kernel<<<1, ((q.n_g + 31) / 32) * 32, 0, s>>>(q, alpha, -1.0, yg, 0, gx_cand, scal2, SC_STEP_NORM2,
                                                                     SC_XNORM2, SC_NONFINITE, -1, count_shared);
    count_launch();
    if (lm_hi > lm_lo) {
        if (max_track_len <= 16)
            phong_candidate_kernel<16><<<vertex_grid((lm_hi - lm_lo + 1) / 2), PB_WARPS * 32, 0, s>>>(
                v, q, lm_lo, lm_hi, alpha, yv, poses_cand, gx_cand, points_cand, normals_cand, scal2);
        else
            phong_candidate_kernel<32><<<vertex_grid(lm_hi - lm_lo), PB_WARPS * 32, 0, s>>>(
                v, q, lm_lo, lm_hi, alpha, yv, poses_cand, gx_cand, points_cand, normals_cand, scal2);
        count_launch();
        if (max_track_len > 32)
            launch_phong_candidate_long(s, v, q, lm_lo, lm_hi, alpha, yv, poses_cand, gx_cand, points_cand, normals_cand, scal2);
    }
    CSLAM_CUDA(cudaGetLastError());
}

void launch_phong_gradnorm(cudaStream_t s, const DevView& v, const PhongSolveView& q, int lm_lo, int lm_hi, const double* gv,
                           const double* gg, double* scal, int count_shared) {
    if (lm_hi > lm_lo) {
        phong_gradnorm_kernel<<<(lm_hi - lm_lo + 127) / 128, 128, 0, s>>>(v, q, lm_lo, lm_hi, gv, scal);
        count_launch();
    }
    phong_global_plus_kernel<<<1, ((q.n_g + 31) / 32) * 32, 0, s>>>(q, 1.0, -1.0, gg, 1, nullptr, scal, -1, SC_XNORM2_CUR, -1,
                                                                     SC_GRADMAX, count_shared);
    count_launch();
    CSLAM_CUDA(cudaGetLastError());
}

void launch_phong_project_initial(cudaStream_t s, const PhongSolveView& q, int n_lm, double* normals, double* gx, double* zero_g) {
    if (n_lm > 0) {
        phong_renormalize_kernel<<<(n_lm + 127) / 128, 128, 0, s>>>(n_lm, normals);
        count_launch();
    }
    // Plus(x, 0) on the shared blocks: projection onto the box / re-normalised light direction
    phong_global_plus_kernel<<<1, ((q.n_g + 31) / 32) * 32, 0, s>>>(q, 1.0, 1.0, zero_g, 0, gx, nullptr, -1, -1, -1, -1, 0);
    count_launch();
    CSLAM_CUDA(cudaGetLastError());
}

void launch_dot(cudaStream_t s, const double* a, const double* b, long long n, double* dst) {
    if (n <= 0) return;
    const int grid = int(std::min<long long>((n + 255) / 256, 148 * 4));
    dot_kernel<<<grid, 256, 0, s>>>(a, b, n, dst);
    count_launch();
    CSLAM_CUDA(cudaGetLastError());
}
void launch_absmax_scaled(cudaStream_t s, const double* y, const double* sc, long long n, double* dst) {
    if (n <= 0) return;
    const int grid = int(std::min<long long>((n + 255) / 256, 148 * 4));
    absmax_scaled_kernel<<<grid, 256, 0, s>>>(y, sc, n, dst);
    count_launch();
    CSLAM_CUDA(cudaGetLastError());
}

}  // namespace cslam
