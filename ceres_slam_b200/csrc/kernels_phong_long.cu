// K2p / K4p for vertices observed MORE than 32 times (dataset_ba_phong.cpp:110-140 puts no bound on a track).
//
// kernels_phong_solve.cu gives every vertex one warp and every observation one lane, which keeps the observation's
// blocks in registers from the reduction over the vertex to its camera-pair products — and stops at 32 observations.
// The kernels here take the long vertices of the same range (they skip every vertex the lane-per-observation kernels
// take, and those skip these): still one warp per vertex, the observations walked in chunks of 32.  The sums over the
// vertex are accumulated per lane across the chunks and reduced once; what the second half of each kernel needs per
// observation is evaluated again (a long vertex is rare: this path is about being exact, not fast); the camera-pair
// products of the elimination run chunk against chunk, the second chunk's W staged in shared memory as before.
// Same closed forms, same arithmetic per observation, FP64.
#include "kernels.cuh"
#include "phong_common.cuh"

namespace cslam {

namespace {

__device__ __forceinline__ void fill_r7(const PhObs& ob, double* r7) {
    r7[0] = ob.rs[0], r7[1] = ob.rs[1], r7[2] = ob.rs[2], r7[3] = ob.rI;
    r7[4] = ob.rN[0], r7[5] = ob.rN[1], r7[6] = ob.rN[2];
}

// W = A_c^T A_v (6 x 6) of one observation
__device__ __forceinline__ void obs_W(const PhObs& ob, double* W) {
#pragma unroll
    for (int a = 0; a < 6; ++a) {
#pragma unroll
        for (int b = 0; b < 3; ++b) {
            W[6 * a + b] = ob.Jcs[a] * ob.S[b] + ob.Jcs[6 + a] * ob.S[3 + b] + ob.Jcs[12 + a] * ob.S[6 + b] + ob.JIc[a] * ob.ip[b];
            W[6 * a + 3 + b] = ob.JIc[a] * ob.in[b] + ob.JNc[a] * ob.N[b] + ob.JNc[6 + a] * ob.N[3 + b] + ob.JNc[12 + a] * ob.N[6 + b];
        }
    }
}

// (J y) restricted to the camera and shared-block columns
__device__ __forceinline__ void obs_Jy(const PhObs& ob, const VertexCtx& c, const double* __restrict__ yp,
                                       const double* __restrict__ yg, double* Jy) {
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
}

// S block (a <= b in free-camera order) -= Y_x W_y^T; the same camera seen twice by the vertex adds both orders
__device__ __forceinline__ void pair_update(const DevView& v, double* __restrict__ S, int fx, int fy, const double* Y,
                                            const double* Wy, bool both_orders) {
    const bool swap = fx > fy;
    const int a = swap ? fy : fx, b = swap ? fx : fy;
    double* B = S + 36ll * find_block(v.s_rowptr, v.s_col, a, b);
    if (fx == fy) {
#pragma unroll
        for (int pp = 0; pp < 6; ++pp)
#pragma unroll
            for (int qq = pp; qq < 6; ++qq) {
                double val = 0.0;
#pragma unroll
                for (int k = 0; k < 6; ++k) val += Y[6 * pp + k] * Wy[(6 * qq + k) * 32];
                if (both_orders)
#pragma unroll
                    for (int k = 0; k < 6; ++k) val += Y[6 * qq + k] * Wy[(6 * pp + k) * 32];
                red_add(&B[6 * pp + qq], -val);
            }
    } else {
#pragma unroll
        for (int pp = 0; pp < 6; ++pp)
#pragma unroll
            for (int qq = 0; qq < 6; ++qq) {
                double val = 0.0;
#pragma unroll
                for (int k = 0; k < 6; ++k) val += Y[6 * pp + k] * Wy[(6 * qq + k) * 32];
                red_add(swap ? &B[6 * qq + pp] : &B[6 * pp + qq], -val);
            }
    }
}

template <bool kSchur>
__global__ void __launch_bounds__(PB_WARPS * 32)
    phong_build_long_kernel(DevView v, PhongSolveView q, int lm_lo, int lm_hi, LmDiag dg, PhongSystem o) {
    __shared__ double sW[PB_WARPS][36 * 32];  // W of every lane of the staged chunk, [k][lane]
    __shared__ double sG[PB_WARPS][72];
    __shared__ int sF[PB_WARPS][32];
    __shared__ double s_red[32];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    double* myW = sW[wib];
    double* myG = sG[wib];
    int* myF = sF[wib];
    const int nf6 = 6 * v.n_free;
    double cost = 0.0, fixed = 0.0;
    for (int j = lm_lo + blockIdx.x * PB_WARPS + wib; j < lm_hi; j += gridDim.x * PB_WARPS) {
        const int L = int(v.lm_cnt[j]);
        if (L <= 32) continue;  // (warp-uniform) taken by phong_build_kernel
        const long long e0 = v.lm_base[j], es = v.lm_stride[j];
        VertexCtx c;
        load_vertex(v, q, j, c);
        // ---- sums over the vertex: V, g_v, the shared-block gradient / diagonal / pair sums, G ----
        double V21[21], gv[6], ggl[7], hd[7], G[42], hp[21];
#pragma unroll
        for (int k = 0; k < 21; ++k) V21[k] = hp[k] = 0.0;
#pragma unroll
        for (int k = 0; k < 6; ++k) gv[k] = 0.0;
#pragma unroll
        for (int k = 0; k < 7; ++k) ggl[k] = hd[k] = 0.0;
#pragma unroll
        for (int k = 0; k < 42; ++k) G[k] = 0.0;
        for (int x = lane; x < L; x += 32) {
            PhObs ob;
            eval_phong_obs(v, q, e0 + x * es, c, ob);
            double r7[7], t21[21], t6[6];
            fill_r7(ob, r7);
#pragma unroll
            for (int k = 0; k < 7; ++k) cost += 0.5 * r7[k] * r7[k];
            if (!kSchur && q.hold_positions && ob.f < 0) {
                const long long e = e0 + x * es;
                double rs[3];
                stereo_block<false>(v.cam, v.poses + 12ll * v.obs_cam[e], c.p, v.obs_u[e], v.obs_v[e], v.obs_d[e], v.obs_W, rs,
                                    nullptr, nullptr);
                fixed += 0.5 * (rs[0] * rs[0] + rs[1] * rs[1] + rs[2] * rs[2]);
            }
            vertex_normal_eq(ob, r7, t21, t6);
#pragma unroll
            for (int k = 0; k < 21; ++k) V21[k] += t21[k];
#pragma unroll
            for (int k = 0; k < 6; ++k) gv[k] += t6[k];
#pragma unroll
            for (int k = 0; k < 7; ++k) {
                ggl[k] += ob.ag[k] * ob.rI;
                hd[k] += ob.ag[k] * ob.ag[k];
            }
            if (kSchur) {
                int idx = 0;
#pragma unroll
                for (int k = 0; k < 7; ++k) {
#pragma unroll
                    for (int a = 0; a < 3; ++a) {
                        G[6 * k + a] += ob.ag[k] * ob.ip[a];
                        G[6 * k + 3 + a] += ob.ag[k] * ob.in[a];
                    }
#pragma unroll
                    for (int k2 = k + 1; k2 < 7; ++k2) hp[idx++] += ob.ag[k] * ob.ag[k2];
                }
            } else if (ob.f >= 0) {
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
        }
#pragma unroll
        for (int k = 0; k < 21; ++k) V21[k] = seg_sum<32>(V21[k]);
#pragma unroll
        for (int k = 0; k < 6; ++k) gv[k] = seg_sum<32>(gv[k]);
#pragma unroll
        for (int k = 0; k < 7; ++k) {
            ggl[k] = seg_sum<32>(ggl[k]);
            hd[k] = seg_sum<32>(hd[k]);
        }
        if (!kSchur) {
            if (lane == 0) {
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
            continue;
        }
        // ---- Schur pass: the vertex block, the shared-block part ----------------------------------
#pragma unroll
        for (int k = 0; k < 42; ++k) G[k] = seg_sum<32>(G[k]);
#pragma unroll
        for (int k = 0; k < 21; ++k) hp[k] = seg_sum<32>(hp[k]);
        double V[36], Vi[36];
        unpack_sym6(V21, V);
        add_lm_diag(V, dg);
        const bool pd = spd6_inverse(V, Vi);
        if (lane == 0) {
#pragma unroll
            for (int a = 0; a < 6; ++a) o.gv[6ll * j + a] = gv[a];
            if (!pd) red_add(&o.scal[SC_INVALID], 1.0);
        }
        if (!pd) continue;  // (warp-uniform)
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
        __syncwarp();
        if (lane == 0) {
            int idx = 0, ip = 0;
#pragma unroll
            for (int k = 0; k < 7; ++k)
#pragma unroll
                for (int k2 = k; k2 < 7; ++k2) {
                    double h = (k2 == k) ? hd[k] : hp[ip++];
#pragma unroll
                    for (int a = 0; a < 6; ++a) h -= GV[6 * k + a] * G[6 * k2 + a];
                    myG[idx++] = h;
                }
#pragma unroll
            for (int k = 0; k < 7; ++k) {
                double bb = ggl[k];
#pragma unroll
                for (int a = 0; a < 6; ++a) bb -= GV[6 * k + a] * gv[a];
                myG[28 + k] = bb;
                myG[35 + k] = ggl[k];
                myG[42 + k] = hd[k];
            }
        }
        __syncwarp();
        if (lane < 28) {
            int k = 0, rem = lane;
            while (rem >= 7 - k) {
                rem -= 7 - k;
                ++k;
            }
            const int k2 = k + rem;
            const double h = myG[lane];
            const int gk = c.gi[k], gk2 = c.gi[k2];
            red_add(&o.Sgg[(long long)gk * q.n_g + gk2], h);
            if (gk != gk2) red_add(&o.Sgg[(long long)gk2 * q.n_g + gk], h);
        }
        if (lane < 7) {
            red_add(&o.bg[c.gi[lane]], myG[28 + lane]);
            red_add(&o.gg[c.gi[lane]], myG[35 + lane]);
            red_add(&o.hg[c.gi[lane]], myG[42 + lane]);
        }
        // ---- camera part: chunk bx against itself and against every later chunk ------------------------
        for (int bx = 0; bx < L; bx += 32) {
            const int x = bx + lane;
            double Y[36];
            int fx = -1;
            __syncwarp();
            if (x < L) {
                PhObs ob;
                eval_phong_obs(v, q, e0 + x * es, c, ob);
                fx = ob.f;
                if (fx >= 0) {
                    double r7[7], W[36];
                    fill_r7(ob, r7);
                    obs_W(ob, W);
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
                    double* Bd = o.Bdiag + 36ll * fx;
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
                        red_add(&o.gp[6ll * fx + a], ga);
                        red_add(&o.bp[6ll * fx + a], ga - yg);
#pragma unroll
                        for (int k = 0; k < 7; ++k) {
                            double s = ob.JIc[a] * ob.ag[k];
#pragma unroll
                            for (int b = 0; b < 6; ++b) s -= Y[6 * a + b] * G[6 * k + b];
                            red_add(&o.Scg[(long long)c.gi[k] * nf6 + 6 * fx + a], s);
                        }
                    }
                }
            }
            myF[lane] = fx;
            __syncwarp();
            // pairs inside the chunk: lane x takes (x, (x + s) mod Lc), s = 0 .. Lc/2 — every unordered pair once
            const int Lc = min(32, L - bx);
            if (fx >= 0) {
                for (int s = 0; s <= Lc / 2; ++s) {
                    if (2 * s == Lc && lane >= s) break;  // even Lc: the antipodal pairs appear twice
                    int y = lane + s;
                    if (y >= Lc) y -= Lc;
                    const int fy = myF[y];
                    if (fy < 0) continue;
                    pair_update(v, o.S, fx, fy, Y, myW + y, s != 0);
                }
            }
            // pairs with the later chunks: every (x, y) once
            for (int by = bx + 32; by < L; by += 32) {
                __syncwarp();
                const int y0 = by + lane;
                int fy0 = -1;
                if (y0 < L) {
                    PhObs oy;
                    eval_phong_obs(v, q, e0 + y0 * es, c, oy);
                    fy0 = oy.f;
                    if (fy0 >= 0) {
                        double W[36];
                        obs_W(oy, W);
#pragma unroll
                        for (int k = 0; k < 36; ++k) myW[k * 32 + lane] = W[k];
                    }
                }
                myF[lane] = fy0;
                __syncwarp();
                const int Ly = min(32, L - by);
                if (fx >= 0) {
                    for (int s = 0; s < Ly; ++s) {
                        const int y = (lane + s) % Ly;
                        const int fy = myF[y];
                        if (fy < 0) continue;
                        pair_update(v, o.S, fx, fy, Y, myW + y, true);
                    }
                }
            }
        }
        __syncwarp();
    }
    block_atomic_sum(cost, &o.scal[SC_COST], s_red);
    if (!kSchur) block_atomic_sum(fixed, &o.scal[SC_FIXED], s_red);
}

__global__ void __launch_bounds__(PB_WARPS * 32)
    phong_backsub_long_kernel(DevView v, PhongSolveView q, int lm_lo, int lm_hi, LmDiag dg, const double* __restrict__ yp,
                              const double* __restrict__ yg, const double* __restrict__ gv, double* __restrict__ yv_out,
                              double* __restrict__ scal2) {
    __shared__ double s_red[32];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    double model = 0.0, bad = 0.0, gy = 0.0, dmax = 0.0;
    for (int j = lm_lo + blockIdx.x * PB_WARPS + wib; j < lm_hi; j += gridDim.x * PB_WARPS) {
        const int L = int(v.lm_cnt[j]);
        if (L <= 32) continue;
        const long long e0 = v.lm_base[j], es = v.lm_stride[j];
        VertexCtx c;
        load_vertex(v, q, j, c);
        double V21[21], t6[6];
#pragma unroll
        for (int k = 0; k < 21; ++k) V21[k] = 0.0;
#pragma unroll
        for (int k = 0; k < 6; ++k) t6[k] = 0.0;
        for (int x = lane; x < L; x += 32) {
            PhObs ob;
            eval_phong_obs(v, q, e0 + x * es, c, ob);
            double r7[7], Jy[7], w[7], a21[21], a6[6];
            fill_r7(ob, r7);
            obs_Jy(ob, c, yp, yg, Jy);
#pragma unroll
            for (int k = 0; k < 7; ++k) w[k] = r7[k] - Jy[k];
            vertex_normal_eq(ob, w, a21, a6);
#pragma unroll
            for (int k = 0; k < 21; ++k) V21[k] += a21[k];
#pragma unroll
            for (int k = 0; k < 6; ++k) t6[k] += a6[k];
        }
#pragma unroll
        for (int k = 0; k < 21; ++k) V21[k] = seg_sum<32>(V21[k]);
#pragma unroll
        for (int k = 0; k < 6; ++k) t6[k] = seg_sum<32>(t6[k]);
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
        if (lane == 0) {
#pragma unroll
            for (int a = 0; a < 6; ++a) {
                yv_out[6ll * j + a] = yv[a];
                if (isnan(yv[a]) || isinf(yv[a])) bad = 1.0;
                gy += gv[6ll * j + a] * yv[a];
                dmax = fmax(dmax, fabs(yv[a] * (a < 3 ? c.sl[a] : c.sn[a - 3])));
            }
        }
        for (int x = lane; x < L; x += 32) {
            PhObs ob;
            eval_phong_obs(v, q, e0 + x * es, c, ob);
            double r7[7], Jy[7], m[7];
            fill_r7(ob, r7);
            obs_Jy(ob, c, yp, yg, Jy);
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

__global__ void __launch_bounds__(PB_WARPS * 32)
    phong_dogleg_products_long_kernel(DevView v, PhongSolveView q, int lm_lo, int lm_hi, LmDiag dg, const double* __restrict__ gp,
                                      const double* __restrict__ diag_p, const double* __restrict__ yp,
                                      const double* __restrict__ gg, const double* __restrict__ diag_g,
                                      const double* __restrict__ yg, const double* __restrict__ gv, const double* __restrict__ yv,
                                      double* __restrict__ diag_v_out, double* __restrict__ sc_v_out, double* __restrict__ sums) {
    __shared__ double s_red[32];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int j = lm_lo + blockIdx.x * PB_WARPS + wib; j < lm_hi; j += gridDim.x * PB_WARPS) {
        const int L = int(v.lm_cnt[j]);
        if (L <= 32) continue;
        const long long e0 = v.lm_base[j], es = v.lm_stride[j];
        VertexCtx c;
        load_vertex(v, q, j, c);
        double d2[6] = {0, 0, 0, 0, 0, 0};
        for (int x = lane; x < L; x += 32) {
            PhObs ob;
            eval_phong_obs(v, q, e0 + x * es, c, ob);
            double r7[7], a21[21], a6[6];
            fill_r7(ob, r7);
            vertex_normal_eq(ob, r7, a21, a6);
            d2[0] += a21[0], d2[1] += a21[6], d2[2] += a21[11], d2[3] += a21[15], d2[4] += a21[18], d2[5] += a21[20];
        }
        double tg[6], ty[6];
#pragma unroll
        for (int a = 0; a < 6; ++a) {
            d2[a] = fmin(fmax(seg_sum<32>(d2[a]), dg.min_diag), dg.max_diag);
            const double g = gv[6ll * j + a], y = yv[6ll * j + a];
            tg[a] = g / d2[a];
            ty[a] = y;
            if (lane == 0) {
                diag_v_out[6ll * j + a] = d2[a];
                sc_v_out[6ll * j + a] = a < 3 ? c.sl[a] : c.sn[a - 3];
                acc[0] += g * g / d2[a];
                acc[1] -= g * y;
                acc[2] += d2[a] * y * y;
            }
        }
        for (int x = lane; x < L; x += 32) {
            PhObs ob;
            eval_phong_obs(v, q, e0 + x * es, c, ob);
            double r7[7], jg[7], jy[7];
            fill_r7(ob, r7);
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

__global__ void __launch_bounds__(PB_WARPS * 32)
    phong_candidate_long_kernel(DevView v, PhongSolveView q, int lm_lo, int lm_hi, double alpha, const double* __restrict__ yv,
                                const double* __restrict__ poses_cand, const double* __restrict__ gx_cand,
                                double* __restrict__ points_cand, double* __restrict__ normals_cand, double* __restrict__ scal2) {
    __shared__ double s_red[32];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    double ccost = 0.0, sn = 0.0, xn = 0.0;
    for (int j = lm_lo + blockIdx.x * PB_WARPS + wib; j < lm_hi; j += gridDim.x * PB_WARPS) {
        const int L = int(v.lm_cnt[j]);
        if (L <= 32) continue;
        const long long e0 = v.lm_base[j], es = v.lm_stride[j];
        VertexCtx c;
        load_vertex(v, q, j, c);
        double pn[3], dn[3], nn[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            pn[k] = c.p[k] + alpha * (-yv[6ll * j + k] * c.sl[k]);
            dn[k] = alpha * (-yv[6ll * j + 3 + k] * c.sn[k]);
        }
        unit_plus(c.n, dn, nn);
        if (lane == 0) {
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                points_cand[3ll * j + k] = pn[k];
                normals_cand[3ll * j + k] = nn[k];
                sn += (c.p[k] - pn[k]) * (c.p[k] - pn[k]) + (c.n[k] - nn[k]) * (c.n[k] - nn[k]);
                xn += (q.hold_positions ? 0.0 : pn[k] * pn[k]) + nn[k] * nn[k];
            }
        }
        double phong[3], light[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            phong[k] = gx_cand[c.gi[k]];
            light[k] = gx_cand[c.gi[4 + k]];
        }
        for (int x = lane; x < L; x += 32)
            ccost += phong_obs_cost(v, q, e0 + x * es, poses_cand, pn, nn, phong, gx_cand[c.gi[3]], light);
    }
    block_atomic_sum(ccost, &scal2[SC_CAND_COST], s_red);
    block_atomic_sum(sn, &scal2[SC_STEP_NORM2], s_red);
    block_atomic_sum(xn, &scal2[SC_XNORM2], s_red);
}

inline void count_launch() { g_kernel_launches.fetch_add(1, std::memory_order_relaxed); }
inline int long_grid(int n) {
    const int blocks = (n + PB_WARPS - 1) / PB_WARPS;
    return blocks < 1 ? 1 : (blocks < 148 * 8 ? blocks : 148 * 8);
}

}  // namespace

void launch_phong_build_long(cudaStream_t s, const DevView& v, const PhongSolveView& q, int lm_lo, int lm_hi, LmDiag dg,
                             const PhongSystem& o, bool schur) {
    if (lm_hi <= lm_lo) return;
    if (schur)
        phong_build_long_kernel<true><<<long_grid(lm_hi - lm_lo), PB_WARPS * 32, 0, s>>>(v, q, lm_lo, lm_hi, dg, o);
    else
        phong_build_long_kernel<false><<<long_grid(lm_hi - lm_lo), PB_WARPS * 32, 0, s>>>(v, q, lm_lo, lm_hi, dg, o);
    count_launch();
    CSLAM_CUDA(cudaGetLastError());
}

void launch_phong_backsub_long(cudaStream_t s, const DevView& v, const PhongSolveView& q, int lm_lo, int lm_hi, LmDiag dg,
                               const double* yp, const double* yg, const double* gv, double* yv, double* scal2) {
    if (lm_hi <= lm_lo) return;
    phong_backsub_long_kernel<<<long_grid(lm_hi - lm_lo), PB_WARPS * 32, 0, s>>>(v, q, lm_lo, lm_hi, dg, yp, yg, gv, yv, scal2);
    count_launch();
    CSLAM_CUDA(cudaGetLastError());
}

void launch_phong_dogleg_products_long(cudaStream_t s, const DevView& v, const PhongSolveView& q, int lm_lo, int lm_hi, LmDiag dg,
                                       const double* gp, const double* diag_p, const double* yp, const double* gg,
                                       const double* diag_g, const double* yg, const double* gv, const double* yv, double* diag_v,
                                       double* sc_v, double* sums) {
    if (lm_hi <= lm_lo) return;
    phong_dogleg_products_long_kernel<<<long_grid(lm_hi - lm_lo), PB_WARPS * 32, 0, s>>>(v, q, lm_lo, lm_hi, dg, gp, diag_p, yp, gg,
                                                                                          diag_g, yg, gv, yv, diag_v, sc_v, sums);
    count_launch();
    CSLAM_CUDA(cudaGetLastError());
}

void launch_phong_candidate_long(cudaStream_t s, const DevView& v, const PhongSolveView& q, int lm_lo, int lm_hi, double alpha,
                                 const double* yv, const double* poses_cand, const double* gx_cand, double* points_cand,
                                 double* normals_cand, double* scal2) {
    if (lm_hi <= lm_lo) return;
    phong_candidate_long_kernel<<<long_grid(lm_hi - lm_lo), PB_WARPS * 32, 0, s>>>(v, q, lm_lo, lm_hi, alpha, yv, poses_cand, gx_cand,
                                                                                    points_cand, normals_cand, scal2);
    count_launch();
    CSLAM_CUDA(cudaGetLastError());
}

}  // namespace cslam
