// K2p for GROUPED vertices — config 3 (dataset_ba_phong's joint solve, tests/dataset_ba_phong.cpp:100-252).
//
// A group is G vertices seen by the same L cameras (engine.cu groups landmarks with identical camera
// lists; observations of a group are stored slot-major).  Eliminating the 6x6 vertex blocks (position +
// normal) of a group is, for the camera part of the reduced system, a symmetric rank-6G update of ONE
// 6L x 6L tile:
//        S_tile -= Z Z^T ,   Z_a = W_a A^T (6x6 per observation),  V = C C^T,  A = C^-1,  W_a = A_c^T A_v
// so instead of ~2000 RED.ADD.F64 per vertex (the pair loop of phong_build_kernel) the tile stays in
// registers as FP64 tensor-core accumulators (mma.sync m8n8k4, SASS DMMA.8x8x4) for a whole work item
// (<= 128 vertices) and is flushed once.
//
// One CTA (4 warps) per work item, rounds of 8 vertices:
//   produce  every half-warp takes one vertex, lane = camera slot: the three blocks of the observation in
//            closed form (phong_common.cuh), segmented-shuffle sums for V, the vertex gradient and the
//            couplings to the shared blocks (materials, textures, light), 6x6 Cholesky inverse, then per
//            lane Z (into a shared tile, transposed: [k][row]), the slot's U = A_c^T A_c, gradient and
//            reduced right-hand side (lane-private accumulators in shared memory, flushed once per item)
//            and the border rows  S_cg += A_c[I]^T a_g - Z (G A^T)^T  (RED: their columns depend on the
//            vertex's material);
//   consume  the four warps apply the round's 48 columns of Z to the tile: warp w owns MMA row tiles
//            {w, NT-1-w (, 4+w)}; the A and B fragments are both single loads from the transposed tile.
// Arithmetic is that of phong_build_kernel (same closed forms, same elimination, V^-1 = A^T A).
#include "kernels.cuh"
#include "phong_common.cuh"

namespace cslam {

namespace {

constexpr int PG_THREADS = 128;
constexpr int PG_NV = 8;          // vertices per round: two per warp
constexpr int PG_K = 6 * PG_NV;   // columns of Z per round
constexpr int PG_SLOT_ACC = 33;   // per slot: U (21) | gradient (6) | reduced rhs (6)

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

// V (symmetric positive definite, full storage) = C C^T; returns A = C^-1 (lower, zeros above).
__device__ __forceinline__ bool chol6_inverse_lower(const double* V, double* A) {
    double L[36];
#pragma unroll
    for (int i = 0; i < 36; ++i) L[i] = V[i];
    bool ok = true;
#pragma unroll
    for (int j = 0; j < 6; ++j) {
        double d = L[6 * j + j];
#pragma unroll
        for (int k = 0; k < j; ++k) d -= L[6 * j + k] * L[6 * j + k];
        if (!(d > 0.0) || !(d < 1.7976931348623157e308)) ok = false;
        d = sqrt(d);
        L[6 * j + j] = d;
        const double id = 1.0 / d;
#pragma unroll
        for (int i = j + 1; i < 6; ++i) {
            double s = L[6 * i + j];
#pragma unroll
            for (int k = 0; k < j; ++k) s -= L[6 * i + k] * L[6 * j + k];
            L[6 * i + j] = s * id;
        }
    }
#pragma unroll
    for (int c = 0; c < 6; ++c) {
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            if (i < c) {
                A[6 * i + c] = 0.0;
            } else if (i == c) {
                A[6 * i + c] = 1.0 / L[6 * i + i];
            } else {
                double s = 0.0;
#pragma unroll
                for (int k = c; k < i; ++k) s -= L[6 * i + k] * A[6 * k + c];
                A[6 * i + c] = s / L[6 * i + i];
            }
        }
    }
    return ok;
}

template <int NT>  // MMA row tiles of the 6L x 6L tile: 8 (L <= 10) or 12 (L <= 16)
__global__ void __launch_bounds__(PG_THREADS)
    phong_grouped_kernel(DevView v, PhongSolveView q, GroupView gv, int item_lo, int item_hi, LmDiag dg, PhongSystem o) {
    constexpr int LD = 8 * NT + 4;               // row stride of the transposed Z tile: conflict-free fragment loads
    constexpr int LMAX = NT == 8 ? 10 : 16;
    constexpr int NACC = NT == 8 ? 9 : 21;       // MMA tiles per warp
    extern __shared__ __align__(16) double smem_pg[];
    double* Zt = smem_pg;                         // [PG_K][LD]   Zt[k][row] = Z[row][k]
    double* sAcc = Zt + PG_K * LD;                // [PG_NV][LMAX][PG_SLOT_ACC]
    double* sW = sAcc + PG_NV * LMAX * PG_SLOT_ACC;   // [36][PG_THREADS]  W = A_c^T A_v of the lane's observation
    double* sTile = sW + 36 * PG_THREADS;         // [2 NACC][PG_THREADS]  the MMA accumulators, parked during `produce`
    __shared__ double s_pose_unused[1];
    __shared__ double sG[PG_NV][56];              // a vertex's shared-block contributions, staged for the REDs
    __shared__ int s_free[16];
    __shared__ int s_blk[16 * 17 / 2];
    __shared__ double s_red[32];
    (void)s_pose_unused;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int lv = lane & 15, sub = lane >> 4, qv = 2 * warp + sub;   // slot, half, vertex of the round
    const int nf6 = 6 * v.n_free;
    double* myG = sG[qv];
    double* myAcc = sAcc + (qv * LMAX + (lv < LMAX ? lv : 0)) * PG_SLOT_ACC;
    double cost = 0.0;

    // consumer role: rows r0 = warp, r1 = NT-1-warp, (r2 = 4+warp for NT = 12)
    const int r0 = warp, r1 = NT - 1 - warp, r2 = 4 + warp;
    const int c0 = NT - r0, c1 = NT - r1, c2 = NT == 12 ? NT - r2 : 0;
    const int fg = lane >> 2, ft = lane & 3;      // fragment coordinates

    for (int w = item_lo + blockIdx.x; w < item_hi; w += gridDim.x) {
        const int g = gv.item_group[w];
        const int L = gv.g_L[g];
        const int j0 = gv.item_j0[w], nj = gv.item_n[w];
        const int lm0 = gv.g_lm0[g] + j0;
        const int* __restrict__ cams = gv.g_cams + gv.g_off[g];
        const int* __restrict__ blk = gv.g_blk + gv.g_blk_off[g];
        const int P = L * (L + 1) / 2;
        __syncthreads();   // the previous item's flush has read Zt / sAcc / s_blk
        if (tid < L) s_free[tid] = v.cam_free[cams[tid]];
        for (int k = tid; k < P; k += PG_THREADS) s_blk[k] = blk[k];
        for (int k = tid; k < PG_K * LD; k += PG_THREADS) Zt[k] = 0.0;
        for (int k = tid; k < PG_NV * LMAX * PG_SLOT_ACC; k += PG_THREADS) sAcc[k] = 0.0;
#pragma unroll
        for (int u = 0; u < 2 * NACC; ++u) sTile[u * PG_THREADS + tid] = 0.0;
        __syncthreads();

        for (int rb = 0; rb < nj; rb += PG_NV) {
            // ================= produce =================
            const int jl = rb + qv;
            const bool vok = jl < nj;
            const int j = lm0 + (vok ? jl : nj - 1);
            const long long e0 = v.lm_base[j], es = v.lm_stride[j];
            const bool act = vok && lv < L;
            VertexCtx c;
            load_vertex(v, q, j, c);
            PhObs ob;
            ob.f = -1;
            double gvv[6], r7[7];
            double V[36];
            {
                double V21[21];
                if (act) {
                    eval_phong_obs(v, q, e0 + lv * es, c, ob);
                    r7[0] = ob.rs[0], r7[1] = ob.rs[1], r7[2] = ob.rs[2], r7[3] = ob.rI;
                    r7[4] = ob.rN[0], r7[5] = ob.rN[1], r7[6] = ob.rN[2];
#pragma unroll
                    for (int k = 0; k < 7; ++k) cost += 0.5 * r7[k] * r7[k];
                    vertex_normal_eq(ob, r7, V21, gvv);
                    if (ob.f >= 0) {
                        // everything that needs the pose columns, now: W = A_c^T A_v (parked in shared memory),
                        // U = A_c^T A_c and the gradient (slot accumulators).  The pose Jacobians die here.
                        int u = 0;
#pragma unroll
                        for (int a = 0; a < 6; ++a) {
#pragma unroll
                            for (int b = 0; b < 3; ++b) {
                                sW[(6 * a + b) * PG_THREADS + tid] = ob.Jcs[a] * ob.S[b] + ob.Jcs[6 + a] * ob.S[3 + b] +
                                                                     ob.Jcs[12 + a] * ob.S[6 + b] + ob.JIc[a] * ob.ip[b];
                                sW[(6 * a + 3 + b) * PG_THREADS + tid] = ob.JIc[a] * ob.in[b] + ob.JNc[a] * ob.N[b] +
                                                                         ob.JNc[6 + a] * ob.N[3 + b] + ob.JNc[12 + a] * ob.N[6 + b];
                            }
#pragma unroll
                            for (int b = a; b < 6; ++b, ++u)
                                myAcc[u] += ob.Jcs[a] * ob.Jcs[b] + ob.Jcs[6 + a] * ob.Jcs[6 + b] + ob.Jcs[12 + a] * ob.Jcs[12 + b] +
                                            ob.JIc[a] * ob.JIc[b] + ob.JNc[a] * ob.JNc[b] + ob.JNc[6 + a] * ob.JNc[6 + b] +
                                            ob.JNc[12 + a] * ob.JNc[12 + b];
                            const double ga = ob.Jcs[a] * r7[0] + ob.Jcs[6 + a] * r7[1] + ob.Jcs[12 + a] * r7[2] + ob.JIc[a] * r7[3] +
                                              ob.JNc[a] * r7[4] + ob.JNc[6 + a] * r7[5] + ob.JNc[12 + a] * r7[6];
                            myAcc[21 + a] += ga;
                            myAcc[27 + a] += ga;      // reduced rhs: - Z t follows once the vertex block is factored
                        }
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < 21; ++k) V21[k] = 0.0;
#pragma unroll
                    for (int k = 0; k < 6; ++k) gvv[k] = 0.0;
#pragma unroll
                    for (int k = 0; k < 7; ++k) ob.ag[k] = 0.0;
#pragma unroll
                    for (int k = 0; k < 3; ++k) ob.ip[k] = ob.in[k] = 0.0;
                    ob.rI = 0.0;
                }
#pragma unroll
                for (int k = 0; k < 21; ++k) V21[k] = seg_sum<16>(V21[k]);
                unpack_sym6(V21, V);
            }
#pragma unroll
            for (int k = 0; k < 6; ++k) gvv[k] = seg_sum<16>(gvv[k]);
            add_lm_diag(V, dg);
            double A[36];
            const bool pd = chol6_inverse_lower(V, A) && vok;
            if (lv == 0 && vok) {
#pragma unroll
                for (int a = 0; a < 6; ++a) o.gv[6ll * j + a] = gvv[a];
                if (!pd) red_add(&o.scal[SC_INVALID], 1.0);
            }
            double tq[6];   // t = A g_v
#pragma unroll
            for (int a = 0; a < 6; ++a) {
                double s = 0.0;
#pragma unroll
                for (int b = 0; b <= a; ++b) s += A[6 * a + b] * gvv[b];
                tq[a] = s;
            }
            // GA = G A^T, G[k][p] = sum_obs ag[k] * A_v[intensity row][p]
            double GA[42];
#pragma unroll
            for (int k = 0; k < 7; ++k) {
                double Gk[6];
#pragma unroll
                for (int a = 0; a < 3; ++a) {
                    Gk[a] = seg_sum<16>(ob.ag[k] * ob.ip[a]);
                    Gk[3 + a] = seg_sum<16>(ob.ag[k] * ob.in[a]);
                }
#pragma unroll
                for (int a = 0; a < 6; ++a) {
                    double s = 0.0;
#pragma unroll
                    for (int b = 0; b <= a; ++b) s += Gk[b] * A[6 * a + b];
                    GA[6 * k + a] = s;
                }
            }
            // shared blocks: H_gg = sum a_g a_g^T - GA GA^T, b_g = g_g - GA t, gradient, diagonal
            {
                int idx = 0;
#pragma unroll
                for (int k = 0; k < 7; ++k)
#pragma unroll
                    for (int k2 = k; k2 < 7; ++k2) {
                        const double hs = seg_sum<16>(ob.ag[k] * ob.ag[k2]);
                        double h = hs;
#pragma unroll
                        for (int a = 0; a < 6; ++a) h -= GA[6 * k + a] * GA[6 * k2 + a];
                        if (lv == 0) {
                            myG[idx] = h;
                            if (k2 == k) myG[42 + k] = hs;
                        }
                        ++idx;
                    }
#pragma unroll
                for (int k = 0; k < 7; ++k) {
                    const double gk = seg_sum<16>(ob.ag[k] * ob.rI);
                    double bb = gk;
#pragma unroll
                    for (int a = 0; a < 6; ++a) bb -= GA[6 * k + a] * tq[a];
                    if (lv == 0) {
                        myG[28 + k] = bb;
                        myG[35 + k] = gk;
                    }
                }
            }
            __syncwarp();
            for (int en = lv; en < 28 && pd; en += 16) {
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
            // ---- camera part: Z into the shared tile, slot accumulators, border rows ----
            if (lv < L) {
                double* zc = Zt + (6 * qv) * LD + 6 * lv;   // Zt[6 qv + c][6 lv + a]
                if (act && pd && ob.f >= 0) {
#pragma unroll
                    for (int a = 0; a < 6; ++a) {
                        double Wr[6], Zr[6];
#pragma unroll
                        for (int b = 0; b < 6; ++b) Wr[b] = sW[(6 * a + b) * PG_THREADS + tid];
                        double zt = 0.0;
#pragma unroll
                        for (int cc = 0; cc < 6; ++cc) {
                            double sacc = 0.0;
#pragma unroll
                            for (int b = 0; b <= cc; ++b) sacc += Wr[b] * A[6 * cc + b];
                            Zr[cc] = sacc;
                            zc[cc * LD + a] = sacc;
                            zt += sacc * tq[cc];
                        }
                        myAcc[27 + a] -= zt;
#pragma unroll
                        for (int k = 0; k < 7; ++k) {
                            double sb = ob.JIc[a] * ob.ag[k];
#pragma unroll
                            for (int b = 0; b < 6; ++b) sb -= Zr[b] * GA[6 * k + b];
                            red_add(&o.Scg[(long long)c.gi[k] * nf6 + 6 * ob.f + a], sb);
                        }
                    }
                } else {
#pragma unroll
                    for (int cc = 0; cc < 6; ++cc)
#pragma unroll
                        for (int a = 0; a < 6; ++a) zc[cc * LD + a] = 0.0;
                }
            }
            __syncthreads();
            // ================= consume: tile += Z Z^T over the round's 48 columns =================
            double acc[NACC][2];
#pragma unroll
            for (int u = 0; u < NACC; ++u) {
                acc[u][0] = sTile[(2 * u) * PG_THREADS + tid];
                acc[u][1] = sTile[(2 * u + 1) * PG_THREADS + tid];
            }
#pragma unroll 2
            for (int ks = 0; ks < PG_K / 4; ++ks) {
                const double* zr = Zt + (4 * ks + ft) * LD + fg;
                const double a0 = zr[8 * r0], a1 = zr[8 * r1];
                const double a2 = NT == 12 ? zr[8 * r2] : 0.0;
#pragma unroll
                for (int u = 0; u < NACC; ++u) {
                    int J;
                    double a;
                    if (u < c0) {
                        J = r0 + u;
                        a = a0;
                    } else if (u < c0 + c1) {
                        J = r1 + (u - c0);
                        a = a1;
                    } else {
                        J = r2 + (u - c0 - c1);
                        a = a2;
                        if (NT != 12 || u >= c0 + c1 + c2) continue;
                    }
                    dmma884(acc[u][0], acc[u][1], a, zr[8 * J]);
                }
            }
#pragma unroll
            for (int u = 0; u < NACC; ++u) {
                sTile[(2 * u) * PG_THREADS + tid] = acc[u][0];
                sTile[(2 * u + 1) * PG_THREADS + tid] = acc[u][1];
            }
            __syncthreads();
        }

        // ================= flush =================
#pragma unroll
        for (int u = 0; u < NACC; ++u) {
            int I, J;
            if (u < c0) {
                I = r0;
                J = r0 + u;
            } else if (u < c0 + c1) {
                I = r1;
                J = r1 + (u - c0);
            } else {
                I = r2;
                J = r2 + (u - c0 - c1);
                if (NT != 12 || u >= c0 + c1 + c2) continue;
            }
            const int R = 8 * I + fg;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int Cc = 8 * J + 2 * ft + h;
                if (R > Cc || Cc >= 6 * L) continue;      // upper storage; rows / columns beyond the tile are padding
                const int sa = R / 6, sb = Cc / 6;
                const int e = s_blk[sa * L - sa * (sa - 1) / 2 + (sb - sa)];
                if (e >= 0) red_add(&o.S[36ll * e + 6 * (R - 6 * sa) + (Cc - 6 * sb)], -sTile[(2 * u + h) * PG_THREADS + tid]);
            }
        }
        for (int idx = tid; idx < L * PG_SLOT_ACC; idx += PG_THREADS) {
            const int sl = idx / PG_SLOT_ACC, k = idx - sl * PG_SLOT_ACC;
            const int f = s_free[sl];
            if (f < 0) continue;
            double s = 0.0;
#pragma unroll
            for (int ow = 0; ow < PG_NV; ++ow) s += sAcc[(ow * LMAX + sl) * PG_SLOT_ACC + k];
            if (k < 21) {
                int a = 0, rem = k;
                while (rem >= 6 - a) {
                    rem -= 6 - a;
                    ++a;
                }
                red_add(&o.Bdiag[36ll * f + 6 * a + (a + rem)], s);
            } else if (k < 27) {
                red_add(&o.gp[6ll * f + (k - 21)], s);
            } else {
                red_add(&o.bp[6ll * f + (k - 27)], s);
            }
        }
    }
    block_atomic_sum(cost, &o.scal[SC_COST], s_red);
}

template <int NT>
void launch_pg(cudaStream_t s, const DevView& v, const PhongSolveView& q, const GroupView& g, int lo, int hi, LmDiag dg,
               const PhongSystem& o) {
    if (hi <= lo) return;
    constexpr int LD = 8 * NT + 4, LMAX = NT == 8 ? 10 : 16;
    constexpr int NACC = NT == 8 ? 9 : 21;
    const size_t smem = sizeof(double) * (size_t(PG_K) * LD + size_t(PG_NV) * LMAX * PG_SLOT_ACC + 36 * size_t(PG_THREADS) +
                                          2 * size_t(NACC) * PG_THREADS);
    static PerDevice attr;
    const int dev = PerDevice::current();
    if (attr.first_use(dev)) {
        CSLAM_CUDA(cudaFuncSetAttribute(phong_grouped_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
        int per_sm = 0;
        CSLAM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, phong_grouped_kernel<NT>, PG_THREADS, smem));
        attr.value[dev] = per_sm < 1 ? 1 : per_sm;
        attr.mark(dev);
    }
    const int n = hi - lo, cap = 148 * attr.value[dev];
    phong_grouped_kernel<NT><<<n < cap ? n : cap, PG_THREADS, smem, s>>>(v, q, g, lo, hi, dg, o);
    g_kernel_launches.fetch_add(1, std::memory_order_relaxed);
    CSLAM_CUDA(cudaGetLastError());
}

}  // namespace

// Work items [0, n_items_small) have L <= 10, the rest L <= 16 (engine.cu's grouping).
void launch_phong_build_grouped(cudaStream_t s, const DevView& v, const PhongSolveView& q, const GroupView& g, int n_items_small,
                                LmDiag dg, const PhongSystem& o) {
    launch_pg<8>(s, v, q, g, 0, n_items_small, dg, o);
    launch_pg<12>(s, v, q, g, n_items_small, g.n_items, dg, o);
}

}  // namespace cslam
