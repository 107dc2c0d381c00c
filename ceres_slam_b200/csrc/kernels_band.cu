// K3b — direct solve of a block-BANDED reduced camera system (the SPARSE_SCHUR-equivalent exact
// solve for sequential tracks: visual odometry / sliding windows / config 5, where camera k only
// shares landmarks with cameras k-w..k+w).  FP64 throughout, no atomics: deterministic.
//
// A banded Cholesky is a chain of n dependent 6x6 block steps; run naively it occupies one SM.
// The chain is cut into P leaves (substructuring with P-1 separators of W block rows each):
//
//   level 1  band_leaf_kernel<W, spike>   (P CTAs)  each leaf factors its interior A = L L^T
//            right-looking with a (W+1)x(W+1) block window in shared memory.  The coupling to the
//            separator BEFORE the leaf (the "left spike", 6W columns) and the right-hand side ride
//            along as a border G, so X = L^-1 [r | B_left] comes out of the same sweep; the
//            separator AFTER the leaf is the tail of the window and receives its Schur complement
//            in place.  X^T X (contribution to the separator before) accumulates in registers.
//            The 6x6 factorisation of the NEXT pivot runs on warp 0 while the other warps apply
//            the trailing update of the current one (look-ahead), so the sqrt/div chain is hidden.
//   assemble band_assemble_kernel          the separator system is itself block-banded
//            (half-bandwidth 2W-1); it is written in dense band storage.
//   level 2  band_leaf_kernel<2W-1, false> (1 CTA)  the same sweep on the separator system.
//   backsub  band_backsub_kernel<W>        (1 CTA for level 2, then P CTAs for level 1)
//            y_k = Lkk^-T (z_k - X_left,k y_left - sum_d L_{k+d,k}^T y_{k+d}).
//
// Work is ~n (W+1)^2 6^3 flops — negligible; everything is bound by the latency of the dependent
// block steps, which is why the chain is cut and why the step is kept to two CTA barriers.
#include <algorithm>
#include <cmath>

#include "kernels.cuh"

namespace cslam {

namespace {

constexpr int BL_THREADS = 256;
constexpr int BL_WORKERS = BL_THREADS - 32;  // warps 1..7 do the bulk update, warp 0 the look-ahead

// 6x6 Cholesky of a symmetric block (lower triangle read from `A`, row-major) entirely in
// registers: L (lower, zeros above) and Li = L^-1.  Returns false when a pivot is not positive.
__device__ __forceinline__ bool potrf6_inv_reg(const double* A, double* Lout, double* Liout) {
    double L[6][6];
#pragma unroll
    for (int i = 0; i < 6; ++i)
#pragma unroll
        for (int j = 0; j < 6; ++j) L[i][j] = (j <= i) ? A[6 * i + j] : 0.0;
    bool ok = true;
#pragma unroll
    for (int j = 0; j < 6; ++j) {
        double d = L[j][j];
#pragma unroll
        for (int k = 0; k < j; ++k) d -= L[j][k] * L[j][k];
        if (!(d > 0.0) || !(d < 1.7976931348623157e308)) ok = false;
        const double sd = sqrt(d);
        const double id = 1.0 / sd;
        L[j][j] = sd;
#pragma unroll
        for (int i = j + 1; i < 6; ++i) {
            double s = L[i][j];
#pragma unroll
            for (int k = 0; k < j; ++k) s -= L[i][k] * L[j][k];
            L[i][j] = s * id;
        }
    }
    double Li[6][6];
#pragma unroll
    for (int c = 0; c < 6; ++c) {
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            if (i < c) {
                Li[i][c] = 0.0;
            } else {
                double s = (i == c) ? 1.0 : 0.0;
#pragma unroll
                for (int k = c; k < i; ++k) s -= L[i][k] * Li[k][c];
                Li[i][c] = s / L[i][i];
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 6; ++i)
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            Lout[6 * i + j] = L[i][j];
            Liout[6 * i + j] = Li[i][j];
        }
    return ok;
}

// A_{i,j}[r][c] (i >= j) of the symmetric band matrix; zero when the block is absent.
// Storage is upper: block (j, i) as a 6x6 row-major tile, A_ij[r][c] = blk(j,i)[c][r]; the
// diagonal tile is full symmetric.  `band_idx == nullptr` means dense band storage [n][W+1][36].
template <int W>
__device__ __forceinline__ double band_entry(const BandView& B, int i, int j, int r, int c) {
    if (B.band_idx) {
        const int d = i - j;
        if (d > B.w) return 0.0;  // storage width of the index table is the true half-bandwidth
        const int e = B.band_idx[(long long)j * (B.w + 1) + d];
        return e < 0 ? 0.0 : B.S[36ll * e + 6 * c + r];
    }
    return B.S[((long long)j * (W + 1) + (i - j)) * 36 + 6 * c + r];
}

template <int W, bool kSpike>
__global__ void __launch_bounds__(BL_THREADS) band_leaf_kernel(BandView B) {
    constexpr int W1 = W + 1, GC = kSpike ? 1 + 6 * W : 1, b = 6 * W;
    constexpr int NPAIR = W * (W + 1) / 2;
    constexpr int ACC = kSpike ? (b * GC + BL_WORKERS - 1) / BL_WORKERS : 1;
    constexpr int NROW = W1 * 36 + 6 * GC;  // values of one window row (blocks + border)
    constexpr int NXT = (NROW + BL_WORKERS - 1) / BL_WORKERS;
    extern __shared__ __align__(16) double smem_bl[];
    double* Awin = smem_bl;              // [W1][W1][36]
    double* Gwin = Awin + W1 * W1 * 36;  // [W1][6][GC]
    double* Xk = Gwin + W1 * 6 * GC;     // [6][GC]
    double* Lcol = Xk + 6 * GC;          // [W1][36]  d >= 1: L_{k+d,k}
    double* Ldg = Lcol + W1 * 36;        // [2][36]   Lkk of the current / next pivot
    double* Liv = Ldg + 72;              // [2][36]   their inverses
    __shared__ int s_ok;
    __shared__ unsigned char pair_i[NPAIR], pair_j[NPAIR];

    const int tid = threadIdx.x;
    const int p = blockIdx.x;
    const int s = p * B.m;
    const int e = min(B.n, s + B.m);
    const bool has_left = kSpike && p > 0, has_right = p < B.P - 1;
    const int ie = has_right ? e - W : e;
    if (tid == 0) {
        s_ok = 1;
        int t = 0;
        for (int di = 1; di <= W; ++di)
            for (int dj = 1; dj <= di; ++dj) {
                pair_i[t] = (unsigned char)di;
                pair_j[t] = (unsigned char)dj;
                ++t;
            }
    }
    // value `idx` of window row i: blocks A_{i,j} for j = i-W..i (zero before the leaf start),
    // then the border G_i = [r_i | B_left row i]
    auto row_value = [&](int i, int idx) -> double {
        if (idx < W1 * 36) {
            const int j = i - W + idx / 36, r = (idx % 36) / 6, c = idx % 6;
            return j < s ? 0.0 : band_entry<W>(B, i, j, r, c);
        }
        const int g = idx - W1 * 36, r = g / GC, col = g % GC;
        if (col == 0) return B.rhs[6ll * i + r];
        if (!has_left) return 0.0;
        const int cc = s - W + (col - 1) / 6;  // block row of the separator before the leaf
        return (i - cc <= W) ? band_entry<W>(B, i, cc, r, (col - 1) % 6) : 0.0;
    };
    auto row_store = [&](int i, int idx, double v) {
        const int si = i % W1;
        if (idx < W1 * 36) {
            const int j = i - W + idx / 36;
            Awin[(si * W1 + ((j % W1) + W1) % W1) * 36 + idx % 36] = v;
        } else {
            Gwin[si * 6 * GC + (idx - W1 * 36)] = v;
        }
    };
    for (int i = s; i <= min(s + W, e - 1); ++i)
        for (int idx = tid; idx < NROW; idx += BL_THREADS) row_store(i, idx, row_value(i, idx));
    double acc[ACC];
#pragma unroll
    for (int u = 0; u < ACC; ++u) acc[u] = 0.0;
    __syncthreads();
    if (tid == 0 && ie > s) {
        if (!potrf6_inv_reg(Awin + ((s % W1) * W1 + (s % W1)) * 36, Ldg, Liv)) s_ok = 0;
    }
    __syncthreads();

    for (int k = s; k < ie; ++k) {
        if (!s_ok) break;
        const int sk = k % W1;
        const int cur = (k - s) & 1;
        const int nb = min(k + W, e - 1) - k;  // sub-diagonal blocks of column k inside the leaf
        const double* Linv = Liv + 36 * cur;
        // ---- phase 1: L_ik = A_ik Lkk^-T,  X_k = Lkk^-1 G_k -----------------------------------
        for (int idx = tid; idx < nb * 36 + 6 * GC; idx += BL_THREADS) {
            if (idx < nb * 36) {
                const int d = idx / 36 + 1, r = (idx % 36) / 6, c = idx % 6;
                const double* Aik = Awin + (((k + d) % W1) * W1 + sk) * 36 + 6 * r;
                double v = 0.0;
#pragma unroll
                for (int q = 0; q < 6; ++q)
                    if (q <= c) v += Aik[q] * Linv[6 * c + q];
                Lcol[d * 36 + 6 * r + c] = v;
            } else {
                const int i2 = idx - nb * 36, r = i2 / GC, col = i2 % GC;
                double v = 0.0;
#pragma unroll
                for (int q = 0; q < 6; ++q)
                    if (q <= r) v += Linv[6 * r + q] * Gwin[(sk * 6 + q) * GC + col];
                Xk[r * GC + col] = v;
            }
        }
        __syncthreads();
        // ---- phase 2 -----------------------------------------------------------------------------
        if (tid < 32) {
            // look-ahead: bring the next pivot up to date and factor it while the other warps
            // apply the bulk of the trailing update
            if (nb >= 1) {
                double* A11 = Awin + (((k + 1) % W1) * W1 + ((k + 1) % W1)) * 36;
                const double* L1 = Lcol + 36;
                for (int idx = tid; idx < 36; idx += 32) {
                    const int r = idx / 6, c = idx % 6;
                    double v = 0.0;
#pragma unroll
                    for (int q = 0; q < 6; ++q) v += L1[6 * r + q] * L1[6 * c + q];
                    A11[idx] -= v;
                }
                __syncwarp();
                if (tid == 0 && k + 1 < ie) {
                    if (!potrf6_inv_reg(A11, Ldg + 36 * (cur ^ 1), Liv + 36 * (cur ^ 1))) s_ok = 0;
                }
            }
            // factor column (inverse of the pivot block first) and X row for the back-substitution
            double* Lg = B.Lbuf + ((long long)k * W1) * 36;
            for (int idx = tid; idx < 36; idx += 32) Lg[idx] = Linv[idx];
            for (int idx = tid; idx < nb * 36; idx += 32) Lg[36 + idx] = Lcol[36 + idx];
            double* Xg = B.Xbuf + (long long)k * 6 * GC;
            for (int idx = tid; idx < 6 * GC; idx += 32) Xg[idx] = Xk[idx];
        } else {
            const int t = tid - 32;
            // the row entering the window: issue its global loads first, park them in registers
            const int inew = k + W + 1;
            double nxt[NXT];
            if (inew < e) {
#pragma unroll
                for (int u = 0; u < NXT; ++u) {
                    const int idx = t + BL_WORKERS * u;
                    nxt[u] = idx < NROW ? row_value(inew, idx) : 0.0;
                }
            }
            // trailing window: A_ij -= L_ik L_jk^T (pair (1,1) is warp 0's)
            const int npairs = nb * (nb + 1) / 2;
            for (int idx = 36 + t; idx < npairs * 36; idx += BL_WORKERS) {
                const int pr = idx / 36, rc = idx % 36, r = rc / 6, c = rc % 6;
                const int di = pair_i[pr], dj = pair_j[pr];
                const double* Li = Lcol + di * 36 + 6 * r;
                const double* Lj = Lcol + dj * 36 + 6 * c;
                double v = 0.0;
#pragma unroll
                for (int q = 0; q < 6; ++q) v += Li[q] * Lj[q];
                Awin[(((k + di) % W1) * W1 + ((k + dj) % W1)) * 36 + rc] -= v;
            }
            // border: G_i -= L_ik X_k
            for (int idx = t; idx < nb * 6 * GC; idx += BL_WORKERS) {
                const int d = idx / (6 * GC) + 1, rem = idx % (6 * GC), r = rem / GC, col = rem % GC;
                const double* Li = Lcol + d * 36 + 6 * r;
                double v = 0.0;
#pragma unroll
                for (int q = 0; q < 6; ++q) v += Li[q] * Xk[q * GC + col];
                Gwin[((k + d) % W1) * 6 * GC + rem] -= v;
            }
            // contribution to the separator before this leaf: X^T X over (left column, any column)
            if (kSpike && has_left) {
#pragma unroll
                for (int u = 0; u < ACC; ++u) {
                    const int flat = t + BL_WORKERS * u;
                    if (flat < b * GC) {
                        const int c1 = 1 + flat / GC, c2 = flat % GC;
                        double v = 0.0;
#pragma unroll
                        for (int q = 0; q < 6; ++q) v += Xk[q * GC + c1] * Xk[q * GC + c2];
                        acc[u] += v;
                    }
                }
            }
            if (inew < e) {
#pragma unroll
                for (int u = 0; u < NXT; ++u) {
                    const int idx = t + BL_WORKERS * u;
                    if (idx < NROW) row_store(inew, idx, nxt[u]);
                }
            }
        }
        __syncthreads();
    }
    if (!s_ok) {
        if (tid == 0) *B.fail = 1;
        return;
    }
    if (has_right) {
        // separator after this leaf: the window tail holds S[sep,sep] - B_right^T A^-1 B_right,
        // its border the coupling to the separator before (if any) and the right-hand side
        for (int idx = tid; idx < b * b; idx += BL_THREADS) {
            const int R = idx / b, Cc = idx % b;
            const int i = ie + R / 6, j = ie + Cc / 6;
            double v;
            if (j <= i)
                v = Awin[((i % W1) * W1 + (j % W1)) * 36 + 6 * (R % 6) + (Cc % 6)];
            else
                v = Awin[((j % W1) * W1 + (i % W1)) * 36 + 6 * (Cc % 6) + (R % 6)];
            B.Ta[(long long)p * b * b + idx] = v;
            B.Ca[(long long)p * b * b + idx] = has_left ? Gwin[((i % W1) * 6 + R % 6) * GC + (kSpike ? 1 + Cc : 0)] : 0.0;
        }
        for (int R = tid; R < b; R += BL_THREADS) B.fa[(long long)p * b + R] = Gwin[(((ie + R / 6) % W1) * 6 + R % 6) * GC];
    }
    if (kSpike && has_left && tid >= 32) {
        const int t = tid - 32;
#pragma unroll
        for (int u = 0; u < ACC; ++u) {
            const int flat = t + BL_WORKERS * u;
            if (flat < b * GC) {
                const int c1 = flat / GC, c2 = flat % GC;
                if (c2 == 0)
                    B.fb[(long long)(p - 1) * b + c1] = acc[u];
                else
                    B.Tb[(long long)(p - 1) * b * b + c1 * b + (c2 - 1)] = acc[u];
            }
        }
    }
}

// Separator system in dense band storage: n2 = (P-1) W block rows, half-bandwidth W2 = 2W-1.
//   block (a, a+d), a = q W + i:  inside separator q (i + d < W): (Ta - Tb)[q] tile (i, i+d);
//   into separator q+1 (j = i + d - W < W): T[sep q row i, sep q+1 row j] = Ca[q+1] tile (j, i)^T.
template <int W>
__global__ void band_assemble_kernel(BandView B, double* T2, double* rhs2) {
    constexpr int W2 = 2 * W - 1, b = 6 * W;
    const int n2 = (B.P - 1) * W;
    const long long total = (long long)n2 * (W2 + 1) * 36;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total + 6ll * n2;
         idx += (long long)gridDim.x * blockDim.x) {
        if (idx >= total) {
            const int row = int(idx - total);  // scalar row of the separator system
            const int q = row / b, rr = row % b;
            rhs2[row] = B.fa[(long long)q * b + rr] - B.fb[(long long)q * b + rr];
            continue;
        }
        const int rc = int(idx % 36), r = rc / 6, c = rc % 6;
        const long long blkid = idx / 36;
        const int d = int(blkid % (W2 + 1)), a = int(blkid / (W2 + 1));
        const int q = a / W, i = a % W;
        double v = 0.0;
        if (a + d < n2) {
            if (i + d < W) {
                const long long o = (long long)q * b * b + (long long)(6 * i + r) * b + 6 * (i + d) + c;
                v = B.Ta[o] - B.Tb[o];
            } else if (i + d - W < W) {
                const int j = i + d - W;
                v = B.Ca[(long long)(q + 1) * b * b + (long long)(6 * j + c) * b + 6 * i + r];
            }
        }
        T2[idx] = v;
    }
}

// Interior unknowns of every leaf.  `ysep` holds the separator unknowns of the level (block row
// q W + i of the separator system = row i of separator q); nullptr when the level has one leaf.
template <int W, bool kSpike>
__global__ void __launch_bounds__(128) band_backsub_kernel(BandView B, const double* ysep) {
    constexpr int W1 = W + 1, GC = kSpike ? 1 + 6 * W : 1, b = 6 * W;
    constexpr int DG = (W + 4) / 5;  // block columns per lane group (5 groups x 6 lanes)
    extern __shared__ __align__(16) double smem_bb[];
    const int tid = threadIdx.x;
    const int p = blockIdx.x;
    const int s = p * B.m;
    const int e = min(B.n, s + B.m);
    const bool has_left = kSpike && p > 0, has_right = p < B.P - 1;
    const int ie = has_right ? e - W : e;
    const int ni = ie - s;
    double* z = smem_bb;           // [6 (ni + W)] leaf rows: z_k - X_left,k y_left, then the solution
    double* yl = z + 6 * (B.m + W);  // [b] separator before
    __shared__ double s_part[5][6], s_t[6];
    if (*B.fail) return;
    if (has_left)
        for (int c = tid; c < b; c += 128) yl[c] = ysep[(long long)(p - 1) * b + c];
    if (has_right)
        for (int c = tid; c < b; c += 128) {
            const double v = ysep[(long long)p * b + c];
            z[6 * ni + c] = v;
            B.y[6ll * ie + c] = v;
        }
    __syncthreads();
    for (int idx = tid; idx < 6 * ni; idx += 128) {
        const double* X = B.Xbuf + (long long)(6 * s + idx) * GC;
        double v = X[0];
        if (kSpike && has_left)
            for (int c = 0; c < b; ++c) v -= X[1 + c] * yl[c];
        z[idx] = v;
    }
    __syncthreads();
    if (tid < 32) {
        const int lane = tid, c = lane % 6, g = lane / 6;  // lanes 30, 31 idle
        double Lr[DG][6], Li[6];
        auto fetch = [&](int k) {
            const double* L = B.Lbuf + ((long long)k * W1) * 36;
            const int nb = min(k + W, e - 1) - k;
            if (g < 5) {
#pragma unroll
                for (int u = 0; u < DG; ++u) {
                    const int d = 1 + g * DG + u;
#pragma unroll
                    for (int r = 0; r < 6; ++r) Lr[u][r] = (d <= nb) ? L[d * 36 + 6 * r + c] : 0.0;
                }
            }
            if (lane < 6) {
#pragma unroll
                for (int i = 0; i < 6; ++i) Li[i] = L[6 * i + c];  // column c of Lkk^-1
            }
        };
        if (ni > 0) fetch(ie - 1);
        for (int k = ie - 1; k >= s; --k) {
            // partial sums with this step's factor column (already in registers)
            double part = 0.0;
            if (g < 5) {
#pragma unroll
                for (int u = 0; u < DG; ++u) {
                    const int d = 1 + g * DG + u;
                    if (d <= W && k + d < e) {
                        const double* yv = z + 6 * (k + d - s);
#pragma unroll
                        for (int r = 0; r < 6; ++r) part += Lr[u][r] * yv[r];
                    }
                }
                s_part[g][c] = part;
            }
            double Lic[6];
#pragma unroll
            for (int i = 0; i < 6; ++i) Lic[i] = Li[i];
            if (k > s) fetch(k - 1);  // next step's loads are in flight during the dependent part
            __syncwarp();
            if (lane < 6) {
                double t = z[6 * (k - s) + lane];
#pragma unroll
                for (int gg = 0; gg < 5; ++gg) t -= s_part[gg][lane];
                s_t[lane] = t;
            }
            __syncwarp();
            if (lane < 6) {
                double yv = 0.0;
#pragma unroll
                for (int i = 0; i < 6; ++i)
                    if (i >= lane) yv += Lic[i] * s_t[i];  // (Lkk^-T t)_c = sum_{i >= c} Linv[i][c] t_i
                z[6 * (k - s) + lane] = yv;
            }
            __syncwarp();
        }
    }
    __syncthreads();
    for (int idx = tid; idx < 6 * ni; idx += 128) B.y[6ll * s + idx] = z[idx];
}

__global__ void band_status_kernel(const int* fail, double* ps) {
    ps[PS_ITERS] = 1.0;
    ps[PS_FAIL] = *fail ? 2.0 : 0.0;
}

template <int W, bool kSpike>
size_t leaf_smem() {
    constexpr int W1 = W + 1, GC = kSpike ? 1 + 6 * W : 1;
    return sizeof(double) * (size_t(W1) * W1 * 36 + size_t(W1) * 6 * GC + 6 * GC + size_t(W1) * 36 + 144);
}

template <int W, bool kSpike>
void run_leaf(cudaStream_t s, const BandView& V) {
    static bool attr = false;
    if (!attr) {
        CSLAM_CUDA(cudaFuncSetAttribute(band_leaf_kernel<W, kSpike>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        int(leaf_smem<W, kSpike>())));
        attr = true;
    }
    band_leaf_kernel<W, kSpike><<<V.P, BL_THREADS, leaf_smem<W, kSpike>(), s>>>(V);
    CSLAM_CUDA(cudaGetLastError());
}

template <int W, bool kSpike>
void run_backsub(cudaStream_t s, const BandView& V, const double* ysep) {
    static bool attr = false;
    if (!attr) {
        CSLAM_CUDA(cudaFuncSetAttribute(band_backsub_kernel<W, kSpike>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        attr = true;
    }
    const size_t smem = sizeof(double) * (6 * size_t(V.m + W) + 6 * size_t(W));
    band_backsub_kernel<W, kSpike><<<V.P, 128, smem, s>>>(V, ysep);
    CSLAM_CUDA(cudaGetLastError());
}

// One solve with storage width W (>= true half-bandwidth V.w).
template <int W>
int band_solve_w(cudaStream_t s, const BandView& V, const BandScratch& K) {
    constexpr int W2 = 2 * W - 1;
    if (V.P == 1) {
        run_leaf<W, false>(s, V);
        run_backsub<W, false>(s, V, nullptr);
        return 2;
    }
    run_leaf<W, true>(s, V);
    const int n2 = (V.P - 1) * W;
    band_assemble_kernel<W><<<std::min(4 * 148, (n2 * (W2 + 1) * 36 + 255) / 256), 256, 0, s>>>(V, K.T2, K.rhs2);
    CSLAM_CUDA(cudaGetLastError());
    BandView V2 = V;
    V2.n = n2;
    V2.w = W2;
    V2.P = 1;
    V2.m = n2;
    V2.band_idx = nullptr;
    V2.S = K.T2;
    V2.rhs = K.rhs2;
    V2.Lbuf = K.L2;
    V2.Xbuf = K.X2;
    V2.y = K.y2;
    run_leaf<W2, false>(s, V2);
    run_backsub<W2, false>(s, V2, nullptr);
    run_backsub<W, true>(s, V, K.y2);
    return 5;
}

}  // namespace

int band_storage_width(int w) { return w <= 3 ? 3 : w <= 6 ? 6 : w <= 9 ? 9 : 12; }

void launch_band_solve(cudaStream_t s, const BandView& V, const BandScratch& K, double* ps) {
    CSLAM_CUDA(cudaMemsetAsync(V.fail, 0, sizeof(int), s));
    int launched = 0;
    switch (band_storage_width(V.w)) {
        case 3: launched = band_solve_w<3>(s, V, K); break;
        case 6: launched = band_solve_w<6>(s, V, K); break;
        case 9: launched = band_solve_w<9>(s, V, K); break;
        default: launched = band_solve_w<12>(s, V, K); break;
    }
    band_status_kernel<<<1, 1, 0, s>>>(V.fail, ps);
    CSLAM_CUDA(cudaGetLastError());
    g_kernel_launches.fetch_add(launched + 1, std::memory_order_relaxed);
}

}  // namespace cslam
