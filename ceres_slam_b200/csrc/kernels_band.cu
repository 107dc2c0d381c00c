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
#include <cstdlib>

#include "kernels.cuh"
#include "chol_chain.cuh"

namespace cslam {

namespace {

constexpr int BL_THREADS = 512;
constexpr int BL_WORKERS = BL_THREADS - 32;  // warps 1..7 do the bulk update, warp 0 the look-ahead

// 6x6 Cholesky of a symmetric block (lower triangle read from `A`, row-major) entirely in
// registers: L (lower, zeros above) and Li = L^-1.  Returns false when a pivot is not positive.
// Right-looking: the only dependent chain is  rsqrt -> scale the next row's entry -> its diagonal
// (~85 cycles per pivot; the left-looking dot products it replaces were chains of up to five FMAs
// in front of every rsqrt).
__device__ __forceinline__ bool potrf6_inv_reg(const double* A, double* Lout, double* Liout) {
    double L[6][6], inv[6];
#pragma unroll
    for (int i = 0; i < 6; ++i)
#pragma unroll
        for (int j = 0; j < 6; ++j) L[i][j] = (j <= i) ? A[6 * i + j] : 0.0;
    bool ok = true;
#pragma unroll
    for (int j = 0; j < 6; ++j) {
        const double d = L[j][j];
        ok = ok && pivot_ok(d);
        const double id = rsqrt_nr(d);
        inv[j] = id;
        L[j][j] = d * id;
#pragma unroll
        for (int i = j + 1; i < 6; ++i) L[i][j] *= id;
#pragma unroll
        for (int i = j + 1; i < 6; ++i)
#pragma unroll
            for (int c = j + 1; c <= i; ++c) L[i][c] -= L[i][j] * L[c][j];
    }
    double Li[6][6];
#pragma unroll
    for (int c = 0; c < 6; ++c) {
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            if (i < c) {
                Li[i][c] = 0.0;
            } else if (i == c) {
                Li[i][c] = inv[c];
            } else {
                double s = 0.0;
#pragma unroll
                for (int k = c; k < i; ++k) s -= L[i][k] * Li[k][c];
                Li[i][c] = s * inv[i];
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
    constexpr int ACC = kSpike ? (b + (BL_WORKERS / GC) - 1) / (BL_WORKERS / GC) : 1;  // left columns per thread
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
            // trailing window: A_ij -= L_ik L_jk^T (pair (1,1) is warp 0's).  One item = one row r
            // of one pair and three of its columns: the L_ik row stays in registers.
            const int npairs = nb * (nb + 1) / 2;
            for (int item = 12 + t; item < npairs * 12; item += BL_WORKERS) {
                const int pr = item / 12, rh = item % 12, r = rh >> 1, c0 = (rh & 1) * 3;
                const int di = pair_i[pr], dj = pair_j[pr];
                const double2* Li2 = reinterpret_cast<const double2*>(Lcol + di * 36 + 6 * r);
                const double2 l0 = Li2[0], l1 = Li2[1], l2 = Li2[2];
                int si = sk + di, sj = sk + dj;
                si -= si >= W1 ? W1 : 0;
                sj -= sj >= W1 ? W1 : 0;
                double* Arow = Awin + (si * W1 + sj) * 36 + 6 * r + c0;
#pragma unroll
                for (int cc = 0; cc < 3; ++cc) {
                    const double2* Lj2 = reinterpret_cast<const double2*>(Lcol + dj * 36 + 6 * (c0 + cc));
                    const double2 m0 = Lj2[0], m1 = Lj2[1], m2 = Lj2[2];
                    Arow[cc] -= l0.x * m0.x + l0.y * m0.y + l1.x * m1.x + l1.y * m1.y + l2.x * m2.x + l2.y * m2.y;
                }
            }
            // border G_i -= L_ik X_k and X^T X: a thread owns one border column (its X_k column
            // stays in registers) and every NG-th (block, row) / left column
            constexpr int NG = BL_WORKERS / GC;
            const int col = t % GC, grp = t / GC;
            if (grp < NG) {
                double x[6];
#pragma unroll
                for (int q = 0; q < 6; ++q) x[q] = Xk[q * GC + col];
                for (int dr = grp; dr < nb * 6; dr += NG) {
                    const int d = dr / 6 + 1, r = dr % 6;
                    const double2* Li2 = reinterpret_cast<const double2*>(Lcol + d * 36 + 6 * r);
                    const double2 l0 = Li2[0], l1 = Li2[1], l2 = Li2[2];
                    int si = sk + d;
                    si -= si >= W1 ? W1 : 0;
                    Gwin[(si * 6 + r) * GC + col] -= l0.x * x[0] + l0.y * x[1] + l1.x * x[2] + l1.y * x[3] + l2.x * x[4] + l2.y * x[5];
                }
                if (kSpike && has_left) {
#pragma unroll
                    for (int u = 0; u < ACC; ++u) {
                        const int c1 = 1 + grp + NG * u;
                        if (c1 < GC) {
                            double v = 0.0;
#pragma unroll
                            for (int q = 0; q < 6; ++q) v += Xk[q * GC + c1] * x[q];
                            acc[u] += v;
                        }
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
            // cyclic reduction (second generation) reads the coupling of an even separator transposed
            const int cidx = (B.sep_solver == 2 && W <= 9 && !(p & 1)) ? Cc * b + R : idx;
            B.Ca[(long long)p * b * b + cidx] = has_left ? Gwin[((i % W1) * 6 + R % 6) * GC + (kSpike ? 1 + Cc : 0)] : 0.0;
        }
        for (int R = tid; R < b; R += BL_THREADS) B.fa[(long long)p * b + R] = Gwin[(((ie + R / 6) % W1) * 6 + R % 6) * GC];
    }
    if (kSpike && has_left && tid >= 32) {
        constexpr int NG = BL_WORKERS / GC;
        const int t = tid - 32, c2 = t % GC, grp = t / GC;
        if (grp < NG) {
#pragma unroll
            for (int u = 0; u < ACC; ++u) {
                const int c1 = 1 + grp + NG * u;  // left column (1-based border column)
                if (c1 < GC) {
                    if (c2 == 0)
                        B.fb[(long long)(p - 1) * b + (c1 - 1)] = acc[u];
                    else
                        B.Tb[(long long)(p - 1) * b * b + (c1 - 1) * b + (c2 - 1)] = acc[u];
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Leaf kernel, second generation (leaves with a left spike, W <= 9).
//
// band_leaf_kernel above keeps the window in shared memory and spends ~7 instructions per FMA on index
// arithmetic and shared-memory round trips (ncu: 43 % of the issue slots, 17 % of the FP64 pipe, 2.8 us
// per block step).  Here every COLUMN of the moving window belongs to one thread that keeps it in
// registers, statically indexed, sliding down by six registers per block step (the same trick as the
// cyclic-reduction kernel), and the step is a pipeline of five warp roles joined by named barriers:
//   A (2 warps)  one thread per scalar column of the band window [6 (W+1) rows]: rank-6 trailing update
//                a[i] -= L_col[i] . L_col[j]; the columns of the NEXT pivot block are updated first and
//                handed to the pivot warp while the rest of the update is still running (look-ahead);
//   pivot        6x6 Cholesky of the pivot block and its inverse (potrf6_inv_reg);
//   panel        L_ik = A_ik Lkk^-T for the block column (written to shared memory and to Lbuf);
//   G (2 warps)  one thread per border column [rhs | left spike]: X_k = Lkk^-1 G_k, then the same
//                rank-6 update of its column;
//   T (2 warps)  one thread per border column: X_left^T [x_r | X_left] accumulated in registers.
// A, G and T each issue 6 (6W) FMAs per thread and step against L broadcast from shared memory:
// ~1 LDS.128 per 4 FMAs and no address arithmetic.
// ---------------------------------------------------------------------------------------------
template <int W>
struct Leaf2 {
    static constexpr int W1 = W + 1, NR = 6 * W1, NB = 6 * W, GC = 1 + 6 * W;
    static constexpr int XLD = 2 + NB;  // X_k row in shared memory: [x_r, pad, spike columns] (even stride)
    static constexpr int THREADS = 256;
};
__device__ __forceinline__ void nbar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void nbar_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }

template <int W>
__global__ void __launch_bounds__(256, 1) band_leaf2_kernel(BandView B) {
    using C = Leaf2<W>;
    constexpr int W1 = C::W1, NR = C::NR, NB = C::NB, GC = C::GC, XLD = C::XLD, b = 6 * W;
    // barrier ids and thread counts (arrivals + waiters)
    constexpr int B1A = 1, B1B = 2, B2 = 3, B3 = 4, B4 = 5, B5 = 6;
    constexpr int N1A = 64 + 32, N1B = 64 + 64 + 32 + 32, N2 = 32 + 32 + 64, N3 = 32 + 64 + 64, N4 = 128, N5 = 128;
    __shared__ __align__(16) double Pcol[NR * 6];   // the pivot block's columns: Pcol[row][c]
    __shared__ __align__(16) double Linv[36];
    __shared__ __align__(16) double Lcol[NB * 6];   // L_{k+d,k} rows below the pivot block
    __shared__ __align__(16) double Xk[6 * XLD];
    __shared__ int s_bad;
    const int tid = threadIdx.x, role = tid >> 6, t = tid & 63;
    const int p = blockIdx.x;
    const int s = p * B.m;
    const int e = min(B.n, s + B.m);
    const bool has_left = p > 0, has_right = p < B.P - 1;
    const int ie = has_right ? e - W : e;
    if (tid == 0) s_bad = 0;
    // A[6 I + rr][6 J + c] for block row I >= block column J inside the band (0 outside the leaf or the band)
    auto band_val = [&](int I, int J, int rr, int c) -> double {
        if (I >= e || I - J > B.w || J < 0) return 0.0;
        const int ent = B.band_idx[(long long)J * (B.w + 1) + (I - J)];
        return ent < 0 ? 0.0 : B.S[36ll * ent + 6 * c + rr];
    };
    __syncthreads();
    if (ie <= s) {
        // nothing to eliminate (cannot happen with the leaf sizes the planner picks); fall through to the outputs
    }
    if (role == 0) {
        // ================= A: window columns =================
        const bool live = t < NR;
        int jj = t;                      // window-relative column index
        int J = s + t / 6;               // block column
        const int c = t % 6;
        double a[NR];
#pragma unroll
        for (int ii = 0; ii < NR; ++ii) a[ii] = (live && ii >= jj) ? band_val(s + ii / 6, J, ii % 6, c) : 0.0;
        if (ie > s) {
            if (live && jj < 6) {
#pragma unroll
                for (int ii = 0; ii < NR; ++ii) Pcol[ii * 6 + jj] = a[ii];
            }
            nbar_arrive(B1A, N1A);
            nbar_arrive(B1B, N1B);
        }
        for (int k = s; k < ie; ++k) {
            // the row block entering the window: A[k+W+1][J] (issued before the wait)
            const int In = k + W1;
            const bool retiring = jj < 6;
            const int Jn = retiring ? In : J;  // a retiring thread takes over the first column block of the new rows
            double nw[6];
#pragma unroll
            for (int rr = 0; rr < 6; ++rr) nw[rr] = 0.0;
            if (live && In < e && In - Jn <= B.w) {
                const int ent = B.band_idx[(long long)Jn * (B.w + 1) + (In - Jn)];
                if (ent >= 0) {
                    const double2* src = reinterpret_cast<const double2*>(B.S + 36ll * ent + 6 * c);
                    const double2 v0 = src[0], v1 = src[1], v2 = src[2];
                    nw[0] = v0.x, nw[1] = v0.y, nw[2] = v1.x, nw[3] = v1.y, nw[4] = v2.x, nw[5] = v2.y;
                }
            }
            nbar_sync(B3, N3);  // L_col of step k
            double myL[6];
            {
                const int rj = (live && !retiring) ? jj - 6 : 0;
                const double2* lp = reinterpret_cast<const double2*>(Lcol + rj * 6);
                const double2 l0 = lp[0], l1 = lp[1], l2 = lp[2];
                const double m = (live && !retiring) ? 1.0 : 0.0;
                myL[0] = m * l0.x, myL[1] = m * l0.y, myL[2] = m * l1.x, myL[3] = m * l1.y, myL[4] = m * l2.x, myL[5] = m * l2.y;
            }
            const bool next_pivot = live && jj >= 6 && jj < 12 && k + 1 < ie;
            // part 1: the rows of the next pivot block
#pragma unroll
            for (int ii = 6; ii < 12; ++ii) {
                const double2* lp = reinterpret_cast<const double2*>(Lcol + (ii - 6) * 6);
                const double2 l0 = lp[0], l1 = lp[1], l2 = lp[2];
                a[ii - 6] = a[ii] - (l0.x * myL[0] + l0.y * myL[1] + l1.x * myL[2] + l1.y * myL[3] + l2.x * myL[4] + l2.y * myL[5]);
            }
            if (next_pivot) {
#pragma unroll
                for (int ii = 0; ii < 6; ++ii) Pcol[ii * 6 + (jj - 6)] = a[ii];
            }
            if (k + 1 < ie) nbar_arrive(B1A, N1A);
            // part 2: the rest
#pragma unroll
            for (int ii = 12; ii < NR; ++ii) {
                const double2* lp = reinterpret_cast<const double2*>(Lcol + (ii - 6) * 6);
                const double2 l0 = lp[0], l1 = lp[1], l2 = lp[2];
                a[ii - 6] = a[ii] - (l0.x * myL[0] + l0.y * myL[1] + l1.x * myL[2] + l1.y * myL[3] + l2.x * myL[4] + l2.y * myL[5]);
            }
            if (retiring) {
#pragma unroll
                for (int ii = 0; ii < NR - 6; ++ii) a[ii] = 0.0;
            }
#pragma unroll
            for (int rr = 0; rr < 6; ++rr) a[NR - 6 + rr] = nw[rr];
            if (next_pivot) {
#pragma unroll
                for (int ii = 6; ii < NR; ++ii) Pcol[ii * 6 + (jj - 6)] = a[ii];
            }
            if (k + 1 < ie) nbar_arrive(B1B, N1B);
            jj -= 6;
            if (jj < 0) {
                jj += NR;
                J += W1;
            }
        }
        // the window now starts at block row ie: separator block (lower triangle, mirrored)
        if (has_right && live && jj < NB) {
#pragma unroll
            for (int ii = 0; ii < NB; ++ii)
                if (ii >= jj) {
                    B.Ta[(long long)p * b * b + ii * b + jj] = a[ii];
                    B.Ta[(long long)p * b * b + jj * b + ii] = a[ii];
                }
        }
    } else if (role == 1) {
        // ================= G: border columns [rhs | left spike] =================
        const bool live = t < GC;
        const int xc = t == 0 ? 0 : 1 + t;  // position in an X_k row of shared memory
        const int cb = s - W + (t - 1) / 6, cq = (t - 1) % 6;  // spike column: block / component of the separator before
        double g[NR];
#pragma unroll
        for (int ii = 0; ii < NR; ++ii) {
            const int I = s + ii / 6, rr = ii % 6;
            double v = 0.0;
            if (live && I < e) {
                if (t == 0)
                    v = B.rhs[6ll * I + rr];
                else if (has_left)
                    v = band_val(I, cb, rr, cq);
            }
            g[ii] = v;
        }
        if (ie > s) nbar_arrive(B1B, N1B);
        for (int k = s; k < ie; ++k) {
            const int In = k + W1;
            double nw[6];
#pragma unroll
            for (int rr = 0; rr < 6; ++rr) nw[rr] = (live && t == 0 && In < e) ? B.rhs[6ll * In + rr] : 0.0;
            nbar_sync(B2, N2);  // Lkk^-1
            double x[6];
#pragma unroll
            for (int r = 0; r < 6; ++r) {
                double v = 0.0;
#pragma unroll
                for (int q = 0; q <= r; ++q) v += Linv[6 * r + q] * g[q];
                x[r] = v;
            }
            if (k > s) nbar_sync(B5, N5);  // the T warps have read the previous X_k
            if (live) {
#pragma unroll
                for (int q = 0; q < 6; ++q) {
                    Xk[q * XLD + xc] = x[q];
                    B.Xbuf[(long long)k * 6 * GC + q * GC + t] = x[q];
                }
            }
            nbar_arrive(B4, N4);
            nbar_sync(B3, N3);  // L_col of step k
#pragma unroll
            for (int ii = 6; ii < NR; ++ii) {
                const double2* lp = reinterpret_cast<const double2*>(Lcol + (ii - 6) * 6);
                const double2 l0 = lp[0], l1 = lp[1], l2 = lp[2];
                g[ii - 6] = g[ii] - (l0.x * x[0] + l0.y * x[1] + l1.x * x[2] + l1.y * x[3] + l2.x * x[4] + l2.y * x[5]);
            }
#pragma unroll
            for (int rr = 0; rr < 6; ++rr) g[NR - 6 + rr] = nw[rr];
            if (k + 1 < ie) nbar_arrive(B1B, N1B);
        }
        if (has_right && live) {
            if (t == 0) {
#pragma unroll
                for (int ii = 0; ii < NB; ++ii) B.fa[(long long)p * b + ii] = g[ii];
            } else {
                // cyclic reduction (second generation) reads the coupling of an even separator transposed
                const bool tr = B.sep_solver == 2 && !(p & 1);
#pragma unroll
                for (int ii = 0; ii < NB; ++ii)
                    B.Ca[(long long)p * b * b + (tr ? (t - 1) * b + ii : ii * b + (t - 1))] = has_left ? g[ii] : 0.0;
            }
        }
    } else if (role == 2) {
        // ================= T: X_left^T [x_r | X_left] =================
        const bool live = t < GC;
        const int xc = t == 0 ? 0 : 1 + t;
        double acc[NB];
#pragma unroll
        for (int c1 = 0; c1 < NB; ++c1) acc[c1] = 0.0;
        for (int k = s; k < ie; ++k) {
            nbar_sync(B4, N4);
            if (has_left && live) {
#pragma unroll
                for (int q = 0; q < 6; ++q) {
                    const double xq = Xk[q * XLD + xc];
                    const double2* xr = reinterpret_cast<const double2*>(Xk + q * XLD + 2);
#pragma unroll
                    for (int c1 = 0; c1 < NB; c1 += 2) {
                        const double2 xv = xr[c1 >> 1];
                        acc[c1] += xv.x * xq;
                        acc[c1 + 1] += xv.y * xq;
                    }
                }
            }
            nbar_arrive(B5, N5);
        }
        if (has_left && live) {
            if (t == 0) {
#pragma unroll
                for (int c1 = 0; c1 < NB; ++c1) B.fb[(long long)(p - 1) * b + c1] = acc[c1];
            } else {
#pragma unroll
                for (int c1 = 0; c1 < NB; ++c1) B.Tb[(long long)(p - 1) * b * b + c1 * b + (t - 1)] = acc[c1];
            }
        }
    } else if (t < 32) {
        // ================= panel warp =================
        for (int k = s; k < ie; ++k) {
            nbar_sync(B1B, N1B);  // the pivot block's columns are complete
            nbar_sync(B2, N2);    // Lkk^-1
            const int nb = min(k + W, e - 1) - k;
#pragma unroll
            for (int u = 0; u < (NB + 31) / 32; ++u) {
                const int r = t + 32 * u;
                if (r < NB) {
                    const double2* pp = reinterpret_cast<const double2*>(Pcol + (6 + r) * 6);
                    const double2 p0 = pp[0], p1 = pp[1], p2 = pp[2];
                    const double pr[6] = {p0.x, p0.y, p1.x, p1.y, p2.x, p2.y};
                    double l[6];
#pragma unroll
                    for (int c = 0; c < 6; ++c) {
                        double v = 0.0;
#pragma unroll
                        for (int q = 0; q <= c; ++q) v += pr[q] * Linv[6 * c + q];
                        l[c] = v;
                    }
                    double2* lo = reinterpret_cast<double2*>(Lcol + r * 6);
                    lo[0] = make_double2(l[0], l[1]);
                    lo[1] = make_double2(l[2], l[3]);
                    lo[2] = make_double2(l[4], l[5]);
                    if (r / 6 < nb) {
                        double2* go = reinterpret_cast<double2*>(B.Lbuf + ((long long)k * W1 + 1 + r / 6) * 36 + 6 * (r % 6));
                        go[0] = make_double2(l[0], l[1]);
                        go[1] = make_double2(l[2], l[3]);
                        go[2] = make_double2(l[4], l[5]);
                    }
                }
            }
            nbar_arrive(B3, N3);
        }
    } else {
        // ================= pivot warp =================
        const int lane = t - 32;
        for (int k = s; k < ie; ++k) {
            nbar_sync(B1A, N1A);  // rows 0..5 of the pivot block's columns
            if (lane == 0) {
                double Lkk[36], Li[36];
                if (!potrf6_inv_reg(Pcol, Lkk, Li)) s_bad = 1;
#pragma unroll
                for (int q = 0; q < 36; ++q) Linv[q] = Li[q];
            }
            __syncwarp();
            for (int idx = lane; idx < 36; idx += 32) B.Lbuf[((long long)k * W1) * 36 + idx] = Linv[idx];  // inverse of the pivot block first
            nbar_sync(B1B, N1B);  // everyone is done with step k-1 (keeps the arrivals below one phase apart)
            nbar_arrive(B2, N2);
        }
        if (lane == 0 && s_bad) *B.fail = 1;
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
//
// The recurrence y_k = Lkk^-T (z_k - sum_{d>=1} L_{k+d,k}^T y_{k+d}) is sequential in k; one warp
// walks it.  Only the d = 1 term depends on the value finished in the previous step, so the
// d >= 2 terms of step k-1 are formed (lane groups 1..4) while step k is completed (lane group 0),
// and the factor columns stream through a cp.async ring several steps ahead of their use.
constexpr int BBS_T = 512;  // one warp walks the recurrence; the rest only helps with the two parallel passes around it
template <int W, bool kSpike>
__global__ void __launch_bounds__(BBS_T) band_backsub_kernel(BandView B, const double* ysep) {
    constexpr int W1 = W + 1, GC = kSpike ? 1 + 6 * W : 1, b = 6 * W;
    constexpr int COLD = W1 * 36;           // doubles of one factor column
    constexpr int NST = 8, PD = 6;          // ring stages, prefetch distance
    constexpr int DGB = (W - 1 + 3) / 4;    // d >= 2 blocks per lane group
    extern __shared__ __align__(16) double smem_bb[];
    const int tid = threadIdx.x;
    const int p = blockIdx.x;
    const int s = p * B.m;
    const int e = min(B.n, s + B.m);
    const bool has_left = kSpike && p > 0, has_right = p < B.P - 1;
    const int ie = has_right ? e - W : e;
    const int ni = ie - s;
    double* ring = smem_bb;                  // [NST][COLD]
    double* z = ring + NST * COLD;           // [6 (m + W)] leaf rows: right-hand side, then the solution
    double* yl = z + 6 * (B.m + W);          // [b] separator before
    if (*B.fail) return;
    if (has_left)
        for (int c = tid; c < b; c += BBS_T) yl[c] = ysep[(long long)(p - 1) * b + c];
    if (has_right)
        for (int c = tid; c < b; c += BBS_T) {
            const double v = ysep[(long long)p * b + c];
            z[6 * ni + c] = v;
            B.y[6ll * ie + c] = v;
        }
    __syncthreads();
    if (kSpike && has_left) {
        // z = x_r - X_left y_left: a warp per row, the row's b spike entries read coalesced (a thread per row walked
        // its 1 + b values one dependent global load at a time: 60 % of this kernel's time), four rows in flight
        const int wp = tid >> 5, ln = tid & 31;
        constexpr int NC = (b + 31) / 32;
        double ylr[NC];
#pragma unroll
        for (int u = 0; u < NC; ++u) ylr[u] = ln + 32 * u < b ? yl[ln + 32 * u] : 0.0;
        for (int row0 = 4 * wp; row0 < 6 * ni; row0 += 4 * (BBS_T / 32)) {
            double v[4], x0[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int row = row0 + q;
                v[q] = 0.0;
                x0[q] = 0.0;
                if (row < 6 * ni) {
                    const double* X = B.Xbuf + (long long)(6 * s + row) * GC;
                    x0[q] = X[0];
#pragma unroll
                    for (int u = 0; u < NC; ++u)
                        if (ln + 32 * u < b) v[q] += X[1 + ln + 32 * u] * ylr[u];
                }
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const double sum = warp_sum(v[q]);
                if (ln == 0 && row0 + q < 6 * ni) z[row0 + q] = x0[q] - sum;
            }
        }
    } else {
        for (int idx = tid; idx < 6 * ni; idx += BBS_T) z[idx] = B.Xbuf[(long long)(6 * s + idx) * GC];
    }
    __syncthreads();
    if (tid < 32 && ni > 0) {
        const int lane = tid, g = lane / 6, c = lane % 6;
        auto issue = [&](int kk) {
            if (kk >= s) {
                const double* src = B.Lbuf + (long long)kk * COLD;
                double* dst = ring + ((kk - s) % NST) * COLD;
                for (int idx = lane; idx < COLD / 2; idx += 32) cp_async16(dst + 2 * idx, src + 2 * idx);
            }
            cp_async_commit();
        };
        for (int q = 0; q < PD; ++q) issue(ie - 1 - q);
        double zt = 0.0;  // lanes 0..5: z_k[c] minus the d >= 2 terms of step k
        for (int k = ie; k >= s; --k) {
            issue(k - 1 - PD);
            cp_async_wait<PD>();  // column k-1 has landed (k was waited for one step earlier)
            __syncwarp();
            const int kk = k - 1;
            // lane groups 1..4: d >= 2 terms of step kk (every y they need is final)
            double part = 0.0;
            if (kk >= s && g >= 1 && g <= 4) {
                const double* Lc = ring + ((kk - s) % NST) * COLD;
                const int nbk = min(kk + W, e - 1) - kk;
#pragma unroll
                for (int u = 0; u < DGB; ++u) {
                    const int d = 2 + (g - 1) * DGB + u;
                    if (d <= nbk) {
                        const double* yv = z + 6 * (kk + d - s);
#pragma unroll
                        for (int r = 0; r < 6; ++r) part += Lc[d * 36 + 6 * r + c] * yv[r];
                    }
                }
            }
            // lane group 0: finish step k
            double t = 0.0;
            const double* Lk = ring + (((k < ie ? k : ie - 1) - s) % NST) * COLD;
            if (k < ie && lane < 6) {
                t = zt;
                if (k + 1 < e) {
                    const double* yv = z + 6 * (k + 1 - s);
#pragma unroll
                    for (int r = 0; r < 6; ++r) t -= Lk[36 + 6 * r + c] * yv[r];
                }
            }
            double ycur = 0.0;
#pragma unroll
            for (int i = 0; i < 6; ++i) {
                const double ti = __shfl_sync(0xffffffffu, t, i);
                if (k < ie && lane < 6 && i >= c) ycur += Lk[6 * i + c] * ti;  // (Lkk^-T t)_c
            }
            if (k < ie && lane < 6) z[6 * (k - s) + c] = ycur;
            // right-hand side of the next step: z_kk minus the partial sums of groups 1..4
            double psum = 0.0;
#pragma unroll
            for (int gg = 1; gg <= 4; ++gg) psum += __shfl_sync(0xffffffffu, part, (c + 6 * gg) & 31);
            if (lane < 6 && kk >= s) zt = z[6 * (kk - s) + c] - psum;
            __syncwarp();
        }
        cp_async_wait<0>();
    }
    __syncthreads();
    for (int idx = tid; idx < 6 * ni; idx += BBS_T) B.y[6ll * s + idx] = z[idx];
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

template <int W>
void run_leaf2(cudaStream_t s, const BandView& V) {
    band_leaf2_kernel<W><<<V.P, 256, 0, s>>>(V);
    CSLAM_CUDA(cudaGetLastError());
}

template <int W, bool kSpike>
void run_leaf(cudaStream_t s, const BandView& V) {
    static PerDevice attr;
    const int dev_ = PerDevice::current();
    if (attr.first_use(dev_)) {
        CSLAM_CUDA(cudaFuncSetAttribute(band_leaf_kernel<W, kSpike>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        int(leaf_smem<W, kSpike>())));
        attr.mark(dev_);
    }
    band_leaf_kernel<W, kSpike><<<V.P, BL_THREADS, leaf_smem<W, kSpike>(), s>>>(V);
    CSLAM_CUDA(cudaGetLastError());
}

template <int W, bool kSpike>
void run_backsub(cudaStream_t s, const BandView& V, const double* ysep) {
    static PerDevice attr;
    const int dev_ = PerDevice::current();
    if (attr.first_use(dev_)) {
        CSLAM_CUDA(cudaFuncSetAttribute(band_backsub_kernel<W, kSpike>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        attr.mark(dev_);
    }
    const size_t smem = sizeof(double) * (8 * size_t(W + 1) * 36 + 6 * size_t(V.m + W) + 6 * size_t(W));
    band_backsub_kernel<W, kSpike><<<V.P, BBS_T, smem, s>>>(V, ysep);
    CSLAM_CUDA(cudaGetLastError());
}


// ---------------------------------------------------------------------------------------------
// Separator system by block cyclic reduction.
//
// The separators form a block-TRIDIAGONAL system with b x b blocks, b = 6 W: separator q couples
// only to q-1 and q+1 (through the leaf between them): D_q = Ta[q] - Tb[q], coupling
// E(q, q-1) = Ca[q], right-hand side fa[q] - fb[q].  Factoring it as a band (above) is a chain of
// (P-1) W dependent block steps on ONE CTA; cyclic reduction eliminates every other separator at
// once — log2(P-1) levels of independent b x b factorisations — so the band can be cut into many
// more, shorter leaves.  Level with stride s, "odd" rows i = s, 3s, 5s, ...:
//     D_i = L L^T,  A = L^-1 E(i, i-s),  Bm = L^-1 E(i+s, i)^T,  g = L^-1 f_i,  Li = L^-1
// "even" rows i = 0, 2s, 4s, ...:
//     D_i -= Bm_{i-s}^T Bm_{i-s} + A_{i+s}^T A_{i+s},  f_i -= Bm_{i-s}^T g_{i-s} + A_{i+s}^T g_{i+s},
//     E(i, i-2s) = -Bm_{i-s}^T A_{i-s}
// and back down:  y_i = Li^T (g - A y_{i-s} - Bm y_{i+s}).
// Storage reuses the band path's buffers: D in place of Ta, E in place of Ca, f = rhs2, A | Bm = T2,
// Li = L2, g = X2, y = y2.
// ---------------------------------------------------------------------------------------------
struct BcrView {
    int N, b;
    double *D, *E, *f;         // [N][b*b], [N][b*b] (E[i] = coupling (i, i - s) of the current level), [N][b]
    double *A, *Bm, *Li, *g;   // per eliminated row
    double* y;                 // [N][b]
    int* fail;
};

__global__ void bcr_init_kernel(BcrView R, const double* __restrict__ Tb, const double* __restrict__ fa,
                                const double* __restrict__ fb) {
    const long long bb = (long long)R.b * R.b, total = (long long)R.N * bb;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total + (long long)R.N * R.b;
         idx += (long long)gridDim.x * blockDim.x) {
        if (idx < total)
            R.D[idx] -= Tb[idx];
        else
            R.f[idx - total] = fa[idx - total] - fb[idx - total];
    }
}

// rows i = first + blockIdx.x * step; left / right neighbours at distance s (absent when out of range
// or when s == 0: the last row standing).  Right-looking BLOCKED Cholesky of D_i (6x6 pivots, as
// in the leaf kernel) with the border [E_left | E_right^T | f | I] riding along: W block steps of
//   (1) pivot      Lkk, Lkk^-1 of the 6x6 pivot in registers (one thread)
//   (2) panel      L_rk = A_rk Lkk^-T for the rows below, X_k = Lkk^-1 G_k for the border row block
//   (3) trailing   A_rc -= L_rk . L_ck   (symmetric part)   /   G_rc -= L_rk . X_kc   (border)
// i.e. 3 barriers per SIX pivots and rank-6 updates.  X_k is final after (2) and is written to
// A | Bm | g | Li while (3) runs.
constexpr int BCR_T = 512;
__global__ void __launch_bounds__(BCR_T) bcr_odd_kernel(BcrView R, int first, int step, int s) {
    extern __shared__ __align__(16) double smem_bcr[];
    __shared__ double sLi[72];  // Lkk^-1 of the current / next pivot block
    __shared__ int s_ok;
    const int b = R.b, ld = 4 * b + 1, nbord = 3 * b + 1, tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
    constexpr int TY = BCR_T / 32;
    const int i = first + blockIdx.x * step;
    const bool has_l = s > 0 && i - s >= 0, has_r = s > 0 && i + s < R.N;
    const long long bb = (long long)b * b;
    double* M = smem_bcr;  // [b][ld]: D | E_left | E_right^T | f | I
    if (*R.fail) return;
    const double* Di = R.D + i * bb;
    const double* El = R.E + i * bb;
    const double* Er = R.E + (long long)(i + s) * bb;
    for (int r = ty; r < b; r += TY) {
        double* Mr = M + r * ld;
        for (int c = tx; c < b; c += 32) {
            Mr[c] = Di[r * b + c];
            Mr[b + c] = has_l ? El[r * b + c] : 0.0;
            Mr[3 * b + 1 + c] = (c == r) ? 1.0 : 0.0;
        }
        if (tx == 0) Mr[3 * b] = R.f[(long long)i * b + r];
    }
    // E_right^T: read E(i+s, i) row-wise (coalesced), store transposed
    for (int r = ty; r < b; r += TY)
        for (int c = tx; c < b; c += 32) M[c * ld + 2 * b + r] = has_r ? Er[r * b + c] : 0.0;
    if (tid == 0) s_ok = 1;
    __syncthreads();
    double* Ai = R.A + i * bb;
    double* Bi = R.Bm + i * bb;
    double* Li = R.Li + i * bb;
    const int nblk = b / 6;
    // pivot of block step 0
    if (tid == 0) {
        double P[36], L[36], Linv[36];
#pragma unroll
        for (int a = 0; a < 6; ++a)
#pragma unroll
            for (int q = 0; q < 6; ++q) P[6 * a + q] = M[a * ld + q];
        if (!potrf6_inv_reg(P, L, Linv)) s_ok = 0;
#pragma unroll
        for (int q = 0; q < 36; ++q) sLi[q] = Linv[q];
    }
    __syncthreads();
    for (int kb = 0; kb < nblk; ++kb) {
        if (!s_ok) break;
        const int k0 = 6 * kb;
        const double* Lcur = sLi + 36 * (kb & 1);
        double* Lnext = sLi + 36 * ((kb + 1) & 1);
        // ---- (2) panel: rows below the pivot, and the pivot's border row block --------------------
        {
            const int nrow = b - k0 - 6;
            for (int t = tid; t < nrow + nbord; t += BCR_T) {
                double v[6], o[6];
                if (t < nrow) {
                    double* p = M + (k0 + 6 + t) * ld + k0;  // A_rk, 6 contiguous entries of row r
#pragma unroll
                    for (int q = 0; q < 6; ++q) v[q] = p[q];
#pragma unroll
                    for (int a = 0; a < 6; ++a) {
                        double acc = 0.0;
#pragma unroll
                        for (int q = 0; q < 6; ++q)
                            if (q <= a) acc += v[q] * Lcur[6 * a + q];
                        o[a] = acc;
                    }
#pragma unroll
                    for (int q = 0; q < 6; ++q) p[q] = o[q];
                } else {
                    double* p = M + k0 * ld + b + (t - nrow);  // G_k column c, 6 entries with stride ld
#pragma unroll
                    for (int q = 0; q < 6; ++q) v[q] = p[q * ld];
#pragma unroll
                    for (int a = 0; a < 6; ++a) {
                        double acc = 0.0;
#pragma unroll
                        for (int q = 0; q < 6; ++q)
                            if (q <= a) acc += Lcur[6 * a + q] * v[q];
                        o[a] = acc;
                    }
#pragma unroll
                    for (int q = 0; q < 6; ++q) p[q * ld] = o[q];
                }
            }
        }
        __syncthreads();
        // ---- outputs: X_k (6 x border) is final -------------------------------------------------
        for (int t = tid; t < 6 * nbord; t += BCR_T) {
            const int a = t / nbord, c = t - a * nbord;
            const double v = M[(k0 + a) * ld + b + c];
            const int k = k0 + a;
            if (c < b)
                Ai[k * b + c] = v;
            else if (c < 2 * b)
                Bi[k * b + (c - b)] = v;
            else if (c == 2 * b)
                R.g[(long long)i * b + k] = v;
            else
                Li[k * b + (c - 2 * b - 1)] = v;
        }
        // ---- look-ahead: thread 0 brings the NEXT pivot block up to date and factors it while the
        // other warps apply the trailing update (which leaves that 6x6 block to it) ---------------
        if (tid == 0 && kb + 1 < nblk) {
            double P[36], L[36], Linv[36];
            const int k1 = k0 + 6;
#pragma unroll
            for (int a = 0; a < 6; ++a)
#pragma unroll
                for (int q = 0; q <= a; ++q) {
                    double acc = M[(k1 + a) * ld + k1 + q];
#pragma unroll
                    for (int m = 0; m < 6; ++m) acc -= M[(k1 + a) * ld + k0 + m] * M[(k1 + q) * ld + k0 + m];
                    P[6 * a + q] = acc;
                }
#pragma unroll
            for (int a = 0; a < 6; ++a)
#pragma unroll
                for (int q = a + 1; q < 6; ++q) P[6 * a + q] = 0.0;
            if (!potrf6_inv_reg(P, L, Linv)) s_ok = 0;
#pragma unroll
            for (int q = 0; q < 36; ++q) Lnext[q] = Linv[q];
        }
        // ---- (3) trailing update (two columns per pass: independent chains) -----------------------
        // columns k0+6 .. k0+11 of rows k0+6 .. k0+11 belong to the look-ahead thread; the rest of those
        // columns (rows below) is updated here as usual
        for (int r = k0 + 6 + (ty - 1); ty > 0 && r < b; r += TY - 1) {  // warp 0 is the look-ahead warp
            double* Mr = M + r * ld;
            double lr[6];
#pragma unroll
            for (int q = 0; q < 6; ++q) lr[q] = Mr[k0 + q];
            const bool in_next = r < k0 + 12;
            int c = k0 + 6 + tx;
            for (; c < b; c += 32) {
                if (in_next && c < k0 + 12) continue;  // next pivot block: thread 0's
                const double* Lc = M + c * ld + k0;
                double acc = 0.0;
#pragma unroll
                for (int q = 0; q < 6; ++q) acc += lr[q] * Lc[q];
                Mr[c] -= acc;
            }
            for (; c + 32 < ld; c += 64) {
                const double* X0 = M + k0 * ld + c;
                const double* X1 = X0 + 32;
                double a0 = 0.0, a1 = 0.0;
#pragma unroll
                for (int q = 0; q < 6; ++q) {
                    a0 += lr[q] * X0[q * ld];
                    a1 += lr[q] * X1[q * ld];
                }
                Mr[c] -= a0;
                Mr[c + 32] -= a1;
            }
            for (; c < ld; c += 32) {
                const double* Xc = M + k0 * ld + c;
                double acc = 0.0;
#pragma unroll
                for (int q = 0; q < 6; ++q) acc += lr[q] * Xc[q * ld];
                Mr[c] -= acc;
            }
        }
        __syncthreads();
    }
    if (!s_ok && tid == 0) *R.fail = 1;
}

// rows i = blockIdx.x * 2s.  The three b x b products are computed in 2 x 3 register tiles (six
// independent accumulation chains per product and thread).
__global__ void __launch_bounds__(BCR_T) bcr_even_kernel(BcrView R, int s) {
    extern __shared__ __align__(16) double smem_bcr[];
    const int b = R.b, tid = threadIdx.x;
    const int i = blockIdx.x * 2 * s;
    const bool has_l = i - s >= 0, has_r = i + s < R.N, has_ll = i - 2 * s >= 0;
    const long long bb = (long long)b * b;
    if (*R.fail) return;
    double* X1 = smem_bcr;          // Bm of the left odd neighbour
    double* X2 = X1 + b * b;        // A of the left odd neighbour
    double* X3 = X2 + b * b;        // A of the right odd neighbour
    double* gl = X3 + b * b;
    double* gr = gl + b;
    for (int idx = tid; idx < b * b; idx += BCR_T) {
        X1[idx] = has_l ? R.Bm[(long long)(i - s) * bb + idx] : 0.0;
        X2[idx] = has_l ? R.A[(long long)(i - s) * bb + idx] : 0.0;
        X3[idx] = has_r ? R.A[(long long)(i + s) * bb + idx] : 0.0;
    }
    for (int r = tid; r < b; r += BCR_T) {
        gl[r] = has_l ? R.g[(long long)(i - s) * b + r] : 0.0;
        gr[r] = has_r ? R.g[(long long)(i + s) * b + r] : 0.0;
    }
    __syncthreads();
    double* Di = R.D + i * bb;
    double* Ei = R.E + i * bb;
    const int tr = b / 2, tc = b / 3;  // b = 6 W: both divide
    for (int t = tid; t < tr * tc; t += BCR_T) {
        const int r0 = 2 * (t / tc), c0 = 3 * (t % tc);
        double dd[2][3] = {{0, 0, 0}, {0, 0, 0}}, ee[2][3] = {{0, 0, 0}, {0, 0, 0}};
#pragma unroll 2
        for (int k = 0; k < b; ++k) {
            const double* x1 = X1 + k * b;
            const double* x2 = X2 + k * b;
            const double* x3 = X3 + k * b;
            const double a1[2] = {x1[r0], x1[r0 + 1]}, a3[2] = {x3[r0], x3[r0 + 1]};
            const double b1[3] = {x1[c0], x1[c0 + 1], x1[c0 + 2]};
            const double b2[3] = {x2[c0], x2[c0 + 1], x2[c0 + 2]};
            const double b3[3] = {x3[c0], x3[c0 + 1], x3[c0 + 2]};
#pragma unroll
            for (int u = 0; u < 2; ++u)
#pragma unroll
                for (int w = 0; w < 3; ++w) {
                    dd[u][w] += a1[u] * b1[w] + a3[u] * b3[w];
                    ee[u][w] += a1[u] * b2[w];
                }
        }
#pragma unroll
        for (int u = 0; u < 2; ++u)
#pragma unroll
            for (int w = 0; w < 3; ++w) {
                const int idx = (r0 + u) * b + c0 + w;
                Di[idx] -= dd[u][w];
                if (has_ll) Ei[idx] = -ee[u][w];
            }
    }
    for (int r = tid; r < b; r += BCR_T) {
        double ff = 0.0;
        for (int k = 0; k < b; ++k) ff += X1[k * b + r] * gl[k] + X3[k * b + r] * gr[k];
        R.f[(long long)i * b + r] -= ff;
    }
}

// y_i = Li^T (g - A y_{i-s} - Bm y_{i+s}).  A_i, Bm_i and Li_i are staged in shared memory with one
// wave of coalesced loads (the products themselves are tiny; the kernel is load latency).
__global__ void __launch_bounds__(256) bcr_backsub_kernel(BcrView R, int first, int step, int s) {
    extern __shared__ __align__(16) double smem_bcr[];
    const int b = R.b, tid = threadIdx.x;
    const int i = first + blockIdx.x * step;
    const bool has_l = s > 0 && i - s >= 0, has_r = s > 0 && i + s < R.N;
    const long long bb = (long long)b * b;
    if (*R.fail) return;
    double* sA = smem_bcr;
    double* sB = sA + b * b;
    double* sL = sB + b * b;
    double* t = sL + b * b;
    double* yl = t + b;
    double* yr = yl + b;
    const double* Ai = R.A + i * bb;
    const double* Bi = R.Bm + i * bb;
    const double* Li = R.Li + i * bb;
    for (int idx = tid; idx < b * b; idx += 256) {
        sA[idx] = Ai[idx];
        sB[idx] = Bi[idx];
        sL[idx] = Li[idx];
    }
    for (int r = tid; r < b; r += 256) {
        yl[r] = has_l ? R.y[(long long)(i - s) * b + r] : 0.0;
        yr[r] = has_r ? R.y[(long long)(i + s) * b + r] : 0.0;
        t[r] = R.g[(long long)i * b + r];
    }
    __syncthreads();
    // 4 threads per row for the two products
    {
        const int r = tid >> 2, part = tid & 3;
        double v = 0.0;
        if (r < b)
            for (int k = part; k < b; k += 4) v += sA[r * b + k] * yl[k] + sB[r * b + k] * yr[k];
        v += __shfl_xor_sync(0xffffffffu, v, 1);
        v += __shfl_xor_sync(0xffffffffu, v, 2);
        __syncthreads();
        if (r < b && part == 0) t[r] -= v;
        // rows beyond 64 (b = 72)
        for (int r2 = 64 + (tid >> 2); r2 < b; r2 += 64) {
            double v2 = 0.0;
            for (int k = part; k < b; k += 4) v2 += sA[r2 * b + k] * yl[k] + sB[r2 * b + k] * yr[k];
            v2 += __shfl_xor_sync(0xffffffffu, v2, 1);
            v2 += __shfl_xor_sync(0xffffffffu, v2, 2);
            if (part == 0) t[r2] -= v2;
        }
    }
    __syncthreads();
    for (int r = tid; r < b; r += 256) {
        double v = 0.0;
#pragma unroll 6
        for (int k = r; k < b; ++k) v += sL[k * b + r] * t[k];  // (L^-T t)_r, L^-1 lower triangular
        R.y[(long long)i * b + r] = v;
    }
}


// ---------------------------------------------------------------------------------------------
// Cyclic reduction, second generation (sep_solver == 2).  One level used to cost ~65 us (odd 34 +
// even 24 + back-substitution 7): everything in it is latency, so the kernels are organised around
// the one chain that cannot be shortened — the B dependent pivots of the B x B Cholesky.
//
// bcr_odd2_kernel<B>: the first (B+31)/32 warps own the ROWS of D_i in registers (lane r: a[c] =
//   D[r][c], c <= r, statically indexed because the pivot loop is fully unrolled) and run a scalar
//   right-looking Cholesky whose per-pivot chain is  1/L_kk broadcast -> l = a[k] / L_kk ->
//   diag -= l^2 -> rsqrt  (the row's own diagonal lives in a register of its own, so the next pivot
//   never waits for shared memory); the rank-1 update of the other columns is issued behind the
//   rsqrt of the next pivot and fills its latency.  Column k of L is published in shared memory
//   (Lt[k][r]) together with a progress counter (release / acquire).
//   The border [E_left | E_right^T | f | I] is NOT part of that factorisation: every border column
//   belongs to one thread of the remaining warps, which keeps the whole column (B doubles) in
//   registers and performs the forward substitution x = L^-1 g right-looking, trailing the factor
//   warps by one pivot: x_k *= 1/L_kk, x_r -= L_rk x_k (r > k) with L_rk broadcast from shared
//   memory.  No barrier between the two roles, no shared-memory traffic for the border itself.
// bcr_even2_kernel<B>: the three B x B products of a surviving row are cut into (row, D | E,
//   strip-of-rows) tasks, one CTA each, so that late levels (a handful of rows) still use the chip.
// The coupling block E[j] is stored row-major when row j is eliminated at the level it is built
//   for ("odd": read as E_left) and transposed otherwise (read as E_right^T by its odd neighbour):
//   both reads of the odd kernel are then coalesced across the column threads.
// ---------------------------------------------------------------------------------------------
// raw != 0: the level-0 inputs have not been combined yet (D = Ta - Tb, f = fa - fb).
template <int B>
__global__ void __launch_bounds__(Odd2<B>::THREADS, 1)
    bcr_odd2_kernel(BcrView R, int first, int step, int s, int raw, const double* __restrict__ Tb,
                    const double* __restrict__ fa, const double* __restrict__ fb) {
    using C = Odd2<B>;
    constexpr int FW = C::FW, R0 = C::R0;
    // Lt2[(k + 1) * B + j] = L[k + 1 + j][k]; row 0 is zeros (pivot "-1"), one spare row behind
    __shared__ __align__(16) double Lt2[(B + 2) * B];
    __shared__ double sInv[B + 2];                 // 1 / L[k][k]
    __shared__ __align__(8) uint64_t done0[B + 1];  // done0[k + 1]: warp 0 published column k; done0[0]: 1 / L_00
    __shared__ __align__(8) uint64_t done1[B];      // done1[k]: warp 1 published column k (B > 32)
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int i = first + blockIdx.x * step;
    const bool has_l = s > 0 && i - s >= 0, has_r = s > 0 && i + s < R.N;
    constexpr long long bb = (long long)B * B;
    if (*R.fail) return;
    for (int j = tid; j < B + 1; j += C::THREADS) {
        if (j < B) {
            Lt2[j] = 0.0;
            mbar_init(&done1[j], 32);
        }
        mbar_init(&done0[j], 32);
    }
    fence_mbar_init();
    __syncthreads();
    uint64_t* col_done = FW > 1 ? done1 : done0 + 1;  // what the border waits on: column k complete
    const double* Di = R.D + i * bb;
    const double* Tbi = Tb + i * bb;
#ifdef CSLAM_ODD2_PROF
    const long long t_entry = clock64();
#define ODD2_T(name) const long long name = clock64();
#define ODD2_P(...) if (lane == 0 && blockIdx.x == 0 && s == 4) printf(__VA_ARGS__);
#else
#define ODD2_T(name)
#define ODD2_P(...)
#endif
    if (warp == 0) {
        // ---------------- factor warp 0: rows 0 .. R0-1, it never needs anything from other warps ----------------
        const int r = lane;
        const bool act = r < R0;
        const int rr = act ? r : 0;
        double b[R0 + 1];
        b[0] = 0.0;
#pragma unroll
        for (int c = 0; c < R0; ++c) b[c + 1] = Di[c * B + rr];  // D symmetric: column read, coalesced over the lanes
        double diag = Di[rr * B + rr];
        if (raw) {
#pragma unroll
            for (int c = 0; c < R0; ++c) b[c + 1] -= Tbi[c * B + rr];
            diag -= Tbi[rr * B + rr];
        }
        if (!act) diag = 1.0;
        double nid = rsqrt_nr(diag);
        bool bad = false;
        if (r == 0) {
            bad = !pivot_ok(diag);
            sInv[0] = nid;
        }
        __syncwarp();
        mbar_arrive(done0);
        double lprev = 0.0;
        constexpr int KH = R0 / 2;
        ODD2_T(t0)
        odd2_factor_phase<B, R0, R0 + 1, true>(b, diag, lprev, nid, 0, KH, r, 0, act, Lt2, sInv, nullptr, done0 + 1, bad);
        ODD2_T(t1)
        odd2_factor_phase<B, R0, R0 + 2 - KH, true>(b, diag, lprev, nid, KH, FW > 1 ? R0 - 1 : R0, r, 0, act, Lt2, sInv, nullptr,
                                                    done0 + 1, bad);
        ODD2_T(t2)
        ODD2_P("odd2 warp0: load %lld  k0-15 %lld  k16-30 %lld\n", t0 - t_entry, t1 - t0, t2 - t1)
        __syncwarp();
        mbar_arrive(done0 + (FW > 1 ? R0 - 1 : R0));  // the last column (arrivals run one pivot behind)
        if (__any_sync(0xffffffffu, bad) && lane == 0) *R.fail = 1;
    } else if (FW > 1 && warp == 1) {
        // ---------------- factor warp 1: rows 32 .. B-1; trails warp 0, then owns the chain ----------------
        const int r = 32 + lane;
        const bool act = r < B;
        const int rr = act ? r : 0;
        double b[B + 1];
        b[0] = 0.0;
#pragma unroll
        for (int c = 0; c < B; ++c) b[c + 1] = Di[c * B + rr];
        double diag = Di[rr * B + rr];
        if (raw) {
#pragma unroll
            for (int c = 0; c < B; ++c) b[c + 1] -= Tbi[c * B + rr];
            diag -= Tbi[rr * B + rr];
        }
        if (!act) diag = 1.0;
        double nid = 1.0, lprev = 0.0;
        bool bad = false;
        ODD2_T(t0)
        odd2_factor_phase<B, B, B + 1, false>(b, diag, lprev, nid, 0, 16, r, 32, act, Lt2, sInv, done0, done1, bad);
        ODD2_T(t1)
        odd2_factor_phase<B, B, B + 1 - 15, false>(b, diag, lprev, nid, 16, 32, r, 32, act, Lt2, sInv, done0, done1, bad);
        ODD2_T(t2)
        odd2_factor_phase<B, B, B + 1 - 31, true>(b, diag, lprev, nid, 32, B, r, 32, act, Lt2, sInv, nullptr, done1, bad);
        ODD2_T(t3)
        ODD2_P("odd2 warp1: load %lld  k0-15 %lld  k16-31 %lld  k32-53 %lld\n", t0 - t_entry, t1 - t0, t2 - t1, t3 - t2)
        __syncwarp();
        mbar_arrive(done1 + B - 1);
        if (__any_sync(0xffffffffu, bad) && lane == 0) *R.fail = 1;
    } else {
        // ---------------- border warps: one column of [E_left | E_right^T | f | I] per thread ----------------
        const int col = 32 * (warp - FW) + lane;
        double x[B];
        double* outp = nullptr;
        int ostride = B;
        if (col < B) {
            const double* El = R.E + i * bb;  // row-major
#pragma unroll
            for (int r = 0; r < B; ++r) x[r] = has_l ? El[r * B + col] : 0.0;
            outp = R.A + i * bb + col;
        } else if (col < 2 * B) {
            const double* Ert = R.E + (long long)(i + s) * bb;  // stored transposed
#pragma unroll
            for (int r = 0; r < B; ++r) x[r] = has_r ? Ert[r * B + (col - B)] : 0.0;
            outp = R.Bm + i * bb + (col - B);
        } else if (col == 2 * B) {
            if (raw) {
#pragma unroll
                for (int r = 0; r < B; ++r) x[r] = fa[(long long)i * B + r] - fb[(long long)i * B + r];
            } else {
#pragma unroll
                for (int r = 0; r < B; ++r) x[r] = R.f[(long long)i * B + r];
            }
            outp = R.g + (long long)i * B;
            ostride = 1;
        } else {
#pragma unroll
            for (int r = 0; r < B; ++r) x[r] = (r == col - 2 * B - 1) ? 1.0 : 0.0;
            if (col < C::NCOL) outp = R.Li + i * bb + (col - 2 * B - 1);
        }
        constexpr int K1 = B / 3, K2 = 2 * B / 3;
        ODD2_T(t0)
        odd2_border_phase<B, B>(x, 0, K1, Lt2, sInv, col_done, outp, ostride);
        odd2_border_phase<B, B - K1>(x, K1, K2, Lt2, sInv, col_done, outp, ostride);
        odd2_border_phase<B, B - K2>(x, K2, B, Lt2, sInv, col_done, outp, ostride);
        ODD2_T(t3)
        ODD2_P("odd2 border warp %d: load %lld  solve %lld\n", warp, t0 - t_entry, t3 - t0)
    }
}

// One CTA = (surviving row i = e * 2s, task D | E, strip of B / nstrip rows).
//   task 0:  D_i[strip, :] -= Bm_{i-s}^T Bm_{i-s} + A_{i+s}^T A_{i+s},  f_i[strip] -= Bm_{i-s}^T g_{i-s} + A_{i+s}^T g_{i+s}
//   task 1:  E(i, i-2s)[strip, :] = -Bm_{i-s}^T A_{i-s}
template <int B>
__global__ void __launch_bounds__(256) bcr_even2_kernel(BcrView R, int s, int nstrip, int raw, const double* __restrict__ Tb,
                                                         const double* __restrict__ fa, const double* __restrict__ fb) {
    extern __shared__ __align__(16) double smem_bcr[];
    const int tid = threadIdx.x;
    const int strip = blockIdx.x % nstrip, task = (blockIdx.x / nstrip) & 1, e = blockIdx.x / (2 * nstrip);
    const int i = e * 2 * s;
    const bool has_l = i - s >= 0, has_r = i + s < R.N, has_ll = i - 2 * s >= 0;
    constexpr long long bb = (long long)B * B;
    if (*R.fail) return;
    if (task == 1 && !has_ll) return;
    double* X = smem_bcr;   // task 0: Bm of the left odd neighbour   task 1: the same
    double* Y = X + B * B;  // task 0: A of the right odd neighbour   task 1: A of the left odd neighbour
    double* gx = Y + B * B;
    double* gy = gx + B;
    const bool use_x = has_l, use_y = task == 0 ? has_r : has_l;
    {
        const double2* Xs = reinterpret_cast<const double2*>(R.Bm + (long long)(i - s) * bb);
        const double2* Ys = reinterpret_cast<const double2*>(task == 0 ? R.A + (long long)(i + s) * bb : R.A + (long long)(i - s) * bb);
        double2* X2 = reinterpret_cast<double2*>(X);
        double2* Y2 = reinterpret_cast<double2*>(Y);
        const double2 z = make_double2(0.0, 0.0);
        for (int idx = tid; idx < B * B / 2; idx += 256) {
            X2[idx] = use_x ? Xs[idx] : z;
            Y2[idx] = use_y ? Ys[idx] : z;
        }
        if (task == 0)
            for (int r = tid; r < B; r += 256) {
                gx[r] = has_l ? R.g[(long long)(i - s) * B + r] : 0.0;
                gy[r] = has_r ? R.g[(long long)(i + s) * B + r] : 0.0;
            }
    }
    __syncthreads();
    const int rs = B / nstrip, row0 = strip * rs;  // rs is a multiple of 6
    constexpr int TC = B / 3;
    const int ntile = (rs / 2) * TC;
    for (int t = tid; t < ntile; t += 256) {
        const int r0 = row0 + 2 * (t / TC), c0 = 3 * (t % TC);
        double acc[2][3] = {{0, 0, 0}, {0, 0, 0}};
        if (task == 0) {
#pragma unroll 3
            for (int k = 0; k < B; ++k) {
                const double* xr = X + k * B;
                const double* yr = Y + k * B;
                const double2 ax = *reinterpret_cast<const double2*>(xr + r0);
                const double2 ay = *reinterpret_cast<const double2*>(yr + r0);
                const double bx0 = xr[c0], bx1 = xr[c0 + 1], bx2 = xr[c0 + 2];
                const double by0 = yr[c0], by1 = yr[c0 + 1], by2 = yr[c0 + 2];
                acc[0][0] += ax.x * bx0 + ay.x * by0;
                acc[0][1] += ax.x * bx1 + ay.x * by1;
                acc[0][2] += ax.x * bx2 + ay.x * by2;
                acc[1][0] += ax.y * bx0 + ay.y * by0;
                acc[1][1] += ax.y * bx1 + ay.y * by1;
                acc[1][2] += ax.y * bx2 + ay.y * by2;
            }
            double* Di = R.D + i * bb;
#pragma unroll
            for (int u = 0; u < 2; ++u)
#pragma unroll
                for (int w = 0; w < 3; ++w) {
                    const int idx = (r0 + u) * B + c0 + w;
                    double d = Di[idx];
                    if (raw) d -= Tb[i * bb + idx];
                    Di[idx] = d - acc[u][w];
                }
        } else {
#pragma unroll 3
            for (int k = 0; k < B; ++k) {
                const double* xr = X + k * B;
                const double* yr = Y + k * B;
                const double2 ax = *reinterpret_cast<const double2*>(xr + r0);
                const double by0 = yr[c0], by1 = yr[c0 + 1], by2 = yr[c0 + 2];
                acc[0][0] += ax.x * by0;
                acc[0][1] += ax.x * by1;
                acc[0][2] += ax.x * by2;
                acc[1][0] += ax.y * by0;
                acc[1][1] += ax.y * by1;
                acc[1][2] += ax.y * by2;
            }
            // row-major when row i is eliminated at the next level (stride 2s), transposed otherwise
            const bool odd_next = ((i / (2 * s)) & 1) != 0;
            double* Ei = R.E + i * bb;
#pragma unroll
            for (int u = 0; u < 2; ++u)
#pragma unroll
                for (int w = 0; w < 3; ++w) {
                    const int rr = r0 + u, cc = c0 + w;
                    Ei[odd_next ? rr * B + cc : cc * B + rr] = -acc[u][w];
                }
        }
    }
    if (task == 0) {
        for (int r = row0 + tid; r < row0 + rs; r += 256) {
            double ff = 0.0;
            for (int k = 0; k < B; ++k) ff += X[k * B + r] * gx[k] + Y[k * B + r] * gy[k];
            double f0 = raw ? fa[(long long)i * B + r] - fb[(long long)i * B + r] : R.f[(long long)i * B + r];
            R.f[(long long)i * B + r] = f0 - ff;
        }
    }
}

template <int B>
static int bcr2_solve(cudaStream_t st, const BandView& V, const BandScratch& K) {
    BcrView R;
    R.N = V.P - 1;
    R.b = B;
    const long long bb = (long long)B * B;
    R.D = V.Ta;
    R.E = V.Ca;
    R.f = K.rhs2;
    R.A = K.T2;
    R.Bm = K.T2 + R.N * bb;
    R.Li = K.L2;
    R.g = K.X2;
    R.y = K.y2;
    R.fail = V.fail;
    const size_t smem_even = sizeof(double) * (2 * size_t(bb) + 2 * B);
    CSLAM_CUDA(cudaFuncSetAttribute(bcr_even2_kernel<B>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem_even)));
    int launched = 0;
    int levels[32], nl = 0;
    int raw = 1;
    constexpr int W = B / 6;
    for (int s = 1; s < R.N; s *= 2) {
        const int n_odd = (R.N - s - 1) / (2 * s) + 1, n_even = (R.N - 1) / (2 * s) + 1;
        bcr_odd2_kernel<B><<<n_odd, Odd2<B>::THREADS, 0, st>>>(R, s, 2 * s, s, raw, V.Tb, V.fa, V.fb);
        int nstrip = 1;
        for (int d = 1; d <= W; ++d)
            if (W % d == 0 && n_even * 2 * d <= 444) nstrip = d;
        bcr_even2_kernel<B><<<n_even * 2 * nstrip, 256, smem_even, st>>>(R, s, nstrip, raw, V.Tb, V.fa, V.fb);
        raw = 0;
        launched += 2;
        levels[nl++] = s;
    }
    bcr_odd2_kernel<B><<<1, Odd2<B>::THREADS, 0, st>>>(R, 0, 1, 0, raw, V.Tb, V.fa, V.fb);  // the last row standing
    const size_t smem_bs = sizeof(double) * (3 * size_t(bb) + 3 * B);
    CSLAM_CUDA(cudaFuncSetAttribute(bcr_backsub_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem_bs)));
    bcr_backsub_kernel<<<1, 256, smem_bs, st>>>(R, 0, 1, 0);
    launched += 2;
    for (int l = nl - 1; l >= 0; --l) {
        const int s = levels[l], n_odd = (R.N - s - 1) / (2 * s) + 1;
        bcr_backsub_kernel<<<n_odd, 256, smem_bs, st>>>(R, s, 2 * s, s);
        ++launched;
    }
    CSLAM_CUDA(cudaGetLastError());
    return launched;
}

static int bcr_solve(cudaStream_t st, const BandView& V, const BandScratch& K, int W) {
    BcrView R;
    R.N = V.P - 1;
    R.b = 6 * W;
    const long long bb = (long long)R.b * R.b;
    R.D = V.Ta;
    R.E = V.Ca;
    R.f = K.rhs2;
    R.A = K.T2;
    R.Bm = K.T2 + R.N * bb;
    R.Li = K.L2;
    R.g = K.X2;
    R.y = K.y2;
    R.fail = V.fail;
    const size_t smem_odd = sizeof(double) * size_t(R.b) * (4 * R.b + 1);
    const size_t smem_even = sizeof(double) * (3 * size_t(bb) + 2 * R.b);
    // per device and size dependent: set on every solve (a host-side call of ~1 us; no process-wide cache)
    CSLAM_CUDA(cudaFuncSetAttribute(bcr_odd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem_odd)));
    CSLAM_CUDA(cudaFuncSetAttribute(bcr_even_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem_even)));
    int launched = 0;
    const long long init_n = R.N * bb + (long long)R.N * R.b;
    bcr_init_kernel<<<int(std::min<long long>((init_n + 255) / 256, 4 * 148)), 256, 0, st>>>(R, V.Tb, V.fa, V.fb);
    ++launched;
    int levels[32], nl = 0;
    for (int s = 1; s < R.N; s *= 2) {
        const int n_odd = (R.N - s - 1) / (2 * s) + 1, n_even = (R.N - 1) / (2 * s) + 1;
        bcr_odd_kernel<<<n_odd, BCR_T, smem_odd, st>>>(R, s, 2 * s, s);
        bcr_even_kernel<<<n_even, BCR_T, smem_even, st>>>(R, s);
        launched += 2;
        levels[nl++] = s;
    }
    bcr_odd_kernel<<<1, BCR_T, smem_odd, st>>>(R, 0, 1, 0);  // the last row standing
    const size_t smem_bs = sizeof(double) * (3 * size_t(bb) + 3 * R.b);
    CSLAM_CUDA(cudaFuncSetAttribute(bcr_backsub_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem_bs)));
    bcr_backsub_kernel<<<1, 256, smem_bs, st>>>(R, 0, 1, 0);
    launched += 2;
    for (int l = nl - 1; l >= 0; --l) {
        const int s = levels[l], n_odd = (R.N - s - 1) / (2 * s) + 1;
        bcr_backsub_kernel<<<n_odd, 256, smem_bs, st>>>(R, s, 2 * s, s);
        ++launched;
    }
    CSLAM_CUDA(cudaGetLastError());
    return launched;
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
    // second-generation leaf kernel (register-resident window columns); CSLAM_BAND_LEAF1=1 keeps the first one (A/B runs)
    static const bool leaf1 = [] {
        const char* e = std::getenv("CSLAM_BAND_LEAF1");
        return e && std::atoi(e) != 0;
    }();
    if constexpr (W <= 9) {
        if (leaf1 || V.band_idx == nullptr)
            run_leaf<W, true>(s, V);
        else
            run_leaf2<W>(s, V);
    } else {
        run_leaf<W, true>(s, V);
    }
    if (V.sep_solver == 2 || V.sep_solver == 3) {
        // W = 12 (B = 72) keeps the first-generation kernels: its unrolled columns do not fit the register file
        int launched;
        if constexpr (W <= 9)
            launched = V.sep_solver == 2 ? bcr2_solve<6 * W>(s, V, K) : bcr_solve(s, V, K, W);
        else
            launched = bcr_solve(s, V, K, W);
        run_backsub<W, true>(s, V, K.y2);
        return 2 + launched;
    }
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
