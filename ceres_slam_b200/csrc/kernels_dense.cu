// K3d — dense FP64 Cholesky of the reduced camera system, for the exact (SPARSE_SCHUR-equivalent)
// solve when S is NOT a narrow band: loop closures put blocks far from the diagonal
// (/root/reference/scripts/ba_all_sims.sh:8-13 runs closed tracks, dataset_vo.cpp:118-121 solves them
// as one batch), long tracks widen the band past what kernels_band.cu takes.  north_star (3): "dense
// FP64 Cholesky using DMMA tensor cores when the reduced camera matrix is dense enough".
//
// Storage: the lower triangle, column-major, padded to a multiple of the panel width with an identity
// diagonal; the right-hand side rides along as ONE EXTRA ROW (index n_pad), so the forward substitution
// y = L^-1 b falls out of the factorisation itself.  Blocked right-looking, 48 columns per panel:
//
//   dense_panel_kernel   one CTA per 192 rows of the panel.  Every CTA factors the 48 x 48 diagonal
//                        block itself (two warps own its rows in registers and run the scalar pivot
//                        chain of chol_chain.cuh — the same code as the cyclic-reduction kernel), and
//                        every other thread owns one ROW of the panel below it (48 doubles in
//                        registers) and solves  l L11^T = a  right-looking, trailing the factor warps
//                        by a pivot.  Redundant factorisation instead of a grid-wide hand-over: the
//                        chain is latency (~7 us), the SMs would idle anyway.
//   dense_syrk_kernel    trailing update C -= L21 L21^T on the FP64 tensor cores (mma.sync m8n8k4,
//                        SASS DMMA.8x8x4): 64 x 64 tiles of the lower triangle, K = 48, the two panel
//                        slabs staged in shared memory; the 8 x 8 accumulator tile is computed
//                        TRANSPOSED so that a lane's two values are consecutive rows of one column,
//                        i.e. one 16-byte read-modify-write of the column-major matrix.
//   dense_backsolve_kernel  L^T x = y from the last panel up: each CTA forms L11^-T times the panel's entries
//                        (the panel kernel solved 48 unit-vector rows along with the others, so the inverse
//                        of the diagonal block is there: a product instead of a 48-step triangular solve) and
//                        subtracts the panel's contribution from its 256 columns to the left.
//
// Work is n^3 / 3 flops (n = 6 x free poses), all of the O(n^3) part inside the DMMA kernel; it is
// chosen over PCG only when the system is small or dense (engine.cu plan_dense_solver).
#include <algorithm>

#include "chol_chain.cuh"
#include "kernels.cuh"

namespace cslam {

namespace {

constexpr int DNB = 48;                     // panel width
constexpr int DPW = 6;                      // row warps per panel CTA
constexpr int DPT = 32 * (2 + DPW);         // threads of a panel CTA
constexpr int DPR = 32 * DPW;               // panel rows per CTA
constexpr int DSL = 68;                     // row stride of a staged panel slab ([k][64 rows]); 4 mod 16: the fragment load of lane
                                            // (g, q) reads word q * 68 + g (+ tile), distinct banks over a half-warp (72 was 2-way)

// One CTA per block row of the upper block-CSR; block 0 also writes the right-hand-side row and the
// identity padding.
__global__ void dense_fill_kernel(DenseView V) {
    const int a = blockIdx.x;
    const long long ld = V.ld;
    for (int idx = threadIdx.x; idx < (V.rowptr[a + 1] - V.rowptr[a]) * 36; idx += blockDim.x) {
        const int e = V.rowptr[a] + idx / 36, r = (idx % 36) / 6, c = idx % 6;
        const int b = V.col[e];
        V.A[(6ll * a + r) * ld + 6 * b + c] = V.S[36ll * e + 6 * r + c];  // A[6b+c][6a+r] = S_ab[r][c]
    }
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j <= V.n_pad; j += gridDim.x * blockDim.x) {
        if (j < V.n)
            V.A[j * ld + V.n_pad] = V.rhs[j];
        else if (j < V.n_pad)
            V.A[j * ld + j] = 1.0;
    }
}

__global__ void __launch_bounds__(DPT, 1) dense_panel_kernel(DenseView V, int j0) {
    constexpr int B = DNB;
    __shared__ __align__(16) double Lt2[(B + 2) * B];
    __shared__ double sInv[B + 2];
    __shared__ __align__(8) uint64_t done0[B + 1];
    __shared__ __align__(8) uint64_t done1[B];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long ld = V.ld;
    if (*V.fail) return;
    for (int j = tid; j < B + 1; j += DPT) {
        if (j < B) {
            Lt2[j] = 0.0;
            mbar_init(&done1[j], 32);
        }
        mbar_init(&done0[j], 32);
    }
    fence_mbar_init();
    __syncthreads();
    double* Ap = V.A + j0 * ld;  // column j0
    if (warp == 0) {
        const int r = lane;
        double b[33];
        b[0] = 0.0;
#pragma unroll
        for (int c = 0; c < 32; ++c) b[c + 1] = Ap[c * ld + j0 + r];
        double diag = Ap[r * ld + j0 + r];
        double nid = rsqrt_nr(diag);
        bool bad = false;
        if (r == 0) {
            bad = !pivot_ok(diag);
            sInv[0] = nid;
        }
        __syncwarp();
        mbar_arrive(done0);
        double lprev = 0.0;
        odd2_factor_phase<B, 32, 33, true>(b, diag, lprev, nid, 0, 16, r, 0, true, Lt2, sInv, nullptr, done0 + 1, bad);
        odd2_factor_phase<B, 32, 18, true>(b, diag, lprev, nid, 16, 31, r, 0, true, Lt2, sInv, nullptr, done0 + 1, bad);
        __syncwarp();
        mbar_arrive(done0 + 31);
        if (__any_sync(0xffffffffu, bad) && lane == 0) *V.fail = 1;
    } else if (warp == 1) {
        const int r = 32 + lane;
        const bool act = r < B;
        const int rr = act ? r : 0;
        double b[B + 1];
        b[0] = 0.0;
#pragma unroll
        for (int c = 0; c < B; ++c) b[c + 1] = Ap[c * ld + j0 + rr];
        double diag = Ap[rr * ld + j0 + rr];
        if (!act) diag = 1.0;
        double nid = 1.0, lprev = 0.0;
        bool bad = false;
        odd2_factor_phase<B, B, B + 1, false>(b, diag, lprev, nid, 0, 16, r, 32, act, Lt2, sInv, done0, done1, bad);
        odd2_factor_phase<B, B, B + 1 - 15, false>(b, diag, lprev, nid, 16, 32, r, 32, act, Lt2, sInv, done0, done1, bad);
        odd2_factor_phase<B, B, B + 1 - 31, true>(b, diag, lprev, nid, 32, B, r, 32, act, Lt2, sInv, nullptr, done1, bad);
        __syncwarp();
        mbar_arrive(done1 + B - 1);
        if (__any_sync(0xffffffffu, bad) && lane == 0) *V.fail = 1;
    } else {
        // one row of the panel below the diagonal block per thread (the last one is the right-hand side); 48 more
        // "rows" are the unit vectors: their solutions are the rows of L11^-T, which the back-substitution applies as
        // a product instead of a 48-step triangular solve.  Ld[panel][c][i] = (L11^-T)[i][c]
        const int rows = V.n_pad + 1 - (j0 + B);
        const int rl = blockIdx.x * DPR + (warp - 2) * 32 + lane;
        const bool valid = rl < rows + B;
        double* rowp = rl < rows ? Ap + j0 + B + rl : V.Ldiag + (long long)(j0 / B) * B * B + (rl - rows);
        const int stride = rl < rows ? int(ld) : B;
        double x[B];
#pragma unroll
        for (int c = 0; c < B; ++c) x[c] = !valid ? 0.0 : rl < rows ? rowp[(long long)c * stride] : (c == rl - rows ? 1.0 : 0.0);
        double* outp = valid ? rowp : nullptr;
        odd2_border_phase<B, B>(x, 0, 16, Lt2, sInv, done1, outp, stride);
        odd2_border_phase<B, B - 16>(x, 16, 32, Lt2, sInv, done1, outp, stride);
        odd2_border_phase<B, B - 32>(x, 32, B, Lt2, sInv, done1, outp, stride);
    }
}

__device__ __forceinline__ void dmma_m8n8k4(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// Trailing update behind panel j0: rows / columns t0 = j0 + 48 .. n_pad (the last row is the right-hand side).
// Four CTAs per SM (64 registers): more of a launch's latency-bound tiles are resident at once.
__global__ void __launch_bounds__(256, 4) dense_syrk_kernel(DenseView V, int j0) {
    extern __shared__ __align__(16) double smem_syrk[];
    double* Lr = smem_syrk;           // Lr[k][i]: panel rows of the tile's row range
    double* Lc = Lr + DNB * DSL;      // panel rows of the tile's column range
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long ld = V.ld;
    if (*V.fail) return;
    const int t0 = j0 + DNB, m = V.n_pad + 1 - t0;
    // linear CTA index -> tile (I, J), I >= J
    int I = int((sqrt(8.0 * blockIdx.x + 1.0) - 1.0) * 0.5);
    while ((long long)(I + 1) * (I + 2) / 2 <= blockIdx.x) ++I;
    while ((long long)I * (I + 1) / 2 > blockIdx.x) --I;
    const int J = blockIdx.x - I * (I + 1) / 2;
    const double* Ap = V.A + j0 * ld + t0;
    {
        // every thread stages 12 entries of each slab: row i = tid & 63, panel columns k = (tid >> 6) + 4 t.  All 24
        // loads are issued before the first store (a loop with a store per load serialises on the load latency)
        const int i = tid & 63, kq = tid >> 6;
        const bool okr = 64 * I + i < m, okc = 64 * J + i < m;
        const double* pr = Ap + kq * ld + 64 * I + i;
        const double* pc = Ap + kq * ld + 64 * J + i;
        double vr[12], vc[12];
#pragma unroll
        for (int t = 0; t < 12; ++t) {
            vr[t] = okr ? pr[4 * t * ld] : 0.0;
            vc[t] = okc ? pc[4 * t * ld] : 0.0;
        }
#pragma unroll
        for (int t = 0; t < 12; ++t) {
            Lr[(kq + 4 * t) * DSL + i] = vr[t];
            Lc[(kq + 4 * t) * DSL + i] = vc[t];
        }
    }
    __syncthreads();
    // warp = column tile jt of the 64 x 64 block; the transposed accumulator D[m][n] = C[i = 8 it + n][j = 8 jt + m]
    const int jt = warp, g = lane >> 2, q = lane & 3;
    double acc[8][2];
#pragma unroll
    for (int it = 0; it < 8; ++it) acc[it][0] = acc[it][1] = 0.0;
#pragma unroll 4
    for (int k0 = 0; k0 < DNB; k0 += 4) {
        const double a = Lc[(k0 + q) * DSL + 8 * jt + g];
#pragma unroll
        for (int it = 0; it < 8; ++it) {
            const double b = Lr[(k0 + q) * DSL + 8 * it + g];
            dmma_m8n8k4(acc[it][0], acc[it][1], a, b);
        }
    }
    const int j = 64 * J + 8 * jt + g;
    if (j < m) {
        double* Cc = V.A + (long long)(t0 + j) * ld + t0;
        // read-modify-write of the lane's eight row pairs, four at a time: their loads first, then their stores
#pragma unroll
        for (int h0 = 0; h0 < 8; h0 += 4) {
            double2 cv[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = 64 * I + 8 * (h0 + u) + 2 * q;  // two consecutive rows of column j
                cv[u] = make_double2(0.0, 0.0);
                if (i + 1 < m)
                    cv[u] = *reinterpret_cast<const double2*>(Cc + i);
                else if (i < m)
                    cv[u].x = Cc[i];
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = 64 * I + 8 * (h0 + u) + 2 * q;
                cv[u].x -= acc[h0 + u][0];
                cv[u].y -= acc[h0 + u][1];
                if (i + 1 < m)
                    *reinterpret_cast<double2*>(Cc + i) = cv[u];
                else if (i < m)
                    Cc[i] = cv[u].x;
            }
        }
    }
}

// Panel j0 of L^T x = y.  xw: work vector (entries right of the panel are final, the panel's own have every
// later panel's contribution subtracted already).  Every CTA forms x_panel = L11^-T xw_panel (the inverse the panel
// kernel left behind); CTA c then updates columns [256 c, 256 c + 256) left of the panel.
__global__ void __launch_bounds__(256) dense_backsolve_kernel(DenseView V, int j0, double* xw, double* y) {
    constexpr int B = DNB;
    __shared__ double Ls[B * (B + 1)];   // Ls[c][i] = (L11^-T)[i][c] (zero for c < i), row stride B + 1
    __shared__ double ts[B], xs[B];
    const int tid = threadIdx.x;
    const long long ld = V.ld;
    if (*V.fail) return;
    const double* Li = V.Ldiag + (long long)(j0 / B) * B * B;
    {
        double lv[9];
#pragma unroll
        for (int t = 0; t < 9; ++t) lv[t] = Li[tid + 256 * t];   // 48 * 48 = 9 * 256
        if (tid < B) ts[tid] = xw[j0 + tid];
#pragma unroll
        for (int t = 0; t < 9; ++t) {
            const int idx = tid + 256 * t;
            Ls[(idx / B) * (B + 1) + idx % B] = lv[t];
        }
    }
    __syncthreads();
    if (tid < 4 * B) {
        // four lanes per row i: c = q, q + 4, ...
        const int i = tid >> 2, q = tid & 3;
        double a = 0.0;
#pragma unroll
        for (int c = 0; c < B; c += 4) a = fma(Ls[(c + q) * (B + 1) + i], ts[c + q], a);
        a += __shfl_xor_sync(0xffffffffu, a, 1);
        a += __shfl_xor_sync(0xffffffffu, a, 2);
        if (q == 0) xs[i] = a;
    }
    __syncthreads();
    if (blockIdx.x == 0 && tid < B && j0 + tid < V.n) y[j0 + tid] = xs[tid];
    const int j = blockIdx.x * 256 + tid;
    if (j < j0) {
        const double2* Lj = reinterpret_cast<const double2*>(V.A + j * ld + j0);  // L[j0 .. j0 + 47][j]: contiguous
        double2 l[B / 2];
#pragma unroll
        for (int k = 0; k < B / 2; ++k) l[k] = Lj[k];
        double d0 = 0.0, d1 = 0.0;
#pragma unroll
        for (int k = 0; k < B / 2; ++k) {
            d0 = fma(l[k].x, xs[2 * k], d0);
            d1 = fma(l[k].y, xs[2 * k + 1], d1);
        }
        xw[j] -= d0 + d1;
    }
}

__global__ void dense_status_kernel(const int* fail, double* ps) {
    ps[PS_ITERS] = 1.0;
    ps[PS_FAIL] = *fail ? 2.0 : 0.0;
}

}  // namespace

int dense_panel_width() { return DNB; }

void launch_dense_factor(cudaStream_t s, const DenseView& V, double* xw) {
    int launched = 0;
    constexpr size_t smem_syrk = sizeof(double) * 2 * DNB * DSL;
    CSLAM_CUDA(cudaFuncSetAttribute(dense_syrk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem_syrk)));
    // (four 52 KB CTAs per SM need the large carve-out; the default left room for two)
    CSLAM_CUDA(cudaFuncSetAttribute(dense_syrk_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    for (int j0 = 0; j0 < V.n_pad; j0 += DNB) {
        const int rows = V.n_pad + 1 - (j0 + DNB);  // rows below the diagonal block (>= 1: the right-hand side)
        dense_panel_kernel<<<(rows + DNB + DPR - 1) / DPR, DPT, 0, s>>>(V, j0);   // + the 48 unit-vector rows
        ++launched;
        if (j0 + DNB < V.n_pad) {
            const long long T = (rows + 63) / 64;
            dense_syrk_kernel<<<int(T * (T + 1) / 2), 256, smem_syrk, s>>>(V, j0);
            ++launched;
        }
    }
    // y = L^-1 b is the extra row; gather it into the work vector (a strided copy)
    CSLAM_CUDA(cudaMemcpy2DAsync(xw, sizeof(double), V.A + V.n_pad, sizeof(double) * size_t(V.ld), sizeof(double), size_t(V.n_pad),
                                 cudaMemcpyDeviceToDevice, s));
    for (int j0 = V.n_pad - DNB; j0 >= 0; j0 -= DNB) {
        dense_backsolve_kernel<<<std::max(1, (j0 + 255) / 256), 256, 0, s>>>(V, j0, xw, V.y);
        ++launched;
    }
    CSLAM_CUDA(cudaGetLastError());
    g_kernel_launches.fetch_add(launched, std::memory_order_relaxed);
}

void launch_dense_solve(cudaStream_t s, const DenseView& V, double* xw, double* ps) {
    CSLAM_CUDA(cudaMemsetAsync(V.fail, 0, sizeof(int), s));
    CSLAM_CUDA(cudaMemsetAsync(V.A, 0, sizeof(double) * size_t(V.ld) * size_t(V.n_pad + 1), s));
    dense_fill_kernel<<<V.n / 6, 128, 0, s>>>(V);
    launch_dense_factor(s, V, xw);
    dense_status_kernel<<<1, 1, 0, s>>>(V.fail, ps);
    CSLAM_CUDA(cudaGetLastError());
    g_kernel_launches.fetch_add(2, std::memory_order_relaxed);
}

}  // namespace cslam
