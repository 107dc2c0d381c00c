// K3c — exact solve of a reduced camera system that is "a band plus a little": tracks longer than the banded
// direct solver's window or a few re-observations put blocks outside the band, and the system is too large for
// the dense factorisation.  (A realistic stereo front end produces exactly that: most tracks are short, a few
// run for 30 frames.)  Conjugate gradients preconditioned with the EXACT banded solve of
//     M = sum over the landmarks whose cameras lie within the window  (U_j - W_j V_j^-1 W_j^T)  + damping.
// Dropping whole landmarks instead of far blocks keeps M positive definite (S - M is a sum of positive
// semi-definite per-landmark Schur terms) AND keeps the near-null space of S — the slowly varying, gauge-like
// modes every landmark term annihilates — out of S - M; truncating blocks and compensating the diagonal was
// tried first and converges at 0.93 per iteration because it stiffens exactly those modes.  The Schur kernels
// accumulate the two landmark classes into two buffers (engine.cu schur_pass), M is the first one in dense band
// storage, S their sum.  Measured on 5 k poses x 100 landmarks / frame, track lengths 2..30 with drop-outs:
// 1.55 s (block-Jacobi PCG to 1e-15) -> 27 ms per solve.
#include <algorithm>

#include "kernels.cuh"

namespace cslam {

namespace {

// M in dense band storage from the short-track part of the system BEFORE finalize: off-diagonal blocks as they are,
// diagonal block = sym(Schur part (upper triangle) + U_short) + clamp(diag(U_short + U_long)) / radius — what
// finalize_kernel makes of the whole system, restricted to the landmarks inside the window.
__global__ void bpc_build_kernel(int n, int W, const int* __restrict__ rowptr, const int* __restrict__ col,
                                 const double* __restrict__ S1, const double* __restrict__ U1, const double* __restrict__ U2,
                                 LmDiag dg, double* __restrict__ Sband) {
    const int a = blockIdx.x;
    double* dst = Sband + (long long)a * (W + 1) * 36;
    for (int i = threadIdx.x; i < (W + 1) * 36; i += blockDim.x) dst[i] = 0.0;
    __syncthreads();
    for (int idx = threadIdx.x; idx < (rowptr[a + 1] - rowptr[a]) * 36; idx += blockDim.x) {
        const int e = rowptr[a] + idx / 36, d = col[e] - a, rc = idx % 36;
        if (d == 0) {
            const int r = rc / 6, c = rc % 6, lo = r < c ? r : c, hi = r < c ? c : r;   // upper triangles -> full block
            double v = S1[36ll * e + 6 * lo + hi] + U1[36ll * a + 6 * lo + hi];
            if (r == c) {
                const double uu = U1[36ll * a + 7 * r] + U2[36ll * a + 7 * r];
                v += fmin(fmax(uu, dg.min_diag), dg.max_diag) * dg.inv_radius;
            }
            dst[rc] = v;
        } else if (d <= W) {
            dst[d * 36 + rc] = S1[36ll * e + rc];
        }
    }
}
__global__ void bpc_add_kernel(long long n, const double* __restrict__ a, double* __restrict__ b) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) b[i] += a[i];
}

// q = S p with the upper block-CSR (diagonal blocks full symmetric) and the mirrored lists; one warp per block row
__global__ void __launch_bounds__(256) bpc_spmv_kernel(int n, const int* __restrict__ rowptr, const int* __restrict__ col,
                                                       const int* __restrict__ ent_ptr, const int* __restrict__ ent_cb,
                                                       const double* __restrict__ S, const double* __restrict__ p,
                                                       double* __restrict__ q) {
    const int a = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (a >= n) return;
    // lane = (entry slot t, row r): 5 blocks in flight, 6 lanes each
    const int t = lane / 6, r = lane % 6;
    double acc = 0.0;
    if (t < 5) {
        for (int e = rowptr[a] + t; e < rowptr[a + 1]; e += 5) {
            const double* B = S + 36ll * e + 6 * r;
            const double* pb = p + 6ll * col[e];
#pragma unroll
            for (int c = 0; c < 6; ++c) acc += B[c] * pb[c];
        }
        for (int m = ent_ptr[a] + t; m < ent_ptr[a + 1]; m += 5) {
            const int src = ent_cb[2 * m], e = ent_cb[2 * m + 1];
            const double* B = S + 36ll * e;
            const double* pb = p + 6ll * src;
#pragma unroll
            for (int c = 0; c < 6; ++c) acc += B[6 * c + r] * pb[c];   // S_e^T
        }
    }
    // sum the five slots of each row r
    double tot = 0.0;
#pragma unroll
    for (int k = 0; k < 5; ++k) tot += __shfl_sync(0xffffffffu, acc, 6 * k + r < 30 ? 6 * k + r : 0);
    if (lane < 6) q[6ll * a + lane] = tot;
}

__global__ void bpc_xpby_kernel(long long n, const double* __restrict__ z, double beta, double* __restrict__ p) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        p[i] = beta == 0.0 ? z[i] : z[i] + beta * p[i];   // (the first direction: p is uninitialised memory, 0 * NaN = NaN)
}
__global__ void bpc_update_kernel(long long n, double alpha, const double* __restrict__ p, const double* __restrict__ q,
                                  double* __restrict__ x, double* __restrict__ r) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        x[i] += alpha * p[i];
        r[i] -= alpha * q[i];
    }
}

}  // namespace

void launch_bpc_build(cudaStream_t s, int n, int W, const int* rowptr, const int* col, const double* S1, const double* U1,
                      const double* U2, LmDiag dg, double* Sband) {
    bpc_build_kernel<<<n, 128, 0, s>>>(n, W, rowptr, col, S1, U1, U2, dg, Sband);
    CSLAM_CUDA(cudaGetLastError());
    g_kernel_launches.fetch_add(1, std::memory_order_relaxed);
}
void launch_bpc_add(cudaStream_t s, long long n, const double* a, double* b) {
    if (n <= 0) return;
    bpc_add_kernel<<<int(std::min<long long>((n + 255) / 256, 8 * 148)), 256, 0, s>>>(n, a, b);
    CSLAM_CUDA(cudaGetLastError());
    g_kernel_launches.fetch_add(1, std::memory_order_relaxed);
}
void launch_bpc_spmv(cudaStream_t s, int n, const int* rowptr, const int* col, const int* ent_ptr, const int* ent_cb,
                     const double* S, const double* p, double* q) {
    bpc_spmv_kernel<<<(n + 7) / 8, 256, 0, s>>>(n, rowptr, col, ent_ptr, ent_cb, S, p, q);
    CSLAM_CUDA(cudaGetLastError());
    g_kernel_launches.fetch_add(1, std::memory_order_relaxed);
}
void launch_bpc_xpby(cudaStream_t s, long long n, const double* z, double beta, double* p) {
    bpc_xpby_kernel<<<int(std::min<long long>((n + 255) / 256, 4 * 148)), 256, 0, s>>>(n, z, beta, p);
    CSLAM_CUDA(cudaGetLastError());
    g_kernel_launches.fetch_add(1, std::memory_order_relaxed);
}
void launch_bpc_update(cudaStream_t s, long long n, double alpha, const double* p, const double* q, double* x, double* r) {
    bpc_update_kernel<<<int(std::min<long long>((n + 255) / 256, 4 * 148)), 256, 0, s>>>(n, alpha, p, q, x, r);
    CSLAM_CUDA(cudaGetLastError());
    g_kernel_launches.fetch_add(1, std::memory_order_relaxed);
}

}  // namespace cslam
