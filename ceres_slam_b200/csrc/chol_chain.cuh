// Device building blocks of a dense B x B FP64 Cholesky organised around its pivot chain, shared by the
// cyclic-reduction kernel of the banded solver (kernels_band.cu) and the dense reduced solve
// (kernels_dense.cu).  See the notes in front of bcr_odd2_kernel for the design.
#pragma once
#include <cstdint>

#include "common.cuh"

namespace cslam {

// Reciprocal square root without the library's special-case branch: the approximation instruction
// (MUFU.RSQ64H, ~2^-22) and one third-order correction, the same five FP64 operations rsqrt() performs on
// its fast path.  Branch-free matters more than the count: the compiler can then interleave independent
// work with this dependent sequence, which is what every pivot chain in this file waits for.  Callers
// reject pivots outside [1e-290, 1e290] (where flush-to-zero / overflow handling would matter).
__device__ __forceinline__ double rsqrt_nr(double d) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
    const double e = fma(-(y * y), d, 1.0);
    return fma(fma(e, 0.375, 0.5), y * e, y);
}
__device__ __forceinline__ bool pivot_ok(double d) { return d > 1e-290 && d < 1e290; }

template <int B>
struct Odd2 {
    static constexpr int FW = (B + 31) / 32;        // factor warps (B <= 64)
    static constexpr int R0 = B < 32 ? B : 32;      // rows of warp 0
    static constexpr int NCOL = 3 * B + 1;          // border columns
    static constexpr int BW = (NCOL + 31) / 32;     // border warps
    static constexpr int THREADS = 32 * (FW + BW);
};

// Pivots [k_lo, k_hi) of one factor warp; lane = row `r` of D_i, NC = number of columns its rows can
// reach (32 for warp 0, B for warp 1).  Registers: b[j] = A[r][k - 1 + j], not yet updated by pivot
// k-1 — the rank-1 update writes b[j] <- b[j+1] - l L, i.e. the row slides down by one register per
// pivot, so every register index is static while k is a run-time loop variable: the body is ~100
// instructions and stays in the instruction cache (a fully unrolled version of this kernel spent
// most of its time fetching instructions).
//   CHAIN: the pivot row belongs to this warp.  The chain per pivot is  shuffle 1/L_kk and L_{k,k-1}
//     from the pivot lane -> l = a / L_kk -> diag -= l^2 -> rsqrt;  the rest of the previous pivot's
//     update is issued behind the rsqrt and hides in its latency.  Nothing on the chain touches
//     shared memory or a CTA barrier.
//   !CHAIN (warp 1 while the pivot is still among warp 0's rows): waits for warp 0's column.
// After publishing column k (Lt2 row k+1; 1/L of the next pivot in sInv) every lane arrives on
// `done[k]`: consumers (warp 1, the border warps) wait on these single-use mbarriers — a hardware
// wait, no polling traffic that would sit in front of the chain's own shared-memory accesses.
template <int B, int NC, int WID, bool CHAIN>
__device__ __forceinline__ void odd2_factor_phase(double (&b)[NC + 1], double& diag, double& lprev, double& nid, int k_lo,
                                                  int k_hi, int r, int lane_base, bool act, double* Lt2, double* sInv,
                                                  uint64_t* wait_on, uint64_t* done, bool& bad) {
    for (int k = k_lo; k < k_hi; ++k) {
        const double* Lp = Lt2 + k * B;  // column k-1 of L from row k on (row 0 of Lt2 is zeros)
        double id, lpo;
        if (CHAIN) {
            id = __shfl_sync(0xffffffffu, nid, k - lane_base);
            lpo = __shfl_sync(0xffffffffu, lprev, k - lane_base);
        } else {
            // warp 0 has published column k-1 and 1 / L_kk — and column k as well, so that this warp's
            // own arrival on done[k] tells the border that column k is complete
            mbar_wait(wait_on + (k + 1 < Odd2<B>::R0 ? k + 1 : Odd2<B>::R0 - 1), 0);
            id = sInv[k];
            lpo = Lp[0];
        }
        const double a0 = b[1] - lprev * lpo;
        const double l = a0 * id;
        if (r > k) diag -= l * l;
        if (CHAIN || k + 1 >= lane_base) nid = rsqrt_nr(diag);  // the next pivot's reciprocal root (row k+1 owns it)
        b[0] = a0;
#pragma unroll
        for (int j = 1; j + 1 < WID; ++j) b[j] = b[j + 1] - lprev * Lp[j];
        // column k-1 (stored one pivot ago, long complete): the release of this arrive costs nothing now
        if (k > 0) mbar_arrive(done + k - 1);
        if (act && r > k) Lt2[(k + 1) * B + (r - k - 1)] = l;
        if (act && r == k + 1) sInv[k + 1] = nid;
        bad |= act && (r == k + 1) && !pivot_ok(diag);
        __syncwarp();
        lprev = l;
    }
}

// Border columns: x[j] = X[k + j]; pivot k finishes x[0], writes it out and slides the rest.
template <int B, int WID>
__device__ __forceinline__ void odd2_border_phase(double (&x)[B], int k_lo, int k_hi, const double* Lt2, const double* sInv,
                                                  uint64_t* col_done, double* outp, int ostride) {
    for (int k = k_lo; k < k_hi; ++k) {
        // a completed barrier still costs ~90 cycles to test: look two columns ahead, every third pivot
        if ((k - k_lo) % 3 == 0) mbar_wait(col_done + (k + 2 < B ? k + 2 : B - 1), 0);
        const double xk = x[0] * sInv[k];
        if (outp) outp[(long long)k * ostride] = xk;
        const double2* Lp = reinterpret_cast<const double2*>(Lt2 + (k + 1) * B);  // column k of L from row k+1 on
#pragma unroll
        for (int j = 0; j + 1 < WID; j += 2) {
            const double2 lv = Lp[j >> 1];
            x[j] = x[j + 1] - lv.x * xk;
            if (j + 2 < WID) x[j + 1] = x[j + 2] - lv.y * xk;
        }
    }
}

}  // namespace cslam
