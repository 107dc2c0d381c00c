// K3e — exact (SPARSE_SCHUR-equivalent) solve of a WIDE block-banded reduced camera system: half-bandwidth w of
// 13 .. 64 pose blocks.  That is what tracks longer than 13 frames produce (a real stereo front end: lengths 2 .. 30,
// /root/reference/tests/dataset_vo.cpp:118-121 solves the whole trajectory as one batch); kernels_band.cu keeps its
// window in registers and stops at w = 12, the dense factorisation (kernels_dense.cu) is O(n^3) and stops at 12 288
// unknowns.  Before this file such a system went to conjugate gradients preconditioned with the narrow-band solver
// (~60 banded solves per LM iteration).
//
// A band Cholesky is one dependent chain of n pivots (~150 ns each); the only parallelism is across the chain.  So
// the poses are cut into C chunks separated by C - 1 separators of w poses (two chunks never couple), and every chunk
// is factored as a BORDERED band
//
//        [ B_c   Y^T ]      B_c: the chunk's interior (band, LAPACK-style band storage: element (i, j) at A[j ld + i])
//        [ Y     D'  ]      border rows: [ left separator (6w) | right-hand side (1) | right separator (6w) ]
//
// with the dense solver's machinery, 48 columns per panel, ALL chunks in every launch (blockIdx.y = chunk):
//   wband_panel_kernel   the 48 x 48 diagonal block on the pivot-chain code of chol_chain.cuh, every other thread owns
//                        one row below it — band rows and border rows alike — and solves l L11^T = a in registers;
//                        48 more "rows" are the unit vectors: their solutions are L11^-T, which turns the
//                        back-substitution's triangular solve into a product;
//   wband_syrk_kernel    trailing update on the FP64 tensor cores (mma.sync m8n8k4 / DMMA.8x8x4) over the panel's
//                        local index space [band rows within reach | border rows]: band x band, border x band, and
//                        border x border, which accumulates the chunk's Schur complement D' on its separators.
// The separator system T = S_sep,sep + sum of the chunks' D' is block tridiagonal with blocks of 6w; it is small
// ((C - 1) 6w unknowns) and goes through the dense solver as it is.  Then L^T x = y - Y_sep^T x_sep per chunk: ONE
// launch, a CTA per chunk walks its panels from the last one up (48 x 48 product with L11^-T, then the update of the
// bwr columns to the left).  The right separator couples to the chunk's last w poses only, so its border rows are
// structurally zero — and skipped — until panel r_start.  C ~ sqrt(n / 7w) balances the chunk chains against the
// separator solve.
// Deterministic: every entry has one owner per launch, the separator assembly is a gather.
//
// scripts/wband_model.py is the numpy model of exactly this layout and loop structure.
#include <algorithm>

#include "chol_chain.cuh"
#include "kernels.cuh"

namespace cslam {

namespace {

constexpr int WNB = 48;                     // panel width (the dense solver's)
constexpr int WPW = 6;                      // row warps per panel CTA
constexpr int WPT = 32 * (2 + WPW);         // threads of a panel CTA
constexpr int WPR = 32 * WPW;               // panel rows per CTA
constexpr int WSL = 68;                     // row stride of a staged panel slab ([k][64 rows]); 4 mod 16: the fragment load of lane
                                            // (g, q) reads word q * 68 + g (+ tile), distinct banks over a half-warp (72 was 2-way)

// One CTA per block row of the upper block-CSR.  Where a block goes depends on what its two poses are:
// interior x interior -> the chunk's band, interior x separator -> the chunk's border rows, separator x separator ->
// the separator system.  Only the lower triangle is written: in band storage an entry above the diagonal would alias
// the tail of the previous column.
__global__ void wband_fill_kernel(WbandView V) {
    const int a = blockIdx.x;
    const int oa = V.owner[a], la = V.local[a];
    const long long ldT = V.T.ld;
    for (int idx = threadIdx.x; idx < (V.rowptr[a + 1] - V.rowptr[a]) * 36; idx += blockDim.x) {
        const int e = V.rowptr[a] + idx / 36, r = (idx % 36) / 6, c = idx % 6;
        const int b = V.col[e];
        const int ob = V.owner[b], lb = V.local[b];
        const double v = V.S[36ll * e + 6 * r + c];   // S[6a + r][6b + c], b >= a
        if (oa >= 0 && ob == oa) {
            const int i = 6 * lb + c, j = 6 * la + r;
            if (i >= j) V.A[oa * V.a_stride + (long long)j * V.ld + i] = v;
        } else if (oa >= 0 && ob == -(oa + 1)) {      // the separator right of a's chunk
            V.Bd[oa * V.b_stride + (long long)(6 * la + r) * V.ldB + V.sepw + 1 + 6 * lb + c] = v;
        } else if (oa < 0 && ob == -oa) {             // a in separator s = -oa - 1, b in chunk s + 1: its left separator
            V.Bd[ob * V.b_stride + (long long)(6 * lb + c) * V.ldB + 6 * la + r] = v;
        } else if (oa < 0 && ob == oa) {
            const int s = -oa - 1;
            const int i = s * V.sepw + 6 * lb + c, j = s * V.sepw + 6 * la + r;
            if (i >= j) V.T.A[j * ldT + i] = v;
        } else {
            *V.fail = 1;  // a coupling the chunk layout cannot hold (the plan guarantees there is none)
        }
    }
    if (threadIdx.x < 6) {
        const int r = threadIdx.x;
        if (oa >= 0)
            V.Bd[oa * V.b_stride + (long long)(6 * la + r) * V.ldB + V.sepw] = V.rhs[6ll * a + r];
        else
            V.T.A[((-oa - 1) * V.sepw + 6 * la + r) * ldT + V.T.n_pad] = V.rhs[6ll * a + r];
    }
}

// identity diagonal of the padding columns (chunks and separator system)
__global__ void wband_pad_kernel(WbandView V) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int per = V.m_pad;
    if (t < V.C * per) {
        const int c = t / per, j = t % per;
        if (j >= V.unit * V.chunk_len[c]) V.A[c * V.a_stride + (long long)j * V.ld + j] = 1.0;
    } else {
        const int j = t - V.C * per;
        if (V.C > 1 && j >= V.T.n && j < V.T.n_pad) V.T.A[(long long)j * V.T.ld + j] = 1.0;
    }
}

// Separator level: the dense separator system T of view P (lower triangle, block tridiagonal with blocks of P.sepw =
// V.unit scalars, rhs as row n_pad) is chunked again.  One CTA per column gj of T: its entries down to the end of the
// next block go to V's band / border / separator storage, by the same rules as wband_fill_kernel with blocks of
// V.unit scalars instead of 6.
__global__ void wband_refill_kernel(WbandView P, WbandView V) {
    const int gj = blockIdx.x;
    const int u = V.unit, ns = P.T.n;
    const int bj = gj / u, oj = gj % u;
    const long long ldT = P.T.ld, ldT2 = V.T.ld;
    const double* Tc = P.T.A + gj * ldT;
    const int oa = V.owner[bj], la = V.local[bj];
    const int hi = min(ns, (bj + 2) * u);
    for (int gi = gj + threadIdx.x; gi < hi; gi += blockDim.x) {
        const int bi = gi / u, oi = gi % u;
        const int ob = V.owner[bi], lb = V.local[bi];
        const double v = Tc[gi];
        if (oa >= 0 && ob == oa) {
            V.A[oa * V.a_stride + (long long)(u * la + oj) * V.ld + u * lb + oi] = v;
        } else if (oa >= 0 && ob == -(oa + 1)) {      // bi is the separator right of bj's chunk
            V.Bd[oa * V.b_stride + (long long)(u * la + oj) * V.ldB + V.sepw + 1 + u * lb + oi] = v;
        } else if (oa < 0 && ob == -oa) {             // bj is separator s = -oa - 1, bi in chunk s + 1: its left separator
            V.Bd[ob * V.b_stride + (long long)(u * lb + oi) * V.ldB + u * la + oj] = v;
        } else if (oa < 0 && ob == oa) {
            const int s = -oa - 1;
            V.T.A[(s * V.sepw + u * la + oj) * ldT2 + s * V.sepw + u * lb + oi] = v;
        } else {
            *V.fail = 1;
        }
    }
    if (threadIdx.x == 0) {
        const double r = Tc[P.T.n_pad];
        if (oa >= 0)
            V.Bd[oa * V.b_stride + (long long)(u * la + oj) * V.ldB + V.sepw] = r;
        else
            V.T.A[((-oa - 1) * V.sepw + u * la + oj) * ldT2 + V.T.n_pad] = r;
    }
}

// rows of the panel at j0: band rows [t0, t0 + mb) of the chunk, then the border rows [left separator | rhs | right
// separator] — the right separator's only from panel r_start on (structurally zero before)
struct PanelRows {
    int t0, mb, m;
};
__device__ __forceinline__ PanelRows panel_rows(const WbandView& V, int j0) {
    PanelRows p;
    p.t0 = j0 + WNB;
    p.mb = min(V.bwr, V.m_pad - p.t0);
    p.m = p.mb + (j0 >= V.r_start ? V.nbr : V.sepw + 1);
    return p;
}

__global__ void __launch_bounds__(WPT, 1) wband_panel_kernel(WbandView V, int j0) {
    constexpr int B = WNB;
    __shared__ __align__(16) double Lt2[(B + 2) * B];
    __shared__ double sInv[B + 2];
    __shared__ __align__(8) uint64_t done0[B + 1];
    __shared__ __align__(8) uint64_t done1[B];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int ch = blockIdx.y;
    const long long ld = V.ld;
    if (*V.fail) return;
    for (int j = tid; j < B + 1; j += WPT) {
        if (j < B) {
            Lt2[j] = 0.0;
            mbar_init(&done1[j], 32);
        }
        mbar_init(&done0[j], 32);
    }
    fence_mbar_init();
    __syncthreads();
    double* Ap = V.A + ch * V.a_stride + j0 * ld;  // column j0 of the chunk's band
    if (warp == 0) {
        const int r = lane;
        double b[33];
        b[0] = 0.0;
#pragma unroll
        for (int c = 0; c < 32; ++c) b[c + 1] = Ap[c * ld + j0 + r];   // (entries above the diagonal are read, never used)
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
        // one row below the diagonal block per thread: a band row, or a border row (separator / right-hand side)
        const PanelRows P = panel_rows(V, j0);
        const int rl = blockIdx.x * WPR + (warp - 2) * 32 + lane;
        const bool valid = rl < P.m + B;
        double* rowp;
        int stride;
        if (rl < P.mb) {
            rowp = Ap + P.t0 + rl;
            stride = int(ld);
        } else if (rl < P.m) {
            rowp = V.Bd + ch * V.b_stride + (long long)j0 * V.ldB + (rl - P.mb);
            stride = V.ldB;
        } else {
            // unit vector e_i, i = rl - m: its row of solutions is row i of L11^-T; Ld[chunk][panel][c][i]
            rowp = V.Ldiag + ((long long)ch * (V.m_pad / B) + j0 / B) * B * B + (rl - P.m);
            stride = B;
        }
        double x[B];
#pragma unroll
        for (int c = 0; c < B; ++c) x[c] = !valid ? 0.0 : rl < P.m ? rowp[(long long)c * stride] : (c == rl - P.m ? 1.0 : 0.0);
        double* outp = valid ? rowp : nullptr;
        odd2_border_phase<B, B>(x, 0, 16, Lt2, sInv, done1, outp, stride);
        odd2_border_phase<B, B - 16>(x, 16, 32, Lt2, sInv, done1, outp, stride);
        odd2_border_phase<B, B - 32>(x, 32, B, Lt2, sInv, done1, outp, stride);
    }
}

__device__ __forceinline__ void wb_dmma_m8n8k4(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// Trailing update behind panel j0 over the local index space [0, m): local row / column l < mb is the chunk's band
// row / column t0 + l, l >= mb is border row l - mb (as a COLUMN: column m_pad + l - mb of the border array, the
// border x border Schur complement).  Lower triangle only.
// Four CTAs per SM (64 registers): the launch is one wave of latency-bound tiles — with 90 registers it was two.
__global__ void __launch_bounds__(256, 4) wband_syrk_kernel(WbandView V, int j0) {
    extern __shared__ __align__(16) double smem_wsyrk[];
    double* Lr = smem_wsyrk;          // Lr[k][i]: panel rows of the tile's row range
    double* Lc = Lr + WNB * WSL;      // panel rows of the tile's column range
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int ch = blockIdx.y;
    const long long ld = V.ld, ldB = V.ldB;
    if (*V.fail) return;
    const PanelRows P = panel_rows(V, j0);
    const int t0 = P.t0, mb = P.mb, m = P.m;
    int I = int((sqrt(8.0 * blockIdx.x + 1.0) - 1.0) * 0.5);
    while ((long long)(I + 1) * (I + 2) / 2 <= blockIdx.x) ++I;
    while ((long long)I * (I + 1) / 2 > blockIdx.x) --I;
    const int J = blockIdx.x - I * (I + 1) / 2;
    double* Ab = V.A + ch * V.a_stride;
    double* Bb = V.Bd + ch * V.b_stride;
    const double* Apan = Ab + j0 * ld + t0;        // band rows of the panel: Apan[k * ld + l]
    const double* Bpan = Bb + j0 * ldB - mb;       // border rows of the panel: Bpan[k * ldB + l]
    {
        // every thread stages 12 entries of each slab: row i = tid & 63, panel columns k = (tid >> 6) + 4 t.  All 24
        // loads are issued before the first store (a loop with a store per load serialises on the load latency)
        const int i = tid & 63, kq = tid >> 6;
        const int li = 64 * I + i, lj = 64 * J + i;
        const double* pr = li < mb ? Apan + kq * ld + li : Bpan + kq * ldB + li;
        const double* pc = lj < mb ? Apan + kq * ld + lj : Bpan + kq * ldB + lj;
        const long long sr = 4 * (li < mb ? ld : ldB), sc = 4 * (lj < mb ? ld : ldB);
        double vr[12], vc[12];
#pragma unroll
        for (int t = 0; t < 12; ++t) {
            vr[t] = li < m ? pr[t * sr] : 0.0;
            vc[t] = lj < m ? pc[t * sc] : 0.0;
        }
#pragma unroll
        for (int t = 0; t < 12; ++t) {
            Lr[(kq + 4 * t) * WSL + i] = vr[t];
            Lc[(kq + 4 * t) * WSL + i] = vc[t];
        }
    }
    __syncthreads();
    // warp = column tile jt of the 64 x 64 block; the transposed accumulator D[m][n] = C[i = 8 it + n][j = 8 jt + m]
    const int jt = warp, g = lane >> 2, q = lane & 3;
    double acc[8][2];
#pragma unroll
    for (int it = 0; it < 8; ++it) acc[it][0] = acc[it][1] = 0.0;
#pragma unroll 4
    for (int k0 = 0; k0 < WNB; k0 += 4) {
        const double a = Lc[(k0 + q) * WSL + 8 * jt + g];
#pragma unroll
        for (int it = 0; it < 8; ++it) {
            const double b = Lr[(k0 + q) * WSL + 8 * it + g];
            wb_dmma_m8n8k4(acc[it][0], acc[it][1], a, b);
        }
    }
    const int j = 64 * J + 8 * jt + g;
    if (j < m) {
        // column j of the target: rows l < mb at colA[l], rows l >= mb at colB[l]
        double* colA = Ab + (long long)(t0 + j) * ld + t0;                                   // only when j < mb
        double* colB = Bb + (j < mb ? (long long)(t0 + j) : (long long)(V.m_pad + j - mb)) * ldB - mb;
        // read-modify-write of the lane's eight row pairs, four at a time: their loads first, then their stores
#pragma unroll
        for (int h0 = 0; h0 < 8; h0 += 4) {
            double2 cv[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = 64 * I + 8 * (h0 + u) + 2 * q;  // rows i, i + 1 of column j (mb is even: both band or both border)
                cv[u] = make_double2(0.0, 0.0);
                if (i + 1 < j || i >= m) continue;            // above the diagonal / beyond the last row
                const double* p = (i < mb ? colA : colB) + i;
                if (i >= j && i + 1 < m)
                    cv[u] = *reinterpret_cast<const double2*>(p);
                else if (i >= j)
                    cv[u].x = p[0];
                else if (i + 1 < m)                           // i + 1 == j: the diagonal entry alone
                    cv[u].y = p[1];
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = 64 * I + 8 * (h0 + u) + 2 * q;
                if (i + 1 < j || i >= m) continue;
                double* p = (i < mb ? colA : colB) + i;
                cv[u].x -= acc[h0 + u][0];
                cv[u].y -= acc[h0 + u][1];
                if (i >= j && i + 1 < m)
                    *reinterpret_cast<double2*>(p) = cv[u];
                else if (i >= j)
                    p[0] = cv[u].x;
                else if (i + 1 < m)
                    p[1] = cv[u].y;
            }
        }
    }
}

// T += the chunks' Schur complements.  One CTA per column gj of the separator system (separator sj): rows down to
// the end of separator sj + 1, and the right-hand-side row.  A gather: every entry has one writer and a fixed order.
__global__ void wband_sep_assemble_kernel(WbandView V) {
    const int gj = blockIdx.x;
    const int sepw = V.sepw, ns = V.T.n;
    const int sj = gj / sepw, oj = gj % sepw;
    const long long ldB = V.ldB, ldT = V.T.ld;
    const double* D0 = V.Bd + sj * V.b_stride + (long long)V.m_pad * ldB;         // chunk sj: separator sj is its RIGHT one
    const double* D1 = V.Bd + (sj + 1) * V.b_stride + (long long)V.m_pad * ldB;   // chunk sj + 1: its LEFT one
    double* Tc = V.T.A + gj * ldT;
    const int hi = min(ns, (sj + 2) * sepw);
    const int R = sepw + 1;   // first right-separator border row
    for (int gi = gj + threadIdx.x; gi < hi; gi += blockDim.x) {
        const int si = gi / sepw, oi = gi % sepw;
        if (si == sj)
            Tc[gi] += D0[(long long)(R + oj) * ldB + R + oi] + D1[(long long)oj * ldB + oi];
        else
            Tc[gi] += D1[(long long)oj * ldB + R + oi];   // (separator sj + 1 is chunk sj + 1's right one)
    }
    // right-hand side: border row sepw; against a right-separator row it is the COLUMN of the stored lower triangle
    if (threadIdx.x == 0) Tc[V.T.n_pad] += D0[(long long)sepw * ldB + R + oj] + D1[(long long)oj * ldB + sepw];
}

// xw = y - Y_sep^T x_sep: one warp per column of a chunk
__global__ void __launch_bounds__(256) wband_backinit_kernel(WbandView V) {
    const int ch = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int j = blockIdx.x * 8 + warp;
    if (*V.fail || j >= V.m_pad) return;
    const int sepw = V.sepw;
    const double* col = V.Bd + ch * V.b_stride + (long long)j * V.ldB;
    double d = 0.0;
    if (ch > 0)
        for (int b = lane; b < sepw; b += 32) d += col[b] * V.xsep[(ch - 1) * sepw + b];
    if (ch + 1 < V.C)
        for (int b = lane; b < sepw; b += 32) d += col[sepw + 1 + b] * V.xsep[ch * sepw + b];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
    if (lane == 0) V.xw[(long long)ch * V.m_pad + j] = col[sepw] - d;
}

// L^T x = xw, one CTA per chunk, panels from the last one up: x_panel = L11^-T xw_panel (a 48 x 48 product with the
// inverse the panel kernel left behind, staged in shared memory one panel ahead), then the bwr columns left of the
// panel lose the panel's contribution (further left the panel's rows are structurally zero).  One launch instead of
// one per panel; a step costs two barriers and one round of global loads.
constexpr int WBT = 512;   // threads of the back-substitution CTA
__global__ void __launch_bounds__(WBT) wband_backsolve_kernel(WbandView V) {
    constexpr int B = WNB;
    constexpr int NL = (B * B + WBT - 1) / WBT;   // entries of L11^-T per thread
    __shared__ double Ls[B * (B + 1)];            // Ls[c][i] = (L11^-T)[i][c] (zero for c < i), row stride B + 1
    __shared__ double ts[B], xs[B];
    const int tid = threadIdx.x, ch = blockIdx.x;
    const long long ld = V.ld;
    if (*V.fail) return;
    double* xw = V.xw + (long long)ch * V.m_pad;
    const double* Ach = V.A + ch * V.a_stride;
    const double* Lch = V.Ldiag + (long long)ch * (V.m_pad / B) * B * B;
    const int n_own = V.unit * V.chunk_len[ch];
    double* yo = V.y + (long long)V.unit * V.chunk_p0[ch];
    double lnext[NL];
    auto load_L = [&](int j0) {
        const double* Li = Lch + (long long)(j0 / B) * B * B;
#pragma unroll
        for (int t = 0; t < NL; ++t) lnext[t] = tid + t * WBT < B * B ? Li[tid + t * WBT] : 0.0;
    };
    auto store_L = [&]() {
#pragma unroll
        for (int t = 0; t < NL; ++t) {
            const int idx = tid + t * WBT;
            if (idx < B * B) Ls[(idx / B) * (B + 1) + idx % B] = lnext[t];
        }
    };
    load_L(V.m_pad - B);
    for (int j0 = V.m_pad - B; j0 >= 0; j0 -= B) {
        store_L();
        if (tid < B) ts[tid] = xw[j0 + tid];
        __syncthreads();
        if (j0 >= B) load_L(j0 - B);   // the next panel's inverse flies during this panel's product and update
        if (tid < 4 * B) {
            // four lanes per row i: c = q, q + 4, ...
            const int i = tid >> 2, q = tid & 3;
            double a = 0.0;
#pragma unroll
            for (int c = 0; c < B; c += 4) a = fma(Ls[(c + q) * (B + 1) + i], ts[c + q], a);
            a += __shfl_xor_sync(0xffffffffu, a, 1);
            a += __shfl_xor_sync(0xffffffffu, a, 2);
            if (q == 0) {
                xs[i] = a;
                if (j0 + i < n_own) yo[j0 + i] = a;
            }
        }
        __syncthreads();
        // two lanes per column j left of the panel: rows j0 .. j0 + 23 and j0 + 24 .. j0 + 47 (contiguous in memory)
        for (int base = 0; base < 2 * V.bwr; base += WBT) {   // (warp-uniform trip count: the pair sum is a shuffle)
            const int it = base + tid;
            const int j = it < 2 * V.bwr ? j0 - V.bwr + (it >> 1) : -1, h = it & 1;
            double d0 = 0.0, d1 = 0.0;
            if (j >= 0) {
                const double2* Lj = reinterpret_cast<const double2*>(Ach + j * ld + j0 + 24 * h);
                double2 l[B / 4];
#pragma unroll
                for (int k = 0; k < B / 4; ++k) l[k] = Lj[k];
#pragma unroll
                for (int k = 0; k < B / 4; ++k) {
                    d0 = fma(l[k].x, xs[24 * h + 2 * k], d0);
                    d1 = fma(l[k].y, xs[24 * h + 2 * k + 1], d1);
                }
            }
            double d = d0 + d1;
            d += __shfl_xor_sync(0xffffffffu, d, 1);
            if (h == 0 && j >= 0) xw[j] -= d;
        }
        // (the next iteration's first barrier orders these xw updates before its reads of ts and the stores to Ls / xs)
        __syncthreads();
    }
}

__global__ void wband_sep_scatter_kernel(WbandView V) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= V.T.n) return;
    const int s = t / V.sepw, o = t % V.sepw;
    V.y[(long long)V.unit * (V.chunk_p0[s] + V.chunk_len[s]) + o] = V.xsep[t];
}

__global__ void wband_status_kernel(const int* fail, double* ps) {
    ps[PS_ITERS] = 1.0;
    ps[PS_FAIL] = *fail ? 2.0 : 0.0;
}

}  // namespace

namespace {

void wband_clear(cudaStream_t s, const WbandView& V) {
    CSLAM_CUDA(cudaMemsetAsync(V.A, 0, sizeof(double) * size_t(V.a_stride) * size_t(V.C), s));
    CSLAM_CUDA(cudaMemsetAsync(V.Bd, 0, sizeof(double) * size_t(V.b_stride) * size_t(V.C), s));
    if (V.C > 1) CSLAM_CUDA(cudaMemsetAsync(V.T.A, 0, sizeof(double) * size_t(V.T.ld) * size_t(V.T.n_pad + 1), s));
}

void wband_pad(cudaStream_t s, const WbandView& V) {
    const int total = V.C * V.m_pad + (V.C > 1 ? V.T.n_pad : 0);
    wband_pad_kernel<<<(total + 255) / 256, 256, 0, s>>>(V);
}

// Everything after the storage of V has been written: factorisation of the chunks, the separator system (dense, or
// the next level), back-substitution into V.y.  Returns the number of kernels launched.
int wband_factor_solve(cudaStream_t s, const WbandView& V) {
    int launched = 0;
    constexpr size_t smem_syrk = sizeof(double) * 2 * WNB * WSL;
    CSLAM_CUDA(cudaFuncSetAttribute(wband_syrk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem_syrk)));
    // (four 52 KB CTAs per SM need the large carve-out; the default left room for two)
    CSLAM_CUDA(cudaFuncSetAttribute(wband_syrk_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    for (int j0 = 0; j0 < V.m_pad; j0 += WNB) {
        const int mb = std::min(V.bwr, V.m_pad - (j0 + WNB)), m = mb + (j0 >= V.r_start ? V.nbr : V.sepw + 1);
        wband_panel_kernel<<<dim3((m + WNB + WPR - 1) / WPR, V.C), WPT, 0, s>>>(V, j0);   // + the 48 unit-vector rows
        const int T = (m + 63) / 64;
        wband_syrk_kernel<<<dim3(T * (T + 1) / 2, V.C), 256, smem_syrk, s>>>(V, j0);
        launched += 2;
    }
    if (V.C > 1) {
        wband_sep_assemble_kernel<<<V.T.n, 128, 0, s>>>(V);
        ++launched;
        if (V.next) {
            const WbandView& N = *V.next;   // (N.y == V.xsep: the next level solves T in place of the dense solver)
            wband_clear(s, N);
            wband_refill_kernel<<<V.T.n, 128, 0, s>>>(V, N);
            wband_pad(s, N);
            launched += 2 + wband_factor_solve(s, N);
        } else {
            launch_dense_factor(s, V.T, V.Txw);
        }
    }
    wband_backinit_kernel<<<dim3((V.m_pad + 7) / 8, V.C), 256, 0, s>>>(V);
    wband_backsolve_kernel<<<V.C, WBT, 0, s>>>(V);
    launched += 2;
    if (V.C > 1) {
        wband_sep_scatter_kernel<<<(V.T.n + 255) / 256, 256, 0, s>>>(V);
        ++launched;
    }
    return launched;
}

}  // namespace

void launch_wband_solve(cudaStream_t s, const WbandView& V, double* ps) {
    CSLAM_CUDA(cudaMemsetAsync(V.fail, 0, sizeof(int), s));
    wband_clear(s, V);
    wband_fill_kernel<<<V.n_free, 128, 0, s>>>(V);
    wband_pad(s, V);
    const int launched = 2 + wband_factor_solve(s, V);
    wband_status_kernel<<<1, 1, 0, s>>>(V.fail, ps);
    CSLAM_CUDA(cudaGetLastError());
    g_kernel_launches.fetch_add(launched + 1, std::memory_order_relaxed);
}

}  // namespace cslam
