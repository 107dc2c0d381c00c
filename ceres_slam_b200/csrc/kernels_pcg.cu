// K3a — block-Jacobi preconditioned conjugate gradients on the block-sparse reduced camera
// system, as ONE persistent cooperative kernel (sm_100a, FP64).
//
// The whole solve runs inside a single launch.  Each CTA (one per SM, 32 warps) owns a contiguous
// range of block rows, processed in tiles.  One CG iteration is
//   (1) tile SpMV  q = S p  reading every stored 6x6 block of S exactly ONCE:
//         a warp takes a row a, four blocks in flight (lane = slot x entry); for block (a,b) it
//         forms the direct product  S_ab p_b  (accumulated into row a) and the mirrored product
//         S_ab^T p_a (6 values, parked in a shared-memory scratch slot of that block);
//         after a CTA barrier row b gathers its mirrored slots from shared memory.
//         Mirrored blocks whose source row lies outside the tile (previous CTA, tile seams) are
//         read transposed from global memory — a few per cent of a banded system.
//         p = z + beta p_old is formed once per tile into a shared-memory window, so p_b comes
//         from shared memory for every in-window column.
//   (2) grid barrier fused with the reduction of p.q (one atomic per CTA),
//   (3) element-parallel update  x += a p, r -= a q, z = Minv r  over the CTA's own rows,
//   (4) grid barrier fused with the reductions of r.z, Q = -x.(b + r), |r|^2.
// S (tens of MB) stays in the 126 MB L2 across iterations; the kernel is bound by L2->SM
// bandwidth and latency, which is why each block is fetched once and index lists live on chip.
// Iteration rules are Ceres' conjugate_gradients_solver: Q-based
// stopping rule with q_tolerance = eta, optional residual rule, residual recomputed from scratch
// every `reset_period` iterations.  Control flow is identical in all threads because every
// decision derives from sums broadcast after a barrier.  No atomics touch vector data, so the
// result is deterministic.
#include "kernels.cuh"

namespace cslam {

namespace {

struct Rec {
    double pq, rho_next, q1, nr2;
};

__device__ __forceinline__ bool zero_or_inf(double x) { return x == 0.0 || isinf(x) || isnan(x); }

constexpr int PCG_THREADS = 1024;   // one CTA per SM, 32 warps
constexpr int PCG_TILE_ROWS = 160;  // block rows per tile
constexpr int PCG_MAXB = 2048;      // upper blocks per tile with a mirrored-product scratch slot
constexpr int PCG_WIN = 256;        // rows of p held in the shared window (tile rows + halo)
constexpr int PCG_MAXM = 2048;      // mirrored-entry list entries cached per tile
constexpr int PCG_MAX_TILES = 256;

struct PcgSmem {
    double m[PCG_MAXB][6];          // mirrored products S_ab^T p_a, one slot per upper block
    double d[PCG_TILE_ROWS][6];     // direct products of the tile's rows
    double v[PCG_WIN][6];           // p window: rows [t0, t0 + PCG_WIN)
    double r[PCG_THREADS + 8];      // residual exchange of the update phase
    double part[33][4];             // reduction scratch / broadcast
    int col[PCG_MAXB];              // cached upper column indices of tile 0
    int2 ment[PCG_MAXM];            // cached mirrored lists of tile 0: (source row, block)
    int rowptr[PCG_TILE_ROWS + 1];
    int mptr[PCG_TILE_ROWS + 1];
    int tile_lo[PCG_MAX_TILES + 1];
    int n_tiles;
    int cached;                     // tile 0 index lists are in shared memory
};

#ifdef PCG_TIMING
__device__ unsigned long long g_pcg_t[148 * 8];
__device__ __forceinline__ unsigned long long gtime() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
#define PCG_T(slot)                                           \
    do {                                                      \
        if (threadIdx.x == 0) {                               \
            const unsigned long long now = gtime();           \
            g_pcg_t[blockIdx.x * 8 + (slot)] += now - t_last; \
            t_last = now;                                     \
        }                                                     \
    } while (0)
#else
#define PCG_T(slot) \
    do {            \
    } while (0)
#endif

// Grid-wide barrier for a co-resident (cooperatively launched) grid: one arrival per CTA on a
// monotonically increasing counter, release/acquire at GPU scope (the acquire invalidates L1, so
// plain loads after the barrier see other CTAs' writes).
struct GridBarrier {
    unsigned int* counter;
    unsigned int target;
    double (*part)[4];
    __device__ __forceinline__ void arrive_and_wait() {
        target += gridDim.x;
        __threadfence();
        atomicAdd(counter, 1u);
        unsigned int seen;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(counter) : "memory");
        } while (seen < target);
    }
    __device__ __forceinline__ void sync() {
        __syncthreads();
        if (threadIdx.x == 0) arrive_and_wait();
        __syncthreads();
    }
    // Barrier fused with up to four global sums: warp shuffle -> shared -> ONE atomic per CTA per
    // sum; afterwards one thread reads the finished sums and broadcasts them through shared
    // memory (thousands of warps loading one global word serialise at a single L2 slice).
    template <int N>
    __device__ __forceinline__ void sync_sum(const double (&v)[N], double* const (&dst)[N], double (&out)[N]) {
        const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
        for (int k = 0; k < N; ++k) {
            const double s = warp_sum(v[k]);
            if (lane == 0) part[w][k] = s;
        }
        __syncthreads();
        if (w == 0) {
            const int nw = blockDim.x >> 5;
#pragma unroll
            for (int k = 0; k < N; ++k) {
                const double s = warp_sum(lane < nw ? part[lane][k] : 0.0);
                if (lane == 0 && s != 0.0) atomicAdd(dst[k], s);
            }
            if (lane == 0) {
                arrive_and_wait();
#pragma unroll
                for (int k = 0; k < N; ++k) part[32][k] = __ldcg(dst[k]);
            }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < N; ++k) out[k] = part[32][k];
        __syncthreads();
    }
};

// q = S v over the CTA's rows, tile by tile, with v = vz + beta * vp (vz may be null: v = vp).
// When pnew != null the CTA's own rows of v are also written there.
// Returns this thread's share of v.q.
__device__ __forceinline__ double cta_spmv(const PcgBufs& B, PcgSmem& sm, const double* vz, const double* vp,
                                           double beta, double* pnew, double* q) {
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int s = lane >> 3, i = lane & 7;
    const int nf = B.nf;
    const int2* ment_all = reinterpret_cast<const int2*>(B.ent_cb);
    double vq = 0.0;
    for (int t = 0; t < sm.n_tiles; ++t) {
        const int t0 = sm.tile_lo[t], t1 = sm.tile_lo[t + 1];
        const int e_lo = B.rowptr[t0];
        const bool cached = (t == 0) && sm.cached;
        const int win_hi = min(nf, t0 + PCG_WIN);
        // ---- p window: v = vz + beta vp for rows [t0, win_hi) ----
        for (int idx = threadIdx.x; idx < 6 * (win_hi - t0); idx += blockDim.x) {
            const int g = 6 * t0 + idx;
            const double val = vz ? vz[g] + beta * vp[g] : vp[g];
            (&sm.v[0][0])[idx] = val;
            if (pnew && idx < 6 * (t1 - t0)) pnew[g] = val;
        }
        __syncthreads();
        // ---- phase 1: direct products into sm.d, mirrored products into sm.m ----
        for (int a = t0 + wib; a < t1; a += 32) {
            const int r = a - t0;
            const int e0 = cached ? sm.rowptr[r] : B.rowptr[a];
            const int e1 = cached ? sm.rowptr[r + 1] : B.rowptr[a + 1];
            double acc = 0.0;
            if (i < 6) {
                const double2* va2 = reinterpret_cast<const double2*>(sm.v[r]);
                const double2 a0 = va2[0], a1 = va2[1], a2 = va2[2];
                for (int e = e0 + s; e < e1; e += 4) {
                    const int b = cached ? sm.col[e - e_lo] : B.col[e];
                    double2 v0, v1, v2;
                    if (b < win_hi) {
                        const double2* vb2 = reinterpret_cast<const double2*>(sm.v[b - t0]);
                        v0 = vb2[0];
                        v1 = vb2[1];
                        v2 = vb2[2];
                    } else {
                        const double2* pb = reinterpret_cast<const double2*>(vp + 6 * b);
                        v0 = pb[0];
                        v1 = pb[1];
                        v2 = pb[2];
                        if (vz) {
                            const double2* zb = reinterpret_cast<const double2*>(vz + 6 * b);
                            const double2 z0 = zb[0], z1 = zb[1], z2 = zb[2];
                            v0 = make_double2(z0.x + beta * v0.x, z0.y + beta * v0.y);
                            v1 = make_double2(z1.x + beta * v1.x, z1.y + beta * v1.y);
                            v2 = make_double2(z2.x + beta * v2.x, z2.y + beta * v2.y);
                        }
                    }
                    const double* blk = B.S + 36ll * e;
                    const double2* row = reinterpret_cast<const double2*>(blk + 6 * i);
                    const double2 r0 = row[0], r1 = row[1], r2 = row[2];
                    acc += r0.x * v0.x + r0.y * v0.y + r1.x * v1.x + r1.y * v1.y + r2.x * v2.x + r2.y * v2.y;
                    if (b != a && e - e_lo < PCG_MAXB) {
                        // column i of the block (the lines were just fetched by the row loads)
                        const double* c = blk + i;
                        sm.m[e - e_lo][i] =
                            c[0] * a0.x + c[6] * a0.y + c[12] * a1.x + c[18] * a1.y + c[24] * a2.x + c[30] * a2.y;
                    }
                }
            }
            acc += __shfl_down_sync(0xffffffffu, acc, 16);
            acc += __shfl_down_sync(0xffffffffu, acc, 8);
            if (lane < 6) sm.d[r][lane] = acc;
        }
        __syncthreads();
        // ---- phase 2: gather the mirrored products of every row ----
        for (int b = t0 + wib; b < t1; b += 32) {
            const int r = b - t0;
            const int m0 = cached ? sm.mptr[r] : B.ent_ptr[b];
            const int m1 = cached ? sm.mptr[r + 1] : B.ent_ptr[b + 1];
            const int m_base = cached ? sm.mptr[0] : 0;
            double y = 0.0;
            if (i < 6) {
                for (int m = m0 + s; m < m1; m += 4) {
                    const int2 ae = cached ? sm.ment[m - m_base] : ment_all[m];
                    const int el = ae.y - e_lo;
                    if (ae.x >= t0 && el >= 0 && el < PCG_MAXB) {
                        y += sm.m[el][i];
                    } else {
                        // source row outside the tile: read the block transposed
                        const int a = ae.x;
                        const double* c = B.S + 36ll * ae.y + i;
#pragma unroll
                        for (int j = 0; j < 6; ++j) {
                            const double va = vz ? vz[6 * a + j] + beta * vp[6 * a + j] : vp[6 * a + j];
                            y += c[6 * j] * va;
                        }
                    }
                }
            }
            y += __shfl_down_sync(0xffffffffu, y, 16);
            y += __shfl_down_sync(0xffffffffu, y, 8);
            if (lane < 6) {
                const double qv = y + sm.d[r][lane];
                q[6 * b + lane] = qv;
                vq += sm.v[r][lane] * qv;
            }
        }
        __syncthreads();
    }
    return vq;
}

__global__ void __launch_bounds__(PCG_THREADS, 1)
    pcg_persistent_kernel(PcgBufs B, double* pbuf2, Rec* rec, unsigned int* bar_counter, double q_tol, double r_tol,
                          int min_iters, int max_iters, int reset_period) {
    extern __shared__ __align__(16) unsigned char pcg_smem_raw[];
    PcgSmem& sm = *reinterpret_cast<PcgSmem*>(pcg_smem_raw);
    GridBarrier grid{bar_counter, 0u, sm.part};
    const int nf = B.nf;
    double* ps = B.ps;
    const int rows_per_cta = (nf + gridDim.x - 1) / gridDim.x;
    const int row_lo = min(nf, int(blockIdx.x) * rows_per_cta);
    const int row_hi = min(nf, row_lo + rows_per_cta);
    const int2* ment_all = reinterpret_cast<const int2*>(B.ent_cb);

    // ---- static tiling of the CTA's rows; index lists of tile 0 cached on chip ----
    if (threadIdx.x == 0) {
        int nt = 0, a = row_lo;
        sm.tile_lo[0] = a;
        while (a < row_hi && nt < PCG_MAX_TILES) {
            int end = a + 1;
            while (end < row_hi && end - a < PCG_TILE_ROWS && B.rowptr[end + 1] - B.rowptr[a] <= PCG_MAXB) ++end;
            a = end;
            sm.tile_lo[++nt] = a;
        }
        sm.n_tiles = nt;  // the launcher refuses systems that would need more tiles per CTA
        int cached = 0;
        if (nt > 0) {
            const int t0 = sm.tile_lo[0], t1 = sm.tile_lo[1];
            cached = (t1 - t0 <= PCG_TILE_ROWS) && (B.rowptr[t1] - B.rowptr[t0] <= PCG_MAXB) &&
                     (B.ent_ptr[t1] - B.ent_ptr[t0] <= PCG_MAXM);
        }
        sm.cached = cached;
    }
    __syncthreads();
    if (sm.cached) {
        const int t0 = sm.tile_lo[0], t1 = sm.tile_lo[1];
        const int e_lo = B.rowptr[t0], m_lo = B.ent_ptr[t0];
        for (int k = threadIdx.x; k <= t1 - t0; k += blockDim.x) {
            sm.rowptr[k] = B.rowptr[t0 + k];
            sm.mptr[k] = B.ent_ptr[t0 + k];
        }
        for (int k = threadIdx.x; k < B.rowptr[t1] - e_lo; k += blockDim.x) sm.col[k] = B.col[e_lo + k];
        for (int k = threadIdx.x; k < B.ent_ptr[t1] - m_lo; k += blockDim.x) sm.ment[k] = ment_all[m_lo + k];
    }
    __syncthreads();

    const int n_own = 6 * (row_hi - row_lo);

    // ---- init: x = 0, r = b, z = Minv r, rho = r.z, |b|^2 ----
    double init_sums[2];
    {
        double rz = 0, nb = 0;
        for (int base = 0; base < n_own; base += 1020) {
            const int idx = base + threadIdx.x;
            const bool on = threadIdx.x < 1020 && idx < n_own;
            const int g = 6 * row_lo + idx;
            const double bi = on ? B.b[g] : 0.0;
            sm.r[threadIdx.x] = bi;
            __syncthreads();
            if (on) {
                const int a = g / 6, row = g - 6 * a, l0 = threadIdx.x - row;
                const double* mi = B.Minv + 36ll * a + 6 * row;
                double z = 0;
#pragma unroll
                for (int j = 0; j < 6; ++j) z += mi[j] * sm.r[l0 + j];
                B.x[g] = 0.0;
                B.r[g] = bi;
                B.z[g] = z;
                B.p[g] = 0.0;
                rz += bi * z;
                nb += bi * bi;
            }
            __syncthreads();
        }
        const double v[2] = {rz, nb};
        double* const d[2] = {&rec[0].rho_next, &ps[PS_NORMB2]};
        grid.sync_sum(v, d, init_sums);
    }
    const double normb2 = init_sums[1];
    double rho = init_sums[0], last_rho = 1.0, Q0 = 0.0;
    const double r_tol2 = r_tol < 0 ? -1.0 : r_tol * r_tol * normb2;
    int k = 0, fail = 0;
#ifdef PCG_TIMING
    unsigned long long t_last = gtime();
#endif
    if (normb2 != 0.0) {
        for (k = 1;; ++k) {
            // conjugate_gradients_solver: rho is checked before the direction update
            double beta = 0.0;
            if (zero_or_inf(rho)) {
                fail = 2;
                --k;
                break;
            }
            if (k > 1) {
                beta = rho / last_rho;
                if (zero_or_inf(beta)) {
                    fail = 2;
                    --k;
                    break;
                }
            }
            Rec* cur = rec + (k % 3);
            if (blockIdx.x == 0 && threadIdx.x == 0) {
                Rec* nxt = rec + ((k + 1) % 3);
                nxt->pq = nxt->rho_next = nxt->q1 = nxt->nr2 = 0.0;
            }
            const double* pold = (k & 1) ? B.p : pbuf2;
            double* pnew = (k & 1) ? pbuf2 : B.p;
            // ---- q = S p with p = z + beta p_old; p.q ----
            const double pq_part = cta_spmv(B, sm, B.z, pold, beta, pnew, B.q);
            PCG_T(0);
            double pq_sum[1];
            {
                const double v[1] = {pq_part};
                double* const d[1] = {&cur->pq};
                grid.sync_sum(v, d, pq_sum);
            }
            PCG_T(1);
            const double pq = pq_sum[0];
            if (pq <= 0.0 || isinf(pq) || isnan(pq)) {
                fail = 1;  // indefinite: NO_CONVERGENCE, x keeps the previous iterate
                break;
            }
            const double alpha = rho / pq;
            if (isinf(alpha) || isnan(alpha)) {
                fail = 2;
                break;
            }
            const bool reset = reset_period > 0 && (k % reset_period == 0);
            if (reset) {
                // r = b - S x from scratch: x first, then q = S x
                for (int idx = threadIdx.x; idx < n_own; idx += blockDim.x) {
                    const int g = 6 * row_lo + idx;
                    B.x[g] += alpha * pnew[g];
                }
                grid.sync();
                cta_spmv(B, sm, nullptr, B.x, 0.0, nullptr, B.q);
            }
            double rz = 0, q1 = 0, nr2 = 0;
            for (int base = 0; base < n_own; base += 1020) {
                const int idx = base + threadIdx.x;
                const bool on = threadIdx.x < 1020 && idx < n_own;
                const int g = 6 * row_lo + idx;
                double ri = 0, xi = 0, bi = 0;
                if (on) {
                    bi = B.b[g];
                    if (reset) {
                        xi = B.x[g];
                        ri = bi - B.q[g];
                    } else {
                        xi = B.x[g] + alpha * pnew[g];
                        B.x[g] = xi;
                        ri = B.r[g] - alpha * B.q[g];
                    }
                }
                sm.r[threadIdx.x] = ri;
                __syncthreads();
                if (on) {
                    const int a = g / 6, row = g - 6 * a, l0 = threadIdx.x - row;
                    const double* mi = B.Minv + 36ll * a + 6 * row;
                    double z = 0;
#pragma unroll
                    for (int j = 0; j < 6; ++j) z += mi[j] * sm.r[l0 + j];
                    B.r[g] = ri;
                    B.z[g] = z;
                    rz += ri * z;
                    q1 -= xi * (bi + ri);
                    nr2 += ri * ri;
                }
                __syncthreads();
            }
            PCG_T(2);
            double upd[3];
            {
                const double v[3] = {rz, q1, nr2};
                double* const d[3] = {&cur->rho_next, &cur->q1, &cur->nr2};
                grid.sync_sum(v, d, upd);
            }
            PCG_T(3);
            const double Q1 = upd[1];
            const double zeta = k * (Q1 - Q0) / Q1;
            bool done = false;
            if (zeta < q_tol && k >= min_iters) done = true;
            if (upd[2] <= r_tol2 && k >= min_iters) done = true;
            if (k >= max_iters) done = true;
            if (done) break;
            Q0 = Q1;
            last_rho = rho;
            rho = upd[0];
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        ps[PS_DONE] = 1.0;
        ps[PS_ITERS] = double(k);
        ps[PS_FAIL] = double(fail);
    }
}

}  // namespace

void launch_pcg_persistent(cudaStream_t s, const PcgBufs& B, double* pbuf2, double* rec3, double q_tol, double r_tol,
                           int min_iters, int max_iters, int reset_period) {
    static PerDevice cache;   // value = SM count, value2 = resident CTAs per SM, both of the current device
    const int smem = int(sizeof(PcgSmem));
    const int dev = PerDevice::current();
    if (cache.first_use(dev)) {
        int sms = 0, per_sm = 0;
        CSLAM_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        CSLAM_CUDA(cudaFuncSetAttribute(pcg_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        CSLAM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pcg_persistent_kernel, PCG_THREADS, smem));
        if (per_sm < 1) throw CudaError("pcg_persistent_kernel does not fit on an SM");
        cache.value[dev] = sms;
        cache.value2[dev] = per_sm;
        cache.mark(dev);
    }
    const int n_sms = cache.value[dev];
    CSLAM_CUDA(cudaMemsetAsync(B.ps, 0, PS_COUNT * sizeof(double), s));
    CSLAM_CUDA(cudaMemsetAsync(rec3, 0, 3 * sizeof(Rec) + 16, s));
    // at least 32 block rows per CTA (one per warp); never more CTAs than can be co-resident
    int want = (B.nf + 31) / 32;
    int grid = want < n_sms ? want : n_sms;
    if (grid < 1) grid = 1;
    if ((B.nf + grid - 1) / grid > PCG_MAX_TILES * 32)
        throw CudaError("reduced camera system too large for the persistent PCG kernel");
    PcgBufs Bc = B;
    Rec* rec = reinterpret_cast<Rec*>(rec3);
    unsigned int* bar = reinterpret_cast<unsigned int*>(rec + 3);
    void* args[] = {&Bc, &pbuf2, &rec, &bar, &q_tol, &r_tol, &min_iters, &max_iters, &reset_period};
    CSLAM_CUDA(cudaLaunchCooperativeKernel((void*)pcg_persistent_kernel, dim3(grid), dim3(PCG_THREADS), args, size_t(smem), s));
    g_kernel_launches.fetch_add(1, std::memory_order_relaxed);
}

}  // namespace cslam
