// K1-phong — materialised residuals and Jacobians of the lighting blocks of dataset_ba_phong
// (tests/dataset_ba_phong.cpp:100-205): per observation one IntensityError{Point,Directional}Light
// block (intensity_error_point_light.hpp:24-113, intensity_error_directional_light.hpp:24-113) and
// one NormalError block (normal_error.hpp:16-54), Jacobians in tangent coordinates (SE3Perturbation
// on the pose, UnitVectorPerturbation on the normal and on a directional light).
//
// One thread per observation, closed forms of closed_form.h (no Jets, no heap: the reference
// allocates two shared_ptr objects per functor call).  Per observation the kernel reads 12 B of
// indices + 32 B of observations (+ pose / vertex / material gathers that live in L2) and writes
// 50 doubles: HBM-bound.  Outputs are staged per 128-observation tile in shared memory and leave
// as TMA bulk stores (cp.async.bulk.global.shared::cta); several CTAs per SM overlap tiles.
#include "kernels.cuh"

namespace cslam {

namespace {

constexpr int PH_TILE = 128;
constexpr int PH_STAGE = PH_TILE * 50;  // r_I 1 | J_I 19 | r_N 3 | Jpose_N 18 | Jn_N 9 per observation
constexpr size_t PH_SMEM = PH_STAGE * sizeof(double);  // one staging tile: 51 KB -> 3-4 CTAs per SM

__global__ void __launch_bounds__(PH_TILE)
    phong_eval_kernel(PhongView v, double* __restrict__ out_rI, double* __restrict__ out_JI, double* __restrict__ out_rN,
                      double* __restrict__ out_JNc, double* __restrict__ out_JNn, double* __restrict__ cost_out) {
    extern __shared__ __align__(128) unsigned char smem_ph[];
    double* s_out = reinterpret_cast<double*>(smem_ph);
    __shared__ double s_red[32];
    const int tid = threadIdx.x;
    const long long n_tiles = (v.n + PH_TILE - 1) / PH_TILE;
    double cost = 0.0;
    double Wn[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) Wn[k] = v.W_normal[k];
    const double light[3] = {v.light[0], v.light[1], v.light[2]};
    int it = 0;
    for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
        const long long base = t * PH_TILE, i = base + tid;
        const bool active = i < v.n;
        double rI = 0.0, JI[19], rN[3] = {0, 0, 0}, JNc[18], JNn[9];
        if (active) {
            const uint32_t c = v.cam[i], j = v.vertex[i];
            double pose[12];
#pragma unroll
            for (int k = 0; k < 12; ++k) pose[k] = v.poses[12ll * c + k];
            const double p[3] = {v.points[3ll * j], v.points[3ll * j + 1], v.points[3ll * j + 2]};
            const double nrm[3] = {v.normals[3ll * j], v.normals[3ll * j + 1], v.normals[3ll * j + 2]};
            const uint32_t m = v.material_id[j];
            const double phong[3] = {v.phong[3ll * m], v.phong[3ll * m + 1], v.phong[3ll * m + 2]};
            intensity_block(pose, p, nrm, phong, v.texture[j], light, v.intensity[i], v.int_stiffness, v.directional != 0,
                            &rI, JI, JI + 6, JI + 9, JI + 12, JI + 15, JI + 16);
            const double nobs[3] = {v.normal_obs[3 * i], v.normal_obs[3 * i + 1], v.normal_obs[3 * i + 2]};
            normal_block(pose, nrm, nobs, Wn, rN, JNc, JNn);
            if (v.cam_free[c] < 0) {
                // constant pose block: its columns are dropped (dataset_ba_phong.cpp:76)
#pragma unroll
                for (int k = 0; k < 6; ++k) JI[k] = 0.0;
#pragma unroll
                for (int k = 0; k < 18; ++k) JNc[k] = 0.0;
            }
            cost += 0.5 * (rI * rI + rN[0] * rN[0] + rN[1] * rN[1] + rN[2] * rN[2]);
        }
        if (base + PH_TILE <= v.n) {
            double* st = s_out;
            if (tid == 0) tma_store_wait_read<0>();  // the bulk stores of the previous tile have read the stage
            __syncthreads();
            st[tid] = rI;
#pragma unroll
            for (int k = 0; k < 19; ++k) st[PH_TILE + 19 * tid + k] = JI[k];
#pragma unroll
            for (int k = 0; k < 3; ++k) st[PH_TILE * 20 + 3 * tid + k] = rN[k];
#pragma unroll
            for (int k = 0; k < 18; ++k) st[PH_TILE * 23 + 18 * tid + k] = JNc[k];
#pragma unroll
            for (int k = 0; k < 9; ++k) st[PH_TILE * 41 + 9 * tid + k] = JNn[k];
            fence_proxy_async();
            __syncthreads();
            if (tid == 0) {
                if (out_rI) tma_store_1d(out_rI + base, st, PH_TILE * 8);
                if (out_JI) tma_store_1d(out_JI + 19 * base, st + PH_TILE, PH_TILE * 19 * 8);
                if (out_rN) tma_store_1d(out_rN + 3 * base, st + PH_TILE * 20, PH_TILE * 3 * 8);
                if (out_JNc) tma_store_1d(out_JNc + 18 * base, st + PH_TILE * 23, PH_TILE * 18 * 8);
                if (out_JNn) tma_store_1d(out_JNn + 9 * base, st + PH_TILE * 41, PH_TILE * 9 * 8);
                tma_store_commit();
            }
        } else if (active) {
            if (out_rI) out_rI[i] = rI;
            if (out_JI)
                for (int k = 0; k < 19; ++k) out_JI[19 * i + k] = JI[k];
            if (out_rN)
                for (int k = 0; k < 3; ++k) out_rN[3 * i + k] = rN[k];
            if (out_JNc)
                for (int k = 0; k < 18; ++k) out_JNc[18 * i + k] = JNc[k];
            if (out_JNn)
                for (int k = 0; k < 9; ++k) out_JNn[9 * i + k] = JNn[k];
        }
    }
    if (tid == 0) tma_store_wait<0>();
    block_atomic_sum(cost, cost_out, s_red);
}

}  // namespace

void launch_phong_eval(cudaStream_t s, const PhongView& v, double* r_int, double* J_int, double* r_normal,
                       double* Jpose_normal, double* Jn_normal, double* cost) {
    if (v.n <= 0) return;
    static PerDevice attr_done;
    const int dev_ = PerDevice::current();
    if (attr_done.first_use(dev_)) {
        CSLAM_CUDA(cudaFuncSetAttribute(phong_eval_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(PH_SMEM)));
        attr_done.mark(dev_);
    }
    const long long tiles = (v.n + PH_TILE - 1) / PH_TILE;
    const int grid = int(tiles < 3ll * 148 ? tiles : 3ll * 148);  // persistent: 3 CTAs per SM (registers)
    phong_eval_kernel<<<grid, PH_TILE, PH_SMEM, s>>>(v, r_int, J_int, r_normal, Jpose_normal, Jn_normal, cost);
    g_kernel_launches.fetch_add(1, std::memory_order_relaxed);
    CSLAM_CUDA(cudaGetLastError());
}

}  // namespace cslam
