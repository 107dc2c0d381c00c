// Microbenchmark of the persistent PCG kernel on a synthetic block-banded SPD system
// (nf block rows, half-bandwidth `band` blocks) — isolates K3a from the BA pipeline.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <random>
#define PCG_TIMING 1
#include "../../ceres_slam_b200/csrc/kernels_pcg.cu"
namespace cslam { std::atomic<unsigned long long> g_kernel_launches{0}; }
using namespace cslam;
int main(int argc, char** argv) {
    int nf = argc > 1 ? atoi(argv[1]) : 19999, band = argc > 2 ? atoi(argv[2]) : 9, iters = argc > 3 ? atoi(argv[3]) : 200;
    std::mt19937 rng(1);
    std::normal_distribution<double> nd(0, 1);
    std::vector<int> rowptr(nf + 1, 0), col;
    for (int a = 0; a < nf; ++a) { for (int b = a; b < std::min(nf, a + band + 1); ++b) col.push_back(b); rowptr[a + 1] = (int)col.size(); }
    int nnz = (int)col.size();
    std::vector<double> S(36ull * nnz), Minv(36ull * nf, 0.0), b(6ull * nf);
    for (int a = 0; a < nf; ++a)
        for (int e = rowptr[a]; e < rowptr[a + 1]; ++e)
            for (int k = 0; k < 36; ++k) {
                double v = 0.01 * nd(rng);
                if (col[e] == a) { int r = k / 6, c = k % 6; v = (r == c) ? 10.0 : 0.0; }
                S[36ull * e + k] = v;
            }
    for (int a = 0; a < nf; ++a) for (int k = 0; k < 6; ++k) Minv[36ull * a + 7 * k] = 0.1;
    for (auto& x : b) x = nd(rng);
    std::vector<int> ep(nf + 1, 0);
    for (int a = 0; a < nf; ++a) for (int e = rowptr[a]; e < rowptr[a + 1]; ++e) { if (col[e] != a) ep[col[e] + 1]++; }
    for (int a = 0; a < nf; ++a) ep[a + 1] += ep[a];
    std::vector<int> ecb(2ull * ep[nf] + 2), fill(ep.begin(), ep.end() - 1);
    for (int a = 0; a < nf; ++a) for (int e = rowptr[a]; e < rowptr[a + 1]; ++e) {
        int bb = col[e];
        if (bb != a) { ecb[2 * fill[bb]] = a; ecb[2 * fill[bb]++ + 1] = e; }
    }
    auto up = [](const void* h, size_t bytes) { void* d; cudaMalloc(&d, bytes); cudaMemcpy(d, h, bytes, cudaMemcpyHostToDevice); return d; };
    PcgBufs B;
    B.rowptr = (int*)up(rowptr.data(), rowptr.size() * 4); B.col = (int*)up(col.data(), col.size() * 4);
    B.ent_ptr = (int*)up(ep.data(), ep.size() * 4); B.ent_cb = (int*)up(ecb.data(), ecb.size() * 4);
    B.S = (double*)up(S.data(), S.size() * 8); B.Minv = (double*)up(Minv.data(), Minv.size() * 8); B.b = (double*)up(b.data(), b.size() * 8);
    double *x, *r, *z, *p, *q, *ps, *p2, *rec;
    size_t nv = 6ull * nf * 8;
    cudaMalloc(&x, nv); cudaMalloc(&r, nv); cudaMalloc(&z, nv); cudaMalloc(&p, nv); cudaMalloc(&q, nv); cudaMalloc(&p2, nv);
    cudaMalloc(&ps, 16 * 8); cudaMalloc(&rec, 16 * 8);
    B.x = x; B.r = r; B.z = z; B.p = p; B.q = q; B.ps = ps; B.nf = nf;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        launch_pcg_persistent(0, B, p2, rec, -1.0, -1.0, 0, iters, 10);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double hps[16]; cudaMemcpy(hps, ps, 128, cudaMemcpyDeviceToHost);
        printf("nf=%d band=%d nnzU=%d (%.1f MB) iters=%d  %.3f ms  %.2f us/iter  fail=%g err=%s\n", nf, band, nnz, 288.0 * nnz / 1e6,
               (int)hps[PS_ITERS], ms, 1e3 * ms / hps[PS_ITERS], hps[PS_FAIL], cudaGetErrorString(cudaGetLastError()));
    }
    unsigned long long t[148 * 8];
    cudaMemcpyFromSymbol(t, g_pcg_t, sizeof(t));
    double mx[4] = {0, 0, 0, 0}, av[4] = {0, 0, 0, 0};
    int nb = std::min(148, (nf + 31) / 32);
    for (int b2 = 0; b2 < nb; ++b2) for (int s = 0; s < 4; ++s) { double v = t[b2 * 8 + s] / (3.0 * iters) * 1e-3; av[s] += v / nb; if (v > mx[s]) mx[s] = v; }
    printf("per-iteration us (avg over CTAs / max): spmv %.2f/%.2f  barrier1 %.2f/%.2f  update %.2f/%.2f  barrier2 %.2f/%.2f\n",
           av[0], mx[0], av[1], mx[1], av[2], mx[2], av[3], mx[3]);
    return 0;
}
