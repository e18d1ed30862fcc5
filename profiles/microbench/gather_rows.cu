// Microbenchmark: achievable bandwidth of random fp64-row gathers on B200 (what the sweep's upwind reads look like).
// Each warp reads rows of `nlam` doubles at random row indices from a large array (>> L2) and accumulates them.
// usage: gather_rows <nlam> <rows_million> <tma:0|1>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("%s: %s\n",#x,cudaGetErrorString(e)); exit(1);} }while(0)

__global__ void k_ldg(const double* __restrict__ a, const uint32_t* __restrict__ idx, int64_t nidx, int nlam, double* out, int unroll_rows) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    double acc = 0;
    for (int64_t i = warp * 8; i + 8 <= nidx; i += nwarps * 8) {
        uint32_t r[8];
#pragma unroll
        for (int k = 0; k < 8; k++) r[k] = __ldg(idx + i + k);
        for (int l = lane; l < nlam; l += 32) {
            double v[8];
#pragma unroll
            for (int k = 0; k < 8; k++) v[k] = __ldg(a + (size_t)r[k] * nlam + l);
#pragma unroll
            for (int k = 0; k < 8; k++) acc += v[k];
        }
    }
    if (acc == 12345.678) out[0] = acc;
}

int main(int argc, char** argv) {
    int nlam = argc > 1 ? atoi(argv[1]) : 91;
    int64_t nrows = (int64_t)(argc > 2 ? atof(argv[2]) : 20) * 1000000;   // 20M rows x 728 B = 14.6 GB
    int64_t nidx = 64ll << 20;
    double* a; uint32_t* idx; double* out;
    CK(cudaMalloc(&a, sizeof(double) * nrows * nlam));
    CK(cudaMemset(a, 0, sizeof(double) * nrows * nlam));
    CK(cudaMalloc(&idx, sizeof(uint32_t) * nidx));
    CK(cudaMalloc(&out, 8));
    uint32_t* h = (uint32_t*)malloc(sizeof(uint32_t) * nidx);
    uint64_t s = 88172645463325252ull;
    // locality pattern like the sweep: rows drawn from a sliding window of ~60k rows (two layers)
    for (int mode = 0; mode < 2; mode++) {
        for (int64_t i = 0; i < nidx; i++) {
            s ^= s << 13; s ^= s >> 7; s ^= s << 17;
            if (mode == 0) h[i] = (uint32_t)(s % (uint64_t)nrows);
            else { int64_t base = (i * (nrows - 60000)) / nidx; h[i] = (uint32_t)(base + (int64_t)(s % 60000)); }
        }
        CK(cudaMemcpy(idx, h, sizeof(uint32_t) * nidx, cudaMemcpyHostToDevice));
        for (int bs_per_sm = 1; bs_per_sm <= 8; bs_per_sm *= 2) {
            cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
            int grid = 148 * bs_per_sm;
            k_ldg<<<grid, 256>>>(a, idx, nidx, nlam, out, 8);
            cudaEventRecord(e0);
            k_ldg<<<grid, 256>>>(a, idx, nidx, nlam, out, 8);
            cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            printf("mode=%s nlam=%d blocks/SM=%d (warps/SM=%d): %.2f ms  %.1f GB/s  %.2f Grows/s\n", mode ? "window" : "uniform", nlam, bs_per_sm, bs_per_sm * 8, ms,
                   (double)nidx * nlam * 8 / ms / 1e6, nidx / ms / 1e6);
        }
    }
    return 0;
}
