// Microbenchmark: TMEM -> register load throughput per SM (tcgen05.ld 32x32b.x32), W warps looping over 512 columns.
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
#include "../osml_b200/csrc/knn_ptx.cuh"
using namespace osml;

template <int X>
__device__ __forceinline__ void ld(uint32_t taddr, float (&v)[32]) { ptx::tmem_ld_32x32(taddr, v); }

__global__ void bw_kernel(int iters, int warps_active, long long* cycles, float* sink) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) ptx::tmem_alloc(&slot, 512);
    ptx::tc_fence_before(); __syncthreads(); ptx::tc_fence_after();
    const uint32_t base = slot;
    float acc = 0.f;
    __syncthreads();
    long long t0 = clock64();
    if (warp < warps_active) {
        const uint32_t taddr = base + ((uint32_t)((warp & 3) * 32) << 16);
        for (int it = 0; it < iters; ++it) {
#pragma unroll 1
            for (int c = 0; c < 512; c += 64) {
                float va[32], vb[32];
                ptx::tmem_ld_32x32(taddr + c, va);
                ptx::tmem_ld_32x32(taddr + c + 32, vb);
                ptx::tmem_ld_wait();
                acc += va[0] + va[17] + va[31] + vb[3] + vb[30];
            }
        }
    }
    __syncthreads();
    long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    if (acc == 1234.5f) sink[0] = acc;
    ptx::tc_fence_before(); __syncthreads();
    if (warp == 0) ptx::tmem_dealloc(base, 512);
}

int main() {
    long long* d_c; float* d_s;
    cudaMalloc(&d_c, 148 * 8); cudaMalloc(&d_s, 4);
    for (int warps : {1, 2, 4, 8, 16}) {
        const int iters = 2000;
        bw_kernel<<<148, 512, 0>>>(iters, warps, d_c, d_s);
        cudaError_t e = cudaDeviceSynchronize();
        long long c[148]; cudaMemcpy(c, d_c, 148 * 8, cudaMemcpyDeviceToHost);
        double bytes = (double)warps * iters * 512 * 32 * 4;
        printf("warps=%2d  cycles=%lld  bytes/SM=%.0f  -> %.1f B/cycle/SM (%s)\n", warps, c[0], bytes, bytes / c[0],
               cudaGetErrorString(e));
    }
    return 0;
}
