// Microbenchmark: tcgen05.mma.cta_group::2 (M256 x N256 x K16, bf16) tile rate as K1's pair kernel issues it:
// `ksteps` MMAs per accumulator tile (the first one overwrites), alternating between two TMEM stages, one commit per
// tile; no operand traffic (stale shared memory), no epilogue.  Prints cycles per tile for several K depths.
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
#include "../osml_b200/csrc/knn_ptx.cuh"
using namespace osml;

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(384, 1)
pair_rate_kernel(int ksteps, int tiles, int commit_every, int mode, int n_cols, long long* cycles) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar, bar2, bar3, done;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = ptx::cluster_ctarank();
    for (int i = threadIdx.x; i < (16384 * 2 + 16384 * 2 * 5) / 4; i += blockDim.x)
    {
        uint32_t x = (i + 1u + blockIdx.x * 7919u) * 2654435761u; x ^= x >> 15; x *= 2246822519u; x ^= x >> 13;
        reinterpret_cast<uint32_t*>(smem)[i] = (mode & 4) ? ((x & 0x81ff81ffu) | 0x3e003e00u)      // random sign, exponent, mantissa
                                                          : 0x3c003c00u + (i * 2654435761u & 0x007f007fu);
    }
    if (threadIdx.x == 0) { ptx::mbar_init(&bar, 1); ptx::mbar_init(&bar2, 1); ptx::mbar_init(&bar3, 1); ptx::mbar_init(&done, 1); ptx::fence_barrier_init(); }
    if (warp == 0) ptx::tmem_alloc_pair(&slot, 512);
    ptx::fence_proxy_async();
    ptx::tc_fence_before(); ptx::cluster_sync_all(); ptx::tc_fence_after();
    const uint32_t tmem = slot;
    if (warp == 1 && lane == 0 && rank == 0) {
        const uint32_t idesc = ptx::umma_idesc_bf16(256, n_cols);
        const uint32_t hi = ptx::umma_desc_hi_k_sw128();
        const uint32_t a_lo0 = ptx::umma_desc_lo(ptx::smem_u32(smem)), b_lo0 = ptx::umma_desc_lo(ptx::smem_u32(smem) + 16384 * 2);
        long long t0 = clock64();
        for (int t = 0; t < tiles; ++t) {
            const uint32_t d = tmem + (t & 1) * 256;
            const uint32_t b_lo = b_lo0 + ((mode & 1) ? (uint32_t)(t % 5) * (32768 >> 4) : 0u);   // ring of 5 stages of 32 KB
            for (int ks = 0; ks < ksteps; ++ks) {
                const int kb = ks >> 2, kk = ks & 3;
                ptx::umma_bf16_ss_pair_lohi(d, a_lo0 + kb * (16384 >> 4) + kk * 2, b_lo + kb * (16384 >> 4) + kk * 2, hi, idesc,
                                            (uint32_t)(ks != 0));
            }
            if (commit_every == 1) {          // like K1: two commits per tile, nobody waits on them here
                ptx::umma_commit_pair(&bar);
                ptx::umma_commit_pair(&bar2);
            }
        }
        ptx::umma_commit_pair(&bar3);         // fresh barrier: completes once everything issued has retired
        ptx::mbar_wait(&bar3, 0);
        long long t1 = clock64();
        cycles[blockIdx.x >> 1] = t1 - t0;
    }
    if ((mode & 2) && warp >= 4) {          // 8 warps polling a barrier that completes when the issuer is done
        while (!__all_sync(0xffffffffu, ptx::mbar_try_wait(&done, 0))) {}
    }
    if (warp == 1 && lane == 0) { if (rank == 0) { ptx::mbar_arrive(&done); ptx::mbar_arrive_remote(&done, 1); } }
    ptx::tc_fence_before(); ptx::cluster_sync_all();
    if (warp == 0) { ptx::tc_fence_after(); ptx::tmem_dealloc_pair(tmem, 512); }
}

int main() {
    long long* d_c; cudaMalloc(&d_c, 74 * 8);
    const size_t smem = 16384 * 2 + 32768 * 5 + 1024;
    cudaFuncSetAttribute(pair_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int n_cols : {256, 128, 64})
        for (int ks : {6, 8}) {
            const int grid = 148, tiles = 40000, commit_every = 1, mode = 4;
            pair_rate_kernel<<<grid, 384, smem>>>(ks, tiles, commit_every, mode, n_cols, d_c);
            cudaError_t e = cudaDeviceSynchronize();
            long long c[74]; cudaMemcpy(c, d_c, 74 * 8, cudaMemcpyDeviceToHost);
            long long mx = 0; for (int i = 0; i < grid / 2; ++i) mx = c[i] > mx ? c[i] : mx;
            printf("N=%d mode=%d commit_every=%d grid=%3d ksteps=%2d: %.1f cycles/tile, %.1f cycles/MMA (%s)\n", n_cols, mode, commit_every, grid, ks, (double)mx / tiles,
                   (double)mx / tiles / ks, cudaGetErrorString(e));
        }
    return 0;
}
