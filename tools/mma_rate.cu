// Microbenchmark: issue rate of tcgen05.mma kind::f16 (bf16 -> fp32) for the shapes K1 uses.
//   mode 0: SS, cta_group::1, M128 x N x K16, operands in shared memory (what K1 does)
//   mode 1: TS, cta_group::1, A from TMEM
// One thread per CTA issues `iters` groups of 9 MMAs (one K1 tile) back to back, then commits and waits.
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
#include "../osml_b200/csrc/knn_ptx.cuh"
using namespace osml;

__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(acc) : "memory");
}

__global__ void __launch_bounds__(128, 1) rate_kernel(int mode, int n_cols, int iters, int kblocks, long long* cycles) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // fill operands with a bit pattern (finite bf16 values)
    for (int i = threadIdx.x; i < (16384 * 3 + 32768 * 3) / 4; i += blockDim.x)
        reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u + (i * 2654435761u & 0x007f007fu);
    if (threadIdx.x == 0) { ptx::mbar_init(&bar, 1); ptx::fence_barrier_init(); }
    if (warp == 0) ptx::tmem_alloc(&slot, 512);
    ptx::fence_proxy_async();
    ptx::tc_fence_before(); __syncthreads(); ptx::tc_fence_after();
    const uint32_t tmem = slot;
    if (warp == 1 && lane == 0) {
        const uint32_t idesc = ptx::umma_idesc_bf16(128, n_cols);
        const uint32_t a_addr = ptx::smem_u32(smem), b_addr = a_addr + 16384 * 3;
        long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            const uint32_t d = tmem + (it & 1) * 256;
            for (int kb = 0; kb < kblocks; ++kb) {
                const int nks = kb == 2 ? 1 : 4;
                for (int ks = 0; ks < nks; ++ks) {
                    const uint64_t db = ptx::umma_desc_k_sw128(b_addr + kb * 32768 + ks * 32);
                    if (mode == 0) {
                        const uint64_t da = ptx::umma_desc_k_sw128(a_addr + kb * 16384 + ks * 32);
                        ptx::umma_bf16_ss(d, da, db, idesc, (uint32_t)((kb | ks) != 0));
                    } else {
                        umma_bf16_ts(d, tmem + 256 * 0 + 448 + kb * 32 + ks * 8, db, idesc, (uint32_t)((kb | ks) != 0));
                    }
                }
            }
        }
        ptx::umma_commit(&bar);
        ptx::mbar_wait(&bar, 0);
        long long t1 = clock64();
        cycles[blockIdx.x] = t1 - t0;
    }
    ptx::tc_fence_before(); __syncthreads();
    if (warp == 0) ptx::tmem_dealloc(tmem, 512);
}

int main() {
    long long* d_c; cudaMalloc(&d_c, 148 * 8);
    const size_t smem = 16384 * 3 + 32768 * 3 + 1024;
    cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int mode = 0; mode < 2; ++mode)
        for (int n : {256, 128, 64}) {
            const int iters = 2000, kblocks = 3;
            rate_kernel<<<148, 128, smem>>>(mode, n, iters, kblocks, d_c);
            cudaError_t e = cudaDeviceSynchronize();
            long long c[148]; cudaMemcpy(c, d_c, 148 * 8, cudaMemcpyDeviceToHost);
            long long mx = 0; for (int i = 0; i < 148; ++i) mx = c[i] > mx ? c[i] : mx;
            printf("mode=%d (%s) M128 N%d: %lld cycles for %d tiles x 9 MMAs -> %.1f cycles/MMA, %.1f cycles/tile (%s)\n",
                   mode, mode ? "TS" : "SS", n, mx, iters, (double)mx / (iters * 9), (double)mx / iters, cudaGetErrorString(e));
        }
    // one CTA only (no chip-level power / clock interaction)
    rate_kernel<<<1, 128, smem>>>(0, 256, 2000, 3, d_c);
    cudaDeviceSynchronize();
    long long c0; cudaMemcpy(&c0, d_c, 8, cudaMemcpyDeviceToHost);
    printf("single CTA SS N256: %.1f cycles/MMA\n", (double)c0 / (2000 * 9));
    return 0;
}
