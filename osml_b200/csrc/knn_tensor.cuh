// K0 / K1 / K3: tcgen05 tensor-core candidate filter with a provably complete exact rescoring.
//
// The reference computes sqrt(sum((t - q)^2)) for every (query, train row) pair in fp32 (predict.py:742) and keeps the
// k+1 smallest (predict.py:750).  Doing that arithmetic literally costs 3 un-fusable CUDA-core instructions per
// element (K2).  Here the tensor cores only *filter*:
//
//   K0  knn_prep_rows      one streaming pass: centre by the train column means, round to bf16, append three bf16
//                          columns holding -0.5*|t^|^2 (hi/mid/lo split, exact to 2^-24) to each train row and
//                          (1, 1, 1) to each query row, record |q^|^2 and the rounding residuals |q^ - (q - c)|.
//   K1  knn_tc_filter      persistent tcgen05 kernel: acc[q][t] = q^ . t^ - 0.5*|t^|^2 (so |q^ - t^|^2 = |q^|^2 - 2*acc)
//                          with TMA-staged, 128B-swizzled operand tiles and FP32 accumulators in TMEM; the epilogue
//                          reads each 128 x 256 accumulator tile straight out of TMEM and keeps, per query, the K'
//                          largest acc per train segment.  The distance matrix never leaves the SM.
//   K3  knn_rescore        per query: exact distances of the candidates in the oracle's arithmetic and summation
//                          order, top-(k+1) by (distance, index), and a completeness proof -- every row the filter
//                          dropped has acc <= the segment's K'-th best, which bounds its true distance from below
//                          by sqrt(|q^|^2 - 2*(thr + eps)) - (Eq + Et); if that bound exceeds the exact (k+1)-th
//                          distance the result equals the brute-force one.  Queries that cannot be proven are
//                          re-run on K2.  So the output is bit-identical to K2 / the oracle whatever the filter does.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <algorithm>
#include <cmath>
#include <functional>
#include <string>
#include <vector>

#include "knn_common.cuh"
#include "knn_ptx.cuh"

namespace osml {

// ---------------------------------------------------------------------------------------------------------------------
// geometry
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kTcBM = 128;            // queries per tile (UMMA M)
constexpr int kTcBN = 256;            // train rows per tile (UMMA N)
constexpr int kTcKB = 64;             // bf16 per K block = one 128-byte swizzled row
constexpr int kTcABlockBytes = kTcBM * 128;   // 16 KB
constexpr int kTcBStageBytes = kTcBN * 128;   // 32 KB
constexpr int kTcThreads = 384;       // warp 0 TMA, 1 MMA, 2 TMEM alloc, 3 idle, 4..11 epilogue
constexpr int kTcMaxKPrime = 64;
constexpr int kTcMaxCand = 512;       // candidates per query K3 handles (slots * K')
constexpr float kTcPadScore = -3.0e38f;

struct TcParams {
    int nkb;             // K blocks of the padded feature dimension
    int ksteps;          // UMMA K steps (16 bf16 each) actually needed: ceil((d + 3) / 16)
    int nqt;             // query tiles
    int nseg;            // train segments
    int tiles_per_seg;   // 256-row tiles per segment
    int tiles_total;
    int kprime;          // candidates kept per (query, slot)
    int epi_warps;       // 4 or 8
    int stages;          // B ring depth
    int a_resident;      // 1: the query tile stays in shared memory for a whole sweep; 0 (wide rows): its K blocks
                         //    are streamed through the ring next to the train blocks
    int nslot;           // nseg * halves
    float* cand_val;     // [qpad][nslot][kprime]
    int* cand_idx;
    float* cand_thr;     // [qpad][nslot]: every row the slot dropped has acc <= this (-inf: nothing dropped)
    int debug;           // ablation switch for profiling: 1 = epilogue skips TMEM loads, 2 = loads but no selection
};

// ---------------------------------------------------------------------------------------------------------------------
// K0: preparation
// ---------------------------------------------------------------------------------------------------------------------
__global__ void tc_col_partial_sums_kernel(const float* __restrict__ x, int64_t n, int d, int rows_per_block,
                                           double* __restrict__ partial) {
    __shared__ double red[8][33];
    const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
    const int64_t r1 = min(n, r0 + rows_per_block);
    for (int j0 = 0; j0 < d; j0 += 32) {
        const int j = j0 + threadIdx.x;
        double s = 0.0;
        if (j < d)
            for (int64_t r = r0 + threadIdx.y; r < r1; r += 8) s += (double)x[r * d + j];
        red[threadIdx.y][threadIdx.x] = s;
        __syncthreads();
        if (threadIdx.y == 0 && j < d) {
            double t = 0.0;
            for (int y = 0; y < 8; ++y) t += red[y][threadIdx.x];
            partial[(int64_t)blockIdx.x * d + j] = t;
        }
        __syncthreads();
    }
}

__global__ void tc_col_mean_kernel(const double* __restrict__ partial, int nblocks, int d, int64_t n,
                                   float* __restrict__ center) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= d) return;
    double s = 0.0;
    for (int b = 0; b < nblocks; ++b) s += partial[(int64_t)b * d + j];
    const float c = (float)(s / (double)n);
    center[j] = (c == c && fabsf(c) < INFINITY) ? c : 0.0f;
}

struct TcPrepStats {          // device-side reductions over the train rows
    unsigned long long et_max_bits;      // max |t^ - (t - c)|     (double bits; non-negative doubles order like uint64)
    unsigned long long tnorm_max_bits;   // max |t^|
    int nonfinite;
};

// one warp per row
template <bool kTrain>
__global__ void tc_prep_rows_kernel(const float* __restrict__ x, int64_t n, int64_t npad, int d, int dp,
                                    const float* __restrict__ center, __nv_bfloat16* __restrict__ out,
                                    double* __restrict__ norm2, double* __restrict__ err, TcPrepStats* stats) {
    const int lane = threadIdx.x & 31;
    const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= npad) return;
    __nv_bfloat16* o = out + row * dp;
    if (row >= n) {                       // padding row: can never be selected (train) / is ignored (queries)
        for (int j = lane; j < dp; j += 32) o[j] = __float2bfloat16(0.0f);
        if (kTrain && lane == 0) o[d] = __float2bfloat16(kTcPadScore);
        if (!kTrain && lane == 0 && norm2) { norm2[row] = 0.0; err[row] = 0.0; }
        return;
    }
    const float* xr = x + row * d;
    double n2 = 0.0, e2 = 0.0;
    for (int j = lane; j < d; j += 32) {
        const float v = xr[j], c = center[j];
        const __nv_bfloat16 h = __float2bfloat16_rn(__fsub_rn(v, c));
        o[j] = h;
        const double hd = (double)__bfloat162float(h);
        const double diff = hd - ((double)v - (double)c);
        n2 += hd * hd;
        e2 += diff * diff;
    }
    for (int off = 16; off > 0; off >>= 1) {
        n2 += __shfl_xor_sync(0xffffffffu, n2, off);
        e2 += __shfl_xor_sync(0xffffffffu, e2, off);
    }
    for (int j = d + 3 + lane; j < dp; j += 32) o[j] = __float2bfloat16(0.0f);
    if (lane == 0) {
        const double e = sqrt(e2) * (1.0 + 1e-12);
        if (kTrain) {
            const double h = -0.5 * n2;
            const __nv_bfloat16 h1 = __double2bfloat16(h);
            const double r1 = h - (double)__bfloat162float(h1);
            const __nv_bfloat16 h2 = __double2bfloat16(r1);
            const double r2 = r1 - (double)__bfloat162float(h2);
            const __nv_bfloat16 h3 = __double2bfloat16(r2);
            o[d] = h1; o[d + 1] = h2; o[d + 2] = h3;
            const double tn = sqrt(n2) * (1.0 + 1e-12);
            if (!(n2 < 1.0e37) || !(e2 < 1.0e37)) atomicExch(&stats->nonfinite, 1);
            else {
                atomicMax(&stats->et_max_bits, (unsigned long long)__double_as_longlong(e));
                atomicMax(&stats->tnorm_max_bits, (unsigned long long)__double_as_longlong(tn));
            }
        } else {
            o[d] = __float2bfloat16(1.0f); o[d + 1] = __float2bfloat16(1.0f); o[d + 2] = __float2bfloat16(1.0f);
            norm2[row] = n2;
            err[row] = e;     // NaN / inf propagate and make the query fall back
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// K1: persistent tcgen05 filter
// ---------------------------------------------------------------------------------------------------------------------
// Per-thread candidate list in shared memory (slot e of a thread lives e * stride_bytes further on).  The insert runs
// out of line with scalar register arguments and 32-bit shared addresses, so the hot loop only pays a call when a
// candidate actually beats the thread's current K'-th best; it returns the new minimum and its slot.
__device__ __forceinline__ float lds_f32(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts_b32(uint32_t addr, uint32_t v) {
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __noinline__ unsigned long long tc_list_insert(uint32_t val_addr, uint32_t idx_addr, int kprime,
                                                          uint32_t stride_bytes, uint32_t pos_min, float v, int row) {
    sts_b32(val_addr + pos_min * stride_bytes, __float_as_uint(v));
    sts_b32(idx_addr + pos_min * stride_bytes, (uint32_t)row);
    float m = INFINITY;
    uint32_t pm = 0;
#pragma unroll 8
    for (int e = 0; e < kprime; ++e) {
        const float x = (uint32_t)e == pos_min ? v : lds_f32(val_addr + e * stride_bytes);
        if (x < m) { m = x; pm = (uint32_t)e; }
    }
    return ((unsigned long long)pm << 32) | (unsigned long long)__float_as_uint(m);
}

__device__ __forceinline__ float tc_max32(const float (&v)[32]) {
    float a0 = ptx::fmax3(v[0], v[1], v[2]), a1 = ptx::fmax3(v[3], v[4], v[5]), a2 = ptx::fmax3(v[6], v[7], v[8]);
    float a3 = ptx::fmax3(v[9], v[10], v[11]), a4 = ptx::fmax3(v[12], v[13], v[14]), a5 = ptx::fmax3(v[15], v[16], v[17]);
    float a6 = ptx::fmax3(v[18], v[19], v[20]), a7 = ptx::fmax3(v[21], v[22], v[23]), a8 = ptx::fmax3(v[24], v[25], v[26]);
    float a9 = ptx::fmax3(v[27], v[28], v[29]);
    float b0 = ptx::fmax3(a0, a1, a2), b1 = ptx::fmax3(a3, a4, a5), b2 = ptx::fmax3(a6, a7, a8);
    float b3 = ptx::fmax3(a9, v[30], v[31]);
    return fmaxf(ptx::fmax3(b0, b1, b2), b3);
}

__device__ __forceinline__ float tc_sel_below(float x, float cur) { return x < cur ? x : -INFINITY; }

__device__ __forceinline__ float tc_max8(float a0, float a1, float a2, float a3, float a4, float a5, float a6, float a7) {
    return ptx::fmax3(ptx::fmax3(a0, a1, a2), ptx::fmax3(a3, a4, a5), fmaxf(a6, a7));
}

__device__ __forceinline__ float fmin3(float a, float b, float c) {
    float d;
    asm("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}

template <int KP>
__device__ __forceinline__ float tc_list_min(const float (&lv)[KP]) {
    float m = lv[0];
#pragma unroll
    for (int e = 1; e + 1 < KP; e += 2) m = fmin3(m, lv[e], lv[e + 1]);
    if ((KP & 1) == 0) m = fminf(m, lv[KP - 1]);
    return m;
}

// Register-resident candidate list (KP = 16 or 32 entries per thread).
// Values are "packed": the low log2(KP) mantissa bits of a list entry hold its slot number, and the low 5 bits of an
// accumulator read from TMEM are replaced by its column inside the 32-column chunk, so a max / min tree returns the
// value and where it lives in one go.  The perturbation (<= 2^-18 relative) is part of the error budget K3 proves
// completeness with (TcBounds::c_mma).  All 32 lanes of a warp run the insert rounds together, so a chunk costs
// max-over-lanes (not sum-over-lanes) of the number of hits.
template <int KP>
struct TcRegList {
    static constexpr uint32_t kSlotMask = KP - 1;
    float lv[KP];
    int lr[KP];
    float thr;
    // empty slots hold -FLT_MAX, or a seed threshold proven elsewhere (values at or below it are never inserted)
    __device__ __forceinline__ void reset(float seed = -INFINITY) {
        const uint32_t base = seed > -3.0e38f ? __float_as_uint(seed) : 0xFF7FFFFFu;
#pragma unroll
        for (int e = 0; e < KP; ++e) { lv[e] = __uint_as_float((base & ~kSlotMask) | (uint32_t)e); lr[e] = -1; }
        thr = tc_list_min<KP>(lv);
    }
    __device__ __forceinline__ void insert(float cur, int row) {
        const uint32_t slot = __float_as_uint(thr) & kSlotMask;
        const float nv = __uint_as_float((__float_as_uint(cur) & ~kSlotMask) | slot);
#pragma unroll
        for (int e = 0; e < KP; ++e)
            if (slot == (uint32_t)e) { lv[e] = nv; lr[e] = row; }
        thr = tc_list_min<KP>(lv);
    }
    // one group of 8 columns: only entered (warp-uniformly) when some lane has a value above its threshold there
    template <int G>
    __device__ __forceinline__ void scan_group(const float (&v)[32], int row0) {
        float w[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) w[j] = __uint_as_float((__float_as_uint(v[8 * G + j]) & ~7u) | (uint32_t)j);
        float cur = tc_max8(w[0], w[1], w[2], w[3], w[4], w[5], w[6], w[7]);
        while (__any_sync(0xffffffffu, cur > thr)) {
            if (cur > thr) insert(cur, row0 + 8 * G + (int)(__float_as_uint(cur) & 7u));
            cur = tc_max8(tc_sel_below(w[0], cur), tc_sel_below(w[1], cur), tc_sel_below(w[2], cur),
                          tc_sel_below(w[3], cur), tc_sel_below(w[4], cur), tc_sel_below(w[5], cur),
                          tc_sel_below(w[6], cur), tc_sel_below(w[7], cur));
        }
    }
    __device__ __forceinline__ void scan(const float (&v)[32], int row0) {
        const float g0 = tc_max8(v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7]);
        const float g1 = tc_max8(v[8], v[9], v[10], v[11], v[12], v[13], v[14], v[15]);
        const float g2 = tc_max8(v[16], v[17], v[18], v[19], v[20], v[21], v[22], v[23]);
        const float g3 = tc_max8(v[24], v[25], v[26], v[27], v[28], v[29], v[30], v[31]);
        const float mx = fmaxf(fmaxf(g0, g1), fmaxf(g2, g3));
        if (__any_sync(0xffffffffu, mx > thr)) {
            if (__any_sync(0xffffffffu, g0 > thr)) scan_group<0>(v, row0);
            if (__any_sync(0xffffffffu, g1 > thr)) scan_group<1>(v, row0);
            if (__any_sync(0xffffffffu, g2 > thr)) scan_group<2>(v, row0);
            if (__any_sync(0xffffffffu, g3 > thr)) scan_group<3>(v, row0);
        }
    }
};

template <int KP>
__global__ void __launch_bounds__(kTcThreads, 1)
knn_tc_filter_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                     const TcParams p) {
    extern __shared__ __align__(1024) unsigned char smem[];
    // carve: [A blocks][B ring][lists][barriers]
    unsigned char* a_smem = smem;
    unsigned char* b_smem = a_smem + (p.a_resident ? (size_t)p.nkb * kTcABlockBytes : 0);
    const uint32_t ring_stage_bytes = p.a_resident ? kTcBStageBytes : kTcABlockBytes + kTcBStageBytes;
    const uint32_t ring_b_offset = p.a_resident ? 0 : kTcABlockBytes;      // streamed mode: [A block][B block]
    const int epi_threads = p.epi_warps * 32;
    const size_t list_elems = KP > 0 ? 0 : (size_t)p.kprime * epi_threads;   // register lists need no shared memory
    float* list_val = reinterpret_cast<float*>(b_smem + (size_t)p.stages * ring_stage_bytes);
    int* list_idx = reinterpret_cast<int*>(list_val + list_elems);
    uint64_t* bars = reinterpret_cast<uint64_t*>(list_idx + list_elems);
    uint64_t* full_bar = bars;                    // [stages]
    uint64_t* empty_bar = bars + 8;               // [stages]
    uint64_t* tmem_full = bars + 16;              // [2]
    uint64_t* tmem_empty = bars + 18;             // [2]
    uint64_t* a_full = bars + 20;
    uint64_t* a_empty = bars + 21;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 22);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nitems = p.nqt * p.nseg;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tensormap(&map_a);
        ptx::prefetch_tensormap(&map_b);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < p.stages; ++s) { ptx::mbar_init(&full_bar[s], 1); ptx::mbar_init(&empty_bar[s], 1); }
        for (int s = 0; s < 2; ++s) { ptx::mbar_init(&tmem_full[s], 1); ptx::mbar_init(&tmem_empty[s], p.epi_warps); }
        ptx::mbar_init(a_full, 1);
        ptx::mbar_init(a_empty, 1);
        ptx::fence_barrier_init();
    }
    if (warp == 2) ptx::tmem_alloc(tmem_slot, 512);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            // ===== TMA producer =====
            uint32_t stage = 0, phase = 0, it = 0;
            for (int item = blockIdx.x; item < nitems; item += gridDim.x, ++it) {
                const int seg = item / p.nqt, qt = item % p.nqt;
                const int tile0 = seg * p.tiles_per_seg;
                const int ntiles = min(p.tiles_per_seg, p.tiles_total - tile0);
                if (p.a_resident) {
                    ptx::mbar_wait(a_empty, (it & 1) ^ 1);
                    ptx::mbar_expect_tx(a_full, (uint32_t)p.nkb * kTcABlockBytes);
                    for (int kb = 0; kb < p.nkb; ++kb)
                        ptx::tma_load_2d(a_smem + (size_t)kb * kTcABlockBytes, &map_a, a_full, kb * kTcKB, qt * kTcBM);
                }
                for (int t = 0; t < ntiles; ++t) {
                    const int row0 = (tile0 + t) * kTcBN;
                    for (int kb = 0; kb < p.nkb; ++kb) {
                        ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
                        unsigned char* ring = b_smem + (size_t)stage * ring_stage_bytes;
                        if (p.debug == 3) {                      // ablation: no operand traffic, MMA on stale smem
                            ptx::mbar_arrive(&full_bar[stage]);
                        } else {
                            ptx::mbar_expect_tx(&full_bar[stage], ring_stage_bytes);
                            if (!p.a_resident) ptx::tma_load_2d(ring, &map_a, &full_bar[stage], kb * kTcKB, qt * kTcBM);
                            ptx::tma_load_2d(ring + ring_b_offset, &map_b, &full_bar[stage], kb * kTcKB, row0);
                        }
                        if (++stage == (uint32_t)p.stages) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ===== MMA issuer =====
            const uint32_t idesc = ptx::umma_idesc_bf16(kTcBM, kTcBN);
            const uint32_t hi = ptx::umma_desc_hi_k_sw128();
            const uint32_t a_lo0 = ptx::umma_desc_lo(ptx::smem_u32(a_smem)), b_lo0 = ptx::umma_desc_lo(ptx::smem_u32(b_smem));
            uint32_t stage = 0, phase = 0, it = 0, acc_it = 0;
            for (int item = blockIdx.x; item < nitems; item += gridDim.x, ++it) {
                const int seg = item / p.nqt;
                const int tile0 = seg * p.tiles_per_seg;
                const int ntiles = min(p.tiles_per_seg, p.tiles_total - tile0);
                if (p.a_resident) {
                    ptx::mbar_wait(a_full, it & 1);
                    ptx::tc_fence_after();
                }
                for (int t = 0; t < ntiles; ++t, ++acc_it) {
                    const uint32_t as = acc_it & 1, aphase = (acc_it >> 1) & 1;
                    ptx::mbar_wait(&tmem_empty[as], aphase ^ 1);
                    ptx::tc_fence_after();
                    const uint32_t tmem_d = tmem_base + as * kTcBN;
                    for (int kb = 0; kb < p.nkb; ++kb) {
                        ptx::mbar_wait(&full_bar[stage], phase);
                        ptx::tc_fence_after();
                        const int nks = p.ksteps - kb * 4;
                        const uint32_t ring_lo = b_lo0 + stage * (ring_stage_bytes >> 4);
                        const uint32_t a_lo = p.a_resident ? a_lo0 + kb * (kTcABlockBytes >> 4) : ring_lo;
                        const uint32_t b_lo = ring_lo + (ring_b_offset >> 4);
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks)
                            if (ks < nks)
                                ptx::umma_bf16_ss_lohi(tmem_d, a_lo + ks * 2, b_lo + ks * 2, hi, idesc,
                                                       (uint32_t)((kb | ks) != 0));
                        ptx::umma_commit(&empty_bar[stage]);       // frees the smem stage when these MMAs retire
                        if (++stage == (uint32_t)p.stages) { stage = 0; phase ^= 1; }
                    }
                    ptx::umma_commit(&tmem_full[as]);              // accumulator tile ready for the epilogue
                }
                if (p.a_resident) ptx::umma_commit(a_empty);       // query tile may be overwritten
            }
        }
    } else if (warp >= 4 && warp < 4 + p.epi_warps) {
        // ===== epilogue: TMEM -> registers -> per-query running top-K' =====
        const int ew = warp - 4;
        const int quarter = ew & 3;                     // == warp % 4: the TMEM lane quarter this warp may read
        const int half = ew >> 2;                       // column half when 8 warps share the tile
        const int halves = p.epi_warps >> 2;
        const int cols = kTcBN / halves;                // columns of each tile this thread scans
        const int et = ew * 32 + lane;                  // epilogue thread id
        const int m = quarter * 32 + lane;              // query row inside the tile == TMEM lane
        uint32_t acc_it = 0;
        if constexpr (KP > 0) {
            TcRegList<KP> list;
            for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
                const int seg = item / p.nqt, qt = item % p.nqt;
                const int tile0 = seg * p.tiles_per_seg;
                const int ntiles = min(p.tiles_per_seg, p.tiles_total - tile0);
                list.reset();
                if (p.debug == 4) list.thr = INFINITY;   // ablation: fast path only
                for (int t = 0; t < ntiles; ++t, ++acc_it) {
                    const uint32_t as = acc_it & 1, aphase = (acc_it >> 1) & 1;
                    ptx::mbar_wait(&tmem_full[as], aphase);
                    ptx::tc_fence_after();
                    const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + as * kTcBN + half * cols;
                    const int row_base = (tile0 + t) * kTcBN + half * cols;
                    if (p.debug != 1 && p.debug != 3) {
                        if constexpr (KP <= 16) {
                            float va[32], vb[32];
                            ptx::tmem_ld_32x32(taddr, va);
                            for (int c = 0; c < cols; c += 64) {   // two chunks per trip, loads one chunk ahead
                                ptx::tmem_ld_wait();
                                ptx::tmem_ld_32x32(taddr + c + 32, vb);
                                if (p.debug != 2) list.scan(va, row_base + c);
                                else if (va[0] + va[31] == 12345.678f) list.thr = va[0];
                                ptx::tmem_ld_wait();
                                if (c + 64 < cols) ptx::tmem_ld_32x32(taddr + c + 64, va);
                                if (p.debug != 2) list.scan(vb, row_base + c + 32);
                                else if (vb[0] + vb[31] == 12345.678f) list.thr = vb[0];
                            }
                        } else {
                            for (int c = 0; c < cols; c += 32) {
                                float va[32];
                                ptx::tmem_ld_32x32(taddr + c, va);
                                ptx::tmem_ld_wait();
                                list.scan(va, row_base + c);
                            }
                        }
                    }
                    ptx::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) ptx::mbar_arrive(&tmem_empty[as]);
                }
                const size_t q = (size_t)qt * kTcBM + m;
                const size_t slot = (size_t)seg * halves + half;
                float* ov = p.cand_val + (q * p.nslot + slot) * KP;
                int* oi = p.cand_idx + (q * p.nslot + slot) * KP;
                bool full = true;
#pragma unroll
                for (int e = 0; e < KP; ++e) { ov[e] = list.lv[e]; oi[e] = list.lr[e]; full = full && list.lr[e] >= 0; }
                p.cand_thr[q * p.nslot + slot] = full ? list.thr : -INFINITY;
            }
        } else {
            const uint32_t stride_bytes = (uint32_t)epi_threads * 4;
            const uint32_t val_addr = ptx::smem_u32(list_val + et), idx_addr = ptx::smem_u32(list_idx + et);
            for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
                const int seg = item / p.nqt, qt = item % p.nqt;
                const int tile0 = seg * p.tiles_per_seg;
                const int ntiles = min(p.tiles_per_seg, p.tiles_total - tile0);
                for (int e = 0; e < p.kprime; ++e) {
                    sts_b32(val_addr + e * stride_bytes, __float_as_uint(-INFINITY));
                    sts_b32(idx_addr + e * stride_bytes, 0xffffffffu);
                }
                float thr = -INFINITY;
                uint32_t pos_min = 0;
                auto scan = [&](const float (&v)[32], int row0) {
                    const float mx = tc_max32(v);
                    if (mx > thr) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            if (v[i] > thr) {
                                const unsigned long long r =
                                    tc_list_insert(val_addr, idx_addr, p.kprime, stride_bytes, pos_min, v[i], row0 + i);
                                thr = __uint_as_float((uint32_t)r);
                                pos_min = (uint32_t)(r >> 32);
                            }
                        }
                    }
                };
                for (int t = 0; t < ntiles; ++t, ++acc_it) {
                    const uint32_t as = acc_it & 1, aphase = (acc_it >> 1) & 1;
                    ptx::mbar_wait(&tmem_full[as], aphase);
                    ptx::tc_fence_after();
                    const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + as * kTcBN + half * cols;
                    const int row_base = (tile0 + t) * kTcBN + half * cols;
                    float va[32], vb[32];
                    if (p.debug != 1 && p.debug != 3) {
                        ptx::tmem_ld_32x32(taddr, va);
                        for (int c = 0; c < cols; c += 64) {      // two 32-column chunks per trip, loads one chunk ahead
                            ptx::tmem_ld_wait();
                            ptx::tmem_ld_32x32(taddr + c + 32, vb);
                            if (p.debug != 2) scan(va, row_base + c);
                            else if (va[0] + va[31] == 12345.678f) thr = va[0];
                            ptx::tmem_ld_wait();
                            if (c + 64 < cols) ptx::tmem_ld_32x32(taddr + c + 64, va);
                            if (p.debug != 2) scan(vb, row_base + c + 32);
                            else if (vb[0] + vb[31] == 12345.678f) thr = vb[0];
                        }
                    }
                    ptx::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) ptx::mbar_arrive(&tmem_empty[as]);
                }
                // publish this (query, slot) list and the bound on everything it dropped
                const size_t q = (size_t)qt * kTcBM + m;
                const size_t slot = (size_t)seg * halves + half;
                float* ov = p.cand_val + (q * p.nslot + slot) * p.kprime;
                int* oi = p.cand_idx + (q * p.nslot + slot) * p.kprime;
                for (int e = 0; e < p.kprime; ++e) {
                    ov[e] = lds_f32(val_addr + e * stride_bytes);
                    oi[e] = (int)__float_as_uint(lds_f32(idx_addr + e * stride_bytes));
                }
                p.cand_thr[q * p.nslot + slot] = thr;
            }
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, 512);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// K1, CTA-pair variant (cta_group::2): two CTAs of a cluster (= the two SMs of a TPC) work on one item.  Each keeps
// its own 128-query tile, loads HALF of every 256-row train tile, and the leader issues one M=256 x N=256 UMMA per K
// step that reads A and B from both SMs.  Per SM this halves the L2 -> shared-memory operand traffic and the
// shared-memory bytes the tensor core reads per FLOP, which is what limits the single-CTA kernel.
// Barrier topology (as in CUTLASS / DeepGEMM 2-SM kernels): TMA bytes of both CTAs are credited to the LEADER's
// full / a_full barriers; tcgen05.commit multicasts to the empty / tmem_full / a_empty barriers of both CTAs; the
// epilogue threads of both CTAs arrive on the leader's tmem_empty.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kTcBHalfBytes = (kTcBN / 2) * 128;      // 16 KB: this CTA's half of one K block of a train tile
// a ring stage of the pair kernel holds ALL K blocks of this CTA's half tile (nkb * 16 KB): one barrier round trip and
// one wait per tile keep the single MMA-issuing thread off the critical path

template <int KP>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kTcThreads, 1)
knn_tc_filter_pair_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                          const TcParams p) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* a_smem = smem;
    unsigned char* b_smem = a_smem + (size_t)p.nkb * kTcABlockBytes;
    const uint32_t stage_bytes = (uint32_t)p.nkb * kTcBHalfBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(b_smem + (size_t)p.stages * stage_bytes);
    uint64_t* full_bar = bars;                    // [stages]  (used in the leader)
    uint64_t* empty_bar = bars + 8;               // [stages]
    uint64_t* tmem_full = bars + 16;              // [2]
    uint64_t* tmem_empty = bars + 18;             // [2]       (used in the leader)
    uint64_t* a_full = bars + 20;                 //           (used in the leader)
    uint64_t* a_empty = bars + 21;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 22);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = ptx::cluster_ctarank();
    const bool leader = rank == 0;
    const int epi_threads = p.epi_warps * 32;
    const int nqp = p.nqt / 2;                    // query-tile pairs (host pads nqt to even)
    const int nitems = nqp * p.nseg;
    const int cluster_id = blockIdx.x >> 1, nclusters = gridDim.x >> 1;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tensormap(&map_a);
        ptx::prefetch_tensormap(&map_b);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < p.stages; ++s) { ptx::mbar_init(&full_bar[s], 2); ptx::mbar_init(&empty_bar[s], 1); }
        for (int s = 0; s < 2; ++s) { ptx::mbar_init(&tmem_full[s], 1); ptx::mbar_init(&tmem_empty[s], 2 * p.epi_warps); }
        ptx::mbar_init(a_full, 2);
        ptx::mbar_init(a_empty, 1);
        ptx::fence_barrier_init();
    }
    if (warp == 2) ptx::tmem_alloc_pair(tmem_slot, 512);
    ptx::tc_fence_before();
    ptx::cluster_sync_all();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            // ===== TMA producer (both CTAs): own query tile, own half of each train tile =====
            uint32_t stage = 0, phase = 0, it = 0;
            for (int item = cluster_id; item < nitems; item += nclusters, ++it) {
                const int seg = item / nqp, qt = 2 * (item % nqp) + (int)rank;
                const int tile0 = seg * p.tiles_per_seg;
                const int ntiles = min(p.tiles_per_seg, p.tiles_total - tile0);
                ptx::mbar_wait(a_empty, (it & 1) ^ 1);
                if (leader) ptx::mbar_expect_tx(a_full, 2u * (uint32_t)p.nkb * kTcABlockBytes);
                else ptx::mbar_arrive_remote(a_full, 0);
                for (int kb = 0; kb < p.nkb; ++kb)
                    ptx::tma_load_2d_pair(a_smem + (size_t)kb * kTcABlockBytes, &map_a, a_full, kb * kTcKB, qt * kTcBM);
                for (int t = 0; t < ntiles; ++t) {
                    const int row0 = (tile0 + t) * kTcBN + (int)rank * (kTcBN / 2);
                    ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
                    if (p.debug == 3) {                          // ablation: no operand traffic
                        if (leader) ptx::mbar_arrive(&full_bar[stage]);
                        else ptx::mbar_arrive_remote(&full_bar[stage], 0);
                    } else {
                        if (leader) ptx::mbar_expect_tx(&full_bar[stage], 2u * stage_bytes);
                        else ptx::mbar_arrive_remote(&full_bar[stage], 0);
                        for (int kb = 0; kb < p.nkb; ++kb)
                            ptx::tma_load_2d_pair(b_smem + (size_t)stage * stage_bytes + (size_t)kb * kTcBHalfBytes,
                                                  &map_b, &full_bar[stage], kb * kTcKB, row0);
                    }
                    if (++stage == (uint32_t)p.stages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && leader) {
            // ===== MMA issuer (leader CTA only) =====
            const uint32_t idesc = ptx::umma_idesc_bf16(2 * kTcBM, kTcBN);
            const uint32_t hi = ptx::umma_desc_hi_k_sw128();
            const uint32_t a_lo0 = ptx::umma_desc_lo(ptx::smem_u32(a_smem)), b_lo0 = ptx::umma_desc_lo(ptx::smem_u32(b_smem));
            uint32_t stage = 0, phase = 0, it = 0, acc_it = 0;
            for (int item = cluster_id; item < nitems; item += nclusters, ++it) {
                const int seg = item / nqp;
                const int tile0 = seg * p.tiles_per_seg;
                const int ntiles = min(p.tiles_per_seg, p.tiles_total - tile0);
                ptx::mbar_wait(a_full, it & 1);
                ptx::tc_fence_after();
                for (int t = 0; t < ntiles; ++t, ++acc_it) {
                    const uint32_t as = acc_it & 1, aphase = (acc_it >> 1) & 1;
                    ptx::mbar_wait(&tmem_empty[as], aphase ^ 1);
                    ptx::mbar_wait(&full_bar[stage], phase);
                    ptx::tc_fence_after();
                    const uint32_t tmem_d = tmem_base + as * kTcBN;
                    const uint32_t b_lo_stage = b_lo0 + stage * (stage_bytes >> 4);
                    for (int kb = 0; kb < p.nkb; ++kb) {
                        const int nks = p.ksteps - kb * 4;
                        const uint32_t a_lo = a_lo0 + kb * (kTcABlockBytes >> 4);
                        const uint32_t b_lo = b_lo_stage + kb * (kTcBHalfBytes >> 4);
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks)
                            if (ks < nks)
                                ptx::umma_bf16_ss_pair_lohi(tmem_d, a_lo + ks * 2, b_lo + ks * 2, hi, idesc,
                                                            (uint32_t)((kb | ks) != 0));
                    }
                    ptx::umma_commit_pair(&empty_bar[stage]);      // both CTAs may refill this stage
                    ptx::umma_commit_pair(&tmem_full[as]);         // both CTAs' epilogues may read the tile
                    if (++stage == (uint32_t)p.stages) { stage = 0; phase ^= 1; }
                }
                ptx::umma_commit_pair(a_empty);
            }
        }
    } else if (warp >= 4 && warp < 4 + p.epi_warps) {
        // ===== epilogue (both CTAs): own 128 query rows x 256 columns of every accumulator tile =====
        const int ew = warp - 4;
        const int quarter = ew & 3, half = ew >> 2, halves = p.epi_warps >> 2;
        const int cols = kTcBN / halves;
        const int m = quarter * 32 + lane;
        uint32_t acc_it = 0;
        TcRegList<KP> list;
        for (int item = cluster_id; item < nitems; item += nclusters) {
            const int seg = item / nqp, qt = 2 * (item % nqp) + (int)rank;
            const int tile0 = seg * p.tiles_per_seg;
            const int ntiles = min(p.tiles_per_seg, p.tiles_total - tile0);
            // seed the threshold with what earlier segments of the same query already proved: their K'-th best
            // bounds this query's answer from below, so rows scoring under it cannot matter here either.  The slots
            // are written by other CTAs of this launch; an unfinished one still reads -inf (memset before launch).
            const size_t q = (size_t)qt * kTcBM + m;
            float seed = -INFINITY;
            for (int s2 = 0; s2 < seg; ++s2)
                seed = fmaxf(seed, *(const volatile float*)(p.cand_thr + q * p.nslot + (size_t)s2 * halves + half));
            list.reset(seed);
            if (p.debug == 4) list.thr = INFINITY;       // ablation: fast path only
            for (int t = 0; t < ntiles; ++t, ++acc_it) {
                const uint32_t as = acc_it & 1, aphase = (acc_it >> 1) & 1;
                ptx::mbar_wait(&tmem_full[as], aphase);
                ptx::tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + as * kTcBN + half * cols;
                const int row_base = (tile0 + t) * kTcBN + half * cols;
                bool released = false;
                if (p.debug != 1 && p.debug != 3) {
                    if constexpr (KP <= 16) {
                        // two register buffers; the accumulator stage is handed back to the MMA as soon as the last
                        // TMEM load has landed, the last two chunks are scanned after the release
                        float va[32], vb[32];
                        ptx::tmem_ld_32x32(taddr, va);
                        ptx::tmem_ld_32x32(taddr + 32, vb);
                        for (int c = 0; c < cols; c += 64) {
                            ptx::tmem_ld_wait();
                            const bool last = c + 64 >= cols;
                            if (last) {
                                ptx::tc_fence_before();
                                __syncwarp();
                                if (lane == 0) ptx::mbar_arrive_remote(&tmem_empty[as], 0);
                                released = true;
                            }
                            if (p.debug != 2) list.scan(va, row_base + c);
                            else if (va[0] + va[31] == 12345.678f) list.thr = va[0];
                            if (!last) ptx::tmem_ld_32x32(taddr + c + 64, va);
                            if (p.debug != 2) list.scan(vb, row_base + c + 32);
                            else if (vb[0] + vb[31] == 12345.678f) list.thr = vb[0];
                            if (!last) ptx::tmem_ld_32x32(taddr + c + 96, vb);
                        }
                    } else {
                        for (int c = 0; c < cols; c += 32) {
                            float va[32];
                            ptx::tmem_ld_32x32(taddr + c, va);
                            ptx::tmem_ld_wait();
                            list.scan(va, row_base + c);
                        }
                    }
                }
                if (!released) {
                    ptx::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) ptx::mbar_arrive_remote(&tmem_empty[as], 0);
                }
            }
            const size_t slot = (size_t)seg * halves + half;
            float* ov = p.cand_val + (q * p.nslot + slot) * KP;
            int* oi = p.cand_idx + (q * p.nslot + slot) * KP;
            bool full = true;
#pragma unroll
            for (int e = 0; e < KP; ++e) { ov[e] = list.lv[e]; oi[e] = list.lr[e]; full = full && list.lr[e] >= 0; }
            // everything this slot dropped scores <= thr (placeholders keep thr at the seed while the list is not full)
            p.cand_thr[q * p.nslot + slot] = (full || seed > -INFINITY) ? list.thr : -INFINITY;
        }
    }
    ptx::tc_fence_before();
    ptx::cluster_sync_all();            // the peer's remote arrives / multicast commits have landed before anyone exits
    if (warp == 2) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc_pair(tmem_base, 512);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// K3: exact rescoring + completeness proof, one warp per query
// ---------------------------------------------------------------------------------------------------------------------
struct TcBounds {
    double et_max;        // max over train rows of |t^ - (t - c)|
    double tnorm_max;     // max |t^|
    double c_mma;         // relative accumulation error of the tensor core sum, per unit of sum |terms|
    double gamma;         // relative error of the oracle's fp32 distance
};

constexpr int kRescoreWarps = 4;

// exact fp32 distance, 16-byte loads when the row length allows (same left-to-right order as sum_sequential)
__device__ __forceinline__ float tc_exact_distance(const float* __restrict__ t, const float* __restrict__ q, int d,
                                                   int sum_order) {
    if (sum_order == 1 && (d & 3) == 0) {
        using E = Elem<float>;
        const float4* t4 = reinterpret_cast<const float4*>(t);
        const float4* q4 = reinterpret_cast<const float4*>(q);
        float acc = 0.0f;
#pragma unroll 4
        for (int j = 0; j < d / 4; ++j) {
            const float4 a = __ldg(t4 + j), b = __ldg(q4 + j);
            acc = E::add(acc, E::sq_diff(a.x, b.x));
            acc = E::add(acc, E::sq_diff(a.y, b.y));
            acc = E::add(acc, E::sq_diff(a.z, b.z));
            acc = E::add(acc, E::sq_diff(a.w, b.w));
        }
        return E::finish(acc);
    }
    return exact_distance<float>(t, q, d, sum_order);
}

// One warp per query.  Exact distances are only computed where they can matter: first for the k+1 candidates with the
// best filter score, whose largest exact distance D0 caps the answer; any other candidate whose score-derived lower
// bound already exceeds D0 cannot be among the k+1 nearest and is never gathered from HBM.
__global__ void __launch_bounds__(32 * kRescoreWarps)
knn_rescore_kernel(const float* __restrict__ train, int64_t n, const float* __restrict__ queries, int d, int sum_order,
                   int64_t q_count, int ksel, const float* __restrict__ cand_val, const int* __restrict__ cand_idx,
                   const float* __restrict__ cand_thr, int nslot, int kprime, const double* __restrict__ qnorm2,
                   const double* __restrict__ qerr, TcBounds bounds, int64_t index_base, float* __restrict__ out_d,
                   int64_t* __restrict__ out_i, int* __restrict__ flags) {
    __shared__ float s_acc[kRescoreWarps][kTcMaxCand];
    __shared__ float s_dist[kRescoreWarps][kTcMaxCand];
    __shared__ int s_idx[kRescoreWarps][kTcMaxCand];
    __shared__ float s_best[kRescoreWarps];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t q = (int64_t)blockIdx.x * kRescoreWarps + warp;
    if (q >= q_count) return;
    const int ncand = nslot * kprime;
    const float* cv = cand_val + (size_t)q * ncand;
    const int* ci = cand_idx + (size_t)q * ncand;
    const float* qrow = queries + (size_t)q * d;
    float* acc = s_acc[warp];
    float* dist = s_dist[warp];
    int* idx = s_idx[warp];

    const double n2 = qnorm2[q], eq = qerr[q];
    const double eps = bounds.c_mma * (sqrt(n2) * bounds.tnorm_max + 0.5 * bounds.tnorm_max * bounds.tnorm_max);
    // every row whose filter score is <= a has an exact (oracle-arithmetic) distance >= lower(a)
    auto lower = [&](float a) {
        const double l2 = n2 - 2.0 * ((double)a + eps);
        return (sqrt(fmax(l2, 0.0)) - (eq + bounds.et_max)) * (1.0 - bounds.gamma);
    };

    int nvalid = 0;
    for (int c = lane; c < ncand; c += 32) {
        const int row = ci[c];
        const bool valid = row >= 0 && row < n;       // negative: empty slot; >= n: a padding row of the filter copy
        idx[c] = valid ? row : kIdxSentinel;
        acc[c] = valid ? cv[c] : -INFINITY;
        dist[c] = nan_of<float>();                    // "not computed": sorts after everything
        nvalid += valid ? 1 : 0;
    }
    for (int off = 16; off > 0; off >>= 1) nvalid += __shfl_xor_sync(0xffffffffu, nvalid, off);
    __syncwarp();
    bool ok = nvalid >= ksel;
    // pass 1: the k+1 best-scoring candidates (selection first, then the gathers, so ranks see a stable score array)
    uint32_t selected = 0;                            // bit i: candidate lane + 32 * i
    for (int c = lane, i = 0; c < ncand; c += 32, ++i) {
        if (idx[c] == kIdxSentinel) continue;
        const float ac = acc[c];
        int rank = 0;
        for (int o = 0; o < ncand; ++o) {
            const float ao = acc[o];
            rank += (ao > ac || (ao == ac && o < c)) ? 1 : 0;
        }
        if (rank < ksel) selected |= 1u << i;
    }
    float d0 = 0.0f;
    bool d0_nan = false;
    for (int c = lane, i = 0; c < ncand; c += 32, ++i) {
        if (!((selected >> i) & 1u)) continue;
        const float dc = tc_exact_distance(train + (size_t)idx[c] * d, qrow, d, sum_order);
        dist[c] = dc;
        d0 = fmaxf(d0, dc);
        d0_nan = d0_nan || dc != dc;
    }
    for (int off = 16; off > 0; off >>= 1) d0 = fmaxf(d0, __shfl_xor_sync(0xffffffffu, d0, off));
    d0_nan = __any_sync(0xffffffffu, d0_nan) || !ok;
    const double cap = d0_nan ? (double)INFINITY : (double)d0 * (1.0 + 9.5367431640625e-07);
    __syncwarp();
    // pass 2: whatever the cap cannot rule out
    for (int c = lane; c < ncand; c += 32) {
        if (idx[c] == kIdxSentinel || dist[c] == dist[c]) continue;         // empty, or already computed in pass 1
        if (lower(acc[c]) > cap) { idx[c] = kIdxSentinel; continue; }       // provably beyond the (k+1)-th nearest
        dist[c] = tc_exact_distance(train + (size_t)idx[c] * d, qrow, d, sum_order);
    }
    __syncwarp();
    // rank the computed candidates by (distance, index); keys are unique because the slots cover disjoint row ranges
    for (int c = lane; c < ncand; c += 32) {
        const int ic = idx[c];
        if (ic == kIdxSentinel) continue;
        const float dc = dist[c];
        int rank = 0;
        for (int o = 0; o < ncand; ++o) rank += (idx[o] != kIdxSentinel && key_less(dist[o], idx[o], dc, ic)) ? 1 : 0;
        if (rank < ksel) {
            out_d[q * ksel + rank] = dc;
            out_i[q * ksel + rank] = index_base + ic;
            if (rank == ksel - 1) s_best[warp] = dc;
        }
    }
    __syncwarp();
    if (ok) {
        const double dk = (double)s_best[warp] * (1.0 + 9.5367431640625e-07);   // (k+1)-th exact distance, 2^-20 margin
        for (int s = lane; s < nslot; s += 32) {
            const float thr = cand_thr[(size_t)q * nslot + s];
            if (thr == -INFINITY) continue;   // the slot never filled its list: nothing was dropped
            if (!(lower(thr) > dk)) ok = false;    // also catches NaN
        }
    }
    ok = __all_sync(0xffffffffu, ok);
    if (lane == 0) flags[q] = ok ? 0 : 1;
}

// ---- fallback plumbing ------------------------------------------------------------------------------------------------
__global__ void tc_compact_flags_kernel(const int* __restrict__ flags, int64_t q_count, int* __restrict__ list,
                                        int* __restrict__ count) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q < q_count && flags[q]) list[atomicAdd(count, 1)] = (int)q;
}

__global__ void tc_gather_rows_kernel(const float* __restrict__ src, const int* __restrict__ list, int count, int d,
                                      float* __restrict__ dst) {
    const int r = blockIdx.x;
    if (r >= count) return;
    const float* s = src + (size_t)list[r] * d;
    for (int j = threadIdx.x; j < d; j += blockDim.x) dst[(size_t)r * d + j] = s[j];
}

__global__ void tc_scatter_lists_kernel(const float* __restrict__ src_d, const int64_t* __restrict__ src_i,
                                        const int* __restrict__ list, int count, int ksel, float* __restrict__ dst_d,
                                        int64_t* __restrict__ dst_i) {
    const int r = blockIdx.x;
    if (r >= count) return;
    const size_t q = (size_t)list[r];
    for (int e = threadIdx.x; e < ksel; e += blockDim.x) {
        dst_d[q * ksel + e] = src_d[(size_t)r * ksel + e];
        dst_i[q * ksel + e] = src_i[(size_t)r * ksel + e];
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------------
struct TcBuffer {
    void* ptr = nullptr; size_t cap = 0;
    cudaError_t reserve(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (ptr) cudaFree(ptr);
        ptr = nullptr; cap = 0;
        cudaError_t e = cudaMalloc(&ptr, bytes);
        if (e == cudaSuccess) cap = bytes;
        return e;
    }
    void release() { if (ptr) cudaFree(ptr); ptr = nullptr; cap = 0; }
    template <typename T> T* as() const { return reinterpret_cast<T*>(ptr); }
};

struct TensorState {
    bool ready = false;
    bool usable = false;          // false if the train set has non-finite rows (then only K2 is exact)
    int64_t npad = 0; int dp = 0;
    double et_max = 0.0, tnorm_max = 0.0;
    TcBuffer train_bf16, center, stats, partial;
    TcBuffer q_bf16, q_norm2, q_err, cand_val, cand_idx, cand_thr, flags, fb_list, fb_count, fb_queries, fb_d, fb_i;
    CUtensorMap map_b;
    CUtensorMap map_b_half;   // 128-row boxes for the CTA-pair kernel
    void* encode_fn = nullptr;
    int sm_count = 0;
    int smem_optin = 0;
    cudaEvent_t ev_k1_begin = nullptr, ev_k1_end = nullptr;
    double filter_ms = 0.0;       // device time of the K1 launches of the last neighbour search (CUDA events)
    int64_t filter_launches = 0;
};

struct TensorLaunch {
    const void* train; int dtype; int64_t n, d; int k; int sum_order; int64_t index_base; int slack;
    int opt_epi_warps = 0;   // 0 = auto, else 4 or 8
    int opt_nseg = 0;        // 0 = auto
    int opt_smem_lists = 0;  // 1 = force the shared-memory list epilogue
    int opt_cta_group = 0;   // 0 = auto, 1 = single CTA per item, 2 = CTA pairs (cta_group::2)
    int debug = 0;
};

using DirectFallback = std::function<int(const void*, int64_t, void*, int64_t*)>;

inline int tc_default_slack(int k) { return std::max(5, (k + 1) / 3); }

// List length per (query, slot).  With register lists two epilogue threads share a query row (one per column half), so
// a segment keeps 2 * K' candidates; each half must still be able to hold its share of the true k+1 nearest, which is
// Binomial(k+1, 1/2): K' >= (k+1)/2 + 2*sqrt(k+1) keeps the overflow probability below 1e-4.
inline int tc_kprime(int k, int slack) {
    if (slack < 0) slack = tc_default_slack(k);
    const int need = k + 1 + slack;
    const double half_share = 0.5 * (k + 1) + 2.0 * std::sqrt((double)(k + 1));
    if (need <= 32 && half_share <= 16.0) return 16;      // register lists, 16 per half
    if (need <= 64 && half_share <= 32.0) return 32;      // register lists, 32 per half
    return (need + 7) / 8 * 8;                            // shared-memory lists, one thread per query row
}

inline bool tensor_path_supported(int dtype, int64_t n, int64_t d, int k) {
    if (dtype != 0 /* OSML_KNN_F32 */) return false;
    if (n < 256 || d < 1) return false;
    if (d > 8189) return false;                   // 128 K blocks of 64 bf16
    return tc_kprime(k, -1) <= kTcMaxKPrime;
}

inline size_t tensor_bytes(const TensorState& s) {
    return s.train_bf16.cap + s.center.cap + s.stats.cap + s.partial.cap + s.q_bf16.cap + s.q_norm2.cap + s.q_err.cap +
           s.cand_val.cap + s.cand_idx.cap + s.cand_thr.cap + s.flags.cap + s.fb_list.cap + s.fb_count.cap + s.fb_queries.cap +
           s.fb_d.cap + s.fb_i.cap;
}

inline void tensor_release(TensorState& s) {
    for (TcBuffer* b : {&s.train_bf16, &s.center, &s.stats, &s.partial, &s.q_bf16, &s.q_norm2, &s.q_err, &s.cand_val,
                        &s.cand_idx, &s.cand_thr, &s.flags, &s.fb_list, &s.fb_count, &s.fb_queries, &s.fb_d, &s.fb_i})
        b->release();
    if (s.ev_k1_begin) cudaEventDestroy(s.ev_k1_begin);
    if (s.ev_k1_end) cudaEventDestroy(s.ev_k1_end);
    s.ev_k1_begin = s.ev_k1_end = nullptr;
    s.ready = false;
}

typedef CUresult (*TcEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline int tc_make_map(TensorState& st, CUtensorMap* map, void* base, int64_t rows, int dp, int box_rows,
                       std::string* err) {
    const cuuint64_t dims[2] = {(cuuint64_t)dp, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)dp * 2};
    const cuuint32_t box[2] = {(cuuint32_t)kTcKB, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = reinterpret_cast<TcEncodeTiledFn>(st.encode_fn)(
        map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        if (err) *err = "cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r);
        return 2;
    }
    return 0;
}

#define TC_TRY(expr)                                                                           \
    do {                                                                                       \
        cudaError_t e__ = (expr);                                                              \
        if (e__ != cudaSuccess) {                                                              \
            cudaGetLastError();                                                                \
            if (err) *err = std::string(#expr) + ": " + cudaGetErrorString(e__);               \
            return e__ == cudaErrorMemoryAllocation ? 3 : 2;                                   \
        }                                                                                      \
    } while (0)

// one-time K0 over the train set
inline int tc_prepare(TensorState& st, const TensorLaunch& tl, cudaStream_t stream, int64_t* launches,
                      std::string* err) {
    if (st.ready) return 0;
    int dev = 0;
    TC_TRY(cudaGetDevice(&dev));
    TC_TRY(cudaDeviceGetAttribute(&st.sm_count, cudaDevAttrMultiProcessorCount, dev));
    TC_TRY(cudaDeviceGetAttribute(&st.smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    cudaDriverEntryPointQueryResult qres;
    TC_TRY(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &st.encode_fn, cudaEnableDefault, &qres));
    if (!st.encode_fn || qres != cudaDriverEntryPointSuccess) {
        if (err) *err = "cuTensorMapEncodeTiled is not available from the driver";
        return 2;
    }
    const int d = (int)tl.d;
    st.dp = (d + 3 + kTcKB - 1) / kTcKB * kTcKB;
    st.npad = (tl.n + kTcBN - 1) / kTcBN * kTcBN;
    const float* train = static_cast<const float*>(tl.train);
    const int rows_per_block = 2048;
    const int nblocks = (int)((tl.n + rows_per_block - 1) / rows_per_block);
    TC_TRY(st.partial.reserve((size_t)nblocks * d * sizeof(double)));
    TC_TRY(st.center.reserve((size_t)d * sizeof(float)));
    TC_TRY(st.stats.reserve(sizeof(TcPrepStats)));
    TC_TRY(st.train_bf16.reserve((size_t)st.npad * st.dp * sizeof(__nv_bfloat16)));
    TC_TRY(cudaMemsetAsync(st.stats.ptr, 0, sizeof(TcPrepStats), stream));
    tc_col_partial_sums_kernel<<<nblocks, dim3(32, 8), 0, stream>>>(train, tl.n, d, rows_per_block,
                                                                    st.partial.as<double>());
    tc_col_mean_kernel<<<(d + 127) / 128, 128, 0, stream>>>(st.partial.as<double>(), nblocks, d, tl.n,
                                                            st.center.as<float>());
    const int64_t threads = st.npad * 32;
    tc_prep_rows_kernel<true><<<(unsigned)((threads + 255) / 256), 256, 0, stream>>>(
        train, tl.n, st.npad, d, st.dp, st.center.as<float>(), st.train_bf16.as<__nv_bfloat16>(), nullptr, nullptr,
        st.stats.as<TcPrepStats>());
    *launches += 3;
    TC_TRY(cudaGetLastError());
    TcPrepStats hs;
    TC_TRY(cudaMemcpyAsync(&hs, st.stats.ptr, sizeof(hs), cudaMemcpyDeviceToHost, stream));
    TC_TRY(cudaStreamSynchronize(stream));
    st.partial.release();
    double et, tn;
    static_assert(sizeof(double) == sizeof(unsigned long long), "double bits");
    memcpy(&et, &hs.et_max_bits, 8);
    memcpy(&tn, &hs.tnorm_max_bits, 8);
    st.et_max = et; st.tnorm_max = tn;
    st.usable = hs.nonfinite == 0;
    if (st.usable) {
        int rc = tc_make_map(st, &st.map_b, st.train_bf16.ptr, st.npad, st.dp, kTcBN, err);
        if (rc) return rc;
        rc = tc_make_map(st, &st.map_b_half, st.train_bf16.ptr, st.npad, st.dp, kTcBN / 2, err);
        if (rc) return rc;
    }
    st.ready = true;
    return 0;
}

// neighbour search for one block of queries (device, row-major fp32).  out_d/out_i: [q][k+1] final lists.
inline int tc_neighbours_block(TensorState& st, const TensorLaunch& tl, const float* d_queries, int64_t q,
                               float* out_d, int64_t* out_i, cudaStream_t stream, int64_t* fallback,
                               int64_t* launches, std::string* err, const DirectFallback& direct) {
    const int d = (int)tl.d;
    const int ksel = tl.k + 1;
    const int kprime = tc_kprime(tl.k, tl.slack);
    if (kprime > kTcMaxKPrime) { if (err) *err = "k + 1 + slack exceeds 64"; return 4; }
    const bool reg_lists = kprime <= 32 && !tl.opt_smem_lists;
    int epi_warps = reg_lists ? 8 : 4;
    if (tl.opt_epi_warps == 4 || (tl.opt_epi_warps == 8 && (reg_lists || kprime <= 24))) epi_warps = tl.opt_epi_warps;
    const int halves = epi_warps / 4;
    // CTA pairs (cta_group::2) need register lists (no shared-memory list code in that kernel)
    // (the pair kernel stages whole half-tiles; for wide rows that no longer fits twice, so use single CTAs there)
    const size_t pair_stage = (size_t)(st.dp / kTcKB) * kTcBHalfBytes;
    const bool pair_fits = (size_t)st.smem_optin >= (size_t)(st.dp / kTcKB) * kTcABlockBytes + 2 * pair_stage + 1280;
    const int tc_cta_group = (reg_lists && tl.opt_cta_group != 1 && pair_fits) ? 2 : 1;
    const int64_t qtile = (int64_t)kTcBM * tc_cta_group;
    const int64_t qpad = (q + qtile - 1) / qtile * qtile;
    const int nqt = (int)(qpad / kTcBM);
    const int tiles_total = (int)(st.npad / kTcBN);
    const int nkb = st.dp / kTcKB;
    const int ksteps = (d + 3 + 15) / 16;

    // every extra segment restarts the per-query threshold (more list inserts, more candidates for K3), so use the
    // fewest that still give every SM an item; with more query tiles than SMs a single sweep per tile is best
    int max_seg_by_cand = std::max(1, kTcMaxCand / (kprime * halves));
    int nseg = 1;
    {
        // work items = nqt * nseg over sm_count persistent CTAs: pick the split with the least idle tail, charging
        // 1 % per extra segment (later segments start from the thresholds earlier ones proved, so they are cheap)
        double best = 1e30;
        const int units = std::max(1, st.sm_count / tc_cta_group);
        for (int cand = 1; cand <= 32; ++cand) {
            const double items = (double)(nqt / tc_cta_group) * cand;
            const double cost = std::ceil(items / units) / (items / units) + 0.01 * (cand - 1);
            if (cost < best - 1e-9) { best = cost; nseg = cand; }
        }
    }
    if (tl.opt_nseg > 0) nseg = tl.opt_nseg;
    nseg = std::max(1, std::min(nseg, std::min(max_seg_by_cand, std::max(1, tiles_total / 8))));
    int tiles_per_seg = (tiles_total + nseg - 1) / nseg;
    nseg = (tiles_total + tiles_per_seg - 1) / tiles_per_seg;
    const int nslot = nseg * halves;

    const size_t list_bytes = reg_lists ? 0 : (size_t)kprime * epi_warps * 32 * 8;
    // the query tile stays resident when it leaves room for at least two ring stages; wide rows stream it instead
    const bool a_resident = tc_cta_group == 2 ||
        (size_t)st.smem_optin >= (size_t)nkb * kTcABlockBytes + list_bytes + 2 * (size_t)kTcBStageBytes + 1280;
    const size_t a_bytes = a_resident ? (size_t)nkb * kTcABlockBytes : 0;
    const size_t stage_bytes = tc_cta_group == 2 ? (size_t)nkb * kTcBHalfBytes
                               : (a_resident ? (size_t)kTcBStageBytes : (size_t)kTcABlockBytes + kTcBStageBytes);
    const size_t fixed = a_bytes + list_bytes + 256 + 1024;
    int stages = (int)(((size_t)st.smem_optin - fixed) / stage_bytes);
    stages = std::min(stages, tc_cta_group == 2 ? 8 : 6);
    if (stages < 2) { if (err) *err = "not enough shared memory for the tensor filter"; return 4; }
    const size_t smem = a_bytes + (size_t)stages * stage_bytes + list_bytes + 256;

    TC_TRY(st.q_bf16.reserve((size_t)qpad * st.dp * sizeof(__nv_bfloat16)));
    TC_TRY(st.q_norm2.reserve((size_t)qpad * sizeof(double)));
    TC_TRY(st.q_err.reserve((size_t)qpad * sizeof(double)));
    TC_TRY(st.cand_val.reserve((size_t)qpad * nslot * kprime * sizeof(float)));
    TC_TRY(st.cand_idx.reserve((size_t)qpad * nslot * kprime * sizeof(int)));
    TC_TRY(st.cand_thr.reserve((size_t)qpad * nslot * sizeof(float)));
    TC_TRY(st.flags.reserve((size_t)q * sizeof(int)));
    TC_TRY(st.fb_list.reserve((size_t)q * sizeof(int)));
    TC_TRY(st.fb_count.reserve(sizeof(int)));

    const int64_t threads = qpad * 32;
    tc_prep_rows_kernel<false><<<(unsigned)((threads + 255) / 256), 256, 0, stream>>>(
        d_queries, q, qpad, d, st.dp, st.center.as<float>(), st.q_bf16.as<__nv_bfloat16>(), st.q_norm2.as<double>(),
        st.q_err.as<double>(), nullptr);
    ++*launches;
    TC_TRY(cudaGetLastError());

    CUtensorMap map_a;
    int rc = tc_make_map(st, &map_a, st.q_bf16.ptr, qpad, st.dp, kTcBM, err);
    if (rc) return rc;

    TcParams p;
    p.nkb = nkb; p.ksteps = ksteps; p.nqt = nqt; p.nseg = nseg; p.tiles_per_seg = tiles_per_seg;
    p.tiles_total = tiles_total; p.kprime = kprime; p.epi_warps = epi_warps; p.stages = stages; p.nslot = nslot; p.a_resident = a_resident ? 1 : 0;
    p.cand_val = st.cand_val.as<float>(); p.cand_idx = st.cand_idx.as<int>();
    p.cand_thr = st.cand_thr.as<float>();
    p.debug = tl.debug;
    auto kernel = !reg_lists ? knn_tc_filter_kernel<0> : (kprime == 16 ? knn_tc_filter_kernel<16> : knn_tc_filter_kernel<32>);
    if (tc_cta_group == 2) kernel = kprime == 16 ? knn_tc_filter_pair_kernel<16> : knn_tc_filter_pair_kernel<32>;
    TC_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int nitems = (nqt / tc_cta_group) * nseg;
    const int grid = tc_cta_group * std::min(nitems, st.sm_count / tc_cta_group);
    if (!st.ev_k1_begin) { TC_TRY(cudaEventCreate(&st.ev_k1_begin)); TC_TRY(cudaEventCreate(&st.ev_k1_end)); }
    TC_TRY(cudaMemsetAsync(st.cand_thr.ptr, 0xFF, (size_t)qpad * nslot * sizeof(float), stream));   // NaN bits: "not written"
    TC_TRY(cudaEventRecord(st.ev_k1_begin, stream));
    kernel<<<grid, kTcThreads, smem, stream>>>(map_a, tc_cta_group == 2 ? st.map_b_half : st.map_b, p);
    TC_TRY(cudaEventRecord(st.ev_k1_end, stream));
    ++*launches;
    ++st.filter_launches;
    TC_TRY(cudaGetLastError());
    if (tl.debug) return 0;   // profiling ablations: results are meaningless, skip rescoring

    TcBounds b;
    b.et_max = st.et_max; b.tnorm_max = st.tnorm_max;
    // (K + 8) * 2^-22 accumulation + 2^-23 norm split + 2^-17 for the index bits packed into list values
    b.c_mma = (double)(ksteps * 16 + 8) * 2.384185791015625e-07 + 1.2e-7 + 7.62939453125e-06;
    b.gamma = (double)(d + 4) * 1.1920928955078125e-07;                      // (D + 4) * 2^-23
    TC_TRY(cudaMemsetAsync(st.fb_count.ptr, 0, sizeof(int), stream));
    knn_rescore_kernel<<<(unsigned)((q + kRescoreWarps - 1) / kRescoreWarps), 32 * kRescoreWarps, 0, stream>>>(
        static_cast<const float*>(tl.train), tl.n, d_queries, d, tl.sum_order, q, ksel, st.cand_val.as<float>(),
        st.cand_idx.as<int>(), st.cand_thr.as<float>(), nslot, kprime, st.q_norm2.as<double>(), st.q_err.as<double>(), b, tl.index_base, out_d,
        out_i, st.flags.as<int>());
    tc_compact_flags_kernel<<<(unsigned)((q + 255) / 256), 256, 0, stream>>>(st.flags.as<int>(), q, st.fb_list.as<int>(),
                                                                             st.fb_count.as<int>());
    *launches += 2;
    TC_TRY(cudaGetLastError());
    int count = 0;
    TC_TRY(cudaMemcpyAsync(&count, st.fb_count.ptr, sizeof(int), cudaMemcpyDeviceToHost, stream));
    TC_TRY(cudaStreamSynchronize(stream));
    {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, st.ev_k1_begin, st.ev_k1_end) == cudaSuccess) st.filter_ms += ms;
    }
    if (count > 0) {
        TC_TRY(st.fb_queries.reserve((size_t)count * d * sizeof(float)));
        TC_TRY(st.fb_d.reserve((size_t)count * ksel * sizeof(float)));
        TC_TRY(st.fb_i.reserve((size_t)count * ksel * sizeof(int64_t)));
        tc_gather_rows_kernel<<<count, 128, 0, stream>>>(d_queries, st.fb_list.as<int>(), count, d,
                                                         st.fb_queries.as<float>());
        ++*launches;
        rc = direct(st.fb_queries.ptr, count, st.fb_d.ptr, st.fb_i.as<int64_t>());
        if (rc) return rc;
        tc_scatter_lists_kernel<<<count, 64, 0, stream>>>(st.fb_d.as<float>(), st.fb_i.as<int64_t>(),
                                                          st.fb_list.as<int>(), count, ksel, out_d, out_i);
        ++*launches;
        TC_TRY(cudaGetLastError());
        *fallback += count;
    }
    return 0;
}

inline int tensor_neighbours(TensorState& st, const TensorLaunch& tl, const void* d_queries, int64_t q, void* out_d,
                             int64_t* out_i, cudaStream_t stream, int64_t* fallback, int64_t* launches,
                             std::string* err, const DirectFallback& direct) {
    int rc = tc_prepare(st, tl, stream, launches, err);
    if (rc) return rc;
    st.filter_ms = 0.0;
    st.filter_launches = 0;
    if (!st.usable) {                       // non-finite train rows: the filter's bounds do not hold
        *fallback += q;
        return direct(d_queries, q, out_d, out_i);
    }
    const int64_t block = 131072;
    const int ksel = tl.k + 1;
    for (int64_t q0 = 0; q0 < q; q0 += block) {
        const int64_t qb = std::min(block, q - q0);
        rc = tc_neighbours_block(st, tl, static_cast<const float*>(d_queries) + q0 * tl.d, qb,
                                 static_cast<float*>(out_d) + q0 * ksel, out_i + q0 * ksel, stream, fallback, launches,
                                 err, direct);
        if (rc) return rc;
    }
    return 0;
}

}  // namespace osml
