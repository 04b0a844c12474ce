// K0 / K1 / K3: tcgen05 tensor-core candidate filter with a provably complete exact rescoring.
//
// The reference computes sqrt(sum((t - q)^2)) for every (query, train row) pair in fp32 (predict.py:742) and keeps the
// k+1 smallest (predict.py:750).  Doing that arithmetic literally costs 3 un-fusable CUDA-core instructions per
// element (K2).  Here the tensor cores only *filter*:
//
//   K0  knn_prep_rows      centre by the train column means, round to bf16, h = 0.5*|t^|^2 per train row; the filter
//                          copy holds the rows sorted by h (perm maps positions back to train rows) and, per 256-row
//                          tile, an "h block" (h, its minimum per 8 rows / per half tile).  Per call: the queries'
//                          bf16 copy, |q^|^2 and the rounding residuals |q^ - (q - c)|.
//   K1  knn_tc_filter      persistent tcgen05 kernels: score[q][t] = q^ . t^ - h[t] (so |q^ - t^|^2 = |q^|^2 - 2*score)
//                          with TMA-staged, 128B-swizzled operand tiles and FP32 accumulators in TMEM; the epilogue
//                          reads each 128 x 256 accumulator tile straight out of TMEM and keeps, per query, the K'
//                          largest scores per (train segment, column half).  The distance matrix never leaves the SM.
//                          The tensor core only computes the D-long dot products (K = D, no padding columns for the
//                          norms): the hot path compares a max tree over the raw dot products with thr + min(h), and
//                          only the rare groups of 8 columns that pass subtract h element by element.
//                            knn_tc_filter_pair_kernel  CTA pairs (cta_group::2): resident query tile for D <= 192, a K-block
//                                                       granular ring (kWide) for wider rows
//                            knn_tc_filter_kernel       one CTA per tile: shared-memory lists for k > 21
//   K3  knn_rescore        per query: exact distances of the candidates in the oracle's arithmetic and summation
//                          order, top-(k+1) by (distance, index), and a completeness proof -- every row the filter
//                          dropped has score <= the slot's K'-th best, which bounds its true distance from below
//                          by sqrt(|q^|^2 - 2*(thr + eps)) - (Eq + Et); if that bound exceeds the exact (k+1)-th
//                          distance the result equals the brute-force one.  Queries that cannot be proven go through
//                          a second, finer-grained filter pass and, failing that, K2.  So the output is bit-identical
//                          to K2 / the oracle whatever the filter does.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <cub/device/device_radix_sort.cuh>
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <functional>
#include <string>
#include <vector>

#include "knn_common.cuh"
#include "knn_ptx.cuh"

// -DOSML_TC_COUNTERS=1 builds the instrumented variant of the CTA-pair kernel (tools/ablate.sh); the product has none
#ifndef OSML_TC_COUNTERS
#define OSML_TC_COUNTERS 0
#endif
// -DOSML_TC_DEBUG=1 keeps the profiling ablations of the K1 kernels (TcParams::debug); the product compiles them out --
// each test was a constant-bank load and a branch per tile on the epilogue's critical path
// K3 is bound by the latency of its row gathers (each warp walks its candidate rows in ~400 dependent round trips), so
// what counts is how many warps are resident: the register budget asks for 10 CTAs of 4 warps per SM (48 registers, a few
// bytes of spills) and the candidate arrays come in three sizes.  Measured at C3 (interleaved A/B, ms per 10k-query call):
// 7 CTAs 1.12, 10 CTAs 1.02, 12 CTAs 1.03, 16 CTAs 1.12; the query row in shared memory (one copy per warp) 1.09.
#ifndef OSML_K3_MIN_CTAS
#define OSML_K3_MIN_CTAS 10
#endif
#ifndef OSML_TC_DEBUG
#define OSML_TC_DEBUG 0
#endif
#define OSML_TC_DBG(p) (OSML_TC_DEBUG ? (p).debug : 0)

namespace osml {

// ---------------------------------------------------------------------------------------------------------------------
// geometry
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kTcBM = 128;            // queries per tile (UMMA M)
constexpr int kTcBN = 256;            // train rows per tile (UMMA N)
constexpr int kTcKB = 64;             // bf16 per K block = one 128-byte swizzled row
constexpr int kTcABlockBytes = kTcBM * 128;   // 16 KB
constexpr int kTcBStageBytes = kTcBN * 128;   // 32 KB
constexpr int kTcThreads = 384;       // warp 0 TMA, 1 MMA (3 too in the pair kernel), 2 TMEM alloc, 4..11 epilogue
constexpr int kTcMaxKPrime = 64;
constexpr int kTcMaxCand = 512;       // candidates per query K3 handles (slots * K')
constexpr int kTcMaxCandRetry = 1024; // ... in the second filter pass (its own K3 instance, 2 queries per CTA)

struct TcParams {
    int nkb;             // K blocks of the padded feature dimension
    int ksteps;          // UMMA K steps (16 bf16 each) actually needed: ceil(d / 16)
    int nqt;             // query tiles
    int nseg;            // train segments
    int nq_single;       // pair kernel: query-tile pairs swept in one item (see TcItemIter); the others use nseg
    int tiles_per_seg;   // 256-row tiles per segment (segment s owns tiles s, s + nseg, ...: every segment sees the
                         // whole range of row norms, in ascending order)
    int tiles_total;
    int kprime;          // candidates kept per (query, slot)
    int f16;             // operands are fp16 (else bf16)
    int kactive;         // pair kernel: list entries in use, <= kprime (see TcRegList::reset)
    int epi_warps;       // 4 or 8
    int stages;          // B ring depth
    int a_resident;      // 1: the query tile stays in shared memory for a whole sweep; 0 (wide rows): its K blocks
                         //    are streamed through the ring next to the train blocks
    int nslot;           // nseg * halves
    float* cand_val;     // [qpad][nslot][kprime]
    int* cand_idx;
    float* cand_thr;     // [qpad][nslot]: every row the slot dropped has score <= this (-inf: nothing dropped)
    int nhot;            // pair kernel, split items: segments 0 .. nhot-1 interleave over tiles [0, thot) (the small-norm rows,
    int thot;            //    where the neighbours are); the others are contiguous runs of what follows (see TcItemIter)
    int live_all;        // whole-sweep items follow the gate too (their two column halves share it)
    unsigned int* gate;  // [qpad] pair kernel, split items: the best list threshold any slot of the query has proven so far,
                         //    order-preserving bits (tc_gate_enc); 0 = none yet
    const float* h;      // [npad] 0.5*|t^|^2 (+inf for padding rows)
    const float* hmin8;  // [npad / 8] minimum of h over each group of 8 rows
    const float* hblk;   // [tiles][288]: the same per 256-row tile, h then the 32 group minima (pair kernel, bulk copies)
    unsigned long long* dbg;   // counters != 0: cycle counters (see tc_neighbours_block)
    int counters;
    int debug;           // ablation switch for profiling: 1 = epilogue skips TMEM loads, 2 = loads but no selection
};

// ---------------------------------------------------------------------------------------------------------------------
// K0: preparation
// ---------------------------------------------------------------------------------------------------------------------
// X = element type of the train / query matrix: float, or double (the filter copy is 16-bit either way; exactness comes
// from K3, which rescoring in the matrix's own arithmetic)
template <typename X>
__global__ void tc_col_partial_sums_kernel(const X* __restrict__ x, int64_t n, int d, int rows_per_block,
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

// Element type of the filter copies (TensorState::f16).  bf16: the default.  fp16 of the centred values times a
// power-of-two scale: 11 significand bits, so rounding residuals 8x smaller and the completeness proof has 8x more
// room -- the choice when many queries need the second filter pass (it costs ~3 % more K1 time through the power
// cap, measured at the same cycle count, so it only pays when it saves retries).  Both are 16-bit: tc_elem is raw bits.
using tc_elem = unsigned short;
template <bool kF16>
__device__ __forceinline__ tc_elem tc_round(float x) {
    if (kF16) {     // subnormal results are flushed to zero so that the tensor core sees exactly what the residuals assume
        const __half h = __float2half_rn(x);
        return fabsf(__half2float(h)) < 6.103515625e-05f ? (tc_elem)0 : __half_as_ushort(h);
    }
    return __bfloat16_as_ushort(__float2bfloat16_rn(x));
}
template <bool kF16>
__device__ __forceinline__ float tc_value(tc_elem h) {
    return kF16 ? __half2float(__ushort_as_half(h)) : __bfloat162float(__ushort_as_bfloat16(h));
}

// centred, scaled value as the float that gets rounded to 16 bits, and the same in (all but) exact arithmetic
__device__ __forceinline__ float tc_centred(float v, float c, float scale) { return __fmul_rn(__fsub_rn(v, c), scale); }
__device__ __forceinline__ float tc_centred(double v, float c, float scale) {
    return __double2float_rn((v - (double)c) * (double)scale);
}
template <typename X>
__device__ __forceinline__ double tc_centred_exact(X v, float c, float scale) {
    return ((double)v - (double)c) * (double)scale;
}

// max |t - c| over the train set (finite values only): fixes the power-of-two scale of the fp16 filter copy
template <typename X>
__global__ void tc_max_abs_kernel(const X* __restrict__ x, int64_t n, int d, const float* __restrict__ center,
                                  unsigned int* __restrict__ max_bits) {
    const int64_t total = n * d;
    float m = 0.0f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const float a = fabsf(tc_centred(x[i], center[i % d], 1.0f));
        if (a < INFINITY) m = fmaxf(m, a);
    }
    for (int off = 16; off > 0; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
    if ((threadIdx.x & 31) == 0) atomicMax(max_bits, __float_as_uint(m));
}

// Pass A over the train rows (one warp per row): h = 0.5*|t^|^2 of the centred, rounded row, the rounding residual
// |t^ - (t - c)| and the maxima the completeness bound needs.  h is the sort key of the filter copy.
template <bool kF16, typename X>
__global__ void tc_train_norms_kernel(const X* __restrict__ x, int64_t n, int d, const float* __restrict__ center,
                                      float scale, float* __restrict__ half_norm, int* __restrict__ iota,
                                      TcPrepStats* stats) {
    const int lane = threadIdx.x & 31;
    const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= n) return;
    const X* xr = x + row * d;
    double n2 = 0.0, e2 = 0.0;
    for (int j = lane; j < d; j += 32) {
        const X v = xr[j];
        const float c = center[j];
        const double hd = (double)tc_value<kF16>(tc_round<kF16>(tc_centred(v, c, scale)));
        const double diff = hd - tc_centred_exact(v, c, scale);
        n2 += hd * hd;
        e2 += diff * diff;
    }
    for (int off = 16; off > 0; off >>= 1) {
        n2 += __shfl_xor_sync(0xffffffffu, n2, off);
        e2 += __shfl_xor_sync(0xffffffffu, e2, off);
    }
    if (lane == 0) {
        half_norm[row] = (float)(0.5 * n2);          // round to nearest: 2^-24 relative, part of TcBounds::c_mma
        iota[row] = (int)row;
        if (!(n2 < 1.0e37) || !(e2 < 1.0e37)) atomicExch(&stats->nonfinite, 1);
        else {
            atomicMax(&stats->et_max_bits, (unsigned long long)__double_as_longlong(sqrt(e2) * (1.0 + 1e-12)));
            atomicMax(&stats->tnorm_max_bits, (unsigned long long)__double_as_longlong(sqrt(n2) * (1.0 + 1e-12)));
        }
    }
}

// Pass B (one warp per row of the filter copy): position i holds train row perm[i] (rows sorted by ascending h, so the
// rows nearest to the centre -- the likeliest neighbours of any query -- are scanned first and the per-query
// thresholds settle early; groups of 8 consecutive rows also get nearly equal h, which keeps the group test tight).
// Queries (kTrain == false) keep their order and also record |q^|^2 and |q^ - (q - c)|.
template <bool kTrain, bool kF16, typename X>
__global__ void tc_prep_rows_kernel(const X* __restrict__ x, int64_t n, int64_t npad, int d, int dp,
                                    const float* __restrict__ center, float scale, const int* __restrict__ perm,
                                    tc_elem* __restrict__ out, double* __restrict__ norm2,
                                    double* __restrict__ err, float* __restrict__ half_norm_sorted,
                                    int* __restrict__ perm_pad) {
    const int lane = threadIdx.x & 31;
    const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= npad) return;
    tc_elem* o = out + row * dp;
    if (row >= n) {                       // padding row: can never be selected (train) / is ignored (queries)
        for (int j = lane; j < dp; j += 32) o[j] = 0;
        if (kTrain && lane == 0) { half_norm_sorted[row] = INFINITY; perm_pad[row] = -1; }
        if (!kTrain && lane == 0 && norm2) { norm2[row] = 0.0; err[row] = 0.0; }
        return;
    }
    const X* xr = x + (kTrain ? (int64_t)perm[row] : row) * d;
    double n2 = 0.0, e2 = 0.0;
    for (int j = lane; j < d; j += 32) {
        const X v = xr[j];
        const float c = center[j];
        const tc_elem h = tc_round<kF16>(tc_centred(v, c, scale));
        o[j] = h;
        if (!kTrain) {
            const double hd = (double)tc_value<kF16>(h);
            const double diff = hd - tc_centred_exact(v, c, scale);
            n2 += hd * hd;
            e2 += diff * diff;
        }
    }
    for (int j = d + lane; j < dp; j += 32) o[j] = 0;
    if (!kTrain) {
        for (int off = 16; off > 0; off >>= 1) {
            n2 += __shfl_xor_sync(0xffffffffu, n2, off);
            e2 += __shfl_xor_sync(0xffffffffu, e2, off);
        }
        if (lane == 0) {
            norm2[row] = n2;
            err[row] = sqrt(e2) * (1.0 + 1e-12);     // NaN / inf propagate and make the query fall back
        }
    }
}

// per 256-row tile: [256 x h][32 x min(h) over groups of 8 rows][min(h) of each half and of the tile] -- one 1168-byte
// bulk copy per tile in the pair kernel
constexpr int kTcHBlock = kTcBN + kTcBN / 8 + 4;     // floats per tile: h, group minima, min of columns 0-127 / 128-255 / all, pad
__global__ void tc_hblock_kernel(const float* __restrict__ h, int64_t npad, float* __restrict__ hmin8,
                                 float* __restrict__ hblk) {
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;       // group of 8 rows
    if (g >= npad / 8) return;
    const float4 a = reinterpret_cast<const float4*>(h)[2 * g], b = reinterpret_cast<const float4*>(h)[2 * g + 1];
    const float m = fminf(fminf(fminf(a.x, a.y), fminf(a.z, a.w)), fminf(fminf(b.x, b.y), fminf(b.z, b.w)));
    hmin8[g] = m;
    const int64_t tile = g / (kTcBN / 8), gi = g % (kTcBN / 8);
    float* blk = hblk + tile * kTcHBlock;
    reinterpret_cast<float4*>(blk)[2 * gi] = a;
    reinterpret_cast<float4*>(blk)[2 * gi + 1] = b;
    blk[kTcBN + gi] = m;
    if (gi == 0) {                                   // one thread per tile: minima over 128 / 128 / 256 rows
        float m0 = INFINITY, m1 = INFINITY;
        for (int j = 0; j < kTcBN / 2; ++j) m0 = fminf(m0, h[tile * kTcBN + j]);
        for (int j = kTcBN / 2; j < kTcBN; ++j) m1 = fminf(m1, h[tile * kTcBN + j]);
        blk[kTcBN + kTcBN / 8] = m0; blk[kTcBN + kTcBN / 8 + 1] = m1;
        blk[kTcBN + kTcBN / 8 + 2] = fminf(m0, m1); blk[kTcBN + kTcBN / 8 + 3] = 0.0f;
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

// Register-resident candidate list (KP = 16 or 32 entries per thread) fed through a small per-thread shared-memory
// buffer.
// List values are "packed": the low log2(KP) mantissa bits of an entry hold its slot number, so a min tree returns the
// threshold and where it lives in one go.  The perturbation (<= 2^-19 relative) is part of the error budget K3 proves
// completeness with (TcBounds::c_mma).
// The hot path only asks "does any lane of the warp see a value above its threshold in this group of 8 columns?" and
// collects the answers in a warp-uniform bit mask.  The rare groups that hit go through ONE out-of-line-sized code
// region (so the instruction cache holds a single copy): the group's 8 values are appended to the thread's buffer (a
// predicated 8-byte store each), and the warp then flushes the buffers into the lists in rounds -- round r inserts
// entry r of every lane that has one.  Everything dropped anywhere is <= the list minimum at that time <= the final
// threshold.
// Live gate of a query (pair kernel, queries whose sweep is split over several items): every slot publishes its list
// threshold -- K' rows of the query score above it -- and drops whatever scores at or below the best threshold ANY slot of
// the query has published.  K3's proof already asks lower(thr) > d_(k+1) of every slot, i.e. of the largest threshold, so a
// slot that drops up to that largest threshold loses nothing the proof could have used; it just stops inserting rows
// that cannot matter.  Any stale value is a valid gate (thresholds only grow), so relaxed accesses do.
__device__ __forceinline__ uint32_t tc_gate_enc(float f) {
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float tc_gate_dec(uint32_t u) {        // 0 (nothing published) decodes to a NaN, which fmaxf ignores
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u);
}
__device__ __forceinline__ uint32_t tc_gate_load(const unsigned int* addr) {
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void tc_gate_publish(unsigned int* addr, float thr) {
    asm volatile("red.relaxed.gpu.global.max.u32 [%0], %1;" ::"l"(addr), "r"(tc_gate_enc(thr)) : "memory");
}

constexpr int kTcHSlots = 8;          // ring of per-tile h blocks in the pair kernel
#ifndef OSML_TC_BUFCAP
#define OSML_TC_BUFCAP 16
#endif
constexpr int kTcBufCap = OSML_TC_BUFCAP;   // buffered candidates per thread; a flush is forced above kTcBufCap - 8

__device__ __forceinline__ void sts_v2(uint32_t addr, uint32_t a, uint32_t b) {
    asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ float4 lds_f4(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void lds_v2(uint32_t addr, uint32_t& a, uint32_t& b) {
    asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(a), "=r"(b) : "r"(addr) : "memory");
}

// kSmemRows: the row indices of the entries live in shared memory ([KP][threads] ints) instead of registers -- the
// pair kernel with K' = 32 needs the registers to stage a whole tile share (128 values) next to the 32 list values.
template <int KP, bool kSmemRows = false>
struct TcRegList {
    static constexpr uint32_t kSlotMask = KP - 1;
    float lv[KP];
    int lr[kSmemRows ? 1 : KP];
    uint32_t rows_addr, rows_stride;   // kSmemRows: shared address of this thread's entry 0, bytes between entries
    float thr;             // the list's own minimum
    float gate;            // best threshold another slot of the same query has proven (-inf: none)
    float eff;             // max(thr, gate): what a score must beat to matter
    uint32_t buf_addr;     // shared-memory address of this thread's buffer entry 0
    uint32_t buf_stride;   // bytes between consecutive entries of one thread (= epilogue threads * 8)
    int cnt;               // buffered candidates
    int rd;                // ... of which already fed to the list
    int buf_shift;         // log2(buf_stride): the epilogue has 128 or 256 threads
    __device__ __forceinline__ void attach(uint32_t addr, uint32_t stride) {
        buf_addr = addr; buf_stride = stride; buf_shift = 31 - __clz((int)stride); cnt = 0; rd = 0;
    }
    // empty slots hold -FLT_MAX (or a seed threshold proven elsewhere: values at or below it are never inserted; the pair
    // kernel uses the live gate instead and always resets to -FLT_MAX)
    // `active` <= KP entries take part; the others are parked at +FLT_MAX so that they are never the minimum (a list
    // length between the powers of two the slot packing allows: fewer entries -> higher threshold -> fewer hits)
    __device__ __forceinline__ void reset(float seed = -INFINITY, int active = KP) {
        const uint32_t base = seed > -3.0e38f ? __float_as_uint(seed) : 0xFF7FFFFFu;
#pragma unroll
        for (int e = 0; e < KP; ++e) {
            lv[e] = __uint_as_float((((e < active) ? base : 0x7F7FFFFFu) & ~kSlotMask) | (uint32_t)e);
            if (kSmemRows) sts_b32(rows_addr + (uint32_t)e * rows_stride, 0xffffffffu);
            else lr[e] = -1;
        }
        thr = tc_list_min<KP>(lv);
        gate = -INFINITY;
        eff = thr;
        cnt = 0;
        rd = 0;
    }
    __device__ __forceinline__ void raise_gate(float g) {     // NaN (nothing published yet) leaves it as it is
        gate = fmaxf(gate, g);
        eff = fmaxf(thr, gate);
    }
    // branch-free insert (a select per entry): lanes with take == false match no slot
    __device__ __forceinline__ void insert_sel(bool take, float cur, int row) {
        const uint32_t slot = take ? (__float_as_uint(thr) & kSlotMask) : 0xFFFFFFFFu;
        const uint32_t nvb = (__float_as_uint(cur) & ~kSlotMask) | (slot & kSlotMask);
        if (kSmemRows) {
            if (take) sts_b32(rows_addr + (slot & kSlotMask) * rows_stride, (uint32_t)row);
#pragma unroll
            for (int e = 0; e < KP; ++e) {
                uint32_t vb = __float_as_uint(lv[e]);
                asm("{\n\t.reg .pred p;\n\tsetp.eq.u32 p, %1, %2;\n\tselp.b32 %0, %3, %0, p;\n\t}"
                    : "+r"(vb) : "r"(slot), "r"((uint32_t)e), "r"(nvb));
                lv[e] = __uint_as_float(vb);
            }
        } else {
#pragma unroll
            for (int e = 0; e < KP; ++e) {
                uint32_t vb = __float_as_uint(lv[e]);
                asm("{\n\t.reg .pred p;\n\tsetp.eq.u32 p, %2, %3;\n\tselp.b32 %0, %4, %0, p;\n\tselp.b32 %1, %5, %1, p;\n\t}"
                    : "+r"(vb), "+r"(lr[e]) : "r"(slot), "r"((uint32_t)e), "r"(nvb), "r"(row));
                lv[e] = __uint_as_float(vb);
            }
        }
        thr = tc_list_min<KP>(lv);
        eff = fmaxf(thr, gate);
    }
    __device__ __forceinline__ int row_of(int e) const {
        return kSmemRows ? (int)__float_as_uint(lds_f32(rows_addr + (uint32_t)e * rows_stride)) : lr[kSmemRows ? 0 : e];
    }
    // warp-collective: every lane with unread buffer entries feeds its oldest one to its list.  Callers run rounds
    // while they wait for the next accumulator tile, so list maintenance rides on otherwise idle issue slots.
    __device__ __forceinline__ void flush_round() {
        uint32_t vb = 0xFF800000u, rb = 0;          // -inf: lanes without an unread entry never insert
        const bool have = rd < cnt;
        if (have) lds_v2(buf_addr + (uint32_t)rd * buf_stride, vb, rb);
        rd += have ? 1 : 0;
        const float val = __uint_as_float(vb);
        const bool take = val > eff;
        if (__any_sync(0xffffffffu, take)) insert_sel(take, val, (int)rb);
        if (!__any_sync(0xffffffffu, rd < cnt)) { cnt = 0; rd = 0; }    // linear buffers: rewind once all are drained
    }
    __device__ __forceinline__ bool pending() const { return __any_sync(0xffffffffu, rd < cnt); }
    __device__ __forceinline__ void flush() {
#pragma unroll 1
        while (pending()) flush_round();
    }
    // appends the values above the threshold (branch-free: a predicated 8-byte store and a predicated bump each)
    // columns OFF .. OFF+7 of chunk v (train rows row0 .. row0+7, their h in ha | hb)
    template <int OFF>
    __device__ __forceinline__ void append_group(const float (&v)[32], const float4 ha, const float4 hb, int row0) {
        const float hh[8] = {ha.x, ha.y, ha.z, ha.w, hb.x, hb.y, hb.z, hb.w};
        uint32_t waddr = buf_addr + (uint32_t)cnt * buf_stride;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float w = v[OFF + j] - hh[j];
            asm volatile(
                "{\n\t.reg .pred p;\n\tsetp.gt.f32 p, %1, %2;\n\t@p st.shared.v2.b32 [%0], {%1, %3};\n\t"
                "@p add.u32 %0, %0, %4;\n\t}"
                : "+r"(waddr)
                : "f"(w), "f"(eff), "r"(row0 + j), "r"(buf_stride)
                : "memory");
        }
        cnt = (int)((waddr - buf_addr) >> buf_shift);
    }
    // warp-uniform 4-bit mask: bit g set when some lane may have a score above its threshold in columns 8g .. 8g+7 of
    // the chunk.  v holds the raw dot products; score = v - h[column], and hm holds min(h) of the chunk's 4 groups, so
    // max(v over the group) <= thr + min(h) rules the whole group out without touching h.
    __device__ __forceinline__ uint32_t hits(const float (&v)[32], const float4 hm) const {
        const float g0 = tc_max8(v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7]);
        const float g1 = tc_max8(v[8], v[9], v[10], v[11], v[12], v[13], v[14], v[15]);
        const float g2 = tc_max8(v[16], v[17], v[18], v[19], v[20], v[21], v[22], v[23]);
        const float g3 = tc_max8(v[24], v[25], v[26], v[27], v[28], v[29], v[30], v[31]);
        const bool p0 = g0 > eff + hm.x, p1 = g1 > eff + hm.y, p2 = g2 > eff + hm.z, p3 = g3 > eff + hm.w;
        uint32_t m = 0;
        if (__any_sync(0xffffffffu, p0 | p1 | p2 | p3)) {
            m = (__any_sync(0xffffffffu, p0) ? 1u : 0u) | (__any_sync(0xffffffffu, p1) ? 2u : 0u) |
                (__any_sync(0xffffffffu, p2) ? 4u : 0u) | (__any_sync(0xffffffffu, p3) ? 8u : 0u);
        }
        return m;
    }
    // per-lane version of the same test (no votes): bit g set when THIS lane may have a hit in group g of the chunk.
    // OR-ing the four chunks' bits (shifted) and one __reduce_or_sync gives the warp-uniform mask of a whole tile.
    __device__ __forceinline__ uint32_t lane_bits(const float (&v)[32], const float4 hm) const {
        const float g0 = tc_max8(v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7]);
        const float g1 = tc_max8(v[8], v[9], v[10], v[11], v[12], v[13], v[14], v[15]);
        const float g2 = tc_max8(v[16], v[17], v[18], v[19], v[20], v[21], v[22], v[23]);
        const float g3 = tc_max8(v[24], v[25], v[26], v[27], v[28], v[29], v[30], v[31]);
        return (g0 > eff + hm.x ? 1u : 0u) | (g1 > eff + hm.y ? 2u : 0u) | (g2 > eff + hm.z ? 4u : 0u) |
               (g3 > eff + hm.w ? 8u : 0u);
    }
    // slow path over the chunks held in registers; `mask` has one bit per group of 8 columns of v0 | v1 | v2 | v3;
    // h of column 0 of v0 is at hrow (global memory) or, with kSmemH, at shared-memory address h_saddr
    template <bool kSmemH>
    __device__ __forceinline__ void drain(uint32_t mask, const float (&v0)[32], const float (&v1)[32],
                                          const float (&v2)[32], const float (&v3)[32], int row0,
                                          const float* __restrict__ hrow, uint32_t h_saddr) {
        if (__any_sync(0xffffffffu, cnt > kTcBufCap - 8)) flush();
#pragma unroll 1
        while (mask != 0) {
            const int g = __ffs((int)mask) - 1;
            mask &= mask - 1;
            float4 ha, hb;
            if (kSmemH) {
                ha = lds_f4(h_saddr + 32u * (uint32_t)g);
                hb = lds_f4(h_saddr + 32u * (uint32_t)g + 16u);
            } else {
                ha = __ldg(reinterpret_cast<const float4*>(hrow + 8 * g));
                hb = __ldg(reinterpret_cast<const float4*>(hrow + 8 * g + 4));
            }
            // one statically indexed copy of the 8-element append per group (registers cannot be indexed dynamically,
            // and gathering the group into a common array first costs more than the append itself)
            // (an explicit two-level branch tree on the warp-uniform g: a 16-way jump table costs a constant-bank load)
#define OSML_TC_QUAD(ARR, BASE)                                                                     \
            {                                                                                        \
                const int gq = g & 3;                                                                \
                if (gq < 2) {                                                                        \
                    if (gq == 0) append_group<0>(ARR, ha, hb, row0 + 8 * (BASE));                    \
                    else append_group<8>(ARR, ha, hb, row0 + 8 * (BASE) + 8);                        \
                } else {                                                                             \
                    if (gq == 2) append_group<16>(ARR, ha, hb, row0 + 8 * (BASE) + 16);              \
                    else append_group<24>(ARR, ha, hb, row0 + 8 * (BASE) + 24);                      \
                }                                                                                    \
            }
            if (g < 8) {
                if (g < 4) OSML_TC_QUAD(v0, 0) else OSML_TC_QUAD(v1, 4)
            } else {
                if (g < 12) OSML_TC_QUAD(v2, 8) else OSML_TC_QUAD(v3, 12)
            }
#undef OSML_TC_QUAD
            // forced catch-up when a buffer nears its capacity (mostly in the first tiles of a sweep, when every
            // value beats the threshold); otherwise the lists catch up while the warp waits for its next tile
            // (a group adds at most 8 entries, and the loop is entered with at most kTcBufCap - 8: no check needed
            // after the last group)
            if (mask != 0 && __any_sync(0xffffffffu, cnt > kTcBufCap - 8)) flush();
        }
    }
    // one 32-column chunk starting at train row row0 (hm: the 4 group minima of h)
    template <bool kSmemH>
    __device__ __forceinline__ void scan(const float (&v)[32], int row0, const float4 hm, const float* __restrict__ hrow,
                                         uint32_t h_saddr) {
        const uint32_t m = hits(v, hm);
        if (m) drain<kSmemH>(m, v, v, v, v, row0, hrow, h_saddr);
    }
    __device__ __forceinline__ void scan(const float (&v)[32], int row0, const float* __restrict__ h,
                                         const float* __restrict__ hmin8) {
        scan<false>(v, row0, __ldg(reinterpret_cast<const float4*>(hmin8 + (row0 >> 3))), h + row0, 0u);
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
    const size_t list_elems = KP > 0 ? 0 : (size_t)p.kprime * epi_threads;   // register lists only need the buffer
    unsigned char* cand_buf = b_smem + (size_t)p.stages * ring_stage_bytes;  // [kTcBufCap][epi_threads] x 8 B
    float* list_val = reinterpret_cast<float*>(cand_buf + (KP > 0 ? (size_t)kTcBufCap * epi_threads * 8 : 0));
    int* list_idx = reinterpret_cast<int*>(list_val + list_elems);
    uint64_t* bars = reinterpret_cast<uint64_t*>(list_idx + list_elems);
    uint64_t* full_bar = bars;                    // [stages]
    uint64_t* empty_bar = bars + 8;               // [stages]
    uint64_t* tmem_full = bars + 16;              // [2]
    uint64_t* tmem_empty = bars + 18;             // [2]
    uint64_t* a_full = bars + 20;
    uint64_t* a_empty = bars + 21;
    uint64_t* first_done = bars + 22;             // [2]       (wide rows, leader: K block 0 of the tile in this stage is issued)
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
                const int tile0 = seg;                                   // segments interleave: tiles seg, seg + nseg, ...
                const int ntiles = (p.tiles_total - seg + p.nseg - 1) / p.nseg;
                if (p.a_resident) {
                    ptx::mbar_wait(a_empty, (it & 1) ^ 1);
                    ptx::mbar_expect_tx(a_full, (uint32_t)p.nkb * kTcABlockBytes);
                    for (int kb = 0; kb < p.nkb; ++kb)
                        ptx::tma_load_2d(a_smem + (size_t)kb * kTcABlockBytes, &map_a, a_full, kb * kTcKB, qt * kTcBM);
                }
                for (int t = 0; t < ntiles; ++t) {
                    const int row0 = (tile0 + t * p.nseg) * kTcBN;
                    for (int kb = 0; kb < p.nkb; ++kb) {
                        ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
                        unsigned char* ring = b_smem + (size_t)stage * ring_stage_bytes;
                        if (OSML_TC_DBG(p) == 3) {                      // ablation: no operand traffic, MMA on stale smem
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
            const uint32_t idesc = (p.f16 ? ptx::umma_idesc_f16(kTcBM, kTcBN) : ptx::umma_idesc_bf16(kTcBM, kTcBN));
            const uint32_t hi = ptx::umma_desc_hi_k_sw128();
            const uint32_t a_lo0 = ptx::umma_desc_lo(ptx::smem_u32(a_smem)), b_lo0 = ptx::umma_desc_lo(ptx::smem_u32(b_smem));
            uint32_t stage = 0, phase = 0, it = 0, acc_it = 0;
            for (int item = blockIdx.x; item < nitems; item += gridDim.x, ++it) {
                const int seg = item / p.nqt;
                const int tile0 = seg;                                   // segments interleave: tiles seg, seg + nseg, ...
                const int ntiles = (p.tiles_total - seg + p.nseg - 1) / p.nseg;
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
            list.attach(ptx::smem_u32(cand_buf) + (uint32_t)et * 8u, (uint32_t)epi_threads * 8u);
            for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
                const int seg = item / p.nqt, qt = item % p.nqt;
                const int tile0 = seg;                                   // segments interleave: tiles seg, seg + nseg, ...
                const int ntiles = (p.tiles_total - seg + p.nseg - 1) / p.nseg;
                list.reset();
                if (OSML_TC_DBG(p) == 4) list.thr = list.eff = INFINITY;   // ablation: fast path only
                for (int t = 0; t < ntiles; ++t, ++acc_it) {
                    const uint32_t as = acc_it & 1, aphase = (acc_it >> 1) & 1;
                    ptx::mbar_wait(&tmem_full[as], aphase);
                    ptx::tc_fence_after();
                    const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + as * kTcBN + half * cols;
                    const int row_base = (tile0 + t * p.nseg) * kTcBN + half * cols;
                    if (OSML_TC_DBG(p) != 1 && OSML_TC_DBG(p) != 3) {
                        if constexpr (KP <= 16) {
                            float va[32], vb[32];
                            ptx::tmem_ld_32x32(taddr, va);
                            for (int c = 0; c < cols; c += 64) {   // two chunks per trip, loads one chunk ahead
                                ptx::tmem_ld_wait();
                                ptx::tmem_ld_32x32(taddr + c + 32, vb);
                                if (OSML_TC_DBG(p) != 2) list.scan(va, row_base + c, p.h, p.hmin8);
                                else if (va[0] + va[31] == 12345.678f) list.thr = va[0];
                                ptx::tmem_ld_wait();
                                if (c + 64 < cols) ptx::tmem_ld_32x32(taddr + c + 64, va);
                                if (OSML_TC_DBG(p) != 2) list.scan(vb, row_base + c + 32, p.h, p.hmin8);
                                else if (vb[0] + vb[31] == 12345.678f) list.thr = vb[0];
                            }
                        } else {
                            for (int c = 0; c < cols; c += 32) {
                                float va[32];
                                ptx::tmem_ld_32x32(taddr + c, va);
                                ptx::tmem_ld_wait();
                                list.scan(va, row_base + c, p.h, p.hmin8);
                            }
                        }
                    }
                    ptx::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) ptx::mbar_arrive(&tmem_empty[as]);
                }
                list.flush();
                const size_t q = (size_t)qt * kTcBM + m;
                const size_t slot = (size_t)seg * halves + half;
                float* ov = p.cand_val + (q * p.nslot + slot) * KP;
                int* oi = p.cand_idx + (q * p.nslot + slot) * KP;
                bool full = true;
#pragma unroll
                for (int e = 0; e < KP; ++e) { const int r = list.row_of(e); ov[e] = list.lv[e]; oi[e] = r; full = full && r >= 0; }
                p.cand_thr[q * p.nslot + slot] = full ? list.thr : -INFINITY;
            }
        } else {
            const uint32_t stride_bytes = (uint32_t)epi_threads * 4;
            const uint32_t val_addr = ptx::smem_u32(list_val + et), idx_addr = ptx::smem_u32(list_idx + et);
            for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
                const int seg = item / p.nqt, qt = item % p.nqt;
                const int tile0 = seg;                                   // segments interleave: tiles seg, seg + nseg, ...
                const int ntiles = (p.tiles_total - seg + p.nseg - 1) / p.nseg;
                for (int e = 0; e < p.kprime; ++e) {
                    sts_b32(val_addr + e * stride_bytes, __float_as_uint(-INFINITY));
                    sts_b32(idx_addr + e * stride_bytes, 0xffffffffu);
                }
                float thr = -INFINITY;
                uint32_t pos_min = 0;
                auto scan = [&](const float (&v)[32], int row0) {
                    const float4 hm = __ldg(reinterpret_cast<const float4*>(p.hmin8 + (row0 >> 3)));
                    const float mx = tc_max32(v);
                    if (mx > thr + fminf(fminf(hm.x, hm.y), fminf(hm.z, hm.w))) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            const float sc = v[i] - __ldg(p.h + row0 + i);
                            if (sc > thr) {
                                const unsigned long long r =
                                    tc_list_insert(val_addr, idx_addr, p.kprime, stride_bytes, pos_min, sc, row0 + i);
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
                    const int row_base = (tile0 + t * p.nseg) * kTcBN + half * cols;
                    float va[32], vb[32];
                    if (OSML_TC_DBG(p) != 1 && OSML_TC_DBG(p) != 3) {
                        ptx::tmem_ld_32x32(taddr, va);
                        for (int c = 0; c < cols; c += 64) {      // two 32-column chunks per trip, loads one chunk ahead
                            ptx::tmem_ld_wait();
                            ptx::tmem_ld_32x32(taddr + c + 32, vb);
                            if (OSML_TC_DBG(p) != 2) scan(va, row_base + c);
                            else if (va[0] + va[31] == 12345.678f) thr = va[0];
                            ptx::tmem_ld_wait();
                            if (c + 64 < cols) ptx::tmem_ld_32x32(taddr + c + 64, va);
                            if (OSML_TC_DBG(p) != 2) scan(vb, row_base + c + 32);
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

// Work items of the pair kernel.  Query-tile pairs 0 .. nq_single-1 are swept over the whole train set in ONE item
// (one warm-up of the candidate lists, the tightest thresholds); the remaining pairs -- those that would leave most of
// the machine idle in a last partial wave -- are split into nseg interleaved segments (tiles seg, seg + nseg, ...).
// Work items of a CTA pair, in the order it processes them (round robin over the item list): the whole-sweep items of
// pairs 0 .. nq_single-1, then the split pairs' interleaved segments (tiles seg, seg + nseg, ...: every segment sees
// the whole range of row norms in ascending order), numbered segment-major -- all pairs' segment 0, then all pairs'
// segment 1, ... -- so that by the time a later segment of a query starts its earlier segments have usually finished
// and the query's gate (see tc_gate_enc) already holds what they proved.  Every item writes candidate slot `slot` of
// its queries (x halves).
// Hot-first split (option `hot`, nhot < nseg; not the default -- measured, see tc_filter_block): the filter copy is sorted
// by row norm and score = q.t - h, so the rows that end up in a query's lists sit overwhelmingly in the first tiles, and
// interleaving ALL segments hands every item its own share of those rows.  With nhot < nseg the first nhot segments
// interleave over the first thot tiles only and the later ones take contiguous runs of the rest behind a settled gate.
// (Measured and dropped: cutting the pair-major tile sequence into one contiguous run per CTA pair.  It balances
// perfectly and makes sweeps as long as possible, but a run in the middle of the norm-sorted order never sees the
// small-norm rows that settle the thresholds: 10-60 % slower on every shape tried.)
struct TcItem { int qp, tile0, stride, ntiles, slot; };
struct TcItemIter {
    int item, nitems, step;
    __device__ __forceinline__ TcItemIter(const TcParams& p, int cluster_id, int nclusters) {
        item = cluster_id; step = nclusters;
        nitems = p.nq_single + (p.nqt / 2 - p.nq_single) * p.nseg;
    }
    __device__ __forceinline__ bool next(const TcParams& p, TcItem& r) {
        if (item >= nitems) return false;
        if (item < p.nq_single) { r.qp = item; r.tile0 = 0; r.stride = 1; r.slot = 0; r.ntiles = p.tiles_total; }
        else {
            const int j = item - p.nq_single, nrest = p.nqt / 2 - p.nq_single;
            r.qp = p.nq_single + j % nrest; r.slot = j / nrest;
            if (r.slot < p.nhot) {                     // hot segments interleave over the first thot tiles
                r.tile0 = r.slot; r.stride = p.nhot;
                r.ntiles = (p.thot - r.tile0 + p.nhot - 1) / p.nhot;
            } else {                                   // the others: contiguous, equal runs of the remaining tiles
                const long long c = r.slot - p.nhot, ncold = p.nseg - p.nhot, rest = p.tiles_total - p.thot;
                const int lo = p.thot + (int)(rest * c / ncold), hi = p.thot + (int)(rest * (c + 1) / ncold);
                r.tile0 = lo; r.stride = 1; r.ntiles = hi - lo;
            }
        }
        item += step;
        return true;
    }
};

// kWide (rows too long for a resident query tile next to two whole half-tile stages: D > 192): the ring works at K-block
// granularity instead -- a stage holds [this CTA's 128 x 64 query block][its 128 x 64 half of the train block], one
// barrier round trip per K block, K blocks alternating between two issuer threads (the round trips of a lone issuer were
// tensor-pipe idle time: see the issuers below).
// Everything else -- the CTA pair's M = 256 MMA that halves the L2 -> shared-memory traffic per FLOP of the single-CTA
// kernel, the staged epilogue with early TMEM release, the h ring, the live gates -- is the same code.
template <int KP, bool kWide>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kTcThreads, 1)
knn_tc_filter_pair_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                          const TcParams p) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* a_smem = smem;
    unsigned char* b_smem = a_smem + (kWide ? (size_t)0 : (size_t)p.nkb * kTcABlockBytes);
    const uint32_t stage_bytes = kWide ? (uint32_t)(kTcABlockBytes + kTcBHalfBytes) : (uint32_t)p.nkb * kTcBHalfBytes;
    const int epi_threads = p.epi_warps * 32;
    unsigned char* cand_buf = b_smem + (size_t)p.stages * stage_bytes;         // [kTcBufCap][epi_threads] x 8 B
    unsigned char* h_ring = cand_buf + (size_t)kTcBufCap * epi_threads * 8;     // [kTcHSlots][kTcHBlock] floats
    unsigned char* row_lists = h_ring + (size_t)kTcHSlots * kTcHBlock * 4;     // [K'][epi_threads] ints
    uint64_t* bars = reinterpret_cast<uint64_t*>(row_lists + (size_t)KP * epi_threads * 4);
    uint64_t* full_bar = bars;                    // [stages]  (used in the leader)
    uint64_t* empty_bar = bars + 8;               // [stages]
    uint64_t* tmem_full = bars + 16;              // [2]
    uint64_t* tmem_empty = bars + 18;             // [2]       (used in the leader)
    uint64_t* a_full = bars + 20;                 //           (used in the leader)
    uint64_t* a_empty = bars + 21;
    uint64_t* first_done = bars + 22;             // [2]       (wide rows, leader: K block 0 of the tile in this stage is issued)
    uint64_t* h_full = bars + 24;                 // [kTcHSlots]  this CTA's copy of the tile's h block has landed
    uint64_t* h_empty = bars + 24 + kTcHSlots;    // [kTcHSlots]  ... and every epilogue warp is done with it
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 24 + 2 * kTcHSlots);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = ptx::cluster_ctarank();
    const bool leader = rank == 0;
    const int cluster_id = blockIdx.x >> 1, nclusters = gridDim.x >> 1;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tensormap(&map_a);
        ptx::prefetch_tensormap(&map_b);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < p.stages; ++s) { ptx::mbar_init(&full_bar[s], 2); ptx::mbar_init(&empty_bar[s], 1); }
        for (int s = 0; s < 2; ++s) {
            ptx::mbar_init(&tmem_full[s], kWide ? 2 : 1);     // wide rows: one arrival per issuer thread
            ptx::mbar_init(&tmem_empty[s], 2 * p.epi_warps);
            ptx::mbar_init(&first_done[s], 1);
        }
        ptx::mbar_init(a_full, 2);
        ptx::mbar_init(a_empty, 2);
        for (int s = 0; s < kTcHSlots; ++s) { ptx::mbar_init(&h_full[s], 1); ptx::mbar_init(&h_empty[s], p.epi_warps); }
        ptx::fence_barrier_init();
    }
    if (warp == 2) ptx::tmem_alloc_pair(tmem_slot, 512);
    ptx::tc_fence_before();
    ptx::cluster_sync_all();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // 384 threads x 168 registers at launch: warps 0-3 keep 56 each, the two epilogue warpgroups get 224 each
    if (warp < 4) {
        ptx::setmaxnreg_dec<56>();

        if (warp == 0) {
            if (lane == 0) {
                // ===== TMA producer (both CTAs): own query tile, own half of each train tile =====
                uint32_t stage = 0, phase = 0, it = 0, hslot = 0, hphase = 0;
                TcItemIter items(p, cluster_id, nclusters);
                for (TcItem wi; items.next(p, wi); ++it) {
                    const int nsg = wi.stride, qt = 2 * wi.qp + (int)rank;
                    const int tile0 = wi.tile0, ntiles = wi.ntiles;
                    if (!kWide) {
                        ptx::mbar_wait(a_empty, (it & 1) ^ 1);
                        if (leader) ptx::mbar_expect_tx(a_full, 2u * (uint32_t)p.nkb * kTcABlockBytes);
                        else ptx::mbar_arrive_remote(a_full, 0);
                        for (int kb = 0; kb < p.nkb; ++kb)
                            ptx::tma_load_2d_pair(a_smem + (size_t)kb * kTcABlockBytes, &map_a, a_full, kb * kTcKB, qt * kTcBM);
                    }
                    for (int t = 0; t < ntiles; ++t) {
                        const int row0 = (tile0 + t * nsg) * kTcBN + (int)rank * (kTcBN / 2);
                        // this CTA's copy of the tile's h block (256 x h, 32 group minima) for its epilogue warps
                        ptx::mbar_wait(&h_empty[hslot], hphase ^ 1);
                        ptx::mbar_expect_tx(&h_full[hslot], kTcHBlock * 4);
                        ptx::bulk_load_1d(h_ring + (size_t)hslot * kTcHBlock * 4, p.hblk + (size_t)(tile0 + t * nsg) * kTcHBlock,
                                          kTcHBlock * 4, &h_full[hslot]);
                        if (++hslot == (uint32_t)kTcHSlots) { hslot = 0; hphase ^= 1; }
                        if (kWide) {
                            for (int kb = 0; kb < p.nkb; ++kb) {     // one stage per K block: query block + train half block
                                ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
                                if (leader) ptx::mbar_expect_tx(&full_bar[stage], 2u * stage_bytes);
                                else ptx::mbar_arrive_remote(&full_bar[stage], 0);
                                unsigned char* ring = b_smem + (size_t)stage * stage_bytes;
                                ptx::tma_load_2d_pair(ring, &map_a, &full_bar[stage], kb * kTcKB, qt * kTcBM);
                                ptx::tma_load_2d_pair(ring + kTcABlockBytes, &map_b, &full_bar[stage], kb * kTcKB, row0);
                                if (++stage == (uint32_t)p.stages) { stage = 0; phase ^= 1; }
                            }
                            continue;
                        }
                        ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
                        if (OSML_TC_DBG(p) == 3) {                          // ablation: no operand traffic
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
        } else if (kWide && (warp == 1 || warp == 3)) {
            if (lane == 0 && leader) {
                // ===== MMA issuers (leader CTA only), K-block granular ring =====
                // Two threads share every tile: K blocks are numbered through the whole run of this CTA pair (they sit in
                // the ring in that order) and issuer 0 (warp 1) takes the even ones, issuer 1 (warp 3) the odd ones.  The
                // tensor core takes an MMA when its thread gets to it, so a lone issuer's barrier round trips (a full-wait
                // and a commit per K block, ~150 cycles against 512 cycles of MMA) were pipe idle time; now one thread waits
                // and commits while the other's four MMAs run.  Within a tile the order of the K blocks is free (the sum
                // is the same up to the rounding the error bound already covers) EXCEPT that the block with K block 0
                // overwrites the accumulator: its issuer signals `first_done` once those MMAs are issued, the other
                // waits for it before its first block of the tile.  Both commit to tmem_full (two arrivals).
                const uint32_t issuer = (uint32_t)warp >> 1;
                const uint32_t idesc = (p.f16 ? ptx::umma_idesc_f16(2 * kTcBM, kTcBN) : ptx::umma_idesc_bf16(2 * kTcBM, kTcBN));
                const uint32_t hi = ptx::umma_desc_hi_k_sw128();
                const uint32_t ring_lo0 = ptx::umma_desc_lo(ptx::smem_u32(b_smem));
                const uint32_t stage_step = stage_bytes >> 4;
                int ntot = 0;                                    // tiles this CTA pair computes, all items together
                {
                    TcItemIter items(p, cluster_id, nclusters);
                    for (TcItem wi; items.next(p, wi);) ntot += wi.ntiles;
                }
                const int nkb = p.nkb, nstages = p.stages;
                // this thread's next K block: number kb of tile `tile`, sitting in ring stage `stage`
                int tile = 0, kb = (int)issuer;
                uint32_t stage = issuer, phase = 0;
                if (stage >= (uint32_t)nstages) { stage -= nstages; phase ^= 1; }
                bool enter = true;                               // the next block is this thread's first of its tile
                long long dbg_wait_acc = 0, dbg_wait_ops = 0;
                const long long dbg_t0 = clock64();
                while (tile < ntot) {
                    const uint32_t as = (uint32_t)tile & 1, aphase = ((uint32_t)tile >> 1) & 1;
                    const uint32_t tmem_d = tmem_base + as * kTcBN;
                    if (enter) {
                        const long long dbg_a0 = (OSML_TC_COUNTERS && p.counters) ? clock64() : 0;
                        ptx::mbar_wait(&tmem_empty[as], aphase ^ 1);
                        if (kb != 0) ptx::mbar_wait(&first_done[as], aphase);      // K block 0 has been issued
                        if ((OSML_TC_COUNTERS && p.counters)) dbg_wait_acc += clock64() - dbg_a0;
                        enter = false;
                    }
                    const long long dbg_o0 = (OSML_TC_COUNTERS && p.counters) ? clock64() : 0;
                    ptx::mbar_wait(&full_bar[stage], phase);
                    if ((OSML_TC_COUNTERS && p.counters)) dbg_wait_ops += clock64() - dbg_o0;
                    ptx::tc_fence_after();
                    const int nks = p.ksteps - kb * 4;
                    const uint32_t a_lo = ring_lo0 + stage * stage_step;
                    const uint32_t b_lo = a_lo + (kTcABlockBytes >> 4);
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks)
                        if (ks < nks)
                            ptx::umma_bf16_ss_pair_lohi(tmem_d, a_lo + ks * 2, b_lo + ks * 2, hi, idesc,
                                                        (uint32_t)((kb | ks) != 0));
                    ptx::umma_commit_pair(&empty_bar[stage]);      // both CTAs may refill this stage
                    if (kb == 0) {
                        ptx::tc_fence_before();
                        ptx::mbar_arrive(&first_done[as]);
                    }
                    kb += 2; stage += 2;
                    if (stage >= (uint32_t)nstages) { stage -= nstages; phase ^= 1; }
                    if (kb >= nkb) {                               // that was this thread's last block of the tile
                        ptx::umma_commit_pair(&tmem_full[as]);     // one of the two arrivals the epilogues wait for
                        kb -= nkb; ++tile; enter = true;
                    }
                }
                if ((OSML_TC_COUNTERS && p.counters) && issuer == 0) {
                    atomicAdd(p.dbg + 0, (unsigned long long)dbg_wait_acc);
                    atomicAdd(p.dbg + 1, (unsigned long long)dbg_wait_ops);
                    atomicAdd(p.dbg + 2, (unsigned long long)(clock64() - dbg_t0));
                }
            }
        } else if (!kWide && (warp == 1 || warp == 3)) {
            if (lane == 0 && leader) {
                // ===== MMA issuers (leader CTA only) =====
                // Two threads, one per accumulator stage: issuer 0 (warp 1) takes the even tiles, issuer 1 (warp 3) the
                // odd ones.  The tensor pipe holds only a couple of MMAs in flight, so a single issuer's per-tile
                // bookkeeping (two barrier waits, two commits: ~400 cycles) showed up as pipe bubbles; with two, one
                // thread waits and commits while the other keeps the pipe fed.  Tiles of different parity write
                // different TMEM stages and read different shared-memory stages, so their relative order is free.
                const uint32_t issuer = (uint32_t)warp >> 1;
                const uint32_t idesc = (p.f16 ? ptx::umma_idesc_f16(2 * kTcBM, kTcBN) : ptx::umma_idesc_bf16(2 * kTcBM, kTcBN));
                const uint32_t hi = ptx::umma_desc_hi_k_sw128();
                const uint32_t a_lo0 = ptx::umma_desc_lo(ptx::smem_u32(a_smem)), b_lo0 = ptx::umma_desc_lo(ptx::smem_u32(b_smem));
                uint32_t stage = 0, phase = 0, it = 0, acc_it = 0;
                long long dbg_wait_acc = 0, dbg_wait_ops = 0, dbg_issue = 0, dbg_commit = 0, dbg_t2 = 0;
                const long long dbg_t0 = clock64();
                TcItemIter items(p, cluster_id, nclusters);
                for (TcItem wi; items.next(p, wi); ++it) {
                    const int ntiles = wi.ntiles;
                    ptx::mbar_wait(a_full, it & 1);
                    ptx::tc_fence_after();
                    for (int t = 0; t < ntiles; ++t, ++acc_it) {
                        const uint32_t as = acc_it & 1, aphase = (acc_it >> 1) & 1;
                        if (as != issuer) {                          // the other issuer's tile
                            if (++stage == (uint32_t)p.stages) { stage = 0; phase ^= 1; }
                            continue;
                        }
                        if ((OSML_TC_COUNTERS && p.counters)) {      // where does the issuer wait: accumulator stage (epilogue) or operands (TMA)?
                            const long long t0 = clock64();
                            ptx::mbar_wait(&tmem_empty[as], aphase ^ 1);
                            const long long t1 = clock64();
                            ptx::mbar_wait(&full_bar[stage], phase);
                            const long long t2 = clock64();
                            dbg_wait_acc += t1 - t0; dbg_wait_ops += t2 - t1; dbg_t2 = t2;
                        } else {
                            ptx::mbar_wait(&tmem_empty[as], aphase ^ 1);
                            ptx::mbar_wait(&full_bar[stage], phase);
                        }
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
                        const long long dbg_t3 = (OSML_TC_COUNTERS && p.counters) ? clock64() : 0;
                        ptx::umma_commit_pair(&empty_bar[stage]);      // both CTAs may refill this stage
                        ptx::umma_commit_pair(&tmem_full[as]);         // both CTAs' epilogues may read the tile
                        if ((OSML_TC_COUNTERS && p.counters)) { const long long t4 = clock64(); dbg_commit += t4 - dbg_t3; dbg_issue += dbg_t3 - dbg_t2; }
                        if (++stage == (uint32_t)p.stages) { stage = 0; phase ^= 1; }
                    }
                    ptx::umma_commit_pair(a_empty);                  // one arrival per issuer: all MMAs on the query tile retired
                }
                if ((OSML_TC_COUNTERS && p.counters) && issuer == 0) {
                    atomicAdd(p.dbg + 0, (unsigned long long)dbg_wait_acc);
                    atomicAdd(p.dbg + 1, (unsigned long long)dbg_wait_ops);
                    atomicAdd(p.dbg + 2, (unsigned long long)(clock64() - dbg_t0));
                    atomicAdd(p.dbg + 13, (unsigned long long)dbg_issue);
                    atomicAdd(p.dbg + 14, (unsigned long long)dbg_commit);
                }
            }
        }
    } else {
        ptx::setmaxnreg_inc<224>();
        if (warp < 4 + p.epi_warps) {
            // ===== epilogue (both CTAs): own 128 query rows x 256 columns of every accumulator tile =====
            const int ew = warp - 4;
            const int quarter = ew & 3, half = ew >> 2, halves = p.epi_warps >> 2;
            const int cols = kTcBN / halves;
            const int m = quarter * 32 + lane;
            uint32_t acc_it = 0, hslot = 0, hphase = 0;
            // row indices of the list entries live in shared memory: an insert touches K' value registers only (2
            // instead of 3 instructions per entry: -2 % kernel cycles at K' = 16) and K' = 32 fits next to the staged tile
            constexpr bool kSmemRows = true;
            TcRegList<KP, kSmemRows> list;
            list.rows_addr = ptx::smem_u32(row_lists) + (uint32_t)(ew * 32 + lane) * 4u;
            list.rows_stride = (uint32_t)epi_threads * 4u;
            list.attach(ptx::smem_u32(cand_buf) + (uint32_t)(ew * 32 + lane) * 8u, (uint32_t)epi_threads * 8u);
            // loop-invariant shared-memory addresses, kept in registers
            const uint32_t bar_tmem_full = ptx::smem_u32(tmem_full);                          // + 8 * stage
            const uint32_t bar_tmem_empty = ptx::mapa_u32(ptx::smem_u32(tmem_empty), 0);      // the leader's, + 8 * stage
            const uint32_t bar_h_full = ptx::smem_u32(h_full), bar_h_empty = ptx::smem_u32(h_empty);   // + 8 * slot
            const uint32_t h_base = ptx::smem_u32(h_ring) + (uint32_t)(half * cols) * 4u;      // h of this thread's columns
            const uint32_t hm_base = ptx::smem_u32(h_ring) + kTcBN * 4 + (uint32_t)(half * cols / 8) * 4u;   // their group minima
            const uint32_t hmin_base = ptx::smem_u32(h_ring) + (kTcBN + kTcBN / 8 + (cols == 128 ? half : 2)) * 4;  // their minimum
            const uint32_t taddr0 = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(half * cols);
            const uint32_t lane0 = lane == 0 ? 1u : 0u;
            // made opaque to the compiler: it would otherwise re-derive each of these from %tid / the kernel parameters
            // once per tile (S2R, constant-bank loads and dependent integer chains in front of the TMEM loads)
            {
                uint32_t* keep[] = {const_cast<uint32_t*>(&bar_tmem_full), const_cast<uint32_t*>(&bar_tmem_empty),
                                    const_cast<uint32_t*>(&bar_h_full), const_cast<uint32_t*>(&bar_h_empty),
                                    const_cast<uint32_t*>(&h_base), const_cast<uint32_t*>(&hm_base),
                                    const_cast<uint32_t*>(&hmin_base), const_cast<uint32_t*>(&taddr0),
                                    const_cast<uint32_t*>(&lane0)};
#pragma unroll
                for (uint32_t* r : keep) asm volatile("mov.b32 %0, %0;" : "+r"(*r));
            }
            long long dbg_epi_wait = 0, dbg_epi_drain = 0, dbg_epi_hits = 0, dbg_epi_hit_tiles = 0, dbg_epi_load = 0, dbg_big_cyc = 0, dbg_big_cnt = 0, dbg_mid_cyc = 0, dbg_mid_cnt = 0;
            const long long dbg_e0 = clock64();
            TcItemIter items(p, cluster_id, nclusters);
            for (TcItem wi; items.next(p, wi);) {
                const int seg = wi.slot, nsg = wi.stride, qt = 2 * wi.qp + (int)rank;
                const int tile0 = wi.tile0, ntiles = wi.ntiles;
                // A split query starts from the best threshold the other slots of the same query have proven so far (its
                // gate, see tc_gate_enc) and keeps following it, one relaxed load per tile, issued in front of the wait
                // for the accumulator; its own threshold is published whenever it has grown.
                const size_t q = (size_t)qt * kTcBM + m;
                const bool live = p.gate != nullptr && (p.live_all || nsg > 1 || (p.nseg > 1 && wi.qp >= p.nq_single));   // warp-uniform
                unsigned int* const gate_addr = p.gate + q;
                list.reset(-INFINITY, p.kactive);
                if (live) list.raise_gate(tc_gate_dec(tc_gate_load(gate_addr)));
                float published = list.gate;
                if (OSML_TC_DBG(p) == 4) list.thr = list.eff = INFINITY;       // ablation: fast path only
                int row_base = tile0 * kTcBN + half * cols;
                for (int t = 0; t < ntiles; ++t, ++acc_it, row_base += nsg * kTcBN) {
                    const uint32_t as = acc_it & 1, aphase = (acc_it >> 1) & 1;
                    // list maintenance rides on the wait for the next accumulator tile
                    const long long dbg_w0 = (OSML_TC_COUNTERS && p.counters) ? clock64() : 0;
                    uint32_t gate_bits = 0;
                    if (live) {
                        if (list.thr > published) { tc_gate_publish(gate_addr, list.thr); published = list.thr; }
                        gate_bits = tc_gate_load(gate_addr);
                    }
                    while (!__all_sync(0xffffffffu, ptx::mbar_try_wait_addr(bar_tmem_full + as * 8, aphase)))
                        if (list.pending()) list.flush_round();
                    if ((OSML_TC_COUNTERS && p.counters)) dbg_epi_wait += clock64() - dbg_w0;
                    ptx::tc_fence_after();
                    const uint32_t taddr = taddr0 + as * kTcBN;
                    // the tile's h block in shared memory (landed long ago: it was requested with the operand tiles)
                    // (128-column shares check its barrier under the TMEM load latency instead: -2 % K1 at C4, A/B)
                    if (cols != 128) ptx::mbar_wait_addr(bar_h_full + hslot * 8, hphase);
                    const uint32_t hs = hslot * (kTcHBlock * 4);
                    bool released = false;
                    if (OSML_TC_DBG(p) != 1 && OSML_TC_DBG(p) != 3) {
                        if (cols == 128) {
                            // the thread's whole 128-column share of the tile moves to registers in one go and the
                            // accumulator stage goes straight back to the MMA: TMEM is held for the load time only, and
                            // the scan (with its irregular hit handling) gets a full extra tile time of slack
                            float v0[32], v1[32], v2[32], v3[32];
                            ptx::tmem_ld_32x32(taddr, v0);
                            ptx::tmem_ld_32x32(taddr + 32, v1);
                            ptx::tmem_ld_32x32(taddr + 64, v2);
                            ptx::tmem_ld_32x32(taddr + 96, v3);
                            ptx::mbar_wait_addr(bar_h_full + hslot * 8, hphase);     // under the TMEM load latency
                            const float hmin = lds_f32(hmin_base + hs);      // min(h) over the thread's 128 columns
                            ptx::tmem_ld_wait();
                            if (live) list.raise_gate(tc_gate_dec(gate_bits));
                            if ((OSML_TC_COUNTERS && p.counters)) dbg_epi_load += clock64() - dbg_w0 - 0;
                            ptx::tc_fence_before();
                            __syncwarp();
                            if (lane0) ptx::mbar_arrive_cluster_addr(bar_tmem_empty + as * 8);
                            released = true;
                            if (OSML_TC_DBG(p) != 2) {
                                // hot path: one max tree over the 128 raw dot products, one compare against
                                // thr + min(h), one vote.  (The rows are sorted by h, so min(h) is tight for the whole
                                // tile and the test rarely fires without a real candidate.)
                                const float m0 = tc_max32(v0), m1 = tc_max32(v1), m2 = tc_max32(v2), m3 = tc_max32(v3);
                                const float tile_max = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
                                const float bar = list.eff + hmin;
                                if (__any_sync(0xffffffffu, tile_max > bar)) {
                                    // which 32-column chunks (warp-uniform), then which groups of 8 inside those only
                                    const uint32_t cmask = __reduce_or_sync(0xffffffffu, (m0 > bar ? 1u : 0u) | (m1 > bar ? 2u : 0u) |
                                                                                         (m2 > bar ? 4u : 0u) | (m3 > bar ? 8u : 0u));
                                    uint32_t bits = 0;
                                    if (cmask & 1u) bits |= list.lane_bits(v0, lds_f4(hm_base + hs));
                                    if (cmask & 2u) bits |= list.lane_bits(v1, lds_f4(hm_base + hs + 16)) << 4;
                                    if (cmask & 4u) bits |= list.lane_bits(v2, lds_f4(hm_base + hs + 32)) << 8;
                                    if (cmask & 8u) bits |= list.lane_bits(v3, lds_f4(hm_base + hs + 48)) << 12;
                                    const uint32_t mask = __reduce_or_sync(0xffffffffu, bits);
                                    const long long dbg_d0 = (OSML_TC_COUNTERS && p.counters) ? clock64() : 0;
                                    list.template drain<true>(mask, v0, v1, v2, v3, row_base, nullptr, h_base + hs);
                                    if ((OSML_TC_COUNTERS && p.counters)) {
                                        const long long dd = clock64() - dbg_d0;
                                        dbg_epi_drain += dd; dbg_epi_hits += __popc(mask); ++dbg_epi_hit_tiles;
                                        if (dd > 3000) { dbg_big_cyc += dd; ++dbg_big_cnt; }
                                        else if (dd > 1000) { dbg_mid_cyc += dd; ++dbg_mid_cnt; }
                                    }
                                }
                            } else if (v0[0] + v1[31] + v2[5] + v3[7] == 12345.678f) list.thr = v0[0];
                        } else {
                            // two register buffers; the accumulator stage is handed back to the MMA as soon as the last
                            // TMEM load has landed, the last two chunks are scanned after the release
                            float va[32], vb[32];
                            ptx::tmem_ld_32x32(taddr, va);
                            ptx::tmem_ld_32x32(taddr + 32, vb);
                            if (live) list.raise_gate(tc_gate_dec(gate_bits));
                            for (int c = 0; c < cols; c += 64) {
                                const float4 ha = lds_f4(hm_base + hs + (c >> 1)), hb = lds_f4(hm_base + hs + (c >> 1) + 16);
                                ptx::tmem_ld_wait();
                                const bool last = c + 64 >= cols;
                                if (last) {
                                    ptx::tc_fence_before();
                                    __syncwarp();
                                    if (lane0) ptx::mbar_arrive_cluster_addr(bar_tmem_empty + as * 8);
                                    released = true;
                                }
                                if (OSML_TC_DBG(p) != 2) list.template scan<true>(va, row_base + c, ha, nullptr, h_base + hs + c * 4);
                                else if (va[0] + va[31] == 12345.678f) list.thr = va[0];
                                if (!last) ptx::tmem_ld_32x32(taddr + c + 64, va);
                                if (OSML_TC_DBG(p) != 2) list.template scan<true>(vb, row_base + c + 32, hb, nullptr, h_base + hs + (c + 32) * 4);
                                else if (vb[0] + vb[31] == 12345.678f) list.thr = vb[0];
                                if (!last) ptx::tmem_ld_32x32(taddr + c + 96, vb);
                            }
                        }
                    }
                    if (!released) {
                        ptx::tc_fence_before();
                        __syncwarp();
                        if (lane0) ptx::mbar_arrive_cluster_addr(bar_tmem_empty + as * 8);
                    }
                    __syncwarp();
                    if (lane0) ptx::mbar_arrive_addr(bar_h_empty + hslot * 8);
                    if (++hslot == (uint32_t)kTcHSlots) { hslot = 0; hphase ^= 1; }
                }
                list.flush();
                const size_t slot = (size_t)seg * halves + half;
                float* ov = p.cand_val + (q * p.nslot + slot) * KP;
                int* oi = p.cand_idx + (q * p.nslot + slot) * KP;
                int filled = 0;
    #pragma unroll
                for (int e = 0; e < KP; ++e) { const int r = list.row_of(e); ov[e] = list.lv[e]; oi[e] = r; filled += r >= 0 ? 1 : 0; }
                // everything this slot dropped scores <= its own threshold (once the list is full) or <= the last gate it used
                const float bound = fmaxf(filled >= p.kactive ? list.thr : -INFINITY, list.gate);
                p.cand_thr[q * p.nslot + slot] = bound;
                if (live && filled >= p.kactive && list.thr > published) tc_gate_publish(gate_addr, list.thr);
            }
            if ((OSML_TC_COUNTERS && p.counters) && lane == 0) {
                atomicAdd(p.dbg + 3, (unsigned long long)dbg_epi_wait);
                atomicAdd(p.dbg + 4, (unsigned long long)dbg_epi_drain);
                atomicAdd(p.dbg + 5, (unsigned long long)dbg_epi_hits);
                atomicAdd(p.dbg + 6, (unsigned long long)dbg_epi_hit_tiles);
                atomicAdd(p.dbg + 7, (unsigned long long)(clock64() - dbg_e0));
                atomicAdd(p.dbg + 8, (unsigned long long)dbg_epi_load);
                atomicAdd(p.dbg + 9, (unsigned long long)dbg_big_cyc);
                atomicAdd(p.dbg + 10, (unsigned long long)dbg_big_cnt);
                atomicAdd(p.dbg + 11, (unsigned long long)dbg_mid_cyc);
                atomicAdd(p.dbg + 12, (unsigned long long)dbg_mid_cnt);
            }
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
    double inv_scale;     // 1 / scale of the filter copy: scores, norms and residuals above are in scaled units
};

constexpr int kRescoreWarps = 4;
constexpr int kRescoreWarpsRetry = 2;

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
__device__ __forceinline__ double tc_exact_distance(const double* __restrict__ t, const double* __restrict__ q, int d,
                                                    int sum_order) {
    return exact_distance<double>(t, q, d, sum_order);
}

// Exact distances of the candidate rows work[j0 .. j1) in numpy's left-to-right order (sum_order == 1), FOUR LANES PER
// ROW.  With a lane per row every 16-byte load of a warp touches one cache line per lane, and K3 is bound by those L1
// wavefronts (47 M per call at 10 000 x 784, profiles/r2_c3_k1_k3_ncu.md).  Here a lane loads 32 contiguous bytes of its
// row per trip, the four lanes of a group 128: a quarter of the wavefronts per row.  The sum stays ONE sequential chain
// per row, exactly the oracle's: the partial sum travels lane 0 -> 1 -> 2 -> 3 -> lane 0 of the next trip by shuffles,
// and each lane adds its eight (four for doubles) squared differences in order when the sum is with it.  Rows shorter
// than a whole trip are padded with zeros, which a sum of squares absorbs exactly (x + 0 == x for x >= +0, NaN, inf).
// Needs rows that are a whole number of 16-byte pieces; eight rows per pass of the warp.
template <typename X> struct TcVec16;
template <> struct TcVec16<float> { using type = float4; };
template <> struct TcVec16<double> { using type = double2; };
__device__ __forceinline__ void tc_sq_diffs(const float4 t, const float4 q, float* e) {
    using E = Elem<float>;
    e[0] = E::sq_diff(t.x, q.x); e[1] = E::sq_diff(t.y, q.y); e[2] = E::sq_diff(t.z, q.z); e[3] = E::sq_diff(t.w, q.w);
}
__device__ __forceinline__ void tc_sq_diffs(const double2 t, const double2 q, double* e) {
    using E = Elem<double>;
    e[0] = E::sq_diff(t.x, q.x); e[1] = E::sq_diff(t.y, q.y);
}
template <typename X>
__device__ __forceinline__ void tc_exact_distances_seq4(const X* __restrict__ train, const int* __restrict__ idx,
                                                        const short* __restrict__ work, int j0, int j1,
                                                        const X* __restrict__ qrow, int d, X* __restrict__ dist, int lane) {
    using E = Elem<X>;
    using V = typename TcVec16<X>::type;
    constexpr int kPer = 16 / (int)sizeof(X);       // elements per 16-byte piece
    const int sub = lane & 3, grp = lane >> 2;
    const int nv = d / kPer;                         // pieces per row
    const int nfull = nv >> 3;                       // trips in which all four lanes hold two pieces (a group covers 8)
    const int ntrip = (nv + 7) >> 3;
    V zero;
    memset(&zero, 0, sizeof(V));
    const V* q0 = reinterpret_cast<const V*>(qrow) + sub * 2;
    const uint32_t src = (uint32_t)((lane & ~3) | ((sub + 3) & 3));      // the lane in front of this one in the chain
    // kept opaque: the compiler would otherwise rebuild these from %tid and the kernel parameters on every trip
    {
        int* keep[] = {const_cast<int*>(&sub), const_cast<int*>(&nv), const_cast<int*>(&nfull), const_cast<int*>(&ntrip)};
#pragma unroll
        for (int* r : keep) asm volatile("mov.b32 %0, %0;" : "+r"(*r));
        asm volatile("mov.b32 %0, %0;" : "+r"(*const_cast<uint32_t*>(&src)));
    }
    for (int jb = j0; jb < j1; jb += 8) {
        const int j = jb + grp;
        const bool valid = j < j1;                   // idle groups re-read the first row of the pass (same cache lines)
        const V* tp = reinterpret_cast<const V*>(train + (size_t)idx[work[valid ? j : j0]] * d) + sub * 2;
        const V* qp = q0;
        X acc = (X)0;
        V ta, tb, qa, qb;
        auto load_checked = [&](int f) {             // the one trip at the end of a row that is not a whole trip long
            ta = f < nv ? __ldg(tp) : zero; tb = f + 1 < nv ? __ldg(tp + 1) : zero;
            qa = f < nv ? __ldg(qp) : zero; qb = f + 1 < nv ? __ldg(qp + 1) : zero;
        };
        if (nfull > 0) { ta = __ldg(tp); tb = __ldg(tp + 1); qa = __ldg(qp); qb = __ldg(qp + 1); }
        else load_checked(sub * 2);
#pragma unroll 1
        for (int it = 0; it < ntrip; ++it) {
            X e[2 * kPer];
            tc_sq_diffs(ta, qa, e);
            tc_sq_diffs(tb, qb, e + kPer);
            tp += 8; qp += 8;                        // next trip's pieces are on their way while the chain runs
            if (it + 1 < nfull) { ta = __ldg(tp); tb = __ldg(tp + 1); qa = __ldg(qp); qb = __ldg(qp + 1); }
            else if (it + 1 < ntrip) load_checked((it + 1) * 8 + sub * 2);
#pragma unroll
            for (int s = 0; s < 4; ++s) {
                const X in = __shfl_sync(0xffffffffu, acc, src);
                if (sub == s) {
                    X a = in;
#pragma unroll
                    for (int i = 0; i < 2 * kPer; ++i) a = E::add(a, e[i]);
                    acc = a;
                }
            }
        }
        if (valid && sub == 3) dist[j] = E::finish(acc);
    }
}

// One warp per query.  Exact distances are only computed where they can matter: first for the k+1 candidates with the
// best filter score, whose largest exact distance D0 caps the answer; any other candidate whose score-derived lower
// bound already exceeds D0 cannot be among the k+1 nearest and is never gathered from HBM.  Work is kept proportional
// to the candidates that matter (a few dozen) even when many (segment, half) slots feed the query: the k+1 best scores
// are picked by k+1 warp arg-max rounds, and everything after that runs on compacted lists.
// kCand: candidates per query the shared-memory lists hold; kWarps: queries (= warps) per CTA; X: element type of the
// train / query matrix (float or double) -- distances come out as X
template <int kCand, int kWarps, typename X>
__global__ void __launch_bounds__(32 * kWarps, sizeof(X) == 4 ? OSML_K3_MIN_CTAS : 5)     // (float64 would spill at 48 registers)
knn_rescore_kernel(const X* __restrict__ train, int64_t n, const X* __restrict__ queries, int d, int sum_order,
                   int64_t q_count, int ksel, const float* __restrict__ cand_val, const int* __restrict__ cand_idx,
                   const float* __restrict__ cand_thr, const int* __restrict__ perm, int64_t npad, int nslot,
                   int64_t q_single, int slots_single, int kprime, const double* __restrict__ qnorm2,
                   const double* __restrict__ qerr, TcBounds bounds, int64_t index_base, X* __restrict__ out_d,
                   int64_t* __restrict__ out_i, int* __restrict__ flags, const unsigned char* __restrict__ force) {
    __shared__ float s_acc[kWarps][kCand];
    __shared__ int s_idx[kWarps][kCand];
    __shared__ X s_dist[kWarps][kCand];           // exact distances, by position in s_work
    __shared__ short s_work[kWarps][kCand];       // compacted candidate positions: selected, then the rest
    __shared__ X s_best[kWarps];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t q = (int64_t)blockIdx.x * kWarps + warp;
    if (q >= q_count) return;
    const float* cv = cand_val + (size_t)q * nslot * kprime;
    const int* ci = cand_idx + (size_t)q * nslot * kprime;
    const float* ct = cand_thr + (size_t)q * nslot;
    if (q < q_single) nslot = slots_single;           // this query was swept in one item: only its first slots exist
    const int ncand = nslot * kprime;
    const X* qrow = queries + (size_t)q * d;
    float* acc = s_acc[warp];
    X* dist = s_dist[warp];
    int* idx = s_idx[warp];
    short* work = s_work[warp];

    const double n2 = qnorm2[q], eq = qerr[q];
    const double eps = bounds.c_mma * (sqrt(n2) * bounds.tnorm_max + 0.5 * bounds.tnorm_max * bounds.tnorm_max);
    // every row whose filter score is <= a has an exact (oracle-arithmetic) distance >= lower(a)
    auto lower = [&](float a) {
        const double l2 = n2 - 2.0 * ((double)a + eps);
        return (sqrt(fmax(l2, 0.0)) - (eq + bounds.et_max)) * bounds.inv_scale * (1.0 - bounds.gamma);
    };

    // Selection keys live in registers: the order-preserving bits of the filter score with the candidate number in the
    // low bits (ties and the dropped score bits only decide WHICH rows are looked at first -- any k+1 exact distances
    // cap the answer), 0 = empty / already picked.  One warp arg-max is then a register max tree and one REDUX.
    constexpr int kPerLane = kCand / 32;
    constexpr uint32_t kIdMask = (uint32_t)kCand - 1u;
    static_assert((kCand & (kCand - 1)) == 0, "candidate numbers are packed into the low bits of the keys");
    uint32_t key[kPerLane];
    int nvalid = 0;
#pragma unroll
    for (int i = 0; i < kPerLane; ++i) {
        const int c = lane + 32 * i;
        key[i] = 0u;
        if (c < ncand) {
            const int pos = ci[c];                    // position in the (norm-sorted) filter copy; negative: empty slot
            const int row = (pos >= 0 && pos < npad) ? perm[pos] : -1;    // train row; -1 for padding positions
            const bool valid = row >= 0 && row < n;
            const float score = cv[c];
            idx[c] = valid ? row : kIdxSentinel;
            acc[c] = valid ? score : -INFINITY;
            nvalid += valid ? 1 : 0;
            if (valid) key[i] = max(tc_gate_enc(score) & ~kIdMask, kIdMask + 1u) | (kIdMask - (uint32_t)c);
        }
    }
    for (int off = 16; off > 0; off >>= 1) nvalid += __shfl_xor_sync(0xffffffffu, nvalid, off);
    __syncwarp();
    bool ok = nvalid >= ksel;
    // rows that are a whole number of 16-byte pieces, summed left to right: four lanes per row (tc_exact_distances_seq4)
    const bool coop = sum_order == 1 && (d % (16 / (int)sizeof(X))) == 0;
    // pass 1: the best-scoring candidates -- k+1 of them, or the next multiple of the eight rows a cooperative pass takes
    const int nsel = min(coop ? ((ksel + 7) & ~7) : ksel, nvalid);
    for (int s = 0; s < nsel; ++s) {
        uint32_t mx = key[0];
#pragma unroll
        for (int i = 1; i < kPerLane; ++i) mx = max(mx, key[i]);
        const uint32_t best = __reduce_max_sync(0xffffffffu, mx);
        const int c = (int)(kIdMask - (best & kIdMask));
        const int mine = (c & 31) == lane ? (c >> 5) : -1;
#pragma unroll
        for (int i = 0; i < kPerLane; ++i) key[i] = i == mine ? 0u : key[i];
        if (mine >= 0) acc[c] = __int_as_float(0x7fc00000);       // picked (K1 never lists a NaN score: they fail `> thr`)
        if (lane == 0) work[s] = (short)c;
    }
    __syncwarp();
    if (coop) tc_exact_distances_seq4<X>(train, idx, work, 0, nsel, qrow, d, dist, lane);
    else
        for (int j = lane; j < nsel; j += 32)
            dist[j] = tc_exact_distance(train + (size_t)idx[work[j]] * d, qrow, d, sum_order);
    __syncwarp();
    // the (k+1)-th smallest of these exact distances caps the answer
    for (int j = lane; j < nsel; j += 32) {
        const X dj = dist[j];
        int rank = 0;
        for (int o = 0; o < nsel; ++o) rank += key_less(dist[o], o, dj, j) ? 1 : 0;
        if (rank == ksel - 1) s_best[warp] = dj;
    }
    __syncwarp();
    const X d0 = ok ? s_best[warp] : (X)0;
    const bool d0_nan = d0 != d0 || !ok;
    __syncwarp();
    const double cap = d0_nan ? (double)INFINITY : (double)d0 * (1.0 + 9.5367431640625e-07);
    // pass 2: whatever the cap cannot rule out, compacted so that the gathers are spread over the lanes
    int m = nsel;
#pragma unroll
    for (int i = 0; i < kPerLane; ++i) {
        if (32 * i < ncand) {                         // warp-uniform
            const int c = 32 * i + lane;
            const float a = c < ncand ? acc[c] : 0.0f;
            const bool need = c < ncand && a == a && idx[c] != kIdxSentinel && !(lower(a) > cap);   // valid, not picked in pass 1
            const uint32_t ballot = __ballot_sync(0xffffffffu, need);
            if (need) work[m + __popc(ballot & ((1u << lane) - 1u))] = (short)c;
            m += __popc(ballot);
        }
    }
    __syncwarp();
    if (coop) tc_exact_distances_seq4<X>(train, idx, work, nsel, m, qrow, d, dist, lane);
    else
        for (int j = nsel + lane; j < m; j += 32)
            dist[j] = tc_exact_distance(train + (size_t)idx[work[j]] * d, qrow, d, sum_order);
    __syncwarp();
    // rank the computed candidates by (distance, index); keys are unique because the slots cover disjoint row sets
    for (int j = lane; j < m; j += 32) {
        const X dj = dist[j];
        const int ij = idx[work[j]];
        int rank = 0;
        for (int o = 0; o < m; ++o) rank += key_less(dist[o], idx[work[o]], dj, ij) ? 1 : 0;
        if (rank < ksel) {
            out_d[q * ksel + rank] = dj;
            out_i[q * ksel + rank] = index_base + ij;
            if (rank == ksel - 1) s_best[warp] = dj;
        }
    }
    __syncwarp();
    if (ok) {
        const double dk = (double)s_best[warp] * (1.0 + 9.5367431640625e-07);   // (k+1)-th exact distance, 2^-20 margin
        for (int s = lane; s < nslot; s += 32) {
            const float thr = ct[s];
            if (thr == -INFINITY) continue;   // the slot never filled its list: nothing was dropped
            if (!(lower(thr) > dk)) ok = false;    // also catches NaN
        }
    }
    ok = __all_sync(0xffffffffu, ok);
    if (force != nullptr && force[q]) ok = false;     // the caller knows this query needs the exact direct kernel
    if (lane == 0) flags[q] = ok ? 0 : 1;
}

// ---- fallback plumbing ------------------------------------------------------------------------------------------------
// appends the flagged queries of one block (block-local flags, first query = q_offset of the call) to the call's list
__global__ void tc_compact_flags_kernel(const int* __restrict__ flags, int64_t q_count, int64_t q_offset,
                                        int* __restrict__ list, int* __restrict__ count) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q < q_count && flags[q]) list[atomicAdd(count, 1)] = (int)(q + q_offset);
}

template <typename X>
__global__ void tc_gather_rows_kernel(const X* __restrict__ src, const int* __restrict__ list, int count, int d,
                                      X* __restrict__ dst, const unsigned char* __restrict__ force_src,
                                      unsigned char* __restrict__ force_dst) {
    const int r = blockIdx.x;
    if (r >= count) return;
    const X* s = src + (size_t)list[r] * d;
    for (int j = threadIdx.x; j < d; j += blockDim.x) dst[(size_t)r * d + j] = s[j];
    if (force_src != nullptr && threadIdx.x == 0) force_dst[r] = force_src[list[r]];
}

template <typename X>
__global__ void tc_scatter_lists_kernel(const X* __restrict__ src_d, const int64_t* __restrict__ src_i,
                                        const int* __restrict__ list, int count, int ksel, X* __restrict__ dst_d,
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
    double et_max = 0.0, tnorm_max = 0.0;   // in units of the scaled filter copy
    float scale = 1.0f;                     // power of two applied to t - c and q - c before rounding
    bool f16 = false;                       // element type of the filter copies (see tc_round)
    bool want_f16 = false;                  // set when a call needed too many second passes: rebuild in fp16 next time
    TcBuffer train_bf16, half_norm, hmin8, hblk, perm, center, stats, partial;
    TcBuffer dbg;
    // per-call scratch; level 0 = the caller's queries, level 1 = the retry of the few queries K3 could not prove
    struct Scratch {
        TcBuffer q_bf16, q_norm2, q_err, cand_val, cand_idx, cand_thr, gate, flags, fb_list, fb_count, fb_queries, fb_d, fb_i, fb_force;
        size_t bytes() const {
            return q_bf16.cap + q_norm2.cap + q_err.cap + cand_val.cap + cand_idx.cap + cand_thr.cap + gate.cap + flags.cap +
                   fb_list.cap + fb_count.cap + fb_queries.cap + fb_d.cap + fb_i.cap + fb_force.cap;
        }
        void release() {
            for (TcBuffer* b : {&q_bf16, &q_norm2, &q_err, &cand_val, &cand_idx, &cand_thr, &gate, &flags, &fb_list, &fb_count,
                                &fb_queries, &fb_d, &fb_i, &fb_force})
                b->release();
        }
    };
    Scratch scratch[2];
    int64_t retry_queries = 0;    // queries that went through the level-1 retry since the last reset
    CUtensorMap map_b;
    CUtensorMap map_b_half;   // 128-row boxes for the CTA-pair kernel
    void* encode_fn = nullptr;
    int sm_count = 0;
    int smem_optin = 0;
    std::vector<cudaEvent_t> ev_k1;   // (begin, end) per level-0 K1 launch of the call in flight
    int ev_k1_used = 0;
    double filter_ms = 0.0;       // device time of the K1 launches of the last neighbour search (CUDA events)
    int64_t filter_launches = 0;
};

struct TensorLaunch {
    const void* train; int dtype; int64_t n, d; int k; int sum_order; int64_t index_base; int slack;
    int opt_epi_warps = 0;   // 0 = auto, else 4 or 8
    int opt_nseg = 0;        // 0 = auto
    int opt_smem_lists = 0;  // 1 = force the shared-memory list epilogue
    int opt_cta_group = 0;   // 0 = auto, 1 = single CTA per item, 2 = CTA pairs (cta_group::2)
    int opt_active = 0;      // 0 = auto (= K'), else list entries in use per slot (experiments)
    int opt_precision = 0;   // 0 = auto (bf16, fp16 once a call needed many second passes), 1 = bf16, 2 = fp16
    int opt_kprime = 0;      // 0 = auto (tc_kprime), 16 / 32 = list length per slot (experiments: with more segments a slot's
                             //   share of the true neighbours shrinks; whatever a short list loses, the proof catches)
    int opt_hot = 0;         // split items: 0 = every segment interleaves; n = the first n segments interleave over the
                             //   small-norm tiles and the rest take contiguous runs (TcItemIter; experiments)
    int opt_live = 1;        // 0 = split items neither publish nor follow their query's gate (A/B)
    int opt_wide_pair = 1;   // 0 = rows too wide for the resident-tile pair kernel go to the single-CTA kernel (A/B)
    int debug = 0;
    int counters = 0;        // 1 = the CTA-pair kernel records where its issuer / epilogue warps wait (printed to stderr)
    // A pipelined host call hands over its queries in blocks: block b covers queries [block_ends[b-1], block_ends[b]) and
    // only block 0 is on the device when the search starts.  before_block(b), b >= 1, runs once block b-1's kernels are
    // queued: the caller uploads block b there (under block b-1's K1) and makes `stream` wait for it; 0 = ok, else a status.
    const std::function<int(int)>* before_block = nullptr;
    const int64_t* block_ends = nullptr;
    int n_blocks = 0;
    // one byte per query of the call (device, nullable): non-zero = this query must go to the exact direct kernel whatever
    // the completeness proof says (integer matrices: a query outside the range in which numpy's arithmetic cannot wrap)
    const unsigned char* q_force = nullptr;
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

// Larger k (up to 250): register lists of 32 per slot and at least `min_seg` interleaved train segments, so that the
// 2 * min_seg slots of a query hold k + 1 + slack candidates between them and each slot's share of the true k + 1
// nearest -- Binomial(k + 1, 1 / (2 * min_seg)) -- fits its list with four standard deviations to spare.  K3 takes up
// to 512 candidates per query (8 segments) in the first pass.  Returns the list length, or a value above kTcMaxKPrime
// when the filter cannot serve this k.
inline int tc_plan_lists(int k, int slack, int* min_seg) {
    *min_seg = 1;
    const int kprime = tc_kprime(k, slack);
    if (kprime <= kTcMaxKPrime) return kprime;
    if (slack < 0) slack = tc_default_slack(k);
    const int need = k + 1 + slack;
    for (int seg = 2; seg <= kTcMaxCand / 64; ++seg) {
        const double p = 0.5 / seg, mu = (k + 1) * p, sd = std::sqrt((k + 1) * p * (1.0 - p));
        if (2 * seg * 32 >= need && mu + 4.0 * sd <= 32.0) { *min_seg = seg; return 32; }
    }
    return kTcMaxKPrime + 1;
}

inline bool tensor_path_supported(int dtype, int64_t n, int64_t d, int k) {
    if (dtype != 0 /* OSML_KNN_F32 */ && dtype != 1 /* OSML_KNN_F64 */) return false;
    if (n < 256 || d < 1) return false;
    if (d > 8189) return false;                   // 128 K blocks of 64 bf16
    int min_seg = 1;
    if (tc_plan_lists(k, -1, &min_seg) > kTcMaxKPrime) return false;
    return min_seg == 1 || n >= (int64_t)min_seg * 8 * kTcBN;       // a segment is at least 8 tiles long
}

inline size_t tensor_bytes(const TensorState& s) {
    return s.train_bf16.cap + s.half_norm.cap + s.hmin8.cap + s.hblk.cap + s.perm.cap + s.center.cap + s.stats.cap +
           s.partial.cap + s.scratch[0].bytes() + s.scratch[1].bytes();
}

inline void tensor_release(TensorState& s) {
    for (TcBuffer* b : {&s.dbg, &s.train_bf16, &s.half_norm, &s.hmin8, &s.hblk, &s.perm, &s.center, &s.stats, &s.partial})
        b->release();
    s.scratch[0].release();
    s.scratch[1].release();
    for (cudaEvent_t e : s.ev_k1) cudaEventDestroy(e);
    s.ev_k1.clear();
    s.ev_k1_used = 0;
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
        map, st.f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
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

// one-time K0 over the train set (X: float or double train matrix)
template <typename X>
inline int tc_prepare(TensorState& st, const TensorLaunch& tl, cudaStream_t stream, int64_t* launches,
                      std::string* err) {
    const bool f16 = tl.opt_precision == 2 || (tl.opt_precision == 0 && st.want_f16);
    if (st.ready && st.f16 == f16) return 0;
    st.ready = false;                    // first build, or a rebuild of the filter copy in the other element type
    st.f16 = f16;
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
    st.dp = (d + kTcKB - 1) / kTcKB * kTcKB;
    st.npad = (tl.n + kTcBN - 1) / kTcBN * kTcBN;
    const X* train = static_cast<const X*>(tl.train);
    const int rows_per_block = 2048;
    const int nblocks = (int)((tl.n + rows_per_block - 1) / rows_per_block);
    TC_TRY(st.partial.reserve((size_t)nblocks * d * sizeof(double)));
    TC_TRY(st.center.reserve((size_t)d * sizeof(float)));
    TC_TRY(st.stats.reserve(sizeof(TcPrepStats) + 16));
    TC_TRY(st.train_bf16.reserve((size_t)st.npad * st.dp * sizeof(tc_elem)));
    TC_TRY(cudaMemsetAsync(st.stats.ptr, 0, sizeof(TcPrepStats) + 16, stream));
    tc_col_partial_sums_kernel<<<nblocks, dim3(32, 8), 0, stream>>>(train, tl.n, d, rows_per_block,
                                                                    st.partial.as<double>());
    tc_col_mean_kernel<<<(d + 127) / 128, 128, 0, stream>>>(st.partial.as<double>(), nblocks, d, tl.n,
                                                            st.center.as<float>());
    TC_TRY(st.half_norm.reserve((size_t)st.npad * sizeof(float)));
    TC_TRY(st.hmin8.reserve((size_t)(st.npad / 8) * sizeof(float)));
    TC_TRY(st.hblk.reserve((size_t)(st.npad / kTcBN) * kTcHBlock * sizeof(float)));
    TC_TRY(st.perm.reserve((size_t)st.npad * sizeof(int)));
    {
        // sort the rows by h = 0.5*|t^|^2 (fit-time only; CUB radix sort of (key, row id) pairs)
        TcBuffer keys_in, vals_in, temp;
        TC_TRY(keys_in.reserve((size_t)tl.n * sizeof(float)));
        TC_TRY(vals_in.reserve((size_t)tl.n * sizeof(int)));
        st.scale = 1.0f;
        if (st.f16) {
            // power-of-two scale that puts max |t - c| near 2^14: fp16 keeps 11 significant bits down to 2^-14, and
            // queries up to 4x farther out than any train row still fit (beyond that they fall back through K3)
            unsigned int* max_bits = reinterpret_cast<unsigned int*>(st.stats.as<unsigned char>() + sizeof(TcPrepStats));
            tc_max_abs_kernel<<<148 * 8, 256, 0, stream>>>(train, tl.n, d, st.center.as<float>(), max_bits);
            unsigned int hbits = 0;
            TC_TRY(cudaMemcpyAsync(&hbits, max_bits, sizeof(hbits), cudaMemcpyDeviceToHost, stream));
            TC_TRY(cudaStreamSynchronize(stream));
            float max_abs;
            memcpy(&max_abs, &hbits, 4);
            if (max_abs > 0.0f && max_abs < INFINITY) {
                int e = 0;
                std::frexp(max_abs, &e);                          // max_abs = m * 2^e, m in [0.5, 1)
                st.scale = std::ldexp(1.0f, std::max(-100, std::min(100, 14 - e)));
            }
            ++*launches;
        }
        (st.f16 ? tc_train_norms_kernel<true, X> : tc_train_norms_kernel<false, X>)<<<(unsigned)((tl.n * 32 + 255) / 256), 256, 0, stream>>>(
            train, tl.n, d, st.center.as<float>(), st.scale, keys_in.as<float>(), vals_in.as<int>(),
            st.stats.as<TcPrepStats>());
        size_t temp_bytes = 0;
        TC_TRY(cub::DeviceRadixSort::SortPairs(nullptr, temp_bytes, keys_in.as<float>(), st.half_norm.as<float>(),
                                               vals_in.as<int>(), st.perm.as<int>(), (int)tl.n, 0, 32, stream));
        TC_TRY(temp.reserve(temp_bytes));
        TC_TRY(cub::DeviceRadixSort::SortPairs(temp.ptr, temp_bytes, keys_in.as<float>(), st.half_norm.as<float>(),
                                               vals_in.as<int>(), st.perm.as<int>(), (int)tl.n, 0, 32, stream));
        const int64_t threads = st.npad * 32;
        (st.f16 ? tc_prep_rows_kernel<true, true, X> : tc_prep_rows_kernel<true, false, X>)<<<(unsigned)((threads + 255) / 256), 256, 0, stream>>>(
            train, tl.n, st.npad, d, st.dp, st.center.as<float>(), st.scale, st.perm.as<int>(),
            st.train_bf16.as<tc_elem>(), nullptr, nullptr, st.half_norm.as<float>(), st.perm.as<int>());
        tc_hblock_kernel<<<(unsigned)((st.npad / 8 + 255) / 256), 256, 0, stream>>>(
            st.half_norm.as<float>(), st.npad, st.hmin8.as<float>(), st.hblk.as<float>());
        *launches += 6;
        TC_TRY(cudaGetLastError());
        TC_TRY(cudaStreamSynchronize(stream));
        keys_in.release(); vals_in.release(); temp.release();
    }
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

// Filter + rescoring of one block of queries (device, row-major fp32), queued on `stream` without waiting for anything.
// out_d/out_i: the block's [q][k+1] final lists.  Queries K3 could not prove are appended (as q_offset + local number) to
// the level's fb_list / fb_count, which the caller reserved and zeroed; tc_resolve_unproven deals with them afterwards.
template <typename X>
inline int tc_filter_block(TensorState& st, const TensorLaunch& tl, const X* d_queries, int64_t q,
                           X* out_d, int64_t* out_i, cudaStream_t stream, int64_t* launches, std::string* err,
                           int level, int64_t q_offset) {
    TensorState::Scratch& sc = st.scratch[level];
    const int d = (int)tl.d;
    const int ksel = tl.k + 1;
    int min_seg = 1;
    const int planned = tc_plan_lists(tl.k, tl.slack, &min_seg);
    const int kprime = (tl.opt_kprime == 16 || tl.opt_kprime == 32) ? tl.opt_kprime : planned;
    if (kprime > kTcMaxKPrime) { if (err) *err = "k + 1 + slack is beyond what the filter's lists hold"; return 4; }
    const bool reg_lists = kprime <= 32 && !tl.opt_smem_lists;
    int epi_warps = reg_lists ? 8 : 4;
    if (tl.opt_epi_warps == 4 || (tl.opt_epi_warps == 8 && (reg_lists || kprime <= 24))) epi_warps = tl.opt_epi_warps;
    const int halves = epi_warps / 4;
    // CTA pairs (cta_group::2) need register lists (no shared-memory list code in that kernel)
    // (the pair kernel stages whole half-tiles; for wide rows that no longer fits twice, so use single CTAs there)
    const size_t pair_stage = (size_t)(st.dp / kTcKB) * kTcBHalfBytes;
    // pair kernel only: h ring + its barriers, and the lists' row indices
    const size_t pair_extra = (size_t)kTcHSlots * kTcHBlock * 4 + 256 + (size_t)kprime * epi_warps * 32 * 4;
    const bool pair_fits = (size_t)st.smem_optin >=
        (size_t)(st.dp / kTcKB) * kTcABlockBytes + 2 * pair_stage + (size_t)kTcBufCap * 8 * 32 * 8 + pair_extra + 1280;
    // rows too wide for that run the same kernel with a K-block granular ring (kWide): query block + train half block per
    // stage, at least three stages
    const size_t wide_stage = (size_t)kTcABlockBytes + kTcBHalfBytes;
    const bool wide_fits = (size_t)st.smem_optin >= 3 * wide_stage + (size_t)kTcBufCap * 8 * 32 * 8 + pair_extra + 1280;
    const bool pair_wide = reg_lists && tl.opt_cta_group != 1 && !pair_fits && wide_fits && tl.opt_wide_pair != 0;
    const int tc_cta_group = (reg_lists && tl.opt_cta_group != 1 && (pair_fits || pair_wide)) ? 2 : 1;
    const int64_t qtile = (int64_t)kTcBM * tc_cta_group;
    const int64_t qpad = (q + qtile - 1) / qtile * qtile;
    const int nqt = (int)(qpad / kTcBM);
    const int tiles_total = (int)(st.npad / kTcBN);
    const int nkb = st.dp / kTcKB;
    const int ksteps = (d + 15) / 16;

    // Work split.  Every extra segment restarts the per-query lists (more warm-up, more candidates for K3), so a query
    // tile is swept in ONE item whenever that keeps the machine busy: with U persistent units and T query tiles, the
    // first floor(T / U) * U tiles are single-item sweeps and only the remainder -- which would otherwise occupy a
    // fraction of the machine for a whole sweep -- is cut into nseg interleaved segments.  (Single-CTA kernels split
    // every tile the same way: nq_single = 0.)
    // the second filter pass (level 1, a handful of queries) may cut finer: its K3 instance holds twice the candidates
    // (so may a call of one or two query-tile pairs: its K3 cost is nothing, and the more items, the shorter the sweeps)
    const int cand_cap = (level == 0 && q > 512) ? kTcMaxCand : kTcMaxCandRetry;
    const int max_seg_by_cand = std::max(1, cand_cap / (kprime * halves));
    const int max_seg = std::max(1, std::min(max_seg_by_cand, std::max(1, tiles_total / 8)));
    const int units = std::max(1, st.sm_count / tc_cta_group);
    const int nq_units = nqt / tc_cta_group;                 // query tiles (single CTA) or tile pairs
    int nseg = 1, nq_single = 0;
    if (max_seg < min_seg) { if (err) *err = "too few training rows for the segments this k needs"; return 4; }
    if (tl.opt_nseg > 0) {
        nseg = std::max(min_seg, std::min(tl.opt_nseg, max_seg));
    } else {
        // (large k: every pair is split, a whole-sweep item would have two slots only)
        const int rem = (tc_cta_group == 2 && min_seg == 1) ? nq_units % units : nq_units;     // units of work left for the split part
        nq_single = nq_units - rem;
        nseg = min_seg;
        // A handful of pairs (a small call, or a short tail): as many segments as still fit ONE wave of items -- the
        // sweeps are as short as they can be and, sharing a gate, cost little to start (1M x 128 rows, 32 / 256 / 1024
        // queries: 0.47 / 0.50 / 0.50 ms with the rule below, 0.37 / 0.39 / 0.39 ms with 16 segments).  Otherwise weigh
        // the idle share of the last wave against a per-segment cost (measured: 16 pairs are best cut in 4).
        const int one_wave = rem > 0 ? std::min(max_seg, units / rem) : 0;
        if (one_wave >= 8 && one_wave >= min_seg) nseg = one_wave;
        else {
            double best = 1e30;
            for (int cand = min_seg; cand <= max_seg && rem > 0; ++cand) {
                const double cost = std::ceil((double)rem * cand / units) / cand + 0.01 * (cand - 1);
                if (cost < best - 1e-9) { best = cost; nseg = cand; }
            }
        }
        if (nseg == 1) nq_single = tc_cta_group == 2 ? nq_units : 0;         // nothing gained by splitting
    }
    const int tiles_per_seg = (tiles_total + nseg - 1) / nseg;
    const int nslot = nseg * halves;
    // hot-first split (TcItemIter; option `hot` = n): the first n segments interleave over their share of the small-norm
    // tiles, the later ones take contiguous runs of the rest.  Measured at 10 000 x 784 queries against 60 000 rows with the
    // two first-wave segments hot: K1 unchanged (0.695 vs 0.692 ms), but the hot slots' thresholds sit so close to the
    // (k+1)-th neighbour that the proof fails for a few queries per call and the second pass costs 0.2 ms -- so the
    // default keeps every segment interleaved.  Large k (min_seg > 1) must: each slot needs an equal share of the
    // true neighbours.
    int nhot = nseg, thot = tiles_total;
    if (tc_cta_group == 2 && nseg > 1 && min_seg == 1 && nq_units > nq_single) {
        int h = tl.opt_hot > 0 ? tl.opt_hot : nseg;      // default: every segment interleaves (see below)
        h = std::max(1, std::min(h, nseg));
        if (h < nseg) {
            nhot = h;
            thot = std::max(nhot, (int)(((long long)tiles_total * h + nseg / 2) / nseg));
        }
    }

    const size_t list_bytes = reg_lists ? (size_t)kTcBufCap * epi_warps * 32 * 8 : (size_t)kprime * epi_warps * 32 * 8;
    // the query tile stays resident when it leaves room for at least two ring stages; wide rows stream it instead
    const bool a_resident = (tc_cta_group == 2 && !pair_wide) ||
        (tc_cta_group == 1 &&
         (size_t)st.smem_optin >= (size_t)nkb * kTcABlockBytes + list_bytes + 2 * (size_t)kTcBStageBytes + 1280);
    const size_t a_bytes = a_resident ? (size_t)nkb * kTcABlockBytes : 0;
    const size_t stage_bytes = pair_wide ? wide_stage
                               : tc_cta_group == 2 ? (size_t)nkb * kTcBHalfBytes
                               : (a_resident ? (size_t)kTcBStageBytes : (size_t)kTcABlockBytes + kTcBStageBytes);
    const size_t extra = tc_cta_group == 2 ? pair_extra : 0;
    const size_t fixed = a_bytes + list_bytes + extra + 256 + 1024;
    int stages = (int)(((size_t)st.smem_optin - fixed) / stage_bytes);
    stages = std::min(stages, tc_cta_group == 2 ? 8 : 6);
    if (stages < 2) { if (err) *err = "not enough shared memory for the tensor filter"; return 4; }
    const size_t smem = a_bytes + (size_t)stages * stage_bytes + list_bytes + extra + 256;

    TC_TRY(sc.q_bf16.reserve((size_t)qpad * st.dp * sizeof(tc_elem)));
    TC_TRY(sc.q_norm2.reserve((size_t)qpad * sizeof(double)));
    TC_TRY(sc.q_err.reserve((size_t)qpad * sizeof(double)));
    TC_TRY(sc.cand_val.reserve((size_t)qpad * nslot * kprime * sizeof(float)));
    TC_TRY(sc.cand_idx.reserve((size_t)qpad * nslot * kprime * sizeof(int)));
    TC_TRY(sc.cand_thr.reserve((size_t)qpad * nslot * sizeof(float)));
    TC_TRY(sc.gate.reserve((size_t)qpad * sizeof(unsigned int)));
    TC_TRY(sc.flags.reserve((size_t)q * sizeof(int)));

    const int64_t threads = qpad * 32;
    (st.f16 ? tc_prep_rows_kernel<false, true, X> : tc_prep_rows_kernel<false, false, X>)<<<(unsigned)((threads + 255) / 256), 256, 0, stream>>>(
        d_queries, q, qpad, d, st.dp, st.center.as<float>(), st.scale, nullptr, sc.q_bf16.as<tc_elem>(),
        sc.q_norm2.as<double>(), sc.q_err.as<double>(), nullptr, nullptr);
    ++*launches;
    TC_TRY(cudaGetLastError());

    CUtensorMap map_a;
    int rc = tc_make_map(st, &map_a, sc.q_bf16.ptr, qpad, st.dp, kTcBM, err);
    if (rc) return rc;

    TcParams p;
    p.nkb = nkb; p.ksteps = ksteps; p.nqt = nqt; p.nseg = nseg; p.nq_single = nq_single; p.tiles_per_seg = tiles_per_seg;
    p.tiles_total = tiles_total; p.kprime = kprime; p.f16 = st.f16 ? 1 : 0; p.kactive = (tl.opt_active > 0 && tl.opt_active <= kprime) ? tl.opt_active : kprime; p.epi_warps = epi_warps; p.stages = stages; p.nslot = nslot; p.a_resident = a_resident ? 1 : 0;
    p.cand_val = sc.cand_val.as<float>(); p.cand_idx = sc.cand_idx.as<int>();
    p.cand_thr = sc.cand_thr.as<float>();
    p.nhot = nhot; p.thot = thot;
    p.live_all = tl.opt_live == 2 ? 1 : 0;
    p.gate = (tc_cta_group == 2 && (nseg > 1 || p.live_all) && tl.opt_live != 0) ? sc.gate.as<unsigned int>() : nullptr;
    p.h = st.half_norm.as<float>(); p.hmin8 = st.hmin8.as<float>(); p.hblk = st.hblk.as<float>();
    p.debug = OSML_TC_DEBUG ? tl.debug : 0;
    p.counters = tl.counters;
    p.dbg = nullptr;
    if (tl.counters) {
        TC_TRY(st.dbg.reserve(16 * sizeof(unsigned long long)));
        TC_TRY(cudaMemsetAsync(st.dbg.ptr, 0, 16 * sizeof(unsigned long long), stream));
        p.dbg = st.dbg.as<unsigned long long>();
    }
    auto kernel = !reg_lists ? knn_tc_filter_kernel<0> : (kprime == 16 ? knn_tc_filter_kernel<16> : knn_tc_filter_kernel<32>);
    if (tc_cta_group == 2 && pair_wide) kernel = kprime == 16 ? knn_tc_filter_pair_kernel<16, true> : knn_tc_filter_pair_kernel<32, true>;
    else if (tc_cta_group == 2) kernel = kprime == 16 ? knn_tc_filter_pair_kernel<16, false> : knn_tc_filter_pair_kernel<32, false>;
    TC_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int nitems = nq_single + (nq_units - nq_single) * nseg;
    const int grid = tc_cta_group * std::min(nitems, st.sm_count / tc_cta_group);
    if (level == 0 && (size_t)st.ev_k1_used + 2 > st.ev_k1.size()) {
        cudaEvent_t e0 = nullptr, e1 = nullptr;
        TC_TRY(cudaEventCreate(&e0));
        st.ev_k1.push_back(e0);
        TC_TRY(cudaEventCreate(&e1));
        st.ev_k1.push_back(e1);
    }
    if (p.gate) TC_TRY(cudaMemsetAsync(sc.gate.ptr, 0, (size_t)qpad * sizeof(unsigned int), stream));   // 0: nothing published
    if (level == 0) TC_TRY(cudaEventRecord(st.ev_k1[st.ev_k1_used], stream));
    kernel<<<grid, kTcThreads, smem, stream>>>(map_a, tc_cta_group == 2 ? st.map_b_half : st.map_b, p);
    if (level == 0) { TC_TRY(cudaEventRecord(st.ev_k1[st.ev_k1_used + 1], stream)); st.ev_k1_used += 2; }
    ++*launches;
    if (level == 0) ++st.filter_launches;
    TC_TRY(cudaGetLastError());
    if (tl.counters) {      // cycle counters of the CTA-pair kernel (sums over MMA issuers / epilogue warps)
        unsigned long long h[16];
        TC_TRY(cudaMemcpyAsync(h, st.dbg.ptr, sizeof(h), cudaMemcpyDeviceToHost, stream));
        TC_TRY(cudaStreamSynchronize(stream));
        const double issuers = grid / 2.0, ewarps = (double)grid * epi_warps;
        fprintf(stderr, "[K1 dbg] issuer: total %.0f cyc, wait accumulator %.1f%%, wait operands %.1f%% | epilogue warp: total %.0f cyc, "
                        "wait tile %.1f%%, wait+load %.1f%%, drain %.1f%%, hit groups/tile %.3f, tiles with hits %.1f%%, drain cyc/hit tile %.0f\n",
                h[2] / issuers, 100.0 * h[0] / h[2], 100.0 * h[1] / h[2], h[7] / ewarps, 100.0 * h[3] / h[7], 100.0 * h[8] / h[7],
                100.0 * h[4] / h[7], (double)h[5] / ((double)tiles_total * (nqt / tc_cta_group) * 2 * epi_warps),
                100.0 * h[6] / ((double)tiles_total * (nqt / tc_cta_group) * 2 * epi_warps), (double)h[4] / std::max(1ull, h[6]));
        fprintf(stderr, "[K1 dbg] issuer: MMA issue %.1f%%, commits %.1f%%\n", 100.0 * h[13] / h[2], 100.0 * h[14] / h[2]);
        fprintf(stderr, "[K1 dbg] drains: %llu total, %llu > 3000 cyc (%.1f%% of drain time, mean %.0f), %llu in 1000..3000 (%.1f%%, mean %.0f), "
                        "rest mean %.0f\n", h[6], h[10], 100.0 * h[9] / std::max(1ull, h[4]), (double)h[9] / std::max(1ull, h[10]), h[12],
                100.0 * h[11] / std::max(1ull, h[4]), (double)h[11] / std::max(1ull, h[12]),
                (double)(h[4] - h[9] - h[11]) / std::max(1ull, h[6] - h[10] - h[12]));
    }
    if (OSML_TC_DEBUG && tl.debug) return 0;   // profiling ablations: results are meaningless, skip rescoring

    TcBounds b;
    b.et_max = st.et_max; b.tnorm_max = st.tnorm_max; b.inv_scale = 1.0 / (double)st.scale;
    // (K + 8) * 2^-22 accumulation + 2^-22 for rounding h, subtracting it and the group test thr + min(h) +
    // 2^-17 for the slot bits packed into list values
    b.c_mma = (double)(ksteps * 16 + 8) * 2.384185791015625e-07 + 2.4e-7 + 7.62939453125e-06;
    b.gamma = (double)(d + 4) * 1.1920928955078125e-07;                      // (D + 4) * 2^-23
    const bool wide_k3 = nslot * kprime > kTcMaxCand;
    const int k3_warps = wide_k3 ? kRescoreWarpsRetry : kRescoreWarps;
    auto k3 = wide_k3 ? knn_rescore_kernel<kTcMaxCandRetry, kRescoreWarpsRetry, X> : knn_rescore_kernel<kTcMaxCand, kRescoreWarps, X>;
    if (!wide_k3 && nslot * kprime <= 128) k3 = knn_rescore_kernel<128, kRescoreWarps, X>;    // less shared memory: more CTAs per SM
    else if (!wide_k3 && nslot * kprime <= 256) k3 = knn_rescore_kernel<256, kRescoreWarps, X>;
    k3<<<(unsigned)((q + k3_warps - 1) / k3_warps), 32 * k3_warps, 0, stream>>>(
        static_cast<const X*>(tl.train), tl.n, d_queries, d, tl.sum_order, q, ksel, sc.cand_val.as<float>(),
        sc.cand_idx.as<int>(), sc.cand_thr.as<float>(), st.perm.as<int>(), st.npad, nslot,
        (int64_t)nq_single * kTcBM * tc_cta_group, halves, kprime, sc.q_norm2.as<double>(), sc.q_err.as<double>(), b, tl.index_base, out_d,
        out_i, sc.flags.as<int>(), tl.q_force ? tl.q_force + q_offset : nullptr);
    tc_compact_flags_kernel<<<(unsigned)((q + 255) / 256), 256, 0, stream>>>(sc.flags.as<int>(), q, q_offset,
                                                                             sc.fb_list.as<int>(), sc.fb_count.as<int>());
    *launches += 2;
    TC_TRY(cudaGetLastError());
    return 0;
}

template <typename X>
inline int tc_neighbours_block(TensorState& st, const TensorLaunch& tl, const X* d_queries, int64_t q,
                               X* out_d, int64_t* out_i, cudaStream_t stream, int64_t* fallback,
                               int64_t* launches, std::string* err, const DirectFallback& direct, int level);

// Waits for the blocks queued so far and settles the queries on the level's fb_list (numbers relative to d_queries /
// out_d / out_i): level 0 re-runs them through a much finer filter pass, level 1 hands them to the exact direct kernel.
template <typename X>
inline int tc_resolve_unproven(TensorState& st, const TensorLaunch& tl, const X* d_queries, int64_t q,
                               X* out_d, int64_t* out_i, cudaStream_t stream, int64_t* fallback,
                               int64_t* launches, std::string* err, const DirectFallback& direct, int level) {
    TensorState::Scratch& sc = st.scratch[level];
    const int d = (int)tl.d;
    const int ksel = tl.k + 1;
    int rc = 0;
    int count = 0;
    TC_TRY(cudaMemcpyAsync(&count, sc.fb_count.ptr, sizeof(int), cudaMemcpyDeviceToHost, stream));
    TC_TRY(cudaStreamSynchronize(stream));
    if (level == 0) {
        for (int e = 0; e + 1 < st.ev_k1_used; e += 2) {
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, st.ev_k1[e], st.ev_k1[e + 1]) == cudaSuccess) st.filter_ms += ms;
        }
        st.ev_k1_used = 0;
    }
    // many unproven queries mean the bf16 rounding residuals are too large for this data: use fp16 from the next call on
    if (level == 0 && !st.f16 && tl.opt_precision == 0 && q >= 4096 && (int64_t)count * 400 > q) st.want_f16 = true;
    if (count > 0) {
        TC_TRY(sc.fb_queries.reserve((size_t)count * d * sizeof(X)));
        TC_TRY(sc.fb_d.reserve((size_t)count * ksel * sizeof(X)));
        TC_TRY(sc.fb_i.reserve((size_t)count * ksel * sizeof(int64_t)));
        if (tl.q_force) TC_TRY(sc.fb_force.reserve((size_t)count));
        tc_gather_rows_kernel<X><<<count, 128, 0, stream>>>(d_queries, sc.fb_list.as<int>(), count, d,
                                                            sc.fb_queries.as<X>(), tl.q_force,
                                                            tl.q_force ? sc.fb_force.as<unsigned char>() : nullptr);
        ++*launches;
        if (level == 0) {
            // second chance on the filter: the same kernels over the few unproven queries, cut into as many train
            // segments as K3 can take -- every (segment, half) slot keeps its own K' best, so the kept set reaches
            // many times deeper and the proof almost always goes through.  Whatever still fails goes to K2 from there.
            TensorLaunch retry = tl;
            retry.opt_nseg = 1 << 20;                  // clamped to the maximum
            if (tl.q_force) retry.q_force = sc.fb_force.as<unsigned char>();     // follows the gathered queries
            rc = tc_neighbours_block<X>(st, retry, sc.fb_queries.as<X>(), count, sc.fb_d.as<X>(),
                                        sc.fb_i.as<int64_t>(), stream, fallback, launches, err, direct, 1);
            st.retry_queries += count;
        } else {
            rc = direct(sc.fb_queries.ptr, count, sc.fb_d.ptr, sc.fb_i.as<int64_t>());
            *fallback += count;
        }
        if (rc) return rc;
        tc_scatter_lists_kernel<X><<<count, 64, 0, stream>>>(sc.fb_d.as<X>(), sc.fb_i.as<int64_t>(),
                                                             sc.fb_list.as<int>(), count, ksel, out_d, out_i);
        ++*launches;
        TC_TRY(cudaGetLastError());
    }
    return 0;
}

inline int tc_begin_unproven_list(TensorState& st, int level, int64_t q, cudaStream_t stream, std::string* err) {
    TensorState::Scratch& sc = st.scratch[level];
    TC_TRY(sc.fb_list.reserve((size_t)q * sizeof(int)));
    TC_TRY(sc.fb_count.reserve(sizeof(int)));
    TC_TRY(cudaMemsetAsync(sc.fb_count.ptr, 0, sizeof(int), stream));
    return 0;
}

// one block, start to finish
template <typename X>
inline int tc_neighbours_block(TensorState& st, const TensorLaunch& tl, const X* d_queries, int64_t q,
                               X* out_d, int64_t* out_i, cudaStream_t stream, int64_t* fallback,
                               int64_t* launches, std::string* err, const DirectFallback& direct, int level) {
    int rc = tc_begin_unproven_list(st, level, q, stream, err);
    if (rc) return rc;
    rc = tc_filter_block<X>(st, tl, d_queries, q, out_d, out_i, stream, launches, err, level, 0);
    if (rc || (OSML_TC_DEBUG && tl.debug)) return rc;
    return tc_resolve_unproven<X>(st, tl, d_queries, q, out_d, out_i, stream, fallback, launches, err, direct, level);
}

template <typename X>
inline int tensor_neighbours_typed(TensorState& st, const TensorLaunch& tl, const void* d_queries, int64_t q, void* out_d,
                                   int64_t* out_i, cudaStream_t stream, int64_t* fallback, int64_t* launches,
                                   std::string* err, const DirectFallback& direct) {
    int rc = tc_prepare<X>(st, tl, stream, launches, err);
    if (rc) return rc;
    st.filter_ms = 0.0;
    st.filter_launches = 0;
    const bool blocks_given = tl.n_blocks > 1 && tl.block_ends != nullptr && tl.before_block && *tl.before_block;
    if (!st.usable) {                       // non-finite train rows: the filter's bounds do not hold
        for (int b = 1; blocks_given && b < tl.n_blocks; ++b) {
            rc = (*tl.before_block)(b);
            if (rc) return rc;
        }
        *fallback += q;
        return direct(d_queries, q, out_d, out_i);
    }
    // Blocks of queries are queued back to back and the host waits once, at the end, for the count of unproven queries
    // of the whole call.  The blocks are the caller's (a pipelined osml_knn_predict: see TensorLaunch::before_block), cut
    // further where one exceeds 131 072 queries.
    const int64_t block = 131072;
    const int ksel = tl.k + 1;
    const X* dq = static_cast<const X*>(d_queries);
    X* od = static_cast<X*>(out_d);
    rc = tc_begin_unproven_list(st, 0, q, stream, err);
    if (rc) return rc;
    st.ev_k1_used = 0;
    const int nb = blocks_given ? tl.n_blocks : 1;
    int64_t q0 = 0;
    for (int b = 0; b < nb; ++b) {
        if (b > 0) {
            rc = (*tl.before_block)(b);
            if (rc) return rc;
        }
        const int64_t q1 = blocks_given ? std::min(q, tl.block_ends[b]) : q;
        for (; q0 < q1; q0 += block) {
            const int64_t qb = std::min(block, q1 - q0);
            rc = tc_filter_block<X>(st, tl, dq + q0 * tl.d, qb, od + q0 * ksel, out_i + q0 * ksel, stream, launches, err, 0, q0);
            if (rc) return rc;
        }
        q0 = q1;
    }
    if (OSML_TC_DEBUG && tl.debug) return 0;
    return tc_resolve_unproven<X>(st, tl, dq, q, od, out_i, stream, fallback, launches, err, direct, 0);
}

// float32 or float64 train / query matrices (tl.dtype); distances come out in the same type
inline int tensor_neighbours(TensorState& st, const TensorLaunch& tl, const void* d_queries, int64_t q, void* out_d,
                             int64_t* out_i, cudaStream_t stream, int64_t* fallback, int64_t* launches,
                             std::string* err, const DirectFallback& direct) {
    if (tl.dtype == 1 /* OSML_KNN_F64 */)
        return tensor_neighbours_typed<double>(st, tl, d_queries, q, out_d, out_i, stream, fallback, launches, err, direct);
    return tensor_neighbours_typed<float>(st, tl, d_queries, q, out_d, out_i, stream, fallback, launches, err, direct);
}

}  // namespace osml
