// Column transform in front of kNN: the reference's Centering (osml/transform.py:53-66) and Standardization
// (osml/transform.py:69-84), y[i][j] = (x[i][j] - shift[j]) / denom[j].
//
// HBM-bound streaming kernel: reads n*d elements, writes n*d elements, per-column vectors stay in L1.  The arithmetic
// is numpy's: float32 data stays float32, every other dtype is promoted to float64 first (numpy subtracts a float64
// mean from integer data); one rounding for the subtraction, one for the division (IEEE, never contracted).
// Standalone (host matrix in, host matrix out: Standardization.transform) and fused into osml_knn_predict (the queries
// are uploaded raw and transformed in place on the device, which removes the host pass over Q x D).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace osml {

template <typename TOut> struct TransformOps;
template <> struct TransformOps<float> {
    static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
    static __device__ __forceinline__ float div(float a, float b) { return __fdiv_rn(a, b); }
};
template <> struct TransformOps<double> {
    static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
    static __device__ __forceinline__ double div(double a, double b) { return __ddiv_rn(a, b); }
};

// Flat element i of an n x d matrix: column = i / n when column-major, i % d when row-major.  in == out is allowed
// when TIn == TOut.  Grid-stride; consecutive threads touch consecutive elements in either layout.
template <typename TIn, typename TOut>
__global__ void __launch_bounds__(256)
column_transform_kernel(const TIn* __restrict__ in, TOut* __restrict__ out, int64_t n, int64_t d, int col_major,
                        const TOut* __restrict__ shift, const TOut* __restrict__ denom) {
    using Ops = TransformOps<TOut>;
    const int64_t total = n * d;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int64_t j = col_major ? i / n : i % d;
        TOut v = Ops::sub((TOut)in[i], __ldg(shift + j));
        if (denom) v = Ops::div(v, __ldg(denom + j));
        out[i] = v;
    }
}

template <typename TIn, typename TOut>
inline cudaError_t launch_column_transform(const void* in, void* out, int64_t n, int64_t d, int col_major,
                                           const void* shift, const void* denom, cudaStream_t stream) {
    const int64_t total = n * d;
    if (total <= 0) return cudaSuccess;
    // 8 elements per thread keeps enough loads in flight per SM; capped at a few waves of 148 SMs x 8 CTAs
    const int64_t want = (total + 256 * 8 - 1) / (256 * 8);
    const unsigned grid = (unsigned)std::max<int64_t>(1, std::min<int64_t>(want, 148 * 16));
    column_transform_kernel<TIn, TOut><<<grid, 256, 0, stream>>>(static_cast<const TIn*>(in), static_cast<TOut*>(out), n, d,
                                                                 col_major, static_cast<const TOut*>(shift),
                                                                 static_cast<const TOut*>(denom));
    return cudaGetLastError();
}

}  // namespace osml
