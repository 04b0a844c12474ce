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

// Column-major (the layout of DataFrame.values): blockIdx.y is the column, so shift / denom are per-block scalars and
// no index arithmetic is needed; each thread streams kTransformPerThread coalesced elements of its column chunk.
constexpr int kTransformPerThread = 8;

template <typename TIn, typename TOut>
__global__ void __launch_bounds__(256)
column_transform_colmajor_kernel(const TIn* __restrict__ in, TOut* __restrict__ out, int64_t n,
                                 const TOut* __restrict__ shift, const TOut* __restrict__ denom) {
    using Ops = TransformOps<TOut>;
    const int64_t col = blockIdx.y;
    const TOut sh = __ldg(shift + col);
    const TOut dn = denom ? __ldg(denom + col) : (TOut)1;
    const TIn* __restrict__ src = in + col * n;
    TOut* __restrict__ dst = out + col * n;
    const int64_t r0 = (int64_t)blockIdx.x * (256 * kTransformPerThread) + threadIdx.x;
    TIn v[kTransformPerThread];
#pragma unroll
    for (int u = 0; u < kTransformPerThread; ++u) {
        const int64_t r = r0 + (int64_t)u * 256;
        if (r < n) v[u] = src[r];
    }
#pragma unroll
    for (int u = 0; u < kTransformPerThread; ++u) {
        const int64_t r = r0 + (int64_t)u * 256;
        if (r < n) {
            TOut y = Ops::sub((TOut)v[u], sh);
            if (denom) y = Ops::div(y, dn);
            dst[r] = y;
        }
    }
}

// Row-major (device-resident query matrices, C-ordered host arrays): flat index i, column = i % d kept incrementally.
template <typename TIn, typename TOut>
__global__ void __launch_bounds__(256)
column_transform_rowmajor_kernel(const TIn* __restrict__ in, TOut* __restrict__ out, int64_t total, int d,
                                 const TOut* __restrict__ shift, const TOut* __restrict__ denom) {
    using Ops = TransformOps<TOut>;
    const int64_t i0 = (int64_t)blockIdx.x * (256 * kTransformPerThread) + threadIdx.x;
    int j = (int)(i0 % d);
    const int step = 256 % d;
    TIn v[kTransformPerThread];
#pragma unroll
    for (int u = 0; u < kTransformPerThread; ++u) {
        const int64_t i = i0 + (int64_t)u * 256;
        if (i < total) v[u] = in[i];
    }
#pragma unroll
    for (int u = 0; u < kTransformPerThread; ++u) {
        const int64_t i = i0 + (int64_t)u * 256;
        if (i < total) {
            TOut y = Ops::sub((TOut)v[u], __ldg(shift + j));
            if (denom) y = Ops::div(y, __ldg(denom + j));
            out[i] = y;
        }
        j += step;
        if (j >= d) j -= d;
    }
}

// in == out is allowed when TIn == TOut (every element is read before it is written, by the same thread)
template <typename TIn, typename TOut>
inline cudaError_t launch_column_transform(const void* in, void* out, int64_t n, int64_t d, int col_major,
                                           const void* shift, const void* denom, cudaStream_t stream) {
    const int64_t total = n * d;
    if (total <= 0) return cudaSuccess;
    const int64_t per_block = 256 * kTransformPerThread;
    if (col_major && d <= 65535) {
        dim3 grid((unsigned)((n + per_block - 1) / per_block), (unsigned)d);
        column_transform_colmajor_kernel<TIn, TOut><<<grid, 256, 0, stream>>>(
            static_cast<const TIn*>(in), static_cast<TOut*>(out), n, static_cast<const TOut*>(shift),
            static_cast<const TOut*>(denom));
    } else if (!col_major) {
        column_transform_rowmajor_kernel<TIn, TOut><<<(unsigned)((total + per_block - 1) / per_block), 256, 0, stream>>>(
            static_cast<const TIn*>(in), static_cast<TOut*>(out), total, (int)d, static_cast<const TOut*>(shift),
            static_cast<const TOut*>(denom));
    } else {
        return cudaErrorInvalidValue;      // more than 65535 columns in a column-major matrix: not a kNN frame
    }
    return cudaGetLastError();
}

}  // namespace osml
