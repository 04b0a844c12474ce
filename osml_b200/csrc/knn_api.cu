// C ABI of the kNN predict path (include/osml_knn.h): index lifetime, host <-> device staging, kernel dispatch.
// No torch types, no CPU fallback: every arithmetic step of the path runs in the CUDA kernels of this directory.
#include "../../include/osml_knn.h"

#include <cuda_runtime.h>
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <functional>
#include <new>
#include <string>
#include <vector>

#include "knn_aggregate.cuh"
#include "knn_common.cuh"
#include "knn_direct.cuh"
#include "knn_tensor.cuh"
#include "knn_transform.cuh"
#include "host_stager.cuh"

namespace {

thread_local std::string g_last_error;

int fail(int status, const std::string& msg) {
    g_last_error = msg;
    return status;
}

#define CUDA_TRY(expr)                                                                                      \
    do {                                                                                                    \
        cudaError_t err__ = (expr);                                                                         \
        if (err__ != cudaSuccess) {                                                                         \
            cudaGetLastError();                                                                             \
            return fail(err__ == cudaErrorMemoryAllocation ? OSML_KNN_ENOMEM : OSML_KNN_ECUDA,              \
                        std::string(#expr) + ": " + cudaGetErrorString(err__));                            \
        }                                                                                                   \
    } while (0)

size_t dtype_size(int dtype) {
    switch (dtype) {
        case OSML_KNN_F32: case OSML_KNN_I32: case OSML_KNN_U32: return 4;
        case OSML_KNN_F64: case OSML_KNN_I64: case OSML_KNN_U64: return 8;
        case OSML_KNN_I8: case OSML_KNN_U8: return 1;
        case OSML_KNN_I16: case OSML_KNN_U16: return 2;
        default: return 0;
    }
}

struct DeviceBuffer {
    void* ptr = nullptr;
    size_t cap = 0;
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

// tiled transpose of a column-major (n x d, element (i, j) at j * n + i) matrix into row-major, by element size
template <typename W>
__global__ void transpose_colmajor_kernel(const W* __restrict__ in, W* __restrict__ out, int64_t n, int64_t d) {
    __shared__ W tile[32][33];
    const int64_t i0 = (int64_t)blockIdx.x * 32, j0 = (int64_t)blockIdx.y * 32;
    for (int jj = threadIdx.y; jj < 32; jj += blockDim.y) {
        const int64_t i = i0 + threadIdx.x, j = j0 + jj;
        if (i < n && j < d) tile[jj][threadIdx.x] = in[j * n + i];
    }
    __syncthreads();
    for (int ii = threadIdx.y; ii < 32; ii += blockDim.y) {
        const int64_t i = i0 + ii, j = j0 + threadIdx.x;
        if (i < n && j < d) out[i * d + j] = tile[threadIdx.x][ii];
    }
}

// ---- integer matrices on the tensor filter ----------------------------------------------------------------------------
// numpy evaluates sqrt(sum(square(t - q))) of an integer matrix in the matrix's own type: subtract and square wrap, the
// sum widens to (u)int64 (osml/predict.py:742).  While no subtraction, square or partial sum can wrap
// -- |t - q| <= L for every pair of values, L from the type and the feature count -- every intermediate is a small
// integer, and float64 arithmetic reproduces it EXACTLY (integers below 2^53, whatever the order of the adds): the
// float64 tensor path (K0 + K1 + K3) then returns the integer oracle's bits, sqrt(float64(sum)).  So an integer index
// keeps a float64 shadow of its rows for that path; queries are converted per call, and a query that takes some
// |t - q| beyond L is flagged for the exact direct kernel (which reproduces the wrap-around).
template <typename T>
__global__ void int_minmax_kernel(const T* __restrict__ x, int64_t total, double* __restrict__ lo_hi) {
    double lo = INFINITY, hi = -INFINITY;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const double v = (double)x[i];
        lo = fmin(lo, v); hi = fmax(hi, v);
    }
    for (int off = 16; off > 0; off >>= 1) {
        lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, off));
        hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, off));
    }
    if ((threadIdx.x & 31) == 0) {          // doubles of either sign: order-preserving map to unsigned for atomicMin/Max
        auto key = [](double v) { unsigned long long u = (unsigned long long)__double_as_longlong(v);
                                  return (u >> 63) ? ~u : (u | 0x8000000000000000ull); };
        atomicMin(reinterpret_cast<unsigned long long*>(lo_hi), key(lo));
        atomicMax(reinterpret_cast<unsigned long long*>(lo_hi) + 1, key(hi));
    }
}

template <typename T>
__global__ void int_to_f64_kernel(const T* __restrict__ x, int64_t total, double* __restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = (double)x[i];
}

// one warp per query row: float64 copy + "outside the no-wrap range" flag
template <typename T>
__global__ void int_queries_to_f64_kernel(const T* __restrict__ x, int64_t q, int d, double train_lo, double train_hi,
                                          double limit, double* __restrict__ out, unsigned char* __restrict__ force) {
    const int lane = threadIdx.x & 31;
    const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= q) return;
    double lo = train_lo, hi = train_hi;
    for (int j = lane; j < d; j += 32) {
        const double v = (double)x[row * d + j];
        out[row * d + j] = v;
        lo = fmin(lo, v); hi = fmax(hi, v);
    }
    for (int off = 16; off > 0; off >>= 1) {
        lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, off));
        hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, off));
    }
    if (lane == 0) force[row] = (hi - lo <= limit) ? 0 : 1;
}

template <typename T>
__global__ void f64_to_int_kernel(const double* __restrict__ x, int64_t total, T* __restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = (T)x[i];                    // integer-valued by construction (int_queries_to_f64_kernel)
}

}  // namespace

struct osml_knn_index {
    int device = 0;
    int dtype = 0;
    int64_t n = 0, d = 0;
    int sum_order = 1;
    int k = 0;
    double gamma = 1.0;
    int gamma_is_f32 = 0;
    double epsilon = 1e-8;
    int64_t index_base = 0;
    int64_t n_labels = 0;
    int label_kind = 0;
    int label_is_f32 = 0;      // float32 regression labels: the reference's weighted sum then runs in float32 (K4)
    int path = OSML_KNN_PATH_AUTO;
    int slack = -1;
    int opt_epi_warps = 0;
    int opt_nseg = 0;
    int opt_smem_lists = 0;
    int opt_cta_group = 0;
    int opt_active = 0;
    int opt_precision = 0;
    int opt_wide_pair = 1;
    int opt_hot = 0;
    int opt_live = 1;
    int opt_kprime = 0;
    int debug = 0;
    int counters = 0;
    int64_t query_block = 0;
    int opt_pipeline = 1;      // osml_knn_predict: 1 = upload the tail of a large query matrix under the first block's K1
    int opt_stager = 1;        // pageable host matrices of 2 MB and more go up through the pinned staging ring (host_stager.cuh)
    int opt_pipeline_blocks = 0;   // equal blocks of a call below two waves: 0 = two (more: experiments)
    int sm_count = 0;
    const std::function<int(int)>* before_block = nullptr;   // both set only while a pipelined osml_knn_predict is in flight
    std::vector<int64_t> block_ends;
    int64_t seen_queries = 0;  // since create: auto builds the filter once enough queries have been asked for
    bool auto_direct = false;  // auto: the filter could not prove most queries of a call, stay on the direct kernel

    DeviceBuffer train;        // row-major n x d, original dtype
    DeviceBuffer label_bits;   // int64 per label
    DeviceBuffer stage;        // raw upload staging (column-major inputs)
    DeviceBuffer queries;      // row-major q x d
    DeviceBuffer part_d, part_i;   // K2 partial lists
    DeviceBuffer nn_d, nn_i;       // final lists
    DeviceBuffer pred;
    DeviceBuffer train_f64;    // integer dtypes: float64 shadow of the rows for the tensor path (built on first use)
    DeviceBuffer q_f64, q_force, q_back;   // ... per call: converted queries, direct-kernel flags, rows converted back for K2
    double int_lo = 0.0, int_hi = 0.0, int_limit = -1.0;   // value range of the train set; largest |t - q| that cannot wrap
    bool int_tensor_ok = false;
    DeviceBuffer q_shift, q_denom;   // fused query transform (osml_knn_set_query_transform), index dtype
    bool has_q_transform = false, has_q_denom = false;
    osml::TensorState tensor;  // K0/K1/K3 state (bf16 copy, norms, tensor maps, scratch)

    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;      // second block's upload (pipelined osml_knn_predict)
    cudaEvent_t ev_copy = nullptr, ev_first = nullptr;   // second block uploaded / first block uploaded
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr, ev3 = nullptr;
    osml_knn_stats stats{};

    size_t dist_size() const { return dtype == OSML_KNN_F32 ? 4 : 8; }
    int ksel() const { return k + 1; }
};

namespace {

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

// Pageable host memory goes through the pinned staging ring (host_stager.cuh) from 2 MB up: chunks of about a megabyte,
// gathered by a few threads, one cudaMemcpyAsync each.  `dev` receives `bytes` contiguous bytes; piece `i` of them comes
// from host address src(i_offset) -- contiguous for row-major input, one column segment after the other otherwise.
int staged_copy_rowmajor(int device, const char* host, char* dev, size_t bytes, cudaStream_t stream) {
    std::vector<osml::HostStager::Chunk> chunks;
    for (size_t off = 0; off < bytes; off += osml::HostStager::kSlotBytes) {
        const size_t nb = std::min(osml::HostStager::kSlotBytes, bytes - off);
        const char* src = host + off;
        chunks.push_back({dev + off, nb, [src, nb](char* slot) { memcpy(slot, src, nb); }});
    }
    CUDA_TRY(osml::HostStager::instance().copy(device, chunks, stream));
    return OSML_KNN_OK;
}

// column-major host matrix, rows [r0, r1) -> device staging region [cols][nr] (element (i, j) at j * nr + i)
int staged_copy_colmajor(osml_knn_index* ix, const char* host, int64_t rows, int64_t cols, int64_t r0, int64_t r1,
                         size_t es, char* dev, cudaStream_t stream) {
    const int64_t nr = r1 - r0;
    const size_t col_bytes = (size_t)nr * es, slot = osml::HostStager::kSlotBytes;
    std::vector<osml::HostStager::Chunk> chunks;
    if (col_bytes <= slot) {                         // whole column segments, as many as fit a slot
        const int64_t per = std::max<int64_t>(1, (int64_t)(slot / col_bytes));
        for (int64_t j0 = 0; j0 < cols; j0 += per) {
            const int64_t j1 = std::min(cols, j0 + per);
            const char* src = host + ((size_t)j0 * rows + r0) * es;
            const size_t stride = (size_t)rows * es;
            chunks.push_back({dev + (size_t)j0 * col_bytes, (size_t)(j1 - j0) * col_bytes,
                              [src, stride, col_bytes, j0, j1](char* out) {
                                  for (int64_t j = j0; j < j1; ++j)
                                      memcpy(out + (size_t)(j - j0) * col_bytes, src + (size_t)(j - j0) * stride, col_bytes);
                              }});
        }
    } else {                                         // long columns: pieces of one column
        for (int64_t j = 0; j < cols; ++j)
            for (size_t off = 0; off < col_bytes; off += slot) {
                const size_t nb = std::min(slot, col_bytes - off);
                const char* src = host + ((size_t)j * rows + r0) * es + off;
                chunks.push_back({dev + (size_t)j * col_bytes + off, nb, [src, nb](char* out) { memcpy(out, src, nb); }});
            }
    }
    CUDA_TRY(osml::HostStager::instance().copy(ix->device, chunks, stream));
    return OSML_KNN_OK;
}

// Rows [r0, r1) of a host matrix of `rows` rows into the same rows of the row-major device matrix `dst` (already
// reserved for all rows).  Column-major input goes through its own region of the staging buffer (already reserved) as a
// pitched copy and is transposed on the same stream, so blocks on different streams do not touch each other's memory.
int upload_rows(osml_knn_index* ix, const void* host, int layout, int64_t rows, int64_t cols, int64_t r0, int64_t r1,
                DeviceBuffer& dst, cudaStream_t stream) {
    const size_t es = dtype_size(ix->dtype);
    const int64_t nr = r1 - r0;
    char* out = dst.as<char>() + (size_t)r0 * cols * es;
    const size_t bytes = (size_t)nr * cols * es;
    const bool staged = ix->opt_stager && bytes >= ((size_t)2 << 20) && osml::is_pageable_host(host);
    if (layout == OSML_KNN_ROW_MAJOR || rows == 1 || cols == 1) {
        const char* src = static_cast<const char*>(host) + (size_t)r0 * cols * es;
        if (staged) return staged_copy_rowmajor(ix->device, src, out, bytes, stream);
        CUDA_TRY(cudaMemcpyAsync(out, src, bytes, cudaMemcpyHostToDevice, stream));
        return OSML_KNN_OK;
    }
    char* st = ix->stage.as<char>() + (size_t)r0 * cols * es;      // [cols][nr], element (i, j) at j * nr + i
    if (staged) {
        int rc = staged_copy_colmajor(ix, static_cast<const char*>(host), rows, cols, r0, r1, es, st, stream);
        if (rc != OSML_KNN_OK) return rc;
    } else {
        CUDA_TRY(cudaMemcpy2DAsync(st, (size_t)nr * es, static_cast<const char*>(host) + (size_t)r0 * es, (size_t)rows * es,
                                   (size_t)nr * es, (size_t)cols, cudaMemcpyHostToDevice, stream));
    }
    dim3 block(32, 8), grid((unsigned)((nr + 31) / 32), (unsigned)((cols + 31) / 32));
    switch (es) {
        case 1: transpose_colmajor_kernel<uint8_t><<<grid, block, 0, stream>>>((const uint8_t*)st, (uint8_t*)out, nr, cols); break;
        case 2: transpose_colmajor_kernel<uint16_t><<<grid, block, 0, stream>>>((const uint16_t*)st, (uint16_t*)out, nr, cols); break;
        case 4: transpose_colmajor_kernel<uint32_t><<<grid, block, 0, stream>>>((const uint32_t*)st, (uint32_t*)out, nr, cols); break;
        default: transpose_colmajor_kernel<uint64_t><<<grid, block, 0, stream>>>((const uint64_t*)st, (uint64_t*)out, nr, cols); break;
    }
    ix->stats.kernel_launches++;
    CUDA_TRY(cudaGetLastError());
    return OSML_KNN_OK;
}

// whole matrix (the train set at create)
int upload_matrix(osml_knn_index* ix, const void* host, int layout, int64_t rows, int64_t cols, DeviceBuffer& dst,
                  cudaStream_t stream) {
    const size_t bytes = (size_t)rows * cols * dtype_size(ix->dtype);
    CUDA_TRY(dst.reserve(bytes));
    if (layout != OSML_KNN_ROW_MAJOR && rows > 1 && cols > 1) CUDA_TRY(ix->stage.reserve(bytes));
    return upload_rows(ix, host, layout, rows, cols, 0, rows, dst, stream);
}

template <typename T>
int labels_to_bits(const void* labels, int64_t n, std::vector<int64_t>& out) {
    const T* p = static_cast<const T*>(labels);
    out.resize((size_t)n);
    for (int64_t i = 0; i < n; ++i) out[(size_t)i] = (int64_t)p[i];
    return 0;
}

int convert_labels(const void* labels, int label_dtype, int64_t n, std::vector<int64_t>& out, int* kind) {
    switch (label_dtype) {
        case OSML_KNN_I8: *kind = osml::kLabelSigned; return labels_to_bits<int8_t>(labels, n, out);
        case OSML_KNN_I16: *kind = osml::kLabelSigned; return labels_to_bits<int16_t>(labels, n, out);
        case OSML_KNN_I32: *kind = osml::kLabelSigned; return labels_to_bits<int32_t>(labels, n, out);
        case OSML_KNN_I64: *kind = osml::kLabelSigned; return labels_to_bits<int64_t>(labels, n, out);
        case OSML_KNN_U8: *kind = osml::kLabelUnsigned; return labels_to_bits<uint8_t>(labels, n, out);
        case OSML_KNN_U16: *kind = osml::kLabelUnsigned; return labels_to_bits<uint16_t>(labels, n, out);
        case OSML_KNN_U32: *kind = osml::kLabelUnsigned; return labels_to_bits<uint32_t>(labels, n, out);
        case OSML_KNN_U64: *kind = osml::kLabelUnsigned; return labels_to_bits<uint64_t>(labels, n, out);
        case OSML_KNN_F32: {
            *kind = osml::kLabelFloat;
            const float* p = static_cast<const float*>(labels);
            out.resize((size_t)n);
            for (int64_t i = 0; i < n; ++i) { double v = (double)p[i]; std::memcpy(&out[(size_t)i], &v, 8); }
            return 0;
        }
        case OSML_KNN_F64: {
            *kind = osml::kLabelFloat;
            out.resize((size_t)n);
            std::memcpy(out.data(), labels, (size_t)n * 8);
            return 0;
        }
        default: return 1;
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// K2 launch
// ---------------------------------------------------------------------------------------------------------------------
template <typename T>
int launch_direct_typed(osml_knn_index* ix, const void* d_queries, int64_t q, void* out_d, int64_t* out_i,
                        cudaStream_t stream) {
    using Dist = typename osml::Elem<T>::Dist;
    const int ksel = ix->ksel();
    const int n = (int)ix->n, d = (int)ix->d;
    const bool generic = osml::Elem<T>::kIsFloat && ix->sum_order == OSML_KNN_SUM_PAIRWISE && d > 1;
    // the generic kernel keeps k + 1 (distance, index) entries per thread in shared memory: fewer threads for large k
    int generic_qt = osml::kDirectQT;
    while (generic_qt > 8 && (size_t)ksel * generic_qt * (sizeof(Dist) + sizeof(int)) > 200 * 1024) generic_qt /= 2;
    const int qgroup = generic ? generic_qt : osml::kRowQB;
    const int ctas_q = (int)((q + qgroup - 1) / qgroup);
    // split the train rows until the grid covers the machine a few times over
    const int row_quant = generic ? 64 : osml::kRowThreads;
    int nsplit = 1;
    const int target_ctas = generic ? 296 : 148 * 8;
    if (ctas_q < target_ctas) {
        nsplit = (target_ctas + ctas_q - 1) / ctas_q;
        const int max_split = std::max(1, n / (2 * row_quant));
        nsplit = std::max(1, std::min(std::min(nsplit, 64), max_split));
    }
    int rows_per_split = (n + nsplit - 1) / nsplit;
    rows_per_split = ((rows_per_split + row_quant - 1) / row_quant) * row_quant;
    nsplit = (n + rows_per_split - 1) / rows_per_split;

    Dist* part_d = static_cast<Dist*>(out_d);
    int64_t* part_i = out_i;
    if (nsplit > 1) {
        CUDA_TRY(ix->part_d.reserve((size_t)q * nsplit * ksel * sizeof(Dist)));
        CUDA_TRY(ix->part_i.reserve((size_t)q * nsplit * ksel * sizeof(int64_t)));
        part_d = ix->part_d.as<Dist>();
        part_i = ix->part_i.as<int64_t>();
    }
    dim3 grid((unsigned)ctas_q, (unsigned)nsplit);
    if (generic) {
        const size_t smem = (size_t)ksel * generic_qt * (sizeof(Dist) + sizeof(int));
        auto kern = osml::knn_direct_generic_kernel<T>;
        CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, generic_qt, smem, stream>>>(ix->train.as<T>(), static_cast<const T*>(d_queries), n, d, (int)q,
                                                      ksel, rows_per_split, ix->index_base, ix->sum_order, part_d,
                                                      part_i);
    } else {
        // rows too wide for the shared-memory query tile go through it in chunks of d_chunk features (a multiple of 16
        // elements, so 16-byte row loads stay aligned); the reference has no width limit (predict.py:742)
        const size_t budget = 200 * 1024;
        int d_chunk = d;
        if (osml::RowSmem<T>::bytes(d, ksel) > budget) {
            const size_t fixed = osml::RowSmem<T>::bytes(0, ksel);
            if (fixed + 16 * osml::kRowQB * sizeof(T) > budget)
                return fail(OSML_KNN_EUNSUPPORTED, "k too large for the direct kernel's shared-memory lists");
            d_chunk = (int)((budget - fixed) / (osml::kRowQB * sizeof(T))) / 16 * 16;
        }
        const size_t smem = osml::RowSmem<T>::bytes(d_chunk, ksel);
        // 16-byte row loads need 16-byte aligned rows (the base comes from cudaMalloc)
        const bool vec = ((size_t)d * sizeof(T)) % 16 == 0;
        auto kern = vec ? osml::knn_direct_rows_kernel<T, true> : osml::knn_direct_rows_kernel<T, false>;
        CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, osml::kRowThreads, smem, stream>>>(ix->train.as<T>(), static_cast<const T*>(d_queries), n, d, d_chunk,
                                                        (int)q, ksel, rows_per_split, ix->index_base, part_d, part_i);
    }
    ix->stats.kernel_launches++;
    ix->stats.last_filter_launches++;
    CUDA_TRY(cudaGetLastError());
    if (nsplit > 1) {
        const int threads = 128;
        osml::knn_merge_lists_kernel<Dist><<<(unsigned)((q + threads - 1) / threads), threads, 0, stream>>>(
            part_d, part_i, nsplit, (int64_t)ksel, (int64_t)nsplit * ksel, q, ksel, static_cast<Dist*>(out_d), out_i);
        ix->stats.kernel_launches++;
        CUDA_TRY(cudaGetLastError());
    }
    return OSML_KNN_OK;
}

int launch_direct(osml_knn_index* ix, const void* d_queries, int64_t q, void* out_d, int64_t* out_i,
                  cudaStream_t stream) {
    switch (ix->dtype) {
        case OSML_KNN_F32: return launch_direct_typed<float>(ix, d_queries, q, out_d, out_i, stream);
        case OSML_KNN_F64: return launch_direct_typed<double>(ix, d_queries, q, out_d, out_i, stream);
        case OSML_KNN_I8: return launch_direct_typed<int8_t>(ix, d_queries, q, out_d, out_i, stream);
        case OSML_KNN_I16: return launch_direct_typed<int16_t>(ix, d_queries, q, out_d, out_i, stream);
        case OSML_KNN_I32: return launch_direct_typed<int32_t>(ix, d_queries, q, out_d, out_i, stream);
        case OSML_KNN_I64: return launch_direct_typed<int64_t>(ix, d_queries, q, out_d, out_i, stream);
        case OSML_KNN_U8: return launch_direct_typed<uint8_t>(ix, d_queries, q, out_d, out_i, stream);
        case OSML_KNN_U16: return launch_direct_typed<uint16_t>(ix, d_queries, q, out_d, out_i, stream);
        case OSML_KNN_U32: return launch_direct_typed<uint32_t>(ix, d_queries, q, out_d, out_i, stream);
        case OSML_KNN_U64: return launch_direct_typed<uint64_t>(ix, d_queries, q, out_d, out_i, stream);
        default: return fail(OSML_KNN_EUNSUPPORTED, "unsupported dtype");
    }
}

bool is_int_dtype(int dtype) { return dtype != OSML_KNN_F32 && dtype != OSML_KNN_F64; }

// dtype the tensor path computes in: the matrix's own for floats, float64 (an exact shadow) for integers
constexpr int64_t kAutoTensorRows = 4096;      // path = auto: smaller training sets stay on the direct kernel

int tensor_dtype(const osml_knn_index* ix) { return is_int_dtype(ix->dtype) ? (int)OSML_KNN_F64 : ix->dtype; }

bool use_tensor_path(const osml_knn_index* ix, int64_t q) {
    if (ix->path == OSML_KNN_PATH_DIRECT) return false;
    if (is_int_dtype(ix->dtype) && !ix->int_tensor_ok) return false;
    if (!osml::tensor_path_supported(tensor_dtype(ix), ix->n, ix->d, ix->k)) return false;
    if (ix->path == OSML_KNN_PATH_TENSOR) return true;
    // auto.  Measured on a fitted index (tools/crossover.py, profiles/r1h_crossover.jsonl): once K0 has run, the filter
    // beats the direct kernel for every call on the grid, from 20 000 x 32 rows / 32 queries (0.147 vs 0.159 ms) to
    // 1M x 128 / 2048 queries (0.68 vs 46.9 ms).  K0 itself (two passes over the train set, a sort, the allocations) costs
    // about what the direct kernel needs for 150 queries, so it is built once that many have been asked for.  Data on
    // which the completeness proof keeps failing (massive exact ties) would pay for the filter AND the direct kernel:
    // after such a call auto stays on the direct kernel.
    // (round 2, after the work split for small calls: the filter also wins from 4096 rows on -- 4096 x 128: 0.11 / 0.12 /
    // 0.14 ms for 16 / 1024 / 8192 queries against 0.12 / 0.32 / 1.08 ms; 8192 x 784, 8192 queries: 0.38 vs 12.5 ms; at
    // 2048 rows only from ~1000 queries on)
    if (ix->n < kAutoTensorRows || ix->auto_direct) return false;
    return ix->tensor.ready || q + ix->seen_queries >= 128;
}

// ---- integer matrices: float64 shadow for the tensor path --------------------------------------------------------------
template <typename F>
int dispatch_int(int dtype, F&& f) {
    switch (dtype) {
        case OSML_KNN_I8: return f(int8_t{});
        case OSML_KNN_I16: return f(int16_t{});
        case OSML_KNN_I32: return f(int32_t{});
        case OSML_KNN_I64: return f(int64_t{});
        case OSML_KNN_U8: return f(uint8_t{});
        case OSML_KNN_U16: return f(uint16_t{});
        case OSML_KNN_U32: return f(uint32_t{});
        case OSML_KNN_U64: return f(uint64_t{});
        default: return fail(OSML_KNN_EUNSUPPORTED, "not an integer dtype");
    }
}

// largest |t - q| for which numpy's subtract and square cannot wrap in the element type
double int_type_limit(int dtype) {
    switch (dtype) {
        case OSML_KNN_I8: return 11.0;             // 11^2 = 121 <= 127
        case OSML_KNN_I16: return 181.0;           // 181^2 = 32761 <= 32767
        case OSML_KNN_I32: return 46340.0;         // 46340^2 <= 2^31 - 1
        case OSML_KNN_I64: return 3037000499.0;    // floor(sqrt(2^63 - 1))
        case OSML_KNN_U8: return 15.0;             // (t - q) wraps mod 2^8 when t < q, but its square mod 2^8 is (t - q)^2
        case OSML_KNN_U16: return 255.0;           //   as long as that is below 2^b
        case OSML_KNN_U32: return 65535.0;
        case OSML_KNN_U64: return 4294967295.0;
        default: return -1.0;
    }
}

// value range of an integer train matrix (at create) and whether the tensor path can serve it
int int_scan_train(osml_knn_index* ix) {
    DeviceBuffer lo_hi;
    CUDA_TRY(lo_hi.reserve(16));
    const unsigned long long init[2] = {~0ull, 0ull};
    CUDA_TRY(cudaMemcpyAsync(lo_hi.ptr, init, 16, cudaMemcpyHostToDevice, ix->stream));
    const int64_t total = ix->n * ix->d;
    int rc = dispatch_int(ix->dtype, [&](auto tag) {
        using T = decltype(tag);
        int_minmax_kernel<T><<<148 * 8, 256, 0, ix->stream>>>(ix->train.as<T>(), total, lo_hi.as<double>());
        return (int)OSML_KNN_OK;
    });
    if (rc != OSML_KNN_OK) return rc;
    ix->stats.kernel_launches++;
    unsigned long long keys[2];
    CUDA_TRY(cudaMemcpyAsync(keys, lo_hi.ptr, 16, cudaMemcpyDeviceToHost, ix->stream));
    CUDA_TRY(cudaStreamSynchronize(ix->stream));
    lo_hi.release();
    auto unkey = [](unsigned long long u) { u = (u >> 63) ? (u & 0x7fffffffffffffffull) : ~u; double v; std::memcpy(&v, &u, 8); return v; };
    ix->int_lo = unkey(keys[0]); ix->int_hi = unkey(keys[1]);
    // float64 reproduces the integer arithmetic while every partial sum stays below 2^53: d * L^2 <= 2^53
    const double exact = std::floor(std::sqrt(9007199254740992.0 / (double)ix->d));
    ix->int_limit = std::min(int_type_limit(ix->dtype), exact);
    ix->int_tensor_ok = std::fabs(ix->int_lo) <= 9007199254740992.0 && std::fabs(ix->int_hi) <= 9007199254740992.0 &&
                        ix->int_hi - ix->int_lo <= ix->int_limit;
    return OSML_KNN_OK;
}

int int_ensure_shadow(osml_knn_index* ix, cudaStream_t stream) {
    if (ix->train_f64.ptr) return OSML_KNN_OK;
    const int64_t total = ix->n * ix->d;
    CUDA_TRY(ix->train_f64.reserve((size_t)total * 8));
    int rc = dispatch_int(ix->dtype, [&](auto tag) {
        using T = decltype(tag);
        int_to_f64_kernel<T><<<148 * 8, 256, 0, stream>>>(ix->train.as<T>(), total, ix->train_f64.as<double>());
        return (int)OSML_KNN_OK;
    });
    ix->stats.kernel_launches++;
    if (rc != OSML_KNN_OK) return rc;
    CUDA_TRY(cudaGetLastError());
    return OSML_KNN_OK;
}

// neighbour search into (out_d, out_i): final ascending lists, global indices
int neighbours(osml_knn_index* ix, const void* d_queries, int64_t q, void* out_d, int64_t* out_i,
               cudaStream_t stream) {
    if (q < 1) return fail(OSML_KNN_EINVAL, "no queries");
    if (q > 0x7fffffff / 2) return fail(OSML_KNN_EUNSUPPORTED, "more than 2^30 queries in one call");
    ix->stats.last_filter_launches = 0;
    if (use_tensor_path(ix, q)) {
        osml::TensorLaunch tl;
        tl.train = ix->train.ptr; tl.dtype = ix->dtype; tl.n = ix->n; tl.d = ix->d; tl.k = ix->k;
        const bool ints = is_int_dtype(ix->dtype);
        if (ints) {                       // float64 shadow rows, float64 copy of the queries, per-query range flags
            int rc0 = int_ensure_shadow(ix, stream);
            if (rc0 != OSML_KNN_OK) return rc0;
            CUDA_TRY(ix->q_f64.reserve((size_t)q * ix->d * 8));
            CUDA_TRY(ix->q_force.reserve((size_t)q));
            rc0 = dispatch_int(ix->dtype, [&](auto tag) {
                using T = decltype(tag);
                int_queries_to_f64_kernel<T><<<(unsigned)((q * 32 + 255) / 256), 256, 0, stream>>>(
                    static_cast<const T*>(d_queries), q, (int)ix->d, ix->int_lo, ix->int_hi, ix->int_limit,
                    ix->q_f64.as<double>(), ix->q_force.as<unsigned char>());
                return (int)OSML_KNN_OK;
            });
            if (rc0 != OSML_KNN_OK) return rc0;
            ix->stats.kernel_launches++;
            CUDA_TRY(cudaGetLastError());
            tl.train = ix->train_f64.ptr; tl.dtype = OSML_KNN_F64; tl.q_force = ix->q_force.as<unsigned char>();
            d_queries = ix->q_f64.ptr;
        }
        tl.sum_order = ix->sum_order; tl.index_base = ix->index_base; tl.slack = ix->slack;
        tl.opt_epi_warps = ix->opt_epi_warps; tl.opt_nseg = ix->opt_nseg; tl.debug = ix->debug; tl.counters = ix->counters;
        tl.opt_smem_lists = ix->opt_smem_lists; tl.opt_cta_group = ix->opt_cta_group; tl.opt_active = ix->opt_active; tl.opt_precision = ix->opt_precision; tl.opt_wide_pair = ix->opt_wide_pair; tl.opt_kprime = ix->opt_kprime; tl.opt_hot = ix->opt_hot; tl.opt_live = ix->opt_live;
        tl.before_block = ix->before_block; tl.block_ends = ix->block_ends.data(); tl.n_blocks = (int)ix->block_ends.size();
        int64_t fallback = 0, launches = 0;
        std::string err;
        int rc = osml::tensor_neighbours(ix->tensor, tl, d_queries, q, out_d, out_i, stream, &fallback, &launches, &err,
                                         [&](const void* dq, int64_t qq, void* od, int64_t* oi) {
                                             if (!ints) return launch_direct(ix, dq, qq, od, oi, stream);
                                             // the exact kernel works on the matrix's own type: integer-valued
                                             // float64 rows back to it
                                             const int64_t total = qq * ix->d;
                                             CUDA_TRY(ix->q_back.reserve((size_t)total * dtype_size(ix->dtype)));
                                             int rcb = dispatch_int(ix->dtype, [&](auto tag) {
                                                 using T = decltype(tag);
                                                 f64_to_int_kernel<T><<<(unsigned)std::min<int64_t>((total + 255) / 256, 148 * 16), 256, 0, stream>>>(
                                                     static_cast<const double*>(dq), total, ix->q_back.as<T>());
                                                 return (int)OSML_KNN_OK;
                                             });
                                             if (rcb != OSML_KNN_OK) return rcb;
                                             ix->stats.kernel_launches++;
                                             return launch_direct(ix, ix->q_back.ptr, qq, od, oi, stream);
                                         });
        ix->stats.kernel_launches += launches;
        if (rc != OSML_KNN_OK) return err.empty() ? rc : fail(rc, err);
        ix->stats.fallback_queries += fallback;
        ix->seen_queries += q;
        if (q >= 64 && fallback * 2 > q) ix->auto_direct = true;
        ix->stats.retry_queries += ix->tensor.retry_queries;
        ix->tensor.retry_queries = 0;
        ix->stats.tensor_queries += q;
        ix->stats.last_filter_ms = ix->tensor.filter_ms;
        ix->stats.last_filter_launches = ix->tensor.filter_launches;
        return OSML_KNN_OK;
    }
    ix->seen_queries += q;
    return launch_direct(ix, d_queries, q, out_d, out_i, stream);
}

int aggregate(osml_knn_index* ix, const void* nn_d, const int64_t* nn_i, int64_t q, int task, double* d_out_pred,
              cudaStream_t stream) {
    if (task != OSML_KNN_REGRESSION && task != OSML_KNN_CLASSIFICATION) return fail(OSML_KNN_EINVAL, "bad task");
    const int warps = osml::agg_warps_per_cta(ix->k);
    const unsigned grid = (unsigned)((q + warps - 1) / warps);
    const int threads = 32 * warps;
    const size_t smem = osml::agg_smem_bytes(ix->k, warps);
    if (ix->dist_size() == 4) {
        auto kern = osml::knn_aggregate_kernel<float>;
        if (smem > 48 * 1024) CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, threads, smem, stream>>>(
            static_cast<const float*>(nn_d), nn_i, q, ix->k, ix->label_bits.as<int64_t>(), ix->label_kind,
            ix->label_is_f32, ix->n_labels, ix->gamma, ix->gamma_is_f32, ix->epsilon, task, d_out_pred);
    } else {
        auto kern = osml::knn_aggregate_kernel<double>;
        if (smem > 48 * 1024) CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, threads, smem, stream>>>(
            static_cast<const double*>(nn_d), nn_i, q, ix->k, ix->label_bits.as<int64_t>(), ix->label_kind,
            ix->label_is_f32, ix->n_labels, ix->gamma, ix->gamma_is_f32, ix->epsilon, task, d_out_pred);
    }
    ix->stats.kernel_launches++;
    CUDA_TRY(cudaGetLastError());
    return OSML_KNN_OK;
}

int check_predict_args(const osml_knn_index* ix, int64_t q) {
    if (!ix) return fail(OSML_KNN_EINVAL, "null index");
    if (q < 1) return fail(OSML_KNN_EINVAL, "no queries");
    // np.argpartition(distances, k + 1) needs k + 1 < n; n_labels is the global row count
    if (ix->n_labels < (int64_t)ix->k + 2) return fail(OSML_KNN_EINVAL, "fewer than k + 2 training rows");
    return OSML_KNN_OK;
}

}  // namespace

extern "C" {

int osml_knn_abi_version(void) { return 1004; }

const char* osml_knn_last_error(void) { return g_last_error.c_str(); }

int osml_knn_create(osml_knn_index** out, int device, const void* train, int dtype, int layout, int64_t n, int64_t d,
                    int sum_order, const void* labels, int label_dtype, int64_t n_labels, int64_t index_base, int k,
                    double rbf_gamma, int gamma_is_f32, double epsilon) {
    if (!out) return fail(OSML_KNN_EINVAL, "null out");
    *out = nullptr;
    if (!train || !labels) return fail(OSML_KNN_EINVAL, "null input");
    if (k < 1) return fail(OSML_KNN_EINVAL, "k < 1");                       // predict.py:729-730
    if (!(rbf_gamma > 0)) return fail(OSML_KNN_EINVAL, "rbf_gamma <= 0");   // predict.py:731-732
    if (!(epsilon > 0)) return fail(OSML_KNN_EINVAL, "epsilon <= 0");       // model.py:11-12
    if (n < 1 || d < 1) return fail(OSML_KNN_EINVAL, "empty train matrix"); // model.py:19-23
    if (index_base < 0 || n_labels < index_base + n) return fail(OSML_KNN_EINVAL, "labels shorter than the shard");
    if (n_labels < (int64_t)k + 1) return fail(OSML_KNN_EINVAL, "fewer than k + 1 training rows");  // predict.py:768
    if (dtype_size(dtype) == 0) return fail(OSML_KNN_EUNSUPPORTED, "unsupported train dtype");
    if (layout != OSML_KNN_ROW_MAJOR && layout != OSML_KNN_COL_MAJOR) return fail(OSML_KNN_EINVAL, "bad layout");
    if (sum_order != OSML_KNN_SUM_PAIRWISE && sum_order != OSML_KNN_SUM_SEQUENTIAL)
        return fail(OSML_KNN_EINVAL, "bad sum_order");
    if (k + 1 > osml::kAggMaxK) return fail(OSML_KNN_EUNSUPPORTED, "k > 1024 is not supported on the GPU path");
    if (n >= 0x7fffffff || d >= (1 << 24)) return fail(OSML_KNN_EUNSUPPORTED, "shard too large (n < 2^31, d < 2^24)");

    int count = 0;
    CUDA_TRY(cudaGetDeviceCount(&count));
    if (device < 0 || device >= count) return fail(OSML_KNN_ECUDA, "no such CUDA device");
    DeviceGuard guard(device);

    osml_knn_index* ix = new (std::nothrow) osml_knn_index();
    if (!ix) return fail(OSML_KNN_ENOMEM, "host allocation failed");
    ix->device = device; ix->dtype = dtype; ix->n = n; ix->d = d; ix->sum_order = sum_order; ix->k = k;
    ix->gamma = rbf_gamma; ix->gamma_is_f32 = gamma_is_f32 ? 1 : 0; ix->epsilon = epsilon;
    ix->index_base = index_base; ix->n_labels = n_labels;

    auto cleanup = [&](int rc) { osml_knn_destroy(ix); return rc; };
    std::vector<int64_t> bits;
    if (convert_labels(labels, label_dtype, n_labels, bits, &ix->label_kind))
        return cleanup(fail(OSML_KNN_EUNSUPPORTED, "unsupported label dtype"));
    ix->label_is_f32 = label_dtype == OSML_KNN_F32 ? 1 : 0;
    cudaError_t e;
    if ((e = cudaStreamCreateWithFlags(&ix->stream, cudaStreamNonBlocking)) != cudaSuccess ||
        (e = cudaStreamCreateWithFlags(&ix->copy_stream, cudaStreamNonBlocking)) != cudaSuccess ||
        (e = cudaEventCreateWithFlags(&ix->ev_copy, cudaEventDisableTiming)) != cudaSuccess ||
        (e = cudaEventCreateWithFlags(&ix->ev_first, cudaEventDisableTiming)) != cudaSuccess ||
        (e = cudaDeviceGetAttribute(&ix->sm_count, cudaDevAttrMultiProcessorCount, device)) != cudaSuccess ||
        (e = cudaEventCreate(&ix->ev0)) != cudaSuccess || (e = cudaEventCreate(&ix->ev1)) != cudaSuccess ||
        (e = cudaEventCreate(&ix->ev2)) != cudaSuccess || (e = cudaEventCreate(&ix->ev3)) != cudaSuccess)
        return cleanup(fail(OSML_KNN_ECUDA, cudaGetErrorString(e)));
    int rc = upload_matrix(ix, train, layout, n, d, ix->train, ix->stream);
    if (rc != OSML_KNN_OK) return cleanup(rc);
    if ((e = ix->label_bits.reserve((size_t)n_labels * 8)) != cudaSuccess ||
        (e = cudaMemcpyAsync(ix->label_bits.ptr, bits.data(), (size_t)n_labels * 8, cudaMemcpyHostToDevice,
                             ix->stream)) != cudaSuccess)
        return cleanup(fail(e == cudaErrorMemoryAllocation ? OSML_KNN_ENOMEM : OSML_KNN_ECUDA, cudaGetErrorString(e)));
    if ((e = cudaStreamSynchronize(ix->stream)) != cudaSuccess) return cleanup(fail(OSML_KNN_ECUDA, cudaGetErrorString(e)));
    ix->stage.release();
    if (is_int_dtype(dtype) && osml::tensor_path_supported(OSML_KNN_F64, n, d, k) && n >= kAutoTensorRows) {
        rc = int_scan_train(ix);          // can the tensor path serve this integer matrix? (value range, see int_type_limit)
        if (rc != OSML_KNN_OK) return cleanup(rc);
    }
    *out = ix;
    return OSML_KNN_OK;
}

int osml_knn_destroy(osml_knn_index* ix) {
    if (!ix) return OSML_KNN_OK;
    DeviceGuard guard(ix->device);
    osml::tensor_release(ix->tensor);
    ix->train.release(); ix->label_bits.release(); ix->stage.release(); ix->queries.release();
    ix->part_d.release(); ix->part_i.release(); ix->nn_d.release(); ix->nn_i.release(); ix->pred.release();
    ix->train_f64.release(); ix->q_f64.release(); ix->q_force.release(); ix->q_back.release();
    ix->q_shift.release(); ix->q_denom.release();
    if (ix->ev0) cudaEventDestroy(ix->ev0);
    if (ix->ev1) cudaEventDestroy(ix->ev1);
    if (ix->ev2) cudaEventDestroy(ix->ev2);
    if (ix->ev3) cudaEventDestroy(ix->ev3);
    if (ix->ev_copy) cudaEventDestroy(ix->ev_copy);
    if (ix->ev_first) cudaEventDestroy(ix->ev_first);
    if (ix->copy_stream) cudaStreamDestroy(ix->copy_stream);
    if (ix->stream) cudaStreamDestroy(ix->stream);
    delete ix;
    return OSML_KNN_OK;
}

int osml_knn_neighbours_device(osml_knn_index* ix, const void* d_queries, int64_t q, int64_t* d_out_idx,
                               void* d_out_dist, void* stream) {
    if (!ix || !d_queries || !d_out_idx || !d_out_dist) return fail(OSML_KNN_EINVAL, "null argument");
    int rc = check_predict_args(ix, q);
    if (rc != OSML_KNN_OK) return rc;
    DeviceGuard guard(ix->device);
    return neighbours(ix, d_queries, q, d_out_dist, d_out_idx, static_cast<cudaStream_t>(stream));
}

}  // extern "C"

namespace {

// y = (x - shift) / denom on a row-major device matrix of the index dtype (in == out allowed)
int apply_query_transform(osml_knn_index* ix, const void* d_in, void* d_out, int64_t q, cudaStream_t stream) {
    const void* denom = ix->has_q_denom ? ix->q_denom.ptr : nullptr;
    cudaError_t e = ix->dtype == OSML_KNN_F32
        ? osml::launch_column_transform<float, float>(d_in, d_out, q, ix->d, 0, ix->q_shift.ptr, denom, stream)
        : osml::launch_column_transform<double, double>(d_in, d_out, q, ix->d, 0, ix->q_shift.ptr, denom, stream);
    ix->stats.kernel_launches++;
    if (e != cudaSuccess) return fail(OSML_KNN_ECUDA, cudaGetErrorString(e));
    return OSML_KNN_OK;
}

int predict_device_impl(osml_knn_index* ix, const void* d_queries, int64_t q, int task, double* d_out_pred,
                        int64_t* d_out_idx, void* d_out_dist, cudaStream_t stream);

}  // namespace

extern "C" {

int osml_knn_predict_device(osml_knn_index* ix, const void* d_queries, int64_t q, int task, double* d_out_pred,
                            int64_t* d_out_idx, void* d_out_dist, void* stream_v) {
    int rc = check_predict_args(ix, q);
    if (rc != OSML_KNN_OK) return rc;
    if (!d_queries || !d_out_pred) return fail(OSML_KNN_EINVAL, "null argument");
    cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
    if (ix->has_q_transform) {            // the caller's buffer stays untouched: transform into the index's copy
        DeviceGuard guard(ix->device);
        CUDA_TRY(ix->queries.reserve((size_t)q * ix->d * dtype_size(ix->dtype)));
        rc = apply_query_transform(ix, d_queries, ix->queries.ptr, q, stream);
        if (rc != OSML_KNN_OK) return rc;
        d_queries = ix->queries.ptr;
    }
    return predict_device_impl(ix, d_queries, q, task, d_out_pred, d_out_idx, d_out_dist, stream);
}

}  // extern "C"

namespace {

int predict_device_impl(osml_knn_index* ix, const void* d_queries, int64_t q, int task, double* d_out_pred,
                        int64_t* d_out_idx, void* d_out_dist, cudaStream_t stream) {
    int rc = check_predict_args(ix, q);
    if (rc != OSML_KNN_OK) return rc;
    if (!d_queries || !d_out_pred) return fail(OSML_KNN_EINVAL, "null argument");
    if (ix->n != ix->n_labels || ix->index_base != 0)
        return fail(OSML_KNN_EINVAL, "index holds a shard: use neighbours + merge_predict");
    DeviceGuard guard(ix->device);
    const size_t ksel = (size_t)ix->ksel();
    void* nn_d = d_out_dist;
    int64_t* nn_i = d_out_idx;
    if (!nn_d) { CUDA_TRY(ix->nn_d.reserve((size_t)q * ksel * ix->dist_size())); nn_d = ix->nn_d.ptr; }
    if (!nn_i) { CUDA_TRY(ix->nn_i.reserve((size_t)q * ksel * 8)); nn_i = ix->nn_i.as<int64_t>(); }
    CUDA_TRY(cudaEventRecord(ix->ev1, stream));
    rc = neighbours(ix, d_queries, q, nn_d, nn_i, stream);
    if (rc != OSML_KNN_OK) return rc;
    CUDA_TRY(cudaEventRecord(ix->ev2, stream));
    rc = aggregate(ix, nn_d, nn_i, q, task, d_out_pred, stream);
    if (rc != OSML_KNN_OK) return rc;
    ix->stats.queries += q;
    return OSML_KNN_OK;
}

}  // namespace

extern "C" {

int osml_knn_predict(osml_knn_index* ix, const void* queries, int layout, int64_t q, int task, double* out_pred,
                     int64_t* out_idx, void* out_dist) {
    int rc = check_predict_args(ix, q);
    if (rc != OSML_KNN_OK) return rc;
    if (!queries || !out_pred) return fail(OSML_KNN_EINVAL, "null argument");
    if (layout != OSML_KNN_ROW_MAJOR && layout != OSML_KNN_COL_MAJOR) return fail(OSML_KNN_EINVAL, "bad layout");
    DeviceGuard guard(ix->device);
    const size_t ksel = (size_t)ix->ksel();
    const size_t es = dtype_size(ix->dtype);
    CUDA_TRY(ix->queries.reserve((size_t)q * ix->d * es));
    if (layout == OSML_KNN_COL_MAJOR && q > 1 && ix->d > 1) CUDA_TRY(ix->stage.reserve((size_t)q * ix->d * es));
    CUDA_TRY(ix->nn_d.reserve((size_t)q * ksel * ix->dist_size()));
    CUDA_TRY(ix->nn_i.reserve((size_t)q * ksel * 8));
    CUDA_TRY(ix->pred.reserve((size_t)q * 8));
    // Pipelined upload: queries are independent, so a large matrix goes up in blocks and every block but the first is
    // copied on a second stream once the block in front of it has its kernels queued, i.e. under that block's K1 instead
    // of in front of everything.  A call of two waves of K1 work items or more: the first block is one full wave
    // (sm_count x 128 queries: every CTA pair sweeps one tile pair, nothing is split), the rest follows in one piece.
    // Smaller calls whose upload is worth hiding go up in two equal blocks: K1 cuts either block into enough train
    // segments to fill the machine (and the segments of a query share their thresholds), so the split costs a few list
    // warm-ups and hides half of the copy.
    const int64_t wave = (int64_t)ix->sm_count * 128;
    std::vector<int64_t> ends;
    if (ix->opt_pipeline && wave > 0) {
        if (q >= 2 * wave) ends = {wave, q};
        else if (q >= 4096 && (size_t)q * ix->d * es >= ((size_t)4 << 20)) {
            // ... but only where the copy is a sizeable part of the call: ~40 GB/s up against ~1.2 PFLOP/s of filter
            // (measured at 1M x 128 rows: 12 500 queries 3.3 ms in halves, 2.8 ms whole; 10 000 x 784 against 60 000
            // rows: 1.9 ms in halves, 2.0 ms whole)
            const double copy_ms = (double)q * ix->d * es / 40e6;
            const double filter_ms = 2.0 * (double)ix->n * ix->d * (double)q / 1.2e12;
            if (copy_ms >= 0.3 * filter_ms) {
                // (two blocks: measured at 10 000 x 784 queries against 60 000 rows, pinned: whole 1.50 ms, two blocks 1.33,
                // three 1.38, four 1.49 -- every extra block restarts all its lists at once; option pipeline_blocks)
                // a pageable matrix (a pandas frame) is bound by its gathering into pinned memory, not by K1: three blocks
                // leave a shorter tail behind the last copy (1.60 ms against 1.81 in two, same shape)
                const bool slow_source = ix->opt_stager && osml::is_pageable_host(queries);
                const int nb = ix->opt_pipeline_blocks > 1 ? std::min(ix->opt_pipeline_blocks, 16) : (slow_source ? 3 : 2);
                const int64_t per = ((q + nb - 1) / nb + 255) / 256 * 256;
                for (int64_t e = per; e < q; e += per) ends.push_back(e);
                ends.push_back(q);
            }
        }
        int64_t smallest = q, prev = 0;
        for (int64_t e : ends) { smallest = std::min(smallest, e - prev); prev = e; }
        if (ends.size() > 1 && (!use_tensor_path(ix, smallest) || is_int_dtype(ix->dtype))) ends.clear();
    }
    const bool pipelined = ends.size() > 1;
    const int64_t qa = pipelined ? ends[0] : q;
    CUDA_TRY(cudaEventRecord(ix->ev0, ix->stream));
    auto transform_rows = [&](int64_t r0, int64_t r1) -> int {   // raw queries: Centering / Standardization in place
        if (!ix->has_q_transform) return OSML_KNN_OK;
        char* dq = ix->queries.as<char>() + (size_t)r0 * ix->d * es;
        return apply_query_transform(ix, dq, dq, r1 - r0, ix->stream);
    };
    rc = upload_rows(ix, queries, layout, q, ix->d, 0, qa, ix->queries, ix->stream);
    if (rc != OSML_KNN_OK) { cudaStreamSynchronize(ix->stream); return rc; }
    if (pipelined) CUDA_TRY(cudaEventRecord(ix->ev_first, ix->stream));     // the second block's copy starts here
    rc = transform_rows(0, qa);
    if (rc != OSML_KNN_OK) return rc;
    int sent = 1;                                   // blocks uploaded so far
    const std::function<int(int)> send_block = [&](int b) -> int {
        if (b != sent || b >= (int)ends.size()) return fail(OSML_KNN_ECUDA, "internal: query blocks out of order");
        if (b == 1) CUDA_TRY(cudaStreamWaitEvent(ix->copy_stream, ix->ev_first, 0));   // the first block keeps the whole link
        ++sent;
        int rc2 = upload_rows(ix, queries, layout, q, ix->d, ends[b - 1], ends[b], ix->queries, ix->copy_stream);
        if (rc2 != OSML_KNN_OK) return rc2;
        CUDA_TRY(cudaEventRecord(ix->ev_copy, ix->copy_stream));
        CUDA_TRY(cudaStreamWaitEvent(ix->stream, ix->ev_copy, 0));
        return transform_rows(ends[b - 1], ends[b]);
    };
    if (pipelined) {
        ix->before_block = &send_block;
        ix->block_ends = ends;
    }
    rc = predict_device_impl(ix, ix->queries.ptr, q, task, ix->pred.as<double>(), ix->nn_i.as<int64_t>(), ix->nn_d.ptr,
                             ix->stream);
    ix->before_block = nullptr;
    ix->block_ends.clear();
    if (rc != OSML_KNN_OK) {
        // the caller's buffer must not be read after we return: uploads may still be in flight on either stream
        cudaStreamSynchronize(ix->stream);
        if (sent > 1) cudaStreamSynchronize(ix->copy_stream);
        return rc;
    }
    if (pipelined && sent != (int)ends.size()) return fail(OSML_KNN_ECUDA, "internal: not every block of queries was uploaded");
    CUDA_TRY(cudaMemcpyAsync(out_pred, ix->pred.ptr, (size_t)q * 8, cudaMemcpyDeviceToHost, ix->stream));
    if (out_idx) CUDA_TRY(cudaMemcpyAsync(out_idx, ix->nn_i.ptr, (size_t)q * ksel * 8, cudaMemcpyDeviceToHost, ix->stream));
    if (out_dist)
        CUDA_TRY(cudaMemcpyAsync(out_dist, ix->nn_d.ptr, (size_t)q * ksel * ix->dist_size(), cudaMemcpyDeviceToHost,
                                 ix->stream));
    CUDA_TRY(cudaEventRecord(ix->ev3, ix->stream));
    CUDA_TRY(cudaStreamSynchronize(ix->stream));
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, ix->ev0, ix->ev3) == cudaSuccess) ix->stats.last_total_ms = ms;
    if (ix->stats.tensor_queries == 0 || ix->tensor.filter_launches == 0)
        if (cudaEventElapsedTime(&ms, ix->ev1, ix->ev2) == cudaSuccess) ix->stats.last_filter_ms = ms;
    return OSML_KNN_OK;
}

int osml_knn_merge_predict_device(osml_knn_index* ix, const int64_t* d_part_idx, const void* d_part_dist, int parts,
                                  int64_t q, int task, double* d_out_pred, int64_t* d_out_idx, void* d_out_dist,
                                  void* stream_v) {
    int rc = check_predict_args(ix, q);
    if (rc != OSML_KNN_OK) return rc;
    if (!d_part_idx || !d_part_dist || !d_out_pred) return fail(OSML_KNN_EINVAL, "null argument");
    if (parts < 1 || parts > 64) return fail(OSML_KNN_EINVAL, "parts must be in 1..64");
    DeviceGuard guard(ix->device);
    cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
    const int ksel = ix->ksel();
    void* nn_d = d_out_dist;
    int64_t* nn_i = d_out_idx;
    if (!nn_d) { CUDA_TRY(ix->nn_d.reserve((size_t)q * ksel * ix->dist_size())); nn_d = ix->nn_d.ptr; }
    if (!nn_i) { CUDA_TRY(ix->nn_i.reserve((size_t)q * ksel * 8)); nn_i = ix->nn_i.as<int64_t>(); }
    const int threads = 128;
    const unsigned grid = (unsigned)((q + threads - 1) / threads);
    if (ix->dist_size() == 4)
        osml::knn_merge_lists_kernel<float><<<grid, threads, 0, stream>>>(
            static_cast<const float*>(d_part_dist), d_part_idx, parts, q * ksel, (int64_t)ksel, q, ksel,
            static_cast<float*>(nn_d), nn_i);
    else
        osml::knn_merge_lists_kernel<double><<<grid, threads, 0, stream>>>(
            static_cast<const double*>(d_part_dist), d_part_idx, parts, q * ksel, (int64_t)ksel, q, ksel,
            static_cast<double*>(nn_d), nn_i);
    ix->stats.kernel_launches++;
    CUDA_TRY(cudaGetLastError());
    rc = aggregate(ix, nn_d, nn_i, q, task, d_out_pred, stream);
    if (rc == OSML_KNN_OK) ix->stats.queries += q;
    return rc;
}

int osml_knn_pack_slices_device(osml_knn_index* ix, const int64_t* d_idx, const void* d_dist, int64_t q, int parts,
                                void* d_records, void* stream_v) {
    if (!ix || !d_idx || !d_dist || !d_records) return fail(OSML_KNN_EINVAL, "null argument");
    if (q < 1) return fail(OSML_KNN_EINVAL, "no queries");
    if (parts < 1 || parts > 64) return fail(OSML_KNN_EINVAL, "parts must be in 1..64");
    DeviceGuard guard(ix->device);
    cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
    const int ksel = ix->ksel();
    const int64_t slice_q = (q + parts - 1) / parts;
    const int64_t total = (int64_t)parts * slice_q * ksel;
    const unsigned grid = (unsigned)std::min<int64_t>((total + 255) / 256, (int64_t)ix->sm_count * 16);
    auto* out = static_cast<osml::NeighbourRecord*>(d_records);
    if (ix->dist_size() == 4)
        osml::knn_pack_records_kernel<float><<<grid, 256, 0, stream>>>(static_cast<const float*>(d_dist), d_idx, q, slice_q,
                                                                       parts, ksel, out);
    else
        osml::knn_pack_records_kernel<double><<<grid, 256, 0, stream>>>(static_cast<const double*>(d_dist), d_idx, q,
                                                                        slice_q, parts, ksel, out);
    ix->stats.kernel_launches++;
    CUDA_TRY(cudaGetLastError());
    return OSML_KNN_OK;
}

int osml_knn_merge_predict_packed_device(osml_knn_index* ix, const void* d_records, int parts, int64_t slice_q,
                                         int64_t q_valid, int task, double* d_out_pred, int64_t* d_out_idx,
                                         void* d_out_dist, void* stream_v) {
    int rc = check_predict_args(ix, q_valid);
    if (rc != OSML_KNN_OK) return rc;
    if (!d_records || !d_out_pred) return fail(OSML_KNN_EINVAL, "null argument");
    if (parts < 1 || parts > 64) return fail(OSML_KNN_EINVAL, "parts must be in 1..64");
    if (q_valid > slice_q) return fail(OSML_KNN_EINVAL, "more valid queries than the slice holds");
    DeviceGuard guard(ix->device);
    cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
    const int ksel = ix->ksel();
    void* nn_d = d_out_dist;
    int64_t* nn_i = d_out_idx;
    if (!nn_d) { CUDA_TRY(ix->nn_d.reserve((size_t)q_valid * ksel * ix->dist_size())); nn_d = ix->nn_d.ptr; }
    if (!nn_i) { CUDA_TRY(ix->nn_i.reserve((size_t)q_valid * ksel * 8)); nn_i = ix->nn_i.as<int64_t>(); }
    const int threads = 128;
    const unsigned grid = (unsigned)((q_valid + threads - 1) / threads);
    const auto* in = static_cast<const osml::NeighbourRecord*>(d_records);
    if (ix->dist_size() == 4)
        osml::knn_merge_records_kernel<float><<<grid, threads, 0, stream>>>(in, parts, slice_q, q_valid, ksel,
                                                                            static_cast<float*>(nn_d), nn_i);
    else
        osml::knn_merge_records_kernel<double><<<grid, threads, 0, stream>>>(in, parts, slice_q, q_valid, ksel,
                                                                             static_cast<double*>(nn_d), nn_i);
    ix->stats.kernel_launches++;
    CUDA_TRY(cudaGetLastError());
    rc = aggregate(ix, nn_d, nn_i, q_valid, task, d_out_pred, stream);
    if (rc == OSML_KNN_OK) ix->stats.queries += q_valid;
    return rc;
}

int osml_knn_set_query_transform(osml_knn_index* ix, const void* shift, const void* denom) {
    if (!ix) return fail(OSML_KNN_EINVAL, "null index");
    if (!shift) {
        if (denom) return fail(OSML_KNN_EINVAL, "denom without shift");
        ix->has_q_transform = ix->has_q_denom = false;
        return OSML_KNN_OK;
    }
    if (ix->dtype != OSML_KNN_F32 && ix->dtype != OSML_KNN_F64)
        return fail(OSML_KNN_EUNSUPPORTED, "the fused query transform needs a float32 or float64 index");
    DeviceGuard guard(ix->device);
    const size_t bytes = (size_t)ix->d * dtype_size(ix->dtype);
    CUDA_TRY(ix->q_shift.reserve(bytes));
    CUDA_TRY(cudaMemcpy(ix->q_shift.ptr, shift, bytes, cudaMemcpyHostToDevice));
    if (denom) {
        CUDA_TRY(ix->q_denom.reserve(bytes));
        CUDA_TRY(cudaMemcpy(ix->q_denom.ptr, denom, bytes, cudaMemcpyHostToDevice));
    }
    ix->has_q_transform = true;
    ix->has_q_denom = denom != nullptr;
    return OSML_KNN_OK;
}

int osml_knn_column_transform(int device, const void* in, int dtype, int layout, int64_t n, int64_t d,
                              const void* shift, const void* denom, void* out) {
    if (!in || !out || !shift) return fail(OSML_KNN_EINVAL, "null argument");
    if (n < 1 || d < 1) return fail(OSML_KNN_EINVAL, "empty matrix");
    if (layout != OSML_KNN_ROW_MAJOR && layout != OSML_KNN_COL_MAJOR) return fail(OSML_KNN_EINVAL, "bad layout");
    const size_t in_es = dtype_size(dtype);
    if (in_es == 0) return fail(OSML_KNN_EUNSUPPORTED, "unsupported dtype");
    int count = 0;
    CUDA_TRY(cudaGetDeviceCount(&count));
    if (device < 0 || device >= count) return fail(OSML_KNN_ECUDA, "no such CUDA device");
    DeviceGuard guard(device);
    const size_t out_es = dtype == OSML_KNN_F32 ? 4 : 8;
    const size_t elems = (size_t)n * d;
    if (d >= (1 << 24)) return fail(OSML_KNN_EUNSUPPORTED, "too many columns");
    // device staging buffers are kept per thread between calls (cudaMalloc / cudaFree cost more than the kernel) and
    // given back when they grow past 1 GB
    struct Scratch { int device = -1; DeviceBuffer in, out, shift, denom; };
    static thread_local Scratch sc;
    if (sc.device != device) {
        if (sc.device >= 0) { DeviceGuard g2(sc.device); sc.in.release(); sc.out.release(); sc.shift.release(); sc.denom.release(); }
        sc.device = device;
    }
    DeviceBuffer &d_in = sc.in, &d_out = sc.out, &d_shift = sc.shift, &d_denom = sc.denom;
    auto done = [&](int rc) {
        if (d_in.cap + d_out.cap > ((size_t)1 << 30) || rc != OSML_KNN_OK) {
            d_in.release(); d_out.release(); d_shift.release(); d_denom.release();
        }
        return rc;
    };
    cudaError_t e;
    if ((e = d_in.reserve(elems * in_es)) != cudaSuccess || (e = d_out.reserve(elems * out_es)) != cudaSuccess ||
        (e = d_shift.reserve((size_t)d * out_es)) != cudaSuccess ||
        (denom && (e = d_denom.reserve((size_t)d * out_es)) != cudaSuccess))
        return done(fail(e == cudaErrorMemoryAllocation ? OSML_KNN_ENOMEM : OSML_KNN_ECUDA, cudaGetErrorString(e)));
    // the matrix goes up as one flat run of bytes whatever its layout (the kernel reads either); pageable memory -- a
    // frame's values -- through the pinned staging ring (host_stager.cuh)
    if (elems * in_es >= ((size_t)2 << 20) && osml::is_pageable_host(in)) {
        int rc = staged_copy_rowmajor(device, static_cast<const char*>(in), d_in.as<char>(), elems * in_es, nullptr);
        if (rc != OSML_KNN_OK) return done(rc);
        e = cudaSuccess;
    } else e = cudaMemcpy(d_in.ptr, in, elems * in_es, cudaMemcpyHostToDevice);
    if (e != cudaSuccess ||
        (e = cudaMemcpy(d_shift.ptr, shift, (size_t)d * out_es, cudaMemcpyHostToDevice)) != cudaSuccess ||
        (denom && (e = cudaMemcpy(d_denom.ptr, denom, (size_t)d * out_es, cudaMemcpyHostToDevice)) != cudaSuccess))
        return done(fail(OSML_KNN_ECUDA, cudaGetErrorString(e)));
    const int cm = layout == OSML_KNN_COL_MAJOR ? 1 : 0;
    const void* dn = denom ? d_denom.ptr : nullptr;
    switch (dtype) {
        case OSML_KNN_F32: e = osml::launch_column_transform<float, float>(d_in.ptr, d_out.ptr, n, d, cm, d_shift.ptr, dn, nullptr); break;
        case OSML_KNN_F64: e = osml::launch_column_transform<double, double>(d_in.ptr, d_out.ptr, n, d, cm, d_shift.ptr, dn, nullptr); break;
        case OSML_KNN_I8: e = osml::launch_column_transform<int8_t, double>(d_in.ptr, d_out.ptr, n, d, cm, d_shift.ptr, dn, nullptr); break;
        case OSML_KNN_I16: e = osml::launch_column_transform<int16_t, double>(d_in.ptr, d_out.ptr, n, d, cm, d_shift.ptr, dn, nullptr); break;
        case OSML_KNN_I32: e = osml::launch_column_transform<int32_t, double>(d_in.ptr, d_out.ptr, n, d, cm, d_shift.ptr, dn, nullptr); break;
        case OSML_KNN_I64: e = osml::launch_column_transform<int64_t, double>(d_in.ptr, d_out.ptr, n, d, cm, d_shift.ptr, dn, nullptr); break;
        case OSML_KNN_U8: e = osml::launch_column_transform<uint8_t, double>(d_in.ptr, d_out.ptr, n, d, cm, d_shift.ptr, dn, nullptr); break;
        case OSML_KNN_U16: e = osml::launch_column_transform<uint16_t, double>(d_in.ptr, d_out.ptr, n, d, cm, d_shift.ptr, dn, nullptr); break;
        case OSML_KNN_U32: e = osml::launch_column_transform<uint32_t, double>(d_in.ptr, d_out.ptr, n, d, cm, d_shift.ptr, dn, nullptr); break;
        default: e = osml::launch_column_transform<uint64_t, double>(d_in.ptr, d_out.ptr, n, d, cm, d_shift.ptr, dn, nullptr); break;
    }
    if (e != cudaSuccess) return done(fail(OSML_KNN_ECUDA, cudaGetErrorString(e)));
    // ... and comes back the same way: pageable results (a fresh numpy array) through the ring, copied out by the workers
    if (elems * out_es >= ((size_t)2 << 20) && osml::is_pageable_host(out)) {
        std::vector<osml::HostStager::OutChunk> chunks;
        const size_t bytes = elems * out_es;
        for (size_t off = 0; off < bytes; off += osml::HostStager::kSlotBytes) {
            const size_t nb = std::min(osml::HostStager::kSlotBytes, bytes - off);
            char* dst = static_cast<char*>(out) + off;
            chunks.push_back({d_out.as<char>() + off, nb, [dst, nb](const char* slot) { memcpy(dst, slot, nb); }});
        }
        e = osml::HostStager::instance().copy_out(device, chunks, nullptr);
    } else e = cudaMemcpy(out, d_out.ptr, elems * out_es, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) return done(fail(OSML_KNN_ECUDA, cudaGetErrorString(e)));
    return done(OSML_KNN_OK);
}

int osml_knn_set_option(osml_knn_index* ix, const char* name, double value) {
    if (!ix || !name) return fail(OSML_KNN_EINVAL, "null argument");
    const std::string key(name);
    if (key == "path") {
        const int v = (int)value;
        if (v < OSML_KNN_PATH_AUTO || v > OSML_KNN_PATH_TENSOR) return fail(OSML_KNN_EINVAL, "bad path");
        if (v == OSML_KNN_PATH_TENSOR && is_int_dtype(ix->dtype) && ix->int_limit < 0 &&
            osml::tensor_path_supported(OSML_KNN_F64, ix->n, ix->d, ix->k)) {
            DeviceGuard guard(ix->device);            // small integer matrices are not scanned at create
            int rc = int_scan_train(ix);
            if (rc != OSML_KNN_OK) return rc;
        }
        if (v == OSML_KNN_PATH_TENSOR && (!osml::tensor_path_supported(tensor_dtype(ix), ix->n, ix->d, ix->k) ||
                                          (is_int_dtype(ix->dtype) && !ix->int_tensor_ok)))
            return fail(OSML_KNN_EUNSUPPORTED, "tensor path needs n >= 256 (2048 x segments for k > 47), d <= 8189, k <= 250 and float data, or "
                                               "integer data whose value range cannot wrap in numpy's arithmetic");
        ix->path = v;
        return OSML_KNN_OK;
    }
    if (key == "slack") { ix->slack = (int)value; return OSML_KNN_OK; }
    if (key == "epi_warps") { ix->opt_epi_warps = (int)value; return OSML_KNN_OK; }
    if (key == "segments") { ix->opt_nseg = (int)value; return OSML_KNN_OK; }
    if (key == "debug") { ix->debug = (int)value; return OSML_KNN_OK; }
    if (key == "counters") { ix->counters = (int)value; return OSML_KNN_OK; }
    if (key == "smem_lists") { ix->opt_smem_lists = (int)value; return OSML_KNN_OK; }
    if (key == "cta_group") { ix->opt_cta_group = (int)value; return OSML_KNN_OK; }
    if (key == "active") { ix->opt_active = (int)value; return OSML_KNN_OK; }
    if (key == "precision") {           // 0 = auto, 1 = bf16 filter copy, 2 = fp16 filter copy
        const int v = (int)value;
        if (v < 0 || v > 2) return fail(OSML_KNN_EINVAL, "bad precision");
        ix->opt_precision = v;
        return OSML_KNN_OK;
    }
    if (key == "kprime") { ix->opt_kprime = (int)value; return OSML_KNN_OK; }
    if (key == "wide_pair") { ix->opt_wide_pair = (int)value; return OSML_KNN_OK; }
    if (key == "hot") { ix->opt_hot = (int)value; return OSML_KNN_OK; }
    if (key == "live") { ix->opt_live = (int)value; return OSML_KNN_OK; }
    if (key == "query_block") { ix->query_block = (int64_t)value; return OSML_KNN_OK; }
    if (key == "stager") { ix->opt_stager = value != 0.0; return OSML_KNN_OK; }
    if (key == "pipeline_blocks") { ix->opt_pipeline_blocks = (int)value; return OSML_KNN_OK; }
    if (key == "pipeline") { ix->opt_pipeline = value != 0.0; return OSML_KNN_OK; }
    return fail(OSML_KNN_EINVAL, "unknown option: " + key);
}

int osml_knn_get_stats(osml_knn_index* ix, osml_knn_stats* out) {
    if (!ix || !out) return fail(OSML_KNN_EINVAL, "null argument");
    ix->stats.device_bytes = (int64_t)(ix->train.cap + ix->label_bits.cap + ix->stage.cap + ix->queries.cap +
                                       ix->part_d.cap + ix->part_i.cap + ix->nn_d.cap + ix->nn_i.cap + ix->pred.cap +
                                       ix->train_f64.cap + ix->q_f64.cap + ix->q_force.cap + ix->q_back.cap +
                                       osml::tensor_bytes(ix->tensor));
    ix->stats.filter_precision = ix->tensor.ready ? (ix->tensor.f16 ? 2 : 1) : 0;
    *out = ix->stats;
    return OSML_KNN_OK;
}

int osml_knn_reset_stats(osml_knn_index* ix) {
    if (!ix) return fail(OSML_KNN_EINVAL, "null argument");
    ix->stats = osml_knn_stats{};
    return OSML_KNN_OK;
}

int osml_knn_describe(osml_knn_index* ix, int64_t* n, int64_t* d, int* k, int* dtype, int* dist_elem_size) {
    if (!ix) return fail(OSML_KNN_EINVAL, "null argument");
    if (n) *n = ix->n;
    if (d) *d = ix->d;
    if (k) *k = ix->k;
    if (dtype) *dtype = ix->dtype;
    if (dist_elem_size) *dist_elem_size = (int)ix->dist_size();
    return OSML_KNN_OK;
}

}  // extern "C"

#include "svm_api.cuh"
