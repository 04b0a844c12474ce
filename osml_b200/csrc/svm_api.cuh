// C ABI of the support-vector decision function (include/osml_svm.h).  Included by knn_api.cu: one translation unit,
// one shared library, one error string.
#pragma once
#include "../../include/osml_svm.h"
#include "svm_decision.cuh"

struct osml_svm_machine {
    int device = 0, dtype = 0, kernel = 0, sum_order = 1, sm_count = 0;
    int64_t n = 0, d = 0;
    int n_active = 0;
    double intercept = 0.0;
    osml::SvmParams params{};
    DeviceBuffer vectors, labels, alphas, active, queries, stage, partial, out, plan;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    double last_kernel_ms = 0.0;
    int64_t kernel_launches = 0;
};

namespace {

// host matrix (either layout) -> row-major device matrix; column-major input is transposed on the device
int svm_upload(osml_svm_machine* m, const void* host, int layout, int64_t rows, int64_t cols, DeviceBuffer& dst) {
    const size_t es = dtype_size(m->dtype);
    const size_t bytes = (size_t)rows * cols * es;
    CUDA_TRY(dst.reserve(bytes));
    if (layout == OSML_KNN_ROW_MAJOR || rows == 1 || cols == 1) {
        CUDA_TRY(cudaMemcpyAsync(dst.ptr, host, bytes, cudaMemcpyHostToDevice, m->stream));
        return OSML_KNN_OK;
    }
    CUDA_TRY(m->stage.reserve(bytes));
    CUDA_TRY(cudaMemcpyAsync(m->stage.ptr, host, bytes, cudaMemcpyHostToDevice, m->stream));
    dim3 block(32, 8), grid((unsigned)((rows + 31) / 32), (unsigned)((cols + 31) / 32));
    switch (es) {
        case 1: transpose_colmajor_kernel<uint8_t><<<grid, block, 0, m->stream>>>(m->stage.as<uint8_t>(), dst.as<uint8_t>(), rows, cols); break;
        case 2: transpose_colmajor_kernel<uint16_t><<<grid, block, 0, m->stream>>>(m->stage.as<uint16_t>(), dst.as<uint16_t>(), rows, cols); break;
        case 4: transpose_colmajor_kernel<uint32_t><<<grid, block, 0, m->stream>>>(m->stage.as<uint32_t>(), dst.as<uint32_t>(), rows, cols); break;
        default: transpose_colmajor_kernel<uint64_t><<<grid, block, 0, m->stream>>>(m->stage.as<uint64_t>(), dst.as<uint64_t>(), rows, cols); break;
    }
    m->kernel_launches++;
    CUDA_TRY(cudaGetLastError());
    return OSML_KNN_OK;
}

int svm_upload_multipliers(osml_svm_machine* m, const double* alphas) {
    std::vector<int> active;
    active.reserve((size_t)m->n);
    // Rows whose multiplier is exactly zero add exact zeros and are skipped -- for float data only: integer data can wrap
    // to a negative sum of squares, whose NaN distance makes the reference's term NaN * 0 = NaN, so those rows stay.
    const bool skip_zero = m->dtype == OSML_KNN_F32 || m->dtype == OSML_KNN_F64;
    for (int64_t i = 0; i < m->n; ++i)
        if (!skip_zero || alphas[i] != 0.0) active.push_back((int)i);          // NaN multipliers stay in (NaN != 0)
    m->n_active = (int)active.size();
    CUDA_TRY(m->alphas.reserve((size_t)m->n * 8));
    CUDA_TRY(m->active.reserve((size_t)m->n * 4));
    CUDA_TRY(cudaMemcpyAsync(m->alphas.ptr, alphas, (size_t)m->n * 8, cudaMemcpyHostToDevice, m->stream));
    if (m->n_active > 0)
        CUDA_TRY(cudaMemcpyAsync(m->active.ptr, active.data(), (size_t)m->n_active * 4, cudaMemcpyHostToDevice, m->stream));
    CUDA_TRY(cudaStreamSynchronize(m->stream));                  // `active` is a local
    return OSML_KNN_OK;
}

template <typename T>
cudaError_t svm_launch_typed(osml_svm_machine* m, const void* d_queries, int q, int* nsplit, cudaStream_t stream) {
    return osml::launch_svm_decision<T>(m->kernel, m->vectors.ptr, m->active.as<int>(), m->n_active,
                                        m->alphas.as<double>(), m->labels.as<double>(), d_queries, (int)m->d, q,
                                        m->sum_order == OSML_KNN_SUM_PAIRWISE, m->params, m->sm_count,
                                        m->plan.as<osml::SvmPlan>(), m->partial.as<double>(), nsplit, stream);
}

// d_queries row-major on the device -> d_out; queued on `stream`
int svm_decide(osml_svm_machine* m, const void* d_queries, int64_t q, double* d_out, cudaStream_t stream) {
    if (q < 1) return fail(OSML_KNN_EINVAL, "no queries");
    if (q > 0x7fffffff / 16) return fail(OSML_KNN_EUNSUPPORTED, "more than 2^27 queries in one call");
    const int splits = osml::svm_partial_splits((int)q, m->n_active, m->sm_count);
    CUDA_TRY(m->partial.reserve((size_t)splits * q * 8));
    int nsplit = 1;
    CUDA_TRY(cudaEventRecord(m->ev0, stream));
    cudaError_t e;
    switch (m->dtype) {
        case OSML_KNN_F32: e = svm_launch_typed<float>(m, d_queries, (int)q, &nsplit, stream); break;
        case OSML_KNN_F64: e = svm_launch_typed<double>(m, d_queries, (int)q, &nsplit, stream); break;
        case OSML_KNN_I8: e = svm_launch_typed<int8_t>(m, d_queries, (int)q, &nsplit, stream); break;
        case OSML_KNN_I16: e = svm_launch_typed<int16_t>(m, d_queries, (int)q, &nsplit, stream); break;
        case OSML_KNN_I32: e = svm_launch_typed<int32_t>(m, d_queries, (int)q, &nsplit, stream); break;
        case OSML_KNN_I64: e = svm_launch_typed<int64_t>(m, d_queries, (int)q, &nsplit, stream); break;
        case OSML_KNN_U8: e = svm_launch_typed<uint8_t>(m, d_queries, (int)q, &nsplit, stream); break;
        case OSML_KNN_U16: e = svm_launch_typed<uint16_t>(m, d_queries, (int)q, &nsplit, stream); break;
        case OSML_KNN_U32: e = svm_launch_typed<uint32_t>(m, d_queries, (int)q, &nsplit, stream); break;
        default: e = svm_launch_typed<uint64_t>(m, d_queries, (int)q, &nsplit, stream); break;
    }
    if (e != cudaSuccess) return fail(OSML_KNN_ECUDA, cudaGetErrorString(e));
    // the launch writes nsplit * q partial sums: never more than was reserved (compute-sanitizer is closed on this pool,
    // so the bound is checked here)
    if (nsplit > splits) return fail(OSML_KNN_ECUDA, "internal: decision kernel used more row splits than reserved");
    CUDA_TRY(cudaEventRecord(m->ev1, stream));
    osml::svm_finish_kernel<<<(unsigned)((q + 255) / 256), 256, 0, stream>>>(m->partial.as<double>(), nsplit, (int)q,
                                                                             m->intercept, d_out);
    m->kernel_launches += 2;
    CUDA_TRY(cudaGetLastError());
    return OSML_KNN_OK;
}

int svm_finish_host_call(osml_svm_machine* m, int64_t q, double* out) {
    CUDA_TRY(cudaMemcpyAsync(out, m->out.ptr, (size_t)q * 8, cudaMemcpyDeviceToHost, m->stream));
    CUDA_TRY(cudaStreamSynchronize(m->stream));
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, m->ev0, m->ev1) == cudaSuccess) m->last_kernel_ms = ms;
    return OSML_KNN_OK;
}

}  // namespace

extern "C" {

int osml_svm_create(osml_svm_machine** out, int device, const void* vectors, int dtype, int layout, int64_t n,
                    int64_t d, int sum_order, const double* labels, const double* alphas, double intercept,
                    int kernel, double gamma, double degree, double poly_coefficient) {
    if (!out) return fail(OSML_KNN_EINVAL, "null out");
    *out = nullptr;
    if (!vectors || !labels || !alphas) return fail(OSML_KNN_EINVAL, "null input");
    if (n < 1 || d < 1) return fail(OSML_KNN_EINVAL, "empty vector matrix");
    if (kernel < OSML_SVM_LINEAR || kernel > OSML_SVM_GAUSSIAN) return fail(OSML_KNN_EINVAL, "bad kernel");   // predict.py:1069-1070
    if (kernel == OSML_SVM_GAUSSIAN && !(gamma > 0)) return fail(OSML_KNN_EINVAL, "gamma <= 0");              // predict.py:1066-1067
    if (kernel == OSML_SVM_POLYNOMIAL && !(degree >= 1)) return fail(OSML_KNN_EINVAL, "degree < 1");          // predict.py:1062-1063
    if (dtype_size(dtype) == 0) return fail(OSML_KNN_EUNSUPPORTED, "unsupported dtype");
    if (layout != OSML_KNN_ROW_MAJOR && layout != OSML_KNN_COL_MAJOR) return fail(OSML_KNN_EINVAL, "bad layout");
    if (sum_order != OSML_KNN_SUM_PAIRWISE && sum_order != OSML_KNN_SUM_SEQUENTIAL)
        return fail(OSML_KNN_EINVAL, "bad sum_order");
    if (kernel != OSML_SVM_GAUSSIAN && dtype != OSML_KNN_F32 && dtype != OSML_KNN_F64)
        return fail(OSML_KNN_EUNSUPPORTED, "the lin and poly kernels need float32 or float64 data on the GPU path");
    if (n >= 0x7fffffff || d >= (1 << 24)) return fail(OSML_KNN_EUNSUPPORTED, "matrix too large (n < 2^31, d < 2^24)");

    int count = 0;
    CUDA_TRY(cudaGetDeviceCount(&count));
    if (device < 0 || device >= count) return fail(OSML_KNN_ECUDA, "no such CUDA device");
    DeviceGuard guard(device);
    int smem_optin = 0;
    CUDA_TRY(cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
    if (osml::svm_smem_bytes((int)d, dtype_size(dtype)) + 1024 > (size_t)smem_optin)
        return fail(OSML_KNN_EUNSUPPORTED, "rows too wide for the decision kernel's query tile");

    osml_svm_machine* m = new (std::nothrow) osml_svm_machine();
    if (!m) return fail(OSML_KNN_ENOMEM, "host allocation failed");
    m->device = device; m->dtype = dtype; m->kernel = kernel; m->sum_order = sum_order; m->n = n; m->d = d;
    m->intercept = intercept;
    m->params.gamma = gamma; m->params.degree = degree; m->params.coefficient = poly_coefficient;
    auto cleanup = [&](int rc) { osml_svm_destroy(m); return rc; };
    cudaError_t e;
    if ((e = cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking)) != cudaSuccess ||
        (e = cudaEventCreate(&m->ev0)) != cudaSuccess || (e = cudaEventCreate(&m->ev1)) != cudaSuccess ||
        (e = cudaDeviceGetAttribute(&m->sm_count, cudaDevAttrMultiProcessorCount, device)) != cudaSuccess)
        return cleanup(fail(OSML_KNN_ECUDA, cudaGetErrorString(e)));
    int rc = svm_upload(m, vectors, layout, n, d, m->vectors);
    if (rc != OSML_KNN_OK) return cleanup(rc);
    if ((e = m->labels.reserve((size_t)n * 8)) != cudaSuccess ||
        (e = cudaMemcpyAsync(m->labels.ptr, labels, (size_t)n * 8, cudaMemcpyHostToDevice, m->stream)) != cudaSuccess)
        return cleanup(fail(e == cudaErrorMemoryAllocation ? OSML_KNN_ENOMEM : OSML_KNN_ECUDA, cudaGetErrorString(e)));
    {   // leaf plan of numpy's pairwise sum over d features (used by the rbf kernel on C-ordered float vectors)
        osml::SvmPlan plan;
        if (!osml::svm_make_plan((int)d, plan)) return cleanup(fail(OSML_KNN_EUNSUPPORTED, "rows too wide for the pairwise plan"));
        if ((e = m->plan.reserve(sizeof(plan))) != cudaSuccess ||
            (e = cudaMemcpyAsync(m->plan.ptr, &plan, sizeof(plan), cudaMemcpyHostToDevice, m->stream)) != cudaSuccess ||
            (e = cudaStreamSynchronize(m->stream)) != cudaSuccess)
            return cleanup(fail(OSML_KNN_ECUDA, cudaGetErrorString(e)));
    }
    rc = svm_upload_multipliers(m, alphas);
    if (rc != OSML_KNN_OK) return cleanup(rc);
    m->stage.release();
    *out = m;
    return OSML_KNN_OK;
}

int osml_svm_destroy(osml_svm_machine* m) {
    if (!m) return OSML_KNN_OK;
    DeviceGuard guard(m->device);
    for (DeviceBuffer* b : {&m->vectors, &m->labels, &m->alphas, &m->active, &m->queries, &m->stage, &m->partial, &m->out,
                            &m->plan})
        b->release();
    if (m->ev0) cudaEventDestroy(m->ev0);
    if (m->ev1) cudaEventDestroy(m->ev1);
    if (m->stream) cudaStreamDestroy(m->stream);
    delete m;
    return OSML_KNN_OK;
}

int osml_svm_set_multipliers(osml_svm_machine* m, const double* alphas, double intercept) {
    if (!m || !alphas) return fail(OSML_KNN_EINVAL, "null argument");
    DeviceGuard guard(m->device);
    m->intercept = intercept;
    return svm_upload_multipliers(m, alphas);
}

int osml_svm_decision(osml_svm_machine* m, const void* queries, int layout, int64_t q, double* out) {
    if (!m || !queries || !out) return fail(OSML_KNN_EINVAL, "null argument");
    if (q < 1) return fail(OSML_KNN_EINVAL, "no queries");
    if (layout != OSML_KNN_ROW_MAJOR && layout != OSML_KNN_COL_MAJOR) return fail(OSML_KNN_EINVAL, "bad layout");
    DeviceGuard guard(m->device);
    int rc = svm_upload(m, queries, layout, q, m->d, m->queries);
    if (rc != OSML_KNN_OK) return rc;
    CUDA_TRY(m->out.reserve((size_t)q * 8));
    rc = svm_decide(m, m->queries.ptr, q, m->out.as<double>(), m->stream);
    if (rc != OSML_KNN_OK) return rc;
    return svm_finish_host_call(m, q, out);
}

int osml_svm_decision_self(osml_svm_machine* m, double* out) {
    if (!m || !out) return fail(OSML_KNN_EINVAL, "null argument");
    DeviceGuard guard(m->device);
    CUDA_TRY(m->out.reserve((size_t)m->n * 8));
    int rc = svm_decide(m, m->vectors.ptr, m->n, m->out.as<double>(), m->stream);
    if (rc != OSML_KNN_OK) return rc;
    return svm_finish_host_call(m, m->n, out);
}

int osml_svm_gram(osml_svm_machine* m, double* out) {
    if (!m || !out) return fail(OSML_KNN_EINVAL, "null argument");
    if (m->n > 65535) return fail(OSML_KNN_EUNSUPPORTED, "kernel matrix of more than 65535 vectors");
    DeviceGuard guard(m->device);
    const size_t bytes = (size_t)m->n * m->n * 8;
    DeviceBuffer gram;
    CUDA_TRY(gram.reserve(bytes));
    cudaError_t e;
    const int n = (int)m->n, d = (int)m->d;
    switch (m->dtype) {
        case OSML_KNN_F32: e = osml::launch_svm_gram<float>(m->kernel, m->vectors.ptr, n, d, m->params, gram.as<double>(), m->stream); break;
        case OSML_KNN_F64: e = osml::launch_svm_gram<double>(m->kernel, m->vectors.ptr, n, d, m->params, gram.as<double>(), m->stream); break;
        case OSML_KNN_I8: e = osml::launch_svm_gram<int8_t>(m->kernel, m->vectors.ptr, n, d, m->params, gram.as<double>(), m->stream); break;
        case OSML_KNN_I16: e = osml::launch_svm_gram<int16_t>(m->kernel, m->vectors.ptr, n, d, m->params, gram.as<double>(), m->stream); break;
        case OSML_KNN_I32: e = osml::launch_svm_gram<int32_t>(m->kernel, m->vectors.ptr, n, d, m->params, gram.as<double>(), m->stream); break;
        case OSML_KNN_I64: e = osml::launch_svm_gram<int64_t>(m->kernel, m->vectors.ptr, n, d, m->params, gram.as<double>(), m->stream); break;
        case OSML_KNN_U8: e = osml::launch_svm_gram<uint8_t>(m->kernel, m->vectors.ptr, n, d, m->params, gram.as<double>(), m->stream); break;
        case OSML_KNN_U16: e = osml::launch_svm_gram<uint16_t>(m->kernel, m->vectors.ptr, n, d, m->params, gram.as<double>(), m->stream); break;
        case OSML_KNN_U32: e = osml::launch_svm_gram<uint32_t>(m->kernel, m->vectors.ptr, n, d, m->params, gram.as<double>(), m->stream); break;
        default: e = osml::launch_svm_gram<uint64_t>(m->kernel, m->vectors.ptr, n, d, m->params, gram.as<double>(), m->stream); break;
    }
    m->kernel_launches++;
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, gram.ptr, bytes, cudaMemcpyDeviceToHost, m->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(m->stream);
    gram.release();
    if (e != cudaSuccess) return fail(OSML_KNN_ECUDA, cudaGetErrorString(e));
    return OSML_KNN_OK;
}

int osml_svm_decision_device(osml_svm_machine* m, const void* d_queries, int64_t q, double* d_out, void* stream_v) {
    if (!m || !d_queries || !d_out) return fail(OSML_KNN_EINVAL, "null argument");
    DeviceGuard guard(m->device);
    cudaStream_t stream = stream_v ? static_cast<cudaStream_t>(stream_v) : m->stream;
    int rc = svm_decide(m, d_queries, q, d_out, stream);
    if (rc != OSML_KNN_OK) return rc;
    if (!stream_v) {
        CUDA_TRY(cudaStreamSynchronize(m->stream));
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, m->ev0, m->ev1) == cudaSuccess) m->last_kernel_ms = ms;
    }
    return OSML_KNN_OK;
}

int osml_svm_get_stats(osml_svm_machine* m, double* last_kernel_ms, int64_t* kernel_launches) {
    if (!m) return fail(OSML_KNN_EINVAL, "null argument");
    if (last_kernel_ms) *last_kernel_ms = m->last_kernel_ms;
    if (kernel_launches) *kernel_launches = m->kernel_launches;
    return OSML_KNN_OK;
}

}  // extern "C"
