// K0 / K1 / K3: tcgen05 tensor-core candidate filter with exact rescoring (interface; kernels in knn_tensor_impl.cuh).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <functional>
#include <string>

namespace osml {

struct TensorState {
    bool ready = false;
    size_t bytes = 0;
};

struct TensorLaunch {
    const void* train; int dtype; int64_t n, d; int k; int sum_order; int64_t index_base; int slack;
};

using DirectFallback = std::function<int(const void*, int64_t, void*, int64_t*)>;

inline bool tensor_path_supported(int, int64_t, int64_t, int) { return false; }
inline size_t tensor_bytes(const TensorState& s) { return s.bytes; }
inline void tensor_release(TensorState&) {}
inline int tensor_neighbours(TensorState&, const TensorLaunch&, const void*, int64_t, void*, int64_t*, cudaStream_t,
                             int64_t*, int64_t*, std::string* err, const DirectFallback&) {
    if (err) *err = "tensor path not built";
    return 4;
}

}  // namespace osml
