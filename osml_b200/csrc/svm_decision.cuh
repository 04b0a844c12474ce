// Support-vector decision function: osml/predict.py:1086-1099 of the reference,
//     raw[x] = intercept + sum_i K(V_i, x) * alpha_i * label_i,
// for the three kernels of SupportVectorMachine, in the reference's dtype-preserving arithmetic:
//     lin   dot(V_i, x)                                 predict.py:1086-1087   (float dtypes)
//     poly  (coefficient + dot(V_i, x)) ** degree       predict.py:1089-1090   (float dtypes)
//     rbf   exp(-gamma * sqrt(sum_j (V_ij - x_j)^2))    predict.py:1092-1093   (every dtype; the distance is the same
//           expression as KNearestNeighbors._distances, so it reuses Elem<T> / exact_distance of the kNN path: float32
//           stays float32 up to and including exp, integers wrap like numpy, both summation orders)
// The products with the float64 multipliers and labels and the sum over the vectors are float64.
//
// Layout: a CTA of 256 threads serves 8 queries (shared memory) and a contiguous split of the ACTIVE vectors (rows
// whose multiplier is non-zero; during SMO most are zero and their terms are exact zeros).  Thread t walks rows t, t+256,
// ... of its split, evaluates the kernel against the 8 queries with one pass over its row (sequential order) and keeps
// 8 float64 partial sums; a fixed shared-memory tree adds the threads, a second tiny kernel adds the splits in order
// and the intercept -- so the result is deterministic.  It is not numpy's own pairwise order over the vectors, and
// exp / pow differ from numpy's SIMD loops in the last bit: parity with the reference is to ~1e-13 relative (float64
// data; 1e-6 for float32 data, whose kernel values are float32 by the reference's own rules), not bit-exact.
// Bound: CUDA cores, 3 un-fusable instructions per (element, query) for rbf as in K2; lin / poly are one FMA.
#pragma once
#include "knn_common.cuh"

namespace osml {

enum SvmKernel { kSvmLinear = 0, kSvmPolynomial = 1, kSvmGaussian = 2 };

constexpr int kSvmQB = 8;
constexpr int kSvmThreads = 256;

struct SvmParams {
    double gamma, degree, coefficient;
};

// kernel value from the accumulated quantity (dot product or sum of squared differences), in the data's own type
template <typename T, int kKind> struct SvmValue;

template <int kKind> struct SvmValue<float, kKind> {
    __device__ __forceinline__ static double of(float acc, const SvmParams& p) {
        if (kKind == kSvmLinear) return (double)acc;
        if (kKind == kSvmPolynomial) return (double)powf(__fadd_rn((float)p.coefficient, acc), (float)p.degree);
        return (double)expf(__fmul_rn(-(float)p.gamma, Elem<float>::finish(acc)));
    }
};
template <int kKind> struct SvmValue<double, kKind> {
    __device__ __forceinline__ static double of(double acc, const SvmParams& p) {
        if (kKind == kSvmLinear) return acc;
        if (kKind == kSvmPolynomial) return pow(__dadd_rn(p.coefficient, acc), p.degree);
        return exp(__dmul_rn(-p.gamma, Elem<double>::finish(acc)));
    }
};
// integer data: rbf only; the wrapped int64 sum becomes a float64 distance (NaN if the sum wrapped negative)
template <typename T, int kKind> struct SvmValue {
    __device__ __forceinline__ static double of(typename Elem<T>::Acc acc, const SvmParams& p) {
        return exp(__dmul_rn(-p.gamma, Elem<T>::finish(acc)));
    }
};

template <typename T> struct SvmDot {              // accumulator of the lin / poly kernels
    using Acc = T;
    __device__ __forceinline__ static T step(T acc, T t, T q) { return acc + t * q; }     // contracts to one FMA
};

inline size_t svm_smem_bytes(int d, size_t elem) { return ((size_t)d * kSvmQB * elem + 15) / 16 * 16 + 64; }

// numpy's pairwise_sum over d elements, linearised: the recursion (split at n/2 rounded down to a multiple of 8 until a
// block has at most 128 elements) visits its leaves left to right, and after a leaf some number of pending (left, right)
// pairs are complete.  The plan lists (offset, length, merges after it) per leaf; the device keeps a small stack.
constexpr int kSvmMaxLeaves = 256;      // d <= 128 * 256 / 2 and far beyond what the query tile lets through
constexpr int kSvmMaxDepth = 20;
struct SvmLeaf { int lo, n, merges; };
struct SvmPlan { int leaves; SvmLeaf leaf[kSvmMaxLeaves]; };

inline void svm_plan_rec(int lo, int n, SvmPlan& plan, bool& ok) {
    if (n <= 128) {
        if (plan.leaves >= kSvmMaxLeaves) { ok = false; return; }
        plan.leaf[plan.leaves++] = SvmLeaf{lo, n, 0};
        return;
    }
    int n2 = n / 2;
    n2 -= n2 % 8;
    svm_plan_rec(lo, n2, plan, ok);
    svm_plan_rec(lo + n2, n - n2, plan, ok);
    if (ok) plan.leaf[plan.leaves - 1].merges++;        // this node is complete once its right subtree is
}

inline bool svm_make_plan(int d, SvmPlan& plan) {
    plan.leaves = 0;
    bool ok = true;
    svm_plan_rec(0, d, plan, ok);
    return ok;
}

// kPairwise: numpy's pairwise tree along the features (C-ordered vectors, float data, rbf); otherwise feature by feature.
// kQB queries per CTA (8; 4 for the float64 pairwise kernel, whose 8 accumulators per query would not fit otherwise).
template <typename T, int kKind, bool kPairwise, bool kVec, int kQB>
__global__ void __launch_bounds__(kSvmThreads)
svm_decision_kernel(const T* __restrict__ vectors, const int* __restrict__ active, int n_active,
                    const double* __restrict__ alphas, const double* __restrict__ labels,
                    const T* __restrict__ queries, int d, int q_count, int rows_per_split, SvmParams p,
                    const SvmPlan* __restrict__ plan, double* __restrict__ partial) {
    using E = Elem<T>;
    using Acc = typename std::conditional<kKind == kSvmGaussian, typename E::Acc, T>::type;
    constexpr int kVecElems = kVec ? (int)(16 / sizeof(T)) : 1;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* q_s = reinterpret_cast<T*>(smem_raw);          // [feature][query]
    __shared__ double red[kQB][kSvmThreads / 32];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int q0 = blockIdx.x * kQB;
    const int r_begin = min(n_active, (int)blockIdx.y * rows_per_split);
    const int r_end = min(n_active, r_begin + rows_per_split);

    for (int e = tid; e < d * kQB; e += kSvmThreads) {
        const int qi = e % kQB, j = e / kQB;
        q_s[e] = (q0 + qi < q_count) ? queries[(size_t)(q0 + qi) * d + j] : (T)0;
    }
    __syncthreads();

    double sum[kQB];
#pragma unroll
    for (int qi = 0; qi < kQB; ++qi) sum[qi] = 0.0;

    for (int r = r_begin + tid; r < r_end; r += kSvmThreads) {
        const int row = active[r];
        const T* __restrict__ trow = vectors + (size_t)row * d;
        const double a = alphas[row], y = labels[row];
        double v[kQB];
        if constexpr (kPairwise) {
            // one pass over the row for all kQB queries; per query the 8 strided accumulators of numpy's leaf loop
            Acc stack[kQB][kSvmMaxDepth];
            int sp = 0;
            const int leaves = plan->leaves;
            for (int l = 0; l < leaves; ++l) {
                const int lo = plan->leaf[l].lo, n = plan->leaf[l].n;
                Acc res[kQB];
                if (n < 8) {
#pragma unroll
                    for (int qi = 0; qi < kQB; ++qi) res[qi] = E::sq_diff(trow[lo], q_s[(size_t)lo * kQB + qi]);
                    for (int i = 1; i < n; ++i) {
                        const T tv = trow[lo + i];
#pragma unroll
                        for (int qi = 0; qi < kQB; ++qi) res[qi] = E::add(res[qi], E::sq_diff(tv, q_s[(size_t)(lo + i) * kQB + qi]));
                    }
                } else {
                    Acc acc[kQB][8];
                    const int n8 = n - (n % 8);
                    T tv[8];
                    auto load8 = [&](int at) {     // leaves start at multiples of 8 elements: 16-byte aligned when the rows are
                        if (kVec) {
#pragma unroll
                            for (int w = 0; w < 8 / kVecElems; ++w)
                                *reinterpret_cast<int4*>(tv + w * kVecElems) =
                                    __ldg(reinterpret_cast<const int4*>(trow + at + w * kVecElems));
                        } else {
#pragma unroll
                            for (int j = 0; j < 8; ++j) tv[j] = trow[at + j];
                        }
                    };
                    load8(lo);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const T* qv = q_s + (size_t)(lo + j) * kQB;
#pragma unroll
                        for (int qi = 0; qi < kQB; ++qi) acc[qi][j] = E::sq_diff(tv[j], qv[qi]);
                    }
                    for (int i = 8; i < n8; i += 8) {
                        load8(lo + i);
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const T* qv = q_s + (size_t)(lo + i + j) * kQB;
#pragma unroll
                            for (int qi = 0; qi < kQB; ++qi) acc[qi][j] = E::add(acc[qi][j], E::sq_diff(tv[j], qv[qi]));
                        }
                    }
#pragma unroll
                    for (int qi = 0; qi < kQB; ++qi)
                        res[qi] = E::add(E::add(E::add(acc[qi][0], acc[qi][1]), E::add(acc[qi][2], acc[qi][3])),
                                         E::add(E::add(acc[qi][4], acc[qi][5]), E::add(acc[qi][6], acc[qi][7])));
                    for (int i = n8; i < n; ++i) {
                        const T tv = trow[lo + i];
#pragma unroll
                        for (int qi = 0; qi < kQB; ++qi) res[qi] = E::add(res[qi], E::sq_diff(tv, q_s[(size_t)(lo + i) * kQB + qi]));
                    }
                }
                if (leaves == 1) {
#pragma unroll
                    for (int qi = 0; qi < kQB; ++qi) stack[qi][0] = res[qi];
                } else {
#pragma unroll
                    for (int qi = 0; qi < kQB; ++qi) stack[qi][sp] = res[qi];
                    ++sp;
                    for (int m = plan->leaf[l].merges; m > 0; --m) {
                        --sp;
#pragma unroll
                        for (int qi = 0; qi < kQB; ++qi) stack[qi][sp - 1] = E::add(stack[qi][sp - 1], stack[qi][sp]);
                    }
                }
            }
#pragma unroll
            for (int qi = 0; qi < kQB; ++qi) v[qi] = SvmValue<T, kKind>::of(stack[qi][0], p);
        } else {
            Acc acc[kQB];
#pragma unroll
            for (int qi = 0; qi < kQB; ++qi) acc[qi] = (Acc)0;
            for (int j0 = 0; j0 < d; j0 += kVecElems) {
                T tv[kVecElems];
                if (kVec) {
                    *reinterpret_cast<int4*>(tv) = __ldg(reinterpret_cast<const int4*>(trow + j0));
                } else {
                    tv[0] = trow[j0];
                }
#pragma unroll
                for (int jj = 0; jj < kVecElems; ++jj) {
                    const T* qv = q_s + (size_t)(j0 + jj) * kQB;
#pragma unroll
                    for (int qi = 0; qi < kQB; ++qi) {
                        if (kKind == kSvmGaussian) acc[qi] = (Acc)E::add((typename E::Acc)acc[qi], E::sq_diff(tv[jj], qv[qi]));
                        else acc[qi] = (Acc)SvmDot<T>::step((T)acc[qi], tv[jj], qv[qi]);
                    }
                }
            }
#pragma unroll
            for (int qi = 0; qi < kQB; ++qi) v[qi] = SvmValue<T, kKind>::of(acc[qi], p);
        }
#pragma unroll
        for (int qi = 0; qi < kQB; ++qi) sum[qi] = __dadd_rn(sum[qi], __dmul_rn(__dmul_rn(v[qi], a), y));
    }

    // threads -> warp (shuffle tree) -> CTA (one thread per query adds the 8 warp sums in order): fixed order
#pragma unroll
    for (int qi = 0; qi < kQB; ++qi) {
        double s = sum[qi];
        for (int off = 16; off > 0; off >>= 1) s = __dadd_rn(s, __shfl_xor_sync(0xffffffffu, s, off));
        if (lane == 0) red[qi][warp] = s;
    }
    __syncthreads();
    if (tid < kQB && q0 + tid < q_count) {
        double s = red[tid][0];
        for (int w = 1; w < kSvmThreads / 32; ++w) s = __dadd_rn(s, red[tid][w]);
        partial[(size_t)blockIdx.y * q_count + q0 + tid] = s;
    }
}

__global__ void svm_finish_kernel(const double* __restrict__ partial, int nsplit, int q_count, double intercept,
                                  double* __restrict__ out) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= q_count) return;
    double s = partial[q];
    for (int sp = 1; sp < nsplit; ++sp) s = __dadd_rn(s, partial[(size_t)sp * q_count + q]);
    out[q] = __dadd_rn(intercept, s);          // predictions[i] = float(intercept); predictions[i] += np.sum(...)
}

inline int svm_max_splits(int q_count, int n_active, int sm_count, int qb) {
    const int qtiles = (q_count + qb - 1) / qb;
    // enough CTAs for two waves when the query count alone does not give them; a split is at least one batch of rows
    int nsplit = 1;
    if (qtiles < 2 * sm_count) nsplit = std::min((2 * sm_count + qtiles - 1) / qtiles, (n_active + kSvmThreads - 1) / kSvmThreads);
    return std::max(1, std::min(nsplit, 65535));
}

// The row splits a launch may use whatever query group (4 or 8 queries per CTA) its instance picks: fewer queries per
// CTA -> more query tiles -> fewer splits, so the larger group bounds both.
inline int svm_partial_splits(int q_count, int n_active, int sm_count) {
    return std::max(svm_max_splits(q_count, n_active, sm_count, 4), svm_max_splits(q_count, n_active, sm_count, 8));
}

// partial must hold svm_partial_splits(q_count, n_active, sm_count) * q_count doubles
template <typename T, int kKind, bool kPairwise, int kQB>
cudaError_t launch_svm_decision_qb(const void* vectors, const int* active, int n_active, const double* alphas,
                                   const double* labels, const void* queries, int d, int q_count, const SvmParams& p,
                                   int sm_count, const SvmPlan* plan, double* partial, int* nsplit_out,
                                   cudaStream_t stream) {
    const size_t smem = ((size_t)d * kQB * sizeof(T) + 15) / 16 * 16 + 64;
    const int qtiles = (q_count + kQB - 1) / kQB;
    const int nsplit = svm_max_splits(q_count, n_active, sm_count, kQB);
    const int rows_per_split = (n_active + nsplit - 1) / nsplit;
    *nsplit_out = nsplit;
    const bool vec = ((size_t)d * sizeof(T)) % 16 == 0 && ((uintptr_t)vectors % 16) == 0;
    auto kernel = vec ? svm_decision_kernel<T, kKind, kPairwise, true, kQB> : svm_decision_kernel<T, kKind, kPairwise, false, kQB>;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kernel<<<dim3((unsigned)qtiles, (unsigned)nsplit), kSvmThreads, smem, stream>>>(
        static_cast<const T*>(vectors), active, n_active, alphas, labels, static_cast<const T*>(queries), d, q_count,
        rows_per_split, p, plan, partial);
    return cudaGetLastError();
}

template <typename T, int kKind>
cudaError_t launch_svm_decision_kind(const void* vectors, const int* active, int n_active, const double* alphas,
                                     const double* labels, const void* queries, int d, int q_count, bool pairwise,
                                     const SvmParams& p, int sm_count, const SvmPlan* plan, double* partial,
                                     int* nsplit_out, cudaStream_t stream) {
    if constexpr (kKind == kSvmGaussian && Elem<T>::kIsFloat) {
        if (pairwise)
            return launch_svm_decision_qb<T, kKind, true, sizeof(T) == 8 ? 4 : 8>(
                vectors, active, n_active, alphas, labels, queries, d, q_count, p, sm_count, plan, partial, nsplit_out,
                stream);
    }
    return launch_svm_decision_qb<T, kKind, false, kSvmQB>(vectors, active, n_active, alphas, labels, queries, d,
                                                           q_count, p, sm_count, plan, partial, nsplit_out, stream);
}

// K(V_i, V_j) for all pairs of the machine's own vectors (float64 out, values computed in the data's dtype): the scalar
// kernel values SMO's step needs (osml/predict.py:1140-1142, evaluated there on (1, d)-shaped rows -- a single row is
// contiguous along the features, so numpy adds the squared differences with its pairwise tree whatever the matrix order).
template <typename T, int kKind>
__global__ void svm_gram_kernel(const T* __restrict__ vectors, int n, int d, SvmParams p, double* __restrict__ out) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
    if (j >= n) return;
    const T* a = vectors + (size_t)i * d;
    const T* b = vectors + (size_t)j * d;
    double v;
    if constexpr (kKind == kSvmGaussian) {
        typename Elem<T>::Acc acc;
        if (Elem<T>::kIsFloat) acc = sum_pairwise<T>(a, b, d);
        else acc = sum_sequential<T>(a, b, d);
        v = SvmValue<T, kKind>::of(acc, p);
    } else {
        T acc = (T)0;
        for (int f = 0; f < d; ++f) acc = SvmDot<T>::step(acc, a[f], b[f]);
        v = SvmValue<T, kKind>::of(acc, p);
    }
    out[(size_t)i * n + j] = v;
}

template <typename T>
cudaError_t launch_svm_gram(int kind, const void* vectors, int n, int d, const SvmParams& p, double* out,
                            cudaStream_t stream) {
    const dim3 grid((unsigned)((n + 127) / 128), (unsigned)n);
    const T* v = static_cast<const T*>(vectors);
    if (kind == kSvmGaussian) svm_gram_kernel<T, kSvmGaussian><<<grid, 128, 0, stream>>>(v, n, d, p, out);
    else if constexpr (Elem<T>::kIsFloat) {
        if (kind == kSvmLinear) svm_gram_kernel<T, kSvmLinear><<<grid, 128, 0, stream>>>(v, n, d, p, out);
        else svm_gram_kernel<T, kSvmPolynomial><<<grid, 128, 0, stream>>>(v, n, d, p, out);
    } else return cudaErrorInvalidValue;
    return cudaGetLastError();
}

template <typename T>
cudaError_t launch_svm_decision(int kind, const void* vectors, const int* active, int n_active, const double* alphas,
                                const double* labels, const void* queries, int d, int q_count, bool pairwise,
                                const SvmParams& p, int sm_count, const SvmPlan* plan, double* partial, int* nsplit,
                                cudaStream_t stream) {
    if (kind == kSvmGaussian)
        return launch_svm_decision_kind<T, kSvmGaussian>(vectors, active, n_active, alphas, labels, queries, d, q_count,
                                                         pairwise && Elem<T>::kIsFloat && d > 1, p, sm_count, plan,
                                                         partial, nsplit, stream);
    if constexpr (Elem<T>::kIsFloat) {
        if (kind == kSvmLinear)
            return launch_svm_decision_kind<T, kSvmLinear>(vectors, active, n_active, alphas, labels, queries, d,
                                                           q_count, false, p, sm_count, plan, partial, nsplit, stream);
        return launch_svm_decision_kind<T, kSvmPolynomial>(vectors, active, n_active, alphas, labels, queries, d,
                                                           q_count, false, p, sm_count, plan, partial, nsplit, stream);
    }
    return cudaErrorInvalidValue;
}

}  // namespace osml
