// K2: exact direct-difference distance kernel with a fused per-query top-(k+1) selection.
//
// Replaces predict.py:742 (_distances) + predict.py:750 (argpartition) for every dtype the reference accepts, in the
// reference's own arithmetic (Elem<T>) and summation order, so distances are bit-identical to numpy's and the
// neighbour order is the documented (distance, index) rule.  The N x Q distance matrix is never written: each
// thread owns one query, keeps its running list in shared memory and only touches it for the rare candidate that
// beats its current (k+1)-th best.
//
// Tiling (sequential order): a CTA holds QT=128 queries (one per thread) and walks the train rows in tiles of TN rows;
// features are staged through shared memory DCH at a time (train chunk row-major, read as warp-wide broadcasts; query
// chunk transposed so thread t reads column t conflict-free).  Each thread carries TN accumulators in registers, so
// one broadcast LDS feeds 3 arithmetic instructions per (row, feature) -- the minimum the un-fused
// subtract/multiply/add of the oracle allows.
#pragma once
#include "knn_common.cuh"

namespace osml {

constexpr int kDirectQT = 128;   // queries (= threads) per CTA

template <typename Dist>
struct TopList {
    // shared-memory list of one thread: slot e of thread t lives at d[e * kDirectQT + t]
    Dist* d; int* i; int ksel;
    __device__ __forceinline__ void init(Dist* dbase, int* ibase, int ksel_, int tid) {
        d = dbase + tid; i = ibase + tid; ksel = ksel_;
        for (int e = 0; e < ksel; ++e) { d[e * kDirectQT] = nan_of<Dist>(); i[e * kDirectQT] = kIdxSentinel; }
    }
    __device__ __forceinline__ Dist worst_d() const { return d[(ksel - 1) * kDirectQT]; }
    __device__ __forceinline__ int worst_i() const { return i[(ksel - 1) * kDirectQT]; }
    // precondition: key_less(dc, ic, worst)
    __device__ __forceinline__ void insert(Dist dc, int ic) {
        int pos = ksel - 1;
        while (pos > 0 && key_less(dc, ic, d[(pos - 1) * kDirectQT], i[(pos - 1) * kDirectQT])) {
            d[pos * kDirectQT] = d[(pos - 1) * kDirectQT];
            i[pos * kDirectQT] = i[(pos - 1) * kDirectQT];
            --pos;
        }
        d[pos * kDirectQT] = dc; i[pos * kDirectQT] = ic;
    }
};

template <typename T, int TN, int DCH>
struct DirectSmem {
    static constexpr int kQPitch = kDirectQT + (sizeof(T) >= 4 ? 1 : 4 / sizeof(T));
    static size_t bytes(int ksel) {
        size_t b = 0;
        b += (size_t)ksel * kDirectQT * sizeof(typename Elem<T>::Dist);
        b += (size_t)ksel * kDirectQT * sizeof(int);
        b += (size_t)TN * DCH * sizeof(T);
        b += (size_t)DCH * kQPitch * sizeof(T);
        return b + 16;
    }
};

// sum_order == sequential (or integer data): register-tiled kernel
template <typename T, int TN, int DCH>
__global__ void __launch_bounds__(kDirectQT)
knn_direct_seq_kernel(const T* __restrict__ train, const T* __restrict__ queries, int n, int d, int q_count, int ksel,
                      int rows_per_split, int64_t index_base,
                      typename Elem<T>::Dist* __restrict__ part_d, int64_t* __restrict__ part_i) {
    using E = Elem<T>;
    using Acc = typename E::Acc;
    using Dist = typename E::Dist;
    constexpr int kQPitch = DirectSmem<T, TN, DCH>::kQPitch;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Dist* list_d = reinterpret_cast<Dist*>(smem_raw);
    int* list_i = reinterpret_cast<int*>(list_d + (size_t)ksel * kDirectQT);
    T* t_s = reinterpret_cast<T*>(reinterpret_cast<uintptr_t>(list_i + (size_t)ksel * kDirectQT + 3) & ~(uintptr_t)15);
    T* q_s = t_s + TN * DCH;

    const int tid = threadIdx.x;
    const int q0 = blockIdx.x * kDirectQT;
    const int split = blockIdx.y, nsplit = gridDim.y;
    const int n_begin = min(n, split * rows_per_split);
    const int n_end = min(n, n_begin + rows_per_split);
    const bool q_valid = q0 + tid < q_count;
    const bool single_chunk = d <= DCH;

    TopList<Dist> list;
    list.init(list_d, list_i, ksel, tid);
    Dist wd = nan_of<Dist>();
    int wi = kIdxSentinel;
    Acc thr = E::threshold(wd);

    for (int row0 = n_begin; row0 < n_end; row0 += TN) {
        Acc acc[TN];
#pragma unroll
        for (int r = 0; r < TN; ++r) acc[r] = (Acc)0;
        for (int j0 = 0; j0 < d; j0 += DCH) {
            __syncthreads();
            // train chunk: TN x DCH, zero padded
            for (int e = tid; e < TN * DCH; e += kDirectQT) {
                const int r = e / DCH, jj = e % DCH;
                const int row = row0 + r, j = j0 + jj;
                T v = (T)0;
                if (row < n_end && j < d) v = train[(size_t)row * d + j];
                t_s[e] = v;
            }
            if (!single_chunk || row0 == n_begin) {
                for (int e = tid; e < kDirectQT * DCH; e += kDirectQT) {
                    const int qi = e / DCH, jj = e % DCH;
                    const int qq = q0 + qi, j = j0 + jj;
                    T v = (T)0;
                    if (qq < q_count && j < d) v = queries[(size_t)qq * d + j];
                    q_s[jj * kQPitch + qi] = v;
                }
            }
            __syncthreads();
            T qreg[DCH];
#pragma unroll
            for (int jj = 0; jj < DCH; ++jj) qreg[jj] = q_s[jj * kQPitch + tid];
#pragma unroll
            for (int r = 0; r < TN; ++r) {
#pragma unroll
                for (int jj = 0; jj < DCH; ++jj) acc[r] = E::add(acc[r], E::sq_diff(t_s[r * DCH + jj], qreg[jj]));
            }
        }
        if (q_valid) {
#pragma unroll
            for (int r = 0; r < TN; ++r) {
                const int row = row0 + r;
                if (row < n_end && E::may_beat(acc[r], thr)) {
                    const Dist dc = E::finish(acc[r]);
                    if (key_less(dc, row, wd, wi)) {
                        list.insert(dc, row);
                        wd = list.worst_d(); wi = list.worst_i();
                        thr = E::threshold(wd);
                    }
                }
            }
        }
    }
    if (q_valid) {
        const size_t base = ((size_t)(q0 + tid) * nsplit + split) * ksel;
        for (int e = 0; e < ksel; ++e) {
            const int li = list.i[e * kDirectQT];
            part_d[base + e] = list.d[e * kDirectQT];
            part_i[base + e] = li == kIdxSentinel ? kIdxSentinel64 : index_base + li;
        }
    }
}

// any summation order, no tiling: one thread per query, rows streamed from L2.  Only used for the pairwise (C-ordered)
// float case, which pandas >= 3 never produces (DataFrame.values is F-ordered there).
template <typename T>
__global__ void __launch_bounds__(kDirectQT)
knn_direct_generic_kernel(const T* __restrict__ train, const T* __restrict__ queries, int n, int d, int q_count,
                          int ksel, int rows_per_split, int64_t index_base, int sum_order,
                          typename Elem<T>::Dist* __restrict__ part_d, int64_t* __restrict__ part_i) {
    using E = Elem<T>;
    using Dist = typename E::Dist;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Dist* list_d = reinterpret_cast<Dist*>(smem_raw);
    int* list_i = reinterpret_cast<int*>(list_d + (size_t)ksel * kDirectQT);
    const int tid = threadIdx.x;
    const int qq = blockIdx.x * kDirectQT + tid;
    const int split = blockIdx.y, nsplit = gridDim.y;
    const int n_begin = min(n, split * rows_per_split);
    const int n_end = min(n, n_begin + rows_per_split);
    TopList<Dist> list;
    list.init(list_d, list_i, ksel, tid);
    if (qq >= q_count) return;
    Dist wd = nan_of<Dist>();
    int wi = kIdxSentinel;
    const T* qrow = queries + (size_t)qq * d;
    for (int row = n_begin; row < n_end; ++row) {
        const Dist dc = exact_distance<T>(train + (size_t)row * d, qrow, d, sum_order);
        if (key_less(dc, row, wd, wi)) {
            list.insert(dc, row);
            wd = list.worst_d(); wi = list.worst_i();
        }
    }
    const size_t base = ((size_t)qq * nsplit + split) * ksel;
    for (int e = 0; e < ksel; ++e) {
        const int li = list.i[e * kDirectQT];
        part_d[base + e] = list.d[e * kDirectQT];
        part_i[base + e] = li == kIdxSentinel ? kIdxSentinel64 : index_base + li;
    }
}

// Merges `parts` ascending lists per query into one ascending list of ksel entries by (distance, index).
// element (p, q, e) lives at  p * part_stride + q * q_stride + e.  One thread per query.
template <typename Dist>
__global__ void knn_merge_lists_kernel(const Dist* __restrict__ in_d, const int64_t* __restrict__ in_i, int parts,
                                       int64_t part_stride, int64_t q_stride, int64_t q_count, int ksel,
                                       Dist* __restrict__ out_d, int64_t* __restrict__ out_i) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= q_count) return;
    constexpr int kMaxParts = 64;
    unsigned char head[kMaxParts];
    for (int p = 0; p < parts; ++p) head[p] = 0;
    const int64_t qbase = q * q_stride;
    for (int e = 0; e < ksel; ++e) {
        int best = -1;
        Dist bd = nan_of<Dist>();
        int64_t bi = kIdxSentinel64;
        for (int p = 0; p < parts; ++p) {
            if (head[p] >= ksel) continue;
            const int64_t at = p * part_stride + qbase + head[p];
            const Dist cd = in_d[at];
            const int64_t ci = in_i[at];
            if (best < 0 || key_less(cd, ci, bd, bi)) { best = p; bd = cd; bi = ci; }
        }
        if (best >= 0) ++head[best];
        out_d[q * ksel + e] = bd;
        out_i[q * ksel + e] = bi;
    }
}

}  // namespace osml
