// K2: exact direct-difference distance kernels with a fused per-query top-(k+1) selection.
//
// Replace predict.py:742 (_distances) + predict.py:750 (argpartition) for every dtype the reference accepts, in the
// reference's own arithmetic (Elem<T>) and summation order, so distances are bit-identical to numpy's and the
// neighbour order is the documented (distance, index) rule.  The N x Q distance matrix is never written.
//   knn_direct_rows_kernel     sequential summation order (F-ordered DataFrame.values, and all integer dtypes)
//   knn_direct_generic_kernel  numpy's pairwise order (C-ordered float matrices; pandas >= 3 never produces them)
//   knn_merge_lists_kernel     exact merge of partial lists (row splits here, GPU shards in the train-sharded mode)
#pragma once
#include "knn_common.cuh"

namespace osml {

constexpr int kDirectQT = 128;   // queries (= threads) per CTA

template <typename Dist>
struct TopList {
    // shared-memory list of one thread: slot e of thread t lives at d[e * stride + t] (stride = threads per CTA)
    Dist* d; int* i; int ksel; int stride;
    __device__ __forceinline__ void init(Dist* dbase, int* ibase, int ksel_, int tid, int stride_) {
        d = dbase + tid; i = ibase + tid; ksel = ksel_; stride = stride_;
        for (int e = 0; e < ksel; ++e) { d[e * stride] = nan_of<Dist>(); i[e * stride] = kIdxSentinel; }
    }
    __device__ __forceinline__ Dist worst_d() const { return d[(ksel - 1) * stride]; }
    __device__ __forceinline__ int worst_i() const { return i[(ksel - 1) * stride]; }
    // precondition: key_less(dc, ic, worst)
    __device__ __forceinline__ void insert(Dist dc, int ic) {
        int pos = ksel - 1;
        while (pos > 0 && key_less(dc, ic, d[(pos - 1) * stride], i[(pos - 1) * stride])) {
            d[pos * stride] = d[(pos - 1) * stride];
            i[pos * stride] = i[(pos - 1) * stride];
            --pos;
        }
        d[pos * stride] = dc; i[pos * stride] = ic;
    }
};

// ---------------------------------------------------------------------------------------------------------------------
// Row-parallel exact kernel (sequential summation order, every dtype).
//
// A CTA of 256 threads serves QB = 8 queries and a contiguous range of train rows.  Thread t owns one train row per
// batch of 256 rows and accumulates its distance to all 8 queries at once: per feature one (vectorised) load of its own
// row element, the 8 query values as broadcast shared-memory reads, and 8 x (subtract, multiply, add) -- the three
// un-fusable instructions per element that numpy parity dictates.  Train rows are read straight from global memory
// (each thread walks its own row; consecutive threads take consecutive rows, so a warp streams one contiguous 32-row
// block through L1).
//
// Selection: a distance that beats the query's current (k+1)-th best is appended to that query's shared-memory buffer
// (one atomicAdd); after each batch warp w folds buffer w into the sorted list of query w and refreshes its threshold.
// Partial lists per (query, row split) are merged by knn_merge_lists_kernel.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kRowQB = 8;          // queries per CTA (= warps per CTA)
constexpr int kRowThreads = 256;

template <typename T>
struct RowSmem {
    using Dist = typename Elem<T>::Dist;
    static size_t bytes(int d, int ksel) {
        size_t b = 0;
        b += ((size_t)d * kRowQB * sizeof(T) + 15) / 16 * 16;                 // query tile, [feature][query]
        b += (size_t)kRowQB * kRowThreads * (sizeof(Dist) + sizeof(int));     // candidate buffers
        b += (size_t)kRowQB * ksel * (sizeof(Dist) + sizeof(int));            // sorted lists
        return b + 64;
    }
};

template <typename T, bool kVec>
__global__ void __launch_bounds__(kRowThreads)
knn_direct_rows_kernel(const T* __restrict__ train, const T* __restrict__ queries, int n, int d, int d_chunk, int q_count,
                       int ksel, int rows_per_split, int64_t index_base, typename Elem<T>::Dist* __restrict__ part_d,
                       int64_t* __restrict__ part_i) {
    using E = Elem<T>;
    using Acc = typename E::Acc;
    using Dist = typename E::Dist;
    constexpr int kVecElems = kVec ? (int)(16 / sizeof(T)) : 1;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* q_s = reinterpret_cast<T*>(smem_raw);                                              // [d_chunk][kRowQB]
    const size_t q_bytes = ((size_t)d_chunk * kRowQB * sizeof(T) + 15) / 16 * 16;
    Dist* buf_d = reinterpret_cast<Dist*>(smem_raw + q_bytes);                            // [kRowQB][kRowThreads]
    Dist* list_d = buf_d + kRowQB * kRowThreads;                                          // [kRowQB][ksel]
    int* buf_i = reinterpret_cast<int*>(list_d + kRowQB * ksel);                          // [kRowQB][kRowThreads]
    int* list_i = buf_i + kRowQB * kRowThreads;                                           // [kRowQB][ksel]
    __shared__ int s_cnt[kRowQB];
    __shared__ Acc s_thr[kRowQB];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int q0 = blockIdx.x * kRowQB;
    const int split = blockIdx.y, nsplit = gridDim.y;
    const int n_begin = min(n, split * rows_per_split);
    const int n_end = min(n, n_begin + rows_per_split);

    // The query tile holds d_chunk features of the 8 queries: the whole row (d_chunk == d, loaded once) or, for rows
    // too wide for shared memory, one chunk at a time (reloaded per batch of rows; every thread keeps its 8 running
    // sums across the chunks, so the summation order is unchanged).
    const bool chunked = d_chunk < d;
    auto load_queries = [&](int c0, int cw) {
        for (int e = tid; e < cw * kRowQB; e += kRowThreads) {
            const int j = e / kRowQB, qi = e % kRowQB;
            q_s[e] = (q0 + qi < q_count) ? queries[(size_t)(q0 + qi) * d + c0 + j] : (T)0;
        }
    };
    if (!chunked) load_queries(0, d);
    for (int e = lane; e < ksel; e += 32) { list_d[warp * ksel + e] = nan_of<Dist>(); list_i[warp * ksel + e] = kIdxSentinel; }
    if (tid < kRowQB) { s_cnt[tid] = 0; s_thr[tid] = E::threshold(nan_of<Dist>()); }
    __syncthreads();

    for (int row0 = n_begin; row0 < n_end; row0 += kRowThreads) {
        const int row = row0 + tid;
        Acc acc[kRowQB];
#pragma unroll
        for (int qi = 0; qi < kRowQB; ++qi) acc[qi] = (Acc)0;
        for (int c0 = 0; c0 < d; c0 += d_chunk) {
            const int cw = min(d_chunk, d - c0);
            if (chunked) {
                __syncthreads();                       // everyone is done with the previous chunk
                load_queries(c0, cw);
                __syncthreads();
            }
            if (row < n_end) {
                const T* __restrict__ trow = train + (size_t)row * d + c0;
                for (int j0 = 0; j0 < cw; j0 += kVecElems) {
                    T tv[kVecElems];
                    if (kVec) {
                        const int4 raw = __ldg(reinterpret_cast<const int4*>(trow + j0));
                        *reinterpret_cast<int4*>(tv) = raw;
                    } else {
                        tv[0] = trow[j0];
                    }
#pragma unroll
                    for (int jj = 0; jj < kVecElems; ++jj) {
                        const T* qv = q_s + (size_t)(j0 + jj) * kRowQB;
#pragma unroll
                        for (int qi = 0; qi < kRowQB; ++qi) acc[qi] = E::add(acc[qi], E::sq_diff(tv[jj], qv[qi]));
                    }
                }
            }
        }
        if (row < n_end) {
#pragma unroll
            for (int qi = 0; qi < kRowQB; ++qi) {
                if (E::may_beat(acc[qi], s_thr[qi])) {
                    const int slot = atomicAdd(&s_cnt[qi], 1);
                    buf_d[qi * kRowThreads + slot] = E::finish(acc[qi]);
                    buf_i[qi * kRowThreads + slot] = row;
                }
            }
        }
        __syncthreads();
        // warp w folds the candidates of query w into its sorted list (lane 0; after warm-up the buffer is almost empty)
        const int cnt = s_cnt[warp];
        if (cnt > 0 && lane == 0) {
            Dist* ld = list_d + warp * ksel;
            int* li = list_i + warp * ksel;
            for (int c = 0; c < cnt; ++c) {
                const Dist dc = buf_d[warp * kRowThreads + c];
                const int ic = buf_i[warp * kRowThreads + c];
                if (!key_less(dc, ic, ld[ksel - 1], li[ksel - 1])) continue;
                int pos = ksel - 1;
                while (pos > 0 && key_less(dc, ic, ld[pos - 1], li[pos - 1])) {
                    ld[pos] = ld[pos - 1]; li[pos] = li[pos - 1];
                    --pos;
                }
                ld[pos] = dc; li[pos] = ic;
            }
            s_thr[warp] = E::threshold(ld[ksel - 1]);
            s_cnt[warp] = 0;
        }
        __syncthreads();
    }
    if (q0 + warp < q_count) {
        const size_t base = ((size_t)(q0 + warp) * nsplit + split) * ksel;
        for (int e = lane; e < ksel; e += 32) {
            const int li = list_i[warp * ksel + e];
            part_d[base + e] = list_d[warp * ksel + e];
            part_i[base + e] = li == kIdxSentinel ? kIdxSentinel64 : index_base + li;
        }
    }
}

// any summation order, no tiling: one thread per query, rows streamed from L2.  Only used for the pairwise (C-ordered)
// float case, which pandas >= 3 never produces (DataFrame.values is F-ordered there).  blockDim.x = queries per CTA:
// kDirectQT, fewer when k + 1 entries per thread would not fit in shared memory otherwise.
template <typename T>
__global__ void __launch_bounds__(kDirectQT)
knn_direct_generic_kernel(const T* __restrict__ train, const T* __restrict__ queries, int n, int d, int q_count,
                          int ksel, int rows_per_split, int64_t index_base, int sum_order,
                          typename Elem<T>::Dist* __restrict__ part_d, int64_t* __restrict__ part_i) {
    using E = Elem<T>;
    using Dist = typename E::Dist;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Dist* list_d = reinterpret_cast<Dist*>(smem_raw);
    const int qt = blockDim.x;
    int* list_i = reinterpret_cast<int*>(list_d + (size_t)ksel * qt);
    const int tid = threadIdx.x;
    const int qq = blockIdx.x * qt + tid;
    const int split = blockIdx.y, nsplit = gridDim.y;
    const int n_begin = min(n, split * rows_per_split);
    const int n_end = min(n, n_begin + rows_per_split);
    TopList<Dist> list;
    list.init(list_d, list_i, ksel, tid, qt);
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
        const int li = list.i[e * qt];
        part_d[base + e] = list.d[e * qt];
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
    unsigned short head[kMaxParts];
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

// ---------------------------------------------------------------------------------------------------------------------
// Train-sharded exchange: every GPU holds exact partial lists for ALL queries over its own rows,
// but only has to merge the queries of its own slice.  The lists travel as 16-byte records {int64 index, distance bits}
// (a float32 distance in the low word), laid out [destination rank][query of its slice][entry], so that ONE
// all-to-all with equal splits hands every rank the G partial lists of its ceil(Q / G) queries.
// ---------------------------------------------------------------------------------------------------------------------
struct alignas(16) NeighbourRecord { int64_t index; unsigned long long dist_bits; };

template <typename Dist> __device__ __forceinline__ unsigned long long dist_to_bits(Dist d);
template <> __device__ __forceinline__ unsigned long long dist_to_bits<float>(float d) { return (unsigned long long)__float_as_uint(d); }
template <> __device__ __forceinline__ unsigned long long dist_to_bits<double>(double d) { return (unsigned long long)__double_as_longlong(d); }
template <typename Dist> __device__ __forceinline__ Dist dist_from_bits(unsigned long long b);
template <> __device__ __forceinline__ float dist_from_bits<float>(unsigned long long b) { return __uint_as_float((uint32_t)b); }
template <> __device__ __forceinline__ double dist_from_bits<double>(unsigned long long b) { return __longlong_as_double((long long)b); }

// lists [q][ksel] -> records [parts][slice_q][ksel]; queries past q (padding of the last slice) get empty entries
template <typename Dist>
__global__ void knn_pack_records_kernel(const Dist* __restrict__ in_d, const int64_t* __restrict__ in_i, int64_t q_count,
                                        int64_t slice_q, int parts, int ksel, NeighbourRecord* __restrict__ out) {
    const int64_t total = (int64_t)parts * slice_q * ksel;
    for (int64_t at = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; at < total; at += (int64_t)gridDim.x * blockDim.x) {
        const int64_t q = at / ksel;                  // [parts][slice_q] flattened == the query number
        NeighbourRecord r;
        if (q < q_count) { r.index = in_i[at]; r.dist_bits = dist_to_bits<Dist>(in_d[at]); }
        else { r.index = kIdxSentinel64; r.dist_bits = dist_to_bits<Dist>(nan_of<Dist>()); }
        out[at] = r;
    }
}

// records [parts][slice_q][ksel] (part p = the list rank p computed over its rows) -> merged ascending lists of the
// first q_count queries of the slice, by (distance, index).  One thread per query.
template <typename Dist>
__global__ void knn_merge_records_kernel(const NeighbourRecord* __restrict__ in, int parts, int64_t slice_q,
                                         int64_t q_count, int ksel, Dist* __restrict__ out_d,
                                         int64_t* __restrict__ out_i) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= q_count) return;
    constexpr int kMaxParts = 64;
    unsigned short head[kMaxParts];
    for (int p = 0; p < parts; ++p) head[p] = 0;
    for (int e = 0; e < ksel; ++e) {
        int best = -1;
        Dist bd = nan_of<Dist>();
        int64_t bi = kIdxSentinel64;
        for (int p = 0; p < parts; ++p) {
            if (head[p] >= ksel) continue;
            const NeighbourRecord r = in[((int64_t)p * slice_q + q) * ksel + head[p]];
            const Dist cd = dist_from_bits<Dist>(r.dist_bits);
            if (best < 0 || key_less(cd, r.index, bd, bi)) { best = p; bd = cd; bi = r.index; }
        }
        if (best >= 0) ++head[best];
        out_d[q * ksel + e] = bd;
        out_i[q * ksel + e] = bi;
    }
}

}  // namespace osml
