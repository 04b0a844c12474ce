// K4: RBF-weighted neighbour aggregation, one warp per query.
//
// Replaces predict.py:744-746 (_weight), 753-765 (the weight loop and the per-label dict) and the reductions at
// predict.py:788-793 (regression: sum(label * W) / sum(W)) and 809-810 (classification: stable arg-max of W).
//
// Arithmetic follows the reference exactly where it matters:
//  * pandas boxes the distances to Python floats on iteration (predict.py:755), so ratio, square and exponent are
//    fp64 for every data dtype;
//  * insertion order is farthest first: list entry k (the largest by (distance, index)) is the normaliser and does
//    not vote, entries k-1 ... 0 vote in that order;
//  * equal labels merge in first-appearance order and their weights add left to right; the final sums run over the
//    distinct labels in first-appearance order; an exact weight tie goes to the label that appeared first.
// Deviations (documented in DESIGN.md): r**2 is the correctly rounded r*r (libm pow(r, 2.0) differs in the last bit
// for ~0.08 % of inputs) and exp() is CUDA's (<= 1 ulp from glibc's).
#pragma once
#include "knn_common.cuh"

namespace osml {

constexpr int kAggMaxK = 1025;         // k + 1 entries per query the GPU path supports (k <= 1024)
constexpr int kAggWarpsPerCta = 4;     // queries (= warps) per CTA while their voters fit in 48 KB of shared memory
constexpr int kAggBytesPerVoter = 8 + 8 + 8 + 4;

inline int agg_warps_per_cta(int k) { return k <= 384 ? kAggWarpsPerCta : (k <= 768 ? 2 : 1); }
inline size_t agg_smem_bytes(int k, int warps) { return (size_t)warps * ((size_t)k * kAggBytesPerVoter + 16); }

enum LabelKind { kLabelSigned = 0, kLabelUnsigned = 1, kLabelFloat = 2 };

__device__ __forceinline__ double label_value(int64_t bits, int kind) {
    if (kind == kLabelFloat) return __longlong_as_double(bits);
    if (kind == kLabelUnsigned) return __ull2double_rn((unsigned long long)bits);
    return __ll2double_rn(bits);
}

__device__ __forceinline__ bool label_equal(int64_t a, int64_t b, int kind) {
    if (kind == kLabelFloat) return __longlong_as_double(a) == __longlong_as_double(b);   // -0 == +0, NaN != NaN
    return a == b;
}

// label_is_f32: float32 regression labels.  The reference's `key * value` is then np.float32 * Python float, which numpy 2
// (NEP 50) evaluates in float32 -- the weight is rounded to float32 first -- `weighted_labels_sum` accumulates in float32
// and the final division by the (fp64) weight sum rounds the divisor to float32 too (predict.py:788-793).
template <typename Dist>
__global__ void __launch_bounds__(32 * kAggWarpsPerCta)
knn_aggregate_kernel(const Dist* __restrict__ nn_d, const int64_t* __restrict__ nn_i, int64_t q_count, int k,
                     const int64_t* __restrict__ label_bits, int label_kind, int label_is_f32, int64_t n_labels,
                     double gamma, int gamma_is_f32, double epsilon, int task, double* __restrict__ out_pred) {
    extern __shared__ __align__(16) unsigned char agg_smem[];      // per warp: [k] weights, sums, labels, leaders
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int warps = blockDim.x >> 5;
    const int64_t q = (int64_t)blockIdx.x * warps + warp;
    if (q >= q_count) return;
    const int ksel = k + 1;
    const Dist* d = nn_d + q * ksel;
    const int64_t* idx = nn_i + q * ksel;
    unsigned char* mine = agg_smem + (size_t)warp * (((size_t)k * kAggBytesPerVoter + 15) / 16 * 16);
    double* w = reinterpret_cast<double*>(mine);
    double* gsum = w + k;
    int64_t* lab = reinterpret_cast<int64_t*>(gsum + k);
    int* leader = reinterpret_cast<int*>(lab + k);

    const double dn = (double)d[k];
    const double denom = dn + epsilon;                    // predict.py:759  distance_to_normalize_by + self.epsilon
    for (int v = lane; v < k; v += 32) {                  // voter v in insertion order <-> ascending entry k-1-v
        const int e = k - 1 - v;
        const double r = (double)d[e] / denom;
        const double sq = __dmul_rn(r, r);                         // predict.py:746  distance ** 2
        double x;
        if (gamma_is_f32) x = (double)__fmul_rn((float)(-gamma), (float)sq);
        else x = __dmul_rn(-gamma, sq);
        w[v] = exp(x);
        const int64_t ii = idx[e];
        lab[v] = (ii >= 0 && ii < n_labels) ? label_bits[ii] : 0;
    }
    __syncwarp();
    for (int v = lane; v < k; v += 32) {                  // first voter carrying the same label
        int ld = v;
        for (int u = 0; u < v; ++u) if (label_equal(lab[u], lab[v], label_kind)) { ld = u; break; }
        leader[v] = ld;
    }
    __syncwarp();
    for (int v = lane; v < k; v += 32) {                  // per-label weight, added in insertion order
        if (leader[v] == v) {
            double acc = w[v];
            for (int u = v + 1; u < k; ++u) if (leader[u] == v) acc = __dadd_rn(acc, w[u]);
            gsum[v] = acc;
        }
    }
    __syncwarp();
    if (lane == 0) {
        double result;
        if (task == 0 && label_is_f32) {
            float wl = 0.0f;
            double ws = 0.0;
            for (int v = 0; v < k; ++v) if (leader[v] == v) {
                wl = __fadd_rn(wl, __fmul_rn((float)label_value(lab[v], label_kind), (float)gsum[v]));
                ws = __dadd_rn(ws, gsum[v]);
            }
            result = (double)__fdiv_rn(wl, (float)ws);
        } else if (task == 0) {
            double wl = 0.0, ws = 0.0;
            for (int v = 0; v < k; ++v) if (leader[v] == v) {
                wl = __dadd_rn(wl, __dmul_rn(label_value(lab[v], label_kind), gsum[v]));   // key * value, then +=
                ws = __dadd_rn(ws, gsum[v]);
            }
            result = wl / ws;
        } else {
            int best = 0;
            for (int v = 1; v < k; ++v) if (leader[v] == v && gsum[v] > gsum[best]) best = v;
            result = label_value(lab[best], label_kind);
        }
        out_pred[q] = result;
    }
}

}  // namespace osml
