/*
 * osml_svm.h -- C ABI of the B200-native support-vector decision function (widening row f4 of SURVEY.md section 8).
 *
 * ViktorC/OSML's SupportVectorMachine evaluates the same sqrt(sum(square(x1 - x2))) expression as the kNN hot path
 * against every (support) vector per query.  The reference has no FFI; these entry points are what a binding for
 * that path would bind, one per reference hook:
 *
 *   osml_svm_create            holds what _set_parameters keeps (support vectors, their -1 / +1 labels, multipliers,
 *                              intercept; osml/predict.py:1107-1116) -- or, while fitting, ALL training rows
 *   osml_svm_decision          replaces SupportVectorMachine._raw_predict           (osml/predict.py:1095-1099)
 *                              with _linear_kernel / _polynomial_kernel / _gaussian_kernel (1086-1093) inside
 *   osml_svm_decision_self     the same with the machine's own vectors as queries = the values _update_error_cache
 *                              recomputes after every SMO step, before the labels are subtracted (1101-1105)
 *   osml_svm_gram              K(x_i, x_j) for all training pairs: the k_ii, k_jj, k_ij of _optimize_ij (1140-1142)
 *   osml_svm_set_multipliers   new multipliers / intercept after an SMO step (the alphas[i], alphas[j], intercept[0]
 *                              assignments of _optimize_ij, 1151-1161)
 *
 * Status codes, dtype / layout / sum-order enums and osml_knn_last_error() are those of osml_knn.h.
 * Plain pointers and sizes only; the caller owns every buffer it passes, the handle owns its device memory.
 * A handle is bound to one CUDA device and is not thread-safe.
 *
 * Arithmetic follows the reference's dtypes (float32 data: float32 kernel values; integers wrap like numpy; float64
 * products and sums).  The sum over the vectors runs in a fixed GPU order, not numpy's pairwise order, and exp / pow are
 * CUDA's: results agree with the reference to ~1e-13 relative for float64 data and ~1e-6 for float32 data (inside the
 * 1e-5 of BASELINE.json), not bit for bit.  For float data, vectors whose multiplier is exactly zero are skipped: their
 * terms are exact zeros unless the data itself holds NaN / inf (then the reference's NaN * 0 = NaN is not reproduced).
 */
#ifndef OSML_SVM_H
#define OSML_SVM_H

#include "osml_knn.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct osml_svm_machine osml_svm_machine;

/* SupportVectorMachine(kernel=...) (osml/predict.py:1057-1070) */
enum osml_svm_kernel { OSML_SVM_LINEAR = 0 /* 'lin' */, OSML_SVM_POLYNOMIAL = 1 /* 'poly' */, OSML_SVM_GAUSSIAN = 2 /* 'rbf' */ };

/* vectors: n x d matrix of `dtype` in `layout`; sum_order = order of numpy's sum(axis=1) over the features for the
 * reference's copy of the matrix (OSML_KNN_SUM_PAIRWISE for the C-ordered observations[alphas > 0] used at predict
 * time, OSML_KNN_SUM_SEQUENTIAL for the F-ordered DataFrame.values used while fitting).  labels: n doubles, -1 / +1.
 * alphas: n doubles.  lin / poly need a float dtype (EUNSUPPORTED otherwise); rbf takes every dtype.
 * EINVAL: null pointers, n < 1, d < 1, bad enum, gamma <= 0 for rbf, degree < 1 for poly (predict.py:1062-1068). */
int osml_svm_create(osml_svm_machine** out, int device, const void* vectors, int dtype, int layout, int64_t n,
                    int64_t d, int sum_order, const double* labels, const double* alphas, double intercept,
                    int kernel, double gamma, double degree, double poly_coefficient);

int osml_svm_destroy(osml_svm_machine* machine);

/* replaces all n multipliers and the intercept */
int osml_svm_set_multipliers(osml_svm_machine* machine, const double* alphas, double intercept);

/* out[i] = intercept + sum_j K(vectors_j, queries_i) * alphas_j * labels_j; queries: q x d host matrix of the machine's
 * dtype; out: q doubles (host) */
int osml_svm_decision(osml_svm_machine* machine, const void* queries, int layout, int64_t q, double* out);

/* queries = the machine's own vectors; out: n doubles (host) */
int osml_svm_decision_self(osml_svm_machine* machine, double* out);

/* out[i * n + j] = K(vectors_i, vectors_j) for all pairs of the machine's vectors (n * n doubles, host): the scalar kernel
 * values _optimize_ij evaluates on single rows (osml/predict.py:1140-1142), once per fit instead of three per attempt.
 * Values are computed in the data's dtype (single rows: numpy's pairwise order over the features) and widened. */
int osml_svm_gram(osml_svm_machine* machine, double* out);

/* device pointers: d_queries row-major q x d, d_out q doubles; runs on `stream` (a cudaStream_t; NULL = the handle's
 * own stream, synchronised before returning) */
int osml_svm_decision_device(osml_svm_machine* machine, const void* d_queries, int64_t q, double* d_out,
                             void* stream);

/* milliseconds the decision kernel of the last call took on the device (CUDA events), kernel launches so far */
int osml_svm_get_stats(osml_svm_machine* machine, double* last_kernel_ms, int64_t* kernel_launches);

#ifdef __cplusplus
}
#endif
#endif /* OSML_SVM_H */
