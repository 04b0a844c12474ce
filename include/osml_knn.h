/*
 * osml_knn.h -- C ABI of the B200-native brute-force weighted kNN predict path.
 *
 * Drop-in boundary for ViktorC/OSML's kNN hot path.  The reference has no FFI (it is pure Python); the entry points
 * below are what a cgo/ctypes binding for that path would bind, one per reference hook:
 *
 *   osml_knn_create            replaces KNearestNeighbors._fit                      (osml/predict.py:767-771)
 *   osml_knn_predict           replaces KNearestNeighbors{Regression,Classification}._predict
 *                                                                                   (osml/predict.py:785-794, 805-811)
 *                              i.e. _distances (740-742) + _select_weighted_nearest_neighbors (748-765) + aggregation
 *   osml_knn_neighbours_*      exposes what predict.py:748-752 computes internally (indices/distances of the k+1
 *                              nearest), needed for parity checks and for the train-sharded multi-GPU mode
 *   osml_knn_merge_predict_*   the exchange step of the train-sharded mode; no reference twin
 *
 * Plain pointers and sizes only; the caller owns every buffer it passes, the handle owns its device memory.
 * A handle is bound to one CUDA device and is not thread-safe; distinct handles are independent.
 * Every function returns a status; OSML_KNN_EINVAL maps to the reference's bare ValueError (predict.py:729-732,
 * 768-769; model.py:19-34), everything else to RuntimeError.  osml_knn_last_error() gives the message.
 *
 * Tie rule (documented deviation): neighbours are ordered by (distance, train index) ascending with
 * NaN distances last; the reference's order among exactly equal distances is implementation-defined
 * (numpy introselect + pandas quicksort).
 */
#ifndef OSML_KNN_H
#define OSML_KNN_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct osml_knn_index osml_knn_index;

/* element type of the train / query matrix == dtype of DataFrame.values (predict.py:770) */
enum osml_knn_dtype {
    OSML_KNN_F32 = 0, OSML_KNN_F64 = 1,
    OSML_KNN_I8 = 2, OSML_KNN_I16 = 3, OSML_KNN_I32 = 4, OSML_KNN_I64 = 5,
    OSML_KNN_U8 = 6, OSML_KNN_U16 = 7, OSML_KNN_U32 = 8, OSML_KNN_U64 = 9
};

/* memory layout of a host matrix handed to the library */
enum osml_knn_layout { OSML_KNN_ROW_MAJOR = 0 /* C order */, OSML_KNN_COL_MAJOR = 1 /* F order */ };

/* order in which numpy's sum(axis=1) adds the D squared differences (predict.py:742); depends on the layout of the
 * reference's stored matrix: F order -> feature by feature, C order -> numpy pairwise_sum */
enum osml_knn_sum_order { OSML_KNN_SUM_PAIRWISE = 0, OSML_KNN_SUM_SEQUENTIAL = 1 };

enum osml_knn_task { OSML_KNN_REGRESSION = 0, OSML_KNN_CLASSIFICATION = 1 };

enum osml_knn_status {
    OSML_KNN_OK = 0,
    OSML_KNN_EINVAL = 1,       /* the reference raises ValueError here */
    OSML_KNN_ECUDA = 2,        /* CUDA runtime / driver failure */
    OSML_KNN_ENOMEM = 3,
    OSML_KNN_EUNSUPPORTED = 4  /* valid for the reference, not implemented on the GPU path (fails loudly) */
};

/* which distance/selection kernel osml_knn_predict uses */
enum osml_knn_path {
    OSML_KNN_PATH_AUTO = 0,    /* tensor filter for float32 / float64 data with at least 4096 rows, built once 128 queries
                                  have been asked for; direct otherwise, and after a call most of whose queries the
                                  filter could not prove (tie-ridden data) */
    OSML_KNN_PATH_DIRECT = 1,  /* K2: exact direct-difference kernel on CUDA cores, any dtype */
    OSML_KNN_PATH_TENSOR = 2   /* K1: tcgen05 bf16 filter + K3 exact rescoring (+ K2 for unproven queries) */
};

typedef struct osml_knn_stats {
    int64_t kernel_launches;     /* kernels launched by this handle since create / last reset */
    int64_t queries;             /* queries answered */
    int64_t fallback_queries;    /* queries whose tensor-filter candidate list could not be proven complete and
                                    were re-run on the exact direct kernel */
    int64_t tensor_queries;      /* queries that went through the tcgen05 filter */
    double last_filter_ms;       /* device time of the dominant kernel (K1 or K2) in the last predict call */
    double last_total_ms;        /* device time of the whole last predict call (CUDA events) */
    int64_t last_filter_launches;
    int64_t device_bytes;        /* device memory held by the handle */
    int64_t retry_queries;       /* queries K3 could not prove on the first filter pass and that went through the
                                    finer-grained second filter pass (those it still cannot prove count as fallback) */
    int64_t filter_precision;    /* element type of the tensor filter copy in use: 0 not built yet, 1 bf16, 2 fp16 (option
                                    "precision": 0 auto -- bf16, fp16 once a call sent > 0.25 % of its queries to the
                                    second pass --, 1 bf16, 2 fp16) */
} osml_knn_stats;

/*
 * Builds an index on `device` from a host matrix (KNearestNeighbors._fit, predict.py:767-771).
 *   train        n x d elements of `dtype`, laid out as `layout` says (contiguous)
 *   sum_order    see enum; normally SEQUENTIAL for COL_MAJOR and PAIRWISE for ROW_MAJOR
 *   labels       n_labels elements of `label_dtype`; n_labels == n unless the train rows are a shard of a larger
 *                set (then labels are the full replicated label vector and index_base is this shard's first row)
 *   k, rbf_gamma, epsilon   constructor arguments (predict.py:727-738, 781, 801; epsilon predict.py:17)
 *   gamma_is_f32 non-zero if rbf_gamma was a numpy float32 scalar (then -gamma*r^2 rounds to fp32 as numpy would)
 * EINVAL: k < 1, rbf_gamma <= 0, epsilon <= 0, n < k + 1 (predict.py:729-732, 768-769), n_labels < index_base + n.
 */
int osml_knn_create(osml_knn_index** out, int device,
                    const void* train, int dtype, int layout, int64_t n, int64_t d, int sum_order,
                    const void* labels, int label_dtype, int64_t n_labels, int64_t index_base,
                    int k, double rbf_gamma, int gamma_is_f32, double epsilon);

int osml_knn_destroy(osml_knn_index* index);

/*
 * Full predict from host buffers (the reference's _predict).  queries: q x d elements of the index dtype.
 * out_pred: q doubles (what _predict returns: fp64 predictions; class labels as fp64).
 * out_idx / out_dist (nullable): q x (k+1) neighbour train indices / distances, ascending by (distance, index);
 * out_dist elements are float for F32 data, double otherwise (the dtype numpy's sqrt yields).
 * EINVAL: q < 1, or n < k + 2 (np.argpartition(d, k+1) raises ValueError).
 */
int osml_knn_predict(osml_knn_index* index, const void* queries, int layout, int64_t q, int task,
                     double* out_pred, int64_t* out_idx, void* out_dist);

/* Same with device-resident, row-major queries and device outputs, enqueued on `stream` (a cudaStream_t; NULL =
 * default stream).  Does not synchronise unless the tensor path needs its fallback count. */
int osml_knn_predict_device(osml_knn_index* index, const void* d_queries, int64_t q, int task,
                            double* d_out_pred, int64_t* d_out_idx, void* d_out_dist, void* stream);

/* Neighbour search only: this handle's k+1 nearest per query with global indices (index_base added), device
 * buffers, ascending.  Lists of a shard with fewer than k+1 rows are padded with (NaN, INT64_MAX). */
int osml_knn_neighbours_device(osml_knn_index* index, const void* d_queries, int64_t q,
                               int64_t* d_out_idx, void* d_out_dist, void* stream);

/* Exchange step of the train-sharded mode: merges `parts` neighbour lists per query (layout [part][q][k+1], e.g. the
 * output of an all-gather of osml_knn_neighbours_device results) by (distance, index) and aggregates with the
 * handle's replicated labels.  d_out_idx / d_out_dist nullable. */
int osml_knn_merge_predict_device(osml_knn_index* index, const int64_t* d_part_idx, const void* d_part_dist,
                                  int parts, int64_t q, int task,
                                  double* d_out_pred, int64_t* d_out_idx, void* d_out_dist, void* stream);

/* Slice-wise exchange of the train-sharded mode ("each GPU merges ... for its ceil(Q/G) query
 * slice"; the reference keeps the whole matrix in one process, osml/predict.py:770).  pack: this handle's lists for
 * all q queries ([q][k+1], as osml_knn_neighbours_device wrote them) -> 16-byte records {int64 index, distance bits
 * (a float32 distance in the low word)} laid out [parts][slice_q][k+1] with slice_q = ceil(q / parts); queries past q
 * are padded with empty entries.  One all-to-all with equal splits (parts * ... -> every rank receives
 * [parts][slice_q][k+1]: the partial lists of ITS slice from every rank) feeds merge_predict_packed, which merges the
 * first q_valid queries of the slice by (distance, index) and aggregates them.  d_out_idx / d_out_dist nullable. */
int osml_knn_pack_slices_device(osml_knn_index* index, const int64_t* d_idx, const void* d_dist, int64_t q, int parts,
                                void* d_records, void* stream);
int osml_knn_merge_predict_packed_device(osml_knn_index* index, const void* d_records, int parts, int64_t slice_q,
                                         int64_t q_valid, int task, double* d_out_pred, int64_t* d_out_idx,
                                         void* d_out_dist, void* stream);

/* Options: "path" (enum osml_knn_path), "slack" (extra candidates the tensor filter keeps beyond k+1),
 * "query_block" (queries per internal batch), "pipeline" (1 = default: osml_knn_predict uploads a query matrix of
 * 4 MB or more in blocks, every block but the first under the filter kernel of the block in front of it; 0 = one block),
 * "pipeline_blocks" (0 = default: two blocks from pinned memory, three from pageable memory; n = that many),
 * "stager" (1 = default: pageable host matrices of 2 MB or more go up through a ring of pinned staging slots filled
 * by worker threads; 0 = plain cudaMemcpyAsync).  The remaining names are tuning knobs of the tensor filter for
 * experiments.  Unknown name -> EINVAL. */
int osml_knn_set_option(osml_knn_index* index, const char* name, double value);

int osml_knn_get_stats(osml_knn_index* index, osml_knn_stats* out);
int osml_knn_reset_stats(osml_knn_index* index);

/* Shape of the index: n, d, k, dtype, bytes per distance element (4 or 8). */
int osml_knn_describe(osml_knn_index* index, int64_t* n, int64_t* d, int* k, int* dtype, int* dist_elem_size);

/* Message of the last failure on the calling thread ("" if none). */
const char* osml_knn_last_error(void);

/*
 * Column transform in front of kNN: Centering (osml/transform.py:53-66, denom == NULL) and Standardization
 * (osml/transform.py:69-84), out[i][j] = (in[i][j] - shift[j]) / denom[j], in numpy's arithmetic: OSML_KNN_F32 input
 * gives float32 output, every other dtype float64 output (numpy subtracts a float64 mean from integer data); one
 * rounding for the subtraction, one for the division.
 *   in, out      n x d host matrices in the same `layout`; out has the OUTPUT dtype
 *   shift, denom d elements of the output dtype (mean, and sd + epsilon as numpy rounds it); host pointers
 * Replaces TransformativeModel._transform (osml/transform.py:64-65, 83-84) for the two classes.
 */
int osml_knn_column_transform(int device, const void* in, int dtype, int layout, int64_t n, int64_t d,
                              const void* shift, const void* denom, void* out);

/*
 * Fuses the same transform into osml_knn_predict / osml_knn_predict_device: the queries are passed RAW (dtype of the
 * index, which must be OSML_KNN_F32 or OSML_KNN_F64) and are transformed on the device right after the upload, so
 * the host pass over Q x D of `standardization.transform(query_df)` disappears.  shift == NULL removes the
 * transform; denom may be NULL (centering).  The vectors are copied.
 */
int osml_knn_set_query_transform(osml_knn_index* ix, const void* shift, const void* denom);

/* ABI version of this header: major * 1000 + minor. */
int osml_knn_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* OSML_KNN_H */
