"""ctypes binding of the C ABI in include/osml_knn.h and include/osml_svm.h (libosml_knn.so, built in-tree by
osml_b200/build.py).

This is the only door to the arithmetic of the path.  It fails loudly: a missing library raises ImportError with the
build command, a non-zero status raises ValueError (where the reference raises ValueError) or RuntimeError.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# OSML_KNN_LIB: development override for A/B timing of two builds of the same library (tools/ab.sh)
LIB_PATH = os.environ.get('OSML_KNN_LIB') or os.path.join(_HERE, 'libosml_knn.so')

OK, EINVAL, ECUDA, ENOMEM, EUNSUPPORTED = 0, 1, 2, 3, 4
ROW_MAJOR, COL_MAJOR = 0, 1
SUM_PAIRWISE, SUM_SEQUENTIAL = 0, 1
REGRESSION, CLASSIFICATION = 0, 1
PATH_AUTO, PATH_DIRECT, PATH_TENSOR = 0, 1, 2
PATHS = {'auto': PATH_AUTO, 'direct': PATH_DIRECT, 'tensor': PATH_TENSOR}

DTYPE_CODES = {
    np.dtype(np.float32): 0, np.dtype(np.float64): 1,
    np.dtype(np.int8): 2, np.dtype(np.int16): 3, np.dtype(np.int32): 4, np.dtype(np.int64): 5,
    np.dtype(np.uint8): 6, np.dtype(np.uint16): 7, np.dtype(np.uint32): 8, np.dtype(np.uint64): 9,
}

# every symbol include/osml_knn.h declares; tests/test_abi.py checks the header against this table
SYMBOLS = ('osml_knn_create', 'osml_knn_destroy', 'osml_knn_predict', 'osml_knn_predict_device',
           'osml_knn_neighbours_device', 'osml_knn_merge_predict_device', 'osml_knn_set_option',
           'osml_knn_get_stats', 'osml_knn_reset_stats', 'osml_knn_describe', 'osml_knn_last_error',
           'osml_knn_abi_version', 'osml_knn_column_transform', 'osml_knn_set_query_transform',
           'osml_knn_pack_slices_device', 'osml_knn_merge_predict_packed_device')
# every symbol include/osml_svm.h declares
SVM_SYMBOLS = ('osml_svm_create', 'osml_svm_destroy', 'osml_svm_set_multipliers', 'osml_svm_decision',
               'osml_svm_decision_self', 'osml_svm_decision_device', 'osml_svm_get_stats', 'osml_svm_gram')
SVM_KERNELS = {'lin': 0, 'poly': 1, 'rbf': 2}


class Stats(ctypes.Structure):
    _fields_ = [('kernel_launches', ctypes.c_int64), ('queries', ctypes.c_int64),
                ('fallback_queries', ctypes.c_int64), ('tensor_queries', ctypes.c_int64),
                ('last_filter_ms', ctypes.c_double), ('last_total_ms', ctypes.c_double),
                ('last_filter_launches', ctypes.c_int64), ('device_bytes', ctypes.c_int64),
                ('retry_queries', ctypes.c_int64), ('filter_precision', ctypes.c_int64)]

    def as_dict(self):
        return {name: getattr(self, name) for name, _ in self._fields_}


_lib = None


def library_path():
    """Where the C-ABI library lives (what a binding in the reference would hand to ctypes.CDLL)."""
    return LIB_PATH


def lib():
    """Loads libosml_knn.so once.  No fallback: raises ImportError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError('%s is missing: build it with `python -m osml_b200.build` (nvcc, sm_100a). '
                          'osml_b200 has no CPU fallback.' % LIB_PATH)
    cdll = ctypes.CDLL(LIB_PATH)
    c_i64, c_int, c_dbl, c_vp = ctypes.c_int64, ctypes.c_int, ctypes.c_double, ctypes.c_void_p
    cdll.osml_knn_create.argtypes = [ctypes.POINTER(c_vp), c_int, c_vp, c_int, c_int, c_i64, c_i64, c_int, c_vp, c_int,
                                     c_i64, c_i64, c_int, c_dbl, c_int, c_dbl]
    cdll.osml_knn_destroy.argtypes = [c_vp]
    cdll.osml_knn_predict.argtypes = [c_vp, c_vp, c_int, c_i64, c_int, c_vp, c_vp, c_vp]
    cdll.osml_knn_predict_device.argtypes = [c_vp, c_vp, c_i64, c_int, c_vp, c_vp, c_vp, c_vp]
    cdll.osml_knn_neighbours_device.argtypes = [c_vp, c_vp, c_i64, c_vp, c_vp, c_vp]
    cdll.osml_knn_merge_predict_device.argtypes = [c_vp, c_vp, c_vp, c_int, c_i64, c_int, c_vp, c_vp, c_vp, c_vp]
    cdll.osml_knn_pack_slices_device.argtypes = [c_vp, c_vp, c_vp, c_i64, c_int, c_vp, c_vp]
    cdll.osml_knn_merge_predict_packed_device.argtypes = [c_vp, c_vp, c_int, c_i64, c_i64, c_int, c_vp, c_vp, c_vp, c_vp]
    cdll.osml_knn_set_option.argtypes = [c_vp, ctypes.c_char_p, c_dbl]
    cdll.osml_knn_get_stats.argtypes = [c_vp, ctypes.POINTER(Stats)]
    cdll.osml_knn_reset_stats.argtypes = [c_vp]
    cdll.osml_knn_describe.argtypes = [c_vp, ctypes.POINTER(c_i64), ctypes.POINTER(c_i64), ctypes.POINTER(c_int),
                                       ctypes.POINTER(c_int), ctypes.POINTER(c_int)]
    cdll.osml_knn_column_transform.argtypes = [c_int, c_vp, c_int, c_int, c_i64, c_i64, c_vp, c_vp, c_vp]
    cdll.osml_knn_set_query_transform.argtypes = [c_vp, c_vp, c_vp]
    cdll.osml_knn_last_error.restype = ctypes.c_char_p
    cdll.osml_knn_last_error.argtypes = []
    cdll.osml_knn_abi_version.argtypes = []
    cdll.osml_svm_create.argtypes = [ctypes.POINTER(c_vp), c_int, c_vp, c_int, c_int, c_i64, c_i64, c_int, c_vp, c_vp,
                                     c_dbl, c_int, c_dbl, c_dbl, c_dbl]
    cdll.osml_svm_destroy.argtypes = [c_vp]
    cdll.osml_svm_set_multipliers.argtypes = [c_vp, c_vp, c_dbl]
    cdll.osml_svm_decision.argtypes = [c_vp, c_vp, c_int, c_i64, c_vp]
    cdll.osml_svm_decision_self.argtypes = [c_vp, c_vp]
    cdll.osml_svm_gram.argtypes = [c_vp, c_vp]
    cdll.osml_svm_decision_device.argtypes = [c_vp, c_vp, c_i64, c_vp, c_vp]
    cdll.osml_svm_get_stats.argtypes = [c_vp, ctypes.POINTER(c_dbl), ctypes.POINTER(c_i64)]
    for name in SYMBOLS + SVM_SYMBOLS:
        if name not in ('osml_knn_last_error',):
            getattr(cdll, name).restype = c_int
    _lib = cdll
    return _lib


def check(status):
    """Maps a status to the reference's exception types."""
    if status == OK:
        return
    message = lib().osml_knn_last_error().decode('utf-8', 'replace')
    if status == EINVAL:
        raise ValueError(message)
    if status == EUNSUPPORTED:
        raise NotImplementedError(message)
    if status == ENOMEM:
        raise MemoryError(message)
    raise RuntimeError('osml_knn: ' + message)


def dtype_code(dtype):
    try:
        return DTYPE_CODES[np.dtype(dtype)]
    except KeyError:
        raise NotImplementedError('dtype %s is not supported by the GPU kNN path' % dtype)


def matrix_layout(values):
    """(contiguous array, layout code) for a 2-D host matrix, copying only when it is neither C nor F contiguous."""
    if values.flags['C_CONTIGUOUS']:
        return values, ROW_MAJOR
    if values.flags['F_CONTIGUOUS']:
        return values, COL_MAJOR
    return np.ascontiguousarray(values), ROW_MAJOR


def sum_order_of(values):
    """Order in which numpy's ``sum(axis=1)`` adds the squared differences of ``values - q`` (predict.py:742): the
    subtraction result inherits the stride order of ``values``; feature by feature when axis 0 is the fast axis (the
    F-ordered ``DataFrame.values`` of pandas 3), numpy's pairwise tree when axis 1 is."""
    n, d = values.shape
    if d == 1:
        return SUM_SEQUENTIAL
    if n == 1:          # a single row is contiguous along the features whatever its flags say: numpy's pairwise_sum
        return SUM_PAIRWISE
    s0, s1 = (abs(s) for s in values.strides)
    return SUM_SEQUENTIAL if s0 <= s1 else SUM_PAIRWISE


def transform_dtype(dtype):
    """Output dtype of the reference's Centering / Standardization for a frame of `dtype` (numpy promotion rules)."""
    return np.dtype(np.float32) if np.dtype(dtype) == np.float32 else np.dtype(np.float64)


def column_transform(values, shift, denom=None, device=0):
    """out = (values - shift) / denom per column on the GPU (osml_knn_column_transform); same memory order as
    ``values``.  ``denom=None``: centering only."""
    values = np.asarray(values)
    if values.ndim != 2:
        raise ValueError
    mat, layout = matrix_layout(values)
    out_dtype = transform_dtype(mat.dtype)
    shift = np.ascontiguousarray(shift, dtype=out_dtype)
    denom_p = None
    if denom is not None:
        denom = np.ascontiguousarray(denom, dtype=out_dtype)
        denom_p = denom.ctypes.data
    n, d = mat.shape
    if shift.shape != (d,) or (denom is not None and denom.shape != (d,)):
        raise ValueError
    out = np.empty((n, d), dtype=out_dtype, order='F' if layout == COL_MAJOR else 'C')
    check(lib().osml_knn_column_transform(int(device), mat.ctypes.data, dtype_code(mat.dtype), layout, int(n), int(d),
                                          shift.ctypes.data, denom_p, out.ctypes.data))
    return out


class Index:
    """Owner of one ``osml_knn_index`` handle."""

    def __init__(self, train, labels, k, rbf_gamma, epsilon, device=0, index_base=0, sum_order=None,
                 gamma_is_f32=False):
        self._handle = ctypes.c_void_p()
        train = np.asarray(train)
        labels = np.ascontiguousarray(labels)
        if train.ndim != 2:
            raise ValueError
        if sum_order is None:
            sum_order = sum_order_of(train)
        mat, layout = matrix_layout(train)
        self.dtype = mat.dtype
        self.k = int(k)
        self.n, self.d = int(mat.shape[0]), int(mat.shape[1])
        self.dist_dtype = np.dtype(np.float32) if mat.dtype == np.float32 else np.dtype(np.float64)
        self.device = int(device)
        check(lib().osml_knn_create(ctypes.byref(self._handle), self.device, mat.ctypes.data, dtype_code(mat.dtype),
                                    layout, self.n, self.d, int(sum_order), labels.ctypes.data,
                                    dtype_code(labels.dtype), int(labels.shape[0]), int(index_base), self.k,
                                    float(rbf_gamma), int(bool(gamma_is_f32)), float(epsilon)))

    def close(self):
        if getattr(self, '_handle', None) is not None and self._handle.value:
            lib().osml_knn_destroy(self._handle)
            self._handle = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_option(self, name, value):
        check(lib().osml_knn_set_option(self._handle, name.encode(), float(value)))

    def set_query_transform(self, shift, denom=None):
        """Queries are passed raw from now on and become (q - shift) / denom on the device (None: remove)."""
        if shift is None:
            check(lib().osml_knn_set_query_transform(self._handle, None, None))
            return
        shift = np.ascontiguousarray(shift, dtype=self.dtype)
        denom_p = None
        if denom is not None:
            denom = np.ascontiguousarray(denom, dtype=self.dtype)
            denom_p = denom.ctypes.data
        if shift.shape != (self.d,) or (denom is not None and denom.shape != (self.d,)):
            raise ValueError
        check(lib().osml_knn_set_query_transform(self._handle, shift.ctypes.data, denom_p))

    def stats(self):
        st = Stats()
        check(lib().osml_knn_get_stats(self._handle, ctypes.byref(st)))
        return st.as_dict()

    def reset_stats(self):
        check(lib().osml_knn_reset_stats(self._handle))

    def predict(self, queries, task, return_neighbours=False):
        """Host buffers in, host buffers out (osml_knn_predict)."""
        queries = np.asarray(queries)
        if queries.ndim != 2 or queries.shape[1] != self.d or queries.dtype != self.dtype:
            raise ValueError
        mat, layout = matrix_layout(queries)
        q = int(mat.shape[0])
        pred = np.empty(q, dtype=np.float64)
        idx = dist = None
        idx_p = dist_p = None
        if return_neighbours:
            idx = np.empty((q, self.k + 1), dtype=np.int64)
            dist = np.empty((q, self.k + 1), dtype=self.dist_dtype)
            idx_p, dist_p = idx.ctypes.data, dist.ctypes.data
        check(lib().osml_knn_predict(self._handle, mat.ctypes.data, layout, q, int(task), pred.ctypes.data, idx_p,
                                     dist_p))
        if return_neighbours:
            return pred, idx, dist
        return pred

    # device-pointer entry points (raw addresses: torch ``tensor.data_ptr()`` or any CUDA allocation) -------------
    def predict_device(self, d_queries, q, task, d_out_pred, d_out_idx=None, d_out_dist=None, stream=None):
        check(lib().osml_knn_predict_device(self._handle, d_queries, int(q), int(task), d_out_pred, d_out_idx,
                                            d_out_dist, stream))

    def neighbours_device(self, d_queries, q, d_out_idx, d_out_dist, stream=None):
        check(lib().osml_knn_neighbours_device(self._handle, d_queries, int(q), d_out_idx, d_out_dist, stream))

    def merge_predict_device(self, d_part_idx, d_part_dist, parts, q, task, d_out_pred, d_out_idx=None,
                             d_out_dist=None, stream=None):
        check(lib().osml_knn_merge_predict_device(self._handle, d_part_idx, d_part_dist, int(parts), int(q),
                                                  int(task), d_out_pred, d_out_idx, d_out_dist, stream))

    def pack_slices_device(self, d_idx, d_dist, q, parts, d_records, stream=None):
        """[q][k+1] lists -> 16-byte records [parts][ceil(q / parts)][k+1] (the all-to-all send buffer)."""
        check(lib().osml_knn_pack_slices_device(self._handle, d_idx, d_dist, int(q), int(parts), d_records, stream))

    def merge_predict_packed_device(self, d_records, parts, slice_q, q_valid, task, d_out_pred, d_out_idx=None,
                                    d_out_dist=None, stream=None):
        check(lib().osml_knn_merge_predict_packed_device(self._handle, d_records, int(parts), int(slice_q),
                                                         int(q_valid), int(task), d_out_pred, d_out_idx, d_out_dist,
                                                         stream))


class Machine:
    """Owner of one ``osml_svm_machine`` handle (include/osml_svm.h): vectors, -1 / +1 labels, multipliers and
    intercept of a kernel machine on the GPU, and its decision function."""

    def __init__(self, vectors, labels, alphas, intercept, kernel, gamma=.5, degree=2, poly_coefficient=1., device=0,
                 sum_order=None):
        self._handle = ctypes.c_void_p()
        vectors = np.asarray(vectors)
        if vectors.ndim != 2:
            raise ValueError
        if sum_order is None:
            sum_order = sum_order_of(vectors)
        mat, layout = matrix_layout(vectors)
        labels = np.ascontiguousarray(labels, dtype=np.float64)
        alphas = np.ascontiguousarray(alphas, dtype=np.float64)
        self.dtype = mat.dtype
        self.n, self.d = int(mat.shape[0]), int(mat.shape[1])
        if labels.shape != (self.n,) or alphas.shape != (self.n,):
            raise ValueError
        if kernel not in SVM_KERNELS:
            raise ValueError
        self.device = int(device)
        check(lib().osml_svm_create(ctypes.byref(self._handle), self.device, mat.ctypes.data, dtype_code(mat.dtype),
                                    layout, self.n, self.d, int(sum_order), labels.ctypes.data, alphas.ctypes.data,
                                    float(intercept), SVM_KERNELS[kernel], float(gamma), float(degree),
                                    float(poly_coefficient)))

    def close(self):
        if getattr(self, '_handle', None) is not None and self._handle.value:
            lib().osml_svm_destroy(self._handle)
            self._handle = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_multipliers(self, alphas, intercept):
        alphas = np.ascontiguousarray(alphas, dtype=np.float64)
        if alphas.shape != (self.n,):
            raise ValueError
        check(lib().osml_svm_set_multipliers(self._handle, alphas.ctypes.data, float(intercept)))

    def decision(self, queries):
        """Host matrix in, float64 decision values out (osml_svm_decision)."""
        queries = np.asarray(queries)
        if queries.ndim != 2 or queries.shape[1] != self.d or queries.dtype != self.dtype:
            raise ValueError
        mat, layout = matrix_layout(queries)
        out = np.empty(mat.shape[0], dtype=np.float64)
        check(lib().osml_svm_decision(self._handle, mat.ctypes.data, layout, int(mat.shape[0]), out.ctypes.data))
        return out

    def decision_self(self):
        """Decision value of every vector of the machine against the machine (osml_svm_decision_self)."""
        out = np.empty(self.n, dtype=np.float64)
        check(lib().osml_svm_decision_self(self._handle, out.ctypes.data))
        return out

    def gram(self):
        """K(v_i, v_j) for all pairs of the machine's vectors, as an (n, n) array of the dtype the reference's kernel
        function returns for this data (float32 for float32 data, else float64) (osml_svm_gram)."""
        out = np.empty((self.n, self.n), dtype=np.float64)
        check(lib().osml_svm_gram(self._handle, out.ctypes.data))
        return out.astype(np.float32) if self.dtype == np.float32 else out

    def decision_device(self, d_queries, q, d_out, stream=None):
        check(lib().osml_svm_decision_device(self._handle, d_queries, int(q), d_out, stream))

    def stats(self):
        ms, launches = ctypes.c_double(), ctypes.c_int64()
        check(lib().osml_svm_get_stats(self._handle, ctypes.byref(ms), ctypes.byref(launches)))
        return {'last_kernel_ms': ms.value, 'kernel_launches': launches.value}
