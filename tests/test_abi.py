"""CPU tests of the drop-in boundary: the C-ABI library builds, loads, and exports every symbol include/osml_knn.h and
include/osml_svm.h declare (no compute calls -- there is no GPU here), and the ctypes tables agree with the headers."""
import ctypes
import os
import re

import numpy as np
import pytest

from osml_b200 import _native, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, 'include', 'osml_knn.h')
SVM_HEADER = os.path.join(ROOT, 'include', 'osml_svm.h')


def _declared_symbols(header=HEADER, prefix='osml_knn_'):
    text = open(header).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(%s[a-z_]+)\s*\(' % prefix, text)))


@pytest.fixture(scope='module')
def lib():
    build.build()
    return _native.lib()


def test_header_and_binding_agree():
    assert _declared_symbols() == sorted(_native.SYMBOLS)
    assert _declared_symbols(SVM_HEADER, 'osml_svm_') == sorted(_native.SVM_SYMBOLS)
    assert sorted(os.listdir(os.path.join(ROOT, 'include'))) == ['osml_knn.h', 'osml_svm.h']


def test_library_exports_every_declared_symbol(lib):
    raw = ctypes.CDLL(_native.LIB_PATH)
    for name in _declared_symbols() + _declared_symbols(SVM_HEADER, 'osml_svm_'):
        assert getattr(raw, name) is not None
    assert lib.osml_knn_abi_version() == 1004


def test_enum_values_match_the_header():
    text = open(HEADER).read()
    for name, value in [('OSML_KNN_F32', 0), ('OSML_KNN_F64', 1), ('OSML_KNN_I8', 2), ('OSML_KNN_U64', 9),
                        ('OSML_KNN_ROW_MAJOR', _native.ROW_MAJOR), ('OSML_KNN_COL_MAJOR', _native.COL_MAJOR),
                        ('OSML_KNN_SUM_PAIRWISE', _native.SUM_PAIRWISE),
                        ('OSML_KNN_SUM_SEQUENTIAL', _native.SUM_SEQUENTIAL),
                        ('OSML_KNN_REGRESSION', _native.REGRESSION),
                        ('OSML_KNN_CLASSIFICATION', _native.CLASSIFICATION), ('OSML_KNN_EINVAL', _native.EINVAL),
                        ('OSML_KNN_PATH_TENSOR', _native.PATH_TENSOR)]:
        assert re.search(r'\b%s\s*=\s*%d\b' % (name, value), text), name
    assert _native.dtype_code(np.float32) == 0 and _native.dtype_code(np.uint64) == 9


def test_argument_validation_needs_no_gpu(lib):
    """The reference's ValueError conditions are checked before any CUDA call (predict.py:729-732, 768-769)."""
    handle = ctypes.c_void_p()
    train = np.zeros((4, 2), np.float32)
    labels = np.zeros(4, np.int64)

    def create(k=1, gamma=1.0, eps=1e-8, n=4, n_labels=4, base=0):
        return lib.osml_knn_create(ctypes.byref(handle), 0, train.ctypes.data, 0, 0, n, 2, 1, labels.ctypes.data, 5,
                                   n_labels, base, k, gamma, 0, eps)

    assert create(k=0) == _native.EINVAL
    assert create(gamma=0.0) == _native.EINVAL
    assert create(eps=0.0) == _native.EINVAL
    assert create(k=4) == _native.EINVAL                 # n < k + 1
    assert create(n_labels=3) == _native.EINVAL
    assert b'' != lib.osml_knn_last_error()
    with pytest.raises(ValueError):
        _native.check(_native.EINVAL)
    with pytest.raises(NotImplementedError):
        _native.check(_native.EUNSUPPORTED)
    assert create(k=2000, n=4, n_labels=3000) in (_native.EUNSUPPORTED, _native.EINVAL)   # k > 1024


def test_create_fails_loudly_without_a_gpu(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip('a GPU is present')
    with pytest.raises(RuntimeError):
        _native.Index(np.zeros((8, 2), np.float32), np.zeros(8, np.int64), 3, 1, 1e-8)


def test_sum_order_and_layout_detection():
    c = np.zeros((10, 4), np.float32)
    f = np.asfortranarray(c)
    assert _native.sum_order_of(c) == _native.SUM_PAIRWISE and _native.sum_order_of(f) == _native.SUM_SEQUENTIAL
    assert _native.sum_order_of(c[:, :1]) == _native.SUM_SEQUENTIAL
    assert _native.matrix_layout(c)[1] == _native.ROW_MAJOR and _native.matrix_layout(f)[1] == _native.COL_MAJOR
    sliced = np.zeros((20, 8), np.float32)[::2, ::2]
    mat, layout = _native.matrix_layout(sliced)
    assert mat.flags['C_CONTIGUOUS'] and layout == _native.ROW_MAJOR
    assert _native.sum_order_of(sliced) == _native.SUM_PAIRWISE
    with pytest.raises(NotImplementedError):
        _native.dtype_code(np.float16)


def test_svm_argument_validation_needs_no_gpu(lib):
    """SupportVectorMachine's ValueError conditions (predict.py:1062-1070) and the unsupported combinations are checked
    before any CUDA call; with valid arguments and no device the call fails loudly."""
    text = open(SVM_HEADER).read()
    for name, value in [('OSML_SVM_LINEAR', 0), ('OSML_SVM_POLYNOMIAL', 1), ('OSML_SVM_GAUSSIAN', 2)]:
        assert re.search(r'\b%s\s*=\s*%d\b' % (name, value), text), name
    assert _native.SVM_KERNELS == {'lin': 0, 'poly': 1, 'rbf': 2}
    handle = ctypes.c_void_p()
    vectors = np.zeros((4, 2), np.float64)
    ints = np.zeros((4, 2), np.int8)
    labels = np.ones(4)
    alphas = np.ones(4)

    def create(mat=vectors, dtype=1, n=4, d=2, kernel=2, gamma=.5, degree=2., layout=0, order=1):
        return lib.osml_svm_create(ctypes.byref(handle), 0, mat.ctypes.data, dtype, layout, n, d, order,
                                   labels.ctypes.data, alphas.ctypes.data, 0., kernel, gamma, degree, 1.)

    assert create(gamma=0.) == _native.EINVAL
    assert create(kernel=3) == _native.EINVAL
    assert create(kernel=1, degree=0.) == _native.EINVAL
    assert create(n=0) == _native.EINVAL
    assert create(layout=2) == _native.EINVAL
    assert create(order=5) == _native.EINVAL
    assert create(mat=ints, dtype=2, kernel=0) == _native.EUNSUPPORTED       # lin on int8: not on the GPU path
    assert create(dtype=77) == _native.EUNSUPPORTED
    assert lib.osml_svm_create(None, 0, vectors.ctypes.data, 1, 0, 4, 2, 1, labels.ctypes.data, alphas.ctypes.data, 0.,
                               2, .5, 2., 1.) == _native.EINVAL
    assert lib.osml_svm_decision(None, vectors.ctypes.data, 0, 4, alphas.ctypes.data) == _native.EINVAL
    assert lib.osml_svm_destroy(None) == _native.OK
    import torch
    if not torch.cuda.is_available():
        assert create() == _native.ECUDA and not handle.value
        with pytest.raises(RuntimeError):
            _native.Machine(vectors, labels, alphas, 0., 'rbf')
