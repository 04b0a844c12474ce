"""CPU tests of the drop-in boundary: the C-ABI library builds, loads, and exports every symbol include/osml_knn.h
declares (no compute calls -- there is no GPU here), and the ctypes table agrees with the header."""
import ctypes
import os
import re

import numpy as np
import pytest

from osml_b200 import _native, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, 'include', 'osml_knn.h')


def _declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(osml_knn_[a-z_]+)\s*\(', text)))


@pytest.fixture(scope='module')
def lib():
    build.build()
    return _native.lib()


def test_header_and_binding_agree():
    assert _declared_symbols() == sorted(_native.SYMBOLS)


def test_library_exports_every_declared_symbol(lib):
    raw = ctypes.CDLL(_native.LIB_PATH)
    for name in _declared_symbols():
        assert getattr(raw, name) is not None
    assert lib.osml_knn_abi_version() == 1003


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
    assert create(k=200, n=4, n_labels=300) in (_native.EUNSUPPORTED, _native.EINVAL)


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
