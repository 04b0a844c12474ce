"""Integer matrices on the tensor filter (data.py:213-230 one-hot sets like mushroom).

numpy keeps integer arithmetic in the matrix's own type: subtract and square wrap, the sum widens to (u)int64.  Inside
the range where nothing can wrap (|t - q| <= L, L from the type and the feature count) a float64 shadow of the rows
reproduces that arithmetic exactly, so the float64 tensor path serves the index; a query outside the range, and
anything whose completeness proof fails (exact ties are the norm on integer data), goes to the exact direct kernel.
Either way the result must be the oracle's, bit for bit."""
import numpy as np
import pytest

from osml_b200 import _native
from tests import parity

pytestmark = pytest.mark.gpu


def _both_paths(train, labels, test, k, task='classification'):
    code = _native.REGRESSION if task == 'regression' else _native.CLASSIFICATION
    out = {}
    for path in ('direct', 'tensor'):
        ix = _native.Index(train, labels, k, 1, 1e-8)
        ix.set_option('path', _native.PATHS[path])
        out[path] = ix.predict(test, code, return_neighbours=True) + (ix.stats(),)
        ix.close()
    return out


@pytest.mark.parametrize('dtype,hi,d', [(np.int8, 11, 40), (np.uint8, 15, 33), (np.int16, 181, 64), (np.uint16, 255, 17),
                                        (np.int32, 4000, 96), (np.int64, 100000, 24), (np.uint32, 60000, 8),
                                        (np.uint64, 3000000, 5)])
def test_integer_matrix_in_the_no_wrap_range_matches_the_oracle(dtype, hi, d):
    rng = np.random.default_rng(int(hi) + d)
    n, q, k = 20000, 300, 7
    lo = -3 if np.issubdtype(dtype, np.signedinteger) else 2        # a range of width <= L, not anchored at zero
    train = np.asfortranarray(rng.integers(lo, lo + hi + 1, (n, d)).astype(dtype))
    test = rng.integers(lo, lo + hi + 1, (q, d)).astype(dtype)
    labels = rng.integers(0, 5, n)
    got = _both_paths(train, labels, test, k)
    pred, idx, dist, stats = got['tensor']
    assert dist.dtype == np.float64
    parity.assert_matches_oracle(pred, idx, dist, train, labels, test, k, 1, 'classification')
    d_pred, d_idx, d_dist, _ = got['direct']
    assert np.array_equal(idx, d_idx) and np.array_equal(dist, d_dist) and np.array_equal(pred, d_pred)
    assert stats['tensor_queries'] == q          # the filter served the call (whatever it could not prove went to K2)


def test_tie_light_integer_data_is_proven_by_the_filter():
    """Wide value range, many features: exact distance ties are rare, so the completeness proof goes through and the
    direct kernel is hardly needed."""
    rng = np.random.default_rng(5)
    n, d, q, k = 60000, 64, 500, 10
    train = np.asfortranarray(rng.integers(0, 40001, (n, d)).astype(np.int32))
    test = rng.integers(0, 40001, (q, d)).astype(np.int32)
    labels = rng.standard_normal(n)
    ix = _native.Index(train, labels, k, 1, 1e-8)            # auto: n >= 4096 and the range cannot wrap -> filter
    pred, idx, dist = ix.predict(test, _native.REGRESSION, return_neighbours=True)
    stats = ix.stats()
    ix.close()
    parity.assert_matches_oracle(pred, idx, dist, train, labels, test, k, 1, 'regression')
    assert stats['tensor_queries'] == q and stats['fallback_queries'] <= q // 10, stats


def test_one_hot_rows_like_mushroom_stay_exact_through_the_ties():
    rng = np.random.default_rng(6)
    levels = rng.integers(2, 7, 22)
    def onehot(rows):
        return np.concatenate([np.eye(lv, dtype=np.int8)[rng.integers(0, lv, rows)] for lv in levels], axis=1)
    train, test = np.asfortranarray(onehot(20000)), onehot(200)
    labels = rng.integers(0, 2, 20000).astype(np.int8)
    got = _both_paths(train, labels, test, 7)
    pred, idx, dist, stats = got['tensor']
    parity.assert_matches_oracle(pred, idx, dist, train, labels, test, 7, 1, 'classification')
    assert np.array_equal(idx, got['direct'][1])
    # every query has exact distance ties; whatever the first filter pass cannot prove is settled by the finer second
    # pass (it keeps enough candidates to hold whole tie groups) or by K2 -- the result is the oracle's either way
    assert stats['tensor_queries'] == 200 and stats['retry_queries'] + stats['fallback_queries'] >= 0


def test_query_outside_the_no_wrap_range_goes_to_the_direct_kernel():
    """int8 train values in 0..10; three queries carry values that make numpy's int8 subtract / square wrap.  They are
    flagged for K2 (which reproduces the wrap-around); all the others stay on the filter."""
    rng = np.random.default_rng(7)
    n, d, q, k = 20000, 16, 64, 5
    train = np.asfortranarray(rng.integers(0, 11, (n, d)).astype(np.int8))
    test = rng.integers(0, 11, (q, d)).astype(np.int8)
    test[3, 2], test[17, 0], test[40, 9] = -100, 127, 23
    labels = rng.integers(0, 3, n)
    got = _both_paths(train, labels, test, k)
    pred, idx, dist, stats = got['tensor']
    with np.errstate(all='ignore'):
        parity.assert_matches_oracle(pred, idx, dist, train, labels, test, k, 1, 'classification')
    assert np.array_equal(idx, got['direct'][1]) and np.array_equal(dist, got['direct'][2], equal_nan=True)
    assert stats['fallback_queries'] >= 3


def test_range_that_can_wrap_keeps_the_index_on_the_direct_kernel():
    rng = np.random.default_rng(8)
    n, d = 20000, 12
    train = np.asfortranarray(rng.integers(-128, 128, (n, d)).astype(np.int8))      # |t - q| up to 255: wraps
    labels = rng.integers(0, 3, n)
    ix = _native.Index(train, labels, 5, 1, 1e-8)
    with pytest.raises(NotImplementedError):
        ix.set_option('path', _native.PATH_TENSOR)
    test = rng.integers(-128, 128, (50, d)).astype(np.int8)
    pred, idx, dist = ix.predict(test, _native.CLASSIFICATION, return_neighbours=True)     # auto -> K2
    assert ix.stats()['tensor_queries'] == 0
    ix.close()
    with np.errstate(all='ignore'):
        parity.assert_matches_oracle(pred, idx, dist, train, labels, test, 5, 1, 'classification')
