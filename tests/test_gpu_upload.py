"""Host -> device staging of osml_knn_predict / osml_knn_create (osml_b200/csrc/host_stager.cuh): pageable matrices of
2 MB and more are gathered into a pinned ring by worker threads and copied chunk by chunk.  Whatever the route, the
device must see the same bytes: results are compared bit for bit with the driver's own copy (`stager` = 0), for both
memory orders, short and long columns, ragged sizes, and pinned memory (which never takes the ring)."""
import numpy as np
import pytest

from osml_b200 import _native
from tests import parity

pytestmark = pytest.mark.gpu


def _predict(train, labels, test, k, stager, blocks=0):
    ix = _native.Index(train, labels, k, 1, 1e-8)
    ix.set_option('stager', stager)
    ix.set_option('pipeline_blocks', blocks)
    out = ix.predict(test, _native.CLASSIFICATION, return_neighbours=True)
    again = ix.predict(test, _native.CLASSIFICATION, return_neighbours=True)      # the ring is reused across calls
    assert all(np.array_equal(a, b) for a, b in zip(out, again))
    ix.close()
    return out


@pytest.mark.parametrize('order', ['C', 'F'])
@pytest.mark.parametrize('q,d,n', [(9001, 160, 20000), (3, 200000, 600), (300001, 3, 4096), (5000, 131, 4096)])
def test_staged_upload_equals_the_plain_copy(order, q, d, n):
    """5.8 MB in three blocks; three rows of 800 kB (row pieces / column groups of one); 300 001 rows of 12 bytes (F order:
    columns of 1.2 MB, longer than a ring slot); an odd row length."""
    rng = np.random.default_rng(q + d)
    k = 4
    train = np.asarray(rng.standard_normal((n, d), dtype=np.float32), order='F')
    test = np.asarray(rng.standard_normal((q, d), dtype=np.float32), order=order)
    labels = rng.integers(0, 5, n)
    staged = _predict(train, labels, test, k, 1)
    plain = _predict(train, labels, test, k, 0)
    assert all(np.array_equal(a, b) for a, b in zip(staged, plain))
    sample = rng.choice(q, min(q, 32), replace=False)
    pred, idx, dist = (r[sample] for r in staged)
    parity.assert_matches_oracle(pred, idx, dist, train, labels, np.ascontiguousarray(test[sample]), k, 1, 'classification')


def test_float64_and_int8_matrices_through_the_ring():
    rng = np.random.default_rng(5)
    n, d, q, k = 3000, 40, 70000, 3
    for dtype in (np.float64, np.int8):
        if dtype == np.int8:
            train = np.asfortranarray(rng.integers(0, 12, (n, d)).astype(dtype))
            test = np.asfortranarray(rng.integers(0, 12, (q, d)).astype(dtype))
        else:
            train = np.asfortranarray(rng.standard_normal((n, d)).astype(dtype))
            test = np.asfortranarray(rng.standard_normal((q, d)).astype(dtype))
        labels = rng.integers(0, 3, n)
        out = []
        for stager in (1, 0):
            ix = _native.Index(train, labels, k, 1, 1e-8)
            ix.set_option('stager', stager)
            out.append(ix.predict(test, _native.CLASSIFICATION, return_neighbours=True))
            ix.close()
        assert all(np.array_equal(a, b) for a, b in zip(*out))


def test_pinned_queries_do_not_take_the_ring_and_agree():
    import torch
    rng = np.random.default_rng(9)
    n, d, q, k = 8192, 160, 9001, 5
    train = np.asfortranarray(rng.standard_normal((n, d), dtype=np.float32))
    labels = rng.integers(0, 4, n)
    pageable = rng.standard_normal((q, d), dtype=np.float32)
    pinned = torch.from_numpy(pageable.copy()).pin_memory().numpy()
    ix = _native.Index(train, labels, k, 1, 1e-8)
    ix.set_option('path', _native.PATH_TENSOR)
    a = ix.predict(pageable, _native.CLASSIFICATION, return_neighbours=True)
    assert ix.stats()['last_filter_launches'] == 3          # pageable: three blocks
    b = ix.predict(pinned, _native.CLASSIFICATION, return_neighbours=True)
    assert ix.stats()['last_filter_launches'] == 2          # pinned: two
    ix.close()
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
