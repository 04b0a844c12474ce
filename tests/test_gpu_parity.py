"""GPU parity tests proper: the CUDA path (through the Python boundary and the C ABI) against the oracle and the
golden vectors of the real reference.  Run on the B200 box with ``pytest -m gpu``."""
import copy
import pickle
import zlib

import numpy as np
import pandas as pd
import pytest

import osml_b200 as ob
from osml_b200 import _native
from tests import parity

pytestmark = pytest.mark.gpu


def _model(task, k, gamma, path='direct'):
    cls = ob.KNearestNeighborsRegression if task == 'regression' else ob.KNearestNeighborsClassification
    m = cls(k=k, rbf_gamma=gamma)
    m.path = path
    return m


def _frames(g):
    train_df = pd.DataFrame(g['train'])
    test_df = pd.DataFrame(g['test'])
    labels = pd.Series(g['train_labels'], name='label')
    return train_df, test_df, labels


def test_golden_through_the_model_api(golden):
    """fit(DataFrame, Series) / predict(DataFrame) on the reference's own data sets (BASELINE configs 1-2)."""
    g = golden
    if not bool(g['train_is_f_order']):
        pytest.skip('C-ordered matrices cannot come out of a pandas 3 DataFrame; covered through the C ABI below')
    train_df, test_df, labels = _frames(g)
    assert _native.sum_order_of(train_df.values) == _native.SUM_SEQUENTIAL
    m = _model(g['task'], g['k'], g['rbf_gamma'])
    m.fit(train_df, labels)
    pred = m.predict(test_df)
    assert pred.dtype == labels.dtype and pred.name == 'label' and len(pred) == len(test_df)
    idx, dist = m.kneighbors(test_df)
    parity.assert_matches_oracle(pred.values, idx, dist, train_df.values, labels.values, test_df.values, g['k'],
                                 g['rbf_gamma'], g['task'])
    parity.assert_matches_reference_golden(pred.values, idx, dist, g)
    if g['test_labels'].size:
        metric = m.evaluate(pred, pd.Series(g['test_labels'], name='label'))
        if not g['any_tie'].any():
            assert metric == pytest.approx(float(g['ref_metric']), rel=1e-12)


def test_golden_through_the_c_abi_both_layouts(golden):
    """Same data straight through osml_knn_create / osml_knn_predict, in the layout the reference saw."""
    g = golden
    task = _native.REGRESSION if g['task'] == 'regression' else _native.CLASSIFICATION
    ix = _native.Index(g['train'], g['train_labels'], g['k'], g['rbf_gamma'], 1e-8)
    ix.set_option('path', _native.PATH_DIRECT)
    pred, idx, dist = ix.predict(g['test'], task, return_neighbours=True)
    parity.assert_matches_oracle(pred, idx, dist, g['train'], g['train_labels'], g['test'], g['k'], g['rbf_gamma'],
                                 g['task'])
    parity.assert_matches_reference_golden(pred, idx, dist, g)
    assert ix.stats()['kernel_launches'] >= 2
    ix.close()


@pytest.mark.parametrize('dtype', ['float32', 'float64', 'int8', 'int16', 'int32', 'int64', 'uint8', 'uint16',
                                   'uint32', 'uint64'])
@pytest.mark.parametrize('order', ['F', 'C'])
def test_random_shapes_every_dtype(dtype, order):
    rng = np.random.default_rng(zlib.crc32((dtype + order).encode()))
    for n, d, q, k in [(40, 1, 5, 3), (300, 7, 33, 7), (1000, 16, 130, 1), (2500, 37, 64, 10), (700, 150, 20, 32),
                       (5000, 300, 12, 5), (129, 17, 257, 63)]:
        if np.issubdtype(np.dtype(dtype), np.floating):
            train = (rng.standard_normal((n, d)) * 2).astype(dtype)
            test = (rng.standard_normal((q, d)) * 2).astype(dtype)
        else:
            info = np.iinfo(dtype)
            lo, hi = (max(info.min, -300), min(info.max, 300)) if d > 20 else (info.min, info.max)
            train = rng.integers(lo, hi, (n, d), dtype=dtype, endpoint=True)
            test = rng.integers(lo, hi, (q, d), dtype=dtype, endpoint=True)
        train = np.asfortranarray(train) if order == 'F' else np.ascontiguousarray(train)
        for task, labels in (('classification', rng.integers(-3, 4, n)), ('regression', rng.standard_normal(n))):
            ix = _native.Index(train, labels, k, 0.75, 1e-8)
            ix.set_option('path', _native.PATH_DIRECT)
            code = _native.REGRESSION if task == 'regression' else _native.CLASSIFICATION
            pred, idx, dist = ix.predict(test, code, return_neighbours=True)
            parity.assert_matches_oracle(pred, idx, dist, train, labels, test, k, 0.75, task)
            ix.close()


def test_ties_everywhere_binary_int8():
    """One-hot-like data: every query has boundary ties (SURVEY.md tie census); the (distance, index) rule decides."""
    rng = np.random.default_rng(5)
    train = np.asfortranarray(rng.integers(0, 2, (4000, 24)).astype(np.int8))
    test = rng.integers(0, 2, (200, 24)).astype(np.int8)
    labels = rng.integers(0, 2, 4000).astype(np.int8)
    ix = _native.Index(train, labels, 7, 1, 1e-8)
    pred, idx, dist = ix.predict(test, _native.CLASSIFICATION, return_neighbours=True)
    parity.assert_matches_oracle(pred, idx, dist, train, labels, test, 7, 1, 'classification')
    ix.close()


def test_nan_and_inf_features():
    rng = np.random.default_rng(6)
    train = rng.standard_normal((500, 5)).astype(np.float32)
    train[rng.integers(0, 500, 40), rng.integers(0, 5, 40)] = np.nan
    train[rng.integers(0, 500, 10), rng.integers(0, 5, 10)] = np.inf
    train = np.asfortranarray(train)
    test = rng.standard_normal((30, 5)).astype(np.float32)
    test[3, 2] = np.inf
    labels = rng.integers(0, 3, 500)
    ix = _native.Index(train, labels, 7, 1, 1e-8)
    pred, idx, dist = ix.predict(test, _native.CLASSIFICATION, return_neighbours=True)
    parity.assert_matches_oracle(pred, idx, dist, train, labels, test, 7, 1, 'classification')
    ix.close()


def test_error_convention_matches_the_reference():
    X = pd.DataFrame(np.random.default_rng(0).random((8, 3)))
    y = pd.Series(np.arange(8) % 2, name='y')
    with pytest.raises(ValueError):
        ob.KNearestNeighborsClassification(k=8).fit(X, y)           # N < k + 1      (predict.py:768-769)
    m = ob.KNearestNeighborsClassification(k=7)
    m.fit(X, y)                                                      # N == k + 1 fits ...
    with pytest.raises(ValueError):
        m.predict(X)                                                 # ... but predict raises
    m = ob.KNearestNeighborsClassification(k=3)
    m.fit(X, y)
    with pytest.raises(ValueError):
        m.predict(X.astype(np.float32))                              # dtype vector differs from fit
    with pytest.raises(ValueError):
        m.predict(X.iloc[:0])                                        # empty frame
    with pytest.raises(ValueError):
        ob.KNearestNeighborsRegression(k=3).fit(X, y)                # integer labels on a regression model
    assert m.test(X, y) is None


def test_non_range_index_scatters_like_the_reference():
    rng = np.random.default_rng(1)
    X = pd.DataFrame(rng.random((50, 4)))
    y = pd.Series(rng.integers(0, 3, 50), name='y')
    m = ob.KNearestNeighborsClassification(k=5)
    m.fit(X, y)
    Q = pd.DataFrame(rng.random((6, 4)))
    base = m.predict(Q).values
    shuffled = Q.copy()
    shuffled.index = [5, 4, 3, 2, 1, 0]
    assert np.array_equal(m.predict(shuffled).values, base[::-1])
    bad = Q.copy()
    bad.index = [0, 1, 2, 3, 4, 17]
    with pytest.raises(IndexError):
        m.predict(bad)


def test_pickle_and_deepcopy_round_trip(tmp_path):
    rng = np.random.default_rng(2)
    X = pd.DataFrame(rng.standard_normal((300, 6)).astype(np.float32))
    y = pd.Series(rng.standard_normal(300), name='target')
    Q = pd.DataFrame(rng.standard_normal((40, 6)).astype(np.float32))
    m = ob.KNearestNeighborsRegression(k=9, rbf_gamma=0.5)
    m.fit(X, y)
    expect = m.predict(Q)
    clone = copy.deepcopy(m)
    assert clone._index is None
    assert clone.predict(Q).equals(expect)
    again = pickle.loads(pickle.dumps(m))
    assert again.predict(Q).equals(expect)
    ob.save(m, str(tmp_path / 'm.pkl'))
    assert ob.load(str(tmp_path / 'm.pkl')).predict(Q).equals(expect)


def test_train_sharded_merge_equals_unsharded_bit_for_bit():
    """shards produce exact partial lists with global indices; the merge is exact."""
    import torch
    rng = np.random.default_rng(3)
    n, d, q, k = 6000, 24, 150, 10
    train = rng.standard_normal((n, d)).astype(np.float32)
    test = rng.standard_normal((q, d)).astype(np.float32)
    labels = rng.integers(0, 10, n)
    whole = _native.Index(train, labels, k, 1, 1e-8)
    w_pred, w_idx, w_dist = whole.predict(test, _native.CLASSIFICATION, return_neighbours=True)
    parts = 3
    bounds = [0, 1500, 1507, n]
    dq = torch.from_numpy(test).cuda()
    part_idx = torch.empty((parts, q, k + 1), dtype=torch.int64, device='cuda')
    part_dist = torch.empty((parts, q, k + 1), dtype=torch.float32, device='cuda')
    shards = []
    for p in range(parts):
        sh = _native.Index(train[bounds[p]:bounds[p + 1]], labels, k, 1, 1e-8, index_base=bounds[p])
        sh.neighbours_device(dq.data_ptr(), q, part_idx[p].data_ptr(), part_dist[p].data_ptr())
        shards.append(sh)
    torch.cuda.synchronize()
    out = torch.empty(q, dtype=torch.float64, device='cuda')
    o_idx = torch.empty((q, k + 1), dtype=torch.int64, device='cuda')
    o_dist = torch.empty((q, k + 1), dtype=torch.float32, device='cuda')
    shards[0].merge_predict_device(part_idx.data_ptr(), part_dist.data_ptr(), parts, q, _native.CLASSIFICATION,
                                   out.data_ptr(), o_idx.data_ptr(), o_dist.data_ptr())
    torch.cuda.synchronize()
    assert np.array_equal(o_idx.cpu().numpy(), w_idx)
    assert np.array_equal(o_dist.cpu().numpy(), w_dist)
    assert np.array_equal(out.cpu().numpy(), w_pred)
    for sh in shards:
        sh.close()
    whole.close()
