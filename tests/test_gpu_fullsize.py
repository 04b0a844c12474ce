"""Parity at BASELINE.json's full sizes through properties that do not need the (far too slow) oracle on every query:
the tensor path against the exact direct kernel on a sample, the oracle itself on a handful, sortedness, determinism,
and a checksum that the host-buffer and device-buffer entry points agree on.  Everything goes through the C ABI."""
import hashlib

import numpy as np
import pytest

from oracle import knn_oracle as ko
from osml_b200 import _native

pytestmark = pytest.mark.gpu


def _full_size_case(n, d, q, k, seed, sample, oracle_sample):
    rng = np.random.default_rng(seed)
    train = np.asfortranarray(rng.standard_normal((n, d), dtype=np.float32))      # DataFrame.values is F-ordered
    labels = rng.integers(0, 10, n)
    queries = np.random.default_rng(seed + 1).standard_normal((q, d), dtype=np.float32)
    ix = _native.Index(train, labels, k, 1, 1e-8)
    pred, idx, dist = ix.predict(queries, _native.CLASSIFICATION, return_neighbours=True)
    stats = ix.stats()
    assert stats['tensor_queries'] == q                                    # the tcgen05 path carried the load ...
    assert stats['fallback_queries'] <= q // 1000, stats                   # ... and K2 saw next to nothing
    # every list ascending by (distance, index), indices valid and distinct
    assert np.all(np.diff(dist, axis=1) >= 0)
    ties = np.diff(dist, axis=1) == 0
    assert np.all(np.diff(idx, axis=1)[ties] > 0)
    assert idx.min() >= 0 and idx.max() < n
    assert np.all(np.sort(idx, axis=1)[:, 1:] != np.sort(idx, axis=1)[:, :-1])
    # determinism: a second call is bit-identical (checksum of checksums over predictions, indices, distances)
    pred2, idx2, dist2 = ix.predict(queries, _native.CLASSIFICATION, return_neighbours=True)
    digest = [hashlib.sha256(a.tobytes()).hexdigest() for a in (pred, idx, dist)]
    assert digest == [hashlib.sha256(a.tobytes()).hexdigest() for a in (pred2, idx2, dist2)]
    # the exact direct kernel (no filter at all) on a sample of the same queries
    pick = np.random.default_rng(seed + 2).choice(q, sample, replace=False)
    ix.set_option('path', _native.PATH_DIRECT)
    d_pred, d_idx, d_dist = ix.predict(queries[pick], _native.CLASSIFICATION, return_neighbours=True)
    assert np.array_equal(idx[pick], d_idx) and np.array_equal(dist[pick], d_dist) and np.array_equal(pred[pick], d_pred)
    ix.close()
    # the oracle (numpy restatement of predict.py:740-811) on a handful
    few = pick[:oracle_sample]
    o_pred, o_idx, o_dist = ko.predict_classification(train, labels, queries[few], k=k, return_neighbours=True)
    assert np.array_equal(idx[few], o_idx) and np.array_equal(dist[few], o_dist)
    assert np.array_equal(pred[few], o_pred)


def test_config4_full_size():
    """BASELINE config 4: 1M x 128 fp32 train, 100k queries, k = 10."""
    _full_size_case(1_000_000, 128, 100_000, 10, seed=4, sample=2048, oracle_sample=32)


def test_config3_full_size():
    """BASELINE config 3: 60k x 784 fp32 train, 10k queries, k = 10."""
    _full_size_case(60_000, 784, 10_000, 10, seed=3, sample=2048, oracle_sample=32)


def test_config5_one_shard_full_size():
    """BASELINE config 5, one GPU's share: 2.5M x 96 fp32 rows, k = 32 (a 20k-query sample of the 1M)."""
    _full_size_case(2_500_000, 96, 20_000, 32, seed=5, sample=512, oracle_sample=8)


def test_adversarial_binary_int8_ties_everywhere():
    """SURVEY.md section 8d's adversarial set: rng.integers(0, 2, (100000, 128)).astype(int8) -- exact distance ties on
    every query; int8 arithmetic (subtract and square wrap, int64 sum, fp64 sqrt) on the direct kernel, ties resolved
    by the documented (distance, index) rule."""
    from tests import parity
    rng = np.random.default_rng(8)
    n, d, q, k = 100_000, 128, 120, 10
    train = np.asfortranarray(rng.integers(0, 2, (n, d)).astype(np.int8))
    queries = rng.integers(0, 2, (q, d)).astype(np.int8)
    labels = rng.integers(0, 2, n).astype(np.int8)
    ix = _native.Index(train, labels, k, 1, 1e-8)
    pred, idx, dist = ix.predict(queries, _native.CLASSIFICATION, return_neighbours=True)
    ix.close()
    assert dist.dtype == np.float64
    assert np.all(np.diff(dist, axis=1) >= 0)
    assert (np.diff(dist, axis=1) == 0).any(axis=1).all()                 # every query really has ties
    parity.assert_matches_oracle(pred, idx, dist, train, labels, queries, k, 1, 'classification')
