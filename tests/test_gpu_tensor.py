"""GPU parity of the tcgen05 filter path (K0 + K1 + K3, a finer second filter pass and finally K2 for unproven
queries) against the oracle.

The path is exact by construction (completeness proof or fallback), so besides bit-parity these tests assert that the
filter actually carries the load: the share of queries that had to be re-run on the direct kernel must stay small."""
import numpy as np
import pytest

from osml_b200 import _native
from tests import parity

pytestmark = pytest.mark.gpu


def _run(train, labels, test, k, task='classification', gamma=1, slack=None, segments=None):
    ix = _native.Index(train, labels, k, gamma, 1e-8)
    ix.set_option('path', _native.PATH_TENSOR)
    if slack is not None:
        ix.set_option('slack', slack)
    if segments is not None:
        ix.set_option('segments', segments)
    code = _native.REGRESSION if task == 'regression' else _native.CLASSIFICATION
    pred, idx, dist = ix.predict(test, code, return_neighbours=True)
    stats = ix.stats()
    ix.close()
    return pred, idx, dist, stats


@pytest.mark.parametrize('n,d,q,k', [
    (20000, 128, 1000, 10),      # config-4 shape, scaled down
    (5000, 96, 300, 32),         # config-5 feature width and k
    (777, 1, 130, 3), (3001, 3, 257, 7), (4096, 17, 128, 1), (10000, 61, 64, 10), (10000, 62, 64, 10),
    (2500, 200, 100, 5), (3000, 317, 50, 10),
])
def test_gaussian_matches_oracle_bit_for_bit(n, d, q, k):
    rng = np.random.default_rng(n + d)
    train = np.asfortranarray(rng.standard_normal((n, d), dtype=np.float32))
    test = rng.standard_normal((q, d), dtype=np.float32)
    labels = rng.integers(0, 10, n)
    pred, idx, dist, stats = _run(train, labels, test, k)
    parity.assert_matches_oracle(pred, idx, dist, train, labels, test, k, 1, 'classification')
    assert stats['tensor_queries'] == q
    if d >= 16:
        assert stats['fallback_queries'] <= max(2, q // 20), stats


def test_regression_and_pairwise_order():
    rng = np.random.default_rng(9)
    train = rng.standard_normal((6000, 150), dtype=np.float32)          # C order -> numpy pairwise summation
    test = rng.standard_normal((200, 150), dtype=np.float32)
    labels = rng.standard_normal(6000)
    pred, idx, dist, stats = _run(train, labels, test, 10, task='regression', gamma=0.5)
    parity.assert_matches_oracle(pred, idx, dist, train, labels, test, 10, 0.5, 'regression')
    assert stats['fallback_queries'] <= 10, stats


def test_uncentred_data_is_centred_by_k0():
    rng = np.random.default_rng(10)
    train = np.asfortranarray((rng.standard_normal((20000, 64)) + 50.0).astype(np.float32))
    test = (rng.standard_normal((256, 64)) + 50.0).astype(np.float32)
    labels = rng.integers(0, 4, 20000)
    pred, idx, dist, stats = _run(train, labels, test, 10)
    parity.assert_matches_oracle(pred, idx, dist, train, labels, test, 10, 1, 'classification')
    assert stats['fallback_queries'] <= 26, stats


def test_duplicates_and_ties_fall_back_but_stay_exact():
    rng = np.random.default_rng(11)
    base = rng.integers(0, 2, (20, 32)).astype(np.float32)
    train = np.asfortranarray(base[rng.integers(0, 20, 8000)])           # every row duplicated ~400 times
    test = base[rng.integers(0, 20, 150)]
    labels = rng.integers(0, 3, 8000)
    pred, idx, dist, stats = _run(train, labels, test, 10)
    parity.assert_matches_oracle(pred, idx, dist, train, labels, test, 10, 1, 'classification')
    assert stats['fallback_queries'] > 0


def test_multi_segment_sweep():
    rng = np.random.default_rng(12)
    n, d, q, k = 300000, 32, 700, 10
    train = np.asfortranarray(rng.standard_normal((n, d), dtype=np.float32))
    test = rng.standard_normal((q, d), dtype=np.float32)
    labels = rng.integers(0, 10, n)
    pred, idx, dist, stats = _run(train, labels, test, k)
    d_pred, d_idx, d_dist = None, None, None
    ix = _native.Index(train, labels, k, 1, 1e-8)
    ix.set_option('path', _native.PATH_DIRECT)
    d_pred, d_idx, d_dist = ix.predict(test, _native.CLASSIFICATION, return_neighbours=True)
    ix.close()
    assert np.array_equal(idx, d_idx) and np.array_equal(dist, d_dist) and np.array_equal(pred, d_pred)
    sub = slice(0, 40)
    parity.assert_matches_oracle(pred[sub], idx[sub], dist[sub], train, labels, test[sub], k, 1, 'classification')
    assert stats['fallback_queries'] <= q // 20, stats


def test_nonfinite_queries_and_train_rows():
    rng = np.random.default_rng(13)
    train = np.asfortranarray(rng.standard_normal((4000, 40), dtype=np.float32))
    test = rng.standard_normal((140, 40), dtype=np.float32)
    test[5, 3] = np.nan
    test[77, 0] = np.inf
    labels = rng.integers(0, 5, 4000)
    pred, idx, dist, stats = _run(train, labels, test, 7)
    parity.assert_matches_oracle(pred, idx, dist, train, labels, test, 7, 1, 'classification')
    train2 = train.copy(order='F')
    train2[100, 7] = np.inf
    pred, idx, dist, stats = _run(train2, labels, test, 7)
    parity.assert_matches_oracle(pred, idx, dist, train2, labels, test, 7, 1, 'classification')


def test_wide_rows_stream_the_query_tile():
    """MNIST-shaped rows (BASELINE config 3: d = 784): the query tile no longer fits in shared memory next to the ring,
    so its K blocks are streamed with the train blocks."""
    rng = np.random.default_rng(14)
    n, d, q, k = 6000, 784, 300, 10
    train = np.asfortranarray(rng.standard_normal((n, d), dtype=np.float32))
    test = rng.standard_normal((q, d), dtype=np.float32)
    labels = rng.integers(0, 10, n)
    pred, idx, dist, stats = _run(train, labels, test, k)
    parity.assert_matches_oracle(pred, idx, dist, train, labels, test, k, 1, 'classification')
    assert stats['fallback_queries'] <= q // 20, stats


@pytest.mark.parametrize('k', [14, 15, 20, 37, 38, 45])
def test_list_length_plans(k):
    """k decides between 16- and 32-entry register lists and the shared-memory lists."""
    rng = np.random.default_rng(100 + k)
    train = np.asfortranarray(rng.standard_normal((9000, 24), dtype=np.float32))
    test = rng.standard_normal((200, 24), dtype=np.float32)
    labels = rng.standard_normal(9000)
    pred, idx, dist, stats = _run(train, labels, test, k, task='regression')
    parity.assert_matches_oracle(pred, idx, dist, train, labels, test, k, 1, 'regression')
    assert stats['fallback_queries'] <= 20, stats


def test_single_sweep_items_with_a_split_tail():
    """More query-tile pairs than CTA pairs on the machine: the first full waves are swept in one item each, the
    remainder is cut into interleaved segments (tc_pair_item).  Bit-identical to the direct kernel."""
    rng = np.random.default_rng(21)
    n, d, q, k = 20000, 16, 256 * 80 + 37, 5          # 81 pairs on 74 clusters: 74 single sweeps + 7 split pairs
    train = np.asfortranarray(rng.standard_normal((n, d), dtype=np.float32))
    test = rng.standard_normal((q, d), dtype=np.float32)
    labels = rng.integers(0, 4, n)
    pred, idx, dist, stats = _run(train, labels, test, k)
    ix = _native.Index(train, labels, k, 1, 1e-8)
    ix.set_option('path', _native.PATH_DIRECT)
    d_pred, d_idx, d_dist = ix.predict(test, _native.CLASSIFICATION, return_neighbours=True)
    ix.close()
    assert np.array_equal(idx, d_idx) and np.array_equal(dist, d_dist) and np.array_equal(pred, d_pred)
    assert stats['fallback_queries'] <= q // 50, stats


@pytest.mark.parametrize('segments', [1, 2, 5, 16])
def test_forced_segment_counts(segments):
    rng = np.random.default_rng(22)
    n, d, q, k = 70000, 20, 300, 9
    train = np.asfortranarray(rng.standard_normal((n, d), dtype=np.float32))
    test = rng.standard_normal((q, d), dtype=np.float32)
    labels = rng.integers(0, 4, n)
    pred, idx, dist, stats = _run(train, labels, test, k, segments=segments)
    parity.assert_matches_oracle(pred, idx, dist, train, labels, test, k, 1, 'classification')
    assert stats['fallback_queries'] <= q // 20, stats


def test_unproven_queries_take_the_second_filter_pass():
    """Clustered data: many rows at nearly the same distance make the completeness proof fail on the first pass for some
    queries; they must come back exact from the finer second pass (retry_queries), not only from the direct kernel."""
    rng = np.random.default_rng(23)
    centres = rng.standard_normal((30, 48)).astype(np.float32)
    train = np.asfortranarray(centres[rng.integers(0, 30, 60000)] +
                              1e-3 * rng.standard_normal((60000, 48)).astype(np.float32))
    test = centres[rng.integers(0, 30, 400)] + 1e-3 * rng.standard_normal((400, 48)).astype(np.float32)
    labels = rng.integers(0, 5, 60000)
    pred, idx, dist, stats = _run(train, labels, test, 10)
    parity.assert_matches_oracle(pred, idx, dist, train, labels, test, 10, 1, 'classification')
    assert stats['retry_queries'] > 0, stats


def test_rows_with_very_different_norms():
    """The filter copy is sorted by row norm and the hot path tests against thr + min(h): rows whose norms span orders
    of magnitude (and a few all-zero rows) must not disturb exactness."""
    rng = np.random.default_rng(24)
    n, d, q, k = 50000, 33, 500, 10
    scale = np.exp(rng.uniform(-4, 4, (n, 1))).astype(np.float32)
    train = rng.standard_normal((n, d), dtype=np.float32) * scale
    train[rng.integers(0, n, 50)] = 0.0
    train = np.asfortranarray(train)
    test = rng.standard_normal((q, d), dtype=np.float32) * np.exp(rng.uniform(-3, 3, (q, 1))).astype(np.float32)
    labels = rng.integers(0, 7, n)
    pred, idx, dist, stats = _run(train, labels, test, k)
    parity.assert_matches_oracle(pred, idx, dist, train, labels, test, k, 1, 'classification')


def test_random_shapes_tensor_path_equals_direct_kernel():
    """Property test over shapes the fixed cases do not hit: ragged row / query counts, odd feature counts on both
    sides of the 64-column K blocks, k around the list-length plan boundaries, offset and scaled data.  The filter path
    must be bit-identical to the exact direct kernel whatever it keeps or drops."""
    rng = np.random.default_rng(2024)
    for case in range(24):
        n = int(rng.integers(300, 40000))
        d = int(rng.choice([1, 2, 5, 16, 31, 63, 64, 65, 100, 127, 128, 129, 191, 193, 253, 254, 300]))
        q = int(rng.integers(1, 900))
        k = int(rng.choice([1, 2, 5, 7, 10, 13, 14, 21, 22, 32, 40]))
        if n < k + 2:
            continue
        scale = float(np.exp(rng.uniform(-3, 3)))
        train = np.asfortranarray((rng.standard_normal((n, d)) * scale + rng.uniform(-5, 5, d)).astype(np.float32))
        test = (rng.standard_normal((q, d)) * scale + rng.uniform(-5, 5, d)).astype(np.float32)
        labels = rng.standard_normal(n)
        pred, idx, dist, stats = _run(train, labels, test, k, task='regression')
        ix = _native.Index(train, labels, k, 1, 1e-8)
        ix.set_option('path', _native.PATH_DIRECT)
        d_pred, d_idx, d_dist = ix.predict(test, _native.REGRESSION, return_neighbours=True)
        ix.close()
        where = 'case %d: n=%d d=%d q=%d k=%d' % (case, n, d, q, k)
        assert np.array_equal(idx, d_idx), where
        assert np.array_equal(dist, d_dist), where
        assert np.array_equal(pred, d_pred), where


def test_more_queries_than_one_block():
    """The filter processes queries in blocks of 131072: two blocks, the second one ragged."""
    rng = np.random.default_rng(31)
    n, d, q, k = 30000, 16, 131072 + 8929, 5
    train = np.asfortranarray(rng.standard_normal((n, d), dtype=np.float32))
    test = rng.standard_normal((q, d), dtype=np.float32)
    labels = rng.integers(0, 6, n)
    pred, idx, dist, stats = _run(train, labels, test, k)
    ix = _native.Index(train, labels, k, 1, 1e-8)
    ix.set_option('path', _native.PATH_DIRECT)
    d_pred, d_idx, d_dist = ix.predict(test, _native.CLASSIFICATION, return_neighbours=True)
    ix.close()
    assert np.array_equal(idx, d_idx) and np.array_equal(dist, d_dist) and np.array_equal(pred, d_pred)
    assert stats['tensor_queries'] == q


@pytest.mark.parametrize('precision', [1, 2])
def test_both_filter_precisions_are_exact(precision):
    """The filter copy is bf16 (precision=1) or scaled fp16 (precision=2); either way the result is the exact one."""
    rng = np.random.default_rng(41)
    n, d, q, k = 120000, 96, 4000, 32
    train = np.asfortranarray(rng.standard_normal((n, d), dtype=np.float32) * 3 + 7)
    test = rng.standard_normal((q, d), dtype=np.float32) * 3 + 7
    labels = rng.integers(0, 5, n)
    ix = _native.Index(train, labels, k, 1, 1e-8)
    ix.set_option('path', _native.PATH_TENSOR)
    ix.set_option('precision', precision)
    pred, idx, dist = ix.predict(test, _native.CLASSIFICATION, return_neighbours=True)
    stats = ix.stats()
    assert stats['filter_precision'] == precision
    ix.set_option('path', _native.PATH_DIRECT)
    d_pred, d_idx, d_dist = ix.predict(test, _native.CLASSIFICATION, return_neighbours=True)
    ix.close()
    assert np.array_equal(idx, d_idx) and np.array_equal(dist, d_dist) and np.array_equal(pred, d_pred)
    if precision == 2:                       # 8x smaller rounding residuals: (almost) nothing left for the second pass
        assert stats['retry_queries'] <= q // 200, stats


def test_precision_switches_to_fp16_when_bf16_needs_many_second_passes():
    """Auto mode starts with bf16; a call in which more than 0.25 % of the queries needed the second pass makes the next
    call rebuild the filter copy in fp16.  Results are identical before and after."""
    rng = np.random.default_rng(42)
    centres = rng.standard_normal((40, 64)).astype(np.float32) * 30
    train = np.asfortranarray(centres[rng.integers(0, 40, 50000)] +
                              0.05 * rng.standard_normal((50000, 64)).astype(np.float32))
    test = centres[rng.integers(0, 40, 6000)] + 0.05 * rng.standard_normal((6000, 64)).astype(np.float32)
    labels = rng.integers(0, 5, 50000)
    ix = _native.Index(train, labels, 10, 1, 1e-8)
    ix.set_option('path', _native.PATH_TENSOR)
    first = ix.predict(test, _native.CLASSIFICATION, return_neighbours=True)
    stats_first = ix.stats()
    second = ix.predict(test, _native.CLASSIFICATION, return_neighbours=True)
    stats_second = ix.stats()
    ix.close()
    for a, b in zip(first, second):
        assert np.array_equal(a, b)
    assert stats_first['filter_precision'] == 1 and stats_first['retry_queries'] > len(test) // 400, stats_first
    assert stats_second['filter_precision'] == 2, stats_second


@pytest.mark.parametrize('order', ['C', 'F'])
@pytest.mark.parametrize('fused_transform', [False, True])
def test_pipelined_upload_equals_the_single_block_call(order, fused_transform):
    """osml_knn_predict sends a large query matrix up in two blocks (one wave of K1 work items, then the rest on a
    second stream under the first block's K1).  Row- and column-major host buffers, with and without the fused query
    transform, ragged tail: everything must equal the unpipelined call bit for bit."""
    rng = np.random.default_rng(51)
    n, d, q, k = 20000, 24, 2 * 148 * 128 + 3331, 6
    train = np.asfortranarray(rng.standard_normal((n, d), dtype=np.float32))
    test = np.asarray(rng.standard_normal((q, d), dtype=np.float32) * 1.5 + 0.25, order=order)
    labels = rng.integers(0, 7, n)
    results = []
    for pipeline in (1, 0):
        ix = _native.Index(train, labels, k, 1, 1e-8)
        ix.set_option('path', _native.PATH_TENSOR)
        ix.set_option('pipeline', pipeline)
        if fused_transform:
            ix.set_query_transform(np.full(d, 0.25, np.float32), np.full(d, 1.5, np.float32))
        results.append(ix.predict(test, _native.CLASSIFICATION, return_neighbours=True))
        # a second call reuses every buffer and stream
        again = ix.predict(test, _native.CLASSIFICATION, return_neighbours=True)
        assert all(np.array_equal(a, b) for a, b in zip(results[-1], again))
        stats = ix.stats()
        assert stats['tensor_queries'] == 2 * q
        assert stats['last_filter_launches'] == (2 if pipeline else 1)
        ix.close()
    for a, b in zip(*results):
        assert np.array_equal(a, b)
    if not fused_transform:
        sample = rng.choice(q, 64, replace=False)
        pred, idx, dist = (r[sample] for r in results[0])
        parity.assert_matches_oracle(pred, idx, dist, train, labels, np.ascontiguousarray(test[sample]), k, 1,
                                     'classification')


def test_unproven_queries_of_every_block_are_settled_in_one_pass_at_the_end():
    """Three blocks in one call (the pipelined first wave, a full block of 131072, a ragged tail) on clustered data: each
    block leaves unproven queries behind; they are collected with call-wide numbers and re-run once, after the last
    block.  Result = the direct kernel's, bit for bit."""
    rng = np.random.default_rng(61)
    n, d, k = 40000, 32, 8
    q = 148 * 128 + 131072 + 517
    centres = rng.standard_normal((25, d)).astype(np.float32)
    train = np.asfortranarray(centres[rng.integers(0, 25, n)] + 1e-3 * rng.standard_normal((n, d)).astype(np.float32))
    test = centres[rng.integers(0, 25, q)] + 1e-3 * rng.standard_normal((q, d)).astype(np.float32)
    labels = rng.integers(0, 5, n)
    pred, idx, dist, stats = _run(train, labels, test, k)
    assert stats['tensor_queries'] == q and stats['last_filter_launches'] == 3, stats
    assert stats['retry_queries'] > 0, stats
    ix = _native.Index(train, labels, k, 1, 1e-8)
    ix.set_option('path', _native.PATH_DIRECT)
    d_pred, d_idx, d_dist = ix.predict(test, _native.CLASSIFICATION, return_neighbours=True)
    ix.close()
    assert np.array_equal(idx, d_idx) and np.array_equal(dist, d_dist) and np.array_equal(pred, d_pred)


@pytest.mark.parametrize('extra', [-1, 0, 1])
def test_pipelined_upload_boundaries(extra):
    """Two waves of K1 work items (2 x 148 x 128 queries on a B200) is where osml_knn_predict starts to upload in two
    blocks: one query fewer (single block), exactly two waves (the second block is one full wave), one more."""
    rng = np.random.default_rng(71 + extra)
    n, d, k = 4096, 8, 3
    q = 2 * 148 * 128 + extra
    train = np.asfortranarray(rng.standard_normal((n, d), dtype=np.float32))
    test = np.asfortranarray(rng.standard_normal((q, d), dtype=np.float32))
    labels = rng.standard_normal(n)
    ix = _native.Index(train, labels, k, 1, 1e-8)
    ix.set_option('path', _native.PATH_TENSOR)
    pred, idx, dist = ix.predict(test, _native.REGRESSION, return_neighbours=True)
    launches = ix.stats()['last_filter_launches']
    ix.set_option('path', _native.PATH_DIRECT)
    d_pred, d_idx, d_dist = ix.predict(test, _native.REGRESSION, return_neighbours=True)
    ix.close()
    assert np.array_equal(idx, d_idx) and np.array_equal(dist, d_dist) and np.array_equal(pred, d_pred)
    import torch
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    assert launches == (2 if q >= 2 * sms * 128 else 1)


def test_small_calls_upload_in_equal_blocks():
    """Below two waves of queries a call of 4 MB and more is uploaded in equal blocks (two, or what the pipeline_blocks
    option says; every block but the first under the K1 of the block in front of it); smaller calls go up at once.  Same
    bits as the unpipelined call either way."""
    rng = np.random.default_rng(81)
    n, d, k = 8192, 160, 5
    train = np.asfortranarray(rng.standard_normal((n, d), dtype=np.float32))
    labels = rng.integers(0, 4, n)
    # 5.8, 4.5, 2.6, 0.4 MB of pageable queries (three blocks by default; pinned memory: two)
    for q, launches, blocks in ((9001, 3, 0), (9001, 4, 4), (7000, 2, 2), (4100, 1, 0), (700, 1, 4)):
        test = rng.standard_normal((q, d), dtype=np.float32)
        out = []
        for pipeline in (1, 0):
            ix = _native.Index(train, labels, k, 1, 1e-8)
            ix.set_option('path', _native.PATH_TENSOR)
            ix.set_option('pipeline', pipeline)
            ix.set_option('pipeline_blocks', blocks)
            out.append(ix.predict(test, _native.CLASSIFICATION, return_neighbours=True))
            assert ix.stats()['last_filter_launches'] == (launches if pipeline else 1)
            ix.close()
        assert all(np.array_equal(a, b) for a, b in zip(*out))


@pytest.mark.parametrize('n,d,q,k,order', [
    (20000, 128, 600, 10, 'F'), (5000, 96, 300, 32, 'F'), (3001, 3, 257, 7, 'F'), (10000, 61, 64, 10, 'F'),
    (6000, 150, 200, 10, 'C'), (2500, 200, 100, 5, 'C'),
])
def test_float64_data_takes_the_tensor_path_and_stays_bit_exact(n, d, q, k, order):
    """pandas' default dtype.  The 16-bit filter copy is built from the float64 matrix, K3 rescoring runs in float64 in
    numpy's order: distances (float64), indices and labels equal the oracle's bit for bit, as for float32."""
    rng = np.random.default_rng(n + d + 1)
    train = np.asarray(rng.standard_normal((n, d)) * 3.0 + 1.5, order=order)
    test = rng.standard_normal((q, d)) * 3.0 + 1.5
    labels = rng.integers(0, 10, n)
    assert train.dtype == np.float64
    pred, idx, dist, stats = _run(train, labels, test, k)
    assert dist.dtype == np.float64
    parity.assert_matches_oracle(pred, idx, dist, train, labels, test, k, 1, 'classification')
    assert stats['tensor_queries'] == q
    if d >= 16:
        assert stats['fallback_queries'] <= max(2, q // 20), stats


def test_float64_regression_ties_and_second_pass_equal_the_direct_kernel():
    """float64: clustered rows (second filter pass), exact duplicates (direct-kernel fallback), values that differ only
    below float32 resolution (the filter cannot tell them apart, K3 must), regression output -- against K2."""
    rng = np.random.default_rng(91)
    n, d, q, k = 30000, 40, 3000, 9
    centres = rng.standard_normal((20, d))
    train = centres[rng.integers(0, 20, n)] + 1e-3 * rng.standard_normal((n, d))
    train[:2000] = train[2000:4000]                                  # exact duplicates
    train[4000:6000] = train[6000:8000] + 1e-11                      # distinct only in float64
    train = np.asfortranarray(train)
    test = centres[rng.integers(0, 20, q)] + 1e-3 * rng.standard_normal((q, d))
    labels = rng.standard_normal(n)
    pred, idx, dist, stats = _run(train, labels, test, k, task='regression', gamma=0.7)
    assert stats['tensor_queries'] == q and stats['retry_queries'] > 0, stats
    ix = _native.Index(train, labels, k, 0.7, 1e-8)
    ix.set_option('path', _native.PATH_DIRECT)
    d_pred, d_idx, d_dist = ix.predict(test, _native.REGRESSION, return_neighbours=True)
    ix.close()
    assert np.array_equal(idx, d_idx) and np.array_equal(dist, d_dist) and np.array_equal(pred, d_pred)
    sample = rng.choice(q, 48, replace=False)
    parity.assert_matches_oracle(pred[sample], idx[sample], dist[sample], train, labels, test[sample], k, 0.7,
                                 'regression')


def test_float64_through_the_model_api_with_pipelined_upload():
    """DataFrame of float64 columns, more than two waves of queries (two-block upload)."""
    import pandas as pd
    import osml_b200 as ob
    rng = np.random.default_rng(92)
    n, d, q, k = 60000, 24, 2 * 148 * 128 + 777, 7
    train_df = pd.DataFrame(rng.standard_normal((n, d)))
    labels = pd.Series(rng.integers(0, 5, n), name='y')
    query_df = pd.DataFrame(rng.standard_normal((q, d)))
    model = ob.KNearestNeighborsClassification(k=k)
    model.path = 'tensor'
    model.fit(train_df, labels)
    pred = model.predict(query_df)
    stats = model._index.stats()
    assert stats['tensor_queries'] == q and stats['last_filter_launches'] == 2, stats
    model.path = 'direct'
    model._drop_index()
    assert np.array_equal(model.predict(query_df).values, pred.values)


def test_auto_path_builds_the_filter_after_128_queries_and_gives_up_on_tie_ridden_data():
    """path = auto on a fitted index (osml_knn_set_option is never called): the first few queries go to the direct
    kernel, the filter is built once 128 have been asked for and serves every later call, however small
    (tools/crossover.py: it wins from 32 queries up); on data whose exact ties defeat the completeness proof for most
    queries auto returns to the direct kernel after the first such call.  Results are exact either way."""
    rng = np.random.default_rng(101)
    n, d, k = 20000, 32, 5
    train = np.asfortranarray(rng.standard_normal((n, d), dtype=np.float32))
    labels = rng.integers(0, 4, n)
    ix = _native.Index(train, labels, k, 1, 1e-8)
    seen = []
    for q in (50, 100, 10):
        test = rng.standard_normal((q, d), dtype=np.float32)
        pred, idx, dist = ix.predict(test, _native.CLASSIFICATION, return_neighbours=True)
        parity.assert_matches_oracle(pred, idx, dist, train, labels, test, k, 1, 'classification')
        seen.append(ix.stats()['tensor_queries'])
    ix.close()
    assert seen == [0, 100, 110]

    base = rng.integers(0, 2, (12, d)).astype(np.float32)
    train = np.asfortranarray(base[rng.integers(0, 12, n)])              # every row repeated ~1700 times
    ix = _native.Index(train, labels, k, 1, 1e-8)
    test = base[rng.integers(0, 12, 300)]
    first = ix.predict(test, _native.CLASSIFICATION, return_neighbours=True)
    after_first = ix.stats()
    second = ix.predict(test, _native.CLASSIFICATION, return_neighbours=True)
    after_second = ix.stats()
    ix.close()
    assert after_first['tensor_queries'] == 300 and after_first['fallback_queries'] > 150, after_first
    assert after_second['tensor_queries'] == 300, after_second                  # the second call never saw the filter
    assert all(np.array_equal(a, b) for a, b in zip(first, second))
    parity.assert_matches_oracle(first[0][:40], first[1][:40], first[2][:40], train, labels, test[:40], k, 1,
                                 'classification')


# ---- adversarial inputs for the completeness bound ("argued in comments, not tested adversarially") --------
# The filter's error budget (TcBounds: c_mma, et_max, the 2^-20 tie margin) decides which queries may skip the exact
# kernel.  Whatever the data does to it, the answer must stay the oracle's: a budget that is too optimistic shows up
# here as a wrong neighbour, one that is merely conservative as extra fallbacks.
def test_near_ties_one_ulp_apart_on_a_ray():
    """Train rows on a ray from the origin, radii one float32 ulp apart: hundreds of rows whose distances to the query
    differ in the last bit only, far below what a bf16 score can resolve."""
    rng = np.random.default_rng(21)
    n, d, k = 20000, 48, 10
    direction = rng.standard_normal(d)
    direction /= np.linalg.norm(direction)
    radii = (1.0 + np.arange(n, dtype=np.float64) * 2.0 ** -23) * 3.0
    rng.shuffle(radii)
    train = np.asfortranarray((radii[:, None] * direction[None, :]).astype(np.float32))
    test = np.concatenate([np.zeros((1, d), np.float32), (rng.standard_normal((40, d)) * 1e-4).astype(np.float32),
                           (direction[None, :] * radii[:23, None]).astype(np.float32)])
    labels = rng.integers(0, 4, n)
    pred, idx, dist, stats = _run(train, labels, test, k)
    parity.assert_matches_oracle(pred, idx, dist, train, labels, test, k, 1, 'classification')


def test_two_distant_clusters_defeat_the_centring():
    """Two clusters 2000 apart in every feature: centring leaves +-1000, whose bf16 rounding error (~4 per feature)
    dwarfs the within-cluster distances (~10).  The bound must notice and send the queries to the exact kernel."""
    rng = np.random.default_rng(22)
    n, d, k = 20000, 32, 7
    side = rng.integers(0, 2, n) * 2.0 - 1.0
    train = np.asfortranarray((side[:, None] * 1000.0 + rng.standard_normal((n, d))).astype(np.float32))
    test = ((rng.integers(0, 2, 120) * 2.0 - 1.0)[:, None] * 1000.0 + rng.standard_normal((120, d))).astype(np.float32)
    labels = rng.standard_normal(n)
    pred, idx, dist, stats = _run(train, labels, test, k, task='regression')
    parity.assert_matches_oracle(pred, idx, dist, train, labels, test, k, 1, 'regression')


def test_feature_scales_twelve_orders_of_magnitude_apart():
    rng = np.random.default_rng(23)
    n, d, k = 30000, 24, 10
    scale = (10.0 ** rng.uniform(-6, 6, d)).astype(np.float32)
    train = np.asfortranarray(rng.standard_normal((n, d), dtype=np.float32) * scale)
    test = rng.standard_normal((200, d), dtype=np.float32) * scale
    labels = rng.integers(0, 3, n)
    pred, idx, dist, stats = _run(train, labels, test, k)
    parity.assert_matches_oracle(pred, idx, dist, train, labels, test, k, 1, 'classification')


def test_huge_norms_tiny_gaps_float64():
    """float64 rows of norm ~1e8 whose mutual distances are ~1: the filter copy cannot see any of it."""
    rng = np.random.default_rng(24)
    n, d, k = 20000, 16, 5
    centre = rng.standard_normal(d) * 1e8
    train = np.asfortranarray(centre[None, :] + rng.standard_normal((n, d)))
    test = centre[None, :] + rng.standard_normal((64, d))
    labels = rng.integers(0, 3, n)
    pred, idx, dist, stats = _run(train, labels, test, k)
    assert dist.dtype == np.float64
    parity.assert_matches_oracle(pred, idx, dist, train, labels, test, k, 1, 'classification')


def test_queries_far_outside_the_train_range_and_on_top_of_train_rows():
    rng = np.random.default_rng(25)
    n, d, k = 25000, 40, 10
    train = np.asfortranarray(rng.standard_normal((n, d), dtype=np.float32))
    test = np.concatenate([train[:30] * 1.0, train[30:60] + np.float32(1e-6),                  # exact hits, near hits
                           rng.standard_normal((30, d), dtype=np.float32) * np.float32(1e4),    # far outside
                           np.full((2, d), 3e37, np.float32)]).astype(np.float32)                # squares overflow to inf
    labels = rng.integers(0, 5, n)
    pred, idx, dist, stats = _run(train, labels, test, k)
    with np.errstate(all='ignore'):
        parity.assert_matches_oracle(pred, idx, dist, train, labels, test, k, 1, 'classification')


@pytest.mark.parametrize('k,n,d', [(64, 40000, 24), (100, 50000, 32), (250, 60000, 16)])
def test_large_k_runs_on_the_filter_with_forced_segments(k, n, d):
    """k beyond what two lists of 32 per query hold: register lists of 32 per slot and enough interleaved train segments
    that the slots hold k + 1 + slack candidates between them (tc_plan_lists).  Same bits as the oracle; the direct
    kernel sees next to nothing."""
    rng = np.random.default_rng(k)
    q = 300
    train = np.asfortranarray(rng.standard_normal((n, d), dtype=np.float32))
    test = rng.standard_normal((q, d), dtype=np.float32)
    labels = rng.integers(0, 6, n)
    pred, idx, dist, stats = _run(train, labels, test, k)
    parity.assert_matches_oracle(pred, idx, dist, train, labels, test, k, 1, 'classification')
    assert stats['tensor_queries'] == q and stats['fallback_queries'] <= q // 10, stats


def test_k_beyond_the_filter_stays_on_the_direct_kernel():
    rng = np.random.default_rng(77)
    n, d, k = 30000, 8, 300
    train = np.asfortranarray(rng.standard_normal((n, d), dtype=np.float32))
    labels = rng.standard_normal(n)
    ix = _native.Index(train, labels, k, 1, 1e-8)
    with pytest.raises(NotImplementedError):
        ix.set_option('path', _native.PATH_TENSOR)
    test = rng.standard_normal((40, d), dtype=np.float32)
    pred, idx, dist = ix.predict(test, _native.REGRESSION, return_neighbours=True)
    assert ix.stats()['tensor_queries'] == 0
    ix.close()
    parity.assert_matches_oracle(pred, idx, dist, train, labels, test, k, 1, 'regression')


@pytest.mark.parametrize('n,d,q,k', [(4096, 128, 1500, 10), (5000, 20, 300, 7), (9000, 300, 700, 10)])
def test_auto_takes_the_filter_from_4096_rows_on_and_small_calls_use_one_wave_of_segments(n, d, q, k):
    """path = auto: a fitted float index of 4096 rows or more answers through the tensor filter once 128 queries have
    been asked for (smaller ones stay on the direct kernel).  Calls of a few query-tile pairs are cut into as many
    segments as fit one wave of work items (up to 32 for one or two pairs); the result is the direct kernel's, bit for
    bit, and the oracle's."""
    rng = np.random.default_rng(n + q)
    train = np.asfortranarray(rng.standard_normal((n, d), dtype=np.float32))
    test = rng.standard_normal((q, d), dtype=np.float32)
    labels = rng.integers(0, 6, n)
    ix = _native.Index(train, labels, k, 1, 1e-8)
    auto = ix.predict(test, _native.CLASSIFICATION, return_neighbours=True)
    assert ix.stats()['tensor_queries'] == q
    few = ix.predict(test[:40], _native.CLASSIFICATION, return_neighbours=True)          # one pair: the most segments
    ix.set_option('path', _native.PATH_DIRECT)
    direct = ix.predict(test, _native.CLASSIFICATION, return_neighbours=True)
    ix.close()
    assert all(np.array_equal(a, b) for a, b in zip(auto, direct))
    assert all(np.array_equal(a, b[:40]) for a, b in zip(few, direct))
    sample = rng.choice(q, 48, replace=False)
    parity.assert_matches_oracle(auto[0][sample], auto[1][sample], auto[2][sample], train, labels,
                                 np.ascontiguousarray(test[sample]), k, 1, 'classification')
    small = _native.Index(train[:4000], labels[:4000], k, 1, 1e-8)                        # below the threshold: direct
    small.predict(test, _native.CLASSIFICATION)
    assert small.stats()['tensor_queries'] == 0
    small.close()
