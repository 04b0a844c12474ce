"""Tie-aware comparison of the CUDA path with the oracle (SURVEY.md section 8d "parity gates").

Rules (north_star): neighbour indices and integer class predictions bit-exact; distances and regression outputs within
1e-5 relative (we assert far tighter: distances bit-exact, regression 1e-12, because K2/K3 mirror numpy's arithmetic);
exact-distance ties are resolved by the documented (distance, index) rule, which the deterministic oracle shares, so
the GPU is compared with ``tie_rule='deterministic'`` on every query and with the real reference's outputs on the
tie-free ones.
"""
import numpy as np

from oracle import knn_oracle as ko

REG_RTOL = 1e-12      # measured ~1e-16; north_star allows 1e-5
NORTH_STAR_RTOL = 1e-5


def oracle_predict(train, labels, queries, k, gamma, task, return_neighbours=True):
    fn = ko.predict_regression if task == 'regression' else ko.predict_classification
    with np.errstate(all='ignore'):
        return fn(train, labels, queries, k=k, rbf_gamma=gamma, tie_rule='deterministic',
                  return_neighbours=return_neighbours)


def assert_matches_oracle(gpu_pred, gpu_idx, gpu_dist, train, labels, queries, k, gamma, task):
    """GPU output vs the deterministic-rule oracle on the same inputs: indices and distances bit-exact for every
    query; class labels exact; regression within REG_RTOL."""
    o_pred, o_idx, o_dist = oracle_predict(train, labels, queries, k, gamma, task)
    assert gpu_idx.shape == o_idx.shape
    bad = np.nonzero((gpu_idx != o_idx).any(axis=1))[0]
    assert bad.size == 0, 'neighbour indices differ on queries %s: gpu %s oracle %s' % (
        bad[:5], gpu_idx[bad[:1]], o_idx[bad[:1]])
    assert gpu_dist.dtype == o_dist.dtype
    assert np.array_equal(gpu_dist, o_dist, equal_nan=True), 'distances are not bit-identical'
    assert_predictions_match(gpu_pred, o_pred, task)
    return o_pred, o_idx, o_dist


def assert_predictions_match(gpu_pred, o_pred, task):
    gpu_pred = np.asarray(gpu_pred, dtype=np.float64)
    o_pred = np.asarray(o_pred, dtype=np.float64)
    if task == 'regression':
        both_nan = np.isnan(gpu_pred) & np.isnan(o_pred)
        ok = both_nan | (np.abs(gpu_pred - o_pred) <= REG_RTOL * np.maximum(np.abs(o_pred), 1e-300))
        assert ok.all(), 'regression outputs differ: max rel %g' % np.nanmax(
            np.abs(gpu_pred - o_pred) / np.maximum(np.abs(o_pred), 1e-300))
    else:
        assert np.array_equal(gpu_pred, o_pred), 'class predictions differ at %s' % np.nonzero(gpu_pred != o_pred)[0][:8]


def assert_matches_reference_golden(gpu_pred, gpu_idx, gpu_dist, g):
    """GPU output vs the real reference's recorded output: strict on tie-free queries; on tie-flagged queries the
    neighbour set must be a valid completion and the prediction must equal the oracle's aggregation of the GPU's own
    neighbour order."""
    clean = ~g['any_tie']
    ref_pred = g['ref_predictions'].astype(np.float64)
    assert np.array_equal(gpu_idx[clean], g['ref_nn_idx_desc'][clean][:, ::-1])
    assert np.array_equal(gpu_dist[clean], g['ref_nn_dist_desc'][clean][:, ::-1])
    assert_predictions_match(np.asarray(gpu_pred, np.float64)[clean], ref_pred[clean], g['task'])
    tied = np.nonzero(~clean)[0]
    with np.errstate(all='ignore'):
        for qi in tied[:200]:
            d = ko.distances(g['train'], np.ascontiguousarray(g['test'][qi]))
            assert ko.is_valid_neighbour_set(d, g['k'], gpu_idx[qi]) or np.isnan(d[gpu_idx[qi]]).any()
        again = ko.predict_from_neighbours(gpu_idx[tied], gpu_dist[tied], g['train_labels'], g['task'],
                                           rbf_gamma=g['rbf_gamma'])
    assert_predictions_match(np.asarray(gpu_pred, np.float64)[tied], again, g['task'])
    # multiset of distances agrees with the reference even where the tie resolution differs
    assert np.array_equal(np.sort(gpu_dist, axis=1), np.sort(g['ref_nn_dist_desc'], axis=1), equal_nan=True)
    return int(clean.sum()), int(tied.size)
