"""CPU tests: the oracle (oracle/knn_oracle.py) against the golden vectors produced by the real reference
(oracle/make_golden.py), and its explicit summation-order restatements against numpy itself."""
import hashlib
import math

import numpy as np
import pandas as pd
import pytest

from oracle import knn_oracle as ko

SURVEY_PINS = {  # SURVEY.md section 8c "known-answer values"
    'iris_cls_k7': ('b9f47435de9ab836', 0.9777777777777777),
    'boston_reg_k5': ('814c7c8a0ecc8755', 6.195528043302151),
    'boston_reg_k7': ('25f5d5273e6e491e', 6.366717991618547),
    'mushroom_cls_k7': ('c44862d8290c6cd5', 1.0),
    'exam_cls_k7': ('987b374780811e11', 0.9166666666666666),
    'titanic_cls_k7': ('5d2c6164478141b5', 0.7218045112781954),
}


def _same_versions(g):
    return str(g['numpy_version']) == np.__version__ and str(g['pandas_version']) == pd.__version__


def _predict(g, tie_rule, return_neighbours=False):
    fn = ko.predict_regression if g['task'] == 'regression' else ko.predict_classification
    return fn(g['train'], g['train_labels'], g['test'], k=g['k'], rbf_gamma=g['rbf_gamma'], tie_rule=tie_rule,
              return_neighbours=return_neighbours)


def test_golden_files_carry_survey_pins(golden):
    if golden['name'] in SURVEY_PINS:
        sha, metric = SURVEY_PINS[golden['name']]
        assert str(golden['pred_sha256_16']) == sha
        assert float(golden['ref_metric']) == metric
        cast = golden['ref_predictions']
        assert hashlib.sha256(cast.tobytes()).hexdigest()[:16] == sha


def test_oracle_reference_rule_reproduces_reference_bit_for_bit(golden):
    """Same arithmetic, same argpartition/pandas tie behaviour -> identical bits (when numpy/pandas match the
    versions that generated the goldens; otherwise only tie-free queries are comparable)."""
    pred = _predict(golden, 'reference')
    ref = golden['ref_predictions'].astype(np.float64)
    mask = np.ones(len(ref), bool) if _same_versions(golden) else ~golden['any_tie']
    assert np.array_equal(pred[mask], ref[mask], equal_nan=True)


def test_oracle_deterministic_rule_matches_reference_off_ties(golden):
    pred, idx_asc, dist_asc = _predict(golden, 'deterministic', return_neighbours=True)
    ref = golden['ref_predictions'].astype(np.float64)
    clean = ~golden['any_tie']
    assert np.array_equal(pred[clean], ref[clean])
    assert np.array_equal(idx_asc[clean], golden['ref_nn_idx_desc'][clean][:, ::-1])
    assert np.array_equal(dist_asc[clean], golden['ref_nn_dist_desc'][clean][:, ::-1])
    # on tie-flagged queries the reference's own set and ours must both be valid completions
    for qi in np.nonzero(golden['any_tie'])[0][:64]:
        d = ko.distances(golden['train'], golden['test'][qi])
        assert ko.is_valid_neighbour_set(d, golden['k'], idx_asc[qi])
        assert ko.is_valid_neighbour_set(d, golden['k'], golden['ref_nn_idx_desc'][qi][::-1])
    # distances agree as multisets even on ties
    assert np.array_equal(np.sort(dist_asc, axis=1), np.sort(golden['ref_nn_dist_desc'], axis=1), equal_nan=True)


def test_aggregation_from_given_neighbours_matches(golden):
    pred, idx_asc, dist_asc = _predict(golden, 'deterministic', return_neighbours=True)
    again = ko.predict_from_neighbours(idx_asc, dist_asc, golden['train_labels'], golden['task'],
                                       rbf_gamma=golden['rbf_gamma'])
    assert np.array_equal(pred, again, equal_nan=True)


def test_explicit_summation_order_matches_numpy(golden):
    train = golden['train']
    order = ko.sum_order_of(train)
    assert order == ('sequential' if bool(golden['train_is_f_order']) else 'pairwise')
    explicit = ko.distances_sequential if order == 'sequential' else ko.distances_pairwise
    with np.errstate(invalid='ignore'):
        for qi in range(min(8, golden['test'].shape[0])):
            q = np.ascontiguousarray(golden['test'][qi])
            assert np.array_equal(ko.distances(train, q), explicit(train, q), equal_nan=True)


@pytest.mark.parametrize('dtype', [np.float32, np.float64])
@pytest.mark.parametrize('d', [1, 2, 4, 6, 7, 8, 9, 13, 15, 16, 17, 96, 112, 127, 128, 129, 136, 255, 257, 784, 1031])
def test_summation_orders_random(dtype, d):
    rng = np.random.default_rng(d)
    t = (rng.standard_normal((257, d)) * 3).astype(dtype)
    q = (rng.standard_normal(d) * 3).astype(dtype)
    tf = np.asfortranarray(t)
    assert np.array_equal(ko.distances(tf, q), ko.distances_sequential(tf, q))
    assert np.array_equal(ko.distances(t, q), ko.distances_pairwise(t, q))
    # a dataframe hands the reference the F-ordered matrix
    assert ko.sum_order_of(pd.DataFrame(t).values) == 'sequential'


def test_int8_wraps_like_the_reference():
    t = np.asfortranarray(np.array([[100, -100], [3, 4], [127, -128]], dtype=np.int8))
    q = np.array([-100, 100], dtype=np.int8)
    with np.errstate(invalid='ignore'):
        d = ko.distances(t, q)
        assert d.dtype == np.float64
        assert np.array_equal(d, ko.distances_sequential(t, q), equal_nan=True)
    # 100 - (-100) wraps to -56; (-56)^2 = 3136 wraps to 64
    assert d[0] == np.sqrt(np.float64(64 + 64))


def test_weights_are_fp64_python_scalars_whatever_the_data_dtype():
    """pandas boxes Series values to Python floats on iteration (predict.py:755), so an fp32 distance is widened
    before the ratio; SURVEY.md row a4's "fp32 ratio" does not hold for pandas 3.0.2."""
    s = pd.Series(np.array([1.5, 2.5], np.float32), index=np.array([5, 3]))
    assert all(type(v) is float and type(i) is int for i, v in s.items())
    w = ko.weighted_labels(np.array([2, 1, 0]), np.array([2.0, 1.7, 0.3], np.float32), np.array([7, 5, 9]))
    r1 = float(np.float32(1.7)) / (2.0 + 1e-8)
    r0 = float(np.float32(0.3)) / (2.0 + 1e-8)
    assert list(w.keys()) == [5, 7]
    assert w[5] == math.exp(-1 * r1 ** 2)
    assert w[7] == math.exp(-1 * r0 ** 2)


def test_select_deterministic_rule_and_errors():
    d = np.array([3.0, 1.0, 2.0, 1.0, np.nan, 2.0, 0.5])
    idx, dd = ko.select_deterministic(d, 3)
    assert idx.tolist() == [2, 3, 1, 6]
    with pytest.raises(ValueError):
        ko.select_deterministic(np.array([1.0, 2.0, 3.0]), 2)  # N == k+1
    idx, _ = ko.select_deterministic(d, 5)
    assert idx.tolist()[0] == 4 or not np.isnan(d[idx[0]])  # NaN last in ascending order = first in descending
    assert ko.tie_flags(d, 3) == (True, True)
    assert ko.tie_flags(np.array([1., 2., 3., 4., 5., 6.]), 3) == (False, False)


def test_single_row_matrices_sum_pairwise_whatever_their_flags_say():
    """A (1, d) matrix is contiguous along the features in C and in F order, so numpy reduces it with pairwise_sum
    (the support-vector path meets this: one support vector, and the (1, d)-shaped rows of SMO's kernel calls)."""
    from osml_b200 import _native
    rng = np.random.default_rng(5)
    q = rng.standard_normal(130).astype(np.float32)
    for order in ('C', 'F'):
        row = np.asarray(rng.standard_normal((1, 130)).astype(np.float32), order=order)
        assert ko.sum_order_of(row) == 'pairwise'
        assert _native.sum_order_of(row) == _native.SUM_PAIRWISE
        assert np.array_equal(ko.distances(row, q), ko.distances_pairwise(row, q))
    column = rng.standard_normal((130, 1)).astype(np.float32)
    assert ko.sum_order_of(column) == 'sequential' and _native.sum_order_of(column) == _native.SUM_SEQUENTIAL
