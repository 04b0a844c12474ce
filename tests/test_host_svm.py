"""Host-side contract of the SVM mirror (osml_b200/svm.py) that needs no GPU: constructor validation, label validation,
the binary metrics, pickling, and that nothing computes on the CPU when there is no device."""
import copy
import os
import pickle

import numpy as np
import pandas as pd
import pytest
import torch

import osml_b200 as ob
import osml_b200.svm as ob_svm
from oracle import svm_oracle as so

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def test_mro_matches_the_reference():
    # this package's own classes come first and in the reference's order; with the reference package reachable its
    # classes of the same names follow as bases (osml_b200.model.reference_modules), so isinstance holds both ways
    names = [c.__name__ for c in ob.SupportVectorMachine.__mro__ if c.__module__.startswith('osml_b200')]
    assert names == ['SupportVectorMachine', 'BinaryClassificationModel', 'PredictiveModel', 'Model']
    assert all(c is object or c.__module__.startswith(('osml_b200', 'osml.')) for c in ob.SupportVectorMachine.__mro__)


@pytest.mark.parametrize('kwargs', [
    dict(tol=-1e-3), dict(c=-1.), dict(kernel='sigmoid'), dict(kernel='rbf', gamma=0.), dict(kernel='rbf', gamma=None),
    dict(kernel='poly', degree=0), dict(kernel='poly', degree=None), dict(kernel='poly', poly_coefficient=None),
])
def test_constructor_rejects_what_the_reference_rejects(kwargs):
    with pytest.raises(ValueError):                       # osml/predict.py:1057-1070
        ob.SupportVectorMachine(**kwargs)


def test_constructor_defaults_and_attributes():
    m = ob.SupportVectorMachine()
    assert (m.tol, m.c, m.kernel, m.gamma, m.degree, m.poly_coefficient, m.epsilon) == (1e-3, 1., 'rbf', .5, 2, 1., 1e-8)
    assert m._support_vectors is None and m._intercept is None
    ob.SupportVectorMachine(kernel='lin', gamma=None)      # parameters of the kernels not in use are not checked


def test_labels_must_be_zero_one_integers():
    X = pd.DataFrame(np.arange(12.).reshape(6, 2))
    m = ob.SupportVectorMachine()
    with pytest.raises(ValueError):
        m.fit(X, pd.Series([0, 1, 2, 0, 1, 2], name='y'))          # osml/predict.py:119-122
    with pytest.raises(ValueError):
        m.fit(X, pd.Series([0., 1., 0., 1., 0., 1.], name='y'))
    with pytest.raises(ValueError):
        m.fit(X.iloc[:1], pd.Series([1], name='y'))                # fewer than two rows (osml/predict.py:1208-1209)


def test_fit_fails_loudly_without_a_gpu():
    if torch.cuda.is_available():
        pytest.skip('needs a box without a GPU')
    X = pd.DataFrame(np.arange(12.).reshape(6, 2))
    with pytest.raises(RuntimeError):
        ob.SupportVectorMachine().fit(X, pd.Series([0, 1, 0, 1, 0, 1], name='y'))


def test_binary_metrics_match_the_reference_definitions():
    m = ob.SupportVectorMachine()
    m._label_type = np.dtype(np.int64)
    pred = pd.Series([1, 1, 0, 1, 0, 0, 1, 0])
    true = pd.Series([1, 0, 0, 1, 1, 0, 1, 1])
    assert m.evaluate_precision(pred, true) == 3 / 4               # osml/predict.py:152-175
    assert m.evaluate_recall(pred, true) == 3 / 5                  # osml/predict.py:177-201
    assert m.evaluate_fall_out(pred, true) == 1 / 3                # osml/predict.py:203-225
    p, r, eps = 3 / 4, 3 / 5, 1e-8
    assert m.evaluate(pred, true) == (2 * p * r + eps) / (p + r + eps)          # osml/predict.py:129-133
    zeros = pd.Series([0] * 8)
    assert m.evaluate_precision(zeros, true) == 1. and m.evaluate_recall(pred, zeros) == 1.
    with pytest.raises(ValueError):
        m.evaluate_precision(pred.iloc[:3], true)
    with pytest.raises(ValueError):
        m.evaluate_recall(pd.Series([2, 1, 0, 1, 0, 0, 1, 0]), true)
    fuzzy = pd.Series([.9, .6, .4, .8, .3, .2, .7, .45])
    roc, auc = m.evaluate_receiver_operating_characteristic(fuzzy, true, threshold_increment=.25)
    assert [t for t, _, _ in roc] == [0., .25, .5, .75, 1.]
    assert roc[0][1:] == (1., 1.) and roc[-1][1:] == (0., 0.)
    # thresholds .25 / .5 / .75: recall 1, .6, .4; fall-out 2/3, 1/3, 0 -- trapezoids between points that are both
    # non-zero (the reference skips a segment whose previous point has a zero coordinate)
    assert auc == pytest.approx((1 + 1) / 2 * (1 - 2 / 3) + (.6 + 1) / 2 * (2 / 3 - 1 / 3) + (.4 + .6) / 2 * (1 / 3 - 0))
    with pytest.raises(ValueError):
        m.evaluate_receiver_operating_characteristic(fuzzy, true, threshold_increment=1.)


def test_state_survives_pickle_and_deepcopy_without_device_handles():
    m = ob.SupportVectorMachine(kernel='poly', degree=3)
    m._support_vectors = np.arange(6.).reshape(3, 2)
    m._support_vector_labels = np.array([1., -1., 1.])
    m._support_vector_alphas = np.array([.5, 1., .5])
    m._intercept = .25
    m._machine = object()                                           # stands in for a live device handle
    for clone in (copy.deepcopy(m), pickle.loads(pickle.dumps(m))):
        assert clone._machine is None
        assert (clone.kernel, clone.degree, clone._intercept) == ('poly', 3, .25)
        assert np.array_equal(clone._support_vectors, m._support_vectors)
    m._machine = None


class _OracleMachine:
    """Stands in for ``_native.Machine`` with the numpy oracle behind it, so that the host SMO logic can be checked on the
    CPU against the reference's fitted state.  Test infrastructure only."""

    def __init__(self, vectors, labels, alphas, intercept, kernel, gamma=.5, degree=2, poly_coefficient=1., device=0,
                 sum_order=None):
        self.vectors, self.labels = vectors, labels
        self.alphas, self.intercept = np.array(alphas), intercept
        self.args = dict(kernel=kernel, gamma=gamma, degree=degree, poly_coefficient=poly_coefficient)

    def set_multipliers(self, alphas, intercept):
        self.alphas, self.intercept = np.array(alphas), intercept

    def decision(self, queries):
        return so.raw_predict(queries, self.vectors, self.labels, self.alphas, self.intercept, **self.args)

    def decision_self(self):
        return self.decision(self.vectors)

    def gram(self):
        # pair by pair on (1, d)-shaped rows, exactly as the reference evaluates them (osml/predict.py:1140-1142)
        n, d = self.vectors.shape
        rows = [self.vectors[i].reshape((1, d)) for i in range(n)]
        first = so.kernel_values(rows[0], rows[0][0], **self.args).reshape(-1)[0]
        out = np.empty((n, n), dtype=np.asarray(first).dtype)
        for i in range(n):
            for j in range(n):
                out[i, j] = so.kernel_values(rows[i], rows[j][0], **self.args).reshape(-1)[0]
        return out

    def close(self):
        pass


def test_smo_host_logic_walks_the_reference_path_step_for_step(monkeypatch):
    """With an error cache that is bit-identical to the reference's (the oracle), the SMO restatement must end exactly
    where the REAL reference ended for the same data and ``np.random`` seed (tests/golden/svm_exam_rbf.npz): same
    support vectors, multipliers, intercept, predictions -- including the reference's quirk of using a position in the
    unbounded list as a row number (osml/predict.py:1188-1192)."""
    monkeypatch.setattr(ob_svm._native, 'Machine', _OracleMachine)
    g = np.load(os.path.join(GOLDEN, 'svm_exam_rbf.npz'))
    train_df = pd.DataFrame(np.asfortranarray(g['train']))
    train_sr = pd.Series(g['train_labels'], name=str(g['label_name']))
    test_df = pd.DataFrame(g['test'])
    test_sr = pd.Series(g['test_labels'], name=str(g['label_name']))
    model = ob.SupportVectorMachine()
    np.random.seed(int(g['fit_seed']))
    model.fit(train_df, train_sr)
    assert model.smo_steps == 968
    assert model._support_vectors.flags['C_CONTIGUOUS']
    assert np.array_equal(model._support_vectors, g['support_vectors'])
    assert np.array_equal(model._support_vector_labels, g['support_labels'])
    assert np.array_equal(model._support_vector_alphas, g['support_alphas'])
    assert model._intercept == g['intercept'].item()
    assert np.array_equal(model.decision_function(test_df), g['raw'])
    assert np.array_equal(model.fuzzy_predict(test_df).values, g['fuzzy'])
    pred = model.predict(test_df)
    assert pred.dtype == g['pred'].dtype and np.array_equal(pred.values, g['pred'])
    assert model.test(test_df, test_sr) == g['hinge'].item()
    assert model.evaluate(pred, test_sr) == g['f1'].item()
