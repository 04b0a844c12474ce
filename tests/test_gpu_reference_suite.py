"""The reference's own acceptance tests for this path (tests/integration/test_predict.py of ViktorC/OSML), run against
the GPU models on the data splits stored in tests/golden/ (the splits the real reference saw after ``np.random.seed(0)``).

Same structure, same thresholds: fit, predict, evaluate, compare with a bound or with scikit-learn.  The reference
builds its data sets from CSV files that do not travel to the GPU box; the stored splits stand in for them."""
import numpy as np
import pandas as pd
import pytest
import sklearn.neighbors as sknn
import sklearn.svm as sks

import osml_b200 as osp
from tests.conftest import load_golden
from tests.test_oracle_svm import GOLDEN

pytestmark = pytest.mark.gpu


class StoredDataSet:
    """get_training_observations / get_training_labels / get_test_observations / get_test_labels of osml/data.py:13-69
    over a stored split."""

    def __init__(self, train, train_labels, test, test_labels, label_name='label'):
        self._train = pd.DataFrame(np.asfortranarray(train))
        self._train_labels = pd.Series(train_labels, name=label_name)
        self._test = pd.DataFrame(np.asfortranarray(test))
        self._test_labels = pd.Series(test_labels, name=label_name)

    def get_training_observations(self):
        return self._train

    def get_training_labels(self):
        return self._train_labels

    def get_test_observations(self):
        return self._test

    def get_test_labels(self):
        return self._test_labels


def knn_data_set(name):
    g = load_golden(name)
    return StoredDataSet(g['train'], g['train_labels'], g['test'], g['test_labels']), g


def svm_data_set(name):
    import os
    g = np.load(os.path.join(GOLDEN, name + '.npz'))
    return StoredDataSet(g['train'], g['train_labels'], g['test'], g['test_labels'], str(g['label_name'])), g


def test_regression_model_error():                                   # tests/integration/test_predict.py:37, 43-46
    data_set, g = knn_data_set('boston_reg_k7')
    model = osp.KNearestNeighborsRegression()
    model.fit(data_set.get_training_observations(), data_set.get_training_labels())
    error = model.evaluate(model.predict(data_set.get_test_observations()), data_set.get_test_labels())
    assert error <= 7.9
    assert error == pytest.approx(float(g['ref_metric']), rel=1e-12)


@pytest.mark.parametrize('name,min_accuracy', [('iris_cls_k7', .95), ('exam_cls_k7', .7)])
def test_classification_accuracy(name, min_accuracy):               # tests/integration/test_predict.py:57-58, 68-71
    data_set, g = knn_data_set(name)
    model = osp.KNearestNeighborsClassification()
    model.fit(data_set.get_training_observations(), data_set.get_training_labels())
    accuracy = model.evaluate(model.predict(data_set.get_test_observations()), data_set.get_test_labels())
    assert accuracy >= min_accuracy


def test_regression_error_compared_to_sklearn():                     # tests/integration/test_predict.py:92-93, 103-111
    data_set, _ = knn_data_set('boston_reg_k7')
    model = osp.KNearestNeighborsRegression()
    sk_model = sknn.KNeighborsRegressor(n_neighbors=7, weights='distance')
    model.fit(data_set.get_training_observations(), data_set.get_training_labels())
    sk_model.fit(data_set.get_training_observations(), data_set.get_training_labels())
    predictions = model.predict(data_set.get_test_observations())
    sk_predictions = pd.Series(sk_model.predict(data_set.get_test_observations().values), dtype=predictions.dtype)
    model_error = model.evaluate(predictions, data_set.get_test_labels())
    sk_model_error = model.evaluate(sk_predictions, data_set.get_test_labels())
    assert abs(model_error - sk_model_error) < .1 or model_error / sk_model_error <= 1.1


@pytest.mark.parametrize('name', ['iris_cls_k7', 'exam_cls_k7'])
def test_classification_accuracy_compared_to_sklearn(name):        # tests/integration/test_predict.py:125-128, 139-149
    data_set, g = knn_data_set(name)
    model = osp.KNearestNeighborsClassification()
    sk_model = sknn.KNeighborsClassifier(n_neighbors=7, weights='distance')
    model.fit(data_set.get_training_observations(), data_set.get_training_labels())
    sk_model.fit(data_set.get_training_observations(), data_set.get_training_labels())
    predictions = model.predict(data_set.get_test_observations())
    sk_predictions = pd.Series(sk_model.predict(data_set.get_test_observations().values), dtype=predictions.dtype)
    model_accuracy = model.evaluate(predictions, data_set.get_test_labels())
    sk_model_accuracy = model.evaluate(sk_predictions, data_set.get_test_labels())
    # The reference's criterion.  On the stored exam split the REAL reference itself misses it (0.917 against
    # scikit-learn's 0.833: neither within .05 nor 1.2x better), so the outcome is held to the reference's, not to the bound.
    passes = abs(model_accuracy - sk_model_accuracy) < .05 or model_accuracy / sk_model_accuracy >= 1.2
    ref_accuracy = float(g['ref_metric'])
    ref_passes = abs(ref_accuracy - sk_model_accuracy) < .05 or ref_accuracy / sk_model_accuracy >= 1.2
    assert model_accuracy == ref_accuracy and passes == ref_passes


@pytest.mark.parametrize('name,kwargs,min_f1_score', [
    ('svm_titanic_lin', dict(tol=1e-2, c=.1, kernel='lin'), .66),
    ('svm_exam_rbf', dict(), .54),
])
def test_binary_classification_f1_score(name, kwargs, min_f1_score):       # tests/integration/test_predict.py:78-79, 82-85
    data_set, g = svm_data_set(name)
    model = osp.SupportVectorMachine(**kwargs)
    np.random.seed(int(g['fit_seed']))
    model.fit(data_set.get_training_observations(), data_set.get_training_labels())
    f1_score = model.evaluate(model.predict(data_set.get_test_observations()), data_set.get_test_labels())
    assert f1_score >= min_f1_score


@pytest.mark.parametrize('name,kwargs,sk_model', [
    ('svm_titanic_lin', dict(tol=1e-2, c=.1, kernel='lin'), sks.SVC(tol=1e-2, C=.1, kernel='linear')),
    ('svm_exam_rbf', dict(), sks.SVC(tol=1e-3, gamma=.1, C=1., kernel='rbf')),
])
def test_binary_classification_f1_score_compared_to_sklearn(name, kwargs, sk_model):    # test_predict.py:154-157, 161-170
    data_set, g = svm_data_set(name)
    model = osp.SupportVectorMachine(**kwargs)
    np.random.seed(int(g['fit_seed']))
    model.fit(data_set.get_training_observations(), data_set.get_training_labels())
    sk_model.fit(data_set.get_training_observations(), data_set.get_training_labels())
    predictions = model.predict(data_set.get_test_observations())
    sk_predictions = pd.Series(sk_model.predict(data_set.get_test_observations().values), dtype=predictions.dtype)
    model_f1_score = model.evaluate(predictions, data_set.get_test_labels())
    sk_model_f1_score = model.evaluate(sk_predictions, data_set.get_test_labels())
    assert abs(model_f1_score - sk_model_f1_score) < .05 or model_f1_score / sk_model_f1_score >= 1.2


def test_model_loss():                                                       # tests/integration/test_predict.py:22, 27-30
    """The reference bounds the hinge loss of the exam model by 17.1 under its own random state; with the seeds of the
    stored fixture the REAL reference ends at 18.30, so the GPU fit is held to that."""
    data_set, g = svm_data_set('svm_exam_rbf')
    model = osp.SupportVectorMachine()
    np.random.seed(int(g['fit_seed']))
    model.fit(data_set.get_training_observations(), data_set.get_training_labels())
    test_loss = model.test(data_set.get_test_observations(), data_set.get_test_labels())
    assert test_loss <= 1.05 * g['hinge'].item()
