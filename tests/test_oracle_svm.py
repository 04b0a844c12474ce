"""The SVM oracle (oracle/svm_oracle.py) against what the REAL reference returned (tests/golden/svm_*.npz, written by
oracle/make_golden_svm.py in the authoring container): bit for bit, NaN positions included."""
import glob
import os

import numpy as np
import pytest

from oracle import svm_oracle as so

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
DECISION_CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, 'svm_synth_*.npz')))
FITTED_CASES = ['svm_exam_rbf', 'svm_titanic_lin']


def _same(a, b):
    return np.array_equal(np.asarray(a), np.asarray(b), equal_nan=True)


def load_decision_case(name):
    g = np.load(os.path.join(GOLDEN, name + '.npz'))
    vectors = np.asarray(g['vectors'], order=str(g['order']))
    args = dict(kernel=str(g['kernel']), gamma=g['gamma'].item(), degree=g['degree'].item(),
                poly_coefficient=g['poly_coefficient'].item())
    return g, vectors, args


def test_the_fixture_set_is_complete():
    assert len(DECISION_CASES) == 10
    kinds = {(str(np.load(os.path.join(GOLDEN, n + '.npz'))['kernel'])) for n in DECISION_CASES}
    assert kinds == {'lin', 'poly', 'rbf'}


@pytest.mark.parametrize('name', DECISION_CASES)
def test_decision_values_match_the_reference_bit_for_bit(name):
    g, vectors, args = load_decision_case(name)
    raw = so.raw_predict(g['queries'], vectors, g['labels'], g['alphas'], g['intercept'].item(), **args)
    assert _same(raw, g['raw'])


def test_the_summation_order_of_the_vectors_matrix_matters():
    """Same numbers, other memory order: the reference's own result changes in the last bits (d >= 8), which is why
    the C ABI takes a sum_order."""
    g, vectors, args = load_decision_case('svm_synth_f32_rbf_corder')
    other = so.raw_predict(g['queries'], np.asfortranarray(vectors), g['labels'], g['alphas'], g['intercept'].item(),
                           **args)
    assert not _same(other, g['raw'])
    np.testing.assert_allclose(other, g['raw'], rtol=1e-5)


@pytest.mark.parametrize('name', FITTED_CASES)
def test_fitted_state_reproduces_the_reference_outputs(name):
    g = np.load(os.path.join(GOLDEN, name + '.npz'))
    args = dict(kernel=str(g['kernel']), gamma=g['gamma'].item(), degree=g['degree'].item(),
                poly_coefficient=g['poly_coefficient'].item())
    raw = so.raw_predict(g['test'], g['support_vectors'], g['support_labels'], g['support_alphas'],
                         g['intercept'].item(), **args)
    assert _same(raw, g['raw'])
    fuzzy = so.fuzzy_from_raw(raw)
    assert _same(fuzzy, g['fuzzy'])
    assert _same(so.labels_from_fuzzy(fuzzy).astype(g['pred'].dtype), g['pred'])
    assert so.hinge_loss(raw, g['test_labels']) == g['hinge'].item()
    # the SMO error cache of the final state: all training rows (F-ordered, as DataFrame.values is), zero multipliers in
    train = np.asfortranarray(g['train'])
    alphas = np.zeros(len(train))
    alphas[g['support_rows']] = g['support_alphas']
    errors = so.error_cache(train, g['train_labels'] * 2. - 1., alphas, g['fit_intercept'].item(), **args)
    assert _same(errors, g['errors'])
