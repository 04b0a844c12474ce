"""GPU parity of the support-vector decision function (include/osml_svm.h, osml_b200/svm.py) with what the REAL
reference returned (tests/golden/svm_*.npz) and with the oracle.

Tolerances (stated in include/osml_svm.h): the GPU adds the per-vector terms in its own fixed order and uses CUDA's exp /
pow, so float64 data agree to 1e-12 and float32 data (float32 kernel values by the reference's own rules) to 1e-5 of
the scale |intercept| + sum_i |K_i * alpha_i|; NaN results must be NaN in the same places."""
import os

import numpy as np
import pandas as pd
import pytest

import osml_b200 as ob
from oracle import svm_oracle as so
from osml_b200 import _native
from tests.test_oracle_svm import DECISION_CASES, FITTED_CASES, GOLDEN, load_decision_case

pytestmark = pytest.mark.gpu


def _assert_close(got, want, vectors, queries, alphas, intercept, args):
    tol = 1e-5 if vectors.dtype == np.float32 else 1e-12
    assert np.array_equal(np.isnan(got), np.isnan(want))
    ok = ~np.isnan(want)
    with np.errstate(invalid='ignore'):
        scale = np.array([abs(intercept) + np.nansum(np.abs(so.kernel_values(vectors, x, **args) * alphas))
                          for x in queries])
    assert np.all(np.abs(got[ok] - want[ok]) <= tol * scale[ok]), float(np.max(np.abs(got[ok] - want[ok]) / scale[ok]))


@pytest.mark.parametrize('name', DECISION_CASES)
def test_decision_values_match_the_reference(name):
    g, vectors, args = load_decision_case(name)
    machine = _native.Machine(vectors, g['labels'], g['alphas'], g['intercept'].item(), **args)
    for queries in (g['queries'], np.asfortranarray(g['queries'])):          # both host layouts of the query matrix
        got = machine.decision(queries)
        _assert_close(got, g['raw'], vectors, g['queries'], g['alphas'], g['intercept'].item(), args)
    assert machine.stats()['kernel_launches'] >= 2 and machine.stats()['last_kernel_ms'] > 0
    # deterministic: a second call returns the same bits
    assert np.array_equal(machine.decision(g['queries']), got, equal_nan=True)
    machine.close()


def test_query_count_and_vector_splits_do_not_change_the_contract():
    """One query (many vector splits), a ragged tile, and more queries than two waves of CTAs (a single split)."""
    rng = np.random.default_rng(3)
    n, d = 5000, 16
    vectors = rng.standard_normal((n, d))
    labels = rng.integers(0, 2, n) * 2. - 1.
    alphas = rng.uniform(0, 1, n)
    args = dict(kernel='rbf', gamma=.3, degree=2, poly_coefficient=1.)
    machine = _native.Machine(vectors, labels, alphas, -.75, **args)
    for q in (1, 13, 3000):
        queries = rng.standard_normal((q, d))
        want = so.raw_predict(queries[:40], vectors, labels, alphas, -.75, **args)
        _assert_close(machine.decision(queries)[:40], want, vectors, queries[:40], alphas, -.75, args)
    machine.close()


@pytest.mark.parametrize('dtype', [np.float32, np.float64])
@pytest.mark.parametrize('d', [1, 5, 13, 24, 130, 300, 1027])
def test_pairwise_feature_sum_of_c_ordered_vectors(dtype, d):
    """observations[alphas > 0] is C-ordered, so at predict time numpy adds the squared differences with its pairwise
    tree (8 strided accumulators per block of <= 128, blocks split at multiples of 8).  Feature counts that leave a
    remainder, that fit one block, and that need a tree of blocks -- against the oracle (= numpy itself).  (The leaf and
    tree arithmetic is the one K2's generic kernel is pinned with bit for bit, tests/test_gpu_parity.py.)"""
    rng = np.random.default_rng(d)
    n, q = 300, 37
    vectors = rng.standard_normal((n, d)).astype(dtype)              # C order
    queries = rng.standard_normal((q, d)).astype(dtype)
    labels = rng.integers(0, 2, n) * 2. - 1.
    alphas = rng.uniform(0, 1, n)
    args = dict(kernel='rbf', gamma=.07, degree=2, poly_coefficient=1.)
    machine = _native.Machine(vectors, labels, alphas, .5, **args)
    want = so.raw_predict(queries, vectors, labels, alphas, .5, **args)
    _assert_close(machine.decision(queries), want, vectors, queries, alphas, .5, args)
    machine.close()
    # one vector, multiplier 1, label 1, intercept 0: the decision value IS exp(arg), arg = -gamma * distance in the
    # data's dtype.  numpy gives the reference's arg; the exact exp of it (float64) must be within CUDA's documented
    # error of exp (2 ulp float32, 1 ulp float64; + the rounding to dtype) -- a distance off by one ulp would move
    # the result by |arg| ulps and more.
    one = _native.Machine(vectors[:1], np.ones(1), np.ones(1), 0., **args)
    got = one.decision(queries)
    arg = np.array([(-args['gamma'] * np.sqrt(np.square(vectors[:1] - x).sum(axis=1)))[0] for x in queries])
    assert arg.dtype == dtype
    exact = np.exp(arg.astype(np.float64))
    ulp = np.spacing(exact.astype(dtype)).astype(np.float64)
    assert np.all(np.abs(got - exact) <= 2.5 * ulp), float(np.max(np.abs(got - exact) / ulp))
    one.close()


@pytest.mark.parametrize('name', ['svm_synth_f32_rbf_forder', 'svm_synth_f64_rbf_corder', 'svm_synth_i8_rbf_wrap',
                                  'svm_synth_f64_poly', 'svm_synth_f32_poly', 'svm_synth_f64_lin', 'svm_synth_f32_lin'])
def test_kernel_table_matches_the_pairwise_evaluations_of_the_reference(name):
    """osml_svm_gram against K(x_i, x_j) evaluated pair by pair on (1, d)-shaped rows, as _optimize_ij does
    (osml/predict.py:1140-1142): same dtype; rbf to 2.5 ulp of the exact exp of numpy's argument, NaN in the same places
    (wrapped int8 sums); lin / poly to the accuracy of a d-term dot product in the data's dtype."""
    g, vectors, args = load_decision_case(name)
    vectors = vectors[:60]
    n, d = vectors.shape
    machine = _native.Machine(vectors, np.ones(n), np.ones(n), 0., **args)
    got = machine.gram()
    machine.close()
    rows = [vectors[i].reshape((1, d)) for i in range(n)]
    with np.errstate(invalid='ignore'):
        want = np.array([[so.kernel_values(rows[i], rows[j][0], **args).reshape(-1)[0] for j in range(n)]
                         for i in range(n)])
    assert got.dtype == want.dtype and got.shape == (n, n)
    assert np.array_equal(np.isnan(got), np.isnan(want))
    ok = ~np.isnan(want)
    if args['kernel'] == 'rbf':
        with np.errstate(invalid='ignore'):
            arg = np.array([[(-args['gamma'] * np.sqrt(np.square(rows[i] - rows[j][0]).sum(axis=1)))[0]
                             for j in range(n)] for i in range(n)])
        exact = np.exp(arg.astype(np.float64))
        ulp = np.spacing(np.abs(exact).astype(want.dtype)).astype(np.float64)
        assert np.all(np.abs(got.astype(np.float64) - exact)[ok] <= 2.5 * ulp[ok])
    else:
        eps = np.finfo(want.dtype).eps
        mag = np.abs(vectors.astype(np.float64)) @ np.abs(vectors.astype(np.float64)).T + abs(args['poly_coefficient'])
        bound = 4 * d * eps * (mag if args['kernel'] == 'lin' else args['degree'] * mag ** args['degree'])
        assert np.all(np.abs(got.astype(np.float64) - want.astype(np.float64)) <= bound)


@pytest.mark.parametrize('name', FITTED_CASES)
def test_fitted_state_of_the_reference_through_the_model_api(name):
    """The reference's fitted state loaded into the GPU model: decision values, fuzzy memberships, labels, hinge loss and
    F1 on the test split, and the SMO error cache of the final state."""
    g = np.load(os.path.join(GOLDEN, name + '.npz'))
    args = dict(kernel=str(g['kernel']), gamma=g['gamma'].item(), degree=g['degree'].item(),
                poly_coefficient=g['poly_coefficient'].item())
    train_df = pd.DataFrame(g['train'])
    test_df = pd.DataFrame(g['test'])
    test_sr = pd.Series(g['test_labels'], name=str(g['label_name']))
    model = ob.SupportVectorMachine(tol=g['tol'].item(), c=g['c'].item(), **args)
    model._observation_types, model._label_type, model._label_name = train_df.dtypes, test_sr.dtype, test_sr.name
    model._support_vectors = np.ascontiguousarray(g['support_vectors'])          # observations[alphas > 0]: C order
    model._support_vector_labels, model._support_vector_alphas = g['support_labels'], g['support_alphas']
    model._intercept = g['intercept'].item()
    raw = model.decision_function(test_df)
    _assert_close(raw, g['raw'], model._support_vectors, g['test'], g['support_alphas'], model._intercept, args)
    np.testing.assert_allclose(model.fuzzy_predict(test_df).values, g['fuzzy'], rtol=0, atol=1e-12)
    pred = model.predict(test_df)
    assert pred.dtype == g['pred'].dtype and np.array_equal(pred.values, g['pred'])
    assert model.test(test_df, test_sr) == pytest.approx(g['hinge'].item(), rel=1e-12)
    assert model.evaluate(pred, test_sr) == g['f1'].item()
    # the reference's private signature with explicit matrices (a temporary device copy)
    raw2 = model._raw_predict(g['test'], g['support_vectors'], g['support_labels'], g['support_alphas'],
                              g['intercept'].item())
    assert np.array_equal(raw2, raw)
    # error cache: every training row (F order, as DataFrame.values) against all of them, zero multipliers included
    train = np.asfortranarray(g['train'])
    labels = g['train_labels'] * 2. - 1.
    alphas = np.zeros(len(train))
    alphas[g['support_rows']] = g['support_alphas']
    machine = _native.Machine(train, labels, np.zeros(len(train)), 0., **args)
    machine.set_multipliers(alphas, g['fit_intercept'].item())
    errors = machine.decision_self() - labels
    scale = np.abs(g['errors']).max() + np.abs(alphas).sum() * (np.abs(g['train']).max() ** 2 * train.shape[1]
                                                                if args['kernel'] == 'lin' else 1.)
    assert np.max(np.abs(errors - g['errors'])) <= 1e-12 * scale
    machine.close()


def _check_dual_feasible(model, n_train):
    alphas, labels = model._support_vector_alphas, model._support_vector_labels
    assert np.all(alphas > 0) and np.all(alphas <= model.c * (1 + 1e-12))
    assert abs(np.sum(alphas * labels)) <= 1e-9 * max(1., np.sum(alphas))          # sum_i alpha_i y_i = 0 is kept by SMO
    assert 0 < len(alphas) <= n_train


def test_smo_fit_on_exam():
    """Same data and ``np.random`` seed as oracle/make_golden_svm.py.  The host SMO logic is pinned step for step on the
    CPU (tests/test_host_svm.py, oracle-backed error cache); on the GPU the error cache differs from numpy's in the last
    bits (summation order, exp), and over ~1000 steps SMO may fork at a comparison decided by those bits -- so the
    fitted state is held to the reference's when the path was the same, and to SMO's invariants and the reference's
    acceptance bar (tests/integration/test_predict.py:79) either way."""
    g = np.load(os.path.join(GOLDEN, 'svm_exam_rbf.npz'))
    train_df = pd.DataFrame(np.asfortranarray(g['train']))
    train_sr = pd.Series(g['train_labels'], name=str(g['label_name']))
    test_df = pd.DataFrame(g['test'])
    test_sr = pd.Series(g['test_labels'], name=str(g['label_name']))
    model = ob.SupportVectorMachine()
    np.random.seed(int(g['fit_seed']))
    model.fit(train_df, train_sr)
    assert model.smo_steps > 0
    assert model._support_vectors.flags['C_CONTIGUOUS']
    _check_dual_feasible(model, len(train_df))
    pred = model.predict(test_df)
    hinge, f1 = model.test(test_df, test_sr), model.evaluate(pred, test_sr)
    same_path = (model._support_vectors.shape == g['support_vectors'].shape and
                 np.array_equal(model._support_vectors, g['support_vectors']) and
                 np.allclose(model._support_vector_alphas, g['support_alphas'], rtol=1e-6, atol=1e-9))
    print('exam: %d SMO steps, %d support vectors, same path as the reference: %s, hinge %.4f (reference %.4f), '
          'F1 %.4f (reference %.4f)' % (model.smo_steps, len(model._support_vectors), same_path, hinge,
                                        g['hinge'].item(), f1, g['f1'].item()))
    if same_path:
        assert np.array_equal(pred.values, g['pred'])
        assert hinge == pytest.approx(g['hinge'].item(), rel=1e-6)
    assert f1 >= .54 and hinge <= 1.25 * g['hinge'].item()
    import copy
    import pickle
    for clone in (copy.deepcopy(model), pickle.loads(pickle.dumps(model))):
        assert np.array_equal(clone.predict(test_df).values, pred.values)


def test_smo_fit_on_titanic_linear_kernel():
    """621 rows, linear kernel, c = 0.1, tol = 1e-2 (tests/integration/test_predict.py:78): thousands of SMO steps, each
    followed by a full error-cache refresh on the GPU.  Must reach the reference's quality bar; the fitted state is
    compared with the reference's when the run did not fork on a last-bit comparison."""
    g = np.load(os.path.join(GOLDEN, 'svm_titanic_lin.npz'))
    train_df = pd.DataFrame(np.asfortranarray(g['train']))
    train_sr = pd.Series(g['train_labels'], name=str(g['label_name']))
    test_df = pd.DataFrame(g['test'])
    test_sr = pd.Series(g['test_labels'], name=str(g['label_name']))
    model = ob.SupportVectorMachine(tol=1e-2, c=.1, kernel='lin')
    np.random.seed(int(g['fit_seed']))
    model.fit(train_df, train_sr)
    _check_dual_feasible(model, len(train_df))
    pred = model.predict(test_df)
    assert model.evaluate(pred, test_sr) >= .66
    same_path = (model._support_vectors.shape == g['support_vectors'].shape and
                 np.array_equal(model._support_vectors, g['support_vectors']))
    print('titanic: %d SMO steps, %d support vectors, same path as the reference: %s, F1 %.4f (reference %.4f)' %
          (model.smo_steps, len(model._support_vectors), same_path, model.evaluate(pred, test_sr), g['f1'].item()))
    if same_path:
        np.testing.assert_allclose(model._support_vector_alphas, g['support_alphas'], rtol=1e-6, atol=1e-9)
        assert np.array_equal(pred.values, g['pred'])
