"""Generates tests/golden/svm_*.npz by running the REAL reference's SupportVectorMachine (oracle/ref_shim.py).

Run:  python oracle/make_golden_svm.py        (needs /root/reference; not runnable on the GPU box)

Fitted cases (the two configurations of the reference's own tests, tests/integration/test_predict.py:22, 78-79):
``np.random.seed(0)`` before the data set is built (its shuffle / split), ``np.random.seed(FIT_SEED)`` right before
``fit`` (SMO draws its start indices from ``np.random``, osml/predict.py:1183, 1194).  Stored: the matrices the
reference saw, the fitted state (support vectors, their -1 / +1 labels, multipliers, intercept), and what the
reference returned on the test split: ``_raw_predict``, ``fuzzy_predict``, ``predict``, ``test`` (hinge loss),
``evaluate`` (F1), and the error cache ``_update_error_cache`` computes for the final multipliers.

Decision-function cases (no fit): random vectors / multipliers pushed through the reference's ``_raw_predict`` for each
kernel, float32 / float64 / int8 data, C- and F-ordered vector matrices (the two summation orders).
"""
import os
import sys

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle.ref_shim import data_root, import_reference  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), 'tests', 'golden')
FIT_SEED = 7


def versions():
    return dict(numpy_version=np.__version__, pandas_version=pd.__version__)


def fitted_case(name, osp, data_set, **svm_args):
    model = osp.SupportVectorMachine(**svm_args)
    train_df, train_sr = data_set.get_training_observations(), data_set.get_training_labels()
    test_df, test_sr = data_set.get_test_observations(), data_set.get_test_labels()
    np.random.seed(FIT_SEED)
    model.fit(train_df, train_sr)
    raw = model._raw_predict(test_df.values, model._support_vectors, model._support_vector_labels,
                             model._support_vector_alphas, model._intercept)
    fuzzy = model.fuzzy_predict(test_df)
    pred = model.predict(test_df)
    # the error cache of the final SMO state: all training rows, multipliers of non-support rows are zero
    train = train_df.values
    labels_pm = train_sr.values * 2. - 1.
    alphas_full = np.zeros(len(train))
    # rows are unique in these data sets after the shuffle, so the support rows can be located by value
    sv_rows = []
    for v in model._support_vectors:
        hits = np.flatnonzero((train == v).all(axis=1))
        sv_rows.append(int([h for h in hits if h not in sv_rows][0]))
    alphas_full[sv_rows] = model._support_vector_alphas
    n_sv = len(model._support_vectors)
    fit_intercept = model._intercept * n_sv / len(train)          # undo predict.py:1116
    errors = np.zeros(len(train))
    model._update_error_cache(train, labels_pm, alphas_full, errors, np.full(1, fit_intercept))
    path = os.path.join(OUT, name + '.npz')
    np.savez_compressed(
        path, train=np.ascontiguousarray(train), train_labels=train_sr.values, test=np.ascontiguousarray(test_df.values),
        test_labels=test_sr.values, label_name=str(train_sr.name), label_dtype=str(train_sr.dtype),
        column_dtypes=np.array([str(t) for t in train_df.dtypes]),
        kernel=svm_args.get('kernel', 'rbf'), tol=model.tol, c=model.c, gamma=model.gamma, degree=model.degree,
        poly_coefficient=model.poly_coefficient, fit_seed=FIT_SEED,
        support_vectors=np.ascontiguousarray(model._support_vectors), support_rows=np.array(sv_rows),
        support_labels=model._support_vector_labels, support_alphas=model._support_vector_alphas,
        intercept=model._intercept, fit_intercept=fit_intercept, errors=errors,
        raw=raw, fuzzy=fuzzy.values, pred=pred.values, hinge=model.test(test_df, test_sr),
        f1=model.evaluate(pred, test_sr), **versions())
    print(name, 'n_train', len(train), 'support vectors', n_sv, 'hinge', float(model.test(test_df, test_sr)),
          'f1', float(model.evaluate(pred, test_sr)))


def decision_case(name, osp, dtype, kernel, order, n, d, q, seed, int_range=5, **svm_args):
    rng = np.random.default_rng(seed)
    if np.issubdtype(dtype, np.integer):
        # |difference| <= 2 * int_range: squares stay below 128 for int_range = 5 and wrap in int8 above that
        vectors = rng.integers(-int_range, int_range + 1, (n, d)).astype(dtype)
        queries = rng.integers(-int_range, int_range + 1, (q, d)).astype(dtype)
    else:
        vectors = rng.standard_normal((n, d)).astype(dtype)
        queries = rng.standard_normal((q, d)).astype(dtype)
    vectors = np.asarray(vectors, order=order)
    labels = rng.integers(0, 2, n) * 2. - 1.
    alphas = rng.uniform(0, 1, n)
    alphas[rng.integers(0, n, n // 10)] = 0.          # some multipliers are exactly zero, as in SMO's error cache
    intercept = float(rng.standard_normal())
    model = osp.SupportVectorMachine(kernel=kernel, **svm_args)
    raw = model._raw_predict(queries, vectors, labels, alphas, intercept)
    np.savez_compressed(os.path.join(OUT, name + '.npz'), vectors=np.ascontiguousarray(vectors), order=order,
                        queries=queries, labels=labels, alphas=alphas, intercept=intercept, kernel=kernel,
                        gamma=model.gamma, degree=model.degree, poly_coefficient=model.poly_coefficient, raw=raw,
                        **versions())
    print(name, raw[:3])


def main():
    osp, osd, _ = import_reference()
    root = data_root()
    os.makedirs(OUT, exist_ok=True)
    os.chdir(root)
    np.random.seed(0)
    fitted_case('svm_exam_rbf', osp, osd.ExamDataSet('data/exam/ex4x.dat', 'data/exam/ex4y.dat'))
    np.random.seed(0)
    fitted_case('svm_titanic_lin', osp, osd.TitanicDataSet('data/titanic/titanic.csv'), tol=1e-2, c=.1, kernel='lin')
    decision_case('svm_synth_f32_rbf_corder', osp, np.float32, 'rbf', 'C', 700, 24, 200, 1, gamma=.05)
    decision_case('svm_synth_f32_rbf_forder', osp, np.float32, 'rbf', 'F', 700, 24, 200, 2, gamma=.05)
    decision_case('svm_synth_f64_rbf_corder', osp, np.float64, 'rbf', 'C', 400, 136, 100, 3, gamma=.02)
    decision_case('svm_synth_f64_rbf_forder', osp, np.float64, 'rbf', 'F', 333, 7, 129, 4, gamma=.5)
    decision_case('svm_synth_i8_rbf', osp, np.int8, 'rbf', 'F', 400, 30, 100, 5, gamma=.01)
    # int8 squares that wrap (numpy keeps int8 through subtract and square): negative sums give NaN distances
    decision_case('svm_synth_i8_rbf_wrap', osp, np.int8, 'rbf', 'F', 24, 3, 200, 10, int_range=8, gamma=.05)
    decision_case('svm_synth_f64_poly', osp, np.float64, 'poly', 'C', 500, 12, 120, 6, degree=3, poly_coefficient=.5)
    decision_case('svm_synth_f32_poly', osp, np.float32, 'poly', 'F', 300, 9, 80, 7, degree=2, poly_coefficient=1.)
    decision_case('svm_synth_f64_lin', osp, np.float64, 'lin', 'F', 600, 40, 100, 8)
    decision_case('svm_synth_f32_lin', osp, np.float32, 'lin', 'C', 600, 40, 100, 9)


if __name__ == '__main__':
    main()
