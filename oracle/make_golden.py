"""Generates tests/golden/*.npz by running the REAL reference (oracle/ref_shim.py) in the authoring container.

Run:  python oracle/make_golden.py          (needs /root/reference; not runnable on the GPU box)

For each data set of BASELINE.json configs 1-2 (plus exam / titanic, which the reference's own kNN tests use:
tests/integration/test_predict.py:57-58, 125-128) it stores
  * the exact train / test matrices and labels the reference saw after ``np.random.seed(0)`` + its shuffle/split,
  * the reference's predictions (``model.predict``), its metric (``model.evaluate``),
  * the reference's neighbour lists lifted from its internals (``model._distances`` + predict.py:750-752),
  * per-query tie flags, and the numpy / pandas versions that produced them.
It also stores three synthetic cases (fp32 gaussian F- and C-ordered, int8 wrap-around) run through the reference
classes directly, so that the summation-order and integer wrap semantics are pinned by the reference itself.
"""
import hashlib
import os
import sys

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import knn_oracle as ko  # noqa: E402
from oracle.ref_shim import data_root, import_reference  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), 'tests', 'golden')


def run_reference(osp, model, train_df, train_sr, test_df, test_sr):
    model.fit(train_df, train_sr)
    pred = model.predict(test_df)
    metric = model.evaluate(pred, test_sr) if test_sr is not None else np.nan
    k = model.k
    q_values = test_df.values
    nn_idx = np.zeros((len(test_df.index), k + 1), dtype=np.int64)
    nn_dist = None
    any_tie = np.zeros(len(test_df.index), dtype=bool)
    boundary_tie = np.zeros(len(test_df.index), dtype=bool)
    for qi in range(len(test_df.index)):
        d = model._distances(np.ascontiguousarray(q_values[qi]))       # predict.py:740
        idx = np.argpartition(d, k + 1)[:k + 1]                        # predict.py:750
        ser = pd.Series(d[idx], index=idx).sort_values(ascending=False)  # predict.py:751-752
        if nn_dist is None:
            nn_dist = np.zeros((len(test_df.index), k + 1), dtype=d.dtype)
        nn_idx[qi] = ser.index.values
        nn_dist[qi] = ser.values
        any_tie[qi], boundary_tie[qi] = ko.tie_flags(d, k)
    return pred, metric, nn_idx, nn_dist, any_tie, boundary_tie


def save_case(name, osp, model, task, train_df, train_sr, test_df, test_sr):
    pred, metric, nn_idx, nn_dist, any_tie, boundary_tie = run_reference(osp, model, train_df, train_sr, test_df,
                                                                         test_sr)
    train = train_df.values
    test = test_df.values
    path = os.path.join(OUT, name + '.npz')
    np.savez_compressed(
        path,
        train=np.ascontiguousarray(train), test=np.ascontiguousarray(test),
        train_is_f_order=np.array(ko.sum_order_of(train) == 'sequential'),
        train_labels=train_sr.values, test_labels=(test_sr.values if test_sr is not None else np.zeros(0)),
        k=np.array(model.k), rbf_gamma=np.array(model.rbf_gamma), task=np.array(task),
        ref_predictions=pred.values, ref_metric=np.array(metric),
        ref_nn_idx_desc=nn_idx, ref_nn_dist_desc=nn_dist,
        any_tie=any_tie, boundary_tie=boundary_tie,
        pred_sha256_16=np.array(hashlib.sha256(pred.values.tobytes()).hexdigest()[:16]),
        numpy_version=np.array(np.__version__), pandas_version=np.array(pd.__version__))
    print('%-28s train %s %s (%s) test %s  metric %.16g  sha %s  ties any/boundary %d/%d  %d bytes' % (
        name, train.shape, train.dtype, ko.sum_order_of(train), test.shape, metric,
        hashlib.sha256(pred.values.tobytes()).hexdigest()[:16], any_tie.sum(), boundary_tie.sum(),
        os.path.getsize(path)))


def dataset_case(name, osp, ds, model, task):
    save_case(name, osp, model, task, ds.get_training_observations(), ds.get_training_labels(),
              ds.get_test_observations(), ds.get_test_labels())


def main():
    osp, osd, _ = import_reference()
    root = data_root()
    os.makedirs(OUT, exist_ok=True)
    data = os.path.join(root, 'data')

    def seeded(factory):
        np.random.seed(0)  # tests/integration/__init__.py:3-4, applied per data set
        return factory()

    dataset_case('iris_cls_k7', osp, seeded(lambda: osd.IrisDataSet(os.path.join(data, 'iris/iris.csv'))),
                 osp.KNearestNeighborsClassification(k=7), 'classification')
    dataset_case('boston_reg_k5', osp, seeded(lambda: osd.BostonDataSet(os.path.join(data, 'boston/boston.csv'))),
                 osp.KNearestNeighborsRegression(k=5), 'regression')
    dataset_case('boston_reg_k7', osp, seeded(lambda: osd.BostonDataSet(os.path.join(data, 'boston/boston.csv'))),
                 osp.KNearestNeighborsRegression(), 'regression')
    dataset_case('mushroom_cls_k7', osp,
                 seeded(lambda: osd.MushroomDataSet(os.path.join(data, 'mushroom/mushroom.csv'))),
                 osp.KNearestNeighborsClassification(), 'classification')
    dataset_case('exam_cls_k7', osp,
                 seeded(lambda: osd.ExamDataSet(os.path.join(data, 'exam/ex4x.dat'),
                                                os.path.join(data, 'exam/ex4y.dat'))),
                 osp.KNearestNeighborsClassification(), 'classification')
    dataset_case('titanic_cls_k7', osp,
                 seeded(lambda: osd.TitanicDataSet(os.path.join(data, 'titanic/titanic.csv'))),
                 osp.KNearestNeighborsClassification(), 'classification')

    # synthetic cases through the reference classes: gaussian fp32 (no ties), both memory orders, both tasks
    rng = np.random.default_rng(11)
    n, d, q = 1000, 37, 48
    tr = rng.standard_normal((n, d), dtype=np.float32)
    te = rng.standard_normal((q, d), dtype=np.float32)
    ycls = pd.Series(rng.integers(0, 10, n), name='y')
    yreg = pd.Series(rng.standard_normal(n), name='y')
    save_case('synth_f32_forder_cls_k10', osp, osp.KNearestNeighborsClassification(k=10), 'classification',
              pd.DataFrame(tr), ycls, pd.DataFrame(te), None)
    save_case('synth_f32_forder_reg_k10', osp, osp.KNearestNeighborsRegression(k=10, rbf_gamma=0.5), 'regression',
              pd.DataFrame(tr), yreg, pd.DataFrame(te), None)

    # C-ordered train matrix: pandas 3 always hands out F order, so drive the reference's hooks directly
    # (predict.py:767-771 stores whatever .values returns; older pandas returned the C-ordered array)
    class COrderFrame(pd.DataFrame):
        @property
        def values(self):
            return np.ascontiguousarray(super().values)
    n2, d2 = 320, 200
    tr2 = rng.standard_normal((n2, d2), dtype=np.float32)
    te2 = rng.standard_normal((24, d2), dtype=np.float32)
    save_case('synth_f32_corder_reg_k10', osp, osp.KNearestNeighborsRegression(k=10), 'regression',
              COrderFrame(tr2), pd.Series(rng.standard_normal(n2), name='y'), pd.DataFrame(te2), None)

    # int8 with wrap-around in subtract and square
    tr3 = rng.integers(-128, 128, (1500, 9)).astype(np.int8)
    te3 = rng.integers(-128, 128, (48, 9)).astype(np.int8)
    save_case('synth_i8_wrap_cls_k7', osp, osp.KNearestNeighborsClassification(k=7), 'classification',
              pd.DataFrame(tr3), pd.Series(rng.integers(0, 4, 1500).astype(np.int8), name='y'), pd.DataFrame(te3),
              None)
    # fp64 gaussian, k=32 (config-5-like selection width)
    tr4 = rng.standard_normal((900, 24))
    te4 = rng.standard_normal((40, 24))
    save_case('synth_f64_cls_k32', osp, osp.KNearestNeighborsClassification(k=32, rbf_gamma=2.5), 'classification',
              pd.DataFrame(tr4), pd.Series(rng.integers(0, 5, 900), name='y'), pd.DataFrame(te4), None)


if __name__ == '__main__':
    main()
