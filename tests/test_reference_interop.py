"""Interplay with the reference package itself: GPU models as base models of the
reference's meta-models, pickles that travel between a reference-only install and this package, and INTEGRATION.md
section B's two-hook patch.  CPU tests use doubles for the device; the `gpu` ones run the real thing.  Everything here
needs the reference package (baseline/_ref or /root/reference) and is skipped, with the reason, where it is absent."""
import copy
import os
import pickle
import re
import subprocess
import sys
import types

import numpy as np
import pandas as pd
import pytest

import osml_b200 as ob
from oracle import knn_oracle as ko
from tests.conftest import ROOT, load_golden

try:
    from oracle.ref_shim import import_reference
    osp, _, REF_ROOT = import_reference()
except ImportError:                                    # pragma: no cover
    osp, REF_ROOT = None, None

needs_reference = pytest.mark.skipif(osp is None or ob.reference_modules() is None,
                                     reason='the reference package (osml) is not reachable: no baseline/_ref, '
                                            '/root/reference or $OSML_REF')


# ---- f1: isinstance contract of the meta-models ------------------------------------------------------------------------
@needs_reference
def test_gpu_models_are_instances_of_the_reference_bases():
    cls, reg = ob.KNearestNeighborsClassification(k=3), ob.KNearestNeighborsRegression(k=3)
    assert isinstance(cls, osp.PredictiveModel) and isinstance(cls, osp.ClassificationModel)
    assert isinstance(reg, osp.PredictiveModel) and isinstance(reg, osp.RegressionModel)
    assert not isinstance(cls, osp.RegressionModel) and not isinstance(reg, osp.ClassificationModel)
    assert isinstance(ob.SupportVectorMachine(), osp.BinaryClassificationModel)
    # the constructor checks at osml/predict.py:1326, 1393, 1410, 1427 accept them ...
    osp.BootstrapAggregatingClassification(cls, 4, number_of_processes=1)
    osp.BootstrapAggregatingRegression(reg, 4, number_of_processes=1)
    osp.GradientBoostingRegression(reg, number_of_models=2)
    # ... and still reject the wrong kind
    with pytest.raises(ValueError):
        osp.BootstrapAggregatingClassification(reg, 4, number_of_processes=1)
    with pytest.raises(ValueError):
        osp.BootstrapAggregatingRegression(cls, 4, number_of_processes=1)
    # every method of the contract is this package's own, whatever the base
    for name in ('fit', 'predict', 'evaluate', 'test', '_validate_observations', '_validate_labels'):
        assert getattr(type(cls), name).__module__.startswith('osml_b200')


def test_standalone_bases_without_the_reference():
    code = ('import osml_b200 as ob, sys; '
            'assert ob.reference_modules() is None; '
            'assert all(c.__module__.startswith("osml_b200") or c is object '
            '           for c in ob.KNearestNeighborsRegression.__mro__); '
            'assert "osml" not in sys.modules; print("ok")')
    env = dict(os.environ, OSML_B200_STANDALONE='1')
    res = subprocess.run([sys.executable, '-c', code], cwd=ROOT, env=env, stdout=subprocess.PIPE,
                         stderr=subprocess.PIPE, text=True, timeout=300)
    assert res.returncode == 0 and res.stdout.strip() == 'ok', res.stderr[-2000:]


# ---- f2: pickles between a reference-only install and this package ----------------------------------------------------
def _fitted_state_without_a_device(model, g):
    """What fit() leaves behind (osml_b200/predict.py `_fit`, osml/model.py:39-59) minus the device index."""
    train_df = pd.DataFrame(g['train'])
    labels_sr = pd.Series(g['train_labels'], name='label')
    model._observation_types = train_df.dtypes
    model._label_type, model._label_name = labels_sr.dtype, labels_sr.name
    model._observations, model._labels = train_df.values, labels_sr.values
    return train_df, labels_sr


@needs_reference
@pytest.mark.parametrize('name', ['iris_cls_k7', 'boston_reg_k5'])
def test_portable_pickle_loads_in_a_reference_only_process(tmp_path, name):
    g = load_golden(name)
    cls = ob.KNearestNeighborsRegression if g['task'] == 'regression' else ob.KNearestNeighborsClassification
    model = cls(k=g['k'], rbf_gamma=g['rbf_gamma'])
    _fitted_state_without_a_device(model, g)
    path = str(tmp_path / 'model.pkl')
    ob.save(model, path, portable=True)
    np.save(str(tmp_path / 'queries.npy'), g['test'])
    # a process that can import the reference but NOT this package loads the file and predicts on the CPU
    code = '''
import sys, pickle, numpy as np, pandas as pd
assert not any('osml_b200' in m for m in sys.modules)
if not hasattr(np, 'float_'): np.float_ = np.float64
if not hasattr(pd.Series, 'iteritems'): pd.Series.iteritems = pd.Series.items
sys.path.insert(0, %r)
import osml.model
model = osml.model.load(%r)
assert type(model).__module__ == 'osml.predict', type(model)
pred = model.predict(pd.DataFrame(np.load(%r)))
np.save(%r, pred.values)
assert not any('osml_b200' in m for m in sys.modules)
''' % (REF_ROOT, path, str(tmp_path / 'queries.npy'), str(tmp_path / 'pred.npy'))
    res = subprocess.run([sys.executable, '-c', code], cwd=str(tmp_path), stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                         text=True, timeout=600, env={k: v for k, v in os.environ.items() if k != 'PYTHONPATH'})
    assert res.returncode == 0, res.stderr[-3000:]
    assert np.array_equal(np.load(str(tmp_path / 'pred.npy')), g['ref_predictions'], equal_nan=True)


@needs_reference
def test_reference_pickle_loads_as_the_gpu_model(tmp_path):
    g = load_golden('iris_cls_k7')
    ref_model = osp.KNearestNeighborsClassification(k=g['k'])
    ref_model.fit(pd.DataFrame(g['train']), pd.Series(g['train_labels'], name='label'))
    path = str(tmp_path / 'ref.pkl')
    with open(path, 'wb') as f:
        pickle.dump(ref_model, f)                      # = osml.model.save (osml/model.py:62-71)
    model = ob.load(path)
    assert type(model) is ob.KNearestNeighborsClassification and model._index is None
    assert model.k == g['k'] and model._label_name == 'label' and model._observations is not None
    assert np.array_equal(model._observations, g['train']) and model._observation_types.equals(ref_model._observation_types)
    back = ob.to_reference(model)
    assert type(back) is osp.KNearestNeighborsClassification
    assert np.array_equal(back.predict(pd.DataFrame(g['test'])).values, g['ref_predictions'])
    # GPU-class pickles keep working as before and deep copies stay GPU models
    again = pickle.loads(pickle.dumps(model))
    assert type(again) is ob.KNearestNeighborsClassification and type(copy.deepcopy(model)) is type(model)


# ---- the two-hook patch of the reference ----------------------------------------------------
def _patched_reference_classes(knn_gpu):
    """The reference's own classes with the two hooks replaced as INTEGRATION.md section B shows (subclasses here, so
    the imported reference stays untouched for the other tests)."""
    def _fit(self, observations_df, labels_sr):                        # was predict.py:767-771
        if len(observations_df.index) < self.k + 1:
            raise ValueError
        self._observations = observations_df.values
        self._labels = labels_sr.values
        self._gpu = knn_gpu.create(self._observations, self._labels, self.k, self.rbf_gamma, self.epsilon)

    def make_predict(task):
        def _predict(self, observations_df):                           # was predict.py:785-794 / 805-811
            positional = knn_gpu.predict(self._gpu, observations_df.values, task)
            predictions = np.zeros((len(observations_df.index)), )
            predictions[np.asarray(observations_df.index)] = positional
            return predictions
        return _predict
    reg = type('KNearestNeighborsRegression', (osp.KNearestNeighborsRegression,), {'_fit': _fit, '_predict': make_predict(0)})
    cls = type('KNearestNeighborsClassification', (osp.KNearestNeighborsClassification,),
               {'_fit': _fit, '_predict': make_predict(1)})
    return reg, cls


def _integration_stub_source():
    text = open(os.path.join(ROOT, 'INTEGRATION.md')).read()
    blocks = re.findall(r'```python\n(.*?)```', text, flags=re.S)
    stub = [b for b in blocks if b.startswith('# osml/_knn_gpu.py')]
    assert len(stub) == 1
    return stub[0]


def _run_patched(knn_gpu, g):
    reg, cls = _patched_reference_classes(knn_gpu)
    model = (reg if g['task'] == 'regression' else cls)(k=g['k'], rbf_gamma=g['rbf_gamma'])
    model.fit(pd.DataFrame(g['train']), pd.Series(g['train_labels'], name='label'))
    return model.predict(pd.DataFrame(g['test']))


@needs_reference
@pytest.mark.parametrize('name', ['iris_cls_k7', 'boston_reg_k5', 'synth_f32_f32labels_reg_k10'])
def test_integration_patch_with_a_fake_device_module(name):
    """The patch's plumbing (what is passed to create / predict, positional scatter, Series wrapping by the untouched
    template methods) with a CPU double behind the `_knn_gpu` interface."""
    g = load_golden(name)
    fake = types.SimpleNamespace()
    fake.create = lambda obs, labels, k, gamma, eps, device=0: dict(obs=obs, labels=labels, k=k, gamma=gamma, eps=eps)
    fake.predict = lambda h, q, task: (ko.predict_regression if task == 0 else ko.predict_classification)(
        h['obs'], h['labels'], q, k=h['k'], rbf_gamma=h['gamma'], epsilon=h['eps'], tie_rule='reference')
    pred = _run_patched(fake, g)
    assert pred.dtype == g['ref_predictions'].dtype and pred.name == 'label'
    assert np.array_equal(pred.values, g['ref_predictions'], equal_nan=True)


def test_integration_stub_matches_the_header():
    """The ctypes stub printed in INTEGRATION.md binds the argument lists include/osml_knn.h declares."""
    src = _integration_stub_source()
    header = open(os.path.join(ROOT, 'include', 'osml_knn.h')).read()
    for symbol in ('osml_knn_create', 'osml_knn_predict'):
        proto = re.search(r'int %s\((.*?)\);' % symbol, header, flags=re.S).group(1)
        n_args = len([a for a in proto.split(',') if a.strip()])
        bound = re.search(r'_lib\.%s\.argtypes = \[(.*?)\]' % symbol, src, flags=re.S).group(1)
        assert len([a for a in bound.replace('\n', ' ').split(',') if a.strip()]) == n_args, symbol


@needs_reference
@pytest.mark.gpu
@pytest.mark.parametrize('name', ['iris_cls_k7', 'boston_reg_k5', 'mushroom_cls_k7', 'synth_f32_f32labels_reg_k10'])
def test_integration_stub_from_the_document_drives_the_gpu(name):
    """INTEGRATION.md's stub, executed as printed (library path filled in), behind the patched reference classes."""
    from osml_b200 import _native
    from tests import parity
    g = load_golden(name)
    src = _integration_stub_source().replace("ctypes.CDLL('libosml_knn.so')", 'ctypes.CDLL(%r)' % _native.library_path())
    knn_gpu = types.ModuleType('_knn_gpu')
    exec(compile(src, 'INTEGRATION.md:_knn_gpu', 'exec'), knn_gpu.__dict__)
    pred = _run_patched(knn_gpu, g)
    assert pred.dtype == g['ref_predictions'].dtype
    clean = ~g['any_tie']
    parity.assert_predictions_match(pred.values.astype(np.float64)[clean],
                                    g['ref_predictions'].astype(np.float64)[clean], g['task'])


# ---- f1 for real: a GPU kNN as the base model of the reference's bagging meta-model -----------------------------------
def _bagging_data():
    rng = np.random.default_rng(99)
    n, d, q = 600, 8, 120
    train = pd.DataFrame(rng.standard_normal((n, d), dtype=np.float32))
    labels = pd.Series(rng.integers(0, 3, n), name='label')
    queries = pd.DataFrame(rng.standard_normal((q, d), dtype=np.float32))
    return train, labels, queries


@needs_reference
@pytest.mark.gpu
def test_bagging_meta_model_over_a_gpu_knn_equals_the_cpu_reference_single_process():
    train, labels, queries = _bagging_data()
    preds = []
    for base in (ob.KNearestNeighborsClassification(k=5), osp.KNearestNeighborsClassification(k=5)):
        np.random.seed(5)                              # the bootstrap samples come from np.random (predict.py:1349)
        meta = osp.BootstrapAggregatingClassification(base, 4, number_of_processes=1)
        meta.fit(train, labels)
        assert all(type(m) is type(base) for m in meta._models)
        preds.append(meta.predict(queries).values)
    assert np.array_equal(preds[0], preds[1])          # gaussian data: no distance ties, labels bit-exact


@needs_reference
@pytest.mark.gpu
def test_bagging_meta_model_over_a_gpu_knn_two_spawned_processes():
    """BootstrapAggregatingClassification(GPU kNN, 4 models, 2 processes) with the `spawn` start method (a forked child
    of a CUDA-initialised parent cannot use the device: INTEGRATION.md): the workers fit and predict on the GPU, models
    travel as pickles without their device index.  Each worker draws its own bootstrap sample, so the check is per
    sub-model: the reference's CPU kNN fitted on the SAME sample must predict the same labels, and the meta-model's
    vote must be the reference's aggregation of those."""
    import multiprocessing as mp
    train, labels, queries = _bagging_data()
    old_method = mp.get_start_method(allow_none=True)
    old_path = os.environ.get('PYTHONPATH')
    os.environ['PYTHONPATH'] = os.pathsep.join([ROOT, REF_ROOT] + ([old_path] if old_path else []))
    mp.set_start_method('spawn', force=True)
    try:
        meta = osp.BootstrapAggregatingClassification(ob.KNearestNeighborsClassification(k=5), 4, number_of_processes=2)
        meta.fit(train, labels)
        got = meta.predict(queries)
    finally:
        mp.set_start_method(old_method, force=True)
        if old_path is None:
            os.environ.pop('PYTHONPATH', None)
        else:
            os.environ['PYTHONPATH'] = old_path
    assert len(meta._models) == 4 and all(type(m) is ob.KNearestNeighborsClassification for m in meta._models)
    assert all(m._index is None for m in meta._models)           # device handles never travel
    per_model = []
    for m in meta._models:
        twin = ob.to_reference(m)                                 # the reference's CPU model on the same bootstrap sample
        want = twin.predict(queries)
        assert np.array_equal(m.predict(queries).values, want.values)
        per_model.append(want)
    assert np.array_equal(got.values, pd.Series(meta._aggregate_predictions(per_model), dtype=got.dtype).values)
