"""Predictive-model contract and the two kNN models of the reference, on the GPU.

Mirrors osml/predict.py of ViktorC/OSML for the kNN path only:

* ``PredictiveModel``                   osml/predict.py:13-89   (predict / evaluate / test template methods)
* ``RegressionModel``                   osml/predict.py:92-106  (float labels, RMSE)
* ``ClassificationModel``               osml/predict.py:264-278 (integer labels, accuracy)
* ``KNearestNeighbors``                 osml/predict.py:721-774
* ``KNearestNeighborsRegression``       osml/predict.py:777-794
* ``KNearestNeighborsClassification``   osml/predict.py:797-811

Same constructor arguments, same MRO, same bare ``ValueError`` conditions, same ``pd.Series`` wrapping.  What changes
is where the arithmetic runs: ``_fit`` uploads the train matrix to the GPU (``osml_knn_create``) and ``_predict`` is one
call into the CUDA path (``osml_knn_predict``).  There is no CPU fallback.
"""
import os

import numpy as np
import pandas as pd

import osml_b200.model as osm
from osml_b200 import _native


class PredictiveModel(osm.Model, osm._base(1, 'PredictiveModel')):
    """Abstract predictive model (osml/predict.py:13-89)."""

    def __init__(self, epsilon=1e-8):
        super(PredictiveModel, self).__init__(epsilon)

    def _predict(self, observations_df):
        pass

    def _evaluate(self, predictions_sr, labels_sr):
        pass

    def _test(self, observations_df, labels_sr):
        pass

    def predict(self, observations_df):
        """Predicts a label per row (osml/predict.py:29-46).  ValueError for an empty frame, non-numeric columns or
        feature dtypes that differ from the ones the model was fitted on."""
        self._validate_observations(observations_df)
        predictions_sr = pd.Series(self._predict(observations_df), name=self._label_name, dtype=self._label_type)
        self._validate_labels(predictions_sr)
        return predictions_sr

    def evaluate(self, predictions_sr, labels_sr):
        """Scores predictions against labels (osml/predict.py:48-69)."""
        if predictions_sr.count() != labels_sr.count():
            raise ValueError
        self._validate_labels(predictions_sr)
        self._validate_labels(labels_sr)
        return self._evaluate(predictions_sr, labels_sr)

    def test(self, observations_df, labels_sr):
        """Loss on a test set (osml/predict.py:71-89); kNN has none and returns None."""
        self._validate_observations_and_labels(observations_df, labels_sr)
        return self._test(observations_df, labels_sr)


class RegressionModel(PredictiveModel, osm._base(1, 'RegressionModel')):
    """Regression base (osml/predict.py:92-106)."""

    def __init__(self):
        super(RegressionModel, self).__init__()

    def _validate_labels(self, labels_sr):
        super(RegressionModel, self)._validate_labels(labels_sr)
        if not osm.is_continuous(labels_sr.dtype):
            raise ValueError

    def _evaluate(self, predictions_sr, labels_sr):
        # root mean squared error (osml/predict.py:104-106)
        return np.sqrt(np.square(predictions_sr.values - labels_sr.values).mean())


class ClassificationModel(PredictiveModel, osm._base(1, 'ClassificationModel')):
    """Classification base (osml/predict.py:264-278)."""

    def __init__(self):
        super(ClassificationModel, self).__init__()

    def _validate_labels(self, labels_sr):
        super(ClassificationModel, self)._validate_labels(labels_sr)
        if not osm.is_categorical(labels_sr.dtype):
            raise ValueError

    def _evaluate(self, predictions_sr, labels_sr):
        # accuracy by index label (osml/predict.py:276-278; ``iteritems`` was removed in pandas 2, ``items`` is it).
        # With the same unique index on both series ``labels_sr[i]`` is the element at the same position, so the
        # reference's per-item loop (a second per 100 000 predictions) is one vector comparison with the same count.
        if predictions_sr.index.is_unique and predictions_sr.index.equals(labels_sr.index):
            hits = int(np.count_nonzero(predictions_sr.values == labels_sr.values))
        else:
            hits = sum(1 for i, p in predictions_sr.items() if p == labels_sr[i])
        return hits / predictions_sr.count()


def _default_device():
    local_rank = os.environ.get('LOCAL_RANK')
    return int(local_rank) if local_rank is not None else 0


class KNearestNeighbors(PredictiveModel):
    """Weighted k-nearest-neighbours base (osml/predict.py:721-774): RBF-weighted votes of the k nearest, distances
    normalised by the (k+1)-th (Hechenbichler & Schliep).

    GPU-only knobs are attributes with defaults (not constructor arguments, so the reference's signature is kept):
    ``device`` (CUDA ordinal; default ``$LOCAL_RANK`` or 0) and ``path`` ('auto' | 'direct' | 'tensor').
    """
    _task = None

    def __init__(self, k, rbf_gamma):
        super(KNearestNeighbors, self).__init__()
        if k is None or k < 1:
            raise ValueError
        if rbf_gamma <= 0:
            raise ValueError
        self.k = k
        self.rbf_gamma = rbf_gamma
        self._observations_mean = None   # dead fields of the reference (osml/predict.py:735-736), kept for pickles
        self._observations_sd = None
        self._observations = None
        self._labels = None
        self.device = None
        self.path = 'auto'
        self._index = None
        self._query_transform = None     # (shift, denom) fused into predict, see set_query_transform

    # -- device state ---------------------------------------------------------------------------------------------
    def _gamma_is_f32(self):
        gamma = self.rbf_gamma
        return isinstance(gamma, np.floating) and gamma.dtype.itemsize == 4

    def _ensure_index(self):
        """(Re)creates the device index from the host copy; lazy so that unpickled / deep-copied models work."""
        if self._index is None:
            if self._observations is None:
                raise ValueError
            device = self.device if self.device is not None else _default_device()
            self._index = _native.Index(self._observations, self._labels, int(self.k), float(self.rbf_gamma),
                                        float(self.epsilon), device=device, gamma_is_f32=self._gamma_is_f32())
            self._index.set_option('path', _native.PATHS[self.path])
            if getattr(self, '_query_transform', None) is not None:
                self._index.set_query_transform(*self._query_transform)
        return self._index

    def _drop_index(self):
        if self._index is not None:
            self._index.close()
            self._index = None

    def __getstate__(self):
        state = self.__dict__.copy()
        state['_index'] = None          # device handles never travel (pickle, copy.deepcopy, multiprocessing)
        return state

    def __setstate__(self, state):
        self.__dict__.update(state)
        self._index = None

    # -- hooks ------------------------------------------------------------------------------------------------------
    def _fit(self, observations_df, labels_sr):
        if len(observations_df.index) < self.k + 1:
            raise ValueError
        self._drop_index()
        self._observations = observations_df.values      # osml/predict.py:770 (layout decides the summation order)
        self._labels = labels_sr.values                   # osml/predict.py:771
        self._ensure_index()

    def _test(self, observations_df, labels_sr):
        return None

    def _predict(self, observations_df):
        index = self._ensure_index()
        positional = index.predict(observations_df.values, self._task)
        # the reference writes predictions[i] with i the index *label* of each row (osml/predict.py:787-793)
        labels = observations_df.index
        if isinstance(labels, pd.RangeIndex) and labels.start == 0 and labels.step == 1:
            return positional
        predictions = np.zeros((len(labels)), )
        predictions[np.asarray(labels)] = positional
        return predictions

    def set_query_transform(self, transform_model):
        """Not in the reference's API: fuses a fitted ``osml_b200.Centering`` / ``Standardization`` into ``predict``.
        The model must have been fitted on the TRANSFORMED train frame (as in the reference's pipeline:
        ``model.fit(std.fit_transform(x, y), y)``); afterwards ``model.predict(raw_query_df)`` equals
        ``model.predict(std.transform(raw_query_df))`` bit for bit, without the host pass over the query frame.
        Needs float32 / float64 features of one dtype.  ``None`` removes the transform."""
        if transform_model is None:
            self._query_transform = None
        else:
            shift, denom = transform_model.shift_and_denominator()
            self._query_transform = (np.array(shift, copy=True), None if denom is None else np.array(denom, copy=True))
        if self._index is not None:
            self._index.set_query_transform(*(self._query_transform or (None, None)))

    def kneighbors(self, observations_df):
        """Not in the reference's public API: the (k+1) nearest train indices and distances per query, ascending by
        (distance, index) -- what osml/predict.py:748-752 computes internally.  Used by the parity tests."""
        self._validate_observations(observations_df)
        _, idx, dist = self._ensure_index().predict(observations_df.values, self._task, return_neighbours=True)
        return idx, dist


class KNearestNeighborsRegression(KNearestNeighbors, RegressionModel):
    """Weighted kNN regression (osml/predict.py:777-794)."""
    _task = _native.REGRESSION

    def __init__(self, k=7, rbf_gamma=1):
        KNearestNeighbors.__init__(self, k, rbf_gamma)
        RegressionModel.__init__(self)


class KNearestNeighborsClassification(KNearestNeighbors, ClassificationModel):
    """Weighted kNN classification, binary labels included (osml/predict.py:797-811)."""
    _task = _native.CLASSIFICATION

    def __init__(self, k=7, rbf_gamma=1):
        KNearestNeighbors.__init__(self, k, rbf_gamma)
        ClassificationModel.__init__(self)


# -- cross-install pickles (osml/model.py:62-85) ------------------------------------------------
_STATE = ('epsilon', '_observation_types', '_label_type', '_label_name', 'k', 'rbf_gamma', '_observations_mean',
          '_observations_sd', '_observations', '_labels')


def is_reference_knn(model):
    """True for an instance of the reference's own CPU classes osml.predict.KNearestNeighbors{Regression,Classification}."""
    cls = type(model)
    return cls.__module__ == 'osml.predict' and cls.__name__ in ('KNearestNeighborsRegression',
                                                                  'KNearestNeighborsClassification')


def to_reference(model):
    """The reference's CPU model (``osml.predict.KNearestNeighbors*``) with this model's logical state -- what the
    reference's own ``fit`` on the same frame would have left (predict.py:733-738, 770-771; model.py:14-17).  Needs the
    reference package (``reference_modules()``); the result pickles without any trace of this package."""
    mods = osm.reference_modules()
    if mods is None:
        raise ImportError('the reference package (osml) is not importable: set OSML_REF or install it')
    if not isinstance(model, KNearestNeighbors):
        raise ValueError
    cls = getattr(mods[1], type(model).__name__)
    twin = cls(k=model.k, rbf_gamma=model.rbf_gamma)
    for name in _STATE:
        setattr(twin, name, getattr(model, name))
    return twin


def from_reference(model):
    """The GPU model with the state of a reference CPU kNN model (fitted or not); the device index is built lazily."""
    if not is_reference_knn(model):
        raise ValueError
    cls = KNearestNeighborsRegression if type(model).__name__.endswith('Regression') else KNearestNeighborsClassification
    twin = cls(k=model.k, rbf_gamma=model.rbf_gamma)
    for name in _STATE:
        setattr(twin, name, getattr(model, name))
    return twin
