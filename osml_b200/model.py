"""Model contract of the reference (osml/model.py:7-111), restated for the GPU kNN path.

Same template method (``fit`` records dtypes and label metadata, validates, then calls the ``_fit`` hook), same bare
``ValueError`` at the same conditions, same pickle-based ``save`` / ``load``.
"""
import importlib
import os
import pickle
import sys

import numpy as np


def reference_modules():
    """The reference's own ``osml.model`` / ``osml.predict`` / ``osml.transform`` when that package is importable --
    installed next to this one, or at ``$OSML_REF`` -- else ``None``.

    When it is there, every class of this package is built ON the reference's class of the same name (see ``_base``),
    so the ``isinstance(base_model, PredictiveModel / RegressionModel / ClassificationModel)`` checks of the reference's
    meta-models (osml/predict.py:1326, 1393, 1410, 1427) accept a GPU model as their base model.  Every method of the
    contract is overridden here, so behaviour does not depend on which base was picked.
    ``OSML_B200_STANDALONE=1`` forces the standalone bases."""
    if os.environ.get('OSML_B200_STANDALONE'):
        return None
    if 'mods' not in _REFERENCE:
        _REFERENCE['mods'] = _import_reference()
    return _REFERENCE['mods']


_REFERENCE = {}


def _import_reference():
    def attempt():
        mods = tuple(importlib.import_module('osml.' + name) for name in ('model', 'predict', 'transform'))
        ok = all(hasattr(mods[1], name) for name in ('PredictiveModel', 'RegressionModel', 'ClassificationModel'))
        return mods if ok and hasattr(mods[0], 'Model') else None
    try:
        return attempt()
    except Exception:
        pass
    root = os.environ.get('OSML_REF')
    if root and os.path.isfile(os.path.join(root, 'osml', 'predict.py')):
        sys.path.append(root)
        try:
            return attempt()
        except Exception:
            pass
        sys.path.remove(root)
    return None


def _base(module_index, name):
    """The reference class to build on, or ``object``."""
    mods = reference_modules()
    return getattr(mods[module_index], name, object) if mods else object


def is_continuous(data_type):
    """True for floating point dtypes (osml/model.py:88-98)."""
    return np.issubdtype(data_type, np.floating)


def is_categorical(data_type):
    """True for integer dtypes (osml/model.py:101-111)."""
    return np.issubdtype(data_type, np.integer)


class Model(_base(0, 'Model')):
    """Base of every model (osml/model.py:7-59)."""

    def __init__(self, epsilon):
        if epsilon is None or epsilon <= 0:
            raise ValueError
        self.epsilon = epsilon
        self._observation_types = None
        self._label_type = None
        self._label_name = ''

    # -- validation (osml/model.py:19-34) ------------------------------------------------------------------------
    def _validate_observations(self, observations_df):
        dtypes = observations_df.dtypes
        if len(dtypes) == 0 or len(observations_df.index) == 0:
            raise ValueError
        # the reference tests every column (osml/model.py:24); the distinct dtypes decide the same thing, and a frame of
        # 784 columns costs 0.02 ms here instead of 0.8 ms -- a third of a 10 000-query predict call on the GPU
        if any(not np.issubdtype(dtype, np.number) for dtype in set(dtypes.values)):
            raise ValueError
        if not dtypes.equals(self._observation_types):
            raise ValueError

    def _validate_labels(self, labels_sr):
        if len(labels_sr.index) == 0:
            raise ValueError
        if not np.issubdtype(labels_sr.dtype, np.number) or labels_sr.dtype != self._label_type:
            raise ValueError

    def _validate_observations_and_labels(self, observations_df, labels_sr):
        self._validate_observations(observations_df)
        self._validate_labels(labels_sr)
        if len(observations_df.index) != len(labels_sr.index):
            raise ValueError

    # -- hooks ------------------------------------------------------------------------------------------------------
    def _fit(self, observations_df, labels_sr):
        pass

    def fit(self, observations_df, labels_sr):
        """Fits the model (osml/model.py:39-59): records the feature dtypes, label dtype and label name, validates,
        then hands over to ``_fit``.  Raises ValueError on empty / non-numeric / length-mismatched input."""
        self._observation_types = observations_df.dtypes
        if labels_sr is None:
            self._validate_observations(observations_df)
        else:
            self._label_type = labels_sr.dtype
            self._label_name = labels_sr.name
            self._validate_observations_and_labels(observations_df, labels_sr)
        return self._fit(observations_df, labels_sr)


def save(model, file_path, portable=False):
    """Pickles the model to ``file_path`` (osml/model.py:62-71).  Device handles are dropped by the model's
    ``__getstate__``; the logical state matches the reference's (``_observations``, ``_labels``, ``k``, ...).
    ``portable=True`` writes a kNN model as the REFERENCE's class (``osml.predict.KNearestNeighbors*``, see
    ``osml_b200.predict.to_reference``), so the file also loads where only the reference is installed."""
    if portable:
        from osml_b200.predict import to_reference
        model = to_reference(model)
    with open(file_path, mode='wb') as file:
        pickle.dump(model, file)


def load(file_path):
    """Inverse of ``save`` (osml/model.py:74-85).  A file that holds the reference's CPU kNN model (written by the
    reference's own ``save``, or by ``save(..., portable=True)``) comes back as the GPU model with the same state."""
    with open(file_path, mode='rb') as file:
        model = pickle.load(file)
    from osml_b200.predict import from_reference, is_reference_knn
    return from_reference(model) if is_reference_knn(model) else model
