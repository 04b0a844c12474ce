"""Model contract of the reference (osml/model.py:7-111), restated for the GPU kNN path.

Same template method (``fit`` records dtypes and label metadata, validates, then calls the ``_fit`` hook), same bare
``ValueError`` at the same conditions, same pickle-based ``save`` / ``load``.
"""
import pickle

import numpy as np


def is_continuous(data_type):
    """True for floating point dtypes (osml/model.py:88-98)."""
    return np.issubdtype(data_type, np.floating)


def is_categorical(data_type):
    """True for integer dtypes (osml/model.py:101-111)."""
    return np.issubdtype(data_type, np.integer)


class Model:
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
        if any(not np.issubdtype(dtype, np.number) for dtype in dtypes):
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


def save(model, file_path):
    """Pickles the model to ``file_path`` (osml/model.py:62-71).  Device handles are dropped by the model's
    ``__getstate__``; the logical state matches the reference's (``_observations``, ``_labels``, ``k``, ...)."""
    with open(file_path, mode='wb') as file:
        pickle.dump(model, file)


def load(file_path):
    """Inverse of ``save`` (osml/model.py:74-85)."""
    with open(file_path, mode='rb') as file:
        return pickle.load(file)
