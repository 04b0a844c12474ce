"""The two pre-processing transforms the reference's README puts in front of kNN, on the GPU.

Mirrors osml/transform.py of ViktorC/OSML for this part only:

* ``TransformativeModel``   osml/transform.py:9-50   (transform / fit_transform template methods)
* ``Centering``             osml/transform.py:53-66
* ``Standardization``       osml/transform.py:69-84

``_fit`` keeps the reference's numpy statistics (``values.mean(axis=0)`` / ``values.std(axis=0)``: fit time, once, and
bit-identical by construction); ``_transform`` -- the per-call pass over the whole frame -- runs in the CUDA kernel
behind ``osml_knn_column_transform`` and returns a frame with the same dtype, values and memory order as the
reference's.  A fitted transform can also be attached to a kNN model (``model.set_query_transform(std)``): ``predict``
then takes the RAW query frame and the transform is fused into the upload on the device.  No CPU fallback.
"""
import numpy as np
import pandas as pd

import osml_b200.model as osm
from osml_b200 import _native


class TransformativeModel(osm.Model, osm._base(2, 'TransformativeModel')):
    """Abstract data transformation model (osml/transform.py:9-50)."""

    def __init__(self, epsilon=1e-8):
        super(TransformativeModel, self).__init__(epsilon)
        self.device = None

    def _transform(self, observations_df):
        pass

    def transform(self, observations_df):
        """Transforms the frame with what ``fit`` learnt (osml/transform.py:19-31)."""
        self._validate_observations(observations_df)
        return pd.DataFrame(self._transform(observations_df))

    def fit_transform(self, observations_df, labels_sr):
        """``fit`` then ``transform`` (osml/transform.py:33-50)."""
        self.fit(observations_df, labels_sr)
        return self.transform(observations_df)

    def _device(self):
        from osml_b200.predict import _default_device
        return self.device if self.device is not None else _default_device()

    def shift_and_denominator(self):
        """(shift, denom) per column as the kernel takes them; denom is None for a pure shift."""
        raise NotImplementedError


class Centering(TransformativeModel):
    """Subtracts the column means of the fit frame (osml/transform.py:53-66)."""

    def __init__(self):
        super(Centering, self).__init__()
        self._mean = None

    def _fit(self, observations_df, labels_sr):
        self._mean = observations_df.values.mean(axis=0)                 # osml/transform.py:61-62

    def shift_and_denominator(self):
        if self._mean is None:
            raise ValueError
        return self._mean, None

    def _transform(self, observations_df):
        shift, _ = self.shift_and_denominator()
        return _native.column_transform(observations_df.values, shift, None, self._device())   # transform.py:64-65


class Standardization(TransformativeModel):
    """Centres and scales by the column standard deviations of the fit frame (osml/transform.py:69-84)."""

    def __init__(self):
        super(Standardization, self).__init__()
        self._mean = None
        self._sd = None

    def _fit(self, observations_df, labels_sr):
        observations = observations_df.values
        self._mean = observations.mean(axis=0)                           # osml/transform.py:79-81
        self._sd = observations.std(axis=0)

    def shift_and_denominator(self):
        if self._mean is None:
            raise ValueError
        # ``sd + epsilon`` exactly as numpy rounds it (the Python-float epsilon is a weak scalar)
        return self._mean, self._sd + self.epsilon

    def _transform(self, observations_df):
        shift, denom = self.shift_and_denominator()
        return _native.column_transform(observations_df.values, shift, denom, self._device())  # transform.py:83-84
