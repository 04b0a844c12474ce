"""CPU restatement (numpy) of the two pre-processing transforms the reference's README puts in front of kNN:
``Centering`` (osml/transform.py:53-66) and ``Standardization`` (osml/transform.py:69-84).

TEST INFRASTRUCTURE ONLY: imported by ``tests/`` (and ``oracle/make_golden_transform.py``), never by ``osml_b200``.
Parity pin: ``tests/golden/transform_*.npz`` were produced by the REAL reference classes in the authoring container
(``oracle/make_golden_transform.py``); ``tests/test_oracle_transform.py`` checks this restatement bit-for-bit against
them.

Arithmetic, as numpy does it for a frame whose columns share one dtype T:
  * ``values.mean(axis=0)`` / ``values.std(axis=0)`` -- numpy reductions in T (pairwise within a column); they stay on
    the host in the build (fit time, once), so parity is by construction,
  * ``values - mean``: one rounding in T for float T; for integer T the mean is float64 and the result is float64,
  * ``/ (sd + epsilon)``: the Python-float epsilon is a weak scalar (NEP 50): ``sd + epsilon`` is rounded once in the
    dtype of ``sd``, and the division is rounded once in that dtype.
The result keeps the memory order of ``values`` (elementwise ufuncs preserve F order), which decides the summation
order of the kNN distances downstream (oracle/knn_oracle.py).
"""
import numpy as np


def centering_fit(values):
    """osml/transform.py:61-62."""
    return values.mean(axis=0)


def centering_transform(values, mean):
    """osml/transform.py:64-65."""
    return values - mean


def standardization_fit(values):
    """osml/transform.py:78-81."""
    return values.mean(axis=0), values.std(axis=0)


def standardization_transform(values, mean, sd, epsilon=1e-8):
    """osml/transform.py:83-84."""
    return (values - mean) / (sd + epsilon)


def shift_and_denominator(mean, sd=None, epsilon=1e-8):
    """The two per-column vectors the GPU kernel takes: ``mean`` and ``sd + epsilon`` exactly as numpy rounds it."""
    return mean, (None if sd is None else sd + epsilon)
