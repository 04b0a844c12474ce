"""osml_b200 -- B200-native brute-force weighted kNN predict behind ViktorC/OSML's model API.

Only the path BASELINE.json's north_star names is here: ``KNearestNeighborsRegression`` and
``KNearestNeighborsClassification`` (osml/predict.py:721-811 of the reference) with the reference's
``fit / predict / evaluate / test`` contract, computed by hand-written sm_100a CUDA kernels behind the C ABI of
``include/osml_knn.h`` -- plus the widening rows of SURVEY.md section 8: the two transforms the reference's README puts
in front of kNN (``Centering``, ``Standardization``: osml/transform.py:53-84), which can be fused into ``predict``, and
``SupportVectorMachine`` (osml/predict.py:1050-1250) with its kernel evaluations against all rows (decision function,
SMO error cache) behind ``include/osml_svm.h``.  There is no CPU fallback: without the built library and a CUDA device
``fit`` raises.
"""
from osml_b200.model import Model, load, save, is_categorical, is_continuous, reference_modules  # noqa: F401
from osml_b200.predict import (  # noqa: F401
    ClassificationModel, KNearestNeighbors, KNearestNeighborsClassification, KNearestNeighborsRegression,
    PredictiveModel, RegressionModel, from_reference, to_reference)
from osml_b200.svm import BinaryClassificationModel, SupportVectorMachine  # noqa: F401
from osml_b200.transform import Centering, Standardization, TransformativeModel  # noqa: F401

__all__ = ['Model', 'PredictiveModel', 'RegressionModel', 'ClassificationModel', 'KNearestNeighbors',
           'KNearestNeighborsRegression', 'KNearestNeighborsClassification', 'save', 'load', 'is_continuous',
           'is_categorical', 'TransformativeModel', 'Centering', 'Standardization', 'BinaryClassificationModel',
           'SupportVectorMachine', 'reference_modules', 'to_reference', 'from_reference']
