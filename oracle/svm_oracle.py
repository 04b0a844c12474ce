"""CPU oracle of the reference's support-vector decision function.

TEST INFRASTRUCTURE ONLY: imported by ``tests/`` (and the golden generator); never by ``osml_b200``.

Restates, in numpy, what ``SupportVectorMachine`` of ViktorC/OSML computes once it is fitted
(osml/predict.py:1086-1099, 1238-1250) and ``BinaryClassificationModel._predict`` (osml/predict.py:126-127):

    kernel value   lin   np.dot(V, x)                                   predict.py:1086-1087
                   poly  (coefficient + np.dot(V, x)) ** degree         predict.py:1089-1090
                   rbf   exp(-gamma * sqrt(sum_j (V[i, j] - x[j])^2))   predict.py:1092-1093  (distance, not squared)
    raw            intercept + sum_i K(V_i, x) * alpha_i * label_i      predict.py:1095-1099
    fuzzy          raw / (2 * max|raw| + epsilon) + 0.5                 predict.py:1245-1250
    label          rint(fuzzy)                                          predict.py:126-127
    hinge loss     sum max(1 - y * raw, 0)                              predict.py:1238-1243

The arithmetic is dtype preserving exactly as in the reference: for float32 vectors the distance, the product with
``-gamma`` (a weak Python scalar) and ``exp`` stay float32 and only the product with the float64 multipliers widens.
The distance sum follows the memory order of ``vectors`` (``observations[alphas > 0]`` is C-ordered: numpy's pairwise
tree along the features; the F-ordered ``DataFrame.values`` used while fitting adds feature by feature).

Parity pin: ``tests/golden/svm_*.npz`` hold what the REAL reference returned for the same inputs in the authoring
container (``oracle/make_golden_svm.py``); ``tests/test_oracle_svm.py`` checks this file against them bit for bit.
"""
import numpy as np

KERNELS = ('lin', 'poly', 'rbf')


def kernel_values(vectors, x, kernel, gamma=.5, degree=2, poly_coefficient=1.):
    """K(V_i, x) for every row of ``vectors`` (osml/predict.py:1086-1093)."""
    if kernel == 'lin':
        return np.dot(vectors, x.T)
    if kernel == 'poly':
        return np.power(poly_coefficient + np.dot(vectors, x.T), degree)
    if kernel == 'rbf':
        return np.exp(-gamma * np.sqrt(np.square(vectors - x).sum(axis=1)))
    raise ValueError(kernel)


def raw_predict(queries, vectors, labels, alphas, intercept, kernel, gamma=.5, degree=2, poly_coefficient=1.):
    """Decision values, one per query row (osml/predict.py:1095-1099).  ``labels`` are -1 / +1 floats."""
    out = np.full(len(queries), float(intercept))
    for i, x in enumerate(queries):
        out[i] += np.sum(kernel_values(vectors, x, kernel, gamma, degree, poly_coefficient) * alphas * labels)
    return out


def fuzzy_from_raw(raw, epsilon=1e-8):
    """Maps decision values to [0, 1] (osml/predict.py:1249)."""
    return raw / (np.abs(raw).max() * 2 + epsilon) + .5


def labels_from_fuzzy(fuzzy):
    """osml/predict.py:126-127."""
    return np.rint(fuzzy)


def hinge_loss(raw, labels01):
    """osml/predict.py:1241-1243 (labels in {0, 1})."""
    return np.sum(np.maximum(1. - (labels01 * 2. - 1.) * raw, .0))


def error_cache(observations, labels, alphas, intercept, kernel, gamma=.5, degree=2, poly_coefficient=1.):
    """The SMO error cache after a step (osml/predict.py:1101-1105): decision value of every training row against ALL
    training rows (multipliers that are zero included) minus its -1 / +1 label."""
    return raw_predict(observations, observations, labels, alphas, intercept, kernel, gamma, degree,
                       poly_coefficient) - labels
