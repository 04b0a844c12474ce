"""Support vector machine of the reference with its kernel evaluations on the GPU.

Mirrors osml/predict.py of ViktorC/OSML:

* ``BinaryClassificationModel``   osml/predict.py:109-262   ({0, 1} integer labels, F1, precision / recall / fall-out / ROC)
* ``SupportVectorMachine``        osml/predict.py:1050-1250 (SMO; 'lin' | 'poly' | 'rbf' kernels)

Same constructor arguments, attributes, bare ``ValueError`` conditions and ``pd.Series`` wrapping.  What moves to the
GPU is every place the reference evaluates its kernel against ALL rows of a matrix -- ``_raw_predict``
(osml/predict.py:1095-1099), i.e. ``predict`` / ``fuzzy_predict`` / ``test`` and the error cache SMO recomputes after
every successful step (``_update_error_cache``, 1101-1105: N decision values over N rows, the fit-time hot spot) -- as
one call into ``osml_svm_decision`` / ``osml_svm_decision_self`` (include/osml_svm.h) -- and the scalar kernel values
of SMO's step attempts, taken from a table of all training pairs computed once on the device (``osml_svm_gram``).  The
SMO bookkeeping itself is host control flow as in the reference (the partner scan tests all candidates of a row in one
vector pass instead of one by one: same first success) and draws from ``np.random`` in the same order, so with the same seed it walks
the same path as long as no comparison is decided by the last bits of an error value (the GPU sums in a different
order: ~1e-16 relative).  There is no CPU fallback for the decision function.
"""
import numpy as np
import pandas as pd

import osml_b200.model as osm
from osml_b200 import _native
from osml_b200.predict import PredictiveModel, _default_device


class BinaryClassificationModel(PredictiveModel, osm._base(1, 'BinaryClassificationModel')):
    """Binary classification base (osml/predict.py:109-262): labels are integers in {0, 1}; ``_fuzzy_predict`` gives
    the positive-class membership in [0, 1] and ``predict`` rounds it."""

    def __init__(self):
        super(BinaryClassificationModel, self).__init__()

    def _fuzzy_predict(self, observations_df):
        pass

    def _validate_labels(self, labels_sr):
        super(BinaryClassificationModel, self)._validate_labels(labels_sr)
        if not osm.is_categorical(labels_sr.dtype) or any(v != 0 and v != 1 for v in labels_sr.unique()):
            raise ValueError

    def _predict(self, observations_df):
        return np.rint(self._fuzzy_predict(observations_df))

    def _evaluate(self, predictions_sr, labels_sr):
        # F1 score, smoothed by epsilon (osml/predict.py:129-133)
        precision = self.evaluate_precision(predictions_sr, labels_sr)
        recall = self.evaluate_recall(predictions_sr, labels_sr)
        return (2 * precision * recall + self.epsilon) / (precision + recall + self.epsilon)

    def fuzzy_predict(self, observations_df):
        """Positive-class membership per row as a float Series (osml/predict.py:135-150)."""
        self._validate_observations(observations_df)
        return pd.Series(self._fuzzy_predict(observations_df), dtype=np.float64)

    def _check_pair(self, predictions_sr, labels_sr):
        if predictions_sr.count() != labels_sr.count():
            raise ValueError
        self._validate_labels(predictions_sr)
        self._validate_labels(labels_sr)

    def evaluate_precision(self, predictions_sr, labels_sr):
        """True positives / predicted positives; 1 when nothing was predicted positive (osml/predict.py:152-175)."""
        self._check_pair(predictions_sr, labels_sr)
        predicted_pos = predictions_sr.sum()
        if predicted_pos == 0:
            return 1.
        return float(labels_sr[predictions_sr == 1].sum()) / predicted_pos

    def evaluate_recall(self, predictions_sr, labels_sr):
        """True positives / positives; 1 when there are no positives (osml/predict.py:177-201)."""
        self._check_pair(predictions_sr, labels_sr)
        pos = labels_sr.sum()
        if pos == 0:
            return 1.
        return float(predictions_sr[labels_sr == 1].sum()) / pos

    def evaluate_fall_out(self, predictions_sr, labels_sr):
        """False positives / negatives (osml/predict.py:203-225)."""
        self._check_pair(predictions_sr, labels_sr)
        neg = len(labels_sr[labels_sr == 0])
        return float(predictions_sr[labels_sr != 1].sum()) / neg

    def evaluate_receiver_operating_characteristic(self, fuzzy_predictions_sr, labels_sr, threshold_increment=.01):
        """ROC points (threshold, recall, fall-out) and the trapezoid area under them (osml/predict.py:227-262)."""
        if not 0 < threshold_increment < 1:
            raise ValueError
        roc, auc = [], .0
        prev_recall = prev_fall_out = None
        for threshold in np.arange(0., 1. + threshold_increment, threshold_increment):
            hard = pd.Series(np.where(fuzzy_predictions_sr.values >= threshold, 1, 0), dtype=labels_sr.dtype)
            recall = self.evaluate_recall(hard, labels_sr)
            fall_out = self.evaluate_fall_out(hard, labels_sr)
            roc.append((threshold, recall, fall_out))
            if prev_recall and prev_fall_out:      # truthiness as in the reference: a zero point adds no area
                auc += (recall + prev_recall) / 2 * (prev_fall_out - fall_out)
            prev_recall, prev_fall_out = recall, fall_out
        return roc, auc


GRAM_MAX_ROWS = 8192      # kernel table of the training set: n * n values (512 MB of float64 at the limit)


class _Smo:
    """State of one sequential-minimal-optimisation run (osml/predict.py:1118-1236) with the error cache on the GPU."""

    def __init__(self, model, observations, labels, device):
        self.model = model
        self.x = observations
        self.y = labels                                    # -1. / +1.
        self.alphas = np.zeros(labels.shape)
        self.intercept = .0
        self.errors = -1 * labels                          # decision value 0 for every row minus its label
        self.kernel, self.gamma, self.degree, self.coefficient = model._kernel_args()
        # every training row is a vector of the machine; rows whose multiplier is zero cost nothing on the device
        self.machine = _native.Machine(observations, labels, self.alphas, self.intercept, self.kernel, self.gamma,
                                       self.degree, self.coefficient, device=device)
        self.steps = 0
        # K(x_i, x_j) for all pairs, computed once on the device (osml_svm_gram) in the data's dtype: the three scalar
        # kernel values of every step attempt become table look-ups, and the partner search can test all candidates of
        # a row at once.  Beyond GRAM_MAX_ROWS the table would not fit comfortably; the scalars are then evaluated on
        # the host like the reference does.
        self.gram = self.machine.gram() if len(labels) <= GRAM_MAX_ROWS else None

    def pair_kernel(self, i, j):
        """K(x_i, x_j) for two single rows as a numpy scalar of the data's dtype (osml/predict.py:1120-1121, 1140-1142):
        a look-up in the device-computed table, or -- for training sets too large for the table -- evaluated on
        (1, d)-shaped rows with the reference's own expressions (three scalars per attempt: bookkeeping, not the hot
        path)."""
        if self.gram is not None:
            return self.gram[i, j]
        a = self.x[i].reshape((1, self.x.shape[1]))
        b = self.x[j].reshape((1, self.x.shape[1]))
        if self.kernel == 'lin':
            value = np.dot(a, b.T)
        elif self.kernel == 'poly':
            value = np.power(self.coefficient + np.dot(a, b.T), self.degree)
        else:
            value = np.exp(-self.gamma * np.sqrt(np.square(a - b).sum(axis=1)))
        return value.reshape(-1)[0]

    def refresh_errors(self):
        """osml/predict.py:1101-1105 in one launch: decision value of every row against all rows, minus the labels."""
        self.machine.set_multipliers(self.alphas, self.intercept)
        self.errors[:] = self.machine.decision_self() - self.y

    def step(self, i, j):
        """Joint optimisation of multipliers i and j (osml/predict.py:1118-1164).  True if they moved."""
        c, tol = self.model.c, self.model.tol
        y_i, y_j = self.y[i], self.y[j]
        a_i, a_j = self.alphas[i], self.alphas[j]
        if y_i == y_j:
            hi = a_i + a_j
            lo = hi - c
        else:
            lo = a_j - a_i
            hi = lo + c
        hi = min(c, hi)
        lo = max(.0, lo)
        if hi == lo:
            return False
        k_ii = self.pair_kernel(i, i)
        k_jj = self.pair_kernel(j, j)
        k_ij = self.pair_kernel(i, j)
        eta = float(k_ii + k_jj - 2 * k_ij)
        if eta == 0:
            return False
        e_i, e_j = self.errors[i], self.errors[j]
        a_j_new = max(lo, min(hi, a_j + y_j * (e_i - e_j) / eta))
        delta_j = a_j_new - a_j
        if abs(delta_j) < tol:
            return False
        delta_i = y_i * y_j * -delta_j
        a_i_new = a_i + delta_i
        self.alphas[i] = a_i_new
        self.alphas[j] = a_j_new
        b = float(self.intercept)
        b_i = b - e_i - delta_i * y_i * k_ii - delta_j * y_j * k_ij
        b_j = b - e_j - delta_i * y_i * k_ij - delta_j * y_j * k_jj
        if 0 < a_i_new < c:
            self.intercept = float(b_i)
        elif 0 < a_j_new < c:
            self.intercept = float(b_j)
        else:
            self.intercept = float((b_i + b_j) / 2)
        self.refresh_errors()
        self.steps += 1
        return True

    def first_partner(self, i, candidates):
        """Position in ``candidates`` of the first row j for which ``step(i, j)`` would move the multipliers, or -1.
        A failed attempt has no side effects, so the reference's one-by-one scan (osml/predict.py:1199-1201, 1209-1211)
        equals: evaluate the three rejection tests of ``step`` for every candidate -- here in one pass of float64 vector
        arithmetic, element for element the operations ``step`` performs -- and take the first that passes."""
        c, tol = self.model.c, self.model.tol
        j = np.asarray(candidates, dtype=np.intp)
        y_i, a_i, e_i = self.y[i], self.alphas[i], self.errors[i]
        y_j, a_j, e_j = self.y[j], self.alphas[j], self.errors[j]
        same = y_j == y_i
        total, gap = a_i + a_j, a_j - a_i
        hi = np.minimum(c, np.where(same, total, gap + c))
        lo = np.maximum(.0, np.where(same, total - c, gap))
        eta = (self.gram[i, i] + self.gram[j, j] - 2 * self.gram[i, j]).astype(np.float64)
        with np.errstate(divide='ignore', invalid='ignore'):
            a_j_new = np.maximum(lo, np.minimum(hi, a_j + y_j * (e_i - e_j) / eta))
            moves = (hi != lo) & (eta != 0) & ~(np.abs(a_j_new - a_j) < tol)
        hits = np.flatnonzero(moves)
        return int(hits[0]) if len(hits) else -1

    def try_partners(self, i, candidates):
        """``any(step(i, j) for j in candidates)`` with the reference's short-circuit order."""
        if self.gram is None:
            return any(self.step(i, j) for j in candidates)
        while candidates:
            pos = self.first_partner(i, candidates)
            if pos < 0:
                return False
            if self.step(i, candidates[pos]):           # the scalar step re-checks and performs the update
                return True
            candidates = candidates[pos + 1:]            # cannot happen with finite data; keeps the scan faithful
        return False

    def examine(self, i):
        """Picks partners for multiplier i if it violates the KKT conditions (osml/predict.py:1166-1204)."""
        c, tol = self.model.c, self.model.tol
        r_i = self.y[i] * self.errors[i]
        if not ((r_i < -tol and self.alphas[i] < c) or (r_i > tol and self.alphas[i] > 0)):
            return False
        others = np.arange(len(self.y)) != i
        unbounded = np.flatnonzero(others & (self.alphas > 0) & (self.alphas < c)).tolist()
        if unbounded:
            order = np.argsort(self.errors[unbounded])
            # As in the reference (osml/predict.py:1188-1192) the winner of the |E_i - E_j| heuristic is a POSITION in
            # the unbounded list and is used as a row number without mapping it back.
            first = order[0] if self.errors[i] > 0 else order[-1]
            if self.step(i, first):
                return True
            start = np.random.randint(0, len(unbounded))
            if self.try_partners(i, [j for j in unbounded[start:] + unbounded[:start] if j != first]):
                return True
        at_zero = np.flatnonzero(others & (self.alphas == .0)).tolist()
        if at_zero:
            start = np.random.randint(0, len(at_zero))
            if self.try_partners(i, at_zero[start:] + at_zero[:start]):
                return True
        return False

    def run(self):
        """Alternates one pass over all rows with passes over the unbounded ones until nothing moves
        (osml/predict.py:1206-1231)."""
        c = self.model.c
        whole_set, moved = True, 0
        while whole_set or moved > 0:
            moved = 0
            rows = range(len(self.y)) if whole_set else np.flatnonzero((self.alphas > 0) & (self.alphas < c)).tolist()
            for i in rows:
                if self.examine(i):
                    moved += 1
            whole_set = not whole_set and moved == 0

    def close(self):
        self.machine.close()


class SupportVectorMachine(BinaryClassificationModel):
    """Support vector machine trained by sequential minimal optimisation (osml/predict.py:1050-1250).

    GPU-only knob (an attribute, so the reference's constructor signature is kept): ``device`` (CUDA ordinal; default
    ``$LOCAL_RANK`` or 0).  ``kernel`` keeps the constructor's kernel name (the reference keeps the bound function).
    """

    def __init__(self, tol=1e-3, c=1., kernel='rbf', gamma=.5, degree=2, poly_coefficient=1.):
        super(SupportVectorMachine, self).__init__()
        if tol < 0 or c < 0:
            raise ValueError
        if kernel == 'poly':
            if degree is None or degree < 1 or poly_coefficient is None:
                raise ValueError
        elif kernel == 'rbf':
            if gamma is None or gamma <= 0:
                raise ValueError
        elif kernel != 'lin':
            raise ValueError
        self.kernel = kernel
        self.tol = tol
        self.c = c
        self.degree = degree
        self.poly_coefficient = poly_coefficient
        self.gamma = gamma
        self._support_vectors = None
        self._support_vector_labels = None
        self._support_vector_alphas = None
        self._intercept = None
        self.device = None
        self._machine = None
        self.smo_steps = 0

    # -- device state ---------------------------------------------------------------------------------------------
    def _device(self):
        return self.device if self.device is not None else _default_device()

    def _kernel_args(self):
        # parameters of the kernels that are not in use may be None in the reference; the C ABI wants numbers
        return (self.kernel, self.gamma if self.gamma is not None else .5, self.degree if self.degree is not None else 2,
                self.poly_coefficient if self.poly_coefficient is not None else 1.)

    def _ensure_machine(self):
        """(Re)creates the device copy of the fitted state; lazy so that unpickled / deep-copied models work."""
        if self._machine is None:
            if self._support_vectors is None:
                raise ValueError
            if len(self._support_vectors) == 0:
                return None
            self._machine = _native.Machine(self._support_vectors, self._support_vector_labels,
                                            self._support_vector_alphas, self._intercept, *self._kernel_args(),
                                            device=self._device())
        return self._machine

    def _drop_machine(self):
        if self._machine is not None:
            self._machine.close()
            self._machine = None

    def __getstate__(self):
        state = self.__dict__.copy()
        state['_machine'] = None        # device handles never travel (pickle, copy.deepcopy, multiprocessing)
        return state

    def __setstate__(self, state):
        self.__dict__.update(state)
        self._machine = None

    # -- hooks ------------------------------------------------------------------------------------------------------
    def _raw_predict(self, new_observations, observations, labels, alphas, intercept):
        """Decision values intercept + sum_i K(observations_i, x) * alphas_i * labels_i per row x of
        ``new_observations`` (osml/predict.py:1095-1099), on the GPU.  Called with the model's own fitted state it uses
        the cached device copy; any other matrices get a temporary one."""
        if len(observations) == 0:      # no vectors: every sum is empty
            return np.full(len(new_observations), float(intercept))
        own = (observations is self._support_vectors and labels is self._support_vector_labels and
               alphas is self._support_vector_alphas and intercept is self._intercept)
        if own:
            return self._ensure_machine().decision(new_observations)
        machine = _native.Machine(observations, labels, alphas, intercept, *self._kernel_args(), device=self._device())
        try:
            return machine.decision(new_observations)
        finally:
            machine.close()

    def _fitted_raw_predict(self, new_observations):
        return self._raw_predict(new_observations, self._support_vectors, self._support_vector_labels,
                                 self._support_vector_alphas, self._intercept)

    def _fit(self, observations_df, labels_sr):
        if len(observations_df.index) < 2:          # the equality constraint needs two rows (osml/predict.py:1208-1209)
            raise ValueError
        self._drop_machine()
        observations = observations_df.values
        labels = labels_sr.values * 2. - 1.         # {0, 1} -> {-1, +1}
        smo = _Smo(self, observations, labels, self._device())
        try:
            smo.run()
        finally:
            smo.close()
        self.smo_steps = smo.steps
        # keep the rows with a positive multiplier (osml/predict.py:1107-1116), intercept rescaled as the reference does
        keep = smo.alphas > .0
        self._support_vectors = observations[keep]
        self._support_vector_labels = labels[keep]
        self._support_vector_alphas = smo.alphas[keep]
        n_support = len(self._support_vectors)
        self._intercept = float(smo.intercept) * len(observations) / n_support if n_support > 0 else 0.

    def _test(self, observations_df, labels_sr):
        # hinge loss (osml/predict.py:1238-1243)
        raw = self._fitted_raw_predict(observations_df.values)
        return np.sum(np.maximum(1. - (labels_sr.values * 2. - 1.) * raw, .0))

    def _fuzzy_predict(self, observations_df):
        # (-inf, inf) -> [0, 1] (osml/predict.py:1245-1250)
        raw = self._fitted_raw_predict(observations_df.values)
        return raw / (np.abs(raw).max() * 2 + self.epsilon) + .5

    def decision_function(self, observations_df):
        """Not in the reference's public API: the raw decision values ``_raw_predict`` returns, for the parity tests."""
        self._validate_observations(observations_df)
        return self._fitted_raw_predict(observations_df.values)
