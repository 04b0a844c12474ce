"""CPU oracle for the OSML brute-force weighted kNN predict path.

TEST INFRASTRUCTURE ONLY.  Nothing in ``osml_b200/`` may import this module; it is used by
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``
as the checker (and as the timed CPU arm), never as the product path.

It is a numpy restatement of the reference's algorithm, function by function:

* ``distances``                 <- /root/reference/osml/predict.py:740-742  (``_distances``)
* ``select_reference``          <- predict.py:748-752  (argpartition + pandas descending sort, implementation-defined ties)
* ``select_deterministic``      <- same selection with the documented (distance, index) tie rule the GPU path uses
* ``weighted_labels``           <- predict.py:744-746, 753-765  (``_weight`` + the dict accumulation loop)
* ``predict_regression``        <- predict.py:785-794
* ``predict_classification``    <- predict.py:805-811
* ``distances_sequential`` / ``distances_pairwise`` <- explicit restatements of the summation order numpy uses for
  ``sum(axis=1)`` on an F-ordered / C-ordered matrix (numpy ``pairwise_sum``, blocks of <=128, 8 accumulators).
  These are what the CUDA kernels implement; tests pin them bit-for-bit against ``distances``.

Parity pin: ``tests/golden/*.npz`` hold outputs of the *real* reference (run in the authoring container by
``oracle/make_golden.py`` through ``oracle/ref_shim.py``) on its own data sets; ``tests/test_oracle.py`` checks every
function here against them.  The third-party arithmetic the reference leans on (numpy 2.3.5, pandas 3.0.2; unpinned in
the reference's requirements.txt:1-8) is recorded in each golden file.
"""
import math

import numpy as np
import pandas as pd

EPSILON = 1e-8  # predict.py:17  PredictiveModel.__init__(epsilon=1e-8)


# ----------------------------------------------------------------------------------------------------------------------
# distances
# ----------------------------------------------------------------------------------------------------------------------
def distances(observations, observation):
    """predict.py:742 verbatim in meaning: dtype-preserving subtract, square, row-sum, sqrt."""
    return np.sqrt(np.square(observations - observation).sum(axis=1))


def sum_order_of(observations):
    """'sequential' if numpy reduces axis 1 of (observations - q) feature by feature (axis 0 has the smaller stride,
    e.g. the F-ordered ``DataFrame.values`` of pandas 3), else 'pairwise' (C order: numpy's pairwise_sum)."""
    if observations.ndim != 2:
        raise ValueError
    n, d = observations.shape
    if d == 1:
        return 'sequential'
    if n == 1:          # a single row is contiguous along the features whatever its flags say: numpy's pairwise_sum
        return 'pairwise'
    s0, s1 = (abs(s) for s in observations.strides)
    return 'sequential' if s0 <= s1 else 'pairwise'


def _elementwise_squares(observations, observation):
    diff = observations - observation  # wraps for narrow ints exactly like the reference
    return np.square(diff)


def distances_sequential(observations, observation):
    """sum_j in the order j = 0, 1, ..., D-1 with one rounding per add (what numpy does for an F-ordered matrix)."""
    sq = _elementwise_squares(np.asarray(observations), observation)
    acc_dtype = _sum_dtype(sq.dtype)
    acc = sq[:, 0].astype(acc_dtype)
    for j in range(1, sq.shape[1]):
        acc = acc + sq[:, j].astype(acc_dtype)
    return np.sqrt(acc)


def _sum_dtype(dtype):
    # np.add.reduce widens ints narrower than the platform int to int64 / uint64; floats stay.
    if np.issubdtype(dtype, np.signedinteger):
        return np.dtype(np.int64)
    if np.issubdtype(dtype, np.unsignedinteger):
        return np.dtype(np.uint64)
    return np.dtype(dtype)


def _pairwise_block(sq, lo, n):
    """numpy's pairwise_sum over columns lo..lo+n of sq (rows vectorised)."""
    if n < 8:
        res = np.zeros(sq.shape[0], dtype=sq.dtype)
        for i in range(n):
            res = res + sq[:, lo + i]
        return res
    if n <= 128:
        r = [sq[:, lo + i].copy() for i in range(8)]
        i = 8
        while i < n - (n % 8):
            for j in range(8):
                r[j] = r[j] + sq[:, lo + i + j]
            i += 8
        res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]))
        while i < n:
            res = res + sq[:, lo + i]
            i += 1
        return res
    n2 = n // 2
    n2 -= n2 % 8
    return _pairwise_block(sq, lo, n2) + _pairwise_block(sq, lo + n2, n - n2)


def distances_pairwise(observations, observation):
    """sum_j in numpy's pairwise order (what numpy does for a C-ordered float matrix)."""
    sq = _elementwise_squares(np.asarray(observations), observation)
    if not np.issubdtype(sq.dtype, np.floating):
        return np.sqrt(sq.astype(_sum_dtype(sq.dtype)).sum(axis=1))
    return np.sqrt(_pairwise_block(sq, 0, sq.shape[1]))


# ----------------------------------------------------------------------------------------------------------------------
# selection
# ----------------------------------------------------------------------------------------------------------------------
def select_reference(dist, k):
    """predict.py:750-752: returns (indices, distances) in the reference's own descending order."""
    k_plus_1_nn_indices = np.argpartition(dist, k + 1)[:k + 1]
    k_plus_1_nn = pd.Series(dist[k_plus_1_nn_indices], index=k_plus_1_nn_indices)
    k_plus_1_nn = k_plus_1_nn.sort_values(ascending=False)
    return k_plus_1_nn.index.values.astype(np.int64), k_plus_1_nn.values


def select_deterministic(dist, k):
    """The GPU path's documented rule: the k+1 smallest by (distance, index) ascending; returned descending, i.e. the
    exact reverse (farthest first; among equal distances the larger index first).  NaN sorts after +inf."""
    if len(dist) <= k + 1:
        # np.argpartition(dist, k + 1) raises ValueError for kth out of bounds
        raise ValueError
    order = np.lexsort((np.arange(len(dist)), dist))[:k + 1]
    order = order[::-1]
    return order.astype(np.int64), dist[order]


# ----------------------------------------------------------------------------------------------------------------------
# weights and aggregation
# ----------------------------------------------------------------------------------------------------------------------
def weighted_labels(indices_desc, dists_desc, labels, rbf_gamma=1, epsilon=EPSILON):
    """predict.py:753-765: first (farthest) entry normalises, the remaining k vote; dict keeps insertion order.
    ``Series.iteritems()`` (predict.py:755) yields *Python* scalars (pandas boxes numeric values on iteration), so the
    ratio, the square (libm ``pow``) and the exponent are all fp64 whatever the data dtype; ``tolist()`` boxes the
    same way."""
    out = {}
    distance_to_normalize_by = None
    for i, value in zip(np.asarray(indices_desc).tolist(), np.asarray(dists_desc).tolist()):
        if distance_to_normalize_by is None:
            distance_to_normalize_by = value
            continue
        distance = value / (distance_to_normalize_by + epsilon)
        weight = math.exp(-rbf_gamma * (distance ** 2))  # predict.py:746
        label = labels[i]
        if label in out:
            out[label] += weight
        else:
            out[label] = weight
    return out


def aggregate_regression(weights):
    """predict.py:788-793."""
    weighted_labels_sum = 0
    weight_sum = 0
    for key, value in weights.items():
        weighted_labels_sum += key * value
        weight_sum += value
    return weighted_labels_sum / weight_sum


def aggregate_classification(weights):
    """predict.py:809-810: stable sort, reverse=True keeps insertion order among equal weights."""
    return sorted(weights.items(), key=lambda kv: kv[1], reverse=True)[0][0]


def _predict(observations, labels, queries, k, rbf_gamma, epsilon, aggregate, select, dist_fn, return_neighbours):
    q_count = queries.shape[0]
    predictions = np.zeros((q_count,), )
    nn_idx = np.zeros((q_count, k + 1), dtype=np.int64) if return_neighbours else None
    nn_dist = None
    for qi in range(q_count):
        d = dist_fn(observations, np.ascontiguousarray(queries[qi]))
        idx_desc, d_desc = select(d, k)
        if return_neighbours:
            if nn_dist is None:
                nn_dist = np.zeros((q_count, k + 1), dtype=d.dtype)
            nn_idx[qi] = idx_desc[::-1]
            nn_dist[qi] = d_desc[::-1]
        with np.errstate(all='ignore'):
            predictions[qi] = aggregate(weighted_labels(idx_desc, d_desc, labels, rbf_gamma, epsilon))
    if return_neighbours:
        return predictions, nn_idx, nn_dist
    return predictions


def predict_regression(observations, labels, queries, k=7, rbf_gamma=1, epsilon=EPSILON, tie_rule='deterministic',
                       return_neighbours=False):
    """predict.py:785-794.  ``tie_rule='reference'`` uses argpartition + pandas sort; 'deterministic' the GPU rule."""
    select = select_reference if tie_rule == 'reference' else select_deterministic
    return _predict(observations, labels, queries, k, rbf_gamma, epsilon, aggregate_regression, select, distances,
                    return_neighbours)


def predict_classification(observations, labels, queries, k=7, rbf_gamma=1, epsilon=EPSILON,
                           tie_rule='deterministic', return_neighbours=False):
    """predict.py:805-811 (returns the fp64 array ``_predict`` returns; the caller casts to the label dtype)."""
    select = select_reference if tie_rule == 'reference' else select_deterministic
    return _predict(observations, labels, queries, k, rbf_gamma, epsilon, aggregate_classification, select, distances,
                    return_neighbours)


def predict_from_neighbours(nn_idx_asc, nn_dist_asc, labels, task, rbf_gamma=1, epsilon=EPSILON):
    """Aggregation half only (predict.py:753-765 + 788-793 / 809-810) on given neighbours (ascending order, as the
    C-ABI reports them).  Lets a test check K4 on the GPU's own neighbour order for tie-flagged queries."""
    agg = aggregate_regression if task == 'regression' else aggregate_classification
    out = np.zeros((nn_idx_asc.shape[0],), )
    for qi in range(nn_idx_asc.shape[0]):
        with np.errstate(all='ignore'):
            out[qi] = agg(weighted_labels(nn_idx_asc[qi][::-1], nn_dist_asc[qi][::-1], labels, rbf_gamma, epsilon))
    return out


# ----------------------------------------------------------------------------------------------------------------------
# tie classification (SURVEY.md section 8 "tie census")
# ----------------------------------------------------------------------------------------------------------------------
def tie_flags(dist, k):
    """Returns (any_tie, boundary_tie): ``any_tie`` if two of the k+2 smallest distances are equal (order inside the
    neighbour list, and hence fp64 summation order / weight-tie winners, is implementation-defined in the reference);
    ``boundary_tie`` if the (k+1)-th and (k+2)-th smallest are equal (set membership itself is ambiguous)."""
    m = min(len(dist), k + 2)
    smallest = np.sort(dist)[:m]
    any_tie = bool(np.any(smallest[1:] == smallest[:-1]))
    boundary_tie = bool(m == k + 2 and smallest[k] == smallest[k + 1])
    return any_tie, boundary_tie


def is_valid_neighbour_set(dist, k, idx_asc):
    """True if idx_asc (ascending (distance, index)) is a valid k+1 nearest set for ``dist`` under *some* tie
    resolution: contains every point strictly closer than the (k+1)-th distance, nothing farther than it."""
    idx_asc = np.asarray(idx_asc)
    if len(set(idx_asc.tolist())) != k + 1:
        return False
    kth = np.sort(dist)[k]
    chosen = dist[idx_asc]
    if np.any(chosen > kth):
        return False
    return int(np.sum(dist < kth)) == int(np.sum(chosen < kth))
