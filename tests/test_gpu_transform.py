"""GPU parity of Centering / Standardization (osml_knn_column_transform) and of the transform fused into predict
(osml_knn_set_query_transform) against the oracle and the fixtures produced by the real reference."""
import numpy as np
import pandas as pd
import pytest

import osml_b200 as ob
from oracle import knn_oracle as ko
from oracle import transform_oracle as to
from osml_b200 import _native

pytestmark = pytest.mark.gpu


def _same(a, b):
    return a.dtype == b.dtype and a.shape == b.shape and np.array_equal(a, b, equal_nan=True)


def test_golden_through_the_model_api(transform_golden):
    g = transform_golden
    fit_df, query_df = pd.DataFrame(g['fit']), pd.DataFrame(g['query'])
    cen, std = ob.Centering(), ob.Standardization()
    got = cen.fit_transform(fit_df, None)
    assert _same(got.values, g['centered_fit'])
    assert got.values.flags['F_CONTIGUOUS'] == g['centered_fit'].flags['F_CONTIGUOUS']
    assert _same(cen.transform(query_df).values, g['centered_query'])
    assert _same(std.fit_transform(fit_df, None).values, g['standardized_fit'])
    assert _same(std.transform(query_df).values, g['standardized_query'])
    assert _same(std._mean, g['mean']) and _same(std._sd, g['sd'])


@pytest.mark.parametrize('dtype', ['float32', 'float64', 'int8', 'int16', 'int32', 'int64', 'uint8', 'uint16', 'uint32',
                                   'uint64'])
@pytest.mark.parametrize('order', ['F', 'C'])
def test_every_dtype_both_layouts_through_the_c_abi(dtype, order):
    rng = np.random.default_rng(hash((dtype, order)) % 2**32)
    n, d = 1234, 7
    if dtype.startswith('float'):
        x = (rng.standard_normal((n, d)) * 50).astype(dtype)
    else:
        info = np.iinfo(dtype)
        x = rng.integers(max(info.min, -10**6), min(info.max, 10**6), (n, d)).astype(dtype)
    x = np.asfortranarray(x) if order == 'F' else np.ascontiguousarray(x)
    mean, sd = to.standardization_fit(x)
    shift, denom = to.shift_and_denominator(mean, sd, 1e-8)
    want = to.standardization_transform(x, mean, sd, 1e-8)
    got = _native.column_transform(x, shift, denom)
    assert _same(got, want) and got.flags['F_CONTIGUOUS'] == want.flags['F_CONTIGUOUS']
    assert _same(_native.column_transform(x, shift, None), to.centering_transform(x, mean))


@pytest.mark.parametrize('dtype,n,d', [('float32', 60001, 20), ('float64', 300007, 3), ('int8', 900000, 3),
                                         ('float32', 9, 60000)])
@pytest.mark.parametrize('order', ['F', 'C'])
def test_large_matrices_through_the_staging_ring(dtype, n, d, order):
    """From 2 MB up a pageable matrix goes to the device -- and the result comes back -- through the ring of pinned slots
    (host_stager.cuh: chunks of a megabyte, the last one ragged; int8 in, float64 out: only the way back is staged)."""
    rng = np.random.default_rng(n + d)
    x = (rng.standard_normal((n, d)) * 9).astype(dtype)
    x = np.asfortranarray(x) if order == 'F' else np.ascontiguousarray(x)
    mean, sd = to.standardization_fit(x)
    shift, denom = to.shift_and_denominator(mean, sd, 1e-8)
    want = to.standardization_transform(x, mean, sd, 1e-8)
    for _ in range(2):                  # the ring is reused
        got = _native.column_transform(x, shift, denom)
        assert _same(got, want) and got.flags['F_CONTIGUOUS'] == want.flags['F_CONTIGUOUS']


def test_degenerate_columns_and_nonfinite_values():
    x = np.asfortranarray(np.random.default_rng(3).standard_normal((300, 5)).astype(np.float32))
    x[:, 2] = 7.0                       # sd == 0: the reference divides by epsilon
    x[5, 0] = np.nan
    x[6, 1] = np.inf
    mean, sd = to.standardization_fit(x)
    with np.errstate(all='ignore'):
        want = to.standardization_transform(x, mean, sd, 1e-8)
    assert _same(_native.column_transform(x, mean, sd + 1e-8), want)


@pytest.mark.parametrize('dtype', ['float32', 'float64'])
def test_transform_fused_into_predict_equals_the_two_step_pipeline(dtype):
    """README pipeline of the reference: standardise, fit kNN on the standardised frame, standardise the queries, predict.
    With the transform fused, predict takes the raw queries and must return the same labels / neighbours."""
    rng = np.random.default_rng(9)
    n, d, q, k = 30000, 24, 2000, 7
    scale = rng.uniform(0.1, 40, d)
    train = pd.DataFrame((rng.standard_normal((n, d)) * scale + rng.uniform(-20, 20, d)).astype(dtype))
    queries = pd.DataFrame((rng.standard_normal((q, d)) * scale + rng.uniform(-20, 20, d)).astype(dtype))
    labels = pd.Series(rng.integers(0, 5, n), name='y')
    std = ob.Standardization()
    train_std = std.fit_transform(train, labels)
    model = ob.KNearestNeighborsClassification(k=k)
    model.fit(train_std, labels)
    two_step = model.predict(std.transform(queries))
    idx2, dist2 = model.kneighbors(std.transform(queries))
    model.set_query_transform(std)
    fused = model.predict(queries)
    idx1, dist1 = model.kneighbors(queries)
    assert np.array_equal(fused.values, two_step.values)
    assert np.array_equal(idx1, idx2) and np.array_equal(dist1, dist2)
    # and both equal the oracle on the reference's own arithmetic
    want = ko.predict_classification(to.standardization_transform(train.values, std._mean, std._sd, std.epsilon),
                                     labels.values, to.standardization_transform(queries.values[:50], std._mean, std._sd,
                                                                                 std.epsilon), k=k)
    assert np.array_equal(fused.values[:50].astype(np.float64), want)
    # the fused transform survives pickling (the device handle does not travel)
    import pickle
    clone = pickle.loads(pickle.dumps(model))
    assert np.array_equal(clone.predict(queries).values, fused.values)
    model.set_query_transform(None)
    assert np.array_equal(model.predict(std.transform(queries)).values, two_step.values)


def test_validation_matches_the_reference():
    std = ob.Standardization()
    df = pd.DataFrame(np.random.default_rng(1).standard_normal((10, 3)))
    with pytest.raises(ValueError):
        std.fit(pd.DataFrame(), None)
    std.fit(df, None)
    with pytest.raises(ValueError):                      # dtype vector differs from the fit frame's
        std.transform(df.astype(np.float32))
    with pytest.raises(NotImplementedError):             # the fused transform needs a float index
        ix = _native.Index(np.zeros((20, 3), np.int8), np.zeros(20, np.int64), 3, 1, 1e-8)
        ix.set_query_transform(np.zeros(3))
