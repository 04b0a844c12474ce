"""The numpy restatement of Centering / Standardization (oracle/transform_oracle.py) against fixtures produced by the
REAL reference classes (osml/transform.py:53-84)."""
import numpy as np

from oracle import transform_oracle as to


def _same(a, b):
    return a.dtype == b.dtype and a.shape == b.shape and np.array_equal(a, b, equal_nan=True)


def test_fit_statistics_match_the_reference(transform_golden):
    g = transform_golden
    mean, sd = to.standardization_fit(g['fit'])
    assert _same(mean, g['mean']) and _same(sd, g['sd'])
    assert _same(to.centering_fit(g['fit']), g['centering_mean'])


def test_transforms_match_the_reference_bit_for_bit(transform_golden):
    g = transform_golden
    for frame, tag in ((g['fit'], 'fit'), (g['query'], 'query')):
        assert _same(to.centering_transform(frame, g['centering_mean']), g['centered_' + tag])
        assert _same(to.standardization_transform(frame, g['mean'], g['sd'], g['epsilon']), g['standardized_' + tag])


def test_output_keeps_the_memory_order_of_the_input(transform_golden):
    g = transform_golden
    out = to.standardization_transform(g['fit'], g['mean'], g['sd'], g['epsilon'])
    assert out.flags['F_CONTIGUOUS'] == g['standardized_fit'].flags['F_CONTIGUOUS']
    assert bool(g['out_is_f_order']) == bool(g['fit_is_f_order'])


def test_weak_scalar_epsilon_is_rounded_in_the_data_dtype():
    sd = np.array([2.5, 1e-9, 0.0], dtype=np.float32)
    _, denom = to.shift_and_denominator(np.zeros(3, np.float32), sd, 1e-8)
    assert denom.dtype == np.float32
    assert denom[0] == np.float32(2.5)                       # 1e-8 vanishes next to 2.5 in float32
    assert denom[2] == np.float32(1e-8)
