"""Ad-hoc: tensor path with option overrides vs the direct path on the same inputs (python tools/check_tensor.py n d q k key=value...)"""
import sys

import numpy as np

sys.path.insert(0, '.')
from osml_b200 import _native  # noqa: E402

n, d, q, k = (int(x) for x in sys.argv[1:5])
opts = dict(a.split('=') for a in sys.argv[5:])
rng = np.random.default_rng(1)
train = np.asfortranarray(rng.standard_normal((n, d), dtype=np.float32))
labels = rng.integers(0, 10, n)
test = rng.standard_normal((q, d), dtype=np.float32)
ref = _native.Index(train, labels, k, 1, 1e-8)
ref.set_option('path', _native.PATH_DIRECT)
rp, ri, rd = ref.predict(test, _native.CLASSIFICATION, return_neighbours=True)
ix = _native.Index(train, labels, k, 1, 1e-8)
ix.set_option('path', _native.PATH_TENSOR)
for key, value in opts.items():
    ix.set_option(key, float(value))
p, i, dd = ix.predict(test, _native.CLASSIFICATION, return_neighbours=True)
st = ix.stats()
print('match idx %s dist %s pred %s  fallback %d / %d  filter_ms %.3f' % (
    np.array_equal(i, ri), np.array_equal(dd, rd), np.array_equal(p, rp), st['fallback_queries'], q,
    st['last_filter_ms']))
