"""Repeated tensor-vs-direct comparisons on wide-row shapes (the two-issuer K-block ring, odd and even numbers of K
blocks, K' = 16 and 32, split sweeps with live gates): python tools/stress_wide.py [rounds]"""
import sys
import numpy as np
sys.path.insert(0, '.')
from osml_b200 import _native

rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 3
shapes = [(60000, 784, 3000, 10), (30000, 200, 5000, 10), (20000, 300, 4000, 5), (25000, 1000, 1500, 10),
          (40000, 400, 3000, 32), (20000, 257, 2500, 20), (70000, 96, 20000, 32), (100000, 128, 30000, 10)]
bad = 0
for r in range(rounds):
    for n, d, q, k in shapes:
        rng = np.random.default_rng(1000 * r + n + d)
        train = np.asfortranarray(rng.standard_normal((n, d), dtype=np.float32))
        test = rng.standard_normal((q, d), dtype=np.float32)
        labels = rng.integers(0, 7, n)
        out = []
        for path in ('tensor', 'direct'):
            ix = _native.Index(train, labels, k, 1, 1e-8)
            ix.set_option('path', _native.PATHS[path])
            out.append(ix.predict(test, _native.CLASSIFICATION, return_neighbours=True))
            st = ix.stats()
            ix.close()
        same = all(np.array_equal(a, b) for a, b in zip(*out))
        bad += 0 if same else 1
        print('round %d  n=%d d=%d q=%d k=%d  %s' % (r, n, d, q, k, 'ok' if same else 'MISMATCH'), flush=True)
print('mismatches:', bad)
sys.exit(1 if bad else 0)
