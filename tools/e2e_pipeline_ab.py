"""End-to-end osml_knn_predict (pinned host queries) with the query matrix uploaded whole, in two or in four blocks:
    python tools/e2e_pipeline_ab.py n d q [q ...]"""
import sys, time
import numpy as np, torch
sys.path.insert(0, '.')
from osml_b200 import _native
rng = np.random.default_rng(0)
n, d = int(sys.argv[1]), int(sys.argv[2])
k = 10
train = np.asfortranarray(rng.standard_normal((n, d), dtype=np.float32))
labels = rng.integers(0, 10, n)
ix = _native.Index(train, labels, k, 1, 1e-8)
for q in (int(a) for a in sys.argv[3:]):
    test = torch.from_numpy(rng.standard_normal((q, d), dtype=np.float32)).pin_memory().numpy()
    for rnd in range(2):
        for pipe, blocks in ((0, 0), (1, 2), (1, 3), (1, 4)):
            ix.set_option('pipeline', pipe)
            ix.set_option('pipeline_blocks', blocks)
            for _ in range(3):
                ix.predict(test, _native.CLASSIFICATION)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(10):
                ix.predict(test, _native.CLASSIFICATION)
            print('q=%d pipeline=%d blocks=%d: %.3f ms  (K1 launches %d, K1 %.3f ms)' % (
                q, pipe, blocks, (time.perf_counter() - t0) * 100, ix.stats()['last_filter_launches'], ix.stats()['last_filter_ms']))
