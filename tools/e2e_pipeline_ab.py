import sys, time
import numpy as np, torch
sys.path.insert(0, '.')
from osml_b200 import _native
rng = np.random.default_rng(0)
n, d, k = 1000000, 128, 10
train = np.asfortranarray(rng.standard_normal((n, d), dtype=np.float32))
labels = rng.integers(0, 10, n)
ix = _native.Index(train, labels, k, 1, 1e-8)
for q in (8192, 12500, 25000, 50000):
    test = torch.from_numpy(rng.standard_normal((q, d), dtype=np.float32)).pin_memory().numpy()
    for rnd in range(2):
        for pipe in (1, 0):
            ix.set_option('pipeline', pipe)
            for _ in range(3):
                ix.predict(test, _native.CLASSIFICATION)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(10):
                ix.predict(test, _native.CLASSIFICATION)
            print('q=%d pipeline=%d: %.3f ms' % (q, pipe, (time.perf_counter() - t0) * 100))
