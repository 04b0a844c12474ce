"""C4-shaped float64 problem (pandas' default dtype): tensor filter path vs the direct kernel, through osml_knn_predict
with pinned host buffers.   python tools/time_f64.py [n] [d] [q] [k]"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from osml_b200 import _native  # noqa: E402


def main():
    n, d, q, k = (int(x) for x in (sys.argv[1:5] + ['1000000', '128', '100000', '10'][len(sys.argv) - 1:]))
    rng = np.random.default_rng(4)
    train = np.asfortranarray(rng.standard_normal((n, d)))
    queries = torch.from_numpy(rng.standard_normal((q, d))).pin_memory().numpy()
    labels = rng.integers(0, 10, n)
    out = {'what': 'float64 %d x %d train, %d queries, k=%d, osml_knn_predict with pinned host buffers' % (n, d, q, k)}
    ix = _native.Index(train, labels, k, 1, 1e-8)
    ix.set_option('path', _native.PATH_TENSOR)
    pred = ix.predict(queries, _native.CLASSIFICATION)
    pred = ix.predict(queries, _native.CLASSIFICATION)
    t0 = time.perf_counter()
    for _ in range(5):
        pred = ix.predict(queries, _native.CLASSIFICATION)
    dt = (time.perf_counter() - t0) / 5
    st = ix.stats()
    out['tensor_path'] = {'queries_per_s': q / dt, 'ms_per_call': dt * 1e3, 'k1_ms': st['last_filter_ms'],
                          'retry_queries_per_call': st['retry_queries'] / 7, 'fallback_queries': st['fallback_queries']}
    ix.set_option('path', _native.PATH_DIRECT)
    sample = min(q, 2048)
    ix.predict(queries[:sample], _native.CLASSIFICATION)
    t0 = time.perf_counter()
    d_pred = ix.predict(queries[:sample], _native.CLASSIFICATION)
    dt = time.perf_counter() - t0
    out['direct_kernel'] = {'queries_per_s': sample / dt, 'sample': sample}
    out['same_predictions_on_the_sample'] = bool(np.array_equal(d_pred, pred[:sample]))
    ix.close()
    print(json.dumps(out))


if __name__ == '__main__':
    main()
