"""Ad-hoc device timing of the neighbour-search paths (not the bench):
    python tools/time_paths.py n d q k path [reps] [option=value ...] [dtype=int8 hi=11]
`dtype` / `hi`: integer data drawn uniformly from 0..hi instead of float32 gaussians (hi = 11 is the widest range int8
arithmetic cannot wrap in: the integer tensor-filter path)."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, '.')
from osml_b200 import _native  # noqa: E402


def main():
    n, d, q, k = (int(x) for x in sys.argv[1:5])
    path = sys.argv[5] if len(sys.argv) > 5 else 'direct'
    reps = int(sys.argv[6]) if len(sys.argv) > 6 else 3
    opts = dict(a.split('=') for a in sys.argv[7:])
    rng = np.random.default_rng(0)
    dtype = np.dtype(opts.pop('dtype', 'float32'))
    hi = int(opts.pop('hi', 11))
    if dtype.kind in 'iu':
        train = np.asfortranarray(rng.integers(0, hi + 1, (n, d)).astype(dtype))
        test = rng.integers(0, hi + 1, (q, d)).astype(dtype)
    else:
        train = np.asfortranarray(rng.standard_normal((n, d)).astype(dtype))
        test = rng.standard_normal((q, d)).astype(dtype)
    labels = rng.integers(0, 10, n)
    t0 = time.time()
    ix = _native.Index(train, labels, k, 1, 1e-8)
    ix.set_option('path', _native.PATHS[path])
    for key, value in opts.items():
        ix.set_option(key, float(value))
    print('create %.3fs' % (time.time() - t0))
    dq = torch.from_numpy(test).cuda()
    out = torch.empty(q, dtype=torch.float64, device='cuda')
    idx = torch.empty((q, k + 1), dtype=torch.int64, device='cuda')
    dist = torch.empty((q, k + 1), dtype=torch.float32 if dtype == np.float32 else torch.float64, device='cuda')
    stream = torch.cuda.current_stream().cuda_stream
    for r in range(reps + 1):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ix.predict_device(dq.data_ptr(), q, _native.CLASSIFICATION, out.data_ptr(), idx.data_ptr(), dist.data_ptr(),
                          stream)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print('%s n=%d d=%d q=%d k=%d: %.3f ms  %.1f q/s  %.2f Tpair-elem/s  stats %s' % (
            path, n, d, q, k, ms, q / ms * 1e3, n * d * q / ms / 1e9, ix.stats()))


if __name__ == '__main__':
    main()
