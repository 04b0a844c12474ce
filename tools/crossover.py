"""Where does the tensor filter start to beat the direct kernel for ONE predict call on a fitted index?  Device-resident
queries, steady state (K0 done), per-call time of both paths over a grid of (n, d, q).
    python tools/crossover.py"""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, '.')
from osml_b200 import _native  # noqa: E402


def time_path(ix, path, dq, q, out, reps=3):
    ix.set_option('path', path)
    stream = torch.cuda.current_stream().cuda_stream
    best = 1e9
    for _ in range(reps + 1):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ix.predict_device(dq.data_ptr(), q, _native.CLASSIFICATION, out.data_ptr(), None, None, stream)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def main():
    rng = np.random.default_rng(0)
    for dtype in (np.float32, np.float64):
        for n, d in ((1000000, 128), (100000, 128), (20000, 32), (60000, 784)):
            if dtype == np.float64 and n * d > 2e7:
                continue
            train = np.asfortranarray(rng.standard_normal((n, d)).astype(dtype))
            ix = _native.Index(train, rng.integers(0, 10, n), 10, 1, 1e-8)
            for q in (32, 128, 512, 2048):
                dq = torch.from_numpy(rng.standard_normal((q, d)).astype(dtype)).cuda()
                out = torch.empty(q, dtype=torch.float64, device='cuda')
                t_tensor = time_path(ix, _native.PATH_TENSOR, dq, q, out)
                t_direct = time_path(ix, _native.PATH_DIRECT, dq, q, out)
                print(json.dumps({'dtype': np.dtype(dtype).name, 'n': n, 'd': d, 'q': q, 'nqd': float(n) * q * d,
                                  'tensor_ms': round(t_tensor, 3), 'direct_ms': round(t_direct, 3)}))
            ix.close()


if __name__ == '__main__':
    main()
