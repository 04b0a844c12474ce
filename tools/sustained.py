"""Same-regime comparison of K1 with the yardstick it is measured against ("the roofline fraction is quoted
against the wrong regime"): MEASURED_PEAKS.json's sustained bf16 figure is a cuBLAS 8192^3 GEMM looped for 4 s under the
1 kW power cap.  This tool loops BOTH on one box, back to back, for `seconds` each -- the cuBLAS GEMM and the C4 predict
step -- with the SM clock sampled during each loop, and prints one JSON object:

    python tools/sustained.py [seconds]          (needs a B200; `gpurun`)

`k1.tflops` over `cublas.tflops` is the fraction of what the tensor pipe sustains on that box at that power cap."""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, '.')
import bench  # noqa: E402
from osml_b200 import _native  # noqa: E402


def loop(fn, seconds, sampler):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    sampler.start()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    reps = 0
    while time.perf_counter() - t0 < seconds:
        for _ in range(5):
            fn()
        reps += 5
        torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    return reps, e0.elapsed_time(e1), sampler.stop()


def main():
    seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 4.0
    out = {'seconds_per_loop': seconds}
    a = torch.randn(8192, 8192, dtype=torch.bfloat16, device='cuda')
    b = torch.randn(8192, 8192, dtype=torch.bfloat16, device='cuda')
    c = torch.empty(8192, 8192, dtype=torch.bfloat16, device='cuda')
    reps, ms, clocks = loop(lambda: torch.matmul(a, b, out=c), seconds, bench.ClockSampler(0))
    out['cublas'] = {'what': 'torch.matmul bf16 8192^3 back to back', 'reps': reps, 'ms_per_rep': ms / reps,
                     'tflops': 2 * 8192 ** 3 * reps / (ms / 1e3) / 1e12, 'clocks': clocks}
    del a, b, c

    train, labels, queries, k = bench.make_data('c4', 0)
    n, d = train.shape
    q = queries.shape[0]
    index = _native.Index(train, labels, k, 1, 1e-8)
    d_queries = torch.from_numpy(queries).cuda()
    d_pred = torch.empty(q, dtype=torch.float64, device='cuda')
    stream = torch.cuda.current_stream().cuda_stream
    k1 = {'ms': 0.0, 'launches': 0}

    def step():
        index.predict_device(d_queries.data_ptr(), q, _native.CLASSIFICATION, d_pred.data_ptr(), None, None, stream)
        st = index.stats()
        k1['ms'] += st['last_filter_ms']
        k1['launches'] += st['last_filter_launches']
    for _ in range(3):
        step()
    k1.update(ms=0.0, launches=0)
    first = {}

    def counted():
        step()
        if not first:
            first['ms'] = k1['ms']
    reps, ms, clocks = loop(counted, seconds, bench.ClockSampler(0))
    # the loop's own warm-up steps are in k1 too: normalise by launches
    out['k1'] = {'what': 'C4 predict step (1M x 128 train, 100k queries) back to back', 'reps': reps,
                 'ms_per_step': ms / reps, 'k1_ms_per_launch': k1['ms'] / k1['launches'],
                 'tflops': 2.0 * n * d * q / (k1['ms'] / k1['launches'] / 1e3) / 1e12,
                 'step_tflops': 2.0 * n * d * q * reps / (ms / 1e3) / 1e12, 'clocks': clocks}
    out['k1_over_cublas_sustained'] = out['k1']['tflops'] / out['cublas']['tflops']
    out['step_over_cublas_sustained'] = out['k1']['step_tflops'] / out['cublas']['tflops']
    peaks = bench.load_peaks()
    out['measured_peaks'] = {'bf16_tflops_burst': peaks[0], 'bf16_tflops_sustained': peaks[1], 'source': peaks[3]}
    print(json.dumps(out))


if __name__ == '__main__':
    main()
