"""Standardization in front of kNN at the C4 shape (not the headline bench): pipeline queries/s for
  (a) the reference's way  -- numpy transform of the query frame on the host, then predict,
  (b) osml_b200.Standardization.transform (GPU kernel, host in / host out), then predict,
  (c) the transform fused into predict (raw queries go straight to the device),
and the HBM roofline of the transform kernel from its CUDA-event time on device-resident data.
    python tools/bench_transform.py [n] [d] [q] [k]"""
import json
import os
import sys
import time

import numpy as np
import pandas as pd
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import osml_b200 as ob  # noqa: E402
from osml_b200 import _native  # noqa: E402


def main():
    n, d, q, k = (int(x) for x in (sys.argv[1:5] + ['1000000', '128', '100000', '10'][len(sys.argv) - 1:]))
    rng = np.random.default_rng(4)
    scale, off = rng.uniform(0.5, 20, d).astype(np.float32), rng.uniform(-30, 30, d).astype(np.float32)
    train = pd.DataFrame(rng.standard_normal((n, d), dtype=np.float32) * scale + off)
    queries = pd.DataFrame(rng.standard_normal((q, d), dtype=np.float32) * scale + off)
    labels = pd.Series(rng.integers(0, 10, n), name='y')
    std = ob.Standardization()
    model = ob.KNearestNeighborsClassification(k=k)
    model.fit(std.fit_transform(train, labels), labels)

    def timed(fn, reps=5):
        fn()
        t0 = time.perf_counter()
        for _ in range(reps):
            out = fn()
        return (time.perf_counter() - t0) / reps, out

    def host_two_step():                       # what the reference does before predict (osml/transform.py:83-84)
        x = pd.DataFrame((queries.values - std._mean) / (std._sd + std.epsilon))
        return model.predict(x)

    def gpu_two_step():
        return model.predict(std.transform(queries))

    t_host, p_host = timed(host_two_step)
    t_gpu, p_gpu = timed(gpu_two_step)
    model.set_query_transform(std)
    t_fused, p_fused = timed(lambda: model.predict(queries))
    assert np.array_equal(p_host.values, p_gpu.values) and np.array_equal(p_host.values, p_fused.values)

    # kernel-only: the fused transform on device-resident queries (difference of two device-timed predicts)
    index = model._ensure_index()
    dq = torch.from_numpy(np.ascontiguousarray(queries.values)).cuda()
    out = torch.empty(q, dtype=torch.float64, device='cuda')
    stream = torch.cuda.current_stream().cuda_stream

    def device_ms(buf, reps=10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        index.predict_device(buf.data_ptr(), q, _native.CLASSIFICATION, out.data_ptr(), None, None, stream)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):
            index.predict_device(buf.data_ptr(), q, _native.CLASSIFICATION, out.data_ptr(), None, None, stream)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    with_t = device_ms(dq)                           # raw queries, transform fused
    index.set_query_transform(None)
    dq_std = torch.from_numpy(np.ascontiguousarray(std.transform(queries).values)).cuda()
    without_t = device_ms(dq_std)                    # already standardised queries, no transform
    peaks = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'MEASURED_PEAKS.json'))) \
        if os.path.exists(os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'MEASURED_PEAKS.json')) else {}
    print(json.dumps({
        'workload': 'Standardization + kNN classify, %dx%d fp32 train, %d queries, k=%d' % (n, d, q, k),
        'pipeline_queries_per_s': {'host numpy transform + predict (reference way)': q / t_host,
                                   'GPU transform (host in/out) + predict': q / t_gpu,
                                   'transform fused into predict': q / t_fused},
        'ms': {'host_two_step': t_host * 1e3, 'gpu_two_step': t_gpu * 1e3, 'fused': t_fused * 1e3,
               'device_predict_raw_queries_fused_transform': with_t,
               'device_predict_prestandardised_queries': without_t},
        'transform_kernel': {'algorithmic_bytes': 2 * q * d * 4, 'hbm_peak_gbs': peaks.get('hbm_gbs'),
                             'note': 'see the ncu launch list for column_transform_kernel<float, float>'}}))


if __name__ == '__main__':
    main()
