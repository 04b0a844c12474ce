"""Support-vector decision function (not the headline bench): throughput of
osml_svm_decision on synthetic data, the CPU oracle port of osml/predict.py:1095-1099 beside it, and the SMO fit of the
reference's titanic configuration with the error cache on the GPU.
    python tools/bench_svm.py [n_vectors] [d] [queries]"""
import json
import os
import sys
import time

import numpy as np
import pandas as pd

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import osml_b200 as ob  # noqa: E402
from oracle import svm_oracle as so  # noqa: E402  (CPU baseline leg only)
from osml_b200 import _native  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests', 'golden')


def decision_line(dtype, order, n, d, q, reps=5):
    rng = np.random.default_rng(1)
    vectors = np.asarray(rng.standard_normal((n, d)).astype(dtype), order=order)
    queries = rng.standard_normal((q, d)).astype(dtype)
    labels = rng.integers(0, 2, n) * 2. - 1.
    alphas = rng.uniform(0, 1, n)
    machine = _native.Machine(vectors, labels, alphas, .1, 'rbf', gamma=.05)
    machine.decision(queries)
    t0 = time.perf_counter()
    kernel_ms = 0.
    for _ in range(reps):
        out = machine.decision(queries)
        kernel_ms += machine.stats()['last_kernel_ms']
    e2e_s = (time.perf_counter() - t0) / reps
    kernel_ms /= reps
    sample = 16
    t0 = time.perf_counter()
    want = so.raw_predict(queries[:sample], vectors, labels, alphas, .1, 'rbf', gamma=.05)
    cpu_s = (time.perf_counter() - t0) / sample
    tol = 1e-5 if dtype == np.float32 else 1e-12
    assert np.max(np.abs(out[:sample] - want)) <= tol * alphas.sum()
    machine.close()
    return {'what': 'osml_svm_decision, rbf', 'dtype': np.dtype(dtype).name,
            'vector_order': order + (' (numpy pairwise sum over features)' if order == 'C' else ' (sequential sum)'),
            'vectors': n, 'features': d, 'queries': q,
            'e2e_queries_per_s': q / e2e_s, 'kernel_ms': kernel_ms,
            'kernel_pair_elements_per_s': n * d * q / (kernel_ms / 1e3),
            'kernel_instr_per_s_3_per_element': 3.0 * n * d * q / (kernel_ms / 1e3),
            'cpu_port_queries_per_s_1_core': 1. / cpu_s, 'cpu_sample': '%d queries through oracle/svm_oracle.py' % sample}


def fit_line():
    g = np.load(os.path.join(GOLDEN, 'svm_titanic_lin.npz'))
    train_df = pd.DataFrame(np.asfortranarray(g['train']))
    train_sr = pd.Series(g['train_labels'], name=str(g['label_name']))
    model = ob.SupportVectorMachine(tol=1e-2, c=.1, kernel='lin')
    np.random.seed(int(g['fit_seed']))
    t0 = time.perf_counter()
    model.fit(train_df, train_sr)
    fit_s = time.perf_counter() - t0
    return {'what': 'SupportVectorMachine(tol=1e-2, c=.1, kernel="lin").fit on titanic (621 x %d)' % train_df.shape[1],
            'fit_s': fit_s, 'smo_steps': model.smo_steps, 'ms_per_step': 1e3 * fit_s / max(1, model.smo_steps),
            'support_vectors': int(len(model._support_vectors)),
            'reference_fit_s_authoring_container': 26.0}


def main():
    n, d, q = (int(x) for x in (sys.argv[1:4] + ['20000', '128', '100000'][len(sys.argv) - 1:]))
    for line in (decision_line(np.float32, 'F', n, d, q), decision_line(np.float32, 'C', n, d, q),
                 decision_line(np.float64, 'C', n, d, q // 4), fit_line()):
        print(json.dumps(line))


if __name__ == '__main__':
    main()
