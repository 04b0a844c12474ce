"""Where the time of model.predict(DataFrame) goes at the C3 shape: python tools/df_path_timing.py [n d q]"""
import sys, time
import numpy as np, pandas as pd, torch
sys.path.insert(0, '.')
import osml_b200
from osml_b200 import _native
n, d, q = (int(a) for a in sys.argv[1:4]) if len(sys.argv) > 3 else (60000, 784, 10000)
rng = np.random.default_rng(0)
train = pd.DataFrame(np.asfortranarray(rng.standard_normal((n, d), dtype=np.float32)))
labels = pd.Series(rng.integers(0, 10, n), name='y')
test = pd.DataFrame(np.asfortranarray(rng.standard_normal((q, d), dtype=np.float32)))
model = osml_b200.KNearestNeighborsClassification(k=10)
model.fit(train, labels)
ix = model._index
def T(f, reps=20):
    for _ in range(3): f()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): f()
    return (time.perf_counter() - t0) / reps * 1e3
vals_f = test.values
vals_c = np.ascontiguousarray(vals_f)
pin_c = torch.from_numpy(vals_c.copy()).pin_memory().numpy()
pin_f = torch.from_numpy(np.ascontiguousarray(vals_f.T).copy()).pin_memory().numpy().T     # F-ordered view of pinned memory
print('F flags', vals_f.flags['F_CONTIGUOUS'], pin_f.flags['F_CONTIGUOUS'])
print('model.predict(DataFrame)        %.3f ms' % T(lambda: model.predict(test)))
print('validate                        %.3f ms' % T(lambda: model._validate_observations(test)))
print('index.predict pageable F        %.3f ms' % T(lambda: ix.predict(vals_f, _native.CLASSIFICATION)))
print('index.predict pageable C        %.3f ms' % T(lambda: ix.predict(vals_c, _native.CLASSIFICATION)))
print('index.predict pinned   F        %.3f ms' % T(lambda: ix.predict(pin_f, _native.CLASSIFICATION)))
print('index.predict pinned   C        %.3f ms' % T(lambda: ix.predict(pin_c, _native.CLASSIFICATION)))
pred = ix.predict(vals_f, _native.CLASSIFICATION)
print('Series + validate_labels        %.3f ms' % T(lambda: model._validate_labels(pd.Series(pred, name='y', dtype=np.int64))))
d_t = torch.empty(q * d, dtype=torch.float32, device='cuda')
src_pageable = torch.from_numpy(vals_c)
src_pinned = torch.from_numpy(pin_c)
print('torch H2D pageable %d MB         %.3f ms' % (q * d * 4 >> 20, T(lambda: (d_t.copy_(src_pageable.view(-1)), torch.cuda.synchronize()))))
print('torch H2D pinned                %.3f ms' % T(lambda: (d_t.copy_(src_pinned.view(-1), non_blocking=True), torch.cuda.synchronize())))
