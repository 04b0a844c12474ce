import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a real B200 (run with -m gpu on the GPU box)')


def golden_names():
    return sorted(f[:-4] for f in os.listdir(GOLDEN_DIR) if f.endswith('.npz'))


def load_golden(name):
    """Returns a dict; 'train' is re-laid-out in the memory order the reference saw (F or C)."""
    import numpy as np
    z = np.load(os.path.join(GOLDEN_DIR, name + '.npz'))
    g = {key: z[key] for key in z.files}
    for key in ('k', 'rbf_gamma', 'task', 'pred_sha256_16', 'numpy_version', 'pandas_version', 'ref_metric'):
        g[key] = g[key][()]
    g['k'] = int(g['k'])
    g['task'] = str(g['task'])
    gamma = g['rbf_gamma']
    g['rbf_gamma'] = int(gamma) if float(gamma).is_integer() else float(gamma)
    if bool(g['train_is_f_order']):
        g['train'] = np.asfortranarray(g['train'])
    else:
        g['train'] = np.ascontiguousarray(g['train'])
    g['name'] = name
    return g


@pytest.fixture(params=golden_names())
def golden(request):
    return load_golden(request.param)
