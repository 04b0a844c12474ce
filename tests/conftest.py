import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, 'tests', 'golden')

# When the reference package is reachable (baseline/_ref travels to the GPU box; /root/reference exists in the authoring
# container only) the GPU classes are built on the reference's bases (osml_b200.model.reference_modules), which is what
# the meta-model tests need.  Must be set before osml_b200 is first imported.
for _root in (os.path.join(ROOT, 'baseline', '_ref'), '/root/reference'):
    if 'OSML_REF' not in os.environ and os.path.isfile(os.path.join(_root, 'osml', 'predict.py')):
        os.environ['OSML_REF'] = _root


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a real B200 (run with -m gpu on the GPU box)')


def golden_names():
    # kNN fixtures; transform_* and svm_* have their own loaders (load_transform_golden, tests/test_oracle_svm.py)
    return sorted(f[:-4] for f in os.listdir(GOLDEN_DIR)
                  if f.endswith('.npz') and not f.startswith(('transform_', 'svm_')))


def transform_golden_names():
    return sorted(f[:-4] for f in os.listdir(GOLDEN_DIR) if f.endswith('.npz') and f.startswith('transform_'))


def load_transform_golden(name):
    """Centering / Standardization fixtures (oracle/make_golden_transform.py); matrices re-laid-out as the reference
    saw / produced them."""
    import numpy as np
    z = np.load(os.path.join(GOLDEN_DIR, name + '.npz'))
    g = {key: z[key] for key in z.files}
    for key in ('fit_is_f_order', 'query_is_f_order', 'out_is_f_order', 'epsilon', 'numpy_version', 'pandas_version'):
        g[key] = g[key][()]
    order = lambda flag: np.asfortranarray if bool(flag) else np.ascontiguousarray   # noqa: E731
    g['fit'] = order(g['fit_is_f_order'])(g['fit'])
    g['query'] = order(g['query_is_f_order'])(g['query'])
    for key in ('centered_fit', 'centered_query', 'standardized_fit', 'standardized_query'):
        g[key] = order(g['out_is_f_order'])(g[key])
    g['epsilon'] = float(g['epsilon'])
    g['name'] = name
    return g


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


@pytest.fixture(params=transform_golden_names())
def transform_golden(request):
    return load_transform_golden(request.param)
