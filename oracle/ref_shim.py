"""Imports the REAL reference (ViktorC/OSML): $OSML_REF, then baseline/_ref (the unmodified package as pip installed
it from /root/reference with oracle/install_reference.sh -- git-ignored, but it travels to the GPU box), then
/root/reference (the authoring container only).

TEST INFRASTRUCTURE ONLY; never imported by ``osml_b200``.  The reference does not run unmodified on numpy 2 /
pandas 3 because it uses ``np.float_`` (data.py:157) and ``Series.iteritems`` (predict.py:278, 755), both removed
upstream.  Two aliases restore them without touching the reference sources.

The reference's data sets (`data/`) exist under /root/reference only, so everything that needs them
(``oracle/make_golden.py``) runs in the authoring container and commits its outputs under ``tests/golden/``.
"""
import os
import sys

import numpy as np
import pandas as pd


_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def reference_root():
    for root in (os.environ.get('OSML_REF'), os.path.join(_REPO, 'baseline', '_ref'), '/root/reference'):
        if root and os.path.isfile(os.path.join(root, 'osml', 'predict.py')):
            return root
    return None


def data_root():
    """Directory that holds the reference's `data/` (the authoring container's checkout only)."""
    for root in (os.environ.get('OSML_REF'), '/root/reference'):
        if root and os.path.isdir(os.path.join(root, 'data')):
            return root
    raise ImportError('the reference data sets are not present (authoring container only)')


def import_reference():
    """Returns (osml.predict, osml.data, root) of the real reference, or raises ImportError."""
    root = reference_root()
    if root is None:
        raise ImportError('reference not present (expected in the authoring container only)')
    if not hasattr(np, 'float_'):
        np.float_ = np.float64
    if not hasattr(pd.Series, 'iteritems'):
        pd.Series.iteritems = pd.Series.items
    if root not in sys.path:
        sys.path.insert(0, root)
    import osml.data as osd
    import osml.predict as osp
    return osp, osd, root
