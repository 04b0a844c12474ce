"""Imports the REAL reference (ViktorC/OSML at /root/reference or $OSML_REF) in the authoring container.

TEST INFRASTRUCTURE ONLY; never imported by ``osml_b200``.  The reference does not run unmodified on numpy 2 /
pandas 3 because it uses ``np.float_`` (data.py:157) and ``Series.iteritems`` (predict.py:278, 755), both removed
upstream.  Two aliases restore them without touching the reference sources.

The reference cannot travel to the GPU box, so everything that needs it (``oracle/make_golden.py``) runs here and
commits its outputs under ``tests/golden/``.
"""
import os
import sys

import numpy as np
import pandas as pd


def reference_root():
    for root in (os.environ.get('OSML_REF'), '/root/reference'):
        if root and os.path.isdir(os.path.join(root, 'osml')):
            return root
    return None


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
