"""Generates tests/golden/transform_*.npz by running the REAL reference's Centering / Standardization
(osml/transform.py:53-84, through oracle/ref_shim.py) in the authoring container.

Run:  python oracle/make_golden_transform.py      (needs /root/reference; not runnable on the GPU box)
Stored per case: the input matrix (and its memory order), the fitted mean / sd, and the transformed matrices of the
fit frame and of a second ("query") frame, plus numpy / pandas versions.
"""
import os
import sys

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle.ref_shim import data_root, import_reference  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), 'tests', 'golden')


def main():
    osp, osd, _ = import_reference()
    root = data_root()
    import osml.transform as ost
    rng = np.random.default_rng(77)
    cases = {}
    np.random.seed(0)
    iris = osd.IrisDataSet(os.path.join(root, 'data', 'iris/iris.csv'))
    cases['iris_f64'] = (iris.get_training_observations(), iris.get_test_observations())
    x32 = (rng.standard_normal((3000, 17)) * rng.uniform(0.1, 30, 17) + rng.uniform(-50, 50, 17)).astype(np.float32)
    cases['synthetic_f32'] = (pd.DataFrame(x32[:2500]), pd.DataFrame(x32[2500:]))
    xi = rng.integers(-100, 100, (1200, 9)).astype(np.int8)
    cases['synthetic_i8'] = (pd.DataFrame(xi[:1000]), pd.DataFrame(xi[1000:]))
    x64c = np.ascontiguousarray(rng.standard_normal((800, 6)) * 1e3)
    cases['synthetic_f64'] = (pd.DataFrame(x64c[:700]), pd.DataFrame(x64c[700:]))
    for name, (fit_df, query_df) in cases.items():
        fit_df = fit_df.reset_index(drop=True)
        query_df = query_df.reset_index(drop=True)
        cen, std = ost.Centering(), ost.Standardization()
        cen_fit = cen.fit_transform(fit_df, None).values
        std_fit = std.fit_transform(fit_df, None).values
        cen_q = cen.transform(query_df).values
        std_q = std.transform(query_df).values
        values = fit_df.values
        np.savez_compressed(
            os.path.join(OUT, 'transform_%s.npz' % name), fit=np.asarray(values), query=np.asarray(query_df.values),
            fit_is_f_order=bool(values.flags['F_CONTIGUOUS'] and not values.flags['C_CONTIGUOUS']),
            query_is_f_order=bool(query_df.values.flags['F_CONTIGUOUS'] and not query_df.values.flags['C_CONTIGUOUS']),
            mean=std._mean, sd=std._sd, centering_mean=cen._mean, epsilon=std.epsilon,
            centered_fit=cen_fit, centered_query=cen_q, standardized_fit=std_fit, standardized_query=std_q,
            out_is_f_order=bool(std_fit.flags['F_CONTIGUOUS'] and not std_fit.flags['C_CONTIGUOUS']),
            numpy_version=np.__version__, pandas_version=pd.__version__)
        print(name, values.dtype, 'F' if values.flags['F_CONTIGUOUS'] else 'C', '->', std_fit.dtype,
              'F' if std_fit.flags['F_CONTIGUOUS'] else 'C', std_fit.shape)


if __name__ == '__main__':
    main()
