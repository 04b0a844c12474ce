"""Small invocations of every kernel family, cross-checked against each other (tensor path vs direct kernel, packed
exchange vs the lists it was packed from), sized for `compute-sanitizer --tool memcheck`:
    compute-sanitizer --tool memcheck --error-exitcode 7 python tools/sanitize_cases.py
(compute-sanitizer is closed on the round-2 GPU pool -- "runs under it have left GPUs needing a reset" -- so round 2 only
ran it plain: gpurun_out/sanitize_plain.log.)"""
import sys

import numpy as np
import torch

sys.path.insert(0, '.')
from osml_b200 import _native  # noqa: E402


def knn(train, labels, test, k, path, task=_native.CLASSIFICATION):
    ix = _native.Index(train, labels, k, 1, 1e-8)
    ix.set_option('path', _native.PATHS[path])
    out = ix.predict(test, task, return_neighbours=True)
    ix.close()
    return out


def main():
    rng = np.random.default_rng(0)
    # K1 pair kernel (resident query tile), K1 wide (K-block ring), K3, K4; K2 as the cross-check
    for n, d, q, k in ((20000, 64, 300, 10), (17000, 300, 200, 5), (18000, 96, 260, 32)):
        train = np.asfortranarray(rng.standard_normal((n, d), dtype=np.float32))
        test = rng.standard_normal((q, d), dtype=np.float32)
        labels = rng.integers(0, 5, n)
        a, b = knn(train, labels, test, k, 'tensor'), knn(train, labels, test, k, 'direct')
        assert all(np.array_equal(x, y) for x, y in zip(a, b)), (n, d)
        print('tensor == direct', n, d, q, k)
    # K2 with a chunked query tile (6000 features), k = 200, float32 regression labels
    train = np.asfortranarray(rng.standard_normal((400, 6000), dtype=np.float32))
    test = rng.standard_normal((9, 6000), dtype=np.float32)
    knn(train, rng.standard_normal(400).astype(np.float32), test, 5, 'direct', _native.REGRESSION)
    train = np.asfortranarray(rng.standard_normal((1200, 21), dtype=np.float32))
    knn(train, rng.integers(0, 6, 1200), rng.standard_normal((24, 21), dtype=np.float32), 200, 'direct')
    knn(np.ascontiguousarray(train), rng.standard_normal(1200), rng.standard_normal((24, 21), dtype=np.float32), 150,
        'direct', _native.REGRESSION)                              # C order: the generic kernel with a narrow CTA
    print('direct kernel variants ok')
    # integer matrix on the filter (shadow, range flags, conversion back for K2)
    train = np.asfortranarray(rng.integers(0, 11, (20000, 16)).astype(np.int8))
    test = rng.integers(0, 11, (64, 16)).astype(np.int8)
    test[3, 2] = -100
    labels = rng.integers(0, 3, 20000)
    a, b = knn(train, labels, test, 5, 'tensor'), knn(train, labels, test, 5, 'direct')
    assert all(np.array_equal(x, y, equal_nan=True) for x, y in zip(a, b))
    print('integer filter ok')
    # packed exchange records: three "ranks" emulated on one index
    n, d, q, k = 3000, 8, 50, 6
    train = np.asfortranarray(rng.standard_normal((n, d), dtype=np.float32))
    test = rng.standard_normal((q, d), dtype=np.float32)
    labels = rng.integers(0, 4, n)
    ix = _native.Index(train, labels, k, 1, 1e-8)
    dq = torch.from_numpy(test).cuda()
    idx = torch.empty((q, k + 1), dtype=torch.int64, device='cuda')
    dist = torch.empty((q, k + 1), dtype=torch.float32, device='cuda')
    ix.neighbours_device(dq.data_ptr(), q, idx.data_ptr(), dist.data_ptr(), None)
    parts, slice_q = 3, -(-q // 3)
    rec = torch.empty((parts, slice_q, k + 1, 2), dtype=torch.int64, device='cuda')
    ix.pack_slices_device(idx.data_ptr(), dist.data_ptr(), q, parts, rec.data_ptr(), None)
    torch.cuda.synchronize()
    # every "rank" sent the same lists: slice s of the merged result must equal the lists of its queries
    for s in range(parts):
        recv = rec[s:s + 1].expand(parts, -1, -1, -1).contiguous()      # [parts][slice_q][k+1][2]
        valid = min(q, (s + 1) * slice_q) - s * slice_q
        pred = torch.zeros(slice_q, dtype=torch.float64, device='cuda')
        m_idx = torch.zeros((slice_q, k + 1), dtype=torch.int64, device='cuda')
        m_dist = torch.zeros((slice_q, k + 1), dtype=torch.float32, device='cuda')
        ix.merge_predict_packed_device(recv.data_ptr(), parts, slice_q, valid, _native.CLASSIFICATION, pred.data_ptr(),
                                       m_idx.data_ptr(), m_dist.data_ptr(), None)
        torch.cuda.synchronize()
        # identical lists in every part: the merge keeps each entry's first copies -> entries repeat `parts` times
        want = idx[s * slice_q:s * slice_q + valid].repeat_interleave(parts, dim=1)[:, :k + 1]
        assert torch.equal(m_idx[:valid], want)
    ix.close()
    print('packed records ok')
    # support-vector decision function: query counts around the split boundaries the advisor named
    vectors = np.ascontiguousarray(rng.standard_normal((700, 12), dtype=np.float32))
    sv_labels = rng.integers(0, 2, 700) * 2. - 1.
    alphas = rng.uniform(0, 1, 700)
    for kernel in ('rbf', 'lin'):
        m = _native.Machine(vectors, sv_labels, alphas, -.3, kernel, gamma=.1)
        for qn in (5, 1184, 1500, 2360):
            m.decision(rng.standard_normal((qn, 12), dtype=np.float32))
        m.close()
    print('svm decision ok')


if __name__ == '__main__':
    main()
