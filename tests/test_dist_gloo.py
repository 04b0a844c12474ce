"""World-size-2 gloo tests (CPU) of the multi-GPU host logic in osml_b200/sharded.py: shard bounds, the all-gather of
partial neighbour lists, the exact merge, and the query-sharded gather.  The CUDA backend is replaced by a test double
built on the oracle (test infrastructure); the collective code under test is the production code."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import knn_oracle as ko
from osml_b200 import sharded

SENTINEL = np.iinfo(np.int64).max


class OracleBackend:
    """Same interface as sharded.NativeBackend, CPU tensors, arithmetic from the oracle."""

    def __init__(self, train_shard, labels_full, k, rbf_gamma, epsilon, index_base, device, sum_order, path='auto'):
        self.train, self.labels, self.k, self.gamma, self.base = train_shard, labels_full, k, rbf_gamma, index_base

    def neighbours(self, queries):
        q, ksel = queries.shape[0], self.k + 1
        idx = np.full((q, ksel), SENTINEL, dtype=np.int64)
        dst = np.full((q, ksel), np.nan, dtype=np.float64 if self.train.dtype != np.float32 else np.float32)
        for i in range(q):
            d = ko.distances(self.train, queries[i])
            order = np.lexsort((np.arange(len(d)), d))[:ksel]
            idx[i, :len(order)] = order + self.base
            dst[i, :len(order)] = d[order]
        return torch.from_numpy(idx), torch.from_numpy(dst)

    def _merge(self, pi, pd_, task):
        g, q, ksel = pi.shape
        flat_i = pi.transpose(1, 0, 2).reshape(q, g * ksel)
        flat_d = pd_.transpose(1, 0, 2).reshape(q, g * ksel)
        out_i = np.empty((q, ksel), np.int64)
        out_d = np.empty((q, ksel), flat_d.dtype)
        for i in range(q):
            order = np.lexsort((flat_i[i], flat_d[i]))[:ksel]     # NaN sorts last, then by index
            out_i[i], out_d[i] = flat_i[i][order], flat_d[i][order]
        name = 'regression' if task == 0 else 'classification'
        return ko.predict_from_neighbours(out_i, out_d, self.labels, name, rbf_gamma=self.gamma), out_i, out_d

    def merge_predict(self, parts_idx, parts_dist, task, keep_on_device=False, return_lists=False):
        pred, out_i, out_d = self._merge(parts_idx.numpy(), parts_dist.numpy(), task)
        if return_lists:
            return torch.from_numpy(pred), torch.from_numpy(out_i), torch.from_numpy(out_d)
        return torch.from_numpy(pred) if keep_on_device else pred

    # the record layout of include/osml_knn.h (osml_knn_pack_slices_device): [parts][slice_q][k+1] x {index, distance bits}
    def pack_slices(self, idx, dist_, parts):
        q, ksel = idx.shape
        slice_q = -(-q // parts)
        rec = np.empty((parts * slice_q, ksel, 2), dtype=np.int64)
        rec[:, :, 0] = SENTINEL
        d = dist_.numpy()
        nan_bits = np.array([np.nan], dtype=d.dtype).view(np.uint32 if d.dtype == np.float32 else np.uint64)[0]
        rec[:, :, 1] = int(nan_bits)
        rec[:q, :, 0] = idx.numpy()
        rec[:q, :, 1] = d.view(np.uint32 if d.dtype == np.float32 else np.uint64).astype(np.int64)
        return torch.from_numpy(rec.reshape(parts, slice_q, ksel, 2))

    def merge_predict_packed(self, records, q_valid, task, return_lists=False):
        rec = records.numpy()
        parts, slice_q, ksel, _ = rec.shape
        f32 = self.train.dtype == np.float32
        out = np.zeros(slice_q)
        out_i = np.zeros((slice_q, ksel), np.int64)
        out_d = np.zeros((slice_q, ksel), np.float32 if f32 else np.float64)
        if q_valid > 0:
            pi = rec[:, :q_valid, :, 0]
            bits = np.ascontiguousarray(rec[:, :q_valid, :, 1])
            pd_ = bits.astype(np.uint32).view(np.float32) if f32 else bits.view(np.float64)
            out[:q_valid], out_i[:q_valid], out_d[:q_valid] = self._merge(np.ascontiguousarray(pi), pd_, task)
        if return_lists:
            return torch.from_numpy(out), torch.from_numpy(out_i), torch.from_numpy(out_d)
        return torch.from_numpy(out)

    def close(self):
        pass


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, results):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(42)
        n, d, q, k = 301, 5, 37, 6
        train = np.asfortranarray(rng.standard_normal((n, d)))
        queries = rng.standard_normal((q, d))
        labels_c = rng.integers(0, 4, n)
        labels_r = rng.standard_normal(n)
        out = {}
        for task, labels, name in ((1, labels_c, 'cls'), (0, labels_r, 'reg')):
            ix = sharded.TrainShardedIndex(train, labels, k, rbf_gamma=0.7, backend_factory=OracleBackend,
                                           query_block=16)
            assert ix.bounds == [(0, 151), (151, 301)] and ix.exchange == 'alltoall'
            out['train_' + name] = ix.predict(queries, task)          # blocks of 16, 16 and 5 queries: odd last slice
            ix = sharded.TrainShardedIndex(train, labels, k, rbf_gamma=0.7, backend_factory=OracleBackend,
                                           query_block=16, exchange='allgather')
            out['train_' + name + '_allgather'] = ix.predict(queries, task)
        # float32 data (4-byte distances in the low word of a record), one query: rank 1's slice is empty
        train32 = np.asfortranarray(train.astype(np.float32))
        ix = sharded.TrainShardedIndex(train32, labels_c, k, rbf_gamma=0.7, backend_factory=OracleBackend)
        out['train_cls_f32'] = ix.predict(queries.astype(np.float32), 1)
        out['train_cls_f32_one'] = ix.predict(queries[:1].astype(np.float32), 1)
        out['train_cls_f32_lists'] = ix.predict(queries.astype(np.float32), 1, return_neighbours=True)
        # shard-local construction: each rank only passes its own rows (uneven split on purpose)
        lo, hi = (0, 100) if rank == 0 else (100, n)
        ix = sharded.TrainShardedIndex(train[lo:hi], labels_c, k, rbf_gamma=0.7, backend_factory=OracleBackend,
                                       query_block=16, shard_of=(lo, n))
        out['train_cls_local'] = ix.predict(queries, 1)
        full = ko.predict_classification(train, labels_c, queries, k=k, rbf_gamma=0.7)

        def block(lo, hi):
            return full[lo:hi] * 1.0
        out['query'] = sharded.predict_query_sharded(block, q)
        out['query_tiny'] = sharded.predict_query_sharded(lambda lo, hi: full[lo:hi], 1)
        results[rank] = out
    finally:
        dist.destroy_process_group()


def test_split_bounds():
    assert sharded.split_bounds(10, 4) == [(0, 3), (3, 6), (6, 9), (9, 10)]
    assert sharded.split_bounds(2, 4) == [(0, 1), (1, 2), (2, 2), (2, 2)]
    assert sharded.split_bounds(20_000_000, 8)[7] == (17_500_000, 20_000_000)


def test_world_size_2_sharded_paths_match_the_unsharded_oracle():
    world = 2
    manager = mp.Manager()
    results = manager.dict()
    mp.spawn(_worker, args=(world, _free_port(), results), nprocs=world, join=True)
    rng = np.random.default_rng(42)
    n, d, q, k = 301, 5, 37, 6
    train = np.asfortranarray(rng.standard_normal((n, d)))
    queries = rng.standard_normal((q, d))
    labels_c = rng.integers(0, 4, n)
    labels_r = rng.standard_normal(n)
    want_c = ko.predict_classification(train, labels_c, queries, k=k, rbf_gamma=0.7)
    want_r = ko.predict_regression(train, labels_r, queries, k=k, rbf_gamma=0.7)
    for rank in range(world):
        got = results[rank]
        assert np.array_equal(got['train_cls'], want_c)
        assert np.array_equal(got['train_reg'], want_r)          # merge is exact: bit-identical to unsharded
        assert np.array_equal(got['train_cls_allgather'], want_c)
        assert np.array_equal(got['train_reg_allgather'], want_r)
        want32 = ko.predict_classification(np.asfortranarray(train.astype(np.float32)), labels_c,
                                           queries.astype(np.float32), k=k, rbf_gamma=0.7)
        assert np.array_equal(got['train_cls_f32'], want32)
        assert np.array_equal(got['train_cls_f32_one'], want32[:1])
        _, want_i, want_d = ko.predict_classification(np.asfortranarray(train.astype(np.float32)), labels_c,
                                                      queries.astype(np.float32), k=k, rbf_gamma=0.7,
                                                      return_neighbours=True)
        pred, got_i, got_d = got['train_cls_f32_lists']
        assert np.array_equal(pred, want32) and np.array_equal(got_i, want_i) and np.array_equal(got_d, want_d)
        assert np.array_equal(got['train_cls_local'], want_c)
        assert np.array_equal(got['query'], want_c)
        assert np.array_equal(got['query_tiny'], want_c[:1])


def test_single_process_degenerates_to_one_shard():
    rng = np.random.default_rng(1)
    train = np.asfortranarray(rng.standard_normal((50, 3)))
    labels = rng.integers(0, 3, 50)
    queries = rng.standard_normal((9, 3))
    ix = sharded.TrainShardedIndex(train, labels, 4, backend_factory=OracleBackend)
    assert np.array_equal(ix.predict(queries, 1), ko.predict_classification(train, labels, queries, k=4))
    with pytest.raises(ValueError):
        sharded.TrainShardedIndex(train[:4], labels[:4], 4, backend_factory=OracleBackend)
    with pytest.raises(ValueError):                              # labels must cover the whole matrix, not the shard
        sharded.TrainShardedIndex(train[:20], labels[:20], 4, backend_factory=OracleBackend, shard_of=(10, 50))
