"""Multi-GPU host logic: one process per GPU (torchrun), `torch.distributed` for the plumbing.

The reference has no multi-device story; the two modes below are the ones north_star names
(SURVEY.md section 8e):

* **query-sharded** (`predict_query_sharded`): the train set is replicated, every rank answers a contiguous block of
  ceil(Q / G) queries, the per-rank predictions are all-gathered (Q doubles -- the only exchange, and only so that
  every rank returns the full Series like the reference would).  No data-path collective.
* **train-sharded** (`TrainShardedIndex`): rank r holds train rows [r * ceil(N / G), ...) plus the full label vector.
  Every rank runs the neighbour search over all queries on its rows (exact distances in the oracle's arithmetic,
  global row indices).  The partial (k+1)-lists are packed as 16-byte records, one block per destination rank, and
  ONE all-to-all over NCCL / NVLink hands rank r the G partial lists of ITS slice of ceil(Q / G) queries only; it
  merges them by (distance, index), aggregates, and a final all-gather of ceil(Q / G) doubles per rank gives every
  rank all Q predictions.  The merge is exact, so the result is bit-identical to the single-GPU one.
  (`exchange='allgather'` keeps the first round's form -- every rank receives and merges all Q lists -- as the
  independent check `bench.py` and the tests compare against.)

The arithmetic is always the CUDA path (`NativeBackend`); the `backend_factory` argument exists so the collective and
slicing logic can be exercised on CPU with gloo in tests/test_dist_gloo.py (with a test double, never in production).
"""
import numpy as np

from osml_b200 import _native


def split_bounds(n, parts):
    """Contiguous blocks of ceil(n / parts): [(lo, hi)] * parts (trailing blocks may be empty)."""
    step = -(-n // parts)
    return [(min(n, r * step), min(n, (r + 1) * step)) for r in range(parts)]


def _dist():
    import torch.distributed as dist
    return dist


def _world(group=None):
    dist = _dist()
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(group), dist.get_rank(group)
    return 1, 0


class NativeBackend:
    """The CUDA path behind the C ABI, speaking torch CUDA tensors for the collective."""

    def __init__(self, train_shard, labels_full, k, rbf_gamma, epsilon, index_base, device, sum_order, path='auto'):
        import torch
        self.torch = torch
        self.device = torch.device('cuda', device)
        self.index = _native.Index(train_shard, labels_full, k, rbf_gamma, epsilon, device=device,
                                   index_base=index_base, sum_order=sum_order)
        self.index.set_option('path', _native.PATHS[path])
        self.k = int(k)
        self.dist_dtype = torch.float32 if self.index.dist_dtype == np.float32 else torch.float64

    def neighbours(self, queries):
        """queries: host (Q, D) array, or a contiguous CUDA tensor already on this device; dtype and width must be
        the index's (the C entry points take raw pointers and cannot check)."""
        torch = self.torch
        q = queries.shape[0]
        if queries.ndim != 2 or queries.shape[1] != self.index.d:
            raise ValueError
        with torch.cuda.device(self.device):
            if isinstance(queries, torch.Tensor):
                if (queries.dtype != getattr(torch, np.dtype(self.index.dtype).name) or not queries.is_contiguous()
                        or queries.device != self.device):
                    raise ValueError
                d_q = queries
            else:
                if queries.dtype != self.index.dtype:
                    raise ValueError
                d_q = torch.from_numpy(np.ascontiguousarray(queries)).to(self.device)
            idx = torch.empty((q, self.k + 1), dtype=torch.int64, device=self.device)
            dist = torch.empty((q, self.k + 1), dtype=self.dist_dtype, device=self.device)
            stream = torch.cuda.current_stream(self.device).cuda_stream
            self.index.neighbours_device(d_q.data_ptr(), q, idx.data_ptr(), dist.data_ptr(), stream)
        return idx, dist

    def merge_predict(self, parts_idx, parts_dist, task, keep_on_device=False, return_lists=False):
        torch = self.torch
        parts, q, ksel = parts_idx.shape
        with torch.cuda.device(self.device):
            pred = torch.empty(q, dtype=torch.float64, device=self.device)
            m_idx = torch.empty((q, ksel), dtype=torch.int64, device=self.device) if return_lists else None
            m_dist = torch.empty((q, ksel), dtype=self.dist_dtype, device=self.device) if return_lists else None
            stream = torch.cuda.current_stream(self.device).cuda_stream
            self.index.merge_predict_device(parts_idx.data_ptr(), parts_dist.data_ptr(), parts, q, task,
                                            pred.data_ptr(), m_idx.data_ptr() if return_lists else None,
                                            m_dist.data_ptr() if return_lists else None, stream)
            if return_lists:
                return pred, m_idx, m_dist
            return pred if keep_on_device else pred.cpu().numpy()

    def pack_slices(self, idx, dist, parts):
        """Lists of all Q queries -> the all-to-all send buffer: int64 tensor [parts, ceil(Q / parts), k+1, 2]
        (16-byte records: global index, distance bits)."""
        torch = self.torch
        q, ksel = idx.shape
        slice_q = -(-q // parts)
        with torch.cuda.device(self.device):
            records = torch.empty((parts, slice_q, ksel, 2), dtype=torch.int64, device=self.device)
            stream = torch.cuda.current_stream(self.device).cuda_stream
            self.index.pack_slices_device(idx.data_ptr(), dist.data_ptr(), q, parts, records.data_ptr(), stream)
        return records

    def merge_predict_packed(self, records, q_valid, task, return_lists=False):
        """records [parts, slice_q, k+1, 2] = every rank's partial lists of this rank's slice -> fp64 predictions of
        the slice as a tensor of slice_q entries (the first q_valid are meaningful); with return_lists also the merged
        (index, distance) lists, slice_q rows each."""
        torch = self.torch
        parts, slice_q, ksel = records.shape[0], records.shape[1], records.shape[2]
        with torch.cuda.device(self.device):
            pred = torch.zeros(slice_q, dtype=torch.float64, device=self.device)
            m_idx = torch.zeros((slice_q, ksel), dtype=torch.int64, device=self.device) if return_lists else None
            m_dist = torch.zeros((slice_q, ksel), dtype=self.dist_dtype, device=self.device) if return_lists else None
            if q_valid > 0:
                stream = torch.cuda.current_stream(self.device).cuda_stream
                self.index.merge_predict_packed_device(records.data_ptr(), parts, slice_q, q_valid, task,
                                                       pred.data_ptr(), m_idx.data_ptr() if return_lists else None,
                                                       m_dist.data_ptr() if return_lists else None, stream)
        return (pred, m_idx, m_dist) if return_lists else pred

    def close(self):
        self.index.close()


class TrainShardedIndex:
    """Train rows sharded across the ranks of `group`; predictions bit-identical to the unsharded index."""

    def __init__(self, train, labels, k, rbf_gamma=1, epsilon=1e-8, device=None, group=None, backend_factory=None,
                 query_block=131072, path='auto', shard_of=None, exchange='alltoall'):
        """`train` is the whole (N, D) matrix (every rank slices its block), or -- with shard_of=(lo, N) -- this rank's
        rows [lo, lo + len(train)) of a matrix of N rows that no single process has to hold.  `labels` always has N
        entries."""
        train = np.asarray(train)
        self.world, self.rank = _world(group)
        self.group = group
        self.n = train.shape[0] if shard_of is None else int(shard_of[1])
        self.k = int(k)
        if self.n < self.k + 1:
            raise ValueError                                  # predict.py:768-769 on the *global* row count
        sum_order = _native.sum_order_of(train)               # layout of the whole matrix decides the order
        if shard_of is None:
            self.bounds = split_bounds(self.n, self.world)
            lo, hi = self.bounds[self.rank]
            if hi <= lo:
                raise ValueError('rank %d would hold no training rows' % self.rank)
            shard = train[lo:hi]
        else:
            lo = int(shard_of[0])
            shard = train
            if len(labels) != self.n or lo < 0 or lo + shard.shape[0] > self.n:
                raise ValueError
        if device is None:
            from osml_b200.predict import _default_device
            device = _default_device()
        factory = backend_factory or NativeBackend
        self.backend = factory(shard, np.ascontiguousarray(labels), k, rbf_gamma, epsilon, lo, device, sum_order,
                               path)
        self.query_block = int(query_block)
        if exchange not in ('alltoall', 'allgather'):
            raise ValueError
        self.exchange = exchange

    def predict(self, queries, task, keep_on_device=False, return_neighbours=False):
        """queries: the same (Q, D) matrix on every rank (host array, or a CUDA tensor on this rank's device).
        Returns the fp64 predictions for all Q on every rank (a CUDA tensor with keep_on_device); with
        return_neighbours also the merged global (index, distance) lists, (Q, k+1) each."""
        on_device = not isinstance(queries, np.ndarray) and hasattr(queries, 'data_ptr')
        if not on_device:
            queries = np.asarray(queries)
        if self.n < self.k + 2:
            raise ValueError                                  # np.argpartition(d, k + 1) needs k + 1 < N
        outs = []
        for q0 in range(0, queries.shape[0], self.query_block):
            block = queries[q0:q0 + self.query_block]
            idx, dist_ = self.backend.neighbours(block)
            if self.world > 1 and self.exchange == 'alltoall':
                res = self._exchange_slices(idx, dist_, task, return_neighbours)
            elif self.world > 1:
                res = self._exchange_all(idx, dist_, task, return_neighbours)
            else:
                res = self.backend.merge_predict(idx[None], dist_[None], task, keep_on_device=True,
                                                 return_lists=return_neighbours)
            outs.append(res if return_neighbours else (res,))
        import torch
        cols = [outs[0][c] if len(outs) == 1 else torch.cat([o[c] for o in outs]) for c in range(len(outs[0]))]
        if not keep_on_device:
            cols = [c.cpu().numpy() for c in cols]
        return tuple(cols) if return_neighbours else cols[0]

    def _exchange_slices(self, idx, dist_, task, want_lists=False):
        """One all-to-all of packed records (each rank receives the G partial lists of its own ceil(Q / G) queries),
        slice-local merge + aggregation, one all-gather of the slice predictions."""
        import torch
        dist = _dist()
        qb = idx.shape[0]
        send = self.backend.pack_slices(idx, dist_, self.world)           # [G][slice_q][k+1][2]
        slice_q = send.shape[1]
        recv = torch.empty_like(send)
        dist.all_to_all_single(recv, send, group=self.group)
        lo = min(qb, self.rank * slice_q)
        mine = self.backend.merge_predict_packed(recv, min(qb, lo + slice_q) - lo, task, return_lists=want_lists)
        gathered = []
        for part in (mine if want_lists else (mine,)):                    # slice_q rows each
            every = torch.empty((self.world * slice_q,) + tuple(part.shape[1:]), dtype=part.dtype, device=part.device)
            dist.all_gather_into_tensor(every, part.contiguous(), group=self.group)
            gathered.append(every[:qb])
        return tuple(gathered) if want_lists else gathered[0]

    def _exchange_all(self, idx, dist_, task, want_lists=False):
        """First-round form: all-gather of every rank's lists for ALL queries, every rank merges all of them."""
        import torch
        dist = _dist()
        qb, ksel = idx.shape
        all_idx = torch.empty((self.world * qb, ksel), dtype=idx.dtype, device=idx.device)
        all_dist = torch.empty((self.world * qb, ksel), dtype=dist_.dtype, device=dist_.device)
        dist.all_gather_into_tensor(all_idx, idx.contiguous(), group=self.group)
        dist.all_gather_into_tensor(all_dist, dist_.contiguous(), group=self.group)
        return self.backend.merge_predict(all_idx.view(self.world, qb, ksel),        # [part][q][k+1]
                                          all_dist.view(self.world, qb, ksel), task, keep_on_device=True,
                                          return_lists=want_lists)

    def close(self):
        self.backend.close()


def predict_query_sharded(predict_block, n_queries, group=None, comm_device=None):
    """Runs `predict_block(lo, hi) -> fp64 array` on this rank's block of queries and all-gathers the predictions so
    that every rank returns all `n_queries` of them.  `predict_block` is typically
    ``lambda lo, hi: model._ensure_index().predict(values[lo:hi], task)``."""
    world, rank = _world(group)
    bounds = split_bounds(n_queries, world)
    lo, hi = bounds[rank]
    mine = np.asarray(predict_block(lo, hi), dtype=np.float64) if hi > lo else np.zeros(0)
    if world == 1:
        return mine
    import torch
    dist = _dist()
    step = bounds[0][1] - bounds[0][0]
    if comm_device is None:
        comm_device = torch.device('cuda', torch.cuda.current_device()) if dist.get_backend(group) == 'nccl' \
            else torch.device('cpu')
    send = torch.zeros(step, dtype=torch.float64, device=comm_device)
    send[:hi - lo] = torch.from_numpy(mine).to(comm_device)
    recv = torch.empty(world * step, dtype=torch.float64, device=comm_device)
    dist.all_gather_into_tensor(recv, send, group=group)
    return recv.cpu().numpy()[:n_queries].copy() if world * step >= n_queries else None
