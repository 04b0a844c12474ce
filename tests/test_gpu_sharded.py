"""Multi-GPU tests (need >= 2 B200s: `gpurun --gpus 2`): the train-sharded mode over NCCL must be bit-identical to one
GPU, and the query-sharded gather must reproduce the single-GPU predictions."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, results):
    import torch.distributed as dist
    from osml_b200 import _native, sharded
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    try:
        rng = np.random.default_rng(7)
        n, d, q, k = 70001, 96, 1500, 32
        train = np.asfortranarray(rng.standard_normal((n, d), dtype=np.float32))
        queries = rng.standard_normal((q, d), dtype=np.float32)
        labels = rng.integers(0, 10, n)
        ix = sharded.TrainShardedIndex(train, labels, k, device=rank, query_block=512)
        got = ix.predict(queries, _native.CLASSIFICATION)
        ix.close()
        whole = _native.Index(train, labels, k, 1, 1e-8, device=rank)
        want = whole.predict(queries, _native.CLASSIFICATION)
        gathered = sharded.predict_query_sharded(
            lambda lo, hi: whole.predict(queries[lo:hi], _native.CLASSIFICATION), q)
        whole.close()
        results[rank] = (bool(np.array_equal(got, want)), bool(np.array_equal(gathered, want)))
    finally:
        dist.destroy_process_group()


def test_train_sharded_and_query_sharded_over_nccl():
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip('needs at least 2 GPUs')
    import torch.multiprocessing as mp
    manager = mp.Manager()
    results = manager.dict()
    mp.spawn(_worker, args=(world, _free_port(), results), nprocs=world, join=True)
    for rank in range(world):
        assert results[rank] == (True, True), (rank, results[rank])
