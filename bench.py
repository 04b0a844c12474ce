#!/usr/bin/env python
"""Benchmark of the kNN predict hot path (BASELINE.json metric: kNN predict queries/sec, k=10, on B200).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c4|c3|c5]

Workload (default, BASELINE.json configs[3], "C4"): synthetic 1,000,000 x 128 fp32 train set resident on every GPU,
100,000 fp32 queries per GPU per step, k = 10, classification labels in 0..9 -- the reference's
KNearestNeighborsClassification.predict (osml/predict.py:805-811).  Queries shard across GPUs with no collective
(SURVEY.md section 8e), so per-GPU work is fixed as N grows: "scaling": "weak".

One step = one pass of the hot path over one batch of 100,000 queries:
  value  queries/s with the queries already in HBM (osml_knn_predict_device, CUDA events, max over ranks)
  e2e    queries/s through the C ABI with HOST buffers (osml_knn_predict: pinned host queries -> H2D -> kernels ->
         D2H predictions), copies inside the timed region
  roofline  of the dominant kernel (K1, the tcgen05 filter): algorithmic 2*N*D FLOP per query
         over its CUDA-event duration, against the measured dense bf16 peaks of MEASURED_PEAKS.json -- the burst figure
         when the timed region is shorter than the 4 s loop the sustained figure was measured over, both fractions
         printed either way
  cpu_baseline  the reference's own predict() (baseline/_ref, see oracle/install_reference.sh; the numpy oracle port if
         it is absent) on a bounded sample of the same queries, 1 core
`--impl reference` times that CPU path with every host core it can use and prints the same JSON shape.

With N > 1 GPUs the same line carries two more legs (keys of the headline line, not lines of their own):
  strong  BASELINE.json configs[3] as written: the SAME 100,000 queries split N ways (12,500 per GPU at N = 8)
  c5      BASELINE.json configs[4] shape: train rows SHARDED across the GPUs (2,500,000 x 96 fp32 rows per GPU, 20M x 96
          at 8 GPUs), every GPU searches the same 131,072 queries per step (a subset of the config's 1M, stated), k = 32;
          the exact per-GPU partial lists go through ONE NCCL all-to-all (each rank receives its slice of the queries
          from every peer), slice-local merge, one all-gather of the predictions (osml_b200/sharded.py).  `matches` is
          a bit-for-bit parity flag: for 64 sampled queries every rank's exact direct-kernel (K2) partial lists are
          all-gathered and merged on the HOST with np.lexsort, and indices, distances and labels must equal the
          GPU-merged result of the tensor path.
Every rank also checks a sample of ITS queries of the headline leg against the numpy oracle (`oracle_check`).
`--workload c5` runs the train-sharded leg alone as its own line (1-GPU profile runs).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (n, d, queries per GPU per step, k, seed)
    'c4': (1_000_000, 128, 100_000, 10, 4),
    'c3': (60_000, 784, 10_000, 10, 3),
}
METRIC = 'kNN predict queries/sec (k=10)'
C5 = {'rows_per_gpu': 2_500_000, 'd': 96, 'q': 131_072, 'k': 32, 'seed': 5}


def make_data(name, rank):
    n, d, q, k, seed = WORKLOADS[name]
    rng = np.random.default_rng(seed)
    train = np.asfortranarray(rng.standard_normal((n, d), dtype=np.float32))   # DataFrame.values is F-ordered
    labels = rng.integers(0, 10, n)
    qrng = np.random.default_rng(seed + 1 + 1000 * rank)
    queries = qrng.standard_normal((q, d), dtype=np.float32)
    return train, labels, queries, k


def workload_description(name, n_gpus):
    n, d, q, k, seed = WORKLOADS[name]
    return {'workload': '%s: synthetic gaussian %dx%d fp32 train (replicated per GPU), %d fp32 queries per GPU per step '
                        '(%d in total), k=%d, KNearestNeighborsClassification, 10 classes'
                        % (name.upper(), n, d, q, q * n_gpus, k),
            'train_rows': n, 'features': d, 'queries_per_gpu_per_step': q, 'k': k, 'task': 'classification',
            'parallelism': 'query-sharded x%d, no collective' % n_gpus,
            'l2': 'inputs larger than L2: %.0f MB fp32 train + bf16 filter copy per pass vs 126 MB L2'
                  % (n * d * 4 / 1e6)}


# ----------------------------------------------------------------------------------------------------------------------
# CPU arms: the reference's own predict() when the package is reachable ($OSML_REF, baseline/_ref, /root/reference;
# oracle/ref_shim.py), else the oracle port of its numpy path.  Test / baseline infrastructure, timed here only.
# ----------------------------------------------------------------------------------------------------------------------
_CPU = {}


def _cpu_init(train, labels, k):
    """Fits the CPU arm once (before any fork): returns 'reference' or 'port'."""
    _CPU['train'], _CPU['labels'], _CPU['k'] = train, labels, k
    _CPU['model'] = None
    try:
        import pandas as pd
        from oracle.ref_shim import import_reference
        osp, _, root = import_reference()
        model = osp.KNearestNeighborsClassification(k=k)
        model.fit(pd.DataFrame(train), pd.Series(labels, name='label'))      # osml/predict.py:767-771: keeps .values
        _CPU['model'] = model
        _CPU['where'] = root
        return 'reference'
    except ImportError:
        return 'port'


def _cpu_predict(queries):
    if _CPU.get('model') is not None:
        import pandas as pd
        return _CPU['model'].predict(pd.DataFrame(queries)).values.astype(np.float64)   # osml/predict.py:29-46, 805-811
    from oracle import knn_oracle as ko
    return ko.predict_classification(_CPU['train'], _CPU['labels'], queries, k=_CPU['k'], tie_rule='reference')


def _cpu_description(kind):
    if kind == 'reference':
        return 'the unmodified reference (osml.predict.KNearestNeighborsClassification.predict from %s)' % _CPU['where']
    return 'numpy oracle port of osml/predict.py:740-811 (the reference package is not reachable here)'


def cpu_baseline_single_core(train, labels, queries, k, budget_s=20.0):
    """The reference on one core (it is single-threaded: predict.py:787, 807); bounded sample."""
    kind = _cpu_init(train, labels, k)
    t0 = time.perf_counter()
    _cpu_predict(queries[:1])
    per_query = time.perf_counter() - t0
    sample = int(max(2, min(64, budget_s / max(per_query, 1e-6))))
    t0 = time.perf_counter()
    pred = _cpu_predict(queries[:sample])
    dt = time.perf_counter() - t0
    return sample / dt, sample, pred, kind


def run_reference_arm(args):
    """`--impl reference`: the reference's CPU path on all host cores (fork pool, queries split across workers)."""
    import multiprocessing as mp
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return 0
    train, labels, queries, k = make_data(args.workload, 0)
    cores = len(os.sched_getaffinity(0))
    # every worker holds two N x D temporaries per query (predict.py:742): bound the pool by the memory that is there
    per_worker_bytes = 3 * train.nbytes
    try:
        import psutil
        by_memory = int(psutil.virtual_memory().available * 0.5 // per_worker_bytes)
    except Exception:
        by_memory = 16
    cores = max(1, min(cores, by_memory, 64))
    kind = _cpu_init(train, labels, k)
    ctx = mp.get_context('fork')
    with ctx.Pool(cores) as pool:
        # bounded sample per step: about 4 s of wall clock per step with every worker busy (calibrated under load:
        # the numpy path is DRAM-bound, so a query takes longer when all cores stream at once)
        pool.map(_cpu_predict, [queries[i:i + 1] for i in range(cores)])
        t0 = time.perf_counter()
        pool.map(_cpu_predict, [queries[i:i + 1] for i in range(cores)])
        per_query = time.perf_counter() - t0
        per_worker = int(max(1, min(16, 4.0 / max(per_query, 1e-6))))
        sample = per_worker * cores
        chunks = [queries[i * per_worker:(i + 1) * per_worker] for i in range(cores)]
        for _ in range(args.warmup):
            pool.map(_cpu_predict, chunks)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            pool.map(_cpu_predict, chunks)
        dt = time.perf_counter() - t0
    value = sample * args.steps / dt
    line = {'metric': METRIC, 'value': value, 'unit': 'queries/s', 'n_gpus': args.gpus, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': dt / args.steps * 1e3, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic', 'impl': 'reference',
            'config': workload_description(args.workload, args.gpus),
            'cpu_baseline': {'value': value, 'unit': 'queries/s', 'cores': cores, 'kind': kind,
                             'sample': '%d queries per step (%d per worker process), %s'
                                       % (sample, per_worker, _cpu_description(kind))},
            'e2e': {'value': value, 'unit': 'queries/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    print(json.dumps(line))
    return 0


# ----------------------------------------------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
              'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
              'clocks_event_reasons.sw_power_cap')

    def __init__(self, device):
        self.device = device
        self.samples = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.device), '--query-gpu=' + self.FIELDS,
                                          '--format=csv,noheader,nounits', '-lms', '100'], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            parts = [p.strip() for p in line.split(',')]
            if len(parts) >= 7:
                self.samples.append(parts)

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm = []
        reasons = set()
        sm_max = None
        for p in self.samples:
            try:
                sm.append(float(p[0]))
                sm_max = float(p[1])
            except ValueError:
                continue
            for name, flag in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), p[3:7]):
                if flag.lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': sm_max, 'reasons': sorted(reasons),
                'samples': len(sm)}


# ----------------------------------------------------------------------------------------------------------------------
# roofline
# ----------------------------------------------------------------------------------------------------------------------
def load_peaks():
    """(burst bf16 TFLOP/s, sustained bf16 TFLOP/s, HBM GB/s, 'measured' | 'fallback')."""
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as f:
            peaks = json.load(f)
        burst = peaks.get('bf16_tflops')
        return burst, peaks.get('bf16_tflops_sustained', burst), peaks.get('hbm_gbs'), 'measured'
    return 1590.0, 1400.0, 6650.0, 'fallback'   # B200_PROFILING.md: 1.59 PFLOP/s burst, ~1.4 sustained, 6.65 TB/s


def load_traffic(workload, q):
    """DRAM bytes per K1 launch from the committed `ncu --set full` capture of the same workload (profiles/), or None."""
    path = os.path.join(ROOT, 'profiles', 'k1_dram_traffic.json')
    if not os.path.exists(path):
        return None, None
    with open(path) as f:
        rec = json.load(f)
    entry = rec.get('%s_q%d' % (workload, q))
    if not entry:
        return None, None
    return entry['dram_bytes_per_launch'], 'profiles/' + entry.get('source', 'k1_dram_traffic.json')


def tensor_roofline(flop_per_launch, k1_ms, launches, timed_ms, workload, q, kernel, clocks, note=None):
    """Roofline object of the K1 launches of a timed region.  The yardstick follows the regime: MEASURED_PEAKS.json's
    sustained figure comes from a 4 s back-to-back cuBLAS loop under the power cap, its burst figure from a best-of-10; a
    timed region shorter than that loop is compared with the burst figure.  Both fractions are printed, and so is the
    fraction of the tensor pipe's own rate at the clock the run saw (8192 dense bf16 FLOP per SM per cycle: what ncu's
    sm__pipe_tensor_cycles_active reports)."""
    burst, sustained, _, kind = load_peaks()
    if not (launches > 0 and k1_ms > 0):
        return None
    achieved = flop_per_launch / (k1_ms / 1e3) / 1e12
    long_run = timed_ms >= 4000.0
    peak = sustained if long_run else burst
    traffic, traffic_source = load_traffic(workload, q)
    sm_mhz = (clocks or {}).get('sm_mhz')
    out = {'bound': 'tensor', 'achieved': achieved, 'peak': peak, 'unit': 'TFLOP/s', 'frac': achieved / peak,
           'traffic': traffic, 'traffic_source': traffic_source, 'kernel': kernel, 'kernel_ms_per_launch': k1_ms,
           'launches_timed': launches, 'kernel_share_of_step': k1_ms * launches / timed_ms,
           'algorithmic_flop_per_launch': flop_per_launch,
           'peak_source': '%s bf16 dense, %s (timed region %.2f s)' % (kind, 'sustained' if long_run else 'burst',
                                                                       timed_ms / 1e3),
           'frac_burst': achieved / burst, 'frac_sustained': achieved / sustained,
           'frac_of_pipe_at_run_clock': (achieved * 1e12 / (148 * 8192 * sm_mhz * 1e6)) if sm_mhz else None,
           'passes': 1}
    if note:
        out['note'] = note
    return out


# ----------------------------------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------------------------------
class Ctx:
    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get('WORLD_SIZE', '1'))
        self.rank = int(os.environ.get('RANK', '0'))
        self.local_rank = int(os.environ.get('LOCAL_RANK', '0'))
        torch.cuda.set_device(self.local_rank)
        if self.world > 1:
            dist.init_process_group('nccl', device_id=torch.device('cuda', self.local_rank))

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def reduce(self, value, op='max'):
        if self.world == 1:
            return float(value)
        t = self.torch.tensor([float(value)], dtype=self.torch.float64, device='cuda')
        self.dist.all_reduce(t, op={'max': self.dist.ReduceOp.MAX, 'min': self.dist.ReduceOp.MIN,
                                    'sum': self.dist.ReduceOp.SUM}[op])
        return float(t.item())

    def close(self):
        if self.world > 1:
            self.dist.destroy_process_group()


def timed(ctx, step_fn, steps, index, use_events):
    """EXACTLY `steps` calls between barrier + synchronize on both sides; device time by CUDA events (or the host clock
    for host-buffer calls), max over ranks; K1 time and launch count of this rank from the index's own events."""
    torch = ctx.torch
    ctx.barrier()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    filt_ms, filt_launches = 0.0, 0
    for _ in range(steps):
        step_fn()
        st = index.stats()
        filt_ms += st['last_filter_ms']
        filt_launches += st['last_filter_launches']
    e1.record()
    ctx.barrier()
    ms = e0.elapsed_time(e1) if use_events else (time.perf_counter() - t0) * 1e3
    return ctx.reduce(ms, 'max'), filt_ms, filt_launches


def query_leg(ctx, args, index, queries, steps, warmup, sampler=None):
    """Resident and end-to-end timing of one block of queries per rank on a fitted index (osml_knn_predict_device /
    osml_knn_predict)."""
    torch = ctx.torch
    from osml_b200 import _native
    q = queries.shape[0]
    stream = torch.cuda.current_stream().cuda_stream
    d_queries = torch.from_numpy(queries).cuda()
    d_pred = torch.empty(q, dtype=torch.float64, device='cuda')
    pinned_q = torch.from_numpy(queries).pin_memory().numpy()

    def resident_step():
        index.predict_device(d_queries.data_ptr(), q, _native.CLASSIFICATION, d_pred.data_ptr(), None, None, stream)

    def e2e_step():
        return index.predict(pinned_q, _native.CLASSIFICATION)

    for _ in range(max(warmup, 3)):
        resident_step()
    torch.cuda.synchronize()
    first_pred = d_pred.cpu().numpy().copy()
    if sampler is not None:
        sampler.start()
    index.reset_stats()
    total_ms, filt_ms, filt_launches = timed(ctx, resident_step, steps, index, True)
    stats = index.stats()
    clocks = sampler.stop() if sampler is not None else None
    for _ in range(2):
        e2e_step()
    e2e_ms, _, _ = timed(ctx, e2e_step, steps, index, False)
    assert np.array_equal(e2e_step(), first_pred), 'host-buffer and device-buffer paths disagree'
    return {'total_ms': total_ms, 'e2e_ms': e2e_ms, 'filt_ms': filt_ms, 'filt_launches': filt_launches,
            'stats': stats, 'clocks': clocks, 'pred': first_pred}


def oracle_sample_check(train, labels, queries, k, pred, sample):
    from oracle import knn_oracle as ko
    want = ko.predict_classification(train, labels, queries[:sample], k=k, tie_rule='reference')
    return bool(np.array_equal(want, pred[:sample]))


def run_ours(args):
    from osml_b200 import _native
    ctx = Ctx()
    torch = ctx.torch
    world, rank, n_gpus = ctx.world, ctx.rank, ctx.world
    train, labels, queries, k = make_data(args.workload, rank)
    n, d = train.shape
    q = queries.shape[0]
    index = _native.Index(train, labels, k, 1, 1e-8, device=ctx.local_rank)      # = KNearestNeighbors._fit
    index.set_option('path', _native.PATHS[args.path])
    if args.no_pipeline:
        index.set_option('pipeline', 0)

    leg = query_leg(ctx, args, index, queries, args.steps, args.warmup, ClockSampler(ctx.local_rank) if rank == 0 else None)
    total_ms, e2e_ms, first_pred = leg['total_ms'], leg['e2e_ms'], leg['pred']
    value = q * n_gpus * args.steps / (total_ms / 1e3)
    e2e_value = q * n_gpus * args.steps / (e2e_ms / 1e3)

    # every rank checks a sample of ITS queries against the numpy oracle
    oracle_n = 8 if world > 1 else 0      # N = 1 does it through the cpu_baseline sample below
    ranks_ok = None
    if oracle_n:
        ok = oracle_sample_check(train, labels, queries, k, first_pred, oracle_n)
        ranks_ok = int(ctx.reduce(1.0 if ok else 0.0, 'sum'))

    strong = c5 = None
    if world > 1 and args.workload == 'c4' and not args.headline_only:
        strong = strong_leg(ctx, args, index, k, total_ms / args.steps)
    dataframe_api = None
    if world == 1:
        dataframe_api = dataframe_leg(ctx, args, index, train, labels, queries, k, first_pred)
    index.close()
    del index
    if world > 1 and args.workload == 'c4' and not args.headline_only:
        c5 = c5_leg(ctx, args)
    if rank != 0:
        ctx.close()
        return 0

    kernel = 'knn_tc_filter_pair_kernel<16, resident>' if d <= 192 else 'knn_tc_filter_pair_kernel<16, wide>'
    roofline = tensor_roofline(2.0 * n * d * q * args.steps / max(leg['filt_launches'], 1),
                               leg['filt_ms'] / max(leg['filt_launches'], 1), leg['filt_launches'], total_ms,
                               args.workload, q, kernel, leg['clocks'],
                               note='single bf16 pass, K = %d (the feature dimension padded to 16; row norms are applied '
                                    'in the epilogue, not as extra K columns)' % ((d + 15) // 16 * 16))
    if roofline is None:   # direct path only: CUDA-core bound; report against the HBM roofline for completeness
        roofline = {'bound': 'hbm', 'achieved': None, 'peak': load_peaks()[2], 'unit': 'GB/s', 'frac': None,
                    'traffic': None, 'note': 'direct (K2) path is CUDA-core bound: 3 un-fusable instructions per element'}

    if n_gpus == 1:
        cpu_value, cpu_sample, cpu_pred, kind = cpu_baseline_single_core(train, labels, queries, k)
        parity_ok = bool(np.array_equal(cpu_pred, first_pred[:cpu_sample]))
        cpu_baseline = {'value': cpu_value, 'unit': 'queries/s', 'cores': 1, 'kind': kind,
                        'sample': '%d of the same queries through %s (the reference is single-threaded)'
                                  % (cpu_sample, _cpu_description(kind)),
                        'matches_gpu_predictions': parity_ok}
    else:
        cpu_baseline = {'value': None, 'unit': 'queries/s', 'cores': 1, 'kind': 'reference',
                        'sample': 'timed at N = 1 only (a per-host figure, it does not change with the GPU count)'}

    stats = leg['stats']
    line = {'metric': METRIC, 'value': value, 'unit': 'queries/s', 'n_gpus': n_gpus, 'steps': args.steps,
            'warmup': max(args.warmup, 3), 'ms_per_step': total_ms / args.steps, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': workload_description(args.workload, n_gpus),
            'e2e': {'value': e2e_value, 'unit': 'queries/s', 'h2d_bytes_per_step': int(queries.nbytes) * n_gpus,
                    'd2h_bytes_per_step': int(q * 8) * n_gpus, 'ms_per_step': e2e_ms / args.steps,
                    'api': 'osml_knn_predict (C ABI, pinned host buffers)',
                    'upload': 'whole matrix first' if args.no_pipeline else
                              'two blocks, the second under the first block\'s filter kernel',
                    'dataframe_api': dataframe_api},
            'gpu_launches': int(stats['kernel_launches']), 'roofline': roofline,
            'cpu_baseline': cpu_baseline,
            'clocks': leg['clocks'], 'retry_queries_per_step': stats['retry_queries'] / args.steps,
            'fallback_queries_per_step': stats['fallback_queries'] / args.steps, 'path': args.path}
    if ranks_ok is not None:
        line['oracle_check'] = {'queries_per_rank': oracle_n, 'ranks_checked': world, 'ranks_ok': ranks_ok,
                                'what': 'each rank: its first queries through the numpy oracle port, labels bit-exact'}
    if strong is not None:
        line['strong'] = strong
    if c5 is not None:
        line['c5'] = c5
    print(json.dumps(line))
    ctx.close()
    return 0


def dataframe_leg(ctx, args, index, train, labels, queries, k, first_pred):
    """The call a user of the reference makes: model.predict(DataFrame) -> Series (pageable F-ordered frame values,
    validation and Series wrapping included); its own fitted model, same data."""
    import pandas as pd
    import osml_b200 as ob
    torch = ctx.torch
    q = queries.shape[0]
    model = ob.KNearestNeighborsClassification(k=k)
    model.path = args.path
    model.fit(pd.DataFrame(train), pd.Series(labels, name='label'))
    if args.no_pipeline:
        model._index.set_option('pipeline', 0)
    query_df = pd.DataFrame(queries)
    for _ in range(2):
        series = model.predict(query_df)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        series = model.predict(query_df)
    df_ms = (time.perf_counter() - t0) * 1e3
    assert np.array_equal(series.values.astype(np.float64), first_pred), 'model API and C ABI disagree'
    model._drop_index()
    return {'value': q * args.steps / (df_ms / 1e3), 'unit': 'queries/s', 'ms_per_step': df_ms / args.steps,
            'api': 'osml_b200.KNearestNeighborsClassification.predict(DataFrame) -> Series'}


def strong_leg(ctx, args, index, k, weak_ms_per_step):
    """BASELINE.json configs[3] as written: the same 100,000 queries split across the ranks (no collective)."""
    from osml_b200 import sharded
    world, rank = ctx.world, ctx.rank
    _, _, all_queries, _ = make_data(args.workload, 0)                    # rank 0's queries, the same on every rank
    q_total = all_queries.shape[0]
    lo, hi = sharded.split_bounds(q_total, world)[rank]
    mine = np.ascontiguousarray(all_queries[lo:hi])
    leg = query_leg(ctx, args, index, mine, args.steps, args.warmup)
    k1_ms = ctx.reduce(leg['filt_ms'] / args.steps, 'max')
    value = q_total * args.steps / (leg['total_ms'] / 1e3)
    ms = leg['total_ms'] / args.steps
    return {'value': value, 'unit': 'queries/s', 'scaling': 'strong', 'queries_total': q_total,
            'queries_per_gpu': int(hi - lo), 'ms_per_step': ms,
            'e2e': {'value': q_total * args.steps / (leg['e2e_ms'] / 1e3), 'unit': 'queries/s',
                    'ms_per_step': leg['e2e_ms'] / args.steps, 'h2d_bytes_per_step': int(all_queries.nbytes),
                    'd2h_bytes_per_step': int(q_total * 8)},
            'k1_ms_per_step_max_over_ranks': k1_ms,
            'efficiency': weak_ms_per_step / (world * ms),
            'efficiency_definition': 'one GPU answering all %d queries (this run\'s own per-GPU step of the headline leg, '
                                     '%.3f ms) / (N x the N-GPU step)' % (q_total, weak_ms_per_step)}


def c5_parity(ctx, ix, d_queries, labels, k, sample, path):
    """Independent check of the tensor path + NCCL exchange: every rank's exact direct-kernel partial lists of `sample`
    queries are all-gathered and merged on the host; indices, distances and labels must match bit for bit."""
    from oracle import knn_oracle as ko
    from osml_b200 import _native
    torch, dist = ctx.torch, ctx.dist
    sub = d_queries[:sample].contiguous()
    pred, idx, dst = ix.predict(sub, _native.CLASSIFICATION, keep_on_device=True, return_neighbours=True)
    pred, idx, dst = pred.cpu().numpy(), idx.cpu().numpy(), dst.cpu().numpy()
    index = ix.backend.index
    index.set_option('path', _native.PATHS['direct'])
    p_idx, p_dst = ix.backend.neighbours(sub)                            # K2 on this rank's rows
    index.set_option('path', _native.PATHS[path])
    if ctx.world > 1:
        all_i = [torch.empty_like(p_idx) for _ in range(ctx.world)]
        all_d = [torch.empty_like(p_dst) for _ in range(ctx.world)]
        dist.all_gather(all_i, p_idx)
        dist.all_gather(all_d, p_dst)
    else:
        all_i, all_d = [p_idx], [p_dst]
    flat_i = np.concatenate([t.cpu().numpy() for t in all_i], axis=1)     # [sample][G * (k+1)]
    flat_d = np.concatenate([t.cpu().numpy() for t in all_d], axis=1)
    ksel = k + 1
    want_i = np.empty((sample, ksel), np.int64)
    want_d = np.empty((sample, ksel), flat_d.dtype)
    for i in range(sample):
        order = np.lexsort((flat_i[i], flat_d[i]))[:ksel]
        want_i[i], want_d[i] = flat_i[i][order], flat_d[i][order]
    want_pred = ko.predict_from_neighbours(want_i, want_d, labels, 'classification')
    ok = np.array_equal(idx, want_i) and np.array_equal(dst, want_d) and np.array_equal(pred, want_pred)
    return bool(ctx.reduce(1.0 if ok else 0.0, 'min') == 1.0)


def c5_leg(ctx, args):
    """Train-sharded mode: rows split across ranks, ONE NCCL all-to-all of the exact partial
    (k+1)-lists per step (each rank receives its slice of the queries), slice-local merge, all-gather of predictions."""
    from osml_b200 import _native, sharded
    torch = ctx.torch
    world, rank = ctx.world, ctx.rank
    rows, d, q, k, seed = C5['rows_per_gpu'], C5['d'], C5['q'], C5['k'], C5['seed']
    n_total = rows * world
    shard = np.asfortranarray(np.random.default_rng(seed + 100 * rank).standard_normal((rows, d), dtype=np.float32))
    labels = np.random.default_rng(seed).integers(0, 10, n_total)
    queries = np.random.default_rng(seed + 1).standard_normal((q, d), dtype=np.float32)     # the same on every rank
    ix = sharded.TrainShardedIndex(shard, labels, k, device=ctx.local_rank, shard_of=(rank * rows, n_total),
                                   path=args.path, exchange=args.exchange)
    index = ix.backend.index
    d_queries = torch.from_numpy(queries).cuda()
    pinned_q = torch.from_numpy(queries).pin_memory().numpy()

    def resident_step():
        return ix.predict(d_queries, _native.CLASSIFICATION, keep_on_device=True)

    def e2e_step():
        return ix.predict(pinned_q, _native.CLASSIFICATION)

    for _ in range(max(args.warmup, 3)):
        first_pred = resident_step()
    torch.cuda.synchronize()
    first_pred = first_pred.cpu().numpy().copy()
    sampler = ClockSampler(ctx.local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    index.reset_stats()
    total_ms, filt_ms, filt_launches = timed(ctx, resident_step, args.steps, index, True)
    stats = index.stats()
    clocks = sampler.stop() if sampler else None
    k1_ms_max = ctx.reduce(filt_ms / args.steps, 'max')
    for _ in range(2):
        e2e_step()
    e2e_ms, _, _ = timed(ctx, e2e_step, args.steps, index, False)
    assert np.array_equal(e2e_step(), first_pred), 'host-buffer and device-buffer paths disagree'
    matches = c5_parity(ctx, ix, d_queries, labels, k, 64, args.path)
    cpu = None
    if rank == 0 and world == 1:
        from oracle import knn_oracle as ko
        sample = 4
        t0 = time.perf_counter()
        cpu_pred = ko.predict_classification(shard, labels[:rows], queries[:sample], k=k, tie_rule='reference')
        cpu_dt = time.perf_counter() - t0
        cpu = {'value': sample / cpu_dt, 'unit': 'queries/s', 'cores': 1, 'kind': 'port',
               'sample': '%d queries against the %d-row shard through the numpy oracle port' % (sample, rows),
               'matches_gpu_predictions': bool(np.array_equal(cpu_pred, first_pred[:sample]))}
    ix.close()
    if rank != 0:
        return None
    k1_ms = filt_ms / max(filt_launches, 1)
    roofline = tensor_roofline(2.0 * rows * d * q * args.steps / max(filt_launches, 1), k1_ms, filt_launches, total_ms,
                               'c5', q, 'knn_tc_filter_pair_kernel<32>', clocks,
                               note='per GPU: its 2,500,000-row shard x all %d queries' % q)
    exchange = ('one NCCL all-to-all of %d bytes per rank per step (16-byte records, each rank receives its slice of '
                '%d queries from every peer) + all-gather of %d predictions per rank'
                % (-(-q // world) * world * (k + 1) * 16, -(-q // world), -(-q // world))) if args.exchange == 'alltoall' \
        else 'two NCCL all-gathers (%d bytes per rank per step), every rank merges all queries' % (q * (k + 1) * 12)
    return {'metric': 'kNN predict queries/sec (k=32), train-sharded', 'value': q * args.steps / (total_ms / 1e3),
            'unit': 'queries/s', 'n_gpus': world, 'steps': args.steps, 'ms_per_step': total_ms / args.steps,
            'scaling': 'weak (rows per GPU fixed)', 'matches': matches,
            'matches_what': '64 queries: host np.lexsort merge of every rank\'s direct-kernel (K2) partial lists vs the '
                            'GPU-merged tensor-path result; indices, distances, labels bit for bit',
            'config': {'workload': 'C5 shape: synthetic gaussian %d x %d fp32 train rows, %d per GPU (train-sharded x%d), '
                                   '%d fp32 queries per step on every GPU (subset of the config\'s 1M), k=%d, '
                                   'KNearestNeighborsClassification' % (n_total, d, rows, world, q, k),
                       'train_rows': n_total, 'features': d, 'queries_per_step': q, 'k': k,
                       'parallelism': 'train-sharded x%d, %s' % (world, exchange),
                       'l2': 'inputs larger than L2: %.0f MB bf16 filter copy per GPU vs 126 MB L2' % (rows * 128 * 2 / 1e6)},
            'e2e': {'value': q * args.steps / (e2e_ms / 1e3), 'unit': 'queries/s',
                    'h2d_bytes_per_step': int(queries.nbytes) * world, 'd2h_bytes_per_step': int(q * 8) * world,
                    'ms_per_step': e2e_ms / args.steps,
                    'api': 'osml_b200.sharded.TrainShardedIndex.predict (host queries -> neighbours -> exchange -> merge)'},
            'gpu_launches': int(stats['kernel_launches']), 'roofline': roofline,
            'k1_ms_per_step_max_over_ranks': k1_ms_max, 'non_k1_ms_per_step': total_ms / args.steps - k1_ms_max,
            'cpu_baseline': cpu, 'clocks': clocks, 'retry_queries_per_step': stats['retry_queries'] / args.steps,
            'fallback_queries_per_step': stats['fallback_queries'] / args.steps, 'path': args.path}


def run_ours_c5(args):
    ctx = Ctx()
    rec = c5_leg(ctx, args)
    if ctx.rank == 0:
        rec.update({'warmup': max(args.warmup, 3), 'higher_is_better': True, 'vs_baseline': None, 'dtype': 'f32',
                    'data': 'synthetic', 'scaling': 'weak'})
        print(json.dumps(rec))
    ctx.close()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='c4', choices=sorted(WORKLOADS) + ['c5'])
    ap.add_argument('--path', default='auto', choices=['auto', 'direct', 'tensor'])
    ap.add_argument('--exchange', default='alltoall', choices=['alltoall', 'allgather'],
                    help='train-sharded leg: slice-wise all-to-all (default) or the first round\'s all-gather of everything')
    ap.add_argument('--headline-only', action='store_true', help='N > 1: skip the strong-scaling and train-sharded legs')
    ap.add_argument('--no-pipeline', action='store_true',
                    help='A/B: osml_knn_predict uploads the whole query matrix before the first kernel')
    args = ap.parse_args()
    if args.workload == 'c5':
        if args.impl == 'reference':
            print(json.dumps({'impl': 'reference', 'unavailable': 'the reference arm is timed on the headline '
                                                                  'workload (c4) only'}))
            return 0
        return run_ours_c5(args)
    if args.impl == 'reference':
        return run_reference_arm(args)
    return run_ours(args)


if __name__ == '__main__':
    sys.exit(main())
