"""CPU checks of bench.py's contract that need no GPU: the reference arm (``--impl reference``) prints one JSON line
with the keys the driver reads, and under a multi-rank launch only rank 0 works."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra_env=None):
    env = dict(os.environ)
    for key in ('RANK', 'WORLD_SIZE', 'LOCAL_RANK'):
        env.pop(key, None)
    env.update(extra_env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--workload', 'c3',
                           '--steps', '1', '--warmup', '0'], cwd=ROOT, env=env, stdout=subprocess.PIPE,
                          stderr=subprocess.PIPE, text=True, timeout=600)


def test_reference_arm_prints_the_contract_line():
    res = _run()
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.strip().splitlines() if l.startswith('{')]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line['impl'] == 'reference' and line['metric'].startswith('kNN predict queries/sec')
    assert line['unit'] == 'queries/s' and line['higher_is_better'] is True and line['n_gpus'] == 1
    assert line['steps'] == 1 and line['value'] > 0 and line['ms_per_step'] > 0
    assert 'workload' in line['config'] and 'model' not in line['config']
    base = line['cpu_baseline']
    assert base['kind'] in ('reference', 'port') and base['cores'] >= 1 and base['value'] == line['value']
    assert base['sample']
    e2e = line['e2e']
    assert e2e['value'] == line['value'] and e2e['unit'] == line['unit']
    assert e2e['h2d_bytes_per_step'] == 0 and e2e['d2h_bytes_per_step'] == 0


def test_reference_arm_runs_on_rank_zero_only():
    res = _run({'RANK': '1', 'WORLD_SIZE': '2', 'LOCAL_RANK': '1'})
    assert res.returncode == 0 and res.stdout.strip() == ''
