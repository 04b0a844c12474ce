"""CPU check of the host-made leaf plan behind the pairwise support-vector kernel (svm_decision.cuh, svm_make_plan): for
every feature count the plan, evaluated the way the device does (8 strided accumulators per leaf of at most 128
elements, a remainder loop, a stack that merges finished subtrees), must add float32 numbers in exactly numpy's
pairwise_sum order -- i.e. give np.sum's bits."""
import os
import shutil
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, 'tests', 'helpers', 'print_pairwise_plans.cu')


def _nvcc():
    for cand in (shutil.which('nvcc'), '/usr/local/cuda/bin/nvcc'):
        if cand and os.path.exists(cand):
            return cand
    return None


def _leaf_sum(x):
    n = len(x)
    if n < 8:
        res = x[0]
        for v in x[1:]:
            res = np.float32(res + v)
        return res
    r = [x[j] for j in range(8)]
    n8 = n - n % 8
    for i in range(8, n8, 8):
        for j in range(8):
            r[j] = np.float32(r[j] + x[i + j])
    res = np.float32(np.float32(np.float32(r[0] + r[1]) + np.float32(r[2] + r[3])) +
                     np.float32(np.float32(r[4] + r[5]) + np.float32(r[6] + r[7])))
    for i in range(n8, n):
        res = np.float32(res + x[i])
    return res


def _plan_sum(x, plan):
    stack = []
    for lo, n, merges in plan:
        stack.append(_leaf_sum(x[lo:lo + n]))
        for _ in range(merges):
            right = stack.pop()
            stack[-1] = np.float32(stack[-1] + right)
    assert len(stack) == 1
    return stack[0]


@pytest.fixture(scope='module')
def plans(tmp_path_factory):
    nvcc = _nvcc()
    if nvcc is None:
        pytest.skip('nvcc not available')
    exe = str(tmp_path_factory.mktemp('plan') / 'print_plans')
    subprocess.run([nvcc, '-gencode', 'arch=compute_100a,code=sm_100a', '-std=c++17', '-O1', SRC, '-o', exe],
                   check=True, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    sizes = sorted(set(list(range(1, 300)) + [383, 384, 385, 511, 512, 513, 784, 1000, 1023, 1024, 1025, 1027, 2047,
                                               2048, 2049, 3000, 4097, 8189]))
    out = subprocess.run([exe] + [str(d) for d in sizes], check=True, stdout=subprocess.PIPE, text=True).stdout
    table = {}
    for line in out.strip().splitlines():
        f = [int(v) for v in line.split()]
        assert f[1] > 0, line
        table[f[0]] = [tuple(f[2 + 3 * l:5 + 3 * l]) for l in range(f[1])]
    assert sorted(table) == sizes
    return table


def test_plans_cover_the_row_with_leaves_numpy_would_use(plans):
    for d, plan in plans.items():
        assert plan[0][0] == 0 and sum(n for _, n, _ in plan) == d
        for (lo, n, _), (lo2, _, _) in zip(plan, plan[1:]):
            assert lo + n == lo2
        assert all(1 <= n <= 128 for _, n, _ in plan)
        assert all(lo % 8 == 0 for lo, _, _ in plan)                 # 16-byte aligned vector loads on the device
        assert sum(m for _, _, m in plan) == len(plan) - 1           # a binary tree over the leaves
        assert len(plan) == 1 or d > 128


def test_plan_evaluation_has_numpy_bits(plans):
    rng = np.random.default_rng(0)
    for d, plan in plans.items():
        for _ in range(3):
            x = (rng.standard_normal(d) * 10 ** rng.uniform(-3, 3, d)).astype(np.float32)
            assert _plan_sum(x, plan) == np.sum(x), d
