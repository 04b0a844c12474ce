"""Builds osml_b200/libosml_knn.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m osml_b200.build [--force] [--verbose]
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
OUT = os.path.join(HERE, 'libosml_knn.so')
SOURCES = ['knn_api.cu']

NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '-Xcompiler', '-fPIC', '-Xcompiler', '-pthread', '-shared']


def _nvcc():
    for cand in (shutil.which('nvcc'), '/usr/local/cuda/bin/nvcc'):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError('nvcc not found')


def _stale():
    if not os.path.exists(OUT):
        return True
    built = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    deps.append(os.path.join(os.path.dirname(HERE), 'include', 'osml_knn.h'))
    return any(os.path.getmtime(p) > built for p in deps)


def build(force=False, verbose=False, out=OUT, defines=()):
    """Compiles every CUDA source of the package into libosml_knn.so; returns the path.
    `out` / `defines` build experimental variants for A/B timing (tools/ab.sh); the product is the default build."""
    if not force and out == OUT and not _stale():
        return OUT
    cmd = [_nvcc()] + NVCC_FLAGS + (['-Xptxas', '-v'] if verbose else []) + ['-D' + d for d in defines]
    cmd += [os.path.join(CSRC, s) for s in SOURCES] + ['-o', out]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout)
    if res.returncode != 0:
        raise RuntimeError('nvcc failed (%d): %s' % (res.returncode, ' '.join(cmd)))
    return out


if __name__ == '__main__':
    _defs = [a[2:] for a in sys.argv if a.startswith('-D')]
    _out = [a[6:] for a in sys.argv if a.startswith('--out=')]
    print(build(force='--force' in sys.argv, verbose='--verbose' in sys.argv, out=_out[0] if _out else OUT, defines=_defs))
