"""Build libsegland_b200.so in-tree with nvcc for sm_100a (the only target).

    python -m segland_b200.build [--force] [--verbose]

The shared library exports the plain C ABI declared in include/segland_b200.h; it links the
static CUDA runtime, so it has no dependency on torch or on a particular libcudart.so.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, 'csrc')
LIB_DIR = os.path.join(PKG, 'lib')
LIB_PATH = os.path.join(LIB_DIR, 'libsegland_b200.so')
OBJ_DIR = os.path.join(LIB_DIR, 'obj')

NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
              '-Xcompiler', '-fPIC', '-Xcompiler', '-fvisibility=hidden', '--expt-relaxed-constexpr']


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith('.cu'))


def _deps():
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.cuh', '.h'))]
    hdrs.append(os.path.join(ROOT, 'include', 'segland_b200.h'))
    return hdrs


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, ab_variants=None):
    """ab_variants (or SL_AB_VARIANTS=1 in the environment): also compile the superseded kernel generations kept for
    A/B measurements (profiles/scripts/pair_probe.py, pair_dedup_probe.py); the product build leaves them out."""
    nvcc = os.environ.get('NVCC', 'nvcc')
    if ab_variants is None:
        ab_variants = os.environ.get('SL_AB_VARIANTS', '') == '1'
    flags = NVCC_FLAGS + (['-DSL_AB_VARIANTS'] if ab_variants else [])
    stamp = os.path.join(OBJ_DIR, 'ab_variants' if ab_variants else 'product')
    os.makedirs(OBJ_DIR, exist_ok=True)
    if not os.path.exists(stamp):                 # switching flavour: rebuild everything
        force = True
        for f in ('ab_variants', 'product'):
            if os.path.exists(os.path.join(OBJ_DIR, f)):
                os.remove(os.path.join(OBJ_DIR, f))
        open(stamp, 'w').close()
    os.makedirs(OBJ_DIR, exist_ok=True)
    hdrs = _deps()
    jobs = []
    for src in sources():
        obj = os.path.join(OBJ_DIR, os.path.basename(src)[:-3] + '.o')
        if force or _stale(obj, [src] + hdrs):
            cmd = [nvcc] + flags + (['-Xptxas', '-v'] if verbose else []) + ['-c', src, '-o', obj]
            jobs.append(cmd)

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or r.returncode != 0:
            sys.stderr.write(' '.join(cmd) + '\n' + r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError(f'nvcc failed for {cmd[-3]}')

    with ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        list(ex.map(run, jobs))
    objs = [os.path.join(OBJ_DIR, os.path.basename(s)[:-3] + '.o') for s in sources()]
    if force or jobs or _stale(LIB_PATH, objs):
        # -lcuda: the tensor-map encoder (cuTensorMapEncodeTiled) is a driver-API entry point;
        # it is resolved lazily through cudaGetDriverEntryPoint, so no link-time dependency.
        run([nvcc, '-shared', '-gencode', 'arch=compute_100a,code=sm_100a', '-o', LIB_PATH] + objs)
    return LIB_PATH


if __name__ == '__main__':
    path = build(force='--force' in sys.argv, verbose='--verbose' in sys.argv,
                 ab_variants=True if '--ab-variants' in sys.argv else None)
    print(path)
