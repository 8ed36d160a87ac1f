"""Build libbehavenet_b200.so (sm_100a only) with nvcc.  In-tree output so it ships with gpurun."""

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
LIB = os.path.join(PKG, 'libbehavenet_b200.so')
SOURCES = ['cae_plan.cu', 'cae_simt.cu', 'cae_misc.cu', 'cae_thin.cu', 'cae_thin_tc.cu', 'cae_tc.cu', 'linae.cu', 'psvae.cu', 'arhmm.cu', 'arhmm_tc.cu']
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
         '-Xcompiler', '-fPIC', '--use_fast_math=false']
FLAGS = [f for f in FLAGS if f != '--use_fast_math=false']


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    objdir = os.path.join(HERE, 'build')
    os.makedirs(objdir, exist_ok=True)
    headers = [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith(('.cuh', '.h'))]
    headers.append(os.path.join(os.path.dirname(PKG), 'include', 'behavenet_b200.h'))
    jobs = []
    for src in SOURCES:
        s = os.path.join(HERE, src)
        o = os.path.join(objdir, src.replace('.cu', '.o'))
        if force or _stale(o, [s] + headers):
            cmd = [NVCC] + FLAGS + (['-Xptxas', '-v'] if verbose else []) + ['-c', s, '-o', o]
            jobs.append(cmd)

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError('nvcc failed: %s\n%s\n%s' % (' '.join(cmd), r.stdout, r.stderr))
        return r.stderr

    with ThreadPoolExecutor(max_workers=8) as ex:
        logs = list(ex.map(run, jobs))
    if verbose:
        for l in logs:
            sys.stderr.write(l)
    objs = [os.path.join(objdir, s.replace('.cu', '.o')) for s in SOURCES]
    if force or jobs or _stale(LIB, objs):
        run([NVCC, '-shared', '-o', LIB] + objs + ['-lcudart'])
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
