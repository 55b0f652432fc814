'''Builds libtfb200.so for sm_100a IN-TREE (transiflow_b200/lib/), so that the binary travels
with the repo snapshot to the GPU box.  nvcc cross-compiles without a GPU.

    python -m transiflow_b200.build [--force]
'''
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIBDIR = os.path.join(HERE, 'lib')
SO = os.path.join(LIBDIR, 'libtfb200.so')

ARCH = ['-gencode', 'arch=compute_100a,code=sm_100a']
COMMON = ['-lineinfo', '-O3', '-std=c++17', '-Xcompiler', '-fPIC', '-Xcompiler', '-O2']
# (source, extra flags).  The assembly kernels must not contract a*b+c into FMA: bit-parity
# with the reference's numpy arithmetic depends on it.
UNITS = [
    ('tfb_core.cu', ['-fmad=false'] + (['-DTFB_ASM_EXPERIMENTS'] if os.environ.get('TFB_ASM_EXPERIMENTS') else [])),
    ('tfb_solver.cu', []),
    ('tfb_comm.cu', []),
    ('tfb_direct.cu', []),
]


def _deps():
    out = []
    for root, _, files in os.walk(CSRC):
        out += [os.path.join(root, f) for f in files if f.endswith(('.h', '.cuh', '.cu'))]
    out.append(os.path.join(os.path.dirname(HERE), 'include', 'tfb200.h'))
    return out


def _nvcc():
    for cand in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc'):
        if cand and os.path.exists(cand):
            return cand
    return 'nvcc'


def build(force=False, verbose=False):
    os.makedirs(LIBDIR, exist_ok=True)
    newest = max(os.path.getmtime(p) for p in _deps())
    if not force and os.path.exists(SO) and os.path.getmtime(SO) >= newest:
        return SO
    nvcc = _nvcc()
    env = dict(os.environ)
    env.pop('CC', None)
    env.pop('CXX', None)

    def compile_unit(unit):
        src, extra = unit
        obj = os.path.join(LIBDIR, src.replace('.cu', '.o'))
        cmd = [nvcc] + ARCH + COMMON + extra + ['-ccbin', '/usr/bin/g++', '-c', os.path.join(CSRC, src), '-o', obj]
        if verbose:
            print(' '.join(cmd), flush=True)
        subprocess.check_call(cmd, env=env)
        return obj

    with ThreadPoolExecutor(max_workers=len(UNITS)) as ex:
        objs = list(ex.map(compile_unit, UNITS))
    cmd = [nvcc] + ARCH + ['-shared', '-ccbin', '/usr/bin/g++', '-o', SO] + objs + ['-ldl']
    if verbose:
        print(' '.join(cmd), flush=True)
    subprocess.check_call(cmd, env=env)
    return SO


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose=True))
