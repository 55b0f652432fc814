#!/usr/bin/env python3
'''y = J x timing (tfb_spmv_bench) and a correctness check against scipy on a ragged grid, for the kernel variant selected
by TFB_SPMV_VARIANT:  python tools/spmv_bench.py [ldc|rb] [grid]'''
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy
from transiflow_b200 import Interface, _lib
from transiflow_b200._lib import check
prob = sys.argv[1] if len(sys.argv) > 1 else 'ldc'
grid = int(sys.argv[2]) if len(sys.argv) > 2 else 128
P = {'ldc': {'Reynolds Number': 100},
     'rb': {'Problem Type': 'Rayleigh-Benard', 'Rayleigh Number': 1000.0, 'Prandtl Number': 10.0, 'Biot Number': 1.0, 'X-max': 10, 'Y-max': 10}}[prob]
it = Interface(dict(P), 37, 9, 19)
x = numpy.random.default_rng(0).uniform(-0.5, 0.5, it.n)
jac = it.jacobian(x)
v = numpy.random.default_rng(1).uniform(-1, 1, it.n)
err = numpy.abs(jac @ v - jac.tocsr() @ v).max() / numpy.abs(jac.tocsr() @ v).max()
it = Interface(dict(P), grid, grid, grid)
x = numpy.random.default_rng(0).uniform(-0.5, 0.5, it.n)
jac = it.jacobian(x)
out = {}
for masked in (0, 1):
    ms = ctypes.c_float()
    check(_lib.lib().tfb_spmv_bench(jac._h, 30, masked, ctypes.byref(ms)))
    out[masked] = ms.value
byts = 8 * it.nnz + 4 * (it.n + 1) + 16 * it.n
print('variant %s %s %d^3: ragged-grid rel err %.1e; full %.1f us (%.0f GB/s of %d MB), masked %.1f us' % (
    os.environ.get('TFB_SPMV_VARIANT', '0'), prob, grid, err, 1e3 * out[0], byts / out[0] / 1e6, byts // 1000000, 1e3 * out[1]))
