#!/usr/bin/env python3
'''One or two Newton steps of the 3-D cavity at grid^3 (default 128) with the default solver; used under ncu for
launch lists:  python tools/solve_probe.py [grid] [steps] [problem]'''
import os
import sys
import time
import numpy
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from transiflow_b200 import Interface  # noqa: E402

grid = int(sys.argv[1]) if len(sys.argv) > 1 else 128
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
problem = sys.argv[3] if len(sys.argv) > 3 else 'ldc'
params = {'Reynolds Number': 100} if problem == 'ldc' else \
    {'Problem Type': 'Rayleigh-Benard', 'Rayleigh Number': 1000.0, 'Prandtl Number': 10.0, 'Biot Number': 1.0, 'X-max': 10, 'Y-max': 10}
if os.environ.get('TFB_MAXIT'):
    params['Iterative Solver'] = {'Maximum Iterations': int(os.environ['TFB_MAXIT'])}
it = Interface(params, grid, grid, grid)
x = it.vector()
import warnings
warnings.simplefilter('ignore')
for k in range(steps):
    t0 = time.perf_counter()
    jac, f = it.jacobian_rhs(x)
    dx = it.solve(jac, -f)
    x = x + dx
    ls = it.last_solve
    print('step %d |F| %.3e %s its %d solve %.1f ms step %.1f ms' % (k, numpy.linalg.norm(f), ls['method'], ls['iterations'],
                                                                   ls['solve_ms'], 1e3 * (time.perf_counter() - t0)), flush=True)
