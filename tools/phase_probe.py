#!/usr/bin/env python3
'''Single-GPU per-phase breakdown of the IDR solve (tfb_solve, Verbose): python tools/phase_probe.py [grid]'''
import os
import sys
import warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from transiflow_b200 import Interface  # noqa: E402
grid = int(sys.argv[1]) if len(sys.argv) > 1 else 128
params = {'Reynolds Number': 100}
it = Interface(params, grid, grid, grid)
warnings.simplefilter('ignore')
x = it.vector()
for k in range(3):
    if k == 2:
        params['Verbose'] = True
    jac, f = it.jacobian_rhs(x)
    x = x + it.solve(jac, -f)
    print('step %d: %s %d products, solve %.1f ms' % (k, it.last_solve['method'], it.last_solve['iterations'], it.last_solve['solve_ms']), flush=True)
