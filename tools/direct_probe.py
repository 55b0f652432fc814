#!/usr/bin/env python3
'''Direct (block-tridiagonal) solve on the 2-D BASELINE configurations: factorisation / substitution time, residual, and
the answer against the Krylov solver on the small ones.  python tools/direct_probe.py [names...]'''
import os, sys, time, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy
from transiflow_b200 import Interface
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import PROBLEMS_2D
warnings.simplefilter('ignore')
names = sys.argv[1:] or ['ldc2d', 'dhc2d', 'qg', 'amoc']
for name in names:
    params, nx, ny, desc = PROBLEMS_2D[name]
    params = dict(params)
    it = Interface(params, nx, ny, 1)
    x = numpy.random.default_rng(0).uniform(-0.01, 0.01, it.n)
    jac, f = it.jacobian_rhs(x)
    for rep in range(2):
        t0 = time.perf_counter()
        y = it.solve(jac, -f)
        dt = time.perf_counter() - t0
        ls = it.last_solve
        print('%-6s %dx%d n=%d: %s factor %.1f ms solve %.1f ms wall %.1f ms relres %.2e conv %s' % (
            name, nx, ny, it.n, ls['method'], ls['setup_ms'], ls['solve_ms'], 1e3 * dt, ls['relres'], ls['converged']), flush=True)
    if it.n <= 20000:
        params['Iterative Solver'] = {'Method': 'FGMRES', 'Maximum Iterations': 4000, 'Restart': 4000}
        t0 = time.perf_counter()
        yk = it.solve(jac, -f)
        print('       Krylov: %d its %.1f ms, |y_direct - y_krylov| / |y| = %.2e' % (
            it.last_solve['iterations'], 1e3 * (time.perf_counter() - t0), numpy.abs(y - yk).max() / numpy.abs(yk).max()), flush=True)
