#!/usr/bin/env python3
'''GPU check of the tensor-core preconditioner path (TFB_PREC_TENSOR) against the fp64 path, and timing of
Newton steps with both.  Run on a B200:  python tools/check_tc_precond.py [grid ...]'''
import ctypes
import os
import sys
import time

import numpy

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from transiflow_b200 import Interface, _lib  # noqa: E402
from transiflow_b200._lib import check, ptr  # noqa: E402


def precond(it, jac, r, flags):
    o = _lib.TfbSolveOpts()
    o.pressure_row = it.pressure_row
    o.precond_flags = flags
    z = numpy.empty_like(r)
    check(_lib.lib().tfb_precond_apply_opts(jac._h, ptr(r), ptr(z), ctypes.byref(o)))
    return z


def case(nx, ny, nz, params, newton=3):
    it = Interface(dict(params), nx, ny, nz)
    it.AUTO_IDR_MIN_UNKNOWNS = 1000
    x = numpy.random.default_rng(0).uniform(-0.1, 0.1, it.n)
    jac, f = it.jacobian_rhs(x)
    it._sync_solver()
    r = numpy.random.default_rng(1).uniform(-1, 1, it.n)
    base = _lib.PREC_SCALED_MASS | _lib.PREC_NO_JOINT
    z64 = precond(it, jac, r, base)
    ztc = precond(it, jac, r, base | _lib.PREC_TENSOR)
    err = numpy.abs(ztc - z64).max() / numpy.abs(z64).max()
    parts = {v: numpy.abs(ztc[v::it.dof] - z64[v::it.dof]).max() / max(numpy.abs(z64[v::it.dof]).max(), 1e-300) for v in range(it.dof)}
    print('%dx%dx%d: precond tensor vs fp64 max rel diff %.3e  per var %s' % (nx, ny, nz, err, {k: '%.1e' % v for k, v in parts.items()}), flush=True)
    ok = err < 2e-4
    for prec in ('double', 'tf32x3'):
        it.parameters['Iterative Solver'] = {'Preconditioner Precision': prec}
        xs = it.vector()
        for k in range(newton):
            t0 = time.perf_counter()
            jac, f = it.jacobian_rhs(xs)
            dx = it.solve(jac, -f)
            dt = time.perf_counter() - t0
            xs = xs + dx
            ls = it.last_solve
            print('   %-7s newton %d: |F| %.3e  %s/%s its %d relres %.2e solve %.1f ms step %.1f ms conv %s' % (
                prec, k, numpy.linalg.norm(f), ls['method'], ls['schur'], ls['iterations'], ls['relres'], ls['solve_ms'], 1e3 * dt,
                ls['converged']), flush=True)
            ok = ok and ls['converged']
    return ok


if __name__ == '__main__':
    grids = [int(a) for a in sys.argv[1:]] or [0]
    ok = True
    if 0 in grids:
        ok &= case(32, 24, 20, {'Reynolds Number': 100})
        ok &= case(40, 48, 16, {'Reynolds Number': 50, 'Grid Stretching Factor': 1.5})
    for g in grids:
        if g:
            ok &= case(g, g, g, {'Reynolds Number': 100}, newton=4)
    print('TC CHECK', 'OK' if ok else 'FAILED')
    sys.exit(0 if ok else 1)
