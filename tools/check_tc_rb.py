#!/usr/bin/env python3
'''GPU check of the tensor-core path for every sub-solve (LSC Poisson solves, scalars, coupled (w,T) solve):
preconditioner application fp64 vs tf32x3, then Newton-like solves.  python tools/check_tc_rb.py [grid]'''
import ctypes
import os
import sys
import time
import warnings

import numpy

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from transiflow_b200 import Interface, _lib  # noqa: E402
from transiflow_b200._lib import check, ptr  # noqa: E402

RB = {'Problem Type': 'Rayleigh-Benard', 'Rayleigh Number': 1000.0, 'Prandtl Number': 10.0, 'Biot Number': 1.0, 'X-max': 10, 'Y-max': 10}
DHC = {'Problem Type': 'Differentially Heated Cavity', 'Rayleigh Number': 1e4, 'Prandtl Number': 1000.0, 'Reynolds Number': 1}


def precond(it, jac, r, flags, inner=0):
    o = _lib.TfbSolveOpts()
    o.pressure_row = it.pressure_row
    o.precond_flags = flags
    o.inner_its = inner
    z = numpy.empty_like(r)
    check(_lib.lib().tfb_precond_apply_opts(jac._h, ptr(r), ptr(z), ctypes.byref(o)))
    return z


def case(name, params, nx, ny, nz, solves=True):
    it = Interface(dict(params), nx, ny, nz)
    x = numpy.random.default_rng(0).uniform(-0.01, 0.01, it.n)
    jac, f = it.jacobian_rhs(x)
    it._sync_solver()
    r = numpy.random.default_rng(1).uniform(-1, 1, it.n)
    ok = True
    for label, base in (('LSC block-triangular', _lib.PREC_NO_JOINT), ('LSC joint (w,T)', 0)):
        if base == 0 and not getattr(it, '_joint', False):
            continue
        z64 = precond(it, jac, r, base)
        ztc = precond(it, jac, r, base | _lib.PREC_TENSOR)
        parts = {v: numpy.abs(ztc[v::it.dof] - z64[v::it.dof]).max() / max(numpy.abs(z64[v::it.dof]).max(), 1e-300) for v in range(it.dof)}
        err = max(parts.values())
        print('%s %dx%dx%d %s: precond tensor vs fp64 per var %s' % (name, nx, ny, nz, label, {k: '%.1e' % v for k, v in parts.items()}), flush=True)
        ok &= err < 1e-3
    if solves:
        b = numpy.random.default_rng(2).uniform(-1, 1, it.n)
        b[it.dim] = 0
        for prec in ('double', 'tf32x3'):
            it.parameters['Iterative Solver'] = {'Preconditioner Precision': prec}
            with warnings.catch_warnings():
                warnings.simplefilter('ignore')
                t0 = time.perf_counter()
                it.solve(jac, b)
                dt = time.perf_counter() - t0
            ls = it.last_solve
            print('   %-7s %s/%s its %d relres %.2e solve %.1f ms wall %.1f ms conv %s' % (
                prec, ls['method'], ls['schur'], ls['iterations'], ls['relres'], ls['solve_ms'], 1e3 * dt, ls['converged']), flush=True)
            ok &= ls['converged']
    return ok


if __name__ == '__main__':
    grid = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    ok = True
    ok &= case('RB', RB, 24, 20, 12)
    ok &= case('DHC', DHC, 16, 20, 12)
    ok &= case('LDC', {'Reynolds Number': 100}, 20, 16, 12)
    if grid:
        ok &= case('RB', RB, grid, grid, grid)
    print('TC RB CHECK', 'OK' if ok else 'FAILED')
    sys.exit(0 if ok else 1)
