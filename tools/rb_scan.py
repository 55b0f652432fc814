'''Measurement helper (run on the GPU box): Rayleigh-Benard Newton-update solves at the conduction state for a
list of grids and Rayleigh numbers; prints Krylov iterations, inner iterations and device time as JSON lines.

    python tools/rb_scan.py 64x64x32:1000,3000 128x128x128:1000
'''
import json
import sys
import time

import numpy

sys.path.insert(0, __file__.rsplit('/', 2)[0])
from transiflow_b200 import Interface  # noqa: E402


def conduction_state(it):
    x = numpy.zeros(it.n)
    dx = it.solve(it.jacobian(x), -it.rhs(x))     # linear at u = 0
    return x + dx, dict(it.last_solve)


def main():
    for spec in sys.argv[1:]:
        grid, ras = spec.split(':')
        nx, ny, nz = (int(a) for a in grid.split('x'))
        for Ra in (float(a) for a in ras.split(',')):
            for coupling in ('joint', 'none'):
                params = {'Problem Type': 'Rayleigh-Benard', 'Rayleigh Number': Ra, 'Prandtl Number': 10.0,
                          'Biot Number': 1.0, 'X-max': 10, 'Y-max': 10,
                          'Iterative Solver': {'Scalar Coupling': coupling, 'Maximum Iterations': 400, 'Restart': 400}}
                it = Interface(params, nx, ny, nz)
                x, first = conduction_state(it)
                b = numpy.random.default_rng(0).standard_normal(it.n)
                b[3] = 0
                jac = it.jacobian(x)
                t0 = time.time()
                y = it.solve(jac, b)
                wall = time.time() - t0
                r = jac @ y - b
                r[3] = 0
                print(json.dumps({'grid': grid, 'Ra': Ra, 'coupling': coupling, 'conduction_solve': first,
                                  'solve': it.last_solve, 'wall_s': round(wall, 3),
                                  'true_relres': float(numpy.linalg.norm(r) / numpy.linalg.norm(b))}), flush=True)
                del it


if __name__ == '__main__':
    main()
