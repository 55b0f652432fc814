'''Generates tests/golden/generated/newton_*.npz: converged Newton states and single Newton
updates computed by the UNMODIFIED reference (SciPy backend: Discretization + SuperLU via
Continuation.newton), build container only.

    python tests/golden/make_golden_newton.py
'''
import contextlib
import io
import os
import sys

import numpy

sys.path.insert(0, '/root/reference')
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from transiflow import Continuation, Interface  # noqa: E402  (the reference)
from cases_newton import NEWTON_CASES  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'generated')


def main():
    for name, (params, nx, ny, nz) in NEWTON_CASES.items():
        it = Interface(dict(params), nx, ny, nz)
        cont = Continuation(it)
        x0 = it.vector()
        with contextlib.redirect_stdout(io.StringIO()):
            x = cont.newton(x0, 1e-10)
        # one more linear solve at the converged state with a generic right-hand side
        jac = it.jacobian(x)
        b = numpy.random.default_rng(7).uniform(-1, 1, x.size)
        y = it.solve(jac, b)
        numpy.savez_compressed(os.path.join(OUT, 'newton_' + name + '.npz'), x=x, b=b, y=y,
                               fnorm=numpy.linalg.norm(it.rhs(x)), iterations=cont.newton_iterations)
        print('%-14s n=%6d newton its=%d |F|=%.2e' % (name, x.size, cont.newton_iterations, numpy.linalg.norm(it.rhs(x))))


if __name__ == '__main__':
    main()
