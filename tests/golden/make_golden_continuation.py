'''Generates tests/golden/generated/continuation_ldc2d.npz with the UNMODIFIED reference:
pseudo-arclength continuation in the Reynolds number (Continuation.continuation, SciPy backend).
Build container only.

    python tests/golden/make_golden_continuation.py
'''
import contextlib
import io
import os
import sys

import numpy

sys.path.insert(0, '/root/reference')
from transiflow import Continuation, Interface  # noqa: E402  (the reference)

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'generated')


def main():
    params = {'Problem Type': 'Lid-driven Cavity', 'Reynolds Number': 0, 'Lid Velocity': 1, 'Grid Stretching Factor': 1.5}
    nx = ny = 16
    it = Interface(params, nx, ny)
    cont = Continuation(it)
    x0 = it.vector()
    with contextlib.redirect_stdout(io.StringIO()):
        x0 = cont.newton(x0)
        x, mu = cont.continuation(x0, 'Reynolds Number', 0, 400, 100)
        x_cont = x.copy()
        it.set_parameter('Reynolds Number', mu)
        x = cont.newton(x, 1e-12)      # the continuation corrector stops at its own tolerance; polish on the branch
    print('reached Re = %g, |F| = %.2e' % (mu, numpy.linalg.norm(it.rhs(x))))
    print('continuation end vs polished state: %.2e' % numpy.abs(x - x_cont).max())
    numpy.savez_compressed(os.path.join(OUT, 'continuation_ldc2d.npz'), x=x, x_continuation=x_cont, mu=mu, nx=nx, ny=ny)


if __name__ == '__main__':
    main()
