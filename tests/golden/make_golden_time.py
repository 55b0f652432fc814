'''Generates tests/golden/generated/time_ldc2d.npz with the UNMODIFIED reference: implicit Euler
(TimeIntegration, theta = 1) of the 2-D lid-driven cavity from rest.  Build container only.

    python tests/golden/make_golden_time.py
'''
import contextlib
import io
import os
import sys

import numpy

sys.path.insert(0, '/root/reference')
from transiflow import Interface, TimeIntegration  # noqa: E402  (the reference)

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'generated')


def main():
    params = {'Problem Type': 'Lid-driven Cavity', 'Reynolds Number': 100, 'Lid Velocity': 1}
    nx = ny = 12
    dt, steps = 0.5, 6
    it = Interface(params, nx, ny)
    ti = TimeIntegration(it, theta=1.0)
    with contextlib.redirect_stdout(io.StringIO()):
        x, t = ti.integration(it.vector(), dt, dt * steps - 1e-9)
    print('t = %g, |x| = %.6f' % (t, numpy.linalg.norm(x)))
    numpy.savez_compressed(os.path.join(OUT, 'time_ldc2d.npz'), x=x, t=t, nx=nx, ny=ny, dt=dt, steps=steps)


if __name__ == '__main__':
    main()
