'''Converged Newton states of two 3-D cavities that are too large for the Python reference's assembly in a test run but
still feasible for SuperLU: generated with the pinned oracle (oracle/tf_oracle.py: bit-identical assembly, direct_solve =
the SciPy backend's pinned spsolve path).  Build container only (3 minutes).

    python tests/golden/make_golden_newton_oracle.py
'''
import os
import sys

import numpy

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.tf_oracle import Oracle, direct_solve  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'generated')
CASES = {'ldc3d_20_re100': ({'Reynolds Number': 100}, 20),
         'ldc3d_16_re400_str': ({'Reynolds Number': 400, 'Grid Stretching Factor': 1.5}, 16)}


def main():
    for name, (params, N) in CASES.items():
        orc = Oracle(dict(params), N, N, N)
        x = numpy.zeros(orc.n)
        for k in range(10):
            f = orc.rhs(x)
            if numpy.linalg.norm(f) < 1e-12:
                break
            x = x + direct_solve(orc.jacobian_csr(x), -f, orc.dim, orc.dof)
        print(name, k, numpy.linalg.norm(orc.rhs(x)))
        numpy.savez_compressed(os.path.join(OUT, 'newton_oracle_' + name + '.npz'), x=x, N=N, steps=k)


if __name__ == '__main__':
    main()
