'''Generates tests/golden/generated/*.npz by importing the UNMODIFIED Python reference from
/root/reference (build container only; the GPU box never runs this).

    python tests/golden/make_golden.py
'''
import os
import sys

import numpy

sys.path.insert(0, '/root/reference')
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from transiflow import Discretization  # noqa: E402  (the reference)
from cases import CASES, CUSTOM_BC_CASES, make_state  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'generated')


def main():
    os.makedirs(OUT, exist_ok=True)
    for name, (params, nx, ny, nz, dim, dof, kind) in CASES.items():
        d = Discretization(dict(params), nx, ny, nz, dim, dof)
        n = nx * ny * nz * d.dof
        state = make_state(kind, n)
        A = d.jacobian(state)
        f = d.rhs(state)
        M = d.mass_matrix()
        nnz, mnnz = A.begA[-1], M.begA[-1]
        numpy.savez_compressed(
            os.path.join(OUT, name + '.npz'),
            dim=d.dim, dof=d.dof, x=d.x, y=d.y, z=d.z, state=state,
            coA=A.coA[:nnz], jcoA=A.jcoA[:nnz], begA=A.begA, rhs=f,
            mcoA=M.coA[:mnnz], mjcoA=M.jcoA[:mnnz], mbegA=M.begA)
        print('%-18s n=%6d nnz=%7d' % (name, n, nnz))
    # user-supplied boundary conditions: the same callbacks the tests hand to the B200 Interface
    for name, (params, nx, ny, nz, dim, dof, kind, callback) in CUSTOM_BC_CASES.items():
        d = Discretization(dict(params), nx, ny, nz, dim, dof, boundary_conditions=callback)
        n = nx * ny * nz * d.dof
        state = make_state(kind, n)
        A = d.jacobian(state)
        f = d.rhs(state)
        nnz = A.begA[-1]
        numpy.savez_compressed(
            os.path.join(OUT, name + '.npz'),
            dim=d.dim, dof=d.dof, x=d.x, y=d.y, z=d.z, state=state,
            coA=A.coA[:nnz], jcoA=A.jcoA[:nnz], begA=A.begA, rhs=f)
        print('%-24s n=%6d nnz=%7d' % (name, n, nnz))


if __name__ == '__main__':
    main()
