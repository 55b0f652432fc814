'''Host-side driver of Interface.eigs (transiflow_b200/eigs.py): shift-and-invert Arnoldi against dense
generalized eigenvalues of the oracle's (J, M) pencil.  The device solve is replaced by SuperLU here; the
GPU test in test_solve_gpu.py runs the same comparison through the CUDA solver.'''
import numpy
import pytest
import scipy.linalg
import scipy.sparse.linalg as spla

from oracle.tf_oracle import Oracle
from transiflow_b200.eigs import shift_invert_arnoldi


def dense_reference(o, state, num, target=0.0):
    J = o.jacobian_csr(state).tolil()
    row = o.dim
    J[row, :] = 0
    J[:, row] = 0
    J[row, row] = -1                                   # SciPy.py:212-216
    mco, mj, mb = o.mass_matrix()
    M = numpy.zeros(o.n)
    M[mj] = mco
    lam = scipy.linalg.eig(J.toarray(), numpy.diag(M), right=False)
    lam = lam[numpy.isfinite(lam)]
    lam = lam[numpy.argsort(numpy.abs(lam - target))[:num]]
    return J.tocsc(), M, numpy.array(sorted(lam, key=lambda x: -x.real))


@pytest.mark.parametrize('case', [
    ({'Problem Type': 'Lid-driven Cavity', 'Reynolds Number': 100.0, 'Lid Velocity': 1.0}, 8, 8, 1, 0.0),
    ({'Problem Type': 'Lid-driven Cavity', 'Reynolds Number': 50.0, 'Lid Velocity': 1.0}, 5, 5, 5, 0.0),
    ({'Problem Type': 'Rayleigh-Benard', 'Rayleigh Number': 2000.0, 'Prandtl Number': 10.0, 'Biot Number': 1.0,
      'X-max': 10.0}, 16, 8, 1, 0.0),
    ({'Problem Type': 'Lid-driven Cavity', 'Reynolds Number': 100.0, 'Lid Velocity': 1.0}, 8, 8, 1, -30.0),
])
def test_arnoldi_matches_dense_generalized_eigenvalues(case):
    params, nx, ny, nz, target = case
    o = Oracle(dict(params), nx, ny, nz)
    rng = numpy.random.default_rng(3)
    state = 0.05 * rng.standard_normal(o.n)
    num = 4
    J, M, want = dense_reference(o, state, num, target)
    lu = spla.splu((J - target * scipy.sparse.diags(M)).tocsc())
    lam, vec, ok = shift_invert_arnoldi(lambda v: lu.solve(M * v), o.n, num=num, target=target, tol=1e-9, max_dim=60)
    assert ok
    # conjugate pairs may be cut differently at the end of the list: compare the leading ones
    dist = numpy.abs(want[:num - 1, None] - lam[None, :]).min(axis=1)
    assert dist.max() <= 1e-6 * max(1.0, numpy.abs(want).max())
    # eigen-residual of the returned vectors
    for i in range(num - 1):
        r = J @ vec[:, i] - lam[i] * (M * vec[:, i])
        assert numpy.linalg.norm(r) <= 1e-6 * numpy.linalg.norm(J @ vec[:, i]) + 1e-9
