'''Parity of the device linear solver / Newton step with the SciPy backend's spsolve path.
Golden vectors were produced by the unmodified reference (tests/golden/make_golden_newton.py);
the oracle's restatement of the pinned direct solve checks larger cases on the box.'''
import os

import numpy
import pytest

from cases_newton import NEWTON_CASES
from golden_io import GEN

pytestmark = pytest.mark.gpu


def newton(it, x0, tol=1e-10, maxit=10):
    '''Continuation.newton (Continuation.py:67-114), residual_check='F'.'''
    x = x0
    for k in range(maxit):
        fval = it.rhs(x)
        if numpy.linalg.norm(fval) < tol:
            break
        jac = it.jacobian(x)
        x = x + it.solve(jac, -fval)
    return x, k


def _iface(params, nx, ny, nz):
    from transiflow_b200 import Interface
    params = dict(params)
    if params.get('Problem Type') in ('AMOC', 'Rayleigh-Benard', 'Double Gyre'):
        # the diffusion-based block preconditioner is weak for strongly coupled buoyancy / Coriolis
        # terms; these small systems are solved by FULL (un-restarted) GMRES -- see DESIGN.md
        params['Iterative Solver'] = {'Maximum Iterations': 2000, 'Restart': 2000}
    return Interface(params, nx, ny, nz)


def test_solve_manufactured_solution():
    '''reference tests/test_interface.py:17-50'''
    nx = 4
    it = _iface({}, nx, nx, nx)
    A = it.jacobian(it.vector())
    x = numpy.arange(1, it.n + 1, dtype=float)
    x[3] = 0
    b = A @ x
    y = it.solve(A, b)
    pressure = y[3] - x[3]
    assert numpy.linalg.norm(y) > 0
    assert (numpy.linalg.norm(y - x) - pressure * nx**3) / numpy.linalg.norm(b) < 1e-7
    assert it.last_solve['converged']


@pytest.mark.parametrize('name', ['ldc3d_8', 'ldc3d_12_str', 'ldc2d_24', 'dhc2d_16', 'rb3d_8', 'qg_16', 'amoc_16'])
def test_linear_solve_matches_reference(name):
    '''J(x*) y = b at the reference's converged state: y within 1e-8 of SuperLU's.'''
    params, nx, ny, nz = NEWTON_CASES[name]
    g = numpy.load(os.path.join(GEN, 'newton_' + name + '.npz'))
    it = _iface(params, nx, ny, nz)
    jac = it.jacobian(g['x'])
    y = it.solve(jac, g['b'])
    assert it.last_solve['converged'], it.last_solve
    scale = numpy.abs(g['y']).max()
    assert numpy.abs(y - g['y']).max() <= 1e-8 * scale, (numpy.abs(y - g['y']).max() / scale, it.last_solve)


@pytest.mark.parametrize('name', ['ldc3d_8', 'ldc3d_12_str', 'ldc2d_24', 'dhc2d_16', 'rb3d_8', 'qg_16', 'amoc_16'])
def test_newton_converges_to_reference_state(name):
    params, nx, ny, nz = NEWTON_CASES[name]
    g = numpy.load(os.path.join(GEN, 'newton_' + name + '.npz'))
    it = _iface(params, nx, ny, nz)
    x, k = newton(it, it.vector())
    # AMOC: the reference's own Newton run stops at its iteration limit with |F| = 1.8e-8 (golden `fnorm`); the bar is the
    # residual the SciPy backend reached, and the same state
    assert numpy.linalg.norm(it.rhs(x)) < max(1e-9, 1.001 * float(g['fnorm']))
    scale = max(numpy.abs(g['x']).max(), 1e-300)
    assert numpy.abs(x - g['x']).max() <= 1e-8 * scale, numpy.abs(x - g['x']).max() / scale


def test_solve_against_oracle_direct_solve_32():
    '''3-D LDC 20^3 (SuperLU still feasible): Newton update within 1e-8 of the spsolve path.'''
    from oracle.tf_oracle import Oracle, direct_solve
    params, N = {'Reynolds Number': 100}, 20
    it = _iface(params, N, N, N)
    orc = Oracle(dict(params), N, N, N)
    x = numpy.zeros(it.n)
    for _ in range(2):
        f = it.rhs(x)
        jac = it.jacobian(x)
        dx = it.solve(jac, -f)
        want = direct_solve(orc.jacobian_csr(x), -orc.rhs(x), orc.dim, orc.dof)
        assert numpy.abs(dx - want).max() <= 1e-8 * numpy.abs(want).max(), it.last_solve
        x = x + dx


def test_bordered_solve():
    '''[J V; W^T C][y1; y2] = [b; b2] (SciPy.py:226-250) by block elimination.'''
    it = _iface({'Reynolds Number': 50}, 6, 6, 6)
    x = numpy.random.default_rng(0).uniform(-0.1, 0.1, it.n)
    jac = it.jacobian(x)
    rng = numpy.random.default_rng(1)
    V, W, b = rng.uniform(-1, 1, it.n), rng.uniform(-1, 1, it.n), rng.uniform(-1, 1, it.n)
    V[3] = W[3] = b[3] = 0
    C, b2 = 0.7, 0.3
    y1, y2 = it.solve(jac, b, b2, V, W, C)
    J = jac.tocsr().tolil()
    J[3, :] = 0
    J[:, 3] = 0
    J[3, 3] = -1
    J = J.tocsr()
    r1 = J @ y1 + V * y2 - b
    r2 = W @ y1 + C * y2 - b2
    assert numpy.linalg.norm(r1) <= 1e-8 * numpy.linalg.norm(b) and abs(r2) < 1e-8


def test_complex_right_hand_side_is_solved_by_parts():
    '''SciPy.Interface._lu_solve (SciPy.py:194-202): a complex rhs with the real Jacobian -> real and imaginary solves.'''
    it = _iface({'Reynolds Number': 50}, 6, 6, 6)
    x = numpy.random.default_rng(0).uniform(-0.1, 0.1, it.n)
    jac = it.jacobian(x)
    rng = numpy.random.default_rng(2)
    b = rng.uniform(-1, 1, it.n) + 1j * rng.uniform(-1, 1, it.n)
    b[3] = 0
    y = it.solve(jac, b)
    assert numpy.iscomplexobj(y) and it.last_solve['converged']
    J = jac.tocsr().tolil()
    J[3, :] = 0
    J[:, 3] = 0
    J[3, 3] = -1
    assert numpy.linalg.norm(J.tocsr() @ y - b) <= 1e-8 * numpy.linalg.norm(b)
    # the same solve as a purely real one (bitwise up to the order of the atomic partial sums of the Krylov reductions)
    again = it.solve(jac, b.real.copy())
    assert numpy.abs(y.real - again).max() <= 1e-12 * numpy.abs(again).max()


def test_complex_shifted_matrices_are_solved_on_the_device():
    """`beta * J - alpha * M` with complex alpha, the matrix the reference's JaDa glue hands to solve() (JaDa.py:90,
    149-151,187; SuperLU on a complex matrix there): solved with a complex Krylov iteration on device operators."""
    import scipy.sparse.linalg
    it = _iface({'Reynolds Number': 50}, 6, 6, 6)
    x = numpy.random.default_rng(0).uniform(-0.1, 0.1, it.n)
    jac, mass = it.jacobian(x), it.mass_matrix()
    shifted = 1.0 * jac - (0.3 + 0.7j) * mass                  # scipy complex matrix, as JaDa builds it
    assert numpy.iscomplexobj(shifted.data)
    rng = numpy.random.default_rng(2)
    b = rng.uniform(-1, 1, it.n) + 1j * rng.uniform(-1, 1, it.n)
    y = it.solve(shifted, b)
    assert it.last_solve['converged'], it.last_solve
    A = shifted.tolil()
    A[3, :] = 0
    A[:, 3] = 0
    A[3, 3] = -1
    bb = b.copy()
    bb[3] = 0
    want = scipy.sparse.linalg.spsolve(A.tocsc(), bb)
    assert numpy.abs(y - want).max() <= 1e-8 * numpy.abs(want).max()


def test_eigs_with_a_complex_target():
    """'Target': 1 + 3j style targets (the reference's test configuration stores one, tests/test_interface.py:97):
    eigenvalues closest to a complex shift, against the dense generalized eigenvalues of the oracle's pinned pencil."""
    import scipy.linalg
    from oracle.tf_oracle import Oracle
    from transiflow_b200 import Interface
    params = {'Problem Type': 'Lid-driven Cavity', 'Reynolds Number': 100.0, 'Lid Velocity': 1.0}
    nx = ny = 8
    o = Oracle(dict(params), nx, ny, 1)
    state = 0.05 * numpy.random.default_rng(3).standard_normal(o.n)
    J = o.jacobian_csr(state).tolil()
    J[o.dim, :] = 0
    J[:, o.dim] = 0
    J[o.dim, o.dim] = -1
    mco, mj, mb = o.mass_matrix()
    M = numpy.zeros(o.n)
    M[mj] = mco
    lam = scipy.linalg.eig(J.toarray(), numpy.diag(M), right=False)
    lam = lam[numpy.isfinite(lam)]
    target, num = -20.0 + 5.0j, 3
    want = lam[numpy.argsort(numpy.abs(lam - target))[:num]]
    params['Eigenvalue Solver'] = {'Target': target, 'Number of Eigenvalues': num, 'Tolerance': 1e-8}
    it = Interface(params, nx, ny, 1)
    got = it.eigs(state)
    dist = numpy.abs(want[:num - 1, None] - got[None, :]).min(axis=1)
    assert dist.max() <= 1e-6 * max(1.0, numpy.abs(want).max())


def test_time_integration_operators():
    """TimeIntegration._newton (TimeIntegration.py:40-73) builds `jacobian(x) - mass / (theta*dt)` and
    `mass @ v` with the backend's matrix types and hands the result to solve()."""
    it = _iface({'Reynolds Number': 100}, 8, 8, 8)
    x = numpy.random.default_rng(0).uniform(-0.1, 0.1, it.n)
    jac, mass = it.jacobian(x), it.mass_matrix()
    theta, dt = 1.0, 0.1
    A = jac - mass / (theta * dt)
    from transiflow_b200 import DeviceMatrix
    assert isinstance(A, DeviceMatrix)          # diagonal update happens on the device, no host round trip
    want = (jac.tocsc() - mass / (theta * dt)).tocsr()
    got = A.tocsr()
    assert abs(got - want).max() <= 1e-15 * abs(want).max()
    v = numpy.random.default_rng(1).uniform(-1, 1, it.n)
    b = mass @ v + it.rhs(x)
    b[3] = 0
    y = it.solve(A, b)
    assert it.last_solve['converged']
    Ah = A.tocsr().tolil()
    Ah[3, :] = 0
    Ah[:, 3] = 0
    Ah[3, 3] = -1
    r = Ah.tocsr() @ y - b
    assert numpy.linalg.norm(r) <= 1e-8 * numpy.linalg.norm(b)


@pytest.mark.parametrize('name', ['ldc2d_24', 'dhc2d_16', 'qg_16', 'amoc_16'])
def test_direct_solve_on_2d_grids_matches_superlu(name):
    """2-D grids are solved directly by default (block-tridiagonal elimination over the grid lines, csrc/tfb_direct.cu):
    spsolve-grade agreement with the reference's SuperLU answer, factors reused for the second solve of a matrix."""
    params, nx, ny, nz = NEWTON_CASES[name]
    g = numpy.load(os.path.join(GEN, 'newton_' + name + '.npz'))
    from transiflow_b200 import Interface
    it = Interface(dict(params), nx, ny, nz)
    jac = it.jacobian(g['x'])
    y = it.solve(jac, g['b'])
    assert it.last_solve['method'] == 'Direct' and it.last_solve['converged'] and it.last_solve['setup_ms'] > 0
    scale = numpy.abs(g['y']).max()
    assert numpy.abs(y - g['y']).max() <= 1e-9 * scale, numpy.abs(y - g['y']).max() / scale
    y2 = it.solve(jac, 2 * g['b'])
    assert it.last_solve['setup_ms'] == 0                      # cached factors, like jac.lu in SciPy.py:142-152
    assert numpy.abs(y2 - 2 * g['y']).max() <= 1e-9 * 2 * scale
    # a new Jacobian is factored again; the Krylov solver remains available on request
    jac2 = it.jacobian(0.5 * g['x'])
    it.solve(jac2, g['b'])
    assert it.last_solve['setup_ms'] > 0
    it.parameters['Iterative Solver'] = {'Method': 'FGMRES', 'Maximum Iterations': 2000, 'Restart': 2000}
    yk = it.solve(jac, g['b'])
    assert it.last_solve['method'] == 'FGMRES'
    assert numpy.abs(yk - g['y']).max() <= 1e-8 * scale


@pytest.mark.parametrize('problem,nx,ny', [('ldc', 64, 12), ('ldc', 55, 9), ('dhc', 52, 10), ('ldc', 40, 7)])
def test_direct_solve_with_line_blocks_beyond_one_panel(problem, nx, ny):
    """Line blocks that do not fit in shared memory whole are eliminated panel by panel (k_gj_blocked: NB columns on one
    CTA, one rank-NB update of the other columns by all CTAs); the last panel may be ragged (55 * 3 = 165 rows).  Against
    SuperLU on the pinned host matrix, as SciPy.py:204-258 solves it."""
    import scipy.sparse.linalg
    params = {'ldc': {'Reynolds Number': 200, 'Grid Stretching Factor': 1.5},
              'dhc': {'Problem Type': 'Differentially Heated Cavity', 'Rayleigh Number': 1e4, 'Prandtl Number': 1000,
                      'Reynolds Number': 1}}[problem]
    it = _iface(params, nx, ny, 1)
    x = numpy.random.default_rng(7).uniform(-0.2, 0.2, it.n)
    jac = it.jacobian(x)
    b = numpy.random.default_rng(8).uniform(-1, 1, it.n)
    b[it.dim] = 0
    y = it.solve(jac, b)
    assert it.last_solve['method'] == 'Direct' and it.last_solve['converged'], it.last_solve
    A = jac.tocsr().tolil()
    A[it.dim, :] = 0
    A[:, it.dim] = 0
    A[it.dim, it.dim] = -1
    want = scipy.sparse.linalg.spsolve(A.tocsc(), b)
    assert numpy.abs(y - want).max() <= 1e-9 * numpy.abs(want).max()


def test_semi_2d_heated_cavity_solve():
    """dim = 3 on an nz = 1 grid (the z-offsets fold onto the cell, Discretization.py:127-128) for the heated cavity, dof 5:
    the folded kernel family `dhc3d_flat` and the direct solve, against SuperLU on the pinned host matrix."""
    import scipy.sparse.linalg
    p = {'Problem Type': 'Differentially Heated Cavity', 'Rayleigh Number': 1e4, 'Prandtl Number': 1000.0,
         'Reynolds Number': 1, 'X-max': 0.051, 'Y-max': 1}
    from transiflow_b200 import Interface
    it = Interface(dict(p), 12, 10, 1, 3, 5)
    x = numpy.random.default_rng(1).uniform(-0.1, 0.1, it.n)
    jac, f = it.jacobian_rhs(x)
    dx = it.solve(jac, -f)
    assert it.last_solve['converged'], it.last_solve
    A = jac.tocsr().tolil()
    pr = it.pressure_row
    A[pr, :] = 0
    A[:, pr] = 0
    A[pr, pr] = -1
    b = -f.copy()
    b[pr] = 0
    want = scipy.sparse.linalg.spsolve(A.tocsc(), b)
    assert numpy.abs(dx - want).max() <= 1e-8 * numpy.abs(want).max()


def test_full_size_128_cubed_newton_update_properties():
    """BASELINE headline size (3-D cavity 128^3, 8.4 M unknowns): SuperLU is out of reach there, so the Newton update of
    the solver that bench.py times (IDR(8), scaled-mass Schur complement, tensor-core sub-solves -- what 'auto' picks at
    this size) is checked through size-independent properties: the residual of the pinned system recomputed on the HOST
    from the downloaded CSR matrix (scipy, independent of every device kernel of the solve), and linearity in the
    right-hand side."""
    N = 128
    it = _iface({'Reynolds Number': 100, 'Lid Velocity': 1}, N, N, N)
    x = numpy.zeros(it.n)
    x[0::it.dof] = 1e-3 * numpy.sin(numpy.arange(it.n // it.dof))          # a state with convection switched on
    jac, f = it.jacobian_rhs(x)
    b = -f
    dx = it.solve(jac, b)
    ls = dict(it.last_solve)
    assert ls['method'] == 'IDR' and ls['schur'] == 'Scaled Mass' and ls['converged'] and ls['relres'] <= 1e-10, ls
    A = jac.tocsr()
    prow = it.pressure_row
    bb = b.copy()
    bb[prow] = 0.0

    def host_residual(y, rhs):
        z = y.copy()
        zp, z[prow] = z[prow], 0.0               # pinned column dropped ...
        r = A @ z
        r[prow] = -zp                            # ... pinned row: -1 on the diagonal (SciPy.py:95-106,212-216)
        return numpy.linalg.norm(r - rhs) / numpy.linalg.norm(rhs)

    assert host_residual(dx, bb) <= 2e-10
    dx2 = it.solve(jac, 2.0 * b)
    assert it.last_solve['converged']
    assert host_residual(dx2, 2.0 * bb) <= 2e-10
    # two Krylov solves to the same RESIDUAL tolerance: the velocities agree far better than the pressures, whose error is
    # the residual times the condition number of the saddle-point system (scaled-mass caveat)
    vel = numpy.ones(it.n, dtype=bool)
    vel[it.dim::it.dof] = False
    assert numpy.abs(dx2 - 2.0 * dx)[vel].max() <= 1e-4 * numpy.abs(dx[vel]).max()


@pytest.mark.parametrize('name', ['ldc2d', 'dhc2d', 'qg', 'amoc'])
def test_direct_solve_at_the_baseline_sizes(name):
    """The 2-D BASELINE configurations at their full sizes (32 x 32 stretched cavity, 64 x 64 heated cavity, 256 x 128
    double gyre and AMOC: line blocks of 96 to 1280 rows, the larger ones eliminated in panels over all SMs): the Newton
    update of the device solve against SuperLU on the pinned host matrix (SciPy.py:204-258), 1e-8 as SURVEY section 8(d)
    asks.  SuperLU needs ~10 s for the AMOC matrix on the box's host."""
    import scipy.sparse.linalg
    from bench import PROBLEMS_2D
    params, nx, ny, _ = PROBLEMS_2D[name]
    it = _iface(dict(params), nx, ny, 1)
    x = numpy.random.default_rng(11).uniform(-0.01, 0.01, it.n)
    jac, f = it.jacobian_rhs(x)
    b = -f
    y = it.solve(jac, b)
    assert it.last_solve['method'] == 'Direct' and it.last_solve['converged'], it.last_solve
    import scipy.sparse
    prow = it.pressure_row
    keep = numpy.ones(it.n)
    keep[prow] = 0.0
    D = scipy.sparse.diags(keep)
    pin = scipy.sparse.csr_matrix(([-1.0], ([prow], [prow])), shape=(it.n, it.n))
    A = (D @ jac.tocsr() @ D + pin).tocsc()           # row and column of the pinned pressure removed, -1 on the diagonal
    bb = b.copy()
    bb[prow] = 0
    want = scipy.sparse.linalg.spsolve(A, bb)
    assert numpy.abs(y - want).max() <= 1e-8 * numpy.abs(want).max()


@pytest.mark.parametrize('grid', [(8, 8, 8), (24, 24, 1)])
def test_mass_shifted_matrices_carry_the_shift_into_the_preconditioner(grid):
    """J - M / (theta dt) with a small time step (and J - sigma M with a large shift) are dominated by the mass term; the
    fast-diagonalisation basis is M-orthonormal, so the block preconditioner solves the shifted diffusion operators
    exactly and the iteration count must not blow up (it did when the shift was ignored)."""
    nx, ny, nz = grid
    it = _iface({'Reynolds Number': 100, 'Iterative Solver': {'Method': 'FGMRES'}}, nx, ny, nz)
    x = numpy.random.default_rng(0).uniform(-0.1, 0.1, it.n)
    jac, mass = it.jacobian(x), it.mass_matrix()
    b = numpy.random.default_rng(1).uniform(-1, 1, it.n)
    b[it.dim] = 0
    it.solve(jac, b)
    base_its = it.last_solve['iterations']
    for dt in (1e-1, 1e-3):
        A = jac - mass / dt
        y = it.solve(A, b)
        assert it.last_solve['converged'], it.last_solve
        assert it.last_solve['iterations'] <= base_its + 10, (dt, it.last_solve['iterations'], base_its)
        Ah = A.tocsr().tolil()
        Ah[it.dim, :] = 0
        Ah[:, it.dim] = 0
        Ah[it.dim, it.dim] = -1
        assert numpy.linalg.norm(Ah.tocsr() @ y - b) <= 1e-8 * numpy.linalg.norm(b)


@pytest.mark.parametrize('params', [{'Reynolds Number': 100},
                                    {'Problem Type': 'Rayleigh-Benard', 'Rayleigh Number': 500.0, 'Prandtl Number': 10.0,
                                     'Biot Number': 1.0, 'X-max': 10}])
def test_semi_2d_grids_solve_like_the_spsolve_path(params):
    """dim = 3 on an nz = 1 grid (the reference compares 2-D and semi-2-D continuations, tests/test_continuation.py:84-93):
    Newton updates within 1e-8 of the pinned SuperLU solve, and the (u, v, p) part equal to the 2-D problem's."""
    from oracle.tf_oracle import Oracle, direct_solve
    from transiflow_b200 import Interface
    nx, ny = 12, 10
    p3 = dict(params, **{'Iterative Solver': {'Maximum Iterations': 2000, 'Restart': 2000}})
    it = Interface(p3, nx, ny, 1, dim=3)
    orc = Oracle(dict(params), nx, ny, 1, dim=3)
    x = numpy.zeros(it.n)
    for _ in range(2):
        f = it.rhs(x)
        jac = it.jacobian(x)
        dx = it.solve(jac, -f)
        assert it.last_solve['converged'], it.last_solve
        want = direct_solve(orc.jacobian_csr(x), -orc.rhs(x), orc.dim, orc.dof)
        assert numpy.abs(dx - want).max() <= 1e-8 * numpy.abs(want).max(), it.last_solve
        x = x + dx


def test_structured_spmv_matches_csr_kernel_and_scipy():
    """3-D grids use the column-index-free marching SpMV; it must agree with scipy on ragged grids."""
    for params, nx, ny, nz in (({'Reynolds Number': 100}, 37, 9, 19),
                               ({'Problem Type': 'Rayleigh-Benard', 'Rayleigh Number': 800.0, 'Prandtl Number': 10.0,
                                 'Biot Number': 1.0, 'X-max': 10, 'Y-max': 10}, 33, 7, 18)):
        it = _iface(params, nx, ny, nz)
        x = numpy.random.default_rng(0).uniform(-0.5, 0.5, it.n)
        jac = it.jacobian(x)
        v = numpy.random.default_rng(1).uniform(-1, 1, it.n)
        want = jac.tocsr() @ v
        got = jac @ v
        assert numpy.abs(got - want).max() <= 1e-13 * numpy.abs(want).max()


def test_parameter_continuation_reaches_the_reference_branch_point():
    """The reference's pseudo-arclength continuation (Continuation.continuation, SciPy backend) ends at the
    steady state for Re = 400; stepping the shared parameter dict through set_parameter and converging
    Newton with the device solver must land on the same state (tests/golden/make_golden_continuation.py)."""
    g = numpy.load(os.path.join(GEN, 'continuation_ldc2d.npz'))
    params = {'Problem Type': 'Lid-driven Cavity', 'Reynolds Number': 0, 'Lid Velocity': 1, 'Grid Stretching Factor': 1.5}
    from transiflow_b200 import Interface
    it = Interface(params, int(g['nx']), int(g['ny']))
    x = it.vector()
    for Re in (0, 100, 200, 300, float(g['mu'])):
        it.set_parameter('Reynolds Number', Re)
        x, k = newton(it, x, tol=1e-11, maxit=12)
    assert numpy.linalg.norm(it.rhs(x)) < 1e-10
    assert numpy.abs(x - g['x']).max() <= 1e-8 * numpy.abs(g['x']).max()
    # the reference's un-polished continuation end point lies within its own corrector tolerance of that state
    assert numpy.abs(g['x_continuation'] - x).max() <= 1e-2 * numpy.abs(x).max()


def test_implicit_euler_matches_reference_time_integration():
    """TimeIntegration.integration with theta = 1 (TimeIntegration.py:40-115) restated on the backend's matrix
    types: mass @ v, jacobian(x) - mass / (theta * dt) (device-side), solve.  Golden: the reference's own run
    (tests/golden/make_golden_time.py)."""
    g = numpy.load(os.path.join(GEN, 'time_ldc2d.npz'))
    params = {'Problem Type': 'Lid-driven Cavity', 'Reynolds Number': 100, 'Lid Velocity': 1}
    from transiflow_b200 import Interface
    it = Interface(params, int(g['nx']), int(g['ny']))
    theta, dt = 1.0, float(g['dt'])
    x = it.vector()
    for _ in range(int(g['steps'])):
        x0 = x
        b0 = it.rhs(x0)
        mass = it.mass_matrix()
        for k in range(10):
            fval = mass @ (x0 - x) + dt * theta * it.rhs(x) + dt * (1 - theta) * b0
            fval /= theta * dt
            if numpy.linalg.norm(fval) < 1e-10:
                break
            jac = it.jacobian(x) - mass / (theta * dt)
            x = x + it.solve(jac, -fval)
    assert numpy.abs(x - g['x']).max() <= 1e-8 * numpy.abs(g['x']).max()


@pytest.mark.parametrize('options', [{'Method': 'BiCGStab'}, {'Basis Precision': 'single', 'Restart': 40},
                                     {'Preconditioner Precision': 'single'}, {'Preconditioner Precision': 'tf32x3'},
                                     {'Schur Complement': 'Scaled Mass', 'Method': 'IDR', 'Preconditioner Precision': 'tf32x3'},
                                     {'Schur Complement': 'Scaled Mass', 'Method': 'IDR', 'Preconditioner Precision': 'double'},
                                     {'Preconditioner Precision': 'single', 'Basis Precision': 'single'},
                                     {'Velocity Iterations': 3}, {'Method': 'IDR'}, {'Method': 'IDR', 'IDR Dimension': 4},
                                     {'Method': 'FGMRES'}, {'Schur Complement': 'Scaled Mass', 'Method': 'FGMRES'},
                                     {'Schur Complement': 'Scaled Mass', 'Method': 'IDR'}])
def test_alternative_krylov_options_reach_the_same_solution(options):
    """Every Krylov method / storage / preconditioner option must deliver the same 1e-10 true residual and the
    SuperLU parity of the default."""
    name = 'ldc3d_12_str'
    params, nx, ny, nz = NEWTON_CASES[name]
    g = numpy.load(os.path.join(GEN, 'newton_' + name + '.npz'))
    params = dict(params)
    params['Iterative Solver'] = dict(options)
    from transiflow_b200 import Interface
    it = Interface(params, nx, ny, nz)
    jac = it.jacobian(g['x'])
    y = it.solve(jac, g['b'])
    assert it.last_solve['converged'], it.last_solve
    assert numpy.abs(y - g['y']).max() <= 1e-8 * numpy.abs(g['y']).max()
    if 'Method' in options:
        assert it.last_solve['method'].lower() == options['Method'].lower()


def test_automatic_method_uses_idr_on_large_grids():
    """'Method': 'auto' -> IDR(8) from AUTO_IDR_MIN_UNKNOWNS unknowns on 3-D grids, FGMRES below."""
    from transiflow_b200 import Interface
    name = 'ldc3d_12_str'
    params, nx, ny, nz = NEWTON_CASES[name]
    g = numpy.load(os.path.join(GEN, 'newton_' + name + '.npz'))
    it = Interface(dict(params), nx, ny, nz)
    jac = it.jacobian(g['x'])
    it.solve(jac, g['b'])
    assert it.last_solve['method'] == 'FGMRES' and it.last_solve['schur'] == 'LSC'          # small grid
    it.AUTO_IDR_MIN_UNKNOWNS = 1000
    y = it.solve(jac, g['b'])
    assert it.last_solve['method'] == 'IDR' and it.last_solve['schur'] == 'Scaled Mass' and it.last_solve['converged']
    assert it.last_solve['precond_precision'] == 'tf32x3'       # tensor-core FDM sub-solves wherever IDR + scaled mass is automatic
    assert numpy.abs(y - g['y']).max() <= 1e-8 * numpy.abs(g['y']).max()


@pytest.mark.parametrize('name,params', [('ldc3d_20_re100', {'Reynolds Number': 100}),
                                         ('ldc3d_16_re400_str', {'Reynolds Number': 400, 'Grid Stretching Factor': 1.5})])
def test_newton_with_the_large_grid_defaults_reaches_the_spsolve_state(name, params):
    """The solver configuration that 'auto' selects on large 3-D grids (IDR(8), scaled-mass Schur complement, CUDA-graph
    replay of the preconditioner), forced onto grids where SuperLU is still feasible: Newton from zero converges in the
    same number of steps to the state of the spsolve path (tests/golden/make_golden_newton_oracle.py) within 1e-8."""
    from transiflow_b200 import Interface
    g = numpy.load(os.path.join(GEN, 'newton_oracle_' + name + '.npz'))
    N = int(g['N'])
    it = Interface(dict(params), N, N, N)
    it.AUTO_IDR_MIN_UNKNOWNS = 1000
    x, k = newton(it, it.vector(), tol=1e-12, maxit=12)
    assert it.last_solve['method'] == 'IDR' and it.last_solve['schur'] == 'Scaled Mass'
    assert k == int(g['steps'])
    assert numpy.linalg.norm(it.rhs(x)) < 1e-12
    assert numpy.abs(x - g['x']).max() <= 1e-8 * numpy.abs(g['x']).max()


@pytest.mark.parametrize('case', [
    ({'Problem Type': 'Lid-driven Cavity', 'Reynolds Number': 100.0, 'Lid Velocity': 1.0}, 8, 8, 1, 0.0),
    ({'Problem Type': 'Lid-driven Cavity', 'Reynolds Number': 50.0, 'Lid Velocity': 1.0}, 6, 6, 6, 0.0),
    ({'Problem Type': 'Lid-driven Cavity', 'Reynolds Number': 100.0, 'Lid Velocity': 1.0}, 8, 8, 1, -30.0),
    ({'Problem Type': 'Differentially Heated Cavity', 'Rayleigh Number': 1e4, 'Prandtl Number': 1000.0,
      'Reynolds Number': 1.0}, 8, 8, 1, 0.0),
    ({'Problem Type': 'Rayleigh-Benard', 'Rayleigh Number': 1000.0, 'Prandtl Number': 10.0, 'Biot Number': 1.0,
      'X-max': 10.0}, 16, 8, 1, 0.0),
])
def test_eigs_matches_dense_generalized_eigenvalues(case):
    """Interface.eigs (BaseInterface.py:363-386): eigenvalues of J v = lambda M v closest to the target, through
    the device solver, against scipy.linalg.eig on the oracle's pinned (J, M) pencil."""
    import scipy.linalg
    from oracle.tf_oracle import Oracle
    from transiflow_b200 import Interface
    params, nx, ny, nz, target = case
    params = dict(params)
    o = Oracle(dict(params), nx, ny, nz)
    state = 0.05 * numpy.random.default_rng(3).standard_normal(o.n)
    J = o.jacobian_csr(state).tolil()
    J[o.dim, :] = 0
    J[:, o.dim] = 0
    J[o.dim, o.dim] = -1
    mco, mj, mb = o.mass_matrix()
    M = numpy.zeros(o.n)
    M[mj] = mco
    lam = scipy.linalg.eig(J.toarray(), numpy.diag(M), right=False)
    lam = lam[numpy.isfinite(lam)]
    num = 4
    want = lam[numpy.argsort(numpy.abs(lam - target))[:num]]
    params['Eigenvalue Solver'] = {'Target': target, 'Number of Eigenvalues': num, 'Tolerance': 1e-8}
    it = Interface(params, nx, ny, nz)
    got, vec = it.eigs(state, return_eigenvectors=True)
    assert numpy.all(numpy.diff(got.real) <= 1e-12)                 # sorted by descending real part
    dist = numpy.abs(want[:num - 1, None] - got[None, :]).min(axis=1)
    assert dist.max() <= 1e-6 * max(1.0, numpy.abs(want).max())
    Jc = J.tocsr()
    for i in range(num - 1):
        r = Jc @ vec[:, i] - got[i] * (M * vec[:, i])
        assert numpy.linalg.norm(r) <= 1e-6 * numpy.linalg.norm(Jc @ vec[:, i]) + 1e-9


def test_eigs_reference_configuration_ldc_6x6_re2000():
    """The reference's own eigenvalue test configuration (tests/jada_fixtures.py:19-85, test_jada.py:108-122):
    6x6 lid-driven cavity continued to Re = 2000, 10 eigenvalues closest to zero; golden = ARPACK/dense QZ on the
    reference's matrices (tests/golden/make_golden_eigs.py); tolerances as in the reference (atol = 100 tol)."""
    from transiflow_b200 import Interface
    g = numpy.load(os.path.join(GEN, 'eigs_ldc2d_6_re2000.npz'))
    num, tol = 10, 1e-7
    params = {'Reynolds Number': float(g['reynolds']),
              'Eigenvalue Solver': {'Number of Eigenvalues': num, 'Tolerance': tol}}
    it = Interface(params, 6, 6)
    got = it.eigs(g['x'])
    got = numpy.array(sorted(got, key=lambda z: abs(z)))
    want = g['eigs_dense']
    numpy.testing.assert_allclose(got.real, want.real, rtol=0, atol=100 * tol)
    numpy.testing.assert_allclose(abs(got.imag), abs(want.imag), rtol=0, atol=100 * tol)
