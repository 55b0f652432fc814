'''Parity of the CUDA assembly path (through the Interface / C ABI) against the committed golden
vectors, the reference's own golden files and the CPU oracle.  Bit-exact for the integer
pattern; fp64 values are required to be within 1e-12 relative (north_star) and are in fact
compared bit-for-bit (the kernels are built with -fmad=false and keep the reference's
operation order).'''
import numpy
import pytest

from cases import CASES, CUSTOM_BC_CASES, _bc_unsupported, make_state
from golden_io import assert_csr_equal, compress, load_case, read_ref_matrix, read_ref_vector

pytestmark = pytest.mark.gpu

SUPPORTED = sorted(CASES)


def _iface(params, nx, ny, nz, dim, dof, x=None, y=None, z=None):
    from transiflow_b200 import Interface
    return Interface(dict(params), nx, ny, nz, dim, dof, x, y, z)


def _rhs_close(f, want, exact):
    if exact:
        assert numpy.array_equal(f, want), 'rhs not bit-identical, max abs diff %.3e' % numpy.abs(f - want).max()
    else:
        assert numpy.allclose(f, want, rtol=1e-12, atol=1e-12 * numpy.abs(want).max())


@pytest.mark.parametrize('name', SUPPORTED)
def test_gpu_matches_reference_generated_golden(name):
    params, nx, ny, nz, dim, dof, kind = CASES[name]
    g = load_case(name)
    it = _iface(params, nx, ny, nz, int(g['dim']), int(g['dof']), g['x'], g['y'], g['z'])
    state = g['state']
    jac = it.jacobian(state)
    row_ptr, col = it.pattern()
    got = compress(jac.values(), col, row_ptr)
    assert_csr_equal(got, (g['coA'], g['jcoA'], g['begA']), 0.0, name)
    exact = params.get('Problem Type') not in ('Double Gyre', 'AMOC')   # host cos() may differ by an ulp
    _rhs_close(it.rhs(state), g['rhs'], exact)
    # fused Jacobian+RHS launch gives the same bits as the separate launches
    jac2, f2 = it.jacobian_rhs(state)
    assert numpy.array_equal(jac2.values(), jac.values())
    assert numpy.array_equal(f2, it.rhs(state))
    # scipy export == reference CrsMatrix
    csr = jac.tocsr()
    assert numpy.array_equal(csr.indptr, g['begA']) and numpy.array_equal(csr.indices, g['jcoA'])
    assert numpy.array_equal(csr.data, g['coA'])
    M = it.mass_matrix().tocsr()
    assert numpy.array_equal(M.indptr, g['mbegA']) and numpy.array_equal(M.indices, g['mjcoA'])
    assert numpy.array_equal(M.data, g['mcoA'])


REF_FILES = [
    ('ldc', {'Reynolds Number': 100}, 3, 4, 4),
    ('ldc_stretched', {'Reynolds Number': 100, 'Grid Stretching': True}, 3, 4, 4),
    ('bous', {'Reynolds Number': 1, 'Rayleigh Number': 100, 'Prandtl Number': 100,
              'Problem Type': 'Rayleigh-Benard'}, 3, 5, 4),
    ('bous_stretched', {'Reynolds Number': 1, 'Rayleigh Number': 100, 'Prandtl Number': 100,
                        'Problem Type': 'Rayleigh-Benard', 'Grid Stretching': True}, 3, 5, 4),
    ('dhc', {'Reynolds Number': 1, 'Rayleigh Number': 100, 'Prandtl Number': 100,
             'Problem Type': 'Differentially Heated Cavity'}, 3, 5, 4),
    ('amoc', {'Reynolds Number': 16, 'Rayleigh Number': 4e4, 'Prandtl Number': 2.25, 'Lewis Number': 1,
              'Temperature Forcing': 1, 'Freshwater Flux': 1, 'Problem Type': 'AMOC'}, 2, 5, 1),
]


@pytest.mark.parametrize('stem,params,dim,dof,nz', REF_FILES, ids=[r[0] for r in REF_FILES])
def test_gpu_matches_reference_golden_files(stem, params, dim, dof, nz):
    '''The reference's own fixtures (tests/test_fvm.py:843-1257), state[i] = i+1 on 4x4x{4,1}.'''
    nx = ny = 4
    it = _iface(params, nx, ny, nz, dim, dof)
    state = make_state('lin', it.n)
    want = read_ref_matrix('%s_%dx%dx%d.txt' % (stem, nx, ny, nz), it.n)
    row_ptr, col = it.pattern()
    assert_csr_equal(compress(it.jacobian(state).values(), col, row_ptr), want, 1e-12, stem)
    want_rhs = read_ref_vector('%s_rhs_%dx%dx%d.txt' % (stem, nx, ny, nz))
    f = it.rhs(state)
    assert numpy.all(numpy.abs(f - want_rhs) <= 1e-12 * numpy.maximum(numpy.abs(want_rhs), numpy.abs(f).max() * 1e-3))


BIG = [
    ('ldc3d_32', {'Reynolds Number': 100}, 32, 32, 32),
    ('ldc3d_ragged', {'Reynolds Number': 100, 'Grid Stretching Factor': 1.5}, 37, 13, 9),
    ('ldc3d_wide', {'Reynolds Number': 100}, 70, 5, 6),
    ('rb3d_20', {'Problem Type': 'Rayleigh-Benard', 'Rayleigh Number': 2000.0, 'Prandtl Number': 10.0,
                 'Biot Number': 1.0, 'X-max': 10, 'Y-max': 10}, 20, 21, 22),
    ('dhc2d_64', {'Problem Type': 'Differentially Heated Cavity', 'Rayleigh Number': 1e4, 'Prandtl Number': 1000.0,
                  'Reynolds Number': 1, 'X-max': 0.051, 'Y-max': 1}, 64, 64, 1),
    ('ldc2d_32', {'Reynolds Number': 100, 'Grid Stretching Factor': 1.5}, 32, 32, 1),
    ('qg_256', {'Problem Type': 'Double Gyre', 'Reynolds Number': 16, 'Rossby Parameter': 1000,
                'Wind Stress Parameter': 1000}, 256, 128, 1),
    ('amoc_256', {'Problem Type': 'AMOC', 'Rayleigh Number': 4e4, 'Prandtl Number': 2.25, 'Lewis Number': 1,
                  'Freshwater Flux': 0.1, 'Temperature Forcing': 1, 'X-max': 5}, 256, 128, 1),
]


@pytest.mark.parametrize('name,params,nx,ny,nz', BIG, ids=[b[0] for b in BIG])
def test_gpu_matches_oracle_at_scale(name, params, nx, ny, nz):
    '''BASELINE config sizes (2D) and mid-size 3D grids, ragged tiles included, vs the C oracle.'''
    from oracle.tf_oracle import Oracle
    orc = Oracle(dict(params), nx, ny, nz)
    it = _iface(params, nx, ny, nz, None, None)
    assert numpy.array_equal(it.x, orc.x) and numpy.array_equal(it.y, orc.y) and numpy.array_equal(it.z, orc.z)
    state = make_state(123, it.n)
    row_ptr, col = it.pattern()
    jac, f = it.jacobian_rhs(state)
    assert_csr_equal(compress(jac.values(), col, row_ptr), orc.jacobian(state), 0.0, name)
    exact = params.get('Problem Type') not in ('Double Gyre', 'AMOC')
    _rhs_close(f, orc.rhs(state), exact)
    # structural nnz formulas fitted on the reference
    if name.startswith('ldc3d_32'):
        N = 32
        assert it.nnz == 57 * N**3 - 96 * N**2 + 36 * N


def test_parameters_are_reread_on_every_call():
    '''The parameter dict is shared and mutable (BaseInterface.py:65): continuation changes it
    between calls.'''
    from oracle.tf_oracle import Oracle
    params = {'Reynolds Number': 10, 'Lid Velocity': 1}
    it = _iface(params, 6, 6, 6, None, None)
    it.parameters = params   # same dict object the caller mutates
    state = make_state(5, it.n)
    for Re, lid in ((10, 1), (200, 1), (200, 3.5), (0, 1)):
        params['Reynolds Number'] = Re
        it.set_parameter('Lid Velocity', lid)
        orc = Oracle(dict(params), 6, 6, 6)
        assert numpy.array_equal(it.rhs(state), orc.rhs(state))
        row_ptr, col = it.pattern()
        assert_csr_equal(compress(it.jacobian(state).values(), col, row_ptr), orc.jacobian(state), 0.0, 'Re=%s' % Re)


def test_matvec_matches_scipy():
    it = _iface({'Reynolds Number': 100}, 8, 7, 6, None, None)
    state = make_state(1, it.n)
    jac = it.jacobian(state)
    v = make_state(2, it.n)
    want = jac.tocsr() @ v
    got = jac @ v
    assert numpy.allclose(got, want, rtol=1e-13, atol=1e-13 * numpy.abs(want).max())
    # complex vectors (JaDa's operators, JaDa.py:24-34): real and imaginary parts separately
    vc = v + 1j * make_state(3, it.n)
    wantc = jac.tocsr() @ vc
    gotc = jac @ vc
    assert numpy.iscomplexobj(gotc) and numpy.allclose(gotc, wantc, rtol=1e-13, atol=1e-13 * numpy.abs(wantc).max())


@pytest.mark.parametrize('name', sorted(CUSTOM_BC_CASES))
def test_user_boundary_conditions_on_the_device(name):
    """``Interface(..., boundary_conditions=callback)`` (Discretization.py:62-66,719; 'custom_reference_test' is the callback
    of the reference's tests/test_interface.py:117-150): the callback is recorded once, its op sequence picks a generated
    kernel family and its constants become kernel arguments.  CSR and RHS bit-identical to the reference run with the same
    callback; the Newton-update solve works on such a matrix like on any other."""
    from transiflow_b200 import Interface
    params, nx, ny, nz, dim, dof, kind, callback = CUSTOM_BC_CASES[name]
    g = load_case(name)
    it = Interface(dict(params), nx, ny, nz, dim, dof, g['x'], g['y'], g['z'], boundary_conditions=callback)
    jac, f = it.jacobian_rhs(g['state'])
    row_ptr, col = it.pattern()
    assert_csr_equal(compress(jac.values(), col, row_ptr), (g['coA'], g['jcoA'], g['begA']), 0.0, name)
    assert numpy.array_equal(f, g['rhs'])
    assert numpy.array_equal(it.rhs(g['state']), g['rhs'])
    # the reference's test only asks for rhs(zero state) to run
    it.rhs(it.vector())
    it.parameters['Iterative Solver'] = {'Maximum Iterations': 2000, 'Restart': 2000}
    dx = it.solve(jac, -f)
    assert it.last_solve['converged'], it.last_solve
    A = jac.tocsr().tolil()
    A[it.dim, :] = 0
    A[:, it.dim] = 0
    A[it.dim, it.dim] = -1
    b = -f.copy()
    b[it.dim] = 0
    r = A.tocsr() @ dx - b
    assert numpy.linalg.norm(r) <= 1e-8 * numpy.linalg.norm(b)


def test_unsupported_configurations_fail_loudly():
    from transiflow_b200 import Interface
    with pytest.raises(NotImplementedError):
        Interface({'Problem Type': 'Double Gyre'}, 4, 4, 1, 2, 4)      # QG with an extra scalar
    with pytest.raises(NotImplementedError):
        Interface({}, 4, 4, 4, boundary_conditions=lambda bc, atom: None)       # applies nothing: no such kernel family
    with pytest.raises(NotImplementedError, match='no kernel family'):
        Interface({}, 6, 6, 1, boundary_conditions=_bc_unsupported)             # free-slip side walls under a moving lid
    with pytest.raises(Exception):
        Interface({'Problem Type': 'nonsense'}, 4, 4, 4)


def test_full_size_128_cubed_properties():
    '''BASELINE headline size (3-D LDC 128^3, 8.4 M unknowns, 118 M non-zeros): size-independent
    properties instead of an element-wise oracle comparison -- the structural nnz polynomial fitted on
    the reference, linearity of J, and the finite-difference consistency of J with F that the
    reference checks in tests/test_jacobian.py:139-230 (F is quadratic, so the defect is O(eps)).'''
    N = 128
    it = _iface({'Reynolds Number': 100}, N, N, N, None, None)
    assert it.nnz == 57 * N**3 - 96 * N**2 + 36 * N
    rng = numpy.random.default_rng(0)
    x = rng.uniform(-0.5, 0.5, it.n)
    p = rng.uniform(-0.5, 0.5, it.n)
    q = rng.uniform(-0.5, 0.5, it.n)
    # wall-normal velocities on the far walls are not free unknowns (the padded state zeroes them,
    # utils.py:119-131): keep them zero like the reference's test does (tests/test_jacobian.py:28-36)
    for v in (x, p):
        g = v.reshape(N, N, N, 4)
        g[:, :, N - 1, 0] = 0
        g[:, N - 1, :, 1] = 0
        g[N - 1, :, :, 2] = 0
    jac, f0 = it.jacobian_rhs(x)
    jp, jq = jac @ p, jac @ q
    lin = jac @ (2.0 * p - 3.0 * q)
    assert numpy.abs(lin - (2.0 * jp - 3.0 * jq)).max() <= 1e-12 * numpy.abs(lin).max()
    errs = []
    for eps in (1e-3, 1e-5):
        fd = (it.rhs(x + eps * p) - f0) / eps
        errs.append(numpy.linalg.norm(fd - jp) / numpy.linalg.norm(jp))
    assert errs[0] < 1e-2 and errs[1] < 1e-4 and errs[1] < errs[0] / 30, errs
    # fused and separate launches agree bit-for-bit at full size
    assert numpy.array_equal(it.rhs(x), f0)


def test_64_cubed_matches_oracle():
    '''BASELINE config 3 (3-D LDC 64^3): element-wise against the C oracle (bit-identical).'''
    from oracle.tf_oracle import Oracle
    N = 64
    params = {'Reynolds Number': 100}
    it = _iface(params, N, N, N, None, None)
    orc = Oracle(dict(params), N, N, N)
    state = make_state(11, it.n)
    jac, f = it.jacobian_rhs(state)
    assert numpy.array_equal(f, orc.rhs(state))
    csr = jac.tocsr()
    coA, jcoA, begA = orc.jacobian(state)
    assert numpy.array_equal(csr.indptr, begA) and numpy.array_equal(csr.indices, jcoA)
    assert numpy.array_equal(csr.data, coA)


RB_PARAMS = {'Problem Type': 'Rayleigh-Benard', 'Rayleigh Number': 1000.0, 'Prandtl Number': 10.0, 'Biot Number': 1.0,
             'X-max': 10, 'Y-max': 10}


def test_full_size_rayleigh_benard_128_cubed_properties():
    '''BASELINE config 4 (3-D Rayleigh-Benard 128^3, dof 5, 10.5 M unknowns, 149 M non-zeros): the structural nnz
    polynomial fitted on the reference, linearity of J, finite-difference consistency of J with
    F (tests/test_jacobian.py:139-230), fused == separate launches.'''
    N = 128
    it = _iface(RB_PARAMS, N, N, N, None, None)
    assert it.nnz == 72 * N**3 - 110 * N**2 + 36 * N
    rng = numpy.random.default_rng(0)
    x = rng.uniform(-0.5, 0.5, it.n)
    p = rng.uniform(-0.5, 0.5, it.n)
    q = rng.uniform(-0.5, 0.5, it.n)
    for v in (x, p):
        g = v.reshape(N, N, N, 5)
        g[:, :, N - 1, 0] = 0
        g[:, N - 1, :, 1] = 0
        g[N - 1, :, :, 2] = 0
    jac, f0 = it.jacobian_rhs(x)
    jp, jq = jac @ p, jac @ q
    lin = jac @ (2.0 * p - 3.0 * q)
    assert numpy.abs(lin - (2.0 * jp - 3.0 * jq)).max() <= 1e-12 * numpy.abs(lin).max()
    errs = []
    for eps in (1e-3, 1e-5):
        fd = (it.rhs(x + eps * p) - f0) / eps
        errs.append(numpy.linalg.norm(fd - jp) / numpy.linalg.norm(jp))
    assert errs[0] < 1e-2 and errs[1] < 1e-4 and errs[1] < errs[0] / 30, errs
    assert numpy.array_equal(it.rhs(x), f0)


@pytest.mark.parametrize('nx,ny,nz', [(64, 48, 40), (33, 35, 37)])
def test_rayleigh_benard_matches_oracle_across_tiles(nx, ny, nz):
    '''Several 32-cell tiles in x, odd line counts and more planes than one z-chunk: every staged pair of cells of the
    72-slot rows (pair-padded staging, csrc/tfb_assemble.cuh) must land on its CSR offset -- bit-identical to the oracle.'''
    from oracle.tf_oracle import Oracle
    it = _iface(RB_PARAMS, nx, ny, nz, None, None)
    orc = Oracle(dict(RB_PARAMS), nx, ny, nz)
    state = make_state(5, it.n)
    jac, f = it.jacobian_rhs(state)
    assert numpy.array_equal(f, orc.rhs(state))
    csr = jac.tocsr()
    coA, jcoA, begA = orc.jacobian(state)
    assert numpy.array_equal(csr.indptr, begA) and numpy.array_equal(csr.indices, jcoA)
    assert numpy.array_equal(csr.data, coA)
