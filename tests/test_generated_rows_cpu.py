'''The generated per-row assembly functions (exactly the code each CUDA thread executes),
compiled for the host, against the oracle and the committed golden vectors.  CPU-only.'''
import numpy
import pytest

from cases import CASES, make_state
from golden_io import assert_csr_equal, compress, load_case
from harness_util import assemble
from oracle.tf_oracle import Oracle


@pytest.mark.parametrize('name', sorted(CASES))
def test_rows_match_golden(name):
    params, nx, ny, nz, dim, dof, kind = CASES[name]
    g = load_case(name)
    dim, dof = int(g['dim']), int(g['dof'])
    state = g['state']
    val, col, ptr, rhs = assemble(dict(params), nx, ny, nz, dim, dof, state, g['x'], g['y'], g['z'])
    # fixed structural pattern -> drop |v| <= 1e-14 like CrsMatrix.compress, then everything is bit-exact
    assert_csr_equal(compress(val, col, ptr), (g['coA'], g['jcoA'], g['begA']), 0.0, name)
    if params.get('Problem Type') in ('Double Gyre', 'AMOC'):
        assert numpy.allclose(rhs, g['rhs'], rtol=1e-13, atol=1e-13 * numpy.abs(g['rhs']).max())
    else:
        assert numpy.array_equal(rhs, g['rhs'])


@pytest.mark.parametrize('name', ['ldc3d_rand', 'rb3d_rand', 'dhc3d_rand', 'amoc_rand', 'qg_rand', 'ldc2d_rand'])
def test_pattern_is_state_independent_superset(name):
    '''The fixed pattern must contain the reference pattern at ANY state, and must itself
    be reached at a generic state (no structurally dead slots).'''
    params, nx, ny, nz, dim, dof, kind = CASES[name]
    orc = Oracle(dict(params), nx, ny, nz, dim, dof)
    for k in ('zero', 'lin', 99):
        state = make_state(k, orc.n)
        val, col, ptr, rhs = assemble(dict(params), nx, ny, nz, orc.dim, orc.dof, state, orc.x, orc.y, orc.z)
        assert_csr_equal(compress(val, col, ptr), orc.jacobian(state), 0.0, name)
        assert numpy.array_equal(rhs, orc.rhs(state)) or numpy.allclose(rhs, orc.rhs(state), rtol=1e-13, atol=1e-13)
    if params.get('Problem Type') not in ('AMOC',):
        assert numpy.all(numpy.abs(val) > 1e-14), 'generic state should populate the full structural pattern'
