'''The generated per-row assembly functions (exactly the code each CUDA thread executes),
compiled for the host, against the oracle and the committed golden vectors.  CPU-only.'''
import numpy
import pytest

from cases import CASES, CUSTOM_BC_CASES, _bc_unsupported, make_state
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


@pytest.mark.parametrize('name', sorted(CUSTOM_BC_CASES))
def test_user_boundary_conditions_match_golden(name):
    '''User-supplied ``boundary_conditions(bc, atom)`` callbacks (Discretization.py:62-66,719; the first case is the
    callback of the reference's own tests/test_interface.py:117-150): the recorded ops select a generated kernel family,
    the callback's constants become run-time arguments -- CSR and RHS bit-identical to the reference run with the same
    callback (fixtures: tests/golden/make_golden.py).'''
    params, nx, ny, nz, dim, dof, kind, callback = CUSTOM_BC_CASES[name]
    g = load_case(name)
    val, col, ptr, rhs = assemble(dict(params), nx, ny, nz, dim, dof, g['state'], g['x'], g['y'], g['z'],
                                  boundary_conditions=callback)
    assert_csr_equal(compress(val, col, ptr), (g['coA'], g['jcoA'], g['begA']), 0.0, name)
    assert numpy.array_equal(rhs, g['rhs'])


def test_boundary_recorder_and_matcher():
    from transiflow_b200 import recipes
    ops = recipes.record_boundary_conditions(CUSTOM_BC_CASES['custom_fast_lid'][-1])
    assert ops[-2:] == [('force', 2, 1, 'u', 'lidv', 3.5), ('wall', 2, 1, -1)]          # moving lid = forcing + no-slip fold
    cfg = recipes.match_recorded(ops, 3, 4, 4)
    assert cfg is not None and cfg.name == 'ldc3d' and cfg.recipe == ops
    assert recipes.find_config(recipes.LDC, 3, 4, 4).recipe != ops                      # the generated config is untouched
    assert recipes.match_recorded(ops, 3, 4, 5) is None                                 # other unknowns: other family
    assert recipes.match_recorded(recipes.record_boundary_conditions(_bc_unsupported), 2, 1, 3) is None
